// Early pass of the two-phase scheme, sparse formulation:
//   VulkanShaders/InitialDrawCull.comp.glsl:12-59, HlslShaders/CS/drawOccFirst.cs.hlsl:12-59  (paths relative to /root/reference/src/Renderer)
//     if (visibility[i] == 0) return;  frustum;  LOD;  append {i, lod.indexCount, 1, lod.firstIndex, 0, 0}
//
// In steady state only a few percent of the objects were visible last frame, so the pass is a 4-B/object stream over the
// visibility buffer plus a gather for the survivors of that test (algorithmic bytes N*4 + v*40 + s*R).  The general
// pipelined kernel (cull_draw.cu, PASS_EARLY) walks every tile through its full machinery and needs 0.23 ms for 16.7 M
// objects at 3.7 % visible (profiles/r01e_*); this kernel is built for the common case instead:
//   * one CTA per tile of 4096 objects (tile id from an atomic ticket, so tile order == start order);
//   * the tile's visibility words arrive as four coalesced 128-bit loads per thread, all issued before the first use;
//   * the ids of the visible objects are compacted IN ORDER into shared memory (popc + warp shuffle scan + 32-entry scan);
//   * the compacted ids are culled densely, 256 at a time: RenderObject -> transform gather -> sphere + frustum -> LOD,
//     ranked with ballot/popc so that the tile's survivors stay in ascending id;
//   * cross-tile offsets: single-pass decoupled look-back (scan_lookback.cuh).  CTAs are short-lived and 8 are resident per
//     SM, so by the time a tile looks back its predecessors have usually published their prefix already;
//   * records are expanded from 4-B descriptors and written as one contiguous span.
// When most objects are visible the id list is as long as the tile and this degenerates into a plain (unpipelined) cull; the
// C-ABI layer then prefers the pipelined kernel (see run_draw_pass in capi.cu).
#include "cull_kernels.cuh"
#include "cull_math.cuh"
#include "scan_lookback.cuh"

namespace blz {

namespace {

constexpr int kEarlyThreads = 256;
constexpr int kEarlyVec = 4;                                     // 128-bit loads per thread
constexpr int kEarlyTile = kEarlyThreads * 4 * kEarlyVec;        // 4096 objects
constexpr int kEarlyWarps = kEarlyThreads / 32;

__device__ __forceinline__ uint4 ldg_nc_u4(const void* p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_cs_u2(void* p, uint2 v) { asm volatile("st.global.cs.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory"); }

} // namespace

// ------------------------------------------------------------------------------------------------------------------------
// Early pass, pipelined (the default): persistent CTAs, the 4-B visibility stream arrives by TMA bulk copies, and the three
// dependent memory round trips of a previously-visible object (visibility word -> RenderObject -> transform) are spread over
// three consecutive iterations, each iteration working on three different BATCHES at once:
//     B  batch i    : candidate id (compacted, ascending) -> RenderObject load goes out
//     C  batch i-1  : RenderObject is there -> the two transform loads go out
//     D  batch i-2  : transform is there -> sphere + frustum + LOD, ballot/popc rank, descriptor into the tile's staging buffer
// plus A: ordered compaction of the NEXT tile's visible ids while the current tile's last batch is in B.  A batch is up to 256
// candidates of one tile of 2048 objects (one per thread; a tile with more candidates simply takes more batches, so a fully visible
// scene degenerates into a correct, unpipelined-per-thread but still latency-overlapped cull).  The one-shot kernel above paid
// those round trips back to back in every CTA: 0.104 ms at 3.7 % visible; the visibility stream itself is 10 us of HBM time.
// Cross-tile offsets as in cull_stream.cu: a tile publishes its AGGREGATE when its last batch is through D, its records are written
// two iterations later from the sum of the aggregates between the CTA's consecutive tiles (all threads, one round trip).
// ------------------------------------------------------------------------------------------------------------------------
namespace {

constexpr int kEsThreads = 256, kEsWarps = kEsThreads / 32;
constexpr int kEsTile = 2048;                         // objects per tile, visibility WORDS as the source: 8 words per thread (2 x 128-bit from shared memory)
constexpr int kEsQ = kEsTile / (kEsThreads * 4);     // 128-bit words per thread
constexpr int kEsTileBits = 4096;                     // objects per tile, visibility BITS as the source: 128 mask words, one per thread of the first four warps
constexpr uint32_t kEsNone = 0xFFFFFFFFu;
constexpr uint32_t kEsLocalBits = 12;

__device__ __forceinline__ uint32_t es_smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void es_mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(es_smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void es_mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(es_smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void es_mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(es_smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void es_tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(es_smem_u32(dst)), "l"(src), "r"(bytes), "r"(es_smem_u32(bar)) : "memory");
}

struct EsBatch { uint32_t tile, seq, first, count, last; };   // tile id, the CTA's sequence number of that tile, first candidate, candidates, last batch of the tile?

} // namespace

// BITS = true: the source is the 1-bit-per-object mask the late pass (cull_stream.cu) leaves next to visibility[] -- 2 MB instead
// of 67 MB for 16.7 M objects; the C-ABI layer rebuilds it with pack_vis_bits_kernel whenever something else wrote visibility[].
template <bool BITS>
__global__ void __launch_bounds__(kEsThreads, BITS ? 3 : 4) early_stream_kernel(const __grid_constant__ DrawCullParams p)
{
    constexpr int TILE = BITS ? kEsTileBits : kEsTile, THREADS = kEsThreads, WARPS = kEsWarps, WORDS = TILE / 32;
    extern __shared__ __align__(128) unsigned char es_smem[];
    uint32_t (*s_vis)[TILE] = reinterpret_cast<uint32_t (*)[TILE]>(es_smem);                               // [2] TMA destination: visibility words of the tile being compacted / the next one
    constexpr size_t VISB = BITS ? 0 : size_t(2) * TILE * 4;                                               // no visibility ring when the source is the bit mask
    uint32_t (*s_stage)[TILE] = reinterpret_cast<uint32_t (*)[TILE]>(es_smem + VISB);             // [3] survivor descriptors of a tile (seq % 3): index in tile | lodId << 12
    uint16_t (*s_ids)[TILE] = reinterpret_cast<uint16_t (*)[TILE]>(es_smem + VISB + size_t(3) * TILE * 4);               // [2] ascending index-in-tile of the visible objects, per tile (seq & 1)
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uint32_t s_cntA[kEsQ * WARPS], s_cntD[WARPS], s_sum[4], s_tiles[8];

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t laneLt = (1u << lane) - 1u;
    const ViewConsts& V = p.view;
    uint32_t epoch;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(epoch) : "l"(&p.ctl->epoch));
    epoch &= 0x3FFFFFFFu;
    const uint32_t cap32 = p.capacity > 0xFFFFFFFFull ? 0xFFFFFFFFu : uint32_t(p.capacity);

    auto issue_vis = [&](uint32_t t, uint32_t slot) {        // thread 0: bulk copy of the tile's visibility words
        const uint32_t first = t * uint32_t(TILE);
        const uint32_t cnt = min(uint32_t(TILE), p.n - first);
        const uint32_t bytes = (cnt * 4u + 15u) & ~15u;       // the visibility buffer is padded by 16 B (capi.cu)
        es_mbar_expect_tx(&s_bar[slot], bytes);
        es_tma_load_1d(&s_vis[slot][0], p.visibility + first, bytes, &s_bar[slot]);
    };
    auto tile_of = [&](uint32_t seq) -> uint32_t { const uint32_t t = s_tiles[seq & 7u]; return t < p.numTiles ? t : kEsNone; };

    // tickets: sequence number m+1 is claimed when the compaction of m starts -- the same claim-to-use distance for every tile
    uint32_t lastClaim = 0u;
    if (tid == 0) {
        es_mbar_init(&s_bar[0], 1u); es_mbar_init(&s_bar[1], 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        lastClaim = atomicAdd(&p.ctl->ticket, 1u);
        s_tiles[0] = lastClaim; s_tiles[1] = kEsNone;
        s_sum[0] = s_sum[1] = s_sum[2] = s_sum[3] = 0u;
    }
    __syncthreads();
    if (!BITS && tid == 0 && tile_of(0u) != kEsNone) issue_vis(tile_of(0u), 0u);
    uint32_t wNext = 0u;                                      // BITS: mask word of the next tile to compact (threads < WORDS), loaded an iteration ahead
    if (BITS && tile_of(0u) != kEsNone && tid < uint32_t(WORDS)) wNext = __ldg(p.visBits + size_t(tile_of(0u)) * WORDS + tid);

    const EsBatch none{ kEsNone, 0u, 0u, 0u, 0u };
    EsBatch bC = none, bD = none;                             // batches in stages C and D of the coming iteration
    // tile whose batches currently enter B
    uint32_t curTile = kEsNone, curSeq = 0u, curCand = 0u, curNext = 0u;   // curNext: first candidate of the next batch
    uint32_t nextSeq = 0u;                                    // sequence number of the next tile to compact
    bool started = false;                                     // has the first compaction been triggered?
    // per-thread pipeline registers
    uint2 ob = make_uint2(0u, 0u); uint32_t locB = 0u;        // B -> C
    float4 ps = make_float4(0.f, 0.f, 0.f, 1.f), qt = make_float4(0.f, 0.f, 0.f, 1.f); uint32_t sidC = 0u, locC = 0u;   // C -> D
    uint32_t tileEmitted = 0u;                                // descriptors staged so far for the tile in D
    uint32_t candSum = 0u;                                    // thread 0: previously-visible objects of the tiles this CTA compacted (feedback for the C-ABI layer's choice of kernel)
    // finished tiles waiting for their prefix: records go out two iterations after the aggregate was published
    uint32_t f1Tile = kEsNone, f1Total = 0u, f1Slot = 0u, f2Tile = kEsNone, f2Total = 0u, f2Slot = 0u;
    uint32_t cum = 0u, nextRead = 0u;

    for (uint32_t it = 0u;; ++it) {
        // ---- which batch enters B, and does the next tile get compacted in this iteration? -----------------------------------
        EsBatch bB = none;
        bool lastB = false;
        if (curTile != kEsNone) {
            const uint32_t cnt = min(uint32_t(THREADS), curCand - min(curCand, curNext));
            bB = EsBatch{ curTile, curSeq, curNext, cnt, (curNext + uint32_t(THREADS) >= curCand) ? 1u : 0u };
            lastB = bB.last != 0u;
        }
        const bool trigger = (lastB || !started || curTile == kEsNone) ;
        const uint32_t aTile = trigger ? tile_of(nextSeq) : kEsNone;        // tile compacted in this iteration
        uint32_t ticket = kEsNone;
        if (trigger && aTile != kEsNone && tid == 0 && lastClaim < p.numTiles) { ticket = atomicAdd(&p.ctl->ticket, 1u); lastClaim = ticket; }

        // prefix of the tile whose records go out now: status loads first, consumed before the barrier
        constexpr int NS = 2;
        uint64_t sw[NS];
        if (f2Tile != kEsNone) {
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const uint32_t t = nextRead + tid + uint32_t(s) * THREADS;
                sw[s] = t < f2Tile ? ld_status(p.status + t) : 0ull;
            }
        }

        // Loads first, each with a whole iteration of lead: B and C issue at the top, D consumes what C issued one iteration ago.
        // ---- B (batch i): candidate id -> RenderObject load goes out ----------------------------------------------------------------------------
        uint2 ob2 = ob; uint32_t locB2 = locB;
        if (bB.tile != kEsNone && tid < bB.count) {
            locB2 = s_ids[bB.seq & 1u][bB.first + tid];
            ob2 = __ldg(reinterpret_cast<const uint2*>(p.objs + size_t(bB.tile) * TILE + locB2));
        }
        // ---- C (batch i-1): the RenderObject is here: the two transform loads go out -----------------------------------------------
        float4 ps2 = ps, qt2 = qt; uint32_t sidC2 = sidC, locC2 = locC;
        if (bC.tile != kEsNone && tid < bC.count) {
            const uint32_t t = ob.x - p.transformIdBase;
            ld_transform(p.xf + t, ps2, qt2);
            sidC2 = ob.y; locC2 = locB;
        }
        // ---- D (batch i-2): the transform is here: sphere + frustum + LOD, rank inside the warp -----------------------------------
        bool emit = false; uint32_t lodId = 0u, rankD = 0u;
        if (bD.tile != kEsNone) {
            if (tid < bD.count) {
                const float4 bs = __ldg(reinterpret_cast<const float4*>(p.surfaces + sidC));
                const Sphere sp = view_space_sphere(bs.x, bs.y, bs.z, bs.w, ps.x, ps.y, ps.z, ps.w, qt.x, qt.y, qt.z, qt.w, V);
                if (frustum_test(sp, V)) {
                    emit = true;
                    const uint4 hi = __ldg(reinterpret_cast<const uint4*>(p.surfaces + sidC) + 1);         // {materialId, lodOffset, lodCount, vertexOffset}
                    const uint32_t rel = lod_select(sp, ps.w, V.lodTarget, hi.y, hi.z, [&](uint32_t li) { return __ldg(&p.lods[li].error); });
                    lodId = (p.flags & kFlagOnpcLodQuirk) ? rel : rel + hi.y;
                }
            }
            const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, emit);
            rankD = uint32_t(__popc(ballot & laneLt));
            if (lane == 0) s_cntD[warp] = uint32_t(__popc(ballot));
        }
        const uint32_t locD = locC;
        // ---- A part 1 (tile aTile): visibility words -> per-thread masks, warp scans, per-(q, warp) counts -----------------------
        uint32_t maskA[kEsQ], offA[kEsQ];
        if (aTile != kEsNone) {
            if (BITS) {
                const uint32_t c = uint32_t(__popc(wNext));
                uint32_t x = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d); if (lane >= uint32_t(d)) x += y; }
                maskA[0] = wNext; offA[0] = x - c;
                if (lane == 31) s_cntA[warp] = x;
            } else {
                const uint32_t slot = nextSeq & 1u;
                es_mbar_wait(&s_bar[slot], (nextSeq >> 1) & 1u);
                const uint32_t left = p.n - aTile * uint32_t(TILE);
#pragma unroll
                for (int q = 0; q < kEsQ; ++q) {
                    const uint32_t l0 = (uint32_t(q) * THREADS + tid) * 4u;
                    const uint4 w = *reinterpret_cast<const uint4*>(&s_vis[slot][l0]);
                    uint32_t m = (w.x != 0u ? 1u : 0u) | (w.y != 0u ? 2u : 0u) | (w.z != 0u ? 4u : 0u) | (w.w != 0u ? 8u : 0u);
                    if (l0 + 3u >= left) m &= (l0 >= left) ? 0u : ((1u << (left - l0)) - 1u);      // ragged tail: stale words in the ring
                    maskA[q] = m;
                    const uint32_t c = uint32_t(__popc(m));
                    uint32_t x = c;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d); if (lane >= uint32_t(d)) x += y; }
                    offA[q] = x - c;
                    if (lane == 31) s_cntA[q * WARPS + warp] = x;
                }
            }
        }
        // prefix sum contribution
        if (f2Tile != kEsNone) {
            uint32_t part = 0u;
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const uint32_t t = nextRead + tid + uint32_t(s) * THREADS;
                if (t < f2Tile) {
                    uint64_t w = sw[s];
                    while (uint32_t(w >> 34) != epoch || (uint32_t(w >> 32) & 3u) == 0u) { __nanosleep(40); w = ld_status(p.status + t); }
                    part += uint32_t(w);
                }
            }
            for (uint32_t t = nextRead + tid + uint32_t(NS) * THREADS; t < f2Tile; t += THREADS) {
                uint64_t w;
                do { w = ld_status(p.status + t); } while (uint32_t(w >> 34) != epoch || (uint32_t(w >> 32) & 3u) == 0u);
                part += uint32_t(w);
            }
            part = __reduce_add_sync(0xFFFFFFFFu, part);
            if (lane == 0 && part != 0u) atomicAdd(&s_sum[it & 3u], part);
        }
        __syncthreads();      // (1) counts of D and A, aggregate sum visible; the ring slot of aTile is consumed

        // ---- ticket -> visibility ring (one tile of lead) --------------------------------------------------------------------------------
        if (trigger && aTile != kEsNone && tid == 0) {
            s_tiles[(nextSeq + 1u) & 7u] = ticket;
            if (!BITS && ticket < p.numTiles) issue_vis(ticket, (nextSeq + 1u) & 1u);          // that slot held tile nextSeq-1, consumed an iteration ago
        }
        // ---- D part 2: descriptors into the tile's staging buffer; last batch -> publish the tile's aggregate -------------------------
        uint32_t f0Tile = kEsNone, f0Total = 0u, f0Slot = 0u;
        if (bD.tile != kEsNone) {
            if (bD.first == 0u) tileEmitted = 0u;
            uint32_t warpOff = 0u, total = 0u;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) { const uint32_t c = s_cntD[w]; if (uint32_t(w) < warp) warpOff += c; total += c; }
            const uint32_t slot = bD.seq % 3u;
            if (emit) s_stage[slot][tileEmitted + warpOff + rankD] = locD | (lodId << kEsLocalBits);
            tileEmitted += total;
            if (bD.last) {
                if (tid == 0) st_status(p.status + bD.tile, pack_status(epoch, kStateAggregate, tileEmitted));
                f0Tile = bD.tile; f0Total = tileEmitted; f0Slot = slot;
            }
        }
        // ---- A part 2: ordered ids of aTile ------------------------------------------------------------------------------------------------
        uint32_t aCand = 0u;
        if (aTile != kEsNone) {
            uint16_t* ids = &s_ids[nextSeq & 1u][0];
            if (BITS) {
                uint32_t o = offA[0];
#pragma unroll
                for (int w = 0; w < WARPS; ++w) { const uint32_t c = s_cntA[w]; if (uint32_t(w) < warp) o += c; aCand += c; }
                uint32_t m = maskA[0];
                while (m != 0u) { ids[o++] = uint16_t(tid * 32u + uint32_t(__ffs(int(m))) - 1u); m &= m - 1u; }
            } else {
                uint32_t base[kEsQ];
#pragma unroll
                for (int q = 0; q < kEsQ; ++q) base[q] = 0u;
#pragma unroll
                for (int i = 0; i < kEsQ * WARPS; ++i) {
                    const uint32_t c = s_cntA[i];
#pragma unroll
                    for (int q = 0; q < kEsQ; ++q) if (uint32_t(i) < uint32_t(q) * WARPS + warp) base[q] += c;
                    aCand += c;
                }
#pragma unroll
                for (int q = 0; q < kEsQ; ++q) {
                    uint32_t o = base[q] + offA[q];
                    const uint32_t l0 = (uint32_t(q) * THREADS + tid) * 4u;
#pragma unroll
                    for (int c = 0; c < 4; ++c) if (maskA[q] & (1u << c)) ids[o++] = uint16_t(l0 + uint32_t(c));
                }
            }
        }
        // ---- records of the tile finished two iterations ago ----------------------------------------------------------------------------------
        if (f2Tile != kEsNone) {
            const uint32_t prefix = cum + s_sum[it & 3u];
            if (f2Tile == p.numTiles - 1u && tid == 0) {
                const uint32_t all = prefix + f2Total;
                p.counts[0] = all < cap32 ? all : cap32;
                p.counts[1] = all;
            }
            const uint32_t room = prefix < cap32 ? cap32 - prefix : 0u;
            const uint32_t nrec = room < f2Total ? room : f2Total;
            uint2* dst = reinterpret_cast<uint2*>(p.draws + size_t(prefix) * p.recWords);
            const uint32_t idBase = p.objectIdBase + f2Tile * uint32_t(TILE);
            const uint32_t wpr = p.recWords >> 1;
            for (uint32_t w = tid; w < nrec * wpr; w += THREADS) {
                const uint32_t r = wpr == 3u ? w / 3u : w >> 2, f = w - r * wpr;
                const uint32_t d = s_stage[f2Slot][r];
                const uint2 L = __ldg(reinterpret_cast<const uint2*>(p.lods + (d >> kEsLocalBits)));       // {indexCount, firstIndex}
                st_cs_u2(dst + w, f == 0u ? make_uint2(idBase + (d & ((1u << kEsLocalBits) - 1u)), L.x) : (f == 1u ? make_uint2(1u, L.y) : make_uint2(0u, 0u)));
                if (f == 2u && p.descs != nullptr) st_cs_u2(p.descs + prefix + r, make_uint2(idBase + (d & ((1u << kEsLocalBits) - 1u)), d >> kEsLocalBits));
            }
            cum = prefix + f2Total;
            nextRead = f2Tile + 1u;
        }
        if (tid == 0) s_sum[(it + 2u) & 3u] = 0u;
        __syncthreads();      // (2) ids of aTile visible (read by B from the next iteration on); staging / counters free for reuse

        // the loads issued at the top of this iteration become the pipeline registers of the next one
        ob = ob2; locB = locB2; ps = ps2; qt = qt2; sidC = sidC2; locC = locC2;
        // ---- advance --------------------------------------------------------------------------------------------------------------------------------
        f2Tile = f1Tile; f2Total = f1Total; f2Slot = f1Slot;
        f1Tile = f0Tile; f1Total = f0Total; f1Slot = f0Slot;
        bD = bC; bC = bB;
        if (curTile != kEsNone) { if (lastB) curTile = kEsNone; else curNext += uint32_t(THREADS); }
        if (trigger) {
            started = true;
            if (aTile != kEsNone) {
                // the tile just compacted becomes current as soon as the previous one has handed its last batch to B (this iteration)
                curTile = aTile; curSeq = nextSeq; curCand = aCand; curNext = 0u;
                candSum += aCand;
                ++nextSeq;
                if (BITS) {                                  // its successor's ticket became visible behind barrier (2): the mask word goes out now
                    const uint32_t nt = tile_of(nextSeq);
                    wNext = (nt != kEsNone && tid < uint32_t(WORDS)) ? __ldg(p.visBits + size_t(nt) * WORDS + tid) : 0u;
                }
            }
        }
        if (curTile == kEsNone && bC.tile == kEsNone && bD.tile == kEsNone && f1Tile == kEsNone && f2Tile == kEsNone && (started && tile_of(nextSeq) == kEsNone)) break;
    }

    if (p.n == 0u && blockIdx.x == 0 && tid == 0) { p.counts[0] = 0u; p.counts[1] = 0u; }
    if (tid == 0) {
        if (p.visTotal != nullptr && candSum != 0u) atomicAdd(p.visTotal, candSum);
        __threadfence();
        const uint32_t prev = atomicAdd(&p.ctl->done, 1u);
        if (prev == gridDim.x - 1u) {
            // last CTA out: the total goes straight to the host-mapped word (one posted write, no memset / memcpy on the stream), accumulator re-armed
            if (p.visTotal != nullptr) { const uint32_t tot = atomicExch(p.visTotal, 0u); *reinterpret_cast<volatile uint32_t*>(p.visTotalOut) = tot; }
            uint32_t e = (epoch + 1u) & 0x3FFFFFFFu;
            p.ctl->epoch = e ? e : 1u;
            p.ctl->ticket = 0u;
            p.ctl->done = 0u;
        }
    }
}

template <bool BITS>
static cudaError_t launch_early_stream_t(const DrawCullParams& p, int numSMs, cudaStream_t stream)
{
    constexpr int TILE = BITS ? kEsTileBits : kEsTile;
    if (p.lodCount >= (1u << (32 - kEsLocalBits))) return cudaErrorInvalidValue;
    DrawCullParams q = p;
    q.numTiles = uint32_t((uint64_t(p.n) + TILE - 1) / TILE);
    const size_t smem = size_t(TILE) * ((BITS ? 0 : 2 * 4) + 3 * 4 + 2 * 2);
    cudaError_t e = cudaFuncSetAttribute(early_stream_kernel<BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    int perSM = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, early_stream_kernel<BITS>, kEsThreads, smem);
    if (e != cudaSuccess) return e;
    if (perSM < 1) perSM = 1;
    uint32_t grid = uint32_t(numSMs) * uint32_t(perSM);
    if (grid > q.numTiles) grid = q.numTiles;
    if (grid < 1) grid = 1;
    early_stream_kernel<BITS><<<grid, kEsThreads, smem, stream>>>(q);
    return cudaGetLastError();
}

// p.visBits != nullptr: the 1-bit mask is the source (and is valid); else the 4-B visibility words
cudaError_t launch_early_stream(const DrawCullParams& p, int numSMs, cudaStream_t stream)
{
    return p.visBits ? launch_early_stream_t<true>(p, numSMs, stream) : launch_early_stream_t<false>(p, numSMs, stream);
}

// visibility[] -> 1 bit per object (word i covers objects 32 i .. 32 i + 31; bits of objects >= n are 0)
__global__ void pack_vis_bits_kernel(const uint32_t* __restrict__ vis, uint32_t* __restrict__ bits, uint32_t n, uint32_t words)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;              // one object per thread, a warp makes one word
    const uint32_t b = __ballot_sync(0xFFFFFFFFu, i < n && vis[i] != 0u);
    if ((threadIdx.x & 31u) == 0u && (i >> 5) < words) bits[i >> 5] = b;
}

// 1 bit per object -> the reference's u32-per-object visibility buffer
__global__ void unpack_vis_bits_kernel(const uint32_t* __restrict__ bits, uint32_t* __restrict__ vis, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) vis[i] = (__ldg(bits + (i >> 5)) >> (i & 31u)) & 1u;
}

// population of the 1-bit visibility mask (2 MB for 16.7 M objects): only launched next to the DENSE early pass, whose kernel does not count
__global__ void popc_vis_bits_kernel(const uint32_t* __restrict__ bits, uint32_t words, uint32_t* __restrict__ total)
{
    uint32_t acc = 0u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < words; i += gridDim.x * blockDim.x) acc += uint32_t(__popc(__ldg(bits + i)));
    acc = __reduce_add_sync(0xFFFFFFFFu, acc);
    if ((threadIdx.x & 31u) == 0u && acc != 0u) atomicAdd(total, acc);
}

cudaError_t launch_popc_vis_bits(const uint32_t* bits, uint32_t n, uint32_t* total, cudaStream_t stream)
{
    const uint32_t words = (n + 31u) / 32u;
    if (words == 0u) return cudaSuccess;
    popc_vis_bits_kernel<<<148, 256, 0, stream>>>(bits, words, total);
    return cudaGetLastError();
}

cudaError_t launch_unpack_vis_bits(const uint32_t* bits, uint32_t* vis, uint32_t n, cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    unpack_vis_bits_kernel<<<(n + 255u) / 256u, 256, 0, stream>>>(bits, vis, n);
    return cudaGetLastError();
}

cudaError_t launch_pack_vis_bits(const uint32_t* vis, uint32_t* bits, uint32_t n, cudaStream_t stream)
{
    const uint32_t words = (n + 31u) / 32u;
    if (words == 0) return cudaSuccess;
    const uint32_t threads = words * 32u;
    pack_vis_bits_kernel<<<(threads + 255u) / 256u, 256, 0, stream>>>(vis, bits, n, words);
    return cudaGetLastError();
}

} // namespace blz
