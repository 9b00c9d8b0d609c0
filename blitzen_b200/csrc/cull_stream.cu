// Draw-cull passes, streaming formulation: frustum / early / late (Hi-Z) / temporal, replacing
//   VulkanShaders/{Initial,Late,Transparent,Onpc}DrawCull.comp.glsl and HlslShaders/CS/{drawCull,drawOccFirst,drawOccLate,drawOccTemporal}
// (paths relative to /root/reference/src/Renderer).
//
// History: the pipelined kernel of cull_draw.cu (v4) reached 52 % of the measured HBM peak and its ncu capture
// (profiles/r01e_*) showed it issue-bound: ~375 warp instructions per 32 objects, of which only ~70 are the FP32 arithmetic of the
// shader -- the rest is the software pipeline (cp.async address arithmetic, a shared-memory round trip of every input, ring
// bookkeeping).  A one-shot-CTA kernel with plain register loads and a CTA-wide look-back (v5, profiles/r01g_*) had the instruction
// count right but a CTA lifetime of ~10 dependent memory round trips: 0.32 ms.  This kernel (v6) keeps v5's instruction economy and
// hides every round trip:
//
//   * persistent co-resident CTAs, tiles of THREADS x ITEMS consecutive objects dealt round-robin (the m-th tile of CTA b is
//     m * grid + b); lane l of warp w owns objects w*32*ITEMS + l + 32k: every global access of a warp is one contiguous run of
//     fully used sectors;
//   * the two CONTIGUOUS input streams (RenderObject 8 B, visibility 4 B) arrive by TMA: one elected thread issues two 1-D bulk
//     copies per tile (cp.async.bulk ... mbarrier::complete_tx) into a 2-stage shared-memory ring, two tiles ahead -- no per-thread
//     load instructions, no address arithmetic, no registers held;
//   * the transform GATHER (by transformId, two float4 streams) is issued straight into registers for tile j+1 as soon as the
//     arithmetic of tile j has released them, and consumed at the top of the next iteration: the queue evaluation, the look-back
//     and the record write-out of the current tile all overlap it;
//   * sphere + frustum planes in registers; the survivors (a few %) go to a CTA-wide queue in shared memory so that the expensive
//     tail (projectSphere: 2 sqrt + 5 IEEE div, Hi-Z fetch, LOD loop) runs with full warps instead of 3 active lanes of 32;
//   * compaction is deterministic: ballot/popc in the warp, scan of the warp counts in the CTA, single-pass decoupled look-back
//     across tiles, chain-free: a tile publishes its AGGREGATE when it is computed; one iteration later its CTA sums the aggregates of
//     every tile between its previous tile and that one (all threads at once: one L2 round trip, overlapped with the next queue
//     evaluation).  A warp-wide look-back that waits for inclusive prefixes was measured at 1-2 ms: persistent CTAs run in
//     lockstep rounds, so the resolved front advances only 32 tiles per round trip;
//   * survivors are staged as 4-B descriptors and leave the CTA as one contiguous span of 8-B stores (24-/32-B records).
#include "cull_kernels.cuh"
#include "cull_math.cuh"
#include "scan_lookback.cuh"

namespace blz {

namespace {

__device__ __forceinline__ uint2 ld_stream_u2(const void* p)
{
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_stream_f4(const void* p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
// visibility is read and (late pass) rewritten by the same thread: not the read-only path
__device__ __forceinline__ uint32_t ld_vis_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.global.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_rec_u2(void* p, uint2 v) { asm volatile("st.global.cs.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory"); }

constexpr uint32_t kSLocalBits = 12;                 // index in tile (TILE <= 4096)
constexpr uint32_t kSLocalMask = (1u << kSLocalBits) - 1u;

// Evaluates queue entry e of a warp's private queue: Hi-Z (late / temporal passes) and LOD selection for emitters.
// Returns visible | emit << 1 | lodId << 2.
template <int PASS, int HIZ>
__device__ __forceinline__ uint32_t eval_entry(const float4 q, const uint32_t scaleBits, const bool visPrev, const uint32_t sidx,
                                               const PrimitiveSurface* surfT, const LodData* lodT, const DrawCullParams& p)
{
    constexpr bool HAS_HIZ = (PASS == PASS_LATE || PASS == PASS_TEMPORAL);
    const ViewConsts& V = p.view;
    const Sphere s{ q.x, q.y, q.z, q.w };
    bool visible = true;
    if (HAS_HIZ) {
        float4 aabb;
        if (project_sphere(s, V.zNear, V.proj0, V.proj5, aabb))
            visible = (HIZ == HIZ_VK) ? hiz_test_vk(aabb, p.pyr, s, V) : hiz_test_dx(aabb, p.pyr, s, V);
    }
    bool emit = visible;
    if (PASS == PASS_LATE) emit = visible && !visPrev;                             // LateDrawCull.comp.glsl:49
    uint32_t lodId = 0u;
    if (emit) {
        const uint32_t lodOffset = surfT[sidx].lodOffset, lodCount = surfT[sidx].lodCount;
        const uint32_t rel = lod_select(s, __uint_as_float(scaleBits), V.lodTarget, lodOffset, lodCount, [&](uint32_t li) { return lodT[li].error; });
        lodId = (p.flags & kFlagOnpcLodQuirk) ? rel : rel + lodOffset;
    }
    return (visible ? 1u : 0u) | (emit ? 2u : 0u) | (lodId << 2);
}

constexpr int kStreamDepth = 2;                      // stages of the TMA input ring (a stage is refilled right after it is consumed: two tiles of lead)
constexpr uint32_t kNoTile = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
// 1-D bulk copy global -> shared through the TMA unit; completion is signalled on the mbarrier (bytes: multiple of 16, both addresses 16-B aligned)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// IsObjectInsideViewFrustum's four comparisons (cull_math.cuh frustum_test) without short-circuit branches: the four items of a
// thread are evaluated as straight-line code and interleave in the FP pipe
__device__ __forceinline__ bool frustum_test_nb(const Sphere& s, const ViewConsts& V)
{
    const bool a = fsub(fmul(s.z, V.frustumLeft), fmul(fabsf(s.x), V.frustumRight)) > -s.r;
    const bool b = fsub(fmul(s.z, V.frustumBottom), fmul(fabsf(s.y), V.frustumTop)) > -s.r;
    const bool c = fadd(s.z, s.r) > V.zNear;
    const bool d = fsub(s.z, s.r) < V.zFar;
    return a & b & c & d;
}

constexpr int kLagS = 2;                             // a tile's records are written this many iterations after its arithmetic
constexpr int kStagesS = kLagS + 1;                  // descriptor buffers

template <int PASS, int HIZ, int THREADS, int ITEMS, int MINB, bool SMEM_TABLES>
__global__ void __launch_bounds__(THREADS, MINB) stream_cull_kernel(const __grid_constant__ DrawCullParams p)
{
    constexpr int TILE = THREADS * ITEMS, WARPS = THREADS / 32, D = kStreamDepth, WSPAN = 32 * ITEMS;   // WSPAN: objects of a tile owned by one warp
    constexpr bool VIS_WORDS = (PASS == PASS_EARLY);     // last frame's visibility arrives as 4-B words (generic early pass) ...
    constexpr bool VIS_BITS = (PASS == PASS_LATE);       // ... or as the 1-bit-per-object mask (late pass): 128 B per tile instead of 4 KB
    constexpr int VIS_STAGE_WORDS = VIS_WORDS ? TILE : (VIS_BITS ? TILE / 32 : 0);
    static_assert(TILE <= (1 << kSLocalBits), "descriptor packing");
    static_assert(!VIS_BITS || (TILE / 32 * 4) % 16 == 0, "bulk copy granularity of the mask words");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_bar[D];
    __shared__ uint32_t s_warpCnt[2][WARPS]; // records of each warp's span of the tile computed in iteration j, at [j & 1]
    __shared__ uint32_t s_totals[4];         // records of the tile computed in iteration j, at [j & 3]
    __shared__ uint32_t s_sum[4];            // sum of the aggregates between this CTA's consecutive tiles, at [j & 3]
    __shared__ uint32_t s_tiles[8];          // dynamic order: the CTA's m-th tile at [m & 7]
    __shared__ uint32_t s_visPrev[2][VIS_BITS ? THREADS * ITEMS / 32 : 1];   // late pass: last frame's mask words of the tile fetched in iteration j-1, at [j & 1] (the ring stage is refilled before the queue is evaluated)

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t laneLt = (1u << lane) - 1u;
    const uint32_t localBase = warp * uint32_t(WSPAN) + lane;           // + 32k = index inside the tile
    const ViewConsts& V = p.view;
    const uint32_t epoch = ld_cg_u32(&p.ctl->epoch) & 0x3FFFFFFFu;      // constant for the whole launch (the last CTA out bumps it)
    const uint32_t cap32 = p.capacity > 0xFFFFFFFFull ? 0xFFFFFFFFu : uint32_t(p.capacity);   // record counts fit 32 bits (n < 2^32)

    // ---- carve shared memory --------------------------------------------------------------------------------------------------
    unsigned char* sp = smem_raw;
    uint2* objRing = reinterpret_cast<uint2*>(sp);       sp += size_t(D) * TILE * sizeof(uint2);      // TMA destination: RenderObject stream
    uint32_t* visRing = reinterpret_cast<uint32_t*>(sp); sp += size_t(D) * VIS_STAGE_WORDS * sizeof(uint32_t);     // TMA destination: visibility stream (words or mask)
    // per-warp private survivor queue, WSPAN entries of 32 B: {view-space sphere} {scale bits, index in tile, surfaceId, D}
    // where D of entry i is the i-th EMITTER descriptor of the span (a compact list threaded through the entries' last words: it is
    // written by the evaluation of entry e >= i and nothing else reads that word)
    uint4* queue = reinterpret_cast<uint4*>(sp) + warp * (WSPAN * 2);   sp += size_t(TILE) * 32;
    uint32_t* stageB = reinterpret_cast<uint32_t*>(sp);  sp += size_t(kStagesS) * TILE * sizeof(uint32_t);   // survivor descriptors of whole tiles
    const PrimitiveSurface* surfT = p.surfaces;
    const LodData* lodT = p.lods;
    if (SMEM_TABLES) {
        uint4* dstS = reinterpret_cast<uint4*>(sp);
        const uint4* srcS = reinterpret_cast<const uint4*>(p.surfaces);
        for (uint32_t i = tid; i < p.surfaceCount * 2u; i += THREADS) dstS[i] = __ldg(srcS + i);
        uint4* dstL = dstS + p.surfaceCount * 2u;
        const uint4* srcL = reinterpret_cast<const uint4*>(p.lods);
        for (uint32_t i = tid; i < p.lodCount * 2u; i += THREADS) dstL[i] = __ldg(srcL + i);
        surfT = reinterpret_cast<const PrimitiveSurface*>(dstS);
        lodT = reinterpret_cast<const LodData*>(dstL);
    }

    // one elected thread feeds the input ring: the two contiguous streams of a tile arrive as two bulk copies on one mbarrier
    auto issue_tile = [&](uint32_t t, uint32_t stage) {
        const uint32_t first = t * uint32_t(TILE);
        const uint32_t cnt = min(uint32_t(TILE), p.n - first);
        const uint32_t ob = (cnt * 8u + 15u) & ~15u;                                                       // buffers are padded by 16 B (capi.cu)
        const uint32_t vb = VIS_WORDS ? ((cnt * 4u + 15u) & ~15u) : (VIS_BITS ? uint32_t(TILE / 32 * 4) : 0u);   // the mask is padded to whole tiles
        mbar_expect_tx(&s_bar[stage], ob + vb);
        tma_load_1d(objRing + stage * TILE, p.objs + first, ob, &s_bar[stage]);
        if (VIS_WORDS) tma_load_1d(visRing + stage * VIS_STAGE_WORDS, p.visibility + first, vb, &s_bar[stage]);
        if (VIS_BITS) tma_load_1d(visRing + stage * VIS_STAGE_WORDS, p.visBits + (first >> 5), vb, &s_bar[stage]);
    };

    // Tile order: an atomic ticket, claimed by thread 0 at the top of iteration j for the tile computed in iteration j+2 -- the SAME
    // claim-to-compute delay for every tile, so tile order follows time order and the lag absorbs the rest.  (Tickets claimed with
    // unequal delays -- two at once in the prologue -- made a CTA's second tile lower than its neighbour's first and chained the
    // waits through all CTAs: measured 1-2 ms.  A static round-robin order was kept as a run-time alternative until r02: the kernel
    // with only this order compiled in needs no spills at 80 registers and is 6.5 % faster.)
    auto tile_of = [&](uint32_t m) -> uint32_t { const uint32_t t = s_tiles[m & 7u]; return t < p.numTiles ? t : kNoTile; };
    uint32_t lastClaim = 0u;                 // thread 0, dynamic order: the most recent ticket
    if (tid == 0) {
        lastClaim = atomicAdd(&p.ctl->ticket, 1u); s_tiles[0] = lastClaim; s_tiles[1] = kNoTile;
#pragma unroll
        for (int s = 0; s < D; ++s) mbar_init(&s_bar[s], 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_sum[0] = s_sum[1] = s_sum[2] = s_sum[3] = 0u;
    }
    __syncthreads();      // barriers initialised, tables visible
    if (tid == 0 && tile_of(0u) != kNoTile) issue_tile(tile_of(0u), 0u);

    // registers carried from the fetch of a tile (iteration j-1) to its arithmetic (iteration j)
    float4 ps[ITEMS], qt[ITEMS];
    uint32_t sid[ITEMS];
    uint32_t actMask = 0u, inMask = 0u;
    uint32_t curTile = kNoTile, prev1 = kNoTile, prev2 = kNoTile;      // tiles of iterations j, j-1, j-2 (prev2's records go out in iteration j)
    uint32_t cum = 0u;                       // records emitted by tiles [0, nextRead)
    uint32_t nextRead = 0u;                  // first tile whose aggregate this CTA has not summed yet
    uint32_t slot3 = 0u;                     // descriptor buffer of iteration j (j mod 3)
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) { ps[k] = make_float4(0.f, 0.f, 0.f, 1.f); qt[k] = make_float4(0.f, 0.f, 0.f, 1.f); sid[k] = 0u; }

    // ONE CTA barrier per iteration: everything up to it is warp-local (arithmetic, the warp's private survivor queue, Hi-Z / LOD of
    // its own survivors, visibility write-back), so the warps of a CTA drift apart by up to an iteration and overlap each other's
    // memory and arithmetic phases.
    for (int j = -1;; ++j) {
        const uint32_t ju = uint32_t(j + 4);                     // j shifted to a non-negative value with the same residues mod 2 and 4
        const uint32_t mm = uint32_t(j + 1);                     // sequence number of the tile fetched in this iteration
        const bool valid = curTile != kNoTile;                   // uniform over the CTA
        const uint32_t tileBase = curTile * uint32_t(TILE);
        uint32_t ticket = kNoTile;
        if (tid == 0 && lastClaim < p.numTiles) { ticket = atomicAdd(&p.ctl->ticket, 1u); lastClaim = ticket; }   // tile of iteration j+2

        // prefix of tile j-2: this CTA sums the aggregates of every tile between its previous tile and that one -- all threads at
        // once (one L2 round trip however many tiles are in flight), two iterations late (the words are there), and the loads are
        // issued here and consumed just before the barrier, behind the whole arithmetic.
        constexpr int NS = 2;
        uint64_t sw[NS];
        if (prev2 != kNoTile) {
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const uint32_t t = nextRead + tid + uint32_t(s) * THREADS;
                sw[s] = t < prev2 ? ld_status(p.status + t) : 0ull;
            }
        }

        uint32_t emitRun = 0u;               // emitters of this warp's span (warp-uniform)
        uint32_t survMask = 0u, qn = 0u;     // frustum survivors of this thread / of this warp's span
        if (valid) {
            // ---- S1: sphere + frustum planes of tile j, straight-line (inputs in registers) -------------------------------------
            {
                Sphere sph[ITEMS];
#pragma unroll
                for (int k = 0; k < ITEMS; ++k) {
                    const float4 bs = *reinterpret_cast<const float4*>(&surfT[sid[k]]);       // {center.xyz, radius}
                    sph[k] = view_space_sphere(bs.x, bs.y, bs.z, bs.w, ps[k].x, ps[k].y, ps[k].z, ps[k].w, qt[k].x, qt[k].y, qt[k].z, qt[k].w, V);
                    survMask |= (frustum_test_nb(sph[k], V) ? 1u : 0u) << k;
                }
                survMask &= actMask;                               // inactive items ran on stale registers
                // survivors -> the warp's private queue, in ascending object order (k-major, then lane)
#pragma unroll
                for (int k = 0; k < ITEMS; ++k) {
                    const uint32_t ball = __ballot_sync(0xFFFFFFFFu, (survMask >> k) & 1u);
                    if ((survMask >> k) & 1u) {
                        uint4* e = queue + 2u * (qn + uint32_t(__popc(ball & laneLt)));
                        e[0] = make_uint4(__float_as_uint(sph[k].x), __float_as_uint(sph[k].y), __float_as_uint(sph[k].z), __float_as_uint(sph[k].r));
                        e[1] = make_uint4(__float_as_uint(ps[k].w), localBase + uint32_t(k) * 32u, sid[k], 0u);   // last iteration's descriptors are staged: D is free
                    }
                    qn += uint32_t(__popc(ball));
                }
            }
        }
        const uint32_t inMaskJ = inMask;
        if (tid == 0) {       // the ticket is back by now; its ring stage (that of tile j) was consumed before the barrier of iteration j-1
            s_tiles[(mm + 1u) & 7u] = ticket;
            if (ticket < p.numTiles) issue_tile(ticket, (mm + 1u) % uint32_t(D));
        }
        // ---- S2: fetch tile j+1: its RenderObject / visibility words are in the ring; the transform gather goes out now and is
        //      consumed at the top of the next iteration ------------------------------------------------------------------------
        const uint32_t nextTile = tile_of(mm);
        if (nextTile != kNoTile) {
            const uint32_t stage = mm % uint32_t(D);
            mbar_wait(&s_bar[stage], (mm / uint32_t(D)) & 1u);
            const uint2* ob = objRing + stage * TILE + localBase;
            const uint32_t* vr = visRing + stage * VIS_STAGE_WORDS + (VIS_WORDS ? localBase : warp * uint32_t(ITEMS));
            const uint32_t left = p.n - nextTile * uint32_t(TILE);         // objects from the tile's first to the end of the list (>= 1)
            // ONE code path with PREDICATED loads: a predicated load keeps its destination registers tied to the carried values; an
            // unpredicated copy of the loads behind a "whole tile in range" branch made the compiler merge the two paths with register
            // moves placed right behind the loads, which wait for the data (measured: +9 % time, long-scoreboard stalls doubled)
            actMask = 0u; inMask = 0u;
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const bool in = localBase + uint32_t(k) * 32u < left;
                const uint2 o = ob[k * 32];
                const bool act = (PASS == PASS_EARLY) ? (in && vr[k * 32] != 0u) : in;             // InitialDrawCull.comp.glsl:21-24
                sid[k] = in ? o.y : 0u;                                                    // the ragged tail of the ring holds stale words
                if (act) ld_transform(p.xf + (o.x - p.transformIdBase), ps[k], qt[k]);
                inMask |= (in ? 1u : 0u) << k;
                if (PASS == PASS_EARLY) actMask |= (act ? 1u : 0u) << k;
            }
            if (PASS != PASS_EARLY) actMask = inMask;
            if (VIS_BITS && lane < uint32_t(ITEMS)) s_visPrev[mm & 1u][warp * uint32_t(ITEMS) + lane] = vr[lane];
        }
        if (valid) {
            __syncwarp();
            // ---- S3: the warp evaluates its own survivors, 32 at a time: Hi-Z + LOD; emitters are compacted in place -----------------
            uint32_t bitsK = 0u;                                   // lane k < ITEMS: new visibility mask of the span's k-th 32 objects
            for (uint32_t e0 = 0u; e0 < qn; e0 += 32u) {
                const uint32_t e = e0 + lane;
                uint32_t res = 0u, local = 0u;
                if (e < qn) {
                    const uint4 a = queue[2u * e], b = queue[2u * e + 1u];
                    local = b.y;
                    bool vp = false;
                    if (PASS == PASS_LATE) { const uint32_t r = local - warp * uint32_t(WSPAN); vp = ((s_visPrev[ju & 1u][warp * uint32_t(ITEMS) + (r >> 5)] >> (r & 31u)) & 1u) != 0u; }
                    res = eval_entry<PASS, HIZ>(make_float4(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w)),
                                                b.x, vp, b.z, surfT, lodT, p);
                }
                const uint32_t eb = __ballot_sync(0xFFFFFFFFu, (res & 2u) != 0u);
                if (res & 2u) reinterpret_cast<uint32_t*>(queue)[8u * (emitRun + uint32_t(__popc(eb & laneLt))) + 7u] = local | ((res >> 2) << kSLocalBits);
                emitRun += uint32_t(__popc(eb));
                if (PASS == PASS_LATE) {                           // the span's visibility as ITEMS mask words: OR of the visible entries' bits
                    const uint32_t rel = local - warp * uint32_t(WSPAN);
                    const uint32_t bit = (res & 1u) << (rel & 31u);
#pragma unroll
                    for (int k = 0; k < ITEMS; ++k) {
                        const uint32_t w = __reduce_or_sync(0xFFFFFFFFu, (rel >> 5) == uint32_t(k) ? bit : 0u);
                        if (lane == uint32_t(k)) bitsK |= w;
                    }
                }
            }
            if (lane == 0) s_warpCnt[ju & 1u][warp] = emitRun;
            if (PASS == PASS_LATE) {
                // LateDrawCull.comp.glsl:70.  Visibility lives as a bit mask (one word per 32 consecutive objects): what the next early and
                // late passes stream; the reference's u32-per-object form is written too when the caller asked for it (option vis_words)
                const uint32_t firstObj = tileBase + warp * uint32_t(WSPAN);
                if (lane < uint32_t(ITEMS) && firstObj + lane * 32u < p.n) p.visBits[(firstObj >> 5) + lane] = bitsK;
                if (p.visibility != nullptr) {
#pragma unroll
                    for (int k = 0; k < ITEMS; ++k) {
                        const uint32_t w = __shfl_sync(0xFFFFFFFFu, bitsK, k);
                        if ((inMaskJ >> k) & 1u) p.visibility[tileBase + localBase + uint32_t(k) * 32u] = (w >> lane) & 1u;
                    }
                }
            }
        }

        if (prev2 != kNoTile) {
            uint32_t part = 0u;
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const uint32_t t = nextRead + tid + uint32_t(s) * THREADS;
                if (t < prev2) {
                    uint64_t w = sw[s];
                    while (uint32_t(w >> 34) != epoch || (uint32_t(w >> 32) & 3u) == 0u) { __nanosleep(40); w = ld_status(p.status + t); }
                    part += uint32_t(w);
                }
            }
            for (uint32_t t = nextRead + tid + uint32_t(NS) * THREADS; t < prev2; t += THREADS) {
                uint64_t w;
                do { w = ld_status(p.status + t); } while (uint32_t(w >> 34) != epoch || (uint32_t(w >> 32) & 3u) == 0u);
                part += uint32_t(w);
            }
            part = __reduce_add_sync(0xFFFFFFFFu, part);
            if (lane == 0 && part != 0u) atomicAdd(&s_sum[ju & 3u], part);
        }
        __syncthreads();      // (B) warp counts of tile j, aggregate sum for tile j-2 visible; ring stage of tile j+1 consumed

        // ---- S4: place tile j's descriptors + publish its aggregate; write out the records of tile j-2 ------------------------------------
        if (valid) {
            uint32_t warpOff = 0u, total = 0u;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) { const uint32_t c = s_warpCnt[ju & 1u][w]; if (uint32_t(w) < warp) warpOff += c; total += c; }
            if (tid == 0) {
                st_status(p.status + curTile, pack_status(epoch, kStateAggregate, total));
                s_totals[ju & 3u] = total;
            }
            uint32_t* st = stageB + slot3 * TILE + warpOff;
            for (uint32_t i = lane; i < emitRun; i += 32u) st[i] = reinterpret_cast<const uint32_t*>(queue)[8u * i + 7u];
        }
        if (prev2 != kNoTile) {
            const uint32_t total = s_totals[(ju - uint32_t(kLagS)) & 3u];
            const uint32_t prefix = cum + s_sum[ju & 3u];                              // records before prev2
            if (prev2 == p.numTiles - 1u && tid == 0) {
                const uint32_t all = prefix + total;
                p.counts[0] = all < cap32 ? all : cap32;                             // the draw count the indirect draw reads
                p.counts[1] = all;
            }
            const uint32_t room = prefix < cap32 ? cap32 - prefix : 0u;
            const uint32_t nrec = room < total ? room : total;
            if (nrec != 0u) {
                const uint32_t slotOut = slot3 == 2u ? 0u : slot3 + 1u;               // (j - 2) mod 3 == (j + 1) mod 3
                const uint32_t* st = stageB + slotOut * TILE;
                uint2* dst = reinterpret_cast<uint2*>(p.draws + size_t(prefix) * p.recWords);
                const uint32_t idBase = p.objectIdBase + prev2 * uint32_t(TILE);
                // {objectId, indexCount} {instanceCount = 1, firstIndex} {vertexOffset = 0, firstInstance = 0} [{pad, pad}]
                if (p.recWords == 2u) {
                    // survivor-list form {objectId, absolute LOD id}: input of the instancing / cluster-expand kernels (cull_list.cu)
                    for (uint32_t w = tid; w < nrec; w += THREADS) { const uint32_t d = st[w]; st_rec_u2(dst + w, make_uint2(idBase + (d & kSLocalMask), d >> kSLocalBits)); }
                } else if (p.recWords == 6u) {
                    for (uint32_t w = tid; w < nrec * 3u; w += THREADS) {
                        const uint32_t r = w / 3u, f = w - r * 3u;
                        const uint32_t d = st[r];
                        const uint2 L = *reinterpret_cast<const uint2*>(&lodT[d >> kSLocalBits]);    // {indexCount, firstIndex}
                        st_rec_u2(dst + w, f == 0u ? make_uint2(idBase + (d & kSLocalMask), L.x) : (f == 1u ? make_uint2(1u, L.y) : make_uint2(0u, 0u)));
                        if (f == 2u && p.descs != nullptr) st_rec_u2(p.descs + prefix + r, make_uint2(idBase + (d & kSLocalMask), d >> kSLocalBits));
                    }
                } else {
                    for (uint32_t w = tid; w < nrec * 4u; w += THREADS) {
                        const uint32_t r = w >> 2, f = w & 3u;
                        const uint32_t d = st[r];
                        const uint2 L = *reinterpret_cast<const uint2*>(&lodT[d >> kSLocalBits]);
                        st_rec_u2(dst + w, f == 0u ? make_uint2(idBase + (d & kSLocalMask), L.x) : (f == 1u ? make_uint2(1u, L.y) : make_uint2(0u, 0u)));
                        if (f == 3u && p.descs != nullptr) st_rec_u2(p.descs + prefix + r, make_uint2(idBase + (d & kSLocalMask), d >> kSLocalBits));
                    }
                }
            }
            cum = prefix + total;
            nextRead = prev2 + 1u;
        }
        // s_sum[(j+2)&3] was last read in iteration j-2 and is next added to before the barrier of iteration j+2 (behind that of j+1)
        if (tid == 0) s_sum[(ju + 2u) & 3u] = 0u;
        prev2 = prev1; prev1 = curTile; curTile = nextTile;
        slot3 = slot3 == 2u ? 0u : slot3 + 1u;
        if (j >= 0 && curTile == kNoTile && prev1 == kNoTile && prev2 == kNoTile) break;
    }

    if (p.n == 0u && blockIdx.x == 0 && tid == 0) { p.counts[0] = 0u; p.counts[1] = 0u; }
    // last CTA out re-arms the control block for the next launch on this stream
    if (tid == 0) {
        __threadfence();
        const uint32_t prev = atomicAdd(&p.ctl->done, 1u);
        if (prev == gridDim.x - 1u) {
            uint32_t e = (epoch + 1u) & 0x3FFFFFFFu;
            p.ctl->epoch = e ? e : 1u;
            p.ctl->ticket = 0u;
            p.ctl->done = 0u;
        }
    }
}

template <int PASS, int HIZ, int THREADS, int ITEMS, int MINB>
cudaError_t launch_cfg(const DrawCullParams& p, int numSMs, cudaStream_t stream)
{
    constexpr int TILE = THREADS * ITEMS;
    constexpr size_t VIS_STAGE_BYTES = (PASS == PASS_EARLY) ? size_t(TILE) * 4 : ((PASS == PASS_LATE) ? size_t(TILE) / 8 : 0);
    if (p.lodCount >= (1u << (30 - kSLocalBits))) return cudaErrorInvalidValue;     // descriptor packing (checked by the C-ABI layer too)
    const size_t tableBytes = (size_t(p.surfaceCount) + p.lodCount) * 32u;
    const bool smemTables = tableBytes <= 8192u;
    const size_t smem = (smemTables ? tableBytes : 0u) + size_t(kStreamDepth) * VIS_STAGE_BYTES + size_t(TILE) * (size_t(kStreamDepth) * 8 + 32 + 4 * kStagesS);
    auto kernel = smemTables ? stream_cull_kernel<PASS, HIZ, THREADS, ITEMS, MINB, true> : stream_cull_kernel<PASS, HIZ, THREADS, ITEMS, MINB, false>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    int perSM = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, THREADS, smem);
    if (e != cudaSuccess) return e;
    if (perSM < 1) perSM = 1;
    DrawCullParams q = p;
    q.numTiles = uint32_t((uint64_t(p.n) + uint64_t(TILE) - 1) / uint64_t(TILE));
    uint32_t grid = uint32_t(numSMs) * uint32_t(perSM);
    if (grid > q.numTiles) grid = q.numTiles;
    if (grid < 1) grid = 1;
    kernel<<<grid, THREADS, smem, stream>>>(q);
    return cudaGetLastError();
}

template <int PASS, int HIZ>
cudaError_t launch_pass(const DrawCullParams& p, int cfg, int numSMs, cudaStream_t stream)
{
#ifdef BLZ_STREAM_MINIMAL       // developer A/B builds (scripts/build_variant.sh): one CTA shape, a fraction of the compile time
    return launch_cfg<PASS, HIZ, 256, 4, 3>(p, numSMs, stream);
#else
    switch (cfg) {
    case 3: return launch_cfg<PASS, HIZ, 224, 4, 3>(p, numSMs, stream);       // 7 warps: 96 registers per thread at 3 CTAs / SM
    case 4: return launch_cfg<PASS, HIZ, 192, 4, 4>(p, numSMs, stream);
    case 6: return launch_cfg<PASS, HIZ, 256, 3, 4>(p, numSMs, stream);
    default: return launch_cfg<PASS, HIZ, 256, 4, 3>(p, numSMs, stream);      // cfg 2
    }
#endif
}

} // namespace

cudaError_t launch_stream_cull(const DrawCullParams& p, int pass, int hiz, int cfg, int numSMs, cudaStream_t stream)
{
    switch (pass) {
    case PASS_FRUSTUM: return launch_pass<PASS_FRUSTUM, HIZ_NONE>(p, cfg, numSMs, stream);
    case PASS_EARLY: return launch_pass<PASS_EARLY, HIZ_NONE>(p, cfg, numSMs, stream);
    case PASS_LATE: return hiz == HIZ_VK ? launch_pass<PASS_LATE, HIZ_VK>(p, cfg, numSMs, stream) : launch_pass<PASS_LATE, HIZ_DX>(p, cfg, numSMs, stream);
    case PASS_TEMPORAL: return hiz == HIZ_VK ? launch_pass<PASS_TEMPORAL, HIZ_VK>(p, cfg, numSMs, stream) : launch_pass<PASS_TEMPORAL, HIZ_DX>(p, cfg, numSMs, stream);
    }
    return cudaErrorInvalidValue;
}

} // namespace blz
