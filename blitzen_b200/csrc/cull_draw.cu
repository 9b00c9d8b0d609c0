// Draw-cull passes: frustum / early / late (Hi-Z) / temporal, one persistent software-pipelined kernel each, replacing
//   VulkanShaders/{Initial,Late,Transparent,Onpc}DrawCull.comp.glsl and HlslShaders/CS/{drawCull,drawOccFirst,drawOccLate,drawOccTemporal}
// (paths relative to /root/reference/src/Renderer).
//
// Mapping to the hardware (B200, 148 SMs, HBM-bound).  History of the measurements that shaped it is in DESIGN.md; the
// ncu captures are profiles/r01a_* (first version: 29 % of the HBM peak, 40-57 % of the stalls at CTA barriers and in the
// look-back spin) and profiles/r01b_* (per-warp pending loops: 240 M warp instructions, 3 of 32 lanes active in Hi-Z).
//
//   * persistent grid = numSMs x 2 CTAs of 512 threads; tiles of 1024 consecutive objects are handed out by an atomic
//     ticket claimed three tiles ahead, so tile order == claim order and every predecessor of a tile belongs to a running CTA.
//   * every input stream is staged through shared memory by cp.async (LDGSTS), issued by the thread that later consumes it
//     (no barrier on the input side): iteration j consumes tile j, then re-fills the slots it just read with the
//     RenderObject / visibility words of tile j+2 and the two 16-B transform halves of tile j+1 (gather by transformId,
//     which arrived one iteration earlier).  ~44 KB of loads per CTA are in flight whatever the CTA is doing.
//   * lane l of a warp owns objects l and l+32 of the warp's 64-object span: every global access of a warp is one
//     contiguous, fully used run of sectors (4 B, 8 B or 16 B per lane).
//   * surface + LOD tables (KB) live in shared memory; view constants are kernel parameters (constant bank).
//   * DENSE step (all objects): view-space sphere + frustum planes, ~70 FP32 instructions, nothing else.
//     Survivors (a few %) are pushed into a CTA-wide queue in shared memory.
//   * SPARSE step (queue entries, one per thread): projectSphere (2 sqrt + 5 IEEE div), Hi-Z fetch, LOD loop.  Runs with
//     full warps instead of 3 active lanes out of 32.
//   * compaction is deterministic and chain-free: ballot/popc inside the warp, a 16-entry scan inside the CTA, and across
//     tiles every CTA sums the published per-tile AGGREGATES between its previous tile and the current one (coalesced
//     reads of 8-B epoch-tagged words, L2 hits).  No tile ever waits for another tile's prefix, only for aggregates, and
//     the sum is taken two tiles late, so the wait is not exposed.
//   * survivors are staged as 4-B descriptors {index in tile, lodId}; the 24-/32-B records are expanded on the way out
//     and leave the CTA as one contiguous, coalesced span of 8-B stores.
#include "cull_kernels.cuh"
#include "cull_math.cuh"
#include "scan_lookback.cuh"

namespace blz {

namespace {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async4(void* dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_cs_u32(uint32_t* p, uint32_t v) { asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_cs_u2(void* p, uint2 v) { asm volatile("st.global.cs.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory"); }

constexpr int kTileRing = 8;             // ring of claimed tile ids (power of two, > prefetch distance + 2)
constexpr uint32_t kNoTile = 0xFFFFFFFFu;
constexpr int kLag = 2;                  // the prefix of a tile is resolved (and its records written) this many tiles later
constexpr int kStages = kLag + 1;        // descriptor buffers
constexpr uint32_t kLocalBits = 10;      // index in tile
constexpr uint32_t kLocalMask = (1u << kLocalBits) - 1u;
static_assert(kDrawTile <= (1 << kLocalBits), "descriptor packing");

} // namespace

// Evaluates queue entry e of a tile: Hi-Z (late / temporal passes) and LOD selection for emitters; the result word
// (visible | emit << 1 | lodId << 2) goes to res[index in tile].
template <int PASS, int HIZ>
__device__ __forceinline__ void sparse_eval(uint32_t e, const float4* qSphere, const uint2* qMeta, const uint32_t* qSurf, uint32_t* res,
                                            const PrimitiveSurface* surfT, const LodData* lodT, const DrawCullParams& p)
{
    constexpr bool HAS_HIZ = (PASS == PASS_LATE || PASS == PASS_TEMPORAL);
    const ViewConsts& V = p.view;
    const float4 q = qSphere[e];
    const uint2 m = qMeta[e];
    const Sphere s{ q.x, q.y, q.z, q.w };
    bool visible = true;
    if (HAS_HIZ) {
        float4 aabb;
        if (project_sphere(s, V.zNear, V.proj0, V.proj5, aabb))
            visible = (HIZ == HIZ_VK) ? hiz_test_vk(aabb, p.pyr, s, V) : hiz_test_dx(aabb, p.pyr, s, V);
    }
    bool emit = visible;
    if (PASS == PASS_LATE) emit = visible && ((m.y >> 16) == 0u);                  // LateDrawCull.comp.glsl:49
    uint32_t lodId = 0u;
    if (emit) {
        const uint32_t sidx = qSurf[e];
        const uint32_t lodOffset = surfT[sidx].lodOffset, lodCount = surfT[sidx].lodCount;
        const uint32_t rel = lod_select(s, __uint_as_float(m.x), V.lodTarget, lodOffset, lodCount, [&](uint32_t li) { return lodT[li].error; });
        lodId = (p.flags & kFlagOnpcLodQuirk) ? rel : rel + lodOffset;
    }
    res[m.y & 0xFFFFu] = (visible ? 1u : 0u) | (emit ? 2u : 0u) | (lodId << 2);
}

template <int PASS, int HIZ, bool SMEM_TABLES>
__global__ void __launch_bounds__(kDrawThreads, 2) draw_cull_kernel(const __grid_constant__ DrawCullParams p)
{
    constexpr int ITEMS = kDrawItems, TILE = kDrawTile, THREADS = kDrawThreads, DWARPS = kDrawDenseWarps, DTHREADS = DWARPS * 32;
    constexpr bool HAS_VIS = (PASS == PASS_EARLY || PASS == PASS_LATE);
    constexpr bool VIS_LIST = (PASS == PASS_LATE);       // the late pass also emits the ascending list of visible ids: next frame's early pass
    constexpr int DV = (PASS == PASS_EARLY) ? 3 : 2;     // prefetch distance of the visibility words (the early pass needs them to ask for objects)
    constexpr int NV = DV;                                // ring depth: the slot of tile j is re-filled with tile j+DV right after it is read

    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_tiles[kTileRing];
    __shared__ uint32_t s_warpCnt[DWARPS], s_warpCntV[DWARPS];
    __shared__ uint32_t s_totals[4];       // emit count | visible count << 16 of the tile finished in iteration j, at [j & 3]
    __shared__ uint32_t s_sum[2], s_sumV[2];   // sums of the aggregates (emitters / visible) between this CTA's consecutive tiles, at [j & 1]
    __shared__ uint32_t s_qCount[2];       // entries pushed into queue [j & 1] by the dense step of iteration j
    __shared__ uint32_t s_qHead[2];        // next unclaimed batch of that queue

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t laneLt = (1u << lane) - 1u;
    const bool dense = warp < uint32_t(DWARPS);

    // ---- carve shared memory --------------------------------------------------------------------------------------------
    unsigned char* sp = smem_raw;
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
        sp = reinterpret_cast<unsigned char*>(dstL + p.lodCount * 2u);
    }
    float4* xfPS = reinterpret_cast<float4*>(sp);          sp += size_t(TILE) * sizeof(float4);        // transforms of the tile about to be consumed
    float4* xfQ = reinterpret_cast<float4*>(sp);           sp += size_t(TILE) * sizeof(float4);
    float4* qSphereB = reinterpret_cast<float4*>(sp);      sp += size_t(2) * TILE * sizeof(float4);    // survivor queues, double buffered
    uint2* qMetaB = reinterpret_cast<uint2*>(sp);          sp += size_t(2) * TILE * sizeof(uint2);     //   {scale bits, index in tile | visPrev << 16}
    uint2* objRing = reinterpret_cast<uint2*>(sp);         sp += size_t(2) * TILE * sizeof(uint2);
    uint32_t* qSurfB = reinterpret_cast<uint32_t*>(sp);    sp += size_t(2) * TILE * sizeof(uint32_t);  //   surfaceId
    uint32_t* sResB = reinterpret_cast<uint32_t*>(sp);     sp += size_t(2) * TILE * sizeof(uint32_t);  // visible | emit << 1 | lodId << 2, per object of the tile
    uint32_t* stage = reinterpret_cast<uint32_t*>(sp);     sp += size_t(kStages) * TILE * sizeof(uint32_t);
    uint32_t* visRing = reinterpret_cast<uint32_t*>(sp);   sp += HAS_VIS ? size_t(NV) * TILE * sizeof(uint32_t) : 0;
    uint16_t* stageV = reinterpret_cast<uint16_t*>(sp);    // kStages * TILE index-in-tile of the visible objects (late pass only)

    const uint32_t epoch = ld_cg_u32(&p.ctl->epoch) & 0x3FFFFFFFu;   // constant for the whole launch (the last CTA out bumps it)
    const ViewConsts& V = p.view;
    const uint32_t localBase = warp * uint32_t(32 * ITEMS) + lane;    // + k*32 = index inside the tile (dense warps only)
    const uint32_t nLast = p.n ? p.n - 1u : 0u;

    if (tid == 0) {
        s_tiles[0] = atomicAdd(&p.ctl->ticket, 1u);
        s_qCount[0] = s_qCount[1] = 0u; s_qHead[0] = s_qHead[1] = 0u; s_sum[0] = s_sum[1] = 0u; s_sumV[0] = s_sumV[1] = 0u;
    }
    __syncthreads();     // first tile + tables visible

    uint64_t cum = 0, cumV = 0;        // records emitted by / visible objects of tiles [0, nextRead)
    uint32_t nextRead = 0;             // first tile whose aggregate this CTA has not summed yet
    uint32_t histTile[kLag + 1];       // [0] = tile of iteration j-1 (its queue is evaluated during iteration j-1 .. j), [kLag] = tile whose records go out now
#pragma unroll
    for (int h = 0; h <= kLag; ++h) histTile[h] = kNoTile;
    uint32_t slotV = 0u, slotS = 0u;   // ring slots of iteration j: visibility (mod NV), descriptors (mod kStages)
    uint32_t prevSurvMask = 0u;        // frustum survivors of this thread in the previous tile

    for (int j = -DV;; ++j) {
        const uint32_t ju = uint32_t(j + 2 * kTileRing * kStages * NV);   // j shifted to a non-negative value with the same residues
        const uint32_t tile = j >= 0 ? s_tiles[ju & (kTileRing - 1)] : kNoTile;
        const bool valid = tile < p.numTiles;          // uniform over the CTA
        const uint32_t qb = ju & 1u;                    // queue / result buffer of this iteration's tile
        float4* qSphere = qSphereB + qb * TILE; uint2* qMeta = qMetaB + qb * TILE; uint32_t* qSurf = qSurfB + qb * TILE;
        uint32_t survMask = 0u;
        uint32_t ticket = kNoTile;

        if (dense) {
            // thread 0 claims the ticket of sequence slot j+DV+1 now and publishes it behind the barrier (latency hidden); it is
            // read for the first time in iteration j+1.  Once a claimed tile is past the end every later one is too.
            if (tid == 0 && s_tiles[(ju + uint32_t(DV)) & (kTileRing - 1)] < p.numTiles) ticket = atomicAdd(&p.ctl->ticket, 1u);

            // ---- D1: (a) pull this thread's inputs of tile j out of shared memory, (b) immediately re-fill those slots with the
            //      loads of tiles j+1 / j+2 (a full iteration of lead time), (c) sphere + frustum; survivors -> queue [j & 1] ----
            cp_async_wait_all();               // everything this thread asked for one iteration ago has landed
            uint32_t visPrevMask = 0u, actMask = 0u;
            uint2 ob[ITEMS]; float4 ps[ITEMS], qt[ITEMS];
            if (valid) {
                const uint32_t tileBase = tile * uint32_t(TILE);
#pragma unroll
                for (int k = 0; k < ITEMS; ++k) {
                    const uint32_t l = localBase + uint32_t(k) * 32u, i = tileBase + l;
                    bool act = i < p.n;
                    if (HAS_VIS) {
                        const uint32_t vp = visRing[slotV * TILE + l];
                        if (act && vp != 0u) visPrevMask |= 1u << k;
                        if (PASS == PASS_EARLY) act = act && (vp != 0u);                     // InitialDrawCull.comp.glsl:21-24
                    }
                    if (act) actMask |= 1u << k;
                    ob[k] = objRing[(ju & 1u) * TILE + l];
                    ps[k] = xfPS[l]; qt[k] = xfQ[l];
                }
            }
            {
                const uint32_t slotV1 = slotV + 1u == uint32_t(NV) ? 0u : slotV + 1u;                    // tile j+1
                const uint32_t slotV2 = slotV1 + 1u == uint32_t(NV) ? 0u : slotV1 + 1u;                  // tile j+2
                const uint32_t tN = j + 1 >= 0 ? s_tiles[(ju + 1u) & (kTileRing - 1)] : kNoTile;
                if (tN < p.numTiles) {
#pragma unroll
                    for (int k = 0; k < ITEMS; ++k) {
                        const uint32_t l = localBase + uint32_t(k) * 32u;
                        // early pass: only the objects that were visible last frame (a clamped visibility word of the ragged tail belongs to another object)
                        if (PASS == PASS_EARLY && (visRing[slotV1 * TILE + l] == 0u || tN * uint32_t(TILE) + l >= p.n)) continue;
                        const uint32_t t = objRing[((ju + 1u) & 1u) * TILE + l].x - p.transformIdBase;
                        cp_async16(xfPS + l, p.xfPosScale + t);
                        cp_async16(xfQ + l, p.xfQuat + t);
                    }
                }
                const uint32_t tO = j + 2 >= 0 ? s_tiles[(ju + 2u) & (kTileRing - 1)] : kNoTile;
                if (tO < p.numTiles) {
#pragma unroll
                    for (int k = 0; k < ITEMS; ++k) {
                        const uint32_t l = localBase + uint32_t(k) * 32u;
                        if (PASS == PASS_EARLY && (visRing[slotV2 * TILE + l] == 0u || tO * uint32_t(TILE) + l >= p.n)) continue;
                        cp_async8(objRing + (ju & 1u) * TILE + l, p.objs + min(tO * uint32_t(TILE) + l, nLast));
                    }
                }
                if (HAS_VIS) {
                    const uint32_t tV = s_tiles[(ju + uint32_t(DV)) & (kTileRing - 1)];
                    if (tV < p.numTiles) {
#pragma unroll
                        for (int k = 0; k < ITEMS; ++k) {
                            const uint32_t l = localBase + uint32_t(k) * 32u;
                            cp_async4(visRing + slotV * TILE + l, p.visibility + min(tV * uint32_t(TILE) + l, nLast));
                        }
                    }
                }
                cp_async_commit();
            }
            if (valid) {
                Sphere sph[ITEMS];
#pragma unroll
                for (int k = 0; k < ITEMS; ++k) {
                    sph[k] = Sphere{ 0.f, 0.f, 0.f, 0.f };
                    if ((actMask >> k) & 1u) {
                        const float4 bs = *reinterpret_cast<const float4*>(&surfT[ob[k].y]);   // {center.xyz, radius}
                        sph[k] = view_space_sphere(bs.x, bs.y, bs.z, bs.w, ps[k].x, ps[k].y, ps[k].z, ps[k].w, qt[k].x, qt[k].y, qt[k].z, qt[k].w, V);
                        if (frustum_test(sph[k], V)) survMask |= 1u << k;
                    }
                }
                // queue push: one shared-memory atomic per warp
                uint32_t ball[ITEMS], cnt = 0u;
#pragma unroll
                for (int k = 0; k < ITEMS; ++k) { ball[k] = __ballot_sync(0xFFFFFFFFu, (survMask >> k) & 1u); cnt += uint32_t(__popc(ball[k])); }
                if (cnt != 0u) {
                    uint32_t base = 0u;
                    if (lane == 0) base = atomicAdd(&s_qCount[qb], cnt);
                    base = __shfl_sync(0xFFFFFFFFu, base, 0);
#pragma unroll
                    for (int k = 0; k < ITEMS; ++k) {
                        if ((survMask >> k) & 1u) {
                            const uint32_t slot = base + uint32_t(__popc(ball[k] & laneLt));
                            qSphere[slot] = make_float4(sph[k].x, sph[k].y, sph[k].z, sph[k].r);
                            qMeta[slot] = make_uint2(__float_as_uint(ps[k].w), (localBase + uint32_t(k) * 32u) | (((visPrevMask >> k) & 1u) << 16));
                            qSurf[slot] = ob[k].y;
                        }
                        base += uint32_t(__popc(ball[k]));
                    }
                }
            }
        }
        __syncthreads();   // (X) queue [j & 1] complete; results of tile j-1 (buffer [(j-1) & 1]) complete

        const uint32_t qn = s_qCount[qb];
        if (dense) {
            // ---- D2: finish tile j-1: read its results back, write visibility, rank the emitters; sum the aggregates for the lagging tile ----
            const uint32_t postTile = histTile[0], outTile = histTile[kLag];
            // thread 0 publishes the ticket it claimed at the top (first read in iteration j+1, behind X) and re-arms the counters
            // of the queue that iteration j+1 fills (drained before this X, pushed to after the dense-warp barrier below)
            if (tid == 0) { s_qCount[qb ^ 1u] = 0u; s_qHead[qb ^ 1u] = 0u; s_tiles[(ju + uint32_t(DV) + 1u) & (kTileRing - 1)] = ticket; }
            uint32_t emitMask = 0u, rank[ITEMS], lodSel[ITEMS], running = 0u;
            uint32_t visMask = 0u, rankV[ITEMS], runningV = 0u;
            if (postTile != kNoTile) {
                const uint32_t* res = sResB + (qb ^ 1u) * TILE;
#pragma unroll
                for (int k = 0; k < ITEMS; ++k) {
                    const uint32_t l = localBase + uint32_t(k) * 32u, i = postTile * uint32_t(TILE) + l;
                    const uint32_t r = ((prevSurvMask >> k) & 1u) ? res[l] : 0u;
                    if (PASS == PASS_LATE && i < p.n) st_cs_u32(p.visibility + i, r & 1u);          // LateDrawCull.comp.glsl:70
                    lodSel[k] = r >> 2;
                    const bool emit = (r & 2u) != 0u;
                    emitMask |= (emit ? 1u : 0u) << k;
                    const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, emit);
                    rank[k] = running + uint32_t(__popc(ballot & laneLt));
                    running += uint32_t(__popc(ballot));
                    if (VIS_LIST) {
                        const uint32_t bv = __ballot_sync(0xFFFFFFFFu, (r & 1u) != 0u);
                        visMask |= (r & 1u) << k;
                        rankV[k] = runningV + uint32_t(__popc(bv & laneLt));
                        runningV += uint32_t(__popc(bv));
                    }
                }
                if (lane == 0) { s_warpCnt[warp] = running; if (VIS_LIST) s_warpCntV[warp] = runningV; }
            }
            if (outTile != kNoTile) {
                uint32_t part = 0u, partV = 0u;                     // a tile's aggregate word: emitters | visible << 16 (both <= TILE)
                for (uint32_t t = nextRead + tid; t < outTile; t += DTHREADS) {
                    uint64_t w;
                    do { w = ld_status(p.status + t); } while (uint32_t(w >> 34) != epoch || (uint32_t(w >> 32) & 3u) == 0u);
                    part += uint32_t(w) & 0xFFFFu;
                    if (VIS_LIST) partV += uint32_t(w) >> 16;
                }
                part = __reduce_add_sync(0xFFFFFFFFu, part);
                if (lane == 0 && part != 0u) atomicAdd(&s_sum[ju & 1u], part);
                if (VIS_LIST) {
                    partV = __reduce_add_sync(0xFFFFFFFFu, partV);
                    if (lane == 0 && partV != 0u) atomicAdd(&s_sumV[ju & 1u], partV);
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(DTHREADS) : "memory");   // dense warps only: warp counts + aggregate sum visible

            // ---- D3: stage tile j-1's descriptors + publish its aggregate; write out the records of the tile finished kLag iterations ago ----
            if (postTile != kNoTile) {
                if (warp == 0) {
                    uint32_t tileTotal = __reduce_add_sync(0xFFFFFFFFu, lane < uint32_t(DWARPS) ? s_warpCnt[lane] : 0u);
                    if (VIS_LIST) tileTotal |= __reduce_add_sync(0xFFFFFFFFu, lane < uint32_t(DWARPS) ? s_warpCntV[lane] : 0u) << 16;
                    if (lane == 0) {
                        st_status(p.status + postTile, pack_status(epoch, kStateAggregate, tileTotal));
                        s_totals[ju & 3u] = tileTotal;
                    }
                }
                if (VIS_LIST && visMask != 0u) {
                    uint32_t warpOffV = 0u;
                    for (uint32_t w = 0; w < warp; ++w) warpOffV += s_warpCntV[w];
                    uint16_t* sv = stageV + slotS * TILE;
#pragma unroll
                    for (int k = 0; k < ITEMS; ++k)
                        if ((visMask >> k) & 1u) sv[warpOffV + rankV[k]] = uint16_t(localBase + uint32_t(k) * 32u);
                }
                if (emitMask != 0u) {
                    uint32_t warpOff = 0u;
                    for (uint32_t w = 0; w < warp; ++w) warpOff += s_warpCnt[w];
                    uint32_t* st = stage + slotS * TILE;
#pragma unroll
                    for (int k = 0; k < ITEMS; ++k)
                        if ((emitMask >> k) & 1u) st[warpOff + rank[k]] = (localBase + uint32_t(k) * 32u) | (lodSel[k] << kLocalBits);
                }
            }
            if (outTile != kNoTile) {
                const uint32_t outPacked = s_totals[(ju - uint32_t(kLag)) & 3u];
                const uint32_t outTotal = outPacked & 0xFFFFu, outVis = outPacked >> 16;
                const uint64_t prefix = cum + s_sum[ju & 1u];                                       // records before outTile
                const uint64_t room = prefix < p.capacity ? p.capacity - prefix : 0ull;
                const uint32_t nrec = uint32_t(room < outTotal ? room : outTotal);
                const uint32_t slotOut = slotS + uint32_t(kStages - kLag) >= uint32_t(kStages) ? slotS + uint32_t(kStages - kLag) - uint32_t(kStages) : slotS + uint32_t(kStages - kLag);
                if (VIS_LIST) {
                    const uint64_t prefixV = cumV + s_sumV[ju & 1u];                                // visible objects before outTile
                    if (p.visList != nullptr) {
                        const uint16_t* sv = stageV + slotOut * TILE;
                        const uint32_t idBaseV = outTile * uint32_t(TILE);                          // LOCAL object index (what the early pass indexes with)
                        for (uint32_t w = tid; w < outVis; w += DTHREADS) p.visList[prefixV + w] = idBaseV + sv[w];
                        if (outTile == p.numTiles - 1u && tid == 0) *p.visCount = uint32_t(prefixV + outVis);
                    }
                    cumV = prefixV + outVis;
                }
                if (nrec != 0u) {
                    const uint32_t* st = stage + slotOut * TILE;                                    // slot of iteration j - kLag
                    uint2* dst = reinterpret_cast<uint2*>(p.draws + prefix * p.recWords);
                    const uint32_t idBase = p.objectIdBase + outTile * uint32_t(TILE);
                    // {objectId, indexCount} {instanceCount = 1, firstIndex} {vertexOffset = 0, firstInstance = 0} [{pad, pad}]
                    if (p.recWords == 6u) {
                        for (uint32_t w = tid; w < nrec * 3u; w += DTHREADS) {
                            const uint32_t r = w / 3u, f = w - r * 3u;
                            const uint32_t d = st[r];
                            const uint2 L = *reinterpret_cast<const uint2*>(&lodT[d >> kLocalBits]);    // {indexCount, firstIndex}
                            st_cs_u2(dst + w, f == 0u ? make_uint2(idBase + (d & kLocalMask), L.x) : (f == 1u ? make_uint2(1u, L.y) : make_uint2(0u, 0u)));
                        }
                    } else {
                        for (uint32_t w = tid; w < nrec * 4u; w += DTHREADS) {
                            const uint32_t r = w >> 2, f = w & 3u;
                            const uint32_t d = st[r];
                            const uint2 L = *reinterpret_cast<const uint2*>(&lodT[d >> kLocalBits]);
                            st_cs_u2(dst + w, f == 0u ? make_uint2(idBase + (d & kLocalMask), L.x) : (f == 1u ? make_uint2(1u, L.y) : make_uint2(0u, 0u)));
                        }
                    }
                }
                if (outTile == p.numTiles - 1u && tid == 0) {                                      // the draw count the indirect draw reads
                    const uint64_t total = prefix + outTotal;
                    p.counts[0] = uint32_t(total < p.capacity ? total : p.capacity);
                    p.counts[1] = uint32_t(total);
                }
                cum = prefix + outTotal;
                nextRead = outTile + 1u;
            }
            // s_sum[(j+1)&1] was last read in D3 of iteration j-1 (before this iteration's X) and is next added to in D2 of iteration j+1 (behind its X)
            if (tid == 0) { s_sum[(ju + 1u) & 1u] = 0u; s_sumV[(ju + 1u) & 1u] = 0u; }
        }
        // ---- S / D4: evaluate the queue of tile j in batches of 32 entries, full warps.  The sparse warps start right behind X, so with a
        //      handful of survivors per tile this overlaps D2/D3 and the dense step of tile j+1; dense warps join when there is a backlog. ----
        if (qn != 0u) {
            uint32_t* res = sResB + qb * TILE;
            while (true) {
                uint32_t b = 0u;
                if (lane == 0) b = atomicAdd(&s_qHead[qb], 32u);
                b = __shfl_sync(0xFFFFFFFFu, b, 0);
                if (b >= qn) break;
                if (b + lane < qn) sparse_eval<PASS, HIZ>(b + lane, qSphere, qMeta, qSurf, res, surfT, lodT, p);
            }
        }
        // bookkeeping for the next iteration
#pragma unroll
        for (int h = kLag; h > 0; --h) histTile[h] = histTile[h - 1];
        histTile[0] = valid ? tile : kNoTile;
        prevSurvMask = survMask;
        slotV = slotV + 1u == uint32_t(NV) ? 0u : slotV + 1u;
        slotS = slotS + 1u == uint32_t(kStages) ? 0u : slotS + 1u;
        if (j >= 0 && !valid) {
            bool pending = false;
#pragma unroll
            for (int h = 0; h <= kLag; ++h) pending = pending || (histTile[h] != kNoTile);
            if (!pending) break;
        }
    }

    cp_async_wait_all();
    if (p.n == 0u && blockIdx.x == 0 && tid == 0) { p.counts[0] = 0u; p.counts[1] = 0u; if (VIS_LIST && p.visCount) *p.visCount = 0u; }
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

static size_t draw_smem_bytes(int pass, bool smemTables, const DrawCullParams& p)
{
    const size_t tile = size_t(kDrawTile);
    const bool hasVis = pass == PASS_EARLY || pass == PASS_LATE;
    const size_t nv = pass == PASS_EARLY ? 3 : 2;
    size_t b = smemTables ? (size_t(p.surfaceCount) + p.lodCount) * 32u : 0u;
    b += tile * 16 * 2;                // transforms (two float4 streams, single buffer)
    b += 2 * tile * (16 + 8 + 4);      // survivor queues: sphere, meta, surfaceId
    b += 2 * tile * 8;                 // RenderObject ring
    b += 2 * tile * 4;                 // per-object results of the sparse step
    b += size_t(kStages) * tile * 4;   // survivor descriptors
    if (hasVis) b += nv * tile * 4;
    if (pass == PASS_LATE) b += size_t(kStages) * tile * 2;   // visible-id descriptors
    return b;
}

template <int PASS, int HIZ>
static cudaError_t launch_variant(const DrawCullParams& p, int numSMs, cudaStream_t stream)
{
    if (p.lodCount >= (1u << (32 - kLocalBits))) return cudaErrorInvalidValue;     // descriptor packing (checked by the C-ABI layer too)
    const size_t tableBytes = (size_t(p.surfaceCount) + p.lodCount) * 32u;
    const bool smemTables = tableBytes <= 8192u;
    const size_t smem = draw_smem_bytes(PASS, smemTables, p);
    auto kernel = smemTables ? draw_cull_kernel<PASS, HIZ, true> : draw_cull_kernel<PASS, HIZ, false>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    int perSM = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, kDrawThreads, smem);
    if (e != cudaSuccess) return e;
    if (perSM < 1) perSM = 1;
    DrawCullParams q = p;
    q.numTiles = uint32_t((uint64_t(p.n) + uint64_t(kDrawTile) - 1) / uint64_t(kDrawTile));
    uint32_t grid = uint32_t(numSMs) * uint32_t(perSM);
    if (grid > q.numTiles) grid = q.numTiles;
    if (grid < 1) grid = 1;
    kernel<<<grid, kDrawThreads, smem, stream>>>(q);
    return cudaGetLastError();
}

cudaError_t launch_draw_cull(const DrawCullParams& p, int pass, int hiz, int numSMs, cudaStream_t stream)
{
    switch (pass) {
    case PASS_FRUSTUM: return launch_variant<PASS_FRUSTUM, HIZ_NONE>(p, numSMs, stream);
    case PASS_EARLY: return launch_variant<PASS_EARLY, HIZ_NONE>(p, numSMs, stream);
    case PASS_LATE: return hiz == HIZ_VK ? launch_variant<PASS_LATE, HIZ_VK>(p, numSMs, stream) : launch_variant<PASS_LATE, HIZ_DX>(p, numSMs, stream);
    case PASS_TEMPORAL: return hiz == HIZ_VK ? launch_variant<PASS_TEMPORAL, HIZ_VK>(p, numSMs, stream) : launch_variant<PASS_TEMPORAL, HIZ_DX>(p, numSMs, stream);
    }
    return cudaErrorInvalidValue;
}

// One-time SoA repack of the reference's AoS MeshTransform array (32 B: pos.xyz, scale, quat) into two float4 streams.
__global__ void repack_transforms_kernel(const float4* __restrict__ aos, float4* __restrict__ posScale, float4* __restrict__ quat, uint32_t first, uint32_t count)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    posScale[first + i] = aos[2 * size_t(i)];
    quat[first + i] = aos[2 * size_t(i) + 1];
}

cudaError_t launch_repack_transforms(const MeshTransform* aos, float4* posScale, float4* quat, uint32_t first, uint32_t count, cudaStream_t stream)
{
    if (count == 0) return cudaSuccess;
    repack_transforms_kernel<<<(count + 255u) / 256u, 256, 0, stream>>>(reinterpret_cast<const float4*>(aos), posScale, quat, first, count);
    return cudaGetLastError();
}

} // namespace blz
