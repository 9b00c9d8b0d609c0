// Draw-cull passes, streaming formulation (v5): frustum / early / late (Hi-Z) / temporal, replacing
//   VulkanShaders/{Initial,Late,Transparent,Onpc}DrawCull.comp.glsl and HlslShaders/CS/{drawCull,drawOccFirst,drawOccLate,drawOccTemporal}
// (paths relative to /root/reference/src/Renderer).
//
// The pipelined kernel of cull_draw.cu (v4) reached 52 % of the measured HBM peak and its ncu capture (profiles/r01e_*) showed it
// issue-bound: ~375 warp instructions per 32 objects, of which only ~70 are the FP32 arithmetic of the shader -- the rest is the
// software pipeline itself (cp.async address arithmetic, shared-memory round trips of every input, ring bookkeeping).  This
// kernel goes the other way: NO software pipeline.  One short-lived CTA per tile of consecutive objects, every input loaded
// straight into registers with all of a thread's loads in flight together, and memory latency hidden by occupancy (6-8 CTAs per
// SM at different points of their life) instead of by staging.  ~100 warp instructions per 32 objects on the streaming side.
//
//   * tile = blockIdx.x (CTAs are dispatched in blockIdx order, the assumption every single-pass scan makes), TILE = THREADS x ITEMS
//     consecutive objects; lane l of warp w owns objects w*32*ITEMS + l + 32k, so every global access of a warp is one contiguous
//     run of fully used sectors (4 B visibility, 8 B RenderObject, 2 x 16 B transform halves per lane);
//   * phase 1: RenderObject + visibility words (coalesced), phase 2: transform gather by transformId (two float4 streams);
//   * sphere + frustum planes in registers; the survivors (a few %) go to a CTA-wide queue in shared memory so that the expensive
//     tail (projectSphere: 2 sqrt + 5 IEEE div, Hi-Z fetch, LOD loop) runs with full warps instead of 3 active lanes of 32;
//   * compaction is deterministic: ballot/popc in the warp, scan of the warp counts in the CTA, single-pass decoupled look-back
//     across tiles (scan_lookback.cuh) examined THREADS predecessors at a time -- one-shot CTAs all finish at about the same time,
//     so a warp-wide window would walk back over every resident tile 32 at a time;
//   * survivors are staged as 4-B descriptors and leave the CTA as one contiguous span of 8-B stores (24-/32-B records).
// No ticket, no exit counter: the launch epoch that validates the per-tile status words is bumped by the last TILE (when its
// look-back has completed every other tile has published, hence started, hence read the epoch).
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

// Evaluates queue entry e: Hi-Z (late / temporal passes) and LOD selection for emitters; the result word
// (visible | emit << 1 | lodId << 2) goes to res[index in tile].
template <int PASS, int HIZ>
__device__ __forceinline__ void eval_entry(uint32_t e, const float4* qSphere, const uint2* qMeta, const uint32_t* qSurf, uint32_t* res,
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

template <int PASS, int HIZ, int THREADS, int ITEMS, int MINB, bool SMEM_TABLES>
__global__ void __launch_bounds__(THREADS, MINB) stream_cull_kernel(const __grid_constant__ DrawCullParams p)
{
    constexpr int TILE = THREADS * ITEMS, WARPS = THREADS / 32;
    constexpr bool HAS_VIS = (PASS == PASS_EARLY || PASS == PASS_LATE);
    static_assert(TILE <= (1 << kSLocalBits), "descriptor packing");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_warpCnt[WARPS];
    __shared__ uint32_t s_scratch[2 * WARPS + 2];
    __shared__ uint32_t s_qCount;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t laneLt = (1u << lane) - 1u;
    const uint32_t tile = blockIdx.x, tileBase = tile * uint32_t(TILE);
    const uint32_t localBase = warp * uint32_t(32 * ITEMS) + lane;     // + 32k = index inside the tile
    const ViewConsts& V = p.view;

    // ---- phase 1 loads: visibility + RenderObject -------------------------------------------------------------------------
    uint32_t actMask = 0u, inMask = 0u, visPrevMask = 0u;
    uint2 ob[ITEMS];
    if (HAS_VIS) {
        uint32_t vp[ITEMS];
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const uint32_t i = tileBase + localBase + uint32_t(k) * 32u;
            vp[k] = i < p.n ? ld_vis_u32(p.visibility + i) : 0u;
        }
        if (PASS == PASS_EARLY) {
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {                          // InitialDrawCull.comp.glsl:21-24: only last frame's visible objects
                const uint32_t i = tileBase + localBase + uint32_t(k) * 32u;
                const bool act = i < p.n && vp[k] != 0u;
                ob[k] = act ? ld_stream_u2(p.objs + i) : make_uint2(p.transformIdBase, 0u);
                actMask |= (act ? 1u : 0u) << k;
            }
        } else {
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const uint32_t i = tileBase + localBase + uint32_t(k) * 32u;
                const bool in = i < p.n;
                ob[k] = in ? ld_stream_u2(p.objs + i) : make_uint2(p.transformIdBase, 0u);
                actMask |= (in ? 1u : 0u) << k;
            }
        }
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) visPrevMask |= (vp[k] != 0u ? 1u : 0u) << k;
    } else {
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const uint32_t i = tileBase + localBase + uint32_t(k) * 32u;
            const bool in = i < p.n;
            ob[k] = in ? ld_stream_u2(p.objs + i) : make_uint2(p.transformIdBase, 0u);
            actMask |= (in ? 1u : 0u) << k;
        }
    }
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) inMask |= (tileBase + localBase + uint32_t(k) * 32u < p.n ? 1u : 0u) << k;
    const uint32_t epoch = ld_cg_u32(&p.ctl->epoch) & 0x3FFFFFFFu;      // consumed by the look-back, far below

    // ---- carve shared memory; surface + LOD tables (KB) ---------------------------------------------------------------------
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
    float4* qSphere = reinterpret_cast<float4*>(sp);     sp += size_t(TILE) * sizeof(float4);     // survivor queue: view-space sphere
    uint2* qMeta = reinterpret_cast<uint2*>(sp);         sp += size_t(TILE) * sizeof(uint2);      //   {scale bits, index in tile | visPrev << 16}
    uint32_t* qSurf = reinterpret_cast<uint32_t*>(sp);   sp += size_t(TILE) * sizeof(uint32_t);   //   surfaceId
    uint32_t* sRes = reinterpret_cast<uint32_t*>(sp);                                             // visible | emit << 1 | lodId << 2 per object of the tile
    uint32_t* stage = reinterpret_cast<uint32_t*>(qSphere);                                       // survivor descriptors (the queue is dead by then)
    if (tid == 0) s_qCount = 0u;

    // ---- phase 2 loads: transform gather -----------------------------------------------------------------------------------------
    float4 ps[ITEMS], qt[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        ps[k] = make_float4(0.f, 0.f, 0.f, 1.f); qt[k] = make_float4(0.f, 0.f, 0.f, 1.f);
        if ((actMask >> k) & 1u) {
            const uint32_t t = ob[k].x - p.transformIdBase;
            ps[k] = ld_stream_f4(p.xfPosScale + t);
            qt[k] = ld_stream_f4(p.xfQuat + t);
        }
    }
    __syncthreads();      // tables + queue counter visible

    // ---- sphere + frustum planes; survivors -> queue (one shared-memory atomic per warp) ---------------------------------------------
    uint32_t survMask = 0u;
    {
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
        uint32_t ball[ITEMS], cnt = 0u;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) { ball[k] = __ballot_sync(0xFFFFFFFFu, (survMask >> k) & 1u); cnt += uint32_t(__popc(ball[k])); }
        if (cnt != 0u) {
            uint32_t base = 0u;
            if (lane == 0) base = atomicAdd(&s_qCount, cnt);
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
    __syncthreads();      // queue complete

    // ---- queue evaluation with full warps: Hi-Z + LOD ----------------------------------------------------------------------------------
    {
        const uint32_t qn = s_qCount;
        for (uint32_t e = tid; e < qn; e += THREADS) eval_entry<PASS, HIZ>(e, qSphere, qMeta, qSurf, sRes, surfT, lodT, p);
    }
    __syncthreads();      // results complete; queue dead

    // ---- results back to their owners: visibility write, ranks --------------------------------------------------------------------------
    uint32_t emitMask = 0u, rank[ITEMS], lodSel[ITEMS], running = 0u;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const uint32_t l = localBase + uint32_t(k) * 32u;
        const uint32_t r = ((survMask >> k) & 1u) ? sRes[l] : 0u;
        if (PASS == PASS_LATE && ((inMask >> k) & 1u)) p.visibility[tileBase + l] = r & 1u;          // LateDrawCull.comp.glsl:70
        lodSel[k] = r >> 2;
        const bool emit = (r & 2u) != 0u;
        emitMask |= (emit ? 1u : 0u) << k;
        const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, emit);
        rank[k] = running + uint32_t(__popc(ballot & laneLt));
        running += uint32_t(__popc(ballot));
    }
    if (lane == 0) s_warpCnt[warp] = running;
    __syncthreads();      // warp counts visible

    uint32_t warpOff = 0u, total = 0u;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) { const uint32_t c = s_warpCnt[w]; if (uint32_t(w) < warp) warpOff += c; total += c; }
#pragma unroll
    for (int k = 0; k < ITEMS; ++k)
        if ((emitMask >> k) & 1u) stage[warpOff + rank[k]] = (localBase + uint32_t(k) * 32u) | (lodSel[k] << kSLocalBits);

    // ---- cross-tile offset (decoupled look-back, CTA-wide windows) + contiguous record span ------------------------------------------------
    const uint64_t prefix = lookback_exclusive_prefix_cta<THREADS>(p.status, tile, total, epoch, s_scratch);
    __syncthreads();      // descriptors visible (tile 0 returns from the look-back without a barrier)
    if (tile == gridDim.x - 1u && tid == 0) {
        const uint64_t all = prefix + total;
        p.counts[0] = uint32_t(all < p.capacity ? all : p.capacity);          // the draw count the indirect draw reads
        p.counts[1] = uint32_t(all);
        const uint32_t e = (epoch + 1u) & 0x3FFFFFFFu;                        // every other tile has published, hence read the epoch
        p.ctl->epoch = e ? e : 1u;
    }
    const uint64_t room = prefix < p.capacity ? p.capacity - prefix : 0ull;
    const uint32_t nrec = uint32_t(room < total ? room : total);
    if (nrec != 0u) {
        uint2* dst = reinterpret_cast<uint2*>(p.draws + prefix * p.recWords);
        const uint32_t idBase = p.objectIdBase + tileBase;
        // {objectId, indexCount} {instanceCount = 1, firstIndex} {vertexOffset = 0, firstInstance = 0} [{pad, pad}]
        if (p.recWords == 6u) {
            for (uint32_t w = tid; w < nrec * 3u; w += THREADS) {
                const uint32_t r = w / 3u, f = w - r * 3u;
                const uint32_t d = stage[r];
                const uint2 L = *reinterpret_cast<const uint2*>(&lodT[d >> kSLocalBits]);    // {indexCount, firstIndex}
                st_rec_u2(dst + w, f == 0u ? make_uint2(idBase + (d & kSLocalMask), L.x) : (f == 1u ? make_uint2(1u, L.y) : make_uint2(0u, 0u)));
            }
        } else {
            for (uint32_t w = tid; w < nrec * 4u; w += THREADS) {
                const uint32_t r = w >> 2, f = w & 3u;
                const uint32_t d = stage[r];
                const uint2 L = *reinterpret_cast<const uint2*>(&lodT[d >> kSLocalBits]);
                st_rec_u2(dst + w, f == 0u ? make_uint2(idBase + (d & kSLocalMask), L.x) : (f == 1u ? make_uint2(1u, L.y) : make_uint2(0u, 0u)));
            }
        }
    }
}

template <int PASS, int HIZ, int THREADS, int ITEMS, int MINB>
cudaError_t launch_cfg(const DrawCullParams& p, cudaStream_t stream)
{
    constexpr int TILE = THREADS * ITEMS;
    if (p.lodCount >= (1u << (30 - kSLocalBits))) return cudaErrorInvalidValue;     // descriptor packing (checked by the C-ABI layer too)
    const size_t tableBytes = (size_t(p.surfaceCount) + p.lodCount) * 32u;
    const bool smemTables = tableBytes <= 8192u;
    const size_t smem = (smemTables ? tableBytes : 0u) + size_t(TILE) * (16 + 8 + 4 + 4);
    auto kernel = smemTables ? stream_cull_kernel<PASS, HIZ, THREADS, ITEMS, MINB, true> : stream_cull_kernel<PASS, HIZ, THREADS, ITEMS, MINB, false>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    const uint32_t numTiles = p.n == 0u ? 1u : uint32_t((uint64_t(p.n) + uint64_t(TILE) - 1) / uint64_t(TILE));
    kernel<<<numTiles, THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

template <int PASS, int HIZ>
cudaError_t launch_pass(const DrawCullParams& p, int cfg, cudaStream_t stream)
{
    switch (cfg) {
    case 1: return launch_cfg<PASS, HIZ, 256, 2, 8>(p, stream);
    case 2: return launch_cfg<PASS, HIZ, 256, 4, 4>(p, stream);
    case 3: return launch_cfg<PASS, HIZ, 512, 2, 3>(p, stream);
    case 4: return launch_cfg<PASS, HIZ, 128, 4, 10>(p, stream);
    case 5: return launch_cfg<PASS, HIZ, 256, 4, 5>(p, stream);
    default: return launch_cfg<PASS, HIZ, 256, 2, 6>(p, stream);
    }
}

} // namespace

uint32_t stream_cull_min_tile() { return 512u; }

cudaError_t launch_stream_cull(const DrawCullParams& p, int pass, int hiz, int cfg, cudaStream_t stream)
{
    switch (pass) {
    case PASS_FRUSTUM: return launch_pass<PASS_FRUSTUM, HIZ_NONE>(p, cfg, stream);
    case PASS_EARLY: return launch_pass<PASS_EARLY, HIZ_NONE>(p, cfg, stream);
    case PASS_LATE: return hiz == HIZ_VK ? launch_pass<PASS_LATE, HIZ_VK>(p, cfg, stream) : launch_pass<PASS_LATE, HIZ_DX>(p, cfg, stream);
    case PASS_TEMPORAL: return hiz == HIZ_VK ? launch_pass<PASS_TEMPORAL, HIZ_VK>(p, cfg, stream) : launch_pass<PASS_TEMPORAL, HIZ_DX>(p, cfg, stream);
    }
    return cudaErrorInvalidValue;
}

} // namespace blz
