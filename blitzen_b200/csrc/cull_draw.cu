// Draw-cull passes: frustum / early / late (Hi-Z) / temporal, one persistent kernel each, replacing
//   VulkanShaders/{Initial,Late,Transparent,Onpc}DrawCull.comp.glsl and HlslShaders/CS/{drawCull,drawOccFirst,drawOccLate,drawOccTemporal}
// (paths relative to /root/reference/src/Renderer).
//
// Mapping to the hardware (B200, 148 SMs, HBM-bound):
//   * persistent grid = numSMs x resident CTAs; tiles of kCullTile consecutive objects are handed out by an atomic ticket,
//     so tile order == object order and every predecessor of a tile is already running (look-back cannot deadlock).
//   * a warp owns 32*ITEMS consecutive objects; lane l handles objects l, l+32, ... -> every global access of a warp is one
//     contiguous, fully used span: RenderObject 8 B/lane (LDG.64), visibility 4 B/lane, transforms 2 x 16 B/lane (LDG.128)
//     from the SoA repack.  All loads of a tile are issued before the first use (ITEMS*3 independent requests per thread).
//   * the surface and LOD tables (a few KB) are copied to shared memory once per CTA.
//   * survivors are ranked with ballot/popc inside the warp, by an 8-entry shared scan inside the CTA and by a decoupled
//     look-back across tiles; records are staged in shared memory and leave the CTA as one contiguous, coalesced span.
//   * view constants arrive as kernel parameters (constant bank operands), not loads.
#include "cull_kernels.cuh"
#include "cull_math.cuh"
#include "scan_lookback.cuh"

namespace blz {

__device__ __forceinline__ uint2 ldg_nc_u2(const void* p)
{
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ldg_nc_f4(const void* p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_cs_u32(uint32_t* p, uint32_t v) { asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_cs_u2(void* p, uint2 v) { asm volatile("st.global.cs.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory"); }

// Surface / LOD table access, either from the shared-memory copy or straight from global (tables too large for smem).
struct Tables {
    const PrimitiveSurface* surf;
    const LodData* lod;
};

template <int PASS, int HIZ, bool SMEM_TABLES, int ITEMS>
__global__ void __launch_bounds__(kCullThreads, 4) draw_cull_kernel(const __grid_constant__ DrawCullParams p)
{
    constexpr int TILE = kCullThreads * ITEMS;
    constexpr int WARPS = kCullThreads / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_warpCnt[WARPS];
    __shared__ uint32_t s_tile, s_prefix;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t laneLt = (1u << lane) - 1u;

    Tables T;
    uint32_t* staging;
    if (SMEM_TABLES) {
        uint4* dstS = reinterpret_cast<uint4*>(smem_raw);
        const uint4* srcS = reinterpret_cast<const uint4*>(p.surfaces);
        for (uint32_t i = tid; i < p.surfaceCount * 2u; i += kCullThreads) dstS[i] = __ldg(srcS + i);
        uint4* dstL = dstS + p.surfaceCount * 2u;
        const uint4* srcL = reinterpret_cast<const uint4*>(p.lods);
        for (uint32_t i = tid; i < p.lodCount * 2u; i += kCullThreads) dstL[i] = __ldg(srcL + i);
        T.surf = reinterpret_cast<const PrimitiveSurface*>(dstS);
        T.lod = reinterpret_cast<const LodData*>(dstL);
        staging = reinterpret_cast<uint32_t*>(dstL + p.lodCount * 2u);
    } else {
        T.surf = p.surfaces;
        T.lod = p.lods;
        staging = reinterpret_cast<uint32_t*>(smem_raw);
    }
    const uint32_t epoch = ld_cg_u32(&p.ctl->epoch);   // constant for the whole launch (only the last CTA to leave bumps it)
    const ViewConsts& V = p.view;

    while (true) {
        if (tid == 0) s_tile = atomicAdd(&p.ctl->ticket, 1u);
        __syncthreads();   // (A) ticket visible; previous tile's staging fully drained; tables loaded
        const uint32_t tile = s_tile;
        if (tile >= p.numTiles) break;
        const uint32_t base = tile * uint32_t(TILE) + warp * uint32_t(32 * ITEMS) + lane;

        // ---- phase 1: issue every load of the tile -------------------------------------------------------------
        bool act[ITEMS];
        uint2 obj[ITEMS];
        uint32_t visPrev[ITEMS];
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const uint32_t i = base + uint32_t(k) * 32u;
            act[k] = i < p.n;
            visPrev[k] = 0u;
            if (PASS == PASS_EARLY || PASS == PASS_LATE) {
                if (act[k]) visPrev[k] = ld_cg_u32(p.visibility + i);
                if (PASS == PASS_EARLY) act[k] = act[k] && (visPrev[k] != 0u);    // InitialDrawCull.comp.glsl:21-24
            }
            obj[k] = make_uint2(0u, 0u);
            if (act[k]) obj[k] = ldg_nc_u2(p.objs + i);
        }
        float4 ps[ITEMS], qt[ITEMS];
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            ps[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            qt[k] = make_float4(0.f, 0.f, 0.f, 1.f);
            if (act[k]) {
                const uint32_t t = obj[k].x - p.transformIdBase;
                ps[k] = ldg_nc_f4(p.xfPosScale + t);
                qt[k] = ldg_nc_f4(p.xfQuat + t);
            }
        }

        // ---- phase 2: cull, LOD select, rank inside the warp ----------------------------------------------------
        uint32_t emitMask = 0u;            // bit k: object k of this lane emits a record
        uint32_t rank[ITEMS];
        uint32_t lodAbs[ITEMS];
        uint32_t running = 0u;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            bool visible = false;
            Sphere s{ 0.f, 0.f, 0.f, 0.f };
            uint32_t lodOffset = 0u, lodCount = 0u;
            if (act[k]) {
                const PrimitiveSurface& sf = T.surf[obj[k].y];
                lodOffset = sf.lodOffset; lodCount = sf.lodCount;
                s = view_space_sphere(sf.center[0], sf.center[1], sf.center[2], sf.radius,
                                      ps[k].x, ps[k].y, ps[k].z, ps[k].w, qt[k].x, qt[k].y, qt[k].z, qt[k].w, V);
                visible = frustum_test(s, V);
                if ((PASS == PASS_LATE || PASS == PASS_TEMPORAL) && visible) {
                    float4 aabb;
                    if (project_sphere(s, V.zNear, V.proj0, V.proj5, aabb))
                        visible = (HIZ == HIZ_VK) ? hiz_test_vk(aabb, p.pyr, s, V) : hiz_test_dx(aabb, p.pyr, s, V);
                }
            }
            bool emit = visible;
            if (PASS == PASS_LATE) {
                emit = visible && (visPrev[k] == 0u);                            // LateDrawCull.comp.glsl:49
                const uint32_t i = base + uint32_t(k) * 32u;
                if (i < p.n) st_cs_u32(p.visibility + i, visible ? 1u : 0u);      // LateDrawCull.comp.glsl:70
            }
            lodAbs[k] = 0u;
            if (emit) {
                const uint32_t rel = lod_select(s, ps[k].w, V.lodTarget, lodOffset, lodCount,
                                                [&](uint32_t li) { return T.lod[li].error; });
                lodAbs[k] = (p.flags & kFlagOnpcLodQuirk) ? rel : rel + lodOffset;
            }
            const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, emit);
            rank[k] = running + uint32_t(__popc(ballot & laneLt));
            running += uint32_t(__popc(ballot));
            emitMask |= (emit ? 1u : 0u) << k;
        }
        if (lane == 0) s_warpCnt[warp] = running;
        __syncthreads();   // (B) warp counts visible

        uint32_t warpOff = 0u, tileTotal = 0u;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const uint32_t c = s_warpCnt[w];
            if (uint32_t(w) < warp) warpOff += c;
            tileTotal += c;
        }
        if (warp == 0) {
            const uint32_t prefix = lookback_exclusive_prefix(p.status, tile, tileTotal, epoch, lane);
            if (lane == 0) {
                s_prefix = prefix;
                if (tile == p.numTiles - 1u) {                                     // the draw count the indirect draw reads
                    const uint64_t total = uint64_t(prefix) + tileTotal;
                    p.counts[0] = uint32_t(total < p.capacity ? total : p.capacity);
                    p.counts[1] = uint32_t(total);
                }
            }
        }
        // ---- phase 3: stage the records of this warp in shared memory ---------------------------------------------
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            if (emitMask & (1u << k)) {
                const LodData& lod = T.lod[lodAbs[k]];
                uint32_t* r = staging + size_t(warpOff + rank[k]) * p.recWords;
                const uint32_t i = base + uint32_t(k) * 32u;
                // {objectId, indexCount, instanceCount = 1, firstIndex, vertexOffset = 0, firstInstance = 0 [, pad, pad]}
                *reinterpret_cast<uint2*>(r + 0) = make_uint2(p.objectIdBase + i, lod.indexCount);
                *reinterpret_cast<uint2*>(r + 2) = make_uint2(1u, lod.firstIndex);
                *reinterpret_cast<uint2*>(r + 4) = make_uint2(0u, 0u);
                if (p.recWords == 8u) *reinterpret_cast<uint2*>(r + 6) = make_uint2(0u, 0u);
            }
        }
        __syncthreads();   // (C) staging complete, prefix known

        // ---- phase 4: one contiguous span per tile ---------------------------------------------------------------
        const uint64_t prefix = s_prefix;
        uint64_t room = prefix < p.capacity ? p.capacity - prefix : 0ull;
        const uint32_t nrec = uint32_t(room < tileTotal ? room : tileTotal);
        const uint32_t nwords64 = nrec * (p.recWords >> 1);
        uint2* dst = reinterpret_cast<uint2*>(p.draws + prefix * p.recWords);
        const uint2* src = reinterpret_cast<const uint2*>(staging);
        for (uint32_t j = tid; j < nwords64; j += kCullThreads) st_cs_u2(dst + j, src[j]);
    }

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

template <int PASS, int HIZ>
static cudaError_t launch_variant(const DrawCullParams& p, int numSMs, cudaStream_t stream)
{
    const size_t tableBytes = (size_t(p.surfaceCount) + p.lodCount) * 32u;
    const bool smemTables = tableBytes <= 16384u;
    const size_t stagingBytes = size_t(kCullTile) * p.recWords * 4u;
    const size_t smem = stagingBytes + (smemTables ? tableBytes : 0u);
    auto kernel = smemTables ? draw_cull_kernel<PASS, HIZ, true, kCullItems> : draw_cull_kernel<PASS, HIZ, false, kCullItems>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    int perSM = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, kCullThreads, smem);
    if (e != cudaSuccess) return e;
    if (perSM < 1) perSM = 1;
    uint32_t grid = uint32_t(numSMs) * uint32_t(perSM);
    if (grid > p.numTiles) grid = p.numTiles;
    if (grid < 1) grid = 1;
    kernel<<<grid, kCullThreads, smem, stream>>>(p);
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
