// Indirect-instancing cull (D3D12 backend) and the cluster path (Vulkan backend), replacing
//   HlslShaders/CS/{drawInstCountReset,drawInstCull,drawInstCmd}.cs.hlsl   (host: BlitzenDX12/dx12Draw.cpp:340-413)
//   VulkanShaders/PreClusterDrawCull.comp.glsl                             (host: BlitzenVulkan/vulkanDraw.cpp:318-370)
//   VulkanShaders/{Initial,Transparent}ClusterCull.comp.glsl               (host: BlitzenVulkan/vulkanDraw.cpp:372-423)
// (paths relative to /root/reference/src/Renderer).  Same persistent / ticketed-tile / look-back skeleton as cull_draw.cu.
#include "cull_kernels.cuh"
#include "cull_math.cuh"
#include "scan_lookback.cuh"

namespace blz {

namespace {

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

// copies the surface + LOD tables to shared memory (or points at global memory when they do not fit)
template <bool SMEM_TABLES>
__device__ __forceinline__ unsigned char* setup_tables(unsigned char* smem, const PrimitiveSurface* gS, uint32_t nS, const LodData* gL, uint32_t nL,
                                                       const PrimitiveSurface*& surf, const LodData*& lod)
{
    if (SMEM_TABLES) {
        uint4* dstS = reinterpret_cast<uint4*>(smem);
        const uint4* srcS = reinterpret_cast<const uint4*>(gS);
        for (uint32_t i = threadIdx.x; i < nS * 2u; i += blockDim.x) dstS[i] = __ldg(srcS + i);
        uint4* dstL = dstS + nS * 2u;
        const uint4* srcL = reinterpret_cast<const uint4*>(gL);
        for (uint32_t i = threadIdx.x; i < nL * 2u; i += blockDim.x) dstL[i] = __ldg(srcL + i);
        surf = reinterpret_cast<const PrimitiveSurface*>(dstS);
        lod = reinterpret_cast<const LodData*>(dstL);
        return reinterpret_cast<unsigned char*>(dstL + nL * 2u);
    }
    surf = gS; lod = gL;
    return smem;
}

__device__ __forceinline__ void leave_kernel(ScanCtl* ctl, uint32_t epoch)
{
    if (threadIdx.x == 0) {
        __threadfence();
        const uint32_t prev = atomicAdd(&ctl->done, 1u);
        if (prev == gridDim.x - 1u) {
            uint32_t e = (epoch + 1u) & 0x3FFFFFFFu;
            ctl->epoch = e ? e : 1u;
            ctl->ticket = 0u;
            ctl->done = 0u;
        }
    }
}

// frustum + LOD for ITEMS objects per lane (the drawCull / PreClusterDrawCull front end); returns emit mask and absolute LOD ids
template <int ITEMS>
__device__ __forceinline__ uint32_t eval_frustum_lod(const RenderObject* objs, const float4* xfPS, const float4* xfQ, uint32_t transformIdBase,
                                                     const PrimitiveSurface* surf, const LodData* lod, const ViewConsts& V,
                                                     uint32_t base, uint32_t n, uint32_t (&lodAbs)[ITEMS])
{
    bool act[ITEMS]; uint2 obj[ITEMS]; float4 ps[ITEMS], qt[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const uint32_t i = base + uint32_t(k) * 32u;
        act[k] = i < n;
        obj[k] = make_uint2(0u, 0u);
        if (act[k]) obj[k] = ldg_nc_u2(objs + i);
    }
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        ps[k] = make_float4(0.f, 0.f, 0.f, 0.f); qt[k] = make_float4(0.f, 0.f, 0.f, 1.f);
        if (act[k]) { const uint32_t t = obj[k].x - transformIdBase; ps[k] = ldg_nc_f4(xfPS + t); qt[k] = ldg_nc_f4(xfQ + t); }
    }
    uint32_t emitMask = 0u;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        lodAbs[k] = 0u;
        if (act[k]) {
            const PrimitiveSurface& sf = surf[obj[k].y];
            const Sphere s = view_space_sphere(sf.center[0], sf.center[1], sf.center[2], sf.radius,
                                               ps[k].x, ps[k].y, ps[k].z, ps[k].w, qt[k].x, qt[k].y, qt[k].z, qt[k].w, V);
            if (frustum_test(s, V)) {
                const uint32_t rel = lod_select(s, ps[k].w, V.lodTarget, sf.lodOffset, sf.lodCount, [&](uint32_t li) { return lod[li].error; });
                lodAbs[k] = rel + sf.lodOffset;
                emitMask |= 1u << k;
            }
        }
    }
    return emitMask;
}

} // namespace

// ------------------------------------------------------------------------------------------------------------------------
// Instancing: survivors are bucketed by their selected LOD; inside a bucket ids ascend.  One look-back chain per LOD.
// ------------------------------------------------------------------------------------------------------------------------
constexpr int kInstMaxLods = 256;

template <bool SMEM_TABLES, int ITEMS>
__global__ void __launch_bounds__(kCullThreads, 3) instance_cull_kernel(const __grid_constant__ InstanceCullParams p)
{
    constexpr int TILE = kCullThreads * ITEMS;
    constexpr int WARPS = kCullThreads / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_tile;
    const PrimitiveSurface* surf; const LodData* lod;
    uint32_t* s_wcnt = reinterpret_cast<uint32_t*>(setup_tables<SMEM_TABLES>(smem_raw, p.surfaces, p.surfaceCount, p.lods, p.lodCount, surf, lod));
    const uint32_t L = p.lodCount;
    uint32_t* s_prefix = s_wcnt + WARPS * L;     // exclusive prefix of this tile per LOD
    uint32_t* s_total = s_prefix + L;            // inclusive totals (last tile only)
    uint32_t* s_off = s_total + L;               // instanceOffset per LOD
    uint32_t* s_cap = s_off + L;                 // bucket capacity per LOD

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t laneLt = (1u << lane) - 1u;
    for (uint32_t l = tid; l < L; l += kCullThreads) { s_off[l] = p.lodInstances[l].instanceOffset; s_cap[l] = p.bucketCapacity[l]; }
    const uint32_t epoch = ld_cg_u32(&p.ctl->epoch);

    while (true) {
        if (tid == 0) s_tile = atomicAdd(&p.ctl->ticket, 1u);
        for (uint32_t i = tid; i < WARPS * L; i += kCullThreads) s_wcnt[i] = 0u;
        __syncthreads();   // (A)
        const uint32_t tile = s_tile;
        if (tile >= p.numTiles) break;
        const uint32_t base = tile * uint32_t(TILE) + warp * uint32_t(32 * ITEMS) + lane;

        uint32_t lodAbs[ITEMS], rankW[ITEMS];
        const uint32_t emitMask = eval_frustum_lod<ITEMS>(p.objs, p.xfPosScale, p.xfQuat, p.transformIdBase, surf, lod, p.view, base, p.n, lodAbs);
        uint32_t* myCnt = s_wcnt + warp * L;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const bool emit = (emitMask >> k) & 1u;
            const uint32_t key = emit ? lodAbs[k] : 0xFFFFFFFFu;
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key);
            const uint32_t r = uint32_t(__popc(peers & laneLt));
            uint32_t b = 0u;
            if (emit) b = myCnt[key];
            __syncwarp();
            if (emit && r == 0u) myCnt[key] = b + uint32_t(__popc(peers));
            __syncwarp();
            rankW[k] = b + r;
        }
        __syncthreads();   // (B) per-warp per-LOD counts complete

        for (uint32_t l = tid; l < L; l += kCullThreads) {
            uint32_t run = 0u;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) { const uint32_t c = s_wcnt[w * L + l]; s_wcnt[w * L + l] = run; run += c; }
            const uint32_t prefix = lookback_exclusive_prefix_serial(p.status, L, l, tile, run, epoch);
            s_prefix[l] = prefix;
            if (tile == p.numTiles - 1u) { s_total[l] = prefix + run; p.lodInstances[l].instanceCount = prefix + run; }
        }
        __syncthreads();   // (C)

#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            if ((emitMask >> k) & 1u) {
                const uint32_t l = lodAbs[k];
                const uint32_t pos = s_prefix[l] + myCnt[l] + rankW[k];
                if (pos < s_cap[l]) p.instanceIndices[size_t(s_off[l]) + pos] = p.objectIdBase + base + uint32_t(k) * 32u;   // drawInstCull.cs.hlsl:41
            }
        }
        if (tile == p.numTiles - 1u && tid == 0) {
            // drawInstCmd.cs.hlsl:9-39, one command per non-empty LOD, ascending LOD id
            uint32_t nCmd = 0u, nTot = 0u;
            for (uint32_t l = 0; l < L; ++l) {
                const uint32_t c = s_total[l];
                if (c == 0u) continue;
                if (uint64_t(nCmd) < p.cmdCapacity) {
                    uint32_t* r = p.cmds + size_t(nCmd) * 8u;
                    r[0] = s_off[l]; r[1] = lod[l].indexCount; r[2] = c < s_cap[l] ? c : s_cap[l]; r[3] = lod[l].firstIndex;
                    r[4] = 0u; r[5] = 0u; r[6] = 0u; r[7] = 0u;
                    ++nCmd;
                }
                ++nTot;
            }
            p.counts[0] = nCmd; p.counts[1] = nTot;
        }
        __syncthreads();   // (D) s_wcnt / s_prefix reads done before the next tile zeroes them
    }
    leave_kernel(p.ctl, epoch);
}

cudaError_t launch_instance_cull(const InstanceCullParams& p, int numSMs, cudaStream_t stream)
{
    if (p.lodCount == 0 || p.lodCount > uint32_t(kInstMaxLods)) return cudaErrorInvalidValue;
    const size_t tableBytes = (size_t(p.surfaceCount) + p.lodCount) * 32u;
    const bool smemTables = tableBytes <= 16384u;
    const size_t smem = (smemTables ? tableBytes : 0u) + size_t(kCullThreads / 32 + 4) * p.lodCount * 4u;
    auto kernel = smemTables ? instance_cull_kernel<true, kCullItems> : instance_cull_kernel<false, kCullItems>;
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

// ------------------------------------------------------------------------------------------------------------------------
// Cluster expand: every surviving object appends lod.clusterCount dispatch records.  The records of a tile are produced
// cooperatively (record -> owning survivor by binary search over the tile's start offsets), not by a per-thread serial loop.
// ------------------------------------------------------------------------------------------------------------------------
template <bool SMEM_TABLES, int ITEMS>
__global__ void __launch_bounds__(kCullThreads, 3) cluster_expand_kernel(const __grid_constant__ ClusterExpandParams p)
{
    constexpr int TILE = kCullThreads * ITEMS;
    constexpr int WARPS = kCullThreads / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_tile, s_prefix;
    __shared__ uint32_t s_warpSurv[WARPS], s_warpRec[WARPS];
    const PrimitiveSurface* surf; const LodData* lod;
    uint32_t* s_start = reinterpret_cast<uint32_t*>(setup_tables<SMEM_TABLES>(smem_raw, p.surfaces, p.surfaceCount, p.lods, p.lodCount, surf, lod));
    uint32_t* s_obj = s_start + (TILE + 1);      // survivor descriptors, indexed by survivor rank inside the tile
    uint32_t* s_lod = s_obj + TILE;
    uint32_t* s_coff = s_lod + TILE;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t laneLt = (1u << lane) - 1u;
    const uint32_t epoch = ld_cg_u32(&p.ctl->epoch);

    while (true) {
        if (tid == 0) s_tile = atomicAdd(&p.ctl->ticket, 1u);
        __syncthreads();   // (A)
        const uint32_t tile = s_tile;
        if (tile >= p.numTiles) break;
        const uint32_t base = tile * uint32_t(TILE) + warp * uint32_t(32 * ITEMS) + lane;

        uint32_t lodAbs[ITEMS], survRank[ITEMS], recStart[ITEMS], cnt[ITEMS];
        const uint32_t emitMask = eval_frustum_lod<ITEMS>(p.objs, p.xfPosScale, p.xfQuat, p.transformIdBase, surf, lod, p.view, base, p.n, lodAbs);
        uint32_t survRun = 0u, recRun = 0u;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const bool emit = (emitMask >> k) & 1u;
            cnt[k] = emit ? lod[lodAbs[k]].clusterCount : 0u;
            const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, emit);
            survRank[k] = survRun + uint32_t(__popc(ballot & laneLt));
            survRun += uint32_t(__popc(ballot));
            uint32_t x = cnt[k];                                   // inclusive warp scan of the record counts
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d); if (lane >= uint32_t(d)) x += y; }
            recStart[k] = recRun + x - cnt[k];
            recRun += __shfl_sync(0xFFFFFFFFu, x, 31);
        }
        if (lane == 0) { s_warpSurv[warp] = survRun; s_warpRec[warp] = recRun; }
        __syncthreads();   // (B)

        uint32_t survOff = 0u, recOff = 0u, tileSurv = 0u, tileRec = 0u;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const uint32_t a = s_warpSurv[w], b = s_warpRec[w];
            if (uint32_t(w) < warp) { survOff += a; recOff += b; }
            tileSurv += a; tileRec += b;
        }
        if (warp == 0) {
            const uint32_t prefix = lookback_exclusive_prefix(p.status, tile, tileRec, epoch, lane);
            if (lane == 0) {
                s_prefix = prefix;
                if (tile == p.numTiles - 1u) {
                    const uint64_t total = uint64_t(prefix) + tileRec;
                    p.counts[0] = uint32_t(total < p.capacity ? total : p.capacity);
                    p.counts[1] = uint32_t(total);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            if ((emitMask >> k) & 1u) {
                const uint32_t s = survOff + survRank[k];
                s_start[s] = recOff + recStart[k];
                s_obj[s] = p.objectIdBase + base + uint32_t(k) * 32u;
                s_lod[s] = lodAbs[k];
                s_coff[s] = lod[lodAbs[k]].clusterOffset;
            }
        }
        if (tid == 0) s_start[tileSurv] = tileRec;
        __syncthreads();   // (C)

        const uint64_t prefix = s_prefix;
        const uint64_t room = prefix < p.capacity ? p.capacity - prefix : 0ull;
        const uint32_t nrec = uint32_t(room < tileRec ? room : tileRec);
        uint32_t* out = p.dispatch + prefix * 3ull;
        for (uint32_t j = tid; j < nrec; j += kCullThreads) {
            // largest s with s_start[s] <= j  (s_start is non-decreasing; empty survivors (count 0) are skipped by taking the last)
            uint32_t lo = 0u, hi = tileSurv;                       // invariant: s_start[lo] <= j < s_start[hi]
            while (hi - lo > 1u) { const uint32_t mid = (lo + hi) >> 1; if (s_start[mid] <= j) lo = mid; else hi = mid; }
            // PreClusterDrawCull.comp.glsl:41-43  {objectId, lodIndex, clusterId = lod.clusterOffset + i}
            out[size_t(j) * 3u + 0u] = s_obj[lo];
            out[size_t(j) * 3u + 1u] = s_lod[lo];
            out[size_t(j) * 3u + 2u] = s_coff[lo] + (j - s_start[lo]);
        }
    }
    leave_kernel(p.ctl, epoch);
}

cudaError_t launch_cluster_expand(const ClusterExpandParams& p, int numSMs, cudaStream_t stream)
{
    const size_t tableBytes = (size_t(p.surfaceCount) + p.lodCount) * 32u;
    const bool smemTables = tableBytes <= 16384u;
    const size_t smem = (smemTables ? tableBytes : 0u) + (size_t(kCullTile) * 4u + 1u) * 4u;
    auto kernel = smemTables ? cluster_expand_kernel<true, kCullItems> : cluster_expand_kernel<false, kCullItems>;
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

// ------------------------------------------------------------------------------------------------------------------------
// Cluster cull.  mode 0 (passthrough) is the reference: every dispatch record becomes one draw record, in record order.
// mode 1/2 add the per-cluster bounding-sphere frustum (+ Hi-Z) test the north star asks for; survivors are compacted with
// the same look-back.  The record count is read from device memory: no CPU round trip between expand and cull.
// ------------------------------------------------------------------------------------------------------------------------
template <int HIZ, int ITEMS>
__global__ void __launch_bounds__(kCullThreads, 3) cluster_cull_kernel(const __grid_constant__ ClusterCullParams p)
{
    constexpr int TILE = kCullThreads * ITEMS;
    constexpr int WARPS = kCullThreads / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_tile, s_prefix;
    __shared__ uint32_t s_warpCnt[WARPS];
    uint32_t* s_in = reinterpret_cast<uint32_t*>(smem_raw);          // TILE dispatch records (3 words each)
    uint32_t* staging = s_in + TILE * 3;                             // TILE output records

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t laneLt = (1u << lane) - 1u;
    const uint32_t epoch = ld_cg_u32(&p.ctl->epoch);
    uint32_t M = ld_cg_u32(p.dispatchCount);
    if (M > p.maxRecords) M = p.maxRecords;
    const uint32_t numTiles = M == 0u ? 1u : (M + uint32_t(TILE) - 1u) / uint32_t(TILE);
    const ViewConsts& V = p.view;

    while (true) {
        if (tid == 0) s_tile = atomicAdd(&p.ctl->ticket, 1u);
        __syncthreads();   // (A)
        const uint32_t tile = s_tile;
        if (tile >= numTiles) break;
        const uint32_t tileBase = tile * uint32_t(TILE);
        const uint32_t tileN = M - tileBase < uint32_t(TILE) ? M - tileBase : uint32_t(TILE);
        // coalesced stage-in of the tile's dispatch records
        for (uint32_t j = tid; j < tileN * 3u; j += kCullThreads) s_in[j] = __ldg(p.dispatch + size_t(tileBase) * 3u + j);
        __syncthreads();   // (A2)

        uint32_t emitMask = 0u, rank[ITEMS], running = 0u;
        uint32_t objId[ITEMS], idxCount[ITEMS], firstIdx[ITEMS];
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const uint32_t local = warp * uint32_t(32 * ITEMS) + uint32_t(k) * 32u + lane;
            bool emit = false;
            objId[k] = 0u; idxCount[k] = 0u; firstIdx[k] = 0u;
            if (local < tileN) {
                objId[k] = s_in[local * 3u + 0u];
                const uint32_t clusterId = s_in[local * 3u + 2u];
                const float4* cp = reinterpret_cast<const float4*>(p.clusters + clusterId);
                const uint4 tail = __ldg(reinterpret_cast<const uint4*>(cp) + 1);      // {cone s8x4, dataOffset, vertexCount|triangleCount<<8, pad}
                idxCount[k] = ((tail.z >> 8) & 0xFFu) * 3u;                              // triangleCount * 3   (InitialClusterCull.comp.glsl:50)
                firstIdx[k] = tail.y;                                                    // dataOffset            (:52)
                emit = true;
                if (p.mode != 0u) {
                    const float4 bs = __ldg(cp);
                    const uint2 obj = __ldg(reinterpret_cast<const uint2*>(p.objs + (objId[k] - p.objectIdBase)));
                    const uint32_t t = obj.x - p.transformIdBase;
                    const float4 ps = __ldg(p.xfPosScale + t), qt = __ldg(p.xfQuat + t);
                    const Sphere s = view_space_sphere(bs.x, bs.y, bs.z, bs.w, ps.x, ps.y, ps.z, ps.w, qt.x, qt.y, qt.z, qt.w, V);
                    emit = frustum_test(s, V);
                    if (HIZ != HIZ_NONE && emit) {
                        float4 aabb;
                        if (project_sphere(s, V.zNear, V.proj0, V.proj5, aabb))
                            emit = (HIZ == HIZ_VK) ? hiz_test_vk(aabb, p.pyr, s, V) : hiz_test_dx(aabb, p.pyr, s, V);
                    }
                }
            }
            const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, emit);
            rank[k] = running + uint32_t(__popc(ballot & laneLt));
            running += uint32_t(__popc(ballot));
            emitMask |= (emit ? 1u : 0u) << k;
        }
        if (lane == 0) s_warpCnt[warp] = running;
        __syncthreads();   // (B)
        uint32_t warpOff = 0u, tileTotal = 0u;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) { const uint32_t c = s_warpCnt[w]; if (uint32_t(w) < warp) warpOff += c; tileTotal += c; }
        if (warp == 0) {
            // passthrough keeps every record, so the prefix is known without a chain
            const uint32_t prefix = p.mode == 0u ? tileBase : lookback_exclusive_prefix(p.status, tile, tileTotal, epoch, lane);
            if (lane == 0) {
                s_prefix = prefix;
                if (tile == numTiles - 1u) {
                    const uint64_t total = uint64_t(prefix) + tileTotal;
                    p.counts[0] = uint32_t(total < p.capacity ? total : p.capacity);
                    p.counts[1] = uint32_t(total);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            if (emitMask & (1u << k)) {
                uint32_t* r = staging + size_t(warpOff + rank[k]) * p.recWords;
                *reinterpret_cast<uint2*>(r + 0) = make_uint2(objId[k], idxCount[k]);
                *reinterpret_cast<uint2*>(r + 2) = make_uint2(1u, firstIdx[k]);
                *reinterpret_cast<uint2*>(r + 4) = make_uint2(0u, 0u);
                if (p.recWords == 8u) *reinterpret_cast<uint2*>(r + 6) = make_uint2(0u, 0u);
            }
        }
        __syncthreads();   // (C)
        const uint64_t prefix = s_prefix;
        const uint64_t room = prefix < p.capacity ? p.capacity - prefix : 0ull;
        const uint32_t nrec = uint32_t(room < tileTotal ? room : tileTotal);
        const uint32_t nwords64 = nrec * (p.recWords >> 1);
        uint2* dst = reinterpret_cast<uint2*>(p.draws + prefix * p.recWords);
        const uint2* src = reinterpret_cast<const uint2*>(staging);
        for (uint32_t j = tid; j < nwords64; j += kCullThreads) dst[j] = src[j];
    }
    leave_kernel(p.ctl, epoch);
}

cudaError_t launch_cluster_cull(const ClusterCullParams& p, int hiz, int numSMs, cudaStream_t stream)
{
    const size_t smem = size_t(kCullTile) * 3u * 4u + size_t(kCullTile) * p.recWords * 4u;
    void (*kernel)(const ClusterCullParams) = nullptr;
    if (p.mode != 2u) kernel = cluster_cull_kernel<HIZ_NONE, kCullItems>;
    else kernel = hiz == HIZ_VK ? cluster_cull_kernel<HIZ_VK, kCullItems> : cluster_cull_kernel<HIZ_DX, kCullItems>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    int perSM = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, kCullThreads, smem);
    if (e != cudaSuccess) return e;
    if (perSM < 1) perSM = 1;
    uint32_t grid = uint32_t(numSMs) * uint32_t(perSM);
    const uint32_t maxTiles = p.maxRecords == 0u ? 1u : (p.maxRecords + uint32_t(kCullTile) - 1u) / uint32_t(kCullTile);
    if (grid > maxTiles) grid = maxTiles;
    if (grid < 1) grid = 1;
    kernel<<<grid, kCullThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

} // namespace blz
