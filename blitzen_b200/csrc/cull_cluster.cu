// The cluster path's cull step (Vulkan backend), replacing
//   VulkanShaders/{Initial,Transparent}ClusterCull.comp.glsl               (host: BlitzenVulkan/vulkanDraw.cpp:372-423)
// (paths relative to /root/reference/src/Renderer).  The expand step (PreClusterDrawCull) lives in cull_list.cu.
#include <cstdlib>
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

// L2 residency hints: the dispatch records and the draw records stream through once (evict-first), the 219-KB cluster table is
// re-read by every record (evict-last) -- without them the 1.2 GB of record stores pushed table sectors out of L2 (ncu: 1.10 GB
// of DRAM reads for 0.70 GB of input)
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 ldg_f4_hint(const void* p, uint64_t pol)
{
    float4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint32_t ldg_u32_hint(const void* p, uint64_t pol)
{
    uint32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_cs_u2(void* p, uint2 v) { asm volatile("st.global.cs.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory"); }

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


} // namespace

// ------------------------------------------------------------------------------------------------------------------------
// Cluster cull.  mode 0 (passthrough) is the reference: every dispatch record becomes one draw record, in record order.
// mode 1/2 add the per-cluster bounding-sphere frustum (+ Hi-Z) test the north star asks for; survivors are compacted with
// the same look-back.  The record count is read from device memory: no CPU round trip between expand and cull.
// ------------------------------------------------------------------------------------------------------------------------
// Structure (profiles/r01k_inst_cluster.jsonl showed the first version -- one-shot tile, per-item dependent load chain record ->
// RenderObject -> transform, 8 round trips per tile, look-back on the critical path -- at 26 % / 12 % of the HBM peak in the sphere modes):
//   * persistent CTAs, software-pipelined over tiles of TILE records.  In iteration `it` a CTA
//       - stages the records of tile it+2 into shared memory (cp.async, 16-B chunks, ring of 3; tickets claimed three tiles ahead),
//       - reads the records of tile it+1 from shared memory and issues their RenderObject gather,
//       - does the arithmetic of tile it (its transforms / cluster spheres were requested at the end of iteration it-1),
//       - then issues the transform + cluster-sphere gathers of tile it+1 into the registers the arithmetic just released,
//       - and writes out the records of tile it-2.
//     Every global round trip is one iteration old when it is consumed.
//   * compaction as in cull_stream.cu, chain-free: a tile publishes its AGGREGATE as soon as it is ranked; its records are written
//     kClLag iterations later from the sum of the aggregates of every tile between the CTA's consecutive tiles (all threads, one L2
//     round trip, issued at the top of the iteration and consumed behind the arithmetic).  Nothing waits for another tile's prefix.
//     (A look-back that waits for inclusive prefixes serialises here: with tickets claimed ahead of time, tile order is no longer
//     start order -- measured 17-27 ms.)
//   * survivors are staged per warp as 8-B descriptors {objectId, clusterId}; each warp writes the records of its own span:
//     {dataOffset, triangleCount} come from the cluster table (second half of the sector the sphere came from), the 32 records of a
//     step are transposed through a per-warp buffer and leave as contiguous runs of 8-B stores.
constexpr int kClLag = 2;
constexpr int kClStages = kClLag + 1;
constexpr uint32_t kClNone = 0xFFFFFFFFu;

template <int HIZ, int ITEMS, int MINB>
__global__ void __launch_bounds__(kCullThreads, MINB) cluster_cull_kernel(const __grid_constant__ ClusterCullParams p)
{
    constexpr int TILE = kCullThreads * ITEMS;
    constexpr int WARPS = kCullThreads / 32;
    constexpr int THREADS = kCullThreads;
    constexpr int WSPAN = 32 * ITEMS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_tiles[4];
    __shared__ uint32_t s_warpCnt[4][WARPS];
    __shared__ uint32_t s_sum[4];
    uint32_t* s_in = reinterpret_cast<uint32_t*>(smem_raw);                        // 3 x TILE dispatch records (3 words each)
    uint2* stageB = reinterpret_cast<uint2*>(s_in + 3 * TILE * 3);                 // kClStages x TILE survivor descriptors, per-warp spans
    uint2* wbuf = stageB + kClStages * TILE;                                       // per warp: 32 records being transposed on the way out
    static_assert((TILE * 12) % 16 == 0, "a tile of records is a whole number of 16-B chunks");

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t laneLt = (1u << lane) - 1u;
    const uint32_t epoch = ld_cg_u32(&p.ctl->epoch) & 0x3FFFFFFFu;
    uint32_t M = ld_cg_u32(p.dispatchCount);
    if (M > p.maxRecords) M = p.maxRecords;
    const uint32_t numTiles = M == 0u ? 1u : (M + uint32_t(TILE) - 1u) / uint32_t(TILE);
    const ViewConsts& V = p.view;
    constexpr bool sphereMode = true;        // the passthrough mode has its own kernel below; as a run-time flag in this one it cost registers (cf. cull_stream.cu, r02)
    const uint32_t cap32 = p.capacity > 0xFFFFFFFFull ? 0xFFFFFFFFu : uint32_t(p.capacity);
    const uint32_t localBase = warp * uint32_t(WSPAN) + lane;
    const uint64_t polStream = l2_policy_evict_first(), polTable = l2_policy_evict_last();

    auto tile_count = [&](uint32_t tile) { const uint32_t b = tile * uint32_t(TILE); return M - b < uint32_t(TILE) ? M - b : uint32_t(TILE); };
    auto stage_in = [&](uint32_t tile, uint32_t buf) {
        const uint32_t words = tile_count(tile) * 3u, full = words >> 2;          // whole 16-B chunks inside the valid records
        const uint32_t* src = p.dispatch + size_t(tile) * (TILE * 3);
        uint32_t* dst = s_in + buf * (TILE * 3);
        for (uint32_t c = tid; c < full; c += THREADS)
            asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(uint32_t(__cvta_generic_to_shared(dst + c * 4u))), "l"(src + c * 4u), "l"(polStream) : "memory");
        for (uint32_t w = full * 4u + tid; w < words; w += THREADS) dst[w] = __ldg(src + w);   // ragged tail of the last tile
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // registers carried from the fetch of a tile to its arithmetic one iteration later
    float4 bs[ITEMS], ps[ITEMS], qt[ITEMS];
    uint32_t inMask = 0u;
    // records of `tile` (in ring slot buf) -> RenderObject gather (returns the transform indices) + cluster ids
    auto fetch_records = [&](uint32_t tile, uint32_t buf, uint32_t (&cid)[ITEMS], uint32_t (&xf)[ITEMS]) -> uint32_t {
        const uint32_t tileN = tile_count(tile);
        const uint32_t* rec = s_in + buf * (TILE * 3);
        uint32_t m = 0u;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const uint32_t local = localBase + uint32_t(k) * 32u;
            const bool in = local < tileN;
            m |= (in ? 1u : 0u) << k;
            cid[k] = in ? rec[local * 3u + 2u] : 0u;
            xf[k] = p.transformIdBase;
            if (sphereMode && in) xf[k] = __ldg(reinterpret_cast<const uint32_t*>(p.objs + (rec[local * 3u + 0u] - p.objectIdBase)));   // RenderObject.transformId
        }
        return m;
    };
    auto issue_gathers = [&](uint32_t m, const uint32_t (&cid)[ITEMS], const uint32_t (&xf)[ITEMS]) {
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            if (sphereMode && ((m >> k) & 1u)) {
                const uint32_t t = xf[k] - p.transformIdBase;
                ld_transform(p.xf + t, ps[k], qt[k]);
                bs[k] = ldg_f4_hint(p.clusters + cid[k], polTable);
            }
        }
    };
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) { bs[k] = make_float4(0.f, 0.f, 0.f, 0.f); ps[k] = make_float4(0.f, 0.f, 0.f, 1.f); qt[k] = make_float4(0.f, 0.f, 0.f, 1.f); }

    uint32_t lastClaim = 0u;                 // thread 0: the most recent ticket
    if (tid == 0) {
        s_tiles[0] = atomicAdd(&p.ctl->ticket, 1u);
        s_tiles[1] = atomicAdd(&p.ctl->ticket, 1u);
        lastClaim = atomicAdd(&p.ctl->ticket, 1u);
        s_tiles[2] = lastClaim;
        s_sum[0] = s_sum[1] = s_sum[2] = s_sum[3] = 0u;
    }
    __syncthreads();
    if (s_tiles[0] < numTiles) {
        stage_in(s_tiles[0], 0u);
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
        if (s_tiles[1] < numTiles) stage_in(s_tiles[1], 1u);                       // waited for at barrier (A) of iteration 0
        uint32_t cid[ITEMS], xf[ITEMS];
        inMask = fetch_records(s_tiles[0], 0u, cid, xf);
        issue_gathers(inMask, cid, xf);
    }

    uint32_t prev1 = kClNone, prev2 = kClNone;       // tiles of iterations it-1, it-2 (prev2's records go out in iteration it)
    uint32_t cum = 0u;                               // records emitted by tiles [0, nextRead)
    uint32_t nextRead = 0u;                          // first tile whose aggregate this CTA has not summed yet
    for (uint32_t it = 0u;; ++it) {
        uint32_t tile = s_tiles[it & 3u];
        if (tile >= numTiles) tile = kClNone;
        const bool valid = tile != kClNone;
        if (!valid && prev1 == kClNone && prev2 == kClNone) break;

        // aggregates of the tiles between this CTA's previous tile and prev2: issued now, consumed behind the arithmetic
        constexpr int NS = 2;
        uint64_t sw[NS];
        if (sphereMode && prev2 != kClNone) {
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const uint32_t t = nextRead + tid + uint32_t(s) * THREADS;
                sw[s] = t < prev2 ? ld_status(p.status + t) : 0ull;
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();   // (A) records of tile it+1 staged; ticket of tile it+2 visible; ring slot (it+2)%3 free
        const uint32_t tile1 = s_tiles[(it + 1u) & 3u], tile2 = s_tiles[(it + 2u) & 3u];
        if (tile2 < numTiles) stage_in(tile2, (it + 2u) % 3u);
        uint32_t ticket = kClNone;
        if (tid == 0 && lastClaim < numTiles) { ticket = atomicAdd(&p.ctl->ticket, 1u); lastClaim = ticket; }   // tile of iteration it + 3
        uint32_t cidN[ITEMS], xfN[ITEMS], inMaskN = 0u;
        if (tile1 < numTiles) inMaskN = fetch_records(tile1, (it + 1u) % 3u, cidN, xfN);

        uint32_t running = 0u;
        if (valid) {
            // ---- sphere + frustum (+ Hi-Z) of tile it, rank by ballot, descriptors into the warp's span ------------------------------
            uint32_t emitMask = inMask;
            if (sphereMode) {
                emitMask = 0u;
#pragma unroll
                for (int k = 0; k < ITEMS; ++k) {
                    const Sphere s = view_space_sphere(bs[k].x, bs[k].y, bs[k].z, bs[k].w, ps[k].x, ps[k].y, ps[k].z, ps[k].w, qt[k].x, qt[k].y, qt[k].z, qt[k].w, V);
                    bool emit = ((inMask >> k) & 1u) && frustum_test(s, V);
                    if (HIZ != HIZ_NONE && emit) {
                        float4 aabb;
                        if (project_sphere(s, V.zNear, V.proj0, V.proj5, aabb))
                            emit = (HIZ == HIZ_VK) ? hiz_test_vk(aabb, p.pyr, s, V) : hiz_test_dx(aabb, p.pyr, s, V);
                    }
                    emitMask |= (emit ? 1u : 0u) << k;
                }
            }
            const uint32_t* rec = s_in + (it % 3u) * (TILE * 3);
            uint2* st = stageB + (it % uint32_t(kClStages)) * TILE + warp * uint32_t(WSPAN);
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const bool emit = (emitMask >> k) & 1u;
                const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, emit);
                if (emit) {
                    const uint32_t local = localBase + uint32_t(k) * 32u;
                    st[running + uint32_t(__popc(ballot & laneLt))] = make_uint2(rec[local * 3u + 0u], rec[local * 3u + 2u]);
                }
                running += uint32_t(__popc(ballot));
            }
            if (lane == 0) s_warpCnt[it & 3u][warp] = running;
        }
        // transform + cluster-sphere gathers of tile it+1 into the registers the arithmetic just released
        issue_gathers(inMaskN, cidN, xfN);
        inMask = inMaskN;

        if (sphereMode && prev2 != kClNone) {
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
            if (lane == 0 && part != 0u) atomicAdd(&s_sum[it & 3u], part);
        }
        __syncthreads();   // (B) warp counts of tile it, aggregate sum for prev2 visible

        if (tid == 0) {
            s_tiles[(it + 3u) & 3u] = ticket;                                      // read after barrier (A) of iteration it + 1
            if (valid && sphereMode) {
                uint32_t total = 0u;
#pragma unroll
                for (int w = 0; w < WARPS; ++w) total += s_warpCnt[it & 3u][w];
                st_status(p.status + tile, pack_status(epoch, kStateAggregate, total));
            }
        }
        if (prev2 != kClNone) {
            uint32_t warpOff = 0u, total = 0u;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) { const uint32_t c = s_warpCnt[(it - uint32_t(kClLag)) & 3u][w]; if (uint32_t(w) < warp) warpOff += c; total += c; }
            const uint32_t mine = s_warpCnt[(it - uint32_t(kClLag)) & 3u][warp];
            // passthrough keeps every record, so the prefix is known without any exchange
            const uint32_t prefix = sphereMode ? cum + s_sum[it & 3u] : prev2 * uint32_t(TILE);
            if (prev2 == numTiles - 1u && tid == 0) {
                const uint32_t all = prefix + total;
                p.counts[0] = all < cap32 ? all : cap32;
                p.counts[1] = all;
            }
            const uint32_t first = prefix + warpOff;                               // index of this warp's first record in the output
            const uint32_t room = first < cap32 ? cap32 - first : 0u;
            const uint32_t nrec = room < mine ? room : mine;
            if (nrec != 0u) {
                const uint2* st = stageB + ((it - uint32_t(kClLag)) % uint32_t(kClStages)) * TILE + warp * uint32_t(WSPAN);
                uint2* dst = reinterpret_cast<uint2*>(p.draws + size_t(first) * p.recWords);
                uint2* wb = wbuf + warp * (32 * 4);
                // {objectId, triangleCount * 3} {instanceCount = 1, dataOffset} {0, 0} [{pad, pad}]   (InitialClusterCull.comp.glsl:46-53)
                // all descriptor + cluster-table loads of the span first (one exposed L2 latency per tile instead of one per 32 records)
                uint32_t oid[ITEMS], first[ITEMS], cnt3[ITEMS];
#pragma unroll
                for (int k = 0; k < ITEMS; ++k) {
                    const uint32_t r = uint32_t(k) * 32u + lane;
                    oid[k] = 0u; first[k] = 0u; cnt3[k] = 0u;
                    if (r < nrec) {
                        const uint2 d = st[r];
                        oid[k] = d.x;
                        const uint32_t* c = reinterpret_cast<const uint32_t*>(p.clusters + d.y);
                        first[k] = ldg_u32_hint(c + 5, polTable);                    // dataOffset
                        cnt3[k] = ldg_u32_hint(c + 6, polTable);                     // vertexCount | triangleCount << 8
                    }
                }
                const uint32_t per = p.recWords >> 1;                              // 8-B words per record: 3 (VK24) or 4 (DX32)
#pragma unroll
                for (int k = 0; k < ITEMS; ++k) {
                    const uint32_t r0 = uint32_t(k) * 32u;
                    if (r0 < nrec) {                                               // warp-uniform
                        wb[lane * per + 0u] = make_uint2(oid[k], ((cnt3[k] >> 8) & 0xFFu) * 3u);
                        wb[lane * per + 1u] = make_uint2(1u, first[k]);
                        wb[lane * per + 2u] = make_uint2(0u, 0u);
                        if (per == 4u) wb[lane * per + 3u] = make_uint2(0u, 0u);
                        __syncwarp();
                        uint2* o = dst + size_t(r0) * per;
                        const uint32_t n8 = (nrec - r0 < 32u ? nrec - r0 : 32u) * per;
                        if (n8 == 96u) { st_cs_u2(o + lane, wb[lane]); st_cs_u2(o + lane + 32u, wb[lane + 32u]); st_cs_u2(o + lane + 64u, wb[lane + 64u]); }
                        else for (uint32_t w = lane; w < n8; w += 32u) st_cs_u2(o + w, wb[w]);
                        __syncwarp();
                    }
                }
            }
            cum = prefix + total;
            nextRead = prev2 + 1u;
        }
        if (tid == 0) s_sum[(it + 2u) & 3u] = 0u;    // last read in iteration it-2; next added to before barrier (B) of iteration it+2
        prev2 = prev1; prev1 = tile;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    leave_kernel(p.ctl, epoch);
}

// Passthrough (the reference's InitialClusterCull as it stands: every dispatch record becomes one draw record, in record order).
// No compaction and no gathers besides the cluster table, so the plain one-shot form is the fastest (0.41 ms for 50 M records =
// 68 % of the HBM peak vs 0.48 ms through the pipelined kernel below): coalesced stage-in, one table read per record, records
// assembled in shared memory and copied out as one contiguous span at tile * TILE.
template <int ITEMS>
__global__ void __launch_bounds__(kCullThreads, 3) cluster_passthrough_kernel(const __grid_constant__ ClusterCullParams p)
{
    constexpr int TILE = kCullThreads * ITEMS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_tile;
    uint32_t* s_in = reinterpret_cast<uint32_t*>(smem_raw);          // TILE dispatch records (3 words each)
    uint32_t* staging = s_in + TILE * 3;                             // TILE output records

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t epoch = ld_cg_u32(&p.ctl->epoch);
    uint32_t M = ld_cg_u32(p.dispatchCount);
    if (M > p.maxRecords) M = p.maxRecords;
    const uint32_t numTiles = M == 0u ? 1u : (M + uint32_t(TILE) - 1u) / uint32_t(TILE);
    const uint32_t cap32 = p.capacity > 0xFFFFFFFFull ? 0xFFFFFFFFu : uint32_t(p.capacity);

    while (true) {
        if (tid == 0) s_tile = atomicAdd(&p.ctl->ticket, 1u);
        __syncthreads();   // (A) also: the previous tile's copy-out has finished with `staging`
        const uint32_t tile = s_tile;
        if (tile >= numTiles) break;
        const uint32_t tileBase = tile * uint32_t(TILE);
        const uint32_t tileN = M - tileBase < uint32_t(TILE) ? M - tileBase : uint32_t(TILE);
        for (uint32_t j = tid; j < tileN * 3u; j += kCullThreads) s_in[j] = __ldg(p.dispatch + size_t(tileBase) * 3u + j);
        __syncthreads();   // (A2)
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const uint32_t local = warp * uint32_t(32 * ITEMS) + uint32_t(k) * 32u + lane;
            if (local < tileN) {
                const uint4 tail = __ldg(reinterpret_cast<const uint4*>(p.clusters + s_in[local * 3u + 2u]) + 1);   // {cone s8x4, dataOffset, vertexCount | triangleCount << 8, -}
                uint32_t* r = staging + size_t(local) * p.recWords;
                *reinterpret_cast<uint2*>(r + 0) = make_uint2(s_in[local * 3u + 0u], ((tail.z >> 8) & 0xFFu) * 3u);   // triangleCount * 3 (InitialClusterCull.comp.glsl:50)
                *reinterpret_cast<uint2*>(r + 2) = make_uint2(1u, tail.y);                                              // dataOffset          (:52)
                *reinterpret_cast<uint2*>(r + 4) = make_uint2(0u, 0u);
                if (p.recWords == 8u) *reinterpret_cast<uint2*>(r + 6) = make_uint2(0u, 0u);
            }
        }
        if (tile == numTiles - 1u && tid == 0) { p.counts[0] = M < cap32 ? M : cap32; p.counts[1] = M; }
        __syncthreads();   // (C)
        const uint32_t room = tileBase < cap32 ? cap32 - tileBase : 0u;
        const uint32_t nrec = room < tileN ? room : tileN;
        const uint32_t nwords64 = nrec * (p.recWords >> 1);
        uint2* dst = reinterpret_cast<uint2*>(p.draws + size_t(tileBase) * p.recWords);
        const uint2* src = reinterpret_cast<const uint2*>(staging);
        for (uint32_t j = tid; j < nwords64; j += kCullThreads) st_cs_u2(dst + j, src[j]);
    }
    leave_kernel(p.ctl, epoch);
}

static cudaError_t launch_cluster_passthrough(const ClusterCullParams& p, int numSMs, cudaStream_t stream)
{
    const size_t smem = size_t(kCullTile) * 3u * 4u + size_t(kCullTile) * p.recWords * 4u;
    auto kernel = cluster_passthrough_kernel<kCullItems>;
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

template <int HIZ, int ITEMS, int MINB>
static cudaError_t launch_cluster_cfg(const ClusterCullParams& p, int numSMs, cudaStream_t stream)
{
    constexpr int TILE = kCullThreads * ITEMS;
    const size_t smem = size_t(TILE) * 3u * 4u * 3u + size_t(TILE) * 8u * kClStages + size_t(kCullThreads / 32) * 32u * 4u * 8u;
    auto kernel = cluster_cull_kernel<HIZ, ITEMS, MINB>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    int perSM = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, kCullThreads, smem);
    if (e != cudaSuccess) return e;
    if (perSM < 1) perSM = 1;
    uint32_t grid = uint32_t(numSMs) * uint32_t(perSM);
    const uint32_t maxTiles = p.maxRecords == 0u ? 1u : (p.maxRecords + uint32_t(TILE) - 1u) / uint32_t(TILE);
    if (grid > maxTiles) grid = maxTiles;
    if (grid < 1) grid = 1;
    kernel<<<grid, kCullThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

template <int HIZ>
static cudaError_t launch_cluster_hiz(const ClusterCullParams& p, int cfg, int numSMs, cudaStream_t stream)
{
    switch (cfg) {
    case 1: return launch_cluster_cfg<HIZ, 4, 3>(p, numSMs, stream);
    case 2: return launch_cluster_cfg<HIZ, 2, 3>(p, numSMs, stream);
    case 3: return launch_cluster_cfg<HIZ, 2, 4>(p, numSMs, stream);
    case 4: return launch_cluster_cfg<HIZ, 3, 3>(p, numSMs, stream);
    default: return launch_cluster_cfg<HIZ, 4, 2>(p, numSMs, stream);
    }
}

cudaError_t launch_cluster_cull(const ClusterCullParams& p, int hiz, int numSMs, cudaStream_t stream)
{
    static const int envCfg = [] { const char* e = getenv("BLZ_CLUSTER_CFG"); return e ? atoi(e) : -1; }();   // tuning aid (scripts/inst_cluster_microbench.py)
    // defaults from the s2x sweeps (profiles/r01l_cluster_sweep.txt): 4 records per thread x 2 CTAs/SM without Hi-Z (no spills, the
    // prefetch hides the gathers), 3 x 3 with Hi-Z (the projection / Hi-Z arithmetic wants resident warps)
    if (p.mode == 0u) return launch_cluster_passthrough(p, numSMs, stream);
    if (p.mode != 2u) return launch_cluster_hiz<HIZ_NONE>(p, envCfg >= 0 ? envCfg : 0, numSMs, stream);
    const int cfg = envCfg >= 0 ? envCfg : 4;
    return hiz == HIZ_VK ? launch_cluster_hiz<HIZ_VK>(p, cfg, numSMs, stream) : launch_cluster_hiz<HIZ_DX>(p, cfg, numSMs, stream);
}

} // namespace blz
