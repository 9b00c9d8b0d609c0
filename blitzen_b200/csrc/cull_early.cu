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

__global__ void __launch_bounds__(kEarlyThreads, 6) early_sparse_kernel(const __grid_constant__ DrawCullParams p)
{
    __shared__ uint16_t s_ids[kEarlyTile];          // index in tile of the objects that were visible last frame, ascending
    __shared__ uint32_t s_desc[kEarlyTile];         // survivors: index in tile | lodId << 12
    __shared__ uint32_t s_part[kEarlyVec * kEarlyWarps];
    __shared__ uint32_t s_warpCnt[kEarlyWarps];
    __shared__ uint32_t s_tile, s_nAct;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t laneLt = (1u << lane) - 1u;
    const ViewConsts& V = p.view;
    uint32_t epoch;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(epoch) : "l"(&p.ctl->epoch));

    if (tid == 0) s_tile = atomicAdd(&p.ctl->ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t tileBase = tile * uint32_t(kEarlyTile);

    // ---- 1. visibility words: 4 x 128-bit per thread, object = tileBase + k*1024 + tid*4 + c -----------------------------
    uint32_t mask[kEarlyVec];
    {
        uint4 w[kEarlyVec];
#pragma unroll
        for (int k = 0; k < kEarlyVec; ++k) {
            const uint32_t i = tileBase + uint32_t(k) * 1024u + tid * 4u;
            w[k] = make_uint4(0u, 0u, 0u, 0u);
            if (i + 3u < p.n) w[k] = ldg_nc_u4(p.visibility + i);                 // the visibility buffer is padded to a multiple of 4 words
            else if (i < p.n) {
                w[k].x = __ldg(p.visibility + i);
                if (i + 1u < p.n) w[k].y = __ldg(p.visibility + i + 1u);
                if (i + 2u < p.n) w[k].z = __ldg(p.visibility + i + 2u);
            }
        }
#pragma unroll
        for (int k = 0; k < kEarlyVec; ++k)
            mask[k] = (w[k].x != 0u ? 1u : 0u) | (w[k].y != 0u ? 2u : 0u) | (w[k].z != 0u ? 4u : 0u) | (w[k].w != 0u ? 8u : 0u);
    }
    // ---- 2. ordered compaction of the visible ids into shared memory --------------------------------------------------------
    uint32_t off[kEarlyVec];
#pragma unroll
    for (int k = 0; k < kEarlyVec; ++k) {
        const uint32_t c = uint32_t(__popc(mask[k]));
        uint32_t x = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d); if (lane >= uint32_t(d)) x += y; }
        off[k] = x - c;
        if (lane == 31) s_part[k * kEarlyWarps + warp] = x;
    }
    __syncthreads();
    if (warp == 0) {                                                              // exclusive scan of the 32 (k, warp) partial counts
        const uint32_t c = s_part[lane];
        uint32_t x = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d); if (lane >= uint32_t(d)) x += y; }
        s_part[lane] = x - c;
        if (lane == 31) s_nAct = x;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kEarlyVec; ++k) {
        uint32_t o = s_part[k * kEarlyWarps + warp] + off[k];
        const uint32_t l0 = uint32_t(k) * 1024u + tid * 4u;
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (mask[k] & (1u << c)) s_ids[o++] = uint16_t(l0 + uint32_t(c));
    }
    __syncthreads();
    const uint32_t nAct = s_nAct;

    // ---- 3. dense cull of the compacted ids, 256 at a time --------------------------------------------------------------------
    uint32_t emitted = 0u;                                                        // survivors of the batches so far (uniform)
    for (uint32_t b0 = 0; b0 < nAct; b0 += kEarlyThreads) {
        const uint32_t e = b0 + tid;
        bool emit = false;
        uint32_t local = 0u, lodId = 0u;
        if (e < nAct) {
            local = s_ids[e];
            const uint2 ob = __ldg(reinterpret_cast<const uint2*>(p.objs + tileBase + local));
            const uint32_t t = ob.x - p.transformIdBase;
            const float4 ps = __ldg(p.xfPosScale + t), qt = __ldg(p.xfQuat + t);
            const float4 bs = __ldg(reinterpret_cast<const float4*>(p.surfaces + ob.y));
            const Sphere s = view_space_sphere(bs.x, bs.y, bs.z, bs.w, ps.x, ps.y, ps.z, ps.w, qt.x, qt.y, qt.z, qt.w, V);
            if (frustum_test(s, V)) {
                emit = true;
                const uint4 hi = __ldg(reinterpret_cast<const uint4*>(p.surfaces + ob.y) + 1);         // {materialId, lodOffset, lodCount, vertexOffset}
                const uint32_t rel = lod_select(s, ps.w, V.lodTarget, hi.y, hi.z, [&](uint32_t li) { return __ldg(&p.lods[li].error); });
                lodId = (p.flags & kFlagOnpcLodQuirk) ? rel : rel + hi.y;
            }
        }
        const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, emit);
        if (lane == 0) s_warpCnt[warp] = uint32_t(__popc(ballot));
        __syncthreads();
        uint32_t warpOff = 0u, batchTotal = 0u;
#pragma unroll
        for (int w = 0; w < kEarlyWarps; ++w) { const uint32_t c = s_warpCnt[w]; if (uint32_t(w) < warp) warpOff += c; batchTotal += c; }
        if (emit) s_desc[emitted + warpOff + uint32_t(__popc(ballot & laneLt))] = local | (lodId << 12);
        emitted += batchTotal;
        __syncthreads();                                                          // s_warpCnt is rewritten by the next batch
    }

    // ---- 4. cross-tile offset (decoupled look-back) + contiguous record span -------------------------------------------------
    {
        const uint64_t prefix = lookback_exclusive_prefix_cta<kEarlyThreads>(p.status, tile, emitted, epoch, s_part);   // s_part: 32 words, free by now
        if (tid == 0 && tile == p.numTiles - 1u) {
            const uint64_t total = prefix + emitted;
            p.counts[0] = uint32_t(total < p.capacity ? total : p.capacity);
            p.counts[1] = uint32_t(total);
        }
        const uint64_t room = prefix < p.capacity ? p.capacity - prefix : 0ull;
        const uint32_t nrec = uint32_t(room < emitted ? room : emitted);
        uint2* dst = reinterpret_cast<uint2*>(p.draws + prefix * p.recWords);
        const uint32_t idBase = p.objectIdBase + tileBase;
        const uint32_t wpr = p.recWords >> 1;
        for (uint32_t w = tid; w < nrec * wpr; w += kEarlyThreads) {
            const uint32_t r = wpr == 3u ? w / 3u : w >> 2, f = w - r * wpr;
            const uint32_t d = s_desc[r];
            const uint2 L = __ldg(reinterpret_cast<const uint2*>(p.lods + (d >> 12)));       // {indexCount, firstIndex}
            st_cs_u2(dst + w, f == 0u ? make_uint2(idBase + (d & 4095u), L.x) : (f == 1u ? make_uint2(1u, L.y) : make_uint2(0u, 0u)));
        }
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

// ------------------------------------------------------------------------------------------------------------------------
// Early pass over the late pass's visible list.  The only writer of the visibility buffer is the late pass
// (LateDrawCull.comp.glsl:70), and cull_draw.cu's PASS_LATE kernel emits, next to visibility[], the ascending list of the ids
// it set to 1.  While nothing else has touched visibility[] since (the C-ABI layer tracks that), "if (visibility[i] == 0)
// return" is the same as "i is in the list", and the early pass needs no stream over N objects at all: v * (4 + 40) bytes
// instead of N * 4 + v * 40.  Persistent CTAs, tiles of 512 list entries (two per thread, both gathers in flight together),
// deterministic order (list order == ascending id), CTA-wide decoupled look-back.
// ------------------------------------------------------------------------------------------------------------------------
constexpr int kListThreads = 256;
constexpr int kListItems = 2;
constexpr int kListTile = kListThreads * kListItems;

__global__ void __launch_bounds__(kListThreads, 6) early_list_kernel(const __grid_constant__ DrawCullParams p)
{
    constexpr int WARPS = kListThreads / 32;
    __shared__ uint2 s_desc[kListTile];             // survivors: {local object index, lodId}
    __shared__ uint32_t s_cnt[kListItems * WARPS];
    __shared__ uint32_t s_scratch[2 * WARPS + 2];
    __shared__ uint32_t s_tile;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t laneLt = (1u << lane) - 1u;
    const ViewConsts& V = p.view;
    uint32_t epoch, v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(epoch) : "l"(&p.ctl->epoch));
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p.visCount));
    const uint32_t numTiles = (v + uint32_t(kListTile) - 1u) / uint32_t(kListTile);
    if (v == 0u && blockIdx.x == 0 && tid == 0) { p.counts[0] = 0u; p.counts[1] = 0u; }

    while (true) {
        __syncthreads();                                                          // previous tile fully written out (s_desc, s_tile reuse)
        if (tid == 0) s_tile = atomicAdd(&p.ctl->ticket, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= numTiles) break;
        uint32_t idx[kListItems]; bool have[kListItems];
#pragma unroll
        for (int k = 0; k < kListItems; ++k) {
            const uint32_t e = tile * uint32_t(kListTile) + uint32_t(k) * kListThreads + tid;
            have[k] = e < v;
            idx[k] = have[k] ? __ldg(p.visList + e) : 0u;
        }
        uint2 ob[kListItems];
#pragma unroll
        for (int k = 0; k < kListItems; ++k) ob[k] = have[k] ? __ldg(reinterpret_cast<const uint2*>(p.objs + idx[k])) : make_uint2(p.transformIdBase, 0u);
        float4 ps[kListItems], qt[kListItems], bs[kListItems];
#pragma unroll
        for (int k = 0; k < kListItems; ++k) {
            const uint32_t t = ob[k].x - p.transformIdBase;
            ps[k] = have[k] ? __ldg(p.xfPosScale + t) : make_float4(0.f, 0.f, 0.f, 1.f);
            qt[k] = have[k] ? __ldg(p.xfQuat + t) : make_float4(0.f, 0.f, 0.f, 1.f);
            bs[k] = have[k] ? __ldg(reinterpret_cast<const float4*>(p.surfaces + ob[k].y)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        uint32_t lodId[kListItems], rank[kListItems], ballots[kListItems];
        bool emit[kListItems];
#pragma unroll
        for (int k = 0; k < kListItems; ++k) {
            emit[k] = false; lodId[k] = 0u;
            if (have[k]) {
                const Sphere s = view_space_sphere(bs[k].x, bs[k].y, bs[k].z, bs[k].w, ps[k].x, ps[k].y, ps[k].z, ps[k].w, qt[k].x, qt[k].y, qt[k].z, qt[k].w, V);
                if (frustum_test(s, V)) {
                    emit[k] = true;
                    const uint4 hi = __ldg(reinterpret_cast<const uint4*>(p.surfaces + ob[k].y) + 1);         // {materialId, lodOffset, lodCount, vertexOffset}
                    const uint32_t rel = lod_select(s, ps[k].w, V.lodTarget, hi.y, hi.z, [&](uint32_t li) { return __ldg(&p.lods[li].error); });
                    lodId[k] = (p.flags & kFlagOnpcLodQuirk) ? rel : rel + hi.y;
                }
            }
            ballots[k] = __ballot_sync(0xFFFFFFFFu, emit[k]);
            rank[k] = uint32_t(__popc(ballots[k] & laneLt));
            if (lane == 0) s_cnt[k * WARPS + warp] = uint32_t(__popc(ballots[k]));
        }
        __syncthreads();
        uint32_t total = 0u, offs[kListItems];
#pragma unroll
        for (int k = 0; k < kListItems; ++k) {
            offs[k] = 0u;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) { const uint32_t c = s_cnt[k * WARPS + w]; if (uint32_t(w) < warp) offs[k] += c; }
        }
#pragma unroll
        for (int k = 0; k < kListItems; ++k) {                                   // entries of item 0 precede those of item 1 in the list
            uint32_t kTotal = 0u;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) kTotal += s_cnt[k * WARPS + w];
            if (emit[k]) s_desc[total + offs[k] + rank[k]] = make_uint2(idx[k], lodId[k]);
            total += kTotal;
        }
        const uint64_t prefix = lookback_exclusive_prefix_cta<kListThreads>(p.status, tile, total, epoch, s_scratch);   // contains the barrier that publishes s_desc
        if (tid == 0 && tile == numTiles - 1u) {
            const uint64_t all = prefix + total;
            p.counts[0] = uint32_t(all < p.capacity ? all : p.capacity);
            p.counts[1] = uint32_t(all);
        }
        __syncthreads();
        const uint64_t room = prefix < p.capacity ? p.capacity - prefix : 0ull;
        const uint32_t nrec = uint32_t(room < total ? room : total);
        uint2* dst = reinterpret_cast<uint2*>(p.draws + prefix * p.recWords);
        const uint32_t wpr = p.recWords >> 1;
        for (uint32_t w = tid; w < nrec * wpr; w += kListThreads) {
            const uint32_t r = wpr == 3u ? w / 3u : w >> 2, f = w - r * wpr;
            const uint2 d = s_desc[r];
            const uint2 L = __ldg(reinterpret_cast<const uint2*>(p.lods + d.y));             // {indexCount, firstIndex}
            st_cs_u2(dst + w, f == 0u ? make_uint2(p.objectIdBase + d.x, L.x) : (f == 1u ? make_uint2(1u, L.y) : make_uint2(0u, 0u)));
        }
    }
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

cudaError_t launch_early_list(const DrawCullParams& p, int numSMs, cudaStream_t stream)
{
    uint32_t maxTiles = p.n == 0 ? 1u : uint32_t((uint64_t(p.n) + kListTile - 1) / kListTile);
    uint32_t grid = uint32_t(numSMs) * 6u;
    if (grid > maxTiles) grid = maxTiles;
    early_list_kernel<<<grid, kListThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_early_sparse(const DrawCullParams& p, cudaStream_t stream)
{
    if (p.lodCount >= (1u << 20)) return cudaErrorInvalidValue;                   // descriptor packing: 12 bits of index + 20 bits of lod id
    DrawCullParams q = p;
    q.numTiles = p.n == 0 ? 1u : uint32_t((uint64_t(p.n) + kEarlyTile - 1) / kEarlyTile);
    early_sparse_kernel<<<q.numTiles, kEarlyThreads, 0, stream>>>(q);
    return cudaGetLastError();
}

} // namespace blz
