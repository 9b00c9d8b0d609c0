// Draw-list consumer: replays on the device what the step AFTER the cull reads from its outputs, without a rasteriser.
//   * indexed-indirect path: vkCmdDrawIndexedIndirectCount walks `count` records at stride 24 / 32 (BlitzenVulkan/vulkanDraw.cpp:470-471); the
//     vertex stage does draws[gl_DrawID].objectId -> RenderObject -> transform / surface (VulkanShaders/MainObjectShader.vert.glsl:26).
//     The consumer follows the same chain for every record, checks that the object exists and that {indexCount, firstIndex} is one of the
//     LODs of that object's surface, and reduces the list to a summary: record count, sum of indexCount (the index fetches the draw would
//     issue), sum / xor of the objectIds, per-LOD histogram, number of adjacent records out of ascending order.
//   * instanced path: one command per LOD; the vertex stage does instIndices[objId + SV_InstanceID] (HlslShaders/VS/opaqueDrawInst.vs.hlsl:11).
//     The consumer walks every bucket the commands reference.
// The summary is a size-independent checksum of a whole frame's output: tests compare it with the same reduction of the oracle's list.
#include "ctx.h"

namespace blz {

namespace {

struct DeviceSummary {
    unsigned long long records, indexSum, instanceSum, idSum, idXor;
    uint32_t badObject, badLod, unsorted, pad;
    uint32_t lodHist[256];
};

constexpr int kConsumeThreads = 256;

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    return v;
}
__device__ __forceinline__ unsigned long long warp_xor(unsigned long long v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v ^= __shfl_xor_sync(0xFFFFFFFFu, v, d);
    return v;
}

struct ConsumeParams {
    const uint32_t* draws; const uint32_t* counts; uint32_t recWords;
    const RenderObject* objs; uint32_t n, objectIdBase;
    const PrimitiveSurface* surfaces; uint32_t surfaceCount;
    const LodData* lods; uint32_t lodCount;
    int kind;                        // 0: object draws (LOD check), 1: cluster draws (no LOD check)
    DeviceSummary* out;
};

__global__ void __launch_bounds__(kConsumeThreads) consume_draws_kernel(const ConsumeParams p)
{
    __shared__ uint32_t s_hist[256];
    for (uint32_t l = threadIdx.x; l < 256u; l += kConsumeThreads) s_hist[l] = 0u;
    __syncthreads();
    const uint32_t count = p.counts[0];
    unsigned long long idx = 0, inst = 0, ids = 0, idx_ = 0;
    uint32_t badObj = 0, badLod = 0, unsorted = 0, recs = 0;
    for (uint32_t r = blockIdx.x * kConsumeThreads + threadIdx.x; r < count; r += gridDim.x * kConsumeThreads) {
        const uint2 a0 = *reinterpret_cast<const uint2*>(p.draws + size_t(r) * p.recWords), a1 = *reinterpret_cast<const uint2*>(p.draws + size_t(r) * p.recWords + 2);
        const uint4 a = make_uint4(a0.x, a0.y, a1.x, a1.y);                                     // {objectId, indexCount, instanceCount, firstIndex}; 24-byte records are 8-byte aligned
        const uint32_t objectId = a.x, indexCount = a.y, instanceCount = a.z, firstIndex = a.w;
        ++recs; idx += indexCount; inst += instanceCount; ids += objectId; idx_ ^= (unsigned long long)objectId * 0x9E3779B97F4A7C15ull;
        if (r > 0u && p.draws[size_t(r - 1u) * p.recWords] > objectId) ++unsorted;          // cluster draws repeat an objectId, object draws never do
        const uint32_t local = objectId - p.objectIdBase;
        if (local >= p.n) { ++badObj; continue; }
        if (p.kind != 0) continue;
        const uint32_t sid = p.objs[local].surfaceId;
        if (sid >= p.surfaceCount) { ++badObj; continue; }
        const uint32_t lo = p.surfaces[sid].lodOffset, lc = p.surfaces[sid].lodCount;
        uint32_t found = 0xFFFFFFFFu;
        for (uint32_t l = lo; l < lo + lc && l < p.lodCount; ++l)
            if (p.lods[l].indexCount == indexCount && p.lods[l].firstIndex == firstIndex) { found = l; break; }
        if (found == 0xFFFFFFFFu) ++badLod; else atomicAdd(&s_hist[found & 255u], 1u);
    }
    const unsigned long long rr = warp_sum(recs), ii = warp_sum(idx), nn = warp_sum(inst), ss = warp_sum(ids), xx = warp_xor(idx_);
    const unsigned long long bo = warp_sum(badObj), bl = warp_sum(badLod), us = warp_sum(unsorted);
    if ((threadIdx.x & 31u) == 0u) {
        atomicAdd(&p.out->records, rr); atomicAdd(&p.out->indexSum, ii); atomicAdd(&p.out->instanceSum, nn); atomicAdd(&p.out->idSum, ss);
        atomicXor(&p.out->idXor, xx);
        if (bo) atomicAdd(&p.out->badObject, uint32_t(bo));
        if (bl) atomicAdd(&p.out->badLod, uint32_t(bl));
        if (us) atomicAdd(&p.out->unsorted, uint32_t(us));
    }
    __syncthreads();
    for (uint32_t l = threadIdx.x; l < 256u; l += kConsumeThreads) if (s_hist[l]) atomicAdd(&p.out->lodHist[l], s_hist[l]);
}

struct ConsumeInstParams {
    const uint32_t* cmds; const uint32_t* counts;            // DX32 commands {instanceOffset, indexCount, instCount, indexOffset, ...}
    const uint32_t* instanceIndices;
    const LodInstanceCounter* lodInstances; const LodData* lods; uint32_t lodCount;
    uint32_t n, objectIdBase;
    DeviceSummary* out;
};

// grid.y = command index; the blocks of a row stride over that command's instances
__global__ void __launch_bounds__(kConsumeThreads) consume_instances_kernel(const ConsumeInstParams p)
{
    const uint32_t nCmd = p.counts[0];
    const uint32_t c = blockIdx.y;
    if (c >= nCmd) return;
    const uint32_t* cmd = p.cmds + size_t(c) * 8u;
    const uint32_t off = cmd[0], indexCount = cmd[1], instCount = cmd[2];
    uint32_t lodId = 0xFFFFFFFFu;                            // the LOD whose bucket starts at `off` (drawInstCmd.cs.hlsl writes objId = instanceOffset)
    for (uint32_t l = 0; l < p.lodCount; ++l) if (p.lodInstances[l].instanceOffset == off && p.lods[l].indexCount == indexCount) { lodId = l; break; }
    unsigned long long ids = 0, xr = 0, idx = 0;
    uint32_t badObj = 0, unsorted = 0, recs = 0;
    for (uint32_t i = blockIdx.x * kConsumeThreads + threadIdx.x; i < instCount; i += gridDim.x * kConsumeThreads) {
        const uint32_t id = p.instanceIndices[size_t(off) + i];
        ++recs; ids += id; xr ^= (unsigned long long)id * 0x9E3779B97F4A7C15ull; idx += indexCount;
        if (id - p.objectIdBase >= p.n) ++badObj;
        if (i > 0u && p.instanceIndices[size_t(off) + i - 1u] >= id) ++unsorted;
    }
    const unsigned long long rr = warp_sum(recs), ss = warp_sum(ids), xx = warp_xor(xr), ii = warp_sum(idx), bo = warp_sum(badObj), us = warp_sum(unsorted);
    if ((threadIdx.x & 31u) == 0u) {
        atomicAdd(&p.out->records, rr); atomicAdd(&p.out->idSum, ss); atomicXor(&p.out->idXor, xx); atomicAdd(&p.out->indexSum, ii); atomicAdd(&p.out->instanceSum, rr);
        if (bo) atomicAdd(&p.out->badObject, uint32_t(bo));
        if (us) atomicAdd(&p.out->unsorted, uint32_t(us));
        if (rr) { if (lodId == 0xFFFFFFFFu) atomicAdd(&p.out->badLod, 1u); else atomicAdd(&p.out->lodHist[lodId & 255u], uint32_t(rr)); }
    }
}

} // namespace

} // namespace blz

using namespace blz;

static_assert(sizeof(blz_consume_summary) == sizeof(DeviceSummary), "summary layout");

extern "C" {

int blz_cull_consume_draws(blz_cull_ctx* c, int list, int kind, blz_consume_summary* out)
{
    if (!c || !out) return fail(BLZ_ERR_INVALID, "null argument");
    if (list < 0 || list > 2 || !c->draws || !c->objs[list]) return fail(BLZ_ERR_INVALID, "no scene / list %d", list);
    if (kind != 0 && kind != 1) return fail(BLZ_ERR_INVALID, "kind %d", kind);
    CU_TRY(cudaSetDevice(c->device));
    DeviceSummary* d = nullptr;
    CU_TRY(cudaMalloc(&d, sizeof(DeviceSummary)));
    CU_TRY(cudaMemsetAsync(d, 0, sizeof(DeviceSummary), c->stream));
    ConsumeParams p{};
    p.draws = c->draws; p.counts = c->drawCounts; p.recWords = c->lastRecWords;
    p.objs = c->objs[list]; p.n = c->nObjs[list]; p.objectIdBase = list == BLZ_LIST_OPAQUE ? c->objectIdBase : 0u;
    p.surfaces = c->surf; p.surfaceCount = c->nSurf; p.lods = c->lods; p.lodCount = c->nLods; p.kind = kind; p.out = d;
    consume_draws_kernel<<<c->numSMs * 4, kConsumeThreads, 0, c->stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d, sizeof(DeviceSummary), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    c->launches++;
    if (e != cudaSuccess) return fail(BLZ_ERR_CUDA, "consume_draws: %s", cudaGetErrorString(e));
    return BLZ_OK;
}

// The same reduction over the list the ranks GATHERED on this (the presenting) context for `epoch` (gather.cu): the presenter does not hold the
// other ranks' objects, so only the record-level part of the summary is produced (kind 1: no object / LOD look-ups).  Equality of this summary
// with the sum (xor for id_xor) of the ranks' own blz_cull_consume_draws(kind 1) summaries proves the gather on the device, at full size.
int blz_cull_consume_gathered(blz_cull_ctx* c, uint32_t epoch, blz_consume_summary* out)
{
    if (!c || !out) return fail(BLZ_ERR_INVALID, "null argument");
    if (!c->gatherOwner || !c->gatherBuf) return fail(BLZ_ERR_STATE, "only the presenting rank (the exporter) holds a gathered list");
    CU_TRY(cudaSetDevice(c->device));
    uint32_t counts[64];
    int rc = blz_cull_gather_read(c, epoch, nullptr, 0, counts);            // waits (on the device) for every rank's push of that epoch
    if (rc) return rc;
    uint64_t total = 0;
    for (int r = 0; r < c->world; ++r) total += counts[r];
    if (total > c->gatherCap) total = c->gatherCap;
    const uint32_t total32 = uint32_t(total);
    DeviceSummary* d = nullptr;
    CU_TRY(cudaMalloc(&d, sizeof(DeviceSummary) + 16));
    CU_TRY(cudaMemsetAsync(d, 0, sizeof(DeviceSummary), c->stream));
    uint32_t* cnt = reinterpret_cast<uint32_t*>(d + 1);
    CU_TRY(cudaMemcpyAsync(cnt, &total32, sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    ConsumeParams p{};
    p.draws = c->gatherBuf + size_t(epoch & 1u) * c->gatherCap * c->gatherRecWords; p.counts = cnt; p.recWords = c->gatherRecWords;
    p.objs = nullptr; p.n = 0xFFFFFFFFu; p.objectIdBase = 0u; p.kind = 1; p.out = d;
    consume_draws_kernel<<<c->numSMs * 4, kConsumeThreads, 0, c->stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d, sizeof(DeviceSummary), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    c->launches++;
    if (e != cudaSuccess) return fail(BLZ_ERR_CUDA, "consume_gathered: %s", cudaGetErrorString(e));
    return BLZ_OK;
}

int blz_cull_consume_instances(blz_cull_ctx* c, int list, blz_consume_summary* out)
{
    if (!c || !out) return fail(BLZ_ERR_INVALID, "null argument");
    if (list < 0 || list > 2 || !c->draws || !c->lodInst || !c->instIdx) return fail(BLZ_ERR_INVALID, "scene was uploaded without lod_instances");
    CU_TRY(cudaSetDevice(c->device));
    DeviceSummary* d = nullptr;
    CU_TRY(cudaMalloc(&d, sizeof(DeviceSummary)));
    CU_TRY(cudaMemsetAsync(d, 0, sizeof(DeviceSummary), c->stream));
    ConsumeInstParams p{};
    p.cmds = c->draws; p.counts = c->drawCounts; p.instanceIndices = c->instIdx; p.lodInstances = c->lodInst; p.lods = c->lods; p.lodCount = c->nLods;
    p.n = c->nObjs[list]; p.objectIdBase = list == BLZ_LIST_OPAQUE ? c->objectIdBase : 0u; p.out = d;
    dim3 grid(uint32_t(c->numSMs), c->nLods ? c->nLods : 1u);
    consume_instances_kernel<<<grid, kConsumeThreads, 0, c->stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d, sizeof(DeviceSummary), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    c->launches++;
    if (e != cudaSuccess) return fail(BLZ_ERR_CUDA, "consume_instances: %s", cudaGetErrorString(e));
    return BLZ_OK;
}

} // extern "C"
