// Private: the context behind the opaque blz_cull_ctx handle, shared by capi.cu and gather.cu.
#pragma once
#include "../../include/blz_cull.h"
#include "cull_kernels.cuh"
#include <string>

namespace blz {
// a device allocation made with the virtual memory management API so that it can be exported as a file descriptor (interop.cu)
struct ExportableBuffer { void* ptr = nullptr; size_t size = 0; uint64_t handle = 0; bool active = false; };
int exportable_alloc(int device, size_t bytes, ExportableBuffer& b);
void exportable_free(ExportableBuffer& b);
}

struct blz_cull_ctx {
    int device = 0, numSMs = 0;
    cudaStream_t stream = nullptr, ownStream = nullptr;
    // scene (device)
    blz::RenderObject* objs[3] = { nullptr, nullptr, nullptr };
    uint32_t nObjs[3] = { 0, 0, 0 };
    blz::MeshTransform* xf = nullptr; uint32_t nXf = 0;       // AoS as uploaded: the kernels read one 32-byte record with one 256-bit load
    blz::PrimitiveSurface* surf = nullptr; uint32_t nSurf = 0;
    blz::LodData* lods = nullptr; uint32_t nLods = 0;
    blz::Cluster* clusters = nullptr; uint32_t nClusters = 0;
    blz::LodInstanceCounter* lodInst = nullptr; uint32_t nLodInst = 0;
    uint32_t* bucketCap = nullptr;
    uint32_t objectIdBase = 0, transformIdBase = 0;
    // allocation sizes in bytes (buffers only ever grow; see grow() in capi.cu)
    size_t capObjs[3] = { 0, 0, 0 }, capXf = 0, capSurf = 0, capLods = 0, capClusters = 0, capLodInst = 0, capBucket = 0;
    size_t capVis = 0, capDraws = 0, capDispatch = 0, capInstIdx = 0;
    // per-object state + outputs
    uint32_t* vis = nullptr;
    uint32_t* draws = nullptr; uint64_t drawCap = 0;
    uint32_t* counts = nullptr;               // [0..1] draws (slot 0), [2..3] cluster dispatch, [5] early-pass density accumulator, [6..7] survivor list, [8..9] draws (slot 1)
    uint32_t* drawCounts = nullptr;           // counts + 0 or counts + 8: {written, total} of the current draw buffer
    uint32_t* visBits = nullptr; size_t capVisBits = 0;   // 1 bit per object, padded to whole early-pass tiles
    bool visBitsValid = false;                // visBits is current
    bool visWordsValid = false;               // vis (u32 per object, the reference's form) is current; at least one of the two always is
    uint32_t* dispatch = nullptr; uint64_t dispatchCap = 0;
    uint32_t* instIdx = nullptr; uint64_t instCap = 0;
    uint32_t* visTotalHost = nullptr;         // pinned: number of previously-visible objects the last COMPLETED early pass saw (lags by a frame; a hint only)
    bool earlyDense = false;                  // early pass currently runs the dense (streaming) kernel instead of the sparse pipelined one
    uint2* survList = nullptr; size_t capSurvList = 0;       // {objectId, absolute LOD id} of the frustum survivors (instancing / cluster expand, cull_list.cu)
    uint32_t* listScratch = nullptr; size_t capListScratch = 0;   // per-tile histograms / record counts of the survivor-list kernels
    blz::ScanCtl* ctl = nullptr;
    uint64_t* status = nullptr; size_t statusEntries = 0;
    // depth + pyramid
    const float* depth = nullptr; float* depthOwned = nullptr; size_t depthOwnedTexels = 0;
    uint32_t depthW = 0, depthH = 0;
    float* pyrData = nullptr; size_t pyrTexels = 0;
    blz::PyramidDesc pyr{};
    int pyrVariant = -1;
    uint32_t* pyrTicket = nullptr;
    // view
    blz::CameraViewData view{}; bool haveView = false;
    uint64_t launches = 0;
    uint32_t epochLaunches = 1, epochWrapAt = (1u << 30) - 1024u, epochRestarts = 0;   // host mirror of ScanCtl::epoch (scan_epoch_guard in capi.cu)
    int64_t optPyramidTma = 1;
    int64_t optEarlyMode = 1;                 // 1 = pipelined visibility-stream kernel (cull_early.cu), 0 = the streaming kernel (cull_stream.cu, PASS_EARLY) whatever the density
    int64_t optEarlyBits = 1;                 // pipelined early pass streams the 1-bit mask (1) or the 4-B visibility words (0)
    int64_t optVisWords = 0;                  // 1: the streaming late pass also writes the u32-per-object visibility buffer every frame (else on demand)
    int64_t optEarlyAuto = 1;                 // early_mode 1: switch to the streaming kernel while more than ~20 % of the objects were visible last frame
    bool sceneRefused = false;                // the last upload_scene failed validation
    int64_t optValidate = 1;                  // upload_scene checks every id the kernels will index with (one pass + one 16-byte read-back)
    int64_t optStreamCfg = 2;                 // CTA shape of the streaming kernel (see launch_pass in cull_stream.cu)
    uint32_t lastRecWords = 6;                // record width (u32 words) of the pass that last wrote `draws`
    // gather (multi-GPU): the presenter owns gatherBuf/gatherFlags; every rank (presenter included) writes through gatherDst*
    uint32_t* gatherBuf = nullptr; uint64_t gatherCap = 0; uint32_t gatherRecWords = 6; uint64_t* gatherFlags = nullptr; bool gatherOwner = false;
    uint32_t* gatherDst = nullptr; uint64_t* gatherDstFlags = nullptr; int rank = 0, world = 1; bool gatherImported = false, gatherPeerMapped = false;
    uint32_t* gatherDone = nullptr;
    uint32_t* gatherErrHost = nullptr; uint32_t* gatherErrDev = nullptr; int64_t optGatherTimeoutMs = 60000;   // bounded device-side waits (gather.cu: SpinGuard)
    uint32_t* instDst = nullptr; bool instDstMapped = false;   // presenter's instance index buffer (instance-list gather)
    // descriptor transport (gather.cu): the draw passes also write {objectId, lodId} per record; the ranks ship those 8 bytes instead of the 24/32-byte
    // records and the presenter expands them with its own LOD table.  descs / descsAlt flip together with draws / drawsAlt.
    uint2* descs = nullptr; uint2* descsAlt = nullptr; size_t capDescs = 0; bool descValid = false, descValidAlt = false;
    int64_t optGatherDesc = 1;
    // asynchronous push: the list just pushed stays readable in `drawsAlt` while the next pass writes `draws` (blz_cull_gather_push_async)
    uint32_t* drawsAlt = nullptr; uint32_t lastRecWordsAlt = 6; int drawSlot = 0;
    // zero-copy export of the outputs (interop.cu): when active, `draws` / `counts` live in these allocations instead of cudaMalloc memory
    blz::ExportableBuffer expDraws, expCounts; uint32_t exportGeneration = 0;
    cudaEvent_t exportFence = nullptr; void* extSemaphore = nullptr; bool extSemaphoreTimeline = false;
    cudaStream_t gatherStream = nullptr; cudaEvent_t evCull = nullptr, evPush[2] = { nullptr, nullptr }; bool evPushValid[2] = { false, false };
    cudaEvent_t evExpand = nullptr; bool evExpandValid = false;    // descriptor mode, presenter: behind the expansion of the last asynchronous push
    bool drawBufferPending = false;           // the current draw buffer's previous asynchronous push has not been waited for yet
};

namespace blz {
int fail(int code, const char* fmt, ...);
int gather_report_timeout(blz_cull_ctx* c);      // gather.cu: BLZ_ERR_TIMEOUT if a device-side wait of the draw-list gather gave up (call behind a host synchronisation)

// the part of the 256-byte view block the kernels read, as kernel-parameter constants
inline ViewConsts make_view_consts(const CameraViewData& v)
{
    ViewConsts c;
    for (int col = 0; col < 4; ++col)
        for (int row = 0; row < 3; ++row) c.m[col * 3 + row] = v.view[col * 4 + row];
    c.frustumRight = v.frustumRight; c.frustumLeft = v.frustumLeft; c.frustumTop = v.frustumTop; c.frustumBottom = v.frustumBottom;
    c.proj0 = v.proj0; c.proj5 = v.proj5; c.zNear = v.zNear; c.zFar = v.zFar;
    c.pyramidWidth = v.pyramidWidth; c.pyramidHeight = v.pyramidHeight; c.lodTarget = v.lodTarget;
    return c;
}
}

#define CU_TRY(expr)                                                                                          \
    do {                                                                                                      \
        cudaError_t e__ = (expr);                                                                             \
        if (e__ != cudaSuccess) return blz::fail(BLZ_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)
