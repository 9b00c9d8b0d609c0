// C-ABI layer (include/blz_cull.h): the context that owns the device-side mirror of the reference's shared scene
// buffers and enqueues the cull kernels.  It plays the role of the backend class's SetupForRendering + the five cull
// dispatch functions inside DrawFrame (BlitzenVulkan/vulkanRendererSetup.cpp:365-666, vulkanDraw.cpp:107-226, :318-423,
// :554-622; BlitzenDX12/dx12Draw.cpp:114-413).  No torch types, no CPU fallback.
#include "ctx.h"
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <new>

namespace blz { void pyramid_plan(PyramidBuildParams& p); void gather_release(blz_cull_ctx* c); void interop_release(blz_cull_ctx* c); }
using namespace blz;

namespace {

thread_local std::string g_lastError;

} // namespace
namespace blz {
int fail(int code, const char* fmt, ...)
{
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    g_lastError = buf;
    return code;
}
} // namespace blz
namespace {


template <class T> void dfree(T*& p) { if (p) { cudaFree(p); p = nullptr; } }

} // namespace

namespace {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_tiled()
{
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}


int ensure_status(blz_cull_ctx* c, size_t entries)
{
    if (entries <= c->statusEntries) return BLZ_OK;
    dfree(c->status);
    c->statusEntries = 0;
    CU_TRY(cudaMalloc(&c->status, entries * sizeof(uint64_t)));
    CU_TRY(cudaMemsetAsync(c->status, 0, entries * sizeof(uint64_t), c->stream));
    c->statusEntries = entries;
    return BLZ_OK;
}

// Tile status words are never cleared: a word is valid only if it carries the current launch's 30-bit tag (ScanCtl::epoch, bumped by the last
// CTA out of every scan kernel).  The tag would wrap after 2^30 launches (about a day at several thousand launches per second) and a slot that
// smaller passes did not rewrite in the meantime would then be accepted with a stale value.  The host counts the same launches and,
// well before the wrap, clears the status array and restarts the tag -- stream-ordered, two tiny operations once per ~10^9 launches.
int scan_epoch_guard(blz_cull_ctx* c)
{
    if (++c->epochLaunches < c->epochWrapAt) return BLZ_OK;
    static const ScanCtl restart{ 0u, 0u, 1u, 0u };
    if (c->status && c->statusEntries) CU_TRY(cudaMemsetAsync(c->status, 0, c->statusEntries * sizeof(uint64_t), c->stream));
    CU_TRY(cudaMemcpyAsync(c->ctl, &restart, sizeof(restart), cudaMemcpyHostToDevice, c->stream));
    c->epochLaunches = 1;
    c->epochRestarts++;
    return BLZ_OK;
}

uint32_t tiles_for(uint64_t n) { return n == 0 ? 1u : uint32_t((n + kCullTile - 1) / kCullTile); }

void pyramid_layout(uint32_t depthW, uint32_t depthH, int variant, PyramidDesc& d, size_t& texels)
{
    uint32_t w, h;
    if (variant == BLZ_HIZ_VK) {
        // BlitML::PreviousPow2 (BlitzenMathLibrary/blitML.h:54-62): largest power of two r with r*2 < v ... i.e. strictly below v for v > 1
        auto prev = [](uint32_t v) { uint32_t r = 1; while (r * 2 < v) r *= 2; return r; };
        w = prev(depthW); h = prev(depthH);
    } else {
        w = depthW >> 1 ? depthW >> 1 : 1u; h = depthH >> 1 ? depthH >> 1 : 1u;      // dx12RNDResources.cpp:103-106
    }
    uint32_t mips = 0;                                                            // GetDepthPyramidMipLevels, blitML.h:95-107
    for (uint32_t a = w, b = h; a > 1 || b > 1; a /= 2, b /= 2) ++mips;
    d.width = w; d.height = h; d.mips = mips;
    size_t off = 0;
    for (uint32_t i = 0; i < 16; ++i) {
        d.offset[i] = uint32_t(off);
        if (i < mips) off += size_t((w >> i) ? (w >> i) : 1u) * ((h >> i) ? (h >> i) : 1u);
    }
    texels = off;
}

int ensure_pyramid_storage(blz_cull_ctx* c, int variant, uint32_t depthW, uint32_t depthH)
{
    PyramidDesc d{}; size_t texels = 0;
    pyramid_layout(depthW, depthH, variant, d, texels);
    if (d.mips == 0 || d.mips > 16) return fail(BLZ_ERR_INVALID, "depth %ux%u gives a pyramid with %u mips (need 1..16)", depthW, depthH, d.mips);
    if (texels > c->pyrTexels) {
        dfree(c->pyrData);
        c->pyrTexels = 0;
        CU_TRY(cudaMalloc(&c->pyrData, texels * sizeof(float)));
        c->pyrTexels = texels;
    }
    d.data = c->pyrData;
    c->pyr = d;
    c->pyrVariant = variant;
    // the backends publish the pyramid extent through the view block (vulkanRendererSetup.cpp:903-904, dx12Draw.cpp:563-564)
    c->view.pyramidWidth = float(d.width);
    c->view.pyramidHeight = float(d.height);
    return BLZ_OK;
}

// grows a device buffer only when it is too small, so that re-uploading a scene of the same shape (bench.py's end-to-end
// leg, streaming scenes) costs copies and one repack launch, not cudaMalloc/cudaFree
template <class T>
int grow(blz_cull_ctx* c, T*& p, size_t& capBytes, size_t needBytes)
{
    if (needBytes == 0) needBytes = 16;
    if (p && capBytes >= needBytes) return BLZ_OK;
    CU_TRY(cudaStreamSynchronize(c->stream));
    if (p) { cudaFree(p); p = nullptr; capBytes = 0; }
    CU_TRY(cudaMalloc(&p, needBytes));
    capBytes = needBytes;
    return BLZ_OK;
}
#define TRY_RC(expr) do { int rc__ = (expr); if (rc__) return rc__; } while (0)

// Visibility state has two forms: the reference's u32 per object (c->vis) and 1 bit per object (c->visBits), which is what the
// streaming late pass and the pipelined early pass read and write.  Each is converted from the other on demand.
int ensure_vis_bits(blz_cull_ctx* c)
{
    if (c->visBitsValid) return BLZ_OK;
    CU_TRY(launch_pack_vis_bits(c->vis, c->visBits, c->nObjs[0], c->stream));
    c->launches++;
    c->visBitsValid = true;
    return BLZ_OK;
}
int ensure_vis_words(blz_cull_ctx* c)
{
    if (c->visWordsValid) return BLZ_OK;
    CU_TRY(launch_unpack_vis_bits(c->visBits, c->vis, c->nObjs[0], c->stream));
    c->launches++;
    c->visWordsValid = true;
    return BLZ_OK;
}

// called by every pass that writes `draws`: after an asynchronous gather push flipped the buffers, the one about to be written may still be
// read by its previous push (gather.cu)
int acquire_draw_buffer(blz_cull_ctx* c)
{
    if (c->drawBufferPending) {
        CU_TRY(cudaStreamWaitEvent(c->stream, c->evPush[c->drawSlot], 0));
        c->drawBufferPending = false;
    }
    return BLZ_OK;
}

int check_list(blz_cull_ctx* c, int list)
{
    if (list < 0 || list > 2) return fail(BLZ_ERR_INVALID, "list %d out of range", list);
    if (!c->surf || !c->lods) return fail(BLZ_ERR_INVALID, "no scene uploaded");
    if (c->sceneRefused) return fail(BLZ_ERR_STATE, "no scene: the last upload was refused by validation");
    if (!c->haveView) return fail(BLZ_ERR_INVALID, "no view set");
    return BLZ_OK;
}

int run_draw_pass(blz_cull_ctx* c, int pass, int list, int fmt, int hiz, uint32_t flags)
{
    if (!c) return fail(BLZ_ERR_INVALID, "null context");
    int rc = check_list(c, list); if (rc) return rc;
    if (fmt != BLZ_REC_VK24 && fmt != BLZ_REC_DX32) return fail(BLZ_ERR_INVALID, "record format %d", fmt);
    if ((pass == PASS_EARLY || pass == PASS_LATE) && list != BLZ_LIST_OPAQUE) return fail(BLZ_ERR_INVALID, "two-phase passes run on the opaque list only");
    if (pass == PASS_LATE || pass == PASS_TEMPORAL) {
        if (hiz != BLZ_HIZ_VK && hiz != BLZ_HIZ_DX) return fail(BLZ_ERR_INVALID, "hiz variant %d", hiz);
        if (!c->pyrData || c->pyrVariant != hiz) return fail(BLZ_ERR_INVALID, "no depth pyramid of variant %d (call build_pyramid / clear_pyramid)", hiz);
    }
    CU_TRY(cudaSetDevice(c->device));
    DrawCullParams p{};
    p.objs = c->objs[list]; p.n = c->nObjs[list];
    p.xf = c->xf; p.surfaces = c->surf; p.lods = c->lods;
    TRY_RC(acquire_draw_buffer(c));
    p.visibility = c->vis; p.draws = c->draws; p.counts = c->drawCounts; p.ctl = c->ctl;
    // multi-GPU: the pass also leaves {objectId, lodId} per record for the gather to ship (8 bytes instead of 24 / 32 per record)
    p.descs = nullptr;
    if (c->gatherImported && c->optGatherDesc) {
        const size_t need = size_t(c->drawCap) * sizeof(uint2) + 16u;
        if (c->capDescs < need) {
            CU_TRY(cudaStreamSynchronize(c->stream));
            if (c->gatherStream) CU_TRY(cudaStreamSynchronize(c->gatherStream));
            if (c->descs) cudaFree(c->descs);
            if (c->descsAlt) cudaFree(c->descsAlt);
            c->descs = c->descsAlt = nullptr; c->capDescs = 0;
            CU_TRY(cudaMalloc(&c->descs, need));
            CU_TRY(cudaMalloc(&c->descsAlt, need));
            c->capDescs = need;
        }
        p.descs = c->descs;
    }
    c->descValid = p.descs != nullptr;
    p.numTiles = tiles_for(p.n);                      // re-derived from `items` by the launcher
    rc = ensure_status(c, p.n / kCullMinTile + 2u); if (rc) return rc;
    p.status = c->status;
    p.objectIdBase = list == BLZ_LIST_OPAQUE ? c->objectIdBase : 0u;
    p.transformIdBase = c->transformIdBase;
    p.surfaceCount = c->nSurf; p.lodCount = c->nLods;
    p.recWords = fmt == BLZ_REC_VK24 ? 6u : 8u;
    p.flags = flags & kFlagOnpcLodQuirk;
    p.capacity = c->drawCap;
    p.view = make_view_consts(c->view);
    p.pyr = c->pyr;
    // Early pass: the pipelined visibility-stream kernel (cull_early.cu) is built for a sparse visible set (0.06 ms at 3.7 % visible, but 0.45 ms
    // when everything was visible); the streaming kernel (cull_stream.cu, PASS_EARLY) takes 0.18 ms whatever the density.  The early pass itself
    // reports how many previously-visible objects it walked (device counter -> pinned host word, asynchronous): the value seen here lags by a
    // frame or two and only steers the choice of kernel.  Option early_mode = 0 forces the streaming kernel.
    bool earlyStream = pass == PASS_EARLY && c->optEarlyMode != 0 && p.lodCount < (1u << 20);
    const bool earlyAuto = earlyStream && c->optEarlyAuto && c->visTotalHost != nullptr;
    if (earlyAuto) {
        const uint32_t seen = *static_cast<volatile uint32_t*>(c->visTotalHost);
        if (!c->earlyDense && uint64_t(seen) * 5u > p.n) c->earlyDense = true;
        else if (c->earlyDense && uint64_t(seen) * 7u < p.n) c->earlyDense = false;
        if (c->earlyDense) earlyStream = false;
    }
    p.visTotal = (earlyAuto && earlyStream) ? c->counts + 5 : nullptr;   // stays 0 between launches (the kernel re-arms it)
    p.visTotalOut = c->visTotalHost;
    const bool usesBits = pass == PASS_LATE || (earlyStream && c->optEarlyBits);
    const bool usesWords = pass == PASS_EARLY && !usesBits;
    p.visBits = nullptr;
    if (usesBits) { TRY_RC(ensure_vis_bits(c)); p.visBits = c->visBits; }
    if (usesWords) TRY_RC(ensure_vis_words(c));
    if (pass == PASS_LATE && !c->optVisWords) p.visibility = nullptr;    // the mask is the state; the u32 form is materialised on demand
    TRY_RC(scan_epoch_guard(c));
    if (earlyStream) CU_TRY(launch_early_stream(p, c->numSMs, c->stream));
    else CU_TRY(launch_stream_cull(p, pass, hiz == BLZ_HIZ_DX ? HIZ_DX : HIZ_VK, int(c->optStreamCfg), c->numSMs, c->stream));
    if (earlyAuto && !earlyStream) {       // dense frame: the streaming kernel does not count; a 2 MB popcount + a 4-byte copy next to a 0.18 ms pass
        TRY_RC(ensure_vis_bits(c));
        CU_TRY(cudaMemsetAsync(c->counts + 5, 0, sizeof(uint32_t), c->stream));
        CU_TRY(launch_popc_vis_bits(c->visBits, p.n, c->counts + 5, c->stream)); c->launches++;
        CU_TRY(cudaMemcpyAsync(c->visTotalHost, c->counts + 5, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(cudaMemsetAsync(c->counts + 5, 0, sizeof(uint32_t), c->stream));
    }
    if (pass == PASS_LATE) {
        c->visBitsValid = true;
        c->visWordsValid = c->optVisWords != 0;
    }
    c->launches++;
    c->lastRecWords = p.recWords;
    return BLZ_OK;
}

// Step 1 of the survivor-list pipeline (cull_list.cu): the streaming frustum + LOD pass with 8-byte records {objectId, absolute LOD id};
// the count stays on the device at counts[6].
int run_survivor_list(blz_cull_ctx* c, int list)
{
    const uint32_t n = c->nObjs[list];
    TRY_RC(grow(c, c->survList, c->capSurvList, size_t(n) * sizeof(uint2) + 16u));
    DrawCullParams p{};
    p.objs = c->objs[list]; p.n = n;
    p.xf = c->xf; p.surfaces = c->surf; p.lods = c->lods;
    p.visibility = nullptr; p.draws = reinterpret_cast<uint32_t*>(c->survList); p.counts = c->counts + 6; p.ctl = c->ctl;
    p.numTiles = tiles_for(p.n);
    TRY_RC(ensure_status(c, p.n / kCullMinTile + 2u));
    p.status = c->status;
    p.objectIdBase = list == BLZ_LIST_OPAQUE ? c->objectIdBase : 0u;
    p.transformIdBase = c->transformIdBase;
    p.surfaceCount = c->nSurf; p.lodCount = c->nLods;
    p.recWords = 2u;
    p.flags = 0u;
    p.capacity = n;
    p.view = make_view_consts(c->view);
    p.pyr = c->pyr;
    TRY_RC(scan_epoch_guard(c));
    CU_TRY(launch_stream_cull(p, PASS_FRUSTUM, HIZ_VK, int(c->optStreamCfg), c->numSMs, c->stream));
    c->launches++;
    return BLZ_OK;
}

} // namespace

extern "C" {

int blz_cull_abi_version(void) { return BLZ_CULL_ABI_VERSION; }
const char* blz_cull_last_error(void) { return g_lastError.c_str(); }

static int create_resources(blz_cull_ctx* c);

int blz_cull_create(int device, blz_cull_ctx** out)
{
    if (!out) return fail(BLZ_ERR_INVALID, "out_ctx is null");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(BLZ_ERR_CUDA, "no CUDA device available (%s); this library has no CPU path", cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(BLZ_ERR_INVALID, "device %d out of range (0..%d)", device, n - 1);
    CU_TRY(cudaSetDevice(device));
    blz_cull_ctx* c = new (std::nothrow) blz_cull_ctx();
    if (!c) return fail(BLZ_ERR_INVALID, "out of host memory");
    c->device = device;
    const int rc = create_resources(c);
    if (rc) { const std::string msg = g_lastError; blz_cull_destroy(c); g_lastError = msg; return rc; }   // nothing allocated so far is leaked
    *out = c;
    return BLZ_OK;
}

static int create_resources(blz_cull_ctx* c)
{
    const int device = c->device;
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(BLZ_ERR_UNSUPPORTED, "device %s is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major, prop.minor);
    c->numSMs = prop.multiProcessorCount;
    CU_TRY(cudaStreamCreateWithFlags(&c->ownStream, cudaStreamNonBlocking));
    c->stream = c->ownStream;
    CU_TRY(cudaMalloc(&c->ctl, sizeof(ScanCtl)));
    ScanCtl init{ 0u, 0u, 1u, 0u };
    CU_TRY(cudaMemcpyAsync(c->ctl, &init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaMalloc(&c->counts, 16 * sizeof(uint32_t)));
    CU_TRY(cudaMemsetAsync(c->counts, 0, 16 * sizeof(uint32_t), c->stream));
    c->drawCounts = c->counts;
    CU_TRY(cudaHostAlloc(reinterpret_cast<void**>(&c->visTotalHost), sizeof(uint32_t), cudaHostAllocMapped));
    *c->visTotalHost = 0u;
    CU_TRY(cudaMalloc(&c->pyrTicket, sizeof(uint32_t)));
    CU_TRY(cudaMemsetAsync(c->pyrTicket, 0, sizeof(uint32_t), c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return BLZ_OK;
}

static void free_scene(blz_cull_ctx* c)
{
    for (int i = 0; i < 3; ++i) { dfree(c->objs[i]); c->nObjs[i] = 0; }
    dfree(c->xf); c->nXf = 0;
    dfree(c->surf); dfree(c->lods); dfree(c->clusters); dfree(c->lodInst); dfree(c->bucketCap);
    if (c->expDraws.active) { blz::exportable_free(c->expDraws); c->draws = nullptr; }
    dfree(c->descs); dfree(c->descsAlt); c->capDescs = 0; c->descValid = c->descValidAlt = false;
    dfree(c->vis); dfree(c->visBits); dfree(c->draws); dfree(c->drawsAlt); dfree(c->dispatch); dfree(c->instIdx); dfree(c->survList); dfree(c->listScratch); c->capSurvList = 0; c->capListScratch = 0;
    c->capVisBits = 0; c->visBitsValid = false;
    c->nSurf = c->nLods = c->nClusters = c->nLodInst = 0; c->drawCap = c->dispatchCap = c->instCap = 0;
    c->capObjs[0] = c->capObjs[1] = c->capObjs[2] = 0;
    c->capXf = c->capSurf = c->capLods = c->capClusters = c->capLodInst = c->capBucket = 0;
    c->capVis = c->capDraws = c->capDispatch = c->capInstIdx = 0;
}

int blz_cull_destroy(blz_cull_ctx* c)
{
    if (!c) return BLZ_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->gatherStream) cudaStreamSynchronize(c->gatherStream);
    free_scene(c);
    if (c->visTotalHost) { cudaFreeHost(c->visTotalHost); c->visTotalHost = nullptr; }
    if (c->expCounts.active) { blz::exportable_free(c->expCounts); c->counts = nullptr; }
    blz::interop_release(c);
    dfree(c->ctl); dfree(c->status); dfree(c->counts); dfree(c->depthOwned); dfree(c->pyrData); dfree(c->pyrTicket);
    blz::gather_release(c);
    if (c->ownStream) cudaStreamDestroy(c->ownStream);
    delete c;
    return BLZ_OK;
}

int blz_cull_set_stream(blz_cull_ctx* c, void* s)
{
    if (!c) return fail(BLZ_ERR_INVALID, "null context");
    c->stream = s ? reinterpret_cast<cudaStream_t>(s) : c->ownStream;
    return BLZ_OK;
}
int blz_cull_get_stream(blz_cull_ctx* c, void** out)
{
    if (!c || !out) return fail(BLZ_ERR_INVALID, "null argument");
    *out = c->stream; return BLZ_OK;
}
int blz_cull_synchronize(blz_cull_ctx* c)
{
    if (!c) return fail(BLZ_ERR_INVALID, "null context");
    CU_TRY(cudaSetDevice(c->device));
    CU_TRY(cudaStreamSynchronize(c->stream));
    if (c->gatherStream) CU_TRY(cudaStreamSynchronize(c->gatherStream));
    return gather_report_timeout(c);
}

namespace {

// Scene validation: the kernels index surfaces / transforms / LODs / clusters with the ids they find in the scene; a malformed
// scene must come back as BLZ_ERR_INVALID from upload_scene, not as an out-of-bounds access in a later pass.  One pass over what was
// just uploaded; the four violation counts travel back in one 16-byte copy.
__global__ void validate_objects_kernel(const RenderObject* objs, uint32_t n, uint32_t nSurf, uint32_t transformIdBase, uint32_t nXf, uint32_t* bad)
{
    uint32_t bs = 0u, bt = 0u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const RenderObject o = objs[i];
        if (o.surfaceId >= nSurf) ++bs;
        if (o.transformId - transformIdBase >= nXf) ++bt;
    }
    bs = __reduce_add_sync(0xFFFFFFFFu, bs); bt = __reduce_add_sync(0xFFFFFFFFu, bt);
    if ((threadIdx.x & 31u) == 0u) { if (bs) atomicAdd(bad + 0, bs); if (bt) atomicAdd(bad + 1, bt); }
}
__global__ void validate_tables_kernel(const PrimitiveSurface* surf, uint32_t nSurf, const LodData* lods, uint32_t nLods, uint32_t nClusters, uint32_t* bad)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nSurf && (surf[i].lodCount == 0u || uint64_t(surf[i].lodOffset) + surf[i].lodCount > nLods)) atomicAdd(bad + 2, 1u);
    if (i < nLods && nClusters != 0u && uint64_t(lods[i].clusterOffset) + lods[i].clusterCount > nClusters) atomicAdd(bad + 3, 1u);
}

int validate_scene(blz_cull_ctx* c)
{
    uint32_t* bad = c->counts + 12;                                  // four spare words of the count block
    CU_TRY(cudaMemsetAsync(bad, 0, 4 * sizeof(uint32_t), c->stream));
    for (int i = 0; i < 3; ++i)
        if (c->nObjs[i]) {
            validate_objects_kernel<<<c->numSMs * 8, 256, 0, c->stream>>>(c->objs[i], c->nObjs[i], c->nSurf, c->transformIdBase, c->nXf, bad);
            c->launches++;
        }
    const uint32_t m = c->nSurf > c->nLods ? c->nSurf : c->nLods;
    validate_tables_kernel<<<(m + 255u) / 256u, 256, 0, c->stream>>>(c->surf, c->nSurf, c->lods, c->nLods, c->nClusters, bad);
    c->launches++;
    CU_TRY(cudaGetLastError());
    uint32_t h[4] = { 0, 0, 0, 0 };
    CU_TRY(cudaMemcpyAsync(h, bad, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    if (h[0] | h[1] | h[2] | h[3])
        return fail(BLZ_ERR_INVALID, "malformed scene: %u objects with surfaceId >= %u, %u with a transformId outside [%u, %u), %u surfaces whose LOD range leaves the %u LODs, "
                    "%u LODs whose cluster range leaves the %u clusters", h[0], c->nSurf, h[1], c->transformIdBase, c->transformIdBase + c->nXf, h[2], c->nLods, h[3], c->nClusters);
    return BLZ_OK;
}

} // namespace

int blz_cull_upload_scene(blz_cull_ctx* c, const blz_scene_desc* d)
{
    if (!c || !d) return fail(BLZ_ERR_INVALID, "null argument");
    if (!d->surfaces || d->surface_count == 0 || !d->lods || d->lod_count == 0) return fail(BLZ_ERR_INVALID, "surfaces and lods are required");
    if (!d->transforms || d->transform_count == 0) return fail(BLZ_ERR_INVALID, "transforms are required");
    if (d->lod_instances && d->lod_instance_count && d->lod_instance_count != d->lod_count)
        return fail(BLZ_ERR_INVALID, "lod_instance_count (%u) must equal lod_count (%u)", d->lod_instance_count, d->lod_count);
    if (d->lod_instances && d->lod_instance_count && d->inputs_on_device && !d->instance_bucket_capacity)
        return fail(BLZ_ERR_INVALID, "device inputs need explicit instance_bucket_capacity");
    CU_TRY(cudaSetDevice(c->device));
    const cudaMemcpyKind kind = d->inputs_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    const void* lists[3] = { d->renders, d->transparent_renders, d->onpc_renders };
    const uint32_t counts[3] = { d->render_count, d->transparent_count, d->onpc_count };
    uint64_t maxList = 0;
    for (int i = 0; i < 3; ++i) {
        if (counts[i] && !lists[i]) return fail(BLZ_ERR_INVALID, "render list %d has a count but no pointer", i);
        if (counts[i] > maxList) maxList = counts[i];
    }
    for (int i = 0; i < 3; ++i) {
        c->nObjs[i] = counts[i];
        if (counts[i] == 0) continue;
        TRY_RC(grow(c, c->objs[i], c->capObjs[i], size_t(counts[i]) * sizeof(RenderObject) + 16u));     // + 16: bulk copies of the ragged last tile are rounded up to 16 B
        CU_TRY(cudaMemcpyAsync(c->objs[i], lists[i], size_t(counts[i]) * sizeof(RenderObject), kind, c->stream));
    }
    // transforms: kept AoS exactly as uploaded (32-byte records; cudaMalloc aligns to 256 B, so every record is one aligned 256-bit load)
    c->nXf = d->transform_count;
    TRY_RC(grow(c, c->xf, c->capXf, size_t(c->nXf) * sizeof(MeshTransform)));
    CU_TRY(cudaMemcpyAsync(c->xf, d->transforms, size_t(c->nXf) * sizeof(MeshTransform), kind, c->stream));
    c->nSurf = d->surface_count; c->nLods = d->lod_count;
    TRY_RC(grow(c, c->surf, c->capSurf, size_t(c->nSurf) * sizeof(PrimitiveSurface)));
    CU_TRY(cudaMemcpyAsync(c->surf, d->surfaces, size_t(c->nSurf) * sizeof(PrimitiveSurface), kind, c->stream));
    TRY_RC(grow(c, c->lods, c->capLods, size_t(c->nLods) * sizeof(LodData)));
    CU_TRY(cudaMemcpyAsync(c->lods, d->lods, size_t(c->nLods) * sizeof(LodData), kind, c->stream));
    c->nClusters = 0;
    if (d->clusters && d->cluster_count) {
        c->nClusters = d->cluster_count;
        TRY_RC(grow(c, c->clusters, c->capClusters, size_t(c->nClusters) * sizeof(Cluster)));
        CU_TRY(cudaMemcpyAsync(c->clusters, d->clusters, size_t(c->nClusters) * sizeof(Cluster), kind, c->stream));
    }
    c->objectIdBase = d->object_id_base; c->transformIdBase = d->transform_id_base;
    // visibility buffer, zero-filled (vulkanRendererSetup.cpp:349)
    const size_t nVis = c->nObjs[0] ? c->nObjs[0] : 1;
    TRY_RC(grow(c, c->vis, c->capVis, nVis * sizeof(uint32_t) + 16u));
    CU_TRY(cudaMemsetAsync(c->vis, 0, nVis * sizeof(uint32_t), c->stream));
    const size_t bitBytes = ((nVis + 4095) / 4096) * 512 + 512;                        // whole 4096-object tiles of the bit-mask early pass
    TRY_RC(grow(c, c->visBits, c->capVisBits, bitBytes));
    CU_TRY(cudaMemsetAsync(c->visBits, 0, c->capVisBits, c->stream));
    c->visBitsValid = true; c->visWordsValid = true;
    // draw buffer (sized for the wider DX32 record), cluster dispatch buffer
    c->drawCap = d->draw_capacity ? d->draw_capacity : (maxList ? maxList : 1);
    {
        if (c->gatherStream) CU_TRY(cudaStreamSynchronize(c->gatherStream));
        const size_t before = c->capDraws;
        const size_t need = size_t(c->drawCap) * 8u * sizeof(uint32_t);
        if (c->expDraws.active) {
            // exported outputs stay exportable: a buffer that has to grow is re-created the same way (blz_cull_export_outputs again)
            if (c->expDraws.size < need) {
                CU_TRY(cudaStreamSynchronize(c->stream));
                blz::exportable_free(c->expDraws); c->draws = nullptr; c->capDraws = 0;
                TRY_RC(blz::exportable_alloc(c->device, need, c->expDraws));
                c->draws = static_cast<uint32_t*>(c->expDraws.ptr); c->capDraws = c->expDraws.size;
                c->exportGeneration++;
            }
        } else
        TRY_RC(grow(c, c->draws, c->capDraws, need));
        if (c->capDraws != before && c->drawsAlt) { cudaFree(c->drawsAlt); c->drawsAlt = nullptr; }   // re-created by the next asynchronous push
    }
    c->dispatchCap = d->cluster_dispatch_capacity;
    if (c->dispatchCap) TRY_RC(grow(c, c->dispatch, c->capDispatch, size_t(c->dispatchCap) * 3u * sizeof(uint32_t)));
    // instancing
    c->nLodInst = 0;
    if (d->lod_instances && d->lod_instance_count) {
        c->nLodInst = d->lod_instance_count;
        TRY_RC(grow(c, c->lodInst, c->capLodInst, size_t(c->nLodInst) * sizeof(LodInstanceCounter)));
        CU_TRY(cudaMemcpyAsync(c->lodInst, d->lod_instances, size_t(c->nLodInst) * sizeof(LodInstanceCounter), kind, c->stream));
        std::vector<uint32_t> cap(c->nLodInst);
        std::vector<LodInstanceCounter> li(c->nLodInst);
        if (d->inputs_on_device) {
            CU_TRY(cudaMemcpy(li.data(), d->lod_instances, li.size() * sizeof(LodInstanceCounter), cudaMemcpyDeviceToHost));
            CU_TRY(cudaMemcpy(cap.data(), d->instance_bucket_capacity, cap.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        } else {
            memcpy(li.data(), d->lod_instances, li.size() * sizeof(LodInstanceCounter));
            for (uint32_t l = 0; l < c->nLodInst; ++l) {
                // default: the reference's fixed bucket, instanceOffset = lodId * Ce_MaxInstanceCountPerLOD
                // (Resources/Mesh/blitzenMeshes.cpp:159-161, Core/blitzenEngine.h:65)
                if (d->instance_bucket_capacity) cap[l] = d->instance_bucket_capacity[l];
                else cap[l] = (l + 1 < c->nLodInst && li[l + 1].instanceOffset > li[l].instanceOffset) ? li[l + 1].instanceOffset - li[l].instanceOffset : 100000u;
            }
        }
        uint64_t end = 0;
        for (uint32_t l = 0; l < c->nLodInst; ++l)
            if (uint64_t(li[l].instanceOffset) + cap[l] > end) end = uint64_t(li[l].instanceOffset) + cap[l];
        c->instCap = end ? end : 1;
        TRY_RC(grow(c, c->bucketCap, c->capBucket, size_t(c->nLodInst) * sizeof(uint32_t)));
        CU_TRY(cudaMemcpyAsync(c->bucketCap, cap.data(), cap.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
        TRY_RC(grow(c, c->instIdx, c->capInstIdx, size_t(c->instCap) * sizeof(uint32_t)));
        CU_TRY(cudaStreamSynchronize(c->stream));   // `cap` leaves scope
    }
    // Asynchronous with respect to the host when the sources are pinned; pageable sources are staged by the runtime before
    // cudaMemcpyAsync returns.  Either way the caller may reuse its arrays after blz_cull_synchronize().
    if (c->optValidate) {
        const int rc = validate_scene(c);                            // synchronises: the verdict is part of this call's return value
        c->sceneRefused = rc != 0;                                   // every pass refuses to run on a malformed scene (check_list)
        if (rc) return rc;
    }
    return BLZ_OK;
}

int blz_cull_update_transforms(blz_cull_ctx* c, uint32_t first, uint32_t count, const void* host)
{
    if (!c || (!host && count)) return fail(BLZ_ERR_INVALID, "null argument");
    if (count == 0) return BLZ_OK;
    if (first < c->transformIdBase || uint64_t(first - c->transformIdBase) + count > c->nXf)
        return fail(BLZ_ERR_INVALID, "transform range [%u, %u) outside this context's [%u, %u)", first, first + count, c->transformIdBase, c->transformIdBase + c->nXf);
    CU_TRY(cudaSetDevice(c->device));
    // UpdateObjectTransform / UpdateBuffers (BlitzenVulkan/vulkanDraw.cpp:46-60, :816-820): a stream-ordered copy straight into place
    CU_TRY(cudaMemcpyAsync(c->xf + (first - c->transformIdBase), host, size_t(count) * sizeof(MeshTransform), cudaMemcpyHostToDevice, c->stream));
    return BLZ_OK;
}

int blz_cull_set_view(blz_cull_ctx* c, const void* v)
{
    if (!c || !v) return fail(BLZ_ERR_INVALID, "null argument");
    memcpy(&c->view, v, sizeof(CameraViewData));
    if (c->pyrData) { c->view.pyramidWidth = float(c->pyr.width); c->view.pyramidHeight = float(c->pyr.height); }
    c->haveView = true;
    return BLZ_OK;
}

int blz_cull_reset_visibility(blz_cull_ctx* c)
{
    if (!c || !c->vis) return fail(BLZ_ERR_INVALID, "no scene uploaded");
    CU_TRY(cudaSetDevice(c->device));
    CU_TRY(cudaMemsetAsync(c->vis, 0, size_t(c->nObjs[0] ? c->nObjs[0] : 1) * sizeof(uint32_t), c->stream));
    CU_TRY(cudaMemsetAsync(c->visBits, 0, c->capVisBits, c->stream));
    c->visBitsValid = true; c->visWordsValid = true;
    return BLZ_OK;
}

int blz_cull_write_visibility(blz_cull_ctx* c, const uint32_t* host)
{
    if (!c || !c->vis || !host) return fail(BLZ_ERR_INVALID, "no scene uploaded / null pointer");
    CU_TRY(cudaSetDevice(c->device));
    CU_TRY(cudaMemcpyAsync(c->vis, host, size_t(c->nObjs[0]) * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    c->visBitsValid = false; c->visWordsValid = true;
    return BLZ_OK;
}

int blz_cull_set_depth(blz_cull_ctx* c, const float* host, uint32_t w, uint32_t h)
{
    if (!c || !host || w == 0 || h == 0) return fail(BLZ_ERR_INVALID, "bad depth image");
    CU_TRY(cudaSetDevice(c->device));
    const size_t texels = size_t(w) * h;
    if (texels > c->depthOwnedTexels) {
        CU_TRY(cudaStreamSynchronize(c->stream));
        dfree(c->depthOwned); c->depthOwnedTexels = 0;
        CU_TRY(cudaMalloc(&c->depthOwned, texels * sizeof(float)));
        c->depthOwnedTexels = texels;
    }
    CU_TRY(cudaMemcpyAsync(c->depthOwned, host, texels * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    c->depth = c->depthOwned; c->depthW = w; c->depthH = h;
    return BLZ_OK;
}

int blz_cull_set_depth_device(blz_cull_ctx* c, const float* dev, uint32_t w, uint32_t h)
{
    if (!c || !dev || w == 0 || h == 0) return fail(BLZ_ERR_INVALID, "bad depth image");
    c->depth = dev; c->depthW = w; c->depthH = h;
    return BLZ_OK;
}

int blz_cull_build_pyramid(blz_cull_ctx* c, int variant)
{
    if (!c) return fail(BLZ_ERR_INVALID, "null context");
    if (variant != BLZ_HIZ_VK && variant != BLZ_HIZ_DX) return fail(BLZ_ERR_INVALID, "hiz variant %d", variant);
    if (!c->depth) return fail(BLZ_ERR_INVALID, "no depth image set");
    CU_TRY(cudaSetDevice(c->device));
    int rc = ensure_pyramid_storage(c, variant, c->depthW, c->depthH); if (rc) return rc;
    PyramidBuildParams p{};
    p.depth = c->depth; p.depthW = c->depthW; p.depthH = c->depthH;
    p.out = c->pyrData; p.width = c->pyr.width; p.height = c->pyr.height; p.mips = c->pyr.mips;
    memcpy(p.offset, c->pyr.offset, sizeof(p.offset));
    p.ticket = c->pyrTicket;
    p.variant = variant == BLZ_HIZ_VK ? HIZ_VK : HIZ_DX;
    pyramid_plan(p);
    CUtensorMap map; const void* mapPtr = nullptr;
    if (c->optPyramidTma && (c->depthW % 4u) == 0 && (reinterpret_cast<uintptr_t>(c->depth) % 16u) == 0 && p.boxW <= 256 && p.boxH <= 256) {
        PFN_encodeTiled enc = get_encode_tiled();
        if (enc) {
            cuuint64_t gdim[2] = { c->depthW, c->depthH };
            cuuint64_t gstr[1] = { cuuint64_t(c->depthW) * 4u };
            cuuint32_t box[2] = { p.boxW, p.boxH };
            cuuint32_t estr[2] = { 1, 1 };
            CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(c->depth), gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r == CUDA_SUCCESS) mapPtr = &map;
        }
    }
    CU_TRY(launch_pyramid_build(p, mapPtr, c->stream));
    c->launches++;
    return BLZ_OK;
}

int blz_cull_clear_pyramid(blz_cull_ctx* c, int variant, uint32_t w, uint32_t h)
{
    if (!c) return fail(BLZ_ERR_INVALID, "null context");
    if (variant != BLZ_HIZ_VK && variant != BLZ_HIZ_DX) return fail(BLZ_ERR_INVALID, "hiz variant %d", variant);
    CU_TRY(cudaSetDevice(c->device));
    int rc = ensure_pyramid_storage(c, variant, w, h); if (rc) return rc;
    CU_TRY(cudaMemsetAsync(c->pyrData, 0, c->pyrTexels * sizeof(float), c->stream));
    return BLZ_OK;
}

int blz_cull_frustum_lod(blz_cull_ctx* c, int list, int fmt, uint32_t flags) { return run_draw_pass(c, PASS_FRUSTUM, list, fmt, BLZ_HIZ_VK, flags); }
int blz_cull_early(blz_cull_ctx* c, int fmt) { return run_draw_pass(c, PASS_EARLY, BLZ_LIST_OPAQUE, fmt, BLZ_HIZ_VK, 0); }
int blz_cull_late(blz_cull_ctx* c, int fmt, int hiz) { return run_draw_pass(c, PASS_LATE, BLZ_LIST_OPAQUE, fmt, hiz, 0); }
int blz_cull_temporal(blz_cull_ctx* c, int list, int fmt, int hiz) { return run_draw_pass(c, PASS_TEMPORAL, list, fmt, hiz, 0); }

int blz_cull_instanced(blz_cull_ctx* c, int list)
{
    if (!c) return fail(BLZ_ERR_INVALID, "null context");
    int rc = check_list(c, list); if (rc) return rc;
    if (!c->lodInst) return fail(BLZ_ERR_INVALID, "scene was uploaded without lod_instances");
    if (c->nLods > 256) return fail(BLZ_ERR_CAPACITY, "instancing supports at most 256 LODs (scene has %u)", c->nLods);
    CU_TRY(cudaSetDevice(c->device));
    if (c->nLods >= (1u << 18)) return fail(BLZ_ERR_CAPACITY, "survivor-list descriptors hold 18 bits of LOD id (scene has %u LODs)", c->nLods);
    // step 1: streaming frustum + LOD pass -> survivor list {objectId, absolute LOD id}; step 2: stable counting sort by LOD (cull_list.cu)
    TRY_RC(run_survivor_list(c, list));
    TRY_RC(grow(c, c->listScratch, c->capListScratch, size_t(c->nLods) * kListMaxTiles * sizeof(uint32_t)));
    ListInstanceParams q{};
    q.list = c->survList; q.listCount = c->counts + 6; q.maxEntries = c->nObjs[list];
    q.lodInstances = c->lodInst; q.bucketCapacity = c->bucketCap; q.instanceIndices = c->instIdx;
    q.lods = c->lods; q.lodCount = c->nLods;
    TRY_RC(acquire_draw_buffer(c));
    c->descValid = false;
    q.cmds = c->draws; q.counts = c->drawCounts; q.cmdCapacity = c->drawCap;
    q.hist = c->listScratch;
    CU_TRY(launch_list_instancing(q, c->stream));
    c->launches += 2;
    c->lastRecWords = 8u;
    return BLZ_OK;
}

int blz_cull_cluster_expand(blz_cull_ctx* c, int list)
{
    if (!c) return fail(BLZ_ERR_INVALID, "null context");
    int rc = check_list(c, list); if (rc) return rc;
    if (!c->dispatch) return fail(BLZ_ERR_INVALID, "scene was uploaded with cluster_dispatch_capacity = 0");
    CU_TRY(cudaSetDevice(c->device));
    if (c->nLods >= (1u << 18)) return fail(BLZ_ERR_CAPACITY, "survivor-list descriptors hold 18 bits of LOD id (scene has %u LODs)", c->nLods);
    TRY_RC(run_survivor_list(c, list));
    TRY_RC(grow(c, c->listScratch, c->capListScratch, list_expand_scratch_words(c->nObjs[list], c->dispatchCap) * sizeof(uint32_t)));
    ListExpandParams q{};
    q.list = c->survList; q.listCount = c->counts + 6; q.maxEntries = c->nObjs[list];
    q.lods = c->lods; q.lodCount = c->nLods;
    q.dispatch = c->dispatch; q.counts = c->counts + 2; q.capacity = c->dispatchCap;
    list_expand_carve(c->listScratch, c->nObjs[list], q); q.ctl = c->ctl;
    CU_TRY(launch_list_expand(q, c->numSMs, c->stream));
    c->launches += 2;
    return BLZ_OK;
}

int blz_cull_cluster_cull(blz_cull_ctx* c, int mode, int fmt, int hiz)
{
    if (!c) return fail(BLZ_ERR_INVALID, "null context");
    if (!c->dispatch) return fail(BLZ_ERR_INVALID, "scene was uploaded with cluster_dispatch_capacity = 0");
    if (!c->clusters) return fail(BLZ_ERR_INVALID, "scene has no clusters");
    if (mode < 0 || mode > 2) return fail(BLZ_ERR_INVALID, "cluster mode %d", mode);
    if (fmt != BLZ_REC_VK24 && fmt != BLZ_REC_DX32) return fail(BLZ_ERR_INVALID, "record format %d", fmt);
    if (mode != BLZ_CLUSTER_PASSTHROUGH && !c->haveView) return fail(BLZ_ERR_INVALID, "no view set");
    if (mode == BLZ_CLUSTER_SPHERE_HIZ && (!c->pyrData || c->pyrVariant != hiz)) return fail(BLZ_ERR_INVALID, "no depth pyramid of variant %d", hiz);
    CU_TRY(cudaSetDevice(c->device));
    ClusterCullParams p{};
    p.dispatch = c->dispatch; p.dispatchCount = c->counts + 2;
    p.objs = c->objs[BLZ_LIST_OPAQUE]; p.xf = c->xf; p.clusters = c->clusters;
    TRY_RC(acquire_draw_buffer(c));
    c->descValid = false;
    p.draws = c->draws; p.counts = c->drawCounts; p.ctl = c->ctl;
    p.maxRecords = uint32_t(c->dispatchCap > 0xFFFFFFFFull ? 0xFFFFFFFFull : c->dispatchCap);
    int rc = ensure_status(c, size_t(p.maxRecords) / kCullMinTile + 2u); if (rc) return rc;   // the kernel's tile is 512..1024 records (launch_cluster_cull)
    p.status = c->status;
    p.objectIdBase = c->objectIdBase; p.transformIdBase = c->transformIdBase; p.clusterCount = c->nClusters;
    p.recWords = fmt == BLZ_REC_VK24 ? 6u : 8u; p.mode = uint32_t(mode); p.capacity = c->drawCap;
    p.view = make_view_consts(c->view); p.pyr = c->pyr;
    TRY_RC(scan_epoch_guard(c));
    CU_TRY(launch_cluster_cull(p, hiz == BLZ_HIZ_DX ? HIZ_DX : HIZ_VK, c->numSMs, c->stream));
    c->launches++;
    c->lastRecWords = p.recWords;
    return BLZ_OK;
}

int blz_cull_set_cluster_dispatch(blz_cull_ctx* c, const void* records, uint64_t count, int onDevice)
{
    if (!c || (!records && count)) return fail(BLZ_ERR_INVALID, "null argument");
    if (!c->dispatch || count > c->dispatchCap) return fail(BLZ_ERR_CAPACITY, "dispatch list of %llu records exceeds capacity %llu", (unsigned long long)count, (unsigned long long)c->dispatchCap);
    CU_TRY(cudaSetDevice(c->device));
    if (count) CU_TRY(cudaMemcpyAsync(c->dispatch, records, size_t(count) * 12u, onDevice ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
    uint32_t cnt[2] = { uint32_t(count), uint32_t(count) };
    CU_TRY(cudaMemcpyAsync(c->counts + 2, cnt, sizeof(cnt), cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return BLZ_OK;
}

int blz_cull_get_outputs(blz_cull_ctx* c, blz_outputs* o)
{
    if (!c || !o) return fail(BLZ_ERR_INVALID, "null argument");
    memset(o, 0, sizeof(*o));
    if (c->vis && c->nObjs[0]) TRY_RC(ensure_vis_words(c));                  // the u32 view is materialised on demand (stream-ordered)
    o->draws = c->draws; o->draw_count = c->drawCounts; o->visibility = c->vis;
    o->cluster_dispatch = c->dispatch; o->cluster_count = c->counts ? c->counts + 2 : nullptr;
    o->instance_indices = c->instIdx; o->instance_counts = reinterpret_cast<uint32_t*>(c->lodInst);
    o->pyramid = c->pyrData; o->pyramid_width = c->pyr.width; o->pyramid_height = c->pyr.height; o->pyramid_mips = c->pyr.mips;
    memcpy(o->pyramid_offset, c->pyr.offset, sizeof(o->pyramid_offset));
    o->draw_capacity = c->drawCap; o->cluster_dispatch_capacity = c->dispatchCap;
    return BLZ_OK;
}

int blz_cull_read_count(blz_cull_ctx* c, uint32_t* written, uint32_t* total)
{
    if (!c) return fail(BLZ_ERR_INVALID, "null context");
    CU_TRY(cudaSetDevice(c->device));
    uint32_t h[2];
    CU_TRY(cudaMemcpyAsync(h, c->drawCounts, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    if (written) *written = h[0];
    if (total) *total = h[1];
    return BLZ_OK;
}

static int read_records(blz_cull_ctx* c, const uint32_t* dev, const uint32_t* devCounts, uint32_t recWords, void* host, uint64_t cap, uint32_t* written, uint32_t* total)
{
    CU_TRY(cudaSetDevice(c->device));
    uint32_t h[2];
    CU_TRY(cudaMemcpyAsync(h, devCounts, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    if (written) *written = h[0];
    if (total) *total = h[1];
    if (host) {
        uint64_t n = h[0] < cap ? h[0] : cap;
        if (n) {
            CU_TRY(cudaMemcpyAsync(host, dev, size_t(n) * recWords * 4u, cudaMemcpyDeviceToHost, c->stream));
            CU_TRY(cudaStreamSynchronize(c->stream));
        }
    }
    return BLZ_OK;
}

// `fmt` is the record layout the caller's buffer is made of; it must be the one the last pass wrote (a DX32 list copied into a buffer
// sized for VK24 records would overrun it)
int blz_cull_read_draws(blz_cull_ctx* c, int fmt, void* host, uint64_t cap, uint32_t* written, uint32_t* total)
{
    if (!c || !c->draws) return fail(BLZ_ERR_INVALID, "no scene uploaded");
    if (fmt != BLZ_REC_VK24 && fmt != BLZ_REC_DX32) return fail(BLZ_ERR_INVALID, "unknown record format %d", fmt);
    const uint32_t words = fmt == BLZ_REC_VK24 ? 6u : 8u;
    if (words != c->lastRecWords) return fail(BLZ_ERR_INVALID, "the last pass wrote %u-byte records, the caller asked for %u-byte records", c->lastRecWords * 4u, words * 4u);
    return read_records(c, c->draws, c->drawCounts, c->lastRecWords, host, cap, written, total);
}

int blz_cull_read_visibility(blz_cull_ctx* c, uint32_t* host)
{
    if (!c || !c->vis || !host) return fail(BLZ_ERR_INVALID, "no scene uploaded / null pointer");
    CU_TRY(cudaSetDevice(c->device));
    TRY_RC(ensure_vis_words(c));
    if (c->nObjs[0]) CU_TRY(cudaMemcpyAsync(host, c->vis, size_t(c->nObjs[0]) * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return BLZ_OK;
}

int blz_cull_read_cluster_dispatch(blz_cull_ctx* c, void* host, uint64_t cap, uint32_t* written, uint32_t* total)
{
    if (!c || !c->dispatch) return fail(BLZ_ERR_INVALID, "cluster path not enabled");
    return read_records(c, c->dispatch, c->counts + 2, 3u, host, cap, written, total);
}

int blz_cull_read_instances(blz_cull_ctx* c, uint32_t* idxHost, uint64_t cap, void* countersHost)
{
    if (!c || !c->lodInst) return fail(BLZ_ERR_INVALID, "instancing not enabled");
    CU_TRY(cudaSetDevice(c->device));
    if (idxHost) {
        uint64_t n = c->instCap < cap ? c->instCap : cap;
        if (n) CU_TRY(cudaMemcpyAsync(idxHost, c->instIdx, size_t(n) * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    }
    if (countersHost) CU_TRY(cudaMemcpyAsync(countersHost, c->lodInst, size_t(c->nLodInst) * sizeof(LodInstanceCounter), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return BLZ_OK;
}

int blz_cull_read_pyramid(blz_cull_ctx* c, float* host, uint64_t cap, uint32_t* whm, uint32_t* offsets)
{
    if (!c || !c->pyrData) return fail(BLZ_ERR_INVALID, "no pyramid");
    CU_TRY(cudaSetDevice(c->device));
    size_t texels = 0; PyramidDesc d{};
    d = c->pyr;
    for (uint32_t i = 0; i < d.mips; ++i) texels += size_t((d.width >> i) ? (d.width >> i) : 1u) * ((d.height >> i) ? (d.height >> i) : 1u);
    if (whm) { whm[0] = d.width; whm[1] = d.height; whm[2] = d.mips; }
    if (offsets) memcpy(offsets, d.offset, sizeof(d.offset));
    if (host) {
        if (cap < texels) return fail(BLZ_ERR_CAPACITY, "pyramid has %zu texels, buffer holds %llu", texels, (unsigned long long)cap);
        CU_TRY(cudaMemcpyAsync(host, c->pyrData, texels * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    }
    CU_TRY(cudaStreamSynchronize(c->stream));
    return BLZ_OK;
}

int blz_cull_launch_count(blz_cull_ctx* c, uint64_t* out)
{
    if (!c || !out) return fail(BLZ_ERR_INVALID, "null argument");
    *out = c->launches; return BLZ_OK;
}

int blz_cull_set_option(blz_cull_ctx* c, const char* name, int64_t value)
{
    if (!c || !name) return fail(BLZ_ERR_INVALID, "null argument");
    if (strcmp(name, "pyramid_tma") == 0) { c->optPyramidTma = value; return BLZ_OK; }
    if (strcmp(name, "early_mode") == 0) { c->optEarlyMode = value; return BLZ_OK; }
    if (strcmp(name, "early_bits") == 0) { c->optEarlyBits = value; return BLZ_OK; }
    if (strcmp(name, "vis_words") == 0) { c->optVisWords = value; return BLZ_OK; }
    if (strcmp(name, "l2_fetch_granularity") == 0) {           // 32 / 64 / 128 bytes: device-wide hint (cudaLimitMaxL2FetchGranularity); the sparse early pass gathers 8-16 B per 200-400 B
        CU_TRY(cudaSetDevice(c->device));
        CU_TRY(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, size_t(value)));
        return BLZ_OK;
    }
    if (strcmp(name, "early_auto") == 0) { c->optEarlyAuto = value; c->earlyDense = false; return BLZ_OK; }
    if (strcmp(name, "stream_cfg") == 0) { c->optStreamCfg = value; return BLZ_OK; }
    if (strcmp(name, "gather_desc") == 0) { c->optGatherDesc = value; return BLZ_OK; }
    if (strcmp(name, "gather_timeout_ms") == 0) { if (value < 0) return fail(BLZ_ERR_INVALID, "gather_timeout_ms < 0"); c->optGatherTimeoutMs = value; return BLZ_OK; }
    if (strcmp(name, "validate_scene") == 0) { c->optValidate = value; return BLZ_OK; }
    if (strcmp(name, "epoch_wrap_at") == 0) { if (value < 4) return fail(BLZ_ERR_INVALID, "epoch_wrap_at < 4"); c->epochWrapAt = uint32_t(value); return BLZ_OK; }   // tests: restart the status tag every `value` launches
    return fail(BLZ_ERR_INVALID, "unknown option '%s'", name);
}

} // extern "C"
