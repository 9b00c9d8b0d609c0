// Zero-copy hand-over of the cull outputs to the consumer of the reference's draw.
// The reference's vkCmdDrawIndexedIndirectCount reads `indirectDrawBuffer` at offset 4 / stride 24 and `indirectCountBuffer` at
// offset 0 (BlitzenVulkan/vulkanDraw.cpp:469-471); both are VkBuffers created in SetupForRendering
// (BlitzenVulkan/vulkanRendererSetup.cpp:365-666).  Here the two outputs are moved into exportable allocations (CUDA virtual memory
// management: cuMemCreate with a POSIX file-descriptor handle type) and handed out as file descriptors -- what
// VkImportMemoryFdInfoKHR (VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT) takes, and what a second CUDA process imports with
// cuMemImportFromShareableHandle (blz_interop_import below; tests/test_interop_gpu.py does exactly that and reads the lists in place).
// Ordering: an interprocess CUDA event for CUDA consumers, an imported Vulkan (timeline) semaphore for the renderer.
// The driver API is reached through cudaGetDriverEntryPoint, so the library still loads on a box without libcuda.
#include "ctx.h"
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstring>
#include <unistd.h>

namespace blz {

namespace {

struct Drv {
    CUresult (*getGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*reserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*addressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*setAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*exportHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*importHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
    bool ok = false;
};

const Drv& drv()
{
    static Drv d;
    static bool tried = false;
    if (!tried) {
        tried = true;
        auto get = [](const char* name, void** fn) {
            cudaDriverEntryPointQueryResult q;
            return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn != nullptr;
        };
        d.ok = get("cuMemGetAllocationGranularity", reinterpret_cast<void**>(&d.getGranularity))
            && get("cuMemCreate", reinterpret_cast<void**>(&d.create)) && get("cuMemRelease", reinterpret_cast<void**>(&d.release))
            && get("cuMemAddressReserve", reinterpret_cast<void**>(&d.reserve)) && get("cuMemAddressFree", reinterpret_cast<void**>(&d.addressFree))
            && get("cuMemMap", reinterpret_cast<void**>(&d.map)) && get("cuMemUnmap", reinterpret_cast<void**>(&d.unmap))
            && get("cuMemSetAccess", reinterpret_cast<void**>(&d.setAccess))
            && get("cuMemExportToShareableHandle", reinterpret_cast<void**>(&d.exportHandle))
            && get("cuMemImportFromShareableHandle", reinterpret_cast<void**>(&d.importHandle));
    }
    return d;
}

#define DRV_TRY(expr)                                                                                                   \
    do {                                                                                                                \
        CUresult r__ = (expr);                                                                                          \
        if (r__ != CUDA_SUCCESS) return blz::fail(BLZ_ERR_CUDA, "%s failed: CUresult %d (%s:%d)", #expr, int(r__), __FILE__, __LINE__); \
    } while (0)

CUmemAllocationProp alloc_prop(int device)
{
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof(prop));
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return prop;
}

int map_handle(int device, CUmemGenericAllocationHandle h, size_t size, void** outPtr)
{
    const Drv& d = drv();
    CUdeviceptr ptr = 0;
    DRV_TRY(d.reserve(&ptr, size, 0, 0, 0));
    CUresult r = d.map(ptr, size, 0, h, 0);
    if (r != CUDA_SUCCESS) { d.addressFree(ptr, size); return fail(BLZ_ERR_CUDA, "cuMemMap failed: CUresult %d", int(r)); }
    CUmemAccessDesc acc;
    memset(&acc, 0, sizeof(acc));
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    r = d.setAccess(ptr, size, &acc, 1);
    if (r != CUDA_SUCCESS) { d.unmap(ptr, size); d.addressFree(ptr, size); return fail(BLZ_ERR_CUDA, "cuMemSetAccess failed: CUresult %d", int(r)); }
    *outPtr = reinterpret_cast<void*>(ptr);
    return BLZ_OK;
}

} // namespace

int exportable_alloc(int device, size_t bytes, ExportableBuffer& b)
{
    const Drv& d = drv();
    if (!d.ok) return fail(BLZ_ERR_CUDA, "the CUDA driver does not expose the virtual memory management API");
    CUmemAllocationProp prop = alloc_prop(device);
    size_t gran = 0;
    DRV_TRY(d.getGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
    const size_t size = ((bytes ? bytes : 1) + gran - 1) / gran * gran;
    CUmemGenericAllocationHandle h = 0;
    DRV_TRY(d.create(&h, size, &prop, 0));
    void* ptr = nullptr;
    int rc = map_handle(device, h, size, &ptr);
    if (rc) { d.release(h); return rc; }
    b.ptr = ptr; b.size = size; b.handle = uint64_t(h); b.active = true;
    return BLZ_OK;
}

void exportable_free(ExportableBuffer& b)
{
    if (!b.active) return;
    const Drv& d = drv();
    d.unmap(CUdeviceptr(reinterpret_cast<uintptr_t>(b.ptr)), b.size);
    d.addressFree(CUdeviceptr(reinterpret_cast<uintptr_t>(b.ptr)), b.size);
    d.release(CUmemGenericAllocationHandle(b.handle));
    b = ExportableBuffer{};
}

int exportable_fd(const ExportableBuffer& b, int* outFd)
{
    int fd = -1;
    DRV_TRY(drv().exportHandle(&fd, CUmemGenericAllocationHandle(b.handle), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    *outFd = fd;
    return BLZ_OK;
}

void interop_release(blz_cull_ctx* c)
{
    if (c->exportFence) { cudaEventDestroy(c->exportFence); c->exportFence = nullptr; }
    if (c->extSemaphore) { cudaDestroyExternalSemaphore(static_cast<cudaExternalSemaphore_t>(c->extSemaphore)); c->extSemaphore = nullptr; }
}

} // namespace blz

using namespace blz;

extern "C" {

int blz_cull_export_outputs(blz_cull_ctx* c, blz_exported_outputs* out)
{
    if (!c || !out) return fail(BLZ_ERR_INVALID, "null argument");
    if (!c->draws || !c->counts) return fail(BLZ_ERR_STATE, "no scene uploaded: the draw buffer does not exist yet");
    if (c->drawsAlt) return fail(BLZ_ERR_STATE, "outputs cannot be exported while asynchronous gather pushes alternate the draw buffer");
    CU_TRY(cudaSetDevice(c->device));
    CU_TRY(cudaStreamSynchronize(c->stream));
    if (!c->expDraws.active) {
        // the record buffer's contents are per-pass: nothing to carry over
        ExportableBuffer nb;
        int rc = exportable_alloc(c->device, c->capDraws, nb);
        if (rc) return rc;
        cudaFree(c->draws);
        c->draws = static_cast<uint32_t*>(nb.ptr);
        c->expDraws = nb;
        c->exportGeneration++;
    }
    if (!c->expCounts.active) {
        // the count block also holds the context's small device-side accumulators: copied over
        ExportableBuffer nb;
        int rc = exportable_alloc(c->device, 16 * sizeof(uint32_t), nb);
        if (rc) return rc;
        CU_TRY(cudaMemcpy(nb.ptr, c->counts, 16 * sizeof(uint32_t), cudaMemcpyDeviceToDevice));
        const ptrdiff_t slot = c->drawCounts - c->counts;
        cudaFree(c->counts);
        c->counts = static_cast<uint32_t*>(nb.ptr);
        c->drawCounts = c->counts + slot;
        c->expCounts = nb;
        c->exportGeneration++;
    }
    memset(out, 0, sizeof(*out));
    int rc = exportable_fd(c->expDraws, &out->draws_fd);
    if (rc) return rc;
    rc = exportable_fd(c->expCounts, &out->counts_fd);
    if (rc) { close(out->draws_fd); out->draws_fd = -1; return rc; }
    out->draws_alloc_bytes = c->expDraws.size;
    out->counts_alloc_bytes = c->expCounts.size;
    out->draw_capacity_records = c->drawCap;
    out->count_offset_bytes = uint64_t(c->drawCounts - c->counts) * sizeof(uint32_t);
    out->generation = c->exportGeneration;
    return BLZ_OK;
}

int blz_cull_export_fence(blz_cull_ctx* c, void* out_ipc_event_handle_64)
{
    if (!c || !out_ipc_event_handle_64) return fail(BLZ_ERR_INVALID, "null argument");
    CU_TRY(cudaSetDevice(c->device));
    if (!c->exportFence) CU_TRY(cudaEventCreateWithFlags(&c->exportFence, cudaEventDisableTiming | cudaEventInterprocess));
    cudaIpcEventHandle_t h;
    CU_TRY(cudaIpcGetEventHandle(&h, c->exportFence));
    static_assert(sizeof(h) == 64, "IPC event handle size");
    memcpy(out_ipc_event_handle_64, &h, sizeof(h));
    return BLZ_OK;
}

int blz_cull_signal_fence(blz_cull_ctx* c)
{
    if (!c) return fail(BLZ_ERR_INVALID, "null context");
    if (!c->exportFence) return fail(BLZ_ERR_STATE, "blz_cull_export_fence has not been called");
    CU_TRY(cudaEventRecord(c->exportFence, c->stream));
    return BLZ_OK;
}

int blz_cull_import_semaphore(blz_cull_ctx* c, int fd, int is_timeline)
{
    if (!c) return fail(BLZ_ERR_INVALID, "null context");
    if (fd < 0) return fail(BLZ_ERR_INVALID, "invalid semaphore file descriptor %d", fd);
    CU_TRY(cudaSetDevice(c->device));
    if (c->extSemaphore) { cudaDestroyExternalSemaphore(static_cast<cudaExternalSemaphore_t>(c->extSemaphore)); c->extSemaphore = nullptr; }
    cudaExternalSemaphoreHandleDesc d;
    memset(&d, 0, sizeof(d));
    d.type = is_timeline ? cudaExternalSemaphoreHandleTypeTimelineSemaphoreFd : cudaExternalSemaphoreHandleTypeOpaqueFd;
    d.handle.fd = fd;
    cudaExternalSemaphore_t s = nullptr;
    CU_TRY(cudaImportExternalSemaphore(&s, &d));          // on success the driver owns the descriptor
    c->extSemaphore = s;
    c->extSemaphoreTimeline = is_timeline != 0;
    return BLZ_OK;
}

int blz_cull_signal_semaphore(blz_cull_ctx* c, uint64_t value)
{
    if (!c) return fail(BLZ_ERR_INVALID, "null context");
    if (!c->extSemaphore) return fail(BLZ_ERR_STATE, "no semaphore imported (blz_cull_import_semaphore)");
    cudaExternalSemaphoreSignalParams p;
    memset(&p, 0, sizeof(p));
    p.params.fence.value = c->extSemaphoreTimeline ? value : 0;
    cudaExternalSemaphore_t s = static_cast<cudaExternalSemaphore_t>(c->extSemaphore);
    CU_TRY(cudaSignalExternalSemaphoresAsync(&s, &p, 1, c->stream));
    return BLZ_OK;
}

// ---- the consumer's side, for CUDA consumers and for the tests: map an exported allocation / wait for the fence --------------------
int blz_interop_import(int cuda_device, int fd, uint64_t alloc_bytes, void** out_device_ptr)
{
    if (!out_device_ptr || fd < 0 || alloc_bytes == 0) return fail(BLZ_ERR_INVALID, "bad argument");
    const Drv& d = drv();
    CU_TRY(cudaSetDevice(cuda_device));
    CU_TRY(cudaFree(nullptr));
    if (!d.ok) return fail(BLZ_ERR_CUDA, "the CUDA driver does not expose the virtual memory management API");
    CUmemGenericAllocationHandle h = 0;
    DRV_TRY(d.importHandle(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    void* ptr = nullptr;
    int rc = map_handle(cuda_device, h, size_t(alloc_bytes), &ptr);
    d.release(h);                                          // the mapping keeps the allocation alive
    if (rc) return rc;
    *out_device_ptr = ptr;
    return BLZ_OK;
}

int blz_interop_release(void* device_ptr, uint64_t alloc_bytes)
{
    if (!device_ptr) return BLZ_OK;
    const Drv& d = drv();
    DRV_TRY(d.unmap(CUdeviceptr(reinterpret_cast<uintptr_t>(device_ptr)), size_t(alloc_bytes)));
    DRV_TRY(d.addressFree(CUdeviceptr(reinterpret_cast<uintptr_t>(device_ptr)), size_t(alloc_bytes)));
    return BLZ_OK;
}

int blz_interop_wait_fence(const void* ipc_event_handle_64, void* cuda_stream)
{
    if (!ipc_event_handle_64) return fail(BLZ_ERR_INVALID, "null handle");
    cudaIpcEventHandle_t h;
    memcpy(&h, ipc_event_handle_64, sizeof(h));
    cudaEvent_t ev = nullptr;
    CU_TRY(cudaIpcOpenEventHandle(&ev, h));
    cudaError_t e = cudaStreamWaitEvent(static_cast<cudaStream_t>(cuda_stream), ev, 0);
    cudaEventDestroy(ev);
    if (e != cudaSuccess) return fail(BLZ_ERR_CUDA, "cudaStreamWaitEvent failed: %s", cudaGetErrorString(e));
    return BLZ_OK;
}

int blz_interop_read(void* host_dst, const void* device_src, uint64_t bytes, void* cuda_stream)
{
    CU_TRY(cudaMemcpyAsync(host_dst, device_src, size_t(bytes), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(cuda_stream)));
    CU_TRY(cudaStreamSynchronize(static_cast<cudaStream_t>(cuda_stream)));
    return BLZ_OK;
}

} // extern "C"
