/*
 * blz_cull.h -- C ABI of the B200-native replacement for Blitzen's GPU cull dispatch.
 *
 * The reference (PanosKappos2000/Blitzen) has no plugin / FFI boundary: the backend is a compile-time alias
 * (src/Renderer/Interface/blitRenderer.h:16-52) and the cull pass is a handful of file-static functions called from
 * DrawFrame.  This header turns that implicit shape into an explicit C boundary; every entry point names the reference
 * function(s) it replaces (paths relative to /root/reference/src/Renderer unless noted).
 *
 * Conventions
 *   - plain pointers and sizes only; the POD layouts are the reference's own GPU-shared structs
 *     (Resources/renderingResourcesTypes.h, Game/blitCamera.h) and are consumed byte-for-byte.
 *   - every call returns 0 on success and a negative blz_status on failure; blz_cull_last_error() returns a
 *     thread-local, human-readable message (the reference returns uint8_t 1/0 + BLIT_ERROR log,
 *     e.g. BlitzenVulkan/vulkanRendererSetup.cpp:395-399).  Nothing aborts.
 *   - one context per GPU, used from one thread at a time (the reference has a single render thread,
 *     Core/blitzenEntry.cpp:71-102).  All device work is enqueued on the context's stream (blz_cull_set_stream);
 *     no call synchronises the device except the blz_cull_read_* family and blz_cull_synchronize.
 *   - there is NO CPU fallback: without a CUDA device every call fails with BLZ_ERR_CUDA.
 *
 * Output order.  The reference appends with a global atomic (arrival order, nondeterministic).  This library emits
 * ascending objectId (ascending record index for cluster cull, ascending LOD id for instanced commands, ascending objId
 * inside each instance bucket): the same multiset as the reference, byte-identical to oracle/cull_oracle.cpp.
 */
#ifndef BLZ_CULL_H
#define BLZ_CULL_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BLZ_CULL_ABI_VERSION 1

typedef struct blz_cull_ctx blz_cull_ctx;

typedef enum blz_status {
    BLZ_OK = 0,
    BLZ_ERR_INVALID = -1,      /* bad argument / call order */
    BLZ_ERR_CUDA = -2,         /* CUDA runtime error (no device, launch failure, out of memory) */
    BLZ_ERR_CAPACITY = -3,     /* a fixed internal capacity would be exceeded (tables too large, ...) */
    BLZ_ERR_UNSUPPORTED = -4,
    BLZ_ERR_STATE = -5,        /* the call needs something an earlier call sets up (scene, export, fence) */
    BLZ_ERR_TIMEOUT = -6       /* multi-GPU gather: a peer did not arrive within the device-side wait budget (option "gather_timeout_ms") */
} blz_status;

/* which render-object list a pass runs over: the reference keeps three (Resources/RenderObject/blitRender.h:14-21) and
 * selects one per dispatch through the push-constant buffer address (BlitzenVulkan/vulkanData.h:460-465). */
typedef enum blz_list { BLZ_LIST_OPAQUE = 0, BLZ_LIST_TRANSPARENT = 1, BLZ_LIST_ONPC = 2 } blz_list;

/* output record layout */
typedef enum blz_record_format {
    BLZ_REC_VK24 = 0,   /* IndirectDraw, 6 x u32: VulkanShaderHeaders/ShaderBuffers.glsl:73-84 == BlitzenVulkan/vulkanData.h:429-433 */
    BLZ_REC_DX32 = 1    /* DrawCmd, 8 x u32 (last two = padding, written as 0): HlslShaders/Headers/cullBuffers.hlsl:1-15 == BlitzenDX12/dx12Data.h:284-295 */
} blz_record_format;

/* Hi-Z flavour: pyramid construction AND the occlusion test differ between the two backends */
typedef enum blz_hiz_variant {
    BLZ_HIZ_VK = 0,     /* PreviousPow2 extent, MIN-reduction bilinear footprint: BlitzenVulkan/vulkanResources.cpp:51-55,85-116;
                           VulkanShaders/DepthPyramidGeneration.comp.glsl:13-20; VulkanShaderHeaders/CullingShaderData.glsl:60-73 */
    BLZ_HIZ_DX = 1      /* half-res extent, 2x2 min of Loads, single point texel test: BlitzenDX12/dx12RNDResources.cpp:96-118;
                           HlslShaders/CS/depthPyramid.cs.hlsl:13-27; HlslShaders/Headers/hlslMath.hlsl:58-78 */
} blz_hiz_variant;

typedef enum blz_cluster_mode {
    BLZ_CLUSTER_PASSTHROUGH = 0,  /* reference-exact: VulkanShaders/InitialClusterCull.comp.glsl:12-55 (no test is applied) */
    BLZ_CLUSTER_SPHERE = 1,       /* extension asked for by BASELINE config 4: per-cluster frustum test */
    BLZ_CLUSTER_SPHERE_HIZ = 2    /* ... plus the Hi-Z test (variant given separately) */
} blz_cluster_mode;

/* pass flags */
#define BLZ_FLAG_ONPC_LOD_QUIRK 1u  /* VulkanShaders/OnpcDrawCull.comp.glsl:38-40 uses the RELATIVE lod index as absolute (reference bug, reproduced on request) */

/* Scene upload descriptor == what SetupForRendering copies out of DrawContext
 * (Interface/blitRendererInterface.h:16-29; BlitzenVulkan/vulkanRendererSetup.cpp:365-666 StaticBuffersInit).
 * All pointers are HOST pointers unless the *_on_device flag is set; arrays are copied, the caller keeps ownership. */
typedef struct blz_scene_desc {
    const void* renders;            uint32_t render_count;            /* RenderObject[ ]  (8 B)  m_renders */
    const void* transparent_renders; uint32_t transparent_count;      /* RenderObject[ ]         m_transparentRenders */
    const void* onpc_renders;       uint32_t onpc_count;              /* RenderObject[ ]         m_onpcRenders */
    const void* transforms;         uint32_t transform_count;         /* MeshTransform[ ] (32 B) m_transforms[0 .. m_staticTransformOffset) */
    const void* surfaces;           uint32_t surface_count;           /* PrimitiveSurface[ ] (32 B) */
    const void* lods;               uint32_t lod_count;               /* LodData[ ] (32 B) */
    const void* clusters;           uint32_t cluster_count;           /* Cluster[ ] (32 B), may be NULL/0 */
    const void* lod_instances;      uint32_t lod_instance_count;      /* LodInstanceCounter[ ] (8 B), may be NULL/0 (instancing only) */
    /* sharding (multi-GPU): this context holds objects [object_id_base, object_id_base + render_count) of the global opaque
     * list and transforms [transform_id_base, transform_id_base + transform_count) of the global transform array.
     * Records carry GLOBAL object ids; RenderObject.transformId stays global.  0/0 for a single GPU. */
    uint32_t object_id_base;
    uint32_t transform_id_base;
    /* capacities (records).  0 = default: draw_capacity = max list length, cluster_dispatch_capacity = 0 (cluster path off),
     * instance_capacity = sum of bucket capacities.  The reference's fixed sizes are 10'000'000 draws (vulkanData.h:182),
     * 1'000'000 (dx12Data.h:296), 10'000'000 cluster records (vulkanRendererSetup.cpp:550), 100'000 per LOD bucket
     * (Core/blitzenEngine.h:65) and never clamp; this library clamps and reports both written and total. */
    uint64_t draw_capacity;
    uint64_t cluster_dispatch_capacity;
    const uint32_t* instance_bucket_capacity;   /* per LOD, lod_instance_count entries; NULL = distance to the next instanceOffset / 100000 for the last */
    uint32_t inputs_on_device;                  /* 1: every pointer above is a DEVICE pointer on this context's GPU (copied device-to-device) */
    uint32_t reserved;
} blz_scene_desc;

/* Device-resident outputs of the last pass (stay on the GPU, like indirectDrawBuffer / indirectCountBuffer). */
typedef struct blz_outputs {
    void*     draws;             /* device: records, VK24 or DX32 */
    uint32_t* draw_count;        /* device: [0] = written = min(total, capacity) -- the value vkCmdDrawIndexedIndirectCount reads; [1] = total */
    uint32_t* visibility;        /* device: u32 per opaque object (VulkanShaderHeaders/CullingShaderData.glsl:149-152) */
    void*     cluster_dispatch;  /* device: ClusterDispatchData[ ] (12 B) */
    uint32_t* cluster_count;     /* device: [0] = written, [1] = total */
    uint32_t* instance_indices;  /* device: u32[ ] bucketed by LOD (HlslShaders/Headers/sharedBuffers.hlsl:38) */
    uint32_t* instance_counts;   /* device: LodInstanceCounter[ ] {instanceOffset, instanceCount} */
    float*    pyramid;           /* device: linear mip chain, level k at pyramid_offset[k] (texels) */
    uint32_t  pyramid_width, pyramid_height, pyramid_mips;
    uint32_t  pyramid_offset[16];
    uint64_t  draw_capacity, cluster_dispatch_capacity;
} blz_outputs;

/* ---- diagnostics -------------------------------------------------------------------------------------------------- */
int         blz_cull_abi_version(void);
const char* blz_cull_last_error(void);

/* ---- lifecycle: Init / Shutdown of the backend class (BlitzenVulkan/vulkanRenderer.h:20-40) ----------------------- */
int blz_cull_create(int cuda_device, blz_cull_ctx** out_ctx);
int blz_cull_destroy(blz_cull_ctx* ctx);
int blz_cull_set_stream(blz_cull_ctx* ctx, void* cuda_stream /* cudaStream_t; NULL = the context's own non-blocking stream */);
int blz_cull_get_stream(blz_cull_ctx* ctx, void** out_cuda_stream);
int blz_cull_synchronize(blz_cull_ctx* ctx);

/* ---- SetupForRendering (BlitzenVulkan/vulkanRendererSetup.cpp:826-907) --------------------------------------------- */
int blz_cull_upload_scene(blz_cull_ctx* ctx, const blz_scene_desc* desc);
/* UpdateObjectTransform + UpdateBuffers (BlitzenVulkan/vulkanDraw.cpp:816-820, :46-60): global transform ids [first, first+count) */
int blz_cull_update_transforms(blz_cull_ctx* ctx, uint32_t first, uint32_t count, const void* transforms_host);
/* the CameraViewData write at the top of DrawFrame (vulkanDraw.cpp:832-840).  256-byte block, Game/blitCamera.h:38-64.
 * pyramidWidth/pyramidHeight are overwritten with the current pyramid extent, as the backends do
 * (vulkanRendererSetup.cpp:903-904, BlitzenDX12/dx12Draw.cpp:563-564). */
int blz_cull_set_view(blz_cull_ctx* ctx, const void* camera_view_data_256);
/* zero-fill of the visibility buffer (vulkanRendererSetup.cpp:349) */
int blz_cull_reset_visibility(blz_cull_ctx* ctx);
int blz_cull_write_visibility(blz_cull_ctx* ctx, const uint32_t* visibility_host /* render_count entries */);

/* ---- depth + GenerateDepthPyramid (vulkanDraw.cpp:554-622, BlitzenDX12/dx12Draw.cpp:224-288) ----------------------- */
/* depth: W x H fp32, reverse-Z (1 = near, 0 = far/cleared), row-major, tightly packed */
int blz_cull_set_depth(blz_cull_ctx* ctx, const float* depth_host, uint32_t width, uint32_t height);
int blz_cull_set_depth_device(blz_cull_ctx* ctx, const float* depth_device, uint32_t width, uint32_t height);  /* borrowed, not copied */
int blz_cull_build_pyramid(blz_cull_ctx* ctx, int hiz_variant);
/* a cleared pyramid (all 0) of the extent derived from (width, height): frame 0 of the reference */
int blz_cull_clear_pyramid(blz_cull_ctx* ctx, int hiz_variant, uint32_t depth_width, uint32_t depth_height);

/* ---- cull passes ---------------------------------------------------------------------------------------------------- */
/* frustum + LOD + compaction: DrawCullFirstPass with the transparent pipeline (vulkanDraw.cpp:107-158, :1061),
 * VulkanShaders/TransparentDrawCull.comp.glsl, OnpcDrawCull.comp.glsl (BLZ_FLAG_ONPC_LOD_QUIRK), HlslShaders/CS/drawCull.cs.hlsl
 * (BlitzenDX12/dx12Draw.cpp:150-184) */
int blz_cull_frustum_lod(blz_cull_ctx* ctx, int list, int record_format, uint32_t flags);
/* early pass of the two-phase scheme: DrawCullFirstPass with the initial pipeline (vulkanDraw.cpp:1015),
 * VulkanShaders/InitialDrawCull.comp.glsl, HlslShaders/CS/drawOccFirst.cs.hlsl (dx12Draw.cpp:186-222).  Opaque list only. */
int blz_cull_early(blz_cull_ctx* ctx, int record_format);
/* late pass: DrawCullOcclusionPass (vulkanDraw.cpp:162-226, :1031), VulkanShaders/LateDrawCull.comp.glsl,
 * HlslShaders/CS/drawOccLate.cs.hlsl (dx12Draw.cpp:290-338).  Reads + rewrites the visibility buffer. */
int blz_cull_late(blz_cull_ctx* ctx, int record_format, int hiz_variant);
/* frustum + Hi-Z without the visibility buffer: HlslShaders/CS/drawOccTemporal.hlsl:13-65 (compiled but never dispatched
 * by the reference, dx12Draw.cpp:579-582); also what the commented block of TransparentDrawCull.comp.glsl:36-44 would do */
int blz_cull_temporal(blz_cull_ctx* ctx, int list, int record_format, int hiz_variant);
/* D3D12 indirect instancing: DrawInstanceCullPass (dx12Draw.cpp:340-413) = drawInstCountReset + drawInstCull + drawInstCmd.
 * Commands are DX32 records. */
int blz_cull_instanced(blz_cull_ctx* ctx, int list);
/* cluster path: PreClusterDrawCull (vulkanDraw.cpp:318-370) and ClusterCull (vulkanDraw.cpp:372-423).  The cluster count
 * stays on the device: the reference's fence + mapped-buffer read-back between the two (vulkanDraw.cpp:874-905) is gone. */
int blz_cull_cluster_expand(blz_cull_ctx* ctx, int list);
int blz_cull_cluster_cull(blz_cull_ctx* ctx, int cluster_mode, int record_format, int hiz_variant);
/* test/bench hook: install an externally produced dispatch list (device or host pointer) as the input of cluster_cull */
int blz_cull_set_cluster_dispatch(blz_cull_ctx* ctx, const void* records, uint64_t count, int on_device);

/* ---- outputs -------------------------------------------------------------------------------------------------------- */
int blz_cull_get_outputs(blz_cull_ctx* ctx, blz_outputs* out);
/* synchronising read-backs (tests, CPU consumers).  Any pointer may be NULL. */
/* record_format = the layout `records_host` is made of; BLZ_ERR_INVALID when the last pass wrote the other one (indirect instancing writes DX32) */
int blz_cull_read_draws(blz_cull_ctx* ctx, int record_format, void* records_host, uint64_t capacity_records, uint32_t* out_written, uint32_t* out_total);
int blz_cull_read_count(blz_cull_ctx* ctx, uint32_t* out_written, uint32_t* out_total);
int blz_cull_read_visibility(blz_cull_ctx* ctx, uint32_t* visibility_host);
int blz_cull_read_cluster_dispatch(blz_cull_ctx* ctx, void* records_host, uint64_t capacity_records, uint32_t* out_written, uint32_t* out_total);
int blz_cull_read_instances(blz_cull_ctx* ctx, uint32_t* instance_indices_host, uint64_t capacity, void* lod_instance_counters_host /* LodInstanceCounter[lod_count] */);
int blz_cull_read_pyramid(blz_cull_ctx* ctx, float* pyramid_host, uint64_t capacity_texels, uint32_t* out_whm /* [3] */, uint32_t* out_offsets /* [16] */);

/* ---- draw-list consumer (new): replays on the device what the indirect draw + vertex stage read from the outputs
 * (records at stride 24 / 32, vulkanDraw.cpp:470-471; draws[gl_DrawID].objectId -> RenderObject -> surface / LOD table,
 * MainObjectShader.vert.glsl:26; instIndices[objId + SV_InstanceID], opaqueDrawInst.vs.hlsl:11) and reduces them to a checksum.
 * Synchronising.  kind 0 = object draws (each {indexCount, firstIndex} must be a LOD of the object's surface), 1 = cluster draws. */
typedef struct blz_consume_summary {
    uint64_t records, index_sum, instance_sum, id_sum, id_xor;   /* id_xor: xor of objectId * 0x9E3779B97F4A7C15 */
    uint32_t bad_object, bad_lod, unsorted, pad;
    uint32_t lod_hist[256];
} blz_consume_summary;
int blz_cull_consume_draws(blz_cull_ctx* ctx, int list, int kind, blz_consume_summary* out_host);
int blz_cull_consume_instances(blz_cull_ctx* ctx, int list, blz_consume_summary* out_host);
/* presenting rank only: the record-level summary (kind 1) of the list gathered for `epoch`; equals the sum (xor for id_xor) of the ranks' own
 * blz_cull_consume_draws(kind 1) summaries when the gather is correct -- a device-side proof at full size (bench.py prints it as gather_ok) */
int blz_cull_consume_gathered(blz_cull_ctx* ctx, uint32_t epoch, blz_consume_summary* out_host);

/* ---- multi-GPU draw-list gather (new; the reference is single-GPU) -------------------------------------------------
 * Each rank culls its shard; the per-rank lists are concatenated in shard order on the presenting rank.
 * Peer buffers are exchanged as CUDA IPC handles by the host layer (blitzen_b200/dist.py, torch.distributed).
 *   blz_cull_gather_export : fills a 2 x 64-byte blob {IPC handle of this context's gather buffer, of its flag block}
 *   blz_cull_gather_import : maps the presenter's blobs, records (rank, world)
 *   blz_cull_gather_push   : after a cull pass, ONE kernel publishes this rank's count to the presenter's flag block, waits for
 *                            the lower ranks' counts (exclusive scan) and stores this rank's records into the presenter's
 *                            gather buffer at that offset over NVLink peer memory.  No NCCL call on the data path.
 *                            `epoch` = 1, 2, 3, ... (consecutive, the same sequence on every rank).  Ranks need no host-side
 *                            ordering between pushes: counts are kept per epoch in a ring of 4 and a rank stalls (on the
 *                            device) rather than run more than 3 epochs ahead of the slowest one.
 *   Transport (option "gather_desc", default 1): behind an object-list pass (early / late / temporal / frustum) the ranks ship 8-byte
 *   {objectId, absolute LOD id} descriptors instead of the 24-/32-byte records and the presenter expands them with its own (replicated)
 *   LOD table into the same record area, behind its own push -- a third of the bytes into the one GPU that ingests everything; the
 *   gathered list is byte-identical either way and is complete once every rank's done flag carries the epoch (the presenter raises its
 *   own after the expansion: blz_cull_gather_read / blz_cull_consume_gathered / blz_cull_gather_join wait for it).  Cluster draw lists
 *   always travel as records.
 *   Failure detection: every device-side wait of the gather is bounded (option "gather_timeout_ms", default 60 000, 0 = for ever).  A rank
 *   whose peer never arrives moves nothing for that push, still raises its own flags (the timeout does not cascade as a hang) and the next
 *   call that synchronises -- blz_cull_synchronize, blz_cull_gather_read, blz_cull_consume_gathered -- returns BLZ_ERR_TIMEOUT naming the
 *   peer and the epoch; the gather must then be set up again on every rank, the context itself stays usable. */
int blz_cull_gather_export(blz_cull_ctx* ctx, uint64_t capacity_records, int record_format, void* out_blob128);
int blz_cull_gather_import(blz_cull_ctx* ctx, const void* presenter_blob128, int rank, int world);
/* ranks that did not export learn the presenter buffer's capacity / record format from the host layer */
int blz_cull_gather_configure(blz_cull_ctx* ctx, uint64_t capacity_records, int record_format);
int blz_cull_gather_push(blz_cull_ctx* ctx, uint32_t epoch);
/* same push on a side stream: the context flips to its second draw buffer so the next pass overlaps the NVLink transfer; afterwards
 * `draws` / `draw_count` refer to the buffer the next pass will write.  The presenter's gather buffer has two halves (epoch parity):
 * the list of epoch e stays intact while epoch e + 1 arrives. */
int blz_cull_gather_push_async(blz_cull_ctx* ctx, uint32_t epoch);
int blz_cull_gather_join(blz_cull_ctx* ctx);   /* main stream waits for the asynchronous pushes issued so far (stream-ordered) */
int blz_cull_gather_read(blz_cull_ctx* ctx, uint32_t epoch, void* records_host, uint64_t capacity_records, uint32_t* out_counts /* world entries */);
/* instance-list gather (indirect instancing, objects sharded by contiguous ranges; the presenter must be rank 0, whose buckets start the
 * global ones): the host layer all-gathers the per-rank per-LOD instanceCount words (NCCL, same stream) and every other rank stores its
 * buckets behind the lower ranks' into the presenter's instance index buffer over NVLink peer memory. */
int blz_cull_instances_export(blz_cull_ctx* ctx, void* out_blob64);
int blz_cull_instances_import(blz_cull_ctx* ctx, const void* presenter_blob64 /* NULL on the presenter */, int rank, int world);
int blz_cull_instances_counts(blz_cull_ctx* ctx, uint32_t* dst_device /* lod_count words, stream-ordered */);
int blz_cull_instances_push(blz_cull_ctx* ctx, const uint32_t* all_counts_device /* [world][lod_count] */, const uint32_t* global_offset_device, const uint32_t* global_cap_device);
int blz_cull_gather_outputs(blz_cull_ctx* ctx, void** out_records_device, uint32_t** out_flags_device);

/* ---- device-side software depth ----------------------------------------------------------------------------
 * Stands in for DrawGeometry filling the depth attachment between the two cull phases (frame order BlitzenVulkan/vulkanDraw.cpp:1015-1036;
 * attachment cleared to 0, reverse-Z: BlitzenVulkan/vulkanResources.cpp:70-71).  Clears the context's own width x height depth target and
 * merges (max) into it, for every record of the CURRENT draw list of `list`, the screen-space bounding box of the drawn object's bounding
 * sphere at the depth of the sphere's far side, zNear / (c.z + r).  The target becomes the depth image blz_cull_build_pyramid reads.
 * Closed loop of a frame: blz_cull_early -> blz_cull_raster_depth -> blz_cull_build_pyramid -> blz_cull_late. */
int blz_cull_raster_depth(blz_cull_ctx* ctx, int list, uint32_t width, uint32_t height);
/* the current depth image (set_depth / set_depth_device / raster_depth); records_host may be NULL to query the extent */
int blz_cull_read_depth(blz_cull_ctx* ctx, float* depth_host, uint64_t capacity_texels, uint32_t* out_wh /* [2] */);

/* ---- zero-copy hand-over of the outputs --------------------------------------------------------------------
 * Replaces: the renderer reading `indirectDrawBuffer` / `indirectCountBuffer` in vkCmdDrawIndexedIndirectCount
 * (BlitzenVulkan/vulkanDraw.cpp:469-471; buffers created in vulkanRendererSetup.cpp:365-666).  The draw-record buffer and the count
 * block are moved into exportable allocations and handed out as POSIX file descriptors (one per allocation, owned by the caller):
 *   Vulkan : VkImportMemoryFdInfoKHR{handleType = VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT, fd}, allocationSize = *_alloc_bytes,
 *            bind to a VkBuffer with INDIRECT_BUFFER usage; the draw reads records at offset 4, stride 24 (VK24) and the count at
 *            count_offset_bytes (INTEGRATION.md shows the snippet)
 *   CUDA   : blz_interop_import below (cuMemImportFromShareableHandle + map)
 * Call after blz_cull_upload_scene; a later upload that has to GROW the draw buffer re-allocates it (still exportable) and bumps
 * `generation`: export again.  Not available together with blz_cull_gather_push_async (which alternates two draw buffers). */
typedef struct blz_exported_outputs {
    int draws_fd, counts_fd;
    uint64_t draws_alloc_bytes, counts_alloc_bytes;   /* sizes of the two allocations (what the importer must map) */
    uint64_t draw_capacity_records;                   /* records the buffer holds (32 B reserved per record) */
    uint64_t count_offset_bytes;                      /* byte offset of {written, total} inside the count block */
    uint32_t generation, pad;
} blz_exported_outputs;
int blz_cull_export_outputs(blz_cull_ctx* ctx, blz_exported_outputs* out);
/* ordering for CUDA consumers: an interprocess event (64-byte cudaIpcEventHandle_t), recorded on the context's stream by signal_fence */
int blz_cull_export_fence(blz_cull_ctx* ctx, void* out_ipc_event_handle_64);
int blz_cull_signal_fence(blz_cull_ctx* ctx);
/* ordering for the renderer: import the VkSemaphore the renderer exported (vkGetSemaphoreFdKHR; timeline or binary) and signal it on
 * the context's stream behind the cull passes -- the renderer's draw submission waits on it */
int blz_cull_import_semaphore(blz_cull_ctx* ctx, int fd, int is_timeline);
int blz_cull_signal_semaphore(blz_cull_ctx* ctx, uint64_t value);
/* the consumer's side for CUDA consumers (no context needed): map an exported allocation, order behind the fence, read */
int blz_interop_import(int cuda_device, int fd, uint64_t alloc_bytes, void** out_device_ptr);
int blz_interop_release(void* device_ptr, uint64_t alloc_bytes);
int blz_interop_wait_fence(const void* ipc_event_handle_64, void* cuda_stream);
int blz_interop_read(void* host_dst, const void* device_src, uint64_t bytes, void* cuda_stream);

/* ---- instrumentation ------------------------------------------------------------------------------------------------ */
/* number of kernels this library has launched on this context since creation (bench.py's gpu_launches) */
int blz_cull_launch_count(blz_cull_ctx* ctx, uint64_t* out);
/* runtime knobs for A/B measurements: name in {"pyramid_tma", ...}; returns BLZ_ERR_INVALID for unknown names */
int blz_cull_set_option(blz_cull_ctx* ctx, const char* name, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* BLZ_CULL_H */
