// Kernel parameter blocks + launch prototypes shared between the kernels and the C-ABI layer.
#pragma once
#include "cull_types.cuh"

namespace blz {

constexpr int kCullThreads = 256;                    // 8 warps per CTA
constexpr int kCullItems = 4;                        // objects per thread
constexpr int kCullTile = kCullThreads * kCullItems; // objects per tile (one ticket) of the instancing / cluster kernels
constexpr int kCullMinTile = 512;                    // smallest tile any kernel uses: sizes the per-tile status array

struct DrawCullParams {
    // inputs
    const RenderObject* objs;        // AoS, 8 B, the reference layout (two u32)
    const MeshTransform* xf;         // AoS, 32 B, the reference layout as uploaded (one 256-bit load per object); 32-byte aligned
    const PrimitiveSurface* surfaces;
    const LodData* lods;
    uint32_t* visibility;            // u32 per object (early: read, late: read + write)
    // outputs
    uint32_t* draws;                 // records, recWords u32 each
    uint32_t* counts;                // [0] = written (clamped to capacity), [1] = total
    uint32_t* visTotal;              // pipelined early pass: += number of previously-visible objects it walked (may be null) ...
    uint32_t* visTotalOut;           // ... and the last CTA out stores the total here (host-mapped pinned word) and zeroes the accumulator
    uint2* descs;                    // optional (may be null): {objectId, absolute LOD id} of every record written, same order and clamp -- what the multi-GPU gather ships instead of the 24/32-byte records
    uint32_t* visBits;               // 1 bit per object (word i = objects 32i..32i+31): written by the streaming late pass, source of the pipelined early pass (may be null)
    // scan state
    ScanCtl* ctl;
    uint64_t* status;
    // sizes
    uint32_t n;                      // objects in the list
    uint32_t numTiles;               // ceil(n / kCullTile), at least 1
    uint32_t objectIdBase, transformIdBase;
    uint32_t surfaceCount, lodCount;
    uint32_t recWords;               // 6 (VK24) or 8 (DX32)
    uint32_t flags;
    uint64_t capacity;               // records
    ViewConsts view;
    PyramidDesc pyr;
};

struct ClusterCullParams {
    const uint32_t* dispatch;             // ClusterDispatchData records
    const uint32_t* dispatchCount;        // device: [0] = number of records (stays on the device)
    const RenderObject* objs; const MeshTransform* xf;
    const Cluster* clusters;
    uint32_t* draws; uint32_t* counts;
    ScanCtl* ctl; uint64_t* status;
    uint32_t maxRecords;                  // capacity of the dispatch buffer (upper bound of *dispatchCount)
    uint32_t objectIdBase, transformIdBase, clusterCount;
    uint32_t recWords, mode;
    uint64_t capacity;
    ViewConsts view;
    PyramidDesc pyr;
};

// survivor-list pipeline (cull_list.cu): step 1 = streaming frustum + LOD pass with recWords = 2 -> {objectId, absolute LOD id}
constexpr uint32_t kListMaxTiles = 1024;   // tiles of the instancing step (bounds the per-tile histogram reduction)
struct ListInstanceParams {
    const uint2* list; const uint32_t* listCount;   // survivors, ascending objectId; count on the device
    uint32_t maxEntries;                             // capacity of `list`
    LodInstanceCounter* lodInstances; const uint32_t* bucketCapacity; uint32_t* instanceIndices;
    const LodData* lods; uint32_t lodCount;
    uint32_t* cmds; uint32_t* counts; uint64_t cmdCapacity;
    uint32_t* hist;                                  // [lodCount][kListMaxTiles]
};
struct ListExpandParams {
    const uint2* list; const uint32_t* listCount; uint32_t maxEntries;
    const LodData* lods; uint32_t lodCount;
    uint32_t* dispatch; uint32_t* counts; uint64_t capacity;
    uint32_t* tileCount;                             // per tile of 2048 survivors: records
    uint32_t* tilePrefix;                            // ... and their exclusive prefix
    uint2* items;                                    // work items of the write step: {tile, slice of the tile's records}
    uint32_t* numItems;
    ScanCtl* ctl;
};
size_t list_expand_scratch_words(uint32_t maxEntries, uint64_t capacity);
void list_expand_carve(uint32_t* scratch, uint32_t maxEntries, ListExpandParams& p);
cudaError_t launch_list_instancing(const ListInstanceParams& p, cudaStream_t stream);
cudaError_t launch_list_expand(const ListExpandParams& p, int numSMs, cudaStream_t stream);

// launchers (return the cudaError_t of the launch)
cudaError_t launch_stream_cull(const DrawCullParams& p, int pass, int hiz, int cfg, int numSMs, cudaStream_t stream);   // cull_stream.cu
cudaError_t launch_early_stream(const DrawCullParams& p, int numSMs, cudaStream_t stream);   // cull_early.cu: pipelined early pass (default)
cudaError_t launch_pack_vis_bits(const uint32_t* vis, uint32_t* bits, uint32_t n, cudaStream_t stream);   // cull_early.cu
cudaError_t launch_unpack_vis_bits(const uint32_t* bits, uint32_t* vis, uint32_t n, cudaStream_t stream); // cull_early.cu
cudaError_t launch_popc_vis_bits(const uint32_t* bits, uint32_t n, uint32_t* total, cudaStream_t stream);   // cull_early.cu
cudaError_t launch_cluster_cull(const ClusterCullParams& p, int hiz, int numSMs, cudaStream_t stream);

struct PyramidBuildParams {
    const float* depth; uint32_t depthW, depthH;
    float* out;                        // mip chain
    uint32_t width, height, mips;      // level-0 extent
    uint32_t offset[16];
    uint32_t* ticket;                  // device counter, self-resetting
    uint32_t tilesX, tilesY;
    uint32_t tileLevels;               // levels produced inside the tile stage (<= 6)
    uint32_t boxW, boxH;               // input box staged per tile (texels)
    int variant;
    uint32_t smemFloats;               // dynamic shared memory of a CTA, in floats (set by the launcher: the tail reuses it)
};
cudaError_t launch_pyramid_build(const PyramidBuildParams& p, const void* tensorMap /* CUtensorMap* on host or null */, cudaStream_t stream);
size_t pyramid_smem_bytes(const PyramidBuildParams& p);

} // namespace blz
