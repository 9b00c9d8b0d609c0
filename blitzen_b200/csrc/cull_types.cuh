// POD layouts shared between the C-ABI layer and the kernels.  These are the reference's GPU-shared structs
// (/root/reference/src/Renderer/Resources/renderingResourcesTypes.h, Game/blitCamera.h); they are consumed byte for byte.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace blz {

struct RenderObject { uint32_t transformId, surfaceId; };                                   // renderingResourcesTypes.h:155-159
struct MeshTransform { float pos[3]; float scale; float q[4]; };                            // :124-129
struct PrimitiveSurface { float center[3]; float radius; uint32_t materialId, lodOffset, lodCount, vertexOffset; }; // :104-116
struct LodData { uint32_t indexCount, firstIndex, clusterOffset, clusterCount; float error; uint32_t pad[3]; };     // :73-96
struct Cluster { float center[3]; float radius; int8_t coneAxis[3]; int8_t coneCutoff; uint32_t dataOffset;
                 uint8_t vertexCount, triangleCount, pad0, pad1; uint32_t tail; };                                    // :27-47
struct LodInstanceCounter { uint32_t instanceOffset, instanceCount; };                       // :98-102
struct ClusterDispatchData { uint32_t objectId, lodIndex, clusterId; };                      // VulkanShaderHeaders/CullingShaderData.glsl:118-123
struct CameraViewData {                                                                      // Game/blitCamera.h:38-64 (alignas 256)
    float view[16]; float projView[16]; float position[3];
    float frustumRight, frustumLeft, frustumTop, frustumBottom;
    float proj0, proj5, zNear, zFar, pyramidWidth, pyramidHeight, lodTarget;
    uint8_t pad[72];
};
static_assert(sizeof(RenderObject) == 8 && sizeof(MeshTransform) == 32 && sizeof(PrimitiveSurface) == 32, "layout");
static_assert(sizeof(LodData) == 32 && sizeof(Cluster) == 32 && sizeof(LodInstanceCounter) == 8, "layout");
static_assert(sizeof(ClusterDispatchData) == 12 && sizeof(CameraViewData) == 256, "layout");

// The part of CameraViewData the cull shaders read, passed BY VALUE in kernel parameters so that every constant is a
// constant-bank operand of the FP instructions instead of a per-thread load.
struct ViewConsts {
    float m[12];          // view matrix rows needed for .xyz: m[0..2]=col0.xyz, m[3..5]=col1.xyz, m[6..8]=col2.xyz, m[9..11]=col3.xyz
    float frustumRight, frustumLeft, frustumTop, frustumBottom;
    float proj0, proj5, zNear, zFar, pyramidWidth, pyramidHeight, lodTarget;
};

// Linear mip chain of the Hi-Z pyramid (R32F): level k is (max(1,width>>k) x max(1,height>>k)) texels at data + offset[k].
struct PyramidDesc {
    const float* data;
    uint32_t width, height, mips;
    uint32_t offset[16];
};

// Device control block of the single-pass (decoupled look-back) compaction; lives in the context, shared by all passes
// (passes are stream-ordered).  The last CTA of every launch resets ticket/done and bumps epoch, so no memset launch is needed.
struct ScanCtl {
    uint32_t ticket;   // dynamic tile id dispenser (tile order == object order; a tile's predecessors have all started)
    uint32_t done;     // CTAs that have left the kernel
    uint32_t epoch;    // validity tag of the tile status words written by the current launch
    uint32_t pad;
};

enum Pass { PASS_FRUSTUM = 0, PASS_EARLY = 1, PASS_LATE = 2, PASS_TEMPORAL = 3 };
enum Hiz { HIZ_VK = 0, HIZ_DX = 1, HIZ_NONE = 2 };

constexpr uint32_t kFlagOnpcLodQuirk = 1u;

} // namespace blz
