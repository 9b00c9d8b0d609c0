// BlitzenCuda::CudaCullRenderer -- the C++ host side of the B200 cull path, above the C ABI (include/blz_cull.h).
//
// It mirrors the part of the reference's backend class that the cull dispatch lives in: same member names, argument meaning
// and error behaviour (uint8_t 1 = ok / 0 = failed + an error line in the log, like BlitzenVulkan::VulkanRenderer, whose
// functions return uint8_t and BLIT_ERROR on failure, e.g. BlitzenVulkan/vulkanRendererSetup.cpp:395-399):
//
//   reference (paths relative to /root/reference/src/Renderer)                              this class
//   ------------------------------------------------------------------------------------   -----------------------------------
//   VulkanRenderer::Init                         BlitzenVulkan/vulkanRenderer.h:20          Init(cudaDevice)
//   VulkanRenderer::SetupForRendering            vulkanRendererSetup.cpp:826-907            SetupForRendering(scene)
//   VulkanRenderer::UpdateObjectTransform        vulkanDraw.cpp:816-820                     UpdateObjectTransform(id, pTransform)
//   VulkanRenderer::Update + viewData write      vulkanDraw.cpp:832-840                     Update(pCameraViewData)
//   DrawCullFirstPass(initial | transparent)     vulkanDraw.cpp:107-158, :1015, :1061       DrawCullFirstPass(pipeline)
//   GenerateDepthPyramid                         vulkanDraw.cpp:554-622                     GenerateDepthPyramid()
//   DrawCullOcclusionPass                        vulkanDraw.cpp:162-226, :1031              DrawCullOcclusionPass()
//   PreClusterDrawCull / ClusterCull             vulkanDraw.cpp:318-423                     PreClusterDrawCull() / ClusterCull(mode)
//   Dx12Renderer::DrawInstanceCullPass           BlitzenDX12/dx12Draw.cpp:340-413           DrawInstanceCullPass()
//   the cull part of DrawFrame                   vulkanDraw.cpp:1015-1069                   DrawFrameCull(stats)
//
// This header has no dependency on the reference tree (the GPU box does not have it): the scene arrives as the plain
// pointer + count view `CullScene`.  blitzenCudaCullAdapter.h builds that view from the reference's own
// BlitzenEngine::DrawContext (BlitCL containers, Blitzen math types) where the reference headers are available.
#pragma once
#include <cstdint>
#include <cstddef>

struct blz_cull_ctx;

namespace BlitzenCuda
{
    // What SetupForRendering copies out of DrawContext (Interface/blitRendererInterface.h:16-29); element layouts are the
    // reference's GPU-shared PODs (Resources/renderingResourcesTypes.h): RenderObject 8 B, MeshTransform 32 B,
    // PrimitiveSurface 32 B, LodData 32 B, Cluster 32 B, LodInstanceCounter 8 B.
    struct CullScene
    {
        const void* pRenders{ nullptr };            uint32_t renderCount{ 0 };              // RenderContainer::m_renders
        const void* pTransparentRenders{ nullptr }; uint32_t transparentRenderCount{ 0 };   // m_transparentRenders
        const void* pOnpcRenders{ nullptr };        uint32_t onpcRenderCount{ 0 };          // m_onpcRenders
        const void* pTransforms{ nullptr };         uint32_t transformCount{ 0 };           // m_transforms[0 .. m_transformCount)
        const void* pSurfaces{ nullptr };           uint32_t surfaceCount{ 0 };             // MeshResources::m_surfaces
        const void* pLods{ nullptr };               uint32_t lodCount{ 0 };                 // m_LODs
        const void* pClusters{ nullptr };           uint32_t clusterCount{ 0 };             // m_clusters
        const void* pLodInstances{ nullptr };       uint32_t lodInstanceCount{ 0 };         // m_lodInstanceList
        // capacities in records; 0 = the list length (the reference's fixed sizes: 10'000'000 draws, vulkanData.h:182;
        // 10'000'000 cluster dispatch records, vulkanRendererSetup.cpp:550)
        uint64_t drawCapacity{ 0 };
        uint64_t clusterDispatchCapacity{ 0 };
        const uint32_t* pInstanceBucketCapacity{ nullptr };
    };

    enum class CullPipeline : uint8_t { Initial = 0, Transparent = 1, Onpc = 2 };           // m_initialDrawCullPipeline / m_transparentDrawCullPipeline / m_onpcDrawCullPipeline
    enum class HiZVariant : uint8_t { Vulkan = 0, D3D12 = 1 };
    enum class RecordFormat : uint8_t { IndirectDrawVK24 = 0, DrawCmdDX32 = 1 };
    enum class ClusterMode : uint8_t { Passthrough = 0, Sphere = 1, SphereHiZ = 2 };

    struct CullStats
    {
        uint32_t earlyDrawCount{ 0 }, lateDrawCount{ 0 }, transparentDrawCount{ 0 };          // what vkCmdDrawIndexedIndirectCount would read
        uint32_t earlyTotal{ 0 }, lateTotal{ 0 }, transparentTotal{ 0 };                     // before the capacity clamp
    };

    // log hook: level 0 = info, 1 = error.  Defaults to stderr; the adapter routes it to BLIT_INFO / BLIT_ERROR.
    using LogFn = void (*)(int level, const char* message);
    void SetLogCallback(LogFn fn);

    class CudaCullRenderer
    {
    public:
        CudaCullRenderer() = default;
        ~CudaCullRenderer();
        CudaCullRenderer(const CudaCullRenderer&) = delete;
        CudaCullRenderer& operator=(const CudaCullRenderer&) = delete;

        uint8_t Init(int cudaDevice, HiZVariant hiz = HiZVariant::Vulkan, RecordFormat format = RecordFormat::IndirectDrawVK24);
        void Shutdown();

        uint8_t SetupForRendering(const CullScene& scene);
        void UpdateObjectTransform(uint32_t transformId, const void* pTransform /* MeshTransform, 32 B */);
        uint8_t UpdateObjectTransforms(uint32_t firstTransformId, uint32_t count, const void* pTransforms);
        void Update(const void* pCameraViewData /* CameraViewData, 256 B */);

        // the depth target of the previous draw pass (fp32, reverse-Z); device pointers are borrowed, host images are copied
        uint8_t SetDepthAttachment(const float* pDepth, uint32_t width, uint32_t height, bool onDevice);
        uint8_t ClearDepthPyramid(uint32_t depthWidth, uint32_t depthHeight);      // frame 0: depth cleared to 0 (vulkanResources.cpp:70-71)

        uint8_t DrawCullFirstPass(CullPipeline pipeline);
        uint8_t GenerateDepthPyramid();
        uint8_t DrawCullOcclusionPass();
        uint8_t PreClusterDrawCull();
        uint8_t ClusterCull(ClusterMode mode = ClusterMode::Passthrough);
        uint8_t DrawInstanceCullPass();

        // early pass -> pyramid -> late pass [-> transparent pass], reading back the draw counts between passes (the reference
        // consumes each list with vkCmdDrawIndexedIndirectCount before the next pass overwrites the shared buffer)
        uint8_t DrawFrameCull(CullStats* pStats);

        // device-resident outputs (indirectDrawBuffer / indirectCountBuffer of the reference) and synchronising read-backs
        void* IndirectDrawBuffer() const;
        uint32_t* IndirectCountBuffer() const;
        uint8_t ReadDrawCount(uint32_t* pWritten, uint32_t* pTotal);
        uint8_t ReadDraws(void* pRecords, uint64_t capacityRecords, uint32_t* pWritten, uint32_t* pTotal);
        uint8_t ReadVisibility(uint32_t* pVisibility);
        uint8_t WaitIdle();

        blz_cull_ctx* Handle() const { return m_ctx; }
        bool bTransparentObjectsExist() const { return m_transparentCount != 0; }

    private:
        blz_cull_ctx* m_ctx{ nullptr };
        HiZVariant m_hiz{ HiZVariant::Vulkan };
        RecordFormat m_format{ RecordFormat::IndirectDrawVK24 };
        uint32_t m_renderCount{ 0 }, m_transparentCount{ 0 }, m_onpcCount{ 0 };
        bool m_havePyramid{ false };
        uint8_t Check(int rc, const char* what);
    };
}
