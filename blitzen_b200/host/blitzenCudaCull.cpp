// See blitzenCudaCull.h.  Thin: every member is one or two C-ABI calls plus the reference's error convention.
#include "blitzenCudaCull.h"
#include "../../include/blz_cull.h"
#include <cstdio>
#include <cstring>

namespace BlitzenCuda
{
    namespace
    {
        void DefaultLog(int level, const char* message) { fprintf(stderr, "%s %s\n", level ? "[ERROR]:" : "[INFO]:", message); }
        LogFn g_log = DefaultLog;
    }

    void SetLogCallback(LogFn fn) { g_log = fn ? fn : DefaultLog; }

    uint8_t CudaCullRenderer::Check(int rc, const char* what)
    {
        if (rc == BLZ_OK) return 1;
        char buf[1200];
        snprintf(buf, sizeof(buf), "CudaCullRenderer::%s failed (%d): %s", what, rc, blz_cull_last_error());
        g_log(1, buf);
        return 0;
    }

    CudaCullRenderer::~CudaCullRenderer() { Shutdown(); }

    uint8_t CudaCullRenderer::Init(int cudaDevice, HiZVariant hiz, RecordFormat format)
    {
        Shutdown();
        m_hiz = hiz; m_format = format;
        if (!Check(blz_cull_create(cudaDevice, &m_ctx), "Init")) { m_ctx = nullptr; return 0; }
        g_log(0, "CudaCullRenderer initialised (B200 cull backend, no CPU fallback)");
        return 1;
    }

    void CudaCullRenderer::Shutdown()
    {
        if (m_ctx) { blz_cull_destroy(m_ctx); m_ctx = nullptr; }
        m_havePyramid = false;
    }

    uint8_t CudaCullRenderer::SetupForRendering(const CullScene& s)
    {
        if (!m_ctx) return Check(BLZ_ERR_INVALID, "SetupForRendering (Init was not called)");
        blz_scene_desc d;
        memset(&d, 0, sizeof(d));
        d.renders = s.pRenders; d.render_count = s.renderCount;
        d.transparent_renders = s.pTransparentRenders; d.transparent_count = s.transparentRenderCount;
        d.onpc_renders = s.pOnpcRenders; d.onpc_count = s.onpcRenderCount;
        d.transforms = s.pTransforms; d.transform_count = s.transformCount;
        d.surfaces = s.pSurfaces; d.surface_count = s.surfaceCount;
        d.lods = s.pLods; d.lod_count = s.lodCount;
        d.clusters = s.pClusters; d.cluster_count = s.clusterCount;
        d.lod_instances = s.pLodInstances; d.lod_instance_count = s.lodInstanceCount;
        d.draw_capacity = s.drawCapacity; d.cluster_dispatch_capacity = s.clusterDispatchCapacity;
        d.instance_bucket_capacity = s.pInstanceBucketCapacity;
        if (!Check(blz_cull_upload_scene(m_ctx, &d), "SetupForRendering")) return 0;
        m_renderCount = s.renderCount; m_transparentCount = s.transparentRenderCount; m_onpcCount = s.onpcRenderCount;
        return Check(blz_cull_synchronize(m_ctx), "SetupForRendering");
    }

    void CudaCullRenderer::UpdateObjectTransform(uint32_t transformId, const void* pTransform) { UpdateObjectTransforms(transformId, 1, pTransform); }

    uint8_t CudaCullRenderer::UpdateObjectTransforms(uint32_t first, uint32_t count, const void* pTransforms)
    {
        return Check(blz_cull_update_transforms(m_ctx, first, count, pTransforms), "UpdateObjectTransform");
    }

    void CudaCullRenderer::Update(const void* pCameraViewData) { Check(blz_cull_set_view(m_ctx, pCameraViewData), "Update"); }

    uint8_t CudaCullRenderer::SetDepthAttachment(const float* pDepth, uint32_t w, uint32_t h, bool onDevice)
    {
        return Check(onDevice ? blz_cull_set_depth_device(m_ctx, pDepth, w, h) : blz_cull_set_depth(m_ctx, pDepth, w, h), "SetDepthAttachment");
    }

    uint8_t CudaCullRenderer::ClearDepthPyramid(uint32_t w, uint32_t h)
    {
        if (!Check(blz_cull_clear_pyramid(m_ctx, int(m_hiz), w, h), "ClearDepthPyramid")) return 0;
        m_havePyramid = true;
        return 1;
    }

    uint8_t CudaCullRenderer::DrawCullFirstPass(CullPipeline pipeline)
    {
        switch (pipeline) {
        case CullPipeline::Initial: return Check(blz_cull_early(m_ctx, int(m_format)), "DrawCullFirstPass(initial)");
        case CullPipeline::Transparent: return Check(blz_cull_frustum_lod(m_ctx, BLZ_LIST_TRANSPARENT, int(m_format), 0), "DrawCullFirstPass(transparent)");
        case CullPipeline::Onpc: return Check(blz_cull_frustum_lod(m_ctx, BLZ_LIST_ONPC, int(m_format), BLZ_FLAG_ONPC_LOD_QUIRK), "DrawCullFirstPass(onpc)");
        }
        return Check(BLZ_ERR_INVALID, "DrawCullFirstPass");
    }

    uint8_t CudaCullRenderer::GenerateDepthPyramid()
    {
        if (!Check(blz_cull_build_pyramid(m_ctx, int(m_hiz)), "GenerateDepthPyramid")) return 0;
        m_havePyramid = true;
        return 1;
    }

    uint8_t CudaCullRenderer::DrawCullOcclusionPass() { return Check(blz_cull_late(m_ctx, int(m_format), int(m_hiz)), "DrawCullOcclusionPass"); }
    uint8_t CudaCullRenderer::PreClusterDrawCull() { return Check(blz_cull_cluster_expand(m_ctx, BLZ_LIST_OPAQUE), "PreClusterDrawCull"); }
    uint8_t CudaCullRenderer::ClusterCull(ClusterMode mode) { return Check(blz_cull_cluster_cull(m_ctx, int(mode), int(m_format), int(m_hiz)), "ClusterCull"); }
    uint8_t CudaCullRenderer::DrawInstanceCullPass() { return Check(blz_cull_instanced(m_ctx, BLZ_LIST_OPAQUE), "DrawInstanceCullPass"); }

    uint8_t CudaCullRenderer::DrawFrameCull(CullStats* st)
    {
        CullStats local; if (!st) st = &local;
        if (!DrawCullFirstPass(CullPipeline::Initial)) return 0;                               // vulkanDraw.cpp:1015
        if (!ReadDrawCount(&st->earlyDrawCount, &st->earlyTotal)) return 0;                    // (first DrawGeometry consumes the list here, :1020)
        if (!GenerateDepthPyramid()) return 0;                                                 // :1026
        if (!DrawCullOcclusionPass()) return 0;                                                // :1031
        if (!ReadDrawCount(&st->lateDrawCount, &st->lateTotal)) return 0;
        if (m_transparentCount) {                                                              // :1059-1061
            if (!DrawCullFirstPass(CullPipeline::Transparent)) return 0;
            if (!ReadDrawCount(&st->transparentDrawCount, &st->transparentTotal)) return 0;
        }
        return 1;
    }

    void* CudaCullRenderer::IndirectDrawBuffer() const { blz_outputs o; return (m_ctx && blz_cull_get_outputs(m_ctx, &o) == BLZ_OK) ? o.draws : nullptr; }
    uint32_t* CudaCullRenderer::IndirectCountBuffer() const { blz_outputs o; return (m_ctx && blz_cull_get_outputs(m_ctx, &o) == BLZ_OK) ? o.draw_count : nullptr; }
    uint8_t CudaCullRenderer::ReadDrawCount(uint32_t* w, uint32_t* t) { return Check(blz_cull_read_count(m_ctx, w, t), "ReadDrawCount"); }
    uint8_t CudaCullRenderer::ReadDraws(void* p, uint64_t cap, uint32_t* w, uint32_t* t) { return Check(blz_cull_read_draws(m_ctx, int(m_format), p, cap, w, t), "ReadDraws"); }
    uint8_t CudaCullRenderer::ReadVisibility(uint32_t* p) { return Check(blz_cull_read_visibility(m_ctx, p), "ReadVisibility"); }
    uint8_t CudaCullRenderer::WaitIdle() { return Check(blz_cull_synchronize(m_ctx), "WaitIdle"); }
}
