// Glue between the REFERENCE's frontend types and BlitzenCuda::CudaCullRenderer.  Only meaningful where the reference's
// headers are on the include path (-I<reference>/src -I<reference>/src/VendorCode, defines as in oracle/Makefile REF_DEFS);
// it uses the reference's own BlitCL::DynamicArray, RenderContainer, MeshResources and CameraViewData, nothing is re-typed.
// INTEGRATION.md shows the three edits a maintainer makes in the reference to route its cull dispatch through this.
#pragma once
#include "blitzenCudaCull.h"
#include "Renderer/Resources/RenderObject/blitRender.h"     // RenderContainer, MeshResources (reference)
#include "Game/blitCamera.h"                                 // Camera, CameraViewData (reference)

namespace BlitzenCuda
{
    static_assert(sizeof(BlitzenEngine::RenderObject) == 8 && sizeof(BlitzenEngine::MeshTransform) == 32, "reference layout changed");
    static_assert(sizeof(BlitzenEngine::PrimitiveSurface) == 32 && sizeof(BlitzenEngine::LodData) == 32, "reference layout changed");
    static_assert(sizeof(BlitzenEngine::Cluster) == 32 && sizeof(BlitzenEngine::LodInstanceCounter) == 8, "reference layout changed");
    static_assert(sizeof(BlitzenEngine::CameraViewData) == 256, "reference layout changed");

    // the reads SetupForRendering performs on DrawContext (BlitzenVulkan/vulkanRendererSetup.cpp:826-907, StaticBuffersInit :365-666)
    inline CullScene MakeCullScene(BlitzenEngine::RenderContainer& renders, BlitzenEngine::MeshResources& meshes)
    {
        CullScene s;
        s.pRenders = renders.m_renders;                       s.renderCount = renders.m_renderCount;
        s.pTransparentRenders = renders.m_transparentRenders; s.transparentRenderCount = renders.m_transparentRenderCount;
        s.pOnpcRenders = renders.m_onpcRenders;               s.onpcRenderCount = renders.m_onpcRenderCount;
        s.pTransforms = renders.m_transforms;                 s.transformCount = renders.m_transformCount;     // vulkanRendererSetup.cpp:284-285 copies [0, m_transformCount)
        s.pSurfaces = meshes.m_surfaces.Data();               s.surfaceCount = uint32_t(meshes.m_surfaces.GetSize());
        s.pLods = meshes.m_LODs.Data();                       s.lodCount = uint32_t(meshes.m_LODs.GetSize());
        s.pClusters = meshes.m_clusters.Data();               s.clusterCount = uint32_t(meshes.m_clusters.GetSize());
        s.pLodInstances = meshes.m_lodInstanceList.Data();    s.lodInstanceCount = uint32_t(meshes.m_lodInstanceList.GetSize());
        return s;
    }

    // the per-frame part of DrawFrame that feeds the cull: view data (frozen-frustum rule of vulkanDraw.cpp:832-840) and
    // the dynamic transforms (UpdateBuffers, vulkanDraw.cpp:46-60 copies the first m_dynamicTransformCount transforms)
    inline void UpdatePerFrame(CudaCullRenderer& r, BlitzenEngine::Camera& camera, BlitzenEngine::RenderContainer& renders)
    {
        if (renders.m_dynamicTransformCount) r.UpdateObjectTransforms(0, renders.m_dynamicTransformCount, renders.m_transforms);
        r.Update(&camera.viewData);
    }
}
