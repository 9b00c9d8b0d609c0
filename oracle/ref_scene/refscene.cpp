// TEST INFRASTRUCTURE (oracle side) -- not part of the shipped product.
//
// Drives the REFERENCE's own CPU frontend (compiled from /root/reference, see
// oracle/Makefile target `_ref/refscene`) to produce the exact inputs the reference
// backend would upload for its stress scenes, and dumps them as a flat blob.
// Nothing here re-types reference logic: it only calls reference functions in
// the order the reference's own startup calls them:
//   RenderingResourcesInit  -> LoadMeshFromObj(bunny)        Renderer/Interface/blitzenRenderer.cpp:49
//   CreateSceneFromArguments-> LoadTestGeometry              Renderer/Interface/blitzenRenderer.cpp:154
//                           -> CreateSingleRender(bunny,5)   :155
//                           -> 1000x RandomizeTransform(100,1) + kitten, dynamic   :128-150
//                           -> LoadGeometryStressTest(3000 | 2000)                 :167 / :183
//   SetupCamera(mainCamera)                                  Core/blitzenEntry.cpp:35
//
// Blob layout (little endian), see blitzen_b200/sceneio.py for the reader:
//   u32 magic 'BLZS', u32 version=1,
//   u32 nObjects, nTransforms, nSurfaces, nLods, nClusters, nLodInst, u32 staticTransformOffset, u32 dynamicTransformCount
//   CameraViewData (256 B)
//   RenderObject[nObjects] (8 B), MeshTransform[nTransforms] (32 B), PrimitiveSurface[nSurfaces] (32 B),
//   LodData[nLods] (32 B), Cluster[nClusters] (32 B), LodInstanceCounter[nLodInst] (8 B)
#include "Game/blitCamera.h"
#include "Renderer/Resources/Mesh/blitMeshes.h"
#include "Renderer/Resources/RenderObject/blitRender.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <memory>

using namespace BlitzenEngine;

static_assert(sizeof(RenderObject) == 8, "");
static_assert(sizeof(MeshTransform) == 32, "");
static_assert(sizeof(PrimitiveSurface) == 32, "");
static_assert(sizeof(LodData) == 32, "");
static_assert(sizeof(Cluster) == 32, "");
static_assert(sizeof(LodInstanceCounter) == 8, "");
static_assert(sizeof(CameraViewData) == 256, "");

static void put(FILE* f, const void* p, size_t n) { if (n && fwrite(p, 1, n, f) != n) { perror("fwrite"); exit(2); } }

int main(int argc, char** argv)
{
    if (argc < 4) {
        fprintf(stderr, "usage: refscene <reference_root> <mode: meshes|stress|instancing|dynamic> <out.blob> [maxObjects]\n");
        return 1;
    }
    std::string root = argv[1];
    std::string mode = argv[2];
    const char* out = argv[3];

    if (mode == "views") {
        // argv[4] = text file, one view per line: fovDeg winW winH zNear px py pz zFar yawRad pitchRad
        // Each line goes through the reference's own SetupCamera (Game/blitzenCamera.cpp:14-48); the blob is
        // u32 count followed by count x CameraViewData (256 B).
        if (argc < 5) { fprintf(stderr, "views mode needs a spec file\n"); return 1; }
        FILE* in = fopen(argv[4], "r");
        if (!in) { perror(argv[4]); return 2; }
        FILE* f = fopen(out, "wb");
        if (!f) { perror(out); return 2; }
        uint32_t count = 0; put(f, &count, 4);
        float v[10];
        while (fscanf(in, "%f %f %f %f %f %f %f %f %f %f", v, v + 1, v + 2, v + 3, v + 4, v + 5, v + 6, v + 7, v + 8, v + 9) == 10) {
            Camera cam{};
            SetupCamera(cam, BlitML::Radians(v[0]), v[1], v[2], v[3], BlitML::vec3{ v[4], v[5], v[6] }, v[7], v[8], v[9]);
            put(f, &cam.viewData, sizeof(CameraViewData));
            ++count;
        }
        fseek(f, 0, SEEK_SET); put(f, &count, 4);
        fclose(f); fclose(in);
        fprintf(stderr, "refscene: wrote %u views\n", count);
        return 0;
    }
    uint32_t maxObjects = argc > 4 ? (uint32_t)strtoul(argv[4], nullptr, 10) : 0xFFFFFFFFu;

    auto meshes = std::make_unique<MeshResources>();
    auto renders = std::make_unique<RenderContainer>();
    Camera camera{};

    SetupCamera(camera);

    // same load order as the reference: bunny (default mesh) first, then LoadTestGeometry's three
    const char* files[4][2] = { {"bunny.obj", "bunny"}, {"dragon.obj", "dragon"}, {"kitten.obj", "kitten"}, {"FinalBaseMesh.obj", "human"} };
    for (auto& f : files) {
        std::string p = root + "/Assets/Meshes/" + f[0];
        if (!LoadMeshFromObj(*meshes, p.c_str(), f[1])) { fprintf(stderr, "failed to load %s\n", p.c_str()); return 3; }
    }

    if (mode != "meshes") {
        CreateSingleRender(*renders, *meshes, BlitzenCore::Ce_DefaultMeshName, 5.f);
        // CreateDynamicObjectRendererTest without the GameObject bookkeeping (no cull-visible effect)
        uint32_t kitten = meshes->m_meshMap["kitten"].meshId;
        for (uint32_t i = 0; i < BlitzenCore::Ce_MaxDynamicObjectCount; ++i) {
            MeshTransform t;
            RandomizeTransform(t, 100.f, 1.f);
            CreateRenderObjectFromMesh(*renders, *meshes, kitten, t, true);
        }
        if (mode == "stress") LoadGeometryStressTest(*renders, *meshes, 3000.f);
        else if (mode == "instancing") LoadGeometryStressTest(*renders, *meshes, 2000.f);
        else if (mode != "dynamic") { fprintf(stderr, "bad mode\n"); return 1; }
    }

    uint32_t nObj = renders->m_renderCount < maxObjects ? renders->m_renderCount : maxObjects;
    // transforms [0, staticTransformOffset) as the backend uploads them (vulkanRendererSetup.cpp:284-285 quirk:
    // the reference copies m_transformCount elements from index 0; exact only with 1000 dynamic objects)
    uint32_t nXf = mode == "meshes" ? 0 : renders->m_staticTransformOffset;
    if (maxObjects != 0xFFFFFFFFu) {
        // truncated dump: keep only transforms referenced by the kept objects
        uint32_t mx = 0;
        for (uint32_t i = 0; i < nObj; ++i) if (renders->m_renders[i].transformId + 1 > mx) mx = renders->m_renders[i].transformId + 1;
        nXf = mx;
    }

    FILE* f = fopen(out, "wb");
    if (!f) { perror(out); return 2; }
    uint32_t hdr[10] = { 0x535A4C42u /*BLZS*/, 1u, nObj, nXf,
        (uint32_t)meshes->m_surfaces.GetSize(), (uint32_t)meshes->m_LODs.GetSize(), (uint32_t)meshes->m_clusters.GetSize(),
        (uint32_t)meshes->m_lodInstanceList.GetSize(), renders->m_staticTransformOffset, renders->m_dynamicTransformCount };
    put(f, hdr, sizeof(hdr));
    put(f, &camera.viewData, sizeof(CameraViewData));
    put(f, renders->m_renders, size_t(nObj) * sizeof(RenderObject));
    put(f, renders->m_transforms, size_t(nXf) * sizeof(MeshTransform));
    put(f, meshes->m_surfaces.Data(), meshes->m_surfaces.GetSize() * sizeof(PrimitiveSurface));
    put(f, meshes->m_LODs.Data(), meshes->m_LODs.GetSize() * sizeof(LodData));
    put(f, meshes->m_clusters.Data(), meshes->m_clusters.GetSize() * sizeof(Cluster));
    put(f, meshes->m_lodInstanceList.Data(), meshes->m_lodInstanceList.GetSize() * sizeof(LodInstanceCounter));
    fclose(f);
    fprintf(stderr, "refscene: mode=%s objects=%u transforms=%u surfaces=%zu lods=%zu clusters=%zu vertices=%zu indices=%zu\n",
        mode.c_str(), nObj, nXf, meshes->m_surfaces.GetSize(), meshes->m_LODs.GetSize(), meshes->m_clusters.GetSize(),
        meshes->m_vertices.GetSize(), meshes->m_indices.GetSize());
    return 0;
}
