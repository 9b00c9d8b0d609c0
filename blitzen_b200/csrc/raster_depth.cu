// Device-side software depth: closes the loop  early pass -> depth -> Hi-Z pyramid -> late pass  without a rasteriser.
// In the reference the depth the pyramid is built from is what DrawGeometry leaves in the depth attachment after drawing the early
// pass's list (frame order BlitzenVulkan/vulkanDraw.cpp:1015-1036; attachment cleared to 0 = far, reverse-Z d = zNear / z,
// BlitzenVulkan/vulkanResources.cpp:70-71).  Here every record of the current draw list is replaced by a conservative proxy of the
// object it draws: the screen-space bounding box of the object's bounding sphere (projectSphere, CullingShaderData.glsl:8-33), filled
// with the depth of the sphere's FAR side, d = zNear / (c.z + r) -- nothing in front of that depth inside the box can belong to a
// farther object -- and merged with max (reverse-Z: nearest wins) into a W x H fp32 target.  Spheres that cross the near plane
// (projectSphere returns false) draw nothing.  Pixel (x, y) is covered when  floor(aabb.x * W) <= x < ceil(aabb.z * W)  and the same in
// y, clamped to the target.  max is order-independent, so the result is deterministic; the CPU restatement used by the tests (oracle_raster_depth) is the
// same arithmetic on the CPU and the two are compared bit for bit (tests/test_raster_depth_gpu.py).
//
// Work distribution: one thread per record.  Boxes of <= 16 pixels (the vast majority: the mean box of the bench scene is ~2 pixels) are
// filled by their thread; larger ones by the whole warp, one after the other; boxes above 4096 pixels (a handful of objects next to the
// camera) are queued in shared memory and filled by the whole CTA at the end.
#include "ctx.h"
#include "cull_math.cuh"

namespace blz {

namespace {

constexpr int kRasterThreads = 256;
constexpr uint32_t kLaneArea = 16u, kWarpArea = 4096u;

struct RasterParams {
    const uint32_t* draws; const uint32_t* counts; uint32_t recWords;
    const RenderObject* objs; uint32_t n, objectIdBase, transformIdBase;
    const MeshTransform* xf; const PrimitiveSurface* surfaces; uint32_t surfaceCount;
    uint32_t* depth; uint32_t W, H;
    ViewConsts view;
};

// clamp(floor(f), 0, n) and clamp(ceil(f), 0, n) of a float pixel coordinate; NaN -> 0
__device__ __forceinline__ uint32_t pix_floor(float f, uint32_t n) { if (!(f >= 0.0f)) return 0u; if (f >= float(n)) return n; return uint32_t(floorf(f)); }
__device__ __forceinline__ uint32_t pix_ceil(float f, uint32_t n) { if (!(f >= 0.0f)) return 0u; if (f >= float(n)) return n; return uint32_t(ceilf(f)); }

struct Box { uint32_t x0, y0, x1, y1, bits; };

__global__ void __launch_bounds__(kRasterThreads) raster_depth_kernel(const __grid_constant__ RasterParams p)
{
    __shared__ Box s_big[kRasterThreads];
    __shared__ uint32_t s_nBig;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint32_t count = p.counts[0];
    const ViewConsts& V = p.view;
    for (uint32_t base = blockIdx.x * kRasterThreads; base < count; base += gridDim.x * kRasterThreads) {     // uniform over the CTA
        if (tid == 0) s_nBig = 0u;
        __syncthreads();
        const uint32_t r = base + tid;
        Box b{ 0u, 0u, 0u, 0u, 0u };
        if (r < count) {
            const uint32_t local = p.draws[size_t(r) * p.recWords] - p.objectIdBase;
            if (local < p.n) {
                const RenderObject o = p.objs[local];
                if (o.surfaceId < p.surfaceCount) {
                    float4 ps, qt;
                    ld_transform(p.xf + (o.transformId - p.transformIdBase), ps, qt);
                    const float4 bs = *reinterpret_cast<const float4*>(&p.surfaces[o.surfaceId]);
                    const Sphere s = view_space_sphere(bs.x, bs.y, bs.z, bs.w, ps.x, ps.y, ps.z, ps.w, qt.x, qt.y, qt.z, qt.w, V);
                    float4 aabb;
                    if (project_sphere(s, V.zNear, V.proj0, V.proj5, aabb)) {
                        b.x0 = pix_floor(fmul(aabb.x, float(p.W)), p.W); b.x1 = pix_ceil(fmul(aabb.z, float(p.W)), p.W);
                        b.y0 = pix_floor(fmul(aabb.y, float(p.H)), p.H); b.y1 = pix_ceil(fmul(aabb.w, float(p.H)), p.H);
                        const float d = fdiv(V.zNear, fadd(s.z, s.r));
                        b.bits = d > 0.0f ? __float_as_uint(d) : 0u;          // 0 = the clear value: such a box changes nothing
                    }
                }
            }
        }
        const uint32_t w = b.x1 > b.x0 ? b.x1 - b.x0 : 0u, h = b.y1 > b.y0 ? b.y1 - b.y0 : 0u;
        const uint32_t area = b.bits ? w * h : 0u;                             // w, h <= 65535 for any sane target: no overflow below 2^32
        if (area != 0u && area <= kLaneArea) {
            for (uint32_t y = b.y0; y < b.y1; ++y)
                for (uint32_t x = b.x0; x < b.x1; ++x) atomicMax(p.depth + size_t(y) * p.W + x, b.bits);
        }
        // medium boxes: the warp fills them together, one box at a time
        uint32_t med = __ballot_sync(0xFFFFFFFFu, area > kLaneArea && area <= kWarpArea);
        while (med) {
            const int src = __ffs(int(med)) - 1;
            med &= med - 1u;
            const uint32_t x0 = __shfl_sync(0xFFFFFFFFu, b.x0, src), y0 = __shfl_sync(0xFFFFFFFFu, b.y0, src);
            const uint32_t ww = __shfl_sync(0xFFFFFFFFu, w, src), aa = __shfl_sync(0xFFFFFFFFu, area, src), bits = __shfl_sync(0xFFFFFFFFu, b.bits, src);
            for (uint32_t i = lane; i < aa; i += 32u) atomicMax(p.depth + size_t(y0 + i / ww) * p.W + (x0 + i % ww), bits);
        }
        // large boxes: queued, then filled by the whole CTA
        if (area > kWarpArea) s_big[atomicAdd(&s_nBig, 1u)] = b;
        __syncthreads();
        const uint32_t nBig = s_nBig;
        for (uint32_t k = 0; k < nBig; ++k) {
            const Box g = s_big[k];
            const uint32_t ww = g.x1 - g.x0, aa = ww * (g.y1 - g.y0);
            for (uint32_t i = tid; i < aa; i += kRasterThreads) atomicMax(p.depth + size_t(g.y0 + i / ww) * p.W + (g.x0 + i % ww), g.bits);
        }
        __syncthreads();
    }
}

} // namespace

} // namespace blz

using namespace blz;

extern "C" {

// Clears the context's own W x H depth target to 0 (far) and splats the CURRENT draw list of `list` into it under the current view;
// the target becomes the depth image blz_cull_build_pyramid reads (as after blz_cull_set_depth).  Stream-ordered, no host sync.
int blz_cull_raster_depth(blz_cull_ctx* c, int list, uint32_t width, uint32_t height)
{
    if (!c) return fail(BLZ_ERR_INVALID, "null context");
    if (list < 0 || list > 2 || !c->draws || !c->objs[list]) return fail(BLZ_ERR_INVALID, "no scene / list %d", list);
    if (!c->haveView) return fail(BLZ_ERR_INVALID, "no view set");
    if (width == 0 || height == 0 || width > 65535u || height > 65535u) return fail(BLZ_ERR_INVALID, "depth target %ux%u", width, height);
    if (c->lastRecWords != 6u && c->lastRecWords != 8u) return fail(BLZ_ERR_STATE, "the current draw list is not a list of draw records");
    CU_TRY(cudaSetDevice(c->device));
    const size_t texels = size_t(width) * height;
    if (texels > c->depthOwnedTexels) {
        CU_TRY(cudaStreamSynchronize(c->stream));
        if (c->depthOwned) { cudaFree(c->depthOwned); c->depthOwned = nullptr; }
        c->depthOwnedTexels = 0;
        CU_TRY(cudaMalloc(&c->depthOwned, texels * sizeof(float)));
        c->depthOwnedTexels = texels;
    }
    CU_TRY(cudaMemsetAsync(c->depthOwned, 0, texels * sizeof(float), c->stream));
    RasterParams p{};
    p.draws = c->draws; p.counts = c->drawCounts; p.recWords = c->lastRecWords;
    p.objs = c->objs[list]; p.n = c->nObjs[list]; p.objectIdBase = list == BLZ_LIST_OPAQUE ? c->objectIdBase : 0u; p.transformIdBase = c->transformIdBase;
    p.xf = c->xf; p.surfaces = c->surf; p.surfaceCount = c->nSurf;
    p.depth = reinterpret_cast<uint32_t*>(c->depthOwned); p.W = width; p.H = height;
    p.view = blz::make_view_consts(c->view);
    raster_depth_kernel<<<c->numSMs * 8, kRasterThreads, 0, c->stream>>>(p);
    CU_TRY(cudaGetLastError());
    c->launches++;
    c->depth = c->depthOwned; c->depthW = width; c->depthH = height;
    return BLZ_OK;
}

int blz_cull_read_depth(blz_cull_ctx* c, float* host, uint64_t capacityTexels, uint32_t* outWH)
{
    if (!c || !c->depth) return fail(BLZ_ERR_INVALID, "no depth image set");
    const uint64_t texels = uint64_t(c->depthW) * c->depthH;
    if (outWH) { outWH[0] = c->depthW; outWH[1] = c->depthH; }
    if (!host) return BLZ_OK;
    if (capacityTexels < texels) return fail(BLZ_ERR_INVALID, "depth image has %llu texels, buffer holds %llu", (unsigned long long)texels, (unsigned long long)capacityTexels);
    CU_TRY(cudaSetDevice(c->device));
    CU_TRY(cudaMemcpyAsync(host, c->depth, texels * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return BLZ_OK;
}

} // extern "C"
