// Device-side arithmetic of the cull shaders.  Every floating-point operation is spelled with a round-to-nearest
// intrinsic (__fmul_rn / __fadd_rn / __fsub_rn / __fdiv_rn / __fsqrt_rn), which nvcc never contracts into FMA and never
// replaces by an approximate sequence, so results are IEEE-754 binary32 exact and independent of -fmad / -use_fast_math.
// The evaluation order is the order written in the reference's shader source, left to right:
//   RotateQuat                  VulkanShaderHeaders/ShaderBuffers.glsl:197-200, HlslShaders/Headers/hlslMath.hlsl:1-4
//   IsObjectInsideViewFrustum   VulkanShaderHeaders/CullingShaderData.glsl:36-58, hlslMath.hlsl:13-27 + CS/drawCull.cs.hlsl:25-27
//   projectSphere               CullingShaderData.glsl:8-33, hlslMath.hlsl:30-56
//   OcclusionCullingPassed      CullingShaderData.glsl:60-73 (Vulkan: MIN-reduction bilinear footprint)
//   OcclusionCheck              hlslMath.hlsl:58-78 (D3D12: one point texel)
//   LODSelection                CullingShaderData.glsl:103-116, HlslShaders/Headers/cullBuffers.hlsl:42-55
// (paths relative to /root/reference/src/Renderer)
#pragma once
#include "cull_types.cuh"

namespace blz {

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }

struct Sphere { float x, y, z, r; };

// The reference's 32-byte MeshTransform {pos.xyz, scale, orientation} (Resources/renderingResourcesTypes.h:124-129) read as uploaded, with
// ONE 256-bit load per object (sm_100 LDG.E.256): one request and exactly one 32-byte sector per gathered transform.
#ifndef BLZ_XF_LOAD
#define BLZ_XF_LOAD 0
#endif
__device__ __forceinline__ void ld_transform(const MeshTransform* p, float4& ps, float4& qt)
{
#if BLZ_XF_LOAD == 0
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(ps.x), "=f"(ps.y), "=f"(ps.z), "=f"(ps.w), "=f"(qt.x), "=f"(qt.y), "=f"(qt.z), "=f"(qt.w) : "l"(p));
#elif BLZ_XF_LOAD == 4
    // "+f": the destination registers are also inputs, so a loop-carried transform keeps ONE set of registers (with "=f" the compiler
    // merged the loaded and the carried values with moves placed right behind the load, which wait for the data)
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "+f"(ps.x), "+f"(ps.y), "+f"(ps.z), "+f"(ps.w), "+f"(qt.x), "+f"(qt.y), "+f"(qt.z), "+f"(qt.w) : "l"(p));
#elif BLZ_XF_LOAD == 1
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(ps.x), "=f"(ps.y), "=f"(ps.z), "=f"(ps.w), "=f"(qt.x), "=f"(qt.y), "=f"(qt.z), "=f"(qt.w) : "l"(p));
#elif BLZ_XF_LOAD == 2
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(ps.x), "=f"(ps.y), "=f"(ps.z), "=f"(ps.w) : "l"(p));
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4+16];" : "=f"(qt.x), "=f"(qt.y), "=f"(qt.z), "=f"(qt.w) : "l"(p));
#elif BLZ_XF_LOAD == 3
    asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(ps.x), "=f"(ps.y), "=f"(ps.z), "=f"(ps.w), "=f"(qt.x), "=f"(qt.y), "=f"(qt.z), "=f"(qt.w) : "l"(p));
#endif
}

// view-space bounding sphere: center = RotateQuat(bc, q) * scale + pos; center = (view * vec4(center, 1)).xyz; radius = br * scale
__device__ __forceinline__ Sphere view_space_sphere(float bcx, float bcy, float bcz, float br,
                                                    float px, float py, float pz, float scale,
                                                    float qx, float qy, float qz, float qw, const ViewConsts& V)
{
    // c1 = cross(q.xyz, v)   (GLSL cross: [a.y*b.z - b.y*a.z, a.z*b.x - b.z*a.x, a.x*b.y - b.x*a.y])
    float c1x = fsub(fmul(qy, bcz), fmul(bcy, qz));
    float c1y = fsub(fmul(qz, bcx), fmul(bcz, qx));
    float c1z = fsub(fmul(qx, bcy), fmul(bcx, qy));
    // t = c1 + q.w * v
    float tx = fadd(c1x, fmul(qw, bcx));
    float ty = fadd(c1y, fmul(qw, bcy));
    float tz = fadd(c1z, fmul(qw, bcz));
    // c2 = cross(q.xyz, t)
    float c2x = fsub(fmul(qy, tz), fmul(ty, qz));
    float c2y = fsub(fmul(qz, tx), fmul(tz, qx));
    float c2z = fsub(fmul(qx, ty), fmul(tx, qy));
    // r = v + 2.0 * c2 ; w = r * scale + pos
    float wx = fadd(fmul(fadd(bcx, fmul(2.0f, c2x)), scale), px);
    float wy = fadd(fmul(fadd(bcy, fmul(2.0f, c2y)), scale), py);
    float wz = fadd(fmul(fadd(bcz, fmul(2.0f, c2z)), scale), pz);
    // column-major mat4 * vec4(w, 1): ((m0*x + m4*y) + m8*z) + m12
    Sphere s;
    s.x = fadd(fadd(fadd(fmul(V.m[0], wx), fmul(V.m[3], wy)), fmul(V.m[6], wz)), V.m[9]);
    s.y = fadd(fadd(fadd(fmul(V.m[1], wx), fmul(V.m[4], wy)), fmul(V.m[7], wz)), V.m[10]);
    s.z = fadd(fadd(fadd(fmul(V.m[2], wx), fmul(V.m[5], wy)), fmul(V.m[8], wz)), V.m[11]);
    s.r = fmul(br, scale);
    return s;
}

__device__ __forceinline__ bool frustum_test(const Sphere& s, const ViewConsts& V)
{
    bool visible = fsub(fmul(s.z, V.frustumLeft), fmul(fabsf(s.x), V.frustumRight)) > -s.r;
    visible = visible && (fsub(fmul(s.z, V.frustumBottom), fmul(fabsf(s.y), V.frustumTop)) > -s.r);
    visible = visible && (fadd(s.z, s.r) > V.zNear) && (fsub(s.z, s.r) < V.zFar);
    return visible;
}

// returns false when the sphere crosses the near plane (the caller then keeps the object without a Hi-Z test)
__device__ __forceinline__ bool project_sphere(const Sphere& c, float znear, float P00, float P11, float4& aabb)
{
    if (c.z < fadd(c.r, znear)) return false;
    float crx = fmul(c.x, c.r), cry = fmul(c.y, c.r), crz = fmul(c.z, c.r);
    float czr2 = fsub(fmul(c.z, c.z), fmul(c.r, c.r));
    float vx = fsqrt(fadd(fmul(c.x, c.x), czr2));
    float minx = fdiv(fsub(fmul(vx, c.x), crz), fadd(fmul(vx, c.z), crx));
    float maxx = fdiv(fadd(fmul(vx, c.x), crz), fsub(fmul(vx, c.z), crx));
    float vy = fsqrt(fadd(fmul(c.y, c.y), czr2));
    float miny = fdiv(fsub(fmul(vy, c.y), crz), fadd(fmul(vy, c.z), cry));
    float maxy = fdiv(fadd(fmul(vy, c.y), crz), fsub(fmul(vy, c.z), cry));
    // aabb = vec4(minx*P00, miny*P11, maxx*P00, maxy*P11).xwzy * vec4(.5,-.5,.5,-.5) + .5
    aabb.x = fadd(fmul(fmul(minx, P00), 0.5f), 0.5f);
    aabb.y = fadd(fmul(fmul(maxy, P11), -0.5f), 0.5f);
    aabb.z = fadd(fmul(fmul(maxx, P00), 0.5f), 0.5f);
    aabb.w = fadd(fmul(fmul(miny, P11), -0.5f), 0.5f);
    return true;
}

// floor(log2(x)) for finite x > 0, exact (exponent field; denormals via clz)
__device__ __forceinline__ int ilog2_floor_pos(float x)
{
    uint32_t b = __float_as_uint(x);
    int e = int((b >> 23) & 0xFFu);
    if (e != 0) return e - 127;
    uint32_t m = b & 0x7FFFFFu;
    return (31 - __clz(int(m))) - 149;
}

__device__ __forceinline__ uint32_t umax(uint32_t a, uint32_t b) { return a > b ? a : b; }
__device__ __forceinline__ float min_keep(float a, float b) { return b < a ? b : a; }

// clamp a float texel index to [0, n-1]; NaN -> 0.  Written as  !(f >= 0) ? 0 : f >= float(n-1) ? n-1 : uint(f)  in the oracle; the
// saturating float -> int conversion (F2I.RZ: NaN -> 0, out of range -> INT_MIN / INT_MAX) followed by an integer clamp gives the same
// value for EVERY float (n - 1 < 2^31): three instructions instead of nine, four times per Hi-Z sample
__device__ __forceinline__ uint32_t clamp_index(float f, uint32_t n)
{
#ifdef BLZ_CLAMP_INDEX_COMPARE
    if (!(f >= 0.0f)) return 0u;
    float hi = float(n - 1);
    if (f >= hi) return n - 1;
    return uint32_t(f);
#else
    const int t = __float2int_rz(f);
    return uint32_t(min(max(t, 0), int(n - 1u)));
#endif
}

// LINEAR + VK_SAMPLER_REDUCTION_MODE_MIN + CLAMP_TO_EDGE sample of one mip (BlitzenVulkan/vulkanResources.cpp:51-55, :394-429):
// u = s*W - 0.5, i0 = floor(u), i1 = i0 + 1; MIN over the footprint texels whose weight is non-zero.
template <class Fetch>
__device__ __forceinline__ float sample_min_linear(Fetch&& fetch, uint32_t W, uint32_t H, float s, float t)
{
    float u = fsub(fmul(s, float(W)), 0.5f);
    float v = fsub(fmul(t, float(H)), 0.5f);
    float fu = floorf(u), fv = floorf(v);
    float au = fsub(u, fu), av = fsub(v, fv);
    uint32_t i0 = clamp_index(fu, W), i1 = clamp_index(fadd(fu, 1.0f), W);
    uint32_t j0 = clamp_index(fv, H), j1 = clamp_index(fadd(fv, 1.0f), H);
    bool useI1 = !(au == 0.0f), useJ1 = !(av == 0.0f);
    float d = fetch(i0, j0);
    if (useI1) d = min_keep(d, fetch(i1, j0));
    if (useJ1) {
        d = min_keep(d, fetch(i0, j1));
        if (useI1) d = min_keep(d, fetch(i1, j1));
    }
    return d;
}

__device__ __forceinline__ bool isinf_f(float x) { return (__float_as_uint(x) & 0x7FFFFFFFu) == 0x7F800000u; }

__device__ __forceinline__ bool hiz_test_vk(const float4& aabb, const PyramidDesc& P, const Sphere& c, const ViewConsts& V)
{
    float width = fmul(fsub(aabb.z, aabb.x), V.pyramidWidth);
    float height = fmul(fsub(aabb.w, aabb.y), V.pyramidHeight);
    float m = (width < height) ? height : width;                 // GLSL max(x, y)
    int level;
    if (!(m > 0.0f)) level = 0;
    else if (isinf_f(m)) level = int(P.mips) - 1;
    else level = ilog2_floor_pos(m);
    level = max(level, 0);
    level = min(level, int(P.mips) - 1);
    uint32_t W = umax(1u, P.width >> level), H = umax(1u, P.height >> level);
    const float* img = P.data + P.offset[level];
    float s = fmul(fadd(aabb.x, aabb.z), 0.5f), t = fmul(fadd(aabb.y, aabb.w), 0.5f);
    float depth = sample_min_linear([&](uint32_t i, uint32_t j) { return __ldg(img + size_t(j) * W + i); }, W, H, s, t);
    float depthSphere = fdiv(V.zNear, fsub(c.z, c.r));
    return depthSphere > depth;
}

// float -> uint: truncate, saturate, NaN -> 0 (what F2I.U32 does; HLSL leaves it undefined -- see DESIGN.md)
__device__ __forceinline__ uint32_t f2u_sat(float f) { return __float2uint_rz(f); }

__device__ __forceinline__ bool hiz_test_dx(float4 aabb, const PyramidDesc& P, const Sphere& c, const ViewConsts& V)
{
    uint32_t pw = f2u_sat(V.pyramidWidth), ph = f2u_sat(V.pyramidHeight);
    float width = fmul(fsub(aabb.z, aabb.x), float(pw));
    float height = fmul(fsub(aabb.w, aabb.y), float(ph));
    float m = (width < height) ? height : width;
    uint32_t level;
    if (!(m > 0.0f)) level = 0u;
    else if (isinf_f(m)) level = 0xFFFFFFFFu;
    else { int e = ilog2_floor_pos(m); level = e < 0 ? 0u : uint32_t(e); }
    uint32_t sh = level & 31u;
    uint32_t mipW = umax(1u, pw >> sh), mipH = umax(1u, ph >> sh);
    float fx = fmul(float(mipW), 0.5f), fy = fmul(float(mipH), 0.5f);
    aabb.x = fmul(aabb.x, fx); aabb.y = fmul(aabb.y, fy); aabb.z = fmul(aabb.z, fx); aabb.w = fmul(aabb.w, fy);
    uint32_t tx = f2u_sat(fadd(aabb.x, aabb.z)), ty = f2u_sat(fadd(aabb.y, aabb.w));
    float depth = 0.0f;                                           // out-of-range Load() returns 0
    if (level < P.mips) {
        uint32_t W = umax(1u, P.width >> level), H = umax(1u, P.height >> level);
        if (tx < W && ty < H) depth = __ldg(P.data + P.offset[level] + size_t(ty) * W + tx);
    }
    float depthSphere = fdiv(V.zNear, fsub(c.z, c.r));
    return depthSphere > depth;
}

// relative LOD index; `err(i)` returns lods[i].error
template <class ErrFn>
__device__ __forceinline__ uint32_t lod_select(const Sphere& c, float scale, float lodTarget, uint32_t lodOffset, uint32_t lodCount, ErrFn&& err)
{
    float len = fsqrt(fadd(fadd(fmul(c.x, c.x), fmul(c.y, c.y)), fmul(c.z, c.z)));
    float d = fsub(len, c.r);
    float distance = (d < 0.0f) ? 0.0f : d;                       // max(x, 0)
    float threshold = fdiv(fmul(distance, lodTarget), scale);
    uint32_t lodIndex = 0;
    for (uint32_t i = 1; i < lodCount; ++i)
        if (err(lodOffset + i) < threshold) lodIndex = i;
    return lodIndex;
}

} // namespace blz
