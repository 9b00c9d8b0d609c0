// ============================================================================
// TEST INFRASTRUCTURE -- CPU oracle for the Blitzen cull dispatch.  NOT shipped.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library.  The product (blitzen_b200/) never does.
//
// What this is: a plain C++ restatement ("transliteration") of the reference's
// cull compute shaders, one function per shader function / main, each citing the
// reference file:line it follows (paths relative to /root/reference/src/Renderer).
// The reference implements this path ONLY as GLSL/HLSL compute shaders; there is
// no CPU implementation and no test in the reference (the reference ships no
// golden vectors).  The oracle is pinned instead by
//   (1) inputs produced by the reference's own compiled frontend (oracle/_ref/refscene),
//   (2) golden outputs obtained by EXECUTING the reference's own GLSL shaders --
//       compiled to SPIR-V with the reference's bundled glslang -- in the interpreter
//       under oracle/spirv_interp (tests/golden/spirv_golden.npz, tests/test_oracle_golden.py):
//       every Vulkan cull shader and the pyramid shader, byte-exact, and
//   (3) hand-derived known-answer tests (tests/test_oracle_kat.py).
// PARITY UNPINNED for the D3D12-only parts (HlslShaders/CS/*.hlsl: the point-sample
// OcclusionCheck, the 2x2-min pyramid, indirect instancing): dxc is a Windows binary,
// so those functions are restatements checked only by (1) and (3).
//
// Float rules: strict IEEE-754 binary32, no contraction
// (-ffp-contract=off, no -ffast-math), evaluation order exactly as written in
// the shader source, left to right.  floor(log2(x)) is taken from the exponent
// bits (mathematically exact), never from a libm log2f.
//
// Output order: ascending objectId (ascending record index in cluster mode,
// ascending LOD id for instanced commands).  The reference's own order is the
// nondeterministic atomic-arrival order; equality with it is as multisets.
// ============================================================================
#include <cstdint>
#include <cstring>
#include <cmath>
#include <array>
#include <vector>
#include <thread>
#include <algorithm>
#include <chrono>

namespace {

// ---- reference POD layouts (Resources/renderingResourcesTypes.h) -------------------
struct RenderObject { uint32_t transformId, surfaceId; };                       // :155-159
struct MeshTransform { float pos[3]; float scale; float q[4]; };                // :124-129
struct PrimitiveSurface { float center[3]; float radius; uint32_t materialId, lodOffset, lodCount, vertexOffset; }; // :104-116
struct LodData { uint32_t indexCount, firstIndex, clusterOffset, clusterCount; float error; uint32_t pad[3]; };     // :73-96
struct Cluster { float center[3]; float radius; int8_t coneAxis[3]; int8_t coneCutoff; uint32_t dataOffset;
                 uint8_t vertexCount, triangleCount, pad0, pad1; uint32_t tail; };                                    // :27-47
struct LodInstanceCounter { uint32_t instanceOffset, instanceCount; };           // :98-102
struct ViewData {                                                                // Game/blitCamera.h:38-64
    float view[16]; float projView[16]; float position[3];
    float frustumRight, frustumLeft, frustumTop, frustumBottom;
    float proj0, proj5, zNear, zFar, pyramidWidth, pyramidHeight, lodTarget;
    uint8_t pad[72];
};
struct ClusterDispatchData { uint32_t objectId, lodIndex, clusterId; };          // VulkanShaderHeaders/CullingShaderData.glsl:118-123
static_assert(sizeof(RenderObject) == 8 && sizeof(MeshTransform) == 32 && sizeof(PrimitiveSurface) == 32, "");
static_assert(sizeof(LodData) == 32 && sizeof(Cluster) == 32 && sizeof(ViewData) == 256, "");

struct vec3 { float x, y, z; };
struct vec4 { float x, y, z, w; };

// GLSL cross(): [x1*y2 - y1*x2, x2*y0 - y2*x0, x0*y1 - y0*x1]
inline vec3 cross(vec3 a, vec3 b)
{
    vec3 r;
    r.x = a.y * b.z - b.y * a.z;
    r.y = a.z * b.x - b.z * a.x;
    r.z = a.x * b.y - b.x * a.y;
    return r;
}

// VulkanShaderHeaders/ShaderBuffers.glsl:197-200, HlslShaders/Headers/hlslMath.hlsl:1-4
//   return v + 2.0 * cross(quat.xyz, cross(quat.xyz, v) + quat.w * v);
inline vec3 RotateQuat(vec3 v, vec4 q)
{
    vec3 qv{ q.x, q.y, q.z };
    vec3 c1 = cross(qv, v);
    vec3 t{ c1.x + q.w * v.x, c1.y + q.w * v.y, c1.z + q.w * v.z };
    vec3 c2 = cross(qv, t);
    return vec3{ v.x + 2.0f * c2.x, v.y + 2.0f * c2.y, v.z + 2.0f * c2.z };
}

// VulkanShaderHeaders/CullingShaderData.glsl:36-58 (IsObjectInsideViewFrustum)
// == hlslMath.hlsl:13-27 (FrustumCheck) + the inline prologue of HlslShaders/CS/drawCull.cs.hlsl:25-27.
// Column-major mat4 * vec4(center,1): x' = ((m0*x + m4*y) + m8*z) + m12   (BlitzenMathLibrary/blitMLTypes.h:184-193 order)
inline bool IsObjectInsideViewFrustum(vec3& center, float& radius, vec3 boundCenter, float boundRadius,
    float scale, vec3 pos, vec4 orientation, const float* view,
    float frustumRight, float frustumLeft, float frustumTop, float frustumBottom, float znear, float zfar)
{
    vec3 r = RotateQuat(boundCenter, orientation);
    vec3 w{ r.x * scale + pos.x, r.y * scale + pos.y, r.z * scale + pos.z };
    center.x = ((view[0] * w.x + view[4] * w.y) + view[8] * w.z) + view[12];
    center.y = ((view[1] * w.x + view[5] * w.y) + view[9] * w.z) + view[13];
    center.z = ((view[2] * w.x + view[6] * w.y) + view[10] * w.z) + view[14];
    radius = boundRadius * scale;
    bool visible = true;
    visible = visible && center.z * frustumLeft - std::fabs(center.x) * frustumRight > -radius;
    visible = visible && center.z * frustumBottom - std::fabs(center.y) * frustumTop > -radius;
    visible = visible && center.z + radius > znear && center.z - radius < zfar;
    return visible;
}

// CullingShaderData.glsl:8-33 (projectSphere) == hlslMath.hlsl:30-56 (ProjectSphere)
inline bool projectSphere(vec3 c, float r, float znear, float P00, float P11, vec4& aabb)
{
    if (c.z < r + znear) return false;
    vec3 cr{ c.x * r, c.y * r, c.z * r };
    float czr2 = c.z * c.z - r * r;
    float vx = std::sqrt(c.x * c.x + czr2);
    float minx = (vx * c.x - cr.z) / (vx * c.z + cr.x);
    float maxx = (vx * c.x + cr.z) / (vx * c.z - cr.x);
    float vy = std::sqrt(c.y * c.y + czr2);
    float miny = (vy * c.y - cr.z) / (vy * c.z + cr.y);
    float maxy = (vy * c.y + cr.z) / (vy * c.z - cr.y);
    // aabb = vec4(minx*P00, miny*P11, maxx*P00, maxy*P11); aabb = aabb.xwzy * vec4(.5,-.5,.5,-.5) + vec4(.5)
    float ax = minx * P00, ay = miny * P11, az = maxx * P00, aw = maxy * P11;
    aabb.x = ax * 0.5f + 0.5f;
    aabb.y = aw * -0.5f + 0.5f;
    aabb.z = az * 0.5f + 0.5f;
    aabb.w = ay * -0.5f + 0.5f;
    return true;
}

// floor(log2(x)) for x > 0, exact, from the exponent field (denormals handled).
inline int ilog2_floor_pos(float x)
{
    uint32_t b; std::memcpy(&b, &x, 4);
    int e = int((b >> 23) & 0xFF);
    if (e != 0) return e - 127;
    uint32_t m = b & 0x7FFFFFu;          // denormal, m != 0 because x > 0
    int hi = 31 - __builtin_clz(m);
    return hi - 149;
}

// ---- depth pyramid (linear mip chain) ---------------------------------------------
struct Pyramid { const float* data; uint32_t width, height, mips; uint32_t offset[16]; };

inline uint32_t maxu(uint32_t a, uint32_t b) { return a > b ? a : b; }
inline float minf(float a, float b) { return b < a ? b : a; }

// clamp a float texel index to [0, n-1]; NaN -> 0
inline uint32_t clamp_index(float f, uint32_t n)
{
    if (!(f >= 0.0f)) return 0u;
    float hi = float(n - 1);
    if (f >= hi) return n - 1;
    return uint32_t(f);
}

// The reference's Hi-Z sampler (BlitzenVulkan/vulkanResources.cpp:51-55 + CreateSampler :394-429):
// filter LINEAR + VK_SAMPLER_REDUCTION_MODE_MIN, mipmap NEAREST, CLAMP_TO_EDGE, normalized coords.
// Vulkan spec (texel filtering + VK_EXT_sampler_filter_minmax): u = s*W - 0.5, i0 = floor(u), i1 = i0 + 1,
// weights (1-frac, frac); the reduction takes the MIN over the texels of the 2x2 footprint that have
// non-zero weight; texel indices are clamped to the edge.
inline float sample_min_linear(const float* img, uint32_t W, uint32_t H, float s, float t)
{
    float u = s * float(W) - 0.5f;
    float v = t * float(H) - 0.5f;
    float fu = std::floor(u), fv = std::floor(v);
    float au = u - fu, av = v - fv;
    uint32_t i0 = clamp_index(fu, W), i1 = clamp_index(fu + 1.0f, W);
    uint32_t j0 = clamp_index(fv, H), j1 = clamp_index(fv + 1.0f, H);
    bool useI1 = !(au == 0.0f), useJ1 = !(av == 0.0f);
    float d = img[size_t(j0) * W + i0];
    if (useI1) d = minf(d, img[size_t(j0) * W + i1]);
    if (useJ1) {
        d = minf(d, img[size_t(j1) * W + i0]);
        if (useI1) d = minf(d, img[size_t(j1) * W + i1]);
    }
    return d;
}

// CullingShaderData.glsl:60-73 (OcclusionCullingPassed), Vulkan variant.
// textureLod(depthPyramid, uv, level): mip = clamp(level, 0, mips-1) (sampler minLod 0 / maxLod 16, mip NEAREST,
// integer-valued level); level = floor(log2(max(w,h))) -- -inf (max<=0) and NaN clamp to mip 0, +inf to the last mip.
inline bool OcclusionCullingPassedVK(vec4 aabb, const Pyramid& p, float pyramidWidth, float pyramidHeight, vec3 center, float radius, float zNear)
{
    float width = (aabb.z - aabb.x) * pyramidWidth;
    float height = (aabb.w - aabb.y) * pyramidHeight;
    float m; if (width < height) m = height; else m = width;     // GLSL max(x,y) = y if x < y else x
    int level;
    if (!(m > 0.0f)) level = 0;
    else if (std::isinf(m)) level = int(p.mips) - 1;
    else level = ilog2_floor_pos(m);
    if (level < 0) level = 0;
    if (level > int(p.mips) - 1) level = int(p.mips) - 1;
    uint32_t W = maxu(1u, p.width >> level), H = maxu(1u, p.height >> level);
    float s = (aabb.x + aabb.z) * 0.5f, t = (aabb.y + aabb.w) * 0.5f;
    float depth = sample_min_linear(p.data + p.offset[level], W, H, s, t);
    float depthSphere = zNear / (center.z - radius);
    return depthSphere > depth;
}

// float -> uint as NVIDIA hardware does it (F2I.U32 / __float2uint_rz): truncate, saturate, NaN -> 0.
// HLSL leaves out-of-range conversions undefined; this is the documented choice.
inline uint32_t f2u_sat(float f)
{
    if (!(f > 0.0f)) return 0u;
    if (f >= 4294967296.0f) return 0xFFFFFFFFu;
    return uint32_t(f);
}

// hlslMath.hlsl:58-78 (OcclusionCheck), D3D12 variant: ONE point texel via Texture2D.Load;
// out-of-range mip or coordinate loads return 0 (D3D out-of-bounds read rule) => visible.
inline bool OcclusionCheckDX(vec4 aabb, const Pyramid& p, uint32_t pyramidWidth, uint32_t pyramidHeight, vec3 center, float radius, float zNear)
{
    float width = (aabb.z - aabb.x) * float(pyramidWidth);
    float height = (aabb.w - aabb.y) * float(pyramidHeight);
    float m; if (width < height) m = height; else m = width;
    uint32_t level;
    if (!(m > 0.0f)) level = 0u;
    else if (std::isinf(m)) level = 0xFFFFFFFFu;
    else { int e = ilog2_floor_pos(m); level = e < 0 ? 0u : uint32_t(e); }
    uint32_t sh = level & 31u;                     // DXIL shift: low 5 bits
    uint32_t mipWidth = maxu(1u, pyramidWidth >> sh);
    uint32_t mipHeight = maxu(1u, pyramidHeight >> sh);
    float fx = float(mipWidth) * 0.5f, fy = float(mipHeight) * 0.5f;
    aabb.x *= fx; aabb.y *= fy; aabb.z *= fx; aabb.w *= fy;
    uint32_t tx = f2u_sat(aabb.x + aabb.z), ty = f2u_sat(aabb.y + aabb.w);
    float depth = 0.0f;
    if (level < p.mips) {
        uint32_t W = maxu(1u, p.width >> level), H = maxu(1u, p.height >> level);
        if (tx < W && ty < H) depth = p.data[p.offset[level] + size_t(ty) * W + tx];
    }
    float depthSphere = zNear / (center.z - radius);
    return depthSphere > depth;
}

// CullingShaderData.glsl:103-116 (returns the RELATIVE index; callers add lodOffset) and
// HlslShaders/Headers/cullBuffers.hlsl:42-55 (returns lodOffset + index).
inline uint32_t LODSelection(vec3 center, float radius, float scale, float lodTarget, uint32_t lodOffset, uint32_t lodCount, const LodData* lods)
{
    float len = std::sqrt((center.x * center.x + center.y * center.y) + center.z * center.z);
    float d = len - radius;
    float distance = d < 0.0f ? 0.0f : d;           // max(x, 0) = (x < 0) ? 0 : x
    float threshold = distance * lodTarget / scale;
    uint32_t lodIndex = 0;
    for (uint32_t i = 1; i < lodCount; ++i)
        if (lods[lodOffset + i].error < threshold) lodIndex = i;
    return lodIndex;
}

enum Pass { PASS_FRUSTUM = 0, PASS_EARLY = 1, PASS_LATE = 2, PASS_TEMPORAL = 3 };
enum Hiz { HIZ_VK = 0, HIZ_DX = 1 };
enum Flags { FLAG_ONPC_LOD_QUIRK = 1 };

struct Scene {
    const RenderObject* objs; uint32_t nObj;
    const MeshTransform* xf; const PrimitiveSurface* surf; const LodData* lods;
    const Cluster* clusters;
    uint32_t objectIdBase;
};

struct ObjResult { bool visible; bool emit; uint32_t lodIndex; vec3 center; float radius; float scale; };

// One invocation of a draw-cull main():
//  PASS_FRUSTUM : VulkanShaders/TransparentDrawCull.comp.glsl:13-65, HlslShaders/CS/drawCull.cs.hlsl:10-52,
//                 VulkanShaders/OnpcDrawCull.comp.glsl:11-52 (with FLAG_ONPC_LOD_QUIRK: relative LOD index used as absolute, :38-40)
//  PASS_EARLY   : VulkanShaders/InitialDrawCull.comp.glsl:12-59, HlslShaders/CS/drawOccFirst.cs.hlsl:12-59
//  PASS_LATE    : VulkanShaders/LateDrawCull.comp.glsl:13-72, HlslShaders/CS/drawOccLate.cs.hlsl:13-68
//  PASS_TEMPORAL: HlslShaders/CS/drawOccTemporal.hlsl:13-65 (frustum + Hi-Z, no visibility buffer)
inline ObjResult eval_object(const Scene& S, const ViewData& V, const Pyramid* pyr, int pass, int hiz, uint32_t flags,
                             uint32_t i, uint32_t* vis)
{
    ObjResult R{}; R.visible = false; R.emit = false; R.lodIndex = 0;
    if (pass == PASS_EARLY && vis[i] == 0) return R;
    RenderObject obj = S.objs[i];
    const MeshTransform& T = S.xf[obj.transformId];
    const PrimitiveSurface& sf = S.surf[obj.surfaceId];
    vec3 center; float radius;
    bool visible = IsObjectInsideViewFrustum(center, radius, vec3{ sf.center[0], sf.center[1], sf.center[2] }, sf.radius,
        T.scale, vec3{ T.pos[0], T.pos[1], T.pos[2] }, vec4{ T.q[0], T.q[1], T.q[2], T.q[3] }, V.view,
        V.frustumRight, V.frustumLeft, V.frustumTop, V.frustumBottom, V.zNear, V.zFar);
    if ((pass == PASS_LATE || pass == PASS_TEMPORAL) && visible) {
        vec4 aabb;
        if (projectSphere(center, radius, V.zNear, V.proj0, V.proj5, aabb)) {
            bool occ = hiz == HIZ_VK
                ? OcclusionCullingPassedVK(aabb, *pyr, V.pyramidWidth, V.pyramidHeight, center, radius, V.zNear)
                : OcclusionCheckDX(aabb, *pyr, f2u_sat(V.pyramidWidth), f2u_sat(V.pyramidHeight), center, radius, V.zNear);
            visible = visible && occ;
        }
    }
    bool emit = visible;
    if (pass == PASS_LATE) emit = visible && vis[i] == 0;
    if (emit) {
        uint32_t rel = LODSelection(center, radius, T.scale, V.lodTarget, sf.lodOffset, sf.lodCount, S.lods);
        R.lodIndex = (flags & FLAG_ONPC_LOD_QUIRK) ? rel : rel + sf.lodOffset;
    }
    if (pass == PASS_LATE) vis[i] = visible ? 1u : 0u;
    R.visible = visible; R.emit = emit; R.center = center; R.radius = radius; R.scale = T.scale;
    return R;
}

template <class F>
void parallel_ranges(uint64_t n, int threads, F&& f)
{
    if (threads <= 1 || n < 1024) { f(0, 0, n); return; }
    std::vector<std::thread> th;
    uint64_t per = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        uint64_t a = std::min<uint64_t>(n, per * t), b = std::min<uint64_t>(n, a + per);
        th.emplace_back([=, &f] { f(t, a, b); });
    }
    for (auto& x : th) x.join();
}

inline void write_record(uint32_t* dst, int recWords, uint32_t objectId, uint32_t indexCount, uint32_t instanceCount, uint32_t firstIndex)
{
    // IndirectDraw {objectId, indexCount, instanceCount, firstIndex, vertexOffset, firstInstance} (ShaderBuffers.glsl:73-84)
    // DrawCmd adds padding0/padding1 (cullBuffers.hlsl:1-15); the shaders never write the padding: the oracle zeroes it.
    dst[0] = objectId; dst[1] = indexCount; dst[2] = instanceCount; dst[3] = firstIndex; dst[4] = 0; dst[5] = 0;
    if (recWords == 8) { dst[6] = 0; dst[7] = 0; }
}

} // namespace

extern "C" {

struct oracle_scene {
    const void* objs; uint32_t nObj;
    const void* transforms; const void* surfaces; const void* lods; const void* clusters;
    uint32_t objectIdBase; uint32_t pad;
};
struct oracle_pyramid { const float* data; uint32_t width, height, mips; uint32_t offset[16]; };

int oracle_hardware_threads() { unsigned n = std::thread::hardware_concurrency(); return n ? int(n) : 1; }

// Pyramid geometry.  variant 0 (Vulkan): extent PreviousPow2(W) x PreviousPow2(H) (BlitzenMathLibrary/blitML.h:54-62,
// BlitzenVulkan/vulkanResources.cpp:93-94).  variant 1 (D3D12): max(1,W>>1) x max(1,H>>1) (BlitzenDX12/dx12RNDResources.cpp:103-106).
// mips = GetDepthPyramidMipLevels (blitML.h:95-107).  Returns total texel count; fills out[0..2] = w,h,mips and out[3..] offsets.
uint64_t oracle_pyramid_layout(uint32_t depthW, uint32_t depthH, int variant, uint32_t* out /*3+16*/)
{
    uint32_t w, h;
    if (variant == 0) {
        auto prev = [](uint32_t v) { uint32_t r = 1; while (r * 2 < v) r *= 2; return r; };
        w = prev(depthW); h = prev(depthH);
    } else { w = maxu(1u, depthW >> 1); h = maxu(1u, depthH >> 1); }
    uint32_t mips = 0; { uint32_t a = w, b = h; while (a > 1 || b > 1) { ++mips; a /= 2; b /= 2; } }
    out[0] = w; out[1] = h; out[2] = mips;
    uint64_t off = 0;
    for (uint32_t i = 0; i < 16; ++i) {
        out[3 + i] = uint32_t(off);
        if (i < mips) off += uint64_t(maxu(1u, w >> i)) * maxu(1u, h >> i);
    }
    return off;
}

// Vulkan pyramid: VulkanShaders/DepthPyramidGeneration.comp.glsl:13-20 run once per mip by
// BlitzenVulkan/vulkanDraw.cpp:579-614: mip 0 samples the depth attachment, mip i samples mip i-1, always at
// (pos + 0.5) / levelSize through the MIN sampler.
// D3D12 pyramid: HlslShaders/CS/depthPyramid.cs.hlsl:13-27 run per mip by BlitzenDX12/dx12Draw.cpp:246-277:
// min of the four Load()s at 2*xy + {0,1}^2 of the previous level (depth target for mip 0); OOB loads give 0.
int oracle_build_pyramid(const float* depth, uint32_t depthW, uint32_t depthH, int variant, float* out, int threads)
{
    uint32_t L[19]; oracle_pyramid_layout(depthW, depthH, variant, L);
    uint32_t w = L[0], h = L[1], mips = L[2];
    const float* src = depth; uint32_t sw = depthW, sh = depthH;
    for (uint32_t i = 0; i < mips; ++i) {
        uint32_t lw = maxu(1u, w >> i), lh = maxu(1u, h >> i);
        float* dst = out + L[3 + i];
        parallel_ranges(lh, threads, [&](int, uint64_t a, uint64_t b) {
            for (uint32_t y = uint32_t(a); y < uint32_t(b); ++y)
                for (uint32_t x = 0; x < lw; ++x) {
                    float d;
                    if (variant == 0) {
                        float s = (float(x) + 0.5f) / float(lw), t = (float(y) + 0.5f) / float(lh);
                        d = sample_min_linear(src, sw, sh, s, t);
                    } else {
                        auto ld = [&](uint32_t xx, uint32_t yy) { return (xx < sw && yy < sh) ? src[size_t(yy) * sw + xx] : 0.0f; };
                        float r = ld(2 * x, 2 * y), g = ld(2 * x + 1, 2 * y), bb = ld(2 * x, 2 * y + 1), aa = ld(2 * x + 1, 2 * y + 1);
                        d = minf(r, minf(g, minf(bb, aa)));   // min(r, min(g, min(b, a)))
                    }
                    dst[size_t(y) * lw + x] = d;
                }
        });
        src = dst; sw = lw; sh = lh;
    }
    return 0;
}

// Draw-cull passes.  vis may be null for PASS_FRUSTUM / PASS_TEMPORAL.  recWords = 6 (Vulkan IndirectDraw) or 8 (D3D12 DrawCmd).
// Writes min(total, capacity) records in ascending objectId; *outTotal = total survivors.  Returns 0.
int oracle_cull(const oracle_scene* sc, const void* view, const oracle_pyramid* pyr, int pass, int hiz, uint32_t flags,
                uint32_t* vis, uint32_t* outRecords, uint32_t recWords, uint64_t capacity, uint32_t* outWritten, uint32_t* outTotal, int threads)
{
    Scene S{ (const RenderObject*)sc->objs, sc->nObj, (const MeshTransform*)sc->transforms, (const PrimitiveSurface*)sc->surfaces,
             (const LodData*)sc->lods, (const Cluster*)sc->clusters, sc->objectIdBase };
    ViewData V; std::memcpy(&V, view, sizeof(V));
    Pyramid P{}; if (pyr) { P.data = pyr->data; P.width = pyr->width; P.height = pyr->height; P.mips = pyr->mips; std::memcpy(P.offset, pyr->offset, sizeof(P.offset)); }
    if (threads < 1) threads = 1;
    std::vector<std::vector<uint32_t>> local(threads);
    parallel_ranges(S.nObj, threads, [&](int t, uint64_t a, uint64_t b) {
        auto& out = local[t];
        for (uint64_t i = a; i < b; ++i) {
            ObjResult R = eval_object(S, V, pyr ? &P : nullptr, pass, hiz, flags, uint32_t(i), vis);
            if (R.emit) {
                const LodData& lod = S.lods[R.lodIndex];
                size_t o = out.size(); out.resize(o + recWords);
                write_record(out.data() + o, int(recWords), S.objectIdBase + uint32_t(i), lod.indexCount, 1u, lod.firstIndex);
            }
        }
    });
    uint64_t total = 0, written = 0;
    for (auto& v : local) {
        uint64_t n = v.size() / recWords;
        uint64_t room = capacity > written ? capacity - written : 0;
        uint64_t take = std::min(n, room);
        if (take) std::memcpy(outRecords + written * recWords, v.data(), take * recWords * 4);
        written += take; total += n;
    }
    *outWritten = uint32_t(written); *outTotal = uint32_t(total);
    return 0;
}

// D3D12 indirect instancing: HlslShaders/CS/drawInstCountReset.cs.hlsl:9-19 + drawInstCull.cs.hlsl:12-43 + drawInstCmd.cs.hlsl:9-39
// (host: BlitzenDX12/dx12Draw.cpp:340-413).  instIndices[instanceOffset[lod] + k] = objId, k ascending objId inside each bucket;
// one DrawCmd {objId = instanceOffset, indexCount, instCount = count, indexOffset, 0, 0} per non-empty LOD, ascending LOD id.
// bucketCapacity[l] bounds k (the reference is unguarded, Ce_MaxInstanceCountPerLOD = 100000): overflowing ids are dropped
// from instIndices but still counted in outCounts[l]; instCount = min(count, capacity).
int oracle_cull_instanced(const oracle_scene* sc, const void* view, const void* lodInstances, uint32_t nLods, const uint32_t* bucketCapacity,
                          uint32_t* instIndices, uint32_t* outCounts, uint32_t* outCmds /*8 words each*/, uint32_t* outCmdCount, int threads)
{
    Scene S{ (const RenderObject*)sc->objs, sc->nObj, (const MeshTransform*)sc->transforms, (const PrimitiveSurface*)sc->surfaces,
             (const LodData*)sc->lods, (const Cluster*)sc->clusters, sc->objectIdBase };
    ViewData V; std::memcpy(&V, view, sizeof(V));
    const LodInstanceCounter* LI = (const LodInstanceCounter*)lodInstances;
    if (threads < 1) threads = 1;
    std::vector<std::vector<uint32_t>> sel(threads);   // (lod, objId) pairs per thread range
    parallel_ranges(S.nObj, threads, [&](int t, uint64_t a, uint64_t b) {
        for (uint64_t i = a; i < b; ++i) {
            ObjResult R = eval_object(S, V, nullptr, PASS_FRUSTUM, 0, 0, uint32_t(i), nullptr);
            if (R.emit) { sel[t].push_back(R.lodIndex); sel[t].push_back(S.objectIdBase + uint32_t(i)); }
        }
    });
    for (uint32_t l = 0; l < nLods; ++l) outCounts[l] = 0;
    for (auto& v : sel)
        for (size_t k = 0; k < v.size(); k += 2) {
            uint32_t l = v[k], id = v[k + 1];
            uint32_t slot = outCounts[l]++;
            if (slot < bucketCapacity[l]) instIndices[LI[l].instanceOffset + slot] = id;
        }
    uint32_t n = 0;
    for (uint32_t l = 0; l < nLods; ++l) {
        if (outCounts[l] == 0) continue;
        uint32_t c = std::min(outCounts[l], bucketCapacity[l]);
        uint32_t* d = outCmds + size_t(n) * 8;
        d[0] = LI[l].instanceOffset; d[1] = S.lods[l].indexCount; d[2] = c; d[3] = S.lods[l].firstIndex; d[4] = 0; d[5] = 0; d[6] = 0; d[7] = 0;
        ++n;
    }
    *outCmdCount = n;
    return 0;
}

// VulkanShaders/PreClusterDrawCull.comp.glsl:13-47: frustum + LOD; each visible object appends lod.clusterCount records
// {objectId, lodIndex (absolute), clusterId = lod.clusterOffset + k}.  Ascending objectId, then k.
int oracle_cluster_expand(const oracle_scene* sc, const void* view, uint32_t* outRecords /*3 words*/, uint64_t capacity,
                          uint32_t* outWritten, uint32_t* outTotal, int threads)
{
    Scene S{ (const RenderObject*)sc->objs, sc->nObj, (const MeshTransform*)sc->transforms, (const PrimitiveSurface*)sc->surfaces,
             (const LodData*)sc->lods, (const Cluster*)sc->clusters, sc->objectIdBase };
    ViewData V; std::memcpy(&V, view, sizeof(V));
    if (threads < 1) threads = 1;
    std::vector<std::vector<uint32_t>> local(threads);
    parallel_ranges(S.nObj, threads, [&](int t, uint64_t a, uint64_t b) {
        auto& out = local[t];
        for (uint64_t i = a; i < b; ++i) {
            ObjResult R = eval_object(S, V, nullptr, PASS_FRUSTUM, 0, 0, uint32_t(i), nullptr);
            if (!R.emit) continue;
            const LodData& lod = S.lods[R.lodIndex];
            for (uint32_t k = 0; k < lod.clusterCount; ++k) {
                out.push_back(S.objectIdBase + uint32_t(i)); out.push_back(R.lodIndex); out.push_back(lod.clusterOffset + k);
            }
        }
    });
    uint64_t total = 0, written = 0;
    for (auto& v : local) {
        uint64_t n = v.size() / 3;
        uint64_t room = capacity > written ? capacity - written : 0;
        uint64_t take = std::min(n, room);
        if (take) std::memcpy(outRecords + written * 3, v.data(), take * 12);
        written += take; total += n;
    }
    *outWritten = uint32_t(written); *outTotal = uint32_t(total);
    return 0;
}

// VulkanShaders/InitialClusterCull.comp.glsl:12-55 (== TransparentClusterCull.comp.glsl): per dispatch record append
// {data.objectId, clusters[id].triangleCount*3, 1, clusters[id].dataOffset, 0, 0}.  mode 0 = passthrough (reference-exact:
// the cone test is commented out, :24-42, and no sphere test exists).  mode 1 = "sphere" (NOT in the reference; BASELINE
// config 4 asks for it): the cluster's bounding sphere goes through IsObjectInsideViewFrustum / projectSphere / Hi-Z under the
// owning object's transform, exactly as an object's surface sphere would.  hiz = -1 disables the Hi-Z part of mode 1.
int oracle_cluster_cull(const oracle_scene* sc, const void* view, const oracle_pyramid* pyr, const uint32_t* dispatch, uint64_t nRecords,
                        int mode, int hiz, uint32_t* outRecords, uint32_t recWords, uint64_t capacity, uint32_t* outWritten, uint32_t* outTotal, int threads)
{
    Scene S{ (const RenderObject*)sc->objs, sc->nObj, (const MeshTransform*)sc->transforms, (const PrimitiveSurface*)sc->surfaces,
             (const LodData*)sc->lods, (const Cluster*)sc->clusters, sc->objectIdBase };
    ViewData V; std::memcpy(&V, view, sizeof(V));
    Pyramid P{}; if (pyr) { P.data = pyr->data; P.width = pyr->width; P.height = pyr->height; P.mips = pyr->mips; std::memcpy(P.offset, pyr->offset, sizeof(P.offset)); }
    if (threads < 1) threads = 1;
    std::vector<std::vector<uint32_t>> local(threads);
    parallel_ranges(nRecords, threads, [&](int t, uint64_t a, uint64_t b) {
        auto& out = local[t];
        for (uint64_t r = a; r < b; ++r) {
            ClusterDispatchData d{ dispatch[r * 3], dispatch[r * 3 + 1], dispatch[r * 3 + 2] };
            const Cluster& c = S.clusters[d.clusterId];
            bool visible = true;
            if (mode == 1) {
                RenderObject obj = S.objs[d.objectId - S.objectIdBase];
                const MeshTransform& T = S.xf[obj.transformId];
                vec3 center; float radius;
                visible = IsObjectInsideViewFrustum(center, radius, vec3{ c.center[0], c.center[1], c.center[2] }, c.radius,
                    T.scale, vec3{ T.pos[0], T.pos[1], T.pos[2] }, vec4{ T.q[0], T.q[1], T.q[2], T.q[3] }, V.view,
                    V.frustumRight, V.frustumLeft, V.frustumTop, V.frustumBottom, V.zNear, V.zFar);
                if (visible && hiz >= 0 && pyr) {
                    vec4 aabb;
                    if (projectSphere(center, radius, V.zNear, V.proj0, V.proj5, aabb)) {
                        bool occ = hiz == HIZ_VK
                            ? OcclusionCullingPassedVK(aabb, P, V.pyramidWidth, V.pyramidHeight, center, radius, V.zNear)
                            : OcclusionCheckDX(aabb, P, f2u_sat(V.pyramidWidth), f2u_sat(V.pyramidHeight), center, radius, V.zNear);
                        visible = visible && occ;
                    }
                }
            }
            if (!visible) continue;
            size_t o = out.size(); out.resize(o + recWords);
            write_record(out.data() + o, int(recWords), d.objectId, uint32_t(c.triangleCount) * 3u, 1u, c.dataOffset);
        }
    });
    uint64_t total = 0, written = 0;
    for (auto& v : local) {
        uint64_t n = v.size() / recWords;
        uint64_t room = capacity > written ? capacity - written : 0;
        uint64_t take = std::min(n, room);
        if (take) std::memcpy(outRecords + written * recWords, v.data(), take * recWords * 4);
        written += take; total += n;
    }
    *outWritten = uint32_t(written); *outTotal = uint32_t(total);
    return 0;
}

// Near-boundary census (north_star: "objects whose bounds lie within a stated epsilon of a frustum plane or Hi-Z texel
// boundary ... are counted and reported").  A real GLSL/HLSL GPU may contract to FMA and quantises sampler weights to
// 8 bits, so only these objects can legitimately differ from the oracle on real graphics hardware.
//   out[0] = objects with any frustum comparison within ulpTol ulps of the compared magnitude
//   out[1] = frustum-visible objects whose Hi-Z sample coordinate is within texelTol of a texel-centre line (weight -> 0)
//   out[2] = frustum-visible objects whose max(w,h) is within ulpTol ulps of a power of two (mip level boundary)
//   out[3] = frustum-visible objects whose depthSphere is within ulpTol ulps of the sampled depth
int oracle_boundary_census(const oracle_scene* sc, const void* view, const oracle_pyramid* pyr, int hiz, float ulpTol, float texelTol, uint64_t* out, int threads)
{
    Scene S{ (const RenderObject*)sc->objs, sc->nObj, (const MeshTransform*)sc->transforms, (const PrimitiveSurface*)sc->surfaces,
             (const LodData*)sc->lods, (const Cluster*)sc->clusters, sc->objectIdBase };
    ViewData V; std::memcpy(&V, view, sizeof(V));
    Pyramid P{}; if (pyr) { P.data = pyr->data; P.width = pyr->width; P.height = pyr->height; P.mips = pyr->mips; std::memcpy(P.offset, pyr->offset, sizeof(P.offset)); }
    const float eps = 1.1920929e-7f * ulpTol;
    auto near = [&](float a, float b) { float m = std::fmax(std::fabs(a), std::fabs(b)); return std::fabs(a - b) <= eps * m; };
    std::vector<std::array<uint64_t, 4>> part(size_t(threads > 1 ? threads : 1), std::array<uint64_t, 4>{ 0, 0, 0, 0 });
    parallel_ranges(S.nObj, threads, [&](int t, uint64_t ra, uint64_t rb) {
    uint64_t* out = part[size_t(t)].data();
    for (uint64_t i = ra; i < rb; ++i) {
        RenderObject obj = S.objs[i];
        const MeshTransform& T = S.xf[obj.transformId];
        const PrimitiveSurface& sf = S.surf[obj.surfaceId];
        vec3 c; float r;
        bool vis = IsObjectInsideViewFrustum(c, r, vec3{ sf.center[0], sf.center[1], sf.center[2] }, sf.radius, T.scale,
            vec3{ T.pos[0], T.pos[1], T.pos[2] }, vec4{ T.q[0], T.q[1], T.q[2], T.q[3] }, V.view,
            V.frustumRight, V.frustumLeft, V.frustumTop, V.frustumBottom, V.zNear, V.zFar);
        bool nb = near(c.z * V.frustumLeft - std::fabs(c.x) * V.frustumRight, -r)
               || near(c.z * V.frustumBottom - std::fabs(c.y) * V.frustumTop, -r)
               || near(c.z + r, V.zNear) || near(c.z - r, V.zFar);
        if (nb) out[0]++;
        if (!vis || !pyr) continue;
        vec4 aabb;
        if (!projectSphere(c, r, V.zNear, V.proj0, V.proj5, aabb)) continue;
        float w = (aabb.z - aabb.x) * V.pyramidWidth, h = (aabb.w - aabb.y) * V.pyramidHeight;
        float m = w < h ? h : w;
        if (!(m > 0.0f) || std::isinf(m)) continue;
        int e = ilog2_floor_pos(m);
        float lo = std::ldexp(1.0f, e), hi2 = std::ldexp(1.0f, e + 1);
        if (near(m, lo) || near(m, hi2)) out[2]++;
        int level = std::min(std::max(e, 0), int(P.mips) - 1);
        uint32_t W = maxu(1u, P.width >> level), H = maxu(1u, P.height >> level);
        float depth;
        if (hiz == HIZ_VK) {
            float s = (aabb.x + aabb.z) * 0.5f, t = (aabb.y + aabb.w) * 0.5f;
            float u = s * float(W) - 0.5f, v = t * float(H) - 0.5f;
            float au = u - std::floor(u), av = v - std::floor(v);
            if (au < texelTol || au > 1.0f - texelTol || av < texelTol || av > 1.0f - texelTol) out[1]++;
            depth = sample_min_linear(P.data + P.offset[level], W, H, s, t);
        } else {
            float fx = float(W) * 0.5f, fy = float(H) * 0.5f;
            float x = aabb.x * fx + aabb.z * fx, y = aabb.y * fy + aabb.w * fy;
            float ax = x - std::floor(x), ay = y - std::floor(y);
            if (ax < texelTol || ax > 1.0f - texelTol || ay < texelTol || ay > 1.0f - texelTol) out[1]++;
            uint32_t tx = f2u_sat(x), ty = f2u_sat(y);
            depth = (e >= 0 && uint32_t(e) < P.mips && tx < W && ty < H) ? P.data[P.offset[level] + size_t(ty) * W + tx] : 0.0f;
        }
        if (near(V.zNear / (c.z - r), depth)) out[3]++;
    }
    });
    out[0] = out[1] = out[2] = out[3] = 0;
    for (auto& q : part) for (int k = 0; k < 4; ++k) out[k] += q[size_t(k)];
    return 0;
}

// Software depth (product: blitzen_b200/csrc/raster_depth.cu).  No reference shader does this -- it stands in for the
// rasteriser between the two cull phases (frame order BlitzenVulkan/vulkanDraw.cpp:1015-1036; clear value 0, reverse-Z:
// BlitzenVulkan/vulkanResources.cpp:70-71) -- so this is the DEFINITION the CUDA kernel is compared with, built from the reference's own
// functions: view-space sphere as in IsObjectInsideViewFrustum, screen box from projectSphere, depth of the sphere's far side
// zNear / (c.z + r), merged with max.  Pixel (x, y) is covered when floor(aabb.x * W) <= x < ceil(aabb.z * W), same in y, clamped.
static inline uint32_t pix_floor(float f, uint32_t n) { if (!(f >= 0.0f)) return 0u; if (f >= float(n)) return n; return uint32_t(std::floor(f)); }
static inline uint32_t pix_ceil(float f, uint32_t n) { if (!(f >= 0.0f)) return 0u; if (f >= float(n)) return n; return uint32_t(std::ceil(f)); }

int oracle_raster_depth(const oracle_scene* sc, const void* view, const uint32_t* records, uint64_t nRecords, uint32_t recWords,
                        uint32_t W, uint32_t H, float* outDepth)
{
    Scene S{ (const RenderObject*)sc->objs, sc->nObj, (const MeshTransform*)sc->transforms, (const PrimitiveSurface*)sc->surfaces,
             (const LodData*)sc->lods, (const Cluster*)sc->clusters, sc->objectIdBase };
    ViewData V; std::memcpy(&V, view, sizeof(V));
    for (size_t i = 0; i < size_t(W) * H; ++i) outDepth[i] = 0.0f;
    for (uint64_t k = 0; k < nRecords; ++k) {
        const uint32_t local = records[k * recWords] - S.objectIdBase;
        if (local >= S.nObj) continue;
        RenderObject obj = S.objs[local];
        const MeshTransform& T = S.xf[obj.transformId];
        const PrimitiveSurface& sf = S.surf[obj.surfaceId];
        vec3 c; float r;
        IsObjectInsideViewFrustum(c, r, vec3{ sf.center[0], sf.center[1], sf.center[2] }, sf.radius, T.scale,
            vec3{ T.pos[0], T.pos[1], T.pos[2] }, vec4{ T.q[0], T.q[1], T.q[2], T.q[3] }, V.view,
            V.frustumRight, V.frustumLeft, V.frustumTop, V.frustumBottom, V.zNear, V.zFar);
        vec4 aabb;
        if (!projectSphere(c, r, V.zNear, V.proj0, V.proj5, aabb)) continue;
        const uint32_t x0 = pix_floor(aabb.x * float(W), W), x1 = pix_ceil(aabb.z * float(W), W);
        const uint32_t y0 = pix_floor(aabb.y * float(H), H), y1 = pix_ceil(aabb.w * float(H), H);
        const float d = V.zNear / (c.z + r);
        if (!(d > 0.0f)) continue;
        for (uint32_t y = y0; y < y1; ++y)
            for (uint32_t x = x0; x < x1; ++x) { float& t = outDepth[size_t(y) * W + x]; if (d > t) t = d; }
    }
    return 0;
}

// Per-object debug probe used by the known-answer tests: runs the frustum / projectSphere / Hi-Z / LOD functions for one
// explicit sphere+transform and returns the intermediate values.
// out: [0]=visible(frustum) [1..3]=center [4]=radius [5]=projected(0/1) [6..9]=aabb [10]=hiz passed(0/1) [11]=lodIndex(relative)
int oracle_probe(const float* boundCenter, float boundRadius, const float* transform8, const void* view, const oracle_pyramid* pyr, int hiz,
                 const void* lods, uint32_t lodOffset, uint32_t lodCount, float* out)
{
    ViewData V; std::memcpy(&V, view, sizeof(V));
    vec3 c; float r;
    bool vis = IsObjectInsideViewFrustum(c, r, vec3{ boundCenter[0], boundCenter[1], boundCenter[2] }, boundRadius, transform8[3],
        vec3{ transform8[0], transform8[1], transform8[2] }, vec4{ transform8[4], transform8[5], transform8[6], transform8[7] }, V.view,
        V.frustumRight, V.frustumLeft, V.frustumTop, V.frustumBottom, V.zNear, V.zFar);
    out[0] = vis ? 1.f : 0.f; out[1] = c.x; out[2] = c.y; out[3] = c.z; out[4] = r;
    vec4 aabb{ 0, 0, 0, 0 };
    bool proj = projectSphere(c, r, V.zNear, V.proj0, V.proj5, aabb);
    out[5] = proj ? 1.f : 0.f; out[6] = aabb.x; out[7] = aabb.y; out[8] = aabb.z; out[9] = aabb.w;
    out[10] = 1.f;
    if (proj && pyr) {
        Pyramid P{}; P.data = pyr->data; P.width = pyr->width; P.height = pyr->height; P.mips = pyr->mips; std::memcpy(P.offset, pyr->offset, sizeof(P.offset));
        bool occ = hiz == HIZ_VK ? OcclusionCullingPassedVK(aabb, P, V.pyramidWidth, V.pyramidHeight, c, r, V.zNear)
                                 : OcclusionCheckDX(aabb, P, f2u_sat(V.pyramidWidth), f2u_sat(V.pyramidHeight), c, r, V.zNear);
        out[10] = occ ? 1.f : 0.f;
    }
    out[11] = lods ? float(LODSelection(c, r, transform8[3], V.lodTarget, lodOffset, lodCount, (const LodData*)lods)) : 0.f;
    return 0;
}

// exact floor(log2(x)) probe for the KATs
int oracle_ilog2_floor(float x) { return x > 0.0f && !std::isinf(x) ? ilog2_floor_pos(x) : -9999; }

} // extern "C"
