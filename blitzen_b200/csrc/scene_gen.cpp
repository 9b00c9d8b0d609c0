// Workload generator for the cull path: synthetic stress scenes shaped like the reference's
// RenderingStressTest / InstancingStressTest, plus synthetic reverse-Z depth images.
//
// It follows what the reference does at startup (all paths relative to /root/reference/src):
//   CreateSingleRender(bunny, 5)                         Renderer/Resources/RenderObject/blitzenRender.cpp:97-106
//   1000 x RandomizeTransform(100, 1) dynamic kittens    Renderer/Interface/blitzenRenderer.cpp:128-150
//   LoadGeometryStressTest(multiplier)                   Renderer/Resources/RenderObject/blitzenRender.cpp:126-166
//   RandomizeTransform                                   blitzenRender.cpp:108-115
//   QuatFromAngleAxis(axis, angle, normalize = 0)        BlitzenMathLibrary/blitML.h:559-570  (orientation is NOT unit length)
// In PRNG mode 0 (glibc rand(), default seed) the output is bit-identical to what the reference's own compiled
// frontend produces (oracle/_ref/refscene; checked by tests/test_scene_vs_reference.py whenever /root/reference
// is present).  g++ evaluates the arguments of the reference's `vec3(rand(), rand(), rand())` constructor calls
// right to left, so the draw order is: pos.z, pos.y, pos.x, angle, axis.z, axis.y, axis.x.
// PRNG mode 1 is a counter-based generator (same distribution, any object range can be generated independently),
// used for the scaled multi-GPU configurations where each rank builds only its own shard.
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct RenderObject { uint32_t transformId, surfaceId; };
struct MeshTransform { float pos[3]; float scale; float q[4]; };

struct Rng {
    int mode; uint64_t key; uint64_t ctr;
    inline int next()
    {
        if (mode == 0) return rand();
        // splitmix64 on (key, counter) -> 31 bits, like rand()
        uint64_t z = key + (ctr++) * 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z = z ^ (z >> 31);
        return int(z >> 33);
    }
};

constexpr float kPi = 3.14159265358979323846f;
constexpr float kDegToRad = kPi / 180.f;

inline float unit(Rng& r) { return float(r.next()) / float(RAND_MAX); }

inline void randomize_transform(Rng& r, MeshTransform& t, float multiplier, float scale)
{
    float pz = unit(r) * multiplier;
    float py = unit(r) * multiplier;
    float px = unit(r) * multiplier;
    t.pos[0] = px; t.pos[1] = py; t.pos[2] = pz;
    t.scale = scale;
    float angle = (unit(r) * 90.f) * kDegToRad;
    float az = unit(r) * 2 - 1;
    float ay = unit(r) * 2 - 1;
    float ax = unit(r) * 2 - 1;
    const float half = 0.5f * angle;
    float s = sinf(half), c = cosf(half);
    t.q[0] = s * ax; t.q[1] = s * ay; t.q[2] = s * az; t.q[3] = c;
}

} // namespace

extern "C" {

struct blz_scene_group { uint32_t surfaceId; float scale; uint64_t count; };

// Generates objects [first, first+count) of the scene
//   [prologue: 1 bunny (surface 0, scale 5, at the default camera position) + nDynamic dynamic objects (dynSurface, scale 1, cube 100)]
//   followed by the groups in order, positions uniform in a cube of side `multiplier`.
// Layout (blitRender.h:6-22): dynamic transforms occupy [0, nDynamic), static transforms start at staticOffset = nDynamic... the
// reference uses staticOffset = Ce_MaxDynamicObjectCount = 1000.  transformId of static object k (k-th static object overall) is
// staticOffset + k.  `objs`/`xforms` receive ONLY the requested range of objects; xforms[i] is the transform of object first+i
// and objs[i].transformId is the GLOBAL transform index (caller places xforms accordingly, see blitzen_b200/scene.py).
// PRNG mode 0 requires first == 0 (sequential glibc stream).
int blz_scene_generate(int prngMode, uint64_t seed, int prologue, uint32_t nDynamic, uint32_t dynSurface, uint32_t staticOffset,
                       const blz_scene_group* groups, uint32_t nGroups, float multiplier,
                       uint64_t first, uint64_t count, void* objsOut, void* xformsOut, int threads)
{
    RenderObject* objs = (RenderObject*)objsOut;
    MeshTransform* xf = (MeshTransform*)xformsOut;
    if (prngMode == 0 && first != 0) return -1;
    uint64_t total = prologue ? 1 + uint64_t(nDynamic) : 0;
    for (uint32_t g = 0; g < nGroups; ++g) total += groups[g].count;
    if (first + count > total) return -2;

    // glibc rand() keeps process-global state: restart the stream so that every call yields the reference's scene
    // (an unseeded process behaves as if srand(1) had been called)
    if (prngMode == 0) srand(seed ? unsigned(seed) : 1u);
    auto gen_range = [&](uint64_t a, uint64_t b) {
        Rng r{ prngMode, seed * 0xD1342543DE82EF95ull + 0x632BE59BD9B4E019ull, 0 };
        for (uint64_t i = a; i < b; ++i) {
            MeshTransform t; RenderObject o;
            uint64_t k = i;
            r.ctr = i * 8;
            if (prologue && k == 0) {
                // CreateSingleRender: pos = initial camera position (20,70,0), orientation = QuatFromAngleAxis(vec3(0), 0, 0) = (0,0,0,1)
                t.pos[0] = 20.f; t.pos[1] = 70.f; t.pos[2] = 0.f; t.scale = 5.f;
                t.q[0] = 0.f; t.q[1] = 0.f; t.q[2] = 0.f; t.q[3] = 1.f;
                o.transformId = staticOffset; o.surfaceId = 0;
            } else if (prologue && k <= nDynamic) {
                randomize_transform(r, t, 100.f, 1.f);
                o.transformId = uint32_t(k - 1); o.surfaceId = dynSurface;
            } else {
                uint64_t s = prologue ? k - 1 - nDynamic : k;      // index among stress objects
                uint64_t staticIdx = prologue ? s + 1 : s;          // the single bunny is static object 0
                uint32_t g = 0; uint64_t acc = 0;
                while (g + 1 < nGroups && s >= acc + groups[g].count) { acc += groups[g].count; ++g; }
                randomize_transform(r, t, multiplier, groups[g].scale);
                o.transformId = uint32_t(staticOffset + staticIdx); o.surfaceId = groups[g].surfaceId;
            }
            objs[i - first] = o; xf[i - first] = t;
        }
    };
    if (prngMode == 0 || threads <= 1) { gen_range(first, first + count); return 0; }
    std::vector<std::thread> th;
    uint64_t per = (count + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        uint64_t a = first + per * t, b = a + per;
        if (a > first + count) a = first + count;
        if (b > first + count) b = first + count;
        th.emplace_back([=] { gen_range(a, b); });
    }
    for (auto& x : th) x.join();
    return 0;
}

// Synthetic reverse-Z depth image: d = zNear / z_view, sky = 0 (depth cleared to 0,
// BlitzenVulkan/vulkanResources.cpp:70-71), nRects axis-aligned screen-space rectangles at view depth z in [zMin, zMax],
// nearer rectangles win (max d).  xorshift32 seeded with `seed` (default 0x00B1172E).
int blz_depth_generate(uint32_t W, uint32_t H, float zNear, uint32_t nRects, float zMin, float zMax, uint32_t seed, float* out)
{
    for (uint64_t i = 0; i < uint64_t(W) * H; ++i) out[i] = 0.0f;
    uint32_t s = seed ? seed : 0x00B1172Eu;
    auto next = [&]() { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; };
    auto uf = [&]() { return float(next() >> 8) * (1.0f / 16777216.0f); };
    for (uint32_t r = 0; r < nRects; ++r) {
        float cx = uf() * float(W), cy = uf() * float(H);
        float hw = (0.02f + 0.10f * uf()) * float(W), hh = (0.02f + 0.10f * uf()) * float(H);
        float z = zMin + (zMax - zMin) * uf();
        float d = zNear / z;
        int x0 = int(cx - hw), x1 = int(cx + hw), y0 = int(cy - hh), y1 = int(cy + hh);
        if (x0 < 0) x0 = 0; if (y0 < 0) y0 = 0; if (x1 > int(W)) x1 = int(W); if (y1 > int(H)) y1 = int(H);
        for (int y = y0; y < y1; ++y)
            for (int x = x0; x < x1; ++x) {
                float& p = out[size_t(y) * W + x];
                if (d > p) p = d;
            }
    }
    return 0;
}

} // extern "C"
