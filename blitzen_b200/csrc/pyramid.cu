// Single-pass Hi-Z pyramid build, replacing the reference's one-dispatch-plus-barrier-per-mip loops:
//   Vulkan: VulkanShaders/DepthPyramidGeneration.comp.glsl:13-20 driven by BlitzenVulkan/vulkanDraw.cpp:554-622
//           (mip 0 samples the full-resolution depth through the MIN sampler at (xy+.5)/size -- a non-2:1 footprint --
//            mip i samples mip i-1, which for the power-of-two pyramid extent is an exact 2x2 min)
//   D3D12:  HlslShaders/CS/depthPyramid.cs.hlsl:13-27 driven by BlitzenDX12/dx12Draw.cpp:224-288
//           (min of the four Load()s at 2*xy+{0,1}^2, out-of-range loads return 0)
// (paths relative to /root/reference/src/Renderer).
//
// One CTA owns a 32x32 tile of mip 0.  Its input box of the depth image is staged in shared memory by ONE TMA 2-D tile
// load (cp.async.bulk.tensor.2d, out-of-range elements zero-filled by the hardware), the CTA reduces up to six mips in
// shared memory and writes each once; the last CTA to finish (atomic ticket) reduces the remaining <= 64x32 tail.
// The depth image is read exactly once and every mip is written exactly once: 11.09 MB at 1080p, 44.4 MB at 4K.
#include "cull_kernels.cuh"
#include "cull_math.cuh"
#include <cuda.h>

namespace blz {

namespace {

constexpr int kPyrThreads = 256;
constexpr int kPyrTile = 32;
constexpr int kPyrMinBlocks = 8;                     // 8 x 148 = 1184 resident CTAs: the 1024 tiles of a 1080p build are ONE wave (at 6 per SM they were 1.15)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int32_t x, int32_t y, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

// first input column/row touched by output coordinate x0 of mip 0 (Vulkan footprint) -- mirrors sample_min_linear()
__host__ __device__ inline int32_t vk_first_texel(uint32_t x0, uint32_t outSize, uint32_t inSize)
{
    float s = (float(x0) + 0.5f) / float(outSize);
    float u = s * float(inSize) - 0.5f;
    float f = floorf(u);
    if (!(f >= 0.0f)) return 0;
    if (f >= float(inSize - 1)) return int32_t(inSize - 1);
    return int32_t(f);
}

__device__ __forceinline__ float ld_cg_f32(const float* p)
{
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// generic per-level step (tail levels): mirrors oracle_build_pyramid; `fetch(i, j)` returns texel (i, j) of the previous level
template <int VARIANT, class Fetch>
__device__ __forceinline__ float tail_texel_from(Fetch&& fetch, uint32_t sw, uint32_t sh, uint32_t x, uint32_t y, uint32_t lw, uint32_t lh)
{
    if (VARIANT == HIZ_VK) {
        float s = fdiv(fadd(float(x), 0.5f), float(lw)), t = fdiv(fadd(float(y), 0.5f), float(lh));
        return sample_min_linear(fetch, sw, sh, s, t);
    } else {
        auto ld = [&](uint32_t xx, uint32_t yy) { return (xx < sw && yy < sh) ? fetch(xx, yy) : 0.0f; };
        float r = ld(2 * x, 2 * y), g = ld(2 * x + 1, 2 * y), b = ld(2 * x, 2 * y + 1), a = ld(2 * x + 1, 2 * y + 1);
        return min_keep(r, min_keep(g, min_keep(b, a)));
    }
}
template <int VARIANT>
__device__ __forceinline__ float tail_texel(const float* src, uint32_t sw, uint32_t sh, uint32_t x, uint32_t y, uint32_t lw, uint32_t lh)
{
    return tail_texel_from<VARIANT>([&](uint32_t i, uint32_t j) { return ld_cg_f32(src + size_t(j) * sw + i); }, sw, sh, x, y, lw, lh);
}

template <int VARIANT, bool USE_TMA>
__global__ void __launch_bounds__(kPyrThreads, kPyrMinBlocks) pyramid_kernel(const __grid_constant__ PyramidBuildParams p, const __grid_constant__ CUtensorMap tmap)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_last;
    float* box = reinterpret_cast<float*>(smem_raw);                               // boxH x boxW input texels
    float* lvl = box + ((size_t(p.boxW) * p.boxH + 31u) & ~size_t(31));            // 32x32 + 16x16 + ... + 1

    const uint32_t tid = threadIdx.x;
    const uint32_t x0 = blockIdx.x * kPyrTile, y0 = blockIdx.y * kPyrTile;
    int32_t bx0, by0;
    if (VARIANT == HIZ_VK) { bx0 = vk_first_texel(x0, p.width, p.depthW); by0 = vk_first_texel(y0, p.height, p.depthH); }
    else { bx0 = int32_t(2 * x0); by0 = int32_t(2 * y0); }

    // ---- stage the input box -------------------------------------------------------------------------------------
    if (USE_TMA) {
        if (tid == 0) {
            mbar_init(&s_bar, 1);
            mbar_expect_tx(&s_bar, p.boxW * p.boxH * 4u);
            tma_load_2d(box, &tmap, bx0, by0, &s_bar);
        }
        __syncthreads();            // barrier init visible to the waiters
        mbar_wait(&s_bar, 0);
    } else {
        for (uint32_t j = tid; j < p.boxW * p.boxH; j += kPyrThreads) {
            const uint32_t bx = j % p.boxW, by = j / p.boxW;
            const int64_t gx = int64_t(bx0) + bx, gy = int64_t(by0) + by;
            box[j] = (gx < int64_t(p.depthW) && gy < int64_t(p.depthH)) ? __ldg(p.depth + size_t(gy) * p.depthW + size_t(gx)) : 0.0f;
        }
        __syncthreads();
    }

    // ---- mip 0 of the tile -------------------------------------------------------------------------------------------
    float* l0 = lvl;
    float* out0 = p.out + p.offset[0];
#pragma unroll
    for (int t = 0; t < (kPyrTile * kPyrTile) / kPyrThreads; ++t) {
        const uint32_t idx = tid + uint32_t(t) * kPyrThreads;
        const uint32_t lx = idx & 31u, ly = idx >> 5;
        const uint32_t x = x0 + lx, y = y0 + ly;
        float d = 0.0f;
        if (x < p.width && y < p.height) {
            if (VARIANT == HIZ_VK) {
                const float s = fdiv(fadd(float(x), 0.5f), float(p.width)), tt = fdiv(fadd(float(y), 0.5f), float(p.height));
                d = sample_min_linear([&](uint32_t i, uint32_t j) { return box[(j - uint32_t(by0)) * p.boxW + (i - uint32_t(bx0))]; },
                                      p.depthW, p.depthH, s, tt);
            } else {
                const float* b = box + (2u * ly) * p.boxW + 2u * lx;   // OOB texels were staged as 0
                const float r = b[0], g = b[1], bb = b[p.boxW], a = b[p.boxW + 1];
                d = min_keep(r, min_keep(g, min_keep(bb, a)));
            }
            out0[size_t(y) * p.width + x] = d;
        }
        l0[idx] = d;
    }
    __syncthreads();

    // ---- mips 1 .. tileLevels-1 inside shared memory (exact 2x2 min of the previous mip) -------------------------------
    float* prev = l0;
    uint32_t prevSize = kPyrTile;
    for (uint32_t k = 1; k < p.tileLevels; ++k) {
        const uint32_t size = kPyrTile >> k;
        float* cur = prev + prevSize * prevSize;
        const uint32_t lw = p.width >> k, lh = p.height >> k;       // no clamping inside the tile stage (host guarantees >= 1)
        float* outk = p.out + p.offset[k];
        for (uint32_t idx = tid; idx < size * size; idx += kPyrThreads) {
            const uint32_t lx = idx % size, ly = idx / size;
            const uint32_t x = (x0 >> k) + lx, y = (y0 >> k) + ly;
            const float* q = prev + (2u * ly) * prevSize + 2u * lx;
            float d;
            if (VARIANT == HIZ_VK) d = min_keep(min_keep(min_keep(q[0], q[1]), q[prevSize]), q[prevSize + 1]);
            else d = min_keep(q[0], min_keep(q[1], min_keep(q[prevSize], q[prevSize + 1])));
            cur[idx] = d;
            if (x < lw && y < lh) outk[size_t(y) * lw + x] = d;
        }
        __syncthreads();
        prev = cur; prevSize = size;
    }

    // ---- tail: the last CTA reduces the remaining mips from global memory ------------------------------------------------
    if (p.tileLevels >= p.mips) {
        return;
    }
    // release of this CTA's mip texels: CTA barrier, then ONE thread fences (cumulative) and takes the ticket.  (A __threadfence() in
    // every thread was 11 % of the kernel's stall samples: profiles/r02c_pyramid_*.)
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const uint32_t t = atomicAdd(p.ticket, 1u);
        s_last = (t == gridDim.x * gridDim.y - 1u) ? 1u : 0u;
        if (s_last) *p.ticket = 0u;                                  // re-arm for the next build on this stream
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // The tail is serial (one CTA, one level after the other): profiles/r01m_pyramid_raw.txt showed the SMs active for only half of the
    // kernel's 19 us -- every tail level paid an L2 round trip per texel fetch plus a fence.  When the first tail level's source fits the
    // shared memory this CTA already owns (32 x 32 texels at 1080p, 64 x 64 at 4K), it is loaded ONCE and the remaining levels are reduced
    // there, ping-pong between two shared buffers, with the same arithmetic; otherwise the levels are stepped through global memory.
    {
        const uint32_t k0 = p.tileLevels;
        const uint32_t sw0 = umax(1u, p.width >> (k0 - 1)), sh0 = umax(1u, p.height >> (k0 - 1));
        const uint32_t lw0 = umax(1u, p.width >> k0), lh0 = umax(1u, p.height >> k0);
        if (size_t(sw0) * sh0 + size_t(lw0) * lh0 <= p.smemFloats) {
            float* a = reinterpret_cast<float*>(smem_raw);                  // previous level
            float* b = a + size_t(sw0) * sh0;                               // current level (never larger than lw0 x lh0)
            const float* src0 = p.out + p.offset[k0 - 1];
            for (uint32_t idx = tid; idx < sw0 * sh0; idx += kPyrThreads) a[idx] = ld_cg_f32(src0 + idx);
            __syncthreads();
            for (uint32_t k = k0; k < p.mips; ++k) {
                const uint32_t lw = umax(1u, p.width >> k), lh = umax(1u, p.height >> k);
                const uint32_t sw = umax(1u, p.width >> (k - 1)), sh = umax(1u, p.height >> (k - 1));
                float* dst = p.out + p.offset[k];
                for (uint32_t idx = tid; idx < lw * lh; idx += kPyrThreads) {
                    const uint32_t x = idx % lw, y = idx / lw;
                    const float d = tail_texel_from<VARIANT>([&](uint32_t i, uint32_t j) { return a[size_t(j) * sw + i]; }, sw, sh, x, y, lw, lh);
                    b[idx] = d;
                    dst[idx] = d;
                }
                __syncthreads();
                float* t = a; a = b; b = t;                                  // the old source (>= 4x larger) becomes the next destination
            }
            return;
        }
    }
    for (uint32_t k = p.tileLevels; k < p.mips; ++k) {
        const uint32_t lw = umax(1u, p.width >> k), lh = umax(1u, p.height >> k);
        const uint32_t sw = umax(1u, p.width >> (k - 1)), sh = umax(1u, p.height >> (k - 1));
        const float* src = p.out + p.offset[k - 1];
        float* dst = p.out + p.offset[k];
        for (uint32_t idx = tid; idx < lw * lh; idx += kPyrThreads) {
            const uint32_t x = idx % lw, y = idx / lw;
            dst[idx] = tail_texel<VARIANT>(src, sw, sh, x, y, lw, lh);
        }
        __threadfence();
        __syncthreads();
    }
}

} // namespace

size_t pyramid_smem_bytes(const PyramidBuildParams& p)
{
    size_t boxFloats = (size_t(p.boxW) * p.boxH + 31u) & ~size_t(31);
    size_t lvlFloats = 0;
    for (int s = kPyrTile; s >= 1; s >>= 1) lvlFloats += size_t(s) * s;
    return (boxFloats + lvlFloats) * sizeof(float);
}

cudaError_t launch_pyramid_build(const PyramidBuildParams& p, const void* tensorMap, cudaStream_t stream)
{
    const size_t smem = pyramid_smem_bytes(p);
    PyramidBuildParams q = p;
    q.smemFloats = uint32_t(smem / sizeof(float));
    dim3 grid(p.tilesX, p.tilesY), block(kPyrThreads);
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    if (tensorMap) memcpy(&map, tensorMap, sizeof(map));
    void (*kernel)(const PyramidBuildParams, const CUtensorMap);
    if (p.variant == HIZ_VK) kernel = tensorMap ? pyramid_kernel<HIZ_VK, true> : pyramid_kernel<HIZ_VK, false>;
    else kernel = tensorMap ? pyramid_kernel<HIZ_DX, true> : pyramid_kernel<HIZ_DX, false>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    kernel<<<grid, block, smem, stream>>>(q, map);
    return cudaGetLastError();
}

// Host helper shared with the C-ABI layer: geometry of the tile stage.
void pyramid_plan(PyramidBuildParams& p)
{
    p.tilesX = (p.width + kPyrTile - 1) / kPyrTile;
    p.tilesY = (p.height + kPyrTile - 1) / kPyrTile;
    uint32_t lv = 0;
    while (lv < p.mips && lv < 6 && (p.width >> lv) >= 1 && (p.height >> lv) >= 1) ++lv;
    p.tileLevels = lv;
    if (p.variant == HIZ_VK) {
        uint32_t bw = 1, bh = 1;
        for (uint32_t tx = 0; tx < p.tilesX; ++tx) {
            uint32_t xa = tx * kPyrTile, xb = xa + kPyrTile - 1; if (xb > p.width - 1) xb = p.width - 1;
            int32_t a = vk_first_texel(xa, p.width, p.depthW), b = vk_first_texel(xb, p.width, p.depthW) + 1;
            if (uint32_t(b - a + 1) > bw) bw = uint32_t(b - a + 1);
        }
        for (uint32_t ty = 0; ty < p.tilesY; ++ty) {
            uint32_t ya = ty * kPyrTile, yb = ya + kPyrTile - 1; if (yb > p.height - 1) yb = p.height - 1;
            int32_t a = vk_first_texel(ya, p.height, p.depthH), b = vk_first_texel(yb, p.height, p.depthH) + 1;
            if (uint32_t(b - a + 1) > bh) bh = uint32_t(b - a + 1);
        }
        p.boxW = (bw + 3u) & ~3u; p.boxH = bh;
    } else {
        p.boxW = 2 * kPyrTile; p.boxH = 2 * kPyrTile;
    }
}

} // namespace blz
