// Single-pass chained scan with decoupled look-back (Merrill & Garland 2016), the deterministic replacement for the
// reference's `atomicAdd(indirectDrawCountBuffer.drawCount, 1)` append (VulkanShaders/InitialDrawCull.comp.glsl:49).
//
// One 64-bit status word per tile:  [63:34] launch epoch  [33:32] state  [31:0] value
//   state 1 = tile aggregate published, state 2 = inclusive prefix published.
// A word is valid only if its epoch equals the current launch's epoch, so the status array never needs clearing.
// Value and flag travel in ONE word, hence relaxed 64-bit accesses are sufficient (no separate fence).
#pragma once
#include <stdint.h>

namespace blz {

constexpr uint32_t kStateAggregate = 1u;
constexpr uint32_t kStateInclusive = 2u;

__device__ __forceinline__ uint64_t pack_status(uint32_t epoch, uint32_t state, uint32_t value)
{
    return (uint64_t((epoch << 2) | state) << 32) | uint64_t(value);
}
__device__ __forceinline__ void st_status(uint64_t* p, uint64_t v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_status(const uint64_t* p)
{
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Called by ONE FULL WARP (all 32 lanes).  Publishes `aggregate` for `tile`, resolves the exclusive prefix of the tile by
// looking back over its predecessors 32 at a time, publishes the inclusive prefix and returns the exclusive prefix (all lanes).
__device__ __forceinline__ uint32_t lookback_exclusive_prefix(uint64_t* status, uint32_t tile, uint32_t aggregate, uint32_t epoch, uint32_t lane)
{
    epoch &= 0x3FFFFFFFu;
    if (tile == 0) {
        if (lane == 0) st_status(status, pack_status(epoch, kStateInclusive, aggregate));
        return 0u;
    }
    if (lane == 0) st_status(status + tile, pack_status(epoch, kStateAggregate, aggregate));
    uint32_t exclusive = 0;
    int64_t look = int64_t(tile) - 1 - int64_t(lane);
    while (true) {
        uint32_t state = kStateInclusive, value = 0;      // virtual tiles before tile 0: inclusive prefix 0
        if (look >= 0) {
            uint64_t w;
            do { w = ld_status(status + look); } while (uint32_t(w >> 34) != epoch || (uint32_t(w >> 32) & 3u) == 0u);
            state = uint32_t(w >> 32) & 3u;
            value = uint32_t(w);
        }
        uint32_t inclMask = __ballot_sync(0xFFFFFFFFu, state == kStateInclusive);
        if (inclMask != 0u) {
            uint32_t first = uint32_t(__ffs(int(inclMask))) - 1u;   // nearest predecessor holding an inclusive prefix
            exclusive += __reduce_add_sync(0xFFFFFFFFu, lane <= first ? value : 0u);
            break;
        }
        exclusive += __reduce_add_sync(0xFFFFFFFFu, value);
        look -= 32;
    }
    if (lane == 0) st_status(status + tile, pack_status(epoch, kStateInclusive, exclusive + aggregate));
    return exclusive;
}

// CTA-wide variant: every thread of the CTA (THREADS = blockDim.x, a multiple of 32) calls it; windows of THREADS predecessors
// per step instead of 32.  One-shot CTAs that all finish at about the same time would otherwise walk back over every
// concurrently running tile 32 at a time (an L2 round trip per step).  `scratch` = 2 * THREADS / 32 + 2 shared words.
// Returns the exclusive prefix to all threads; contains __syncthreads().
template <int THREADS>
__device__ __forceinline__ uint32_t lookback_exclusive_prefix_cta(uint64_t* status, uint32_t tile, uint32_t aggregate, uint32_t epoch, uint32_t* scratch)
{
    constexpr int WARPS = THREADS / 32;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    epoch &= 0x3FFFFFFFu;
    uint32_t* s_first = scratch;            // per warp: lane of the nearest inclusive predecessor, 32 = none
    uint32_t* s_sum = scratch + WARPS;      // per warp: sum up to and including that lane (or of all 32)
    uint32_t* s_out = scratch + 2 * WARPS;  // [0] = running exclusive prefix, [1] = done flag
    if (tile == 0) {
        if (tid == 0) st_status(status, pack_status(epoch, kStateInclusive, aggregate));
        return 0u;
    }
    if (tid == 0) { st_status(status + tile, pack_status(epoch, kStateAggregate, aggregate)); s_out[0] = 0u; s_out[1] = 0u; }
    int64_t base = int64_t(tile);           // predecessors [base - THREADS, base) are examined in this step, nearest first
    while (true) {
        const int64_t look = base - 1 - int64_t(tid);
        uint32_t state = kStateInclusive, value = 0;              // virtual tiles before tile 0: inclusive prefix 0
        if (look >= 0) {
            uint64_t w;
            do { w = ld_status(status + look); } while (uint32_t(w >> 34) != epoch || (uint32_t(w >> 32) & 3u) == 0u);
            state = uint32_t(w >> 32) & 3u;
            value = uint32_t(w);
        }
        const uint32_t inclMask = __ballot_sync(0xFFFFFFFFu, state == kStateInclusive);
        const uint32_t first = inclMask ? uint32_t(__ffs(int(inclMask))) - 1u : 32u;
        const uint32_t sum = __reduce_add_sync(0xFFFFFFFFu, lane <= first ? value : 0u);
        if (lane == 0) { s_first[warp] = first; s_sum[warp] = sum; }
        __syncthreads();
        if (tid == 0) {
            uint32_t acc = s_out[0];
            bool found = false;
            for (int w = 0; w < WARPS && !found; ++w) { acc += s_sum[w]; found = s_first[w] < 32u; }
            s_out[0] = acc; s_out[1] = found ? 1u : 0u;
        }
        __syncthreads();
        if (s_out[1] != 0u) break;
        base -= THREADS;
        __syncthreads();                                           // s_out[1] read by everybody before thread 0 rewrites it
    }
    const uint32_t exclusive = s_out[0];
    if (tid == 0) st_status(status + tile, pack_status(epoch, kStateInclusive, exclusive + aggregate));
    return exclusive;
}

// Serial variant used by one THREAD per independent chain (the per-LOD chains of the instancing pass): chain c of tile t
// lives at status[t * stride + c].
__device__ __forceinline__ uint32_t lookback_exclusive_prefix_serial(uint64_t* status, uint32_t stride, uint32_t chain, uint32_t tile,
                                                                     uint32_t aggregate, uint32_t epoch)
{
    epoch &= 0x3FFFFFFFu;
    uint64_t* mine = status + size_t(tile) * stride + chain;
    if (tile == 0) { st_status(mine, pack_status(epoch, kStateInclusive, aggregate)); return 0u; }
    st_status(mine, pack_status(epoch, kStateAggregate, aggregate));
    uint32_t exclusive = 0;
    for (int64_t look = int64_t(tile) - 1; look >= 0; --look) {
        const uint64_t* p = status + size_t(look) * stride + chain;
        uint64_t w;
        do { w = ld_status(p); } while (uint32_t(w >> 34) != epoch || (uint32_t(w >> 32) & 3u) == 0u);
        exclusive += uint32_t(w);
        if ((uint32_t(w >> 32) & 3u) == kStateInclusive) break;
    }
    st_status(mine, pack_status(epoch, kStateInclusive, exclusive + aggregate));
    return exclusive;
}

} // namespace blz
