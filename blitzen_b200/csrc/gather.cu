// Multi-GPU draw-list gather (new: the reference is single-GPU).
// Each rank culls + compacts its shard of the object list into its own draw buffer.  One kernel per rank then
//   1. publishes the rank's survivor count to the presenting GPU's flag block (one 64-bit peer store over NVLink),
//   2. waits for the counts of all lower ranks (the exclusive scan of <= 8 counts, done in-kernel),
//   3. stores its records into the presenter's gather buffer at that offset with peer stores,
//   4. raises its done flag on the presenter.
// Concatenation in shard order == ascending global objectId, so the gathered list is byte-identical to the 1-GPU list.
// No NCCL call and no host round trip on the data path; NCCL (torch.distributed) only carries the 128-byte IPC handle
// blob at set-up time (blitzen_b200/dist.py).
#include "ctx.h"
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace blz {

namespace {

constexpr int kGatherThreads = 256;
constexpr int kFlagStride = 64;      // ranks per row
constexpr int kEpochSlots = 4;       // count words of epoch e live in row e % 4: rows [0, 4); done words in row 4
constexpr int kDoneRow = kEpochSlots;
constexpr int kFlagWords = (kEpochSlots + 1) * kFlagStride;
// Epochs are consecutive integers starting at 1.  A rank may run ahead of a slower one (nothing on the host orders the
// pushes of different ranks), so the count words are kept per epoch in a ring of 4 rows and a rank does not publish epoch e
// before every rank has finished epoch e - 2: a word is never overwritten while somebody still waits for it.  The gather buffer has
// two halves, epoch e goes to half e & 1: the list of epoch e stays intact while epoch e + 1 is being gathered (the early list is
// still being drawn while the late list arrives); the consumer must be through with it before the ranks push epoch e + 2.

__device__ __forceinline__ void st_release_sys_u64(uint64_t* p, uint64_t v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ uint64_t ld_acquire_sys_u64(const uint64_t* p)
{
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Device-side waits are bounded: a rank that never arrives (a crashed process, a protocol error) must turn into an error code on the host, not
// into kernels that spin for ever on every other GPU of the box.  Each waiting thread reads %globaltimer every 64 polls; when the kernel has waited
// longer than the budget in total it records {what, epoch, peer} in a host-mapped word block, sets the context's sticky device word (so that every
// later gather kernel of this context skips its waits instead of paying the budget again), moves no data, and still raises its flags so that the
// timeout does not cascade as a hang.  The host reports BLZ_ERR_TIMEOUT from the next call that synchronises (blz_cull_synchronize,
// blz_cull_gather_read, blz_cull_consume_gathered).  Option "gather_timeout_ms" (default 60 000: generous, a false alarm costs a run; 0 = wait for ever).
enum : uint32_t { kWaitCount = 1u, kWaitBackPressure = 2u, kWaitExpand = 3u, kWaitRead = 4u };
struct SpinGuard {
    uint32_t* sticky; volatile uint32_t* host; uint64_t budgetNs, t0; uint32_t epoch; bool dead;
    __device__ __forceinline__ void begin(uint32_t* stickyWord, uint32_t* hostWords, uint64_t budget, uint32_t e)
    {
        sticky = stickyWord; host = hostWords; budgetNs = budget; epoch = e;
        uint32_t v;
        asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(stickyWord));
        dead = v != 0u;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    }
    __device__ __forceinline__ bool expired(uint32_t what, uint32_t peer)
    {
        if (budgetNs == 0ull) return false;
        uint64_t t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 <= budgetNs) return false;
        dead = true;
        if (atomicExch(sticky, 1u) == 0u) { host[1] = epoch; host[2] = peer; __threadfence_system(); host[0] = what; __threadfence_system(); }
        return true;
    }
};
// polls *word until ok(value); false = gave up (budget spent now or earlier)
template <class Ok>
__device__ __forceinline__ bool spin_until(const uint64_t* word, Ok ok, uint64_t& value, SpinGuard& g, uint32_t what, uint32_t peer)
{
    if (g.dead) return false;
    for (uint32_t polls = 1u;; ++polls) {
        value = ld_acquire_sys_u64(word);
        if (ok(value)) return true;
        if ((polls & 63u) == 0u && g.expired(what, peer)) return false;
    }
}

struct GatherParams {
    const uint32_t* src; const uint32_t* srcCount;   // this rank's compacted list (device-local)
    uint32_t* dst; uint64_t* flags;                  // presenter's buffer + flag block (peer-mapped, or local on the presenter)
    uint32_t* done;                                  // local CTA-completion counter (self-resetting)
    uint64_t capacity;                               // records the presenter's buffer holds
    uint32_t recWords, rank, world, epoch;
    uint32_t signalDone;                             // 0 on the presenter in descriptor mode: its done flag is raised by gather_expand_kernel
    uint32_t* sticky; uint32_t* errHost; uint64_t timeoutNs;   // bounded waits (SpinGuard)
};

// 64 threads, 32 registers, no shared memory to speak of: at most one CTA per SM, which fits in what the persistent cull kernels leave free
// (a 256-thread shape sat idle until the NEXT cull pass had finished: kernel timeline of the 2-GPU frame, profiles/r02g_trace_n2.txt)
constexpr int kPushThreads = 64;
__global__ void __launch_bounds__(kPushThreads) gather_push_kernel(const GatherParams p)
{
    __shared__ uint64_t s_off;
    const uint32_t tid = threadIdx.x;
    uint32_t count;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(count) : "l"(p.srcCount));
    if (tid == 0) {
        SpinGuard g; g.begin(p.sticky, p.errHost, p.timeoutNs, p.epoch);
        uint64_t* row = p.flags + (p.epoch % uint32_t(kEpochSlots)) * kFlagStride;
        uint64_t w;
        if (blockIdx.x == 0) {
            if (p.epoch > 2u)                                              // back-pressure: everybody is done with epoch e - 2 (same half of the gather buffer)
                for (uint32_t r = 0; r < p.world; ++r)
                    spin_until(p.flags + kDoneRow * kFlagStride + r, [&](uint64_t v) { return int32_t(uint32_t(v) - (p.epoch - 2u)) >= 0; }, w, g, kWaitBackPressure, r);
            st_release_sys_u64(row + p.rank, (uint64_t(p.epoch) << 32) | count);
        }
        uint64_t off = 0;
        for (uint32_t r = 0; r < p.rank; ++r) {
            if (spin_until(row + r, [&](uint64_t v) { return uint32_t(v >> 32) == p.epoch; }, w, g, kWaitCount, r)) off += uint32_t(w);
        }
        s_off = g.dead ? ~0ull : off;                                      // gave up: nothing is moved (room = 0 below), the flags are still raised
    }
    __syncthreads();
    const uint64_t off = s_off;
    const uint64_t room = off < p.capacity ? p.capacity - off : 0ull;
    const uint64_t nrec = room < count ? room : count;
    const uint64_t n64 = nrec * (p.recWords >> 1);
    const uint2* src = reinterpret_cast<const uint2*>(p.src);
    uint2* dst = reinterpret_cast<uint2*>(p.dst + off * p.recWords);
    // eight independent loads in flight per thread: with one, a CTA moves 256 x 8 B per local-memory round trip (measured 3.7 GB/s per
    // CTA -- the push was load-latency-bound, not NVLink-bound)
    const uint64_t stride = uint64_t(gridDim.x) * kPushThreads;
    uint64_t j = uint64_t(blockIdx.x) * kPushThreads + tid;
    for (; j + 7ull * stride < n64; j += 8ull * stride) {
        uint2 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = src[j + uint64_t(u) * stride];
#pragma unroll
        for (int u = 0; u < 8; ++u) dst[j + uint64_t(u) * stride] = v[u];
    }
    for (; j < n64; j += stride) dst[j] = src[j];
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        const uint32_t prev = atomicAdd(p.done, 1u);
        if (prev == gridDim.x - 1u) {
            *p.done = 0u;
            if (p.signalDone) st_release_sys_u64(p.flags + kDoneRow * kFlagStride + p.rank, uint64_t(p.epoch));
        }
    }
}

// A variant of this push that leaves the SM through TMA bulk stores (cp.async.bulk.global.shared::cta from a two-stage shared-memory ring) was built
// and measured in round 2 (profiles/r02h_n8_ab.txt): co-residency worked the same, but the bulk stores share the SM's copy
// engine with the cull kernels' own TMA loads (pyramid build next to a push 15 -> 31 us, late pass 175 -> 197 us).  Register stores stayed; the
// variant and the round-1 256-thread shape (which cannot start next to the persistent cull kernels at all) were deleted.

// Descriptor transport, presenter side.  The ranks shipped {objectId, absolute LOD id} (8 bytes) per record into the descriptor area of the gather
// buffer, concatenated in rank order exactly like records; this kernel waits until every OTHER rank's push of `epoch` has landed (done flags; the
// presenter's own push precedes it on the stream), then expands descriptor r into record r of the record area with the presenter's own LOD table --
// the table is replicated, so the result is byte for byte what the ranks' own passes wrote locally:
// {objectId, indexCount, 1, firstIndex, 0, 0 [, 0, 0]}.  The presenter's done flag of the epoch is raised HERE (last CTA out), not by its push:
// "every rank is done with epoch e" then also means "the descriptors of e have been read", which is what lets the ranks overwrite that half of the
// descriptor area with epoch e + 2 (back-pressure in the push kernel).  Same small shape as the push (64 threads, <= 1 CTA per SM, no shared memory
// to speak of) so that it runs next to the persistent cull kernels; 8 records in flight per thread (one record per round trip measured 0.4 ms for
// 1.2 M records: latency-bound).  8 GPUs: the presenter ingests 35 MB instead of 105 MB per frame.
// It must stay at 32 registers and (next to) no shared memory: 64 threads x 32 registers is what fits beside the push kernel in the 4 096 registers
// the persistent cull kernels leave free per SM, and shared memory beyond the cull kernels' carve-out makes its CTAs wait for a cull kernel to END --
// a variant with the LOD table staged in 8 KB of shared memory measured 0.89 ms per frame at 8 GPUs instead of 0.276 (profiles/r02l_n8_desc.txt).
// The expansion has to keep up with the frame: 48 CTAs instead of 148 measured 0.45 ms per frame (the next list's push queues behind it).
struct ExpandParams {
    const uint2* descs; uint32_t* records; uint64_t* flags; const LodData* lods; uint32_t* done; uint32_t lodCount;
    uint64_t capacity; uint32_t recWords, world, rank, epoch;
    uint32_t* sticky; uint32_t* errHost; uint64_t timeoutNs;
};
constexpr int kExpandThreads = 64;
constexpr int kExpandInFlight = 8;

template <int RECW2>     // uint2 words per record: 3 (VK24) or 4 (DX32)
__device__ __forceinline__ void expand_store(uint2* out, uint64_t r, uint2 d, uint2 L, bool ok)
{
    uint2* o = out + r * uint64_t(RECW2);
    const uint2 a = ok ? make_uint2(d.x, L.x) : make_uint2(0u, 0u), b = ok ? make_uint2(1u, L.y) : make_uint2(0u, 0u);
    asm volatile("st.global.cs.v2.u32 [%0], {%1, %2};" ::"l"(o), "r"(a.x), "r"(a.y) : "memory");
    asm volatile("st.global.cs.v2.u32 [%0], {%1, %2};" ::"l"(o + 1), "r"(b.x), "r"(b.y) : "memory");
#pragma unroll
    for (int f = 2; f < RECW2; ++f) asm volatile("st.global.cs.v2.u32 [%0], {%1, %2};" ::"l"(o + f), "r"(0u), "r"(0u) : "memory");
}

template <int RECW2>
__global__ void __launch_bounds__(kExpandThreads) gather_expand_kernel(const ExpandParams p)
{
    __shared__ uint64_t s_total;
    constexpr int kInFlight = RECW2 == 3 ? kExpandInFlight : kExpandInFlight - 2;    // DX32: 6, which keeps the kernel at 32 registers like the VK24 form
    const uint32_t tid = threadIdx.x;
    if (tid == 0) {
        SpinGuard g; g.begin(p.sticky, p.errHost, p.timeoutNs, p.epoch);
        uint64_t total = 0, w;
        for (uint32_t r = 0; r < p.world; ++r) {
            if (r != p.rank) spin_until(p.flags + kDoneRow * kFlagStride + r, [&](uint64_t v) { return int32_t(uint32_t(v) - p.epoch) >= 0; }, w, g, kWaitExpand, r);
            total += uint32_t(ld_acquire_sys_u64(p.flags + (p.epoch % uint32_t(kEpochSlots)) * kFlagStride + r));
        }
        s_total = g.dead ? 0ull : (total < p.capacity ? total : p.capacity);
    }
    __syncthreads();
    const uint64_t n = s_total;
    uint2* out = reinterpret_cast<uint2*>(p.records);
    const uint64_t stride = uint64_t(gridDim.x) * kExpandThreads;
    uint64_t r = uint64_t(blockIdx.x) * kExpandThreads + tid;
    for (; r + uint64_t(kInFlight - 1) * stride < n; r += uint64_t(kInFlight) * stride) {
        uint2 d[kInFlight], L[kInFlight];
#pragma unroll
        for (int u = 0; u < kInFlight; ++u) asm volatile("ld.global.cg.v2.u32 {%0, %1}, [%2];" : "=r"(d[u].x), "=r"(d[u].y) : "l"(p.descs + r + uint64_t(u) * stride));   // written by peers: L2, never L1
#pragma unroll
        for (int u = 0; u < kInFlight; ++u) L[u] = __ldg(reinterpret_cast<const uint2*>(p.lods + (d[u].y < p.lodCount ? d[u].y : 0u)));                                // {indexCount, firstIndex}
#pragma unroll
        for (int u = 0; u < kInFlight; ++u) expand_store<RECW2>(out, r + uint64_t(u) * stride, d[u], L[u], d[u].y < p.lodCount);
    }
    for (; r < n; r += stride) {
        uint2 d;
        asm volatile("ld.global.cg.v2.u32 {%0, %1}, [%2];" : "=r"(d.x), "=r"(d.y) : "l"(p.descs + r));
        const bool ok = d.y < p.lodCount;
        const uint2 L = __ldg(reinterpret_cast<const uint2*>(p.lods + (ok ? d.y : 0u)));
        expand_store<RECW2>(out, r, d, L, ok);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const uint32_t prev = atomicAdd(p.done, 1u);
        if (prev == gridDim.x - 1u) {
            *p.done = 0u;
            st_release_sys_u64(p.flags + kDoneRow * kFlagStride + p.rank, uint64_t(p.epoch));
        }
    }
}

__global__ void gather_wait_kernel(const uint64_t* flags, uint32_t world, uint32_t epoch, uint32_t* sticky, uint32_t* errHost, uint64_t timeoutNs)
{
    const uint32_t r = threadIdx.x;
    if (r < world) {
        SpinGuard g; g.begin(sticky, errHost, timeoutNs, epoch);
        uint64_t w;
        spin_until(flags + kDoneRow * kFlagStride + r, [&](uint64_t v) { return int32_t(uint32_t(v) - epoch) >= 0; }, w, g, kWaitRead, r);
    }
}

// Instance-list gather (indirect instancing, objects sharded by contiguous ranges): the bucket of LOD l on the presenter is the concatenation of
// the ranks' buckets in rank order (= ascending objectId).  The per-rank per-LOD counts arrive by an NCCL all-gather issued by the host layer on
// the same stream (all-gather of the counts, exclusive scan, variable-length gather); this kernel does the scan (<= 8 ranks) and the
// peer stores.  grid.y = LOD.
struct InstGatherParams {
    const uint32_t* src; const LodInstanceCounter* local;     // this rank's buckets: instanceOffset (local layout) + instanceCount
    const uint32_t* localCap;
    const uint32_t* allCounts;                                // [world][lodCount] instanceCount of every rank (device)
    const uint32_t* globalOffset; const uint32_t* globalCap;  // presenter's bucket layout
    uint32_t* dst;                                            // presenter's instance index buffer (peer-mapped, or local on the presenter)
    uint32_t lodCount, rank, world;
};

__global__ void __launch_bounds__(kGatherThreads) gather_instances_kernel(const InstGatherParams p)
{
    const uint32_t l = blockIdx.y;
    if (l >= p.lodCount) return;
    uint32_t base = 0u;
    for (uint32_t r = 0; r < p.rank; ++r) base += p.allCounts[r * p.lodCount + l];
    uint32_t cnt = p.allCounts[p.rank * p.lodCount + l];
    if (cnt > p.localCap[l]) cnt = p.localCap[l];
    const uint32_t room = base < p.globalCap[l] ? p.globalCap[l] - base : 0u;
    if (cnt > room) cnt = room;
    const uint32_t* s = p.src + p.local[l].instanceOffset;
    uint32_t* d = p.dst + size_t(p.globalOffset[l]) + base;
    for (uint32_t i = blockIdx.x * kGatherThreads + threadIdx.x; i < cnt; i += gridDim.x * kGatherThreads) d[i] = s[i];
    __threadfence_system();      // the peer stores are visible system-wide before the kernel ends: the completion collective that follows on this stream covers them
}

// what a rank's bucket of LOD l really holds: min(instanceCount, local bucket capacity) -- the words the ranks all-gather, so that the
// exclusive scan over the lower ranks (the bucket's base on the presenter) counts exactly the ids that will be stored there: no holes
__global__ void instance_counts_kernel(const LodInstanceCounter* li, const uint32_t* cap, uint32_t lodCount, uint32_t* dst)
{
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < lodCount) { const uint32_t c = li[l].instanceCount; dst[l] = c < cap[l] ? c : cap[l]; }
}

} // namespace

void gather_release(blz_cull_ctx* c)
{
    if (c->gatherPeerMapped) {
        if (c->gatherDst) cudaIpcCloseMemHandle(c->gatherDst);
        if (c->gatherDstFlags) cudaIpcCloseMemHandle(c->gatherDstFlags);
    }
    if (c->gatherOwner) { if (c->gatherBuf) cudaFree(c->gatherBuf); if (c->gatherFlags) cudaFree(c->gatherFlags); }
    if (c->instDstMapped && c->instDst) cudaIpcCloseMemHandle(c->instDst);
    c->instDst = nullptr; c->instDstMapped = false;
    if (c->gatherDone) cudaFree(c->gatherDone);
    if (c->gatherErrHost) { cudaFreeHost(c->gatherErrHost); c->gatherErrHost = nullptr; c->gatherErrDev = nullptr; }
    if (c->gatherStream) { cudaStreamSynchronize(c->gatherStream); cudaStreamDestroy(c->gatherStream); cudaEventDestroy(c->evCull); cudaEventDestroy(c->evPush[0]); cudaEventDestroy(c->evPush[1]); cudaEventDestroy(c->evExpand); c->evExpandValid = false; c->gatherStream = nullptr; c->evPushValid[0] = c->evPushValid[1] = false; }
    c->gatherBuf = nullptr; c->gatherFlags = nullptr; c->gatherDst = nullptr; c->gatherDstFlags = nullptr; c->gatherDone = nullptr;
    c->gatherOwner = c->gatherImported = c->gatherPeerMapped = false;
}

// Called behind a host synchronisation: turns a device-side wait that gave up into BLZ_ERR_TIMEOUT (once) and re-arms the waits.  The ranks'
// epochs no longer agree after a timeout: the gather has to be set up again (export / import) before it is used further; the context stays usable.
int gather_report_timeout(blz_cull_ctx* c)
{
    if (!c->gatherErrHost) return BLZ_OK;
    volatile uint32_t* h = c->gatherErrHost;
    const uint32_t what = h[0];
    if (what == 0u) return BLZ_OK;
    const uint32_t epoch = h[1], peer = h[2];
    h[0] = 0u;
    cudaMemset(c->gatherDone + 2, 0, sizeof(uint32_t));
    static const char* const names[] = { "?", "rank %u's record count of epoch %u", "rank %u to finish with epoch %u - 2 (back-pressure)", "rank %u's push of epoch %u (expansion)", "rank %u's push of epoch %u (read)" };
    char what_s[160];
    snprintf(what_s, sizeof what_s, names[what <= 4u ? what : 0u], peer, epoch);
    return fail(BLZ_ERR_TIMEOUT, "draw-list gather: rank %d waited more than %lld ms on the device for %s; nothing was moved for that push -- set the gather up again on every rank", c->rank, (long long)c->optGatherTimeoutMs, what_s);
}

} // namespace blz

using namespace blz;

extern "C" {

int blz_cull_gather_export(blz_cull_ctx* c, uint64_t capacityRecords, int fmt, void* outBlob128)
{
    if (!c || !outBlob128 || capacityRecords == 0) return fail(BLZ_ERR_INVALID, "bad argument");
    if (fmt != BLZ_REC_VK24 && fmt != BLZ_REC_DX32) return fail(BLZ_ERR_INVALID, "record format %d", fmt);
    CU_TRY(cudaSetDevice(c->device));
    gather_release(c);
    c->gatherRecWords = fmt == BLZ_REC_VK24 ? 6u : 8u;
    c->gatherCap = capacityRecords;
    // two halves (epoch parity) of records, then two halves of 8-byte descriptors (descriptor transport) in the SAME allocation: one IPC handle
    CU_TRY(cudaMalloc(&c->gatherBuf, 2u * size_t(capacityRecords) * c->gatherRecWords * 4u + 2u * size_t(capacityRecords) * 8u));
    CU_TRY(cudaMalloc(&c->gatherFlags, kFlagWords * sizeof(uint64_t)));
    CU_TRY(cudaMemset(c->gatherFlags, 0, kFlagWords * sizeof(uint64_t)));
    c->gatherOwner = true;
    cudaIpcMemHandle_t h[2];
    CU_TRY(cudaIpcGetMemHandle(&h[0], c->gatherBuf));
    CU_TRY(cudaIpcGetMemHandle(&h[1], c->gatherFlags));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    unsigned char* o = static_cast<unsigned char*>(outBlob128);
    memcpy(o, &h[0], 64); memcpy(o + 64, &h[1], 64);
    // trailer-free blob: capacity and format travel separately (blitzen_b200/dist.py broadcasts them with the blob)
    return BLZ_OK;
}

int blz_cull_gather_import(blz_cull_ctx* c, const void* blob128, int rank, int world)
{
    if (!c || rank < 0 || world < 1 || rank >= world || world > kFlagStride) return fail(BLZ_ERR_INVALID, "bad rank/world %d/%d", rank, world);
    CU_TRY(cudaSetDevice(c->device));
    c->rank = rank; c->world = world;
    if (c->gatherOwner) {                       // the presenter writes through its own pointers
        c->gatherDst = c->gatherBuf; c->gatherDstFlags = c->gatherFlags; c->gatherPeerMapped = false;
    } else {
        if (!blob128) return fail(BLZ_ERR_INVALID, "presenter blob is null");
        cudaIpcMemHandle_t h[2];
        memcpy(&h[0], blob128, 64); memcpy(&h[1], static_cast<const unsigned char*>(blob128) + 64, 64);
        void *a = nullptr, *b = nullptr;
        CU_TRY(cudaIpcOpenMemHandle(&a, h[0], cudaIpcMemLazyEnablePeerAccess));
        CU_TRY(cudaIpcOpenMemHandle(&b, h[1], cudaIpcMemLazyEnablePeerAccess));
        c->gatherDst = static_cast<uint32_t*>(a); c->gatherDstFlags = static_cast<uint64_t*>(b); c->gatherPeerMapped = true;
    }
    if (!c->gatherDone) {
        CU_TRY(cudaMalloc(&c->gatherDone, 4 * sizeof(uint32_t)));          // CTA-completion counters: [0] push, [1] expansion (self-resetting); [2] sticky "a wait timed out" word
        CU_TRY(cudaMemset(c->gatherDone, 0, 4 * sizeof(uint32_t)));
        CU_TRY(cudaHostAlloc(reinterpret_cast<void**>(&c->gatherErrHost), 4 * sizeof(uint32_t), cudaHostAllocMapped));
        memset(c->gatherErrHost, 0, 4 * sizeof(uint32_t));
        CU_TRY(cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->gatherErrDev), c->gatherErrHost, 0));
    }
    c->gatherImported = true;
    return BLZ_OK;
}

// capacity/format of the presenter's buffer for ranks that did not export (set by the host layer after the broadcast)
int blz_cull_gather_configure(blz_cull_ctx* c, uint64_t capacityRecords, int fmt)
{
    if (!c) return fail(BLZ_ERR_INVALID, "null context");
    c->gatherCap = capacityRecords; c->gatherRecWords = fmt == BLZ_REC_VK24 ? 6u : 8u;
    return BLZ_OK;
}

static bool gather_desc_mode(const blz_cull_ctx* c) { return c->descValid && c->optGatherDesc != 0; }   // every rank runs the same pass sequence with the same options: the mode is uniform

static int gather_launch(blz_cull_ctx* c, uint32_t epoch, cudaStream_t stream, int grid)
{
    GatherParams p{};
    const bool desc = gather_desc_mode(c);
    uint32_t* descArea = c->gatherDst + 2u * size_t(c->gatherCap) * c->gatherRecWords;   // behind the two record halves (8-byte aligned: recWords is even)
    p.src = desc ? reinterpret_cast<const uint32_t*>(c->descs) : c->draws; p.srcCount = c->drawCounts; p.flags = c->gatherDstFlags; p.done = c->gatherDone;
    p.recWords = desc ? 2u : c->gatherRecWords;
    p.dst = desc ? descArea + size_t(epoch & 1u) * c->gatherCap * 2u : c->gatherDst + size_t(epoch & 1u) * c->gatherCap * c->gatherRecWords;
    p.capacity = c->gatherCap; p.rank = uint32_t(c->rank); p.world = uint32_t(c->world); p.epoch = epoch;
    p.signalDone = (desc && c->gatherOwner) ? 0u : 1u;
    p.sticky = c->gatherDone + 2; p.errHost = c->gatherErrDev; p.timeoutNs = uint64_t(c->optGatherTimeoutMs) * 1000000ull;
    const int perSM = grid > 0 && grid < c->numSMs ? grid : c->numSMs;             // one small co-resident CTA per SM at most
    gather_push_kernel<<<perSM, kPushThreads, 0, stream>>>(p);
    CU_TRY(cudaGetLastError());
    c->launches++;
    return BLZ_OK;
}

// descriptor mode, presenter only: expands what everybody shipped for `epoch` (behind the presenter's own push on the same stream)
static int gather_expand_launch(blz_cull_ctx* c, uint32_t epoch, cudaStream_t stream)
{
    ExpandParams e{};
    e.descs = reinterpret_cast<const uint2*>(c->gatherBuf + 2u * size_t(c->gatherCap) * c->gatherRecWords) + size_t(epoch & 1u) * c->gatherCap;
    e.records = c->gatherBuf + size_t(epoch & 1u) * c->gatherCap * c->gatherRecWords;
    e.flags = c->gatherFlags; e.lods = c->lods; e.lodCount = c->nLods; e.capacity = c->gatherCap; e.recWords = c->gatherRecWords;
    e.done = c->gatherDone + 1; e.world = uint32_t(c->world); e.rank = uint32_t(c->rank); e.epoch = epoch;
    e.sticky = c->gatherDone + 2; e.errHost = c->gatherErrDev; e.timeoutNs = uint64_t(c->optGatherTimeoutMs) * 1000000ull;
    static const int envCtas = [] { const char* v = getenv("BLZ_EXPAND_CTAS"); return v ? atoi(v) : 0; }();   // measurement aid
    const int grid = envCtas > 0 && envCtas < c->numSMs ? envCtas : c->numSMs;                                // one small co-resident CTA per SM at most
    if (c->gatherRecWords == 6u) gather_expand_kernel<3><<<grid, kExpandThreads, 0, stream>>>(e);
    else gather_expand_kernel<4><<<grid, kExpandThreads, 0, stream>>>(e);
    CU_TRY(cudaGetLastError());
    c->launches++;
    return BLZ_OK;
}

static int gather_check(blz_cull_ctx* c, uint32_t epoch)
{
    if (!c || !c->gatherImported) return fail(BLZ_ERR_INVALID, "gather not set up (export/import first)");
    if (epoch == 0) return fail(BLZ_ERR_INVALID, "epoch must be non-zero and change every push");
    if (c->lastRecWords != c->gatherRecWords) return fail(BLZ_ERR_INVALID, "last pass wrote %u-word records, gather buffer holds %u-word records", c->lastRecWords, c->gatherRecWords);
    return BLZ_OK;
}

int blz_cull_gather_push(blz_cull_ctx* c, uint32_t epoch)
{
    int rc = gather_check(c, epoch); if (rc) return rc;
    CU_TRY(cudaSetDevice(c->device));
    // the push / expansion kernels share their CTA-completion counters with the asynchronous form: order this push behind whatever the side stream still runs
    for (int k = 0; k < 2; ++k) if (c->evPushValid[k]) CU_TRY(cudaStreamWaitEvent(c->stream, c->evPush[k], 0));
    if (c->evExpandValid) CU_TRY(cudaStreamWaitEvent(c->stream, c->evExpand, 0));
    const bool expand = gather_desc_mode(c) && c->gatherOwner;
    rc = gather_launch(c, epoch, c->stream, c->numSMs); if (rc) return rc;
    return expand ? gather_expand_launch(c, epoch, c->stream) : BLZ_OK;
}

// Same push, off the critical path: it runs on a side stream behind the pass that produced the list, and the context flips to its
// second draw buffer (+ count words), so the NEXT pass on the main stream overlaps the NVLink transfer instead of waiting for it.
// After the call `draws` / `draw_count` (blz_cull_get_outputs, blz_cull_read_draws) refer to the buffer the next pass will write;
// the pushed list is on the presenter (blz_cull_gather_read).  A pass that would overwrite a buffer still being pushed waits for it.
int blz_cull_gather_push_async(blz_cull_ctx* c, uint32_t epoch)
{
    int rc = gather_check(c, epoch); if (rc) return rc;
    CU_TRY(cudaSetDevice(c->device));
    if (!c->gatherStream) {
        CU_TRY(cudaStreamCreateWithFlags(&c->gatherStream, cudaStreamNonBlocking));
        CU_TRY(cudaEventCreateWithFlags(&c->evCull, cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&c->evPush[0], cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&c->evPush[1], cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&c->evExpand, cudaEventDisableTiming));
    }
    if (c->expDraws.active) return fail(BLZ_ERR_STATE, "asynchronous pushes alternate two draw buffers: not available once the outputs have been exported");
    if (!c->drawsAlt) CU_TRY(cudaMalloc(&c->drawsAlt, c->capDraws));
    CU_TRY(cudaEventRecord(c->evCull, c->stream));
    CU_TRY(cudaStreamWaitEvent(c->gatherStream, c->evCull, 0));
    // How many co-resident 64-thread CTAs push: the presenter's kernels slow down while it ingests (world - 1) lists at once, so the ranks
    // throttle themselves -- the list has the rest of the frame to arrive.  Measured at 8 GPUs (profiles/r02h_n8_ab.txt, ms per frame):
    // 148 CTAs 0.308 (pyramid next to the push 0.053 ms), 16 CTAs 0.290, 4 CTAs 0.746 (the push itself becomes the frame).
    static const int envCtas = [] { const char* e = getenv("BLZ_GATHER_CTAS"); return e ? atoi(e) : 0; }();
    int ctas = 224 / (c->world > 1 ? c->world - 1 : 1);
    ctas = ctas < 16 ? 16 : (ctas > 64 ? 64 : ctas);
    if (envCtas > 0) ctas = envCtas;
    const bool expand = gather_desc_mode(c) && c->gatherOwner;
    rc = gather_launch(c, epoch, c->gatherStream, ctas); if (rc) return rc;
    CU_TRY(cudaEventRecord(c->evPush[c->drawSlot], c->gatherStream));      // the draw (and descriptor) buffer of this slot is free again from here on
    c->evPushValid[c->drawSlot] = true;
    if (expand) {
        // the presenter's expansion waits for the OTHER ranks' pushes: it must not sit in front of evPush, or the presenter's next pass would wait for them too
        rc = gather_expand_launch(c, epoch, c->gatherStream); if (rc) return rc;
        CU_TRY(cudaEventRecord(c->evExpand, c->gatherStream));
        c->evExpandValid = true;
    }
    uint32_t* t = c->draws; c->draws = c->drawsAlt; c->drawsAlt = t;
    { uint2* d = c->descs; c->descs = c->descsAlt; c->descsAlt = d; const bool v = c->descValid; c->descValid = c->descValidAlt; c->descValidAlt = v; }
    const uint32_t w = c->lastRecWords; c->lastRecWords = c->lastRecWordsAlt; c->lastRecWordsAlt = w;
    c->drawSlot ^= 1;
    c->drawCounts = c->counts + (c->drawSlot ? 8 : 0);
    // the buffer the NEXT draw-writing pass will use may still be being pushed (its previous push): that pass waits for it when it starts
    // (acquire_draw_buffer in capi.cu) -- not here, where the wait would also hold up passes that do not write draws (the pyramid build)
    c->drawBufferPending = c->evPushValid[c->drawSlot];
    return BLZ_OK;
}

// makes the context's main stream wait for every asynchronous push issued so far (stream-ordered, no host synchronisation)
int blz_cull_gather_join(blz_cull_ctx* c)
{
    if (!c) return fail(BLZ_ERR_INVALID, "null context");
    CU_TRY(cudaSetDevice(c->device));
    for (int k = 0; k < 2; ++k) if (c->evPushValid[k]) CU_TRY(cudaStreamWaitEvent(c->stream, c->evPush[k], 0));
    if (c->evExpandValid) CU_TRY(cudaStreamWaitEvent(c->stream, c->evExpand, 0));
    return BLZ_OK;
}

int blz_cull_gather_read(blz_cull_ctx* c, uint32_t epoch, void* recordsHost, uint64_t capacityRecords, uint32_t* outCounts)
{
    if (!c || !c->gatherOwner) return fail(BLZ_ERR_INVALID, "only the presenting rank (the exporter) can read the gathered list");
    CU_TRY(cudaSetDevice(c->device));
    if (c->evExpandValid) CU_TRY(cudaStreamWaitEvent(c->stream, c->evExpand, 0));   // asynchronous pushes in descriptor mode: the expansion runs on the side stream
    gather_wait_kernel<<<1, kFlagStride, 0, c->stream>>>(c->gatherFlags, uint32_t(c->world), epoch, c->gatherDone + 2, c->gatherErrDev, uint64_t(c->optGatherTimeoutMs) * 1000000ull);
    CU_TRY(cudaGetLastError());
    c->launches++;
    uint64_t flags[kFlagStride];
    CU_TRY(cudaMemcpyAsync(flags, c->gatherFlags + (epoch % uint32_t(kEpochSlots)) * kFlagStride, sizeof(flags), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    { const int trc = gather_report_timeout(c); if (trc) return trc; }
    uint64_t total = 0;
    for (int r = 0; r < c->world; ++r) { if (outCounts) outCounts[r] = uint32_t(flags[r]); total += uint32_t(flags[r]); }
    if (total > c->gatherCap) total = c->gatherCap;
    if (recordsHost) {
        uint64_t n = total < capacityRecords ? total : capacityRecords;
        if (n) {
            CU_TRY(cudaMemcpyAsync(recordsHost, c->gatherBuf + size_t(epoch & 1u) * c->gatherCap * c->gatherRecWords, size_t(n) * c->gatherRecWords * 4u, cudaMemcpyDeviceToHost, c->stream));
            CU_TRY(cudaStreamSynchronize(c->stream));
        }
    }
    return BLZ_OK;
}

// ---- instance-list gather -------------------------------------------------------------------------------------------------------
int blz_cull_instances_export(blz_cull_ctx* c, void* outBlob64)
{
    if (!c || !outBlob64 || !c->instIdx) return fail(BLZ_ERR_INVALID, "scene was uploaded without lod_instances");
    CU_TRY(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    CU_TRY(cudaIpcGetMemHandle(&h, c->instIdx));
    memcpy(outBlob64, &h, 64);
    return BLZ_OK;
}

int blz_cull_instances_import(blz_cull_ctx* c, const void* presenterBlob64, int rank, int world)
{
    if (!c || rank < 0 || world < 1 || rank >= world) return fail(BLZ_ERR_INVALID, "bad rank/world %d/%d", rank, world);
    CU_TRY(cudaSetDevice(c->device));
    if (c->instDstMapped && c->instDst) { cudaIpcCloseMemHandle(c->instDst); c->instDst = nullptr; c->instDstMapped = false; }
    c->rank = rank; c->world = world;
    if (!presenterBlob64) { c->instDst = c->instIdx; c->instDstMapped = false; return BLZ_OK; }      // the presenter itself
    cudaIpcMemHandle_t h;
    memcpy(&h, presenterBlob64, 64);
    void* a = nullptr;
    CU_TRY(cudaIpcOpenMemHandle(&a, h, cudaIpcMemLazyEnablePeerAccess));
    c->instDst = static_cast<uint32_t*>(a); c->instDstMapped = true;
    return BLZ_OK;
}

// all_counts_device: [world][lod_count] instanceCount of every rank; global_offset / global_cap: the presenter's bucket layout (device arrays).
// On the presenter the destination IS its own instance buffer, whose local buckets would be overwritten while they are read: the presenter
// must be rank 0 (its data stays where it is: global offset == local offset, base 0).
int blz_cull_instances_push(blz_cull_ctx* c, const uint32_t* allCounts, const uint32_t* globalOffset, const uint32_t* globalCap)
{
    if (!c || !c->instDst || !c->lodInst || !allCounts || !globalOffset || !globalCap) return fail(BLZ_ERR_INVALID, "instance gather not set up");
    CU_TRY(cudaSetDevice(c->device));
    if (!c->instDstMapped) return BLZ_OK;                     // rank 0: already in place
    InstGatherParams p{};
    p.src = c->instIdx; p.local = c->lodInst; p.localCap = c->bucketCap; p.allCounts = allCounts; p.globalOffset = globalOffset; p.globalCap = globalCap;
    p.dst = c->instDst; p.lodCount = c->nLods; p.rank = uint32_t(c->rank); p.world = uint32_t(c->world);
    dim3 grid(16u, c->nLods);
    gather_instances_kernel<<<grid, kGatherThreads, 0, c->stream>>>(p);
    CU_TRY(cudaGetLastError());
    c->launches++;
    return BLZ_OK;
}

// the per-LOD number of ids the last instancing pass STORED (instanceCount clamped to the bucket capacity), packed into a caller-owned DEVICE array
// (stream-ordered): what the host layer all-gathers
int blz_cull_instances_counts(blz_cull_ctx* c, uint32_t* dstDevice)
{
    if (!c || !c->lodInst || !dstDevice) return fail(BLZ_ERR_INVALID, "scene was uploaded without lod_instances");
    CU_TRY(cudaSetDevice(c->device));
    if (!c->bucketCap) return fail(BLZ_ERR_INVALID, "scene was uploaded without bucket capacities");
    instance_counts_kernel<<<(c->nLods + 127u) / 128u, 128, 0, c->stream>>>(c->lodInst, c->bucketCap, c->nLods, dstDevice);
    CU_TRY(cudaGetLastError());
    c->launches++;
    return BLZ_OK;
}

int blz_cull_gather_outputs(blz_cull_ctx* c, void** outRecords, uint32_t** outFlags)
{
    if (!c || !c->gatherOwner) return fail(BLZ_ERR_INVALID, "only the presenting rank owns gather outputs");
    if (outRecords) *outRecords = c->gatherBuf;
    if (outFlags) *outFlags = reinterpret_cast<uint32_t*>(c->gatherFlags);
    return BLZ_OK;
}

} // extern "C"
