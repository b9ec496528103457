// C1: collectives.  New surface (the reference has no collective): one communicator per
// context (= per process = per GPU).
//
// Sum of the weight adjoints (dsc_?allreduce_sum), up to 8 GPUs of one NVSwitch node: ONE kernel over
// peer memory.  Every rank owns a symmetric buffer (header | region A | region B) that all peers map
// through CUDA IPC at dsc_comm_init.  The kernel (one CTA per SM on the compute stream, captured in
// the step's CUDA graph like any other kernel) pushes: every byte that crosses NVLink is a store.
//   0. scatter: slice q of the rank's (packed) adjoint buffers -> region A of rank q, slot [rank],
//   2. rank r adds the N slots of its region A (local) in rank order 0..N-1, so the sum is bit-identical
//      on every rank and from run to run, and stores the result slice into region B of EVERY rank,
//   4. unpacks region B into the adjoint buffers.
// Two kernels implement it.  allreduce_ll_kernel (the default, further down): no barrier at all — every
// 128-byte line carries its own flag and the receiver polls the data.  allreduce_p2p_kernel (round 1,
// DSC_P2P_PROTO=barrier): phases 1 and 3 are grid-wide, system-scope barriers (one flag per peer,
// written into the peer's memory; relaxed stores and polls between two system fences).
// No staging copies on a second stream, no event hand-offs, deterministic.  Measured (7.45 MB,
// tools/allreduce_bench.py): see profiles/README.md.
// NCCL remains for the bootstrap (exchange of the IPC handles), for the Jacobian all-gather, and
// as the all-reduce path when peer access is not available (more than 8 ranks, no P2P, payload
// larger than the symmetric buffer).  Collectives issued through NCCL run on the context's
// communication stream, ordered after the producing kernels by an event; dsc_comm_join() orders
// the compute stream after them.
//
// NCCL is resolved with dlopen at dsc_comm_init time so that the library loads (and every
// single-GPU path runs) on machines without libnccl, and so that a process that already carries
// torch's bundled libnccl.so.2 shares that one copy instead of loading a second one.
#include <dlfcn.h>

#include <type_traits>

#include "dsc_internal.cuh"

namespace dsc {

typedef struct { char internal[128]; } nccl_unique_id;
typedef void* nccl_comm_t;
enum { NCCL_SUCCESS = 0 };
enum { NCCL_UINT8 = 1, NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8 };  // ncclDataType_t values (nccl.h)
enum { NCCL_SUM = 0 };

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(nccl_unique_id*) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl() {
    if (g_nccl.lib) return DSC_OK;
    const char* names[] = {getenv("DSC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* nm : names) {
        if (!nm || !*nm) continue;
        lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) return fail(DSC_E_NCCL, "cannot dlopen libnccl.so.2 (set DSC_NCCL_LIB): %s", dlerror());
#define SYM(field, name)                                                        \
    *(void**)(&g_nccl.field) = dlsym(lib, name);                                \
    if (!g_nccl.field) return fail(DSC_E_NCCL, "libnccl is missing symbol %s", name);
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce")
    SYM(AllGather, "ncclAllGather")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.lib = lib;
    return DSC_OK;
}

#define DSC_NCCL_CHECK(expr)                                                                            \
    do {                                                                                                \
        int _r = (expr);                                                                                \
        if (_r != NCCL_SUCCESS) return fail(DSC_E_NCCL, "%s failed: %s", #expr, g_nccl.GetErrorString(_r)); \
    } while (0)

// order the comm stream after everything queued so far on the compute stream
static int comm_fork(Ctx* ctx) {
    DSC_CUDA_CHECK(cudaEventRecord(ctx->ev_compute, ctx->stream));
    DSC_CUDA_CHECK(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_compute, 0));
    ctx->comm_pending = true;
    return DSC_OK;
}

// Pack / unpack up to kMaxPack buffers into one contiguous staging buffer so that the adjoints of a
// whole model cross NVLink as ONE all-reduce (six separate collectives cost ~15 us of latency each).
constexpr int kMaxPack = 64;
template <typename T>
struct PackTable {
    T* ptr[kMaxPack];
    long long start[kMaxPack + 1];  // prefix offsets in the staging buffer (elements)
    int count;
};
template <typename T, bool PACK>
__global__ void pack_kernel(PackTable<T> t, T* __restrict__ stage) {
    const long long total = t.start[t.count];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = t.count - 1;  // binary search for the owning buffer
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (t.start[mid] <= i) lo = mid; else hi = mid - 1;
        }
        if (PACK) stage[i] = t.ptr[lo][i - t.start[lo]];
        else t.ptr[lo][i - t.start[lo]] = stage[i];
    }
}

// ---- peer-memory all-reduce ---------------------------------------------------------------------
constexpr int kP2PMaxRanks = 8;
constexpr int kP2PMaxCtas = 148;
constexpr int kP2PThreads = 512;
constexpr size_t kP2PHeaderBytes = 4096;              // flags[8] at 0, sync words at 256: [0] CTA counter, [1] epoch, [2] error
constexpr size_t kP2PRegionBytes = size_t(32) << 20;  // region A and region B: 32 MB each

struct P2PArgs {
    char* peer[kP2PMaxRanks];  // every rank's symmetric buffer as mapped here
    int nranks, rank;
    size_t region_bytes;
    unsigned int* err_host;            // mapped pinned host word: set by the kernel when a barrier times out, read by the host WITHOUT a copy
    unsigned long long timeout_ns;     // 0: wait for ever (like NCCL)
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void st_relaxed_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_relaxed_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// All CTAs of this rank have finished the phase -> tell every peer (flag[rank] = value in the peer's
// header) -> wait until every peer has told us.  `arrivals` is the cumulative CTA count expected at
// the local counter.
// A peer that does not show up within the time-out (DSC_P2P_TIMEOUT_S, default 30 s; 0 = wait for ever, like NCCL) turns
// into an ERROR the host sees: the kernel sets a word in mapped pinned host memory and the next dsc_sync / dsc_comm_join /
// dsc_?allreduce_sum / dsc_comm_destroy on this rank returns DSC_E_NCCL (the sums of that step are invalid).  Later
// barriers wait again — nothing is skipped silently.
// Cost matters here (two barriers per all-reduce): ONE system fence per CTA on the way in, the flags of
// all peers written in parallel by the lanes of one warp with relaxed stores (a release store per peer
// is a system fence per peer: measured 78 us per all-reduce at 8 ranks against 37 us at 2), relaxed
// polling, one system fence on the way out.
__device__ __forceinline__ void p2p_barrier(const P2PArgs& a, unsigned int value, unsigned int arrivals, bool reset_counter) {
    unsigned int* flags = reinterpret_cast<unsigned int*>(a.peer[a.rank]);
    unsigned int* sync = reinterpret_cast<unsigned int*>(a.peer[a.rank] + 256);
    __syncthreads();
    if (threadIdx.x < 32) {
        const unsigned lane = threadIdx.x;
        unsigned int last = 0;
        if (lane == 0) {
            // release pattern at GPU scope per CTA; the CTA that arrives last issues the ONE system fence,
            // which is cumulative over everything it observed (the other CTAs' stores, local and peer)
            __threadfence();
            last = atomicAdd(&sync[0], 1u) == arrivals - 1u;
            if (last) {
                if (reset_counter) sync[0] = 0;
                __threadfence_system();
            }
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        __syncwarp();   // lane 0's system fence is ordered before the other lanes' flag stores (a shuffle alone is not a memory barrier)
        if (last && lane < (unsigned)a.nranks) st_relaxed_sys(reinterpret_cast<unsigned int*>(a.peer[lane]) + a.rank, value);
        if (lane < (unsigned)a.nranks) {
            const unsigned long long t0 = globaltimer_ns();
            unsigned int spins = 0;
            while ((int)(ld_relaxed_sys(flags + lane) - value) < 0) {
                if ((++spins & 1023u) == 0 && a.timeout_ns != 0 && globaltimer_ns() - t0 > a.timeout_ns) {
                    sync[2] = 1;
                    if (a.err_host) st_relaxed_sys(a.err_host, 1u);
                    break;
                }
            }
        }
        __syncwarp();
        if (lane == 0) __threadfence_system();            // acquire side: peers' data written before their flags is visible
    }
    __syncthreads();
}

// The buffers of one call laid out in the symmetric regions: every segment starts on a 16-byte
// boundary (the gaps stay zero: regions are zeroed once and gaps are never written with anything but
// sums of zeros), so an aligned source buffer is moved with 128-bit accesses.
template <typename T>
struct P2PTable {
    T* ptr[kMaxPack];
    long long start[kMaxPack + 1];  // element offset of each segment in the region; multiples of 16 / sizeof(T)
    int len[kMaxPack];
    int count;
};

// Region B -> buffers.  Grid-stride per segment, 128-bit when the buffer is aligned.
template <typename T>
__device__ __forceinline__ void p2p_unpack(const P2PTable<T>& t, const T* __restrict__ region, long long gtid, long long gsize) {
    using V = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
    constexpr int W = 16 / (int)sizeof(T);
    for (int s = 0; s < t.count; ++s) {
        T* p = t.ptr[s];
        const T* r = region + t.start[s];
        const long long n = t.len[s];
        long long done = 0;
        if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
            const long long nv = n / W;
            V* pv = reinterpret_cast<V*>(p);
            const V* rv = reinterpret_cast<const V*>(r);
#pragma unroll 4
            for (long long i = gtid; i < nv; i += gsize) pv[i] = __ldcg(rv + i);
            done = nv * W;
        }
        for (long long i = done + gtid; i < n; i += gsize) p[i] = __ldcg(r + i);
    }
}

// Push design: every byte that crosses NVLink is a STORE (fire and forget; the system fence of the
// following barrier waits for them), never a load whose round trip a thread would have to sit out.
//   0. scatter: element e of the packed (16-byte aligned) concatenation of this rank's buffers belongs
//      to slice q = e / slice_len; it is written into region A of rank q, slot [rank] — the buffers
//      are read once, locally, with 128-bit loads, and pack + reduce-scatter exchange are one pass,
//   1. barrier,
//   2. rank r adds the N slots of its region A (local reads) in rank order 0..N-1 — the sum is
//      bit-identical on every rank and from run to run — and stores the result slice into region B of
//      EVERY rank (all-gather by stores),
//   3. barrier,
//   4. region B (local) -> buffers.
template <typename T>
__global__ void __launch_bounds__(kP2PThreads) allreduce_p2p_kernel(P2PTable<T> t, P2PArgs a) {
    using V = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
    constexpr int W = 16 / (int)sizeof(T);
    char* local = a.peer[a.rank];
    unsigned int* sync = reinterpret_cast<unsigned int*>(local + 256);
    const unsigned int epoch = sync[1];   // bumped by CTA 0 after the second barrier: every CTA has read it by then
    const long long total = t.start[t.count];   // multiple of W
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gsize = (long long)gridDim.x * blockDim.x;
    const long long nvec = total / W;
    const long long per = (nvec + a.nranks - 1) / a.nranks;   // vectors per slice (the last slice may be short)
    unsigned long long* stamps = reinterpret_cast<unsigned long long*>(local + 512 + 64 * (epoch & 1u));   // phase boundaries as CTA 0 sees them, of the last two calls (dsc_comm_debug_stamps)
    const bool stamp = blockIdx.x == 0 && threadIdx.x == 0;
    if (stamp) stamps[0] = globaltimer_ns();
    // 0. scatter this rank's contribution: slice q -> rank q, region A, slot [rank].
    // Every rank walks the packed space starting at a different slice (rank r begins with slice r + 1), so that
    // at any moment the N ranks write into N different destinations instead of all into the same one.
    for (int s = 0; s < t.count; ++s) {
        const T* p = t.ptr[s];
        const long long v_seg = t.start[s] / W;   // first vector of the segment in the packed space
        const long long n = t.len[s];
        long long done = 0;
        if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
            const long long nv = n / W;
            const V* pv = reinterpret_cast<const V*>(p);
            // rotation of the segment's vector index: position j -> vector (j + rot) mod nv, rot = the offset inside the
            // segment of the first vector that belongs to slice rank + 1 (clamped to the segment)
            long long rot = (long long)((a.rank + 1) % a.nranks) * per - v_seg;
            rot = rot < 0 ? 0 : (rot >= nv ? 0 : rot);
#pragma unroll 4
            for (long long j = gtid; j < nv; j += gsize) {
                const long long i = j + rot < nv ? j + rot : j + rot - nv;
                const V x = pv[i];
                const long long v = v_seg + i;
                const int q = (int)(v / per);
                reinterpret_cast<V*>(a.peer[q] + kP2PHeaderBytes)[(long long)a.rank * per + (v - q * per)] = x;
            }
            done = nv * W;
        }
        for (long long i = done + gtid; i < n; i += gsize) {   // unaligned buffer, or the tail of a segment
            const long long e = t.start[s] + i;
            const long long v = e / W;
            const int q = (int)(v / per);
            reinterpret_cast<T*>(a.peer[q] + kP2PHeaderBytes)[((long long)a.rank * per + (v - q * per)) * W + (e - v * W)] = p[i];
        }
    }
    if (stamp) stamps[1] = globaltimer_ns();
    // 1. every rank's contribution to our slice has landed
    p2p_barrier(a, 2u * epoch + 1u, gridDim.x, false);
    if (stamp) stamps[2] = globaltimer_ns();
    // 2. our slice of the sum (slots added in rank order), stored into region B of every rank
    {
        const V* slots = reinterpret_cast<const V*>(local + kP2PHeaderBytes);
        const long long v0 = per * a.rank;
        const long long mine = max(0ll, min(per, nvec - v0));
        for (long long i = gtid; i < mine; i += gsize) {
            V x[kP2PMaxRanks];
#pragma unroll
            for (int r = 0; r < kP2PMaxRanks; ++r)
                if (r < a.nranks) x[r] = __ldcg(slots + (long long)r * per + i);
            V acc = x[0];
#pragma unroll
            for (int r = 1; r < kP2PMaxRanks; ++r)
                if (r < a.nranks) {
                    acc.x += x[r].x; acc.y += x[r].y;
                    if constexpr (W == 4) { acc.z += x[r].z; acc.w += x[r].w; }
                }
            // destinations in rotated order (rank + 1, rank + 2, ...): no two ranks target the same peer at the same step
#pragma unroll
            for (int j = 0; j < kP2PMaxRanks; ++j)
                if (j < a.nranks) {
                    int q = a.rank + 1 + j;
                    q = q >= a.nranks ? q - a.nranks : q;
                    reinterpret_cast<V*>(a.peer[q] + kP2PHeaderBytes + a.region_bytes)[v0 + i] = acc;
                }
        }
    }
    if (stamp) stamps[3] = globaltimer_ns();
    // 3. every slice has landed everywhere
    p2p_barrier(a, 2u * epoch + 2u, 2u * gridDim.x, true);
    if (stamp) stamps[4] = globaltimer_ns();
    // 4. unpack
    p2p_unpack<T>(t, reinterpret_cast<const T*>(local + kP2PHeaderBytes + a.region_bytes), gtid, gsize);
    if (stamp) stamps[5] = globaltimer_ns();
    // the CTA counter was reset by the last arriver of barrier 3; the epoch is bumped by CTA 0 after
    // ITS threads left that barrier (all CTAs read the epoch before arriving at barrier 1)
    if (blockIdx.x == 0 && threadIdx.x == 0) sync[1] = epoch + 1u;
}

// ---- flag-in-data all-reduce (the default) ----------------------------------------------------------------------------
// The barrier kernel above pays two grid-wide, system-scope barriers per call: ~8 us each on an NVSwitch node even when
// the data phases take 2 us (profiles/README.md), i.e. 16 of the 23 us of a 3 MB all-reduce.  This kernel has no barrier.
// Every 128-byte line that crosses NVLink carries 120 bytes of payload and an 8-byte flag (the call's sequence number) in
// its last word; the eight lanes of a line group store it with one coalesced 128-byte store, which the fabric delivers
// whole (the property NCCL's LL128 protocol rests on), and the receiver polls the line itself until the flag is the one it
// waits for — the same load that sees the flag has the payload.  Same data flow as the barrier kernel (scatter slices,
// add the N slots in rank order 0..N-1 — bit-identical on every rank and from run to run —, broadcast, unpack), but the
// phases of different warps overlap and the only latency on the path is the one-way flight of a line.
//   region A of rank q : N slots x lpr lines (slot s = the lines of slice q sent by rank s)
//   region B           : the TL result lines
// Re-use across calls needs no barrier either: a rank can only have finished call k (and start writing call k + 1) after
// it received every result line of call k, hence after every peer has read all slots of call k; and nobody can produce a
// result line of call k + 1 before every rank has sent its call-k+1 slots, i.e. has finished unpacking call k.
constexpr int kLLPayloadWords = 15;   // 8-byte words of payload per 128-byte line

__device__ __forceinline__ void st_line16(void* p, unsigned long long a, unsigned long long b) {
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void ld_line16(const void* p, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

// The segment (buffer) of the call that packed 8-byte word w falls into, kept in registers: the walk over the packed space
// is monotone, so the table in parameter space is searched only when a line crosses into another buffer.
template <typename T>
struct LLCursor {
    long long e0 = 0, e1 = 0;   // element range [e0, e1) of the current segment in the packed space (e1 = start of the next one)
    long long len = 0;          // valid elements (the rest up to e1 is alignment gap)
    T* ptr = nullptr;
    bool al8 = false;           // ptr is 8-byte aligned: a fully valid word moves with one 8-byte access
    __device__ __forceinline__ void seek(const P2PTable<T>& t, long long e) {
        if (e >= e0 && e < e1) return;
        int lo = 0, hi = t.count - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (t.start[mid] <= e) lo = mid; else hi = mid - 1;
        }
        e0 = t.start[lo];
        e1 = t.start[lo + 1];
        len = t.len[lo];
        ptr = t.ptr[lo];
        al8 = (reinterpret_cast<uintptr_t>(ptr) & 7) == 0;
    }
    // pointer to the first element of word w and the number of valid elements there (0 past the end / in a gap)
    __device__ __forceinline__ T* locate(const P2PTable<T>& t, long long w, long long total, int& valid) {
        constexpr int EPW = 8 / (int)sizeof(T);
        const long long e = w * EPW;
        valid = 0;
        if (e >= total) return nullptr;
        seek(t, e);
        const long long i = e - e0, left = len - i;
        valid = left <= 0 ? 0 : (left >= EPW ? EPW : (int)left);
        return ptr + i;
    }
};
template <typename T>
__device__ __forceinline__ unsigned long long ll_load(const T* p, int valid, bool al8) {
    if constexpr (sizeof(T) == 8) {
        return valid > 0 ? (unsigned long long)__double_as_longlong(*p) : 0ull;
    } else {
        if (al8 && valid == 2) return *reinterpret_cast<const unsigned long long*>(p);
        const unsigned int a = valid > 0 ? __float_as_uint(p[0]) : 0u;
        const unsigned int b = valid > 1 ? __float_as_uint(p[1]) : 0u;
        return (unsigned long long)a | ((unsigned long long)b << 32);
    }
}
template <typename T>
__device__ __forceinline__ void ll_store(T* p, int valid, bool al8, unsigned long long v) {
    if constexpr (sizeof(T) == 8) {
        if (valid > 0) *p = __longlong_as_double((long long)v);
    } else {
        if (al8 && valid == 2) { *reinterpret_cast<unsigned long long*>(p) = v; return; }
        if (valid > 0) p[0] = __uint_as_float((unsigned int)v);
        if (valid > 1) p[1] = __uint_as_float((unsigned int)(v >> 32));
    }
}
template <typename T>
__device__ __forceinline__ unsigned long long ll_add(unsigned long long x, unsigned long long y) {
    if constexpr (sizeof(T) == 8) {
        return (unsigned long long)__double_as_longlong(__longlong_as_double((long long)x) + __longlong_as_double((long long)y));
    } else {
        const float a = __uint_as_float((unsigned int)x) + __uint_as_float((unsigned int)y);
        const float b = __uint_as_float((unsigned int)(x >> 32)) + __uint_as_float((unsigned int)(y >> 32));
        return (unsigned long long)__float_as_uint(a) | ((unsigned long long)__float_as_uint(b) << 32);
    }
}

// Poll U line groups (16 bytes per lane each) until the flag lane of every active group reads `seq`.  All 32 lanes take
// part; the U loads are issued back to back, so U lines are in flight per lane, and only the lines still missing are
// re-read.  Returns false after the time-out (the error word is set; the caller carries on with what it has).
template <int U>
__device__ __forceinline__ bool ll_wait_lines(const char* const (&p)[U], const bool (&active)[U], bool flag_lane, unsigned long long seq, const P2PArgs& a,
                                              unsigned int* sync, unsigned long long (&x0)[U], unsigned long long (&x1)[U]) {
    bool ready[U];   // warp-uniform
#pragma unroll
    for (int u = 0; u < U; ++u) ready[u] = false;
    unsigned int spins = 0;
    unsigned long long t0 = 0;
    for (;;) {
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (!ready[u] && active[u]) ld_line16(p[u], x0[u], x1[u]);
        bool all = true;
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (!ready[u]) {
                ready[u] = __all_sync(0xffffffffu, !active[u] || !flag_lane || x1[u] == seq);
                all = all && ready[u];
            }
        if (all) return true;
        if ((++spins & 1023u) == 0 && a.timeout_ns != 0) {
            if (t0 == 0) t0 = globaltimer_ns();
            else if (globaltimer_ns() - t0 > a.timeout_ns) {
                sync[2] = 1;
                if (a.err_host) st_relaxed_sys(a.err_host, 1u);
                return false;
            }
        }
    }
}

template <typename T, int U>   // U: line groups in flight per lane in the scatter and unpack phases
__global__ void __launch_bounds__(kP2PThreads) allreduce_ll_kernel(P2PTable<T> t, P2PArgs a) {
    constexpr int EPW = 8 / (int)sizeof(T);
    char* local = a.peer[a.rank];
    unsigned int* sync = reinterpret_cast<unsigned int*>(local + 256);
    const unsigned int epoch = sync[1];                        // bumped by the last CTA to leave: every CTA has read it by then
    const unsigned long long seq = (unsigned long long)epoch + 1ull;   // never 0: the regions start out zeroed
    const long long total_words = t.start[t.count] / EPW;      // segments start on 16-byte boundaries: a whole number of words
    const long long TL = (total_words + kLLPayloadWords - 1) / kLLPayloadWords;
    const long long lpr = (TL + a.nranks - 1) / a.nranks;      // lines per slice
    const long long span = lpr * a.nranks;
    const int lane = threadIdx.x & 31, sub = lane & 7, grp = lane >> 3;
    const bool flag_lane = sub == 7;
    const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    unsigned long long* stamps = reinterpret_cast<unsigned long long*>(local + 512 + 64 * (epoch & 1u));
    const bool stamp = blockIdx.x == 0 && threadIdx.x == 0;
    if (stamp) stamps[0] = globaltimer_ns();
    // 0. scatter: line g of the packed buffers -> rank g / lpr, region A, slot [rank].  Every rank starts its walk at a
    //    different slice (rank r with slice r + 1): at any moment the N ranks write into N different destinations.
    const long long total_e = t.start[t.count];
    LLCursor<T> cur;
    for (long long g0 = warp_id * 4 * U; g0 < span; g0 += nwarps * 4 * U) {   // 4 U consecutive lines per warp and round
        unsigned long long w0[U], w1[U];
        char* dst[U];
        const T* s0[U];
        const T* s1[U];
        int v0[U], v1[U];
        bool a0[U], a1[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {   // addresses first ...
            const long long gi = g0 + u * 4 + grp;
            long long g = gi + (long long)((a.rank + 1) % a.nranks) * lpr;
            if (g >= span) g -= span;
            dst[u] = nullptr;
            s0[u] = s1[u] = nullptr;
            v0[u] = v1[u] = 0;
            a0[u] = a1[u] = false;
            if (gi < span && g < TL) {
                const int q = (int)(g / lpr);
                const long long w = g * kLLPayloadWords + sub * 2;
                s0[u] = cur.locate(t, w, total_e, v0[u]);
                a0[u] = cur.al8;
                if (!flag_lane) {
                    s1[u] = cur.locate(t, w + 1, total_e, v1[u]);
                    a1[u] = cur.al8;
                }
                dst[u] = a.peer[q] + kP2PHeaderBytes + (((long long)a.rank * lpr + (g - q * lpr)) << 7) + sub * 16;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {   // ... then every load of the round (2 U words in flight per lane) ...
            w0[u] = ll_load<T>(s0[u], v0[u], a0[u]);
            w1[u] = flag_lane ? seq : ll_load<T>(s1[u], v1[u], a1[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)     // ... then the stores
            if (dst[u]) st_line16(dst[u], w0[u], w1[u]);
    }
    if (stamp) stamps[1] = stamps[2] = globaltimer_ns();
    // 2. our slice: wait for the line of every rank (all N in flight), add the N lines in rank order, store the result line
    //    into region B of every rank (destinations in rotated order)
    {
        const long long first = lpr * a.rank;
        const long long mine = max(0ll, min(lpr, TL - first));
        const char* slots = local + kP2PHeaderBytes;
        bool alive = true;
        for (long long l0 = warp_id * 4; l0 < mine && alive; l0 += nwarps * 4) {
            const long long loc = l0 + grp;
            const bool active = loc < mine;
            const char* src[kP2PMaxRanks];
            bool act[kP2PMaxRanks];
            unsigned long long x0[kP2PMaxRanks], x1[kP2PMaxRanks];
#pragma unroll
            for (int r = 0; r < kP2PMaxRanks; ++r) {
                act[r] = active && r < a.nranks;
                src[r] = slots + (((long long)r * lpr + loc) << 7) + sub * 16;
                x0[r] = x1[r] = 0;
            }
            alive = ll_wait_lines<kP2PMaxRanks>(src, act, flag_lane, seq, a, sync, x0, x1);
            if (active) {
                unsigned long long acc0 = x0[0], acc1 = x1[0];
#pragma unroll
                for (int r = 1; r < kP2PMaxRanks; ++r)
                    if (r < a.nranks) {
                        acc0 = ll_add<T>(acc0, x0[r]);
                        if (!flag_lane) acc1 = ll_add<T>(acc1, x1[r]);
                    }
                if (flag_lane) acc1 = seq;
#pragma unroll
                for (int j = 0; j < kP2PMaxRanks; ++j)
                    if (j < a.nranks) {
                        int q = a.rank + 1 + j;
                        q = q >= a.nranks ? q - a.nranks : q;
                        st_line16(a.peer[q] + kP2PHeaderBytes + a.region_bytes + ((first + loc) << 7) + sub * 16, acc0, acc1);
                    }
            }
        }
    }
    if (stamp) stamps[3] = stamps[4] = globaltimer_ns();
    // 4. unpack: every result line, as it arrives, into the buffers
    {
        const char* res = local + kP2PHeaderBytes + a.region_bytes;
        bool alive = true;
        for (long long g0 = warp_id * 4 * U; g0 < TL && alive; g0 += nwarps * 4 * U) {
            const char* src[U];
            bool act[U];
            unsigned long long x0[U], x1[U];
            T* d0[U];
            T* d1[U];
            int v0[U], v1[U];
            bool a0[U], a1[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {   // destinations first: the lines may still be on their way
                const long long g = g0 + u * 4 + grp;
                act[u] = g < TL;
                src[u] = res + (g << 7) + sub * 16;
                x0[u] = x1[u] = 0;
                d0[u] = d1[u] = nullptr;
                v0[u] = v1[u] = 0;
                a0[u] = a1[u] = false;
                if (act[u]) {
                    const long long w = g * kLLPayloadWords + sub * 2;
                    d0[u] = cur.locate(t, w, total_e, v0[u]);
                    a0[u] = cur.al8;
                    if (!flag_lane) {
                        d1[u] = cur.locate(t, w + 1, total_e, v1[u]);
                        a1[u] = cur.al8;
                    }
                }
            }
            alive = ll_wait_lines<U>(src, act, flag_lane, seq, a, sync, x0, x1);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                ll_store<T>(d0[u], v0[u], a0[u], x0[u]);
                ll_store<T>(d1[u], v1[u], a1[u], x1[u]);
            }
        }
    }
    if (stamp) stamps[5] = globaltimer_ns();
    // the last CTA to leave bumps the epoch (every CTA read it on entry) and resets the departure counter
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&sync[0], 1u) == gridDim.x - 1u) {
            sync[0] = 0;
            sync[1] = epoch + 1u;
        }
    }
}

static int p2p_setup(Ctx* ctx);

// Has a barrier of an earlier peer-memory all-reduce timed out on this rank?  (mapped pinned word: no copy, no synchronisation)
static int p2p_check_error(Ctx* ctx) {
    if (ctx->p2p_err_host && *reinterpret_cast<volatile unsigned int*>(ctx->p2p_err_host) != 0)
        return fail(DSC_E_NCCL, "a peer-memory all-reduce timed out waiting for a peer (the sums of that step are invalid; DSC_P2P_TIMEOUT_S sets the limit)");
    return DSC_OK;
}

// The all-reduce kernel meets at grid-wide barriers: every CTA must be resident.  A cooperative launch makes the driver
// guarantee that (or fail the launch) instead of relying on an idle GPU; the grid is sized from the occupancy query.
// Does a payload of `bytes` (16-byte padded segments) fit the symmetric regions?  (flag-in-data lines carry 120 of 128 bytes;
// the slices are rounded up to whole lines per rank)
static bool p2p_fits(const Ctx* ctx, size_t bytes) {
    return bytes + bytes / 15 + 128 * (size_t)(kP2PMaxRanks + 2) <= ctx->p2p_region_bytes;
}

// DSC_P2P_PROTO=barrier selects the two-barrier kernel (every rank must use the same protocol); default: flag in data
static bool p2p_use_ll() {
    static const bool ll = [] { const char* e = getenv("DSC_P2P_PROTO"); return !(e && (*e == 'b' || *e == 'B')); }();
    return ll;
}

template <typename T>
static int p2p_launch(Ctx* ctx, P2PTable<T>& pt, P2PArgs& a) {
    // one line group in flight per lane: measured on 2 GPUs (7.45 MB), 2 and 4 in flight were slower (23.2 / 25.5 us against
    // 20.9 us): the phases are bound by the NVLink writes outstanding per SM, not by the loads of a lane
    void* const kernel = p2p_use_ll() ? reinterpret_cast<void*>(allreduce_ll_kernel<T, 1>) : reinterpret_cast<void*>(allreduce_p2p_kernel<T>);
    static int max_ctas = -1;
    if (max_ctas < 0) {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, p2p_use_ll() ? allreduce_ll_kernel<T, 1> : allreduce_p2p_kernel<T>, kP2PThreads, 0) != cudaSuccess || per_sm < 1) {
            cudaGetLastError();
            per_sm = 1;
        }
        max_ctas = per_sm * ctx->sm_count;
    }
    int ctas = ctx->sm_count < kP2PMaxCtas ? ctx->sm_count : kP2PMaxCtas;
    if (ctas > max_ctas) ctas = max_ctas;
    static int coop = -1;   // -1 unknown, 1 works, 0 the driver refused (e.g. under this capture mode): plain launch
    if (coop != 0) {
        void* args[] = {&pt, &a};
        cudaError_t e = cudaLaunchCooperativeKernel(kernel, dim3(ctas), dim3(kP2PThreads), args, 0, ctx->stream);
        if (e == cudaSuccess) {
            coop = 1;
            DSC_LAUNCH_CHECK(ctx);
            return DSC_OK;
        }
        cudaGetLastError();
        if (coop == 1) return fail(DSC_E_CUDA, "cooperative launch of the all-reduce kernel failed: %s", cudaGetErrorString(e));
        coop = 0;
    }
    {
        void* args[] = {&pt, &a};
        cudaError_t e = cudaLaunchKernel(kernel, dim3(ctas), dim3(kP2PThreads), args, 0, ctx->stream);
        if (e != cudaSuccess) return fail(DSC_E_CUDA, "launch of the all-reduce kernel failed: %s", cudaGetErrorString(e));
    }
    DSC_LAUNCH_CHECK(ctx);
    return DSC_OK;
}

template <typename T>
static int allreduce_impl(dsc_ctx_t c, int32_t count, const dsc_buf_t* bufs) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    if (count < 0 || (count > 0 && !bufs)) return fail(DSC_E_INVALID, "allreduce: bad buffer list");
    if (ctx->nranks == 1) return DSC_OK;  // a single rank already holds the sum
    if (!ctx->nccl_comm) return fail(DSC_E_NCCL, "allreduce: dsc_comm_init has not been called");
    DSC_TRY(p2p_check_error(ctx));
    DSC_TRY(co_schedule_pending_gemms(ctx));   // the weight adjoints about to be summed are usually still-deferred products: run them side by side
    DSC_TRY(flush_colsums_on_lane(ctx));       // ... and deferred bias column sums belong beside the last weight adjoint, not behind it
    // The collective runs on the compute stream, on every SM: measured (tools/step_timeline.py, 2 GPUs), an all-reduce confined
    // to the few SMs a concurrent GEMM leaves free pushes ~5 GB/s per SM over NVLink (outstanding remote writes per SM), i.e.
    // 40-46 us for the 4.2 MB of the later layers on 32 SMs against 17 us on all of them — running it beside the first layer's
    // backward products (and the first layer's own all-reduce behind it) lost to ONE all-reduce of everything after the sweep.
    DSC_TRY(lane_join(ctx));
    PackTable<T> t;
    t.count = 0;
    t.start[0] = 0;
    int st = DSC_OK;
    for (int i = 0; i < count && st == DSC_OK; ++i) {
        Buf* b;
        st = in_buf(ctx, dtype_of<T>::value, bufs[i], &b, "allreduce: buffer");
        if (st != DSC_OK || b->len == 0) continue;
        st = buf_writable(ctx, b);  // in-place collective
        if (st != DSC_OK) break;
        if (t.count == kMaxPack) { st = fail(DSC_E_INVALID, "allreduce: more than %d buffers in one call", kMaxPack); break; }
        t.ptr[t.count] = b->ptr<T>();
        t.start[t.count + 1] = t.start[t.count] + b->len;
        t.count++;
    }
    if (st != DSC_OK || t.count == 0) return st;
    const int dt = sizeof(T) == 4 ? NCCL_FLOAT32 : NCCL_FLOAT64;
    const long long total = t.start[t.count];
    if (ctx->p2p_ok && p2p_fits(ctx, (size_t)(total + 4 * t.count) * sizeof(T))) {
        // one kernel over peer memory on the compute stream (ordered after the producers by stream order)
        if (ctx->comm_pending) {   // a collective issued through NCCL may still be writing these buffers
            DSC_CUDA_CHECK(cudaEventRecord(ctx->ev_comm, ctx->comm_stream));
            DSC_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0));
        }
        constexpr long long W = 16 / (long long)sizeof(T);
        P2PTable<T> pt;
        pt.count = t.count;
        pt.start[0] = 0;
        for (int i = 0; i < t.count; ++i) {
            pt.ptr[i] = t.ptr[i];
            pt.len[i] = (int)(t.start[i + 1] - t.start[i]);
            pt.start[i + 1] = pt.start[i] + (pt.len[i] + W - 1) / W * W;
        }
        P2PArgs a;
        for (int r = 0; r < kP2PMaxRanks; ++r) a.peer[r] = reinterpret_cast<char*>(ctx->p2p_peer[r]);
        a.nranks = ctx->nranks;
        a.rank = ctx->rank;
        a.region_bytes = ctx->p2p_region_bytes;
        a.err_host = reinterpret_cast<unsigned int*>(ctx->p2p_err_dev);
        a.timeout_ns = ctx->p2p_timeout_ns;
        st = p2p_launch<T>(ctx, pt, a);
        if (st == DSC_OK) ctx->stats.allreduce_p2p++;
        return st;
    }
    ctx->stats.allreduce_nccl++;
    for (int i = 0; i < count; ++i) {   // these blocks are now in use on the communication stream: not recycled / overwritten before it is joined
        Buf* b = as_buf(bufs[i]);
        if (b && b->block) note_other_stream_use(ctx, b->block);
    }
    if (t.count == 1) {
        DSC_TRY(comm_fork(ctx));
        DSC_NCCL_CHECK(g_nccl.AllReduce(t.ptr[0], t.ptr[0], (size_t)total, dt, NCCL_SUM, ctx->nccl_comm, ctx->comm_stream));
        ctx->stats.kernel_launches++;
        return DSC_OK;
    }
    // pack (compute stream) -> one all-reduce (comm stream) -> unpack (comm stream); dsc_comm_join orders readers
    const size_t bytes = (size_t)total * sizeof(T);
    if (ctx->comm_stage_bytes < bytes) {
        if (ctx->capturing || ctx->live_graphs > 0)
            return fail(DSC_E_INVALID, "all-reduce staging buffer cannot grow while a CUDA graph is being captured or is alive");
        if (ctx->comm_stage) { cudaStreamSynchronize(ctx->comm_stream); cudaFree(ctx->comm_stage); }
        ctx->comm_stage = nullptr;
        ctx->comm_stage_bytes = 0;
        DSC_CUDA_CHECK(cudaMalloc(&ctx->comm_stage, bytes));
        ctx->comm_stage_bytes = bytes;
    }
    T* stage = reinterpret_cast<T*>(ctx->comm_stage);
    const int grid = grid_for((size_t)total, 256 * 4, ctx->sm_count * 4);
    if (ctx->comm_pending) {  // the staging buffer may still be in use by a previous collective
        DSC_CUDA_CHECK(cudaEventRecord(ctx->ev_comm, ctx->comm_stream));
        DSC_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0));
    }
    pack_kernel<T, true><<<grid, 256, 0, ctx->stream>>>(t, stage);
    DSC_LAUNCH_CHECK(ctx);
    DSC_TRY(comm_fork(ctx));
    DSC_NCCL_CHECK(g_nccl.AllReduce(stage, stage, (size_t)total, dt, NCCL_SUM, ctx->nccl_comm, ctx->comm_stream));
    ctx->stats.kernel_launches++;
    pack_kernel<T, false><<<grid, 256, 0, ctx->comm_stream>>>(t, stage);
    DSC_LAUNCH_CHECK(ctx);
    return DSC_OK;
}

template <typename T>
static int allgather_impl(dsc_ctx_t c, dsc_buf_t sendh, dsc_buf_t recvh) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *s, *r;
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, sendh, &s, "allgather: send"));
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, recvh, &r, "allgather: recv"));
    if ((int64_t)s->len * ctx->nranks != r->len)
        return fail(DSC_E_SHAPE, "allgather: recv length %d != %d ranks x send length %d", r->len, ctx->nranks, s->len);
    if (s->len == 0) return DSC_OK;
    DSC_TRY(buf_writable(ctx, r));
    if (ctx->nranks == 1) {
        DSC_CUDA_CHECK(cudaMemcpyAsync(r->ptr<T>(), s->ptr<T>(), (size_t)s->len * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
        return DSC_OK;
    }
    if (!ctx->nccl_comm) return fail(DSC_E_NCCL, "allgather: dsc_comm_init has not been called");
    note_other_stream_use(ctx, s->block);
    note_other_stream_use(ctx, r->block);
    DSC_TRY(comm_fork(ctx));
    const int dt = sizeof(T) == 4 ? NCCL_FLOAT32 : NCCL_FLOAT64;
    DSC_NCCL_CHECK(g_nccl.AllGather(s->ptr<T>(), r->ptr<T>(), (size_t)s->len, dt, ctx->nccl_comm, ctx->comm_stream));
    ctx->stats.kernel_launches++;
    return DSC_OK;
}

// Map every rank's symmetric buffer into this process.  Collective: every rank must call it; the
// outcome (p2p_ok) is agreed through two all-gathers so that no rank takes the peer path alone.
static int p2p_setup(Ctx* ctx) {
    ctx->p2p_ok = false;
    const char* off = getenv("DSC_P2P");
    int ok = (ctx->nranks >= 2 && ctx->nranks <= kP2PMaxRanks && !(off && *off == '0')) ? 1 : 0;
    int ndev = 0;
    if (ok && (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < ctx->nranks)) ok = 0;
    const size_t bytes = kP2PHeaderBytes + 2 * kP2PRegionBytes;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (ok) {
        if (cudaMalloc(&ctx->p2p_local, bytes) != cudaSuccess) { cudaGetLastError(); ctx->p2p_local = nullptr; ok = 0; }
    }
    if (ok) {
        if (cudaMemset(ctx->p2p_local, 0, bytes) != cudaSuccess || cudaIpcGetMemHandle(&mine, ctx->p2p_local) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    }
    // exchange (handle, ok) through NCCL: 64-byte handle + 4-byte flag per rank
    constexpr int kRec = 128;
    unsigned char* dev = nullptr;
    DSC_CUDA_CHECK(cudaMalloc(&dev, (size_t)kRec * (ctx->nranks + 1)));
    unsigned char host[kRec * (kP2PMaxRanks + 64)];
    if (ctx->nranks > kP2PMaxRanks + 63) { cudaFree(dev); return DSC_OK; }
    memset(host, 0, sizeof(host));
    memcpy(host, &mine, sizeof(mine));
    host[64] = (unsigned char)ok;
    DSC_CUDA_CHECK(cudaMemcpyAsync(dev, host, kRec, cudaMemcpyHostToDevice, ctx->comm_stream));
    DSC_NCCL_CHECK(g_nccl.AllGather(dev, dev + kRec, kRec, NCCL_UINT8, ctx->nccl_comm, ctx->comm_stream));
    DSC_CUDA_CHECK(cudaMemcpyAsync(host, dev + kRec, (size_t)kRec * ctx->nranks, cudaMemcpyDeviceToHost, ctx->comm_stream));
    DSC_CUDA_CHECK(cudaStreamSynchronize(ctx->comm_stream));
    for (int r = 0; r < ctx->nranks; ++r) ok = ok && host[r * kRec + 64];
    if (ok) {
        for (int r = 0; r < ctx->nranks && ok; ++r) {
            if (r == ctx->rank) { ctx->p2p_peer[r] = ctx->p2p_local; continue; }
            cudaIpcMemHandle_t h;
            memcpy(&h, host + r * kRec, sizeof(h));
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
            ctx->p2p_peer[r] = ptr;
        }
    }
    // second agreement round: did every rank manage to open every peer?
    host[0] = (unsigned char)ok;
    DSC_CUDA_CHECK(cudaMemcpyAsync(dev, host, kRec, cudaMemcpyHostToDevice, ctx->comm_stream));
    DSC_NCCL_CHECK(g_nccl.AllGather(dev, dev + kRec, kRec, NCCL_UINT8, ctx->nccl_comm, ctx->comm_stream));
    DSC_CUDA_CHECK(cudaMemcpyAsync(host, dev + kRec, (size_t)kRec * ctx->nranks, cudaMemcpyDeviceToHost, ctx->comm_stream));
    DSC_CUDA_CHECK(cudaStreamSynchronize(ctx->comm_stream));
    for (int r = 0; r < ctx->nranks; ++r) ok = ok && host[r * kRec];
    cudaFree(dev);
    ctx->p2p_ok = ok != 0;
    ctx->p2p_region_bytes = ok ? kP2PRegionBytes : 0;
    if (ok && !ctx->p2p_err_host) {
        if (cudaHostAlloc(&ctx->p2p_err_host, 64, cudaHostAllocMapped) == cudaSuccess) {
            memset(ctx->p2p_err_host, 0, 64);
            if (cudaHostGetDevicePointer(&ctx->p2p_err_dev, ctx->p2p_err_host, 0) != cudaSuccess) { cudaGetLastError(); ctx->p2p_err_dev = nullptr; }
        } else {
            cudaGetLastError();
            ctx->p2p_err_host = nullptr;
        }
    }
    {
        const char* e = getenv("DSC_P2P_TIMEOUT_S");
        const double sec = e ? atof(e) : 30.0;
        ctx->p2p_timeout_ns = sec <= 0 ? 0ull : (unsigned long long)(sec * 1e9);
    }
    return DSC_OK;
}

static void p2p_teardown(Ctx* ctx) {
    for (int r = 0; r < kP2PMaxRanks; ++r) {
        if (ctx->p2p_peer[r] && ctx->p2p_peer[r] != ctx->p2p_local) cudaIpcCloseMemHandle(ctx->p2p_peer[r]);
        ctx->p2p_peer[r] = nullptr;
    }
    if (ctx->p2p_local) cudaFree(ctx->p2p_local);
    ctx->p2p_local = nullptr;
    if (ctx->p2p_err_host) cudaFreeHost(ctx->p2p_err_host);
    ctx->p2p_err_host = ctx->p2p_err_dev = nullptr;
    ctx->p2p_ok = false;
    ctx->p2p_region_bytes = 0;
}

}  // namespace dsc

using namespace dsc;

extern "C" {

int dsc_comm_unique_id(uint8_t out[128]) {
    if (!out) return fail(DSC_E_INVALID, "null id buffer");
    DSC_TRY(load_nccl());
    nccl_unique_id id;
    DSC_NCCL_CHECK(g_nccl.GetUniqueId(&id));
    memcpy(out, id.internal, 128);
    return DSC_OK;
}

int dsc_comm_init(dsc_ctx_t c, int nranks, int rank, const uint8_t idbytes[128]) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(DSC_E_INVALID, "comm_init: rank %d of %d", rank, nranks);
    if (ctx->nccl_comm) return fail(DSC_E_INVALID, "comm_init: communicator already initialised");
    ctx->nranks = nranks;
    ctx->rank = rank;
    if (nranks == 1) return DSC_OK;
    if (!idbytes) return fail(DSC_E_INVALID, "comm_init: null unique id");
    DSC_TRY(load_nccl());
    DSC_CUDA_CHECK(cudaSetDevice(ctx->device));
    nccl_unique_id id;
    memcpy(id.internal, idbytes, 128);
    nccl_comm_t comm = nullptr;
    DSC_NCCL_CHECK(g_nccl.CommInitRank(&comm, nranks, id, rank));
    ctx->nccl_comm = comm;
    return p2p_setup(ctx);
}

int dsc_comm_debug_stamps_at(dsc_ctx_t c, int back, uint64_t out[6]) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    if (!ctx->p2p_local) return fail(DSC_E_NCCL, "the peer-memory all-reduce is not set up");
    if (back < 0 || back > 1 || !out) return fail(DSC_E_INVALID, "dsc_comm_debug_stamps_at: back must be 0 (the last call) or 1 (the one before)");
    DSC_TRY(flush_pending(ctx));
    DSC_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    unsigned int epoch = 0;   // the NEXT call's epoch: the last one ran with epoch - 1
    DSC_CUDA_CHECK(cudaMemcpy(&epoch, reinterpret_cast<char*>(ctx->p2p_local) + 256 + 4, sizeof(epoch), cudaMemcpyDeviceToHost));
    const unsigned int slot = (epoch - 1u - (unsigned)back) & 1u;
    DSC_CUDA_CHECK(cudaMemcpy(out, reinterpret_cast<char*>(ctx->p2p_local) + 512 + 64 * slot, 6 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return DSC_OK;
}

int dsc_comm_debug_stamps(dsc_ctx_t c, uint64_t out[6]) { return dsc_comm_debug_stamps_at(c, 0, out); }

int dsc_comm_join(dsc_ctx_t c) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    DSC_TRY(p2p_check_error(ctx));
    DSC_TRY(lane_join(ctx));   // "the compute stream is ordered after every side stream": the second compute lane included
    if (!ctx->comm_pending) return DSC_OK;
    DSC_CUDA_CHECK(cudaEventRecord(ctx->ev_comm, ctx->comm_stream));
    DSC_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0));
    ctx->comm_pending = false;
    return DSC_OK;
}

int dsc_comm_destroy(dsc_ctx_t c) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    unsigned int p2p_err = 0;
    if (ctx->nccl_comm) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(ctx->comm_stream);
        if (ctx->p2p_local) cudaMemcpy(&p2p_err, reinterpret_cast<char*>(ctx->p2p_local) + 256 + 8, 4, cudaMemcpyDeviceToHost);
        if (ctx->p2p_err_host && *reinterpret_cast<volatile unsigned int*>(ctx->p2p_err_host)) p2p_err = 1;
        p2p_teardown(ctx);
        g_nccl.CommDestroy(ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    ctx->nranks = 1;
    ctx->rank = 0;
    if (p2p_err) return fail(DSC_E_NCCL, "a peer-memory all-reduce timed out waiting for a peer (results of that step are invalid)");
    return DSC_OK;
}

int dsc_sallreduce_sum(dsc_ctx_t c, int32_t count, const dsc_buf_t* bufs) { DSC_NVTX(c, "allreduce_sum"); return allreduce_impl<float>(c, count, bufs); }
int dsc_dallreduce_sum(dsc_ctx_t c, int32_t count, const dsc_buf_t* bufs) { DSC_NVTX(c, "allreduce_sum"); return allreduce_impl<double>(c, count, bufs); }
int dsc_sallgather(dsc_ctx_t c, dsc_buf_t s, dsc_buf_t r) { return allgather_impl<float>(c, s, r); }
int dsc_dallgather(dsc_ctx_t c, dsc_buf_t s, dsc_buf_t r) { return allgather_impl<double>(c, s, r); }

}  // extern "C"
