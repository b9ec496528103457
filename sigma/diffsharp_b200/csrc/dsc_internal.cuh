// Internal definitions shared by the translation units of libdiffsharp_cuda.so.
// Nothing here crosses the C ABI (include/diffsharp_cuda.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../../include/diffsharp_cuda.h"

namespace dsc {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs. Grids are sized in multiples of this.
constexpr uint32_t kBufMagic = 0xD5CB0F01u;
constexpr uint32_t kCtxMagic = 0xD5CC7801u;
constexpr int kNumTickets = 4096;

struct Ctx;

// One device allocation handed out by the pool; shared by a buffer and the views taken from it.
struct Block {
    void* ptr = nullptr;
    size_t bytes = 0;   // rounded (pool bin) size
    int refs = 0;
    Ctx* ctx = nullptr;
    int arena = 0;      // 0 = main pool; >0 = id of the CUDA-graph private pool it belongs to
    bool has_views = false;   // a true view (dsc_buf_view: shared MUTABLE storage) exists or existed
    bool cow_shared = false;  // copy-on-write aliases (dsc_buf_share) exist or existed
};

// Device twin of ISigmaDiffDataBuffer<'T> (reference Util.fs:133-141): a window over a Block.
struct Buf {
    uint32_t magic = kBufMagic;
    int dtype = DSC_F32;
    Ctx* ctx = nullptr;
    Block* block = nullptr;  // null iff len == 0
    size_t off = 0;          // window start, in elements from block->ptr
    int32_t len = 0;
    template <typename T>
    T* ptr() const {
        return block ? reinterpret_cast<T*>(block->ptr) + off : nullptr;
    }
};

struct Graph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int arena = 0;
    std::vector<Block*> blocks;  // every block carved while capturing; freed with the graph
};

struct Pool {
    // size-binned free lists; all users are on one stream, so a freed block is reusable at once.
    std::multimap<size_t, void*> free_blocks;
    size_t reserved = 0, in_use = 0;
    uint64_t hits = 0, misses = 0;
};

struct Ctx {
    uint32_t magic = kCtxMagic;
    int device = 0;
    uint32_t flags = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t comm_stream = nullptr;
    cudaStream_t copy_stream = nullptr;       // host -> device transfers of the copy-stream API
    cudaStream_t copy_stream_down = nullptr;  // device -> host transfers: PCIe is full duplex, the two directions overlap
    cudaEvent_t ev_copy_a = nullptr, ev_copy_b = nullptr, ev_copy_c = nullptr;
    cudaEvent_t ev_compute = nullptr, ev_comm = nullptr;
    bool comm_pending = false;
    Pool pool;
    // capture state
    bool capturing = false;
    int next_arena = 1;
    int live_graphs = 0;          // scratch / workspace may not be reallocated while a captured graph references them
    int cur_arena = 0;
    Pool arena_pool;              // private pool of the graph being captured
    std::vector<Block*> arena_blocks;
    // scratch for reductions: partials + ticket counter + result slot
    void* red_scratch = nullptr;   // device
    size_t red_scratch_bytes = 0;
    void* red_host = nullptr;      // pinned host, 256 B
    // zero-initialised arrival counters for "last CTA of a group finishes the reduction" kernels
    // (each kernel resets the counters it used; one stream => no two users at a time)
    unsigned int* tickets = nullptr;
    void* flush_buf = nullptr;     // L2 flush scratch
    size_t flush_bytes = 0;
    // GEMM operand-split workspace (3xTF32 hi/lo planes)
    void* gemm_ws = nullptr;
    size_t gemm_ws_bytes = 0;
    int gemm_force_bn = -1, gemm_force_pair = -1, gemm_force_split = -1, gemm_force_fuse = -1;   // dsc_gemm_tune; -1 = planner decides
    unsigned long long* gemm_stamps = nullptr;   // %globaltimer at the phase boundaries of the last tensor-core GEMM (DSC_GEMM_STAMPS=1)
    unsigned int* gemm_tickets = nullptr;   // one arrival counter per output tile (stream-K fix-up), zero between launches
    size_t gemm_tickets_n = 0;
    // NCCL
    void* nccl_comm = nullptr;
    void* comm_stage = nullptr;    // contiguous staging buffer for packed all-reduces
    size_t comm_stage_bytes = 0;
    int nranks = 1, rank = 0;
    // peer-memory all-reduce (dsc_comm.cu): one symmetric buffer per rank, opened on every peer through CUDA IPC
    bool p2p_ok = false;
    void* p2p_local = nullptr;          // this rank's symmetric buffer: [flags | sync words | region A | region B]
    void* p2p_peer[8] = {nullptr};      // the same buffer of every rank as mapped into this process (p2p_peer[rank] == p2p_local)
    size_t p2p_region_bytes = 0;        // capacity of region A (= region B)
    dsc_stats stats{};
    int sm_count = kNumSMs;
    size_t smem_optin = 0;
};

// ---- error plumbing -----------------------------------------------------------------------
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

#define DSC_CUDA_CHECK(expr)                                                                       \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return ::dsc::fail(_e == cudaErrorMemoryAllocation ? DSC_E_OOM : DSC_E_CUDA,           \
                               "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                               __LINE__);                                                          \
    } while (0)

#define DSC_TRY(expr)              \
    do {                           \
        int _s = (expr);           \
        if (_s != DSC_OK) return _s; \
    } while (0)

#define DSC_LAUNCH_CHECK(ctx)                                                                  \
    do {                                                                                       \
        (ctx)->stats.kernel_launches++;                                                        \
        cudaError_t _e = cudaPeekAtLastError();                                                \
        if (_e != cudaSuccess) {                                                               \
            cudaGetLastError();                                                                \
            return ::dsc::fail(DSC_E_CUDA, "kernel launch failed: %s (%s:%d)",                 \
                               cudaGetErrorString(_e), __FILE__, __LINE__);                    \
        }                                                                                      \
    } while (0)

// ---- handle helpers -----------------------------------------------------------------------
inline Buf* as_buf(dsc_buf_t h) {
    Buf* b = reinterpret_cast<Buf*>(h);
    return (b && b->magic == kBufMagic) ? b : nullptr;
}
inline bool ctx_ok(dsc_ctx_t c) {
    return c && reinterpret_cast<Ctx*>(c)->magic == kCtxMagic;
}
inline Ctx* as_ctx(dsc_ctx_t c) { return reinterpret_cast<Ctx*>(c); }

template <typename T>
struct dtype_of;
template <>
struct dtype_of<float> {
    static constexpr int value = DSC_F32;
};
template <>
struct dtype_of<double> {
    static constexpr int value = DSC_F64;
};
inline size_t dtype_size(int dt) { return dt == DSC_F64 ? 8 : 4; }

// allocation (dsc_core.cu)
int pool_alloc(Ctx* ctx, size_t bytes, Block** out);
void block_release(Block* blk);
int buf_new(Ctx* ctx, int dtype, int32_t n, Buf** out);
// Resolve an "out" parameter: allocate when *out == 0, otherwise validate dtype/length.
int out_buf(Ctx* ctx, int dtype, int32_t n, dsc_buf_t* out, Buf** res);
// Validate an input handle of the given dtype (zero-length buffers are valid).
int in_buf(Ctx* ctx, int dtype, dsc_buf_t h, Buf** res, const char* what);
// Copy-on-write: before writing through `b`, give it a private block if its block has CoW aliases.
int buf_writable(Ctx* ctx, Buf* b);
int ensure_red_scratch(Ctx* ctx, size_t bytes);
int ensure_gemm_ws(Ctx* ctx, size_t bytes);
int ensure_gemm_tickets(Ctx* ctx, size_t count);

inline int grid_for(size_t work_items, int per_block, int max_blocks) {
    size_t b = (work_items + per_block - 1) / per_block;
    if (b < 1) b = 1;
    if (b > (size_t)max_blocks) b = max_blocks;
    return (int)b;
}

// ---- typed implementations (one template per op, instantiated for float and double) -------
template <typename T> int gemm_impl(Ctx*, int ta, int tb, int m, int n, int k, Buf* A, Buf* B, Buf* bias, Buf* C);
// f32 tensor-core path (dsc_gemm_tf32x3.cu): returns DSC_OK if it handled the GEMM, 1 if the
// shape is not eligible (caller falls through to the CUDA-core kernel; never a CPU fallback).
int gemm_tf32x3_tcgen05(Ctx*, int ta, int tb, int m, int n, int k, const float* A, const float* B,
                        const float* bias, float* C);
int gemm_f64_dmma(Ctx*, int ta, int tb, int m, int n, int k, const double* A, const double* B,
                  const double* bias, double* C);
// n <= 16 products on CUDA cores (dsc_gemm_skinny.cu); same return convention
template <typename T>
int gemm_skinny(Ctx*, int ta, int tb, int m, int n, int k, const T* A, const T* B, const T* bias, T* C);

}  // namespace dsc
