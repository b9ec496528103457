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
struct LazyNode;

// One device allocation handed out by the pool; shared by a buffer and the views taken from it.
struct Block {
    void* ptr = nullptr;
    size_t bytes = 0;   // rounded (pool bin) size
    int refs = 0;       // buffers (views, copy-on-write aliases) and pending deferred operations that hold the block
    Ctx* ctx = nullptr;
    int arena = 0;      // 0 = main pool; >0 = id of the CUDA-graph private pool it belongs to
    int views = 0;      // live TRUE views (dsc_buf_view: shared MUTABLE storage) besides the owner
    int cow = 0;        // live holders that must never observe a write through another holder: copy-on-write aliases
                        // (dsc_buf_share, results shared by several handles) — a writer detaches while cow > 0
    std::vector<LazyNode*> readers;   // deferred operations that will read this block: run before any in-place write
    // Operand-split companion of a float32 block (3xTF32 GEMM): lo = rna_tf32(x - trunc_tf32(x)) element by element, same
    // layout.  Written by the kernel that produced the block (bias+activation, adjoint chains, GEMM epilogues) or by the
    // split kernel on first use; valid for elements [lo_begin, lo_end) while lo_epoch == ctx->cache_epoch and nobody has
    // written the block in place since.
    Block* lo = nullptr;
    uint64_t lo_epoch = 0;
    size_t lo_begin = 0, lo_end = 0;
    uint64_t last_other_stream_use = 0;   // ctx->other_stream_seq of the last copy / collective queued OFF the compute stream
    // Ordering against the second compute lane (see Ctx::lane_*).  w_seq: 0 = written on the compute stream at a time not yet
    // pinned; otherwise the value of ctx->order_seq pinned the first time a lane operation looked at the block (every write
    // issued through the compute stream resets it to 0).  lane_w / lane_r: ctx->lane_seq of the last lane operation that
    // wrote / read the block.
    uint64_t w_seq = 0, lane_w = 0, lane_r = 0;
};

// Device twin of ISigmaDiffDataBuffer<'T> (reference Util.fs:133-141): a window over a Block, or a value whose
// computation is deferred (`lazy`, see LazyNode) until a consumer that cannot fuse with it needs the bytes.
struct Buf {
    uint32_t magic = kBufMagic;
    int dtype = DSC_F32;
    Ctx* ctx = nullptr;
    Block* block = nullptr;  // null iff len == 0 or the value is still deferred
    size_t off = 0;          // window start, in elements from block->ptr
    int32_t len = 0;
    LazyNode* lazy = nullptr;   // non-null: deferred value (block == nullptr)
    bool is_view = false;       // created by dsc_buf_view
    bool is_cow = false;        // counted in block->cow
    template <typename T>
    T* ptr() const {
        return block ? reinterpret_cast<T*>(block->ptr) + off : nullptr;
    }
};

// ---- deferred evaluation behind the unchanged Backend<'T> call sequence ----------------------------------------
// A call whose result may fuse with the call(s) that follow returns a buffer handle whose value is a LazyNode.  Consumers that
// recognise a pattern take the node's OPERANDS and launch one fused kernel; every other consumer materialises the node with
// the same kernel(s) the eager sequence would have run, so results are bit-identical either way (tests).  Nodes are reference
// counted: handles and other nodes hold them; a node holds its operands (blocks or nodes).
enum LazyKind {
    LZ_ZERO = 1,   // n zeros                        CreateDataBuffer(CreateZeroArray n)           AD.Float32.fs:1834-1835,3746
    LZ_FILL,       // n copies of s                  CreateDataBuffer(CreateValueArray(n, s))      :3430 (Sum_DM adjoint)
    LZ_T,          // transpose of src (m x n)       Transpose_M                                    :4127-4143
    LZ_ROWS,       // m copies of vector src         RepeatReshapeCopy_V_MRows                      :1860-1865
    LZ_ADD_ROWS,   // src0 (m x n) + rows(src1)      Add_M_M(A, RepeatReshapeCopy_V_MRows(m, v))    :2357
    LZ_MAP,        // f_op(src0)                     Map_F_V / Map_F_M (Cosh, Sign: heads of adjoint chains)
    LZ_SECH,       // 1 / cosh(src0)                 Map_F_S(1, Div, Map(Cosh, x))                  :4238
    LZ_HAD,        // src0 .* src1                   Mul_Had_M_M / Map2(Mul)                        :4240,4280-4281
    LZ_SCALAR,     // scalar broadcast op(kind, s, src0)   Add_S / Sub_S / Mul_S                    :4280-4281
    LZ_GEMM,       // op(src0) op(src1) [+ src2 rows]      Mul_M_M                                  :2382
    LZ_COLSUM,     // column sums of src0 (m x n) stored into a fresh zero vector   Add_M_Colwise_V_InPlace  :4255 after :3640
};

struct Operand {          // an input of a deferred operation: a device window, or another deferred value
    Block* block = nullptr;
    size_t off = 0;
    int32_t len = 0;
    LazyNode* node = nullptr;
    bool same(const Operand& o) const { return node ? node == o.node : (o.node == nullptr && block == o.block && off == o.off && len == o.len); }
};

struct LazyNode {
    int refs = 0;
    int kind = 0;
    int dtype = DSC_F32;
    Ctx* ctx = nullptr;
    int32_t len = 0;
    int op = 0;                 // MapOp / dsc_scalar_kind
    double s = 0.0;             // scalar parameter
    int m = 0, n = 0, k = 0, ta = 0, tb = 0;
    Operand src[3];
    int nsrc = 0;
    Block* result = nullptr;    // set once materialised (the node holds one reference)
    size_t result_off = 0;
    bool consumed = false;      // some consumer has taken this value (or its operands, fused): not flushed at sync points
    bool sink = false;          // LZ_GEMM: a weight adjoint (A was a pending transpose) — nothing will fold into its epilogue
    int pending_idx = -1;       // position in ctx->pending, -1 when not listed
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
    int gemm_epilogue = 1;       // dsc_gemm_epilogue: which epilogues the float32 tensor-core kernel applies itself (0 none, 1 the cheap ones, 2 all)
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
    void* p2p_err_host = nullptr;       // mapped pinned word a timed-out barrier sets; p2p_err_dev: the same word as the kernel addresses it
    void* p2p_err_dev = nullptr;
    unsigned long long p2p_timeout_ns = 30000000000ull;
    dsc_stats stats{};
    int sm_count = kNumSMs;
    size_t smem_optin = 0;
    // deferred evaluation
    bool fuse = true;                    // !DSC_FLAG_NO_FUSE
    std::vector<LazyNode*> pending;      // deferred COMPUTE nodes nobody has consumed yet: run at the next synchronisation point
    uint64_t cache_epoch = 1;            // operand-split companions older than this are stale (bumped at graph begin / end)
    uint64_t other_stream_seq = 0;       // counts copies / collectives queued off the compute stream (see Block::last_other_stream_use)
    uint64_t other_stream_joined = 0;    // value of other_stream_seq the compute stream is known to be ordered after
    bool nvtx = false;                   // DSC_FLAG_NVTX: one NVTX range per C-ABI entry
    // Second compute lane: two independent, small float32 tensor-core products that are both pending (the weight adjoints of
    // a batch shard: neither fills the GPU, each is mostly fixed cost) run CONCURRENTLY — one on the compute stream, one on
    // this stream, each planned for half of the SMs, with its own workspace and tickets (co_schedule_pending_gemms).
    //
    // The same lane also runs the SINKS of a reverse sweep ahead of the compute stream: weight-adjoint products (their A
    // operand is a pending transpose, AD.Float32.fs:4131-4143), the column sums of the bias adjoints (:4255) and the
    // all-reduce of adjoints that were produced there (dsc_lane.cu).  Nothing on the compute stream waits for them until a
    // consumer touches their result (lane_gate_read / lane_gate_write) or a synchronisation point joins the lane, so in a
    // captured step they become a parallel branch of the graph: the adjoint all-reduce of the later layers overlaps the
    // backward products of the first one.
    cudaStream_t lane_stream = nullptr;
    cudaEvent_t ev_lane_fork = nullptr, ev_lane_done = nullptr;
    uint64_t order_seq = 0;              // counter behind Block::w_seq pins and lane forks
    uint64_t lane_fork_seq = 0;          // the lane is ordered after every compute-stream write pinned at or below this value
    // A fork point recorded ahead of a long compute-stream kernel (lane_prefork) that the lane has NOT been made to wait for
    // yet: lane operations whose inputs were visible before it keep running ahead; the wait is applied only for an operation
    // that needs what the fork point covers (lane_apply_prefork).
    cudaEvent_t ev_lane_prefork = nullptr;
    bool prefork_pending = false;
    uint64_t prefork_seq = 0;
    uint64_t lane_seq = 0;               // lane operations issued
    uint64_t lane_joined = 0;            // ... of which the compute stream is ordered after this many
    std::vector<Block*> lane_held;       // blocks lane operations touch: referenced (never recycled) until the next join
    bool lane_fork_open = false;         // a fork has put the lane behind the compute stream and no join has followed yet
    bool on_lane = false;                // the context's stream / scratch members currently name the lane's (lane_enter .. lane_leave)
    bool main_gemm_half_once = false;    // a tensor-core product has just gone to the lane with half of the SMs: the next one on the compute stream takes the other half
    bool lanes_enabled = true;           // DSC_LANES=0 switches the sink lane off
    void* lane_red_scratch = nullptr;    // the lane's own reduction scratch and tickets (column sums, skinny products)
    size_t lane_red_scratch_bytes = 0;
    unsigned int* lane_tickets = nullptr;
    void* lane_gemm_ws = nullptr;
    size_t lane_gemm_ws_bytes = 0;
    unsigned int* lane_gemm_tickets = nullptr;
    size_t lane_gemm_tickets_n = 0;
    int gemm_sm_budget = 0;              // > 0: the planner of the float32 GEMM may use this many SMs only
    bool lane_pending = false;           // work queued on the lane stream that the compute stream has not been ordered after
    uint64_t stats_co_scheduled = 0;
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

// ---- NVTX: one range per C-ABI entry point, named after the Backend<'T> member(s) it serves (DSC_FLAG_NVTX / DSC_NVTX=1) ----
void nvtx_push(const char* name);   // dsc_core.cu (nvtx3 is header-only and costs nothing when no tool is attached)
void nvtx_pop();
struct NvtxRange {
    bool on;
    NvtxRange(dsc_ctx_t c, const char* name) : on(c && reinterpret_cast<Ctx*>(c)->magic == kCtxMagic && reinterpret_cast<Ctx*>(c)->nvtx) {
        if (on) nvtx_push(name);
    }
    ~NvtxRange() {
        if (on) nvtx_pop();
    }
};
#define DSC_NVTX(c, name) ::dsc::NvtxRange _nvtx_range((c), (name))

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
// Validate an input handle of the given dtype (zero-length buffers are valid) and MATERIALISE a deferred value.
int in_buf(Ctx* ctx, int dtype, dsc_buf_t h, Buf** res, const char* what);
// The same without materialising: the caller inspects b->lazy (fusing consumers).
int in_buf_peek(Ctx* ctx, int dtype, dsc_buf_t h, Buf** res, const char* what);
// Before writing through `b` in place: materialise it, run the deferred operations that still read its block, give it a
// private block if copy-on-write aliases exist, and drop the operand-split companion.
int buf_writable(Ctx* ctx, Buf* b, bool on_compute_stream = true);
// ---- deferred evaluation (dsc_lazy.cu) ----
LazyNode* lazy_new(Ctx* ctx, int kind, int dtype, int32_t len);
void lazy_set_operand(LazyNode* nd, int i, Buf* b);          // operand i := the value of b (its block window, or its node)
void lazy_set_operand(LazyNode* nd, int i, const Operand& o);
void lazy_unref(LazyNode* nd);
int lazy_buf(Ctx* ctx, LazyNode* nd, dsc_buf_t* out);          // new handle whose value is nd (takes over the caller's reference)
int lazy_materialize(LazyNode* nd);                            // compute nd->result (idempotent)
void lazy_set_result(LazyNode* nd, Block* blk, size_t off);    // a fused consumer produced nd's value as a by-product: record it, release the operands
int buf_resolve(Buf* b);                                       // make b a plain device window
int operand_resolve(Ctx* ctx, Operand& o);                     // make o a plain device window (materialises its node)
int flush_pending(Ctx* ctx);                                   // run every deferred compute node nobody has consumed
int flush_colsums_on_lane(Ctx* ctx);                           // deferred column sums whose matrix the second lane already sees: run them there
int co_schedule_pending_gemms(Ctx* ctx);                       // pairs of small pending float32 products: one per compute lane, concurrently (dsc_gemm.cu)
// ---- the sink lane (dsc_lane.cu) ----
void lane_swap(Ctx* ctx);                                      // the lane's stream / workspace / scratch become the context's current ones (and back)
// Can a lane operation that reads `reads` (written on the compute stream) start without a new fork, i.e. is the lane already
// ordered after the last write of every one of them?  (Pins unpinned blocks: a block looked at for the first time now is not.)
bool lane_sees(Ctx* ctx, Block* const* reads, int n);
// Begin a lane operation: fork from the compute stream if some block in `reads` is not yet visible to the lane (or force_fork:
// in-place writes to blocks the compute stream may still be reading), then make the lane current.  lane_leave stamps the
// blocks read / written, holds them until the next join and makes the compute stream current again.
int lane_enter(Ctx* ctx, Block* const* reads, int n, bool force_fork);
void lane_leave(Ctx* ctx, Block* const* reads, int nr, Block* const* writes, int nw);
int lane_join(Ctx* ctx);                                       // the compute stream waits for everything issued on the lane
// Fork now, before the compute stream is given a long kernel: lane operations issued afterwards whose inputs are among `reads`
// (or were visible before) start beside that kernel instead of behind it.
int lane_prefork(Ctx* ctx, Block* const* reads, int n);
inline bool lane_prefork_covers(const Ctx* ctx, const Block* b) { return ctx->prefork_pending && b && b->w_seq != 0 && b->w_seq <= ctx->prefork_seq; }
int lane_apply_prefork(Ctx* ctx);                              // the lane now waits for the recorded fork point
inline bool lane_busy(const Ctx* ctx) { return ctx->lane_seq > ctx->lane_joined; }
bool lane_debug();                                             // DSC_LANE_DEBUG=1: every lane decision on stderr
// compute-stream work is about to read / write blk
inline int lane_gate_read(Ctx* ctx, Block* blk) { return (blk && !ctx->on_lane && blk->lane_w > ctx->lane_joined) ? lane_join(ctx) : DSC_OK; }
inline int lane_gate_write(Ctx* ctx, Block* blk) {
    if (!blk || ctx->on_lane) return DSC_OK;
    blk->w_seq = 0;
    return (blk->lane_w > ctx->lane_joined || blk->lane_r > ctx->lane_joined) ? lane_join(ctx) : DSC_OK;
}
inline LazyNode* lazy_of(const Buf* b) { return (b && !b->block) ? b->lazy : nullptr; }
inline LazyNode* lazy_of(const Buf* b, int kind) { LazyNode* n = lazy_of(b); return (n && n->kind == kind && !n->result) ? n : nullptr; }
void buf_adopt(Buf* b, Block* blk, size_t off, bool cow);      // b := window over blk (takes a new reference)
int buf_share_into(Buf* dst, Buf* src);                        // dst := copy-on-write alias of src's value (dst must hold nothing)
// the compute stream is about to reuse / the host is about to free memory last touched off the compute stream
int join_other_streams(Ctx* ctx);
inline void note_other_stream_use(Ctx* ctx, Block* b) { if (b) b->last_other_stream_use = ++ctx->other_stream_seq; }
int ensure_red_scratch(Ctx* ctx, size_t bytes);
int ensure_gemm_ws(Ctx* ctx, size_t bytes);
int ensure_gemm_tickets(Ctx* ctx, size_t count);

inline int grid_for(size_t work_items, int per_block, int max_blocks) {
    size_t b = (work_items + per_block - 1) / per_block;
    if (b < 1) b = 1;
    if (b > (size_t)max_blocks) b = max_blocks;
    return (int)b;
}

// The operand split of the 3xTF32 GEMM (dsc_gemm_tf32x3.cu): the tensor core reads the fp32 word itself as the "hi" part
// (kind::tf32 ignores the low 13 mantissa bits), lo = rna_tf32(x - trunc_tf32(x)); x - hi is exact in fp32.  Kernels that
// PRODUCE a matrix a later GEMM will read write this next to the value (Block::lo) so that no separate split pass runs.
__device__ __forceinline__ float dsc_lo_of(float x) {
    const float res = x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(res));
    return (res - res == 0.f) ? __uint_as_float(r) : 0.f;   // non-finite x contributes through hi only
}
__device__ __forceinline__ double dsc_lo_of(double) { return 0.0; }   // never called: float64 GEMMs use DMMA

// ---- raw executors: kernels on device pointers, no handle logic (used by the entry points and by the deferred-evaluation
// engine; each is explicitly instantiated for float and double in the translation unit that owns the kernel) -------------
template <typename T> int ew_map(Ctx*, int op, const T* x, T* y, size_t n);                       // dsc_elementwise.cu
template <typename T> int ew_map_s(Ctx*, int op, T s, const T* x, T* y, size_t n);
template <typename T> int ew_map2(Ctx*, int op, const T* a, const T* b, T* y, size_t n);
template <typename T> int ew_scalar(Ctx*, int kind, T s, const T* x, T* y, size_t n);
template <typename T> int ew_neg(Ctx*, const T* x, T* y, size_t n);
template <typename T> int ew_sech(Ctx*, const T* x, T* y, size_t n);                              // 1 / cosh(x), the two roundings of the unfused pair
// y = chain(dA, p) (dsc_fused_kind); y_lo (float32 only, may be null): the operand-split companion of y, written in the same pass
template <typename T> int ew_fused_bwd(Ctx*, int kind, const T* dA, const T* p, T* y, T* y_lo, size_t n);
// z[i,j] = A[i,j] (bop) v[j] with bop in {ADD, SUB, MUL, DIV}; if act >= 0 additionally h = f_act(z) (and h_lo, float32 only)
template <typename T> int ew_rows(Ctx*, int bop, int act, int m, int n, const T* A, const T* v, T* z, T* h, T* h_lo);
template <typename T> int transpose_launch(Ctx* ctx, const T* A, T* B, int m, int n);             // dsc_layout.cu
template <typename T> int repeat_rows_launch(Ctx* ctx, const T* v, T* out, int m, int n);
template <typename T> int colsum_launch(Ctx* ctx, const T* M, T* v, int m, int n, bool accumulate);   // v[j] (+)= sum_i M[i,j]
int fill_raw(Ctx* ctx, int dtype, void* base, size_t n, double v);                                // dsc_core.cu

// ---- typed implementations (one template per op, instantiated for float and double) -------
// A matrix operand: its device window and the block behind it (the float32 tensor-core GEMM keeps / finds the operand-split
// companion there, see Block::lo).  block may be null (no caching for that operand).
struct MatRef {
    const void* ptr = nullptr;
    Block* block = nullptr;
    size_t off = 0;      // window start in elements from block->ptr
};
// What the GEMM does with an output tile besides storing it (the calls that FOLLOW Mul_M_M in the unchanged sequence):
//   GEMM_EP_BIAS_ACT: z = acc + v[j] -> C;  h = f_op(z) -> out2 (op < 0: no h)      Add_M_M(., rows(v)) -> Map_F_M   (:2357,:2704)
//   GEMM_EP_CHAIN   : y = chain_op(acc, aux[i,j]) -> C  (dsc_fused_kind)             the adjoint chains            (:4238-4240,4280-4281)
// lo (float32, may be null): operand-split companion of the LAST output (h, or z when op < 0, or y), same layout.
enum { GEMM_EP_NONE = 0, GEMM_EP_BIAS_ACT = 1, GEMM_EP_CHAIN = 2 };
struct GemmEpilogue {
    int mode = GEMM_EP_NONE;
    int op = -1;
    const void* vec = nullptr;   // v (n)
    const void* aux = nullptr;   // p (m x n)
    void* out2 = nullptr;        // h (m x n)
    float* lo = nullptr;
};
// C(m x n) = op(A) op(B) [+ bias[i]] [epilogue]: picks the kernel (skinny / tensor core / CUDA core); kernels without a fused
// epilogue run it as a second pass over C.
template <typename T> int gemm_run(Ctx*, int ta, int tb, int m, int n, int k, MatRef A, MatRef B, const T* bias, T* C, const GemmEpilogue* ep);
// f32 tensor-core path (dsc_gemm_tf32x3.cu): returns DSC_OK if it handled the GEMM (epilogue included), 1 if the
// shape is not eligible (caller falls through to the CUDA-core kernel; never a CPU fallback), 2 if the product is stored
// but the epilogue has to run as a second pass.
int gemm_tf32x3_tcgen05(Ctx*, int ta, int tb, int m, int n, int k, MatRef A, MatRef B,
                        const float* bias, float* C, const GemmEpilogue* ep);
int gemm_f64_dmma(Ctx*, int ta, int tb, int m, int n, int k, const double* A, const double* B,
                  const double* bias, double* C);
// n <= 16 products on CUDA cores (dsc_gemm_skinny.cu); same return convention
// bias_col (may be null): v[j] added to every row after bias[i] — Add_M_M(., rows(v)) folded into the product when the
// kernel taken supports it (*bias_col_done = 1), otherwise the caller adds it as a second pass
template <typename T>
int gemm_skinny(Ctx*, int ta, int tb, int m, int n, int k, const T* A, const T* B, const T* bias, T* C, const T* bias_col = nullptr,
                int* bias_col_done = nullptr);

}  // namespace dsc
