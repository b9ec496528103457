/*
 * diffsharp_cuda.h — C ABI of libdiffsharp_cuda.so, the B200 (sm_100a) compute backend that sits
 * behind Sigma.DiffSharp's Backend<'T> plugin interface for 'T = float32 ("s") and float ("d").
 *
 * Every entry point below is what a managed (F#/C#) implementation of
 *     DiffSharp.Backend.Backend<'T>             (reference: src/DiffSharp/Backend.fs:74-139)
 *     DiffSharp.Util.ISigmaDiffDataBuffer<'T>   (reference: src/DiffSharp/Util.fs:133-141)
 * binds with P/Invoke, the same way the reference deployment binds OpenBLAS (cblas_sgemm, ...).
 * Each declaration cites the Backend<'T> member(s) it replaces.  INTEGRATION.md shows the F# stub.
 *
 * Conventions
 *  - plain C: pointers, sizes and opaque handles only; no C++ / torch types.
 *  - every function returns an int32 status: DSC_OK (0) or a negative DSC_E_* class.  A human
 *    readable message for the last failure on the calling thread is at dsc_last_error().
 *    Nothing throws or aborts across the ABI.
 *  - dsc_ctx_t: one context = one GPU = one compute stream (+ one communication stream).  A context
 *    is used by one host thread at a time (the reference has no locks either: Util.fs:218-229).
 *  - dsc_buf_t: opaque handle of a device-resident window (Offset, Length) over a ref-counted device
 *    allocation — the device twin of ISigmaDiffDataBuffer (Data/Offset/Length).  0 is "no buffer".
 *    A zero-LENGTH buffer is legal and means "additive identity / empty" exactly as the reference's
 *    DVector.Zero / DNDArray.Zero do (AD.Float32.fs:706,1833; used at :3919,:4245-4257).
 *  - matrices are row-major and contiguous from the window start, like ShapedDataBufferView
 *    (Util.fs:190-193: [i,j] -> Data[Offset + i*Cols + j]); shapes are passed next to the handle.
 *  - lengths/offsets are int32 like ISigmaDiffDataBuffer; N-d shapes are const int64_t* + rank.
 *  - "out" parameters of type dsc_buf_t*: if *out == 0 on entry the library allocates the result
 *    (the Backend<'T> contract: results are NEW buffers); if *out != 0 it must be a buffer of the
 *    right dtype and length and is overwritten.  The caller owns every handle it receives and
 *    releases it with dsc_buf_free (managed side: SafeHandle / finalizer).
 *  - all work is asynchronous on the context's stream; only functions that return a host scalar or
 *    copy to host memory synchronise.
 *  - the dtype is in the symbol name (s = float32, d = float64), BLAS style.
 *  - MapOp op-codes are the ordinals of the reference's MapOp union (Backend.fs:42-70).
 *  - there is NO CPU fallback anywhere: an unsupported op-code is DSC_E_UNSUPPORTED_OP.
 *
 * Deferred evaluation.  The unchanged AD layer issues fixed call chains whose intermediate results nobody else
 * reads (Transpose_M -> Mul_M_M; RepeatReshapeCopy_V_MRows -> Add_M_M -> Map_F_M; Map(Cosh) -> Map_F_S(1, Div) -> Had -> Had;
 * two zero adjoints per tape node; ...).  A managed shim makes ONE call per Backend<'T> member; behind the handle, a call that
 * may start such a chain returns a buffer whose VALUE is deferred.  A consumer that completes the chain launches one fused
 * kernel with the chain's own sequence of roundings; any other use of the handle (another op, a download, a view) computes
 * the value with exactly the kernel the call would have launched.  Results are bit-identical to the undeferred sequence
 * (DSC_FLAG_NO_FUSE); deferred work that nobody consumed runs at the next dsc_sync / dsc_event_record / dsc_graph_end.
 *
 * Second compute lane.  The sinks of a reverse sweep — the weight-adjoint product Transpose(a.P) * d.A (AD.Float32.fs:4131-4143)
 * and the bias column sums Add_M_Colwise_V_InPlace (:4255) — are issued on a second stream of the context, forked from the
 * compute stream only where a data dependency requires it and joined when a consumer touches their result or at a
 * synchronisation point; captured in a CUDA graph they are a parallel branch.  Invisible to the caller (same results, bit for
 * bit; DSC_LANES=0 keeps everything on the compute stream); dsc_stats.lane_ops counts the operations issued there.
 */
#ifndef DIFFSHARP_CUDA_H
#define DIFFSHARP_CUDA_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define DSC_API __declspec(dllexport)
#else
#define DSC_API __attribute__((visibility("default")))
#endif

#define DSC_VERSION 100 /* 0.1.0 */

/* status codes */
#define DSC_OK 0
#define DSC_E_INVALID (-1)        /* null/dead handle, bad argument                                  */
#define DSC_E_SHAPE (-2)          /* ErrorMessages.InvalidArgVV/MM/MColsMRows (Util.fs:117-131)      */
#define DSC_E_OOM (-3)
#define DSC_E_CUDA (-4)
#define DSC_E_NCCL (-5)
#define DSC_E_UNSUPPORTED_OP (-6) /* managed side raises NotSupportedException, never runs a closure */
#define DSC_E_SINGULAR (-7)       /* managed side returns None (Backend.fs:102-103,125-126)          */
#define DSC_E_DTYPE (-8)

/* dtypes */
#define DSC_F32 0
#define DSC_F64 1

/* dsc_create flags */
#define DSC_FLAG_DEFAULT 0u
#define DSC_FLAG_GEMM_FP32_SIMT 1u /* f32 GEMM on CUDA cores (exact FMA) instead of tcgen05 3xTF32     */
#define DSC_FLAG_GEMM_F64_SIMT 2u  /* f64 GEMM on CUDA cores instead of DMMA                           */
#define DSC_FLAG_NO_POOL 4u        /* cudaMalloc/cudaFree per buffer (debugging, compute-sanitizer)    */
#define DSC_FLAG_NO_FUSE 8u        /* no deferred evaluation: every call launches its own kernel(s) at once (see "Deferred evaluation") */
#define DSC_FLAG_NVTX 16u          /* one NVTX range per entry point, named after the Backend<'T> member it serves (also DSC_NVTX=1)   */

/* MapOp ordinals — Backend.fs:42-70, declaration order */
enum dsc_mapop {
    DSC_OP_MUL = 0, DSC_OP_DIV = 1, DSC_OP_ADD = 2, DSC_OP_SUB = 3,
    DSC_OP_POW_AB = 4, DSC_OP_POW_BA = 5, DSC_OP_ATAN2_AB = 6, DSC_OP_ATAN2_BA = 7,
    DSC_OP_LOG = 8, DSC_OP_LOG10 = 9, DSC_OP_EXP = 10, DSC_OP_SIN = 11, DSC_OP_COS = 12,
    DSC_OP_TAN = 13, DSC_OP_SQRT = 14, DSC_OP_SINH = 15, DSC_OP_COSH = 16, DSC_OP_TANH = 17,
    DSC_OP_ASIN = 18, DSC_OP_ACOS = 19, DSC_OP_ATAN = 20, DSC_OP_ABS = 21, DSC_OP_SIGN = 22,
    DSC_OP_FLOOR = 23, DSC_OP_CEILING = 24, DSC_OP_ROUND = 25, DSC_OP_REL = 26, DSC_OP_SIGMOID = 27,
    DSC_OP__COUNT = 28
};

/* scalar-broadcast kinds for dsc_?scalar_op */
enum dsc_scalar_kind {
    DSC_S_ADD = 0,   /* x[i] + s   Add_S_V / Add_S_M   (Backend.fs:93,114)  */
    DSC_S_SUB_SX = 1,/* s - x[i]   Sub_S_V / Sub_S_M   (Backend.fs:95,118)  */
    DSC_S_SUB_XS = 2,/* x[i] - s   Sub_V_S / Sub_M_S   (Backend.fs:96,117)  */
    DSC_S_MUL = 3    /* s * x[i]   Mul_S_V / Mul_S_M   (Backend.fs:97,120)  */
};

/* fused adjoint chains, recognised INSIDE the library from the unchanged call sequence (dsc_lazy.cu);
 * dsc_?fused_bwd exposes the same kernels directly */
enum dsc_fused_kind {
    DSC_FUSED_TANH_BWD = 0,    /* dA .* (1/cosh aP) .* (1/cosh aP)          AD.Float32.fs:4238-4240 */
    DSC_FUSED_SIGMOID_BWD = 1, /* dA .* dP .* (1 - dP)                      AD.Float32.fs:4281      */
    DSC_FUSED_RELU_BWD = 2     /* dA .* ((sign(aP) + 1) / 2)                AD.Float32.fs:4280      */
};

typedef struct dsc_ctx_s* dsc_ctx_t;
typedef uint64_t dsc_buf_t;
typedef uint64_t dsc_event_t;
typedef uint64_t dsc_graph_t;

typedef struct dsc_stats {
    uint64_t kernel_launches;  /* kernels of this library launched on the context since creation */
    uint64_t memcpy_h2d_bytes;
    uint64_t memcpy_d2h_bytes;
    uint64_t pool_bytes_reserved; /* device bytes held by the pool (in use + cached)               */
    uint64_t pool_bytes_in_use;
    uint64_t pool_hits;
    uint64_t pool_misses;         /* cudaMalloc calls                                              */
    uint64_t gemm_tcgen05;        /* f32 GEMMs served by the tcgen05/TMEM/TMA 3xTF32 kernel         */
    uint64_t gemm_dmma;           /* f64 GEMMs served by the DMMA kernel                            */
    uint64_t gemm_simt;           /* GEMMs served by the CUDA-core kernels                          */
    uint64_t allreduce_p2p;       /* all-reduces served by the peer-memory (NVLink P2P) kernel      */
    uint64_t allreduce_nccl;      /* all-reduces served by NCCL                                     */
    uint64_t lane_ops;            /* operations issued on the second compute lane (sinks of reverse sweeps) */
} dsc_stats;

/* ---------------------------------------------------------------- lifecycle ---------------- */
/* New surface (no Backend<'T> twin): the reference backend is process-global OpenBLAS state.    */
DSC_API int dsc_version(void);
DSC_API int dsc_device_count(int* count);
DSC_API int dsc_create(int device, uint32_t flags, dsc_ctx_t* ctx);
DSC_API int dsc_destroy(dsc_ctx_t ctx);
DSC_API int dsc_sync(dsc_ctx_t ctx);
DSC_API const char* dsc_last_error(void);
DSC_API int dsc_get_stats(dsc_ctx_t ctx, dsc_stats* out);
DSC_API int dsc_get_stream(dsc_ctx_t ctx, void** cuda_stream);
DSC_API int dsc_trim_pool(dsc_ctx_t ctx);
/* Deferred evaluation: switch it at run time (off: everything pending runs first); run everything pending now; forget the
 * cached operand splits of the float32 GEMM (they are keyed on the buffer and dropped on any in-place write: a benchmark loop
 * that re-runs a step on unchanged inputs calls this so that no work is carried over from the previous step). */
DSC_API int dsc_set_fuse(dsc_ctx_t ctx, int on);
DSC_API int dsc_flush(dsc_ctx_t ctx);
DSC_API int dsc_invalidate_caches(dsc_ctx_t ctx);
/* CUDA events on the context's stream (bench.py times kernels with these) */
DSC_API int dsc_event_create(dsc_ctx_t ctx, dsc_event_t* ev);
DSC_API int dsc_event_record(dsc_ctx_t ctx, dsc_event_t ev);
DSC_API int dsc_event_elapsed_ms(dsc_ctx_t ctx, dsc_event_t start, dsc_event_t stop, float* ms);
DSC_API int dsc_event_destroy(dsc_ctx_t ctx, dsc_event_t ev);
DSC_API int dsc_event_record_copy(dsc_ctx_t ctx, dsc_event_t ev); /* record on the upload copy stream    */
DSC_API int dsc_event_record_copy_down(dsc_ctx_t ctx, dsc_event_t ev); /* record on the download copy stream */
DSC_API int dsc_event_sync(dsc_ctx_t ctx, dsc_event_t ev);        /* host waits for a recorded event      */
DSC_API int dsc_compute_wait_event(dsc_ctx_t ctx, dsc_event_t ev);/* compute stream waits for the event   */
/* CUDA-graph capture of a replayed call sequence (static shapes): begin, issue the calls once,
 * end; dsc_graph_launch replays every kernel/memset/collective that was issued.  Buffers
 * allocated while capturing come from a pool private to the graph. */
DSC_API int dsc_graph_begin(dsc_ctx_t ctx);
DSC_API int dsc_graph_end(dsc_ctx_t ctx, dsc_graph_t* graph);
DSC_API int dsc_graph_launch(dsc_ctx_t ctx, dsc_graph_t graph);
DSC_API int dsc_graph_destroy(dsc_ctx_t ctx, dsc_graph_t graph);
/* write `bytes` bytes of scratch (> L2) so that the next timed kernel starts from a cold L2 */
DSC_API int dsc_flush_l2(dsc_ctx_t ctx);
/* Pin the plan of the float32 tensor-core GEMM (tests force every kernel variant; tools sweep them):
 * tile width bn in {256,128,64}, pair = 1 for the 2-CTA (cta_group::2) kernel, split = K ranges per
 * tile when the tiles do not fill the GPU, fuse = in-kernel operand split (0 off, 1 single-CTA
 * kernels, 2 also the 2-CTA 256 x 256 kernel).  -1 = planner decides. */
DSC_API int dsc_gemm_tune(dsc_ctx_t ctx, int bn, int pair, int split, int fuse);
/* Which of the calls that follow Mul_M_M in the unchanged sequence (Add_M_M(., rows(b)) -> Map_F_M; the activation's adjoint
 * chain) the float32 tensor-core kernel applies to the finished tile ITSELF instead of as a second pass over the product:
 * 0 none, 1 (default) the ones measured to pay — bias rows without activation, ReL, the ReL chain —, 2 all (Tanh, Sigmoid and
 * their chains too: correct and bit-identical, but the 8 epilogue warps become the bottleneck; kept for the parity tests). */
DSC_API int dsc_gemm_epilogue(dsc_ctx_t ctx, int mode);
/* The plan the float32 tensor-core GEMM makes for op(A)(m x k) op(B)(k x n) on a device with sm_count SMs — a pure
 * function (no GPU, no context), exported so that the schedule arithmetic is unit-tested on any machine.
 * a_direct / b_direct: the operand is addressable by TMA in place (2: and its operand-split companion exists already: no
 * pre-pass, no in-kernel split for it); bn, pair, split, fuse as in dsc_gemm_tune.
 * out = {eligible, bn, pair, workers, chunks per tile, k-blocks, tiles_m, tiles_n, dp_waves, dp_extra, stream-K units,
 *        perm_s, base, rem, dp_tiles, perm_tiles, K ranges per cluster if the cluster exchange applies, a_fused, b_fused, 0} */
DSC_API int dsc_gemm_plan(int sm_count, int m, int n, int k, int a_direct, int b_direct, int bn, int pair, int split, int fuse, int32_t out[20]);
/* With DSC_GEMM_STAMPS=1 in the environment: %globaltimer (ns) as CTA 0 of the last float32 tensor-core GEMM
 * saw it at kernel entry, after the prologue (barriers, TMEM), when the first operand tiles had landed, when
 * the last MMA was issued, when the last accumulator chunk was drained, and when the epilogue was done;
 * [6], [7] (aligned split-K only): segment parked, all segments of the tile parked. */
DSC_API int dsc_gemm_debug_stamps(dsc_ctx_t ctx, uint64_t out[8]);

/* ---------------------------------------------------------------- buffers ------------------ */
/* CreateDataBuffer / CreateZeroArray / CreateValueArray / CreateUninitialisedArray
 * (Backend.fs:76-79) and ISigmaDiffDataBuffer (Util.fs:133-141). */
DSC_API int dsc_buf_alloc(dsc_ctx_t ctx, int dtype, int32_t n, dsc_buf_t* out);               /* CreateUninitialisedArray+CreateDataBuffer */
DSC_API int dsc_buf_alloc_fill(dsc_ctx_t ctx, int dtype, int32_t n, double v, dsc_buf_t* out);/* CreateZeroArray/CreateValueArray+CreateDataBuffer */
DSC_API int dsc_buf_free(dsc_buf_t h);
DSC_API int dsc_buf_upload(dsc_buf_t h, int32_t off, const void* host, int32_t n);            /* CreateDataBuffer(a) */
DSC_API int dsc_buf_download(dsc_buf_t h, int32_t off, void* host, int32_t n);                /* .Data / .SubData    */
DSC_API int dsc_buf_upload_async(dsc_buf_t h, int32_t off, const void* pinned_host, int32_t n);
DSC_API int dsc_buf_download_async(dsc_buf_t h, int32_t off, void* pinned_host, int32_t n);
/* The same copies on the context's COPY streams (one per direction: PCIe is full duplex), so that the
 * host<->device traffic of minibatch i+1 / i-1 overlaps the kernels of minibatch i.  Uploads are
 * ordered among themselves, downloads among themselves; an upload and a download are ordered only
 * through events (dsc_event_record_copy / dsc_event_record_copy_down + dsc_compute_wait_event).
 * dsc_copy_wait_compute: copies queued after this call start only when the compute work queued so far
 * has finished; dsc_compute_wait_copy: the converse; dsc_copy_sync: the host waits for both streams. */
DSC_API int dsc_copy_upload(dsc_buf_t h, int32_t off, const void* pinned_host, int32_t n);
DSC_API int dsc_copy_download(dsc_buf_t h, int32_t off, void* pinned_host, int32_t n);
DSC_API int dsc_copy_wait_compute(dsc_ctx_t ctx);
DSC_API int dsc_compute_wait_copy(dsc_ctx_t ctx);
DSC_API int dsc_copy_sync(dsc_ctx_t ctx);
DSC_API int dsc_buf_fill(dsc_buf_t h, int32_t off, int32_t n, double v);
DSC_API int dsc_buf_view(dsc_buf_t h, int32_t off, int32_t n, dsc_buf_t* out);                /* GetValues: same storage */
DSC_API int dsc_buf_copy(dsc_buf_t h, dsc_buf_t* out);                                        /* DeepCopy; ReshapeCopy_MRows_V; ReshapeCopy_V_MRows (Backend.fs:109,134) */
DSC_API int dsc_buf_share(dsc_buf_t h, dsc_buf_t* out);                                       /* DeepCopy semantics, copy-on-write: the copy happens only if either handle is written in place later */
DSC_API int dsc_buf_gather2d(dsc_buf_t h, int32_t rows, int32_t cols, int32_t r0, int32_t r1,
                             int32_t c0, int32_t c1, dsc_buf_t* out);                         /* GetStackedValues (Util.fs:138,204-210), inclusive bounds */
DSC_API int dsc_buf_info(dsc_buf_t h, int* dtype, int32_t* length, int32_t* offset,
                         int32_t* alloc_length, void** device_ptr);
DSC_API int dsc_host_alloc_pinned(size_t bytes, void** out);
DSC_API int dsc_host_free_pinned(void* p);

/* ---------------------------------------------------------------- typed surface ------------ */
#define DSC_DECLARE_TYPED(P, T)                                                                                   \
    /* ---- BLAS-3 ----                                                                                        */ \
    /* C(m x n) = op(A)(m x k) * op(B)(k x n) [+ bias[i] on every column].  Mul_M_M (Backend.fs:119) is        */ \
    /* (ta,tb)=(0,0), bias=0; Mul_M_M_Add_V_MCols (Backend.fs:121) passes bias.  ta/tb=1 read the stored       */ \
    /* matrix transposed (k x m / n x k row-major): what Transpose_M followed by Mul_M_M computes              */ \
    /* (AD.Float32.fs:4127-4143) without materialising the transpose.                                          */ \
    DSC_API int dsc_##P##gemm(dsc_ctx_t ctx, int ta, int tb, int32_t m, int32_t n, int32_t k, dsc_buf_t A,        \
                              dsc_buf_t B, dsc_buf_t bias, dsc_buf_t* C);                                         \
    /* ---- BLAS-2 ----                                                                                        */ \
    /* out = op(A) x [+ y]; A is m x n row-major.  Mul_M_V (Backend.fs:98): trans=0; Mul_V_M (Backend.fs:101): */ \
    /* trans=1; Mul_M_V_Add_V (Backend.fs:99): trans=0 with y.                                                 */ \
    DSC_API int dsc_##P##gemv(dsc_ctx_t ctx, int trans, int32_t m, int32_t n, dsc_buf_t A, dsc_buf_t x,           \
                              dsc_buf_t y, dsc_buf_t* out);                                                       \
    /* out(m x n) = x y^T.  Mul_Out_V_V (Backend.fs:111)                                                       */ \
    DSC_API int dsc_##P##ger(dsc_ctx_t ctx, dsc_buf_t x, dsc_buf_t y, dsc_buf_t* out);                            \
    /* ---- scalar-valued (synchronise; deterministic fixed-shape tree) ----                                   */ \
    DSC_API int dsc_##P##dot(dsc_ctx_t ctx, dsc_buf_t x, dsc_buf_t y, T* out);   /* Mul_Dot_V_V  Backend.fs:81 */ \
    DSC_API int dsc_##P##asum(dsc_ctx_t ctx, dsc_buf_t x, T* out);               /* L1Norm_V     Backend.fs:82 */ \
    DSC_API int dsc_##P##nrm2(dsc_ctx_t ctx, dsc_buf_t x, T* out);               /* L2Norm_V     Backend.fs:83 */ \
    DSC_API int dsc_##P##amax(dsc_ctx_t ctx, dsc_buf_t x, T* out);               /* SupNorm_V    Backend.fs:84 */ \
    DSC_API int dsc_##P##sum(dsc_ctx_t ctx, dsc_buf_t x, T* out);                /* Sum_V/Sum_M  Backend.fs:85-86 */ \
    DSC_API int dsc_##P##argmax(dsc_ctx_t ctx, dsc_buf_t x, int32_t* out);       /* MaxIndex_V   Backend.fs:87 */ \
    DSC_API int dsc_##P##argmin(dsc_ctx_t ctx, dsc_buf_t x, int32_t* out);       /* MinIndex_V   Backend.fs:88 */ \
    /* same reductions leaving the scalar in a 1-element device buffer (no sync; used under graph capture)     */ \
    DSC_API int dsc_##P##sum_dev(dsc_ctx_t ctx, dsc_buf_t x, dsc_buf_t* out1);                                    \
    DSC_API int dsc_##P##dot_dev(dsc_ctx_t ctx, dsc_buf_t x, dsc_buf_t y, dsc_buf_t* out1);                       \
    /* ---- elementwise ----                                                                                   */ \
    /* y[i] = f_op(x[i]).  Map_F_V / Map_F_M (Backend.fs:105,130); only the op-code crosses, never a closure.  */ \
    DSC_API int dsc_##P##map(dsc_ctx_t ctx, int op, dsc_buf_t x, dsc_buf_t* y);                                   \
    /* y[i] = f_op(x[i], s) for *_Ab codes and DIV->s/x[i], POW_BA->s**x[i], ATAN2_BA->atan2(s,x[i]).         */ \
    /* Map_F_S_V / Map_F_S_M (Backend.fs:106,131)                                                              */ \
    DSC_API int dsc_##P##map_s(dsc_ctx_t ctx, int op, T s, dsc_buf_t x, dsc_buf_t* y);                            \
    /* y[i] = f_op(a[i], b[i]), op in {MUL, DIV, ADD, SUB, POW_AB, ATAN2_AB}.  Map2_F_V_V / Map2_F_M_M         */ \
    /* (Backend.fs:107,132), Add_V_V/Add_M_M (:90,112), Sub_V_V/Sub_M_M (:94,116), Mul_Had_M_M (:124).         */ \
    /* For ADD/SUB a zero-length operand is the additive identity (AD.Float32.fs:3919,4257).                   */ \
    DSC_API int dsc_##P##map2(dsc_ctx_t ctx, int op, dsc_buf_t a, dsc_buf_t b, dsc_buf_t* y);                     \
    /* y[i] = x[i] (+|-|*) s, see dsc_scalar_kind.                                                             */ \
    DSC_API int dsc_##P##scalar_op(dsc_ctx_t ctx, int kind, T s, dsc_buf_t x, dsc_buf_t* y);                      \
    /* dst[i] += src[i] over the whole windows; zero-length src is a no-op.  Add_M_M_InPlace (Backend.fs:113)  */ \
    DSC_API int dsc_##P##add_inplace(dsc_ctx_t ctx, dsc_buf_t dst, dsc_buf_t src);                                \
    /* dst[doff+i] += src[soff+i], i<n.  Add_V_V_InPlace(src, soff, dst, doff, n) (Backend.fs:91)              */ \
    DSC_API int dsc_##P##add_inplace_range(dsc_ctx_t ctx, dsc_buf_t src, int32_t soff, dsc_buf_t dst,             \
                                           int32_t doff, int32_t n);                                              \
    /* fused adjoint chains, see dsc_fused_kind: out = chain(dA, p)                                            */ \
    DSC_API int dsc_##P##fused_bwd(dsc_ctx_t ctx, int kind, dsc_buf_t dA, dsc_buf_t p, dsc_buf_t* out);           \
    /* z[i,j] = A[i,j] + v[j]  ==  Add_M_M(A, RepeatReshapeCopy_V_MRows(m, v))  (Backend.fs:112,135): how the  */ \
    /* fork broadcasts a bias (AD.Float32.fs:1860-1865) without materialising the m copies.  If op >= 0,       */ \
    /* additionally h = Map_F_M(op, z) (Backend.fs:130) from the same pass; h may be NULL when op < 0.          */ \
    DSC_API int dsc_##P##bias_rows_act(dsc_ctx_t ctx, int op, int32_t m, int32_t n, dsc_buf_t A, dsc_buf_t v,     \
                                       dsc_buf_t* z, dsc_buf_t* h);                                               \
    /* ---- layout ----                                                                                        */ \
    DSC_API int dsc_##P##transpose(dsc_ctx_t ctx, int32_t m, int32_t n, dsc_buf_t A, dsc_buf_t* out);             \
    /* Transpose_M (Backend.fs:127), bit exact                                                                 */ \
    DSC_API int dsc_##P##permute(dsc_ctx_t ctx, int rank, const int64_t* shape, const int32_t* perm,              \
                                 dsc_buf_t A, dsc_buf_t* out);                  /* Permute_M  Backend.fs:128 */  \
    DSC_API int dsc_##P##repeat_rows(dsc_ctx_t ctx, int32_t m, dsc_buf_t v, dsc_buf_t* out);                      \
    /* out[i,j]=v[j], i<m.  RepeatReshapeCopy_V_MRows (Backend.fs:135; golden: tests/DiffSharp.Tests/Script.fsx:7-16) */ \
    DSC_API int dsc_##P##repeat_cols(dsc_ctx_t ctx, int32_t n, dsc_buf_t v, dsc_buf_t* out);                      \
    /* out[i,j]=v[i], j<n.  RepeatReshapeCopy_V_MCols (Backend.fs:136)                                         */ \
    DSC_API int dsc_##P##colsum_inplace(dsc_ctx_t ctx, int32_t m, int32_t n, dsc_buf_t M, dsc_buf_t v);           \
    /* v[j] += sum_i M[i,j].  Add_M_Colwise_V_InPlace (Backend.fs:123; AD.Float32.fs:4255)                     */ \
    DSC_API int dsc_##P##add_v_mcols(dsc_ctx_t ctx, int32_t m, int32_t n, dsc_buf_t v, dsc_buf_t M,               \
                                     dsc_buf_t* out);                           /* Add_V_MCols Backend.fs:115 */ \
    DSC_API int dsc_##P##diag(dsc_ctx_t ctx, int32_t m, int32_t n, dsc_buf_t A, dsc_buf_t* out);                  \
    /* Diagonal_M (Backend.fs:104)                                                                             */ \
    /* ---- LAPACK-class (DSC_E_SINGULAR => managed None) ----                                                 */ \
    DSC_API int dsc_##P##gesv(dsc_ctx_t ctx, int32_t n, dsc_buf_t A, dsc_buf_t b, dsc_buf_t* x);                  \
    /* Solve_M_V (Backend.fs:102)                                                                              */ \
    DSC_API int dsc_##P##sysv(dsc_ctx_t ctx, int32_t n, dsc_buf_t A, dsc_buf_t b, dsc_buf_t* x);                  \
    /* SolveSymmetric_M_V (Backend.fs:103)                                                                     */ \
    DSC_API int dsc_##P##getri(dsc_ctx_t ctx, int32_t n, dsc_buf_t A, dsc_buf_t* Ainv);                           \
    /* Inverse_M (Backend.fs:125)                                                                              */ \
    DSC_API int dsc_##P##det(dsc_ctx_t ctx, int32_t n, dsc_buf_t A, T* out);    /* Det_M       Backend.fs:126 */ \
    /* ---- collectives (new): all-reduce = one peer-memory kernel over NVLink (NCCL when P2P is unavailable) ---- */ \
    DSC_API int dsc_##P##allreduce_sum(dsc_ctx_t ctx, int32_t count, const dsc_buf_t* bufs);                      \
    DSC_API int dsc_##P##allgather(dsc_ctx_t ctx, dsc_buf_t send, dsc_buf_t recv);

DSC_DECLARE_TYPED(s, float)
DSC_DECLARE_TYPED(d, double)

/* ---------------------------------------------------------------- communicator ------------- */
DSC_API int dsc_comm_unique_id(uint8_t out[128]);
DSC_API int dsc_comm_init(dsc_ctx_t ctx, int nranks, int rank, const uint8_t id[128]);
DSC_API int dsc_comm_join(dsc_ctx_t ctx); /* compute stream waits for everything queued on the comm stream */
/* %globaltimer (ns) at the phase boundaries of the last peer-memory all-reduce as CTA 0 saw them:
 * start, packed, barrier 1, reduced + broadcast, barrier 2, unpacked (tools/allreduce_bench.py); the flag-in-data kernel
 * has no barriers: stamps 1 = 2 and 3 = 4. */
DSC_API int dsc_comm_debug_stamps(dsc_ctx_t ctx, uint64_t out[6]);
/* the same for the last call (back = 0) or the one before it (back = 1): a step issues two all-reduces (dist.py) */
DSC_API int dsc_comm_debug_stamps_at(dsc_ctx_t ctx, int back, uint64_t out[6]);
DSC_API int dsc_comm_destroy(dsc_ctx_t ctx);

#ifdef __cplusplus
}
#endif
#endif /* DIFFSHARP_CUDA_H */
