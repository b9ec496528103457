// K4/K5/K6: op-code driven elementwise kernels (unary map, scalar-parameter map, binary map,
// scalar broadcast, in-place accumulate, fused adjoint chains).
//
// The reference passes an F# closure next to a MapOp op-code (Backend.fs:105-107,130-132); the
// closure text at each call site is the semantic spec of the op-code (AD.Float32.fs:1388-1594,
// 2631-2934).  Here only the op-code reaches the device: each op-code is one template
// instantiation, the host switch picks it.
//
// All kernels are HBM-bound streaming kernels: 128-bit loads/stores (float4 / double2) when every
// operand window is 16-byte aligned, 4 independent vectors in flight per thread, one CTA per
// contiguous chunk (see stream_kernel for the measurement behind that choice).
//
// The entry points at the bottom are also where the adjoint chains of the unchanged AD layer are recognised
// (dsc_lazy.cu): Map(Cosh) / Map(Sign) / scalar broadcasts / Hadamard products return deferred values, and the
// Hadamard product that completes a chain launches ONE kernel with the chain's own sequence of roundings.
#include "dsc_internal.cuh"
#include "dsc_ops.cuh"

namespace dsc {

// ---- vector helpers -----------------------------------------------------------------------
template <typename T> struct Vec;
template <> struct Vec<float> {
    using type = float4;
    static constexpr int N = 4;
};
template <> struct Vec<double> {
    using type = double2;
    static constexpr int N = 2;
};

template <typename T>
__device__ __forceinline__ typename Vec<T>::type ldv(const T* p) {
    return *reinterpret_cast<const typename Vec<T>::type*>(p);
}
template <typename T>
__device__ __forceinline__ void stv(T* p, typename Vec<T>::type v) {
    *reinterpret_cast<typename Vec<T>::type*>(p) = v;
}
__device__ __forceinline__ float& lane(float4& v, int i) { return (&v.x)[i]; }
__device__ __forceinline__ double& lane(double2& v, int i) { return (&v.x)[i]; }

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

// Generic streaming kernel.  F is a functor (T a, T b) -> T; NIN = number of array inputs (1 or 2).
// VEC selects the 128-bit path (all windows 16 B aligned); the scalar path handles misaligned
// windows (views at odd offsets, Util.fs:202) and the tail.  LO: additionally write the operand-split
// companion of y (float32; dsc_lo_of) for the tensor-core GEMM that will read y.
//
// Launch shape (measured on B200 with tools/ew_bench.cu, 3 x 64 MiB streams): one CTA per
// CONTIGUOUS chunk of kThreads x kUnroll vectors and a grid that covers the array reaches 5.46 TB/s;
// a persistent grid-stride loop over 8 CTAs/SM tops out at 5.12 TB/s and cudaMemcpy D2D at 4.68 TB/s.
// So: no grid-stride loop, kUnroll independent 128-bit loads in flight per input per thread.
template <typename T, typename F, int NIN, bool VEC, bool LO>
__global__ void __launch_bounds__(kThreads) stream_kernel(F f, const T* __restrict__ a, const T* __restrict__ b,
                                                           T* __restrict__ y, T* __restrict__ ylo, size_t n) {
    if constexpr (VEC) {
        using V = typename Vec<T>::type;
        constexpr int W = Vec<T>::N;
        const size_t nvec = n / W;
        const size_t base = (size_t)blockIdx.x * (kThreads * kUnroll) + threadIdx.x;
        V va[kUnroll], vb[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const size_t i = base + (size_t)u * kThreads;
            if (i < nvec) {
                va[u] = ldv<T>(a + i * W);
                if constexpr (NIN == 2) vb[u] = ldv<T>(b + i * W);
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const size_t i = base + (size_t)u * kThreads;
            if (i < nvec) {
                V r, l;
#pragma unroll
                for (int e = 0; e < W; ++e) {
                    lane(r, e) = f(lane(va[u], e), NIN == 2 ? lane(vb[u], e) : T(0));
                    if constexpr (LO) lane(l, e) = dsc_lo_of(lane(r, e));
                }
                stv<T>(y + i * W, r);
                if constexpr (LO) stv<T>(ylo + i * W, l);
            }
        }
        // scalar tail (< W elements), handled by the first threads of the first CTA
        if (blockIdx.x == 0) {
            const size_t j = nvec * W + threadIdx.x;
            if (j < n) {
                const T r = f(a[j], NIN == 2 ? b[j] : T(0));
                y[j] = r;
                if constexpr (LO) ylo[j] = dsc_lo_of(r);
            }
        }
    } else {
        const size_t base = (size_t)blockIdx.x * (kThreads * kUnroll) + threadIdx.x;
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const size_t j = base + (size_t)u * kThreads;
            if (j < n) {
                const T r = f(a[j], NIN == 2 ? b[j] : T(0));
                y[j] = r;
                if constexpr (LO) ylo[j] = dsc_lo_of(r);
            }
        }
    }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <typename T, typename F, int NIN>
static int launch_stream(Ctx* ctx, F f, const T* a, const T* b, T* y, size_t n, T* ylo = nullptr) {
    if (n == 0) return DSC_OK;
    bool vec = aligned16(a) && aligned16(y) && (NIN == 1 || aligned16(b)) && (!ylo || aligned16(ylo));
    size_t items = vec ? (n / Vec<T>::N > 0 ? n / Vec<T>::N : 1) : n;
    size_t grid = (items + (size_t)kThreads * kUnroll - 1) / ((size_t)kThreads * kUnroll);
    if (ylo) {
        if (vec) stream_kernel<T, F, NIN, true, true><<<(unsigned)grid, kThreads, 0, ctx->stream>>>(f, a, b, y, ylo, n);
        else stream_kernel<T, F, NIN, false, true><<<(unsigned)grid, kThreads, 0, ctx->stream>>>(f, a, b, y, ylo, n);
    } else {
        if (vec) stream_kernel<T, F, NIN, true, false><<<(unsigned)grid, kThreads, 0, ctx->stream>>>(f, a, b, y, nullptr, n);
        else stream_kernel<T, F, NIN, false, false><<<(unsigned)grid, kThreads, 0, ctx->stream>>>(f, a, b, y, nullptr, n);
    }
    DSC_LAUNCH_CHECK(ctx);
    return DSC_OK;
}

// ---- functors -----------------------------------------------------------------------------
template <int OP, typename T>
struct UnaryF {
    __device__ __forceinline__ T operator()(T a, T) const { return unary_op<OP>(a); }
};
template <int OP, typename T>
struct BinaryF {
    __device__ __forceinline__ T operator()(T a, T b) const { return binary_op<OP>(a, b); }
};
// scalar-parameter map: element is always the FIRST operand of the *_Ab codes.
template <int OP, typename T>
struct ScalarMapF {
    T s;
    __device__ __forceinline__ T operator()(T v, T) const {
        if constexpr (OP == DSC_OP_DIV) return s / v;  // fun v -> a / v  (AD.Float32.fs:2496)
        else return binary_op<OP>(v, s);               // POW_AB: v**s, POW_BA: s**v, ATAN2_AB: atan2 v s, ATAN2_BA: atan2 s v
    }
};
template <int KIND, typename T>
struct ScalarOpF {
    T s;
    __device__ __forceinline__ T operator()(T v, T) const {
        if constexpr (KIND == DSC_S_ADD) return v + s;
        else if constexpr (KIND == DSC_S_SUB_SX) return s - v;
        else if constexpr (KIND == DSC_S_SUB_XS) return v - s;
        else return s * v;
    }
};
template <typename T>
struct NegF {
    __device__ __forceinline__ T operator()(T v, T) const { return -v; }
};
// 1 / cosh(v): Map(Cosh) rounded to T, then Map_F_S(1, Div) — the two roundings of the unfused pair (AD.Float32.fs:4238)
template <typename T>
struct SechF {
    __device__ __forceinline__ T operator()(T v, T) const { return T(1) / m_cosh(v); }
};
// Fused adjoint chains.  Each reproduces the reference's sequence of roundings:
//  tanh   : secha = 1/cosh(aP); (dA*secha)*secha                      (AD.Float32.fs:4238-4240)
//  sigmoid: (dA*dP)*(1-dP)                                            (AD.Float32.fs:4281)
//  relu   : dA*((sign(aP)+1)*(1/2))  -> derivative at 0 is 1/2        (AD.Float32.fs:4280)
template <int KIND, typename T>
struct FusedBwdF {
    __device__ __forceinline__ T operator()(T dA, T p) const { return chain_op<KIND>(dA, p); }
};

// ---- raw executors ------------------------------------------------------------------------
template <typename T>
int ew_map(Ctx* ctx, int op, const T* xp, T* yp, size_t n) {
    switch (op) {
#define CASE(OPC) \
    case OPC: return launch_stream<T, UnaryF<OPC, T>, 1>(ctx, UnaryF<OPC, T>(), xp, nullptr, yp, n);
        CASE(DSC_OP_LOG) CASE(DSC_OP_LOG10) CASE(DSC_OP_EXP) CASE(DSC_OP_SIN) CASE(DSC_OP_COS) CASE(DSC_OP_TAN)
        CASE(DSC_OP_SQRT) CASE(DSC_OP_SINH) CASE(DSC_OP_COSH) CASE(DSC_OP_TANH) CASE(DSC_OP_ASIN) CASE(DSC_OP_ACOS)
        CASE(DSC_OP_ATAN) CASE(DSC_OP_ABS) CASE(DSC_OP_SIGN) CASE(DSC_OP_FLOOR) CASE(DSC_OP_CEILING)
        CASE(DSC_OP_ROUND) CASE(DSC_OP_REL) CASE(DSC_OP_SIGMOID)
#undef CASE
    }
    return fail(DSC_E_UNSUPPORTED_OP, "Map_F: op-code %d is not a unary MapOp (closures are not executed on this backend)", op);
}

template <typename T>
int ew_map_s(Ctx* ctx, int op, T s, const T* xp, T* yp, size_t n) {
    switch (op) {
#define CASE(OPC) \
    case OPC: return launch_stream<T, ScalarMapF<OPC, T>, 1>(ctx, ScalarMapF<OPC, T>{s}, xp, nullptr, yp, n);
        CASE(DSC_OP_DIV) CASE(DSC_OP_POW_AB) CASE(DSC_OP_POW_BA) CASE(DSC_OP_ATAN2_AB) CASE(DSC_OP_ATAN2_BA)
#undef CASE
    }
    return fail(DSC_E_UNSUPPORTED_OP, "Map_F_S: op-code %d has no scalar-parameter form", op);
}

template <typename T>
int ew_map2(Ctx* ctx, int op, const T* ap, const T* bp, T* yp, size_t n) {
    switch (op) {
#define CASE(OPC) \
    case OPC: return launch_stream<T, BinaryF<OPC, T>, 2>(ctx, BinaryF<OPC, T>(), ap, bp, yp, n);
        CASE(DSC_OP_MUL) CASE(DSC_OP_DIV) CASE(DSC_OP_ADD) CASE(DSC_OP_SUB) CASE(DSC_OP_POW_AB) CASE(DSC_OP_ATAN2_AB)
#undef CASE
    }
    return fail(DSC_E_UNSUPPORTED_OP, "Map2_F: op-code %d is not a binary MapOp", op);
}

template <typename T>
int ew_scalar(Ctx* ctx, int kind, T s, const T* xp, T* yp, size_t n) {
    switch (kind) {
        case DSC_S_ADD: return launch_stream<T, ScalarOpF<DSC_S_ADD, T>, 1>(ctx, ScalarOpF<DSC_S_ADD, T>{s}, xp, nullptr, yp, n);
        case DSC_S_SUB_SX: return launch_stream<T, ScalarOpF<DSC_S_SUB_SX, T>, 1>(ctx, ScalarOpF<DSC_S_SUB_SX, T>{s}, xp, nullptr, yp, n);
        case DSC_S_SUB_XS: return launch_stream<T, ScalarOpF<DSC_S_SUB_XS, T>, 1>(ctx, ScalarOpF<DSC_S_SUB_XS, T>{s}, xp, nullptr, yp, n);
        case DSC_S_MUL: return launch_stream<T, ScalarOpF<DSC_S_MUL, T>, 1>(ctx, ScalarOpF<DSC_S_MUL, T>{s}, xp, nullptr, yp, n);
    }
    return fail(DSC_E_UNSUPPORTED_OP, "scalar_op: unknown kind %d", kind);
}

template <typename T>
int ew_neg(Ctx* ctx, const T* x, T* y, size_t n) { return launch_stream<T, NegF<T>, 1>(ctx, NegF<T>(), x, nullptr, y, n); }

template <typename T>
int ew_sech(Ctx* ctx, const T* x, T* y, size_t n) { return launch_stream<T, SechF<T>, 1>(ctx, SechF<T>(), x, nullptr, y, n); }

template <typename T>
int ew_fused_bwd(Ctx* ctx, int kind, const T* dA, const T* p, T* y, T* ylo, size_t n) {
    if (sizeof(T) != 4) ylo = nullptr;
    switch (kind) {
        case DSC_FUSED_TANH_BWD: return launch_stream<T, FusedBwdF<DSC_FUSED_TANH_BWD, T>, 2>(ctx, FusedBwdF<DSC_FUSED_TANH_BWD, T>(), dA, p, y, n, ylo);
        case DSC_FUSED_SIGMOID_BWD: return launch_stream<T, FusedBwdF<DSC_FUSED_SIGMOID_BWD, T>, 2>(ctx, FusedBwdF<DSC_FUSED_SIGMOID_BWD, T>(), dA, p, y, n, ylo);
        case DSC_FUSED_RELU_BWD: return launch_stream<T, FusedBwdF<DSC_FUSED_RELU_BWD, T>, 2>(ctx, FusedBwdF<DSC_FUSED_RELU_BWD, T>(), dA, p, y, n, ylo);
    }
    return fail(DSC_E_UNSUPPORTED_OP, "fused_bwd: unknown kind %d", kind);
}

// z = A (bop) rows(v) [; h = f_op(z)] — one pass over A instead of broadcast + binary op (+ map).
// Same roundings as the unfused sequence: z is the fp op, h applies the map to the ROUNDED z.
// LO: the operand-split companion of the last output (h when ACT, else z).
template <typename T, int BOP, int OP, bool ACT, bool VEC, bool LO>
__global__ void __launch_bounds__(kThreads) rows_kernel(const T* __restrict__ A, const T* __restrict__ v, T* __restrict__ z,
                                                         T* __restrict__ h, T* __restrict__ lo, size_t total, int n) {
    constexpr int W = VEC ? Vec<T>::N : 1;
    const size_t nitems = total / W;
    const size_t base = (size_t)blockIdx.x * (kThreads * kUnroll) + threadIdx.x;
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
        const size_t i = base + (size_t)u * kThreads;
        if (i >= nitems) continue;
        if constexpr (VEC) {
            using V = typename Vec<T>::type;
            V a = ldv<T>(A + i * W);
            const int col = (int)((i * W) % (size_t)n);  // n % W == 0: a vector never straddles two rows
            V b = ldv<T>(v + col);
            V zz, hh, ll;
#pragma unroll
            for (int l = 0; l < W; ++l) {
                lane(zz, l) = binary_op<BOP>(lane(a, l), lane(b, l));
                if constexpr (ACT) lane(hh, l) = unary_op<OP>(lane(zz, l));
                if constexpr (LO) lane(ll, l) = dsc_lo_of(ACT ? lane(hh, l) : lane(zz, l));
            }
            stv<T>(z + i * W, zz);
            if constexpr (ACT) stv<T>(h + i * W, hh);
            if constexpr (LO) stv<T>(lo + i * W, ll);
        } else {
            const T zz = binary_op<BOP>(A[i], v[i % (size_t)n]);
            z[i] = zz;
            T last = zz;
            if constexpr (ACT) { last = unary_op<OP>(zz); h[i] = last; }
            if constexpr (LO) lo[i] = dsc_lo_of(last);
        }
    }
}

template <typename T, int BOP, int OP, bool ACT>
static int launch_rows(Ctx* ctx, const T* A, const T* v, T* z, T* h, T* lo, int m, int n) {
    const size_t total = (size_t)m * n;
    const bool vec = (n % Vec<T>::N == 0) && aligned16(A) && aligned16(v) && aligned16(z) && (!ACT || aligned16(h)) && (!lo || aligned16(lo));
    const size_t items = vec ? total / Vec<T>::N : total;
    const unsigned grid = (unsigned)((items + (size_t)kThreads * kUnroll - 1) / ((size_t)kThreads * kUnroll));
    if (lo) {
        if (vec) rows_kernel<T, BOP, OP, ACT, true, true><<<grid, kThreads, 0, ctx->stream>>>(A, v, z, h, lo, total, n);
        else rows_kernel<T, BOP, OP, ACT, false, true><<<grid, kThreads, 0, ctx->stream>>>(A, v, z, h, lo, total, n);
    } else {
        if (vec) rows_kernel<T, BOP, OP, ACT, true, false><<<grid, kThreads, 0, ctx->stream>>>(A, v, z, h, nullptr, total, n);
        else rows_kernel<T, BOP, OP, ACT, false, false><<<grid, kThreads, 0, ctx->stream>>>(A, v, z, h, nullptr, total, n);
    }
    DSC_LAUNCH_CHECK(ctx);
    return DSC_OK;
}

template <typename T>
int ew_rows(Ctx* ctx, int bop, int act, int m, int n, const T* A, const T* v, T* z, T* h, T* lo) {
    if ((size_t)m * n == 0) return DSC_OK;
    if (sizeof(T) != 4) lo = nullptr;
    if (bop != DSC_OP_ADD) {
        if (act >= 0) return fail(DSC_E_UNSUPPORTED_OP, "rows: an activation is fused with the row-broadcast ADD only");
        switch (bop) {
            case DSC_OP_SUB: return launch_rows<T, DSC_OP_SUB, 0, false>(ctx, A, v, z, nullptr, lo, m, n);
            case DSC_OP_MUL: return launch_rows<T, DSC_OP_MUL, 0, false>(ctx, A, v, z, nullptr, lo, m, n);
            case DSC_OP_DIV: return launch_rows<T, DSC_OP_DIV, 0, false>(ctx, A, v, z, nullptr, lo, m, n);
        }
        return fail(DSC_E_UNSUPPORTED_OP, "rows: binary op-code %d has no row-broadcast form", bop);
    }
    switch (act) {
#define CASE(OPC) \
    case OPC: return launch_rows<T, DSC_OP_ADD, OPC, true>(ctx, A, v, z, h, lo, m, n);
        CASE(DSC_OP_LOG) CASE(DSC_OP_LOG10) CASE(DSC_OP_EXP) CASE(DSC_OP_SIN) CASE(DSC_OP_COS) CASE(DSC_OP_TAN)
        CASE(DSC_OP_SQRT) CASE(DSC_OP_SINH) CASE(DSC_OP_COSH) CASE(DSC_OP_TANH) CASE(DSC_OP_ASIN) CASE(DSC_OP_ACOS)
        CASE(DSC_OP_ATAN) CASE(DSC_OP_ABS) CASE(DSC_OP_SIGN) CASE(DSC_OP_FLOOR) CASE(DSC_OP_CEILING)
        CASE(DSC_OP_ROUND) CASE(DSC_OP_REL) CASE(DSC_OP_SIGMOID)
#undef CASE
        default: return launch_rows<T, DSC_OP_ADD, 0, false>(ctx, A, v, z, nullptr, lo, m, n);
    }
}

#define DSC_EW_INST(T)                                                                            \
    template int ew_map<T>(Ctx*, int, const T*, T*, size_t);                                      \
    template int ew_map_s<T>(Ctx*, int, T, const T*, T*, size_t);                                 \
    template int ew_map2<T>(Ctx*, int, const T*, const T*, T*, size_t);                           \
    template int ew_scalar<T>(Ctx*, int, T, const T*, T*, size_t);                                \
    template int ew_neg<T>(Ctx*, const T*, T*, size_t);                                           \
    template int ew_sech<T>(Ctx*, const T*, T*, size_t);                                          \
    template int ew_fused_bwd<T>(Ctx*, int, const T*, const T*, T*, T*, size_t);                  \
    template int ew_rows<T>(Ctx*, int, int, int, int, const T*, const T*, T*, T*, T*);
DSC_EW_INST(float)
DSC_EW_INST(double)

// ---- helpers of the entry points ------------------------------------------------------------
// New handle holding a fresh, uninitialised block of n elements.
template <typename T>
static int new_out(Ctx* ctx, int32_t n, dsc_buf_t* out, Buf** res) {
    *out = 0;
    return out_buf(ctx, dtype_of<T>::value, n, out, res);
}

// Should a kernel that produces an m x n float32 matrix also write its operand-split companion?  Yes when a tensor-core GEMM
// could take it as an operand in place (dsc_gemm_tf32x3.cu: TMA needs 16-byte alignment and a contiguous extent that is a
// multiple of 4; the GEMM is tensor-core eligible from 128 x 32 x 32).  The companion costs 4 bytes per element in this pass
// and saves a separate pass of 8 (read + write) plus a launch in front of every GEMM that reads the matrix.
static bool want_lo(Ctx* ctx, int dtype, int m, int n) {
    if (dtype != DSC_F32 || (ctx->flags & DSC_FLAG_GEMM_FP32_SIMT)) return false;
    static const bool off = [] { const char* e = getenv("DSC_EMIT_LO"); return e && *e == '0'; }();
    if (off) return false;
    return m >= 128 && n >= 32 && n % 4 == 0;
}

// Attach a fresh companion block to `blk` (window [off, off + len) valid once the launching kernel has run).
static float* attach_lo(Ctx* ctx, Block* blk, size_t off, size_t len) {
    if (!blk->lo) {
        Block* lo = nullptr;
        if (pool_alloc(ctx, blk->bytes, &lo) != DSC_OK) return nullptr;   // no companion: the GEMM will split on first use
        blk->lo = lo;
    } else if (lane_gate_write(ctx, blk->lo) != DSC_OK) {   // an old companion the second compute lane may still be reading
        return nullptr;
    }
    blk->lo_epoch = ctx->cache_epoch;
    blk->lo_begin = off;
    blk->lo_end = off + len;
    return reinterpret_cast<float*>(blk->lo->ptr) + off;
}

template <typename T>
static T* maybe_lo(Ctx* ctx, Buf* y, int m, int n) {
    if constexpr (sizeof(T) == 4) {
        if (y && y->block && want_lo(ctx, y->dtype, m, n)) return reinterpret_cast<T*>(attach_lo(ctx, y->block, y->off, (size_t)y->len));
    }
    return nullptr;
}

// operand -> plain device window (materialising a deferred operand)
template <typename T>
static int opnd_ptr(Ctx* ctx, Operand& o, const T** p) {
    DSC_TRY(operand_resolve(ctx, o));
    *p = o.node ? reinterpret_cast<const T*>(o.node->result->ptr) + o.node->result_off
                : (o.block ? reinterpret_cast<const T*>(o.block->ptr) + o.off : nullptr);
    return DSC_OK;
}
static Operand opnd_of(Buf* b) {
    Operand o;
    if (b && b->len > 0) {
        o.len = b->len;
        if (b->block) { o.block = b->block; o.off = b->off; }
        else o.node = b->lazy;
    }
    return o;
}

// out := ROWS(vec) — a deferred m x n matrix of m copies of the (already computed) vector `vec`
static int rows_of(Ctx* ctx, Buf* vec, int m, int n, dsc_buf_t* out) {
    LazyNode* nd = lazy_new(ctx, LZ_ROWS, vec->dtype, m * n);
    nd->m = m; nd->n = n;
    lazy_set_operand(nd, 0, vec);
    return lazy_buf(ctx, nd, out);
}

// out := copy-on-write alias of src's value (x + 0, 1 .* x, ...)
static int alias_of(Ctx* ctx, Buf* src, dsc_buf_t* out) {
    Buf* v = new Buf();
    v->dtype = src->dtype;
    v->ctx = ctx;
    v->len = src->len;
    int s = buf_share_into(v, src);
    if (s == 1) {   // true views on the block: eager copy
        delete v;
        *out = 0;
        return dsc_buf_copy(reinterpret_cast<dsc_buf_t>(src), out);
    }
    *out = reinterpret_cast<dsc_buf_t>(v);
    return DSC_OK;
}

// The adjoint chain `kind` over (dA, p): one kernel; dA may be a deferred GEMM, whose epilogue then applies the chain.
template <typename T>
static int run_chain(Ctx* ctx, int kind, Operand dA, Operand p, int32_t n, dsc_buf_t* out) {
    Buf* y;
    const T *pa, *pp;
    DSC_TRY(opnd_ptr<T>(ctx, p, &pp));
    LazyNode* g = (dA.node && dA.node->kind == LZ_GEMM && !dA.node->result) ? dA.node : nullptr;
    if (g && g->len == n) {
        // dZ = (dH W^T) .* f'(z): the product's epilogue applies the chain, the product itself is never stored
        DSC_TRY(new_out<T>(ctx, n, out, &y));
        T* lo = maybe_lo<T>(ctx, y, g->m, g->n);
        for (int i = 0; i < g->nsrc; ++i) DSC_TRY(operand_resolve(ctx, g->src[i]));
        GemmEpilogue ep;
        ep.mode = GEMM_EP_CHAIN; ep.op = kind; ep.aux = pp; ep.lo = reinterpret_cast<float*>(lo);
        auto mat = [](const Operand& o) {
            MatRef r;
            r.block = o.node ? o.node->result : o.block;
            r.off = o.node ? o.node->result_off : o.off;
            r.ptr = r.block ? reinterpret_cast<const T*>(r.block->ptr) + r.off : nullptr;
            return r;
        };
        const T* bias = nullptr;
        if (g->nsrc > 2) DSC_TRY(opnd_ptr<T>(ctx, g->src[2], &bias));
        g->consumed = true;
        return gemm_run<T>(ctx, g->ta, g->tb, g->m, g->n, g->k, mat(g->src[0]), mat(g->src[1]), bias, y->ptr<T>(), &ep);
    }
    DSC_TRY(opnd_ptr<T>(ctx, dA, &pa));
    DSC_TRY(new_out<T>(ctx, n, out, &y));
    // the shape is not known here (buffers only): a companion is written when the length alone makes a GEMM operand plausible
    T* lo = nullptr;
    if constexpr (sizeof(T) == 4) {
        if (n >= 128 * 32 && n % 4 == 0 && want_lo(ctx, DSC_F32, 128, 32)) lo = reinterpret_cast<T*>(attach_lo(ctx, y->block, y->off, (size_t)n));
    }
    return ew_fused_bwd<T>(ctx, kind, pa, pp, y->ptr<T>(), lo, (size_t)n);
}

// ---- typed entry points -------------------------------------------------------------------
template <typename T>
static int map_impl(dsc_ctx_t c, int op, dsc_buf_t xh, dsc_buf_t* yh) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *x, *y;
    DSC_TRY(in_buf_peek(ctx, dtype_of<T>::value, xh, &x, "map: x"));
    if (op < DSC_OP_LOG || op >= DSC_OP__COUNT)
        return fail(DSC_E_UNSUPPORTED_OP, "Map_F: op-code %d is not a unary MapOp (closures are not executed on this backend)", op);
    if (!yh) return fail(DSC_E_INVALID, "null output handle pointer");
    if (ctx->fuse && *yh == 0 && x->len > 0) {
        if (LazyNode* ar = lazy_of(x, LZ_ADD_ROWS)) {
            // z = A + rows(v) is still pending: produce z and h = f(z) in ONE pass (AD.Float32.fs:2357 -> :2704); when A is a
            // deferred product, in the product's epilogue
            Buf *zb, *hb;
            dsc_buf_t zh = 0;
            DSC_TRY(new_out<T>(ctx, x->len, &zh, &zb));
            int s = new_out<T>(ctx, x->len, yh, &hb);
            if (s != DSC_OK) { dsc_buf_free(zh); return s; }
            T* lo = maybe_lo<T>(ctx, hb, ar->m, ar->n);
            const T* vp;
            s = opnd_ptr<T>(ctx, ar->src[1], &vp);
            LazyNode* g = (s == DSC_OK && ar->src[0].node && ar->src[0].node->kind == LZ_GEMM && !ar->src[0].node->result) ? ar->src[0].node : nullptr;
            if (s == DSC_OK && g) {
                for (int i = 0; i < g->nsrc && s == DSC_OK; ++i) s = operand_resolve(ctx, g->src[i]);
                if (s == DSC_OK) {
                    GemmEpilogue ep;
                    ep.mode = GEMM_EP_BIAS_ACT; ep.op = op; ep.vec = vp; ep.out2 = hb->ptr<T>(); ep.lo = reinterpret_cast<float*>(lo);
                    auto mat = [](const Operand& o) {
                        MatRef r;
                        r.block = o.node ? o.node->result : o.block;
                        r.off = o.node ? o.node->result_off : o.off;
                        r.ptr = r.block ? reinterpret_cast<const T*>(r.block->ptr) + r.off : nullptr;
                        return r;
                    };
                    const T* bias = nullptr;
                    if (g->nsrc > 2) s = opnd_ptr<T>(ctx, g->src[2], &bias);
                    g->consumed = true;
                    if (s == DSC_OK) s = gemm_run<T>(ctx, g->ta, g->tb, g->m, g->n, g->k, mat(g->src[0]), mat(g->src[1]), bias, zb->ptr<T>(), &ep);
                }
            } else if (s == DSC_OK) {
                const T* ap;
                s = opnd_ptr<T>(ctx, ar->src[0], &ap);
                if (s == DSC_OK) s = ew_rows<T>(ctx, DSC_OP_ADD, op, ar->m, ar->n, ap, vp, zb->ptr<T>(), hb->ptr<T>(), lo);
            }
            if (s == DSC_OK) {
                lazy_set_result(ar, zb->block, zb->off);   // every handle of the pending sum now resolves to z
                ar->consumed = true;
            } else {
                dsc_buf_free(*yh);
                *yh = 0;
            }
            dsc_buf_free(zh);
            return s;
        }
        if (LazyNode* r = lazy_of(x, LZ_ROWS)) {
            // a map of m identical rows is m copies of the map of one row (forward-mode primal of a matrix tangent)
            Buf* vy;
            dsc_buf_t vh = 0;
            const T* vp;
            DSC_TRY(opnd_ptr<T>(ctx, r->src[0], &vp));
            DSC_TRY(new_out<T>(ctx, r->n, &vh, &vy));
            int s = ew_map<T>(ctx, op, vp, vy->ptr<T>(), (size_t)r->n);
            if (s == DSC_OK) s = rows_of(ctx, vy, r->m, r->n, yh);
            dsc_buf_free(vh);
            return s;
        }
        if (op == DSC_OP_COSH || op == DSC_OP_SIGN) {
            // possible head of the tanh (Cosh, AD.Float32.fs:4238) or ReLU / Abs (Sign, :4280,:4244) adjoint chain: defer
            LazyNode* nd = lazy_new(ctx, LZ_MAP, x->dtype, x->len);
            nd->op = op;
            lazy_set_operand(nd, 0, x);
            return lazy_buf(ctx, nd, yh);
        }
    }
    DSC_TRY(buf_resolve(x));
    DSC_TRY(out_buf(ctx, dtype_of<T>::value, x->len, yh, &y));
    return ew_map<T>(ctx, op, x->ptr<T>(), y->ptr<T>(), (size_t)x->len);
}

template <typename T>
static int map_s_impl(dsc_ctx_t c, int op, T s, dsc_buf_t xh, dsc_buf_t* yh) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *x, *y;
    DSC_TRY(in_buf_peek(ctx, dtype_of<T>::value, xh, &x, "map_s: x"));
    if (!(op == DSC_OP_DIV || op == DSC_OP_POW_AB || op == DSC_OP_POW_BA || op == DSC_OP_ATAN2_AB || op == DSC_OP_ATAN2_BA))
        return fail(DSC_E_UNSUPPORTED_OP, "Map_F_S: op-code %d has no scalar-parameter form", op);
    if (!yh) return fail(DSC_E_INVALID, "null output handle pointer");
    if (ctx->fuse && *yh == 0 && x->len > 0) {
        LazyNode* mp = lazy_of(x, LZ_MAP);
        if (mp && mp->op == DSC_OP_COSH && op == DSC_OP_DIV && s == T(1)) {
            // secha = 1 / cosh(aP)  (AD.Float32.fs:4238): deferred on aP itself
            LazyNode* nd = lazy_new(ctx, LZ_SECH, x->dtype, x->len);
            lazy_set_operand(nd, 0, mp->src[0]);
            mp->consumed = true;
            return lazy_buf(ctx, nd, yh);
        }
        if (LazyNode* r = lazy_of(x, LZ_ROWS)) {
            Buf* vy;
            dsc_buf_t vh = 0;
            const T* vp;
            DSC_TRY(opnd_ptr<T>(ctx, r->src[0], &vp));
            DSC_TRY(new_out<T>(ctx, r->n, &vh, &vy));
            int st = ew_map_s<T>(ctx, op, s, vp, vy->ptr<T>(), (size_t)r->n);
            if (st == DSC_OK) st = rows_of(ctx, vy, r->m, r->n, yh);
            dsc_buf_free(vh);
            return st;
        }
    }
    DSC_TRY(buf_resolve(x));
    DSC_TRY(out_buf(ctx, dtype_of<T>::value, x->len, yh, &y));
    return ew_map_s<T>(ctx, op, s, x->ptr<T>(), y->ptr<T>(), (size_t)x->len);
}

// Had(a, b) as the LAST call of an adjoint chain?  Fires the chain's kernel and returns DSC_OK, or returns 1.
template <typename T>
static int try_chain(Ctx* ctx, Buf* a, Buf* b, dsc_buf_t* yh) {
    // the reference writes d.A on the left (AD.Float32.fs:4240,4280-4281); accept the factors in either order
    for (int swap = 0; swap < 2; ++swap) {
        Buf* l = swap ? b : a;
        Buf* r = swap ? a : b;
        LazyNode* ln = lazy_of(l);
        LazyNode* rn = lazy_of(r);
        if (!rn || rn->result) continue;
        // tanh:  (dA .* secha) .* secha,  secha = 1 / cosh(aP)
        if (rn->kind == LZ_SECH && ln && ln->kind == LZ_HAD && !ln->result) {
            for (int side = 0; side < 2; ++side) {
                LazyNode* inner = ln->src[side].node;
                if (inner && inner->kind == LZ_SECH && !inner->result && inner->src[0].same(rn->src[0])) {
                    ln->consumed = rn->consumed = inner->consumed = true;
                    return run_chain<T>(ctx, DSC_FUSED_TANH_BWD, ln->src[1 - side], rn->src[0], l->len, yh);
                }
            }
        }
        // sigmoid:  (dA .* dP) .* (1 - dP)
        if (rn->kind == LZ_SCALAR && rn->op == DSC_S_SUB_SX && rn->s == 1.0 && ln && ln->kind == LZ_HAD && !ln->result) {
            for (int side = 0; side < 2; ++side)
                if (ln->src[side].same(rn->src[0])) {
                    ln->consumed = rn->consumed = true;
                    return run_chain<T>(ctx, DSC_FUSED_SIGMOID_BWD, ln->src[1 - side], rn->src[0], l->len, yh);
                }
        }
        // ReLU:  dA .* ((Sign(aP) + 1) * (1/2))
        if (rn->kind == LZ_SCALAR && rn->op == DSC_S_MUL && rn->s == 0.5) {
            LazyNode* add = rn->src[0].node;
            if (add && add->kind == LZ_SCALAR && add->op == DSC_S_ADD && add->s == 1.0 && !add->result) {
                LazyNode* sg = add->src[0].node;
                if (sg && sg->kind == LZ_MAP && sg->op == DSC_OP_SIGN && !sg->result) {
                    rn->consumed = add->consumed = sg->consumed = true;
                    return run_chain<T>(ctx, DSC_FUSED_RELU_BWD, opnd_of(l), sg->src[0], l->len, yh);
                }
            }
        }
    }
    return 1;
}

template <typename T>
static int map2_impl(dsc_ctx_t c, int op, dsc_buf_t ah, dsc_buf_t bh, dsc_buf_t* yh) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *a, *b, *y;
    DSC_TRY(in_buf_peek(ctx, dtype_of<T>::value, ah, &a, "map2: a"));
    DSC_TRY(in_buf_peek(ctx, dtype_of<T>::value, bh, &b, "map2: b"));
    if (!(op == DSC_OP_MUL || op == DSC_OP_DIV || op == DSC_OP_ADD || op == DSC_OP_SUB || op == DSC_OP_POW_AB || op == DSC_OP_ATAN2_AB))
        return fail(DSC_E_UNSUPPORTED_OP, "Map2_F: op-code %d is not a binary MapOp", op);
    if (!yh) return fail(DSC_E_INVALID, "null output handle pointer");
    if (ctx->fuse && *yh == 0 && a->len == b->len && a->len > 0) {
        const int32_t n = a->len;
        LazyNode *na = lazy_of(a), *nb = lazy_of(b);
        if (na && na->result) na = nullptr;
        if (nb && nb->result) nb = nullptr;
        const bool za = na && na->kind == LZ_ZERO, zb = nb && nb->kind == LZ_ZERO;
        if (op == DSC_OP_ADD || op == DSC_OP_SUB) {
            // known-zero adjoints (two per tape node): x +/- 0 is x, 0 + x is x — only the sign of a zero can differ
            if (zb) return za ? dsc_buf_alloc_fill(c, a->dtype, n, 0.0, yh) : alias_of(ctx, a, yh);
            if (za && op == DSC_OP_ADD) return alias_of(ctx, b, yh);
            LazyNode *ra = (na && na->kind == LZ_ROWS) ? na : nullptr, *rb = (nb && nb->kind == LZ_ROWS) ? nb : nullptr;
            if (ra && rb && ra->m == rb->m && ra->n == rb->n) {
                // rows(u) +/- rows(v) = rows(u +/- v)
                Buf* vy;
                dsc_buf_t vh = 0;
                const T *pu, *pv;
                DSC_TRY(opnd_ptr<T>(ctx, ra->src[0], &pu));
                DSC_TRY(opnd_ptr<T>(ctx, rb->src[0], &pv));
                DSC_TRY(new_out<T>(ctx, ra->n, &vh, &vy));
                int s = ew_map2<T>(ctx, op, pu, pv, vy->ptr<T>(), (size_t)ra->n);
                if (s == DSC_OK) s = rows_of(ctx, vy, ra->m, ra->n, yh);
                dsc_buf_free(vh);
                return s;
            }
            if (op == DSC_OP_ADD && (ra != nullptr) != (rb != nullptr)) {
                // A + rows(v): the fork's bias broadcast (AD.Float32.fs:1860-1865 -> :2357); fp add commutes bit for bit
                LazyNode* r = ra ? ra : rb;
                Buf* full = ra ? b : a;
                if ((int64_t)r->m * r->n == n) {
                    LazyNode* nd = lazy_new(ctx, LZ_ADD_ROWS, a->dtype, n);
                    nd->m = r->m; nd->n = r->n;
                    lazy_set_operand(nd, 0, full);
                    lazy_set_operand(nd, 1, r->src[0]);
                    r->consumed = true;
                    return lazy_buf(ctx, nd, yh);
                }
            }
        }
        if (op == DSC_OP_MUL) {
            // the all-ones seed of Sum_DM's adjoint times E (AD.Float32.fs:3430 -> :4144): c .* x
            for (int swap = 0; swap < 2; ++swap) {
                LazyNode* f = swap ? nb : na;
                Buf* other = swap ? a : b;
                if (f && f->kind == LZ_FILL) {
                    if (f->s == 1.0) return alias_of(ctx, other, yh);
                    DSC_TRY(buf_resolve(other));
                    DSC_TRY(new_out<T>(ctx, n, yh, &y));
                    return ew_scalar<T>(ctx, DSC_S_MUL, (T)f->s, other->ptr<T>(), y->ptr<T>(), (size_t)n);
                }
            }
            int s = try_chain<T>(ctx, a, b, yh);
            if (s <= 0) return s;
            LazyNode *ra = (na && na->kind == LZ_ROWS) ? na : nullptr, *rb = (nb && nb->kind == LZ_ROWS) ? nb : nullptr;
            if (ra && rb && ra->m == rb->m && ra->n == rb->n) {
                Buf* vy;
                dsc_buf_t vh = 0;
                const T *pu, *pv;
                DSC_TRY(opnd_ptr<T>(ctx, ra->src[0], &pu));
                DSC_TRY(opnd_ptr<T>(ctx, rb->src[0], &pv));
                DSC_TRY(new_out<T>(ctx, ra->n, &vh, &vy));
                s = ew_map2<T>(ctx, op, pu, pv, vy->ptr<T>(), (size_t)ra->n);
                if (s == DSC_OK) s = rows_of(ctx, vy, ra->m, ra->n, yh);
                dsc_buf_free(vh);
                return s;
            }
            if (!za && !zb && !ra && !rb) {
                // a Hadamard product may be the middle of an adjoint chain, or the argument of Sum (had(E, E).Sum()): defer
                LazyNode* nd = lazy_new(ctx, LZ_HAD, a->dtype, n);
                lazy_set_operand(nd, 0, a);
                lazy_set_operand(nd, 1, b);
                return lazy_buf(ctx, nd, yh);
            }
        }
        if (op == DSC_OP_SUB || op == DSC_OP_MUL || op == DSC_OP_DIV) {
            // A (op) rows(v) without materialising the m copies (forward-mode tangent rules over a replicated primal)
            LazyNode* rb = (nb && nb->kind == LZ_ROWS) ? nb : nullptr;
            if (rb && !(na && na->kind == LZ_ROWS) && (int64_t)rb->m * rb->n == n) {
                const T* pv;
                DSC_TRY(opnd_ptr<T>(ctx, rb->src[0], &pv));
                DSC_TRY(buf_resolve(a));
                DSC_TRY(new_out<T>(ctx, n, yh, &y));
                rb->consumed = true;
                return ew_rows<T>(ctx, op, -1, rb->m, rb->n, a->ptr<T>(), pv, y->ptr<T>(), nullptr, nullptr);
            }
        }
    }
    // x + empty, empty + x, x - empty (the reverse sweep pushes DVector.Zero / DNDArray.Zero as "no contribution" and then adds
    // it, AD.Float32.fs:3919,3956,4257): the result is x — a copy-on-write alias instead of a copy
    if (ctx->fuse && *yh == 0 && a->len != b->len && ((b->len == 0 && (op == DSC_OP_ADD || op == DSC_OP_SUB)) || (a->len == 0 && op == DSC_OP_ADD)))
        return alias_of(ctx, a->len ? a : b, yh);
    DSC_TRY(buf_resolve(a));
    DSC_TRY(buf_resolve(b));
    // Empty operand = additive identity: the reverse sweep pushes DVector.Zero / DNDArray.Zero as "no
    // contribution" and then adds it (AD.Float32.fs:3919,3956,4257).
    if ((op == DSC_OP_ADD || op == DSC_OP_SUB) && a->len != b->len && (a->len == 0 || b->len == 0)) {
        Buf* src = a->len ? a : b;
        DSC_TRY(out_buf(ctx, dtype_of<T>::value, src->len, yh, &y));
        if (op == DSC_OP_SUB && src == b)  // 0 - b
            return ew_neg<T>(ctx, src->ptr<T>(), y->ptr<T>(), (size_t)src->len);
        if (y->ptr<T>() != src->ptr<T>()) {
            DSC_CUDA_CHECK(cudaMemcpyAsync(y->ptr<T>(), src->ptr<T>(), (size_t)src->len * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
            ctx->stats.kernel_launches++;
        }
        return DSC_OK;
    }
    if (a->len != b->len) return fail(DSC_E_SHAPE, "Map2_F: operands have lengths %d and %d", a->len, b->len);
    DSC_TRY(out_buf(ctx, dtype_of<T>::value, a->len, yh, &y));
    return ew_map2<T>(ctx, op, a->ptr<T>(), b->ptr<T>(), y->ptr<T>(), (size_t)a->len);
}

template <typename T>
static int scalar_op_impl(dsc_ctx_t c, int kind, T s, dsc_buf_t xh, dsc_buf_t* yh) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *x, *y;
    DSC_TRY(in_buf_peek(ctx, dtype_of<T>::value, xh, &x, "scalar_op: x"));
    if (kind < DSC_S_ADD || kind > DSC_S_MUL) return fail(DSC_E_UNSUPPORTED_OP, "scalar_op: unknown kind %d", kind);
    if (!yh) return fail(DSC_E_INVALID, "null output handle pointer");
    if (ctx->fuse && *yh == 0 && x->len > 0) {
        LazyNode* nx = lazy_of(x);
        if (nx && nx->result) nx = nullptr;
        if (nx && nx->kind == LZ_ROWS) {
            Buf* vy;
            dsc_buf_t vh = 0;
            const T* vp;
            DSC_TRY(opnd_ptr<T>(ctx, nx->src[0], &vp));
            DSC_TRY(new_out<T>(ctx, nx->n, &vh, &vy));
            int st = ew_scalar<T>(ctx, kind, s, vp, vy->ptr<T>(), (size_t)nx->n);
            if (st == DSC_OK) st = rows_of(ctx, vy, nx->m, nx->n, yh);
            dsc_buf_free(vh);
            return st;
        }
        // 1 - dP (sigmoid adjoint, AD.Float32.fs:4281); Sign(aP) + 1 and (.) * (1/2) (ReLU adjoint, :4280): possible chain links
        const bool sig = kind == DSC_S_SUB_SX && s == T(1);
        const bool relu1 = kind == DSC_S_ADD && s == T(1) && nx && nx->kind == LZ_MAP && nx->op == DSC_OP_SIGN;
        const bool relu2 = kind == DSC_S_MUL && s == T(0.5) && nx && nx->kind == LZ_SCALAR && nx->op == DSC_S_ADD && nx->s == 1.0 &&
                           nx->src[0].node && nx->src[0].node->kind == LZ_MAP && nx->src[0].node->op == DSC_OP_SIGN;
        if (sig || relu1 || relu2) {
            LazyNode* nd = lazy_new(ctx, LZ_SCALAR, x->dtype, x->len);
            nd->op = kind;
            nd->s = (double)s;
            lazy_set_operand(nd, 0, x);
            return lazy_buf(ctx, nd, yh);
        }
    }
    DSC_TRY(buf_resolve(x));
    DSC_TRY(out_buf(ctx, dtype_of<T>::value, x->len, yh, &y));
    return ew_scalar<T>(ctx, kind, s, x->ptr<T>(), y->ptr<T>(), (size_t)x->len);
}

// dst += src where dst must give up its block anyway (copy-on-write aliases exist): ONE pass new = old + src instead of a
// copy followed by an in-place add.
template <typename T>
static int add_out_of_place(Ctx* ctx, Buf* d, const T* sp, size_t s_off_in_d, size_t n) {
    Block* old = d->block;
    Block* nb = nullptr;
    DSC_TRY(pool_alloc(ctx, (size_t)d->len * sizeof(T), &nb));
    const T* op = d->ptr<T>();
    T* np = reinterpret_cast<T*>(nb->ptr);
    int s = DSC_OK;
    if (s_off_in_d == 0 && n == (size_t)d->len) {
        s = ew_map2<T>(ctx, DSC_OP_ADD, op, sp, np, n);
    } else {
        cudaError_t e = cudaMemcpyAsync(np, op, (size_t)d->len * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream);
        if (e != cudaSuccess) s = fail(DSC_E_CUDA, "copy-on-write copy failed: %s", cudaGetErrorString(e));
        if (s == DSC_OK) s = ew_map2<T>(ctx, DSC_OP_ADD, np + s_off_in_d, sp, np + s_off_in_d, n);
    }
    if (s != DSC_OK) { block_release(nb); return s; }
    if (d->is_cow) old->cow--;
    d->is_cow = false;
    block_release(old);
    d->block = nb;
    d->off = 0;
    return DSC_OK;
}

template <typename T>
static int add_inplace_impl(dsc_ctx_t c, dsc_buf_t dh, dsc_buf_t sh) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *d, *s;
    DSC_TRY(in_buf_peek(ctx, dtype_of<T>::value, dh, &d, "add_inplace: dst"));
    DSC_TRY(in_buf_peek(ctx, dtype_of<T>::value, sh, &s, "add_inplace: src"));
    if (s->len == 0) return DSC_OK;  // DNDArray.Zero push (AD.Float32.fs:4245-4248,4272)
    if (s->len != d->len) return fail(DSC_E_SHAPE, "Add_M_M_InPlace: lengths %d and %d", d->len, s->len);
    if (lazy_of(s, LZ_ZERO)) return DSC_OK;   // x + 0
    if (LazyNode* z = lazy_of(d, LZ_ZERO)) {
        // first accumulation into a known-zero adjoint (AD.Float32.fs:4118 after :3746): 0 + v = v — the adjoint becomes a
        // copy-on-write alias of the addend; the copy happens only if either side is written in place later
        Buf tmp;
        tmp.dtype = d->dtype; tmp.ctx = ctx; tmp.len = d->len;
        int st = buf_share_into(&tmp, s);
        if (st == DSC_OK) {
            d->lazy = nullptr;
            lazy_unref(z);
            d->block = tmp.block; d->off = tmp.off; d->is_cow = tmp.is_cow; d->lazy = tmp.lazy;
            return DSC_OK;
        }
        if (st < 0) return st;
        // the addend's block has true views: fall through to zeros + add
    }
    DSC_TRY(buf_resolve(s));
    DSC_TRY(buf_resolve(d));
    // The accumulator shares its block with other holders (it became an alias of its first addend: E + E in the loss tail)
    // and would have to be detached before the write anyway: ONE pass new = old + addend into a fresh block.  The old block
    // is not written, so deferred operations that still read it need not run first.
    if (d->block && !d->is_view && d->block->views == 0 && d->block->cow > 0 && d->block->refs > 1)
        return add_out_of_place<T>(ctx, d, s->ptr<T>(), 0, (size_t)d->len);
    DSC_TRY(buf_writable(ctx, d));
    return ew_map2<T>(ctx, DSC_OP_ADD, d->ptr<T>(), s->ptr<T>(), d->ptr<T>(), (size_t)d->len);
}

template <typename T>
static int add_inplace_range_impl(dsc_ctx_t c, dsc_buf_t sh, int32_t soff, dsc_buf_t dh, int32_t doff, int32_t n) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *d, *s;
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, dh, &d, "add_inplace_range: dst"));
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, sh, &s, "add_inplace_range: src"));
    if (n < 0 || soff < 0 || doff < 0 || (int64_t)soff + n > s->len || (int64_t)doff + n > d->len)
        return fail(DSC_E_SHAPE, "Add_V_V_InPlace: range out of bounds (src %d+%d of %d, dst %d+%d of %d)", soff, n, s->len, doff, n, d->len);
    if (n == 0) return DSC_OK;
    DSC_TRY(buf_writable(ctx, d));
    return ew_map2<T>(ctx, DSC_OP_ADD, d->ptr<T>() + doff, s->ptr<T>() + soff, d->ptr<T>() + doff, (size_t)n);
}

template <typename T>
static int fused_bwd_impl(dsc_ctx_t c, int kind, dsc_buf_t dAh, dsc_buf_t ph, dsc_buf_t* outh) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *dA, *p, *y;
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, dAh, &dA, "fused_bwd: dA"));
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, ph, &p, "fused_bwd: p"));
    if (dA->len != p->len) return fail(DSC_E_SHAPE, "fused_bwd: lengths %d and %d", dA->len, p->len);
    if (kind < DSC_FUSED_TANH_BWD || kind > DSC_FUSED_RELU_BWD) return fail(DSC_E_UNSUPPORTED_OP, "fused_bwd: unknown kind %d", kind);
    DSC_TRY(out_buf(ctx, dtype_of<T>::value, dA->len, outh, &y));
    return ew_fused_bwd<T>(ctx, kind, dA->ptr<T>(), p->ptr<T>(), y->ptr<T>(), nullptr, (size_t)dA->len);
}

template <typename T>
static int bias_rows_act_impl(dsc_ctx_t c, int op, int32_t m, int32_t n, dsc_buf_t Ah, dsc_buf_t vh, dsc_buf_t* zh, dsc_buf_t* hh) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *A, *v, *z, *h = nullptr;
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, Ah, &A, "bias_rows_act: A"));
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, vh, &v, "bias_rows_act: v"));
    if (m < 0 || n < 0 || (int64_t)m * n != A->len) return fail(DSC_E_SHAPE, "Matrices must have same shape.");
    if (v->len != n) return fail(DSC_E_SHAPE, "bias_rows_act: vector length %d, matrix has %d columns", v->len, n);
    if (op >= 0 && (op < DSC_OP_LOG || op >= DSC_OP__COUNT)) return fail(DSC_E_UNSUPPORTED_OP, "bias_rows_act: op-code %d is not a unary MapOp", op);
    DSC_TRY(out_buf(ctx, dtype_of<T>::value, m * n, zh, &z));
    if (op >= 0) DSC_TRY(out_buf(ctx, dtype_of<T>::value, m * n, hh, &h));
    if (m * n == 0) return DSC_OK;
    return ew_rows<T>(ctx, DSC_OP_ADD, op, m, n, A->ptr<T>(), v->ptr<T>(), z->ptr<T>(), h ? h->ptr<T>() : nullptr, nullptr);
}

}  // namespace dsc

using namespace dsc;

extern "C" {
#define DSC_EW_API(P, T)                                                                                                   \
    int dsc_##P##map(dsc_ctx_t c, int op, dsc_buf_t x, dsc_buf_t* y) { DSC_NVTX(c, "Map_F_V/Map_F_M"); return map_impl<T>(c, op, x, y); }                  \
    int dsc_##P##map_s(dsc_ctx_t c, int op, T s, dsc_buf_t x, dsc_buf_t* y) { DSC_NVTX(c, "Map_F_S_V/Map_F_S_M"); return map_s_impl<T>(c, op, s, x, y); }      \
    int dsc_##P##map2(dsc_ctx_t c, int op, dsc_buf_t a, dsc_buf_t b, dsc_buf_t* y) { DSC_NVTX(c, "Map2_F/Add/Sub/Mul_Had"); return map2_impl<T>(c, op, a, b, y); } \
    int dsc_##P##scalar_op(dsc_ctx_t c, int kind, T s, dsc_buf_t x, dsc_buf_t* y) { DSC_NVTX(c, "Add_S/Sub_S/Mul_S"); return scalar_op_impl<T>(c, kind, s, x, y); } \
    int dsc_##P##add_inplace(dsc_ctx_t c, dsc_buf_t d, dsc_buf_t s) { DSC_NVTX(c, "Add_M_M_InPlace"); return add_inplace_impl<T>(c, d, s); }               \
    int dsc_##P##add_inplace_range(dsc_ctx_t c, dsc_buf_t s, int32_t so, dsc_buf_t d, int32_t dof, int32_t n) { DSC_NVTX(c, "Add_V_V_InPlace");            \
        return add_inplace_range_impl<T>(c, s, so, d, dof, n);                                                             \
    }                                                                                                                      \
    int dsc_##P##fused_bwd(dsc_ctx_t c, int kind, dsc_buf_t dA, dsc_buf_t p, dsc_buf_t* out) { DSC_NVTX(c, "fused_bwd"); return fused_bwd_impl<T>(c, kind, dA, p, out); } \
    int dsc_##P##bias_rows_act(dsc_ctx_t c, int op, int32_t m, int32_t n, dsc_buf_t A, dsc_buf_t v, dsc_buf_t* z, dsc_buf_t* h) { DSC_NVTX(c, "bias_rows_act"); \
        return bias_rows_act_impl<T>(c, op, m, n, A, v, z, h);                                                             \
    }
DSC_EW_API(s, float)
DSC_EW_API(d, double)
}
