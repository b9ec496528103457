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
#include "dsc_internal.cuh"

namespace dsc {

// ---- per-element functions ----------------------------------------------------------------
// signummod (Util.fs:88-91): -1 / 0 / +1, NaN -> 0 (both comparisons false).
template <typename T>
__device__ __forceinline__ T sgn(T x) {
    return x < T(0) ? T(-1) : (x > T(0) ? T(1) : T(0));
}

// .NET Math.Round = banker's rounding (half to even) == rint in the default rounding mode.
__device__ __forceinline__ float round_even(float x) { return rintf(x); }
__device__ __forceinline__ double round_even(double x) { return rint(x); }

// F# `max 0 v` (AD.Float32.fs:2927) propagates NaN from either operand.
template <typename T>
__device__ __forceinline__ T relu(T v) {
    return (v != v) ? v : (v > T(0) ? v : T(0));
}

#define DSC_MATH1(name, f32, f64)                                         \
    __device__ __forceinline__ float m_##name(float x) { return f32(x); } \
    __device__ __forceinline__ double m_##name(double x) { return f64(x); }
DSC_MATH1(log, logf, log)
DSC_MATH1(log10, log10f, log10)
DSC_MATH1(exp, expf, exp)
DSC_MATH1(sin, sinf, sin)
DSC_MATH1(cos, cosf, cos)
DSC_MATH1(tan, tanf, tan)
DSC_MATH1(sqrt, sqrtf, sqrt)
DSC_MATH1(sinh, sinhf, sinh)
DSC_MATH1(cosh, coshf, cosh)
DSC_MATH1(tanh, tanhf, tanh)
DSC_MATH1(asin, asinf, asin)
DSC_MATH1(acos, acosf, acos)
DSC_MATH1(atan, atanf, atan)
DSC_MATH1(abs, fabsf, fabs)
DSC_MATH1(floor, floorf, floor)
DSC_MATH1(ceil, ceilf, ceil)
#undef DSC_MATH1
__device__ __forceinline__ float m_pow(float a, float b) { return powf(a, b); }
__device__ __forceinline__ double m_pow(double a, double b) { return pow(a, b); }
__device__ __forceinline__ float m_atan2(float a, float b) { return atan2f(a, b); }
__device__ __forceinline__ double m_atan2(double a, double b) { return atan2(a, b); }

template <int OP, typename T>
__device__ __forceinline__ T unary_op(T v) {
    if constexpr (OP == DSC_OP_LOG) return m_log(v);
    else if constexpr (OP == DSC_OP_LOG10) return m_log10(v);
    else if constexpr (OP == DSC_OP_EXP) return m_exp(v);
    else if constexpr (OP == DSC_OP_SIN) return m_sin(v);
    else if constexpr (OP == DSC_OP_COS) return m_cos(v);
    else if constexpr (OP == DSC_OP_TAN) return m_tan(v);
    else if constexpr (OP == DSC_OP_SQRT) return m_sqrt(v);
    else if constexpr (OP == DSC_OP_SINH) return m_sinh(v);
    else if constexpr (OP == DSC_OP_COSH) return m_cosh(v);
    else if constexpr (OP == DSC_OP_TANH) return m_tanh(v);
    else if constexpr (OP == DSC_OP_ASIN) return m_asin(v);
    else if constexpr (OP == DSC_OP_ACOS) return m_acos(v);
    else if constexpr (OP == DSC_OP_ATAN) return m_atan(v);
    else if constexpr (OP == DSC_OP_ABS) return m_abs(v);
    else if constexpr (OP == DSC_OP_SIGN) return sgn(v);
    else if constexpr (OP == DSC_OP_FLOOR) return m_floor(v);
    else if constexpr (OP == DSC_OP_CEILING) return m_ceil(v);
    else if constexpr (OP == DSC_OP_ROUND) return round_even(v);
    else if constexpr (OP == DSC_OP_REL) return relu(v);
    else if constexpr (OP == DSC_OP_SIGMOID) return T(1) / (T(1) + m_exp(-v));  // AD.Float32.fs:2934
    else return v;
}

// Binary op-codes.  For the scalar-parameter form the *_Ab codes are f(elem, scalar) and the *_Ba
// codes f(scalar, elem) (AD.Float32.fs:2556,2568,2580,2591); DIV is scalar / elem (:2496).
template <int OP, typename T>
__device__ __forceinline__ T binary_op(T a, T b) {
    if constexpr (OP == DSC_OP_MUL) return a * b;
    else if constexpr (OP == DSC_OP_DIV) return a / b;
    else if constexpr (OP == DSC_OP_ADD) return a + b;
    else if constexpr (OP == DSC_OP_SUB) return a - b;
    else if constexpr (OP == DSC_OP_POW_AB) return m_pow(a, b);
    else if constexpr (OP == DSC_OP_POW_BA) return m_pow(b, a);
    else if constexpr (OP == DSC_OP_ATAN2_AB) return m_atan2(a, b);
    else if constexpr (OP == DSC_OP_ATAN2_BA) return m_atan2(b, a);
    else return a;
}

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
// windows (views at odd offsets, Util.fs:202) and the tail.
//
// Launch shape (measured on B200 with tools/ew_bench.cu, 3 x 64 MiB streams): one CTA per
// CONTIGUOUS chunk of kThreads x kUnroll vectors and a grid that covers the array reaches 5.46 TB/s;
// a persistent grid-stride loop over 8 CTAs/SM tops out at 5.12 TB/s and cudaMemcpy D2D at 4.68 TB/s.
// So: no grid-stride loop, kUnroll independent 128-bit loads in flight per input per thread.
template <typename T, typename F, int NIN, bool VEC>
__global__ void __launch_bounds__(kThreads) stream_kernel(F f, const T* __restrict__ a, const T* __restrict__ b,
                                                           T* __restrict__ y, size_t n) {
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
                V r;
#pragma unroll
                for (int l = 0; l < W; ++l) lane(r, l) = f(lane(va[u], l), NIN == 2 ? lane(vb[u], l) : T(0));
                stv<T>(y + i * W, r);
            }
        }
        // scalar tail (< W elements), handled by the first threads of the first CTA
        if (blockIdx.x == 0) {
            const size_t j = nvec * W + threadIdx.x;
            if (j < n) y[j] = f(a[j], NIN == 2 ? b[j] : T(0));
        }
    } else {
        const size_t base = (size_t)blockIdx.x * (kThreads * kUnroll) + threadIdx.x;
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const size_t j = base + (size_t)u * kThreads;
            if (j < n) y[j] = f(a[j], NIN == 2 ? b[j] : T(0));
        }
    }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <typename T, typename F, int NIN>
static int launch_stream(Ctx* ctx, F f, const T* a, const T* b, T* y, size_t n) {
    if (n == 0) return DSC_OK;
    bool vec = aligned16(a) && aligned16(y) && (NIN == 1 || aligned16(b));
    size_t items = vec ? (n / Vec<T>::N > 0 ? n / Vec<T>::N : 1) : n;
    size_t grid = (items + (size_t)kThreads * kUnroll - 1) / ((size_t)kThreads * kUnroll);
    if (vec)
        stream_kernel<T, F, NIN, true><<<(unsigned)grid, kThreads, 0, ctx->stream>>>(f, a, b, y, n);
    else
        stream_kernel<T, F, NIN, false><<<(unsigned)grid, kThreads, 0, ctx->stream>>>(f, a, b, y, n);
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
// Fused adjoint chains.  Each reproduces the reference's sequence of roundings:
//  tanh   : secha = 1/cosh(aP); (dA*secha)*secha                      (AD.Float32.fs:4238-4240)
//  sigmoid: (dA*dP)*(1-dP)                                            (AD.Float32.fs:4281)
//  relu   : dA*((sign(aP)+1)/2)  -> derivative at 0 is 1/2            (AD.Float32.fs:4280)
template <int KIND, typename T>
struct FusedBwdF {
    __device__ __forceinline__ T operator()(T dA, T p) const {
        if constexpr (KIND == DSC_FUSED_TANH_BWD) {
            T secha = T(1) / m_cosh(p);
            return (dA * secha) * secha;
        } else if constexpr (KIND == DSC_FUSED_SIGMOID_BWD) {
            return (dA * p) * (T(1) - p);
        } else {
            return dA * ((sgn(p) + T(1)) * T(0.5));
        }
    }
};

// ---- typed entry points -------------------------------------------------------------------
template <typename T>
static int map_impl(dsc_ctx_t c, int op, dsc_buf_t xh, dsc_buf_t* yh) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *x, *y;
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, xh, &x, "map: x"));
    if (op < DSC_OP_LOG || op >= DSC_OP__COUNT)
        return fail(DSC_E_UNSUPPORTED_OP, "Map_F: op-code %d is not a unary MapOp (closures are not executed on this backend)", op);
    DSC_TRY(out_buf(ctx, dtype_of<T>::value, x->len, yh, &y));
    const T* xp = x->ptr<T>();
    T* yp = y->ptr<T>();
    size_t n = x->len;
    switch (op) {
#define CASE(OPC) \
    case OPC: return launch_stream<T, UnaryF<OPC, T>, 1>(ctx, UnaryF<OPC, T>(), xp, nullptr, yp, n);
        CASE(DSC_OP_LOG) CASE(DSC_OP_LOG10) CASE(DSC_OP_EXP) CASE(DSC_OP_SIN) CASE(DSC_OP_COS) CASE(DSC_OP_TAN)
        CASE(DSC_OP_SQRT) CASE(DSC_OP_SINH) CASE(DSC_OP_COSH) CASE(DSC_OP_TANH) CASE(DSC_OP_ASIN) CASE(DSC_OP_ACOS)
        CASE(DSC_OP_ATAN) CASE(DSC_OP_ABS) CASE(DSC_OP_SIGN) CASE(DSC_OP_FLOOR) CASE(DSC_OP_CEILING)
        CASE(DSC_OP_ROUND) CASE(DSC_OP_REL) CASE(DSC_OP_SIGMOID)
#undef CASE
    }
    return fail(DSC_E_UNSUPPORTED_OP, "Map_F: unknown op-code %d", op);
}

template <typename T>
static int map_s_impl(dsc_ctx_t c, int op, T s, dsc_buf_t xh, dsc_buf_t* yh) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *x, *y;
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, xh, &x, "map_s: x"));
    if (!(op == DSC_OP_DIV || op == DSC_OP_POW_AB || op == DSC_OP_POW_BA || op == DSC_OP_ATAN2_AB || op == DSC_OP_ATAN2_BA))
        return fail(DSC_E_UNSUPPORTED_OP, "Map_F_S: op-code %d has no scalar-parameter form", op);
    DSC_TRY(out_buf(ctx, dtype_of<T>::value, x->len, yh, &y));
    const T* xp = x->ptr<T>();
    T* yp = y->ptr<T>();
    size_t n = x->len;
    switch (op) {
#define CASE(OPC) \
    case OPC: return launch_stream<T, ScalarMapF<OPC, T>, 1>(ctx, ScalarMapF<OPC, T>{s}, xp, nullptr, yp, n);
        CASE(DSC_OP_DIV) CASE(DSC_OP_POW_AB) CASE(DSC_OP_POW_BA) CASE(DSC_OP_ATAN2_AB) CASE(DSC_OP_ATAN2_BA)
#undef CASE
    }
    return fail(DSC_E_UNSUPPORTED_OP, "Map_F_S: unknown op-code %d", op);
}

template <typename T>
static int map2_impl(dsc_ctx_t c, int op, dsc_buf_t ah, dsc_buf_t bh, dsc_buf_t* yh) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *a, *b, *y;
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, ah, &a, "map2: a"));
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, bh, &b, "map2: b"));
    if (!(op == DSC_OP_MUL || op == DSC_OP_DIV || op == DSC_OP_ADD || op == DSC_OP_SUB || op == DSC_OP_POW_AB || op == DSC_OP_ATAN2_AB))
        return fail(DSC_E_UNSUPPORTED_OP, "Map2_F: op-code %d is not a binary MapOp", op);
    // Empty operand = additive identity: the reverse sweep pushes DVector.Zero / DNDArray.Zero as "no
    // contribution" and then adds it (AD.Float32.fs:3919,3956,4257).
    if ((op == DSC_OP_ADD || op == DSC_OP_SUB) && a->len != b->len && (a->len == 0 || b->len == 0)) {
        Buf* src = a->len ? a : b;
        DSC_TRY(out_buf(ctx, dtype_of<T>::value, src->len, yh, &y));
        if (op == DSC_OP_SUB && src == b)  // 0 - b
            return launch_stream<T, NegF<T>, 1>(ctx, NegF<T>(), src->ptr<T>(), nullptr, y->ptr<T>(), src->len);
        if (y->ptr<T>() != src->ptr<T>()) {
            DSC_CUDA_CHECK(cudaMemcpyAsync(y->ptr<T>(), src->ptr<T>(), (size_t)src->len * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
            ctx->stats.kernel_launches++;
        }
        return DSC_OK;
    }
    if (a->len != b->len) return fail(DSC_E_SHAPE, "Map2_F: operands have lengths %d and %d", a->len, b->len);
    DSC_TRY(out_buf(ctx, dtype_of<T>::value, a->len, yh, &y));
    const T *ap = a->ptr<T>(), *bp = b->ptr<T>();
    T* yp = y->ptr<T>();
    size_t n = a->len;
    switch (op) {
#define CASE(OPC) \
    case OPC: return launch_stream<T, BinaryF<OPC, T>, 2>(ctx, BinaryF<OPC, T>(), ap, bp, yp, n);
        CASE(DSC_OP_MUL) CASE(DSC_OP_DIV) CASE(DSC_OP_ADD) CASE(DSC_OP_SUB) CASE(DSC_OP_POW_AB) CASE(DSC_OP_ATAN2_AB)
#undef CASE
    }
    return fail(DSC_E_UNSUPPORTED_OP, "Map2_F: unknown op-code %d", op);
}

template <typename T>
static int scalar_op_impl(dsc_ctx_t c, int kind, T s, dsc_buf_t xh, dsc_buf_t* yh) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *x, *y;
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, xh, &x, "scalar_op: x"));
    if (kind < DSC_S_ADD || kind > DSC_S_MUL) return fail(DSC_E_UNSUPPORTED_OP, "scalar_op: unknown kind %d", kind);
    DSC_TRY(out_buf(ctx, dtype_of<T>::value, x->len, yh, &y));
    const T* xp = x->ptr<T>();
    T* yp = y->ptr<T>();
    size_t n = x->len;
    switch (kind) {
        case DSC_S_ADD: return launch_stream<T, ScalarOpF<DSC_S_ADD, T>, 1>(ctx, ScalarOpF<DSC_S_ADD, T>{s}, xp, nullptr, yp, n);
        case DSC_S_SUB_SX: return launch_stream<T, ScalarOpF<DSC_S_SUB_SX, T>, 1>(ctx, ScalarOpF<DSC_S_SUB_SX, T>{s}, xp, nullptr, yp, n);
        case DSC_S_SUB_XS: return launch_stream<T, ScalarOpF<DSC_S_SUB_XS, T>, 1>(ctx, ScalarOpF<DSC_S_SUB_XS, T>{s}, xp, nullptr, yp, n);
        default: return launch_stream<T, ScalarOpF<DSC_S_MUL, T>, 1>(ctx, ScalarOpF<DSC_S_MUL, T>{s}, xp, nullptr, yp, n);
    }
}

template <typename T>
static int add_inplace_impl(dsc_ctx_t c, dsc_buf_t dh, dsc_buf_t sh) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *d, *s;
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, dh, &d, "add_inplace: dst"));
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, sh, &s, "add_inplace: src"));
    if (s->len == 0) return DSC_OK;  // DNDArray.Zero push (AD.Float32.fs:4245-4248,4272)
    if (s->len != d->len) return fail(DSC_E_SHAPE, "Add_M_M_InPlace: lengths %d and %d", d->len, s->len);
    DSC_TRY(buf_writable(ctx, d));
    return launch_stream<T, BinaryF<DSC_OP_ADD, T>, 2>(ctx, BinaryF<DSC_OP_ADD, T>(), d->ptr<T>(), s->ptr<T>(), d->ptr<T>(), d->len);
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
    return launch_stream<T, BinaryF<DSC_OP_ADD, T>, 2>(ctx, BinaryF<DSC_OP_ADD, T>(), d->ptr<T>() + doff, s->ptr<T>() + soff, d->ptr<T>() + doff, n);
}

template <typename T>
static int fused_bwd_impl(dsc_ctx_t c, int kind, dsc_buf_t dAh, dsc_buf_t ph, dsc_buf_t* outh) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *dA, *p, *y;
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, dAh, &dA, "fused_bwd: dA"));
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, ph, &p, "fused_bwd: p"));
    if (dA->len != p->len) return fail(DSC_E_SHAPE, "fused_bwd: lengths %d and %d", dA->len, p->len);
    DSC_TRY(out_buf(ctx, dtype_of<T>::value, dA->len, outh, &y));
    size_t n = dA->len;
    switch (kind) {
        case DSC_FUSED_TANH_BWD: return launch_stream<T, FusedBwdF<DSC_FUSED_TANH_BWD, T>, 2>(ctx, FusedBwdF<DSC_FUSED_TANH_BWD, T>(), dA->ptr<T>(), p->ptr<T>(), y->ptr<T>(), n);
        case DSC_FUSED_SIGMOID_BWD: return launch_stream<T, FusedBwdF<DSC_FUSED_SIGMOID_BWD, T>, 2>(ctx, FusedBwdF<DSC_FUSED_SIGMOID_BWD, T>(), dA->ptr<T>(), p->ptr<T>(), y->ptr<T>(), n);
        case DSC_FUSED_RELU_BWD: return launch_stream<T, FusedBwdF<DSC_FUSED_RELU_BWD, T>, 2>(ctx, FusedBwdF<DSC_FUSED_RELU_BWD, T>(), dA->ptr<T>(), p->ptr<T>(), y->ptr<T>(), n);
    }
    return fail(DSC_E_UNSUPPORTED_OP, "fused_bwd: unknown kind %d", kind);
}

// z = A + rows(v) [; h = f_op(z)] — one pass over A instead of broadcast + add (+ map).
// Same roundings as the unfused sequence: z is the fp add, h applies the op to the ROUNDED z.
template <typename T, int OP, bool ACT, bool VEC>
__global__ void __launch_bounds__(kThreads) bias_rows_act_kernel(const T* __restrict__ A, const T* __restrict__ v, T* __restrict__ z,
                                                                  T* __restrict__ h, size_t total, int n) {
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
            V zz, hh;
#pragma unroll
            for (int l = 0; l < W; ++l) {
                lane(zz, l) = lane(a, l) + lane(b, l);
                if constexpr (ACT) lane(hh, l) = unary_op<OP>(lane(zz, l));
            }
            stv<T>(z + i * W, zz);
            if constexpr (ACT) stv<T>(h + i * W, hh);
        } else {
            const T zz = A[i] + v[i % (size_t)n];
            z[i] = zz;
            if constexpr (ACT) h[i] = unary_op<OP>(zz);
        }
    }
}

template <typename T, int OP, bool ACT>
static int launch_bias_rows(Ctx* ctx, const T* A, const T* v, T* z, T* h, int m, int n) {
    const size_t total = (size_t)m * n;
    const bool vec = (n % Vec<T>::N == 0) && aligned16(A) && aligned16(v) && aligned16(z) && (!ACT || aligned16(h));
    const size_t items = vec ? total / Vec<T>::N : total;
    const unsigned grid = (unsigned)((items + (size_t)kThreads * kUnroll - 1) / ((size_t)kThreads * kUnroll));
    if (vec) bias_rows_act_kernel<T, OP, ACT, true><<<grid, kThreads, 0, ctx->stream>>>(A, v, z, h, total, n);
    else bias_rows_act_kernel<T, OP, ACT, false><<<grid, kThreads, 0, ctx->stream>>>(A, v, z, h, total, n);
    DSC_LAUNCH_CHECK(ctx);
    return DSC_OK;
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
    const T* Ap = A->ptr<T>();
    const T* vp = v->ptr<T>();
    T* zp = z->ptr<T>();
    T* hp = h ? h->ptr<T>() : nullptr;
    switch (op) {
#define CASE(OPC) \
    case OPC: return launch_bias_rows<T, OPC, true>(ctx, Ap, vp, zp, hp, m, n);
        CASE(DSC_OP_LOG) CASE(DSC_OP_LOG10) CASE(DSC_OP_EXP) CASE(DSC_OP_SIN) CASE(DSC_OP_COS) CASE(DSC_OP_TAN)
        CASE(DSC_OP_SQRT) CASE(DSC_OP_SINH) CASE(DSC_OP_COSH) CASE(DSC_OP_TANH) CASE(DSC_OP_ASIN) CASE(DSC_OP_ACOS)
        CASE(DSC_OP_ATAN) CASE(DSC_OP_ABS) CASE(DSC_OP_SIGN) CASE(DSC_OP_FLOOR) CASE(DSC_OP_CEILING)
        CASE(DSC_OP_ROUND) CASE(DSC_OP_REL) CASE(DSC_OP_SIGMOID)
#undef CASE
        default: return launch_bias_rows<T, 0, false>(ctx, Ap, vp, zp, nullptr, m, n);
    }
}

}  // namespace dsc

using namespace dsc;

extern "C" {
#define DSC_EW_API(P, T)                                                                                                   \
    int dsc_##P##map(dsc_ctx_t c, int op, dsc_buf_t x, dsc_buf_t* y) { return map_impl<T>(c, op, x, y); }                  \
    int dsc_##P##map_s(dsc_ctx_t c, int op, T s, dsc_buf_t x, dsc_buf_t* y) { return map_s_impl<T>(c, op, s, x, y); }      \
    int dsc_##P##map2(dsc_ctx_t c, int op, dsc_buf_t a, dsc_buf_t b, dsc_buf_t* y) { return map2_impl<T>(c, op, a, b, y); } \
    int dsc_##P##scalar_op(dsc_ctx_t c, int kind, T s, dsc_buf_t x, dsc_buf_t* y) { return scalar_op_impl<T>(c, kind, s, x, y); } \
    int dsc_##P##add_inplace(dsc_ctx_t c, dsc_buf_t d, dsc_buf_t s) { return add_inplace_impl<T>(c, d, s); }               \
    int dsc_##P##add_inplace_range(dsc_ctx_t c, dsc_buf_t s, int32_t so, dsc_buf_t d, int32_t dof, int32_t n) {            \
        return add_inplace_range_impl<T>(c, s, so, d, dof, n);                                                             \
    }                                                                                                                      \
    int dsc_##P##fused_bwd(dsc_ctx_t c, int kind, dsc_buf_t dA, dsc_buf_t p, dsc_buf_t* out) { return fused_bwd_impl<T>(c, kind, dA, p, out); } \
    int dsc_##P##bias_rows_act(dsc_ctx_t c, int op, int32_t m, int32_t n, dsc_buf_t A, dsc_buf_t v, dsc_buf_t* z, dsc_buf_t* h) { \
        return bias_rows_act_impl<T>(c, op, m, n, A, v, z, h);                                                             \
    }
DSC_EW_API(s, float)
DSC_EW_API(d, double)
}
