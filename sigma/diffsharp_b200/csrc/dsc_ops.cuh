// Per-element semantics of the MapOp op-codes and of the adjoint chains, shared by the streaming kernels
// (dsc_elementwise.cu) and by the epilogue of the tensor-core GEMM (dsc_gemm_tf32x3.cu): ONE definition, so that an
// operation fused into a producing kernel performs exactly the roundings of the stand-alone kernel.
// The closure text at each reference call site is the spec of its op-code (AD.Float32.fs:1388-1594, 2631-2934).
#pragma once

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


// Adjoint chains (dsc_fused_kind), each with the reference's own sequence of roundings:
//  tanh   : secha = 1/cosh(aP); (dA*secha)*secha                      (AD.Float32.fs:4238-4240)
//  sigmoid: (dA*dP)*(1-dP)                                            (AD.Float32.fs:4281)
//  relu   : dA*((1/2)*(sign(aP)+1))  -> derivative at 0 is 1/2        (AD.Float32.fs:4280)
template <int KIND, typename T>
__device__ __forceinline__ T chain_op(T dA, T p) {
    if constexpr (KIND == DSC_FUSED_TANH_BWD) {
        const T secha = T(1) / m_cosh(p);
        return (dA * secha) * secha;
    } else if constexpr (KIND == DSC_FUSED_SIGMOID_BWD) {
        return (dA * p) * (T(1) - p);
    } else {
        return dA * (T(0.5) * (sgn(p) + T(1)));
    }
}

}  // namespace dsc
