// K1/K2 dispatch + the CUDA-core GEMM used for shapes the tensor-core kernels do not take
// (tiny / skinny / unaligned operands) — Mul_M_M and Mul_M_M_Add_V_MCols, reference
// Backend.fs:119,121; call site AD.Float32.fs:2382; adjoint call sites :4127-4143.
//
//   f32: tcgen05/TMEM/TMA 3xTF32 kernel (dsc_gemm_tf32x3.cu) when eligible, else this kernel.
//   f64: DMMA kernel (dsc_gemm_dmma.cu) when eligible, else this kernel.
// There is no host fallback: every path is a device kernel of this library.
#include "dsc_internal.cuh"

namespace dsc {

// C(m x n) = op(A) op(B) (+ bias[i]).  Generic strides: A(i,kk) = A[i*sai + kk*sak], B(kk,j) = B[kk*sbk + j*sbj].
// 64x64x16 tiles, 256 threads, 4x4 register micro-tile, FMA accumulation in T (exact IEEE fma chain
// over k in ascending order for every output element => deterministic).
template <typename T, bool A_K_CONTIG, bool B_N_CONTIG>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const T* __restrict__ A, const T* __restrict__ B, const T* __restrict__ bias,
                                                        T* __restrict__ C, int m, int n, int k, long sai, long sak, long sbk, long sbj) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ T As[BK][BM + 4];
    __shared__ T Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads
    const int row0 = blockIdx.y * BM, col0 = blockIdx.x * BN;
    T acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = T(0);

    for (int k0 = 0; k0 < k; k0 += BK) {
        // A tile: BM x BK elements, 1024 = 256 threads x 4
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            int idx = tid + e * 256;
            int i, kk;
            if (A_K_CONTIG) { kk = idx & (BK - 1); i = idx >> 4; }   // consecutive threads walk k
            else { i = idx & (BM - 1); kk = idx >> 6; }              // consecutive threads walk i
            int gi = row0 + i, gk = k0 + kk;
            As[kk][i] = (gi < m && gk < k) ? A[(size_t)gi * sai + (size_t)gk * sak] : T(0);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            int idx = tid + e * 256;
            int j, kk;
            if (B_N_CONTIG) { j = idx & (BN - 1); kk = idx >> 6; }
            else { kk = idx & (BK - 1); j = idx >> 4; }
            int gj = col0 + j, gk = k0 + kk;
            Bs[kk][j] = (gj < n && gk < k) ? B[(size_t)gk * sbk + (size_t)gj * sbj] : T(0);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            T a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int gi = row0 + ty * 4 + i;
        if (gi >= m) continue;
        T bv = bias ? bias[gi] : T(0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int gj = col0 + tx * 4 + j;
            if (gj < n) C[(size_t)gi * n + gj] = acc[i][j] + bv;
        }
    }
}

template <typename T>
__global__ void bias_rows_kernel(const T* __restrict__ bias, T* __restrict__ C, int m, int n) {
    size_t total = (size_t)m * n;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) C[i] = bias ? bias[i / n] : T(0);
}

template <typename T>
static int gemm_simt(Ctx* ctx, int ta, int tb, int m, int n, int k, const T* A, const T* B, const T* bias, T* C) {
    long sai = ta ? 1 : k, sak = ta ? m : 1;   // stored A is k x m when ta
    long sbk = tb ? 1 : n, sbj = tb ? k : 1;   // stored B is n x k when tb
    dim3 grid((n + 63) / 64, (m + 63) / 64);
    if (!ta && !tb) gemm_simt_kernel<T, true, true><<<grid, 256, 0, ctx->stream>>>(A, B, bias, C, m, n, k, sai, sak, sbk, sbj);
    else if (!ta && tb) gemm_simt_kernel<T, true, false><<<grid, 256, 0, ctx->stream>>>(A, B, bias, C, m, n, k, sai, sak, sbk, sbj);
    else if (ta && !tb) gemm_simt_kernel<T, false, true><<<grid, 256, 0, ctx->stream>>>(A, B, bias, C, m, n, k, sai, sak, sbk, sbj);
    else gemm_simt_kernel<T, false, false><<<grid, 256, 0, ctx->stream>>>(A, B, bias, C, m, n, k, sai, sak, sbk, sbj);
    DSC_LAUNCH_CHECK(ctx);
    ctx->stats.gemm_simt++;
    return DSC_OK;
}

template <>
int gemm_impl<float>(Ctx* ctx, int ta, int tb, int m, int n, int k, Buf* A, Buf* B, Buf* bias, Buf* C) {
    const float* bp = bias ? bias->ptr<float>() : nullptr;
    {
        int s = gemm_skinny<float>(ctx, ta, tb, m, n, k, A->ptr<float>(), B->ptr<float>(), bp, C->ptr<float>());
        if (s <= 0) return s;
    }
    if (!(ctx->flags & DSC_FLAG_GEMM_FP32_SIMT)) {
        int s = gemm_tf32x3_tcgen05(ctx, ta, tb, m, n, k, A->ptr<float>(), B->ptr<float>(), bp, C->ptr<float>());
        if (s <= 0) return s;  // handled (0) or failed (<0); 1 = not eligible
    }
    return gemm_simt<float>(ctx, ta, tb, m, n, k, A->ptr<float>(), B->ptr<float>(), bp, C->ptr<float>());
}

template <>
int gemm_impl<double>(Ctx* ctx, int ta, int tb, int m, int n, int k, Buf* A, Buf* B, Buf* bias, Buf* C) {
    const double* bp = bias ? bias->ptr<double>() : nullptr;
    {
        int s = gemm_skinny<double>(ctx, ta, tb, m, n, k, A->ptr<double>(), B->ptr<double>(), bp, C->ptr<double>());
        if (s <= 0) return s;
    }
    if (!(ctx->flags & DSC_FLAG_GEMM_F64_SIMT)) {
        int s = gemm_f64_dmma(ctx, ta, tb, m, n, k, A->ptr<double>(), B->ptr<double>(), bp, C->ptr<double>());
        if (s <= 0) return s;
    }
    return gemm_simt<double>(ctx, ta, tb, m, n, k, A->ptr<double>(), B->ptr<double>(), bp, C->ptr<double>());
}

template <typename T>
static int gemm_entry(dsc_ctx_t c, int ta, int tb, int32_t m, int32_t n, int32_t k, dsc_buf_t Ah, dsc_buf_t Bh, dsc_buf_t biash, dsc_buf_t* Ch) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *A, *B, *bias = nullptr, *C;
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, Ah, &A, "Mul_M_M: A"));
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, Bh, &B, "Mul_M_M: B"));
    if (biash) DSC_TRY(in_buf(ctx, dtype_of<T>::value, biash, &bias, "Mul_M_M_Add_V_MCols: v"));
    if (m < 0 || n < 0 || k < 0) return fail(DSC_E_SHAPE, "Mul_M_M: negative dimension");
    if ((int64_t)m * k > A->len) return fail(DSC_E_SHAPE, "Mul_M_M: A is %d x %d but its buffer holds %d elements", m, k, A->len);
    if ((int64_t)k * n > B->len) return fail(DSC_E_SHAPE, "Mul_M_M: B is %d x %d but its buffer holds %d elements", k, n, B->len);
    if ((int64_t)m * n > 0x7fffffff) return fail(DSC_E_SHAPE, "Mul_M_M: result too large for an int32 length");
    if (bias && bias->len != m) return fail(DSC_E_SHAPE, "Mul_M_M_Add_V_MCols: vector length %d, result has %d rows", bias->len, m);
    DSC_TRY(out_buf(ctx, dtype_of<T>::value, m * n, Ch, &C));
    if (m * n == 0) return DSC_OK;
    if (C->ptr<T>() == A->ptr<T>() || C->ptr<T>() == B->ptr<T>()) return fail(DSC_E_INVALID, "Mul_M_M: output aliases an input");
    if (k == 0) {
        int grid = grid_for((size_t)m * n, 256 * 4, ctx->sm_count * 8);
        bias_rows_kernel<T><<<grid, 256, 0, ctx->stream>>>(bias ? bias->ptr<T>() : nullptr, C->ptr<T>(), m, n);
        DSC_LAUNCH_CHECK(ctx);
        return DSC_OK;
    }
    return gemm_impl<T>(ctx, ta ? 1 : 0, tb ? 1 : 0, m, n, k, A, B, bias, C);
}

}  // namespace dsc

using namespace dsc;

extern "C" {
int dsc_sgemm(dsc_ctx_t c, int ta, int tb, int32_t m, int32_t n, int32_t k, dsc_buf_t A, dsc_buf_t B, dsc_buf_t bias, dsc_buf_t* C) {
    return gemm_entry<float>(c, ta, tb, m, n, k, A, B, bias, C);
}
int dsc_dgemm(dsc_ctx_t c, int ta, int tb, int32_t m, int32_t n, int32_t k, dsc_buf_t A, dsc_buf_t B, dsc_buf_t bias, dsc_buf_t* C) {
    return gemm_entry<double>(c, ta, tb, m, n, k, A, B, bias, C);
}
}
