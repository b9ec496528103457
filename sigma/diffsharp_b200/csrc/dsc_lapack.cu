// K12: dense LU-based solve / inverse / determinant (Solve_M_V, SolveSymmetric_M_V, Inverse_M,
// Det_M — reference Backend.fs:102-103,125-126; callers AD.Float32.fs:2816,2832,2905,2917).
// Latency-bound small-n helpers, not on any benchmark path: one CTA factors the matrix in place
// (partial pivoting, row-major), one CTA per right-hand side substitutes.  A zero pivot reports
// DSC_E_SINGULAR, which the managed side turns into `None`.
// SolveSymmetric_M_V uses the same LU (the reference's LAPACK ?sysv and ?gesv agree to rounding).
#include "dsc_internal.cuh"

namespace dsc {

constexpr int kLuThreads = 1024;

template <typename T>
__global__ void __launch_bounds__(kLuThreads) lu_factor_kernel(T* __restrict__ A, int n, int* __restrict__ piv, int* info, T* det) {
    __shared__ double s_val[kLuThreads / 32];
    __shared__ int s_idx[kLuThreads / 32];
    __shared__ int s_p;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    T d = T(1);
    for (int j = 0; j < n; ++j) {
        // pivot search: first row with the largest |A[i,j]|, i >= j (LAPACK i?amax convention)
        double bv = -1.0;
        int bi = 0x7fffffff;
        for (int i = j + tid; i < n; i += kLuThreads) {
            double v = fabs((double)A[(size_t)i * n + j]);
            if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { s_val[warp] = bv; s_idx[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < kLuThreads / 32; ++w)
                if (s_val[w] > bv || (s_val[w] == bv && s_idx[w] < bi)) { bv = s_val[w]; bi = s_idx[w]; }
            s_p = bi;
            piv[j] = bi;
            if (!(bv > 0.0) && *info == 0) *info = j + 1;  // zero (or NaN) pivot
        }
        __syncthreads();
        const int p = s_p;
        if (p != j && p < n) {
            for (int c = tid; c < n; c += kLuThreads) {
                T t = A[(size_t)j * n + c];
                A[(size_t)j * n + c] = A[(size_t)p * n + c];
                A[(size_t)p * n + c] = t;
            }
            d = -d;
        }
        __syncthreads();
        const T pivot = A[(size_t)j * n + j];
        d *= pivot;
        if (pivot != T(0)) {
            for (int i = j + 1 + tid; i < n; i += kLuThreads) A[(size_t)i * n + j] /= pivot;
        }
        __syncthreads();
        // trailing update, coalesced along columns
        const int rem = n - j - 1;
        for (long e = tid; e < (long)rem * rem; e += kLuThreads) {
            int i = j + 1 + (int)(e / rem), c = j + 1 + (int)(e % rem);
            A[(size_t)i * n + c] -= A[(size_t)i * n + j] * A[(size_t)j * n + c];
        }
        __syncthreads();
    }
    if (tid == 0 && det) *det = d;
}

// Solve LU x = P b for `nrhs` right-hand sides stored as the columns of X (row-major n x nrhs,
// ldx = nrhs).  One CTA per right-hand side.
template <typename T>
__global__ void __launch_bounds__(256) lu_solve_kernel(const T* __restrict__ LU, const int* __restrict__ piv, T* __restrict__ X, int n, int nrhs) {
    const int col = blockIdx.x;
    if (col >= nrhs) return;
    T* x = X + col;
    const size_t ld = nrhs;
    if (threadIdx.x == 0)
        for (int j = 0; j < n; ++j) {
            int p = piv[j];
            if (p != j) { T t = x[(size_t)j * ld]; x[(size_t)j * ld] = x[(size_t)p * ld]; x[(size_t)p * ld] = t; }
        }
    __syncthreads();
    for (int j = 0; j < n; ++j) {  // L y = Pb (unit diagonal)
        const T xj = x[(size_t)j * ld];
        for (int i = j + 1 + threadIdx.x; i < n; i += blockDim.x) x[(size_t)i * ld] -= LU[(size_t)i * n + j] * xj;
        __syncthreads();
    }
    for (int j = n - 1; j >= 0; --j) {  // U x = y
        if (threadIdx.x == 0) x[(size_t)j * ld] /= LU[(size_t)j * n + j];
        __syncthreads();
        const T xj = x[(size_t)j * ld];
        for (int i = threadIdx.x; i < j; i += blockDim.x) x[(size_t)i * ld] -= LU[(size_t)i * n + j] * xj;
        __syncthreads();
    }
}

template <typename T>
__global__ void eye_kernel(T* __restrict__ X, int n) {
    size_t total = (size_t)n * n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        X[i] = (i / n == i % n) ? T(1) : T(0);
}

struct LuScratch {
    void* lu = nullptr;
    int* piv = nullptr;
    int* info = nullptr;
    void* det = nullptr;
};

template <typename T>
static int lu_factor(Ctx* ctx, Buf* A, int n, LuScratch& s, Block** lu_block, Block** aux_block) {
    if (ctx->capturing) return fail(DSC_E_INVALID, "LAPACK-class ops cannot be captured in a graph (they return a status)");
    DSC_TRY(pool_alloc(ctx, (size_t)n * n * sizeof(T), lu_block));
    int st = pool_alloc(ctx, (size_t)(n + 2) * sizeof(int) + 16, aux_block);
    if (st != DSC_OK) { block_release(*lu_block); return st; }
    s.lu = (*lu_block)->ptr;
    s.info = reinterpret_cast<int*>((*aux_block)->ptr);
    s.det = reinterpret_cast<char*>((*aux_block)->ptr) + 8;
    s.piv = reinterpret_cast<int*>(reinterpret_cast<char*>((*aux_block)->ptr) + 16);
    cudaMemcpyAsync(s.lu, A->ptr<T>(), (size_t)n * n * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream);
    cudaMemsetAsync(s.info, 0, 16, ctx->stream);
    lu_factor_kernel<T><<<1, kLuThreads, 0, ctx->stream>>>(reinterpret_cast<T*>(s.lu), n, s.piv, s.info, reinterpret_cast<T*>(s.det));
    DSC_LAUNCH_CHECK(ctx);
    return DSC_OK;
}

static int read_info(Ctx* ctx, const int* dinfo, int* info) {
    DSC_CUDA_CHECK(cudaMemcpyAsync(ctx->red_host, dinfo, 16, cudaMemcpyDeviceToHost, ctx->stream));
    DSC_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    *info = *reinterpret_cast<int*>(ctx->red_host);
    return DSC_OK;
}

template <typename T>
static int gesv_impl(dsc_ctx_t c, int32_t n, dsc_buf_t Ah, dsc_buf_t bh, dsc_buf_t* xh, const char* name) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *A, *b, *x;
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, Ah, &A, name));
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, bh, &b, name));
    if (n <= 0 || (int64_t)n * n > A->len) return fail(DSC_E_SHAPE, "%s: matrix is not %d x %d", name, n, n);
    if (b->len != n) return fail(DSC_E_SHAPE, "%s: right-hand side has length %d, expected %d", name, b->len, n);
    LuScratch s;
    Block *lub = nullptr, *auxb = nullptr;
    DSC_TRY(lu_factor<T>(ctx, A, n, s, &lub, &auxb));
    int info = 0;
    int st = read_info(ctx, s.info, &info);
    if (st == DSC_OK && info != 0) st = fail(DSC_E_SINGULAR, "%s: matrix is singular (zero pivot at column %d)", name, info - 1);
    if (st == DSC_OK) st = out_buf(ctx, dtype_of<T>::value, n, xh, &x);
    if (st == DSC_OK) {
        cudaMemcpyAsync(x->ptr<T>(), b->ptr<T>(), (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream);
        lu_solve_kernel<T><<<1, 256, 0, ctx->stream>>>(reinterpret_cast<const T*>(s.lu), s.piv, x->ptr<T>(), n, 1);
        ctx->stats.kernel_launches++;
    }
    block_release(lub);
    block_release(auxb);
    return st;
}

template <typename T>
static int getri_impl(dsc_ctx_t c, int32_t n, dsc_buf_t Ah, dsc_buf_t* outh) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *A, *o;
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, Ah, &A, "Inverse_M"));
    if (n <= 0 || (int64_t)n * n > A->len) return fail(DSC_E_SHAPE, "Inverse_M: matrix is not %d x %d", n, n);
    LuScratch s;
    Block *lub = nullptr, *auxb = nullptr;
    DSC_TRY(lu_factor<T>(ctx, A, n, s, &lub, &auxb));
    int info = 0;
    int st = read_info(ctx, s.info, &info);
    if (st == DSC_OK && info != 0) st = fail(DSC_E_SINGULAR, "Inverse_M: matrix is singular (zero pivot at column %d)", info - 1);
    if (st == DSC_OK) st = out_buf(ctx, dtype_of<T>::value, n * n, outh, &o);
    if (st == DSC_OK) {
        eye_kernel<T><<<grid_for((size_t)n * n, 1024, ctx->sm_count * 4), 256, 0, ctx->stream>>>(o->ptr<T>(), n);
        lu_solve_kernel<T><<<n, 256, 0, ctx->stream>>>(reinterpret_cast<const T*>(s.lu), s.piv, o->ptr<T>(), n, n);
        ctx->stats.kernel_launches += 2;
    }
    block_release(lub);
    block_release(auxb);
    return st;
}

template <typename T>
static int det_impl(dsc_ctx_t c, int32_t n, dsc_buf_t Ah, T* out) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    if (!out) return fail(DSC_E_INVALID, "Det_M: null result pointer");
    Ctx* ctx = as_ctx(c);
    Buf* A;
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, Ah, &A, "Det_M"));
    if (n <= 0 || (int64_t)n * n > A->len) return fail(DSC_E_SHAPE, "Det_M: matrix is not %d x %d", n, n);
    LuScratch s;
    Block *lub = nullptr, *auxb = nullptr;
    DSC_TRY(lu_factor<T>(ctx, A, n, s, &lub, &auxb));
    int info = 0;
    int st = read_info(ctx, s.info, &info);
    if (st == DSC_OK) {
        // red_host holds {info, pad, det}
        *out = *reinterpret_cast<T*>(reinterpret_cast<char*>(ctx->red_host) + 8);
        if (info != 0) *out = T(0);  // a singular matrix has determinant 0; Det_M itself never fails
    }
    block_release(lub);
    block_release(auxb);
    return st;
}

}  // namespace dsc

using namespace dsc;

extern "C" {
#define DSC_LAPACK_API(P, T)                                                                                              \
    int dsc_##P##gesv(dsc_ctx_t c, int32_t n, dsc_buf_t A, dsc_buf_t b, dsc_buf_t* x) { DSC_NVTX(c, "Solve_M_V"); return gesv_impl<T>(c, n, A, b, x, "Solve_M_V"); } \
    int dsc_##P##sysv(dsc_ctx_t c, int32_t n, dsc_buf_t A, dsc_buf_t b, dsc_buf_t* x) { DSC_NVTX(c, "SolveSymmetric_M_V"); return gesv_impl<T>(c, n, A, b, x, "SolveSymmetric_M_V"); } \
    int dsc_##P##getri(dsc_ctx_t c, int32_t n, dsc_buf_t A, dsc_buf_t* out) { DSC_NVTX(c, "Inverse_M"); return getri_impl<T>(c, n, A, out); }         \
    int dsc_##P##det(dsc_ctx_t c, int32_t n, dsc_buf_t A, T* out) { DSC_NVTX(c, "Det_M"); return det_impl<T>(c, n, A, out); }
DSC_LAPACK_API(s, float)
DSC_LAPACK_API(d, double)
}
