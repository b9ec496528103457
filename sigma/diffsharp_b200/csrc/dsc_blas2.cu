// K3: GEMV / transposed GEMV / outer product (Mul_M_V, Mul_M_V_Add_V, Mul_V_M, Mul_Out_V_V —
// reference Backend.fs:98-101,111).  All HBM-bound on the matrix; the vector stays in L1/L2.
#include "dsc_internal.cuh"

namespace dsc {

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// out[i] = sum_j A[i,j] x[j] (+ y[i]).  One warp per row, lanes stride the row with 128-bit loads,
// fixed shuffle tree => deterministic.
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) gemv_n_kernel(const T* __restrict__ A, const T* __restrict__ x,
                                                      const T* __restrict__ y, T* __restrict__ out, int m, int n) {
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < m; row += gridDim.x * warps_per_block) {
        const T* a = A + (size_t)row * n;
        T acc0 = T(0), acc1 = T(0);
        if constexpr (VEC) {
            constexpr int W = sizeof(T) == 4 ? 4 : 2;
            using V = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
            const int nvec = n / W;
            int j = lane;
            for (; j + 32 < nvec; j += 64) {
                V a0 = reinterpret_cast<const V*>(a)[j], a1 = reinterpret_cast<const V*>(a)[j + 32];
                V x0 = reinterpret_cast<const V*>(x)[j], x1 = reinterpret_cast<const V*>(x)[j + 32];
                const T* pa0 = reinterpret_cast<const T*>(&a0); const T* px0 = reinterpret_cast<const T*>(&x0);
                const T* pa1 = reinterpret_cast<const T*>(&a1); const T* px1 = reinterpret_cast<const T*>(&x1);
#pragma unroll
                for (int l = 0; l < W; ++l) { acc0 += pa0[l] * px0[l]; acc1 += pa1[l] * px1[l]; }
            }
            for (; j < nvec; j += 32) {
                V a0 = reinterpret_cast<const V*>(a)[j];
                V x0 = reinterpret_cast<const V*>(x)[j];
                const T* pa0 = reinterpret_cast<const T*>(&a0); const T* px0 = reinterpret_cast<const T*>(&x0);
#pragma unroll
                for (int l = 0; l < W; ++l) acc0 += pa0[l] * px0[l];
            }
            for (int jj = nvec * W + lane; jj < n; jj += 32) acc1 += a[jj] * x[jj];
        } else {
            for (int j = lane; j < n; j += 32) acc0 += a[j] * x[j];
        }
        T acc = acc0 + acc1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) out[row] = y ? acc + y[row] : acc;
    }
}

// out[j] = sum_i x[i] A[i,j]: same two-stage column reduction as colsum (dsc_layout.cu) with a
// per-row weight.
template <typename T, int W>
__global__ void __launch_bounds__(256) gemv_t_stage1(const T* __restrict__ A, const T* __restrict__ x, T* __restrict__ partial,
                                                      int m, int n, int rows_per_chunk) {
    __shared__ T sm[8][32 * W + 1];
    const int col0 = (blockIdx.x * 32 + threadIdx.x) * W;
    const int r_begin = blockIdx.y * rows_per_chunk;
    const int r_end = min(m, r_begin + rows_per_chunk);
    T acc[W];
#pragma unroll
    for (int w = 0; w < W; ++w) acc[w] = T(0);
    if (col0 < n) {
        for (int r = r_begin + threadIdx.y; r < r_end; r += 8) {
            const T* row = A + (size_t)r * n + col0;
            const T xr = x[r];
            if constexpr (W == 4) {
                float4 v = *reinterpret_cast<const float4*>(row);
                acc[0] += xr * v.x; acc[1] += xr * v.y; acc[2] += xr * v.z; acc[3] += xr * v.w;
            } else if constexpr (W == 2) {
                double2 v = *reinterpret_cast<const double2*>(row);
                acc[0] += xr * v.x; acc[1] += xr * v.y;
            } else {
                acc[0] += xr * row[0];
            }
        }
    }
#pragma unroll
    for (int w = 0; w < W; ++w) sm[threadIdx.y][threadIdx.x * W + w] = acc[w];
    __syncthreads();
    if (threadIdx.y == 0 && col0 < n) {
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const int cx = threadIdx.x * W + w;
            T s = ((sm[0][cx] + sm[1][cx]) + (sm[2][cx] + sm[3][cx])) + ((sm[4][cx] + sm[5][cx]) + (sm[6][cx] + sm[7][cx]));
            partial[(size_t)blockIdx.y * n + col0 + w] = s;
        }
    }
}
template <typename T>
__global__ void gemv_t_stage2(const T* __restrict__ partial, const T* __restrict__ y, T* __restrict__ out, int n, int chunks) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    T s = T(0);
    int c = 0;
    for (; c + 8 <= chunks; c += 8) {  // 8 independent loads in flight, adds in slab order
        T v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = partial[(size_t)(c + u) * n + j];
#pragma unroll
        for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; c < chunks; ++c) s += partial[(size_t)c * n + j];
    out[j] = y ? s + y[j] : s;
}

template <typename T>
__global__ void ger_kernel(const T* __restrict__ x, const T* __restrict__ y, T* __restrict__ out, int m, int n) {
    size_t total = (size_t)m * n;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) out[i] = x[i / n] * y[i % n];
}

template <typename T>
static int gemv_impl(dsc_ctx_t c, int trans, int32_t m, int32_t n, dsc_buf_t Ah, dsc_buf_t xh, dsc_buf_t yh, dsc_buf_t* outh) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *A, *x, *y = nullptr, *o;
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, Ah, &A, "gemv: A"));
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, xh, &x, "gemv: x"));
    if (yh) DSC_TRY(in_buf(ctx, dtype_of<T>::value, yh, &y, "gemv: y"));
    if (m < 0 || n < 0 || (int64_t)m * n > A->len) return fail(DSC_E_SHAPE, "gemv: %d x %d exceeds buffer length %d", m, n, A->len);
    const int in_len = trans ? m : n, out_len = trans ? n : m;
    if (x->len != in_len)
        return fail(DSC_E_SHAPE, trans ? "Mul_V_M: vector length %d, matrix has %d rows" : "Mul_M_V: vector length %d, matrix has %d columns", x->len, in_len);
    if (y && y->len != out_len) return fail(DSC_E_SHAPE, "Mul_M_V_Add_V: addend length %d, expected %d", y->len, out_len);
    DSC_TRY(out_buf(ctx, dtype_of<T>::value, out_len, outh, &o));
    if (out_len == 0) return DSC_OK;
    if (in_len == 0) {
        if (y) DSC_CUDA_CHECK(cudaMemcpyAsync(o->ptr<T>(), y->ptr<T>(), (size_t)out_len * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
        else DSC_CUDA_CHECK(cudaMemsetAsync(o->ptr<T>(), 0, (size_t)out_len * sizeof(T), ctx->stream));
        return DSC_OK;
    }
    constexpr int W = sizeof(T) == 4 ? 4 : 2;
    const T* yp = y ? y->ptr<T>() : nullptr;
    if (!trans) {
        bool vec = (n % W == 0) && al16(A->ptr<T>()) && al16(x->ptr<T>());
        int grid = grid_for(m, 8, ctx->sm_count * 8);
        if (vec) gemv_n_kernel<T, true><<<grid, 256, 0, ctx->stream>>>(A->ptr<T>(), x->ptr<T>(), yp, o->ptr<T>(), m, n);
        else gemv_n_kernel<T, false><<<grid, 256, 0, ctx->stream>>>(A->ptr<T>(), x->ptr<T>(), yp, o->ptr<T>(), m, n);
        DSC_LAUNCH_CHECK(ctx);
        return DSC_OK;
    }
    bool vec = (n % W == 0) && al16(A->ptr<T>());
    int cols_per_block = vec ? 32 * W : 32;
    int gx = (n + cols_per_block - 1) / cols_per_block;
    int chunks = (ctx->sm_count * 4 + gx - 1) / gx;
    int max_chunks = (m + 63) / 64;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    if (chunks > 256) chunks = 256;
    int rows_per_chunk = (m + chunks - 1) / chunks;
    chunks = (m + rows_per_chunk - 1) / rows_per_chunk;
    DSC_TRY(ensure_red_scratch(ctx, 256 + (size_t)chunks * n * sizeof(T)));
    T* partial = reinterpret_cast<T*>(reinterpret_cast<char*>(ctx->red_scratch) + 256);
    dim3 block(32, 8), grid(gx, chunks);
    if (vec) gemv_t_stage1<T, W><<<grid, block, 0, ctx->stream>>>(A->ptr<T>(), x->ptr<T>(), partial, m, n, rows_per_chunk);
    else gemv_t_stage1<T, 1><<<grid, block, 0, ctx->stream>>>(A->ptr<T>(), x->ptr<T>(), partial, m, n, rows_per_chunk);
    DSC_LAUNCH_CHECK(ctx);
    gemv_t_stage2<T><<<(n + 63) / 64, 64, 0, ctx->stream>>>(partial, yp, o->ptr<T>(), n, chunks);
    DSC_LAUNCH_CHECK(ctx);
    return DSC_OK;
}

template <typename T>
static int ger_impl(dsc_ctx_t c, dsc_buf_t xh, dsc_buf_t yh, dsc_buf_t* outh) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *x, *y, *o;
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, xh, &x, "Mul_Out_V_V: x"));
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, yh, &y, "Mul_Out_V_V: y"));
    int64_t total = (int64_t)x->len * y->len;
    if (total > 0x7fffffff) return fail(DSC_E_SHAPE, "Mul_Out_V_V: result too large for an int32 length");
    DSC_TRY(out_buf(ctx, dtype_of<T>::value, (int32_t)total, outh, &o));
    if (total == 0) return DSC_OK;
    int grid = grid_for((size_t)total, 256 * 4, ctx->sm_count * 8);
    ger_kernel<T><<<grid, 256, 0, ctx->stream>>>(x->ptr<T>(), y->ptr<T>(), o->ptr<T>(), x->len, y->len);
    DSC_LAUNCH_CHECK(ctx);
    return DSC_OK;
}

}  // namespace dsc

using namespace dsc;

extern "C" {
#define DSC_BLAS2_API(P, T)                                                                                          \
    int dsc_##P##gemv(dsc_ctx_t c, int trans, int32_t m, int32_t n, dsc_buf_t A, dsc_buf_t x, dsc_buf_t y, dsc_buf_t* out) { \
        DSC_NVTX(c, "Mul_M_V/Mul_V_M");                                                                               \
        return gemv_impl<T>(c, trans, m, n, A, x, y, out);                                                            \
    }                                                                                                                 \
    int dsc_##P##ger(dsc_ctx_t c, dsc_buf_t x, dsc_buf_t y, dsc_buf_t* out) { DSC_NVTX(c, "Mul_Out_V_V"); return ger_impl<T>(c, x, y, out); }
DSC_BLAS2_API(s, float)
DSC_BLAS2_API(d, double)
}
