// Skinny GEMMs: C(m x n) = op(A) op(B) [+ bias] with n <= 16 — the output layer of the MLP (n = 10):
//   forward      H2 (B x 1024) * W3 (1024 x 10)             A has K contiguous   -> rows kernel
//   weight adj.  H2^T (1024 x B) * dZ (B x 10), K = batch   A has M contiguous   -> cols kernel (split-K)
// These are HBM-bound on A (read once, 128-bit loads); B is tiny and staged in shared memory.
// Tensor cores cannot help (the N extent does not fill a tile and the pre-split would triple the
// traffic), so they run on CUDA cores with a fixed reduction order (deterministic).
#include "dsc_internal.cuh"

namespace dsc {
namespace skinny {

constexpr int NMAX = 16;

template <typename T>
__device__ __forceinline__ T ldB(const T* __restrict__ B, int tb, int k, int n, int kk, int j) {
    return tb ? B[(size_t)j * k + kk] : B[(size_t)kk * n + j];
}

// Stage B^T (n <= NMAX rows of kc k-values) into shared memory as Bs[j][kk], zero padded, with
// coalesced global reads in either stored layout.
template <typename T, int KC>
__device__ __forceinline__ void stage_bt(T* Bs, const T* __restrict__ B, int tb, int k, int n, int k0, int kc) {
    for (int idx = threadIdx.x; idx < NMAX * KC; idx += blockDim.x) Bs[idx] = T(0);
    __syncthreads();
    if (tb) {  // stored n x k: row j is contiguous in kk
        for (int idx = threadIdx.x; idx < n * kc; idx += blockDim.x) {
            const int j = idx / kc, kk = idx % kc;
            Bs[j * KC + kk] = B[(size_t)j * k + k0 + kk];
        }
    } else {   // stored k x n: rows k0..k0+kc are one contiguous span of kc*n elements
        const T* src = B + (size_t)k0 * n;
        for (int idx = threadIdx.x; idx < kc * n; idx += blockDim.x) {
            const int kk = idx / n, j = idx % n;
            Bs[j * KC + kk] = src[idx];
        }
    }
    __syncthreads();
}

// ---- A has K contiguous (ta = 0): one warp per RPW output rows, lanes stride K ----------------
// Persistent CTAs (256 threads = 8 warps) loop over row tiles of 8*RPW rows; when K fits one chunk
// (K <= KC) B^T is staged once per CTA, otherwise once per (tile, chunk).
template <typename T, int KC, int RPW>
__global__ void __launch_bounds__(256) skinny_rows_kernel(const T* __restrict__ A, const T* __restrict__ B, const T* __restrict__ bias,
                                                           T* __restrict__ C, int m, int n, int k, int tb) {
    extern __shared__ unsigned char smem_raw[];
    T* Bs = reinterpret_cast<T*>(smem_raw);  // [NMAX][KC]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles = (m + 8 * RPW - 1) / (8 * RPW);
    const bool single = k <= KC;
    if (single) stage_bt<T, KC>(Bs, B, tb, k, n, 0, k);
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int row0 = (tile * 8 + warp) * RPW;
        T acc[RPW][NMAX];
#pragma unroll
        for (int r = 0; r < RPW; ++r)
#pragma unroll
            for (int j = 0; j < NMAX; ++j) acc[r][j] = T(0);
        for (int k0 = 0; k0 < k; k0 += KC) {
            const int kc = min(KC, k - k0);
            if (!single) stage_bt<T, KC>(Bs, B, tb, k, n, k0, kc);
            // kk outer, rows inner: every B value read from shared memory feeds RPW rows
            const T* arow[RPW];
#pragma unroll
            for (int r = 0; r < RPW; ++r) arow[r] = A + (size_t)min(row0 + r, m - 1) * k + k0;  // clamped; discarded below
#pragma unroll 2
            for (int kk = lane; kk < kc; kk += 32) {
                T av[RPW];
#pragma unroll
                for (int r = 0; r < RPW; ++r) av[r] = arow[r][kk];
#pragma unroll
                for (int j = 0; j < NMAX; ++j) {
                    const T bv = Bs[j * KC + kk];
#pragma unroll
                    for (int r = 0; r < RPW; ++r) acc[r][j] = fma(av[r], bv, acc[r][j]);
                }
            }
            if (!single) __syncthreads();
        }
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int row = row0 + r;
            const T bv = (bias && row < m) ? bias[row] : T(0);
#pragma unroll
            for (int j = 0; j < NMAX; ++j) {
                T v = acc[r][j];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == j && j < n && row < m) C[(size_t)row * n + j] = v + bv;
            }
        }
    }
}

// ---- A has M contiguous (ta = 1, stored k x m): thread per output row, split over K ---------
// grid (ceil(m/256), chunks); block 256 threads; chunk = KCH k-rows; partial[chunk][m][n]
template <typename T, int KCH>
__global__ void __launch_bounds__(256) skinny_cols_stage1(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ partial,
                                                           int m, int n, int k, int tb) {
    __shared__ __align__(16) T Bs[KCH][NMAX];
    const int col = blockIdx.x * 256 + threadIdx.x;  // output row index i (A column)
    const int k0 = blockIdx.y * KCH;
    const int kc = min(KCH, k - k0);
    for (int idx = threadIdx.x; idx < KCH * NMAX; idx += 256) {
        const int kk = idx / NMAX, j = idx % NMAX;
        Bs[kk][j] = (kk < kc && j < n) ? ldB(B, tb, k, n, k0 + kk, j) : T(0);
    }
    __syncthreads();
    if (col >= m) return;
    T acc[NMAX];
#pragma unroll
    for (int j = 0; j < NMAX; ++j) acc[j] = T(0);
    const T* a = A + (size_t)k0 * m + col;
    constexpr int VW = 16 / (int)sizeof(T);  // elements per 128-bit broadcast load of a B row
#pragma unroll 8
    for (int kk = 0; kk < kc; ++kk) {
        const T av = a[(size_t)kk * m];
#pragma unroll
        for (int j = 0; j < NMAX; j += VW) {
            const int4 raw = *reinterpret_cast<const int4*>(&Bs[kk][j]);  // same address in every lane: broadcast
            const T* bv = reinterpret_cast<const T*>(&raw);
#pragma unroll
            for (int v = 0; v < VW; ++v) acc[j + v] = fma(av, bv[v], acc[j + v]);
        }
    }
    T* p = partial + ((size_t)blockIdx.y * m + col) * n;
#pragma unroll
    for (int j = 0; j < NMAX; ++j)
        if (j < n) p[j] = acc[j];
}

template <typename T>
__global__ void skinny_cols_stage2(const T* __restrict__ partial, const T* __restrict__ bias, T* __restrict__ C, int m, int n, int chunks) {
    const size_t total = (size_t)m * n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        T s = T(0);
        int c = 0;
        for (; c + 8 <= chunks; c += 8) {  // 8 independent loads in flight, adds in chunk order
            T v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = partial[(size_t)(c + u) * total + i];
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        for (; c < chunks; ++c) s += partial[(size_t)c * total + i];
        C[i] = s + (bias ? bias[i / n] : T(0));
    }
}

// ---- tiny K (k <= NMAX): C(m x n) = A(m x k) op(B)(k x n), output-write bound -----------------
// (the input adjoint of the output layer: dZ (B x 10) * W3^T (10 x 1024)).  A thread keeps the k
// B-values of its 4 (2 for double) output columns in registers and walks the rows of its CTA's slab.
template <typename T>
__global__ void __launch_bounds__(256) tiny_k_kernel(const T* __restrict__ A, const T* __restrict__ B, const T* __restrict__ bias,
                                                      T* __restrict__ C, int m, int n, int k, int ta, int tb, int rows_per_cta) {
    constexpr int W = 16 / (int)sizeof(T);
    const int c0 = (blockIdx.x * 256 + threadIdx.x) * W;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(m, r0 + rows_per_cta);
    __shared__ T As[64][NMAX];
    T b[NMAX][W];
#pragma unroll
    for (int kk = 0; kk < NMAX; ++kk)
#pragma unroll
        for (int w = 0; w < W; ++w) b[kk][w] = (kk < k && c0 + w < n) ? (tb ? B[(size_t)(c0 + w) * k + kk] : B[(size_t)kk * n + c0 + w]) : T(0);
    const bool vec = (n % W == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    for (int rb = r0; rb < r1; rb += 64) {
        const int nr = min(64, r1 - rb);
        __syncthreads();
        for (int idx = threadIdx.x; idx < 64 * NMAX; idx += 256) {
            const int r = idx / NMAX, kk = idx % NMAX;
            As[r][kk] = (r < nr && kk < k) ? (ta ? A[(size_t)kk * m + rb + r] : A[(size_t)(rb + r) * k + kk]) : T(0);
        }
        __syncthreads();
        if (c0 >= n) continue;
        for (int r = 0; r < nr; ++r) {
            T acc[W];
#pragma unroll
            for (int w = 0; w < W; ++w) acc[w] = T(0);
#pragma unroll
            for (int kk = 0; kk < NMAX; ++kk) {
                const T av = As[r][kk];  // broadcast
#pragma unroll
                for (int w = 0; w < W; ++w) acc[w] = fma(av, b[kk][w], acc[w]);
            }
            const T bv = bias ? bias[rb + r] : T(0);
            T* dst = C + (size_t)(rb + r) * n + c0;
            if (vec) {
                int4 raw;
                T* o = reinterpret_cast<T*>(&raw);
#pragma unroll
                for (int w = 0; w < W; ++w) o[w] = acc[w] + bv;
                *reinterpret_cast<int4*>(dst) = raw;
            } else {
#pragma unroll
                for (int w = 0; w < W; ++w)
                    if (c0 + w < n) dst[w] = acc[w] + bv;
            }
        }
    }
}

}  // namespace skinny

// returns DSC_OK when handled, 1 when the shape is not a skinny / tiny-K product
template <typename T>
int gemm_skinny(Ctx* ctx, int ta, int tb, int m, int n, int k, const T* A, const T* B, const T* bias, T* C) {
    using namespace skinny;
    if (k <= NMAX && n >= 64 && (long)m * n >= 65536) {
        constexpr int W = 16 / (int)sizeof(T);
        const int gx = (n + 256 * W - 1) / (256 * W);
        int gy = (ctx->sm_count * 4 + gx - 1) / gx;
        int rows_per_cta = (m + gy - 1) / gy;
        rows_per_cta = (rows_per_cta + 63) / 64 * 64;
        gy = (m + rows_per_cta - 1) / rows_per_cta;
        tiny_k_kernel<T><<<dim3(gx, gy), 256, 0, ctx->stream>>>(A, B, bias, C, m, n, k, ta, tb, rows_per_cta);
        DSC_LAUNCH_CHECK(ctx);
        ctx->stats.gemm_simt++;
        return DSC_OK;
    }
    if (n > NMAX || (long)m * k < 65536) return 1;
    if (!ta) {
        constexpr int KC = sizeof(T) == 4 ? 1024 : 512;
        constexpr int RPW = 4;
        const int smem = NMAX * KC * (int)sizeof(T);
        static bool attr_set = false;
        if (!attr_set) {
            DSC_CUDA_CHECK(cudaFuncSetAttribute(skinny_rows_kernel<T, KC, RPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            attr_set = true;
        }
        const int tiles = (m + 8 * RPW - 1) / (8 * RPW);
        const int grid = tiles < ctx->sm_count * 2 ? tiles : ctx->sm_count * 2;
        skinny_rows_kernel<T, KC, RPW><<<grid, 256, smem, ctx->stream>>>(A, B, bias, C, m, n, k, tb);
        DSC_LAUNCH_CHECK(ctx);
    } else {
        constexpr int KCH = 128;
        const int chunks = (k + KCH - 1) / KCH;
        const size_t need = 256 + (size_t)chunks * m * n * sizeof(T);
        DSC_TRY(ensure_red_scratch(ctx, need));
        T* partial = reinterpret_cast<T*>(reinterpret_cast<char*>(ctx->red_scratch) + 256);
        dim3 grid((m + 255) / 256, chunks);
        skinny_cols_stage1<T, KCH><<<grid, 256, 0, ctx->stream>>>(A, B, partial, m, n, k, tb);
        DSC_LAUNCH_CHECK(ctx);
        skinny_cols_stage2<T><<<grid_for((size_t)m * n, 128, ctx->sm_count * 4), 128, 0, ctx->stream>>>(partial, bias, C, m, n, chunks);
        DSC_LAUNCH_CHECK(ctx);
    }
    ctx->stats.gemm_simt++;
    return DSC_OK;
}

template int gemm_skinny<float>(Ctx*, int, int, int, int, int, const float*, const float*, const float*, float*);
template int gemm_skinny<double>(Ctx*, int, int, int, int, int, const double*, const double*, const double*, double*);

}  // namespace dsc
