// K2: float64 GEMM on the FP64 tensor path (DMMA: mma.sync.m8n8k4.f64) — Mul_M_M for 'T = float
// (Backend.fs:119; AD.Float64.fs:2369) and the tangent GEMMs of the forward-mode Jacobian (config 4).
// tcgen05 has no f64 kind, so this is the warp-level tensor instruction with a classic cp.async
// multistage shared-memory pipeline:
//   CTA tile 128 x 128 x 16, 8 warps (2 x 4), warp tile 64 x 32 = 8 x 4 m8n8k4 tiles,
//   64 fp64 accumulators (128 registers) per thread, 3 stages (120 KB), one CTA per SM.
// Shared-memory rows are padded so that the 64-bit fragment loads are bank-conflict free
// (row stride = 4 mod 16 doubles).  All four transpose combinations read their operand in its stored
// layout (K-contiguous or MN-contiguous) — a lazily transposed operand costs nothing.
// Results are IEEE fp64 FMA chains over k in ascending order: deterministic, <= 1e-12 vs OpenBLAS dgemm.
#include "dsc_internal.cuh"

namespace dsc {
namespace dmma {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int kStages = 3;
constexpr int kThreads = 256;
constexpr int kStrideK = BK + 4;    // tile stored [mn][k]  : 20 doubles per row
constexpr int kStrideMN = BM + 4;   // tile stored [k][mn]  : 132 doubles per row
constexpr int kTileElemsK = BM * kStrideK;    // 2560
constexpr int kTileElemsMN = BK * kStrideMN;  // 2112
constexpr int kTileElems = kTileElemsK > kTileElemsMN ? kTileElemsK : kTileElemsMN;
constexpr int kSmemBytes = kStages * 2 * kTileElems * (int)sizeof(double);  // 122880

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    const int bytes = pred ? 16 : 0;  // src-size 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// Load one operand tile (128 "mn" rows x 16 k) of a stage.
//  K_CONTIG : the stored matrix has k contiguous: elem(mn, k) = src[mn * ld + k]  -> smem [mn][k], stride 20
//  !K_CONTIG: the stored matrix has mn contiguous: elem(mn, k) = src[k * ld + mn] -> smem [k][mn], stride 132
template <bool K_CONTIG>
__device__ __forceinline__ void load_tile(double* smem, const double* __restrict__ src, int ld, int mn0, int k0, int mn_lim, int k_lim, int tid) {
    if (K_CONTIG) {
        // 128 rows x 8 chunks of 2 doubles = 1024 chunks, 4 per thread
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = tid + e * kThreads;
            const int r = idx >> 3, c = (idx & 7) * 2;
            const bool ok = (mn0 + r < mn_lim) && (k0 + c < k_lim);
            const double* g = ok ? src + (size_t)(mn0 + r) * ld + (k0 + c) : src;
            cp_async16(smem + r * kStrideK + c, g, ok);
        }
    } else {
        // 16 k-rows x 64 chunks of 2 doubles = 1024 chunks
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = tid + e * kThreads;
            const int kk = idx >> 6, c = (idx & 63) * 2;
            const bool ok = (k0 + kk < k_lim) && (mn0 + c < mn_lim);
            const double* g = ok ? src + (size_t)(k0 + kk) * ld + (mn0 + c) : src;
            cp_async16(smem + kk * kStrideMN + c, g, ok);
        }
    }
}

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(kThreads, 1)
gemm_f64_dmma_kernel(const double* __restrict__ A, const double* __restrict__ B, const double* __restrict__ bias, double* __restrict__ C,
                     int m, int n, int k, int lda, int ldb, int vec2) {
    extern __shared__ double smem[];
    double* As = smem;
    double* Bs = smem + kStages * kTileElems;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 2, wn = warp & 3;  // 2 x 4 warps
    const int g = lane >> 2, t = lane & 3;    // groupID, thread-in-group
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int KT = (k + BK - 1) / BK;

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // prologue
#pragma unroll
    for (int s = 0; s < kStages - 1; ++s) {
        if (s < KT) {
            load_tile<A_KC>(As + s * kTileElems, A, lda, m0, s * BK, m, k, tid);
            load_tile<B_KC>(Bs + s * kTileElems, B, ldb, n0, s * BK, n, k, tid);
        }
        cp_async_commit();
    }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<kStages - 2>();
        __syncthreads();
        // prefetch stage kt + kStages - 1 (overwrites the stage consumed in iteration kt - 1)
        {
            const int nk = kt + kStages - 1;
            if (nk < KT) {
                const int s = nk % kStages;
                load_tile<A_KC>(As + s * kTileElems, A, lda, m0, nk * BK, m, k, tid);
                load_tile<B_KC>(Bs + s * kTileElems, B, ldb, n0, nk * BK, n, k, tid);
            }
            cp_async_commit();
        }
        const double* as = As + (kt % kStages) * kTileElems;
        const double* bs = Bs + (kt % kStages) * kTileElems;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = wm * 64 + i * 8 + g;
                a[i] = A_KC ? as[r * kStrideK + kk + t] : as[(kk + t) * kStrideMN + r];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = wn * 32 + j * 8 + g;
                b[j] = B_KC ? bs[c * kStrideK + kk + t] : bs[(kk + t) * kStrideMN + c];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: thread holds C[row = g][cols 2t, 2t+1] of every 8x8 tile
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = m0 + wm * 64 + i * 8 + g;
        if (row >= m) continue;
        const double bv = bias ? bias[row] : 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = n0 + wn * 32 + j * 8 + t * 2;
            double* dst = C + (size_t)row * n + col;
            if (vec2 && col + 1 < n) {
                *reinterpret_cast<double2*>(dst) = make_double2(acc[i][j][0] + bv, acc[i][j][1] + bv);
            } else {
                if (col < n) dst[0] = acc[i][j][0] + bv;
                if (col + 1 < n) dst[1] = acc[i][j][1] + bv;
            }
        }
    }
}

template <bool A_KC, bool B_KC>
static int launch(Ctx* ctx, const double* A, const double* B, const double* bias, double* C, int m, int n, int k, int lda, int ldb, int vec2) {
    static bool attr_set = false;
    if (!attr_set) {
        DSC_CUDA_CHECK(cudaFuncSetAttribute(gemm_f64_dmma_kernel<A_KC, B_KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        attr_set = true;
    }
    dim3 grid((n + BN - 1) / BN, (m + BM - 1) / BM);
    gemm_f64_dmma_kernel<A_KC, B_KC><<<grid, kThreads, kSmemBytes, ctx->stream>>>(A, B, bias, C, m, n, k, lda, ldb, vec2);
    DSC_LAUNCH_CHECK(ctx);
    return DSC_OK;
}

}  // namespace dmma

int gemm_f64_dmma(Ctx* ctx, int ta, int tb, int m, int n, int k, const double* A, const double* B, const double* bias, double* C) {
    using namespace dmma;
    // op(A) is m x k: stored m x k (k contiguous, lda = k) or, with ta, k x m (m contiguous, lda = m)
    // op(B) is k x n: stored k x n (n contiguous, ldb = n) or, with tb, n x k (k contiguous, ldb = k)
    const int lda = ta ? m : k, ldb = tb ? k : n;
    // 16-byte cp.async chunks: leading dimensions even and bases 16-byte aligned; tiny products stay on the CUDA-core kernel
    if ((lda & 1) || (ldb & 1) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return 1;
    if ((long)m * n < 64 * 64 || k < 16) return 1;
    const int vec2 = ((n & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    int s;
    if (!ta && !tb) s = launch<true, false>(ctx, A, B, bias, C, m, n, k, lda, ldb, vec2);
    else if (!ta && tb) s = launch<true, true>(ctx, A, B, bias, C, m, n, k, lda, ldb, vec2);
    else if (ta && !tb) s = launch<false, false>(ctx, A, B, bias, C, m, n, k, lda, ldb, vec2);
    else s = launch<false, true>(ctx, A, B, bias, C, m, n, k, lda, ldb, vec2);
    if (s == DSC_OK) ctx->stats.gemm_dmma++;
    return s;
}

}  // namespace dsc
