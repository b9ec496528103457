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

// ================================================================================================
// 128-bit paths (the ones the MLP's output layer takes; the scalar kernels above stay as the generic
// fallback for unaligned windows / K or M not a multiple of the vector width)
// ================================================================================================
template <typename T> struct V16;
template <> struct V16<float> { using type = float4; static constexpr int W = 4; };
template <> struct V16<double> { using type = double2; static constexpr int W = 2; };
__device__ __forceinline__ float vget(const float4& v, int i) { return (&v.x)[i]; }
__device__ __forceinline__ double vget(const double2& v, int i) { return (&v.x)[i]; }

// Sum 16 per-lane values over the 32 lanes of a warp with 16 shuffles (instead of 16 x 5): at every
// step a lane keeps one half of its values and sends the other half to its partner.  On return
// v[0] of lane l is the total of value index 8*b4 + 4*b3 + 2*b2 + b1 (b_i = bit i of l); the
// combine order is fixed, so the result is deterministic.
template <typename T>
__device__ __forceinline__ void warp_sum16(T (&v)[16], int lane) {
#pragma unroll
    for (int s = 8; s >= 1; s >>= 1) {
        const bool up = (lane & (2 * s)) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const T send = up ? v[i] : v[i + s];
            const T keep = up ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2 * s);
        }
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// ---- A has K contiguous (ta = 0), 128-bit loads ------------------------------------------------
// One warp owns RPW output rows; lane l reads the 16-byte vectors q = it*32 + l of each row (8 per
// K chunk, all issued before the first FMA), B^T sits in shared memory as Bs[j][kk] so the same lane
// reads the matching 16 bytes of every B column without bank conflicts.  NJ = n rounded up to 4.
template <typename T, int RPW, int NJ>
__global__ void __launch_bounds__(256, RPW >= 4 ? 1 : 2) skinny_rows_vec_kernel(const T* __restrict__ A, const T* __restrict__ B, const T* __restrict__ bias,
                                                                  const T* __restrict__ bias_col, T* __restrict__ C, int m, int n, int k, int tb) {
    using V = typename V16<T>::type;
    constexpr int W = V16<T>::W;
    // vectors of a row in flight per lane: 8 (two rows per warp), 4 with four rows per warp — the same 16 loads in flight, but
    // every B vector read from shared memory then feeds four rows: the two-row kernel spends 12 LDS.128 per 8 FMA-quads and
    // is shared-memory bound at ~0.36 of HBM (ncu: short-scoreboard stalls)
    constexpr int ITS = RPW >= 4 ? 4 : 8;
    constexpr int KC = 32 * W * ITS;  // floats / doubles per chunk
    constexpr int KCP = KC + W;       // +16 bytes per row
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* Bs = reinterpret_cast<T*>(smem_raw);  // [NJ][KCP]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles = (m + 8 * RPW - 1) / (8 * RPW);
    const bool single = k <= KC;
    auto stage = [&](int k0, int kc) {
        for (int kk = threadIdx.x; kk < KC; kk += 256) {
            if (kk < kc) {
                if (tb) {
#pragma unroll
                    for (int j = 0; j < NJ; ++j) Bs[j * KCP + kk] = j < n ? B[(size_t)j * k + k0 + kk] : T(0);
                } else {
                    const T* src = B + (size_t)(k0 + kk) * n;
#pragma unroll
                    for (int j = 0; j < NJ; ++j) Bs[j * KCP + kk] = j < n ? src[j] : T(0);
                }
            } else {
#pragma unroll
                for (int j = 0; j < NJ; ++j) Bs[j * KCP + kk] = T(0);
            }
        }
        __syncthreads();
    };
    if (single) stage(0, k);
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int row0 = (tile * 8 + warp) * RPW;
        T acc[RPW][NJ];
#pragma unroll
        for (int r = 0; r < RPW; ++r)
#pragma unroll
            for (int j = 0; j < NJ; ++j) acc[r][j] = T(0);
        for (int k0 = 0; k0 < k; k0 += KC) {
            const int kc = min(KC, k - k0);
            if (!single) {
                __syncthreads();  // every warp is done with the previous chunk of B
                stage(k0, kc);
            }
            V av[ITS][RPW];
#pragma unroll
            for (int it = 0; it < ITS; ++it) {
                const int q = it * 32 + lane;
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    const T* arow = A + (size_t)min(row0 + r, m - 1) * k + k0;  // clamped rows are discarded at the store
                    if (q * W < kc) av[it][r] = *reinterpret_cast<const V*>(arow + q * W);
                    else {
#pragma unroll
                        for (int l = 0; l < W; ++l) (&av[it][r].x)[l] = T(0);
                    }
                }
            }
#pragma unroll
            for (int it = 0; it < ITS; ++it) {
                const int q = it * 32 + lane;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const V bv = *reinterpret_cast<const V*>(Bs + j * KCP + q * W);
#pragma unroll
                    for (int r = 0; r < RPW; ++r)
#pragma unroll
                        for (int l = 0; l < W; ++l) acc[r][j] = fma(vget(av[it][r], l), vget(bv, l), acc[r][j]);
                }
            }
        }
        const int j_out = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            T v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = j < NJ ? acc[r][j < NJ ? j : 0] : T(0);
            warp_sum16(v, lane);
            const int row = row0 + r;
            // (dot + bias[i]) + v[j]: the roundings of Mul_M_M[_Add_V_MCols] followed by Add_M_M(., rows(v))
            if (!(lane & 1) && j_out < n && row < m) {
                const T g = v[0] + (bias ? bias[row] : T(0));
                C[(size_t)row * n + j_out] = bias_col ? g + bias_col[j_out] : g;
            }
        }
    }
}

// ---- A has M contiguous (ta = 1, stored k x m), 128-bit loads -----------------------------------
// grid (ceil(m / 128), chunks).  A CTA owns 128 output rows (= 128 columns of the stored A: one 512-byte segment per k-row,
// read by ONE warp instruction — round 1 read 128-byte segments 4 KB apart and ran at 0.22 of HBM) and one K chunk.
// Warp w, lane l: column vector l (W adjacent output rows) and k-lane w of 8, so the CTA walks 8 k-rows per step.  After the
// K loop the 8 warps are combined through shared memory in warp order and the CTA writes one partial per chunk.  Chunks are
// then summed in two fixed levels by whichever CTA arrives last (groups of 8 chunks, then the groups): one launch, and the
// summation order does not depend on the arrival order.
template <typename T, int NJ>
__global__ void __launch_bounds__(256, 1) skinny_cols_vec_kernel(const T* __restrict__ A, const T* __restrict__ B, const T* __restrict__ bias,
                                                               T* __restrict__ partial, T* __restrict__ partial2, T* __restrict__ C,
                                                               unsigned int* __restrict__ tickets, int m, int n, int k, int tb, int rows_per_chunk) {
    using V = typename V16<T>::type;
    constexpr int W = V16<T>::W;
    constexpr int COLS = 32 * W;   // 128 (float) / 64 (double) output rows per CTA
    constexpr int KL = 8;          // k-lanes per CTA (one per warp)
    constexpr int KCH = 256;       // k-rows of B staged per pass
    constexpr int GROUP = 8;
    // the B chunk and the cross-warp reduction buffer share storage (B is dead after the K loop)
    constexpr int kBsElems = KCH * NJ, kRedElems = 4 * 32 * (W * NJ + 1);
    __shared__ __align__(16) T smem_u[kBsElems > kRedElems ? kBsElems : kRedElems];
    T (*Bs)[NJ] = reinterpret_cast<T (*)[NJ]>(smem_u);
    T (*red)[32][W * NJ + 1] = reinterpret_cast<T (*)[32][W * NJ + 1]>(smem_u);
    __shared__ int s_last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col0 = blockIdx.x * COLS + lane * W;
    const int chunks = gridDim.y;
    const int kbeg = blockIdx.y * rows_per_chunk, kend = min(k, kbeg + rows_per_chunk);
    T acc[W][NJ];
#pragma unroll
    for (int c = 0; c < W; ++c)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[c][j] = T(0);
    for (int k0 = kbeg; k0 < kend; k0 += KCH) {
        const int kc = min(KCH, kend - k0);
        __syncthreads();
        {   // thread t stages k-row t of the chunk (KCH == blockDim.x): n loads, NJ/W vector stores
            const int kk = threadIdx.x;
            T row[NJ];
#pragma unroll
            for (int j = 0; j < NJ; ++j) row[j] = (kk < kc && j < n) ? ldB(B, tb, k, n, k0 + kk, j) : T(0);
#pragma unroll
            for (int j = 0; j < NJ; j += W) {
                V q;
#pragma unroll
                for (int l = 0; l < W; ++l) (&q.x)[l] = row[j + l];
                *reinterpret_cast<V*>(&Bs[kk][j]) = q;
            }
        }
        __syncthreads();
        if (col0 < m) {  // m % W == 0: a column vector is entirely inside or outside
            const T* a = A + (size_t)k0 * m + col0;
#pragma unroll 8
            for (int kk = warp; kk < kc; kk += KL) {
                const V av = *reinterpret_cast<const V*>(a + (size_t)kk * m);
#pragma unroll
                for (int j = 0; j < NJ; j += W) {
                    const V bv = *reinterpret_cast<const V*>(&Bs[kk][j]);   // same address in every lane: broadcast
#pragma unroll
                    for (int c = 0; c < W; ++c)
#pragma unroll
                        for (int l = 0; l < W; ++l) acc[c][j + l] = fma(vget(av, c), vget(bv, l), acc[c][j + l]);
                }
            }
        }
    }
    // combine the 8 warps (k-lanes) in a fixed tree: (w, w + 4), then (w, w + 2), then (0, 1); warp 0 ends up with the CTA's sums
    __syncthreads();  // every warp is done reading Bs
#pragma unroll
    for (int half = 4; half >= 1; half >>= 1) {
        if (warp >= half && warp < 2 * half) {
#pragma unroll
            for (int c = 0; c < W; ++c)
#pragma unroll
                for (int j = 0; j < NJ; ++j) red[warp - half][lane][c * NJ + j] = acc[c][j];
        }
        __syncthreads();
        if (warp < half) {
#pragma unroll
            for (int c = 0; c < W; ++c)
#pragma unroll
                for (int j = 0; j < NJ; ++j) acc[c][j] += red[warp][lane][c * NJ + j];
        }
        __syncthreads();
    }
    if (warp == 0) {
#pragma unroll
        for (int c = 0; c < W; ++c)
#pragma unroll
            for (int j = 0; j < NJ; ++j) red[0][lane][c * NJ + j] = acc[c][j];
    }
    __syncthreads();
    const int blk_cols = min(COLS, m - blockIdx.x * COLS);
    const int E = blk_cols * n;  // this CTA's outputs are contiguous in C and in every partial plane
    const size_t blk_off = (size_t)blockIdx.x * COLS * n;
    T* my_out = chunks == 1 ? C + blk_off : partial + (size_t)blockIdx.y * m * n + blk_off;
    for (int e = threadIdx.x; e < E; e += 256) {
        const int c = e / n, j = e % n;          // output row c of the block = column vector c / W, element c % W
        T s = red[0][c / W][(c % W) * NJ + j];
        if (chunks == 1 && bias) s += bias[blockIdx.x * COLS + c];
        my_out[e] = s;
    }
    if (chunks == 1) return;
    // ---- level 1: the last CTA of a group of 8 chunks adds the group's partials in chunk order ----
    const int ngroups = (chunks + GROUP - 1) / GROUP;
    const int g = blockIdx.y / GROUP;
    const int g_first = g * GROUP, g_size = min(GROUP, chunks - g_first);
    unsigned int* tk = tickets + blockIdx.x * (ngroups + 1);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&tk[g], 1u) == (unsigned)g_size - 1u);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    {
        T* dst = ngroups == 1 ? C + blk_off : partial2 + (size_t)g * m * n + blk_off;
        for (int e = threadIdx.x; e < E; e += 256) {
            T v[GROUP];
#pragma unroll
            for (int c = 0; c < GROUP; ++c) v[c] = c < g_size ? __ldcg(partial + (size_t)(g_first + c) * m * n + blk_off + e) : T(0);
            T s = v[0];
#pragma unroll
            for (int c = 1; c < GROUP; ++c) s += v[c];
            if (ngroups == 1 && bias) s += bias[blockIdx.x * COLS + e / n];
            dst[e] = s;
        }
    }
    if (threadIdx.x == 0) tk[g] = 0;
    if (ngroups == 1) return;
    // ---- level 2: the last group adds the group sums in group order ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&tk[ngroups], 1u) == (unsigned)ngroups - 1u);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int e = threadIdx.x; e < E; e += 256) {
        T s = T(0);
        int c = 0;
        for (; c + 8 <= ngroups; c += 8) {
            T v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcg(partial2 + (size_t)(c + u) * m * n + blk_off + e);
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        for (; c < ngroups; ++c) s += __ldcg(partial2 + (size_t)c * m * n + blk_off + e);
        if (bias) s += bias[blockIdx.x * COLS + e / n];
        C[blk_off + e] = s;
    }
    if (threadIdx.x == 0) tk[ngroups] = 0;
}

// ---- tiny K (k <= NMAX): C(m x n) = A(m x k) op(B)(k x n), output-write bound -----------------
// (Measured in round 2 and not adopted: applying the activation's adjoint chain to the finished element in this kernel instead
// of in the stream kernel that follows — 18.6 us against 6.5 + 4.3 us at 1024 rows, 49.4 us against 17.2 + 19.3 us at 8192:
// a thread walks its slab row by row, so the coshf / division latency of the chain is exposed instead of hidden by 64 warps.)
// (the input adjoint of the output layer: dZ (B x 10) * W3^T (10 x 1024)).  A thread keeps the KP
// (= k rounded up to 4) B-values of its 4 (2 for double) output columns in registers and walks the
// rows of its CTA's slab; the slab's A rows are staged in shared memory and read back as 16-byte
// broadcasts.
template <typename T, int KP>
__global__ void __launch_bounds__(256) tiny_k_kernel(const T* __restrict__ A, const T* __restrict__ B, const T* __restrict__ bias,
                                                      T* __restrict__ C, int m, int n, int k, int ta, int tb, int rows_per_cta) {
    using V = typename V16<T>::type;
    constexpr int W = V16<T>::W;
    constexpr int RS = 64;  // rows staged per pass
    const int c0 = (blockIdx.x * 256 + threadIdx.x) * W;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(m, r0 + rows_per_cta);
    __shared__ __align__(16) T As[RS][KP];
    T b[KP][W];
#pragma unroll
    for (int kk = 0; kk < KP; ++kk)
#pragma unroll
        for (int w = 0; w < W; ++w) b[kk][w] = (kk < k && c0 + w < n) ? (tb ? B[(size_t)(c0 + w) * k + kk] : B[(size_t)kk * n + c0 + w]) : T(0);
    const bool vec = (n % W == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    for (int rb = r0; rb < r1; rb += RS) {
        const int nr = min(RS, r1 - rb);
        __syncthreads();
        for (int idx = threadIdx.x; idx < RS * KP; idx += 256) {
            const int r = idx / KP, kk = idx % KP;
            As[r][kk] = (r < nr && kk < k) ? (ta ? A[(size_t)kk * m + rb + r] : A[(size_t)(rb + r) * k + kk]) : T(0);
        }
        __syncthreads();
        if (c0 >= n) continue;
#pragma unroll 2
        for (int r = 0; r < nr; ++r) {
            T acc[W];
#pragma unroll
            for (int w = 0; w < W; ++w) acc[w] = T(0);
#pragma unroll
            for (int kk = 0; kk < KP; kk += W) {
                const V av = *reinterpret_cast<const V*>(&As[r][kk]);  // same address in every lane: broadcast
#pragma unroll
                for (int l = 0; l < W; ++l)
#pragma unroll
                    for (int w = 0; w < W; ++w) acc[w] = fma(vget(av, l), b[kk + l][w], acc[w]);
            }
            const T bv = bias ? bias[rb + r] : T(0);
            T* dst = C + (size_t)(rb + r) * n + c0;
            if (vec) {
                V o;
#pragma unroll
                for (int w = 0; w < W; ++w) (&o.x)[w] = acc[w] + bv;
                *reinterpret_cast<V*>(dst) = o;
            } else {
#pragma unroll
                for (int w = 0; w < W; ++w)
                    if (c0 + w < n) dst[w] = acc[w] + bv;
            }
        }
    }
}

}  // namespace skinny

template <typename T, int RPW, int NJ>
static int launch_rows_vec_inst(Ctx* ctx, int m, int n, int k, int tb, const T* A, const T* B, const T* bias, const T* bias_col, T* C) {
    using namespace skinny;
    constexpr int W = 16 / (int)sizeof(T);
    constexpr int KCP = 32 * W * (RPW >= 4 ? 4 : 8) + W;
    const int smem = NJ * KCP * (int)sizeof(T);
    static bool attr_set = false;
    if (!attr_set) {
        DSC_CUDA_CHECK(cudaFuncSetAttribute(skinny_rows_vec_kernel<T, RPW, NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    const int tiles = (m + 8 * RPW - 1) / (8 * RPW);
    // every CTA walks the same number of tiles (no straggler wave): at most two CTAs per SM
    const int per_cta = (tiles + ctx->sm_count * 2 - 1) / (ctx->sm_count * 2);
    const int grid = (tiles + per_cta - 1) / per_cta;
    skinny_rows_vec_kernel<T, RPW, NJ><<<grid, 256, smem, ctx->stream>>>(A, B, bias, bias_col, C, m, n, k, tb);
    DSC_LAUNCH_CHECK(ctx);
    return DSC_OK;
}

template <typename T, int RPW>
static int launch_rows_vec_nj(Ctx* ctx, int m, int n, int k, int tb, const T* A, const T* B, const T* bias, const T* bias_col, T* C) {
    if (n <= 4) return launch_rows_vec_inst<T, RPW, 4>(ctx, m, n, k, tb, A, B, bias, bias_col, C);
    if (n <= 8) return launch_rows_vec_inst<T, RPW, 8>(ctx, m, n, k, tb, A, B, bias, bias_col, C);
    if (n <= 12) return launch_rows_vec_inst<T, RPW, 12>(ctx, m, n, k, tb, A, B, bias, bias_col, C);
    return launch_rows_vec_inst<T, RPW, 16>(ctx, m, n, k, tb, A, B, bias, bias_col, C);
}

template <typename T>
static int launch_rows_vec(Ctx* ctx, int m, int n, int k, int tb, const T* A, const T* B, const T* bias, const T* bias_col, T* C) {
    // two rows per warp once that still leaves >= 2 CTAs per SM (B^T is staged once per CTA).  Four rows per warp (half the
    // shared-memory reads per FMA, one CTA per SM) measured SLOWER: 18.1 us against 14.3 us at 8192 x 1024 x 10.
    if (m >= ctx->sm_count * 2 * 16) return launch_rows_vec_nj<T, 2>(ctx, m, n, k, tb, A, B, bias, bias_col, C);
    return launch_rows_vec_nj<T, 1>(ctx, m, n, k, tb, A, B, bias, bias_col, C);
}

template <typename T, int NJ>
static int launch_cols_vec_inst(Ctx* ctx, int m, int n, int k, int tb, const T* A, const T* B, const T* bias, T* C) {
    using namespace skinny;
    constexpr int W = 16 / (int)sizeof(T);
    const int gx = (m + 32 * W - 1) / (32 * W);
    // one full wave at 2 CTAs per SM (rounding the chunk count DOWN: a few CTAs more than a wave
    // would double the run time); every k-lane (8 per CTA) gets at least 4 k-rows
    int chunks = (ctx->sm_count * 2) / gx;
    const int max_chunks = (k + 127) / 128;   // every k-lane (one warp) walks at least 16 k-rows
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    int rows_per_chunk = (k + chunks - 1) / chunks;
    chunks = (k + rows_per_chunk - 1) / rows_per_chunk;
    const int ngroups = (chunks + 7) / 8;
    if ((long)gx * (ngroups + 1) > kNumTickets) return 1;  // caller falls back to the two-stage kernels
    const size_t p1 = (size_t)chunks * m * n * sizeof(T), p2 = (size_t)ngroups * m * n * sizeof(T);
    DSC_TRY(ensure_red_scratch(ctx, 256 + p1 + p2 + 32));
    char* base = reinterpret_cast<char*>(ctx->red_scratch) + 256;
    T* partial = reinterpret_cast<T*>(base);
    T* partial2 = reinterpret_cast<T*>(base + ((p1 + 15) & ~size_t(15)));
    skinny_cols_vec_kernel<T, NJ><<<dim3(gx, chunks), 256, 0, ctx->stream>>>(A, B, bias, partial, partial2, C, ctx->tickets, m, n, k, tb, rows_per_chunk);
    DSC_LAUNCH_CHECK(ctx);
    return DSC_OK;
}

template <typename T>
static int launch_cols_vec(Ctx* ctx, int m, int n, int k, int tb, const T* A, const T* B, const T* bias, T* C) {
    if (n <= 4) return launch_cols_vec_inst<T, 4>(ctx, m, n, k, tb, A, B, bias, C);
    if (n <= 8) return launch_cols_vec_inst<T, 8>(ctx, m, n, k, tb, A, B, bias, C);
    if (n <= 12) return launch_cols_vec_inst<T, 12>(ctx, m, n, k, tb, A, B, bias, C);
    return launch_cols_vec_inst<T, 16>(ctx, m, n, k, tb, A, B, bias, C);
}

// returns DSC_OK when handled, 1 when the shape is not a skinny / tiny-K product
template <typename T>
int gemm_skinny(Ctx* ctx, int ta, int tb, int m, int n, int k, const T* A, const T* B, const T* bias, T* C, const T* bias_col, int* bias_col_done) {
    using namespace skinny;
    if (bias_col_done) *bias_col_done = 0;
    if (k <= NMAX && n >= 64 && (long)m * n >= 65536) {
        constexpr int W = 16 / (int)sizeof(T);
        const int gx = (n + 256 * W - 1) / (256 * W);
        // two CTAs per SM, all resident in one wave (512 CTAs at three per SM ran as 1.15 waves: the straggler wave doubled the
        // time), slabs of at least 16 rows (the B registers are loaded once per CTA)
        int gy = (ctx->sm_count * 2 + gx - 1) / gx;
        int rows_per_cta = (m + gy - 1) / gy;
        rows_per_cta = (rows_per_cta + 15) / 16 * 16;
        gy = (m + rows_per_cta - 1) / rows_per_cta;
        const dim3 grid(gx, gy);
        if (k <= 4) tiny_k_kernel<T, 4><<<grid, 256, 0, ctx->stream>>>(A, B, bias, C, m, n, k, ta, tb, rows_per_cta);
        else if (k <= 8) tiny_k_kernel<T, 8><<<grid, 256, 0, ctx->stream>>>(A, B, bias, C, m, n, k, ta, tb, rows_per_cta);
        else if (k <= 12) tiny_k_kernel<T, 12><<<grid, 256, 0, ctx->stream>>>(A, B, bias, C, m, n, k, ta, tb, rows_per_cta);
        else tiny_k_kernel<T, 16><<<grid, 256, 0, ctx->stream>>>(A, B, bias, C, m, n, k, ta, tb, rows_per_cta);
        DSC_LAUNCH_CHECK(ctx);
        ctx->stats.gemm_simt++;
        return DSC_OK;
    }
    if (n > NMAX || (long)m * k < 65536) return 1;
    constexpr int W = 16 / (int)sizeof(T);
    const bool a16 = (reinterpret_cast<uintptr_t>(A) & 15) == 0;
    int vec_cols = 1;
    if (!ta && a16 && k % W == 0) {
        DSC_TRY((launch_rows_vec<T>(ctx, m, n, k, tb, A, B, bias, bias_col, C)));
        if (bias_col && bias_col_done) *bias_col_done = 1;
    } else if (ta && a16 && m % W == 0 && (vec_cols = launch_cols_vec<T>(ctx, m, n, k, tb, A, B, bias, C)) != 1) {
        if (vec_cols != DSC_OK) return vec_cols;
    } else if (!ta) {
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

template int gemm_skinny<float>(Ctx*, int, int, int, int, int, const float*, const float*, const float*, float*, const float*, int*);
template int gemm_skinny<double>(Ctx*, int, int, int, int, int, const double*, const double*, const double*, double*, const double*, int*);

}  // namespace dsc
