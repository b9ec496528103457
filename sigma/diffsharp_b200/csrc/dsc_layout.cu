// K8/K9/K10: transpose, N-d permute, broadcasts, column reduction, diagonal.
// All are pure data movement or one add per element: HBM-bound, bit-exact where no arithmetic is
// involved (Transpose_M, Permute_M, RepeatReshapeCopy_*, Diagonal_M).
#include "dsc_internal.cuh"

namespace dsc {

// ---- Transpose_M (Backend.fs:127) ---------------------------------------------------------
// 64x64 tiles through shared memory.  Each thread moves 128-bit vectors on both sides when the
// matrix dimensions and windows allow it (m, n multiples of the vector width and 16 B alignment);
// the pad column keeps both the row-wise stores and the column-wise loads conflict-free.
template <typename T, int TILE>
__global__ void __launch_bounds__(256) transpose_kernel(const T* __restrict__ A, T* __restrict__ B, int m, int n) {
    __shared__ T tile[TILE][TILE + 1];
    const int bx = blockIdx.x * TILE;  // column origin in A
    const int by = blockIdx.y * TILE;  // row origin in A
    const int tx = threadIdx.x % TILE, ty = threadIdx.x / TILE;
    constexpr int ROWS_PER_PASS = 256 / TILE;
#pragma unroll
    for (int r = ty; r < TILE; r += ROWS_PER_PASS) {
        int gr = by + r, gc = bx + tx;
        if (gr < m && gc < n) tile[r][tx] = A[(size_t)gr * n + gc];
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < TILE; r += ROWS_PER_PASS) {
        int gr = bx + r, gc = by + tx;  // B is n x m
        if (gr < n && gc < m) B[(size_t)gr * m + gc] = tile[tx][r];
    }
}

// vectorised variant for f32: 64x64 tile, float4 global accesses, scalar shared accesses.
__global__ void __launch_bounds__(256) transpose_f32_v4_kernel(const float* __restrict__ A, float* __restrict__ B, int m, int n) {
    __shared__ float tile[64][65];
    const int bx = blockIdx.x * 64, by = blockIdx.y * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4 floats each along x
#pragma unroll
    for (int r = ty; r < 64; r += 16) {
        int gr = by + r, gc = bx + tx * 4;
        if (gr < m && gc < n) {  // n % 4 == 0 guaranteed by the host
            float4 v = *reinterpret_cast<const float4*>(A + (size_t)gr * n + gc);
            tile[r][tx * 4 + 0] = v.x; tile[r][tx * 4 + 1] = v.y; tile[r][tx * 4 + 2] = v.z; tile[r][tx * 4 + 3] = v.w;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 64; r += 16) {
        int gr = bx + r, gc = by + tx * 4;
        if (gr < n && gc < m) {
            float4 v;
            v.x = tile[tx * 4 + 0][r]; v.y = tile[tx * 4 + 1][r]; v.z = tile[tx * 4 + 2][r]; v.w = tile[tx * 4 + 3][r];
            *reinterpret_cast<float4*>(B + (size_t)gr * m + gc) = v;
        }
    }
}

__global__ void __launch_bounds__(256) transpose_f64_v2_kernel(const double* __restrict__ A, double* __restrict__ B, int m, int n) {
    __shared__ double tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 2 doubles each along x
#pragma unroll
    for (int r = ty; r < 32; r += 16) {
        int gr = by + r, gc = bx + tx * 2;
        if (gr < m && gc < n) {
            double2 v = *reinterpret_cast<const double2*>(A + (size_t)gr * n + gc);
            tile[r][tx * 2 + 0] = v.x; tile[r][tx * 2 + 1] = v.y;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 16) {
        int gr = bx + r, gc = by + tx * 2;
        if (gr < n && gc < m) {
            double2 v;
            v.x = tile[tx * 2 + 0][r]; v.y = tile[tx * 2 + 1][r];
            *reinterpret_cast<double2*>(B + (size_t)gr * m + gc) = v;
        }
    }
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <>
int transpose_launch<float>(Ctx* ctx, const float* A, float* B, int m, int n) {
    if ((m % 4 == 0) && (n % 4 == 0) && al16(A) && al16(B)) {
        dim3 grid((n + 63) / 64, (m + 63) / 64);
        transpose_f32_v4_kernel<<<grid, 256, 0, ctx->stream>>>(A, B, m, n);
    } else {
        dim3 grid((n + 31) / 32, (m + 31) / 32);
        transpose_kernel<float, 32><<<grid, 256, 0, ctx->stream>>>(A, B, m, n);
    }
    DSC_LAUNCH_CHECK(ctx);
    return DSC_OK;
}
template <>
int transpose_launch<double>(Ctx* ctx, const double* A, double* B, int m, int n) {
    dim3 grid((n + 31) / 32, (m + 31) / 32);
    if ((m % 2 == 0) && (n % 2 == 0) && al16(A) && al16(B))
        transpose_f64_v2_kernel<<<grid, 256, 0, ctx->stream>>>(A, B, m, n);
    else
        transpose_kernel<double, 32><<<grid, 256, 0, ctx->stream>>>(A, B, m, n);
    DSC_LAUNCH_CHECK(ctx);
    return DSC_OK;
}

// ---- Permute_M (Backend.fs:128): out.shape[d] = in.shape[perm[d]] --------------------------
struct PermuteArgs {
    int rank;
    long long out_shape[8];
    long long in_stride_for_out[8];  // stride in the input of output axis d
};
template <typename T>
__global__ void permute_kernel(const T* __restrict__ in, T* __restrict__ out, size_t n, PermuteArgs a) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        size_t rem = i, src = 0;
#pragma unroll 1
        for (int d = a.rank - 1; d >= 0; --d) {
            size_t c = rem % (size_t)a.out_shape[d];
            rem /= (size_t)a.out_shape[d];
            src += c * (size_t)a.in_stride_for_out[d];
        }
        out[i] = in[src];
    }
}

// ---- broadcasts ---------------------------------------------------------------------------
// out[i, j] = v[j]  (RepeatReshapeCopy_V_MRows, Backend.fs:135; golden Script.fsx:7-16)
// out[i, j] = v[i]  (RepeatReshapeCopy_V_MCols, Backend.fs:136)
template <typename T, bool ROWS>
__global__ void repeat_kernel(const T* __restrict__ v, T* __restrict__ out, int m, int n) {
    size_t total = (size_t)m * n;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) out[i] = ROWS ? v[i % n] : v[i / n];
}
// 128-bit store variant of the row broadcast (n % W == 0, aligned)
template <typename T, typename V, int W>
__global__ void repeat_rows_vec_kernel(const T* __restrict__ v, T* __restrict__ out, int m, int nvec) {
    size_t total = (size_t)m * nvec;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        V x = reinterpret_cast<const V*>(v)[i % nvec];
        reinterpret_cast<V*>(out)[i] = x;
    }
}

// out[i,j] = M[i,j] + v[i]  (Add_V_MCols, Backend.fs:115)
template <typename T>
__global__ void add_v_mcols_kernel(const T* __restrict__ v, const T* __restrict__ M, T* __restrict__ out, int m, int n) {
    size_t total = (size_t)m * n;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) out[i] = M[i] + v[i / n];
}

template <typename T>
__global__ void diag_kernel(const T* __restrict__ A, T* __restrict__ out, int n, int k) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) out[i] = A[(size_t)i * n + i];
}

// ---- column reduction (Add_M_Colwise_V_InPlace, Backend.fs:123) ---------------------------
// v[j] += sum_i M[i,j] in ONE launch.  The rows are cut into `chunks` contiguous slabs; block
// (bx, chunk) sums its slab for 32*W columns with 8 row lanes (8 independent 128-bit loads in flight
// per thread) and a fixed shared-memory tree over the lanes, and writes one partial row.  The CTA
// that arrives last for a column block (ticket counter) adds the slab partials — row lane ty takes
// the slabs c = ty (mod 8) in increasing order, then the same fixed tree — and accumulates into v.
// Which CTA does the last step varies; the order of the additions does not (deterministic).
template <typename T, int W>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ M, T* __restrict__ partial, T* __restrict__ v,
                                                      unsigned int* __restrict__ tickets, int m, int n, int rows_per_chunk, int accumulate) {
    __shared__ T sm[8][32 * W + 1];
    __shared__ int s_last;
    const int col0 = (blockIdx.x * 32 + threadIdx.x) * W;
    const int r_begin = blockIdx.y * rows_per_chunk;
    const int r_end = min(m, r_begin + rows_per_chunk);
    const int chunks = gridDim.y;
    T acc[W];
#pragma unroll
    for (int w = 0; w < W; ++w) acc[w] = T(0);
    auto tree = [&]() {   // sum of the 8 row lanes, valid in row lane 0
#pragma unroll
        for (int w = 0; w < W; ++w) sm[threadIdx.y][threadIdx.x * W + w] = acc[w];
        __syncthreads();
        if (threadIdx.y == 0) {
#pragma unroll
            for (int w = 0; w < W; ++w) {
                const int c = threadIdx.x * W + w;
                acc[w] = ((sm[0][c] + sm[1][c]) + (sm[2][c] + sm[3][c])) + ((sm[4][c] + sm[5][c]) + (sm[6][c] + sm[7][c]));
            }
        }
    };
    if (col0 < n) {
        auto load8 = [&](int r, T (&x)[8][W]) {   // rows r, r + 8, ..., r + 56 of this thread's column vector
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int rr = r + 8 * u;
#pragma unroll
                for (int w = 0; w < W; ++w) x[u][w] = T(0);
                if (rr < r_end) {
                    const T* row = M + (size_t)rr * n + col0;
                    if constexpr (W == 4) {
                        const float4 q = *reinterpret_cast<const float4*>(row);
                        x[u][0] = q.x; x[u][1] = q.y; x[u][2] = q.z; x[u][3] = q.w;
                    } else if constexpr (W == 2) {
                        const double2 q = *reinterpret_cast<const double2*>(row);
                        x[u][0] = q.x; x[u][1] = q.y;
                    } else {
                        x[u][0] = row[0];
                    }
                }
            }
        };
        // software pipeline: the loads of the next 64 rows are in flight while the current ones are added (same order of additions)
        T xa[8][W], xb[8][W];
        int r = r_begin + threadIdx.y;
        if (r < r_end) load8(r, xa);
        for (; r < r_end; r += 128) {
            if (r + 64 < r_end) load8(r + 64, xb);
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int w = 0; w < W; ++w) acc[w] += xa[u][w];
            if (r + 64 < r_end) {
                if (r + 128 < r_end) load8(r + 128, xa);
#pragma unroll
                for (int u = 0; u < 8; ++u)
#pragma unroll
                    for (int w = 0; w < W; ++w) acc[w] += xb[u][w];
            }
        }
    }
    tree();
    if (chunks == 1) {
        if (threadIdx.y == 0 && col0 < n) {
#pragma unroll
            for (int w = 0; w < W; ++w) v[col0 + w] = accumulate ? v[col0 + w] + acc[w] : acc[w];
        }
        return;
    }
    if (threadIdx.y == 0 && col0 < n) {
#pragma unroll
        for (int w = 0; w < W; ++w) partial[(size_t)blockIdx.y * n + col0 + w] = acc[w];
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) s_last = (atomicAdd(&tickets[blockIdx.x], 1u) == (unsigned)chunks - 1u);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
#pragma unroll
    for (int w = 0; w < W; ++w) acc[w] = T(0);
    if (col0 < n) {
        for (int c = threadIdx.y; c < chunks; c += 64) {
            T x[8][W];
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int w = 0; w < W; ++w) x[u][w] = (c + 8 * u < chunks) ? __ldcg(partial + (size_t)(c + 8 * u) * n + col0 + w) : T(0);
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int w = 0; w < W; ++w) acc[w] += x[u][w];
        }
    }
    tree();
    if (threadIdx.y == 0 && col0 < n) {
#pragma unroll
        for (int w = 0; w < W; ++w) v[col0 + w] = accumulate ? v[col0 + w] + acc[w] : acc[w];
    }
    if (threadIdx.x == 0 && threadIdx.y == 0) tickets[blockIdx.x] = 0;
}

// Narrow matrices (n <= 32: the 10-column output layer).  The main kernel would run 10 of every 32 lanes; here a CTA takes a
// contiguous slab of rows as a FLAT array: thread (rl, c) owns column c and row lane rl of RL = 256 / n lanes, so a warp reads
// ~3 consecutive rows (contiguous bytes).  Row lanes are added in lane order by thread (0, c), slabs in slab order by the CTA
// that arrives last: one launch, fixed order.
template <typename T>
__global__ void __launch_bounds__(256) colsum_narrow_kernel(const T* __restrict__ M, T* __restrict__ partial, T* __restrict__ v,
                                                             unsigned int* __restrict__ ticket, int m, int n, int rows_per_cta, int accumulate) {
    __shared__ T sm[256];
    __shared__ int s_last;
    const int RL = 256 / n;                       // row lanes
    const int rl = threadIdx.x / n, c = threadIdx.x % n;
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(m, r0 + rows_per_cta);
    T acc[4] = {T(0), T(0), T(0), T(0)};
    if (rl < RL) {
        int r = r0 + rl;
        for (; r + 3 * RL < r1; r += 4 * RL) {    // 4 independent loads in flight
            T x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) x[u] = M[(size_t)(r + u * RL) * n + c];
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u] += x[u];
        }
        for (; r < r1; r += RL) acc[0] += M[(size_t)r * n + c];
    }
    sm[threadIdx.x] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    __syncthreads();
    T tot = T(0);
    if (threadIdx.x < n) {
        for (int l = 0; l < RL; ++l) tot += sm[l * n + threadIdx.x];
        if (gridDim.x == 1) v[threadIdx.x] = accumulate ? v[threadIdx.x] + tot : tot;
        else partial[(size_t)blockIdx.x * n + threadIdx.x] = tot;
    }
    if (gridDim.x == 1) return;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1u);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x < n) {
        T s = T(0);
        for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(partial + (size_t)b * n + threadIdx.x);
        v[threadIdx.x] = accumulate ? v[threadIdx.x] + s : s;
    }
    if (threadIdx.x == 0) *ticket = 0;
}

template <typename T>
int colsum_launch(Ctx* ctx, const T* M, T* v, int m, int n, bool accumulate) {
    if (n <= 32) {
        // a CTA takes 4 rows per row lane (one batch of 4 independent loads per thread: the launch is latency-bound, the matrix is
        // a few hundred KB), more only when that would exceed two CTAs per SM
        const int RL = 256 / n;
        int rows_per_cta = 4 * RL;
        int ctas = (m + rows_per_cta - 1) / rows_per_cta;
        if (ctas > 2 * ctx->sm_count) {
            rows_per_cta = (m + 2 * ctx->sm_count - 1) / (2 * ctx->sm_count);
            ctas = (m + rows_per_cta - 1) / rows_per_cta;
        }
        DSC_TRY(ensure_red_scratch(ctx, 256 + (size_t)ctas * n * sizeof(T)));
        T* partial = reinterpret_cast<T*>(reinterpret_cast<char*>(ctx->red_scratch) + 256);
        colsum_narrow_kernel<T><<<ctas, 256, 0, ctx->stream>>>(M, partial, v, ctx->tickets, m, n, rows_per_cta, accumulate ? 1 : 0);
        DSC_LAUNCH_CHECK(ctx);
        return DSC_OK;
    }
    constexpr int W = sizeof(T) == 4 ? 4 : 2;
    bool vec = (n % W == 0) && al16(M);
    int cols_per_block = vec ? 32 * W : 32;
    int gx = (n + cols_per_block - 1) / cols_per_block;
    // Row slabs: a CTA walks its slab 64 rows at a time (every thread has 8 loads of 16 bytes in flight).  The slab count is
    // chosen so that the grid is about four CTAs per SM (all resident): with one CTA per 64-row slab (round 1) a tall matrix put 128 CTAs
    // on every ticket counter and the launch ran at 0.36 of HBM — the fence + atomic + last-arriver pass per CTA, not the
    // loads, set the time.  Slabs are multiples of 64 rows and never smaller than 64.
    int rows_per_chunk = 64;
    {
        const int want_ctas = 4 * ctx->sm_count;
        int chunks_want = want_ctas / gx;
        if (chunks_want < 1) chunks_want = 1;
        int rpc = (m + chunks_want - 1) / chunks_want;
        rpc = (rpc + 63) / 64 * 64;
        if (rpc > rows_per_chunk) rows_per_chunk = rpc;
    }
    int chunks = (m + rows_per_chunk - 1) / rows_per_chunk;
    if (gx > kNumTickets) return fail(DSC_E_SHAPE, "Add_M_Colwise_V_InPlace: %d columns exceed the supported maximum", n);
    size_t need = 256 + (size_t)chunks * n * sizeof(T);
    DSC_TRY(ensure_red_scratch(ctx, need));
    T* partial = reinterpret_cast<T*>(reinterpret_cast<char*>(ctx->red_scratch) + 256);
    dim3 block(32, 8), grid(gx, chunks);
    if (vec) colsum_kernel<T, W><<<grid, block, 0, ctx->stream>>>(M, partial, v, ctx->tickets, m, n, rows_per_chunk, accumulate ? 1 : 0);
    else colsum_kernel<T, 1><<<grid, block, 0, ctx->stream>>>(M, partial, v, ctx->tickets, m, n, rows_per_chunk, accumulate ? 1 : 0);
    DSC_LAUNCH_CHECK(ctx);
    return DSC_OK;
}

// ---- typed entry points -------------------------------------------------------------------
template <typename T>
static int transpose_impl(dsc_ctx_t c, int32_t m, int32_t n, dsc_buf_t Ah, dsc_buf_t* outh) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *A, *o;
    DSC_TRY(in_buf_peek(ctx, dtype_of<T>::value, Ah, &A, "Transpose_M: A"));
    if (m < 0 || n < 0 || (int64_t)m * n > A->len) return fail(DSC_E_SHAPE, "Transpose_M: %d x %d exceeds buffer length %d", m, n, A->len);
    if (!outh) return fail(DSC_E_INVALID, "null output handle pointer");
    if (ctx->fuse && *outh == 0 && (int64_t)m * n > 0 && (int64_t)m * n == A->len) {
        // every weight / input adjoint transposes an operand right before Mul_M_M (AD.Float32.fs:4127-4143): the product reads
        // the stored matrix through a transposed tile map instead; any other consumer materialises the transpose
        LazyNode* nd = lazy_new(ctx, LZ_T, A->dtype, m * n);
        nd->m = m; nd->n = n;
        lazy_set_operand(nd, 0, A);
        return lazy_buf(ctx, nd, outh);
    }
    DSC_TRY(buf_resolve(A));
    DSC_TRY(out_buf(ctx, dtype_of<T>::value, m * n, outh, &o));
    if (m * n == 0) return DSC_OK;
    if (o->ptr<T>() == A->ptr<T>()) return fail(DSC_E_INVALID, "Transpose_M: in-place transpose is not supported");
    return transpose_launch<T>(ctx, A->ptr<T>(), o->ptr<T>(), m, n);
}

template <typename T>
static int permute_impl(dsc_ctx_t c, int rank, const int64_t* shape, const int32_t* perm, dsc_buf_t Ah, dsc_buf_t* outh) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *A, *o;
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, Ah, &A, "Permute_M: A"));
    if (rank < 1 || rank > 8 || !shape || !perm) return fail(DSC_E_INVALID, "Permute_M: rank %d unsupported (1..8)", rank);
    int64_t total = 1, in_stride[8];
    bool seen[8] = {false};
    for (int d = rank - 1; d >= 0; --d) {
        if (shape[d] < 0) return fail(DSC_E_SHAPE, "Permute_M: negative extent");
        in_stride[d] = total;
        total *= shape[d];
    }
    if (total > A->len) return fail(DSC_E_SHAPE, "Permute_M: shape exceeds buffer length");
    PermuteArgs a;
    a.rank = rank;
    for (int d = 0; d < rank; ++d) {
        if (perm[d] < 0 || perm[d] >= rank || seen[perm[d]]) return fail(DSC_E_INVALID, "Permute_M: dims is not a permutation");
        seen[perm[d]] = true;
        a.out_shape[d] = shape[perm[d]];
        a.in_stride_for_out[d] = in_stride[perm[d]];
    }
    DSC_TRY(out_buf(ctx, dtype_of<T>::value, (int32_t)total, outh, &o));
    if (total == 0) return DSC_OK;
    int grid = grid_for((size_t)total, 256 * 4, ctx->sm_count * 8);
    permute_kernel<T><<<grid, 256, 0, ctx->stream>>>(A->ptr<T>(), o->ptr<T>(), (size_t)total, a);
    DSC_LAUNCH_CHECK(ctx);
    return DSC_OK;
}

template <typename T>
int repeat_rows_launch(Ctx* ctx, const T* v, T* out, int m, int n) {
    constexpr int W = sizeof(T) == 4 ? 4 : 2;
    const size_t total = (size_t)m * n;
    if ((n % W == 0) && al16(v) && al16(out)) {
        using V = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
        int grid = grid_for(total / W, 256 * 4, ctx->sm_count * 8);
        repeat_rows_vec_kernel<T, V, W><<<grid, 256, 0, ctx->stream>>>(v, out, m, n / W);
    } else {
        int grid = grid_for(total, 256 * 4, ctx->sm_count * 8);
        repeat_kernel<T, true><<<grid, 256, 0, ctx->stream>>>(v, out, m, n);
    }
    DSC_LAUNCH_CHECK(ctx);
    return DSC_OK;
}
template int colsum_launch<float>(Ctx*, const float*, float*, int, int, bool);
template int colsum_launch<double>(Ctx*, const double*, double*, int, int, bool);
template int repeat_rows_launch<float>(Ctx*, const float*, float*, int, int);
template int repeat_rows_launch<double>(Ctx*, const double*, double*, int, int);

template <typename T, bool ROWS>
static int repeat_impl(dsc_ctx_t c, int32_t reps, dsc_buf_t vh, dsc_buf_t* outh) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *v, *o;
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, vh, &v, "RepeatReshapeCopy: v"));
    if (reps < 0) return fail(DSC_E_SHAPE, "RepeatReshapeCopy: negative repeat count");
    int64_t total = (int64_t)reps * v->len;
    if (total > 0x7fffffff) return fail(DSC_E_SHAPE, "RepeatReshapeCopy: result too large for an int32 length");
    if (!outh) return fail(DSC_E_INVALID, "null output handle pointer");
    if (ROWS && ctx->fuse && *outh == 0 && total > 0) {
        // the fork broadcasts a bias as m copies of the vector and adds them (AD.Float32.fs:1860-1865): the copies are written
        // only if somebody other than Add_M_M reads them
        LazyNode* nd = lazy_new(ctx, LZ_ROWS, v->dtype, (int32_t)total);
        nd->m = reps; nd->n = v->len;
        lazy_set_operand(nd, 0, v);
        return lazy_buf(ctx, nd, outh);
    }
    DSC_TRY(out_buf(ctx, dtype_of<T>::value, (int32_t)total, outh, &o));
    if (total == 0) return DSC_OK;
    int m = ROWS ? reps : v->len, n = ROWS ? v->len : reps;
    if (ROWS) return repeat_rows_launch<T>(ctx, v->ptr<T>(), o->ptr<T>(), m, n);
    int grid = grid_for((size_t)total, 256 * 4, ctx->sm_count * 8);
    repeat_kernel<T, ROWS><<<grid, 256, 0, ctx->stream>>>(v->ptr<T>(), o->ptr<T>(), m, n);
    DSC_LAUNCH_CHECK(ctx);
    return DSC_OK;
}

template <typename T>
static int colsum_impl(dsc_ctx_t c, int32_t m, int32_t n, dsc_buf_t Mh, dsc_buf_t vh) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *M, *v;
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, Mh, &M, "Add_M_Colwise_V_InPlace: M"));
    DSC_TRY(in_buf_peek(ctx, dtype_of<T>::value, vh, &v, "Add_M_Colwise_V_InPlace: v"));
    if (m < 0 || n < 0 || (int64_t)m * n > M->len) return fail(DSC_E_SHAPE, "Add_M_Colwise_V_InPlace: %d x %d exceeds buffer length %d", m, n, M->len);
    if (v->len != n) return fail(DSC_E_SHAPE, "Add_M_Colwise_V_InPlace: vector length %d, matrix has %d columns", v->len, n);
    if (m == 0 || n == 0) return DSC_OK;
    if (LazyNode* z = lazy_of(v, LZ_ZERO)) {
        // a bias adjoint is a fresh zero vector when its column sums arrive (AD.Float32.fs:4255 after :3640-3641): 0 + s = s,
        // so the sums are stored instead of added and the zeros are never written
        v->lazy = nullptr;
        lazy_unref(z);
        DSC_TRY(pool_alloc(ctx, (size_t)n * sizeof(T), &v->block));
        v->off = 0;
        // A sink of the reverse sweep (AD.Float32.fs:4255): when the second compute lane is already ordered after the matrix
        // (the weight adjoint of the same layer read it, or the lane was forked in front of that product), the sums are
        // taken there — beside whatever the compute stream has been given since, not behind it.
        Block* rd[1] = {M->block};
        // The matrix is the newest adjoint of the sweep: a fork point was recorded in front of the weight adjoint that reads it
        // (lane_prefork), but the lane has not been made to wait for it — the column sums of the other layers, which the calls
        // right behind this one issue and whose inputs have been ready for a long time, should not queue behind that wait.
        // These sums are deferred (a pending value like any other); at the next flush they run on the lane, behind the others.
        if (ctx->fuse && ctx->lanes_enabled && !ctx->on_lane && M->block->lane_w <= ctx->lane_joined && !lane_sees(ctx, rd, 1) &&
            lane_prefork_covers(ctx, M->block)) {
            block_release(v->block);
            v->block = nullptr;
            LazyNode* nd = lazy_new(ctx, LZ_COLSUM, v->dtype, n);
            nd->m = m; nd->n = n;
            lazy_set_operand(nd, 0, M);
            v->lazy = nd;
            return DSC_OK;
        }
        if (lane_debug()) fprintf(stderr, "[dsc lane] column sums %d x %d of %p: lane busy %d, lane-written %d\n", m, n, M->block->ptr, (int)lane_busy(ctx), (int)(M->block->lane_w > ctx->lane_joined));
        if (ctx->lanes_enabled && !ctx->on_lane && M->block->lane_w <= ctx->lane_joined && lane_sees(ctx, rd, 1)) {
            DSC_TRY(lane_enter(ctx, rd, 1, false));
            const int st = colsum_launch<T>(ctx, M->ptr<T>(), v->ptr<T>(), m, n, false);
            Block* wr[1] = {v->block};
            lane_leave(ctx, rd, 1, wr, 1);
            return st;
        }
        return colsum_launch<T>(ctx, M->ptr<T>(), v->ptr<T>(), m, n, false);
    }
    DSC_TRY(buf_writable(ctx, v));
    return colsum_launch<T>(ctx, M->ptr<T>(), v->ptr<T>(), m, n, true);
}

template <typename T>
static int add_v_mcols_impl(dsc_ctx_t c, int32_t m, int32_t n, dsc_buf_t vh, dsc_buf_t Mh, dsc_buf_t* outh) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *M, *v, *o;
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, Mh, &M, "Add_V_MCols: M"));
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, vh, &v, "Add_V_MCols: v"));
    if (m < 0 || n < 0 || (int64_t)m * n > M->len) return fail(DSC_E_SHAPE, "Add_V_MCols: %d x %d exceeds buffer length %d", m, n, M->len);
    if (v->len != m) return fail(DSC_E_SHAPE, "Add_V_MCols: vector length %d, matrix has %d rows", v->len, m);
    DSC_TRY(out_buf(ctx, dtype_of<T>::value, m * n, outh, &o));
    if (m * n == 0) return DSC_OK;
    int grid = grid_for((size_t)m * n, 256 * 4, ctx->sm_count * 8);
    add_v_mcols_kernel<T><<<grid, 256, 0, ctx->stream>>>(v->ptr<T>(), M->ptr<T>(), o->ptr<T>(), m, n);
    DSC_LAUNCH_CHECK(ctx);
    return DSC_OK;
}

template <typename T>
static int diag_impl(dsc_ctx_t c, int32_t m, int32_t n, dsc_buf_t Ah, dsc_buf_t* outh) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *A, *o;
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, Ah, &A, "Diagonal_M: A"));
    if (m < 0 || n < 0 || (int64_t)m * n > A->len) return fail(DSC_E_SHAPE, "Diagonal_M: %d x %d exceeds buffer length %d", m, n, A->len);
    int k = m < n ? m : n;
    DSC_TRY(out_buf(ctx, dtype_of<T>::value, k, outh, &o));
    if (k == 0) return DSC_OK;
    diag_kernel<T><<<(k + 255) / 256, 256, 0, ctx->stream>>>(A->ptr<T>(), o->ptr<T>(), n, k);
    DSC_LAUNCH_CHECK(ctx);
    return DSC_OK;
}

}  // namespace dsc

using namespace dsc;

extern "C" {
#define DSC_LAYOUT_API(P, T)                                                                                              \
    int dsc_##P##transpose(dsc_ctx_t c, int32_t m, int32_t n, dsc_buf_t A, dsc_buf_t* out) { DSC_NVTX(c, "Transpose_M"); return transpose_impl<T>(c, m, n, A, out); } \
    int dsc_##P##permute(dsc_ctx_t c, int rank, const int64_t* shape, const int32_t* perm, dsc_buf_t A, dsc_buf_t* out) { DSC_NVTX(c, "Permute_M"); \
        return permute_impl<T>(c, rank, shape, perm, A, out);                                                             \
    }                                                                                                                     \
    int dsc_##P##repeat_rows(dsc_ctx_t c, int32_t m, dsc_buf_t v, dsc_buf_t* out) { DSC_NVTX(c, "RepeatReshapeCopy_V_MRows"); return repeat_impl<T, true>(c, m, v, out); } \
    int dsc_##P##repeat_cols(dsc_ctx_t c, int32_t n, dsc_buf_t v, dsc_buf_t* out) { DSC_NVTX(c, "RepeatReshapeCopy_V_MCols"); return repeat_impl<T, false>(c, n, v, out); } \
    int dsc_##P##colsum_inplace(dsc_ctx_t c, int32_t m, int32_t n, dsc_buf_t M, dsc_buf_t v) { DSC_NVTX(c, "Add_M_Colwise_V_InPlace"); return colsum_impl<T>(c, m, n, M, v); } \
    int dsc_##P##add_v_mcols(dsc_ctx_t c, int32_t m, int32_t n, dsc_buf_t v, dsc_buf_t M, dsc_buf_t* out) { DSC_NVTX(c, "Add_V_MCols");               \
        return add_v_mcols_impl<T>(c, m, n, v, M, out);                                                                   \
    }                                                                                                                     \
    int dsc_##P##diag(dsc_ctx_t c, int32_t m, int32_t n, dsc_buf_t A, dsc_buf_t* out) { DSC_NVTX(c, "Diagonal_M"); return diag_impl<T>(c, m, n, A, out); }
DSC_LAYOUT_API(s, float)
DSC_LAYOUT_API(d, double)
}
