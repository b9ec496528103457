// K7: deterministic full reductions (Sum_V/Sum_M, L1Norm_V, L2Norm_V, SupNorm_V, Mul_Dot_V_V,
// MaxIndex_V, MinIndex_V — reference Backend.fs:81-88).
//
// One launch per reduction.  The element -> thread assignment and the combine tree are functions
// of n only (fixed grid, fixed per-thread stride, warp-shuffle tree, shared-memory tree over warps,
// then two ticketed levels over the per-block partials: groups of 32 blocks, then the groups, each
// added in index order by whichever block arrives last), so the result is bit-identical from run
// to run.  HBM-bound: 128-bit loads, 8 independent accumulators per thread.
#include <float.h>
#include <math.h>
#include <stdlib.h>

#include "dsc_internal.cuh"

namespace dsc {

constexpr int kRThreads = 256;
constexpr int kRUnroll = 8;
constexpr int kMaxPartials = 148 * 4;   // measured (tools/sum_tune.py, 4096^2): 592 -> 4.60 TB/s f32, 1184 -> 4.50, 2368 -> 4.30

enum { R_SUM = 0, R_ASUM = 1, R_SUMSQ = 2, R_AMAX = 3, R_DOT = 4,
       R_PRODSUM = 5,     // Sum(a .* b) with the product ROUNDED first: bit-identical to Mul_Had_M_M followed by Sum_M
       R_SUMSQ_SCALED = 6 };   // Sum((a * scale)^2): the scaled pass of L2Norm_V when the plain sum of squares over- or underflows

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }

template <int KIND, typename T>
__device__ __forceinline__ T r_load(T a, T b) {
    if constexpr (KIND == R_SUMSQ_SCALED) { const T t = a * b; return t * t; }   // b carries the scale
    if constexpr (KIND == R_SUM) return a;
    else if constexpr (KIND == R_ASUM) return a < T(0) ? -a : a;
    else if constexpr (KIND == R_SUMSQ) return a * a;
    else if constexpr (KIND == R_AMAX) return a < T(0) ? -a : a;
    else if constexpr (KIND == R_PRODSUM) return mul_rn(a, b);   // never contracted into an FMA with the accumulation
    else return a * b;
}
template <int KIND, typename T>
__device__ __forceinline__ T r_comb(T x, T y) {
    if constexpr (KIND == R_AMAX) return (x != x) ? x : ((y != y) ? y : (x > y ? x : y));  // NaN-propagating max
    else return x + y;
}

template <typename T> struct RVec;
template <> struct RVec<float> { using type = float4; static constexpr int N = 4; };
template <> struct RVec<double> { using type = double2; static constexpr int N = 2; };
__device__ __forceinline__ float rl(const float4& v, int i) { return (&v.x)[i]; }
__device__ __forceinline__ double rl(const double2& v, int i) { return (&v.x)[i]; }

template <int KIND, typename T>
__device__ __forceinline__ T block_tree(T v, T* smem) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = r_comb<KIND>(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < (kRThreads / 32) ? smem[lane] : T(0);
#pragma unroll
        for (int o = (kRThreads / 64); o > 0; o >>= 1) v = r_comb<KIND>(v, __shfl_xor_sync(0xffffffffu, v, o));
    }
    __syncthreads();
    return v;  // valid in thread 0
}

// scratch layout: [64..) T result, [256..) block partials, then group partials; tickets: ctx->tickets
// Each CTA owns a CONTIGUOUS chunk of `iters` x kRUnroll x kRThreads vectors (same finding as the
// elementwise kernels: contiguous chunks stream faster than a grid-wide stride).
template <int KIND, typename T, bool VEC>
__global__ void __launch_bounds__(kRThreads) reduce_kernel(const T* __restrict__ a, const T* __restrict__ b, size_t n, unsigned int per_cta_items,
                                                            unsigned int* tickets, T* partials, T* result, T scale) {
    __shared__ T smem[kRThreads / 32];
    T acc[kRUnroll];
#pragma unroll
    for (int u = 0; u < kRUnroll; ++u) acc[u] = T(0);
    // per_cta_items: vectors (VEC) or elements owned by one CTA, a multiple of kRThreads chosen by the host so
    // that a large input is cut into a whole number of CTAs per SM (a function of n only: deterministic)
    const unsigned int iters = (per_cta_items + kRUnroll * kRThreads - 1) / (kRUnroll * kRThreads);
    const size_t cta_begin = (size_t)blockIdx.x * per_cta_items, cta_end = cta_begin + per_cta_items;
    const size_t base = cta_begin + threadIdx.x;
    if constexpr (VEC) {
        using V = typename RVec<T>::type;
        constexpr int W = RVec<T>::N;
        const size_t nvec = n / W;
        const size_t end = cta_end < nvec ? cta_end : nvec;
        for (unsigned int it = 0; it < iters; ++it) {
            V va[kRUnroll], vb[kRUnroll];
            bool ok[kRUnroll];
#pragma unroll
            for (int u = 0; u < kRUnroll; ++u) {
                const size_t i = base + (size_t)(it * kRUnroll + u) * kRThreads;
                ok[u] = i < end;
                if (ok[u]) {
                    va[u] = *reinterpret_cast<const V*>(a + i * W);
                    if constexpr (KIND == R_DOT || KIND == R_PRODSUM) vb[u] = *reinterpret_cast<const V*>(b + i * W);
                    else if constexpr (KIND == R_SUMSQ_SCALED) {
#pragma unroll
                        for (int l = 0; l < W; ++l) (&vb[u].x)[l] = scale;
                    } else vb[u] = va[u];
                }
            }
#pragma unroll
            for (int u = 0; u < kRUnroll; ++u)
                if (ok[u]) {
#pragma unroll
                    for (int l = 0; l < W; ++l) acc[u] = r_comb<KIND>(acc[u], r_load<KIND>(rl(va[u], l), rl(vb[u], l)));
                }
        }
        if (blockIdx.x == 0) {  // scalar tail (< W elements)
            const size_t j = nvec * W + threadIdx.x;
            if (j < n) acc[1] = r_comb<KIND>(acc[1], r_load<KIND>(a[j], (KIND == R_DOT || KIND == R_PRODSUM) ? b[j] : (KIND == R_SUMSQ_SCALED ? scale : a[j])));
        }
    } else {
        const size_t end = cta_end < n ? cta_end : n;
        for (unsigned int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < kRUnroll; ++u) {
                const size_t j = base + (size_t)(it * kRUnroll + u) * kRThreads;
                if (j < end) acc[u] = r_comb<KIND>(acc[u], r_load<KIND>(a[j], (KIND == R_DOT || KIND == R_PRODSUM) ? b[j] : (KIND == R_SUMSQ_SCALED ? scale : a[j])));
            }
        }
    }
    T v = r_comb<KIND>(r_comb<KIND>(r_comb<KIND>(acc[0], acc[1]), r_comb<KIND>(acc[2], acc[3])),
                       r_comb<KIND>(r_comb<KIND>(acc[4], acc[5]), r_comb<KIND>(acc[6], acc[7])));
    v = block_tree<KIND>(v, smem);
    // Only warp 0 takes part in the grid-level hand-off: the other warps retire at once, so the round
    // trip of the ticket atomic does not hold the CTA's slots (it cost ~9 us per launch when it did).
    if (threadIdx.x >= 32) return;
    if (gridDim.x == 1) {
        if (threadIdx.x == 0) *result = v;
        return;
    }
    // level 1: groups of 32 blocks; the last block of a group adds the group's partials (one per lane)
    constexpr unsigned G = 32;
    const unsigned ngroups = (gridDim.x + G - 1) / G;
    const unsigned g = blockIdx.x / G, g_first = g * G;
    const unsigned g_size = min(G, gridDim.x - g_first);
    unsigned int t = 0;
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = v;
        __threadfence();
        t = atomicAdd(&tickets[1 + g], 1u);
    }
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t != g_size - 1) return;
    __threadfence();
    T w = threadIdx.x < g_size ? __ldcg(partials + g_first + threadIdx.x) : T(0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w = r_comb<KIND>(w, __shfl_xor_sync(0xffffffffu, w, o));
    if (ngroups == 1) {
        if (threadIdx.x == 0) {
            *result = w;
            tickets[1 + g] = 0;
        }
        return;
    }
    // level 2: the last group adds the group sums (lane-strided in index order, then the shuffle tree)
    T* partials2 = partials + kMaxPartials;
    if (threadIdx.x == 0) {
        partials2[g] = w;
        tickets[1 + g] = 0;
        __threadfence();
        t = atomicAdd(&tickets[0], 1u);
    }
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t != ngroups - 1) return;
    __threadfence();
    w = T(0);
    for (unsigned int i = threadIdx.x; i < ngroups; i += 32) w = r_comb<KIND>(w, __ldcg(partials2 + i));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w = r_comb<KIND>(w, __shfl_xor_sync(0xffffffffu, w, o));
    if (threadIdx.x == 0) {
        *result = w;
        tickets[0] = 0;
    }
}

struct ArgPair {
    double v;  // wide enough for both dtypes (exact for float)
    int idx;
};

// better(x, y): does candidate x beat incumbent y?  first occurrence of the extreme value wins,
// as in the reference's host twin (strict > / <, AD.Float32.fs:1629-1650).
template <bool MAXI>
__device__ __forceinline__ bool better(double xv, int xi, double yv, int yi) {
    if (MAXI) return (xv > yv) || (xv == yv && xi < yi);
    return (xv < yv) || (xv == yv && xi < yi);
}

template <typename T, bool MAXI>
__global__ void __launch_bounds__(kRThreads) argext_kernel(const T* __restrict__ a, int n, unsigned int* ticket,
                                                            ArgPair* partials, int* result) {
    __shared__ double sv[kRThreads / 32];
    __shared__ int si[kRThreads / 32];
    __shared__ bool is_last;
    const double init = MAXI ? -INFINITY : INFINITY;
    double bv = init;
    int bi = 0x7fffffff;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        double v = (double)a[j];
        if (better<MAXI>(v, j, bv, bi)) { bv = v; bi = j; }
    }
    auto tree = [&](double& v, int& i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, v, o);
            int oi = __shfl_xor_sync(0xffffffffu, i, o);
            if (better<MAXI>(ov, oi, v, i)) { v = ov; i = oi; }
        }
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) { sv[warp] = v; si[warp] = i; }
        __syncthreads();
        if (warp == 0) {
            v = lane < (kRThreads / 32) ? sv[lane] : init;
            i = lane < (kRThreads / 32) ? si[lane] : 0x7fffffff;
#pragma unroll
            for (int o = (kRThreads / 64); o > 0; o >>= 1) {
                double ov = __shfl_xor_sync(0xffffffffu, v, o);
                int oi = __shfl_xor_sync(0xffffffffu, i, o);
                if (better<MAXI>(ov, oi, v, i)) { v = ov; i = oi; }
            }
        }
        __syncthreads();
    };
    tree(bv, bi);
    if (threadIdx.x == 0) {
        partials[blockIdx.x].v = bv;
        partials[blockIdx.x].idx = bi;
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    bv = init;
    bi = 0x7fffffff;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x)
        if (better<MAXI>(partials[i].v, partials[i].idx, bv, bi)) { bv = partials[i].v; bi = partials[i].idx; }
    tree(bv, bi);
    if (threadIdx.x == 0) {
        // sequential twin: if every comparison is false (a[0] is NaN, or nothing beats a[0]) the answer is 0
        T a0 = a[0];
        if (a0 != a0 || bi == 0x7fffffff) bi = 0;
        *result = bi;
        *ticket = 0;
    }
}

template <typename T>
__global__ void sqrt1_kernel(T* p) { *p = sqrt(*p); }

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Launch a reduction leaving the scalar at `result` (device pointer).
template <int KIND, typename T>
static int launch_reduce(Ctx* ctx, const T* a, const T* b, size_t n, T* result, T scale = T(1)) {
    unsigned int* ticket = ctx->tickets;
    T* partials = reinterpret_cast<T*>(reinterpret_cast<char*>(ctx->red_scratch) + 256);
    bool vec = al16(a) && ((KIND != R_DOT && KIND != R_PRODSUM) || al16(b));
    size_t items = vec ? n / RVec<T>::N : n;
    if (items == 0) items = 1;
    // grid and per-CTA chunk are functions of n only (determinism)
    // Small inputs: one CTA per 2048 items (8 per thread).  Large inputs: kMaxPartials = 4 CTAs per SM, the
    // items cut evenly between them (rounded up to the block size), so every SM streams the same share.
    const size_t per_iter = (size_t)kRThreads * kRUnroll;
    static const int max_partials = [] { const char* e = getenv("DSC_RED_MAXP"); int v = e ? atoi(e) : 0; return (v >= 1 && v <= kMaxPartials) ? v : kMaxPartials; }();
    size_t per_cta = per_iter;
    if ((items + per_iter - 1) / per_iter > (size_t)max_partials) {
        per_cta = (items + max_partials - 1) / max_partials;
        per_cta = (per_cta + kRThreads - 1) / kRThreads * kRThreads;
    }
    if (per_cta > 0xffffff00u) return fail(DSC_E_SHAPE, "reduction: input too large");
    const int grid = (int)((items + per_cta - 1) / per_cta);
    if (vec) reduce_kernel<KIND, T, true><<<grid, kRThreads, 0, ctx->stream>>>(a, b, n, (unsigned int)per_cta, ticket, partials, result, scale);
    else reduce_kernel<KIND, T, false><<<grid, kRThreads, 0, ctx->stream>>>(a, b, n, (unsigned int)per_cta, ticket, partials, result, scale);
    DSC_LAUNCH_CHECK(ctx);
    return DSC_OK;
}

template <typename T>
static T* result_slot(Ctx* ctx) { return reinterpret_cast<T*>(reinterpret_cast<char*>(ctx->red_scratch) + 64); }

template <typename T>
static int fetch_scalar(Ctx* ctx, const void* dev, T* out) {
    if (ctx->capturing)
        return fail(DSC_E_INVALID, "host-scalar reduction while capturing a graph: use the *_dev variant");
    DSC_TRY(lane_join(ctx));   // the host is about to wait anyway: the second compute lane is drained with the stream (and releases the blocks it holds)
    DSC_CUDA_CHECK(cudaMemcpyAsync(ctx->red_host, dev, sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    DSC_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    *out = *reinterpret_cast<T*>(ctx->red_host);
    ctx->stats.memcpy_d2h_bytes += sizeof(T);
    return DSC_OK;
}

// Sum of a DEFERRED Hadamard product (`had(E, E).Sum()`, the squared-error loss: AD.Float32.fs:2393 -> :2807): reduce the
// rounded products in one pass over the factors instead of writing the product first.  Returns DSC_OK with the factor
// pointers, 1 if x is not such a value.
template <typename T>
static int sum_of_product(Ctx* ctx, Buf* x, const T** pa, const T** pb) {
    LazyNode* h = lazy_of(x, LZ_HAD);
    if (!h || !ctx->fuse) return 1;
    for (int i = 0; i < 2; ++i) DSC_TRY(operand_resolve(ctx, h->src[i]));
    auto ptr = [](const Operand& o) {
        return o.node ? reinterpret_cast<const T*>(o.node->result->ptr) + o.node->result_off : reinterpret_cast<const T*>(o.block->ptr) + o.off;
    };
    *pa = ptr(h->src[0]);
    *pb = ptr(h->src[1]);
    // the element -> thread assignment (hence the summation order) depends on the vector path: fuse only when the factors take
    // the path the freshly allocated product would have taken
    if ((reinterpret_cast<uintptr_t>(*pa) & 15) || (reinterpret_cast<uintptr_t>(*pb) & 15)) return 1;
    h->consumed = true;
    return DSC_OK;
}

template <int KIND, typename T>
static int reduce_scalar(dsc_ctx_t c, dsc_buf_t xh, dsc_buf_t yh, T* out, const char* name) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    if (!out) return fail(DSC_E_INVALID, "%s: null result pointer", name);
    Ctx* ctx = as_ctx(c);
    Buf *x, *y = nullptr;
    if (KIND == R_SUM) {
        DSC_TRY(in_buf_peek(ctx, dtype_of<T>::value, xh, &x, name));
        if (x->len == 0) { *out = T(0); return DSC_OK; }
        const T *pa, *pb;
        int s = sum_of_product<T>(ctx, x, &pa, &pb);
        if (s < 0) return s;
        if (s == DSC_OK) {
            T* slot = result_slot<T>(ctx);
            DSC_TRY((launch_reduce<R_PRODSUM, T>(ctx, pa, pb, x->len, slot)));
            return fetch_scalar<T>(ctx, slot, out);
        }
    }
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, xh, &x, name));
    if (KIND == R_DOT) {
        DSC_TRY(in_buf(ctx, dtype_of<T>::value, yh, &y, name));
        if (x->len != y->len) return fail(DSC_E_SHAPE, "%s: lengths %d and %d", name, x->len, y->len);
    }
    if (x->len == 0) { *out = T(0); return DSC_OK; }
    T* slot = result_slot<T>(ctx);
    DSC_TRY((launch_reduce<KIND, T>(ctx, x->ptr<T>(), y ? y->ptr<T>() : nullptr, x->len, slot)));
    DSC_TRY(fetch_scalar<T>(ctx, slot, out));
    return DSC_OK;
}

template <int KIND, typename T>
static int reduce_dev(dsc_ctx_t c, dsc_buf_t xh, dsc_buf_t yh, dsc_buf_t* outh, const char* name) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *x, *y = nullptr, *o;
    if (KIND == R_SUM) {
        DSC_TRY(in_buf_peek(ctx, dtype_of<T>::value, xh, &x, name));
        const T *pa, *pb;
        int s = x->len > 0 ? sum_of_product<T>(ctx, x, &pa, &pb) : 1;
        if (s < 0) return s;
        if (s == DSC_OK) {
            const bool fresh = *outh == 0;
            DSC_TRY(out_buf(ctx, dtype_of<T>::value, 1, outh, &o));
            // had(E, E).Sum() left on the device is a sink (the loss of a captured step: nobody reads it before the step
            // ends): it is taken on the second compute lane, beside the reverse sweep that starts from E
            LazyNode* h = lazy_of(x, LZ_HAD);
            Block* rd[2] = {nullptr, nullptr};
            if (h) {
                for (int i = 0; i < 2; ++i) rd[i] = h->src[i].node ? h->src[i].node->result : h->src[i].block;
            }
            if (fresh && ctx->lanes_enabled && !ctx->on_lane && rd[0] && rd[1] && x->len >= 4096 && lane_enter(ctx, rd, 2, false) == DSC_OK) {
                const int st = launch_reduce<R_PRODSUM, T>(ctx, pa, pb, x->len, o->ptr<T>());
                Block* wr[1] = {o->block};
                lane_leave(ctx, rd, 2, wr, 1);
                return st;
            }
            return launch_reduce<R_PRODSUM, T>(ctx, pa, pb, x->len, o->ptr<T>());
        }
    }
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, xh, &x, name));
    if (KIND == R_DOT) {
        DSC_TRY(in_buf(ctx, dtype_of<T>::value, yh, &y, name));
        if (x->len != y->len) return fail(DSC_E_SHAPE, "%s: lengths %d and %d", name, x->len, y->len);
    }
    DSC_TRY(out_buf(ctx, dtype_of<T>::value, 1, outh, &o));
    if (x->len == 0) {
        DSC_CUDA_CHECK(cudaMemsetAsync(o->ptr<T>(), 0, sizeof(T), ctx->stream));
        return DSC_OK;
    }
    return launch_reduce<KIND, T>(ctx, x->ptr<T>(), y ? y->ptr<T>() : nullptr, x->len, o->ptr<T>());
}

template <typename T, bool MAXI>
static int argext(dsc_ctx_t c, dsc_buf_t xh, int32_t* out, const char* name) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    if (!out) return fail(DSC_E_INVALID, "%s: null result pointer", name);
    Ctx* ctx = as_ctx(c);
    Buf* x;
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, xh, &x, name));
    if (x->len == 0) return fail(DSC_E_SHAPE, "%s of an empty buffer", name);
    unsigned int* ticket = reinterpret_cast<unsigned int*>(ctx->red_scratch);
    int* slot = reinterpret_cast<int*>(reinterpret_cast<char*>(ctx->red_scratch) + 64);
    ArgPair* partials = reinterpret_cast<ArgPair*>(reinterpret_cast<char*>(ctx->red_scratch) + 256);
    int grid = grid_for(x->len, kRThreads * 8, kMaxPartials);
    argext_kernel<T, MAXI><<<grid, kRThreads, 0, ctx->stream>>>(x->ptr<T>(), x->len, ticket, partials, slot);
    DSC_LAUNCH_CHECK(ctx);
    return fetch_scalar<int32_t>(ctx, slot, out);
}

// L2Norm_V (Backend.fs:83; the reference deployment calls ?nrm2, which scales): sqrt(sum x^2) in one pass; when that sum
// overflowed, or is so small that squares were flushed (the result would be 0 / inf / inaccurate where ?nrm2 is not), a
// second, scaled pass: s = max|x|, s * sqrt(sum (x / s)^2).  The host decides — the call returns a host scalar anyway.
template <typename T>
static int nrm2_impl(dsc_ctx_t c, dsc_buf_t xh, T* out) {
    T ss = T(0);
    DSC_TRY((reduce_scalar<R_SUMSQ, T>(c, xh, 0, &ss, "L2Norm_V")));
    const T tiny = sizeof(T) == 4 ? T(1e-30) : T(1e-290);
    if (ss == ss && ss < T(INFINITY) && ss > tiny) {
        *out = sizeof(T) == 4 ? (T)sqrtf((float)ss) : (T)sqrt((double)ss);
        return DSC_OK;
    }
    T amax = T(0);
    DSC_TRY((reduce_scalar<R_AMAX, T>(c, xh, 0, &amax, "L2Norm_V")));
    if (!(amax > T(0)) || !(amax < T(INFINITY))) {   // all zeros, NaN or an infinite element: ?nrm2 returns 0 / NaN / inf
        *out = amax == amax ? amax : amax;
        return DSC_OK;
    }
    Ctx* ctx = as_ctx(c);
    Buf* x;
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, xh, &x, "L2Norm_V"));
    T* slot = result_slot<T>(ctx);
    DSC_TRY((launch_reduce<R_SUMSQ_SCALED, T>(ctx, x->ptr<T>(), nullptr, x->len, slot, T(1) / amax)));
    T scaled = T(0);
    DSC_TRY(fetch_scalar<T>(ctx, slot, &scaled));
    *out = amax * (sizeof(T) == 4 ? (T)sqrtf((float)scaled) : (T)sqrt((double)scaled));
    return DSC_OK;
}

}  // namespace dsc

using namespace dsc;

extern "C" {
#define DSC_RED_API(P, T)                                                                                           \
    int dsc_##P##dot(dsc_ctx_t c, dsc_buf_t x, dsc_buf_t y, T* out) { DSC_NVTX(c, "Mul_Dot_V_V"); return reduce_scalar<R_DOT, T>(c, x, y, out, "Mul_Dot_V_V"); } \
    int dsc_##P##asum(dsc_ctx_t c, dsc_buf_t x, T* out) { DSC_NVTX(c, "L1Norm_V"); return reduce_scalar<R_ASUM, T>(c, x, 0, out, "L1Norm_V"); }  \
    int dsc_##P##nrm2(dsc_ctx_t c, dsc_buf_t x, T* out) { DSC_NVTX(c, "L2Norm_V"); return nrm2_impl<T>(c, x, out); }                          \
    int dsc_##P##amax(dsc_ctx_t c, dsc_buf_t x, T* out) { DSC_NVTX(c, "SupNorm_V"); return reduce_scalar<R_AMAX, T>(c, x, 0, out, "SupNorm_V"); } \
    int dsc_##P##sum(dsc_ctx_t c, dsc_buf_t x, T* out) { DSC_NVTX(c, "Sum_V/Sum_M"); return reduce_scalar<R_SUM, T>(c, x, 0, out, "Sum_V"); }      \
    int dsc_##P##argmax(dsc_ctx_t c, dsc_buf_t x, int32_t* out) { DSC_NVTX(c, "MaxIndex_V"); return argext<T, true>(c, x, out, "MaxIndex_V"); }   \
    int dsc_##P##argmin(dsc_ctx_t c, dsc_buf_t x, int32_t* out) { DSC_NVTX(c, "MinIndex_V"); return argext<T, false>(c, x, out, "MinIndex_V"); }  \
    int dsc_##P##sum_dev(dsc_ctx_t c, dsc_buf_t x, dsc_buf_t* out) { DSC_NVTX(c, "Sum_V/Sum_M (device scalar)"); return reduce_dev<R_SUM, T>(c, x, 0, out, "Sum_V"); } \
    int dsc_##P##dot_dev(dsc_ctx_t c, dsc_buf_t x, dsc_buf_t y, dsc_buf_t* out) { DSC_NVTX(c, "Mul_Dot_V_V (device scalar)"); return reduce_dev<R_DOT, T>(c, x, y, out, "Mul_Dot_V_V"); }
DSC_RED_API(s, float)
DSC_RED_API(d, double)
}
