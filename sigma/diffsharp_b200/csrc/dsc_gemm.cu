// K1/K2 dispatch + the CUDA-core GEMM used for shapes the tensor-core kernels do not take
// (tiny / skinny / unaligned operands) — Mul_M_M and Mul_M_M_Add_V_MCols, reference
// Backend.fs:119,121; call site AD.Float32.fs:2382; adjoint call sites :4127-4143.
//
//   f32: tcgen05/TMEM/TMA 3xTF32 kernel (dsc_gemm_tf32x3.cu) when eligible, else this kernel.
//   f64: DMMA kernel (dsc_gemm_dmma.cu) when eligible, else this kernel.
// There is no host fallback: every path is a device kernel of this library.
#include <stdlib.h>

#include <algorithm>

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

// The epilogue as a second pass over C, for the kernels that do not fuse it (CUDA-core, skinny, DMMA).
template <typename T>
static int epilogue_pass(Ctx* ctx, int m, int n, T* C, const GemmEpilogue* ep) {
    if (!ep || ep->mode == GEMM_EP_NONE) return DSC_OK;
    T* lo = reinterpret_cast<T*>(ep->lo);
    if (ep->mode == GEMM_EP_BIAS_ACT)
        return ew_rows<T>(ctx, DSC_OP_ADD, ep->op, m, n, C, reinterpret_cast<const T*>(ep->vec), C, reinterpret_cast<T*>(ep->out2), lo);
    return ew_fused_bwd<T>(ctx, ep->op, C, reinterpret_cast<const T*>(ep->aux), C, lo, (size_t)m * n);
}

template <>
int gemm_run<float>(Ctx* ctx, int ta, int tb, int m, int n, int k, MatRef A, MatRef B, const float* bias, float* C, const GemmEpilogue* ep) {
    const float *Ap = reinterpret_cast<const float*>(A.ptr), *Bp = reinterpret_cast<const float*>(B.ptr);
    {
        // the output layer: bias rows without activation ride along in the skinny kernel
        const bool colb = ep && ep->mode == GEMM_EP_BIAS_ACT && ep->op < 0 && !ep->lo;
        int done = 0;
        int s = gemm_skinny<float>(ctx, ta, tb, m, n, k, Ap, Bp, bias, C, colb ? reinterpret_cast<const float*>(ep->vec) : nullptr, &done);
        if (s < 0) return s;
        if (s == 0) return done ? DSC_OK : epilogue_pass<float>(ctx, m, n, C, ep);
    }
    if (!(ctx->flags & DSC_FLAG_GEMM_FP32_SIMT)) {
        int s = gemm_tf32x3_tcgen05(ctx, ta, tb, m, n, k, A, B, bias, C, ep);
        if (s <= 0) return s;  // handled, epilogue included (0) or failed (<0); 1 = not eligible; 2 = product stored, epilogue still to run
        if (s == 2) return epilogue_pass<float>(ctx, m, n, C, ep);
    }
    DSC_TRY(gemm_simt<float>(ctx, ta, tb, m, n, k, Ap, Bp, bias, C));
    return epilogue_pass<float>(ctx, m, n, C, ep);
}

template <>
int gemm_run<double>(Ctx* ctx, int ta, int tb, int m, int n, int k, MatRef A, MatRef B, const double* bias, double* C, const GemmEpilogue* ep) {
    const double *Ap = reinterpret_cast<const double*>(A.ptr), *Bp = reinterpret_cast<const double*>(B.ptr);
    {
        // the output layer: bias rows without activation ride along in the skinny kernel
        const bool colb = ep && ep->mode == GEMM_EP_BIAS_ACT && ep->op < 0 && !ep->lo;
        int done = 0;
        int s = gemm_skinny<double>(ctx, ta, tb, m, n, k, Ap, Bp, bias, C, colb ? reinterpret_cast<const double*>(ep->vec) : nullptr, &done);
        if (s < 0) return s;
        if (s == 0) return done ? DSC_OK : epilogue_pass<double>(ctx, m, n, C, ep);
    }
    if (!(ctx->flags & DSC_FLAG_GEMM_F64_SIMT)) {
        int s = gemm_f64_dmma(ctx, ta, tb, m, n, k, Ap, Bp, bias, C);
        if (s < 0) return s;
        if (s == 0) return epilogue_pass<double>(ctx, m, n, C, ep);
    }
    DSC_TRY(gemm_simt<double>(ctx, ta, tb, m, n, k, Ap, Bp, bias, C));
    return epilogue_pass<double>(ctx, m, n, C, ep);
}

static int try_sink_product(Ctx* ctx, LazyNode* nd);

template <typename T>
static int gemm_entry(dsc_ctx_t c, int ta, int tb, int32_t m, int32_t n, int32_t k, dsc_buf_t Ah, dsc_buf_t Bh, dsc_buf_t biash, dsc_buf_t* Ch) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *A, *B, *bias = nullptr, *C;
    DSC_TRY(in_buf_peek(ctx, dtype_of<T>::value, Ah, &A, "Mul_M_M: A"));
    DSC_TRY(in_buf_peek(ctx, dtype_of<T>::value, Bh, &B, "Mul_M_M: B"));
    if (biash) DSC_TRY(in_buf(ctx, dtype_of<T>::value, biash, &bias, "Mul_M_M_Add_V_MCols: v"));
    if (m < 0 || n < 0 || k < 0) return fail(DSC_E_SHAPE, "Mul_M_M: negative dimension");
    if ((int64_t)m * k > A->len) return fail(DSC_E_SHAPE, "Mul_M_M: A is %d x %d but its buffer holds %d elements", m, k, A->len);
    if ((int64_t)k * n > B->len) return fail(DSC_E_SHAPE, "Mul_M_M: B is %d x %d but its buffer holds %d elements", k, n, B->len);
    if ((int64_t)m * n > 0x7fffffff) return fail(DSC_E_SHAPE, "Mul_M_M: result too large for an int32 length");
    if (bias && bias->len != m) return fail(DSC_E_SHAPE, "Mul_M_M_Add_V_MCols: vector length %d, result has %d rows", bias->len, m);
    // Transpose_M followed by Mul_M_M (every adjoint of a product, AD.Float32.fs:4127-4143): read the stored matrix through a
    // transposed tile map instead of running the transpose.  op(A) is m x k: a pending transpose of a stored k x m matrix.
    Operand oa, ob;
    auto take = [&](Buf* X, int& t, int rows, int cols, Operand& o) {
        LazyNode* tn = lazy_of(X, LZ_T);   // X = transpose of src (tn->m x tn->n), i.e. X is tn->n x tn->m
        if (tn && tn->n == rows && tn->m == cols && (int64_t)rows * cols == X->len) {
            o = tn->src[0];
            t ^= 1;
            tn->consumed = true;
            return;
        }
        o = Operand();
        o.len = X->len;
        if (X->block) { o.block = X->block; o.off = X->off; }
        else o.node = X->lazy;
    };
    int eta = ta ? 1 : 0, etb = tb ? 1 : 0;
    take(A, eta, ta ? k : m, ta ? m : k, oa);
    take(B, etb, tb ? n : k, tb ? k : n, ob);
    if (!Ch) return fail(DSC_E_INVALID, "null output handle pointer");
    if constexpr (sizeof(T) == 8) {
        // rows(x) W = rows(x W): the forward-mode sweep with a matrix tangent carries a primal made of m identical rows
        // (DNDArray.OfRows(m, x), AD.Float64.fs:1846) and multiplies it by every weight matrix — half of the
        // issued flops.  The DMMA kernel accumulates every output element over k in one fixed order whatever the number of rows,
        // so ONE row through the same kernel gives bit for bit the row the full product would repeat m times.
        LazyNode* r = (ctx->fuse && *Ch == 0 && !bias && eta == 0) ? oa.node : nullptr;
        if (r && r->kind == LZ_ROWS && !r->result && r->m == m && r->n == k && m > 1 && n > 16 && k > 16 && (int64_t)m * n >= 64 * 64 && (int64_t)n >= 64 * 64) {
            r->refs++;   // keep the node while its vector and B are resolved
            Operand xv = r->src[0];
            LazyNode* hold = lazy_new(ctx, LZ_GEMM, DSC_F64, n);
            hold->consumed = true;
            lazy_set_operand(hold, 0, xv);
            lazy_set_operand(hold, 1, ob);
            int st = DSC_OK;
            for (int i = 0; i < 2 && st == DSC_OK; ++i) st = operand_resolve(ctx, hold->src[i]);
            Buf* y = nullptr;
            dsc_buf_t yh = 0;
            if (st == DSC_OK) st = out_buf(ctx, DSC_F64, n, &yh, &y);
            int handled = 1;
            if (st == DSC_OK) {
                auto ptr = [](const Operand& o) { return o.node ? reinterpret_cast<const double*>(o.node->result->ptr) + o.node->result_off : reinterpret_cast<const double*>(o.block->ptr) + o.off; };
                const double* Bp = ptr(hold->src[1]);
                // the full product would have to be DMMA-eligible too (same kernel family, same accumulation order)
                const int ldb = etb ? k : n;
                const bool full_ok = !(ctx->flags & DSC_FLAG_GEMM_F64_SIMT) && !(k & 1) && !(ldb & 1) && !(reinterpret_cast<uintptr_t>(Bp) & 15);
                handled = full_ok ? gemm_f64_dmma(ctx, 0, etb, 1, n, k, ptr(hold->src[0]), Bp, nullptr, y->ptr<double>()) : 1;
                if (handled < 0) st = handled;
            }
            lazy_unref(hold);
            if (st == DSC_OK && handled == 0) {
                LazyNode* nd = lazy_new(ctx, LZ_ROWS, DSC_F64, m * n);
                nd->m = m; nd->n = n;
                lazy_set_operand(nd, 0, y);
                r->consumed = true;
                lazy_unref(r);
                dsc_buf_free(yh);
                return lazy_buf(ctx, nd, Ch);
            }
            lazy_unref(r);
            if (yh) dsc_buf_free(yh);
            if (st != DSC_OK) return st;
            // not eligible: fall through to the general path (the replicated rows are materialised)
        }
    }
    if (ctx->fuse && *Ch == 0 && (int64_t)m * n > 0 && k > 0) {
        // The product is DEFERRED: the calls that follow it in the unchanged sequence — Add_M_M(., rows(b)) -> Map_F_M in the
        // forward pass (AD.Float32.fs:2357, :2704), the activation's adjoint chain in the reverse sweep (:4238-4240,
        // :4280-4281) — become its epilogue; any other consumer, or the next synchronisation point, runs the plain product.
        LazyNode* nd = lazy_new(ctx, LZ_GEMM, dtype_of<T>::value, m * n);
        nd->m = m; nd->n = n; nd->k = k; nd->ta = eta; nd->tb = etb;
        lazy_set_operand(nd, 0, oa);
        lazy_set_operand(nd, 1, ob);
        if (bias) lazy_set_operand(nd, 2, bias);
        if (!bias && eta != (ta ? 1 : 0)) {   // A was a pending transpose: a weight adjoint — a sink of the reverse sweep
            nd->refs++;
            nd->sink = true;
            const int ss = try_sink_product(ctx, nd);
            nd->refs--;
            if (ss < 0) {
                lazy_unref(nd);
                return ss;
            }
        }
        return lazy_buf(ctx, nd, Ch);
    }
    // hold the operands while they are being resolved (a deferred operand may be dropped by its handle meanwhile)
    LazyNode* hold = lazy_new(ctx, LZ_GEMM, dtype_of<T>::value, m * n);
    hold->consumed = true;
    lazy_set_operand(hold, 0, oa);
    lazy_set_operand(hold, 1, ob);
    int st = DSC_OK;
    for (int i = 0; i < 2 && st == DSC_OK; ++i) st = operand_resolve(ctx, hold->src[i]);
    if (st == DSC_OK) st = out_buf(ctx, dtype_of<T>::value, m * n, Ch, &C);
    if (st == DSC_OK && m * n > 0) {
        auto mat = [](const Operand& o) {
            MatRef r;
            r.block = o.node ? o.node->result : o.block;
            r.off = o.node ? o.node->result_off : o.off;
            r.ptr = r.block ? reinterpret_cast<const T*>(r.block->ptr) + r.off : nullptr;
            return r;
        };
        MatRef ma = mat(hold->src[0]), mb = mat(hold->src[1]);
        if (C->ptr<T>() == ma.ptr || C->ptr<T>() == mb.ptr) st = fail(DSC_E_INVALID, "Mul_M_M: output aliases an input");
        else if (k == 0) {
            int grid = grid_for((size_t)m * n, 256 * 4, ctx->sm_count * 8);
            bias_rows_kernel<T><<<grid, 256, 0, ctx->stream>>>(bias ? bias->ptr<T>() : nullptr, C->ptr<T>(), m, n);
            ctx->stats.kernel_launches++;
        } else {
            st = gemm_run<T>(ctx, eta, etb, m, n, k, ma, mb, bias ? bias->ptr<T>() : nullptr, C->ptr<T>(), nullptr);
        }
    }
    lazy_unref(hold);
    return st;
}

// ---- two compute lanes for small pending products -------------------------------------------------------------------------
static bool co_schedulable(Ctx* ctx, const LazyNode* nd) {
    if (nd->kind != LZ_GEMM || nd->dtype != DSC_F32 || nd->result || nd->consumed || nd->refs <= 0 || nd->nsrc != 2) return false;
    if (nd->m < 128 || nd->n < 32 || nd->k < 32 || nd->n <= 16 || nd->k <= 16) return false;                 // tensor-core shapes only
    if (((nd->m + 127) / 128) * ((nd->n + 127) / 128) > ctx->sm_count / 2) return false;                  // it would fill the GPU by itself
    for (int i = 0; i < 2; ++i)
        if (nd->src[i].node || !nd->src[i].block) return false;                                          // operands are plain device windows
    return true;
}

// Run the deferred product `nd` on the second compute lane (operands: plain device windows).  sm_budget > 0 limits the SMs its
// planner may use.
static int run_on_lane(Ctx* ctx, LazyNode* nd, int sm_budget) {
    // what it reads: the operand blocks and their operand-split companions; an operand that has no (valid) companion yet gets
    // one from this product's pre-pass — written on the lane: compute-stream products that meet it later wait for the lane
    // (lane_gate_read in the GEMM)
    Block* ops[3];
    bool had_lo[3];
    int nops = 0;
    Block* reads[8];
    int nr = 0;
    for (int i = 0; i < nd->nsrc && i < 3; ++i) {
        Block* b = nd->src[i].node ? nd->src[i].node->result : nd->src[i].block;
        if (!b) continue;
        ops[nops] = b;
        had_lo[nops] = b->lo && b->lo_epoch == ctx->cache_epoch;
        reads[nr++] = b;
        if (had_lo[nops]) reads[nr++] = b->lo;
        b->refs++;   // materialising releases the node's own operand references
        nops++;
    }
    int st = lane_enter(ctx, reads, nr, false);
    Block* writes[4];
    int nw = 0;
    if (st == DSC_OK) {
        const int saved = ctx->gemm_sm_budget;
        ctx->gemm_sm_budget = sm_budget;
        st = lazy_materialize(nd);
        ctx->gemm_sm_budget = saved;
        if (st == DSC_OK && nd->result) writes[nw++] = nd->result;
        for (int i = 0; i < nops; ++i)
            if (!had_lo[i] && ops[i]->lo && ops[i]->lo_epoch == ctx->cache_epoch) writes[nw++] = ops[i]->lo;
        lane_leave(ctx, reads, nr, writes, nw);
    }
    for (int i = 0; i < nops; ++i) block_release(ops[i]);
    return st;
}

int co_schedule_pending_gemms(Ctx* ctx) {
    static const bool off = [] { const char* e = getenv("DSC_GEMM_LANES"); return e && *e == '0'; }();
    if (off || !ctx->fuse || !ctx->lanes_enabled || ctx->on_lane || (ctx->flags & DSC_FLAG_GEMM_FP32_SIMT)) return DSC_OK;
    std::vector<LazyNode*> todo;
    for (LazyNode* nd : ctx->pending)
        if (co_schedulable(ctx, nd)) todo.push_back(nd);
    if (todo.size() < 2) return DSC_OK;
    for (LazyNode* nd : todo) nd->refs++;
    int st = DSC_OK;
    for (size_t i = 0; i + 1 < todo.size() && st == DSC_OK; i += 2) {
        LazyNode *a = todo[i], *b = todo[i + 1];
        if (a->result || b->result) continue;
        // an operand the two products share must already have its operand-split companion: otherwise both lanes would create
        // and read it concurrently
        bool shared_unsplit = false;
        for (int x = 0; x < 2; ++x)
            for (int y = 0; y < 2; ++y)
                if (a->src[x].block == b->src[y].block && !(a->src[x].block->lo && a->src[x].block->lo_epoch == ctx->cache_epoch)) shared_unsplit = true;
        if (shared_unsplit) continue;
        st = run_on_lane(ctx, b, ctx->sm_count / 2);
        if (st == DSC_OK) {
            const int saved = ctx->gemm_sm_budget;
            ctx->gemm_sm_budget = ctx->sm_count / 2;
            st = lazy_materialize(a);
            ctx->gemm_sm_budget = saved;
        }
        if (st == DSC_OK) ctx->stats_co_scheduled++;
    }
    for (LazyNode* nd : todo) lazy_unref(nd);
    return st;
}

// A product whose A operand was a pending Transpose_M is the weight adjoint of a reverse sweep (Transpose(a.P) * d.A,
// AD.Float32.fs:4131-4143): its only consumer is the accumulation into the parameter's adjoint.  While the compute stream still
// has another product pending (the input adjoint d.A * Transpose(b.P) of the same tape node, issued just before it), it runs at
// once on the second compute lane instead of waiting for the end of the sweep: tensor-core products that are both small share
// the SMs half and half, skinny ones (the output layer) simply run beside the compute stream.
static int try_sink_product(Ctx* ctx, LazyNode* nd) {
    if (!ctx->lanes_enabled || ctx->on_lane || nd->dtype != DSC_F32 || nd->nsrc != 2) return 1;
    for (int i = 0; i < 2; ++i)
        if (nd->src[i].node || !nd->src[i].block) return 1;
    LazyNode* partner = nullptr;   // the newest other product still pending on the compute stream
    for (LazyNode* o : ctx->pending)
        if (o != nd && o->kind == LZ_GEMM && !o->result && !o->consumed && o->dtype == DSC_F32) partner = o;
    if (lane_debug()) fprintf(stderr, "[dsc lane] weight adjoint %d x %d x %d: %s\n", nd->m, nd->n, nd->k, partner ? "a product is pending on the compute stream" : "nothing pending: runs now on the compute stream");
    if (!partner) {
        // nothing to run beside: the product still is a sink — no consumer will fold into its epilogue — so it starts now
        // instead of at the end of the sweep (behind the column sums the calls that follow it issue)
        // ... and the lane is forked first: the sinks the calls right behind it issue on the same inputs (the bias column
        // sums of the layer) then run beside this product, not behind it
        Block* rd[2] = {nd->src[0].block, nd->src[1].block};
        DSC_TRY(lane_prefork(ctx, rd, 2));
        nd->refs++;
        const int st = lazy_materialize(nd);
        lazy_unref(nd);
        return st;
    }
    const bool skinny = nd->n <= 16 || nd->k <= 16 || nd->m <= 16;
    if (skinny) {
        if ((int64_t)nd->m * nd->n > (1 << 20)) return 1;
        return run_on_lane(ctx, nd, 0);
    }
    if (ctx->flags & DSC_FLAG_GEMM_FP32_SIMT) return 1;
    auto small = [&](const LazyNode* x) { return x->m >= 128 && x->n >= 32 && x->k >= 32 && ((x->m + 127) / 128) * ((x->n + 127) / 128) <= ctx->sm_count / 2; };
    // The weight adjoint goes ahead on the lane with half of the SMs and the partner (the input adjoint of the same tape node,
    // or an earlier weight adjoint still waiting) takes the other half — also when one of them would fill the GPU by itself:
    // measured A/B on one box (tools/step_timeline.py, DSC_SINK_ALWAYS), step at 2048 / 4096 / 8192 rows 184.3 / 292.7 / 483.6 us
    // against 190.9 / 288.2 / 490.5 us when only products that are both small share the SMs.
    if (nd->m < 128 || nd->n < 32 || nd->k < 32) return 1;
    static const bool small_only = [] { const char* e = getenv("DSC_SINK_ALWAYS"); return e && *e == '0'; }();
    if (small_only && (!small(nd) || !small(partner))) return 1;
    DSC_TRY(run_on_lane(ctx, nd, ctx->sm_count / 2));
    ctx->main_gemm_half_once = true;   // the partner takes the other half when the compute stream reaches it
    ctx->stats_co_scheduled++;
    if (partner->sink && !partner->src[0].node && !partner->src[1].node && partner->nsrc == 2) {
        // the partner is a weight adjoint as well (it found nothing to run beside when it was issued): it starts now
        partner->refs++;
        const int st = lazy_materialize(partner);
        lazy_unref(partner);
        return st;
    }
    return DSC_OK;
}

}  // namespace dsc

using namespace dsc;

extern "C" {
int dsc_sgemm(dsc_ctx_t c, int ta, int tb, int32_t m, int32_t n, int32_t k, dsc_buf_t A, dsc_buf_t B, dsc_buf_t bias, dsc_buf_t* C) {
    DSC_NVTX(c, "Mul_M_M");
    return gemm_entry<float>(c, ta, tb, m, n, k, A, B, bias, C);
}
int dsc_dgemm(dsc_ctx_t c, int ta, int tb, int32_t m, int32_t n, int32_t k, dsc_buf_t A, dsc_buf_t B, dsc_buf_t bias, dsc_buf_t* C) {
    DSC_NVTX(c, "Mul_M_M");
    return gemm_entry<double>(c, ta, tb, m, n, k, A, B, bias, C);
}
}
