// Deferred evaluation behind the unchanged Backend<'T> call sequence, inside the library so that every shim
// (the F# P/Invoke twin as much as the Python one) inherits it with ONE C-ABI call per Backend<'T> member.
//
// The unchanged AD layer issues short, fixed call chains whose intermediate results nobody else reads:
//   Transpose_M -> Mul_M_M                                   (AD.Float32.fs:4127-4143)   transposed GEMM operand, no transpose kernel
//   RepeatReshapeCopy_V_MRows -> Add_M_M [-> Map_F_M]        (:1860-1865, :2357, :2704)  one pass z = A + rows(v); h = f(z)
//   Map(Cosh) -> Map_F_S(1, Div) -> Had -> Had               (:4238-4240)                tanh adjoint, one kernel
//   Had(dA, dP) , Sub_S(1, dP) -> Had                        (:4281)                     sigmoid adjoint, one kernel
//   Map(Sign) -> Add_S(1) -> Mul_S(1/2) -> Had               (:4280)                     ReLU adjoint, one kernel
//   Had(E, E) -> Sum_M                                       (:2393, :2807)              one product-sum reduction
//   CreateZeroArray + CreateDataBuffer (two per tape node)   (:1834-1835, :3746)         nothing until somebody reads the zeros
//   CreateValueArray(n, 1) + CreateDataBuffer -> Had         (:3430, :4144)              1 .* E is E
// An entry point that may start such a chain returns a handle whose value is a LazyNode (dsc_internal.cuh).  A consumer that
// recognises the chain takes the node's operands and launches the fused kernel — which performs the SAME sequence of
// roundings as the separate kernels; any other consumer materialises the node with exactly the kernel the eager call would
// have run.  Either way the bits are those of the eager sequence (tests/test_gpu_configs.py: fused == unfused).
//
// Hazards handled here:
//   * an in-place write to a block that a deferred operation still has to read runs that operation first (Block::readers);
//   * handles that end up sharing one result are copy-on-write aliases of each other (Block::cow);
//   * deferred COMPUTE nodes that nobody consumed run at the next synchronisation point (flush_pending), so a caller that
//     times `Mul_Had_M_M` with events measures the kernel, not the bookkeeping.
#include <algorithm>

#include "dsc_internal.cuh"

namespace dsc {

static bool is_compute_kind(int kind) {
    return kind == LZ_ADD_ROWS || kind == LZ_MAP || kind == LZ_SECH || kind == LZ_HAD || kind == LZ_SCALAR || kind == LZ_GEMM || kind == LZ_COLSUM;
}

LazyNode* lazy_new(Ctx* ctx, int kind, int dtype, int32_t len) {
    LazyNode* nd = new LazyNode();
    nd->refs = 1;
    nd->kind = kind;
    nd->dtype = dtype;
    nd->ctx = ctx;
    nd->len = len;
    if (is_compute_kind(kind)) {
        nd->pending_idx = (int)ctx->pending.size();
        ctx->pending.push_back(nd);
    }
    return nd;
}

static void pending_remove(LazyNode* nd) {
    if (nd->pending_idx < 0) return;
    Ctx* ctx = nd->ctx;
    LazyNode* last = ctx->pending.back();
    ctx->pending[nd->pending_idx] = last;
    last->pending_idx = nd->pending_idx;
    ctx->pending.pop_back();
    nd->pending_idx = -1;
}

static void operand_release(LazyNode* owner, Operand& o) {
    if (o.block) {
        auto& r = o.block->readers;
        auto it = std::find(r.begin(), r.end(), owner);
        if (it != r.end()) r.erase(it);
        block_release(o.block);
    }
    if (o.node) lazy_unref(o.node);
    o = Operand();
}

void lazy_set_operand(LazyNode* nd, int i, const Operand& o) {
    Operand& d = nd->src[i];
    d = o;
    if (d.node) {
        d.node->refs++;
        d.node->consumed = true;
    } else if (d.block) {
        d.block->refs++;
        d.block->readers.push_back(nd);
    }
    if (i >= nd->nsrc) nd->nsrc = i + 1;
}

void lazy_set_operand(LazyNode* nd, int i, Buf* b) {
    Operand o;
    if (b && b->len > 0) {
        if (b->block) {
            o.block = b->block;
            o.off = b->off;
            o.len = b->len;
        } else {
            o.node = b->lazy;
            o.len = b->len;
        }
    }
    lazy_set_operand(nd, i, o);
}

void lazy_unref(LazyNode* nd) {
    if (!nd || --nd->refs > 0) return;
    for (int i = 0; i < nd->nsrc; ++i) operand_release(nd, nd->src[i]);
    if (nd->result) {
        nd->result->cow--;
        block_release(nd->result);
    }
    pending_remove(nd);
    delete nd;
}

int lazy_buf(Ctx* ctx, LazyNode* nd, dsc_buf_t* out) {
    Buf* b = new Buf();
    b->dtype = nd->dtype;
    b->ctx = ctx;
    b->len = nd->len;
    b->lazy = nd;   // takes over the caller's reference
    *out = reinterpret_cast<dsc_buf_t>(b);
    return DSC_OK;
}

void buf_adopt(Buf* b, Block* blk, size_t off, bool cow) {
    b->block = blk;
    b->off = off;
    blk->refs++;
    b->is_cow = cow;
    if (cow) blk->cow++;
}

int buf_resolve(Buf* b) {
    if (b->block) return lane_gate_read(b->ctx, b->block);   // a value the second compute lane produced is about to be used by compute-stream work
    if (b->len == 0 || !b->lazy) return DSC_OK;
    LazyNode* nd = b->lazy;
    DSC_TRY(lazy_materialize(nd));
    // other handles (or deferred operations) may still resolve to the same block: this handle then is a copy-on-write holder
    buf_adopt(b, nd->result, nd->result_off, nd->refs > 1);
    b->lazy = nullptr;
    lazy_unref(nd);
    return lane_gate_read(b->ctx, b->block);   // the value may have been produced on the second compute lane
}

int operand_resolve(Ctx* ctx, Operand& o) {
    if (o.node) {
        DSC_TRY(lazy_materialize(o.node));
        return lane_gate_read(ctx, o.node->result);
    }
    return lane_gate_read(ctx, o.block);
}

template <typename T>
static const T* operand_ptr(const Operand& o) {
    if (o.node) return reinterpret_cast<const T*>(o.node->result->ptr) + o.node->result_off;
    return o.block ? reinterpret_cast<const T*>(o.block->ptr) + o.off : nullptr;
}

template <typename T>
static MatRef operand_mat(const Operand& o) {
    MatRef r;
    r.block = o.node ? o.node->result : o.block;
    r.off = o.node ? o.node->result_off : o.off;
    r.ptr = operand_ptr<T>(o);
    return r;
}

int buf_share_into(Buf* dst, Buf* src) {
    // dst (holding nothing) becomes a copy-on-write alias of src's value
    if (src->len == 0) return DSC_OK;
    if (!src->block) {            // deferred value: both handles name the same node
        dst->lazy = src->lazy;
        dst->lazy->refs++;
        return DSC_OK;
    }
    if (src->block->views > 0 || src->is_view) return 1;   // shared MUTABLE storage: the caller copies eagerly
    if (!src->is_cow) {           // the owner becomes a copy-on-write holder as well: neither side may see the other's writes
        src->is_cow = true;
        src->block->cow++;
    }
    buf_adopt(dst, src->block, src->off, true);
    return DSC_OK;
}

// ---- materialisation -----------------------------------------------------------------------------------------------
template <typename T>
static int materialize_typed(LazyNode* nd, Block* out) {
    Ctx* ctx = nd->ctx;
    T* y = reinterpret_cast<T*>(out->ptr);
    const size_t n = (size_t)nd->len;
    if (nd->kind == LZ_ADD_ROWS && nd->src[0].node && nd->src[0].node->kind == LZ_GEMM && !nd->src[0].node->result) {
        // z = (A B) + rows(v) with the product still deferred (the output layer: its sum feeds Sub_M_M, not an activation):
        // the bias joins the product's epilogue and the bare product is never stored
        LazyNode* g = nd->src[0].node;
        for (int i = 0; i < g->nsrc; ++i) DSC_TRY(operand_resolve(ctx, g->src[i]));
        DSC_TRY(operand_resolve(ctx, nd->src[1]));
        GemmEpilogue ep;
        ep.mode = GEMM_EP_BIAS_ACT;
        ep.op = -1;
        ep.vec = operand_ptr<T>(nd->src[1]);
        g->consumed = true;
        return gemm_run<T>(ctx, g->ta, g->tb, g->m, g->n, g->k, operand_mat<T>(g->src[0]), operand_mat<T>(g->src[1]),
                           g->nsrc > 2 ? operand_ptr<T>(g->src[2]) : nullptr, y, &ep);
    }
    for (int i = 0; i < nd->nsrc; ++i) DSC_TRY(operand_resolve(ctx, nd->src[i]));
    switch (nd->kind) {
        case LZ_ZERO:
            DSC_CUDA_CHECK(cudaMemsetAsync(y, 0, n * sizeof(T), ctx->stream));
            ctx->stats.kernel_launches++;
            return DSC_OK;
        case LZ_FILL: return fill_raw(ctx, nd->dtype, y, n, nd->s);
        case LZ_T: return transpose_launch<T>(ctx, operand_ptr<T>(nd->src[0]), y, nd->m, nd->n);
        case LZ_ROWS: return repeat_rows_launch<T>(ctx, operand_ptr<T>(nd->src[0]), y, nd->m, nd->n);
        case LZ_ADD_ROWS: return ew_rows<T>(ctx, DSC_OP_ADD, -1, nd->m, nd->n, operand_ptr<T>(nd->src[0]), operand_ptr<T>(nd->src[1]), y, nullptr, nullptr);
        case LZ_MAP: return ew_map<T>(ctx, nd->op, operand_ptr<T>(nd->src[0]), y, n);
        case LZ_SECH: return ew_sech<T>(ctx, operand_ptr<T>(nd->src[0]), y, n);
        case LZ_HAD: return ew_map2<T>(ctx, DSC_OP_MUL, operand_ptr<T>(nd->src[0]), operand_ptr<T>(nd->src[1]), y, n);
        case LZ_SCALAR: return ew_scalar<T>(ctx, nd->op, (T)nd->s, operand_ptr<T>(nd->src[0]), y, n);
        case LZ_COLSUM: return colsum_launch<T>(ctx, operand_ptr<T>(nd->src[0]), y, nd->m, nd->n, false);
        case LZ_GEMM: return gemm_run<T>(ctx, nd->ta, nd->tb, nd->m, nd->n, nd->k, operand_mat<T>(nd->src[0]), operand_mat<T>(nd->src[1]),
                                         nd->nsrc > 2 ? operand_ptr<T>(nd->src[2]) : nullptr, y, nullptr);
    }
    return fail(DSC_E_INVALID, "unknown deferred operation %d", nd->kind);
}

int lazy_materialize(LazyNode* nd) {
    if (nd->result) return DSC_OK;
    Ctx* ctx = nd->ctx;
    Block* out = nullptr;
    DSC_TRY(pool_alloc(ctx, (size_t)nd->len * dtype_size(nd->dtype), &out));
    int s = nd->dtype == DSC_F32 ? materialize_typed<float>(nd, out) : materialize_typed<double>(nd, out);
    if (s != DSC_OK) {
        block_release(out);
        return s;
    }
    lazy_set_result(nd, out, 0);
    block_release(out);   // lazy_set_result took the node's own reference
    return DSC_OK;
}

void lazy_set_result(LazyNode* nd, Block* blk, size_t off) {
    nd->result = blk;
    nd->result_off = off;
    blk->refs++;
    blk->cow++;   // the node is a holder that must never observe a write
    for (int i = 0; i < nd->nsrc; ++i) operand_release(nd, nd->src[i]);
    nd->nsrc = 0;
    pending_remove(nd);
}

// Deferred column sums (see colsum_impl) run on the second compute lane when it is already ordered after their matrix.
int flush_colsums_on_lane(Ctx* ctx) {
    if (!ctx->lanes_enabled || ctx->on_lane) return DSC_OK;
    std::vector<LazyNode*> todo;
    for (LazyNode* nd : ctx->pending)
        if (nd->kind == LZ_COLSUM && !nd->result && nd->refs > 0 && !nd->src[0].node && nd->src[0].block) {
            nd->refs++;
            todo.push_back(nd);
        }
    int st = DSC_OK;
    for (LazyNode* nd : todo) {
        Block* rd[1] = {nd->src[0].block};
        if (st == DSC_OK && !nd->result && rd[0]->lane_w <= ctx->lane_joined && !lane_sees(ctx, rd, 1) && lane_prefork_covers(ctx, rd[0])) st = lane_apply_prefork(ctx);
        if (st == DSC_OK && !nd->result && rd[0]->lane_w <= ctx->lane_joined && lane_sees(ctx, rd, 1)) {
            rd[0]->refs++;   // materialising releases the node's operand reference
            st = lane_enter(ctx, rd, 1, false);
            if (st == DSC_OK) {
                st = lazy_materialize(nd);
                Block* wr[1] = {st == DSC_OK ? nd->result : nullptr};
                lane_leave(ctx, rd, 1, wr, 1);
            }
            block_release(rd[0]);
        }
        lazy_unref(nd);
    }
    return st;
}

int flush_pending(Ctx* ctx) {
    DSC_TRY(co_schedule_pending_gemms(ctx));
    DSC_TRY(flush_colsums_on_lane(ctx));
    // materialising a node removes it (and possibly others) from the list: walk a snapshot, holding the nodes
    std::vector<LazyNode*> todo;
    for (LazyNode* nd : ctx->pending)
        if (!nd->consumed && !nd->result) {
            nd->refs++;
            todo.push_back(nd);
        }
    int status = DSC_OK;
    for (LazyNode* nd : todo) {
        if (status == DSC_OK && !nd->result && !nd->consumed && nd->refs > 1) status = lazy_materialize(nd);
        lazy_unref(nd);
    }
    // a synchronisation point of the compute stream covers what the second compute lane was given as well
    const int js = lane_join(ctx);
    return status != DSC_OK ? status : js;
}

}  // namespace dsc
