// The second compute lane: a stream that runs the SINKS of a reverse sweep beside the compute stream.
//
// In the call sequence the unchanged AD layer issues for a reverse sweep (AD.Float32.fs:4118-4257) three kinds of calls
// produce values nobody reads again before the sweep ends — they only land in the adjoint of a parameter:
//   * the weight-adjoint product  Transpose(a.P) * d.A        (:4131-4143, the A operand is a pending Transpose_M),
//   * the bias-adjoint column sum Add_M_Colwise_V_InPlace      (:4255),
//   * (new surface) the all-reduce of those adjoints over the ranks.
// On one stream they sit in the critical path of the sweep (the next layer's input adjoint waits behind the previous layer's
// weight adjoint).  Here they are issued on a second stream — forked from the compute stream only where a true data
// dependency requires it — and joined back only when somebody touches their result (lane_gate_read / lane_gate_write at
// every handle access) or at a synchronisation point.  Captured in the step's CUDA graph the lane is a parallel branch: the
// adjoint all-reduce of the later layers overlaps the backward products of the first layer, and small products that cannot
// fill 148 SMs run two at a time on half of the SMs each.
//
// Ordering rules (all host-side bookkeeping, no device synchronisation):
//   * Block::w_seq pins the position of a block's last compute-stream write in ctx->order_seq the first time a lane operation
//     looks at it; a fork sets ctx->lane_fork_seq.  A lane operation whose inputs are all pinned at or below lane_fork_seq
//     starts without a new fork — which is what lets the column sums and the all-reduce issued at the END of the sweep start
//     right behind the weight adjoint issued in its middle.
//   * Block::lane_w / lane_r stamp lane writes / reads; compute-stream readers of a lane-written block and writers of a
//     lane-touched block join first.  Blocks a lane operation touches stay referenced until the next join, so the pool never
//     hands them to compute-stream work the lane could still race with.
//   * the lane has its own GEMM workspace, tickets and reduction scratch.
//
// Measured and not adopted (round 2): preparing, on the idle lane during the forward pass, the operand-split companions of the
// operands a forward product converted in its kernel (X, W) so that the adjoint products find them ready.  The weight adjoint
// of the first layer got 1.5 us faster (2-CTA kernel instead of in-kernel conversion) and the pre-pass in front of the input
// adjoint disappeared, but that pre-pass already ran beside the lane's weight adjoint and the paired products got ~1 us slower:
// 127.2 us against 126.3 us per 1024-row step, 503 against 497 us at 8192 rows.  Re-measured A/B on single boxes (weights only,
// after the launch or when the product is created, so that the forward product itself can take the 2-CTA kernel): 122.4 / 124.6
// against 125.5 us on one box, 121.1 / 121.9 / 118.4-122.3 against 120.5 us on another — inside the box-to-box and run-to-run
// spread (2-3 %), not worth the code.
#include <stdlib.h>

#include <algorithm>

#include "dsc_internal.cuh"

namespace dsc {

void lane_swap(Ctx* ctx) {
    std::swap(ctx->stream, ctx->lane_stream);
    std::swap(ctx->gemm_ws, ctx->lane_gemm_ws);
    std::swap(ctx->gemm_ws_bytes, ctx->lane_gemm_ws_bytes);
    std::swap(ctx->gemm_tickets, ctx->lane_gemm_tickets);
    std::swap(ctx->gemm_tickets_n, ctx->lane_gemm_tickets_n);
    std::swap(ctx->red_scratch, ctx->lane_red_scratch);
    std::swap(ctx->red_scratch_bytes, ctx->lane_red_scratch_bytes);
    std::swap(ctx->tickets, ctx->lane_tickets);
    ctx->on_lane = !ctx->on_lane;
}

static int lane_ensure(Ctx* ctx) {
    if (ctx->lane_stream) return DSC_OK;
    DSC_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->lane_stream, cudaStreamNonBlocking));
    DSC_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_lane_fork, cudaEventDisableTiming));
    DSC_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_lane_done, cudaEventDisableTiming));
    return DSC_OK;
}

bool lane_debug() {
    static const bool on = [] { const char* e = getenv("DSC_LANE_DEBUG"); return e && *e == '1'; }();
    return on;
}

static uint64_t pin(Ctx* ctx, Block* b) {
    if (b->w_seq == 0) b->w_seq = ++ctx->order_seq;
    return b->w_seq;
}

bool lane_sees(Ctx* ctx, Block* const* reads, int n) {
    bool ok = ctx->lane_stream != nullptr && ctx->lane_fork_seq != 0;   // (the inputs are pinned in any case: a fork that follows covers them)
    for (int i = 0; i < n; ++i) {
        Block* b = reads[i];
        if (!b) continue;
        if (b->lane_w > ctx->lane_joined && b->w_seq == 0) continue;   // written by the lane itself (stream order)
        const bool fresh = b->w_seq == 0;
        if (pin(ctx, b) > ctx->lane_fork_seq) ok = false;
        if (lane_debug()) fprintf(stderr, "[dsc lane]   read %p: %s at %llu, lane fork %llu\n", b->ptr, fresh ? "pinned now" : "pinned", (unsigned long long)b->w_seq, (unsigned long long)ctx->lane_fork_seq);
    }
    return ok;
}

int lane_enter(Ctx* ctx, Block* const* reads, int n, bool force_fork) {
    if (ctx->on_lane) return fail(DSC_E_INVALID, "lane_enter: already on the lane");
    DSC_TRY(lane_ensure(ctx));
    // The blocks lane operations touch stay referenced until the next join.  A host loop that never reaches a synchronisation
    // point (no loss read, no download, no dsc_sync) would pile them up — and every allocation of the following steps would
    // miss the pool: after kLaneMaxUnjoined operations the lane is joined before it is given more.
    constexpr uint64_t kLaneMaxUnjoined = 24;
    if (ctx->lane_seq - ctx->lane_joined >= kLaneMaxUnjoined) DSC_TRY(lane_join(ctx));
    const bool fork = force_fork || !lane_sees(ctx, reads, n);
    if (lane_debug()) fprintf(stderr, "[dsc lane] enter: %s\n", fork ? "fork" : "no fork");
    if (fork) {
        DSC_CUDA_CHECK(cudaEventRecord(ctx->ev_lane_fork, ctx->stream));
        DSC_CUDA_CHECK(cudaStreamWaitEvent(ctx->lane_stream, ctx->ev_lane_fork, 0));
        ctx->lane_fork_seq = ++ctx->order_seq;   // everything pinned so far (the inputs above included) is visible to the lane
        ctx->prefork_pending = false;            // ... which includes whatever an earlier recorded fork point covered
        ctx->lane_fork_open = true;
    }
    lane_swap(ctx);
    return DSC_OK;
}

int lane_prefork(Ctx* ctx, Block* const* reads, int n) {
    if (ctx->on_lane) return DSC_OK;
    DSC_TRY(lane_ensure(ctx));
    if (lane_sees(ctx, reads, n)) return DSC_OK;   // (pins the inputs) nothing to gain: the lane is already ordered after them
    if (lane_debug()) fprintf(stderr, "[dsc lane] fork point recorded ahead of a compute-stream kernel\n");
    if (!ctx->ev_lane_prefork) DSC_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_lane_prefork, cudaEventDisableTiming));
    DSC_CUDA_CHECK(cudaEventRecord(ctx->ev_lane_prefork, ctx->stream));
    ctx->prefork_seq = ++ctx->order_seq;
    ctx->prefork_pending = true;
    return DSC_OK;
}

int lane_apply_prefork(Ctx* ctx) {
    if (!ctx->prefork_pending) return DSC_OK;
    DSC_CUDA_CHECK(cudaStreamWaitEvent(ctx->lane_stream, ctx->ev_lane_prefork, 0));
    if (ctx->prefork_seq > ctx->lane_fork_seq) ctx->lane_fork_seq = ctx->prefork_seq;
    ctx->lane_fork_open = true;
    ctx->prefork_pending = false;
    return DSC_OK;
}

void lane_leave(Ctx* ctx, Block* const* reads, int nr, Block* const* writes, int nw) {
    if (ctx->on_lane) lane_swap(ctx);
    const uint64_t seq = ++ctx->lane_seq;
    for (int i = 0; i < nr; ++i)
        if (reads[i]) {
            reads[i]->lane_r = seq;
            reads[i]->refs++;
            ctx->lane_held.push_back(reads[i]);
        }
    for (int i = 0; i < nw; ++i)
        if (writes[i]) {
            writes[i]->lane_w = seq;
            writes[i]->w_seq = 0;
            writes[i]->refs++;
            ctx->lane_held.push_back(writes[i]);
        }
}

int lane_join(Ctx* ctx) {
    if (ctx->on_lane) return fail(DSC_E_INVALID, "lane_join while on the lane");
    if (ctx->lane_seq == ctx->lane_joined && !ctx->lane_fork_open) return DSC_OK;   // (a fork without work still has to be joined: a captured graph may not end with a dangling branch)
    int st = DSC_OK;
    if (cudaEventRecord(ctx->ev_lane_done, ctx->lane_stream) != cudaSuccess || cudaStreamWaitEvent(ctx->stream, ctx->ev_lane_done, 0) != cudaSuccess) {
        cudaGetLastError();
        st = fail(DSC_E_CUDA, "joining the second compute lane failed");
    }
    if (lane_debug()) fprintf(stderr, "[dsc lane] join (%llu operations)\n", (unsigned long long)(ctx->lane_seq - ctx->lane_joined));
    ctx->lane_joined = ctx->lane_seq;
    ctx->lane_fork_open = false;
    ctx->prefork_pending = false;
    ctx->main_gemm_half_once = false;
    std::vector<Block*> held;
    held.swap(ctx->lane_held);
    for (Block* b : held) block_release(b);
    return st;
}

}  // namespace dsc
