// Context lifecycle, stream-ordered pool allocator, device data buffers (the device twin of
// ISigmaDiffDataBuffer, reference src/DiffSharp/Util.fs:133-165), CUDA events and CUDA-graph capture.
#include <stdarg.h>

#include <algorithm>

#include <nvtx3/nvToolsExt.h>

#include "dsc_internal.cuh"

namespace dsc {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void nvtx_push(const char* name) { nvtxRangePushA(name); }
void nvtx_pop() { nvtxRangePop(); }

// ---- pool ---------------------------------------------------------------------------------
// Bin sizes: multiples of 512 B up to 1 MiB, multiples of 64 KiB up to 64 MiB, multiples of 2 MiB
// above.  The tape of one MLP step asks for a handful of distinct sizes thousands of times
// (two zero adjoints per node, AD.Float32.fs:1834-1835,3746), so exact-bin reuse hits ~100 %.
static size_t round_bin(size_t bytes) {
    if (bytes == 0) bytes = 1;
    if (bytes <= (1u << 20)) return (bytes + 511) & ~size_t(511);
    if (bytes <= (64u << 20)) return (bytes + 65535) & ~size_t(65535);
    return (bytes + (2u << 20) - 1) & ~size_t((2u << 20) - 1);
}

int pool_alloc(Ctx* ctx, size_t bytes, Block** out) {
    size_t bin = round_bin(bytes);
    Pool& pool = ctx->capturing ? ctx->arena_pool : ctx->pool;
    void* p = nullptr;
    if (!(ctx->flags & DSC_FLAG_NO_POOL)) {
        auto it = pool.free_blocks.find(bin);
        if (it != pool.free_blocks.end()) {
            p = it->second;
            pool.free_blocks.erase(it);
            pool.hits++;
        }
    }
    if (!p) {
        cudaError_t e = cudaMalloc(&p, bin);
        if (e != cudaSuccess) {
            cudaGetLastError();
            // release everything cached and retry once
            for (auto& kv : ctx->pool.free_blocks) {
                cudaFree(kv.second);
                ctx->pool.reserved -= kv.first;
            }
            ctx->pool.free_blocks.clear();
            e = cudaMalloc(&p, bin);
            if (e != cudaSuccess) {
                cudaGetLastError();
                return fail(DSC_E_OOM, "device allocation of %zu bytes failed: %s", bin,
                            cudaGetErrorString(e));
            }
        }
        pool.misses++;
        pool.reserved += bin;
    }
    pool.in_use += bin;
    Block* b = new Block();
    b->ptr = p;
    b->bytes = bin;
    b->refs = 1;
    b->ctx = ctx;
    b->arena = ctx->capturing ? ctx->cur_arena : 0;
    if (ctx->capturing) ctx->arena_blocks.push_back(b);
    *out = b;
    return DSC_OK;
}

// Copies on the copy streams and collectives on the communication stream run OFF the compute stream.  Before a block they
// touched is written by compute-stream work, or handed back to the pool (where the next compute-stream kernel may pick it
// up), the compute stream is ordered after those streams.  Cheap (three event records / waits, no host synchronisation) and
// only done when such a use is actually outstanding.
int join_other_streams(Ctx* ctx) {
    if (ctx->other_stream_joined == ctx->other_stream_seq) return DSC_OK;
    if (ctx->capturing) return DSC_OK;   // copy-stream transfers cannot be captured; captured collectives are joined by dsc_comm_join / graph_end
    DSC_CUDA_CHECK(cudaEventRecord(ctx->ev_copy_b, ctx->copy_stream));
    DSC_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy_b, 0));
    DSC_CUDA_CHECK(cudaEventRecord(ctx->ev_copy_c, ctx->copy_stream_down));
    DSC_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy_c, 0));
    DSC_CUDA_CHECK(cudaEventRecord(ctx->ev_comm, ctx->comm_stream));
    DSC_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0));
    ctx->comm_pending = false;
    ctx->other_stream_joined = ctx->other_stream_seq;
    return DSC_OK;
}

void block_release(Block* blk) {
    if (!blk) return;
    if (--blk->refs > 0) return;
    Ctx* ctx = blk->ctx;
    if (blk->lo) {
        block_release(blk->lo);
        blk->lo = nullptr;
    }
    if (blk->last_other_stream_use > ctx->other_stream_joined) join_other_streams(ctx);
    if (blk->arena == 0) {
        ctx->pool.in_use -= blk->bytes;
        if (ctx->flags & DSC_FLAG_NO_POOL) {
            cudaFree(blk->ptr);
            ctx->pool.reserved -= blk->bytes;
        } else {
            ctx->pool.free_blocks.emplace(blk->bytes, blk->ptr);
        }
        delete blk;
    } else if (blk->arena < 0) {
        // orphan of a destroyed graph
        cudaFree(blk->ptr);
        delete blk;
    } else if (ctx->capturing && blk->arena == ctx->cur_arena) {
        // temporaries of the step being captured are recycled inside the capture: the graph is a
        // linear chain on one stream, so reuse is ordered exactly as in eager mode.
        ctx->arena_pool.in_use -= blk->bytes;
        ctx->arena_pool.free_blocks.emplace(blk->bytes, blk->ptr);
        // the Block record itself stays in arena_blocks only through its pointer; mark it dead
        blk->ptr = nullptr;
    }
    // else: block belongs to a finished graph; the graph keeps the memory reserved.
}

int buf_new(Ctx* ctx, int dtype, int32_t n, Buf** out) {
    if (n < 0) return fail(DSC_E_INVALID, "negative buffer length %d", n);
    Buf* b = new Buf();
    b->dtype = dtype;
    b->ctx = ctx;
    b->len = n;
    b->off = 0;
    if (n > 0) {
        int s = pool_alloc(ctx, (size_t)n * dtype_size(dtype), &b->block);
        if (s != DSC_OK) {
            delete b;
            return s;
        }
    }
    *out = b;
    return DSC_OK;
}

int buf_writable(Ctx* ctx, Buf* b, bool on_compute_stream) {
    DSC_TRY(buf_resolve(b));
    Block* old = b->block;
    if (!old) return DSC_OK;
    // deferred operations that still have to read this block see the value it has NOW
    while (!old->readers.empty()) {
        LazyNode* nd = old->readers.back();
        nd->refs++;
        int s = lazy_materialize(nd);   // releases its operands: removes itself from old->readers
        if (s == DSC_OK && !old->readers.empty() && old->readers.back() == nd) s = fail(DSC_E_INVALID, "deferred reader did not release its operand");
        lazy_unref(nd);
        if (s != DSC_OK) return s;
    }
    // a write issued on the compute stream must come after copies / collectives that touch the block on the other streams
    // (a write issued on a copy stream is ordered by that stream and by the caller's events)
    if (on_compute_stream && old->last_other_stream_use > ctx->other_stream_joined) DSC_TRY(join_other_streams(ctx));
    DSC_TRY(lane_gate_write(ctx, old));   // ... and after lane operations that read or wrote it
    if (old->lo) DSC_TRY(lane_gate_write(ctx, old->lo));
    if (old->cow > 0 && old->refs > 1) {
        // copy-on-write: other handles (or deferred results) hold this block and must keep seeing the old value
        Block* nb = nullptr;
        const size_t es = dtype_size(b->dtype);
        DSC_TRY(pool_alloc(ctx, (size_t)b->len * es, &nb));
        cudaError_t e = cudaMemcpyAsync(nb->ptr, reinterpret_cast<const char*>(old->ptr) + b->off * es, (size_t)b->len * es,
                                        cudaMemcpyDeviceToDevice, ctx->stream);
        if (e != cudaSuccess) {
            block_release(nb);
            return fail(DSC_E_CUDA, "copy-on-write copy failed: %s", cudaGetErrorString(e));
        }
        ctx->stats.kernel_launches++;
        if (b->is_cow) old->cow--;
        b->is_cow = false;
        block_release(old);
        b->block = nb;
        b->off = 0;
        return DSC_OK;
    }
    if (b->is_cow && old->refs == 1) {   // the last holder owns the block outright
        b->is_cow = false;
        old->cow--;
    }
    old->lo_epoch = 0;   // the operand-split companion no longer matches
    return DSC_OK;
}

// b is about to be overwritten completely: like buf_writable, but a deferred value is dropped instead of computed and a
// shared block is left to the other holders instead of copied.
static int buf_overwritable(Ctx* ctx, Buf* b) {
    if (b->len == 0) return DSC_OK;
    if (!b->block) {
        if (b->lazy) {
            lazy_unref(b->lazy);
            b->lazy = nullptr;
        }
        return pool_alloc(ctx, (size_t)b->len * dtype_size(b->dtype), &b->block);
    }
    Block* old = b->block;
    if (b->is_view || old->views > 0 || !(old->cow > 0 && old->refs > 1)) return buf_writable(ctx, b);   // shared mutable storage, or sole holder
    Block* nb = nullptr;
    DSC_TRY(pool_alloc(ctx, (size_t)b->len * dtype_size(b->dtype), &nb));
    if (b->is_cow) old->cow--;
    b->is_cow = false;
    block_release(old);
    b->block = nb;
    b->off = 0;
    return DSC_OK;
}

int out_buf(Ctx* ctx, int dtype, int32_t n, dsc_buf_t* out, Buf** res) {
    if (!out) return fail(DSC_E_INVALID, "null output handle pointer");
    if (*out == 0) {
        Buf* b = nullptr;
        DSC_TRY(buf_new(ctx, dtype, n, &b));
        *out = reinterpret_cast<dsc_buf_t>(b);
        *res = b;
        return DSC_OK;
    }
    Buf* b = as_buf(*out);
    if (!b) return fail(DSC_E_INVALID, "output handle is not a live buffer");
    if (b->ctx != ctx) return fail(DSC_E_INVALID, "output buffer belongs to another context");
    if (b->dtype != dtype) return fail(DSC_E_DTYPE, "output buffer has the wrong dtype");
    if (b->len != n) return fail(DSC_E_SHAPE, "output buffer has length %d, expected %d", b->len, n);
    DSC_TRY(buf_overwritable(ctx, b));
    *res = b;
    return DSC_OK;
}

int in_buf_peek(Ctx* ctx, int dtype, dsc_buf_t h, Buf** res, const char* what) {
    Buf* b = as_buf(h);
    if (!b) return fail(DSC_E_INVALID, "%s is not a live buffer handle", what);
    if (b->ctx != ctx) return fail(DSC_E_INVALID, "%s belongs to another context", what);
    if (b->dtype != dtype) return fail(DSC_E_DTYPE, "%s has the wrong dtype", what);
    *res = b;   // no data is touched yet: the callers reach it through buf_resolve / operand_resolve / buf_writable, which order the compute stream after the second lane where needed
    return DSC_OK;
}

int in_buf(Ctx* ctx, int dtype, dsc_buf_t h, Buf** res, const char* what) {
    DSC_TRY(in_buf_peek(ctx, dtype, h, res, what));
    return buf_resolve(*res);
}

int ensure_red_scratch(Ctx* ctx, size_t bytes) {
    if (ctx->red_scratch_bytes >= bytes) return DSC_OK;
    if (ctx->capturing || ctx->live_graphs > 0)
        return fail(DSC_E_INVALID, "reduction scratch cannot grow while a CUDA graph is being captured or is alive (run the step once eagerly first)");
    if (ctx->red_scratch) cudaFree(ctx->red_scratch);
    ctx->red_scratch = nullptr;
    ctx->red_scratch_bytes = 0;
    DSC_CUDA_CHECK(cudaMalloc(&ctx->red_scratch, bytes));
    DSC_CUDA_CHECK(cudaMemsetAsync(ctx->red_scratch, 0, bytes, ctx->stream));
    ctx->red_scratch_bytes = bytes;
    if (ctx->lane_red_scratch_bytes < bytes) {   // the second compute lane's scratch keeps pace (see ensure_gemm_ws)
        if (ctx->lane_red_scratch) {
            DSC_CUDA_CHECK(cudaDeviceSynchronize());
            cudaFree(ctx->lane_red_scratch);
        }
        ctx->lane_red_scratch = nullptr;
        ctx->lane_red_scratch_bytes = 0;
        DSC_CUDA_CHECK(cudaMalloc(&ctx->lane_red_scratch, bytes));
        DSC_CUDA_CHECK(cudaMemsetAsync(ctx->lane_red_scratch, 0, bytes, ctx->stream));
        ctx->lane_red_scratch_bytes = bytes;
    }
    return DSC_OK;
}

int ensure_gemm_ws(Ctx* ctx, size_t bytes) {
    if (ctx->gemm_ws_bytes >= bytes) return DSC_OK;
    if (ctx->capturing || ctx->live_graphs > 0)
        return fail(DSC_E_INVALID, "GEMM workspace cannot grow while a CUDA graph is being captured or is alive (run the step once eagerly first)");
    if (ctx->gemm_ws) {
        DSC_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->gemm_ws);
    }
    ctx->gemm_ws = nullptr;
    ctx->gemm_ws_bytes = 0;
    size_t want = (bytes + (8u << 20) - 1) & ~size_t((8u << 20) - 1);
    DSC_CUDA_CHECK(cudaMalloc(&ctx->gemm_ws, want));
    ctx->gemm_ws_bytes = want;
    // The other compute lane runs a subset of the products this one has seen: keep its workspace as large, so that a pair of
    // products co-scheduled for the first time while a graph is being captured never has to grow anything.
    if (ctx->lane_gemm_ws_bytes < want) {
        if (ctx->lane_gemm_ws) {
            DSC_CUDA_CHECK(cudaDeviceSynchronize());
            cudaFree(ctx->lane_gemm_ws);
        }
        ctx->lane_gemm_ws = nullptr;
        ctx->lane_gemm_ws_bytes = 0;
        DSC_CUDA_CHECK(cudaMalloc(&ctx->lane_gemm_ws, want));
        ctx->lane_gemm_ws_bytes = want;
    }
    return DSC_OK;
}

int ensure_gemm_tickets(Ctx* ctx, size_t count) {
    if (ctx->gemm_tickets_n >= count) return DSC_OK;
    if (ctx->capturing || ctx->live_graphs > 0)
        return fail(DSC_E_INVALID, "GEMM ticket array cannot grow while a CUDA graph is being captured or is alive (run the step once eagerly first)");
    if (ctx->gemm_tickets) {
        DSC_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->gemm_tickets);
    }
    ctx->gemm_tickets = nullptr;
    ctx->gemm_tickets_n = 0;
    size_t want = 4096;
    while (want < count) want *= 2;
    DSC_CUDA_CHECK(cudaMalloc(&ctx->gemm_tickets, want * sizeof(unsigned int)));
    DSC_CUDA_CHECK(cudaMemsetAsync(ctx->gemm_tickets, 0, want * sizeof(unsigned int), ctx->stream));
    ctx->gemm_tickets_n = want;
    if (ctx->lane_gemm_tickets_n < want) {   // see ensure_gemm_ws: the other lane keeps pace
        if (ctx->lane_gemm_tickets) {
            DSC_CUDA_CHECK(cudaDeviceSynchronize());
            cudaFree(ctx->lane_gemm_tickets);
        }
        ctx->lane_gemm_tickets = nullptr;
        ctx->lane_gemm_tickets_n = 0;
        DSC_CUDA_CHECK(cudaMalloc(&ctx->lane_gemm_tickets, want * sizeof(unsigned int)));
        DSC_CUDA_CHECK(cudaMemsetAsync(ctx->lane_gemm_tickets, 0, want * sizeof(unsigned int), ctx->stream));
        ctx->lane_gemm_tickets_n = want;
    }
    return DSC_OK;
}

template <typename T>
__global__ void fill_kernel(T* __restrict__ p, size_t n, T v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

int fill_raw(Ctx* ctx, int dtype, void* base, size_t n, double v) {
    if (n == 0) return DSC_OK;
    if (v == 0.0 && !signbit(v)) {
        DSC_CUDA_CHECK(cudaMemsetAsync(base, 0, n * dtype_size(dtype), ctx->stream));
        ctx->stats.kernel_launches++;
        return DSC_OK;
    }
    int grid = grid_for(n, 256 * 8, ctx->sm_count * 8);
    if (dtype == DSC_F32)
        fill_kernel<float><<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<float*>(base), n, (float)v);
    else
        fill_kernel<double><<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<double*>(base), n, v);
    DSC_LAUNCH_CHECK(ctx);
    return DSC_OK;
}

static int fill_window(Ctx* ctx, Buf* b, int32_t off, int32_t n, double v) {
    if (n == 0) return DSC_OK;
    size_t es = dtype_size(b->dtype);
    return fill_raw(ctx, b->dtype, reinterpret_cast<char*>(b->block->ptr) + (b->off + off) * es, (size_t)n, v);
}

// strided 2-D gather used by GetStackedValues (Util.fs:204-210)
template <typename T>
__global__ void gather2d_kernel(const T* __restrict__ src, T* __restrict__ dst, int cols, int r0,
                                int c0, int orows, int ocols) {
    size_t n = (size_t)orows * ocols;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        int r = (int)(i / ocols), c = (int)(i % ocols);
        dst[i] = src[(size_t)(r0 + r) * cols + (c0 + c)];
    }
}

}  // namespace dsc

using namespace dsc;

extern "C" {

int dsc_version(void) { return DSC_VERSION; }

const char* dsc_last_error(void) { return g_err; }

int dsc_device_count(int* count) {
    if (!count) return fail(DSC_E_INVALID, "null count");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *count = 0;
        return fail(DSC_E_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *count = n;
    return DSC_OK;
}

int dsc_create(int device, uint32_t flags, dsc_ctx_t* out) {
    if (!out) return fail(DSC_E_INVALID, "null ctx pointer");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(DSC_E_CUDA, "no CUDA device available (%s); this backend has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= ndev) return fail(DSC_E_INVALID, "device %d out of range [0,%d)", device, ndev);
    DSC_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    DSC_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(DSC_E_CUDA, "device %d is sm_%d%d; libdiffsharp_cuda is built for sm_100a only", device,
                    prop.major, prop.minor);
    Ctx* ctx = new Ctx();
    ctx->device = device;
    ctx->flags = flags;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    {
        const char* e = getenv("DSC_FUSE");   // DSC_FUSE=0: same as DSC_FLAG_NO_FUSE (debugging)
        ctx->fuse = !(flags & DSC_FLAG_NO_FUSE) && !(e && *e == '0');
        const char* g = getenv("DSC_GEMM_EPILOGUE");
        if (g && *g >= '0' && *g <= '2') ctx->gemm_epilogue = *g - '0';
        const char* v = getenv("DSC_NVTX");
        ctx->nvtx = (flags & DSC_FLAG_NVTX) || (v && *v == '1');
    }
    DSC_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    DSC_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
    DSC_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    DSC_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->copy_stream_down, cudaStreamNonBlocking));
    DSC_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_copy_a, cudaEventDisableTiming));
    DSC_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_copy_b, cudaEventDisableTiming));
    DSC_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_copy_c, cudaEventDisableTiming));
    DSC_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_compute, cudaEventDisableTiming));
    DSC_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_comm, cudaEventDisableTiming));
    DSC_CUDA_CHECK(cudaMallocHost(&ctx->red_host, 256));
    int s = ensure_red_scratch(ctx, 32u << 20);
    if (s != DSC_OK) return s;
    DSC_CUDA_CHECK(cudaMalloc(&ctx->tickets, kNumTickets * sizeof(unsigned int)));
    DSC_CUDA_CHECK(cudaMemsetAsync(ctx->tickets, 0, kNumTickets * sizeof(unsigned int), ctx->stream));
    DSC_CUDA_CHECK(cudaMalloc(&ctx->lane_tickets, kNumTickets * sizeof(unsigned int)));
    DSC_CUDA_CHECK(cudaMemsetAsync(ctx->lane_tickets, 0, kNumTickets * sizeof(unsigned int), ctx->stream));
    {
        const char* e = getenv("DSC_LANES");   // DSC_LANES=0: every kernel on the compute stream (debugging / measurements)
        ctx->lanes_enabled = ctx->fuse && !(e && *e == '0');
    }
    *out = reinterpret_cast<dsc_ctx_t>(ctx);
    return DSC_OK;
}

int dsc_comm_destroy(dsc_ctx_t c);

int dsc_destroy(dsc_ctx_t c) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    cudaSetDevice(ctx->device);
    lane_join(ctx);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->comm_stream);
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamSynchronize(ctx->copy_stream_down);
    cudaEventDestroy(ctx->ev_copy_a);
    cudaEventDestroy(ctx->ev_copy_b);
    cudaEventDestroy(ctx->ev_copy_c);
    cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream_down);
    if (ctx->nccl_comm) dsc_comm_destroy(c);
    for (auto& kv : ctx->pool.free_blocks) cudaFree(kv.second);
    ctx->pool.free_blocks.clear();
    if (ctx->red_scratch) cudaFree(ctx->red_scratch);
    if (ctx->tickets) cudaFree(ctx->tickets);
    if (ctx->red_host) cudaFreeHost(ctx->red_host);
    if (ctx->flush_buf) cudaFree(ctx->flush_buf);
    if (ctx->gemm_ws) cudaFree(ctx->gemm_ws);
    if (ctx->gemm_tickets) cudaFree(ctx->gemm_tickets);
    if (ctx->lane_gemm_ws) cudaFree(ctx->lane_gemm_ws);
    if (ctx->lane_gemm_tickets) cudaFree(ctx->lane_gemm_tickets);
    if (ctx->lane_red_scratch) cudaFree(ctx->lane_red_scratch);
    if (ctx->lane_tickets) cudaFree(ctx->lane_tickets);
    if (ctx->lane_stream) {
        cudaStreamSynchronize(ctx->lane_stream);
        cudaEventDestroy(ctx->ev_lane_fork);
        if (ctx->ev_lane_prefork) cudaEventDestroy(ctx->ev_lane_prefork);
        cudaEventDestroy(ctx->ev_lane_done);
        cudaStreamDestroy(ctx->lane_stream);
    }
    if (ctx->comm_stage) cudaFree(ctx->comm_stage);
    cudaEventDestroy(ctx->ev_compute);
    cudaEventDestroy(ctx->ev_comm);
    cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->comm_stream);
    ctx->magic = 0;
    delete ctx;
    return DSC_OK;
}

int dsc_comm_join(dsc_ctx_t c);

int dsc_sync(dsc_ctx_t c) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    if (ctx->capturing) return fail(DSC_E_INVALID, "dsc_sync while capturing a graph");
    DSC_TRY(flush_pending(ctx));
    DSC_TRY(dsc_comm_join(c));
    DSC_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return dsc_comm_join(c);   // ... and report a barrier time-out the work just synchronised may have hit
}

int dsc_get_stats(dsc_ctx_t c, dsc_stats* out) {
    if (!ctx_ok(c) || !out) return fail(DSC_E_INVALID, "bad context or null stats");
    Ctx* ctx = as_ctx(c);
    ctx->stats.pool_bytes_reserved = ctx->pool.reserved;
    ctx->stats.pool_bytes_in_use = ctx->pool.in_use;
    ctx->stats.pool_hits = ctx->pool.hits;
    ctx->stats.pool_misses = ctx->pool.misses;
    ctx->stats.lane_ops = ctx->lane_seq;
    *out = ctx->stats;
    return DSC_OK;
}

int dsc_get_stream(dsc_ctx_t c, void** s) {
    if (!ctx_ok(c) || !s) return fail(DSC_E_INVALID, "bad context or null stream pointer");
    *s = as_ctx(c)->stream;
    return DSC_OK;
}

int dsc_trim_pool(dsc_ctx_t c) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    if (ctx->capturing) return fail(DSC_E_INVALID, "dsc_trim_pool while capturing");
    DSC_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    for (auto& kv : ctx->pool.free_blocks) {
        cudaFree(kv.second);
        ctx->pool.reserved -= kv.first;
    }
    ctx->pool.free_blocks.clear();
    return DSC_OK;
}

// ---- events -------------------------------------------------------------------------------
int dsc_event_create(dsc_ctx_t c, dsc_event_t* ev) {
    if (!ctx_ok(c) || !ev) return fail(DSC_E_INVALID, "bad context or null event pointer");
    cudaEvent_t e;
    DSC_CUDA_CHECK(cudaEventCreate(&e));
    *ev = reinterpret_cast<dsc_event_t>(e);
    return DSC_OK;
}
int dsc_event_record(dsc_ctx_t c, dsc_event_t ev) {
    if (!ctx_ok(c) || !ev) return fail(DSC_E_INVALID, "bad context or event");
    DSC_TRY(flush_pending(as_ctx(c)));   // what the caller issued so far runs before the event
    DSC_CUDA_CHECK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(ev), as_ctx(c)->stream));
    return DSC_OK;
}
int dsc_event_elapsed_ms(dsc_ctx_t c, dsc_event_t a, dsc_event_t b, float* ms) {
    if (!ctx_ok(c) || !a || !b || !ms) return fail(DSC_E_INVALID, "bad context, event or pointer");
    DSC_CUDA_CHECK(cudaEventSynchronize(reinterpret_cast<cudaEvent_t>(b)));
    DSC_CUDA_CHECK(cudaEventElapsedTime(ms, reinterpret_cast<cudaEvent_t>(a), reinterpret_cast<cudaEvent_t>(b)));
    return DSC_OK;
}
int dsc_event_destroy(dsc_ctx_t c, dsc_event_t ev) {
    if (!ctx_ok(c) || !ev) return fail(DSC_E_INVALID, "bad context or event");
    DSC_CUDA_CHECK(cudaEventDestroy(reinterpret_cast<cudaEvent_t>(ev)));
    return DSC_OK;
}

int dsc_event_record_copy(dsc_ctx_t c, dsc_event_t ev) {
    if (!ctx_ok(c) || !ev) return fail(DSC_E_INVALID, "bad context or event");
    DSC_CUDA_CHECK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(ev), as_ctx(c)->copy_stream));
    return DSC_OK;
}
int dsc_event_record_copy_down(dsc_ctx_t c, dsc_event_t ev) {
    if (!ctx_ok(c) || !ev) return fail(DSC_E_INVALID, "bad context or event");
    DSC_CUDA_CHECK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(ev), as_ctx(c)->copy_stream_down));
    return DSC_OK;
}
int dsc_event_sync(dsc_ctx_t c, dsc_event_t ev) {
    if (!ctx_ok(c) || !ev) return fail(DSC_E_INVALID, "bad context or event");
    DSC_CUDA_CHECK(cudaEventSynchronize(reinterpret_cast<cudaEvent_t>(ev)));
    return DSC_OK;
}

int dsc_compute_wait_event(dsc_ctx_t c, dsc_event_t ev) {
    if (!ctx_ok(c) || !ev) return fail(DSC_E_INVALID, "bad context or event");
    DSC_CUDA_CHECK(cudaStreamWaitEvent(as_ctx(c)->stream, reinterpret_cast<cudaEvent_t>(ev), 0));
    return DSC_OK;
}

int dsc_flush_l2(dsc_ctx_t c) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    const size_t bytes = 256u << 20;  // 2x the 126 MB L2
    if (!ctx->flush_buf) {
        if (ctx->capturing) return fail(DSC_E_INVALID, "first dsc_flush_l2 while capturing");
        DSC_CUDA_CHECK(cudaMalloc(&ctx->flush_buf, bytes));
        ctx->flush_bytes = bytes;
    }
    DSC_CUDA_CHECK(cudaMemsetAsync(ctx->flush_buf, 1, ctx->flush_bytes, ctx->stream));
    return DSC_OK;
}

// ---- graphs -------------------------------------------------------------------------------
int dsc_graph_begin(dsc_ctx_t c) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    if (ctx->capturing) return fail(DSC_E_INVALID, "already capturing");
    DSC_TRY(flush_pending(ctx));   // joins the second compute lane as well
    DSC_TRY(join_other_streams(ctx));
    DSC_TRY(dsc_comm_join(c));
    DSC_CUDA_CHECK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
    ctx->capturing = true;
    ctx->lane_fork_seq = 0;   // the lane is not part of this capture until its first fork
    ctx->prefork_pending = false;
    ctx->cache_epoch++;   // operand-split companions computed before the capture are not trusted inside it: a replay must recompute what depends on replaced inputs
    ctx->cur_arena = ctx->next_arena++;
    ctx->arena_pool = Pool();
    ctx->arena_blocks.clear();
    return DSC_OK;
}

int dsc_graph_end(dsc_ctx_t c, dsc_graph_t* out) {
    if (!ctx_ok(c) || !out) return fail(DSC_E_INVALID, "bad context or null graph pointer");
    Ctx* ctx = as_ctx(c);
    if (!ctx->capturing) return fail(DSC_E_INVALID, "not capturing");
    {
        int fs = flush_pending(ctx);   // deferred work issued inside the capture belongs to the graph
        if (fs != DSC_OK) {
            cudaGraph_t dead = nullptr;
            cudaStreamEndCapture(ctx->stream, &dead);
            if (dead) cudaGraphDestroy(dead);
            cudaGetLastError();
            ctx->capturing = false;
            return fs;
        }
    }
    ctx->cache_epoch++;   // ... and companions computed inside it describe the capture-time values only
    ctx->lane_fork_seq = 0;
    ctx->prefork_pending = false;
    if (ctx->comm_pending) {
        // join the communication stream back into the origin stream before ending the capture
        cudaEventRecord(ctx->ev_comm, ctx->comm_stream);
        cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0);
        ctx->comm_pending = false;
    }
    Graph* g = new Graph();
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &g->graph);
    ctx->capturing = false;
    g->arena = ctx->cur_arena;
    // Blocks released inside the capture were recycled through arena_pool; what remains cached there
    // plus what is still referenced is the graph's private memory.
    for (Block* b : ctx->arena_blocks) {
        if (b->ptr == nullptr) delete b;  // record of a recycled block (memory lives on in another record or the pool)
        else g->blocks.push_back(b);
    }
    for (auto& kv : ctx->arena_pool.free_blocks) {
        Block* b = new Block();
        b->ptr = kv.second;
        b->bytes = kv.first;
        b->refs = 0;
        b->ctx = ctx;
        b->arena = g->arena;
        g->blocks.push_back(b);
    }
    ctx->arena_blocks.clear();
    ctx->arena_pool = Pool();
    ctx->cur_arena = 0;
    if (e != cudaSuccess) {
        cudaGetLastError();
        delete g;
        return fail(DSC_E_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
    }
    e = cudaGraphInstantiate(&g->exec, g->graph, 0);
    if (e != cudaSuccess) {
        cudaGetLastError();
        cudaGraphDestroy(g->graph);
        delete g;
        return fail(DSC_E_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    }
    ctx->live_graphs++;
    *out = reinterpret_cast<dsc_graph_t>(g);
    return DSC_OK;
}

int dsc_graph_launch(dsc_ctx_t c, dsc_graph_t gh) {
    if (!ctx_ok(c) || !gh) return fail(DSC_E_INVALID, "bad context or graph");
    Ctx* ctx = as_ctx(c);
    Graph* g = reinterpret_cast<Graph*>(gh);
    DSC_CUDA_CHECK(cudaGraphLaunch(g->exec, ctx->stream));
    return DSC_OK;
}

int dsc_graph_destroy(dsc_ctx_t c, dsc_graph_t gh) {
    if (!ctx_ok(c) || !gh) return fail(DSC_E_INVALID, "bad context or graph");
    Ctx* ctx = as_ctx(c);
    Graph* g = reinterpret_cast<Graph*>(gh);
    cudaStreamSynchronize(ctx->stream);
    cudaGraphExecDestroy(g->exec);
    cudaGraphDestroy(g->graph);
    ctx->live_graphs--;
    for (Block* b : g->blocks) {
        if (b->refs <= 0) {
            cudaFree(b->ptr);
            delete b;
        } else {
            b->arena = -1;  // still referenced by a live buffer: freed when that buffer goes
        }
    }
    delete g;
    return DSC_OK;
}

// ---- buffers ------------------------------------------------------------------------------
int dsc_buf_alloc(dsc_ctx_t c, int dtype, int32_t n, dsc_buf_t* out) {
    if (!ctx_ok(c) || !out) return fail(DSC_E_INVALID, "bad context or null out");
    if (dtype != DSC_F32 && dtype != DSC_F64) return fail(DSC_E_DTYPE, "unknown dtype %d", dtype);
    Buf* b = nullptr;
    DSC_TRY(buf_new(as_ctx(c), dtype, n, &b));
    *out = reinterpret_cast<dsc_buf_t>(b);
    return DSC_OK;
}

int dsc_buf_alloc_fill(dsc_ctx_t c, int dtype, int32_t n, double v, dsc_buf_t* out) {
    if (ctx_ok(c) && out && as_ctx(c)->fuse && n > 0 && (dtype == DSC_F32 || dtype == DSC_F64) && v == v) {
        // two zero adjoints per tape node (AD.Float32.fs:1834-1835,3746) and the all-ones seed of Sum_DM's adjoint (:3430):
        // nothing is allocated or written until somebody needs the bytes
        const bool zero = v == 0.0 && !signbit(v);
        LazyNode* nd = lazy_new(as_ctx(c), zero ? LZ_ZERO : LZ_FILL, dtype, n);
        nd->s = v;
        return lazy_buf(as_ctx(c), nd, out);
    }
    DSC_TRY(dsc_buf_alloc(c, dtype, n, out));
    Buf* b = as_buf(*out);
    int s = fill_window(as_ctx(c), b, 0, n, v);
    if (s != DSC_OK) {
        dsc_buf_free(*out);
        *out = 0;
    }
    return s;
}

int dsc_buf_free(dsc_buf_t h) {
    if (h == 0) return DSC_OK;
    Buf* b = as_buf(h);
    if (!b) return fail(DSC_E_INVALID, "dsc_buf_free: not a live buffer handle");
    if (b->lazy) lazy_unref(b->lazy);
    if (b->block) {
        if (b->is_cow) b->block->cow--;
        if (b->is_view) b->block->views--;
    }
    block_release(b->block);
    b->magic = 0;
    delete b;
    return DSC_OK;
}

static int check_range(Buf* b, int32_t off, int32_t n, const char* what) {
    if (off < 0 || n < 0 || (int64_t)off + n > b->len)
        return fail(DSC_E_SHAPE, "%s: range [%d,%d) outside buffer of length %d", what, off, off + n, b->len);
    return DSC_OK;
}

int dsc_buf_upload(dsc_buf_t h, int32_t off, const void* host, int32_t n) {
    Buf* b = as_buf(h);
    if (!b) return fail(DSC_E_INVALID, "upload: bad handle");
    DSC_TRY(check_range(b, off, n, "upload"));
    if (n == 0) return DSC_OK;
    if (!host) return fail(DSC_E_INVALID, "upload: null host pointer");
    Ctx* ctx = b->ctx;
    DSC_TRY(buf_writable(ctx, b));
    size_t es = dtype_size(b->dtype);
    char* dst = reinterpret_cast<char*>(b->block->ptr) + (b->off + off) * es;
    DSC_CUDA_CHECK(cudaMemcpyAsync(dst, host, (size_t)n * es, cudaMemcpyHostToDevice, ctx->stream));
    if (!ctx->capturing) DSC_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // pageable source may be reused by the caller
    ctx->stats.memcpy_h2d_bytes += (size_t)n * es;
    return DSC_OK;
}

int dsc_buf_upload_async(dsc_buf_t h, int32_t off, const void* host, int32_t n) {
    Buf* b = as_buf(h);
    if (!b) return fail(DSC_E_INVALID, "upload_async: bad handle");
    DSC_TRY(check_range(b, off, n, "upload_async"));
    if (n == 0) return DSC_OK;
    if (!host) return fail(DSC_E_INVALID, "upload_async: null host pointer");
    Ctx* ctx = b->ctx;
    DSC_TRY(buf_writable(ctx, b));
    size_t es = dtype_size(b->dtype);
    char* dst = reinterpret_cast<char*>(b->block->ptr) + (b->off + off) * es;
    DSC_CUDA_CHECK(cudaMemcpyAsync(dst, host, (size_t)n * es, cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.memcpy_h2d_bytes += (size_t)n * es;
    return DSC_OK;
}

int dsc_buf_download(dsc_buf_t h, int32_t off, void* host, int32_t n) {
    Buf* b = as_buf(h);
    if (!b) return fail(DSC_E_INVALID, "download: bad handle");
    DSC_TRY(check_range(b, off, n, "download"));
    if (n == 0) return DSC_OK;
    if (!host) return fail(DSC_E_INVALID, "download: null host pointer");
    Ctx* ctx = b->ctx;
    if (ctx->capturing) return fail(DSC_E_INVALID, "synchronous download while capturing a graph");
    DSC_TRY(buf_resolve(b));
    DSC_TRY(dsc_comm_join(reinterpret_cast<dsc_ctx_t>(ctx)));   // (joins the second compute lane as well)
    size_t es = dtype_size(b->dtype);
    const char* src = reinterpret_cast<const char*>(b->block->ptr) + (b->off + off) * es;
    DSC_CUDA_CHECK(cudaMemcpyAsync(host, src, (size_t)n * es, cudaMemcpyDeviceToHost, ctx->stream));
    DSC_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->stats.memcpy_d2h_bytes += (size_t)n * es;
    return DSC_OK;
}

int dsc_buf_download_async(dsc_buf_t h, int32_t off, void* host, int32_t n) {
    Buf* b = as_buf(h);
    if (!b) return fail(DSC_E_INVALID, "download_async: bad handle");
    DSC_TRY(check_range(b, off, n, "download_async"));
    if (n == 0) return DSC_OK;
    if (!host) return fail(DSC_E_INVALID, "download_async: null host pointer");
    Ctx* ctx = b->ctx;
    DSC_TRY(buf_resolve(b));
    size_t es = dtype_size(b->dtype);
    const char* src = reinterpret_cast<const char*>(b->block->ptr) + (b->off + off) * es;
    DSC_CUDA_CHECK(cudaMemcpyAsync(host, src, (size_t)n * es, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->stats.memcpy_d2h_bytes += (size_t)n * es;
    return DSC_OK;
}

int dsc_copy_upload(dsc_buf_t h, int32_t off, const void* host, int32_t n) {
    Buf* b = as_buf(h);
    if (!b) return fail(DSC_E_INVALID, "copy_upload: bad handle");
    DSC_TRY(check_range(b, off, n, "copy_upload"));
    if (n == 0) return DSC_OK;
    if (!host) return fail(DSC_E_INVALID, "copy_upload: null host pointer");
    Ctx* ctx = b->ctx;
    if (ctx->capturing) return fail(DSC_E_INVALID, "copy-stream transfers cannot be captured");
    DSC_TRY(buf_writable(ctx, b, false));   // deferred readers run first, copy-on-write aliases keep the old value, the split companion is dropped
    size_t es = dtype_size(b->dtype);
    char* dst = reinterpret_cast<char*>(b->block->ptr) + (b->off + off) * es;
    DSC_CUDA_CHECK(cudaMemcpyAsync(dst, host, (size_t)n * es, cudaMemcpyHostToDevice, ctx->copy_stream));
    note_other_stream_use(ctx, b->block);
    ctx->stats.memcpy_h2d_bytes += (size_t)n * es;
    return DSC_OK;
}

int dsc_copy_download(dsc_buf_t h, int32_t off, void* host, int32_t n) {
    Buf* b = as_buf(h);
    if (!b) return fail(DSC_E_INVALID, "copy_download: bad handle");
    DSC_TRY(check_range(b, off, n, "copy_download"));
    if (n == 0) return DSC_OK;
    if (!host) return fail(DSC_E_INVALID, "copy_download: null host pointer");
    Ctx* ctx = b->ctx;
    if (ctx->capturing) return fail(DSC_E_INVALID, "copy-stream transfers cannot be captured");
    if (!b->block && b->lazy && b->lazy->result) DSC_TRY(buf_resolve(b));   // computed already (e.g. at the end of a captured step): adopt it, nothing is launched
    if (!b->block) return fail(DSC_E_INVALID, "copy_download: the buffer's value is still deferred (dsc_flush / dsc_copy_wait_compute first)");
    size_t es = dtype_size(b->dtype);
    const char* src = reinterpret_cast<const char*>(b->block->ptr) + (b->off + off) * es;
    DSC_CUDA_CHECK(cudaMemcpyAsync(host, src, (size_t)n * es, cudaMemcpyDeviceToHost, ctx->copy_stream_down));
    note_other_stream_use(ctx, b->block);
    ctx->stats.memcpy_d2h_bytes += (size_t)n * es;
    return DSC_OK;
}

int dsc_copy_wait_compute(dsc_ctx_t c) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    DSC_TRY(flush_pending(ctx));
    DSC_TRY(dsc_comm_join(c));
    DSC_CUDA_CHECK(cudaEventRecord(ctx->ev_copy_a, ctx->stream));
    DSC_CUDA_CHECK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy_a, 0));
    DSC_CUDA_CHECK(cudaStreamWaitEvent(ctx->copy_stream_down, ctx->ev_copy_a, 0));
    return DSC_OK;
}

int dsc_compute_wait_copy(dsc_ctx_t c) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    DSC_CUDA_CHECK(cudaEventRecord(ctx->ev_copy_b, ctx->copy_stream));
    DSC_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy_b, 0));
    DSC_CUDA_CHECK(cudaEventRecord(ctx->ev_copy_c, ctx->copy_stream_down));
    DSC_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy_c, 0));
    if (!ctx->comm_pending) ctx->other_stream_joined = ctx->other_stream_seq;
    return DSC_OK;
}

int dsc_copy_sync(dsc_ctx_t c) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    DSC_CUDA_CHECK(cudaStreamSynchronize(as_ctx(c)->copy_stream));
    DSC_CUDA_CHECK(cudaStreamSynchronize(as_ctx(c)->copy_stream_down));
    return DSC_OK;
}

int dsc_buf_fill(dsc_buf_t h, int32_t off, int32_t n, double v) {
    Buf* b = as_buf(h);
    if (!b) return fail(DSC_E_INVALID, "fill: bad handle");
    DSC_TRY(check_range(b, off, n, "fill"));
    if (n > 0) DSC_TRY(buf_writable(b->ctx, b));
    return fill_window(b->ctx, b, off, n, v);
}

int dsc_buf_view(dsc_buf_t h, int32_t off, int32_t n, dsc_buf_t* out) {
    Buf* b = as_buf(h);
    if (!b || !out) return fail(DSC_E_INVALID, "view: bad handle or null out");
    DSC_TRY(check_range(b, off, n, "view"));
    if (n > 0) DSC_TRY(buf_writable(b->ctx, b));  // a true view shares MUTABLE storage: materialise, detach from copy-on-write aliases first
    Buf* v = new Buf();
    v->dtype = b->dtype;
    v->ctx = b->ctx;
    v->len = n;
    if (n > 0) {
        v->block = b->block;
        v->block->refs++;
        v->block->views++;
        v->is_view = true;
        v->off = b->off + off;
    }
    *out = reinterpret_cast<dsc_buf_t>(v);
    return DSC_OK;
}

int dsc_buf_copy(dsc_buf_t h, dsc_buf_t* out);

int dsc_buf_share(dsc_buf_t h, dsc_buf_t* out) {
    Buf* b = as_buf(h);
    if (!b || !out) return fail(DSC_E_INVALID, "share: bad handle or null out");
    Buf* v = new Buf();
    v->dtype = b->dtype;
    v->ctx = b->ctx;
    v->len = b->len;
    int s = buf_share_into(v, b);
    if (s == 1) {   // a block with true views (shared mutable storage) is copied eagerly
        delete v;
        *out = 0;
        return dsc_buf_copy(h, out);
    }
    *out = reinterpret_cast<dsc_buf_t>(v);
    return DSC_OK;
}

int dsc_buf_copy(dsc_buf_t h, dsc_buf_t* out) {
    Buf* b = as_buf(h);
    if (!b || !out) return fail(DSC_E_INVALID, "copy: bad handle or null out");
    Ctx* ctx = b->ctx;
    DSC_TRY(buf_resolve(b));
    Buf* o = nullptr;
    DSC_TRY(out_buf(ctx, b->dtype, b->len, out, &o));
    if (b->len == 0) return DSC_OK;
    size_t es = dtype_size(b->dtype);
    DSC_CUDA_CHECK(cudaMemcpyAsync(reinterpret_cast<char*>(o->block->ptr) + o->off * es,
                                   reinterpret_cast<const char*>(b->block->ptr) + b->off * es,
                                   (size_t)b->len * es, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->stats.kernel_launches++;
    return DSC_OK;
}

int dsc_buf_gather2d(dsc_buf_t h, int32_t rows, int32_t cols, int32_t r0, int32_t r1, int32_t c0,
                     int32_t c1, dsc_buf_t* out) {
    Buf* b = as_buf(h);
    if (!b || !out) return fail(DSC_E_INVALID, "gather2d: bad handle or null out");
    if (rows < 0 || cols < 0 || (int64_t)rows * cols > b->len)
        return fail(DSC_E_SHAPE, "gather2d: %d x %d exceeds buffer length %d", rows, cols, b->len);
    if (r0 < 0 || r1 >= rows || c0 < 0 || c1 >= cols || r1 < r0 - 1 || c1 < c0 - 1)
        return fail(DSC_E_SHAPE, "gather2d: slice [%d..%d, %d..%d] outside %d x %d", r0, r1, c0, c1, rows, cols);
    DSC_TRY(buf_resolve(b));
    int orows = r1 - r0 + 1, ocols = c1 - c0 + 1;
    Ctx* ctx = b->ctx;
    Buf* o = nullptr;
    DSC_TRY(out_buf(ctx, b->dtype, orows * ocols, out, &o));
    if (orows * ocols == 0) return DSC_OK;
    int grid = grid_for((size_t)orows * ocols, 256 * 4, ctx->sm_count * 8);
    if (b->dtype == DSC_F32)
        gather2d_kernel<float><<<grid, 256, 0, ctx->stream>>>(b->ptr<float>(), o->ptr<float>(), cols, r0, c0, orows, ocols);
    else
        gather2d_kernel<double><<<grid, 256, 0, ctx->stream>>>(b->ptr<double>(), o->ptr<double>(), cols, r0, c0, orows, ocols);
    DSC_LAUNCH_CHECK(ctx);
    return DSC_OK;
}

int dsc_buf_info(dsc_buf_t h, int* dtype, int32_t* length, int32_t* offset, int32_t* alloc_length,
                 void** device_ptr) {
    Buf* b = as_buf(h);
    if (!b) return fail(DSC_E_INVALID, "info: bad handle");
    if (device_ptr || alloc_length) DSC_TRY(buf_resolve(b));
    if (dtype) *dtype = b->dtype;
    if (length) *length = b->len;
    if (offset) *offset = (int32_t)b->off;
    if (alloc_length) *alloc_length = b->block ? (int32_t)std::min<size_t>(b->block->bytes / dtype_size(b->dtype), 0x7fffffff) : 0;
    if (device_ptr) *device_ptr = b->block ? reinterpret_cast<char*>(b->block->ptr) + b->off * dtype_size(b->dtype) : nullptr;
    return DSC_OK;
}

int dsc_set_fuse(dsc_ctx_t c, int on) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    if (!on) DSC_TRY(flush_pending(ctx));
    ctx->fuse = on != 0;
    return DSC_OK;
}

int dsc_invalidate_caches(dsc_ctx_t c) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    as_ctx(c)->cache_epoch++;
    if (!as_ctx(c)->capturing) as_ctx(c)->lane_fork_seq = 0;   // stale companions may be rewritten: the second lane's next operation forks afresh
    return DSC_OK;
}

int dsc_flush(dsc_ctx_t c) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    return flush_pending(as_ctx(c));
}

int dsc_host_alloc_pinned(size_t bytes, void** out) {
    if (!out) return fail(DSC_E_INVALID, "null out");
    DSC_CUDA_CHECK(cudaMallocHost(out, bytes ? bytes : 1));
    return DSC_OK;
}

int dsc_host_free_pinned(void* p) {
    if (!p) return DSC_OK;
    DSC_CUDA_CHECK(cudaFreeHost(p));
    return DSC_OK;
}

}  // extern "C"
