// C1: NCCL glue.  New surface (the reference has no collective): one communicator per
// context (= per process = per GPU).  Collectives run on the context's communication stream, ordered
// after the producing kernels by an event, so the all-reduce of one layer's weight adjoints overlaps
// the rest of the reverse sweep; dsc_comm_join() orders the compute stream after them.
//
// NCCL is resolved with dlopen at dsc_comm_init time so that the library loads (and every
// single-GPU path runs) on machines without libnccl, and so that a process that already carries
// torch's bundled libnccl.so.2 shares that one copy instead of loading a second one.
#include <dlfcn.h>

#include "dsc_internal.cuh"

namespace dsc {

typedef struct { char internal[128]; } nccl_unique_id;
typedef void* nccl_comm_t;
enum { NCCL_SUCCESS = 0 };
enum { NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8 };  // ncclDataType_t values (nccl.h)
enum { NCCL_SUM = 0 };

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(nccl_unique_id*) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl() {
    if (g_nccl.lib) return DSC_OK;
    const char* names[] = {getenv("DSC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* nm : names) {
        if (!nm || !*nm) continue;
        lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) return fail(DSC_E_NCCL, "cannot dlopen libnccl.so.2 (set DSC_NCCL_LIB): %s", dlerror());
#define SYM(field, name)                                                        \
    *(void**)(&g_nccl.field) = dlsym(lib, name);                                \
    if (!g_nccl.field) return fail(DSC_E_NCCL, "libnccl is missing symbol %s", name);
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce")
    SYM(AllGather, "ncclAllGather")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.lib = lib;
    return DSC_OK;
}

#define DSC_NCCL_CHECK(expr)                                                                            \
    do {                                                                                                \
        int _r = (expr);                                                                                \
        if (_r != NCCL_SUCCESS) return fail(DSC_E_NCCL, "%s failed: %s", #expr, g_nccl.GetErrorString(_r)); \
    } while (0)

// order the comm stream after everything queued so far on the compute stream
static int comm_fork(Ctx* ctx) {
    DSC_CUDA_CHECK(cudaEventRecord(ctx->ev_compute, ctx->stream));
    DSC_CUDA_CHECK(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_compute, 0));
    ctx->comm_pending = true;
    return DSC_OK;
}

// Pack / unpack up to kMaxPack buffers into one contiguous staging buffer so that the adjoints of a
// whole model cross NVLink as ONE all-reduce (six separate collectives cost ~15 us of latency each).
constexpr int kMaxPack = 64;
template <typename T>
struct PackTable {
    T* ptr[kMaxPack];
    long long start[kMaxPack + 1];  // prefix offsets in the staging buffer (elements)
    int count;
};
template <typename T, bool PACK>
__global__ void pack_kernel(PackTable<T> t, T* __restrict__ stage) {
    const long long total = t.start[t.count];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = t.count - 1;  // binary search for the owning buffer
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (t.start[mid] <= i) lo = mid; else hi = mid - 1;
        }
        if (PACK) stage[i] = t.ptr[lo][i - t.start[lo]];
        else t.ptr[lo][i - t.start[lo]] = stage[i];
    }
}

template <typename T>
static int allreduce_impl(dsc_ctx_t c, int32_t count, const dsc_buf_t* bufs) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    if (count < 0 || (count > 0 && !bufs)) return fail(DSC_E_INVALID, "allreduce: bad buffer list");
    if (ctx->nranks == 1) return DSC_OK;  // a single rank already holds the sum
    if (!ctx->nccl_comm) return fail(DSC_E_NCCL, "allreduce: dsc_comm_init has not been called");
    PackTable<T> t;
    t.count = 0;
    t.start[0] = 0;
    for (int i = 0; i < count; ++i) {
        Buf* b;
        DSC_TRY(in_buf(ctx, dtype_of<T>::value, bufs[i], &b, "allreduce: buffer"));
        if (b->len == 0) continue;
        DSC_TRY(buf_writable(ctx, b));  // in-place collective
        if (t.count == kMaxPack) return fail(DSC_E_INVALID, "allreduce: more than %d buffers in one call", kMaxPack);
        t.ptr[t.count] = b->ptr<T>();
        t.start[t.count + 1] = t.start[t.count] + b->len;
        t.count++;
    }
    if (t.count == 0) return DSC_OK;
    const int dt = sizeof(T) == 4 ? NCCL_FLOAT32 : NCCL_FLOAT64;
    const long long total = t.start[t.count];
    if (t.count == 1) {
        DSC_TRY(comm_fork(ctx));
        DSC_NCCL_CHECK(g_nccl.AllReduce(t.ptr[0], t.ptr[0], (size_t)total, dt, NCCL_SUM, ctx->nccl_comm, ctx->comm_stream));
        ctx->stats.kernel_launches++;
        return DSC_OK;
    }
    // pack (compute stream) -> one all-reduce (comm stream) -> unpack (comm stream); dsc_comm_join orders readers
    const size_t bytes = (size_t)total * sizeof(T);
    if (ctx->comm_stage_bytes < bytes) {
        if (ctx->capturing || ctx->live_graphs > 0)
            return fail(DSC_E_INVALID, "all-reduce staging buffer cannot grow while a CUDA graph is being captured or is alive");
        if (ctx->comm_stage) { cudaStreamSynchronize(ctx->comm_stream); cudaFree(ctx->comm_stage); }
        ctx->comm_stage = nullptr;
        ctx->comm_stage_bytes = 0;
        DSC_CUDA_CHECK(cudaMalloc(&ctx->comm_stage, bytes));
        ctx->comm_stage_bytes = bytes;
    }
    T* stage = reinterpret_cast<T*>(ctx->comm_stage);
    const int grid = grid_for((size_t)total, 256 * 4, ctx->sm_count * 4);
    if (ctx->comm_pending) {  // the staging buffer may still be in use by a previous collective
        DSC_CUDA_CHECK(cudaEventRecord(ctx->ev_comm, ctx->comm_stream));
        DSC_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0));
    }
    pack_kernel<T, true><<<grid, 256, 0, ctx->stream>>>(t, stage);
    DSC_LAUNCH_CHECK(ctx);
    DSC_TRY(comm_fork(ctx));
    DSC_NCCL_CHECK(g_nccl.AllReduce(stage, stage, (size_t)total, dt, NCCL_SUM, ctx->nccl_comm, ctx->comm_stream));
    ctx->stats.kernel_launches++;
    pack_kernel<T, false><<<grid, 256, 0, ctx->comm_stream>>>(t, stage);
    DSC_LAUNCH_CHECK(ctx);
    return DSC_OK;
}

template <typename T>
static int allgather_impl(dsc_ctx_t c, dsc_buf_t sendh, dsc_buf_t recvh) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    Buf *s, *r;
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, sendh, &s, "allgather: send"));
    DSC_TRY(in_buf(ctx, dtype_of<T>::value, recvh, &r, "allgather: recv"));
    if ((int64_t)s->len * ctx->nranks != r->len)
        return fail(DSC_E_SHAPE, "allgather: recv length %d != %d ranks x send length %d", r->len, ctx->nranks, s->len);
    if (s->len == 0) return DSC_OK;
    DSC_TRY(buf_writable(ctx, r));
    if (ctx->nranks == 1) {
        DSC_CUDA_CHECK(cudaMemcpyAsync(r->ptr<T>(), s->ptr<T>(), (size_t)s->len * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
        return DSC_OK;
    }
    if (!ctx->nccl_comm) return fail(DSC_E_NCCL, "allgather: dsc_comm_init has not been called");
    DSC_TRY(comm_fork(ctx));
    const int dt = sizeof(T) == 4 ? NCCL_FLOAT32 : NCCL_FLOAT64;
    DSC_NCCL_CHECK(g_nccl.AllGather(s->ptr<T>(), r->ptr<T>(), (size_t)s->len, dt, ctx->nccl_comm, ctx->comm_stream));
    ctx->stats.kernel_launches++;
    return DSC_OK;
}

}  // namespace dsc

using namespace dsc;

extern "C" {

int dsc_comm_unique_id(uint8_t out[128]) {
    if (!out) return fail(DSC_E_INVALID, "null id buffer");
    DSC_TRY(load_nccl());
    nccl_unique_id id;
    DSC_NCCL_CHECK(g_nccl.GetUniqueId(&id));
    memcpy(out, id.internal, 128);
    return DSC_OK;
}

int dsc_comm_init(dsc_ctx_t c, int nranks, int rank, const uint8_t idbytes[128]) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(DSC_E_INVALID, "comm_init: rank %d of %d", rank, nranks);
    if (ctx->nccl_comm) return fail(DSC_E_INVALID, "comm_init: communicator already initialised");
    ctx->nranks = nranks;
    ctx->rank = rank;
    if (nranks == 1) return DSC_OK;
    if (!idbytes) return fail(DSC_E_INVALID, "comm_init: null unique id");
    DSC_TRY(load_nccl());
    DSC_CUDA_CHECK(cudaSetDevice(ctx->device));
    nccl_unique_id id;
    memcpy(id.internal, idbytes, 128);
    nccl_comm_t comm = nullptr;
    DSC_NCCL_CHECK(g_nccl.CommInitRank(&comm, nranks, id, rank));
    ctx->nccl_comm = comm;
    return DSC_OK;
}

int dsc_comm_join(dsc_ctx_t c) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    if (!ctx->comm_pending) return DSC_OK;
    DSC_CUDA_CHECK(cudaEventRecord(ctx->ev_comm, ctx->comm_stream));
    DSC_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0));
    ctx->comm_pending = false;
    return DSC_OK;
}

int dsc_comm_destroy(dsc_ctx_t c) {
    if (!ctx_ok(c)) return fail(DSC_E_INVALID, "bad context");
    Ctx* ctx = as_ctx(c);
    if (ctx->nccl_comm) {
        cudaStreamSynchronize(ctx->comm_stream);
        g_nccl.CommDestroy(ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    ctx->nranks = 1;
    ctx->rank = 0;
    return DSC_OK;
}

int dsc_sallreduce_sum(dsc_ctx_t c, int32_t count, const dsc_buf_t* bufs) { return allreduce_impl<float>(c, count, bufs); }
int dsc_dallreduce_sum(dsc_ctx_t c, int32_t count, const dsc_buf_t* bufs) { return allreduce_impl<double>(c, count, bufs); }
int dsc_sallgather(dsc_ctx_t c, dsc_buf_t s, dsc_buf_t r) { return allgather_impl<float>(c, s, r); }
int dsc_dallgather(dsc_ctx_t c, dsc_buf_t s, dsc_buf_t r) { return allgather_impl<double>(c, s, r); }

}  // extern "C"
