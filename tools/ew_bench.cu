// Micro-benchmark used to pick the launch shape of the streaming elementwise kernels (not part of the product).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ew_bench tools/ew_bench.cu && /tmp/ew_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

__device__ __forceinline__ float4 ld_stream(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(float* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int U, bool HINT>
__global__ void had_gs(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, size_t nvec) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (size_t)gridDim.x * blockDim.x;
    size_t i = tid;
    for (; i + (U - 1) * nt < nvec; i += U * nt) {
        float4 va[U], vb[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (HINT) { va[u] = ld_stream(a + (i + u * nt) * 4); vb[u] = ld_stream(b + (i + u * nt) * 4); }
            else { va[u] = *(const float4*)(a + (i + u * nt) * 4); vb[u] = *(const float4*)(b + (i + u * nt) * 4); }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float4 r = make_float4(va[u].x * vb[u].x, va[u].y * vb[u].y, va[u].z * vb[u].z, va[u].w * vb[u].w);
            if (HINT) st_stream(y + (i + u * nt) * 4, r); else *(float4*)(y + (i + u * nt) * 4) = r;
        }
    }
    for (; i < nvec; i += nt) {
        float4 va = *(const float4*)(a + i * 4), vb = *(const float4*)(b + i * 4);
        *(float4*)(y + i * 4) = make_float4(va.x * vb.x, va.y * vb.y, va.z * vb.z, va.w * vb.w);
    }
}
// contiguous-chunk variant: each block owns a contiguous chunk; each thread U consecutive-strided vectors
template <int U, bool HINT>
__global__ void had_chunk(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, size_t nvec) {
    size_t base = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x;
    float4 va[U], vb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        size_t i = base + (size_t)u * blockDim.x;
        if (i < nvec) {
            if (HINT) { va[u] = ld_stream(a + i * 4); vb[u] = ld_stream(b + i * 4); }
            else { va[u] = *(const float4*)(a + i * 4); vb[u] = *(const float4*)(b + i * 4); }
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        size_t i = base + (size_t)u * blockDim.x;
        if (i < nvec) {
            float4 r = make_float4(va[u].x * vb[u].x, va[u].y * vb[u].y, va[u].z * vb[u].z, va[u].w * vb[u].w);
            if (HINT) st_stream(y + i * 4, r); else *(float4*)(y + i * 4) = r;
        }
    }
}

template <typename F>
float timeit(F f, void* flush, size_t fb) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    float tot = 0;
    for (int i = 0; i < 10; ++i) {
        cudaMemsetAsync(flush, 1, fb);
        cudaEventRecord(e0);
        f();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); tot += ms;
    }
    return tot / 10;
}

int main() {
    const size_t n = 4096ull * 4096, nvec = n / 4;
    float *a, *b, *y; void* flush; size_t fb = 256u << 20;
    cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 4); cudaMalloc(&y, n * 4); cudaMalloc(&flush, fb);
    cudaMemset(a, 0, n * 4); cudaMemset(b, 0, n * 4);
    const double bytes = 3.0 * n * 4;
    auto rep = [&](const char* name, float ms) { printf("%-44s %8.2f us  %7.1f GB/s\n", name, ms * 1e3, bytes / ms / 1e6); };
    for (int blocks_per_sm : {2, 4, 8, 16}) {
        for (int threads : {128, 256, 512, 1024}) {
            if (blocks_per_sm * threads > 2048) continue;
            int grid = 148 * blocks_per_sm;
            char nm[128];
            snprintf(nm, sizeof nm, "grid-stride U4 plain  %2d x %4d", blocks_per_sm, threads);
            rep(nm, timeit([&] { had_gs<4, false><<<grid, threads>>>(a, b, y, nvec); }, flush, fb));
            snprintf(nm, sizeof nm, "grid-stride U4 hinted %2d x %4d", blocks_per_sm, threads);
            rep(nm, timeit([&] { had_gs<4, true><<<grid, threads>>>(a, b, y, nvec); }, flush, fb));
            snprintf(nm, sizeof nm, "grid-stride U8 hinted %2d x %4d", blocks_per_sm, threads);
            rep(nm, timeit([&] { had_gs<8, true><<<grid, threads>>>(a, b, y, nvec); }, flush, fb));
            snprintf(nm, sizeof nm, "grid-stride U2 hinted %2d x %4d", blocks_per_sm, threads);
            rep(nm, timeit([&] { had_gs<2, true><<<grid, threads>>>(a, b, y, nvec); }, flush, fb));
        }
    }
    for (int threads : {128, 256, 512}) {
        char nm[128];
        snprintf(nm, sizeof nm, "chunk U4 plain  threads %d", threads);
        rep(nm, timeit([&] { had_chunk<4, false><<<(nvec + threads * 4 - 1) / (threads * 4), threads>>>(a, b, y, nvec); }, flush, fb));
        snprintf(nm, sizeof nm, "chunk U4 hinted threads %d", threads);
        rep(nm, timeit([&] { had_chunk<4, true><<<(nvec + threads * 4 - 1) / (threads * 4), threads>>>(a, b, y, nvec); }, flush, fb));
        snprintf(nm, sizeof nm, "chunk U8 hinted threads %d", threads);
        rep(nm, timeit([&] { had_chunk<8, true><<<(nvec + threads * 8 - 1) / (threads * 8), threads>>>(a, b, y, nvec); }, flush, fb));
        snprintf(nm, sizeof nm, "chunk U2 hinted threads %d", threads);
        rep(nm, timeit([&] { had_chunk<2, true><<<(nvec + threads * 2 - 1) / (threads * 2), threads>>>(a, b, y, nvec); }, flush, fb));
        snprintf(nm, sizeof nm, "chunk U1 hinted threads %d", threads);
        rep(nm, timeit([&] { had_chunk<1, true><<<(nvec + threads - 1) / threads, threads>>>(a, b, y, nvec); }, flush, fb));
    }
    // copy reference
    rep("cudaMemcpyAsync D2D (2/3 of the bytes)", timeit([&] { cudaMemcpyAsync(y, a, n * 4, cudaMemcpyDeviceToDevice); }, flush, fb) * 1.5f);
    return 0;
}
