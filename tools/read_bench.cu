// Read-only / write-only HBM stream rates on B200 (yardstick for the reduction kernels; not part of the product).
#include <cuda_runtime.h>
#include <cstdio>
template <int U>
__global__ void read_only(const float4* __restrict__ a, size_t nvec, float* sink) {
    size_t base = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x;
    float4 v[U];
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < U; ++u) { size_t i = base + (size_t)u * blockDim.x; v[u] = i < nvec ? a[i] : make_float4(0, 0, 0, 0); }
#pragma unroll
    for (int u = 0; u < U; ++u) s += v[u].x + v[u].y + v[u].z + v[u].w;
    if (s == 12345.678f) sink[0] = s;
}
template <int U>
__global__ void write_only(float4* __restrict__ y, size_t nvec) {
    size_t base = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x;
#pragma unroll
    for (int u = 0; u < U; ++u) { size_t i = base + (size_t)u * blockDim.x; if (i < nvec) y[i] = make_float4(1, 2, 3, 4); }
}
template <typename F> float timeit(F f, void* flush, size_t fb) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); float tot = 0;
    for (int i = 0; i < 10; ++i) { cudaMemsetAsync(flush, 1, fb); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); tot += ms; }
    return tot / 10;
}
int main() {
    for (size_t mb : {64, 256, 1024}) {
        size_t n = mb * 1024 * 1024 / 4, nvec = n / 4;
        float4 *a; float* sink; void* flush; size_t fb = 256u << 20;
        cudaMalloc(&a, n * 4); cudaMalloc(&sink, 4); cudaMalloc(&flush, fb); cudaMemset(a, 0, n * 4);
        for (int threads : {256, 512}) {
            float ms = timeit([&] { read_only<4><<<(nvec + threads * 4 - 1) / (threads * 4), threads>>>(a, nvec, sink); }, flush, fb);
            printf("read-only  %4zu MiB U4 threads %d: %8.2f us %7.1f GB/s\n", mb, threads, ms * 1e3, n * 4.0 / ms / 1e6);
            ms = timeit([&] { read_only<8><<<(nvec + threads * 8 - 1) / (threads * 8), threads>>>(a, nvec, sink); }, flush, fb);
            printf("read-only  %4zu MiB U8 threads %d: %8.2f us %7.1f GB/s\n", mb, threads, ms * 1e3, n * 4.0 / ms / 1e6);
            ms = timeit([&] { write_only<4><<<(nvec + threads * 4 - 1) / (threads * 4), threads>>>(a, nvec); }, flush, fb);
            printf("write-only %4zu MiB U4 threads %d: %8.2f us %7.1f GB/s\n", mb, threads, ms * 1e3, n * 4.0 / ms / 1e6);
        }
        float ms = timeit([&] { cudaMemsetAsync(a, 0, n * 4); }, flush, fb);
        printf("cudaMemset %4zu MiB: %8.2f us %7.1f GB/s\n", mb, ms * 1e3, n * 4.0 / ms / 1e6);
        cudaFree(a); cudaFree(sink); cudaFree(flush);
    }
}
