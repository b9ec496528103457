// Per-kernel cost of a dependent chain of small kernels replayed from a CUDA graph, with and without
// programmatic dependent launch (griddepcontrol).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pdl_bench.bin pdl_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>

template <bool PDL>
__global__ void step_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n) {
    if (PDL) {
        asm volatile("griddepcontrol.launch_dependents;");
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i] + 1.0f;
}

template <bool PDL>
static float run(cudaStream_t s, float* a, float* b, size_t n, int grid, int chain) {
    cudaGraph_t g;
    cudaGraphExec_t ge;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < chain; ++i) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(256);
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = PDL ? 1 : 0;
        const float* in = (i & 1) ? b : a;
        float* out = (i & 1) ? a : b;
        cudaLaunchKernelEx(&cfg, step_kernel<PDL>, in, out, n);
    }
    cudaStreamEndCapture(s, &g);
    cudaGraphInstantiate(&ge, g, 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaGraphLaunch(ge, s);
    cudaStreamSynchronize(s);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0, s);
        cudaGraphLaunch(ge, s);
        cudaEventRecord(e1, s);
        cudaStreamSynchronize(s);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaGraphExecDestroy(ge);
    cudaGraphDestroy(g);
    return best * 1e3f / chain;
}

int main() {
    cudaStream_t s;
    cudaStreamCreate(&s);
    const size_t nmax = size_t(1) << 23;
    float *a, *b;
    cudaMalloc(&a, nmax * 4);
    cudaMalloc(&b, nmax * 4);
    cudaMemset(a, 0, nmax * 4);
    const int chain = 200;
    struct { size_t n; int grid; } cases[] = {{1024, 4}, {10240, 40}, {1 << 20, 1024}, {1 << 20, 148}, {size_t(1) << 23, 4096}};
    for (auto c : cases) {
        float t0 = run<false>(s, a, b, c.n, c.grid, chain);
        float t1 = run<true>(s, a, b, c.n, c.grid, chain);
        printf("n=%9zu grid=%5d : plain %.2f us/kernel   PDL %.2f us/kernel\n", c.n, c.grid, t0, t1);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    // correctness of the chain under PDL: a[0] after an even number of +1 steps
    return e != cudaSuccess;
}
