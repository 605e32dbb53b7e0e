// Micro-benchmark: sustained DFMA rate of one B200 as a function of resident warps per SM and independent chains per thread.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;
}

template <int ILP>
void run(int warpsPerSM, int nSM) {
    const int threads = 128, blocksPerSM = warpsPerSM * 32 / threads, iters = 4096;
    double* d;
    cudaMalloc(&d, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    dfma_kernel<ILP><<<nSM * blocksPerSM, threads>>>(d, 16, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    dfma_kernel<ILP><<<nSM * blocksPerSM, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = (double)nSM * blocksPerSM * threads * iters * ILP;
    printf("warps/SM %2d ILP %d: %.2f TFLOPS (FMA = 2), %.1f DFMA/clk/SM at 1.965 GHz\n", warpsPerSM, ILP, 2 * fmas / ms / 1e9,
           fmas / (ms * 1e-3) / nSM / 1.965e9);
    cudaFree(d);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int nSM = p.multiProcessorCount;
    for (int w : {4, 8, 12, 16, 32, 64}) {
        run<1>(w, nSM);
        run<2>(w, nSM);
        run<4>(w, nSM);
        run<8>(w, nSM);
    }
    return 0;
}
