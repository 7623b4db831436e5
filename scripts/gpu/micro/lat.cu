// Micro-benchmarks on B200: dependent-issue latency of FP64 ops (one warp alone on an SM) and throughput with 1..8 warps
// per scheduler partition.  nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -o lat lat.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k_chain(double* out, long long* cyc, int n, double a, double b)
{
    double x = a + threadIdx.x, y = b, z = a * 0.5, w = b * 0.25;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        if (MODE == 0) {  // dependent DADD chain, 16 per iteration
#pragma unroll
            for (int k = 0; k < 16; k++) x = x + y;
        } else if (MODE == 1) {  // dependent DMUL
#pragma unroll
            for (int k = 0; k < 16; k++) x = x * y;
        } else if (MODE == 2) {  // DADD -> DSETP -> FSEL -> DADD (Viterbi shape)
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const double v0 = x + y, v1 = z + w;
                const bool p = v1 > v0;
                x = (p ? v1 : v0) + y;
                z = (p ? v0 : v1) + w;
            }
        } else if (MODE == 3) {  // 4 independent DADD chains
#pragma unroll
            for (int k = 0; k < 4; k++) {
                x = x + y; z = z + w; y = y + 1.0; w = w + 2.0;
            }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + y + z + w;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 4096);
    const int n = 4096;
    long long h;
    for (int threads : {32, 128, 256, 512, 768, 1024}) {
        k_chain<0><<<1, threads>>>(out, cyc, n, 1.0, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("threads %4d  dep DADD: %.2f cyc/op (warp0)", threads, (double)h / (n * 16.0));
        k_chain<1><<<1, threads>>>(out, cyc, n, 1.0, 1.0000001); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("  dep DMUL: %.2f", (double)h / (n * 16.0));
        k_chain<2><<<1, threads>>>(out, cyc, n, 1.0, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("  vit step (2 DADD, DSETP, 4 FSEL, 2 DADD): %.2f cyc/step", (double)h / (n * 8.0));
        k_chain<3><<<1, threads>>>(out, cyc, n, 1.0, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("  4 indep chains: %.2f cyc/op\n", (double)h / (n * 16.0));
    }
    return 0;
}
