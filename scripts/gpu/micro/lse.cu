// Micro-benchmark: cycles per residue step of the LUT log-sum-exp recurrences (forward / backward, posteriorl :3349-3411)
// as the per-residue and long-sequence kernels run them: one lane = one dependent chain.
// nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -I../../../plaac_b200/csrc -o lse lse.cu
#include <cmath>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "common.cuh"
#include "summary_kernel_v2.cuh"
using namespace plaac;

template <int MODE>
__global__ void k(const uint8_t* __restrict__ codes, const double* __restrict__ lutg, KScalars ks, int n, double* out, long long* cyc)
{
    extern __shared__ __align__(16) unsigned char smem[];
    double2* lut2 = reinterpret_cast<double2*>(smem);
    double2* le = lut2 + PLAAC_LUT_LEN + 1;
    for (int i = threadIdx.x; i <= PLAAC_LUT_LEN; i += blockDim.x)
        lut2[i] = make_double2(i < PLAAC_LUT_LEN ? lutg[i] : 0.0, i + 1 < PLAAC_LUT_LEN ? lutg[i + 1] : 0.0);
    if (threadIdx.x < 32) le[threadIdx.x] = make_double2(-2.5 - 0.03 * threadIdx.x, -3.1 + 0.02 * threadIdx.x);
    __syncthreads();
    const uint32_t lut = smem_u32(lut2);
    const uint8_t* src = codes + (size_t)threadIdx.x * 37;
    double a0 = ks.li0, a1 = ks.li1, c0 = ks.li0 - 1e-3, c1 = ks.li1 + 1e-3;
    const long long t0 = clock64();
    if (MODE == 0 || MODE == 1) {
#pragma unroll 4
        for (int t = 0; t < n; t++) {
            const double2 l = le[src[t] & 31];
            const double f0 = lse_lut2<MODE == 1>(ks.lt00 + a0, ks.lt10 + a1, lut) + l.x;
            const double f1 = lse_lut2<MODE == 1>(ks.lt01 + a0, ks.lt11 + a1, lut) + l.y;
            a0 = f0, a1 = f1;
        }
    } else if (MODE == 2) {
#pragma unroll 4
        for (int t = n - 1; t >= 0; t--) {
            const double2 l = le[src[t] & 31];
            const double x0 = (ks.lt00 + a0) + l.x, x1 = (ks.lt01 + a1) + l.y;
            const double y0 = (ks.lt10 + a0) + l.x, y1 = (ks.lt11 + a1) + l.y;
            a0 = lse_lut2<false>(x0, x1, lut);
            a1 = lse_lut2<false>(y0, y1, lut);
        }
    } else if (MODE == 3) {
#pragma unroll 2
        for (int t = 0; t < n; t++) {
            const double2 l = le[src[t] & 31];
            const double f0 = lse_lut2<false>(ks.lt00 + a0, ks.lt10 + a1, lut) + l.x;
            const double f1 = lse_lut2<false>(ks.lt01 + a0, ks.lt11 + a1, lut) + l.y;
            const double h0 = lse_lut2<false>(ks.lt00 + c0, ks.lt10 + c1, lut) + l.x;
            const double h1 = lse_lut2<false>(ks.lt01 + c0, ks.lt11 + c1, lut) + l.y;
            a0 = f0, a1 = f1, c0 = h0, c1 = h1;
        }
    } else if (MODE == 4) {  // forward and backward chains in one loop
#pragma unroll 2
        for (int t = 0; t < n; t++) {
            const double2 l = le[src[t] & 31];
            const double2 m = le[src[n - 1 - t] & 31];
            const double f0 = lse_lut2<false>(ks.lt00 + a0, ks.lt10 + a1, lut) + l.x;
            const double f1 = lse_lut2<false>(ks.lt01 + a0, ks.lt11 + a1, lut) + l.y;
            const double x0 = (ks.lt00 + c0) + m.x, x1 = (ks.lt01 + c1) + m.y;
            const double y0 = (ks.lt10 + c0) + m.x, y1 = (ks.lt11 + c1) + m.y;
            a0 = f0, a1 = f1;
            c0 = lse_lut2<false>(x0, x1, lut);
            c1 = lse_lut2<false>(y0, y1, lut);
        }
    } else if (MODE == 5) {  // forward, codes prefetched 8 at a time as one 64-bit word
        for (int t = 0; t < n; t += 8) {
            unsigned long long w = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) w |= (unsigned long long)src[t + i] << (8 * i);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const double2 l = le[(w >> (8 * i)) & 31];
                const double f0 = lse_lut2<false>(ks.lt00 + a0, ks.lt10 + a1, lut) + l.x;
                const double f1 = lse_lut2<false>(ks.lt01 + a0, ks.lt11 + a1, lut) + l.y;
                a0 = f0, a1 = f1;
            }
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + c0 + c1;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, const uint8_t* codes, const double* lut, KScalars ks, int n, double* out, long long* cyc)
{
    const size_t smem = (PLAAC_LUT_LEN + 1 + 32) * 16;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int threads : {32, 128, 512, 1024}) {
        k<MODE><<<1, threads, smem>>>(codes, lut, ks, n, out, cyc);
        k<MODE><<<1, threads, smem>>>(codes, lut, ks, n, out, cyc);
        cudaDeviceSynchronize();
        long long c;
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-34s %4d threads: %7.1f cycles/step (%s)\n", name, threads, (double)c / n, cudaGetErrorString(cudaGetLastError()));
    }
}

int main()
{
    const int n = 4096;
    std::vector<uint8_t> h(n + 1024 * 37 + 64);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)((i * 2654435761u >> 13) % 21 + 1);
    std::vector<double> lut(PLAAC_LUT_LEN + 3);
    for (int i = 0; i < PLAAC_LUT_LEN; i++) lut[i] = std::log1p(std::exp(-i / 100.0));
    uint8_t* dc; double *dl, *out; long long* cyc;
    cudaMalloc(&dc, h.size()); cudaMalloc(&dl, lut.size() * 8); cudaMalloc(&out, 8 * 2048); cudaMalloc(&cyc, 64);
    cudaMemcpy(dc, h.data(), h.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dl, lut.data(), lut.size() * 8, cudaMemcpyHostToDevice);
    KScalars ks = {};
    ks.lt00 = std::log(0.98); ks.lt01 = std::log(0.02); ks.lt10 = std::log(0.001); ks.lt11 = std::log(0.999);
    ks.li0 = std::log(0.95); ks.li1 = std::log(0.05);
    run<0>("forward, lse<false>", dc, dl, ks, n, out, cyc);
    run<1>("forward, lse<true> (no range test)", dc, dl, ks, n, out, cyc);
    run<2>("backward, lse<false>", dc, dl, ks, n, out, cyc);
    run<3>("forward, two frames", dc, dl, ks, n, out, cyc);
    run<4>("forward + backward in one loop", dc, dl, ks, n, out, cyc);
    run<5>("forward, codes 8 at a time", dc, dl, ks, n, out, cyc);
    return 0;
}
