// Benchmark utilities (include/plaac_bench.h): synthetic proteomes + FP64 peak microbenchmark.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/plaac_bench.h"
#include "../../include/plaac_cuda.h"

namespace {

struct U4 {
    uint32_t x, y, z, w;
};

// Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0,k1).
__device__ __forceinline__ U4 philox(U4 c, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        U4 n;
        n.x = hi1 ^ c.y ^ k0;
        n.y = lo1;
        n.z = hi0 ^ c.w ^ k1;
        n.w = lo0;
        c = n;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ double u01(uint32_t a, uint32_t b)
{
    // 53 random bits -> (0,1)
    const uint64_t m = ((uint64_t)a << 21) ^ (uint64_t)b;
    return ((double)(m & ((1ull << 53) - 1)) + 0.5) * (1.0 / 9007199254740992.0);
}

__global__ void k_synth_lengths(uint64_t seed, int64_t first, int64_t nprot, double mu, double sigma, int min_len,
                                int max_len, int64_t* __restrict__ lens)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nprot) return;
    const uint64_t g = (uint64_t)(first + i);
    U4 c = {(uint32_t)g, (uint32_t)(g >> 32), 0u, 0x4c454eu /* "LEN" */};
    U4 r = philox(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double u1 = u01(r.x, r.y), u2 = u01(r.z, r.w);
    const double z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    double len = rint(exp(mu + sigma * z));
    len = fmin(fmax(len, (double)min_len), (double)max_len);
    lens[i] = (int64_t)len;
}

struct Cdf {
    float bg[22], prd[22];
};

__device__ __forceinline__ int draw(const float* cdf, float u)
{
    int k = 0;
#pragma unroll
    for (int j = 0; j < 21; j++) k += (u >= cdf[j]) ? 1 : 0;
    return k;
}

// One warp per protein; lane handles residues lane*4 + 128*it .. +3 (one Philox call = 4 residues).
__global__ void k_synth_residues(uint64_t seed, int64_t first, int64_t nprot, const int64_t* __restrict__ offsets, Cdf cdf,
                                 float prd_rate, float x_rate, uint8_t* __restrict__ codes)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < nprot; p += warps) {
        const uint64_t g = (uint64_t)(first + p);
        const int64_t off = offsets[p];
        const int n = (int)(offsets[p + 1] - off);
        U4 hc = {(uint32_t)g, (uint32_t)(g >> 32), 0u, 0x505244u /* "PRD" */};
        U4 h = philox(hc, (uint32_t)seed, (uint32_t)(seed >> 32));
        int seg_lo = -1, seg_hi = -1;
        if ((h.x >> 8) * (1.0f / 16777216.0f) < prd_rate) {
            int seg = 60 + (int)(h.y % 241u);
            if (seg > n) seg = n;
            seg_lo = (int)(h.z % (uint32_t)(n - seg + 1));
            seg_hi = seg_lo + seg;
        }
        for (int base = lane * 4; base < n; base += 128) {
            U4 c = {(uint32_t)g, (uint32_t)(g >> 32), (uint32_t)(base >> 2), 0x524553u /* "RES" */};
            U4 r = philox(c, (uint32_t)seed, (uint32_t)(seed >> 32));
            const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int t = base + q;
                if (t < n) {
                    const float u = (rr[q] >> 8) * (1.0f / 16777216.0f);
                    const bool in_prd = t >= seg_lo && t < seg_hi;
                    int code = draw(in_prd ? cdf.prd : cdf.bg, u);
                    // a second, decorrelated 24-bit uniform decides X
                    const uint32_t xr = (rr[q] * 2654435761u) >> 8;
                    if (xr * (1.0f / 16777216.0f) < x_rate) code = 0;
                    codes[off + t] = (uint8_t)code;
                }
            }
        }
    }
}

template <bool FMA>
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, double seedv)
{
    double a0 = seedv + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 1.0000001, b = 1e-9;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (FMA) {
                a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
                a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
            } else {
                a0 = __dadd_rn(a0, b); a1 = __dadd_rn(a1, b); a2 = __dadd_rn(a2, b); a3 = __dadd_rn(a3, b);
                a4 = __dadd_rn(a4, b); a5 = __dadd_rn(a5, b); a6 = __dadd_rn(a6, b); a7 = __dadd_rn(a7, b);
            }
        }
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 12345.678) out[0] = s;  // never true; keeps the chains alive
}

void normalise_cdf(const double* f, float* cdf)
{
    double tot = 0;
    for (int i = 0; i < 22; i++) tot += f[i];
    double acc = 0;
    for (int i = 0; i < 22; i++) {
        acc += f[i] / tot;
        cdf[i] = (float)acc;
    }
    cdf[21] = 2.0f;
}

}  // namespace

extern "C" {

int plaac_bench_synth_lengths(void* stream, uint64_t seed, int64_t first_index, int64_t nprot, double mu, double sigma,
                              int32_t min_len, int32_t max_len, int64_t* d_lengths)
{
    if (nprot <= 0) return PLAAC_OK;
    const int tb = 256;
    k_synth_lengths<<<(unsigned)((nprot + tb - 1) / tb), tb, 0, (cudaStream_t)stream>>>(seed, first_index, nprot, mu, sigma,
                                                                                       min_len, max_len, d_lengths);
    return cudaGetLastError() == cudaSuccess ? PLAAC_OK : PLAAC_E_CUDA;
}

int plaac_bench_synth_residues(void* stream, uint64_t seed, int64_t first_index, int64_t nprot, const int64_t* d_offsets,
                               const double* bg_freq, const double* prd_freq, double prd_rate, double x_rate,
                               uint8_t* d_codes)
{
    if (nprot <= 0) return PLAAC_OK;
    Cdf cdf;
    normalise_cdf(bg_freq, cdf.bg);
    normalise_cdf(prd_freq, cdf.prd);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = (nprot + 7) / 8;
    const unsigned grid = (unsigned)(want < (int64_t)sms * 16 ? want : (int64_t)sms * 16);
    k_synth_residues<<<grid, 256, 0, (cudaStream_t)stream>>>(seed, first_index, nprot, d_offsets, cdf, (float)prd_rate,
                                                            (float)x_rate, d_codes);
    return cudaGetLastError() == cudaSuccess ? PLAAC_OK : PLAAC_E_CUDA;
}

int plaac_bench_fp64_peak(int device, int use_fma, double* ops_per_s, float* ms_out)
{
    if (cudaSetDevice(device) != cudaSuccess) return PLAAC_E_CUDA;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    double* d = nullptr;
    if (cudaMalloc(&d, 64) != cudaSuccess) return PLAAC_E_NOMEM;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int iters = 20000, grid = sms * 8, tb = 256;
    for (int rep = 0; rep < 2; rep++) {  // first pass warms up
        cudaEventRecord(a);
        if (use_fma)
            k_fp64_peak<true><<<grid, tb>>>(d, iters, 1.0);
        else
            k_fp64_peak<false><<<grid, tb>>>(d, iters, 1.0);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d);
    if (cudaGetLastError() != cudaSuccess) return PLAAC_E_CUDA;
    const double ops = (double)grid * tb * (double)iters * 64.0;
    if (ops_per_s) *ops_per_s = ops / (ms * 1e-3);
    if (ms_out) *ms_out = ms;
    return PLAAC_OK;
}

}  // extern "C"
