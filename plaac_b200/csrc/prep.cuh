// Length bucketing + transposed packing of the residue stream (HBM-bound integer/byte work).
//
//   offsets --k_len_hist--> hist --k_scan--> cursor --k_scatter--> order (descending length)
//   order   --k_bucket_chunks--> nchunks --k_scan--> chunk_base
//   codes   --k_pack--> stream[(chunk_base[b]+j)*32 + lane] : 16 ext codes of protein order[32b+lane]
//
// After this a warp owns a bucket of 32 proteins of (nearly) equal length and every 16-byte load of
// the scoring kernel is one fully coalesced 512-byte warp transaction.
#pragma once
#include "common.cuh"

namespace plaac {

__device__ __forceinline__ int len_bin(int64_t len)
{
    // bin 0: len >= kHistBins (unsorted among themselves, scheduled first); bin kHistBins: len 0
    return len >= kHistBins ? 0 : (int)(kHistBins - len);
}

__global__ void k_len_hist(const int64_t* __restrict__ offsets, int64_t nprot, int32_t* __restrict__ hist)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nprot) return;
    atomicAdd(&hist[len_bin(offsets[i + 1] - offsets[i])], 1);
}

// Single-CTA exclusive scan, out[0..n] (n+1 values, out[n] = total).  n is small here (bins, buckets).
// Launch with exactly 1024 threads.
template <typename TIn>
__global__ void __launch_bounds__(1024) k_scan_exclusive(const TIn* __restrict__ in, int64_t* __restrict__ out, int64_t n)
{
    constexpr int kPer = 4;
    __shared__ int64_t warp_inc[32];
    __shared__ int64_t warp_ex[32];
    __shared__ int64_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += (int64_t)1024 * kPer) {
        int64_t v[kPer], s = 0;
        const int64_t i0 = base + (int64_t)tid * kPer;
#pragma unroll
        for (int k = 0; k < kPer; k++) {
            v[k] = (i0 + k < n) ? (int64_t)in[i0 + k] : 0;
            s += v[k];
        }
        int64_t inc = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int64_t o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) warp_inc[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int64_t w = warp_inc[lane];
            int64_t winc = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int64_t o = __shfl_up_sync(0xffffffffu, winc, d);
                if (lane >= d) winc += o;
            }
            warp_ex[lane] = winc - w;
        }
        __syncthreads();
        const int64_t carry = carry_s;
        int64_t ex = carry + warp_ex[wid] + (inc - s);
#pragma unroll
        for (int k = 0; k < kPer; k++) {
            if (i0 + k < n) out[i0 + k] = ex;
            ex += v[k];
        }
        __syncthreads();
        if (tid == 1023) carry_s = ex;  // last thread's running value = carry + block total
        __syncthreads();
    }
    if (tid == 0) out[n] = carry_s;
}

__global__ void k_scatter(const int64_t* __restrict__ offsets, int64_t nprot, int64_t* __restrict__ cursor,
                          int32_t* __restrict__ order)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nprot) return;
    int b = len_bin(offsets[i + 1] - offsets[i]);
    unsigned long long r = atomicAdd((unsigned long long*)&cursor[b], 1ull);
    order[r] = (int32_t)i;
}

// One warp per bucket: slots needed = ceil(max length / 16).
__global__ void k_bucket_chunks(const int64_t* __restrict__ offsets, const int32_t* __restrict__ order, int64_t nprot,
                                int64_t nbuckets, int32_t* __restrict__ nchunks)
{
    int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (b >= nbuckets) return;
    int64_t r = b * 32 + lane;
    int64_t len = 0;
    if (r < nprot) {
        int32_t p = order[r];
        len = offsets[p + 1] - offsets[p];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        int64_t o = __shfl_xor_sync(0xffffffffu, len, d);
        len = o > len ? o : len;
    }
    if (lane == 0) nchunks[b] = (int32_t)((len + kChunk - 1) / kChunk);
}

__device__ __forceinline__ uint32_t sel8(const uint4& a, const uint4& b, int i)
{
    // word i (0..7) of the 32-byte pair {a,b}; i is per-lane dynamic
    uint32_t lo = (i & 2) ? ((i & 1) ? a.w : a.z) : ((i & 1) ? a.y : a.x);
    uint32_t hi = (i & 2) ? ((i & 1) ? b.w : b.z) : ((i & 1) ? b.y : b.x);
    return (i & 4) ? hi : lo;
}

// Ext byte layout: bits 4:0 residue code (22 = pad), bit 5 PAPA proline mask, bits 7:6 charge class.
// One warp per bucket.  Lane l streams protein order[32b+l]: aligned 16-byte loads, byte realignment with
// funnel shifts, SWAR sanitising / padding / PAPA proline flags, one coalesced 512-byte store per slot.
__global__ void __launch_bounds__(256)
k_pack(const uint8_t* __restrict__ codes, const int64_t* __restrict__ offsets,
       int64_t off_base, const int32_t* __restrict__ order, const int64_t* __restrict__ chunk_base, int64_t nprot,
       int64_t nbuckets, int adjust_prolines, uint32_t charge_plus, uint32_t charge_minus, uint4* __restrict__ stream,
       int* __restrict__ errflag)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    constexpr uint32_t kPadW = 0x01010101u * kPad;
    for (int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < nbuckets; b += warps) {
        const int64_t r = b * 32 + lane;
        int64_t n = 0;
        const uint8_t* src = codes;
        if (r < nprot) {
            int32_t p = order[r];
            n = offsets[p + 1] - offsets[p];
            src = codes + (offsets[p] - off_base);
        }
        const int64_t cb = chunk_base[b];
        const int nch = (int)(chunk_base[b + 1] - cb);
        const int sh = (int)((uintptr_t)src & 15);
        const uint4* ap = (const uint4*)(src - sh);
        const int s4 = sh >> 2, s1 = (sh & 3) * 8;
        uint4 A = make_uint4(0, 0, 0, 0), B = A;
        if (n > 0) A = __ldg(ap);
        uint32_t carry = 0;  // eq13 bits of the previous word (positions before the current one)
        bool bad_any = false;
        for (int j = 0; j < nch; j++) {
            const int64_t pos0 = (int64_t)j * kChunk;
            // B is needed when the 16 source bytes straddle into the next aligned block and are still in range
            B = make_uint4(0, 0, 0, 0);
            if (sh != 0 && (pos0 + (16 - sh)) < n) B = __ldg(ap + j + 1);
            uint32_t o[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint32_t w0 = sel8(A, B, q + s4);
                uint32_t w1 = sel8(A, B, (q + s4 + 1) & 7);
                uint32_t w = __funnelshift_r(w0, w1, s1);
                // sanitise: bytes > 21 are invalid input -> X (0) and flag the error
                uint32_t bad = (w & 0x80808080u) | (((w & 0x7f7f7f7fu) + 0x6a6a6a6au) & 0x80808080u);
                int64_t rem = n - (pos0 + 4 * q);
                uint32_t vmask = rem >= 4 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << (8 * (int)rem)) - 1u));
                bad &= vmask;
                bad_any |= (bad != 0);
                uint32_t badbytes = (bad >> 7) * 0xffu;
                w &= ~badbytes;
                w = (w & vmask) | (kPadW & ~vmask);
                if (adjust_prolines) {
                    uint32_t x = w ^ 0x0d0d0d0du;
                    uint32_t t = (x | 0x80808080u) - 0x01010101u;
                    uint32_t eq = ~t & 0x80808080u;
                    uint32_t prev1 = (eq << 8) | (carry >> 24);
                    uint32_t prev2 = (eq << 16) | (carry >> 16);
                    uint32_t flag = eq & (prev1 | prev2);
                    carry = eq;
                    w |= flag >> 2;  // bit 7 -> bit 5 (kPapaMaskBit)
                }
                // charge class in bits 7:6 of every byte: 01 = +1, 11 = -1 (so (int8)byte >> 6 is the charge)
                uint32_t cb = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t cd = (w >> (8 * i)) & 31u;
                    const uint32_t pl = (charge_plus >> cd) & 1u, mi = (charge_minus >> cd) & 1u;
                    cb |= ((pl << 6) | (mi * 0xc0u)) << (8 * i);
                }
                w |= cb;
                o[q] = w;
            }
            stream[(cb + j) * 32 + lane] = make_uint4(o[0], o[1], o[2], o[3]);
            // next aligned block becomes A
            if (sh != 0) {
                A = B;
                // if B was not loaded because the protein ended, A is zeros: fine (masked as pad)
                if (!(pos0 + (16 - sh) < n)) A = make_uint4(0, 0, 0, 0);
            } else {
                A = make_uint4(0, 0, 0, 0);
                if (pos0 + 16 < n) A = __ldg(ap + j + 1);
            }
        }
        if (bad_any) atomicOr(errflag, 1);
    }
}

}  // namespace plaac
