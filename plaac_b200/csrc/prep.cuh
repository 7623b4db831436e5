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

// Length histogram.  Lengths cluster in a few hundred bins, so global atomics serialise on a handful of L2
// sectors; each CTA therefore counts its contiguous share of the proteins in a private shared-memory histogram
// (kHistBins+1 counters, dynamic shared memory) and flushes only its non-zero bins.
constexpr int kPrepThreads = 1024;
constexpr size_t kHistSmemBytes = sizeof(int32_t) * (kHistBins + 1);

__global__ void __launch_bounds__(kPrepThreads)
k_len_hist(const int64_t* __restrict__ offsets, int64_t nprot, int64_t long_min, int32_t* __restrict__ hist)
{
    extern __shared__ int32_t sh_hist[];
    for (int i = threadIdx.x; i <= kHistBins; i += blockDim.x) sh_hist[i] = 0;
    __syncthreads();
    const int64_t per = (nprot + gridDim.x - 1) / gridDim.x;
    const int64_t lo = per * blockIdx.x, hi = min(nprot, lo + per);
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x)
        atomicAdd(&sh_hist[len_bin(eff_len(offsets[i + 1] - offsets[i], long_min))], 1);
    __syncthreads();
    for (int i = threadIdx.x; i <= kHistBins; i += blockDim.x) {
        const int32_t v = sh_hist[i];
        if (v) atomicAdd(&hist[i], v);
    }
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

// Multi-CTA exclusive scan for the large inputs (bucket sizes, radix-sort tile histograms: several 10^5 values, where the
// single CTA above takes 0.3-0.6 ms): tile sums -> single-CTA scan of the tile sums -> every tile scans itself from its
// base.  A tile is kScanTile values, 1024 threads x 8.
constexpr int kScanPer = 8;
constexpr int kScanTile = 1024 * kScanPer;

__device__ __forceinline__ int64_t block_reduce_1024(int64_t v, int64_t* sh /* 32 */)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    int64_t t = sh[lane];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
    return t;
}

template <typename TIn>
__global__ void __launch_bounds__(1024) k_scan_tile_sums(const TIn* __restrict__ in, int64_t n, int64_t* __restrict__ tile_sum)
{
    __shared__ int64_t sh[32];
    const int64_t i0 = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanPer;
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanPer; k++)
        if (i0 + k < n) s += (int64_t)in[i0 + k];
    s = block_reduce_1024(s, sh);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = s;
}

template <typename TIn>
__global__ void __launch_bounds__(1024)
k_scan_tiles_apply(const TIn* __restrict__ in, int64_t n, const int64_t* __restrict__ tile_base, int64_t ntiles,
                   int64_t* __restrict__ out)
{
    __shared__ int64_t warp_inc[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t i0 = (int64_t)blockIdx.x * kScanTile + (int64_t)tid * kScanPer;
    int64_t v[kScanPer], s = 0;
#pragma unroll
    for (int k = 0; k < kScanPer; k++) {
        v[k] = (i0 + k < n) ? (int64_t)in[i0 + k] : 0;
        s += v[k];
    }
    int64_t inc = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int64_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) warp_inc[wid] = inc;
    __syncthreads();
    int64_t w = warp_inc[lane], winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int64_t o = __shfl_up_sync(0xffffffffu, winc, d);
        if (lane >= d) winc += o;
    }
    const int64_t wex = __shfl_sync(0xffffffffu, winc - w, wid);
    int64_t ex = tile_base[blockIdx.x] + wex + (inc - s);
#pragma unroll
    for (int k = 0; k < kScanPer; k++) {
        if (i0 + k < n) out[i0 + k] = ex;
        ex += v[k];
    }
    if (blockIdx.x == 0 && tid == 0) out[n] = tile_base[ntiles];
}

// Counting-sort scatter.  A CTA ranks a tile of kScatterTile proteins inside shared memory (one shared atomic
// each), reserves one global range per (tile, bin) with a single global atomic, and writes order[].  The order
// inside a bin is arbitrary (results are per protein; nothing depends on it).
constexpr int kScatterPer = 8;
constexpr int kScatterTile = kPrepThreads * kScatterPer;

__global__ void __launch_bounds__(kPrepThreads)
k_scatter(const int64_t* __restrict__ offsets, int64_t nprot, int64_t long_min, int64_t* __restrict__ cursor,
          int32_t* __restrict__ order)
{
    extern __shared__ int32_t sh_cnt[];  // per bin: count during ranking, then the reserved global base
    for (int i = threadIdx.x; i <= kHistBins; i += blockDim.x) sh_cnt[i] = 0;
    __syncthreads();
    const int64_t ntiles = (nprot + kScatterTile - 1) / kScatterTile;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base_i = tile * kScatterTile;
        int bin[kScatterPer], rk[kScatterPer];
#pragma unroll
        for (int k = 0; k < kScatterPer; k++) {
            const int64_t i = base_i + (int64_t)k * kPrepThreads + threadIdx.x;
            bin[k] = -1;
            rk[k] = 0;
            if (i < nprot) {
                bin[k] = len_bin(eff_len(offsets[i + 1] - offsets[i], long_min));
                rk[k] = atomicAdd(&sh_cnt[bin[k]], 1);
            }
        }
        __syncthreads();
        // the first arrival of every bin reserves the tile's range
        int32_t reserved[kScatterPer];
#pragma unroll
        for (int k = 0; k < kScatterPer; k++) {
            reserved[k] = 0;
            if (bin[k] >= 0 && rk[k] == 0)
                reserved[k] = (int32_t)atomicAdd((unsigned long long*)&cursor[bin[k]], (unsigned long long)sh_cnt[bin[k]]);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kScatterPer; k++)
            if (bin[k] >= 0 && rk[k] == 0) sh_cnt[bin[k]] = reserved[k];
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kScatterPer; k++) {
            const int64_t i = base_i + (int64_t)k * kPrepThreads + threadIdx.x;
            if (bin[k] >= 0) order[sh_cnt[bin[k]] + rk[k]] = (int32_t)i;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kScatterPer; k++)
            if (bin[k] >= 0 && rk[k] == 0) sh_cnt[bin[k]] = 0;
        __syncthreads();
    }
}

// One warp per bucket: slots needed = ceil(max length / 16).
__global__ void k_bucket_chunks(const int64_t* __restrict__ offsets, const int32_t* __restrict__ order, int64_t nprot,
                                int64_t nbuckets, int64_t long_min, int32_t* __restrict__ nchunks)
{
    int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (b >= nbuckets) return;
    int64_t r = b * 32 + lane;
    int64_t len = 0;
    if (r < nprot) {
        int32_t p = order[r];
        len = eff_len(offsets[p + 1] - offsets[p], long_min);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        int64_t o = __shfl_xor_sync(0xffffffffu, len, d);
        len = o > len ? o : len;
    }
    if (lane == 0) nchunks[b] = (int32_t)((len + kChunk - 1) / kChunk);
}

// SWAR helpers on 4 packed bytes that are all < 0x80.
__device__ __forceinline__ uint32_t bytes_zero(uint32_t x)
{
    // 0x80 in every zero byte of x; exact when no byte of x is 0x80 (borrows cannot cross a byte whose top bit
    // is forced on), which holds for sanitised codes xor a code pattern or xor 0xff
    return ~((x | 0x80808080u) - 0x01010101u) & ~x & 0x80808080u;
}
__device__ __forceinline__ uint32_t bytes_eq(uint32_t w, uint32_t code)
{
    // 0x80 in every byte of w that equals `code`
    return bytes_zero(w ^ (code * 0x01010101u));
}
__device__ __forceinline__ uint32_t bytes_in_set(uint32_t w, uint32_t set)
{
    // 0x80 in every byte of w whose value is a member of the 32-bit code set (warp-uniform loop over the members)
    uint32_t r = 0;
    while (set) {
        const uint32_t c = (uint32_t)__ffs((int)set) - 1u;
        set &= set - 1u;
        r |= bytes_eq(w, c);
    }
    return r;
}

// Zero-byte tests for words whose bytes are all < 0x80 (sanitised codes xor a code pattern < 0x80): byte + 0x7f carries
// into bit 7 exactly when the byte is not zero and never into the next byte.  The "matches nothing" pattern is 0x7f7f7f7f.
__device__ __forceinline__ uint32_t bytes_eq7(uint32_t w, uint32_t pat)
{
    return ~((w ^ pat) + 0x7f7f7f7fu) & 0x80808080u;
}
__device__ __forceinline__ uint32_t bytes_eq7_2(uint32_t w, uint32_t pat0, uint32_t pat1)
{
    return ~(((w ^ pat0) + 0x7f7f7f7fu) & ((w ^ pat1) + 0x7f7f7f7fu)) & 0x80808080u;
}

// 0x80 in every byte of w (all bytes < 0x80) whose value lies in [lo, hi]: two carries instead of a zero-byte test per
// member (the ALU pipe is k_pack's busiest: 57 % in the round-2 capture)
__device__ __forceinline__ uint32_t bytes_in_range(uint32_t w, uint32_t lo, uint32_t hi)
{
    const uint32_t ge_lo = w + (0x80u - lo) * 0x01010101u;       // bit 7 set: byte >= lo   (byte + 0x80 - lo < 0x100)
    const uint32_t gt_hi = w + (0x7fu - hi) * 0x01010101u;       // bit 7 set: byte > hi
    return ge_lo & ~gt_hi & 0x80808080u;
}

inline bool pack_is_std(uint32_t charge_plus, uint32_t charge_minus)
{
    if (charge_plus == 0 || __builtin_popcount(charge_plus) > 2 || __builtin_popcount(charge_minus) > 2) return false;
    const uint32_t run = charge_plus >> __builtin_ctz(charge_plus);
    return (run & (run + 1u)) == 0u;
}

// Ext byte layout: bits 4:0 residue code (22 = pad), bit 5 PAPA proline mask, bits 7:6 charge class.
// One warp per bucket.  Lane l streams protein order[32b+l]: aligned 16-byte loads two blocks ahead, a
// two-level word barrel shifter + funnel shifts for the byte realignment, SWAR sanitising / padding / PAPA
// proline flags / charge classes, one coalesced 512-byte store per slot.
// Where the time goes (ncu round 2, profiles/r02_pack_summary.txt): the warps wait for their own 16-byte loads (each lane
// walks its own protein, so a warp's load touches 32 lines; the L1 keeps them for 26 % hits and loads that skip it are
// 30 % slower) with the ALU pipe 61 % busy and DRAM at 52 %.  Measured without effect: a third block in flight, 128-byte L2
// prefetch sizes, 128 / 384 / 512 threads per block (40 / 36 / 32 resident warps).
// kStd: both charge sets have at most two members and the +1 set is a run of consecutive codes (PLAAC's D, E), kPro: the
// PAPA proline rule is on -- the warp-uniform tests of the general kernel decided once by the host (pack_is_std).
template <bool kStd, bool kPro>
__global__ void __launch_bounds__(256, 5)
k_pack(const uint8_t* __restrict__ codes, const int64_t* __restrict__ offsets,
       int64_t off_base, const int32_t* __restrict__ order, const int64_t* __restrict__ chunk_base, int64_t nprot,
       int64_t nbuckets, int64_t long_min, int adjust_prolines, uint32_t charge_plus, uint32_t charge_minus,
       uint4* __restrict__ stream,
       int32_t* __restrict__ slot_bucket, int* __restrict__ errflag)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    constexpr uint32_t kPadW = 0x01010101u * kPad;
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    // up to two codes per charge class are matched with straight-line SWAR compares (PLAAC has D,E / K,R);
    // 0x7f7f7f7f never matches a sanitised byte
    const bool small_sets = kStd || (__popc(charge_plus) <= 2 && __popc(charge_minus) <= 2);
    // a set of consecutive codes (PLAAC: D, E = 3, 4 are +1) is one range test
    const bool plus_run = kStd || (charge_plus != 0 && ((charge_plus >> (__ffs((int)charge_plus) - 1)) & ((charge_plus >> (__ffs((int)charge_plus) - 1)) + 1u)) == 0u);
    const uint32_t plus_lo = charge_plus ? (uint32_t)__ffs((int)charge_plus) - 1u : 0u, plus_hi = charge_plus ? 31u - (uint32_t)__clz((int)charge_plus) : 0u;
    uint32_t cp0 = 0x7f7f7f7fu, cp1 = 0x7f7f7f7fu, cm0 = 0x7f7f7f7fu, cm1 = 0x7f7f7f7fu;
    if (charge_plus) cp0 = ((uint32_t)__ffs((int)charge_plus) - 1u) * 0x01010101u;
    if (charge_plus & (charge_plus - 1u)) cp1 = (31u - (uint32_t)__clz((int)charge_plus)) * 0x01010101u;
    if (charge_minus) cm0 = ((uint32_t)__ffs((int)charge_minus) - 1u) * 0x01010101u;
    if (charge_minus & (charge_minus - 1u)) cm1 = (31u - (uint32_t)__clz((int)charge_minus)) * 0x01010101u;
    for (int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < nbuckets; b += warps) {
        const int64_t r = b * 32 + lane;
        int64_t n = 0;
        const uint8_t* src = codes;
        if (r < nprot) {
            int32_t p = order[r];
            n = eff_len(offsets[p + 1] - offsets[p], long_min);
            src = codes + (offsets[p] - off_base);
        }
        const int64_t cb = chunk_base[b];
        const int nch = (int)(chunk_base[b + 1] - cb);
        const int sh = (int)((uintptr_t)src & 15);
        const uint4* ap = (const uint4*)(src - sh);
        const bool s4b1 = (sh & 8) != 0, s4b0 = (sh & 4) != 0;
        const int s1 = (sh & 3) * 8;
        const int64_t nblk = (n + sh + 15) >> 4;  // aligned 16-byte blocks that hold the protein
        uint4 A = nblk > 0 ? __ldg(ap) : zero4;
        uint4 B = nblk > 1 ? __ldg(ap + 1) : zero4;
        uint32_t carry = 0;  // proline bits of the previous word (positions before the current one)
        uint32_t bad_any = 0;
        uint4* dst = stream + cb * 32 + lane;
        for (int j = lane; j < nch; j += 32) slot_bucket[cb + j] = (int32_t)b;
        for (int j = 0; j < nch; j++) {
            const uint4 C = ((int64_t)j + 2 < nblk) ? __ldg(ap + j + 2) : zero4;  // in flight while slot j is packed
            // bytes sh .. sh+15 of the 32-byte pair {A, B}
            uint32_t W0 = A.x, W1 = A.y, W2 = A.z, W3 = A.w, W4 = B.x, W5 = B.y, W6 = B.z, W7 = B.w;
            if (s4b1) {
                W0 = W2; W1 = W3; W2 = W4; W3 = W5; W4 = W6; W5 = W7;
            }
            if (s4b0) {
                W0 = W1; W1 = W2; W2 = W3; W3 = W4; W4 = W5;
            }
            uint32_t o[4] = {__funnelshift_r(W0, W1, s1), __funnelshift_r(W1, W2, s1), __funnelshift_r(W2, W3, s1),
                             __funnelshift_r(W3, W4, s1)};
            const int64_t rem = n - (int64_t)j * kChunk;  // valid bytes from this slot on
            // flags of one sanitised word (all bytes <= 22): PAPA proline mask and charge class
            auto classify = [&](uint32_t w) -> uint32_t {
                uint32_t ext = w;
                if (kPro || (!kStd && adjust_prolines)) {
                    const uint32_t eq = bytes_eq7(w, 13u * 0x01010101u);
                    const uint32_t prev1 = __funnelshift_l(carry, eq, 8);   // proline one position earlier
                    const uint32_t prev2 = __funnelshift_l(carry, eq, 16);  // two positions earlier
                    carry = eq;
                    ext |= (eq & (prev1 | prev2)) >> 2;  // bit 7 -> bit 5 (kPapaMaskBit)
                }
                // charge class in bits 7:6 of every byte: 01 = +1, 11 = -1 (so (int8)byte >> 6 is the charge)
                uint32_t pl, mi;
                if (small_sets) {
                    pl = plus_run ? bytes_in_range(w, plus_lo, plus_hi) : bytes_eq7_2(w, cp0, cp1);
                    mi = bytes_eq7_2(w, cm0, cm1);
                } else {
                    pl = bytes_in_set(w, charge_plus);
                    mi = bytes_in_set(w, charge_minus);
                }
                return ext | (pl >> 1) | mi | (mi >> 1);
            };
            if (rem >= 16) {
                // a full slot (all but the last one or two of a protein): no padding, and one validity test for the
                // four words -- bit 7 of (byte | (low 7 bits + 0x6a)) is set exactly for bytes > 21
                const uint32_t t0 = o[0] | ((o[0] & 0x7f7f7f7fu) + 0x6a6a6a6au), t1 = o[1] | ((o[1] & 0x7f7f7f7fu) + 0x6a6a6a6au);
                const uint32_t t2 = o[2] | ((o[2] & 0x7f7f7f7fu) + 0x6a6a6a6au), t3 = o[3] | ((o[3] & 0x7f7f7f7fu) + 0x6a6a6a6au);
                if ((t0 | t1 | t2 | t3) & 0x80808080u) {  // invalid input -> X (0), flagged
                    bad_any = 1;
                    o[0] &= ~(((t0 & 0x80808080u) >> 7) * 0xffu);
                    o[1] &= ~(((t1 & 0x80808080u) >> 7) * 0xffu);
                    o[2] &= ~(((t2 & 0x80808080u) >> 7) * 0xffu);
                    o[3] &= ~(((t3 & 0x80808080u) >> 7) * 0xffu);
                }
#pragma unroll
                for (int q = 0; q < 4; q++) o[q] = classify(o[q]);
            } else {
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    uint32_t w = o[q];
                    const int64_t rq = rem - 4 * q;
                    const uint32_t vmask = rq >= 4 ? 0xffffffffu : (rq <= 0 ? 0u : ((1u << (8 * (int)rq)) - 1u));
                    // sanitise: bytes > 21 are invalid input -> X (0), flagged
                    const uint32_t bad = ((w & 0x80808080u) | (((w & 0x7f7f7f7fu) + 0x6a6a6a6au) & 0x80808080u)) & vmask;
                    if (bad) {
                        bad_any = 1;
                        w &= ~((bad >> 7) * 0xffu);
                    }
                    o[q] = classify((w & vmask) | (kPadW & ~vmask));
                }
            }
            __stcs(dst + (size_t)j * 32, make_uint4(o[0], o[1], o[2], o[3]));  // streaming: keep L2 for the reads
            A = B;
            B = C;
        }
        if (bad_any) atomicOr(errflag, 1);
    }
}

}  // namespace plaac
