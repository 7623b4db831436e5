// On-device ranking of the summary records.
//
// The reference's only consumer that orders the table is the web front end: it re-sorts the rows by COREscore
// descending, then LLR descending, NaNs last (web/lib/server.rb:222-229) and paginates.  At proteome scale the
// consumer wants the head of that order, not 160 B x 100 M records, so the order is computed where the records are:
//
//   1. LSD radix sort of (key(LLR), index) over all proteins            8 passes x 8 bits, stable
//   2. one stable 2-way split: proteins with a CORE first               (same pass kernel, 1-bit digit)
//   3. LSD radix sort of the CORE proteins by key(COREscore)            8 passes over ~5 % of the rows
//
// Stability makes the composite order (COREscore desc, LLR desc, input index asc).  All passes are HBM-bound
// integer work: a pass reads 12 B and writes 12 B per row.  Keys are the usual order-preserving map of an IEEE
// double to uint64, complemented for descending order; NaN maps to the largest key.
#pragma once
#include "common.cuh"

namespace plaac {

constexpr int kRankThreads = 256;
constexpr int kRankWarps = kRankThreads / 32;
constexpr int kRankRounds = 16;                                // 32 consecutive rows per warp and round
constexpr int kRankTile = kRankThreads * kRankRounds;          // rows per CTA
constexpr int kRankDigits = 256;

__device__ __forceinline__ uint64_t rank_desc_key(double v)
{
    if (v != v) return ~0ull;  // NaN last
    const uint64_t u = (uint64_t)__double_as_longlong(v);
    const uint64_t asc = (u >> 63) ? ~u : (u | 0x8000000000000000ull);
    return ~asc;               // no finite or infinite double reaches ~0 (that would need u == all ones, a NaN)
}

// field 0: LLR (column 4), field 1: COREscore gathered through vals.  (LLR = -Inf, a protein shorter than the core
// length, sorts last: the table prints it as NaN and server.rb:225 puts "NaN" last.)
__global__ void __launch_bounds__(256)
k_rank_keys(const plaac_summary* __restrict__ rec, int64_t n, int field, uint64_t* __restrict__ keys,
            int32_t* __restrict__ vals)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (field == 0) {
        keys[i] = rank_desc_key(rec[i].llr);
        vals[i] = (int32_t)i;
    } else {
        keys[i] = rank_desc_key(rec[vals[i]].core_score);
    }
}

// SPLIT: digit = "has no CORE" (key == ~0), else byte `shift/8` of the key.
template <bool SPLIT>
__device__ __forceinline__ int rank_digit(uint64_t key, int shift)
{
    if (SPLIT) return key == ~0ull ? 1 : 0;
    return (int)((key >> shift) & 255u);
}

// hist[d * ntiles + tile] = rows of `tile` with digit d (digit-major, so one exclusive scan gives every
// (digit, tile) its output base).
template <bool SPLIT>
__global__ void __launch_bounds__(kRankThreads)
k_rank_hist(const uint64_t* __restrict__ keys, int64_t n, int shift, int32_t* __restrict__ hist, int64_t ntiles)
{
    __shared__ int32_t h[kRankDigits];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kRankTile;
#pragma unroll 4
    for (int r = 0; r < kRankRounds; r++) {
        const int64_t i = base + (int64_t)r * kRankThreads + threadIdx.x;
        if (i < n) atomicAdd(&h[rank_digit<SPLIT>(keys[i], shift)], 1);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

// Stable scatter of one tile.  Warp w owns rows [w*512, (w+1)*512) of the tile and walks them 32 at a time;
// __match_any_sync groups the lanes of a round by digit, the lowest lane of a group is its leader.
template <bool SPLIT>
__global__ void __launch_bounds__(kRankThreads)
k_rank_scatter(const uint64_t* __restrict__ keys_in, const int32_t* __restrict__ vals_in, uint64_t* __restrict__ keys_out,
               int32_t* __restrict__ vals_out, int64_t n, int shift, const int64_t* __restrict__ offs, int64_t ntiles)
{
    __shared__ int32_t wh[kRankWarps][kRankDigits];
    __shared__ int64_t wbase[kRankWarps][kRankDigits];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int i = tid; i < kRankWarps * kRankDigits; i += kRankThreads) (&wh[0][0])[i] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kRankTile + (int64_t)w * (32 * kRankRounds);
    const uint32_t lt = (1u << lane) - 1u;
    for (int r = 0; r < kRankRounds; r++) {
        const int64_t i = base + r * 32 + lane;
        const int d = i < n ? rank_digit<SPLIT>(keys_in[i], shift) : kRankDigits;
        const uint32_t m = __match_any_sync(0xffffffffu, d);
        if (d < kRankDigits && (m & lt) == 0) wh[w][d] += __popc(m);
        __syncwarp();
    }
    __syncthreads();
    {
        int64_t run = offs[(int64_t)tid * ntiles + blockIdx.x];
#pragma unroll
        for (int k = 0; k < kRankWarps; k++) {
            wbase[k][tid] = run;
            run += wh[k][tid];
        }
    }
    __syncthreads();
    for (int r = 0; r < kRankRounds; r++) {
        const int64_t i = base + r * 32 + lane;
        uint64_t key = 0;
        int d = kRankDigits;
        if (i < n) {
            key = keys_in[i];
            d = rank_digit<SPLIT>(key, shift);
        }
        const uint32_t m = __match_any_sync(0xffffffffu, d);
        if (d < kRankDigits) {
            const int64_t pos = wbase[w][d] + __popc(m & lt);
            keys_out[pos] = key;
            vals_out[pos] = vals_in[i];
        }
        __syncwarp();
        if (d < kRankDigits && (m & lt) == 0) wbase[w][d] += __popc(m);
        __syncwarp();
    }
}

// Compact output: out[k] = rec[order[k]], k < count.  20 lanes move one 160-byte record as 8-byte words.
__global__ void __launch_bounds__(256)
k_rank_gather(const plaac_summary* __restrict__ rec, const int32_t* __restrict__ order, int64_t count,
              plaac_summary* __restrict__ out)
{
    constexpr int kWords = sizeof(plaac_summary) / 8;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t k = t / kWords;
    const int wd = (int)(t % kWords);
    if (k >= count) return;
    const uint64_t* src = reinterpret_cast<const uint64_t*>(rec + order[k]);
    reinterpret_cast<uint64_t*>(out + k)[wd] = src[wd];
}

}  // namespace plaac
