// Device side of the transport-lean batch call (include/plaac_cuda.h, plaac_score_packed / plaac_score_hits):
//   k_unpack22      radix-22 words (7 residues per uint32) -> one-byte residue codes, HBM-bound byte work
//   k_hits_flag/... compaction of the records with a CORE for the ranked compact output
// The reference has no counterpart (plaac.java keeps everything on the Java heap); the order of the compact output is
// the web front end's (web/lib/server.rb:222-229), as in rank.cuh.
#pragma once
#include "common.cuh"
#include "rank.cuh"

namespace plaac {

constexpr int kUnpackThreads = 256;
constexpr int kUnpackWordsPerLane = 4;                                   // one 16-byte load
constexpr int kUnpackTileWords = 32 * kUnpackWordsPerLane;               // per warp and round: 128 words = 896 residues
constexpr int kUnpackTileBytes = kUnpackTileWords * PLAAC_PACK_PER_WORD;  // 896 = 56 x 16

// codes[7*w + k] = digit k of words[w].  A warp takes 128 consecutive words per round: one coalesced 16-byte load per
// lane, 28 digits per lane assembled into seven 32-bit words at shared-memory word 7*lane + k (stride 7: conflict
// free), then 56 coalesced 16-byte stores.  `codes` must be 16-byte aligned with room for 7*nwords bytes rounded up to
// a whole tile.  A word >= 22^7 is not a packed word: flagged, its first seven digits are used.
__global__ void __launch_bounds__(kUnpackThreads)
k_unpack22(const uint32_t* __restrict__ words, int64_t nwords, uint8_t* __restrict__ codes, int* __restrict__ errflag)
{
    __shared__ __align__(16) uint32_t sh[kUnpackThreads / 32][kUnpackTileBytes / 4];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t ntiles = (nwords + kUnpackTileWords - 1) / kUnpackTileWords;
    uint32_t* my = sh[wid];
    bool bad = false;
    for (int64_t tile = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); tile < ntiles; tile += warps) {
        const int64_t w0 = tile * kUnpackTileWords + (int64_t)lane * kUnpackWordsPerLane;
        uint32_t v[kUnpackWordsPerLane] = {0, 0, 0, 0};
        if (w0 + kUnpackWordsPerLane <= nwords) {
            const uint4 q = __ldcs(reinterpret_cast<const uint4*>(words + w0));
            v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
        } else {
#pragma unroll
            for (int k = 0; k < kUnpackWordsPerLane; k++)
                if (w0 + k < nwords) v[k] = words[w0 + k];
        }
        uint32_t out[7] = {0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < kUnpackWordsPerLane; k++) {
            uint32_t q = v[k];
            bad |= q >= 2494357888u;  // 22^7
#pragma unroll
            for (int d = 0; d < PLAAC_PACK_PER_WORD; d++) {
                const uint32_t nq = q / 22u;
                const uint32_t c = q - nq * 22u;
                q = nq;
                const int pos = k * PLAAC_PACK_PER_WORD + d;  // byte 0..27 of the lane's run
                out[pos >> 2] |= c << (8 * (pos & 3));
            }
        }
#pragma unroll
        for (int k = 0; k < 7; k++) my[7 * lane + k] = out[k];
        __syncwarp();
        uint4* dst = reinterpret_cast<uint4*>(codes + tile * kUnpackTileBytes);
        const uint4* src = reinterpret_cast<const uint4*>(my);
        dst[lane] = src[lane];
        if (lane < kUnpackTileBytes / 16 - 32) dst[32 + lane] = src[32 + lane];
        __syncwarp();
    }
    if (bad) atomicOr(errflag, 2);
}

__device__ __forceinline__ bool rec_has_core(const plaac_summary& r) { return r.core_score == r.core_score; }

__global__ void __launch_bounds__(256) k_hits_flag(const plaac_summary* __restrict__ rec, int64_t n, int32_t* __restrict__ flag)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = rec_has_core(rec[i]) ? 1 : 0;
}

// idx[pos[i]] = i for the flagged rows (pos = exclusive scan of flag): stable, so equal rows keep input order.
__global__ void __launch_bounds__(256)
k_hits_compact(const int32_t* __restrict__ flag, const int64_t* __restrict__ pos, int64_t n, int32_t* __restrict__ idx)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) idx[pos[i]] = (int32_t)i;
}

// keys of the rows listed in vals: field 0 LLR, field 1 COREscore (rank.cuh's order-preserving map)
__global__ void __launch_bounds__(256)
k_hits_keys(const plaac_summary* __restrict__ rec, const int32_t* __restrict__ vals, int64_t n, int field,
            uint64_t* __restrict__ keys)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const plaac_summary& r = rec[vals[i]];
    keys[i] = rank_desc_key(field ? r.core_score : r.llr);
}

// plaac_score_multi_packed: indices of a shard's rows become indices of the whole batch
__global__ void __launch_bounds__(256) k_hits_add_base(int32_t* __restrict__ idx, int64_t n, int32_t base)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) idx[i] += base;
}

__global__ void __launch_bounds__(256)
k_hits_gather_idx(const int32_t* __restrict__ idx, const int32_t* __restrict__ order, int64_t n, int32_t* __restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = idx[order[i]];
}

}  // namespace plaac
