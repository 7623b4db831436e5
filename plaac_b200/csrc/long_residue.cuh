// Long-sequence path, per-residue mode (plotsomefastas :610-647 on titin-length proteins): the posterior columns and the
// MAP parse of proteins the bucketed kernels would walk in a single lane (posteriorl :3349-3411, mapdecodel :4032-4045).
//
// ONE THREAD-BLOCK CLUSTER per long protein (1 or 8 CTAs x 512 lanes), one lane per chunk of >= 32 residues.  Forward and
// backward LUT recurrences both get the binade-frame treatment of long_kernel.cuh (pass 1: chunk-local frame after a
// warm-up, prefix / suffix sums of the chunk increments; pass 2: two frames one ulp apart in the jar's own binade), then
//
//   carry    one lane per direction walks the chunks in order and accepts the frame that enters the chunk an even number of
//            ulps from the exact value with exactly the bits of d = x1 - x0 the previous chunk left with (else it redoes the
//            chunk sequentially: binade crossings, |x| < 1024, uncoalesced warm-ups).  The exact state at every chunk
//            boundary is kept.  Across the CTAs of the cluster the walk is a relay: the CTA's walker hands its exact state to
//            the next CTA through DISTRIBUTED SHARED MEMORY (forward: rank r -> r+1, backward: r -> r-1), one cluster
//            barrier per stage; the approximate prefix sums of pass 1 get their cross-CTA carry the same way.
//   pass 3   every lane re-runs its chunk from the EXACT boundary state: backward first (b into scratch), then forward,
//            which leaves a + b per state in scratch.  These are the jar's a[][] and b[][] bit for bit.
//   pass 4   all threads, coalesced: pp = exp((a + b) - lpseq) (:3401-3405), MAP byte = pp1 > pp0.
//
// Small chunks make the sequential part cheap: a binade crossing costs its 2-3 chunks of 32-200 residues instead of a
// 288-residue chunk, and chunks whose warm-up reaches the protein's start (end) run the true chain and are exact as they are.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "long_kernel.cuh"
#include "summary_kernel_v2.cuh"

namespace plaac {
namespace cg = cooperative_groups;

constexpr int kLpThreads = 512;   // chunk lanes per CTA
constexpr int kLpMinChunk = 32;
constexpr int kLpBigCluster = 8;  // CTAs per protein of at least kLpBigMin residues
constexpr int kLpBigMin = 24576;
constexpr int kLpBndStride = 2 * (kLpBigCluster * kLpThreads + 8);  // doubles of boundary scratch per listed protein

struct LongPostArgs {
    const uint8_t* codes;
    const int64_t* offsets;
    int64_t off_base;            // codes[offsets[p] - off_base] is residue 0 of protein p
    int64_t res_base;            // out.*[offsets[p] - res_base] likewise
    const int32_t* list;         // k_long_select's list
    const int64_t* scratch_off;  // ... and scratch cursor (>= n + 128 per protein, multiple of 128)
    KScalars ks;
    const DeviceTables* tabs;
    plaac_residue_out out;
    double* S0;                  // scratch planes, indexed scratch_off + t: b0 then a0 + b0
    double* S1;
    double* bnd;                 // per listed protein kLpBndStride doubles: approximate a0 / b0 at the chunk boundaries
    double* lpseq;               // per listed protein
    int warm;                    // warm-up residues (rounded up to whole chunks)
    int cluster;                 // CTAs per protein of this launch; the launch handles proteins of its size class only
    int64_t big_min;             // class boundary: n >= big_min belongs to the kLpBigCluster launch
    unsigned long long* redone;  // statistics: chunks redone sequentially although they had frames
};

struct LpShared {
    double2 lut2[PLAAC_LUT_LEN + 1];
    double2 le[32];
    // pass 1: chunk increments and their local scans
    double inc[2][kLpThreads];
    // pass 2 per direction (0 forward, 1 backward): two frames
    double g_inc[2][2][kLpThreads], g_ea0[2][2][kLpThreads], g_den[2][2][kLpThreads], g_dex[2][2][kLpThreads];
    // exact state at the chunk's entry (forward: at cs-1, backward: at ce), filled by the carry walk
    double ent0[2][kLpThreads], ent1[2][kLpThreads];
    unsigned char mode[2][kLpThreads];  // 0 frames, 1 no frames (redo), 2 true chain (g_ea0/g_den[.][0] = exact entry, [.][1] = exact exit)
    double tot[2];                      // pass 1: this CTA's total increment per direction
    double carry[2][2];                 // relay: exact state handed over by the neighbour CTA
    int redone;
};

__device__ __forceinline__ void lp_fwd_step(double& a0, double& a1, const double2 le, const KScalars& ks, uint32_t lut)
{
    // a[i][t] = LSE_k(lt[k][i] + a[k][t-1]) + le[i][aa[t]], k ascending (:3359-3367)
    const double f0 = lse_lut2<false>(ks.lt00 + a0, ks.lt10 + a1, lut) + le.x;
    const double f1 = lse_lut2<false>(ks.lt01 + a0, ks.lt11 + a1, lut) + le.y;
    a0 = f0;
    a1 = f1;
}
__device__ __forceinline__ void lp_bwd_step(double& b0, double& b1, const double2 le_next, const KScalars& ks, uint32_t lut)
{
    // b[i][t] = LSE_k((lt[i][k] + b[k][t+1]) + le[k][aa[t+1]]), k ascending (:3384-3389)
    const double x0 = (ks.lt00 + b0) + le_next.x, x1 = (ks.lt01 + b1) + le_next.y;
    const double y0 = (ks.lt10 + b0) + le_next.x, y1 = (ks.lt11 + b1) + le_next.y;
    b0 = lse_lut2<false>(x0, x1, lut);
    b1 = lse_lut2<false>(y0, y1, lut);
}
// ulp of the binade of |x| (|x| >= 1024 wherever this is used)
__device__ __forceinline__ double lp_ulp(double x)
{
    return __hiloint2double((((__double2hiint(fabs(x)) >> 20) & 0x7ff) - 52) << 20, 0);
}
// does a value that moves between lo_v and hi_v (any order) leave its binade, margin included?
__device__ __forceinline__ bool lp_cross(double v0, double v1)
{
    const double lo = fmin(fabs(v0), fabs(v1)) - 64.0, hi = fmax(fabs(v0), fabs(v1)) + 64.0;
    return !(lo >= 1024.0) || (__double2hiint(lo) >> 20) != (__double2hiint(hi) >> 20);
}

__global__ void __launch_bounds__(kLpThreads, 1) k_long_post(LongPostArgs g)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int X = g.cluster;
    const int pslot = blockIdx.x / X;
    const int rank = (int)cluster.block_rank();
    const int32_t prot = g.list[pslot];
    const int64_t o = g.offsets[prot];
    const int n = (int)(g.offsets[prot + 1] - o);
    // size classes: the whole cluster leaves together
    if ((X > 1) != (n >= g.big_min)) return;

    extern __shared__ __align__(16) unsigned char lp_smem[];
    LpShared& sm = *reinterpret_cast<LpShared*>(lp_smem);
    const KScalars& ks = g.ks;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint8_t* __restrict__ src = g.codes + (o - g.off_base);
    const int64_t so = g.scratch_off[pslot];
    double* __restrict__ S0 = g.S0 + so;
    double* __restrict__ S1 = g.S1 + so;
    double* __restrict__ gA = g.bnd + (size_t)pslot * kLpBndStride;  // gA[j]: approximate a0 at residue j*C - 1 (j = 1..K)
    double* __restrict__ gB = gA + kLpBndStride / 2;                 // gB[j]: approximate b0 at residue j*C     (j = 0..K-1)

    // geometry: C residues per chunk, K chunks over the cluster's lanes, warm-up of m whole chunks
    const int lanes = X * kLpThreads;
    const int C = max(kLpMinChunk, (n + lanes - 1) / lanes);
    const int K = (n + C - 1) / C;
    const int m = (max(1, g.warm) + C - 1) / C;
    const int R = (K + kLpThreads - 1) / kLpThreads;  // CTAs that own chunks

    {
        const DeviceTables* T = g.tabs;
        for (int i = tid; i <= PLAAC_LUT_LEN; i += kLpThreads) {
            const double l0 = i < PLAAC_LUT_LEN ? T->lut[i] : 0.0;
            const double l1 = i + 1 < PLAAC_LUT_LEN ? T->lut[i + 1] : 0.0;
            sm.lut2[i] = make_double2(l0, l1);
        }
        if (tid < 32) sm.le[tid] = make_double2(T->le0[tid], T->le1[tid]);
        if (tid == 0) sm.redone = 0;
    }
    __syncthreads();
    const uint32_t lut = smem_u32(&sm.lut2[0]);
    // invalid input (> 21) is scored as X, as everywhere else (reported by k_pack)
    auto LE = [&](int t) -> double2 {
        const uint32_t c = src[t];
        return sm.le[c > 21u ? 0u : c];
    };

    const int kg = rank * kLpThreads + tid;  // this lane's chunk
    const bool live = kg < K;
    const int cs = kg * C, ce = min(n, cs + C);
    // forward: state initialised AT residue fi (true init when fi == 0), first step at fi + 1
    const int fi = (kg - m <= 0) ? 0 : (kg - m) * C - 1;
    // backward: state initialised AT residue bi (true init when bi == n-1), first step at bi - 1
    const int bi = (kg + m + 1 >= K) ? n - 1 : (kg + m + 1) * C;

    // ================= pass 1: chunk-local frame =================
    if (live) {
        {
            const double2 l0 = LE(fi);
            double a0 = ks.li0 + l0.x, a1 = ks.li1 + l0.y, e = 0.0;
#pragma unroll 4
            for (int t = fi + 1; t < ce; t++) {
                if (t == cs) e = a0;
                lp_fwd_step(a0, a1, LE(t), ks, lut);
            }
            if (fi + 1 >= ce && cs > 0) e = a0;  // (cannot happen: ce > cs >= fi + 1)
            sm.inc[0][tid] = a0 - e;  // first chunk: e = 0, absolute
        }
        {
            double b0 = ks.lf0, b1 = ks.lf1, e = 0.0;
#pragma unroll 4
            for (int t = bi - 1; t >= cs; t--) {
                if (t == ce - 1) e = b0;
                lp_bwd_step(b0, b1, LE(t + 1), ks, lut);
            }
            sm.inc[1][tid] = b0 - e;  // last chunk: e = 0, absolute
        }
    } else {
        sm.inc[0][tid] = 0.0;
        sm.inc[1][tid] = 0.0;
    }
    __syncthreads();
    // local scans: warp 0 inclusive prefix of the forward increments, warp 1 inclusive suffix of the backward ones
    if (wid < 2) {
        double carry = 0.0;
        for (int r = 0; r < kLpThreads / 32; r++) {
            const int idx = wid == 0 ? r * 32 + lane : kLpThreads - 1 - (r * 32 + lane);
            double v = sm.inc[wid][idx];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double u = __shfl_up_sync(0xffffffffu, v, d);
                if (lane >= d) v += u;
            }
            v += carry;
            sm.inc[wid][idx] = v;
            carry = __shfl_sync(0xffffffffu, v, 31);
        }
        if (lane == 0) sm.tot[wid] = carry;
    }
    cluster.sync();
    {
        // cross-CTA carry of the approximate sums through distributed shared memory
        double offF = 0.0, offB = 0.0;
        for (int r = 0; r < X; r++) {
            const double* t = cluster.map_shared_rank(&sm.tot[0], r);
            if (r < rank) offF += t[0];
            if (r > rank) offB += t[1];
        }
        if (live) {
            gA[kg + 1] = offF + sm.inc[0][tid];
            gB[kg] = offB + sm.inc[1][tid];
        }
    }
    cluster.sync();

    // ================= pass 2: the jar's own binade, two frames one ulp apart =================
    if (live) {
        // ---- forward
        if (fi == 0) {
            // the true chain: exact as it is
            const double2 l0 = LE(0);
            double a0 = ks.li0 + l0.x, a1 = ks.li1 + l0.y, e0 = 0.0, e1 = 0.0;
#pragma unroll 4
            for (int t = 1; t < ce; t++) {
                if (t == cs) e0 = a0, e1 = a1;
                lp_fwd_step(a0, a1, LE(t), ks, lut);
            }
            sm.mode[0][tid] = 2;
            sm.g_ea0[0][0][tid] = e0, sm.g_den[0][0][tid] = e1;
            sm.g_ea0[0][1][tid] = a0, sm.g_den[0][1][tid] = a1;
        } else if (lp_cross(gA[kg], gA[kg + 1])) {
            sm.mode[0][tid] = 1;
        } else {
            const double2 l0 = LE(fi);
            const double dd = (ks.li1 + l0.y) - (ks.li0 + l0.x);
            const double u = lp_ulp(gA[kg]);
            double a0 = gA[kg - m], a1 = a0 + dd, c0 = a0 + u, c1 = c0 + dd;
            double ea = 0, ed = 0, ec = 0, ee = 0;
#pragma unroll 2
            for (int t = fi + 1; t < ce; t++) {
                if (t == cs) {
                    ea = a0, ed = a1 - a0;
                    ec = c0, ee = c1 - c0;
                }
                const double2 le = LE(t);
                lp_fwd_step(a0, a1, le, ks, lut);
                lp_fwd_step(c0, c1, le, ks, lut);
            }
            sm.mode[0][tid] = 0;
            sm.g_inc[0][0][tid] = a0 - ea, sm.g_ea0[0][0][tid] = ea, sm.g_den[0][0][tid] = ed, sm.g_dex[0][0][tid] = a1 - a0;
            sm.g_inc[0][1][tid] = c0 - ec, sm.g_ea0[0][1][tid] = ec, sm.g_den[0][1][tid] = ee, sm.g_dex[0][1][tid] = c1 - c0;
        }
        // ---- backward
        if (bi == n - 1) {
            double b0 = ks.lf0, b1 = ks.lf1, e0 = 0.0, e1 = 0.0;
#pragma unroll 4
            for (int t = n - 2; t >= cs; t--) {
                if (t == ce - 1) e0 = b0, e1 = b1;
                lp_bwd_step(b0, b1, LE(t + 1), ks, lut);
            }
            sm.mode[1][tid] = 2;
            sm.g_ea0[1][0][tid] = e0, sm.g_den[1][0][tid] = e1;
            sm.g_ea0[1][1][tid] = b0, sm.g_den[1][1][tid] = b1;
        } else if (lp_cross(gB[kg + 1], gB[kg])) {
            sm.mode[1][tid] = 1;
        } else {
            const double dd = ks.lf1 - ks.lf0;
            const double u = lp_ulp(gB[kg + 1]);
            double a0 = gB[kg + m + 1], a1 = a0 + dd, c0 = a0 + u, c1 = c0 + dd;
            double ea = 0, ed = 0, ec = 0, ee = 0;
#pragma unroll 2
            for (int t = bi - 1; t >= cs; t--) {
                if (t == ce - 1) {
                    ea = a0, ed = a1 - a0;
                    ec = c0, ee = c1 - c0;
                }
                const double2 le = LE(t + 1);
                lp_bwd_step(a0, a1, le, ks, lut);
                lp_bwd_step(c0, c1, le, ks, lut);
            }
            sm.mode[1][tid] = 0;
            sm.g_inc[1][0][tid] = a0 - ea, sm.g_ea0[1][0][tid] = ea, sm.g_den[1][0][tid] = ed, sm.g_dex[1][0][tid] = a1 - a0;
            sm.g_inc[1][1][tid] = c0 - ec, sm.g_ea0[1][1][tid] = ec, sm.g_den[1][1][tid] = ee, sm.g_dex[1][1][tid] = c1 - c0;
        }
    }
    __syncthreads();

    // ================= carry: exact state at every chunk boundary, in chunk order, relayed over the cluster =================
    for (int stage = 0; stage < R; stage++) {
        const int dir = (tid == 0) ? 0 : (tid == 32 ? 1 : -1);
        const bool mine = dir == 0 ? (rank == stage) : (dir == 1 ? (rank == R - 1 - stage) : false);
        if (mine) {
            const int k_lo = rank * kLpThreads, k_hi = min(K, k_lo + kLpThreads);  // this CTA's chunks
            double x0 = sm.carry[dir][0], x1 = sm.carry[dir][1];                  // (unused by the first walker)
            int nredo = 0;
            for (int q = 0; q < k_hi - k_lo; q++) {
                const int k = dir == 0 ? k_lo + q : k_hi - 1 - q;
                const int l = k - k_lo;
                const int s = k * C, e = min(n, s + C);
                sm.ent0[dir][l] = x0;
                sm.ent1[dir][l] = x1;
                const int md = sm.mode[dir][l];
                if (md == 2) {
                    x0 = sm.g_ea0[dir][1][l];
                    x1 = sm.g_den[dir][1][l];
                    continue;
                }
                int fr = -1;
                if (md == 0) {
                    const double u = lp_ulp(x0);
                    const double dex = x1 - x0;
#pragma unroll
                    for (int f = 0; f < 2; f++) {
                        const double qd = (sm.g_ea0[dir][f][l] - x0) / u;  // exact: a small integer
                        if (fr < 0 && qd == 2.0 * rint(0.5 * qd) &&
                            __double_as_longlong(sm.g_den[dir][f][l]) == __double_as_longlong(dex))
                            fr = f;
                    }
                }
                if (fr >= 0) {
                    x0 = x0 + sm.g_inc[dir][fr][l];
                    x1 = x0 + sm.g_dex[dir][fr][l];
                } else {
                    if (dir == 0) {
#pragma unroll 4
                        for (int t = s; t < e; t++) lp_fwd_step(x0, x1, LE(t), ks, lut);
                    } else {
#pragma unroll 4
                        for (int t = e - 1; t >= s; t--) lp_bwd_step(x0, x1, LE(t + 1), ks, lut);
                    }
                    nredo += md == 0 ? 1 : 0;
                }
            }
            if (nredo) atomicAdd(&sm.redone, nredo);
            // hand the exact state to the neighbour
            const int next = dir == 0 ? rank + 1 : rank - 1;
            if (next >= 0 && next < R) {
                double* c = cluster.map_shared_rank(&sm.carry[dir][0], next);
                c[0] = x0;
                c[1] = x1;
            }
        }
        cluster.sync();
    }

    // ================= pass 3: every chunk from its exact boundary state =================
    if (live) {
        {
            double b0, b1;
            int t = ce - 1;
            if (kg == K - 1) {
                b0 = ks.lf0, b1 = ks.lf1;  // :3379-3381
                S0[n - 1] = b0;
                S1[n - 1] = b1;
                t = n - 2;
            } else {
                b0 = sm.ent0[1][tid], b1 = sm.ent1[1][tid];
            }
#pragma unroll 4
            for (; t >= cs; t--) {
                lp_bwd_step(b0, b1, LE(t + 1), ks, lut);
                S0[t] = b0;
                S1[t] = b1;
            }
        }
        {
            double a0, a1;
            int t = cs;
            if (kg == 0) {
                const double2 l0 = LE(0);
                a0 = ks.li0 + l0.x, a1 = ks.li1 + l0.y;  // :3356-3358
                const double s0 = a0 + S0[0], s1 = a1 + S1[0];
                g.lpseq[pslot] = lse_lut2<false>(s0, s1, lut);  // :3393-3396
                S0[0] = s0;
                S1[0] = s1;
                t = 1;
            } else {
                a0 = sm.ent0[0][tid], a1 = sm.ent1[0][tid];
            }
#pragma unroll 4
            for (; t < ce; t++) {
                lp_fwd_step(a0, a1, LE(t), ks, lut);
                S0[t] = a0 + S0[t];
                S1[t] = a1 + S1[t];
            }
        }
    }
    if (tid == 0 && sm.redone && g.redone) atomicAdd(g.redone, (unsigned long long)sm.redone);
    __threadfence();
    cluster.sync();

    // ================= pass 4: posteriors and MAP bytes, coalesced =================
    {
        const double lpseq = g.lpseq[pslot];
        const int64_t ob = o - g.res_base;
        const int per = (n + X - 1) / X;
        const int t_lo = rank * per, t_hi = min(n, t_lo + per);
        for (int t = t_lo + tid; t < t_hi; t += kLpThreads) {
            const double p0 = exp(S0[t] - lpseq), p1 = exp(S1[t] - lpseq);
            g.out.post_bg[ob + t] = p0;
            g.out.post_prd[ob + t] = p1;
            if (g.out.map) g.out.map[ob + t] = p1 > p0 ? 1 : 0;
        }
    }
}

// Viterbi bits of the long proteins (k_long_score's bit words) -> one byte per residue of out.vit
__global__ void __launch_bounds__(256)
k_long_vit_bytes(const int64_t* __restrict__ offsets, int64_t res_base, const int32_t* __restrict__ list,
                 const int64_t* __restrict__ scratch_off, const uint32_t* __restrict__ vit, uint8_t* __restrict__ out)
{
    const int32_t prot = list[blockIdx.y];
    const int64_t o = offsets[prot];
    const int n = (int)(offsets[prot + 1] - o);
    const uint32_t* w = vit + (scratch_off[blockIdx.y] >> 5);
    uint8_t* dst = out + (o - res_base);
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) dst[t] = (uint8_t)((w[t >> 5] >> (t & 31)) & 1u);
}

}  // namespace plaac
