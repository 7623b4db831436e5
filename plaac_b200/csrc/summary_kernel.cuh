// Summary-mode scoring kernel: one lane per protein, one warp per length bucket.
//
// Follows plaac.java in REFERENCE OPERATION ORDER for everything whose rounding the output depends on:
//   Viterbi (viterbidecodel :3077-3121), LUT forward (posteriorl :3349-3375 with logeapeb :1024-1047),
//   hmm0's sequential log-emission sum, the sequential psum[] of hss2 (:1206-1257) for LLR and for the
//   -1e6-masked CORE search (:816-833), the PRD expansion and score (:851-873), longestrun (:1787).
// The FoldIndex / PAPA / smoothed-LLR tracks (disorderreport :4866-5068, slidingaverage :2585-2662) are
// evaluated as running window sums instead of 41-tap loops: same quantities, different rounding
// (|delta| ~ 1e-13, DESIGN.md "tolerances").
//
// All lanes of a warp advance the residue index t in lock step, so every ring-buffer address is
// warp-uniform and conflict-free; per-lane work ends when t passes the lane's own length.
#pragma once
#include "common.cuh"

namespace plaac {

struct SummarySmem {
    double lut[PLAAC_LUT_LEN + 3];
    double le0[kTabN], le1[kTabN], lebg[kTabN], llr[kTabN], hyd[kTabN], pap[kTabN], hydw[kTabN];
    double rcp[kTabN * 4];  // 1/cnt for cnt in [0, 255]; only used when 2w+1 <= 255
};

// logeapeb, plaac.java:1024-1047, branch-free but with the reference's exact arithmetic.
__device__ __forceinline__ double lse_lut(double a, double b, const double* __restrict__ lut, double ln2)
{
    const bool gt = a > b;
    const bool lt = b > a;
    const double hi = gt ? a : b;
    const double lo = gt ? b : a;
    const double c = hi - lo;
    const bool in = c < 40.0;
    const double x = 100.0 * c;
    int dex = in ? __double2int_rd(x) : 0;
    const double l1 = lut[dex + 1];
    const double l0 = lut[dex];
    const double f1 = x - (double)dex;
    const double f0 = (double)(dex + 1) - x;
    double r = hi + (f1 * l1 + f0 * l0);
    r = in ? r : hi;
    return (gt || lt) ? r : (a + ln2);
}

__device__ __forceinline__ int code_charge(uint32_t plus, uint32_t minus, int c)
{
    return (int)((plus >> c) & 1u) - (int)((minus >> c) & 1u);
}

// Rare tail (about 5 % of proteins): the -1e6-masked CORE window search, PRD expansion and PRD score,
// all in reference order.  Reads the lane's own stream bytes and Viterbi bits back from global memory.
__device__ __noinline__ void core_search(const uint8_t* __restrict__ sbytes /* lane slot 0 */,
                                         const uint32_t* __restrict__ vitw /* lane word 0 */, int n, int c,
                                         const double* __restrict__ llr_tab, double big_neg, int& core_start,
                                         int& core_end, double& core_score, int& prd_start, int& prd_end,
                                         double& prd_score)
{
    // element i of the lane: byte (i>>4)*512 + (i&15) ; Viterbi bit: word (i>>4)*32, bit (i&15)
    auto code_at = [&](int i) -> int { return sbytes[(size_t)(i >> 4) * 512 + (i & 15)] & 31; };
    auto vit_at = [&](int i) -> int { return (vitw[(size_t)(i >> 4) * 32] >> (i & 15)) & 1; };
    double ps = 0.0, psl = 0.0, best = 0.0;
    int bstop = c - 1;
    for (int i = 0; i < n; i++) {
        double x = vit_at(i) ? llr_tab[code_at(i)] : big_neg;
        ps = ps + x;
        if (i >= c) {
            int k = i - c;
            double xl = vit_at(k) ? llr_tab[code_at(k)] : big_neg;
            psl = psl + xl;
        }
        if (i == c - 1) {
            best = ps;
        } else if (i >= c) {
            double d = ps - psl;
            if (d > best) {
                best = d;
                bstop = i;
            }
        }
    }
    if (best > big_neg / 2) {
        int s = bstop - c + 1, e = bstop;
        core_start = s;
        core_end = e;
        core_score = best;
        int a0 = s, a1 = e;
        while (a0 >= 0 && vit_at(a0) == 1) a0--;
        a0++;
        while (a1 < n && vit_at(a1) == 1) a1++;
        a1--;
        double sc = 0.0;
        for (int k = a0; k <= a1; k++) sc = sc + llr_tab[code_at(k)];
        prd_start = a0;
        prd_end = a1;
        prd_score = sc;
    }
}

__global__ void __launch_bounds__(384, 1)
k_score_summary(BatchView bv, KScalars ks, const DeviceTables* __restrict__ tabs, plaac_summary* __restrict__ out,
                int ring_words /* power of two, words per lane */)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SummarySmem& S = *reinterpret_cast<SummarySmem*>(smem_raw);
    uint32_t* ring_all = reinterpret_cast<uint32_t*>(smem_raw + sizeof(SummarySmem));

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    for (int i = tid; i < PLAAC_LUT_LEN + 3; i += blockDim.x) S.lut[i] = tabs->lut[i];
    for (int i = tid; i < kTabN; i += blockDim.x) {
        S.le0[i] = tabs->le0[i];
        S.le1[i] = tabs->le1[i];
        S.lebg[i] = tabs->lebg[i];
        S.llr[i] = tabs->llr[i];
        S.hyd[i] = tabs->hyd[i];
        S.hydw[i] = tabs->hydw[i];
        S.pap[i] = tabs->pap[i];
    }
    for (int i = tid; i < kTabN * 4; i += blockDim.x) S.rcp[i] = i > 0 ? 1.0 / (double)i : 0.0;
    __syncthreads();

    const int64_t b = (int64_t)blockIdx.x * nw + wid;
    if (b >= bv.nbuckets) return;

    uint32_t* ring = ring_all + (size_t)wid * ring_words * 32 + lane;
    const int rmask = ring_words - 1;
    constexpr uint32_t kPadW = 0x01010101u * kPad;
    for (int i = 0; i < ring_words; i++) ring[i * 32] = kPadW;

    const int64_t rank = b * 32 + lane;
    int n = 0;
    int32_t prot = -1;
    if (rank < bv.nprot) {
        prot = bv.order[rank];
        n = (int)(bv.offsets[prot + 1] - bv.offsets[prot]);
    }
    const int64_t cb = bv.chunk_base[b];
    const int nch = (int)(bv.chunk_base[b + 1] - cb);
    const uint4* sp = bv.stream + cb * 32 + lane;
    uint32_t* tbp = bv.tbw + cb * 32 + lane;

    const int w = ks.w, c = ks.core_len, mw = ks.mw_window;
    const int off1 = 2 * w + 1, off2 = 4 * w + 2;
    int nmax = n;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, d));
    const int t_end = nmax + w;  // exclusive

    // ---- per-lane state -------------------------------------------------------------------------
    double s0 = 0, s1 = 0, a0 = 0, a1 = 0, sum0 = 0;  // Viterbi, forward, hmm0
    double ps = 0, psl = 0, llr_best = -INFINITY;      // LLR psum / lagged psum
    int llr_stop = -2;
    double sh = 0;  // sequential sum of hydro2 (meanhydro)
    int csum = 0;   // sum of charges
    int qn = 0, mw_best = 0, mw_stop = -1;
    double SLh = 0, SLl = 0, SLp = 0, SGh = 0, SGl = 0, SGp = 0;  // lead / lag window sums
    int SLc = 0, SGc = 0;
    double Th = 0, Tl = 0, Tp = 0;  // pass-2 running sums
    int Tac = 0, W = 0;
    double pbest = -INFINITY, pfix = 0, pllr2 = 0;
    int pcen = -1;
    int halfw = ks.h_fi;
    if (halfw > n / 2) halfw = n / 2;
    int fi_run_start = -1, fi_numaa = 0, fi_maxrun = 0;
    uint32_t tbacc = 0;

    uint4 nxt = make_uint4(kPadW, kPadW, kPadW, kPadW);
    if (nch > 0) nxt = sp[0];

    for (int t = 0; t < t_end; t++) {
        if ((t & 15) == 0) {
            const int j = t >> 4;
            const int wi = (j * 4) & rmask;
            ring[(wi + 0) * 32] = nxt.x;
            ring[(wi + 1) * 32] = nxt.y;
            ring[(wi + 2) * 32] = nxt.z;
            ring[(wi + 3) * 32] = nxt.w;
            nxt = make_uint4(kPadW, kPadW, kPadW, kPadW);
            if (j + 1 < nch) nxt = sp[(size_t)(j + 1) * 32];
        }
        auto rd = [&](int pos) -> int {
            return (int)((ring[((pos >> 2) & rmask) * 32] >> ((pos & 3) * 8)) & 0xffu);
        };
        const int e0 = rd(t);
        const int e1 = rd(t - off1);
        const int e2 = rd(t - off2);
        const int c0 = e0 & 31, c1 = e1 & 31, c2 = e2 & 31;
        const double lr0 = S.llr[c0];
        const double hy0 = S.hyd[c0];

        // ---------------- HMMs, LLR window, MW window, means (residue t) ----------------
        if (t < n) {
            const double le0 = S.le0[c0], le1 = S.le1[c0], lb = S.lebg[c0];
            if (t == 0) {
                s0 = ks.li0 + le0;
                s1 = ks.li1 + le1;
                a0 = s0;
                a1 = s1;
                sum0 = lb;  // hmm0: liprob[0] + le = 0 + le
            } else {
                // viterbidecodel :3088-3101 (state 1 wins only on strict >)
                const double v00 = ks.lt00 + s0, v10 = ks.lt10 + s1;
                const double v01 = ks.lt01 + s0, v11 = ks.lt11 + s1;
                const bool tb0 = v10 > v00, tb1 = v11 > v01;
                s0 = (tb0 ? v10 : v00) + le0;
                s1 = (tb1 ? v11 : v01) + le1;
                tbacc |= ((uint32_t)tb0 | ((uint32_t)tb1 << 1)) << ((t & 15) * 2);
                // posteriorl forward :3359-3367
                const double f0 = lse_lut(ks.lt00 + a0, ks.lt10 + a1, S.lut, ks.ln2) + le0;
                const double f1 = lse_lut(ks.lt01 + a0, ks.lt11 + a1, S.lut, ks.ln2) + le1;
                a0 = f0;
                a1 = f1;
                sum0 = sum0 + lb;
            }
            ps = ps + lr0;  // hss2 psum :1230-1233
            sh = sh + hy0;  // mean() :1584
            csum += code_charge(ks.charge_plus, ks.charge_minus, c0);
        }
        if ((t & 15) == 15 || t == t_end - 1) {
            if ((t >> 4) < nch) tbp[(size_t)(t >> 4) * 32] = tbacc;
            tbacc = 0;
        }
        {
            const int ec = rd(t - c) & 31;
            psl = psl + S.llr[ec];  // == psum[t-c+1]
            const int em = rd(t - mw) & 31;
            qn += (int)((ks.qn_mask >> c0) & 1u) - (int)((ks.qn_mask >> em) & 1u);
            if (t < n) {
                if (t >= c - 1) {
                    const double d = ps - psl;
                    if (t == c - 1 || d > llr_best) {
                        llr_best = d;
                        llr_stop = t;
                    }
                }
                if (t >= mw - 1) {
                    if (t == mw - 1 || qn > mw_best) {
                        mw_best = qn;
                        mw_stop = t;
                    }
                } else if (t == n - 1) {  // n < mw: one window = the whole protein (:766-767)
                    mw_best = qn;
                    mw_stop = t;
                }
            }
        }

        // ---------------- sliding windows ----------------
        {
            const double hy0w = S.hydw[c0], hy1 = S.hydw[c1], hy2 = S.hydw[c2];  // window sums: exact on the grid
            const double lr1 = S.llr[c1], lr2 = S.llr[c2];
            const double pa0 = S.pap[e0 & 63], pa1 = S.pap[e1 & 63], pa2 = S.pap[e2 & 63];
            const int ch0 = code_charge(ks.charge_plus, ks.charge_minus, c0);
            const int ch1 = code_charge(ks.charge_plus, ks.charge_minus, c1);
            const int ch2 = code_charge(ks.charge_plus, ks.charge_minus, c2);
            SLh = (SLh + hy0w) - hy1;
            SLl = (SLl + lr0) - lr1;
            SLp = (SLp + pa0) - pa1;
            SLc += ch0 - ch1;
            SGh = (SGh + hy1) - hy2;
            SGl = (SGl + lr1) - lr2;
            SGp = (SGp + pa1) - pa2;
            SGc += ch1 - ch2;
        }
        const int p = t - w;  // lead centre: SL* = window sums of [p-w, p+w]
        if (p >= 0) {         // warp-uniform
            const int cnt = min(n - 1, t) - max(0, t - 2 * w) + 1;  // taps inside [0,n) (slidingaverage :2604-2606)
            if (p < n) {
                Th += SLh;
                Tl += SLl;
                Tp += SLp;
                Tac += abs(SLc);
                W += cnt;
                // FoldIndex run scan :5010-5059 over i in [halfw, n-halfw)
                if (p >= halfw && p < n - halfw) {
                    const double hyd = SLh / (double)cnt;
                    const double chg = (double)SLc / (double)cnt;
                    const double fi = (ks.cc0 * hyd + ks.cc1 * fabs(chg)) + ks.cc2;
                    const bool neg = fi < 0;
                    if (neg && fi_run_start < 0) fi_run_start = (p == halfw) ? 0 : p;
                    const bool last = (p == n - halfw - 1);
                    if (fi_run_start >= 0 && (!neg || last)) {
                        const int stop = neg ? (n - 1) : (p - 1);  // neg here implies last
                        const int len = stop - fi_run_start + 1;
                        if (len >= 5) {
                            fi_numaa += len;
                            fi_maxrun = max(fi_maxrun, len);
                        }
                        fi_run_start = -1;
                    }
                }
            }
            const int q = p - off1;  // lag centre leaving the pass-2 window
            if (q >= 0) {            // warp-uniform
                if (q < n) {
                    const int cntg = min(n - 1, q + w) - max(0, q - w) + 1;
                    Th -= SGh;
                    Tl -= SGl;
                    Tp -= SGp;
                    Tac -= abs(SGc);
                    W -= cntg;
                }
            }
            // pass-2 centre k = p - w: T* cover pass-1 centres [k-w, k+w] (slidingaverage weight=true)
            const int k = p - w;
            if (k >= w && k <= n - w - 1) {
                const double Wd = (double)W;
                const double papax2 = Tp / Wd;
                const double vfi = (ks.cc0 * Th + ks.cc1 * (double)Tac) + ks.cc2 * Wd;
                if (papax2 > pbest && vfi < 0) {  // :4943 (fix2[k] < 0  <=>  numerator < 0)
                    pbest = papax2;
                    pcen = k;
                    pfix = vfi / Wd;
                    pllr2 = Tl / Wd;
                }
            }
        }
    }

    if (prot < 0) return;
    plaac_summary r;
    r.prot_len = n;
    if (n < 1) {
        // the jar prints no row (:762); emit a zeroed record with prot_len 0
        r.mw_score = r.mw_start = r.mw_end = r.llr_start = r.llr_end = r.vit_maxrun = 0;
        r.core_start = r.core_end = r.prd_start = r.prd_end = r.fi_numaa = r.fi_maxrun = r.papa_center = 0;
        r.llr = r.core_score = r.prd_score = r.hmm_all = r.hmm_vit = 0;
        r.fi_meanhydro = r.fi_meancharge = r.fi_meancombo = 0;
        r.papa_combo = r.papa_prop = r.papa_fi = r.papa_llr = r.papa_llr2 = 0;
        out[prot] = r;
        return;
    }

    // ---- MW (:764-771): single window when n < 80
    r.mw_start = (n < mw) ? 0 : mw_stop - mw + 1;
    r.mw_score = mw_best;
    r.mw_end = mw_stop;
    // ---- LLR (:782-783)
    if (n < c) {
        r.llr = -INFINITY;
        r.llr_start = -1;
        r.llr_end = -2;
    } else {
        r.llr = llr_best;
        r.llr_start = llr_stop - c + 1;
        r.llr_end = llr_stop;
    }
    // ---- HMM scores (:797-798, :3102-3109, :3369-3375)
    const double e0v = s0 + ks.lf0, e1v = s1 + ks.lf1;
    const int vlast = e1v > e0v ? 1 : 0;
    const double lvit = vlast ? e1v : e0v;
    const double lmarg = lse_lut(a0 + ks.lf0, a1 + ks.lf1, S.lut, ks.ln2);
    r.hmm_all = lmarg - sum0;
    r.hmm_vit = lvit - sum0;
    // ---- FoldIndex means (:4876-4883)
    const double mh = (1.0 * sh) / (double)n;
    const double mc = (1.0 * (double)csum) / (double)n;
    r.fi_meanhydro = mh;
    r.fi_meancharge = mc;
    r.fi_meancombo = (ks.cc2 + ks.cc1 * fabs(mc)) + ks.cc0 * mh;
    r.fi_numaa = fi_numaa;
    r.fi_maxrun = fi_maxrun;
    // ---- PAPA (:4931-4997)
    r.papa_center = pcen;
    r.papa_combo = pbest;
    if (pcen >= 0) {
        r.papa_prop = pbest;
        r.papa_fi = pfix;
        r.papa_llr2 = pllr2;
        // plaacllr[pcen]: 2w+1 taps in reference order (:2604-2620); pcen is interior so all taps are in range
        const uint8_t* sb = reinterpret_cast<const uint8_t*>(sp);
        double sc = 0.0, den = 0.0;
        for (int j = pcen - w; j <= pcen + w; j++) {
            den = den + 1.0;
            sc = sc + 1.0 * S.llr[sb[(size_t)(j >> 4) * 512 + (j & 15)] & 31];
        }
        r.papa_llr = sc / den;
    } else {
        r.papa_prop = r.papa_fi = r.papa_llr = r.papa_llr2 = nan("");
    }

    // ---- traceback (:3110-3113) + longestrun (:1787-1804); Viterbi bits replace the traceback words
    int v = vlast, cur = 0, mx = 0;
    for (int j = (n - 1) >> 4; j >= 0; j--) {
        const uint32_t tw = tbp[(size_t)j * 32];
        uint32_t vb = 0;
        const int hi = (j == ((n - 1) >> 4)) ? ((n - 1) & 15) : 15;
        for (int i = hi; i >= 0; i--) {
            vb |= (uint32_t)v << i;
            cur = v ? cur + 1 : 0;
            mx = max(mx, cur);
            v = (tw >> (2 * i + v)) & 1;  // state at t-1 = tb[state at t][t]
        }
        tbp[(size_t)j * 32] = vb;
    }
    r.vit_maxrun = mx;

    r.core_start = -1;
    r.core_end = -2;
    r.prd_start = -1;
    r.prd_end = -2;
    r.core_score = nan("");
    r.prd_score = 0.0;
    if (mx >= c && n >= c) {
        core_search(reinterpret_cast<const uint8_t*>(sp), tbp, n, c, S.llr, ks.big_neg, r.core_start, r.core_end,
                    r.core_score, r.prd_start, r.prd_end, r.prd_score);
    }
    out[prot] = r;
}

}  // namespace plaac
