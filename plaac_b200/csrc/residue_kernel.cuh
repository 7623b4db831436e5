// Per-residue mode (plotsomefastas, plaac.java:610-647): VIT, MAP, the eight smoothed tracks and the two
// posterior columns for every residue.  One lane per protein on the bucketed stream (same prep as the
// summary mode).  This mode is bound by its 82 output bytes per residue, not by arithmetic.
//
//   k_residue_hmm      Viterbi + traceback (viterbidecodel :3077-3121), forward/backward in log space with
//                      the lookup-table log-sum-exp in REFERENCE ORDER (posteriorl :3349-3411), MAP parse
//                      (mapdecodel :4032-4045).  The forward variables are parked in the posterior output
//                      arrays, overwritten by a+b on the way back, and turned into exp(a+b-lpseq) last.
//   k_residue_windows  disorderreport tracks (:4866-4903) as running window sums.
#pragma once
#include "common.cuh"

namespace plaac {

// logeapeb :1024-1047 with the LUT in global memory (L1-resident), all branches as in the reference.
__device__ __forceinline__ double lse_lut_g(double a, double b, const double* __restrict__ lut, double ln2)
{
    if (a > b) {
        const double c = a - b;
        if (!(c < 40.0)) return a;
        const double x = 100.0 * c;
        const int dex = __double2int_rd(x);
        return a + ((x - (double)dex) * __ldg(lut + dex + 1) + ((double)(dex + 1) - x) * __ldg(lut + dex));
    } else if (b > a) {
        const double c = b - a;
        if (!(c < 40.0)) return b;
        const double x = 100.0 * c;
        const int dex = __double2int_rd(x);
        return b + ((x - (double)dex) * __ldg(lut + dex + 1) + ((double)(dex + 1) - x) * __ldg(lut + dex));
    }
    return a + ln2;
}

struct ResidueCtx {
    const uint8_t* sb;  // lane's slot 0 of the bucketed stream
    uint32_t* tbp;      // lane's word 0 of the traceback scratch
    int n;
    int64_t base;       // index of residue 0 in the output arrays
    __device__ __forceinline__ int ext(int t) const { return sb[(size_t)(t >> 4) * 512 + (t & 15)]; }
};

__device__ __forceinline__ bool residue_ctx(const BatchView& bv, int64_t res_base, int64_t rank, ResidueCtx& c)
{
    if (rank >= bv.nprot) return false;
    const int32_t prot = bv.order[rank];
    const int64_t b = rank >> 5;
    const int lane = (int)(rank & 31);
    const int64_t cb = bv.chunk_base[b];
    c.sb = reinterpret_cast<const uint8_t*>(bv.stream + cb * 32 + lane);
    c.tbp = bv.tbw + cb * 32 + lane;
    c.n = (int)(bv.offsets[prot + 1] - bv.offsets[prot]);
    c.base = bv.offsets[prot] - res_base;
    return c.n > 0;
}

__global__ void __launch_bounds__(128)
k_residue_hmm(BatchView bv, KScalars ks, const DeviceTables* __restrict__ T, plaac_residue_out out, int64_t res_base)
{
    ResidueCtx c;
    if (!residue_ctx(bv, res_base, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, c)) return;
    const int n = c.n;
    const double* lut = T->lut;
    double* A0 = out.post_bg + c.base;
    double* A1 = out.post_prd + c.base;

    // forward (:3356-3367) and Viterbi (:3085-3101) in one sweep
    double s0 = 0, s1 = 0, a0 = 0, a1 = 0;
    uint32_t tbacc = 0;
    for (int t = 0; t < n; t++) {
        const int cd = c.ext(t) & 31;
        const double le0 = T->le0[cd], le1 = T->le1[cd];
        if (t == 0) {
            s0 = ks.li0 + le0;
            s1 = ks.li1 + le1;
            a0 = s0;
            a1 = s1;
        } else {
            const double v00 = ks.lt00 + s0, v10 = ks.lt10 + s1;
            const double v01 = ks.lt01 + s0, v11 = ks.lt11 + s1;
            const bool tb0 = v10 > v00, tb1 = v11 > v01;
            s0 = (tb0 ? v10 : v00) + le0;
            s1 = (tb1 ? v11 : v01) + le1;
            tbacc |= ((uint32_t)tb0 | ((uint32_t)tb1 << 1)) << ((t & 15) * 2);
            // score = logeapeb(logeapeb(-Inf, x0), x1): the inner call returns x0
            const double f0 = lse_lut_g(ks.lt00 + a0, ks.lt10 + a1, lut, ks.ln2) + le0;
            const double f1 = lse_lut_g(ks.lt01 + a0, ks.lt11 + a1, lut, ks.ln2) + le1;
            a0 = f0;
            a1 = f1;
        }
        A0[t] = a0;
        A1[t] = a1;
        if ((t & 15) == 15 || t == n - 1) {
            c.tbp[(size_t)(t >> 4) * 32] = tbacc;
            tbacc = 0;
        }
    }
    // traceback (:3102-3113)
    if (out.vit) {
        uint8_t* vit = out.vit + c.base;
        int v = (s1 + ks.lf1 > s0 + ks.lf0) ? 1 : 0;
        for (int t = n - 1; t >= 0; t--) {
            vit[t] = (uint8_t)v;
            const uint32_t tw = c.tbp[(size_t)(t >> 4) * 32];
            v = (tw >> (2 * (t & 15) + v)) & 1;
        }
    }
    // backward (:3378-3391): b[i][t] = LSE_k( (lt[i][k] + b[k][t+1]) + le[k][aa[t+1]] ), k ascending
    double b0 = ks.lf0, b1 = ks.lf1;
    for (int t = n - 1; t >= 0; t--) {
        if (t < n - 1) {
            const int cd = c.ext(t + 1) & 31;
            const double le0 = T->le0[cd], le1 = T->le1[cd];
            const double nb0 = lse_lut_g((ks.lt00 + b0) + le0, (ks.lt01 + b1) + le1, lut, ks.ln2);
            const double nb1 = lse_lut_g((ks.lt10 + b0) + le0, (ks.lt11 + b1) + le1, lut, ks.ln2);
            b0 = nb0;
            b1 = nb1;
        }
        A0[t] = A0[t] + b0;
        A1[t] = A1[t] + b1;
    }
    // lpseq (:3393-3396) and the posteriors (:3401-3405); MAP (:4036-4040)
    const double lpseq = lse_lut_g(A0[0], A1[0], lut, ks.ln2);
    uint8_t* map = out.map ? out.map + c.base : nullptr;
    for (int t = 0; t < n; t++) {
        const double p0 = exp(A0[t] - lpseq);
        const double p1 = exp(A1[t] - lpseq);
        A0[t] = p0;
        A1[t] = p1;
        if (map) map[t] = (p1 > p0) ? 1 : 0;
    }
}

__global__ void __launch_bounds__(128)
k_residue_windows(BatchView bv, KScalars ks, const DeviceTables* __restrict__ T, plaac_residue_out out, int64_t res_base)
{
    ResidueCtx c;
    if (!residue_ctx(bv, res_base, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, c)) return;
    const int n = c.n, w = ks.w;
    const int off1 = 2 * w + 1, off2 = 4 * w + 2, full = 2 * w + 1, Wfull = full * full;
    double* o_charge = out.charge + c.base;
    double* o_hydro = out.hydro + c.base;
    double* o_fi = out.fi + c.base;
    double* o_plaac = out.plaac + c.base;
    double* o_papa = out.papa + c.base;
    double* o_fix2 = out.fix2 + c.base;
    double* o_plaacx2 = out.plaacx2 + c.base;
    double* o_papax2 = out.papax2 + c.base;

    double SLh = 0, SGh = 0, Th = 0, SLl = 0, SGl = 0, Tl = 0, SLp = 0, SGp = 0, Tp = 0;
    int SLc = 0, SGc = 0, Tac = 0;
    auto ext_or_pad = [&](int t) -> int { return (t >= 0 && t < n) ? c.ext(t) : kPad; };
    const int t_end = n + 2 * w;
    for (int t = 0; t < t_end; t++) {
        const int e0 = ext_or_pad(t), e1 = ext_or_pad(t - off1), e2 = ext_or_pad(t - off2);
        const int ch0 = ((int)(int8_t)e0) >> 6, ch1 = ((int)(int8_t)e1) >> 6, ch2 = ((int)(int8_t)e2) >> 6;
        const double hy0 = T->hyd[e0 & 63], hy1 = T->hyd[e1 & 63], hy2 = T->hyd[e2 & 63];
        const double lr0 = T->llr[e0 & 63], lr1 = T->llr[e1 & 63], lr2 = T->llr[e2 & 63];
        const double pa0 = T->pap[e0 & 63], pa1 = T->pap[e1 & 63], pa2 = T->pap[e2 & 63];
        SLh = (SLh + hy0) - hy1;
        SGh = (SGh + hy1) - hy2;
        Th = (Th + SLh) - SGh;
        SLl = (SLl + lr0) - lr1;
        SGl = (SGl + lr1) - lr2;
        Tl = (Tl + SLl) - SGl;
        SLp = (SLp + pa0) - pa1;
        SGp = (SGp + pa1) - pa2;
        Tp = (Tp + SLp) - SGp;
        SLc += ch0 - ch1;
        SGc += ch1 - ch2;
        Tac += abs(SLc) - abs(SGc);
        const int p = t - w;  // pass-1 centre (slidingaverage shrink=true, :2604-2620)
        double fi_p = 0, ll_p = 0, pa_p = 0;
        if (p >= 0 && p < n) {
            const double cnt = (double)(full - max(0, w - p) - max(0, p + w - (n - 1)));
            const double hyd = SLh / cnt;
            const double chg = (double)SLc / cnt;
            fi_p = (ks.cc0 * hyd + ks.cc1 * fabs(chg)) + ks.cc2;
            ll_p = SLl / cnt;
            pa_p = SLp / cnt;
            o_hydro[p] = hyd;
            o_charge[p] = chg;
            o_fi[p] = fi_p;
            o_plaac[p] = ll_p;
            o_papa[p] = pa_p;
            if (n == 1) {  // w clips to 0: the second pass returns the value itself (:2588-2589)
                o_fix2[0] = fi_p;
                o_plaacx2[0] = ll_p;
                o_papax2[0] = pa_p;
            }
        }
        const int k = t - 2 * w;  // pass-2 centre (weight=true, NaN outside [w, n-w-1])
        if (k >= 0 && k < n && n > 1) {
            if (k >= w && k <= n - 1 - w) {
                const int ml = 2 * w - k, mr = 2 * w - (n - 1 - k);
                const double Wd = (double)(Wfull - (ml > 0 ? (ml * (ml + 1)) >> 1 : 0) - (mr > 0 ? (mr * (mr + 1)) >> 1 : 0));
                o_fix2[k] = ((ks.cc0 * Th + ks.cc1 * (double)Tac) + ks.cc2 * Wd) / Wd;
                o_plaacx2[k] = Tl / Wd;
                o_papax2[k] = Tp / Wd;
            } else {
                o_fix2[k] = nan("");
                o_plaacx2[k] = nan("");
                o_papax2[k] = nan("");
            }
        }
    }
}

inline int residue_setup(const KScalars&, int) { return PLAAC_OK; }

inline int launch_residue(const KScalars& ks, const DeviceTables* tabs, const BatchView& bv, const plaac_residue_out& out,
                          int64_t res_base, int /*sm_count*/, cudaStream_t st, int64_t* launches)
{
    if (!out.post_bg || !out.post_prd || !out.charge || !out.hydro || !out.fi || !out.plaac || !out.papa || !out.fix2 ||
        !out.plaacx2 || !out.papax2)
        return PLAAC_E_INVALID;
    const unsigned grid = (unsigned)((bv.nprot + 127) / 128);
    k_residue_hmm<<<grid, 128, 0, st>>>(bv, ks, tabs, out, res_base);
    k_residue_windows<<<grid, 128, 0, st>>>(bv, ks, tabs, out, res_base);
    if (launches) *launches += 2;
    return cudaGetLastError() == cudaSuccess ? PLAAC_OK : PLAAC_E_CUDA;
}

}  // namespace plaac
