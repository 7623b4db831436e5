// Window tracks and the disorderreport columns for window sizes the streaming kernels do not cover.
//
// plaac.jar takes -w (FoldIndex window, ww1) and -W (PAPA window, ww2; ww3 = ww2, plaac.java:355) independently.  The
// streaming kernels (summary_kernel_v2.cuh, long_kernel.cuh, residue_kernel_v2.cuh) carry ONE half-width through their
// running sums, which covers the jar's default and every ww1/2 == ww2/2 setting.  For the rest these three kernels
// recompute what depends on the windows, tap by tap in plaac.java's own operation order -- slidingaverage :2585-2662 and
// the disorderreport constructor :4866-5068 -- so every value has the jar's bits.  It is the slow path of a rarely used
// option: O(window) work per residue, one warp per protein.
//
//   k_gen_pass1   hydro, charge, fi, plaac, papa                  (first sliding average, shrink at the ends)
//   k_gen_pass2   fix2, plaacx2, papax2                           (second, weighted average; NaN near the ends)
//   k_gen_report  FI runs, PAPA centre and the values there, means -> the 11 window-dependent fields of plaac_summary
#pragma once
#include "common.cuh"

namespace plaac {

struct GenTables {
    double hyd[PLAAC_NAA], chg[PLAAC_NAA], llr[PLAAC_NAA], pap[PLAAC_NAA];  // exact table values (plaac_params)
};

struct GenArgs {
    const uint8_t* codes;
    const int64_t* offsets;
    int64_t code_base;  // residue index of codes[0]
    int64_t out_base;   // residue index of element 0 of the track arrays
    int64_t nprot;
    int ww1, ww2, ww3, adjust_prolines;
    double cc0, cc1, cc2;
    double *hydro, *charge, *fi, *plaac, *papa, *fix2, *plaacx2, *papax2;
    plaac_summary* out;  // k_gen_report only
    GenTables t;
};

__device__ __forceinline__ uint32_t gen_code(const uint8_t* src, int p)
{
    const uint32_t c = src[p];
    return c > 21u ? 0u : c;  // invalid input is scored as X (and reported by the packing kernel)
}

// slidingaverage(arr, ww, shrink = true, weight = false [, mergeme = 13, seq]) :2585-2662 at position i
__device__ __forceinline__ double gen_avg1(const uint8_t* src, int n, int i, int ww, const double* tab, bool prolines)
{
    int w = ww / 2;
    if (w >= n) w = n - 1;
    double score = 0.0, denom = 0.0;
    for (int j = -w; j <= w; j++) {
        const int p = i + j;
        if (p >= 0 && p < n) {
            denom = denom + 1.0;
            const uint32_t c = gen_code(src, p);
            if (prolines && c == 13u && ((p >= 1 && gen_code(src, p - 1) == 13u) || (p >= 2 && gen_code(src, p - 2) == 13u)))
                continue;  // the second proline of PP / PxP is not scored (:2652-2655); it still counts in denom
            score = score + 1.0 * tab[c];
        }
    }
    return score / denom;
}

// slidingaverage(arr, ww, shrink = false, weight = true) :2585-2622 at position i
__device__ __forceinline__ double gen_avg2(const double* arr, int n, int i, int ww)
{
    int w = ww / 2;
    if (w >= n) w = n - 1;
    if (i < w || i > n - w - 1) return nan("");
    double score = 0.0, denom = 0.0;
    for (int j = -w; j <= w; j++) {
        const int p = i + j;  // always inside the protein here
        const int m1 = p < w ? p : w, m2 = (n - p - 1 < w) ? (n - p - 1) : w;
        const double wt = (1.0 + (double)m1) + (double)m2;
        denom = denom + wt;
        score = score + wt * arr[p];
    }
    return score / denom;
}

__global__ void __launch_bounds__(128) k_gen_pass1(GenArgs g)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < g.nprot; p += warps) {
        const int64_t o = g.offsets[p];
        const int n = (int)(g.offsets[p + 1] - o);
        const uint8_t* src = g.codes + (o - g.code_base);
        const int64_t ob = o - g.out_base;
        for (int i = lane; i < n; i += 32) {
            const double hy = gen_avg1(src, n, i, g.ww1, g.t.hyd, false);
            const double ch = gen_avg1(src, n, i, g.ww1, g.t.chg, false);
            g.hydro[ob + i] = hy;
            g.charge[ob + i] = ch;
            g.fi[ob + i] = (g.cc0 * hy + g.cc1 * fabs(ch)) + g.cc2;  // :4885
            g.plaac[ob + i] = gen_avg1(src, n, i, g.ww3, g.t.llr, false);
            g.papa[ob + i] = gen_avg1(src, n, i, g.ww2, g.t.pap, g.adjust_prolines != 0);
        }
    }
}

__global__ void __launch_bounds__(128) k_gen_pass2(GenArgs g)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < g.nprot; p += warps) {
        const int64_t o = g.offsets[p];
        const int n = (int)(g.offsets[p + 1] - o);
        const int64_t ob = o - g.out_base;
        for (int i = lane; i < n; i += 32) {
            g.papax2[ob + i] = gen_avg2(g.papa + ob, n, i, g.ww2);
            g.plaacx2[ob + i] = gen_avg2(g.plaac + ob, n, i, g.ww3);
            g.fix2[ob + i] = gen_avg2(g.fi + ob, n, i, g.ww1);
        }
    }
}

// One lane per protein: the scans are sequential by definition (first strict maximum, maximal runs).
__global__ void __launch_bounds__(128) k_gen_report(GenArgs g)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < g.nprot; p += stride) {
        const int64_t o = g.offsets[p];
        const int n = (int)(g.offsets[p + 1] - o);
        if (n < 1) continue;  // the jar prints no row (:762); the record stays as the scoring kernels left it
        const uint8_t* src = g.codes + (o - g.code_base);
        const int64_t ob = o - g.out_base;
        plaac_summary* r = g.out + p;
        // means :4875-4885 (mean() :1584 is a sequential sum)
        double sh = 0.0, sc = 0.0;
        for (int i = 0; i < n; i++) {
            const uint32_t c = gen_code(src, i);
            sh = sh + g.t.hyd[c];
            sc = sc + g.t.chg[c];
        }
        const double mh = (1.0 * sh) / (double)n, mc = (1.0 * sc) / (double)n;
        r->fi_meanhydro = mh;
        r->fi_meancharge = mc;
        r->fi_meancombo = (g.cc2 + g.cc1 * fabs(mc)) + g.cc0 * mh;
        // PAPA centre :4931-4948: first strict maximum of papax2 among centres with fix2 < 0
        const double* papax2 = g.papax2 + ob;
        const double* fix2 = g.fix2 + ob;
        double best = -INFINITY;
        int cen = -1;
        const int hp = (g.ww2 - 1) / 2;
        for (int k = hp; k < n - hp; k++) {
            const double v = papax2[k];
            if ((v > best) & (fix2[k] < 0)) {
                cen = k;
                best = v;
            }
        }
        r->papa_center = cen;
        r->papa_combo = best;
        if (cen >= 0) {
            r->papa_prop = papax2[cen];
            r->papa_fi = fix2[cen];
            r->papa_llr2 = g.plaacx2[ob + cen];
            r->papa_llr = g.plaac[ob + cen];
        } else
            r->papa_prop = r->papa_fi = r->papa_llr = r->papa_llr2 = nan("");
        // FoldIndex runs :4912-4913, :5010-5059
        const double* fi = g.fi + ob;
        int halfw = (g.ww1 - 1) / 2;
        if (halfw > n / 2) halfw = n / 2;
        int num = 0, mx = 0, i = halfw;
        while (i < n - halfw) {
            if (fi[i] < 0) {
                int s0 = i;
                i++;
                while (i < n - halfw && fi[i] < 0) i++;
                int s1 = i - 1;
                if (s0 == halfw) s0 = 0;
                if (s1 == n - halfw - 1) s1 = n - 1;
                const int len = s1 - s0 + 1;
                if (len >= 5) {
                    num += len;
                    mx = max(mx, len);
                }
            } else
                i++;
        }
        r->fi_numaa = num;
        r->fi_maxrun = mx;
    }
}

}  // namespace plaac
