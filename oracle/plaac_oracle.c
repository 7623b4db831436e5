/*
 * plaac_oracle.c -- CPU restatement of the PLAAC per-protein scoring path.
 * TEST INFRASTRUCTURE ONLY; parity pinned against the jar's own bytecode (see plaac_oracle.h).
 *
 * Written against cli/src/plaac.java of whitehead/plaac ("plaac.java:N" below).
 * Arithmetic is IEEE double, one rounding per Java operator, evaluated left to
 * right as the Java source writes it; compile with -ffp-contract=off so that
 * no multiply-add is fused (Java never fuses).
 */
#include "plaac_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ tables */

/* plaac.java:37-60 */
static const double k_aacharge[ORC_NAA] = {0, 0, 0, 1, 1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 0, -1, 0, 0, 0, 0, 0, 0};
/* plaac.java:64-87 (Kyte-Doolittle) */
static const double k_aahydro[ORC_NAA] = {0.0,  1.8,  2.5,  -3.5, -3.5, 2.8,  -0.4, -3.2, 4.5,  -3.9, 3.8,
                                          1.9,  -3.5, -1.6, -3.5, -4.5, -0.8, -0.7, 4.2,  -0.9, -1.3, 0.0};
/* plaac.java:206-229 */
static const double k_odpapa1[ORC_NAA] = {0.0,        0.67267686, 1.5146198,  0.27887323, 0.5460614, 2.313433,
                                          0.96153843, 0.75686276, 2.2562358,  0.20664589, 0.9607843, 1.9615384,
                                          1.0836071,  0.30196398, 1.0716166,  0.6664044,  1.1432927, 0.8917492,
                                          2.2562358,  1.9478673,  2.1785367,  0.0};
/* plaac.java:261-262 */
static const double k_bg_freq_scer[ORC_NAA] = {0,      0.0550, 0.0126, 0.0586, 0.0655, 0.0441, 0.0498, 0.0217,
                                               0.0655, 0.0735, 0.0950, 0.0207, 0.0615, 0.0438, 0.0396, 0.0444,
                                               0.0899, 0.0592, 0.0556, 0.0104, 0.0337, 0};
/* plaac.java:269-270 */
static const double k_prd_freq_scer_28[ORC_NAA] = {0,       0.04865, 0.00219, 0.01638, 0.00783, 0.02537, 0.07603, 0.0181,
                                                   0.02018, 0.01641, 0.02639, 0.02975, 0.25885, 0.05126, 0.15178, 0.025,
                                                   0.10988, 0.03841, 0.01972, 0.00157, 0.05624, 0};

/* plaac.java:1508-1534 */
int orc_aatoint(int c)
{
    switch (c) {
    case 'A': case 'a': return 1;
    case 'C': case 'c': return 2;
    case 'D': case 'd': return 3;
    case 'E': case 'e': return 4;
    case 'F': case 'f': return 5;
    case 'G': case 'g': return 6;
    case 'H': case 'h': return 7;
    case 'I': case 'i': return 8;
    case 'K': case 'k': return 9;
    case 'L': case 'l': return 10;
    case 'M': case 'm': return 11;
    case 'N': case 'n': return 12;
    case 'P': case 'p': return 13;
    case 'Q': case 'q': return 14;
    case 'R': case 'r': return 15;
    case 'S': case 's': return 16;
    case 'T': case 't': return 17;
    case 'V': case 'v': return 18;
    case 'W': case 'w': return 19;
    case 'Y': case 'y': return 20;
    case '*': return 21;
    default: return 0;
    }
}

/* plaac.java:1933-1941: sequential sum; if sum < 1e-12 divide by 1 */
static void normalize22(const double *arr, double *out)
{
    double sm = 0;
    for (int i = 0; i < ORC_NAA; i++) sm = sm + arr[i];
    sm = 1.0 * sm;
    if (sm < 0.000000000001) sm = 1;
    for (int i = 0; i < ORC_NAA; i++) out[i] = arr[i] / sm;
}

/* plaac.java:2893-2935 for the two fixed 2-state models of :968-1001 */
static void hmm_init(orc_hmm *h, const double tmat[2][2], const double imat[2], const double *e0, const double *e1)
{
    for (int i = 0; i < 2; i++) {
        for (int j = 0; j < 2; j++) h->lt[i][j] = log(tmat[i][j]);
        h->li[i] = log(imat[i]);
    }
    for (int j = 0; j < ORC_NAA; j++) {
        h->le[0][j] = log(e0[j]);
        h->le[1][j] = log(e1[j]);
    }
    /* :2904-2922  fprob = max(0, 1 - rowsum); all <= 1e-4 => freeend => fprob = 1 */
    double fprob[2];
    int freeend = 1;
    for (int i = 0; i < 2; i++) {
        double rs = 0;
        for (int j = 0; j < 2; j++) rs = rs + tmat[i][j];
        fprob[i] = fmax(0.0, 1.0 - rs);
        if (fprob[i] > 0.0001) freeend = 0;
    }
    if (freeend) fprob[0] = fprob[1] = 1.0;
    for (int i = 0; i < 2; i++) h->lf[i] = log(fprob[i]);
}

void orc_params_init(orc_params *P, double alpha, const double *bg_counts, const double *fg_freq, int core_len, int ww1,
                     int ww2, int ww3, int adjust_prolines)
{
    memset(P, 0, sizeof(*P));
    P->core_len = core_len;
    P->ww1 = ww1;
    P->ww2 = ww2;
    P->ww3 = ww3;
    P->adjust_prolines = adjust_prolines;
    /* :444-447 */
    if (alpha > 1 || alpha < 0) alpha = 1.0;
    P->alpha = alpha;
    P->ln2 = log(2); /* :30 */

    /* :282-283 */
    for (int i = 0; i <= ORC_LUTLEN; i++) P->loglut[i] = log(1.0 + exp(-i / 100.0));
    /* :288-291 */
    for (int k = 1; k <= 20; k++) P->papa_lod[k] = log(k_odpapa1[k]);
    /* :90 aahydro2 = axpb(1/9, aahydro, 0.5) */
    {
        double a = 1.0 / 9.0;
        for (int k = 0; k < ORC_NAA; k++) P->hydro2[k] = a * k_aahydro[k] + 0.5;
    }
    memcpy(P->charge, k_aacharge, sizeof(k_aacharge));
    P->fi_cc[0] = 2.785; /* :800 */
    P->fi_cc[1] = -1;
    P->fi_cc[2] = -1.151;

    /* :310 */
    normalize22(k_bg_freq_scer, P->bgscer);

    double fgfreq[ORC_NAA], bgf[ORC_NAA], bgcombo[ORC_NAA], tmp[ORC_NAA];
    memcpy(fgfreq, fg_freq ? fg_freq : k_prd_freq_scer_28, sizeof(fgfreq));
    if (bg_counts)
        memcpy(bgf, bg_counts, sizeof(bgf));
    else
        memset(bgf, 0, sizeof(bgf));

    /* :449-458 */
    fgfreq[0] = 0;
    fgfreq[21] = 0;
    normalize22(fgfreq, tmp);
    memcpy(fgfreq, tmp, sizeof(tmp));
    bgf[0] = 0;
    bgf[21] = 0;
    normalize22(bgf, P->bgthis);
    for (int i = 0; i < ORC_NAA; i++) tmp[i] = alpha * P->bgscer[i] + (1 - alpha) * P->bgthis[i];
    normalize22(tmp, bgcombo);

    /* :490-500 */
    double epsx = 0.00001;
    fgfreq[0] = epsx;
    fgfreq[21] = epsx;
    bgcombo[0] = epsx;
    bgcombo[21] = epsx;
    normalize22(fgfreq, P->fg);
    normalize22(bgcombo, P->bg);
    for (int j = 1; j < 21; j++) P->llr[j] = log(P->fg[j] / P->bg[j]);

    /* :968-981 prionhmm1 (note the second normalisation of fg and bg) */
    {
        const double tmat[2][2] = {{99.9 / 100, 0.1 / 100}, {2.0 / 100, 98.0 / 100}};
        const double imat[2] = {0.9524, 0.0476};
        double e0[ORC_NAA], e1[ORC_NAA];
        normalize22(P->bg, e0);
        normalize22(P->fg, e1);
        hmm_init(&P->hmm1, tmat, imat, e0, e1);
    }
    /* :988-1001 prionhmm0 */
    {
        const double tmat[2][2] = {{1, 0}, {0, 1}};
        const double imat[2] = {1, 0};
        double e0[ORC_NAA];
        normalize22(P->bg, e0);
        hmm_init(&P->hmm0, tmat, imat, e0, e0);
    }
}

/* plaac.java:1024-1047 */
double orc_logeapeb(const orc_params *P, double a, double b)
{
    const double *loglut = P->loglut;
    if (a > b) {
        double c = a - b;
        if (!(c < 40)) return a;
        int dex = (int)floor(100 * c);
        return (a + ((100 * c - dex) * loglut[dex + 1] + (dex + 1 - 100 * c) * loglut[dex]));
    } else if (b > a) {
        double c = b - a;
        if (!(c < 40)) return b;
        int dex = (int)floor(100 * c);
        return (b + ((100 * c - dex) * loglut[dex + 1] + (dex + 1 - 100 * c) * loglut[dex]));
    } else
        return (a + P->ln2);
}

/* plaac.java:1206-1257 (general min/max form kept, including the inner j loop) */
void orc_hss2(const double *seq, int n, int minlength, int maxlength, double score[3])
{
    if (minlength > n || minlength > maxlength) {
        score[0] = -1.0;
        score[1] = -2.0;
        score[2] = -INFINITY;
        return;
    }
    if (maxlength > n) maxlength = n;
    double bestscore;
    int beststart = 0;
    int beststop = minlength - 1;
    int curstart = 0;
    int newstart = 0;
    double *psum = (double *)malloc(sizeof(double) * ((size_t)n + 1));
    psum[0] = 0;
    for (int i = 0; i < n; i++) psum[i + 1] = psum[i] + seq[i];
    double d = psum[minlength];
    bestscore = d;
    for (int i = minlength; i < n; i++) {
        if ((i - curstart) >= maxlength) curstart++;
        d = psum[i + 1] - psum[curstart];
        newstart = curstart;
        for (int j = curstart + 1; j < i - minlength; j++) {
            if (psum[i + 1] - psum[j] >= d) {
                d = psum[i + 1] - psum[j];
                newstart = j;
            }
            curstart = newstart;
        }
        if (d > bestscore) {
            bestscore = d;
            beststop = i;
            beststart = curstart;
        }
    }
    free(psum);
    score[0] = beststart;
    score[1] = beststop;
    score[2] = bestscore;
}

/* plaac.java:3077-3121 */
double orc_viterbi(const orc_hmm *h, const uint8_t *seq, int n, uint8_t *vit)
{
    const int ns = 2;
    double *s = (double *)malloc(sizeof(double) * 2 * (size_t)n);
    uint8_t *tb = (uint8_t *)calloc(2 * (size_t)n, 1);
    double *s0 = s, *s1 = s + n;
    double *sv[2] = {s0, s1};
    uint8_t *tbv[2] = {tb, tb + n};
    for (int i = 0; i < ns; i++) sv[i][0] = h->li[i] + h->le[i][seq[0]];
    for (int t = 1; t < n; t++) {
        for (int i = 0; i < ns; i++) {
            int bestdex = 0;
            double bestscore = h->lt[0][i] + sv[0][t - 1];
            for (int k = 1; k < ns; k++) {
                if (h->lt[k][i] + sv[k][t - 1] > bestscore) {
                    bestscore = h->lt[k][i] + sv[k][t - 1];
                    bestdex = k;
                }
            }
            sv[i][t] = bestscore + h->le[i][seq[t]];
            tbv[i][t] = (uint8_t)bestdex;
        }
    }
    int bestdex = 0;
    double bestscore = sv[0][n - 1] + h->lf[0];
    for (int k = 1; k < ns; k++) {
        if (sv[k][n - 1] + h->lf[k] > bestscore) {
            bestscore = sv[k][n - 1] + h->lf[k];
            bestdex = k;
        }
    }
    vit[n - 1] = (uint8_t)bestdex;
    for (int t = n - 2; t >= 0; t--) vit[t] = tbv[vit[t + 1]][t + 1];
    free(s);
    free(tb);
    return bestscore;
}

/* plaac.java:3349-3411 */
double orc_posterior(const orc_params *P, const orc_hmm *h, const uint8_t *seq, int n, double *pp0, double *pp1,
                     int want_posterior)
{
    const int ns = 2;
    double *a = (double *)malloc(sizeof(double) * 2 * (size_t)n);
    double *av[2] = {a, a + n};
    for (int i = 0; i < ns; i++) av[i][0] = h->li[i] + h->le[i][seq[0]];
    for (int t = 1; t < n; t++) {
        for (int i = 0; i < ns; i++) {
            double score = -INFINITY;
            for (int k = 0; k < ns; k++) score = orc_logeapeb(P, score, h->lt[k][i] + av[k][t - 1]);
            av[i][t] = score + h->le[i][seq[t]];
        }
    }
    double ltotprob = -INFINITY;
    for (int i = 0; i < ns; i++) ltotprob = orc_logeapeb(P, ltotprob, av[i][n - 1] + h->lf[i]);

    if (want_posterior) {
        double *b = (double *)malloc(sizeof(double) * 2 * (size_t)n);
        double *bv[2] = {b, b + n};
        for (int i = 0; i < ns; i++) bv[i][n - 1] = h->lf[i];
        for (int t = n - 2; t >= 0; t--) {
            for (int i = 0; i < ns; i++) {
                double score = -INFINITY;
                for (int k = 0; k < ns; k++)
                    score = orc_logeapeb(P, score, h->lt[i][k] + bv[k][t + 1] + h->le[k][seq[t + 1]]);
                bv[i][t] = score;
            }
        }
        double lpseq = -INFINITY;
        for (int i = 0; i < ns; i++) lpseq = orc_logeapeb(P, lpseq, av[i][0] + bv[i][0]);
        for (int t = 0; t < n; t++) {
            double p0 = exp((av[0][t] + bv[0][t]) - lpseq);
            double p1 = exp((av[1][t] + bv[1][t]) - lpseq);
            if (pp0) pp0[t] = p0;
            if (pp1) pp1[t] = p1;
        }
        free(b);
    }
    free(a);
    return ltotprob;
}

/* plaac.java:2585-2622 (mergeme < 0) and :2626-2662 (mergeme >= 0) */
void orc_slidingaverage(const double *arr, int n, int ww, int shrink, int weight, int mergeme, const uint8_t *seq,
                        double *sa)
{
    if (n == 0) return;
    int w = ww / 2;
    if (w >= n) w = n - 1;
    int mini, maxi;
    if (shrink) {
        mini = 0;
        maxi = n - 1;
    } else {
        mini = w;
        maxi = n - w - 1;
        for (int i = 0; i < mini; i++) sa[i] = NAN;
        for (int i = maxi + 1; i < n; i++) sa[i] = NAN;
    }
    for (int i = mini; i <= maxi; i++) {
        double score = 0.0;
        double denom = 0.0;
        for (int j = -w; j <= w; j++) {
            if ((i + j >= 0) && (i + j < n)) {
                double wt = 1.0;
                if (weight) {
                    int m1 = (i + j < w) ? (i + j) : w;
                    int m2 = (n - i - j - 1 < w) ? (n - i - j - 1) : w;
                    wt = 1.0 + m1 + m2;
                }
                denom = denom + wt;
                if (mergeme >= 0) {
                    if (!((seq[i + j] == mergeme) && (i + j >= 1) && (seq[i + j - 1] == mergeme)) &&
                        !((seq[i + j] == mergeme) && (i + j >= 2) && (seq[i + j - 2] == mergeme))) {
                        score = score + wt * arr[i + j];
                    }
                } else {
                    score = score + wt * arr[i + j];
                }
            }
        }
        sa[i] = score / denom;
    }
}

/* plaac.java:2574-2581 */
static void mapseq(const uint8_t *aa, int n, const double *map, double *out)
{
    for (int i = 0; i < n; i++) out[i] = map[aa[i]];
}

/* plaac.java:1787-1804 */
static int longestrun(const uint8_t *bitvec, int n)
{
    int maxlen = 0;
    int i = 0;
    while (i < n) {
        if (bitvec[i] > 0) {
            int startdex = i;
            i++;
            while (i < n && bitvec[i] > 0) i++;
            int stopdex = i - 1;
            int len = stopdex - startdex + 1;
            if (len >= maxlen) maxlen = len;
        } else {
            i++;
        }
    }
    return maxlen;
}

/* The live part of the disorderreport constructor, plaac.java:4866-5068
 * (dead: numdisordered(strict) :4903-4929 are never printed; hssr/hssr2
 * :5002-5007 are never read; localmean/sd/... are never printed). */
typedef struct {
    double *hydro, *charge, *fi, *plaacllr, *papa, *papax2, *plaacllrx2, *fix2;
    double meanhydro, meancharge, meanfi;
    int numdisorderedstrict2, maxlen;
    double papamaxprop, papamaxscore, papamaxdis, papamaxllr, papamaxllr2;
    int papamaxcenter;
} disorderreport;

static void dr_free(disorderreport *d)
{
    free(d->hydro);
}

static void dr_compute(const orc_params *P, const uint8_t *aa, int n, disorderreport *d)
{
    const int ww1 = P->ww1, ww2 = P->ww2, ww3 = P->ww3;
    const double *cc = P->fi_cc;
    /* one slab: 8 tracks + 2 temporaries */
    double *slab = (double *)malloc(sizeof(double) * 10 * (size_t)n);
    d->hydro = slab;
    d->charge = slab + (size_t)n;
    d->fi = slab + 2 * (size_t)n;
    d->plaacllr = slab + 3 * (size_t)n;
    d->papa = slab + 4 * (size_t)n;
    d->papax2 = slab + 5 * (size_t)n;
    d->plaacllrx2 = slab + 6 * (size_t)n;
    d->fix2 = slab + 7 * (size_t)n;
    double *maa = slab + 8 * (size_t)n;
    double *tmp = slab + 9 * (size_t)n;

    /* :4875-4881 */
    mapseq(aa, n, P->hydro2, maa);
    {
        double mn = 0;
        for (int i = 0; i < n; i++) mn = mn + maa[i];
        d->meanhydro = (1.0 * mn) / n;
    }
    orc_slidingaverage(maa, n, ww1, 1, 0, -1, NULL, d->hydro);
    mapseq(aa, n, P->charge, maa);
    {
        double mn = 0;
        for (int i = 0; i < n; i++) mn = mn + maa[i];
        d->meancharge = (1.0 * mn) / n;
    }
    orc_slidingaverage(maa, n, ww1, 1, 0, -1, NULL, d->charge);
    /* :4883-4885 */
    d->meanfi = cc[2] + cc[1] * fabs(d->meancharge) + cc[0] * d->meanhydro;
    for (int i = 0; i < n; i++) d->fi[i] = cc[0] * d->hydro[i] + cc[1] * fabs(d->charge[i]) + cc[2];
    /* :4887-4895 */
    mapseq(aa, n, P->llr, maa);
    orc_slidingaverage(maa, n, ww3, 1, 0, -1, NULL, d->plaacllr);
    mapseq(aa, n, P->papa_lod, tmp);
    if (P->adjust_prolines)
        orc_slidingaverage(tmp, n, ww2, 1, 0, 13, aa, d->papa);
    else
        orc_slidingaverage(tmp, n, ww2, 1, 0, -1, NULL, d->papa);
    /* :4901-4903 */
    orc_slidingaverage(d->papa, n, ww2, 0, 1, -1, NULL, d->papax2);
    orc_slidingaverage(d->plaacllr, n, ww3, 0, 1, -1, NULL, d->plaacllrx2);
    orc_slidingaverage(d->fi, n, ww1, 0, 1, -1, NULL, d->fix2);

    /* :4931-4948 papamode == 1 */
    d->papamaxscore = -INFINITY;
    d->papamaxcenter = -1;
    for (int k = (ww2 - 1) / 2; k < n - (ww2 - 1) / 2; k++) {
        double papascore = d->papax2[k];
        if ((papascore > d->papamaxscore) & (d->fix2[k] < 0)) {
            d->papamaxcenter = k;
            d->papamaxscore = papascore;
        }
    }
    /* :4985-4997 */
    d->papamaxprop = NAN;
    d->papamaxdis = NAN;
    d->papamaxllr = NAN;
    d->papamaxllr2 = NAN;
    if (d->papamaxcenter >= 0) {
        d->papamaxprop = d->papax2[d->papamaxcenter];
        d->papamaxdis = d->fix2[d->papamaxcenter];
        d->papamaxllr2 = d->plaacllrx2[d->papamaxcenter];
        d->papamaxllr = d->plaacllr[d->papamaxcenter];
    }

    /* :4912-4913 and :5010-5059 */
    int halfw = (ww1 - 1) / 2;
    if (halfw > n / 2) halfw = n / 2;
    const int minlen = 5;
    int i = 0 + halfw;
    d->numdisorderedstrict2 = 0;
    d->maxlen = 0;
    while (i < n - halfw) {
        if (d->fi[i] < 0) {
            int startdex = i;
            i++;
            while (i < n - halfw && d->fi[i] < 0) i++;
            int stopdex = i - 1;
            if (startdex == halfw) startdex = 0;
            if (stopdex == n - halfw - 1) stopdex = n - 1;
            int len = stopdex - startdex + 1;
            if (len >= minlen) {
                d->numdisorderedstrict2 = d->numdisorderedstrict2 + len;
                if (len > d->maxlen) d->maxlen = len; /* maxint(lenaa) :5060 */
            }
        } else
            i++;
    }
}

/* plaac.java:755-948 */
void orc_score_protein(const orc_params *P, const uint8_t *aa, int n, orc_summary *out, int full_jar_work)
{
    memset(out, 0, sizeof(*out));
    out->prot_len = n;
    if (n < 1) return; /* :762 -- the jar prints no row */
    const int corelength = P->core_len;
    double *maa = (double *)malloc(sizeof(double) * (size_t)n);
    uint8_t *mp = (uint8_t *)malloc((size_t)n);
    double hs1[3], hs2[3], hs3[3];

    /* :764-771 MW */
    static const double qnmask[ORC_NAA] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1.0, 0, 1.0, 0, 0, 0, 0, 0, 0, 0};
    mapseq(aa, n, qnmask, maa);
    int mwsize = 80;
    if (n < 80) mwsize = n;
    orc_hss2(maa, n, mwsize, mwsize, hs1);
    /* :782-783 LLR */
    mapseq(aa, n, P->llr, maa);
    orc_hss2(maa, n, corelength, corelength, hs2);

    /* :794-798 */
    double lvit1 = orc_viterbi(&P->hmm1, aa, n, mp);
    double lmarg1, lvit0, lmarg0;
    if (full_jar_work) {
        double *pp = (double *)malloc(sizeof(double) * 2 * (size_t)n);
        uint8_t *mp0 = (uint8_t *)malloc((size_t)n);
        lmarg1 = orc_posterior(P, &P->hmm1, aa, n, pp, pp + n, 1);
        lvit0 = orc_viterbi(&P->hmm0, aa, n, mp0);
        lmarg0 = orc_posterior(P, &P->hmm0, aa, n, pp, pp + n, 1);
        free(pp);
        free(mp0);
    } else {
        uint8_t *mp0 = (uint8_t *)malloc((size_t)n);
        lmarg1 = orc_posterior(P, &P->hmm1, aa, n, NULL, NULL, 0);
        lvit0 = orc_viterbi(&P->hmm0, aa, n, mp0);
        lmarg0 = orc_posterior(P, &P->hmm0, aa, n, NULL, NULL, 0);
        free(mp0);
    }
    double hmmscore = lmarg1 - lmarg0;
    double hmmscorev = lvit1 - lvit0;

    /* :800 */
    disorderreport dr;
    dr_compute(P, aa, n, &dr);

    /* :814-833 */
    int longestprd = longestrun(mp, n);
    mapseq(aa, n, P->llr, maa);
    const double big_neg = -1000000.0;
    for (int i = 0; i < n; i++)
        if (mp[i] == 0) maa[i] = big_neg;
    orc_hss2(maa, n, corelength, corelength, hs3);

    /* :851-880 */
    int corestart = (int)hs3[0];
    int corestop = (int)hs3[1];
    int aastart = corestart;
    int aastop = corestop;
    double prdscore = 0;
    if (hs3[2] > big_neg / 2) {
        while (aastart >= 0 && mp[aastart] == 1) aastart--;
        aastart++;
        while (aastop < n && mp[aastop] == 1) aastop++;
        aastop--;
        /* prd = submatrix(aa, aastart, aastop) :867 (clamps of :1445-1456 are no-ops here) */
        for (int kk = aastart; kk <= aastop; kk++) prdscore = prdscore + P->llr[aa[kk]];
    } else {
        hs3[2] = NAN;
        aastart = -1;
        aastop = -2;
        corestart = -1;
        corestop = -2;
    }

    out->mw_score = (int)hs1[2];
    out->mw_start = (int)hs1[0];
    out->mw_end = (int)hs1[1];
    out->llr = hs2[2];
    out->llr_start = (int)hs2[0];
    out->llr_end = (int)hs2[1];
    out->vit_maxrun = longestprd;
    out->core_score = hs3[2];
    out->core_start = corestart;
    out->core_end = corestop;
    out->prd_score = prdscore;
    out->prd_start = aastart;
    out->prd_end = aastop;
    out->hmm_all = hmmscore;
    out->hmm_vit = hmmscorev;
    out->fi_numaa = dr.numdisorderedstrict2;
    out->fi_meanhydro = dr.meanhydro;
    out->fi_meancharge = dr.meancharge;
    out->fi_meancombo = dr.meanfi;
    out->fi_maxrun = dr.maxlen;
    out->papa_combo = dr.papamaxscore;
    out->papa_prop = dr.papamaxprop;
    out->papa_fi = dr.papamaxdis;
    out->papa_llr = dr.papamaxllr;
    out->papa_llr2 = dr.papamaxllr2;
    out->papa_center = dr.papamaxcenter;
    dr_free(&dr);
    free(maa);
    free(mp);
}

/* plaac.java:610-647 (per-residue table); needs n >= 1 */
void orc_residue_protein(const orc_params *P, const uint8_t *aa, int n, const orc_residue_out *o, int64_t base)
{
    if (n < 1) return;
    uint8_t *vit = (uint8_t *)malloc((size_t)n);
    double *pp = (double *)malloc(sizeof(double) * 2 * (size_t)n);
    /* hmm1.decodealls :628 -> viterbidecodel + mapdecodel(:4032-4045) */
    orc_viterbi(&P->hmm1, aa, n, vit);
    orc_posterior(P, &P->hmm1, aa, n, pp, pp + n, 1);
    disorderreport dr;
    dr_compute(P, aa, n, &dr);
    for (int i = 0; i < n; i++) {
        int map = 0;
        if (pp[n + i] > pp[i]) map = 1;
        if (o->vit) o->vit[base + i] = vit[i];
        if (o->map) o->map[base + i] = (uint8_t)map;
        if (o->charge) o->charge[base + i] = dr.charge[i];
        if (o->hydro) o->hydro[base + i] = dr.hydro[i];
        if (o->fi) o->fi[base + i] = dr.fi[i];
        if (o->plaac) o->plaac[base + i] = dr.plaacllr[i];
        if (o->papa) o->papa[base + i] = dr.papa[i];
        if (o->fix2) o->fix2[base + i] = dr.fix2[i];
        if (o->plaacx2) o->plaacx2[base + i] = dr.plaacllrx2[i];
        if (o->papax2) o->papax2[base + i] = dr.papax2[i];
        if (o->post_bg) o->post_bg[base + i] = pp[i];
        if (o->post_prd) o->post_prd[base + i] = pp[n + i];
    }
    dr_free(&dr);
    free(vit);
    free(pp);
}

void orc_score_batch(const orc_params *P, const uint8_t *codes, const int64_t *offsets, int64_t nprot, orc_summary *out,
                     int full_jar_work, int nthreads)
{
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (int64_t i = 0; i < nprot; i++)
        orc_score_protein(P, codes + offsets[i], (int)(offsets[i + 1] - offsets[i]), &out[i], full_jar_work);
}

void orc_residue_batch(const orc_params *P, const uint8_t *codes, const int64_t *offsets, int64_t nprot,
                       const orc_residue_out *out, int nthreads)
{
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (int64_t i = 0; i < nprot; i++)
        orc_residue_protein(P, codes + offsets[i], (int)(offsets[i + 1] - offsets[i]), out, offsets[i]);
}

/* Test instrumentation (no reference counterpart): how close does any FoldIndex value that the run scan of
 * plaac.java:5010-5059 looks at come to its threshold 0?  Returns min over proteins and scanned positions i of
 * |fi[i]| * (taps in the window of i), the quantity whose sign the CUDA kernels test on running sums (DESIGN.md,
 * "tie classes"); *at receives the protein index of the minimum. */
double orc_fi_min_margin(const orc_params *P, const uint8_t *codes, const int64_t *offsets, int64_t nprot, int nthreads,
                         int64_t *at)
{
    double best = INFINITY;
    int64_t best_at = -1;
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
#endif
    {
        double mine = INFINITY;
        int64_t mine_at = -1;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 64)
#endif
        for (int64_t p = 0; p < nprot; p++) {
            const int n = (int)(offsets[p + 1] - offsets[p]);
            if (n < 1) continue;
            disorderreport d;
            dr_compute(P, codes + offsets[p], n, &d);
            int w = P->ww1 / 2;
            if (w > n - 1) w = n - 1;
            int halfw = (P->ww1 - 1) / 2;
            if (halfw > n / 2) halfw = n / 2;
            for (int i = halfw; i < n - halfw; i++) {
                const int lo = i - w < 0 ? 0 : i - w, hi = i + w > n - 1 ? n - 1 : i + w;
                const double m = fabs(d.fi[i]) * (double)(hi - lo + 1);
                if (m < mine) {
                    mine = m;
                    mine_at = p;
                }
            }
            dr_free(&d);
        }
#ifdef _OPENMP
#pragma omp critical
#endif
        if (mine < best) {
            best = mine;
            best_at = mine_at;
        }
    }
    if (at) *at = best_at;
    return best;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* java.util.Formatter "%.<d>f" (plaac.java:899-945, :638-641).  The JDK does NOT round the exact binary
 * value: Formatter hands the double to sun.misc.FormattedFloatingDecimal, which takes the SHORTEST decimal
 * digit string that identifies the double (the digits of Double.toString) and rounds THAT half-up
 * (applyPrecision).  So 1.0005 (binary 1.000499999999999989...) prints as 1.001 where C prints 1.000.
 * Restated here: shortest round-trip digits via "%.{p}e", then schoolbook half-up on the digit string. */
int orc_java_fmt(char *buf, int buflen, double x, int d)
{
    if (isnan(x)) return snprintf(buf, buflen, "NaN");
    if (isinf(x)) return snprintf(buf, buflen, x > 0 ? "Infinity" : "-Infinity");
    char sci[64];
    double ax = fabs(x);
    int p;
    for (p = 0; p < 17; p++) {
        snprintf(sci, sizeof sci, "%.*e", p, ax);
        if (strtod(sci, NULL) == ax) break;
    }
    char digits[32];
    int nd = 0;
    const char *q = sci;
    for (; *q && *q != 'e'; q++)
        if (*q != '.') digits[nd++] = *q;
    int point = atoi(q + 1) + 1; /* digits before the decimal point */
    int keep = point + d;
    char out[400];
    int no = 0;
    if (keep <= 0) {
        for (int i = 0; i < d + 1; i++) out[no++] = '0';
        if (keep == 0 && nd > 0 && digits[0] >= '5') out[no - 1] = '1';
    } else {
        for (int i = 0; i < keep; i++) out[no++] = i < nd ? digits[i] : '0';
        if (keep < nd && digits[keep] >= '5') {
            int i = no - 1;
            while (i >= 0 && out[i] == '9') out[i--] = '0';
            if (i >= 0)
                out[i]++;
            else {
                memmove(out + 1, out, (size_t)no);
                out[0] = '1';
                no++;
            }
        }
        while (no - d <= 0) {
            memmove(out + 1, out, (size_t)no);
            out[0] = '0';
            no++;
        }
    }
    out[no] = 0;
    int intd = no - d;
    if (d > 0)
        return snprintf(buf, buflen, "%s%.*s.%s", signbit(x) ? "-" : "", intd, out, out + intd);
    return snprintf(buf, buflen, "%s%s", signbit(x) ? "-" : "", out);
}
