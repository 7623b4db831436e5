// Host-side parameter chain behind plaac_params_init() (include/plaac_cuda.h).
// Product code: independent of oracle/.  Follows plaac.java main :310-518.
#include <cmath>
#include <cstring>

#include "../../include/plaac_cuda.h"

namespace {

// plaac.java:37-60
const double kCharge[PLAAC_NAA] = {0, 0, 0, 1, 1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 0, -1, 0, 0, 0, 0, 0, 0};
// plaac.java:64-87 Kyte-Doolittle hydropathy
const double kHydro[PLAAC_NAA] = {0.0, 1.8,  2.5,  -3.5, -3.5, 2.8,  -0.4, -3.2, 4.5,  -3.9, 3.8,
                                  1.9, -3.5, -1.6, -3.5, -4.5, -0.8, -0.7, 4.2,  -0.9, -1.3, 0.0};
// plaac.java:206-229 PAPA odds ratios (Toombs et al. 2010)
const double kPapaOdds[PLAAC_NAA] = {0.0,        0.67267686, 1.5146198, 0.27887323, 0.5460614,  2.313433,  0.96153843, 0.75686276,
                                     2.2562358,  0.20664589, 0.9607843, 1.9615384,  1.0836071,  0.30196398, 1.0716166, 0.6664044,
                                     1.1432927,  0.8917492,  2.2562358, 1.9478673,  2.1785367,  0.0};
// plaac.java:261-262 S. cerevisiae background
const double kBgScer[PLAAC_NAA] = {0,      0.0550, 0.0126, 0.0586, 0.0655, 0.0441, 0.0498, 0.0217, 0.0655, 0.0735, 0.0950,
                                   0.0207, 0.0615, 0.0438, 0.0396, 0.0444, 0.0899, 0.0592, 0.0556, 0.0104, 0.0337, 0};
// plaac.java:269-270 prion-domain composition from 28 domains
const double kPrd28[PLAAC_NAA] = {0,       0.04865, 0.00219, 0.01638, 0.00783, 0.02537, 0.07603, 0.0181,  0.02018, 0.01641, 0.02639,
                                  0.02975, 0.25885, 0.05126, 0.15178, 0.025,   0.10988, 0.03841, 0.01972, 0.00157, 0.05624, 0};

struct Vec22 {
    double v[PLAAC_NAA];
};

// plaac.java:1933-1941
Vec22 normalized(const Vec22& a)
{
    double total = 0;
    for (double x : a.v) total = total + x;
    if (total < 1e-12) total = 1;
    Vec22 r;
    for (int i = 0; i < PLAAC_NAA; i++) r.v[i] = a.v[i] / total;
    return r;
}

}  // namespace

extern "C" int plaac_params_init(plaac_params* out, double alpha, const double* bg_counts, const double* fg_freq,
                                 int core_len, int ww1, int ww2, int ww3, int adjust_prolines, double* info)
{
    if (!out) return PLAAC_E_INVALID;
    plaac_params& P = *out;
    std::memset(&P, 0, sizeof(P));
    P.core_len = core_len;
    P.ww1 = ww1;
    P.ww2 = ww2;
    P.ww3 = ww3;
    P.adjust_prolines = adjust_prolines ? 1 : 0;
    P.mw_window = 80;  // :766
    if (alpha > 1 || alpha < 0) alpha = 1.0;  // :444-447

    P.ln2 = std::log(2.0);
    P.big_neg = -1000000.0;
    for (int i = 0; i < PLAAC_LUT_LEN; i++) P.loglut[i] = std::log(1.0 + std::exp(-i / 100.0));  // :283
    for (int k = 1; k <= 20; k++) P.papa_lod[k] = std::log(kPapaOdds[k]);                         // :288-291
    const double ninth = 1.0 / 9.0;
    for (int k = 0; k < PLAAC_NAA; k++) {
        P.hydro2[k] = ninth * kHydro[k] + 0.5;  // :90 (compiled without FMA contraction)
        P.charge[k] = kCharge[k];
    }
    P.fi_cc[0] = 2.785;
    P.fi_cc[1] = -1;
    P.fi_cc[2] = -1.151;

    Vec22 scer, fg0, bgin;
    std::memcpy(scer.v, kBgScer, sizeof(kBgScer));
    std::memcpy(fg0.v, fg_freq ? fg_freq : kPrd28, sizeof(fg0.v));
    if (bg_counts)
        std::memcpy(bgin.v, bg_counts, sizeof(bgin.v));
    else
        std::memset(bgin.v, 0, sizeof(bgin.v));
    const Vec22 bgscer = normalized(scer);  // :310
    fg0.v[0] = fg0.v[21] = 0;               // :449
    Vec22 fgn = normalized(fg0);            // :452
    bgin.v[0] = bgin.v[21] = 0;             // :454
    const Vec22 bgthis = normalized(bgin);  // :456
    Vec22 mix;
    for (int i = 0; i < PLAAC_NAA; i++) mix.v[i] = alpha * bgscer.v[i] + (1 - alpha) * bgthis.v[i];  // :458
    Vec22 combo = normalized(mix);
    const double epsx = 0.00001;  // :490-496
    fgn.v[0] = fgn.v[21] = epsx;
    combo.v[0] = combo.v[21] = epsx;
    const Vec22 fg = normalized(fgn);
    const Vec22 bg = normalized(combo);
    for (int j = 1; j < 21; j++) P.llr[j] = std::log(fg.v[j] / bg.v[j]);  // :500

    // prionhmm1 :968-981 (emissions are normalised once more), hmm.initialize :2893-2935
    const Vec22 e_bg = normalized(bg), e_fg = normalized(fg);
    const double tmat[2][2] = {{99.9 / 100, 0.1 / 100}, {2.0 / 100, 98.0 / 100}};
    const double imat[2] = {0.9524, 0.0476};
    bool freeend = true;
    double fprob[2];
    for (int i = 0; i < 2; i++) {
        double rs = 0;
        for (int j = 0; j < 2; j++) {
            P.lt[i][j] = std::log(tmat[i][j]);
            rs = rs + tmat[i][j];
        }
        P.li[i] = std::log(imat[i]);
        fprob[i] = std::fmax(0.0, 1.0 - rs);
        if (fprob[i] > 0.0001) freeend = false;
    }
    for (int i = 0; i < 2; i++) P.lf[i] = std::log(freeend ? 1.0 : fprob[i]);
    for (int j = 0; j < PLAAC_NAA; j++) {
        P.le[0][j] = std::log(e_bg.v[j]);
        P.le[1][j] = std::log(e_fg.v[j]);
        P.le0[j] = std::log(e_bg.v[j]);  // prionhmm0 :988-1001 emits normalize(bg) in both states
    }
    if (info) {
        std::memcpy(info, fg.v, sizeof(fg.v));
        std::memcpy(info + 22, bgscer.v, sizeof(fg.v));
        std::memcpy(info + 44, bgthis.v, sizeof(fg.v));
        std::memcpy(info + 66, bg.v, sizeof(fg.v));
    }
    return PLAAC_OK;
}
