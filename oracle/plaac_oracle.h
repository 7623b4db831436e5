/*
 * plaac_oracle.h -- CPU restatement of the PLAAC per-protein scoring path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (plaac_b200/, the
 * C-ABI library, the host CLI) may include, link or call this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it, and only as the checker / CPU baseline.
 *
 * PARITY PINNED AGAINST THE REFERENCE'S OWN BYTECODE (with one stated caveat).  The reference ships no tests
 * and no golden outputs, and there is no JVM in this image, so plaac.jar cannot be *launched*.  Its class files
 * can be interpreted, though: tests/golden/minijvm.py executes web/bin/plaac.jar's plaac.main() unmodified on
 * the reference's own cli/example/four_classic_prions.fasta and on 44 edge-case proteins (summary table and
 * `-p all` per-residue table, default and non-default flags) and tests/golden/jar_vectors.json.gz holds every
 * value the jar handed to System.out.format() at full precision.  This restatement reproduces ALL of them bit
 * for bit (tests/test_jar_vectors.py).  Caveat: the JDK natives behind the bytecode (Math.log/exp/floor...) are
 * this box's libm in both cases; a real JVM's Math.log/exp may differ from libm in the last ulp of the tables.
 * The restatement follows plaac.java line by line in the reference's own operation order (sequential psum,
 * 41-tap window loops, lookup-table log-sum-exp); it is also cross-checked against an independent pure-Python
 * restatement (oracle/plaac_oracle_py.py) and the provisional known answers of SURVEY.md Appendix B.
 *
 * Every function cites the plaac.java lines it follows.
 */
#ifndef PLAAC_ORACLE_H
#define PLAAC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NAA 22
#define ORC_LUTLEN 4000 /* plaac.java:34  loglutlength = 40*100 */

typedef struct {
    double lt[2][2]; /* ltprob  plaac.java:2898 */
    double le[2][ORC_NAA]; /* leprob */
    double li[2]; /* liprob */
    double lf[2]; /* lfprob (== 0 for both PLAAC HMMs, plaac.java:2904-2922) */
} orc_hmm;

typedef struct {
    int32_t core_len, ww1, ww2, ww3, adjust_prolines;
    double alpha;
    double fg[ORC_NAA], bg[ORC_NAA], bgscer[ORC_NAA], bgthis[ORC_NAA], llr[ORC_NAA];
    orc_hmm hmm1, hmm0;
    double papa_lod[ORC_NAA], hydro2[ORC_NAA], charge[ORC_NAA], fi_cc[3];
    double loglut[ORC_LUTLEN + 1];
    double ln2;
} orc_params;

/* Per-protein record: what plaac.java:899-945 prints, before +1 / inf2nan.
 * Same field order as the product's plaac_summary so tests can memcmp ints. */
typedef struct {
    int32_t mw_score, mw_start, mw_end;
    int32_t llr_start, llr_end;
    int32_t vit_maxrun;
    int32_t core_start, core_end, prd_start, prd_end;
    int32_t prot_len;
    int32_t fi_numaa, fi_maxrun;
    int32_t papa_center;
    double llr, core_score, prd_score, hmm_all, hmm_vit;
    double fi_meanhydro, fi_meancharge, fi_meancombo;
    double papa_combo, papa_prop, papa_fi, papa_llr, papa_llr2;
} orc_summary;

/* Per-residue tracks (plaac.java:603-605, 637-641); any pointer may be NULL. */
typedef struct {
    uint8_t *vit, *map;
    double *charge, *hydro, *fi, *plaac, *papa, *fix2, *plaacx2, *papax2;
    double *post_bg, *post_prd;
} orc_residue_out;

/* plaac.java:1508-1534 */
int orc_aatoint(int ch);

/* Parameter chain of plaac.java:279-299 (LUT, lodpapa1), :449-500 (fg/bg/llr),
 * :968-1001 + :2893-2935 (HMMs).  bg_counts = the 22 numbers `bgf` holds at
 * :374-384 (from -B, -b or -i); NULL means all zero.  fg_freq = NULL for the
 * built-in prd_freq_scer_28 (:269). */
void orc_params_init(orc_params *P, double alpha, const double *bg_counts, const double *fg_freq,
                     int core_len, int ww1, int ww2, int ww3, int adjust_prolines);

/* plaac.java:1024-1047 */
double orc_logeapeb(const orc_params *P, double a, double b);

/* plaac.java:1206-1257 */
void orc_hss2(const double *seq, int n, int minlength, int maxlength, double out[3]);

/* plaac.java:3077-3121; returns lviterbiprob */
double orc_viterbi(const orc_hmm *h, const uint8_t *aa, int n, uint8_t *path);

/* plaac.java:3349-3411; pp0/pp1 may be NULL (then backward pass is skipped
 * when want_posterior == 0); returns lmarginalprob */
double orc_posterior(const orc_params *P, const orc_hmm *h, const uint8_t *aa, int n, double *pp0, double *pp1,
                     int want_posterior);

/* plaac.java:2585-2622 and :2626-2662 (mergeme >= 0 selects the second) */
void orc_slidingaverage(const double *arr, int n, int ww, int shrink, int weight, int mergeme, const uint8_t *seq,
                        double *sa);

/* One protein, summary mode: the loop body of scoreallfastas, plaac.java:755-948.
 * full_jar_work != 0 also runs the posterior/backward passes the jar computes
 * but never prints (decodeall :3261-3284), for the CPU-baseline timing. */
void orc_score_protein(const orc_params *P, const uint8_t *aa, int n, orc_summary *out, int full_jar_work);

/* One protein, per-residue mode: loop body of plotsomefastas, plaac.java:610-647. */
void orc_residue_protein(const orc_params *P, const uint8_t *aa, int n, const orc_residue_out *out, int64_t base);

/* Batch drivers (OpenMP over proteins when nthreads > 1). */
void orc_score_batch(const orc_params *P, const uint8_t *codes, const int64_t *offsets, int64_t nprot,
                     orc_summary *out, int full_jar_work, int nthreads);
void orc_residue_batch(const orc_params *P, const uint8_t *codes, const int64_t *offsets, int64_t nprot,
                       const orc_residue_out *out, int nthreads);

/* Java `%.<d>f` (HALF_UP on the exact binary value; NaN -> "NaN", +-Inf ->
 * "Infinity"/"-Infinity"), used to print rows the way plaac.java:899-945 does. */
int orc_java_fmt(char *buf, int buflen, double x, int decimals);

/* test instrumentation: min over the FoldIndex values the run scan looks at of |fi| * window taps (see the .c file) */
double orc_fi_min_margin(const orc_params *P, const uint8_t *codes, const int64_t *offsets, int64_t nprot, int nthreads,
                         int64_t *at);
int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
