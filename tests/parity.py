"""Shared comparison rules for the parity tests (SURVEY.md App. C / BASELINE.json north_star).

Integers (residue codes, window positions, Viterbi/PrD boundaries, run lengths, counts) must be
bit-exact.  Doubles must agree within 1e-9 relative, with a per-column floor so that values that
are legitimately ~0 (FoldIndex-type scores) are not held to an impossible relative bound:
    |x - y| <= 1e-9 * max(|y|, scale[column]).
Columns evaluated in reference operation order are additionally expected to be (nearly) bit-exact.
"""
import numpy as np

RTOL = 1e-9
SCALE = {
    "llr": 1.0, "core_score": 1.0, "prd_score": 1.0, "hmm_all": 1.0, "hmm_vit": 1.0,
    "fi_meanhydro": 0.1, "fi_meancharge": 0.1, "fi_meancombo": 0.1,
    "papa_combo": 0.1, "papa_prop": 0.1, "papa_fi": 0.1, "papa_llr": 0.1, "papa_llr2": 0.1,
    # per-residue tracks
    "charge": 0.1, "hydro": 0.1, "fi": 0.1, "plaac": 0.1, "papa": 0.1, "fix2": 0.1, "plaacx2": 0.1, "papax2": 0.1,
    "post_bg": 1.0, "post_prd": 1.0,
}
# reference-order columns: same operation order as the jar => expect <= 1e-12 relative (libm/exp aside)
REF_ORDER = ("llr", "core_score", "prd_score", "hmm_all", "hmm_vit", "fi_meanhydro", "fi_meancharge", "fi_meancombo",
             "papa_llr")


def close(x, y, scale, rtol=RTOL):
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    same_nan = np.isnan(x) & np.isnan(y)
    same_inf = np.isinf(y) & (x == y)
    with np.errstate(invalid="ignore"):
        ok = np.abs(x - y) <= rtol * np.maximum(np.abs(y), scale)
    return ok | same_nan | same_inf


def compare_summaries(got, ref, int_fields, dbl_fields, rtol=RTOL):
    """Returns a list of human-readable mismatches (empty = parity)."""
    bad = []
    for f in int_fields:
        idx = np.nonzero(got[f] != ref[f])[0]
        for i in idx[:5]:
            bad.append(f"int {f}[{i}]: got {got[f][i]} want {ref[f][i]} (len {ref['prot_len'][i]})")
        if len(idx) > 5:
            bad.append(f"int {f}: {len(idx)} mismatches in total")
    for f in dbl_fields:
        ok = close(got[f], ref[f], SCALE[f], rtol)
        idx = np.nonzero(~ok)[0]
        for i in idx[:5]:
            bad.append(f"dbl {f}[{i}]: got {got[f][i]!r} want {ref[f][i]!r} (len {ref['prot_len'][i]})")
        if len(idx) > 5:
            bad.append(f"dbl {f}: {len(idx)} mismatches in total")
    return bad


def max_rel(got, ref, f):
    x, y = np.asarray(got[f]), np.asarray(ref[f])
    m = np.isfinite(x) & np.isfinite(y)
    if not m.any():
        return 0.0
    return float(np.max(np.abs(x[m] - y[m]) / np.maximum(np.abs(y[m]), SCALE[f])))
