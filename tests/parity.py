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


PAPA_FIELDS = ("papa_center", "papa_combo", "papa_prop", "papa_fi", "papa_llr", "papa_llr2")


def papa_tie_class_ok(P, codes, offsets, i, got_row, rtol=RTOL):
    """Documented exact-tie class (DESIGN.md "tie classes"): PAPAcen is the first strict maximum of the
    twice-smoothed PAPA track.  On plateaus (poly-Q, perfect repeats >= 2*ww-1 long) the jar's own choice
    is decided by its rounding noise, so a different centre is accepted iff, in the ORACLE's tracks, it
    attains the maximum within tolerance and satisfies the same FoldIndex gate, and the four values
    reported at the centre match the oracle's tracks at that centre.  The gate `fix2 < 0` (:4944) is itself a
    threshold on a value that is legitimately ~0: a centre whose |fix2| is below the tolerance (1e-9 * scale) may
    be gated either way (found by tests/test_gpu_fuzz.py: the jar accepted a centre with fix2 = -1.3e-16)."""
    from oracle import orc

    lo, hi = int(offsets[i]), int(offsets[i + 1])
    c, o = orc.pack([codes[lo:hi]])
    tr = orc.residue_batch(P, c, o)
    k = int(got_row["papa_center"])
    n = hi - lo
    h = (P.ww2 - 1) // 2
    if not (h <= k < n - h):
        return False
    px, fx = tr["papax2"], tr["fix2"]
    gate_tol = rtol * SCALE["fix2"]
    with np.errstate(invalid="ignore"):
        inrange = np.zeros(n, bool)
        inrange[h:n - h] = True
        inrange &= ~np.isnan(px)
        valid = inrange & (fx < gate_tol)      # centres that may pass the gate
        firm = inrange & (fx < -gate_tol)      # centres that certainly pass it
    if not valid[k]:
        return False
    best = px[firm].max() if firm.any() else -np.inf
    if not (px[k] >= best - rtol * max(abs(best), SCALE["papa_prop"])):
        return False
    want = dict(papa_combo=px[k], papa_prop=px[k], papa_fi=fx[k], papa_llr=tr["plaac"][k], papa_llr2=tr["plaacx2"][k])
    return all(bool(close(got_row[f], v, SCALE[f], rtol)) for f, v in want.items())


def compare_with_tie_classes(P, codes, offsets, got, ref, int_fields, dbl_fields, rtol=RTOL):
    """compare_summaries + the PAPA tie-class rule.  Returns (mismatches, n_tie_class_rows)."""
    idx = np.nonzero(got["papa_center"] != ref["papa_center"])[0]
    ties = [int(i) for i in idx if got["papa_center"][i] >= 0 and papa_tie_class_ok(P, codes, offsets, int(i), got[i], rtol)]
    # the other direction of the gate: the jar accepted a centre whose fix2 is ~0 and the candidate found none at all
    for i in idx:
        if got["papa_center"][i] < 0 and ref["papa_center"][i] >= 0 and abs(ref["papa_fi"][i]) <= rtol * SCALE["fix2"]:
            ties.append(int(i))
    if ties:
        got = got.copy()
        for f in PAPA_FIELDS:
            got[f][ties] = ref[f][ties]
    return compare_summaries(got, ref, int_fields, dbl_fields, rtol), len(ties)
