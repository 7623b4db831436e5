"""CPU tests of the oracle itself (no GPU): the C restatement against SURVEY.md
Appendix B's known answers, against the committed golden fixture, and against the
independent pure-Python restatement."""
import math

import numpy as np
import pytest

from oracle import orc
from oracle import plaac_oracle_py as opy
from tests import synth

# SURVEY.md Appendix B (1-based positions, 3 decimals) -- derived by an independent restatement.
APPENDIX_B = {
    "Sup35p": dict(MW=38, MWs=4, MWe=83, LLR=51.215, LLRs=5, LLRe=64, VITmaxrun=133, CORE=51.215, COREs=5, COREe=64,
                   PRD=89.773, PRDs=1, PRDe=133, PROTlen=685, HMMall=81.820, HMMvit=79.598, FInumaa=366,
                   FImeanhydro=0.411, FImeancharge=0.004, FImeancombo=-0.010, FImaxrun=262, PAPAprop=0.100,
                   PAPAfi=-0.423, PAPAllr=0.861, PAPAllr2=0.814, PAPAcen=25),
    "Ure2p": dict(MW=38, MWs=1, MWe=80, LLR=29.225, LLRs=17, LLRe=76, VITmaxrun=89, CORE=29.225, COREs=17, COREe=76,
                  PRD=37.909, PRDs=1, PRDe=89, PROTlen=354, HMMall=30.842, HMMvit=28.910, FInumaa=107,
                  FImeanhydro=0.439, FImeancharge=0.017, FImeancombo=0.054, FImaxrun=107, PAPAprop=0.103,
                  PAPAfi=-0.130, PAPAllr=0.228, PAPAllr2=0.317, PAPAcen=21),
    "Rnq1p": dict(MW=43, MWs=238, MWe=317, LLR=47.459, LLRs=218, LLRe=277, VITmaxrun=282, CORE=47.459, COREs=218,
                  COREe=277, PRD=165.340, PRDs=124, PRDe=405, PROTlen=405, HMMall=154.981, HMMvit=152.584,
                  FInumaa=286, FImeanhydro=0.374, FImeancharge=0.007, FImeancombo=-0.118, FImaxrun=205,
                  PAPAprop=0.141, PAPAfi=-0.324, PAPAllr=0.911, PAPAllr2=0.791, PAPAcen=234),
    "Mot3p": dict(MW=38, MWs=83, MWe=162, LLR=40.751, LLRs=98, LLRe=157, VITmaxrun=296, CORE=40.751, COREs=98,
                  COREe=157, PRD=109.847, PRDs=1, PRDe=296, PROTlen=490, HMMall=115.787, HMMvit=107.777, FInumaa=416,
                  FImeanhydro=0.373, FImeancharge=-0.010, FImeancombo=-0.122, FImaxrun=115, PAPAprop=0.102,
                  PAPAfi=-0.116, PAPAllr=0.502, PAPAllr2=0.422, PAPAcen=109),
}
APPENDIX_B_LLR = dict(A=-0.12257, C=-1.74969, D=-1.27456, E=-2.12398, F=-0.55278, G=0.42322, H=-0.18129, I=-1.17725,
                      K=-1.49928, L=-1.28078, M=0.36281, N=1.43732, P=0.15739, Q=1.34371, R=-0.57425, S=0.20080,
                      T=-0.43249, V=-1.03644, W=-1.89062, Y=0.51224, X=0.0)


def _score_golden(golden, **kw):
    P = orc.make_params(**kw)
    codes, offs = orc.pack([orc.encode(p["seq"]) for p in golden["proteins"]])
    return P, codes, offs, orc.score_batch(P, codes, offs)


def test_appendix_b(golden):
    P, _, _, out = _score_golden(golden)
    for ch, v in APPENDIX_B_LLR.items():
        assert round(P.llr[orc.AANAMES.index(ch)], 5) == pytest.approx(v, abs=1e-9)
    for prot, r in zip(golden["proteins"], out):
        k = APPENDIX_B[prot["name"]]
        got = dict(MW=r["mw_score"], MWs=r["mw_start"] + 1, MWe=r["mw_end"] + 1, LLRs=r["llr_start"] + 1,
                   LLRe=r["llr_end"] + 1, VITmaxrun=r["vit_maxrun"], COREs=r["core_start"] + 1,
                   COREe=r["core_end"] + 1, PRDs=r["prd_start"] + 1, PRDe=r["prd_end"] + 1, PROTlen=r["prot_len"],
                   FInumaa=r["fi_numaa"], FImaxrun=r["fi_maxrun"], PAPAcen=r["papa_center"] + 1)
        for name, val in got.items():
            assert int(val) == k[name], (prot["name"], name)
        dbl = dict(LLR="llr", CORE="core_score", PRD="prd_score", HMMall="hmm_all", HMMvit="hmm_vit",
                   FImeanhydro="fi_meanhydro", FImeancharge="fi_meancharge", FImeancombo="fi_meancombo",
                   PAPAprop="papa_prop", PAPAfi="papa_fi", PAPAllr="papa_llr", PAPAllr2="papa_llr2")
        for name, f in dbl.items():
            assert float(orc.java_fmt(r[f], 3)) == pytest.approx(k[name], abs=5e-4), (prot["name"], name)


def test_golden_fixture(golden):
    _, codes, offs, out = _score_golden(golden)
    P = orc.make_params()
    res = orc.residue_batch(P, codes, offs)
    for i, prot in enumerate(golden["proteins"]):
        for k, v in prot["summary"].items():
            if isinstance(v, float) and math.isnan(v):
                assert math.isnan(out[i][k])
            else:
                assert out[i][k] == v, (prot["name"], k)
        lo = int(offs[i])
        for k, vals in prot["residue"].items():
            for j, v in zip(prot["residue_idx"], vals):
                x = res[k][lo + j]
                assert (v is None and x != x) or x == v, (prot["name"], k, j)


def _same(a, b):
    return (a != a and b != b) or a == b


def test_c_vs_python_restatement(golden):
    """Two independent restatements must agree bit for bit (same libm underneath)."""
    P = orc.make_params()
    Q = opy.Params()
    assert list(P.llr) == Q.llr and list(P.loglut) == Q.loglut
    assert [list(r) for r in P.hmm1.le] == Q.hmm1["le"]
    seqs = [orc.encode(p["seq"]) for p in golden["proteins"][:2]]
    codes, offs = synth.edge_cases()
    seqs += [codes[offs[i]:offs[i + 1]] for i in range(len(offs) - 1)]
    codes, offs = orc.pack(seqs)
    out = orc.score_batch(P, codes, offs)
    res = orc.residue_batch(P, codes, offs)
    for i, s in enumerate(seqs):
        ref = opy.score_protein(Q, [int(x) for x in s])
        for k, v in ref.items():
            assert _same(float(out[i][k]), float(v)), (i, len(s), k, out[i][k], v)
        if len(s) <= 130:
            tr = opy.residue_protein(Q, [int(x) for x in s])
            lo = int(offs[i])
            for k, vals in tr.items():
                for j, v in enumerate(vals):
                    assert _same(float(res[k][lo + j]), float(v)), (i, len(s), k, j)


@pytest.mark.parametrize("kw", [dict(alpha=0.5, bg_counts=np.arange(22) * 1000.0 + 500),
                                dict(core_len=30, ww1=21, ww2=31), dict(ww1=40, ww2=40), dict(adjust_prolines=False)])
def test_c_vs_python_other_params(kw):
    P = orc.make_params(**kw)
    Q = opy.Params(**kw)
    codes, offs = synth.proteome(12, seed=3, median=150, sigma=0.5)
    out = orc.score_batch(P, codes, offs)
    for i in range(len(offs) - 1):
        ref = opy.score_protein(Q, [int(x) for x in codes[offs[i]:offs[i + 1]]])
        for k, v in ref.items():
            assert _same(float(out[i][k]), float(v)), (i, k, out[i][k], v)


def test_hss2_against_bruteforce():
    """The reference author's own check pattern (plaac.java:772-791): fast hss2 vs brute force."""
    rng = np.random.default_rng(0)
    L = orc.lib()
    for _ in range(200):
        n = int(rng.integers(1, 120))
        w = int(rng.integers(1, 90))
        x = rng.normal(size=n)
        out = np.zeros(3)
        L.orc_hss2(x.ctypes.data, n, w, w, out.ctypes.data)
        if w > n:
            assert out[0] == -1 and out[1] == -2 and out[2] == -np.inf
            continue
        ps = np.zeros(n + 1)
        for i in range(n):
            ps[i + 1] = ps[i] + x[i]
        sums = [ps[i + w] - ps[i] for i in range(n - w + 1)]
        assert out[2] == max(sums)
        assert int(out[0]) == int(np.argmax(sums)) and int(out[1]) == int(out[0]) + w - 1


def test_lut_lse_properties():
    P = orc.make_params()
    L = orc.lib()
    import ctypes as C
    f = lambda a, b: L.orc_logeapeb(C.byref(P), a, b)
    assert f(-np.inf, -np.inf) == -np.inf
    assert f(-np.inf, -3.5) == -3.5 and f(-3.5, -np.inf) == -3.5
    assert f(1.0, 1.0) == 1.0 + math.log(2)
    assert f(0.0, -40.0) == 0.0 and f(0.0, -30.0) > 0.0
    rng = np.random.default_rng(1)
    for a, b in rng.uniform(-30, 0, size=(2000, 2)):
        exact = np.logaddexp(a, b)
        assert -1e-14 <= f(a, b) - exact < 3.2e-6  # LUT over-estimates by at most 3.12e-6
        assert f(a, b) == f(b, a)


def test_java_fmt_half_up():
    assert orc.java_fmt(0.0625, 3) == "0.063"     # C printf would give 0.062
    assert orc.java_fmt(-0.0625, 3) == "-0.063"
    assert orc.java_fmt(0.1875, 3) == "0.188"
    assert orc.java_fmt(2.5, 0) == "3"
    assert orc.java_fmt(0.1, 3) == "0.100"
    assert orc.java_fmt(float("nan"), 3) == "NaN"
    assert orc.java_fmt(1234.56749999, 3) == "1234.567"


def test_short_and_degenerate_rows():
    P = orc.make_params()
    codes, offs = orc.pack([np.array([14], np.uint8), np.full(59, 12, np.uint8), np.full(60, 12, np.uint8)])
    out = orc.score_batch(P, codes, offs)
    # n < c: LLR = -Inf (printed NaN), start -1 / end -2 (plaac.java:1211-1216)
    assert out[0]["llr"] == -np.inf and out[0]["llr_start"] == -1 and out[0]["llr_end"] == -2
    assert out[1]["llr"] == -np.inf and math.isnan(out[1]["core_score"]) and out[1]["prd_score"] == 0.0
    assert out[1]["vit_maxrun"] == 59 and out[1]["core_start"] == -1 and out[1]["core_end"] == -2
    assert out[2]["llr_start"] == 0 and out[2]["llr_end"] == 59 and out[2]["core_start"] == 0
    assert out[2]["prd_start"] == 0 and out[2]["prd_end"] == 59
    assert out[0]["mw_score"] == 1 and out[0]["mw_start"] == 0 and out[0]["mw_end"] == 0
    # n <= 40: no PAPA centre (plaac.java:4941 range is empty)
    assert out[0]["papa_center"] == -1 and math.isnan(out[0]["papa_prop"]) and out[0]["papa_combo"] == -np.inf
