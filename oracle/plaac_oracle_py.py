"""Pure-Python second restatement of the PLAAC scoring path (small inputs only).

TEST INFRASTRUCTURE ONLY.  The C oracle is pinned against the jar's own bytecode (tests/test_jar_vectors.py);
this file is the second, independent restatement it is cross-checked with.  Written independently of oracle/plaac_oracle.c, from
cli/src/plaac.java directly, with plain Python floats (IEEE double, one
rounding per operator, no numpy reductions) so the two restatements
cross-check each other.  Citations are plaac.java line numbers.
"""
from __future__ import annotations

import math

NEG_INF = float("-inf")
NAN = float("nan")
AANAMES = "XACDEFGHIKLMNPQRSTVWY*"  # :26

AACHARGE = [0, 0, 0, 1, 1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 0, -1, 0, 0, 0, 0, 0, 0]  # :37-60
AAHYDRO = [0.0, 1.8, 2.5, -3.5, -3.5, 2.8, -0.4, -3.2, 4.5, -3.9, 3.8, 1.9, -3.5, -1.6, -3.5, -4.5,
           -0.8, -0.7, 4.2, -0.9, -1.3, 0.0]  # :64-87
ODPAPA1 = [0.0, 0.67267686, 1.5146198, 0.27887323, 0.5460614, 2.313433, 0.96153843, 0.75686276, 2.2562358,
           0.20664589, 0.9607843, 1.9615384, 1.0836071, 0.30196398, 1.0716166, 0.6664044, 1.1432927,
           0.8917492, 2.2562358, 1.9478673, 2.1785367, 0.0]  # :206-229
BG_FREQ_SCER = [0, 0.0550, 0.0126, 0.0586, 0.0655, 0.0441, 0.0498, 0.0217, 0.0655, 0.0735, 0.0950, 0.0207,
                0.0615, 0.0438, 0.0396, 0.0444, 0.0899, 0.0592, 0.0556, 0.0104, 0.0337, 0]  # :261
PRD_FREQ_SCER_28 = [0, 0.04865, 0.00219, 0.01638, 0.00783, 0.02537, 0.07603, 0.0181, 0.02018, 0.01641, 0.02639,
                    0.02975, 0.25885, 0.05126, 0.15178, 0.025, 0.10988, 0.03841, 0.01972, 0.00157, 0.05624, 0]  # :269


def aatoint(ch: str) -> int:  # :1508-1534
    i = AANAMES.find(ch.upper()) if ch != "*" else 21
    if ch in "xX":
        return 0
    return i if i > 0 else 0


def encode(seq: str) -> list[int]:  # :758-759
    if seq.endswith("*"):
        seq = seq[:-1]
    return [aatoint(c) for c in seq]


def _seqsum(v):
    s = 0.0
    for x in v:
        s = s + x
    return s


def normalize(v):  # :1933-1941
    sm = 1.0 * _seqsum(v)
    if sm < 0.000000000001:
        sm = 1
    return [x / sm for x in v]


def _log(x):
    return math.log(x) if x > 0 else NEG_INF


class Params:
    def __init__(self, alpha=1.0, bg_counts=None, fg_freq=None, core_len=60, ww1=41, ww2=41, ww3=None,
                 adjust_prolines=True):
        if alpha > 1 or alpha < 0:  # :444
            alpha = 1.0
        self.core_len, self.ww1, self.ww2 = core_len, ww1, ww2
        self.ww3 = ww2 if ww3 is None else ww3
        self.adjust_prolines = adjust_prolines
        self.ln2 = math.log(2)
        self.loglut = [math.log(1.0 + math.exp(-i / 100.0)) for i in range(4001)]  # :282
        self.papa_lod = [0.0] * 22
        for k in range(1, 21):
            self.papa_lod[k] = math.log(ODPAPA1[k])
        a = 1.0 / 9.0
        self.hydro2 = [a * h + 0.5 for h in AAHYDRO]  # :90
        self.charge = [float(c) for c in AACHARGE]
        self.cc = [2.785, -1, -1.151]
        bgscer = normalize(BG_FREQ_SCER)
        fgfreq = list(PRD_FREQ_SCER_28 if fg_freq is None else fg_freq)
        bgf = [0.0] * 22 if bg_counts is None else [float(x) for x in bg_counts]
        fgfreq[0] = 0
        fgfreq[21] = 0
        fgfreq = normalize(fgfreq)
        bgf[0] = 0
        bgf[21] = 0
        bgthis = normalize(bgf)
        bgcombo = normalize([alpha * x + (1 - alpha) * y for x, y in zip(bgscer, bgthis)])
        eps = 0.00001
        fgfreq[0] = fgfreq[21] = eps
        bgcombo[0] = bgcombo[21] = eps
        self.fg = normalize(fgfreq)
        self.bg = normalize(bgcombo)
        self.llr = [0.0] * 22
        for j in range(1, 21):
            self.llr[j] = math.log(self.fg[j] / self.bg[j])
        # :968-1001
        self.hmm1 = dict(
            lt=[[_log(99.9 / 100), _log(0.1 / 100)], [_log(2.0 / 100), _log(98.0 / 100)]],
            li=[_log(0.9524), _log(0.0476)],
            le=[[_log(x) for x in normalize(self.bg)], [_log(x) for x in normalize(self.fg)]],
            lf=[0.0, 0.0],
        )
        e0 = [_log(x) for x in normalize(self.bg)]
        self.hmm0 = dict(lt=[[0.0, NEG_INF], [NEG_INF, 0.0]], li=[0.0, NEG_INF], le=[e0, e0], lf=[0.0, 0.0])

    def lse(self, a, b):  # :1024-1047
        if a > b:
            c = a - b
            if not (c < 40):
                return a
            dex = int(math.floor(100 * c))
            return a + ((100 * c - dex) * self.loglut[dex + 1] + (dex + 1 - 100 * c) * self.loglut[dex])
        elif b > a:
            c = b - a
            if not (c < 40):
                return b
            dex = int(math.floor(100 * c))
            return b + ((100 * c - dex) * self.loglut[dex + 1] + (dex + 1 - 100 * c) * self.loglut[dex])
        return a + self.ln2


def hss_fixed(x, L):
    """hss2(seq, L, L) :1206-1257 specialised to min == max (inner j loop is empty)."""
    n = len(x)
    if L > n:
        return -1, -2, NEG_INF
    ps = [0.0] * (n + 1)
    for i in range(n):
        ps[i + 1] = ps[i] + x[i]
    best, bstart, bstop = ps[L], 0, L - 1
    for i in range(L, n):
        d = ps[i + 1] - ps[i - L + 1]
        if d > best:
            best, bstart, bstop = d, i - L + 1, i
    return bstart, bstop, best


def viterbi(h, aa):  # :3077-3121
    n = len(aa)
    lt, le, li, lf = h["lt"], h["le"], h["li"], h["lf"]
    s = [[li[i] + le[i][aa[0]]] for i in range(2)]
    tb = [[0], [0]]
    for t in range(1, n):
        for i in range(2):
            bd, bs = 0, lt[0][i] + s[0][t - 1]
            if lt[1][i] + s[1][t - 1] > bs:
                bs, bd = lt[1][i] + s[1][t - 1], 1
            s[i].append(bs + le[i][aa[t]])
            tb[i].append(bd)
    bd, bs = 0, s[0][n - 1] + lf[0]
    if s[1][n - 1] + lf[1] > bs:
        bs, bd = s[1][n - 1] + lf[1], 1
    vit = [0] * n
    vit[n - 1] = bd
    for t in range(n - 2, -1, -1):
        vit[t] = tb[vit[t + 1]][t + 1]
    return vit, bs


def posterior(P, h, aa):  # :3349-3411
    n = len(aa)
    lt, le, li, lf = h["lt"], h["le"], h["li"], h["lf"]
    a = [[li[i] + le[i][aa[0]]] + [0.0] * (n - 1) for i in range(2)]
    for t in range(1, n):
        for i in range(2):
            sc = NEG_INF
            for k in range(2):
                sc = P.lse(sc, lt[k][i] + a[k][t - 1])
            a[i][t] = sc + le[i][aa[t]]
    ltot = NEG_INF
    for i in range(2):
        ltot = P.lse(ltot, a[i][n - 1] + lf[i])
    b = [[0.0] * n for _ in range(2)]
    for i in range(2):
        b[i][n - 1] = lf[i]
    for t in range(n - 2, -1, -1):
        for i in range(2):
            sc = NEG_INF
            for k in range(2):
                sc = P.lse(sc, lt[i][k] + b[k][t + 1] + le[k][aa[t + 1]])
            b[i][t] = sc
    lpseq = NEG_INF
    for i in range(2):
        lpseq = P.lse(lpseq, a[i][0] + b[i][0])
    pp = [[math.exp((a[i][t] + b[i][t]) - lpseq) for t in range(n)] for i in range(2)]
    return pp, ltot


def sliding(arr, ww, shrink, weight, mergeme=None, seq=None):  # :2585-2662
    n = len(arr)
    if n == 0:
        return arr
    w = ww // 2
    if w >= n:
        w = n - 1
    sa = [NAN] * n
    lo, hi = (0, n - 1) if shrink else (w, n - w - 1)
    for i in range(lo, hi + 1):
        score = 0.0
        denom = 0.0
        for j in range(-w, w + 1):
            p = i + j
            if 0 <= p < n:
                wt = 1.0
                if weight:
                    wt = 1.0 + min(p, w) + min(n - p - 1, w)
                denom = denom + wt
                if mergeme is not None and seq[p] == mergeme and (
                        (p >= 1 and seq[p - 1] == mergeme) or (p >= 2 and seq[p - 2] == mergeme)):
                    continue
                score = score + wt * arr[p]
        sa[i] = score / denom
    return sa


def disorder(P, aa):  # :4866-5068 (live parts)
    n = len(aa)
    cc = P.cc
    m1 = [P.hydro2[a] for a in aa]
    meanhydro = (1.0 * _seqsum(m1)) / n
    hydro = sliding(m1, P.ww1, True, False)
    m2 = [P.charge[a] for a in aa]
    meancharge = (1.0 * _seqsum(m2)) / n
    charge = sliding(m2, P.ww1, True, False)
    meanfi = cc[2] + cc[1] * abs(meancharge) + cc[0] * meanhydro
    fi = [cc[0] * h + cc[1] * abs(c) + cc[2] for h, c in zip(hydro, charge)]
    m3 = [P.llr[a] for a in aa]
    plaacllr = sliding(m3, P.ww3, True, False)
    m4 = [P.papa_lod[a] for a in aa]
    papa = sliding(m4, P.ww2, True, False, 13, aa) if P.adjust_prolines else sliding(m4, P.ww2, True, False)
    papax2 = sliding(papa, P.ww2, False, True)
    plaacllrx2 = sliding(plaacllr, P.ww3, False, True)
    fix2 = sliding(fi, P.ww1, False, True)
    best, cen = NEG_INF, -1
    for k in range((P.ww2 - 1) // 2, n - (P.ww2 - 1) // 2):
        if papax2[k] > best and fix2[k] < 0:
            cen, best = k, papax2[k]
    halfw = (P.ww1 - 1) // 2
    if halfw > n // 2:
        halfw = n // 2
    numaa, maxlen, i = 0, 0, halfw
    while i < n - halfw:
        if fi[i] < 0:
            st = i
            i += 1
            while i < n - halfw and fi[i] < 0:
                i += 1
            sp = i - 1
            if st == halfw:
                st = 0
            if sp == n - halfw - 1:
                sp = n - 1
            ln = sp - st + 1
            if ln >= 5:
                numaa += ln
                maxlen = max(maxlen, ln)
        else:
            i += 1
    tracks = dict(charge=charge, hydro=hydro, fi=fi, plaac=plaacllr, papa=papa, fix2=fix2, plaacx2=plaacllrx2,
                  papax2=papax2)
    summ = dict(fi_numaa=numaa, fi_maxrun=maxlen, fi_meanhydro=meanhydro, fi_meancharge=meancharge,
                fi_meancombo=meanfi, papa_combo=best, papa_center=cen,
                papa_prop=papax2[cen] if cen >= 0 else NAN, papa_fi=fix2[cen] if cen >= 0 else NAN,
                papa_llr=plaacllr[cen] if cen >= 0 else NAN, papa_llr2=plaacllrx2[cen] if cen >= 0 else NAN)
    return summ, tracks


def score_protein(P, aa):  # :755-948
    n = len(aa)
    c = P.core_len
    out = dict(prot_len=n)
    qn = [1.0 if a in (12, 14) else 0.0 for a in aa]
    L = 80 if n >= 80 else n
    s, e, v = hss_fixed(qn, L)
    out.update(mw_score=int(v), mw_start=s, mw_end=e)
    llrs = [P.llr[a] for a in aa]
    s, e, v = hss_fixed(llrs, c)
    out.update(llr=v, llr_start=s, llr_end=e)
    vit, lv1 = viterbi(P.hmm1, aa)
    _, lm1 = posterior(P, P.hmm1, aa)
    _, lv0 = viterbi(P.hmm0, aa)
    _, lm0 = posterior(P, P.hmm0, aa)
    out.update(hmm_all=lm1 - lm0, hmm_vit=lv1 - lv0)
    ds, _ = disorder(P, aa)
    out.update(ds)
    # longestrun :1787
    mx = cur = 0
    for b in vit:
        cur = cur + 1 if b else 0
        mx = max(mx, cur)
    out["vit_maxrun"] = mx
    masked = [x if b else -1000000.0 for x, b in zip(llrs, vit)]
    s, e, v = hss_fixed(masked, c)
    if v > -1000000.0 / 2:
        a0, a1 = s, e
        while a0 >= 0 and vit[a0] == 1:
            a0 -= 1
        a0 += 1
        while a1 < n and vit[a1] == 1:
            a1 += 1
        a1 -= 1
        prd = 0.0
        for k in range(a0, a1 + 1):
            prd = prd + P.llr[aa[k]]
        out.update(core_score=v, core_start=s, core_end=e, prd_start=a0, prd_end=a1, prd_score=prd)
    else:
        out.update(core_score=NAN, core_start=-1, core_end=-2, prd_start=-1, prd_end=-2, prd_score=0.0)
    return out


def residue_protein(P, aa):  # :610-647
    vit, _ = viterbi(P.hmm1, aa)
    pp, _ = posterior(P, P.hmm1, aa)
    _, tr = disorder(P, aa)
    tr = dict(tr)
    tr.update(vit=vit, map=[1 if pp[1][t] > pp[0][t] else 0 for t in range(len(aa))], post_bg=pp[0], post_prd=pp[1])
    return tr
