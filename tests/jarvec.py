"""Access to tests/golden/jar_vectors.json.gz: values the reference's own bytecode (web/bin/plaac.jar, interpreted by
tests/golden/minijvm.py) handed to System.out.format(), at full precision (hex floats)."""
import gzip
import json
import math
import os

import numpy as np

from oracle import orc

HERE = os.path.dirname(os.path.abspath(__file__))
_cache = None


def load():
    global _cache
    if _cache is None:
        with gzip.open(os.path.join(HERE, "golden", "jar_vectors.json.gz"), "rb") as f:
            _cache = json.loads(f.read().decode())
    return _cache


_cache_long = None


def load_long_residue():
    """tests/golden/jar_vectors_long_residue.json.gz: `-p all` on the 4 500- and 9 000-residue proteins of long_fasta."""
    global _cache_long
    if _cache_long is None:
        with gzip.open(os.path.join(HERE, "golden", "jar_vectors_long_residue.json.gz"), "rb") as f:
            _cache_long = json.loads(f.read().decode())
    return _cache_long


def val(v):
    if isinstance(v, str):
        if v == "NaN":
            return float("nan")
        if v == "Infinity":
            return float("inf")
        if v == "-Infinity":
            return float("-inf")
        return float.fromhex(v)
    return v


def jar_reader(text: str):
    """(name, unstripped sequence) per record with the jar's reader semantics (fastareader :4302-4375)."""
    import re

    lines = re.split(r"\r\n|\n|\r", text)
    if lines and lines[-1] == "":
        lines.pop()
    recs, i, ondeck, name = [], 0, False, None
    while True:
        if not ondeck:
            while i < len(lines) and not lines[i].startswith(">"):
                i += 1
            if i >= len(lines):
                break
            name = lines[i].strip("".join(map(chr, range(33))))[1:]
            i += 1
        seq, ondeck, nxt = "", False, None
        while i < len(lines):
            ln = lines[i]
            i += 1
            if ln == "":
                break
            if ln.startswith(">"):
                ondeck, nxt = True, ln[1:]
                break
            seq += ln
        recs.append((name, seq))
        if ondeck:
            name = nxt
    return recs


def background_counts(recs):
    """computeaafreq(inputfile) :1655-1739, as main :382-384 does when neither -b nor -B is given."""
    bg = np.zeros(22)
    for _, s in recs:
        if not s:
            continue
        aa = orc.encode(s, strip_stop=False)
        m = len(aa)
        if aa[m - 1] != 0 and not any(aa[i] in (0, 21) for i in range(1, m - 1)):
            bg += np.bincount(aa, minlength=22)
    return bg


def scenario(which):
    """-> (recs with non-empty stripped sequence [(name, codes)], params kwargs, jar rows)"""
    J = load()
    if which == "prions_summary":
        text, kw, rows = J["prions_fasta"], {}, J["prions_summary"]
    elif which == "edge_summary":
        text, kw, rows = J["edge_fasta"], dict(alpha=0.5), J["edge_summary"]
    elif which == "long_summary":
        text, kw, rows = J["long_fasta"], {}, J["long_summary"]
    elif which == "long_residue":
        text, kw, rows = J["long_fasta"], {}, load_long_residue()["long_residue"]
    elif which == "prions_residue":
        text, kw, rows = J["prions_fasta"], {}, J["prions_residue"]
    elif which == "edge_residue":
        text, kw, rows = J["edge_fasta"], dict(alpha=0.5, core_len=40, ww1=21, ww2=21), J["edge_residue"]
    elif which == "edge_alt_summary":
        text, kw, rows = J["edge_fasta"], dict(alpha=0.0, core_len=30, ww1=31, ww2=51), J["edge_alt_summary"]
    elif which == "edge_alt_residue":
        text, kw, rows = J["edge_fasta"], dict(alpha=0.0, ww1=52, ww2=9), J["edge_alt_residue"]
    elif which == "human_b_summary":
        recs = jar_reader(J["human_fasta"])
        enc = [(n, orc.encode(s)) for n, s in recs if s]
        assert len(enc) == len(J[which])
        return enc, dict(alpha=0.3, bg_counts=background_counts(jar_reader(J["edge_fasta"]))), J[which]
    elif which in ("human_summary", "human_F_summary"):
        # -B file: read_aa_params (plaac.java:1923-1941) takes the first number of each of the 22 lines
        bgf = np.array([float(ln.split()[0]) for ln in open(os.path.join(HERE, "golden", "bg_freqs_HUMAN.txt"))][:22])
        recs = jar_reader(J["human_fasta"])
        enc = [(n, orc.encode(s)) for n, s in recs if s]
        assert len(enc) == len(J[which])
        kw = dict(alpha=0.5, bg_counts=bgf)
        if which == "human_F_summary":
            kw["fg_freq"] = bgf  # -F: main :388 reads the -B file for the foreground as well
        return enc, kw, J[which]
    else:
        raise KeyError(which)
    recs = jar_reader(text)
    kw = dict(kw, bg_counts=background_counts(recs))
    enc = [(n, orc.encode(s)) for n, s in recs if s]
    enc = [(n, c) for n, c in enc if len(c) > 0]
    assert len(enc) == len(rows)
    return enc, kw, rows


def expected_cells(r):
    """The numeric cells of one summary row as plaac.java:899-945 computes them from a record."""
    llrlen = int(r["llr_end"] - r["llr_start"] + 1)
    llr = float("nan") if math.isinf(r["llr"]) else float(r["llr"])
    with np.errstate(all="ignore"):
        nllr = float(np.float64(llr) / np.float64(llrlen))
    return {
        "MW": int(r["mw_score"]), "MWstart": int(r["mw_start"]) + 1, "MWend": int(r["mw_end"]) + 1,
        "MWlen": int(r["mw_end"] - r["mw_start"] + 1), "LLR": llr, "LLRstart": int(r["llr_start"]) + 1,
        "LLRend": int(r["llr_end"]) + 1, "LLRlen": llrlen, "NLLR": nllr, "VITmaxrun": int(r["vit_maxrun"]),
        "COREscore": float(r["core_score"]), "COREstart": int(r["core_start"]) + 1, "COREend": int(r["core_end"]) + 1,
        "CORElen": int(r["core_end"] - r["core_start"] + 1), "PRDscore": float(r["prd_score"]),
        "PRDstart": int(r["prd_start"]) + 1, "PRDend": int(r["prd_end"]) + 1, "PRDlen": int(r["prd_end"] - r["prd_start"] + 1),
        "PROTlen": int(r["prot_len"]), "HMMall": float(r["hmm_all"]), "HMMvit": float(r["hmm_vit"]),
        "FInumaa": int(r["fi_numaa"]), "FImeanhydro": float(r["fi_meanhydro"]), "FImeancharge": float(r["fi_meancharge"]),
        "FImeancombo": float(r["fi_meancombo"]), "FImaxrun": int(r["fi_maxrun"]),
        "PAPAcombo": float("nan") if math.isinf(r["papa_combo"]) else float(r["papa_combo"]), "PAPAprop": float(r["papa_prop"]),
        "PAPAfi": float(r["papa_fi"]), "PAPAllr": float(r["papa_llr"]), "PAPAllr2": float(r["papa_llr2"]),
        "PAPAcen": int(r["papa_center"]) + 1,
    }


RES_MAP = {"CHARGE": "charge", "HYDRO": "hydro", "FI": "fi", "PLAAC": "plaac", "PAPA": "papa", "FIx2": "fix2",
           "PLAACx2": "plaacx2", "PAPAx2": "papax2"}
