"""Regenerates tests/golden/classic_prions.json.

Run in the build container only (needs /root/reference for the INPUT
sequences; the GPU box has no /root/reference, which is why the result is
committed).  The reference ships no expected outputs and cannot be executed
(no JVM), so the "expected" block is produced by the CPU oracle
(oracle/plaac_oracle.c) and pinned against SURVEY.md Appendix B's
independently derived known answers (tests/test_oracle.py::test_appendix_b).
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import orc  # noqa: E402

SRC = "/root/reference/cli/example/four_classic_prions.fasta"


def read_fasta(path):
    names, seqs = [], []
    cur = None
    for line in open(path):
        line = line.rstrip("\r\n")
        if line.startswith(">"):
            names.append(line[1:])
            seqs.append([])
            cur = seqs[-1]
        elif not line:
            cur = None
        elif cur is not None:
            cur.append(line)
    return names, ["".join(s) for s in seqs]


def main():
    names, seqs = read_fasta(SRC)
    P = orc.make_params()
    codes, offs = orc.pack([orc.encode(s) for s in seqs])
    out = orc.score_batch(P, codes, offs)
    res = orc.residue_batch(P, codes, offs)
    doc = {"source": "cli/example/four_classic_prions.fasta (input sequences only)",
           "params": {"alpha": 1.0, "core_len": 60, "ww1": 41, "ww2": 41, "ww3": 41, "adjust_prolines": True},
           "llr": [float(x) for x in P.llr],
           "proteins": []}
    for i, (nm, sq) in enumerate(zip(names, seqs)):
        rec = {k: (float(out[i][k]) if out.dtype[k].kind == "f" else int(out[i][k])) for k in out.dtype.names}
        lo, hi = int(offs[i]), int(offs[i + 1])
        # a sparse sample of the per-residue tracks keeps the fixture small
        idx = sorted(set([0, 1, 19, 20, 21, 40, 41, 59, 60, (hi - lo) // 2, hi - lo - 42, hi - lo - 21, hi - lo - 20,
                          hi - lo - 1]))
        tracks = {k: [(None if v[lo + j] != v[lo + j] else float(v[lo + j])) for j in idx] for k, v in res.items()}
        doc["proteins"].append({"name": nm, "seq": sq, "summary": rec, "residue_idx": idx, "residue": tracks,
                                "vit_runs": runs(res["vit"][lo:hi]), "map_runs": runs(res["map"][lo:hi])})
    with open(os.path.join(HERE, "classic_prions.json"), "w") as f:
        json.dump(doc, f, indent=1)


def runs(bits):
    out, start = [], None
    for i, b in enumerate(list(bits) + [0]):
        if b and start is None:
            start = i
        if not b and start is not None:
            out.append([start, i - 1])
            start = None
    return out


if __name__ == "__main__":
    main()
