"""Golden per-residue vectors from the reference's own bytecode for the LONG proteins of jar_vectors.json.gz.

Runs web/bin/plaac.jar's `plaac.main -i long.fa -p all` under tests/golden/minijvm.py on the two proteins of `long_fasta`
(4 500 and 9 000 residues): plotsomefastas :610-647 -- Viterbi and MAP parse, the eight disorderreport tracks and both
posterior columns of every residue, at full precision.  These pin the per-residue long-sequence path (long_residue.cuh)
to the jar directly.

    python tests/golden/make_jar_vectors_long_residue.py   # needs /root/reference; writes jar_vectors_long_residue.json.gz
"""
import gzip
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_jar_vectors as M  # noqa: E402


def main():
    J = json.loads(gzip.open(os.path.join(HERE, "jar_vectors.json.gz")).read())
    txt = J["long_fasta"]
    assert txt == M.long_fasta()
    with tempfile.NamedTemporaryFile("w", suffix=".fa", delete=False) as f:
        f.write(txt)
        path = f.name
    ev, steps = M.run_main(["-i", path, "-p", "all"])
    os.unlink(path)
    prots = M.parse_residue(ev)
    print("long per-residue:", len(prots), "proteins,", [len(p["vit"]) for p in prots], "residues,", steps, "bytecodes")
    out = {"generator": "tests/golden/make_jar_vectors_long_residue.py (reference bytecode web/bin/plaac.jar under minijvm.py)",
           "fasta": "long_fasta of jar_vectors.json.gz", "args": ["-p", "all"], "long_residue": prots}
    p = os.path.join(HERE, "jar_vectors_long_residue.json.gz")
    with gzip.GzipFile(p, "wb", mtime=0) as f:
        f.write(json.dumps(out, indent=0, separators=(",", ":")).encode())
    print("wrote", p, os.path.getsize(p), "bytes")


if __name__ == "__main__":
    main()
