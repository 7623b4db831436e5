"""Golden vectors from the reference's own bytecode.

Runs web/bin/plaac.jar's `plaac.main` under tests/golden/minijvm.py (this image has no JVM) on
  * cli/example/four_classic_prions.fasta of the reference (summary table and `-p all` per-residue table), and
  * a synthetic FASTA of edge cases (written by this script, embedded in the fixture),
  * two long proteins (chunked long-sequence path) and eight proteins of human composition scored with
    `-B bg_freqs_HUMAN.txt -a 0.5` (config 3's background blend),
and stores every value the jar passed to System.out.format()/print() at full precision.

    python tests/golden/make_jar_vectors.py        # needs /root/reference; writes tests/golden/jar_vectors.json.gz

The natives behind the bytecode (Math.log/exp...) are this box's libm, i.e. the same library the oracle links, so
the oracle is expected to reproduce these numbers to the last bit wherever it follows the jar's operation order.
"""
import json
import math
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import minijvm  # noqa: E402

JAR = "/root/reference/web/bin/plaac.jar"
PRIONS = "/root/reference/cli/example/four_classic_prions.fasta"

SUMMARY_INT = ["mw_score", "mw_start", "mw_end", "mw_len"]
COLS1 = ["SEQid", "MW", "MWstart", "MWend", "MWlen", "LLR", "LLRstart", "LLRend", "LLRlen", "NLLR", "VITmaxrun", "COREscore",
         "COREstart", "COREend", "CORElen", "PRDscore", "PRDstart", "PRDend", "PRDlen", "PROTlen", "HMMall", "HMMvit"]
COLS2 = ["FInumaa", "FImeanhydro", "FImeancharge", "FImeancombo", "FImaxrun", "PAPAcombo", "PAPAprop", "PAPAfi", "PAPAllr",
         "PAPAllr2", "PAPAcen", "PAPAaa"]
RES_COLS = ["CHARGE", "HYDRO", "FI", "PLAAC", "PAPA", "FIx2", "PLAACx2", "PAPAx2"]


def jnum(v):
    if isinstance(v, float):
        if v != v:
            return "NaN"
        if math.isinf(v):
            return "Infinity" if v > 0 else "-Infinity"
        return float.hex(v)  # exact
    return v


def run_main(args):
    vm = minijvm.VM(JAR)
    vm.cls("plaac")
    main = vm.classes["plaac"].methods[("main", "([Ljava/lang/String;)V")]
    vm.invoke(main, [minijvm.JArray(list(args), "A")])
    return vm.out.events, vm.steps


def parse_summary(events):
    rows, cur, strings = [], None, []
    params = {}
    for e in events:
        if e[0] == "print" and e[1].startswith("## ") and ": {" in e[1]:
            k = e[1][3:e[1].index(":")]
            params[k] = e[1].strip()
        if e[0] == "format" and e[1].startswith("%s\t%d"):
            cur = dict(zip(COLS1, [jnum(v) for v in e[2]]))
            strings = []
        elif e[0] == "print" and cur is not None and e[1] not in ("\t", "\n"):
            strings.append(e[1])
        elif e[0] == "format" and e[1].startswith("\t%d\t%.3f") and cur is not None:
            vals = [v.s if hasattr(v, "s") else v for v in e[2]]
            cur.update(dict(zip(COLS2, [jnum(v) for v in vals])))
            cur["COREaa"], cur["STARTaa"], cur["ENDaa"], cur["PRDaa"] = strings[:4]
            rows.append(cur)
            cur = None
    return rows, params


def parse_residue(events):
    prots, cur = [], None
    for e in events:
        if e[0] == "print" and e[1].count("\t") == 6 and e[1].endswith("\t"):
            oid, nm, aanum, aa, vit, mp = e[1].rstrip("\t").split("\t")
            if cur is None or cur["name"] != nm or int(aanum) == 1:
                cur = {"order": oid, "name": nm, "aa": "", "vit": [], "map": [], "post": [], **{c: [] for c in RES_COLS}}
                prots.append(cur)
            cur["aa"] += aa
            cur["vit"].append(int(vit))
            cur["map"].append(int(mp))
            cur["post"].append([])
        elif e[0] == "format" and e[1].startswith("%.4f\t%.4f\t%.8f"):
            for c, v in zip(RES_COLS, e[2]):
                cur[c].append(jnum(v))
        elif e[0] == "format" and e[1] == "\t%.4f":
            cur["post"][-1].append(jnum(e[2][0]))
    return prots


def edge_fasta():
    rng = np.random.default_rng(20261017)
    aa = "ACDEFGHIKLMNPQRSTVWY"
    prd = np.array([0.04865, 0.00219, 0.01638, 0.00783, 0.02537, 0.07603, 0.0181, 0.02018, 0.01641, 0.02639, 0.02975,
                    0.25885, 0.05126, 0.15178, 0.025, 0.10988, 0.03841, 0.01972, 0.00157, 0.05624])
    prd /= prd.sum()
    recs = []

    def rnd(n, p=None):
        return "".join(rng.choice(list(aa), n, p=p))

    for n in (1, 2, 5, 20, 21, 40, 41, 42, 59, 60, 61, 79, 80, 81, 83, 100, 130):
        recs.append((f"len{n}", rnd(n)))
        recs.append((f"prd{n}", rnd(n, prd)))
    recs.append(("polyQ", "Q" * 150))
    recs.append(("polyP", "P" * 90))
    recs.append(("PxP", "PAPPN" * 30))
    recs.append(("QN_repeat", "QN" * 100))
    recs.append(("allX", "X" * 70))
    recs.append(("lower_and_junk", (rnd(50) + " 12" + rnd(60, prd)).lower()))
    recs.append(("stop_inside", rnd(40) + "*" + rnd(80, prd) + "*"))
    recs.append(("two_prds", rnd(120) + rnd(90, prd) + rnd(150) + rnd(140, prd) + rnd(60)))
    recs.append(("prd_whole", rnd(260, prd)))
    recs.append(("late_prd", rnd(700) + rnd(120, prd) + rnd(30)))
    lines = []
    for k, (nm, sq) in enumerate(recs):
        lines.append(">" + nm + (" description " if k % 3 == 0 else ""))
        for j in range(0, len(sq), 60):
            lines.append(sq[j:j + 60])
        if k % 7 == 3:
            lines += ["", "text after a blank line is skipped"]
    return "\n".join(lines) + "\n"


def long_fasta():
    """Two proteins above the threshold of the chunked long-sequence path (4096): background composition with
    Q/N-rich segments early, in the middle and reaching the last residue."""
    rng = np.random.default_rng(20261018)
    aa = "ACDEFGHIKLMNPQRSTVWY"
    bg = np.array([0.0550, 0.0126, 0.0586, 0.0655, 0.0441, 0.0498, 0.0217, 0.0655, 0.0735, 0.0950, 0.0207, 0.0615, 0.0438,
                   0.0396, 0.0444, 0.0899, 0.0592, 0.0556, 0.0104, 0.0337])
    bg /= bg.sum()
    prd = np.array([0.04865, 0.00219, 0.01638, 0.00783, 0.02537, 0.07603, 0.0181, 0.02018, 0.01641, 0.02639, 0.02975,
                    0.25885, 0.05126, 0.15178, 0.025, 0.10988, 0.03841, 0.01972, 0.00157, 0.05624])
    prd /= prd.sum()

    def rnd(n, p):
        return "".join(rng.choice(list(aa), n, p=p))

    recs = [("long4500", rnd(300, bg) + rnd(140, prd) + rnd(2000, bg) + rnd(90, prd) + rnd(1970, bg)),
            ("long9000", rnd(4000, bg) + rnd(200, prd) + rnd(4600, bg) + rnd(200, prd))]
    lines = []
    for nm, sq in recs:
        lines.append(">" + nm)
        for j in range(0, len(sq), 60):
            lines.append(sq[j:j + 60])
    return "\n".join(lines) + "\n"


def human_fasta():
    """Config 3's parameter path: proteins of human background composition (half of them with a Q/N-rich segment),
    scored with `-B bg_freqs_HUMAN.txt -a 0.5` (read_aa_params + the alpha blend of main :449-500)."""
    rng = np.random.default_rng(20261019)
    aa = "ACDEFGHIKLMNPQRSTVWY"
    bg = np.array([float(ln.split()[0]) for ln in open(os.path.join(HERE, "bg_freqs_HUMAN.txt"))][1:21])
    bg /= bg.sum()
    prd = np.array([0.04865, 0.00219, 0.01638, 0.00783, 0.02537, 0.07603, 0.0181, 0.02018, 0.01641, 0.02639, 0.02975,
                    0.25885, 0.05126, 0.15178, 0.025, 0.10988, 0.03841, 0.01972, 0.00157, 0.05624])
    prd /= prd.sum()

    def rnd(n, p):
        return "".join(rng.choice(list(aa), n, p=p))

    lines = []
    for k, n in enumerate((90, 150, 233, 310, 415, 520, 640, 800)):
        sq = rnd(n, bg)
        if k % 2:
            st, m = n // 3, min(100, n - n // 3)
            sq = sq[:st] + rnd(m, prd) + sq[st + m:]
        lines.append(f">hs{k}_{n}")
        lines += [sq[j:j + 60] for j in range(0, len(sq), 60)]
    return "\n".join(lines) + "\n"


def main():
    out = {"generator": "tests/golden/make_jar_vectors.py (reference bytecode web/bin/plaac.jar under minijvm.py)",
           "jar_manifest": "Created-By: 1.7.0_55"}
    ev, steps = run_main(["-i", PRIONS])
    rows, params = parse_summary(ev)
    out["prions_fasta"] = open(PRIONS).read()
    out["prions_summary"] = rows
    out["prions_params"] = params
    print("prions summary:", len(rows), "rows,", steps, "bytecodes")
    ev, steps = run_main(["-i", PRIONS, "-p", "all"])
    out["prions_residue"] = parse_residue(ev)
    print("prions per-residue:", len(out["prions_residue"]), "proteins,", steps, "bytecodes")
    txt = edge_fasta()
    with tempfile.NamedTemporaryFile("w", suffix=".fa", delete=False) as f:
        f.write(txt)
        path = f.name
    ev, steps = run_main(["-i", path, "-a", "0.5"])
    rows, params = parse_summary(ev)
    out["edge_fasta"] = txt
    out["edge_args"] = ["-a", "0.5"]
    out["edge_summary"] = rows
    out["edge_params"] = params
    print("edge summary:", len(rows), "rows,", steps, "bytecodes")
    ev, steps = run_main(["-i", path, "-a", "0.5", "-p", "all", "-c", "40", "-w", "21", "-W", "21"])
    out["edge_residue_args"] = ["-a", "0.5", "-c", "40", "-w", "21", "-W", "21"]
    out["edge_residue"] = [p for p in parse_residue(ev)]
    print("edge per-residue:", len(out["edge_residue"]), "proteins,", steps, "bytecodes")
    # -w and -W with different half-widths (FoldIndex window 31, PAPA/LLR windows 51): the windows of disorderreport are
    # independent in the jar (:4875-4903)
    ev, steps = run_main(["-i", path, "-a", "0", "-c", "30", "-w", "31", "-W", "51"])
    rows, params = parse_summary(ev)
    out["edge_alt_args"] = ["-a", "0", "-c", "30", "-w", "31", "-W", "51"]
    out["edge_alt_summary"] = rows
    out["edge_alt_params"] = params
    print("edge summary, -w 31 -W 51:", len(rows), "rows,", steps, "bytecodes")
    ev, steps = run_main(["-i", path, "-a", "0", "-p", "all", "-w", "52", "-W", "9"])
    out["edge_alt_residue_args"] = ["-a", "0", "-w", "52", "-W", "9"]
    out["edge_alt_residue"] = [p for p in parse_residue(ev)]
    print("edge per-residue, -w 52 -W 9:", len(out["edge_alt_residue"]), "proteins,", steps, "bytecodes")
    os.unlink(path)
    txt = long_fasta()
    with tempfile.NamedTemporaryFile("w", suffix=".fa", delete=False) as f:
        f.write(txt)
        path = f.name
    ev, steps = run_main(["-i", path])
    rows, params = parse_summary(ev)
    out["long_fasta"] = txt
    out["long_summary"] = rows
    out["long_params"] = params
    print("long summary:", len(rows), "rows,", steps, "bytecodes")
    os.unlink(path)
    txt = human_fasta()
    with tempfile.NamedTemporaryFile("w", suffix=".fa", delete=False) as f:
        f.write(txt)
        path = f.name
    bgfile = "/root/reference/web/bg_freqs/bg_freqs_HUMAN.txt"
    assert open(bgfile).read() == open(os.path.join(HERE, "bg_freqs_HUMAN.txt")).read()  # the committed copy is the reference's
    ev, steps = run_main(["-i", path, "-B", bgfile, "-a", "0.5"])
    rows, params = parse_summary(ev)
    out["human_fasta"] = txt
    out["human_args"] = ["-B", "bg_freqs_HUMAN.txt", "-a", "0.5"]
    out["human_summary"] = rows
    out["human_params"] = params
    print("human (-B, alpha 0.5) summary:", len(rows), "rows,", steps, "bytecodes")
    # -F: main :388 reads the -B file name for the foreground too (a bug of the jar); the host keeps it behind --compat-F
    ev, steps = run_main(["-i", path, "-B", bgfile, "-F", "/nonexistent/ignored_by_the_jar.txt", "-a", "0.5"])
    rows, params = parse_summary(ev)
    out["human_F_args"] = ["-B", "bg_freqs_HUMAN.txt", "-F", "ignored_by_the_jar.txt", "-a", "0.5"]
    out["human_F_summary"] = rows
    out["human_F_params"] = params
    print("human with -F (the jar reads the -B file):", len(rows), "rows,", steps, "bytecodes")
    # -b: background counted from another FASTA (computeaafreq + isvalidprotein :1655-1739 on the edge-case file)
    with tempfile.NamedTemporaryFile("w", suffix=".fa", delete=False) as f:
        f.write(edge_fasta())
        bpath = f.name
    ev, steps = run_main(["-i", path, "-b", bpath, "-a", "0.3"])
    rows, params = parse_summary(ev)
    out["human_b_args"] = ["-b", "<edge_fasta>", "-a", "0.3"]
    out["human_b_summary"] = rows
    out["human_b_params"] = params
    print("human with -b edge.fa -a 0.3:", len(rows), "rows,", steps, "bytecodes")
    os.unlink(bpath)
    os.unlink(path)
    import gzip

    path = os.path.join(HERE, "jar_vectors.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(out, indent=0, separators=(",", ":")).encode())
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
