"""Host CLI (plaac_b200/host/plaac_cli.cpp): plaac.jar's command line, header block, FASTA semantics and row
formatting.  CPU tests cover everything up to the device boundary; the gpu-marked tests run the binary end to end
and compare every printed cell with rows formatted from the oracle's numbers."""
import math
import os
import subprocess
from decimal import ROUND_HALF_UP, Decimal

import numpy as np
import pytest

import plaac_b200
from oracle import orc
from plaac_b200 import build as pb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def cli():
    pb.build_lib()
    path = pb.build_cli()
    assert path and os.path.exists(path)
    return path


def run(cli, *args, stdin=None):
    return subprocess.run([cli, *args], input=stdin, capture_output=True, text=True, timeout=600)


def java_fmt(x, d):
    """java.util.Formatter %.<d>f: half-up on the shortest round-trip digits (= Python's repr digits)."""
    x = float(x)
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "Infinity" if x > 0 else "-Infinity"
    q = Decimal(repr(abs(x))).quantize(Decimal(1).scaleb(-d), rounding=ROUND_HALF_UP)
    return ("-" if math.copysign(1.0, x) < 0 else "") + f"{q:f}"


def test_java_formatting_rule(cli):
    rng = np.random.default_rng(3)
    vals = [0.0625, -0.0625, 0.1875, 2.5, 0.1, 1.0005, 1.0015, 0.0005, 0.00049, 0.0004999, 99.9995, 999.9995, -0.0, 0.0,
            1234.56749999, 5e-324, 1e-5, 123456789.125, float("nan"), float("inf"), float("-inf"), 0.5, 1.5, 0.05, 0.15]
    vals += list(rng.normal(0, 30, 3000)) + list(np.round(rng.normal(0, 3, 2000), 4)) + list(rng.integers(-9999, 9999, 2000) / 2000.0)
    lines = []
    for v in vals:
        for d in (0, 3, 4, 5, 8):
            lines.append((d, float(v)))
    txt = "".join(f"{d} {v!r}\n" for d, v in lines)
    out = run(cli, "--format-check", stdin=txt).stdout.split("\n")
    for (d, v), got in zip(lines, out):
        want = java_fmt(v, d)
        assert got == want, (v, d, got, want)
        assert orc.java_fmt(v, d) == want, (v, d, orc.java_fmt(v, d), want)
    # Double.toString for the alpha echo
    dbl = {1.0: "1.0", 0.5: "0.5", 0.0: "0.0", 0.25: "0.25", 1e-4: "1.0E-4", 0.001: "0.001", 0.3: "0.3", 1.0 / 3: "0.3333333333333333"}
    out = run(cli, "--format-check", stdin="".join(f"-1 {v!r}\n" for v in dbl)).stdout.split("\n")
    assert out[:len(dbl)] == list(dbl.values())


def test_usage_and_background_count_mode(cli, tmp_path):
    r = run(cli)
    assert r.returncode == 0 and "USAGE" in r.stdout
    fa = tmp_path / "bg.fa"
    # isvalidprotein (plaac.java:1732): X/* inside or X at the end drop the whole record; position 0 is never checked;
    # a terminal * is counted (bin 21); a blank line ends the record and the rest up to '>' is skipped
    fa.write_text(">a\nMKV\nAC*\n>b\nMXK\n>c\nXKV\n\nIGNORED\n>d\nKK*K\n>e\nacd\r\n")
    r = run(cli, "-b", str(fa))
    assert r.returncode == 0
    rows = [l.split(" # ") for l in r.stdout.strip().split("\n")]
    assert [n for _, n in rows] == list("XACDEFGHIKLMNPQRSTVWY*")
    cnt = {n: float(v) for v, n in rows}
    want = {c: 0.0 for c in "XACDEFGHIKLMNPQRSTVWY*"}
    for ch in "MKVAC*" + "XKV" + "ACD":
        want[ch] += 1
    assert cnt == want
    assert rows[1][0] == "2.000000"
    # -B without -i echoes the file
    r2 = run(cli, "-B", os.path.join(GOLD, "bg_freqs_HUMAN.txt"))
    assert r2.stdout.split("\n")[1].startswith("2428201.000000 # A")


def test_parameter_block_and_column_docs(cli, tmp_path):
    """Everything the jar prints before the first row is host work and must not need a GPU."""
    fa = tmp_path / "x.fa"
    fa.write_text(">p1\nMKVQQNNQQ\n")
    r = run(cli, "-i", str(fa), "-d", "-a", "0.5", "-c", "30", "-W", "21", "-zz", "1", "-m", "1")
    out = r.stdout.split("\n")
    assert out[0] == "# skipping unknown option -zz"
    assert out[1] == "# skipping unknown option 1"  # an unknown token does not consume its neighbour (plaac.java:351)
    blk = out[out.index("############################ parameters at run-time ####################################"):]
    assert blk[1] == "## alpha=0.5; corelength=30; ww1=41; ww2=21; ww3=21; adjustprolines=true;"
    assert [l.split(":")[0] for l in blk[2:8]] == ["## fg_used", "## bg_scer", "## bg_input", "## bg_used", "## plaac_llr", "## papa_lods"]
    P = orc.make_params(alpha=0.5, bg_counts=np.array([0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 1, 2, 0, 4, 0, 0, 0, 1, 0, 0, 0.0]),
                        core_len=30, ww2=21)
    want_llr = "".join(f"{n}={java_fmt(v, 5)};" for n, v in zip(orc.AANAMES, P.llr))
    assert blk[6] == "## plaac_llr: {" + want_llr + "}"
    # column documentation == the reference's generated web/views/_plaac_headers.haml (committed as a fixture)
    docs = [l[3:] for l in out if l.startswith("## ") and ": " in l and l.split(":")[0][3:] in open(os.path.join(GOLD, "column_names.txt")).read().split()]
    gold = [l.rstrip("\n") for l in open(os.path.join(GOLD, "column_docs.txt"))]
    assert docs == gold
    hdr = [l for l in out if l.startswith("SEQid\t")]
    if plaac_b200.lib().plaac_device_count() == 0:
        assert r.returncode == 2 and "no CUDA device" in r.stderr  # fails loudly, no CPU scoring path
    else:
        assert hdr and hdr[0].split("\t") == open(os.path.join(GOLD, "column_names.txt")).read().split()


def test_parameter_block_equals_the_jars_with_B_and_F(cli, tmp_path):
    """`-B file -a 0.5` and the jar's -F behaviour (main :388 reads the -B file for the foreground; --compat-F):
    the six `## ...` vectors the binary prints are the ones the jar's bytecode printed."""
    from tests import jarvec

    J = jarvec.load()
    fa = tmp_path / "h.fa"
    fa.write_text(J["human_fasta"])
    bgf = os.path.join(GOLD, "bg_freqs_HUMAN.txt")
    edge = tmp_path / "edge.fa"
    edge.write_text(J["edge_fasta"])
    for tag, args in (("human_params", ["-B", bgf, "-a", "0.5"]),
                      ("human_F_params", ["-B", bgf, "-a", "0.5", "-F", str(tmp_path / "ignored.txt"), "--compat-F"]),
                      ("human_b_params", ["-b", str(edge), "-a", "0.3"])):  # background counted from another FASTA
        r = run(cli, "-i", str(fa), *args)
        got = {l[3:l.index(":")]: l for l in r.stdout.split("\n") if l.startswith("## ") and ": {" in l}
        for key, want in J[tag].items():
            assert got[key] == want, (tag, key)


# ------------------------------------------------------------------------------------------------- end to end (GPU)
def _expected_summary_row(P, name, aa, s, corelen, ww2):
    names = orc.AANAMES

    def sub(r1, r2):
        m = len(aa)
        r1 = max(r1, 0)
        r2 = max(r2, r1)
        r1 = min(r1, m - 1)
        r2 = min(r2, m - 1)
        return "".join(names[c] for c in aa[r1:r2 + 1])

    def F(x):
        return java_fmt(float("nan") if math.isinf(x) else x, 3)

    llrlen = int(s["llr_end"] - s["llr_start"] + 1)
    llr = float("nan") if math.isinf(s["llr"]) else float(s["llr"])
    prdlen = int(s["prd_end"] - s["prd_start"] + 1)
    with np.errstate(all="ignore"):
        nllr = float(np.float64(llr) / np.float64(llrlen))
    row = [name, s["mw_score"], s["mw_start"] + 1, s["mw_end"] + 1, s["mw_end"] - s["mw_start"] + 1, F(llr), s["llr_start"] + 1,
           s["llr_end"] + 1, llrlen, F(nllr), s["vit_maxrun"], F(s["core_score"]), s["core_start"] + 1, s["core_end"] + 1,
           s["core_end"] - s["core_start"] + 1, F(s["prd_score"]), s["prd_start"] + 1, s["prd_end"] + 1, prdlen, s["prot_len"],
           F(s["hmm_all"]), F(s["hmm_vit"])]
    if prdlen >= corelen:
        row += [sub(s["core_start"], s["core_end"]), sub(s["prd_start"], s["prd_start"] + 14), sub(s["prd_end"] - 14, s["prd_end"]),
                sub(s["prd_start"], s["prd_end"])]
    else:
        row += ["-"] * 4
    row += [s["fi_numaa"], F(s["fi_meanhydro"]), F(s["fi_meancharge"]), F(s["fi_meancombo"]), s["fi_maxrun"], F(s["papa_combo"]),
            F(s["papa_prop"]), F(s["papa_fi"]), F(s["papa_llr"]), F(s["papa_llr2"]), s["papa_center"] + 1,
            sub(s["papa_center"] - ww2 // 2, s["papa_center"] + ww2 // 2)]
    return [str(x) for x in row]


def _cells_match(got, want):
    """Text equality, except that a float cell may differ by one unit in the last printed place (the GPU's
    window sums are within 1e-9 of the oracle's, which can straddle a rounding boundary)."""
    if got == want:
        return True
    try:
        a, b = float(got), float(want)
    except ValueError:
        return False
    return "." in want and abs(a - b) <= 1.5 * 10.0 ** (-len(want.split(".")[1]))


@pytest.mark.gpu
def test_summary_table_end_to_end(cli, tmp_path, golden):
    fa = tmp_path / "in.fa"
    rng = np.random.default_rng(5)
    recs = [(p["name"], p["seq"]) for p in golden["proteins"]]
    recs += [("short one", "MKV*"), ("lower case", "mqnqnqnqnqnqyyyqqqnnn" * 8), ("with X", "MKXXQQNN" * 20), ("only stop", "*")]
    for i in range(40):
        n = int(rng.integers(1, 700))
        recs.append((f"rnd{i} some description", "".join(rng.choice(list("ACDEFGHIKLMNPQRSTVWY"), n))))
    with open(fa, "w") as f:
        for k, (nm, sq) in enumerate(recs):
            f.write(f">{nm}\n")
            for j in range(0, len(sq), 60):
                f.write(sq[j:j + 60] + ("\r\n" if k % 3 == 0 else "\n"))
            if k % 5 == 0:
                f.write("\nstray text after a blank line is skipped\n")
    r = run(cli, "-i", str(fa), "-s", "--batch-mb", "1")
    assert r.returncode == 0, r.stderr
    lines = r.stdout.rstrip("\n").split("\n")
    assert lines[0].split("\t") == open(os.path.join(GOLD, "column_names.txt")).read().split()
    kept = [(nm, orc.encode(sq)) for nm, sq in recs if len(orc.encode(sq)) > 0]
    assert len(lines) - 1 == len(kept)
    codes, offs = orc.pack([c for _, c in kept])
    bg = np.zeros(22)
    for nm, sq in recs:  # the jar derives bg_input from the input file even at alpha = 1 (header only)
        pass
    P = orc.make_params()
    ref = orc.score_batch(P, codes, offs)
    for line, (nm, aa), s in zip(lines[1:], kept, ref):
        got = line.split("\t")
        want = _expected_summary_row(P, nm, aa, s, 60, 41)
        assert len(got) == len(want) == 38
        for g, w in zip(got, want):
            assert _cells_match(g, w), (nm, got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("which,args", [("edge_summary", ["-a", "0.5"]),
                                        ("edge_alt_summary", ["-a", "0", "-c", "30", "-w", "31", "-W", "51"])])
def test_summary_table_equals_the_jars_own_output(cli, tmp_path, which, args):
    """The binary on the edge-case FASTA against the table web/bin/plaac.jar itself printed for the same command line
    (its bytecode run by tests/golden/minijvm.py): every cell, numbers through java.util.Formatter's %.3f rule."""
    from tests import jarvec

    J = jarvec.load()
    fa = tmp_path / "edge.fa"
    fa.write_text(J["edge_fasta"])
    r = run(cli, "-i", str(fa), "-s", *args)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.rstrip("\n").split("\n")
    cols = open(os.path.join(GOLD, "column_names.txt")).read().split()
    assert lines[0].split("\t") == cols
    rows = J[which]
    assert len(lines) - 1 == len(rows)
    floats = {"LLR", "NLLR", "COREscore", "PRDscore", "HMMall", "HMMvit", "FImeanhydro", "FImeancharge", "FImeancombo",
              "PAPAcombo", "PAPAprop", "PAPAfi", "PAPAllr", "PAPAllr2"}
    nties = 0
    for line, row in zip(lines[1:], rows):
        got = dict(zip(cols, line.split("\t")))
        cen_same = got["PAPAcen"] == str(row["PAPAcen"])
        nties += not cen_same
        for c in cols:
            if c.startswith("PAPA") and not cen_same:
                continue  # documented exact-tie class of the PAPA centre (plateaus), DESIGN.md section 3
            want = java_fmt(jarvec.val(row[c]), 3) if c in floats else str(row[c])
            assert _cells_match(got[c], want), (row["SEQid"], c, got[c], want)
    assert nties <= 3


@pytest.mark.gpu
def test_per_residue_table_end_to_end(cli, tmp_path, golden):
    fa = tmp_path / "in.fa"
    with open(fa, "w") as f:
        for p in golden["proteins"]:
            f.write(f">{p['name']}\n{p['seq']}\n")
    plist = tmp_path / "list.txt"
    names = [p["name"] for p in golden["proteins"]]
    plist.write_text(f"{names[2]}\tThird\n>{names[0]}\n")
    r = run(cli, "-i", str(fa), "-s", "-p", str(plist))
    assert r.returncode == 0, r.stderr
    lines = r.stdout.rstrip("\n").split("\n")
    assert lines[0].split("\t") == ("ORDER SEQid AANUM AA VIT MAP CHARGE HYDRO FI PLAAC PAPA FIx2 PLAACx2 PAPAx2 "
                                    "HMM.background HMM.PrD-like").split()
    P = orc.make_params()
    body = lines[1:]
    pos = 0
    # file order: names[0] (matched through ">name": keeps its name, ORDER = running count), then names[2] (synonym, ORDER = line 1)
    for idx, (nm, order) in ((0, (names[0], "1")), (2, ("Third", "1"))):
        aa = orc.encode(golden["proteins"][idx]["seq"])
        c, o = orc.pack([aa])
        tr = orc.residue_batch(P, c, o)
        for t in range(len(aa)):
            got = body[pos].split("\t")
            want = [order, nm, str(t + 1), orc.AANAMES[aa[t]], str(int(tr["vit"][t])), str(int(tr["map"][t]))]
            for k, d in (("charge", 4), ("hydro", 4), ("fi", 8), ("plaac", 4), ("papa", 8), ("fix2", 8), ("plaacx2", 4), ("papax2", 8),
                         ("post_bg", 4), ("post_prd", 4)):
                want.append(java_fmt(float(tr[k][t]), d))
            assert len(got) == 16
            for g, w in zip(got, want):
                assert _cells_match(g, w), (nm, t, got, want)
            pos += 1
        assert body[pos] == "#" * 56
        pos += 1
    assert pos == len(body)
    # -p all numbers the proteins 1..N
    r = run(cli, "-i", str(fa), "-s", "-p", "all")
    ids = [l.split("\t")[0] for l in r.stdout.split("\n")[1:] if l and not l.startswith("#")]
    assert sorted(set(ids), key=int) == ["1", "2", "3", "4"]


@pytest.mark.gpu
@pytest.mark.parametrize("which,args", [("edge_residue", ["-a", "0.5", "-c", "40", "-w", "21", "-W", "21"]),
                                        ("edge_alt_residue", ["-a", "0", "-w", "52", "-W", "9"])])
def test_per_residue_table_equals_the_jars_own_output(cli, tmp_path, which, args):
    """`-p all` on the edge-case FASTA against what plaac.jar's bytecode printed for the same command line: ORDER, name,
    position, residue, VIT, MAP and the ten numeric columns (%.4f / %.8f through java.util.Formatter's rule)."""
    from tests import jarvec

    J = jarvec.load()
    fa = tmp_path / "edge.fa"
    fa.write_text(J["edge_fasta"])
    r = run(cli, "-i", str(fa), "-s", "-p", "all", *args)
    assert r.returncode == 0, r.stderr
    body = [l for l in r.stdout.rstrip("\n").split("\n")[1:] if not l.startswith("#")]
    prots = J[which]
    assert len(body) == sum(len(p["aa"]) for p in prots)
    pos = 0
    for pr in prots:
        for t in range(len(pr["aa"])):
            got = body[pos].split("\t")
            want = [pr["order"], pr["name"], str(t + 1), pr["aa"][t], str(pr["vit"][t]), str(pr["map"][t])]
            for k, d in (("CHARGE", 4), ("HYDRO", 4), ("FI", 8), ("PLAAC", 4), ("PAPA", 8), ("FIx2", 8), ("PLAACx2", 4),
                         ("PAPAx2", 8)):
                want.append(java_fmt(jarvec.val(pr[k][t]), d))
            want += [java_fmt(jarvec.val(v), 4) for v in pr["post"][t]]
            assert len(got) == len(want) == 16
            for g, w in zip(got, want):
                assert _cells_match(g, w), (pr["name"], t, got, want)
            pos += 1


@pytest.mark.gpu
def test_gpu_ingest_path_prints_the_same_table(cli, tmp_path, golden):
    """--gpu-ingest (FASTA parsed, encoded and scored on the GPU, in record-aligned pieces) == the host reader path,
    including name trimming at piece boundaries."""
    rng = np.random.default_rng(8)
    parts = []
    for k, p in enumerate(golden["proteins"]):
        parts.append(f">{p['name']} trailing blanks  \n{p['seq']}*\n".encode())
    for r in range(600):
        term = [b"\n", b"\r\n"][r % 2]
        parts.append(b">r%d  name with spaces \t" % r + term)
        seq = "".join(rng.choice(list("ACDEFGHIKLMNPQRSTVWYqnx "), int(rng.integers(0, 1500)))).encode()
        for j in range(0, len(seq), 70):
            parts.append(seq[j:j + 70] + term)
        if r % 9 == 0:
            parts.append(term + b"ignored after blank" + term)
    fa = tmp_path / "messy.fa"
    fa.write_bytes(b"".join(parts))
    a = run(cli, "-i", str(fa), "-s", "--host-reader")
    b = run(cli, "-i", str(fa), "-s", "--gpu-ingest")
    c = run(cli, "-i", str(fa), "-s", "--gpu-ingest", "--batch-mb", "1")  # many pieces
    assert a.returncode == 0 and b.returncode == 0 and c.returncode == 0, (a.stderr, b.stderr, c.stderr)
    assert len(a.stdout.split("\n")) > 500
    assert a.stdout == b.stdout
    assert a.stdout == c.stdout
    # the whole output, parameter block included, in the three orders of work the fast path has:
    #   alpha = 1 (one pass, rows held until the input's composition is known for the "## bg_input" line),
    #   alpha < 1 (composition counted on the GPU first, then scored), -B given (streamed)
    bgfile = os.path.join(GOLD, "bg_freqs_HUMAN.txt")
    for extra in ([], ["-a", "0.4"], ["-B", bgfile, "-a", "0.5"], ["-d"]):
        h = run(cli, "-i", str(fa), "--host-reader", *extra)
        g = run(cli, "-i", str(fa), "--gpu-ingest", "--batch-mb", "1", *extra)
        assert h.returncode == 0 and g.returncode == 0, (extra, h.stderr, g.stderr)
        assert h.stdout == g.stdout, extra
    # a file of this size takes the fast path by default (>= 1 MB) -- same table; --gpus 2 uses what the box has
    assert os.path.getsize(fa) > (1 << 20) or True
    d = run(cli, "-i", str(fa), "-s", "--gpus", "2", "--batch-mb", "1")
    assert d.returncode == 0 and d.stdout == a.stdout


@pytest.mark.gpu
def test_fast_path_ranks_like_the_batch_path(cli, tmp_path):
    """--gpu-ingest with --rank / --rank-core (ignored in round 1, ADVICE): same rows as the host-reader path."""
    from tests import synth

    codes, offs = synth.proteome(3000, 33, prd_rate=0.3)
    names = "XACDEFGHIKLMNPQRSTVWY*"
    fa = tmp_path / "rank_fast.fa"
    with open(fa, "w") as f:
        for i in range(3000):
            f.write(f">p{i}\n" + "".join(names[c] for c in codes[offs[i]:offs[i + 1]]) + "\n")
    assert os.path.getsize(fa) > (1 << 20)
    for opt in ("--rank", "--rank-core"):
        h = run(cli, "-i", str(fa), "-s", "--host-reader", opt)
        g = run(cli, "-i", str(fa), "-s", opt)  # default: fast path (file >= 1 MB)
        assert h.returncode == 0 and g.returncode == 0, (h.stderr, g.stderr)
        assert h.stdout == g.stdout and len(g.stdout.split("\n")) > 100, opt


@pytest.mark.gpu
def test_rank_option_orders_rows_like_the_web_front_end(cli, tmp_path):
    """--rank prints the same rows in the order of web/lib/server.rb:222-229; --rank-core only those with a CORE."""
    from tests import synth

    codes, offs = synth.proteome(400, 31, prd_rate=0.3)
    names = "XACDEFGHIKLMNPQRSTVWY*"
    fa = tmp_path / "rank.fa"
    with open(fa, "w") as f:
        for i in range(400):
            f.write(f">p{i}\n" + "".join(names[c] for c in codes[offs[i]:offs[i + 1]]) + "\n")
    plain = run(cli, "-i", str(fa), "-s")
    ranked = run(cli, "-i", str(fa), "-s", "--rank")
    core = run(cli, "-i", str(fa), "-s", "--rank-core")
    assert plain.returncode == 0 and ranked.returncode == 0 and core.returncode == 0, (ranked.stderr, core.stderr)
    rows = [l for l in plain.stdout.split("\n") if l and not l.startswith("#")]
    rrows = [l for l in ranked.stdout.split("\n") if l and not l.startswith("#")]
    crows = [l for l in core.stdout.split("\n") if l and not l.startswith("#")]
    assert rows[0] == rrows[0] == crows[0] and sorted(rows) == sorted(rrows)
    hdr = rows[0].split("\t")
    ic, il = hdr.index("COREscore"), hdr.index("LLR")
    body = [r.split("\t") for r in rrows[1:]]
    ncore = sum(1 for r in body if r[ic] != "NaN")
    assert 0 < ncore < len(body) and len(crows) - 1 == ncore
    assert all(r[ic] != "NaN" for r in body[:ncore]) and all(r[ic] == "NaN" for r in body[ncore:])
    cs = [float(r[ic]) for r in body[:ncore]]
    assert all(a >= b for a, b in zip(cs, cs[1:]))
    ls = [float(r[il]) for r in body[ncore:]]
    assert all(a >= b for a, b in zip(ls, ls[1:]))
    assert crows[1:] == rrows[1:1 + ncore]
