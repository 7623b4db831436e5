"""The Java host (plaac_b200/host/java/Plaac.java) cannot be compiled in this image (no JDK): check what can be checked
from its source -- its constant tables equal the C++ host's, its struct sizes/offsets equal the header's, and the
column documentation it ships equals the golden text."""
import ctypes as C
import os
import re

import plaac_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JAVA = os.path.join(ROOT, "plaac_b200", "host", "java")


def _arrays(text, pattern):
    out = {}
    for m in re.finditer(pattern, text, flags=re.S):
        out[m.group(1).lower()] = [float(x) for x in re.findall(r"-?\d+\.?\d*(?:[eE]-?\d+)?", m.group(2))]
    return out


def test_java_tables_equal_the_cpp_host():
    java = open(os.path.join(JAVA, "Plaac.java")).read()
    cpp = open(os.path.join(ROOT, "plaac_b200", "csrc", "host_params.cpp")).read()
    ja = _arrays(java, r"static final double\[\] (\w+) = \{(.*?)\};")
    ca = _arrays(cpp, r"const double k(\w+)\[PLAAC_NAA\] = \{(.*?)\};")
    pairs = {"hydro": "hydro", "papa_odds": "papaodds", "bg_scer": "bgscer", "prd_28": "prd28"}
    for j, c in pairs.items():
        assert len(ja[j]) == 22 and ja[j] == ca[c], j
    charge = re.search(r"double\[\] charge = \{(.*?)\};", java).group(1)
    assert [float(x) for x in charge.split(",")] == ca["charge"]
    assert 'AA = "XACDEFGHIKLMNPQRSTVWY*"' in java


def test_java_struct_sizes_match_the_header():
    java = open(os.path.join(JAVA, "Plaac.java")).read()
    assert "SUMMARY_BYTES = 160" in java and C.sizeof(plaac_b200.Summary) == 160
    naa, lut = 22, 4001
    assert "PARAMS_BYTES = 8 * 4 + 8L * (4 + 2 + 2 + 2 * NAA + 5 * NAA + 3 + 1 + 1 + LUT)" in java
    assert 8 * 4 + 8 * (4 + 2 + 2 + 2 * naa + 5 * naa + 3 + 1 + 1 + lut) == C.sizeof(plaac_b200.Params)
    # field order the constructor writes == field order of the ctypes mirror of plaac_params
    order = [n for n, _ in plaac_b200.Params._fields_]
    assert order == ["core_len", "ww1", "ww2", "ww3", "adjust_prolines", "mw_window", "reserved", "lt", "li", "lf", "le", "le0",
                     "llr", "papa_lod", "hydro2", "charge", "fi_cc", "big_neg", "ln2", "loglut"]
    puts = re.findall(r"o = put\(s, o, (.*?)\);", java)
    assert [p.split("//")[0].strip() for p in puts] == [
        "p.lt[0]", "p.lt[1]", "p.li", "p.lf", "p.le[0]", "p.le[1]", "p.le[0]", "p.llr", "p.papaLod", "p.hydro2", "p.charge",
        "new double[] {2.785, -1, -1.151}", "new double[] {-1000000.0, Math.log(2.0)}", "p.logLut"]
    # the 14 ints / 13 doubles are read at 4*k and 56 + 8*k
    assert "r + 4L * k" in java and "r + 56 + 8L * k" in java


def test_java_column_docs_and_headers():
    gold = open(os.path.join(ROOT, "tests", "golden", "column_docs.txt")).read()
    assert open(os.path.join(JAVA, "column_docs.txt")).read() == gold
    java = open(os.path.join(JAVA, "Plaac.java")).read()
    names = open(os.path.join(ROOT, "tests", "golden", "column_names.txt")).read().split()
    hdr = "".join(re.findall(r'"((?:[^"\\]|\\.)*)"', re.search(r"SUMMARY_HEADER = (.*?);", java, flags=re.S).group(1)))
    assert hdr.replace("\\t", "\t").split("\t") == names


def _c_prototypes():
    txt = open(os.path.join(ROOT, "include", "plaac_cuda.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(int64_t|int|void|const char \*|void \*)\s*(plaac_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        protos[name] = (ret, [] if args in ("", "void") else [a.strip() for a in args.split(",")])
    return protos


def _ffm_layout(ctype):
    """C parameter / return type -> the java.lang.foreign layout that binds it."""
    t = ctype.strip()
    if "*" in t:
        return "ADDRESS"
    base = t.split()[0] if t.split()[0] != "const" else t.split()[1]
    return {"int": "JAVA_INT", "int32_t": "JAVA_INT", "int64_t": "JAVA_LONG", "size_t": "JAVA_LONG", "double": "JAVA_DOUBLE",
            "void": None}[base]


def test_java_downcall_descriptors_match_the_header():
    """Every fn("plaac_x", FunctionDescriptor...) of the Java host against the prototype in include/plaac_cuda.h:
    arity and the layout of every argument and of the result.  The multi-GPU, packed, ranking and FASTA entry points
    must be among them (the Java host bound plaac_score and the allocator only)."""
    java = open(os.path.join(JAVA, "Plaac.java")).read()
    protos = _c_prototypes()
    bound = {}
    for m in re.finditer(r'fn\("(plaac_\w+)",\s*FunctionDescriptor\.(of|ofVoid)\((.*?)\)\);', java, flags=re.S):
        layouts = [x.strip() for x in " ".join(m.group(3).split()).split(",") if x.strip()]
        bound[m.group(1)] = (m.group(2), layouts)
    for need in ("plaac_score", "plaac_score_multi", "plaac_score_packed", "plaac_score_multi_packed", "plaac_pack_host",
                 "plaac_packed_words", "plaac_rank", "plaac_score_fasta", "plaac_host_alloc", "plaac_create"):
        assert need in bound, need
    for name, (kind, layouts) in bound.items():
        ret, args = protos[name]
        want_ret = _ffm_layout(ret)
        if kind == "ofVoid":
            assert want_ret is None, name
            got_args = layouts
        else:
            assert layouts[0] == want_ret, (name, layouts[0], want_ret)
            got_args = layouts[1:]
        assert got_args == [_ffm_layout(a) for a in args], (name, got_args, args)
