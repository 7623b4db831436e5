"""ctypes loader for the CPU oracle (oracle/plaac_oracle.c).

TEST INFRASTRUCTURE ONLY -- see oracle/plaac_oracle.h.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
never by the product package plaac_b200.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liborc.so")

NAA = 22
LUTLEN = 4000
AANAMES = "XACDEFGHIKLMNPQRSTVWY*"  # plaac.java:26

SUMMARY_DTYPE = np.dtype(
    [
        ("mw_score", "<i4"), ("mw_start", "<i4"), ("mw_end", "<i4"),
        ("llr_start", "<i4"), ("llr_end", "<i4"), ("vit_maxrun", "<i4"),
        ("core_start", "<i4"), ("core_end", "<i4"), ("prd_start", "<i4"), ("prd_end", "<i4"),
        ("prot_len", "<i4"), ("fi_numaa", "<i4"), ("fi_maxrun", "<i4"), ("papa_center", "<i4"),
        ("llr", "<f8"), ("core_score", "<f8"), ("prd_score", "<f8"), ("hmm_all", "<f8"), ("hmm_vit", "<f8"),
        ("fi_meanhydro", "<f8"), ("fi_meancharge", "<f8"), ("fi_meancombo", "<f8"),
        ("papa_combo", "<f8"), ("papa_prop", "<f8"), ("papa_fi", "<f8"), ("papa_llr", "<f8"), ("papa_llr2", "<f8"),
    ],
    align=False,
)
assert SUMMARY_DTYPE.itemsize == 160
INT_FIELDS = [n for n in SUMMARY_DTYPE.names if SUMMARY_DTYPE[n].kind == "i"]
DBL_FIELDS = [n for n in SUMMARY_DTYPE.names if SUMMARY_DTYPE[n].kind == "f"]

RESIDUE_U8 = ("vit", "map")
RESIDUE_F64 = ("charge", "hydro", "fi", "plaac", "papa", "fix2", "plaacx2", "papax2", "post_bg", "post_prd")


class Hmm(C.Structure):
    _fields_ = [
        ("lt", (C.c_double * 2) * 2),
        ("le", (C.c_double * NAA) * 2),
        ("li", C.c_double * 2),
        ("lf", C.c_double * 2),
    ]


class Params(C.Structure):
    _fields_ = [
        ("core_len", C.c_int32), ("ww1", C.c_int32), ("ww2", C.c_int32), ("ww3", C.c_int32),
        ("adjust_prolines", C.c_int32),
        ("alpha", C.c_double),
        ("fg", C.c_double * NAA), ("bg", C.c_double * NAA), ("bgscer", C.c_double * NAA),
        ("bgthis", C.c_double * NAA), ("llr", C.c_double * NAA),
        ("hmm1", Hmm), ("hmm0", Hmm),
        ("papa_lod", C.c_double * NAA), ("hydro2", C.c_double * NAA), ("charge", C.c_double * NAA),
        ("fi_cc", C.c_double * 3),
        ("loglut", C.c_double * (LUTLEN + 1)),
        ("ln2", C.c_double),
    ]


class ResidueOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in RESIDUE_U8 + RESIDUE_F64]


_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/liborc.so with the committed Makefile."""
    src = os.path.join(_HERE, "plaac_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "liborc.so"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_aatoint.restype = C.c_int
        L.orc_aatoint.argtypes = [C.c_int]
        L.orc_params_init.restype = None
        L.orc_params_init.argtypes = [C.POINTER(Params), C.c_double, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_logeapeb.restype = C.c_double
        L.orc_logeapeb.argtypes = [C.POINTER(Params), C.c_double, C.c_double]
        L.orc_hss2.restype = None
        L.orc_hss2.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_score_batch.restype = None
        L.orc_score_batch.argtypes = [C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int]
        L.orc_residue_batch.restype = None
        L.orc_residue_batch.argtypes = [C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_int64,
                                        C.POINTER(ResidueOut), C.c_int]
        L.orc_java_fmt.restype = C.c_int
        L.orc_java_fmt.argtypes = [C.c_char_p, C.c_int, C.c_double, C.c_int]
        L.orc_fi_min_margin.restype = C.c_double
        L.orc_fi_min_margin.argtypes = [C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_int64)]
        L.orc_max_threads.restype = C.c_int
        L.orc_slidingaverage.restype = None
        L.orc_slidingaverage.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def make_params(alpha=1.0, bg_counts=None, fg_freq=None, core_len=60, ww1=41, ww2=41, ww3=None,
                adjust_prolines=True) -> Params:
    """plaac.java main :302-530 parameter chain (ww3 = ww2 as at :355 unless given)."""
    P = Params()
    if ww3 is None:
        ww3 = ww2
    bgp = fgp = None
    if bg_counts is not None:
        bg = np.ascontiguousarray(bg_counts, dtype=np.float64)
        assert bg.shape == (NAA,)
        bgp = bg.ctypes.data
    if fg_freq is not None:
        fg = np.ascontiguousarray(fg_freq, dtype=np.float64)
        assert fg.shape == (NAA,)
        fgp = fg.ctypes.data
    lib().orc_params_init(C.byref(P), float(alpha), bgp, fgp, core_len, ww1, ww2, ww3, int(bool(adjust_prolines)))
    return P


_ENC = np.zeros(256, dtype=np.uint8)
for _i, _ch in enumerate(AANAMES):
    if _ch == "X":
        continue
    _ENC[ord(_ch)] = _i
    _ENC[ord(_ch.lower())] = _i


def encode(seq: str | bytes, strip_stop: bool = True) -> np.ndarray:
    """string2aa (plaac.java:1764) after the terminal-'*' strip of :758."""
    if isinstance(seq, str):
        seq = seq.encode("latin-1")
    if strip_stop and len(seq) and seq[-1:] == b"*":
        seq = seq[:-1]
    return _ENC[np.frombuffer(seq, dtype=np.uint8)]


def pack(seqs) -> tuple[np.ndarray, np.ndarray]:
    """List of code arrays -> (codes u8, offsets i64[n+1])."""
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    offsets = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    codes = np.concatenate(seqs).astype(np.uint8) if len(seqs) and offsets[-1] > 0 else np.zeros(0, np.uint8)
    return np.ascontiguousarray(codes), offsets


def score_batch(P: Params, codes: np.ndarray, offsets: np.ndarray, full_jar_work=False, nthreads=1) -> np.ndarray:
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    nprot = len(offsets) - 1
    out = np.zeros(nprot, dtype=SUMMARY_DTYPE)
    lib().orc_score_batch(C.byref(P), codes.ctypes.data, offsets.ctypes.data, nprot, out.ctypes.data,
                          int(full_jar_work), int(nthreads))
    return out


def fi_min_margin(P: Params, codes: np.ndarray, offsets: np.ndarray, nthreads=1):
    """min |fi| * taps over every FoldIndex value the run scan tests (test instrumentation) -> (margin, protein index)."""
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    at = C.c_int64(-1)
    m = lib().orc_fi_min_margin(C.byref(P), codes.ctypes.data, offsets.ctypes.data, len(offsets) - 1, int(nthreads), C.byref(at))
    return float(m), int(at.value)


def residue_batch(P: Params, codes: np.ndarray, offsets: np.ndarray, nthreads=1) -> dict:
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    nprot = len(offsets) - 1
    ntot = int(offsets[-1])
    arrs = {n: np.zeros(ntot, dtype=np.uint8) for n in RESIDUE_U8}
    arrs.update({n: np.zeros(ntot, dtype=np.float64) for n in RESIDUE_F64})
    ro = ResidueOut(**{n: a.ctypes.data for n, a in arrs.items()})
    lib().orc_residue_batch(C.byref(P), codes.ctypes.data, offsets.ctypes.data, nprot, C.byref(ro), int(nthreads))
    return arrs


def java_fmt(x: float, decimals: int) -> str:
    buf = C.create_string_buffer(512)
    lib().orc_java_fmt(buf, 512, float(x), decimals)
    return buf.value.decode()


def max_threads() -> int:
    return int(lib().orc_max_threads())
