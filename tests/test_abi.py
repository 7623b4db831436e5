"""CPU-side checks of the boundary: the library loads, exports every symbol include/*.h declares,
the struct layouts match the header, and the host-side parameter chain equals the oracle's."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import plaac_b200
from oracle import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(plaac_[a-z0-9_]+)\s*\(", txt)))


@pytest.mark.parametrize("header", ["plaac_cuda.h", "plaac_bench.h"])
def test_library_exports_every_declared_symbol(header):
    L = plaac_b200.lib()
    names = _declared(header)
    assert len(names) >= 3
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/{header} but not exported"


def test_struct_layouts():
    assert C.sizeof(plaac_b200.Summary) == 160
    assert plaac_b200.SUMMARY_DTYPE.itemsize == 160
    assert [n for n, _ in plaac_b200.Summary._fields_] == list(orc.SUMMARY_DTYPE.names)
    # 8 int32 + (4+2+2+44+22*5+3+2+4001) doubles
    assert C.sizeof(plaac_b200.Params) == 32 + 8 * (4 + 2 + 2 + 44 + 110 + 3 + 2 + 4001)


@pytest.mark.parametrize("kw", [dict(), dict(alpha=0.5, bg_counts=np.arange(22) * 10.0 + 3), dict(alpha=7.0),
                                dict(alpha=0.0, bg_counts=np.ones(22)), dict(core_len=30, ww1=21, ww2=31)])
def test_host_parameter_chain_equals_oracle(kw):
    P, info = plaac_b200.default_params(return_info=True, **kw)
    Q = orc.make_params(**kw)
    for f in ["llr", "papa_lod", "hydro2", "charge", "loglut", "fi_cc"]:
        assert list(getattr(P, f)) == list(getattr(Q, f)), f
    assert [list(r) for r in P.le] == [list(r) for r in Q.hmm1.le]
    assert list(P.le0) == list(Q.hmm0.le[0])
    assert [list(r) for r in P.lt] == [list(r) for r in Q.hmm1.lt]
    assert list(P.li) == list(Q.hmm1.li) and list(P.lf) == list(Q.hmm1.lf) and P.ln2 == Q.ln2
    assert (P.core_len, P.ww1, P.ww2, P.ww3, P.mw_window) == (Q.core_len, Q.ww1, Q.ww2, Q.ww3, 80)
    assert list(info[0]) == list(Q.fg) and list(info[1]) == list(Q.bgscer)
    assert list(info[2]) == list(Q.bgthis) and list(info[3]) == list(Q.bg)


def test_encode_host():
    s = "XACDEFGHIKLMNPQRSTVWY*acdefghiklmnpqrstvwyxBJOUZ -1\t>"
    assert list(plaac_b200.encode(s, strip_stop=False)) == [orc.lib().orc_aatoint(ord(c)) for c in s]


def test_no_gpu_fails_loudly():
    """Without a device the product reports an error code; it never falls back to a CPU path."""
    if plaac_b200.lib().plaac_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(plaac_b200.PlaacError) as e:
        plaac_b200.Scorer()
    assert e.value.code == -5
    # the pinned-memory helpers report an error too (no crash, no silent pageable fallback)
    with pytest.raises(plaac_b200.PlaacError):
        plaac_b200.PinnedBuffer(1024)
    with pytest.raises(plaac_b200.PlaacError):
        plaac_b200.host_register(np.zeros(4096, np.uint8))
