"""On-device ranking against a host restatement of web/lib/server.rb:222-229:
sort_by [COREscore NaN ? 1 : 0, -COREscore, LLR NaN ? 1 : 0, -LLR]; equal rows in input order."""
import numpy as np
import pytest

import plaac_b200
from tests import synth

pytestmark = pytest.mark.gpu


def web_order(rec):
    """The Ruby comparator on full-precision values, made total with the input index.  The jar prints the LLR of a
    protein shorter than the core length (-Inf) as NaN (inf2nan, plaac.java:904), and server.rb:225 sorts "NaN" last
    within a COREscore group -- where -Inf sorts too."""
    core = rec["core_score"].astype(np.float64)
    llr = rec["llr"].astype(np.float64).copy()
    nan_core = np.isnan(core)
    k_core = np.where(nan_core, 0.0, -core)   # "NaN".to_f == 0.0 -> key -0.0; all NaN rows compare equal here
    idx = np.arange(len(rec))
    # np.lexsort: last key is the primary one; it is a stable sort
    return np.lexsort((idx, -llr, k_core, nan_core.astype(np.int8))).astype(np.int32)


@pytest.fixture(scope="module")
def scorer():
    s = plaac_b200.Scorer(device=0)
    yield s
    s.close()


@pytest.mark.parametrize("nprot,seed", [(1, 1), (37, 2), (4096, 3), (4097, 4), (60000, 5)])
def test_rank_matches_web_order(scorer, nprot, seed):
    codes, offs = synth.proteome(nprot, seed, prd_rate=0.2, min_len=16)
    rec = scorer.score(codes, offs)
    order, ncore = scorer.rank(rec)
    assert ncore == int(np.count_nonzero(~np.isnan(rec["core_score"])))
    assert sorted(order.tolist()) == list(range(nprot))
    assert order.tolist() == web_order(rec).tolist()
    assert not np.isnan(rec["core_score"][order[:ncore]]).any()


def test_rank_ties_keep_input_order_and_short_proteins(scorer):
    # duplicated proteins give exactly equal (COREscore, LLR) pairs; proteins shorter than the core have LLR = -Inf
    codes, offs = synth.proteome(300, 11, prd_rate=0.5)
    seqs = [codes[offs[i]:offs[i + 1]] for i in range(300)]
    rng = np.random.default_rng(5)
    seqs = seqs + seqs[:150] + [rng.integers(1, 21, size=n).astype(np.uint8) for n in (1, 5, 30, 59)] + seqs[100:200]
    codes, offs = plaac_b200.pack(seqs)
    rec = scorer.score(codes, offs)
    assert np.isinf(rec["llr"]).sum() == 4
    plain, _ = scorer.rank(rec)
    assert plain.tolist() == web_order(rec).tolist()
    assert set(plain[-4:].tolist()) == set(range(450, 454))  # -Inf LLR (printed NaN) last, as the web sorts it


def test_rank_device_and_gather(scorer):
    import torch

    codes, offs = synth.proteome(20000, 21, prd_rate=0.1)
    rec = scorer.score(codes, offs)
    d_rec = torch.from_numpy(rec.view(np.uint8).reshape(len(rec), 160)).cuda()
    d_order = torch.zeros(len(rec), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ncore = scorer.rank_device(d_rec.data_ptr(), len(rec), d_order.data_ptr())
    ref = web_order(rec)
    assert d_order.cpu().numpy().tolist() == ref.tolist()
    d_top = torch.zeros((ncore, 160), dtype=torch.uint8, device="cuda")
    scorer.gather_device(d_rec.data_ptr(), d_order.data_ptr(), ncore, d_top.data_ptr())
    top = d_top.cpu().numpy().reshape(-1).view(plaac_b200.SUMMARY_DTYPE)
    assert top.tobytes() == rec[ref[:ncore]].tobytes()
    assert (np.diff(top["core_score"]) <= 0).all()


def test_rank_empty(scorer):
    order, ncore = scorer.rank(np.zeros(0, dtype=plaac_b200.SUMMARY_DTYPE))
    assert len(order) == 0 and ncore == 0
