"""Several live contexts on one GPU: the shared-memory limit of a kernel is a property of the function on the device,
not of a ctx: a ctx created later with smaller rings must not lower it under an earlier ctx."""
import numpy as np
import pytest

import plaac_b200
from oracle import orc
from tests import parity, synth


@pytest.mark.gpu
def test_two_live_contexts_large_ring_created_first():
    codes, offsets = synth.proteome(400, seed=21, prd_rate=0.2)
    big = plaac_b200.Scorer(plaac_b200.default_params(core_len=1200))        # v1 kernel, large residue ring
    mid = plaac_b200.Scorer(plaac_b200.default_params(core_len=300, ww1=81, ww2=81))
    small = plaac_b200.Scorer(plaac_b200.default_params(core_len=20, ww1=11, ww2=11))
    try:
        for sc, kw in ((small, dict(core_len=20, ww1=11, ww2=11)), (big, dict(core_len=1200)),
                       (mid, dict(core_len=300, ww1=81, ww2=81)), (big, dict(core_len=1200)), (small, dict(core_len=20, ww1=11, ww2=11))):
            got, res = sc.score(codes, offsets, per_residue=True)
            ref = orc.score_batch(orc.make_params(**kw), codes, offsets)
            assert not parity.compare_summaries(got, ref, orc.INT_FIELDS, orc.DBL_FIELDS)
            assert np.isfinite(res["post_bg"]).all()
    finally:
        for sc in (big, mid, small):
            sc.close()
