"""Randomised parity: random parameter sets (alpha, background, foreground, core / window sizes, proline rule), random
proteomes (length mix incl. proteins on both sides of the long-path threshold, composition mix, repeats, X and '*'
inside) and random entry points (host / device-resident API, chunk sizes, long-path threshold) against the CPU oracle.
Seeded: every case is reproducible from its index."""
import numpy as np
import pytest

import plaac_b200
from oracle import orc
from tests import parity, synth

import os

pytestmark = pytest.mark.gpu
NT = 16
N_SUMMARY = int(os.environ.get("PLAAC_FUZZ_N", "24"))   # PLAAC_FUZZ_N=400 for a soak run
N_RESIDUE = max(6, N_SUMMARY // 4)


def _random_case(seed):
    rng = np.random.default_rng(1000 + seed)
    w1 = int(rng.choice([5, 11, 21, 31, 41, 41, 41, 51, 61]))
    kw = dict(core_len=int(rng.choice([7, 20, 40, 60, 60, 60, 90, 150])), ww1=w1, ww2=w1 if rng.random() < 0.7 else w1 - 1 + (w1 % 2 == 0),
              adjust_prolines=bool(rng.random() < 0.8), alpha=float(rng.choice([1.0, 1.0, 0.0, 0.5, rng.random()])))
    if kw["ww2"] // 2 != kw["ww1"] // 2 or (kw["ww2"] - 1) // 2 != (kw["ww1"] - 1) // 2:
        kw["ww2"] = kw["ww1"]
    if rng.random() < 0.5 or kw["alpha"] < 1.0:  # alpha < 1 without any background is all zeros: log(0) everywhere
        kw["bg_counts"] = rng.integers(1, 5000, size=22).astype(np.float64)
    if rng.random() < 0.3:
        fg = rng.random(22) ** 2
        kw["fg_freq"] = fg / fg.sum()
    bg = rng.random(22) ** 1.5 + 0.01
    bg[[0, 21]] *= 0.01
    bg /= bg.sum()
    prd = synth.PRD_28 / synth.PRD_28.sum()
    seqs = []
    nprot = int(rng.integers(50, 1500))
    lens = np.clip(np.rint(rng.lognormal(np.log(rng.choice([60, 200, 400])), 0.8, nprot)), 1, 9000).astype(int)
    for n in lens:
        s = rng.choice(22, size=n, p=bg).astype(np.uint8)
        r = rng.random()
        if r < 0.25 and n > 30:
            st = int(rng.integers(0, n - 20))
            seg = int(rng.integers(10, min(400, n - st) + 1))
            s[st:st + seg] = rng.choice(22, size=seg, p=prd)
        elif r < 0.30:
            s[:] = np.resize(rng.choice(22, size=int(rng.integers(1, 6)), p=prd), n)  # perfect repeats: ties, plateaus
        seqs.append(s)
    for n in rng.choice([1100, 2100, 4096, 5000, 8200, 13000], size=int(rng.integers(0, 4)), replace=False):
        s = rng.choice(22, size=int(n), p=bg).astype(np.uint8)
        for _ in range(int(rng.integers(0, 4))):
            st = int(rng.integers(0, n - 300))
            s[st:st + 250] = rng.choice(22, size=250, p=prd)
        seqs.append(s)
    order = rng.permutation(len(seqs))
    seqs = [seqs[i] for i in order]
    api = dict(device=bool(rng.random() < 0.4), chunk=int(rng.choice([0, 0, 20000, 200000])),
               long_min=int(rng.choice([4096, 4096, 0, -1, 1024, 2048])))
    return kw, seqs, api


@pytest.mark.parametrize("seed", range(N_SUMMARY))
def test_random_case_against_oracle(seed):
    kw, seqs, api = _random_case(seed)
    codes, offs = plaac_b200.pack(seqs)
    P = orc.make_params(**kw)
    ref = orc.score_batch(P, codes, offs, nthreads=NT)
    sc = plaac_b200.Scorer(plaac_b200.default_params(**kw))
    sc.set_long_path(api["long_min"])
    if api["device"]:
        import torch

        d_codes = torch.from_numpy(codes).cuda() if len(codes) else torch.zeros(1, dtype=torch.uint8, device="cuda")
        d_offs = torch.from_numpy(offs).cuda()
        d_out = torch.zeros((len(seqs), 160), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        sc.score_device(d_codes.data_ptr(), d_offs.data_ptr(), len(seqs), int(offs[-1]), d_out.data_ptr())
        got = d_out.cpu().numpy().reshape(-1).view(plaac_b200.SUMMARY_DTYPE)
    else:
        if api["chunk"]:
            sc.set_chunk(api["chunk"], 300)
        got = sc.score(codes, offs)
    sc.close()
    bad, nties = parity.compare_with_tie_classes(P, codes, offs, got, ref, orc.INT_FIELDS, orc.DBL_FIELDS)
    assert not bad, f"seed {seed} {kw} {api}\n" + "\n".join(bad[:20])
    assert nties <= max(3, len(seqs) // 12), (seed, nties)   # perfect repeats are exact PAPA plateaus (documented tie class)
    for f in parity.REF_ORDER:
        if f != "papa_llr":
            assert parity.max_rel(got, ref, f) <= 1e-13, (seed, f, parity.max_rel(got, ref, f))


@pytest.mark.parametrize("seed", range(N_RESIDUE))
def test_random_case_per_residue_against_oracle(seed):
    """Per-residue mode with random parameters; the long proteins of the case (1.1 k - 13 k residues) stay in, so the
    random long-path threshold sends them through k_long_post (long_residue.cuh) under random HMM tables, window sizes
    and backgrounds -- with records (Viterbi bits from k_long_score) or, device-resident, without (Viterbi parse inside
    k_long_post)."""
    kw, seqs, api = _random_case(100 + seed)
    seqs = [s for s in seqs if len(s) <= 3000 or len(s) >= 1100][:300]
    seqs = [s for s in seqs if len(s) <= 3000][:260] + [s for s in seqs if len(s) > 3000]
    codes, offs = plaac_b200.pack(seqs)
    P = orc.make_params(**kw)
    ref = orc.residue_batch(P, codes, offs, nthreads=NT)
    sc = plaac_b200.Scorer(plaac_b200.default_params(**kw))
    sc.set_long_path(api["long_min"])
    if api["device"]:
        import torch

        ntot = int(offs[-1])
        d_codes = torch.from_numpy(np.concatenate([codes, np.zeros(64, np.uint8)])).cuda()
        d_offs = torch.from_numpy(offs).cuda()
        u8 = torch.zeros(2 * ntot, dtype=torch.uint8, device="cuda")
        f64 = torch.zeros(10 * ntot, dtype=torch.float64, device="cuda")
        ptrs = {"vit": u8.data_ptr(), "map": u8.data_ptr() + ntot}
        for k, nm in enumerate(plaac_b200.RESIDUE_F64):
            ptrs[nm] = f64.data_ptr() + 8 * k * ntot
        sc.score_device(d_codes.data_ptr(), d_offs.data_ptr(), len(seqs), ntot, 0, residue_ptrs=ptrs, sync=True)
        h8, hf = u8.cpu().numpy(), f64.cpu().numpy()
        got = {"vit": h8[:ntot], "map": h8[ntot:]}
        for k, nm in enumerate(plaac_b200.RESIDUE_F64):
            got[nm] = hf[k * ntot:(k + 1) * ntot]
    else:
        if api["chunk"]:
            sc.set_chunk(api["chunk"], 300)
        _, got = sc.score(codes, offs, per_residue=True)
    sc.close()
    assert (got["vit"] == ref["vit"]).all() and (got["map"] == ref["map"]).all(), (seed, kw, api)
    for f in orc.RESIDUE_F64:
        assert (np.isnan(got[f]) == np.isnan(ref[f])).all(), (seed, f)
        assert parity.close(got[f], ref[f], parity.SCALE[f]).all(), (seed, f, kw, api)
