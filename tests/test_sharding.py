"""Multi-GPU host logic on CPU: the shard plan and the ordered gather, world_size 2 over gloo.

The data path has no collective (proteins are independent); what is tested is that shards are contiguous,
balanced, cover the batch exactly once, and that rank 0 ends up with every record in input order."""
import os
import socket

import numpy as np
import pytest

import plaac_b200
from oracle import orc
from tests import synth


def test_shard_plan_properties():
    codes, offs = synth.proteome(5000, seed=3)
    for n in (1, 2, 3, 8, 64):
        b = plaac_b200.shard_plan(offs, n)
        assert b[0] == 0 and b[-1] == len(offs) - 1 and (np.diff(b) >= 0).all()
        res = offs[b[1:]] - offs[b[:-1]]
        cost = res + 64 * np.diff(b)
        assert cost.max() - cost.min() <= offs.max() // 100 + np.diff(offs).max() + 64, (n, cost)
    # more shards than proteins, empty input, one giant protein
    b = plaac_b200.shard_plan(np.array([0, 5, 9], np.int64), 8)
    assert b[0] == 0 and b[-1] == 2 and (np.diff(b) >= 0).all() and np.diff(b).sum() == 2
    assert plaac_b200.shard_plan(np.array([0], np.int64), 4).tolist() == [0, 0, 0, 0, 0]
    b = plaac_b200.shard_plan(np.array([0, 10, 100010, 100020], np.int64), 2)
    assert np.diff(b).sum() == 3
    # offsets that do not start at 0 (a slice of a larger batch)
    b0 = plaac_b200.shard_plan(offs[100:2000], 4)
    b1 = plaac_b200.shard_plan(offs[100:2000] - offs[100], 4)
    assert b0.tolist() == b1.tolist()


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from plaac_b200 import shard

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        codes, offs = synth.proteome(400, seed=21, median=150.0)
        bounds = shard.shard_bounds(offs, world)
        c, o = shard.local_slice(codes, offs, bounds, rank)
        # on a CPU host the "scorer" of this test is the oracle; on a GPU box it is plaac_b200.Scorer
        mine = orc.score_batch(orc.make_params(), c, o)
        out = shard.gather_in_order(mine, bounds, rank, world)
        if rank == 0:
            whole = orc.score_batch(orc.make_params(), codes, offs)
            q.put(("ok", out.tobytes() == whole.tobytes(), bounds.tolist()))
        dist.barrier()
    except Exception as e:  # pragma: no cover
        q.put(("err", repr(e), None))
        raise
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_scoring_gathers_in_input_order():
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    status, same, bounds = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
    assert status == "ok", same
    assert same
    assert bounds[0] == 0 and bounds[-1] == 400 and 100 < bounds[1] < 300


@pytest.mark.gpu
def test_multi_ctx_scoring_matches_single(golden):
    """plaac_score_multi with two ctxs (on one or two GPUs) == plaac_score, summary and per-residue."""
    ndev = plaac_b200.lib().plaac_device_count()
    devices = [0, 1] if ndev >= 2 else [0, 0]
    codes, offs = synth.proteome(3000, seed=13, median=250.0)
    one = plaac_b200.Scorer()
    ref_s, ref_r = one.score(codes, offs, per_residue=True)
    one.close()
    ms = plaac_b200.MultiScorer(devices=devices)
    got_s, got_r = ms.score(codes, offs, per_residue=True)
    got_only = ms.score(codes, offs)
    ms.close()
    assert got_s.tobytes() == ref_s.tobytes() == got_only.tobytes()
    for k in ref_r:
        assert got_r[k].tobytes() == ref_r[k].tobytes(), k


@pytest.mark.gpu
@pytest.mark.parametrize("nctx", [2, 3])
def test_multi_ctx_packed_with_merged_hits(nctx):
    """plaac_score_multi_packed: shards start in the middle of a radix-22 word; the merged ranked hits equal the
    single-ctx ones (which tests/test_lean.py checks against the web order)."""
    ndev = plaac_b200.lib().plaac_device_count()
    devices = [k % max(ndev, 1) for k in range(nctx)]
    codes, offs = synth.proteome(3000, seed=14, median=250.0, prd_rate=0.2)
    lengths = np.diff(offs).astype(np.int32)
    words = plaac_b200.pack_words(codes)
    one = plaac_b200.Scorer()
    ref_s, ref_r = one.score(codes, offs, per_residue=True)
    _, ref_h = one.score_packed(words, lengths, hits="core")
    _, ref_t = one.score_packed(words, lengths, hits="topk", capacity=500)
    one.close()
    ms = plaac_b200.MultiScorer(devices=devices)
    got_s, got_r, got_h = ms.score_packed(words, lengths, per_residue=True, hits="core")
    _, got_t = ms.score_packed(words, lengths, hits="topk", capacity=500)
    ms.close()
    assert got_s.tobytes() == ref_s.tobytes()
    for k in ref_r:
        assert got_r[k].tobytes() == ref_r[k].tobytes(), k
    assert got_h["n_core"] == ref_h["n_core"] > 100
    assert np.array_equal(got_h["index"], ref_h["index"]) and got_h["records"].tobytes() == ref_h["records"].tobytes()
    assert np.array_equal(got_t["index"], ref_t["index"]) and got_t["records"].tobytes() == ref_t["records"].tobytes()


@pytest.mark.gpu
def test_multi_ctx_on_distinct_devices():
    """The single-process multi-GPU path (one ctx + host thread per GPU) on really different devices; skipped on a
    one-GPU box, where test_multi_ctx_scoring_matches_single runs both ctxs on device 0."""
    ndev = plaac_b200.lib().plaac_device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs (plaac_device_count() = %d)" % ndev)
    codes, offs = synth.proteome(20000, seed=15, median=300.0)
    lengths = np.diff(offs).astype(np.int32)
    words = plaac_b200.pack_words(codes)
    one = plaac_b200.Scorer(device=0)
    ref = one.score(codes, offs)
    one.close()
    ms = plaac_b200.MultiScorer(devices=list(range(ndev)))
    got = ms.score(codes, offs)
    got_p, hits = ms.score_packed(words, lengths, hits="core")
    ms.close()
    assert got.tobytes() == ref.tobytes() == got_p.tobytes()
    assert hits["n_core"] == int((~np.isnan(ref["core_score"])).sum())
