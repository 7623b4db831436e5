"""Transport-lean batch call (plaac_score_packed / plaac_score_hits, include/plaac_cuda.h): radix-22 words in,
ranked compact output.  CPU: the host packers.  GPU: records bit-identical to plaac_score's and to the oracle's, the
compact output equal to a numpy restatement of the web order (web/lib/server.rb:222-229) applied to the full table."""
import numpy as np
import pytest

import plaac_b200
from oracle import orc
from tests import parity, synth


def test_pack_roundtrip_and_layout():
    rng = np.random.default_rng(3)
    for n in (0, 1, 6, 7, 8, 13, 14, 15, 1000, 70001):
        codes = rng.integers(0, 22, n).astype(np.uint8)
        words = plaac_b200.pack_words(codes)
        assert len(words) == (n + 6) // 7 == plaac_b200.lib().plaac_packed_words(n)
        # layout: residue r is digit r % 7 of word r // 7, unused digits 0
        ref = np.zeros(len(words) * 7, dtype=np.uint64)
        ref[:n] = codes
        want = (ref.reshape(-1, 7) * (22 ** np.arange(7, dtype=np.uint64))).sum(1).astype(np.uint32)
        assert np.array_equal(words, want)
        assert np.array_equal(plaac_b200.unpack_words(words, 0, n), codes)
        for first, cnt in ((3, max(0, n - 5)), (7, max(0, n - 7)), (n // 2, n - n // 2)):
            if first + cnt <= n:
                assert np.array_equal(plaac_b200.unpack_words(words, first, cnt), codes[first:first + cnt])


def test_pack_append_builds_the_same_words_piece_by_piece():
    rng = np.random.default_rng(6)
    codes = rng.integers(0, 22, 100_003).astype(np.uint8)
    want = plaac_b200.pack_words(codes)
    words = np.full(len(want), 0xFFFFFFFF, dtype=np.uint32)  # (stale contents must not leak into the result)
    words[0] = 0
    pos = 0
    cuts = [0, 1, 2, 9, 10, 24, 700, 701, 7000, 7013, 50_000, 100_003]
    for a, b in zip(cuts, cuts[1:]):
        pos = plaac_b200.pack_append(codes[a:b], words, pos)
    assert pos == len(codes)
    assert np.array_equal(words, want)


def test_pack_threads_agree():
    rng = np.random.default_rng(4)
    codes = rng.integers(0, 22, 5_000_003).astype(np.uint8)
    assert np.array_equal(plaac_b200.pack_words(codes, nthreads=1), plaac_b200.pack_words(codes, nthreads=5))


def test_pack_chars_is_aatoint_then_pack():
    s = b"XACDEFGHIKLMNPQRSTVWY*acdefghiklmnpqrstvwyxBJOUZ -1\t>" * 3
    codes = plaac_b200.encode(s, strip_stop=False)
    assert np.array_equal(plaac_b200.pack_chars(s), plaac_b200.pack_words(codes))


def test_pack_rejects_bad_codes_after_packing_them_as_x():
    codes = np.array([1, 2, 30, 4], dtype=np.uint8)
    words = np.zeros(1, dtype=np.uint32)
    rc = plaac_b200.lib().plaac_pack_host(codes.ctypes.data, 4, words.ctypes.data, 1)
    assert rc == -1
    assert list(plaac_b200.unpack_words(words, 0, 4)) == [1, 2, 0, 4]


def web_order(rec):
    """server.rb:222-229 on the full-precision values: COREscore desc, LLR desc, NaN rows last; ties keep input order."""
    idx = np.arange(len(rec))
    has = ~np.isnan(rec["core_score"])
    key = np.lexsort((idx, -rec["llr"], -np.where(has, rec["core_score"], 0.0), ~has))
    return key, int(has.sum())


@pytest.fixture(scope="module")
def scorer():
    if plaac_b200.lib().plaac_device_count() < 1:
        pytest.skip("no GPU")
    s = plaac_b200.Scorer()
    yield s
    s.close()


@pytest.fixture(scope="module")
def proteome():
    codes, offsets = synth.proteome(3000, seed=77, prd_rate=0.15)
    return codes, offsets


@pytest.mark.gpu
def test_packed_records_equal_byte_call_and_oracle(scorer, proteome):
    codes, offsets = proteome
    lengths = np.diff(offsets).astype(np.int32)
    words = plaac_b200.pack_words(codes)
    full = scorer.score(codes, offsets)
    got = scorer.score_packed(words, lengths)
    assert got.tobytes() == full.tobytes()
    ref = orc.score_batch(orc.make_params(), codes, offsets)
    assert not parity.compare_summaries(got, ref, orc.INT_FIELDS, orc.DBL_FIELDS)


@pytest.mark.gpu
@pytest.mark.parametrize("chunk", [(0, 0), (20000, 0), (1 << 20, 100), (5000, 7)])
def test_packed_is_chunking_invariant(scorer, proteome, chunk):
    """Chunks start at arbitrary digits of a word (residue index % 7 != 0): every alignment is exercised."""
    codes, offsets = proteome
    lengths = np.diff(offsets).astype(np.int32)
    words = plaac_b200.pack_words(codes)
    full = scorer.score(codes, offsets)
    s2 = plaac_b200.Scorer()
    try:
        if chunk != (0, 0):
            s2.set_chunk(*chunk)
        got = s2.score_packed(words, lengths)
        assert got.tobytes() == full.tobytes()
    finally:
        s2.close()


@pytest.mark.gpu
def test_packed_per_residue_equals_byte_call(scorer):
    codes, offsets = synth.proteome(200, seed=5, prd_rate=0.2)
    lengths = np.diff(offsets).astype(np.int32)
    words = plaac_b200.pack_words(codes)
    s1, r1 = scorer.score(codes, offsets, per_residue=True)
    s2, r2 = scorer.score_packed(words, lengths, per_residue=True)
    assert s1.tobytes() == s2.tobytes()
    for k in r1:
        assert r1[k].tobytes() == r2[k].tobytes(), k


@pytest.mark.gpu
def test_hits_core_mode_is_the_head_of_the_web_order(scorer, proteome):
    codes, offsets = proteome
    lengths = np.diff(offsets).astype(np.int32)
    words = plaac_b200.pack_words(codes)
    full = scorer.score(codes, offsets)
    order, ncore = web_order(full)
    assert ncore > 50
    tab, hits = scorer.score_packed(words, lengths, hits="core")
    assert tab.tobytes() == full.tobytes()
    assert hits["n_core"] == ncore and len(hits["index"]) == ncore
    assert np.array_equal(hits["index"], order[:ncore])
    assert hits["records"].tobytes() == full[order[:ncore]].tobytes()
    # the same without the full table, from one-byte codes, and with a capacity below the number of hits
    h2 = scorer.score_hits(codes, offsets, "core")
    assert np.array_equal(h2["index"], order[:ncore]) and h2["records"].tobytes() == hits["records"].tobytes()
    _, h3 = scorer.score_packed(words, lengths, summaries=False, hits="core", capacity=10)
    assert h3["n_core"] == ncore and np.array_equal(h3["index"], order[:10])


@pytest.mark.gpu
def test_hits_topk(scorer, proteome):
    codes, offsets = proteome
    lengths = np.diff(offsets).astype(np.int32)
    words = plaac_b200.pack_words(codes)
    full = scorer.score(codes, offsets)
    order, ncore = web_order(full)
    k = ncore + 100
    _, h = scorer.score_packed(words, lengths, summaries=False, hits="topk", capacity=k)
    assert h["n_core"] == ncore and len(h["index"]) == k
    assert np.array_equal(h["index"], order[:k])
    assert h["records"].tobytes() == full[order[:k]].tobytes()


@pytest.mark.gpu
def test_hits_without_any_core(scorer):
    codes, offsets = synth.proteome(50, seed=9, prd_rate=0.0)
    h = scorer.score_hits(codes, offsets, "core")
    full = scorer.score(codes, offsets)
    _, ncore = web_order(full)
    assert h["n_core"] == ncore and len(h["index"]) == ncore


@pytest.mark.gpu
def test_packed_argument_errors(scorer):
    codes, offsets = synth.proteome(20, seed=1)
    lengths = np.diff(offsets).astype(np.int32)
    words = plaac_b200.pack_words(codes)
    with pytest.raises(plaac_b200.PlaacError):
        scorer.score_packed(words, lengths, nres=int(lengths.sum()) + 1)
    bad = lengths.copy()
    bad[3] = -1
    with pytest.raises(plaac_b200.PlaacError):
        scorer.score_packed(words, bad, nres=int(lengths.sum()))
    # a word that is not a packed word is flagged after the batch
    w2 = words.copy()
    w2[0] = 0xFFFFFFFF
    with pytest.raises(plaac_b200.PlaacError):
        scorer.score_packed(w2, lengths)
    # and the ctx still works
    assert scorer.score_packed(words, lengths).tobytes() == scorer.score(codes, offsets).tobytes()
