"""The device algorithm of the rank phase, modelled in plain Python (tests/algo_model.py), against the reference's
interleave arrays and the oracle -- runs without a GPU."""
import numpy as np
import pytest

import algo_model as M


@pytest.mark.parametrize("name", ["merge_small", "merge_div", "merge_dup"])
@pytest.mark.parametrize("seg_len", [8, 104, 512])
def test_sliced_walk_reproduces_the_reference_interleave(oracle, golden, name, seg_len):
    g = golden(name)
    sym, ln = oracle.plain2runs(g["bwt0"])
    for b in range(1, int(g["n_batches"])):
        bwt = g["bwt%d" % b]
        ka, unres = M.interleave(oracle.runs2plain(sym, ln), bwt, seg_len)
        assert unres == 0 and np.array_equal(M.pack_rb(ka, bwt), g["rb%d" % b]), (name, b)
        sym, ln = oracle.merge_runs(sym, ln, g["rb%d" % b])


@pytest.mark.parametrize("name", ["merge_small", "merge_dup"])
@pytest.mark.parametrize("seg_len,warm", [(16, 0), (16, 8), (64, 16)])
def test_warm_up_rows_and_transfer_masks(oracle, golden, name, seg_len, warm):
    """k_walk_pair's warm-up and the mask-driven fix-up of k_fix_chain: x' = popcount(mask below x) must equal the rank
    chain at every tight row (asserted inside the model) and the result must still be the reference's array."""
    g = golden(name)
    sym, ln = oracle.plain2runs(g["bwt0"])
    for b in range(1, int(g["n_batches"])):
        bwt = g["bwt%d" % b]
        ka, unres = M.interleave(oracle.runs2plain(sym, ln), bwt, seg_len, warm=warm, masks=True)
        assert unres == 0 and np.array_equal(M.pack_rb(ka, bwt), g["rb%d" % b]), (name, b)
        sym, ln = oracle.merge_runs(sym, ln, g["rb%d" % b])


@pytest.mark.parametrize("n_parts", [2, 3, 5])
def test_sharded_slices_cover_the_batch(oracle, golden, n_parts):
    """Every part resolves its own slices from a halo; together they give the whole array (the all-reduce MAX)."""
    g = golden("merge_div")
    a, bwt = g["bwt0"], g["bwt1"]
    whole, _ = M.interleave(a, bwt, 64)
    got = np.full(len(bwt), -1, np.int64)
    for part in range(n_parts):
        ka, unres = M.interleave(a, bwt, 64, part=part, n_parts=n_parts, halo=8)
        assert unres == 0
        assert not np.any((ka >= 0) & (got >= 0))     # no row resolved twice
        got = np.maximum(got, ka)
    assert np.array_equal(got, whole)


@pytest.mark.parametrize("so", [1, 2])
def test_sorted_order_heads(oracle, golden, so):
    """RLO/RCLO: batch BWT by the closed form, merged with the head rule == the reference algorithm (BCR restatement)."""
    g = golden("rb2")
    lines = bytes(g["lines"]).decode().split()
    b = [int(x) for x in g["batch_bounds"]]
    batches = [oracle.encode_batch(lines[b[i]:b[i + 1]]) for i in range(3)]
    cur = oracle.sorted_bwt(batches[0], so)
    assert np.array_equal(cur, g["first_so%d" % so])
    ropes = [[] for _ in range(6)]
    oracle.insert_multi(ropes, batches[0], so)
    for t in batches[1:]:
        bwt = oracle.sorted_bwt(t, so)
        ka, unres = M.interleave(cur, bwt, 32, so=so)
        assert unres == 0 and np.all(np.diff(ka) >= 0)
        pos = ka + np.arange(len(bwt))
        merged = np.empty(len(cur) + len(bwt), np.uint8)
        mask = np.zeros(len(merged), bool)
        mask[pos] = True
        merged[mask] = bwt
        merged[~mask] = cur
        cur = merged
        oracle.insert_multi(ropes, t, so)
        assert np.array_equal(cur, np.array([c for r in ropes for c in r], np.uint8))


def test_ssa_from_walk_order(oracle, golden):
    """The data-parallel SSA rule of the device == ssa_gen1's per-string walk == the reference's .ssa files."""
    g = golden("ssa")
    for name, key in [("merge_small", "fmd"), ("rb2", "fmd_so2")]:
        s, l, _ = oracle.fmd_decode(bytes(golden(name)[key]))
        bwt = oracle.runs2plain(s, l)
        for ss in (0, 3, 8):
            assert M.ssa_image(bwt, ss) == bytes(g["%s_ss%d" % (name, ss)]), (name, ss)
