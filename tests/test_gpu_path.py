"""Parity of the CUDA hot path (through the C ABI) against the CPU oracle, the golden
fixtures made with the unmodified reference, and -- when oracle/_ref travelled to the box --
the reference itself.  Everything here is integer work: the bar is bit-exact."""
import os
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MERGE_SETS = ["merge_small", "merge_div", "merge_dup"]


@pytest.fixture(params=["rle", "bitmap"], autouse=True)
def index_kind(request, rb3):
    """Every test runs on both device layouts: run-length cells and one-hot bitmap cells."""
    rb3.set_param("index_kind", 1 if request.param == "rle" else 2)
    yield request.param
    rb3.set_param("index_kind", 0)


def runs_of(idx, oracle):
    s, l = idx.export_runs()
    s2, l2 = oracle.coalesce(s, l)
    assert np.array_equal(s, s2) and np.array_equal(l, l2), "export_runs must already be canonical"
    return s, l


# ---------------------------------------------------------------- index + rank

@pytest.mark.parametrize("n_runs,max_len,big", [(1, 1, 0), (1, 100000, 0), (47, 9, 0), (48, 9, 0), (49, 9, 0),
                                                (5000, 40, 0), (200000, 6, 0), (3000, 5000, 0), (4000, 30, 17)])
def test_rank1a_random_runs(rb3, oracle, n_runs, max_len, big):
    from ropebwt3_b200 import synth
    rng = np.random.default_rng(n_runs + max_len)
    sym, ln = synth.random_runs(rng, n_runs, max_len, big_every=big)
    idx = rb3.Index.from_runs(sym, ln)
    n = int(ln.sum())
    assert len(idx) == n
    starts = np.concatenate([[0], np.cumsum(ln)])
    k = np.concatenate([rng.integers(0, n, 4000), starts[:2000], np.maximum(starts[1:2001] - 1, 0), [n, n + 1, n + 10**9]]).astype(np.int64)
    ok, ret = idx.rank1a(k)
    ok0, ret0 = oracle.rank1a(sym, ln, k)
    bad = np.flatnonzero((ok != ok0).any(1) | (ret != ret0))
    assert len(bad) == 0, "first mismatch at k=%d: got %s/%d want %s/%d" % (k[bad[0]], ok[bad[0]], ret[bad[0]], ok0[bad[0]], ret0[bad[0]])
    s2, l2 = runs_of(idx, oracle)
    assert np.array_equal(s2, sym) and np.array_equal(l2, ln)
    acc = idx.acc()
    tot = np.zeros(6, np.int64)
    np.add.at(tot, sym, ln)
    assert np.array_equal(np.diff(acc), tot)


def test_overflow_cells(rb3, oracle):
    """Mostly long runs (wide cells) with dense stretches of length-1 runs: cells with > 48 runs use overflow blocks."""
    rng = np.random.default_rng(5)
    sym, ln = [], []
    for rep in range(40):
        k = int(rng.integers(100, 3000))
        s = rng.integers(0, 6, k).astype(np.uint8)
        for i in range(1, k):
            if s[i] == s[i - 1]:
                s[i] = (s[i] + 1) % 6
        sym.append(s); ln.append(np.ones(k, np.int64))
        s2, l2 = np.array([(s[-1] + 1) % 6, (s[-1] + 2) % 6], np.uint8), rng.integers(50000, 2000000, 2)
        sym.append(s2); ln.append(l2.astype(np.int64))
    sym, ln = oracle.coalesce(np.concatenate(sym), np.concatenate(ln))
    rb3.set_param("index_kind", 1)
    idx = rb3.Index.from_runs(sym, ln)
    assert rb3.get_stat("cell_shift") >= 10 and rb3.get_stat("n_ovf_cells") > 0   # wide cells; the dense stretches cap the span at 2^11
    n = int(ln.sum())
    starts = np.concatenate([[0], np.cumsum(ln)])
    k = np.concatenate([rng.integers(0, n, 5000), starts[:-1], starts[1:] - 1, starts[rng.integers(0, len(starts) - 1, 3000)] + 1, [n]]).astype(np.int64)
    k = np.minimum(k, n)
    ok, ret = idx.rank1a(k)
    ok0, ret0 = oracle.rank1a(sym, ln, k)
    bad = np.flatnonzero((ok != ok0).any(1) | (ret != ret0))
    assert len(bad) == 0, "first mismatch at k=%d: got %s/%d want %s/%d" % (k[bad[0]], ok[bad[0]], ret[bad[0]], ok0[bad[0]], ret0[bad[0]])
    s2, l2 = runs_of(idx, oracle)
    assert np.array_equal(s2, sym) and np.array_equal(l2, ln)
    # merging into an index with overflow cells: insert a batch whose rows land everywhere
    text = np.concatenate([rng.integers(1, 5, 3000).astype(np.uint8), [0]])
    bwt = oracle.build_bwt(text)
    # positions do not matter for the writer test: use the oracle for both the interleave and the merge
    rb0, _ = oracle.mg_rank_plain(sym, ln, bwt)
    rb, _ = idx.mg_rank_plain(bwt)
    assert np.array_equal(rb, rb0)
    idx.merge_plain(bwt)
    s3, l3 = runs_of(idx, oracle)
    s4, l4 = oracle.merge_runs(sym, ln, rb0)
    assert np.array_equal(s3, s4) and np.array_equal(l3, l4)


def test_rank1a_golden(rb3, oracle, golden):
    for name in MERGE_SETS + ["long_runs"]:
        g = golden(name)
        sym, ln, _ = oracle.fmd_decode(bytes(g["fmd"]))
        idx = rb3.Index.from_runs(sym, ln)
        ok, ret = idx.rank1a(g["q_k"])
        assert np.array_equal(ok, g["q_ok"]) and np.array_equal(ret, g["q_ret"]), name


def test_from_plain_matches_runs(rb3, oracle, golden):
    g = golden("merge_small")
    bwt = g["bwt0"]
    idx = rb3.Index.from_plain(bwt)
    s, l = runs_of(idx, oracle)
    s0, l0 = oracle.plain2runs(bwt)
    assert np.array_equal(s, s0) and np.array_equal(l, l0)
    # empty and one-symbol BWTs (edge cases of rb3_enc_plain2fmr)
    e = rb3.Index.from_plain(np.zeros(0, np.uint8))
    assert len(e) == 0 and len(e.export_runs()[0]) == 0
    ok, ret = e.rank1a([0, 5])
    assert ret.tolist() == [-1, -1] and ok.sum() == 0
    one = rb3.Index.from_plain(np.array([0], np.uint8))
    ok, ret = one.rank1a([0, 1])
    assert ret.tolist() == [0, -1] and ok[1, 0] == 1
    with pytest.raises(rb3.Rb3bError):
        rb3.Index.from_plain(np.array([1, 2, 6, 0], np.uint8))  # fm-index.c:125 asserts symbols < 6


@pytest.mark.parametrize("variant", [0, 1, 2, 4, 8, 11, 22, 42, 44, 82, 84])
def test_lf_dev(rb3, oracle, variant):
    import torch
    from ropebwt3_b200 import synth
    rng = np.random.default_rng(77)
    sym, ln = synth.random_runs(rng, 30000, 25, big_every=1001)
    idx = rb3.Index.from_runs(sym, ln)
    n = int(ln.sum())
    k = np.concatenate([rng.integers(0, n + 1, 50000), [0, n]]).astype(np.int64)
    c = rng.integers(0, 6, len(k)).astype(np.uint8)
    dk, dc = torch.from_numpy(k).cuda(), torch.from_numpy(c).cuda()
    out = torch.empty(len(k), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    idx.lf_dev(len(k), dk.data_ptr(), dc.data_ptr(), out.data_ptr(), variant)
    rb3.sync()
    ok0, _ = oracle.rank1a(sym, ln, k)
    want = idx.acc()[c] + ok0[np.arange(len(k)), c]
    assert np.array_equal(out.cpu().numpy(), want)


# ---------------------------------------------------------------- the merge path

@pytest.mark.parametrize("seg_len", [16, 100, 2048])
@pytest.mark.parametrize("name", MERGE_SETS)
def test_merge_chain_golden(rb3, oracle, golden, name, seg_len):
    """rb3_mg_rank_plain and rb3_fmi_merge_plain, batch by batch, against reference outputs."""
    g = golden(name)
    rb3.set_param("seg_len", seg_len)
    try:
        idx = rb3.rb3_enc_plain2fmr(g["bwt0"])
        if seg_len == 100:
            idx.reserve(10 * len(g["bwt0"]))   # the size hint must not change any result
        for b in range(1, int(g["n_batches"])):
            bwt = g["bwt%d" % b]
            rb, acc = rb3.rb3_mg_rank_plain(idx, bwt)
            assert np.array_equal(acc, g["acc%d" % b])
            bad = np.flatnonzero(rb != g["rb%d" % b])
            assert len(bad) == 0, "%s batch %d: %d rows differ, first row %d got %d want %d (unresolved=%d rounds=%d)" % (
                name, b, len(bad), bad[0], rb[bad[0]] >> 6, g["rb%d" % b][bad[0]] >> 6, rb3.get_stat("unresolved_rows"), rb3.get_stat("fix_rounds"))
            rb3.rb3_fmi_merge_plain(idx, bwt)
            assert np.array_equal(idx.acc(), g["accA%d" % b])
        sym, ln = runs_of(idx, oracle)
        s0, l0, _ = oracle.fmd_decode(bytes(g["fmd"]))
        assert np.array_equal(sym, s0) and np.array_equal(ln, l0)
        with tempfile.TemporaryDirectory() as d:
            idx.dump_fmd(os.path.join(d, "x.fmd"))
            assert open(os.path.join(d, "x.fmd"), "rb").read() == bytes(g["fmd"])
        ok, ret = idx.rank1a(g["q_k"])
        assert np.array_equal(ok, g["q_ok"]) and np.array_equal(ret, g["q_ret"])
    finally:
        rb3.set_param("seg_len", 0)


@pytest.mark.parametrize("seg_len", [16, 0])
@pytest.mark.parametrize("name", MERGE_SETS)
def test_prepared_batches_golden(rb3, oracle, golden, name, seg_len):
    """The two-step form (rb3b_batch_prepare + rb3b_merge_prepared: walk order straight from the suffix sort, no LF chase)
    must build exactly what the reference built from the same texts."""
    g = golden(name)
    rb3.set_param("seg_len", seg_len)
    try:
        idx = rb3.Index()
        for b in range(int(g["n_batches"])):
            batch = rb3.Batch.prepare(g["text%d" % b])
            assert np.array_equal(batch.bwt(), g["bwt%d" % b]), (name, b)
            rb3.merge_prepared(idx, batch)
            batch.close()
            if b > 0:
                assert np.array_equal(idx.acc(), g["accA%d" % b])
        sym, ln = runs_of(idx, oracle)
        assert rb3.fmd_image(sym, ln) == bytes(g["fmd"])
    finally:
        rb3.set_param("seg_len", 0)


def test_prefetched_batches(rb3, oracle, golden):
    """rb3b_prefetch_batch: the copy of batch i+1 is queued before the merge of batch i; the consuming calls (merge_plain,
    batch_prepare) must find it by pointer + length, and a prefetched batch nobody consumes, or a second prefetch of the same
    slot, must be harmless."""
    from ropebwt3_b200 import capi
    L = capi.lib()
    g = golden("merge_div")
    n = int(g["n_batches"])
    bw = [np.ascontiguousarray(g["bwt%d" % b]) for b in range(n)]
    idx = rb3.Index.from_plain(bw[0])
    capi.check(L.rb3b_prefetch_batch(len(bw[1]), capi.ptr(bw[1])))
    for b in range(1, n):
        if b + 1 < n:
            capi.check(L.rb3b_prefetch_batch(len(bw[b + 1]), capi.ptr(bw[b + 1])))
        capi.check(L.rb3b_merge_plain(idx.h, len(bw[b]), capi.ptr(bw[b])))
        assert np.array_equal(idx.acc(), g["accA%d" % b])
    sym, ln = runs_of(idx, oracle)
    assert rb3.fmd_image(sym, ln) == bytes(g["fmd"])
    # texts through the two-step form, with a stray prefetch in between
    tx = [np.ascontiguousarray(g["text%d" % b]) for b in range(n)]
    idx2 = rb3.Index()
    for b in range(n):
        capi.check(L.rb3b_prefetch_batch(len(tx[b]), capi.ptr(tx[b])))
        capi.check(L.rb3b_prefetch_batch(len(bw[0]), capi.ptr(bw[0])))   # never consumed
        batch = rb3.Batch.prepare(tx[b])
        assert np.array_equal(batch.bwt(), bw[b])
        rb3.merge_prepared(idx2, batch)
        batch.close()
    s2, l2 = runs_of(idx2, oracle)
    assert np.array_equal(s2, sym) and np.array_equal(l2, ln)
    with pytest.raises(rb3.Rb3bError):
        capi.check(L.rb3b_prefetch_batch(0, capi.ptr(bw[0])))


def _fast_runs(rng, n_runs, max_len, big_every=0):
    sym = (np.cumsum(rng.integers(1, 6, n_runs)) % 6).astype(np.uint8)   # neighbours always differ
    ln = rng.integers(1, max_len + 1, n_runs).astype(np.int64)
    if big_every:
        ln[::big_every] = rng.integers(1, 1 << 18, len(ln[::big_every]))
    return sym, ln


@pytest.mark.parametrize("n_runs,max_len,big", [(5000, 9, 0), (300000, 6, 0), (300000, 40, 13), (200000, 3000, 3), (20000, 1 << 17, 0)])
def test_fmd_device_encoder(rb3, oracle, golden, tmp_path, n_runs, max_len, big):
    """rld_enc / enc_next_block / rld_rank_index on the device (rb3b_fmd_dev.cu): the file must equal the host writer's image,
    which tests/test_host.py pins to reference-made .fmd files -- 16-bit and 32-bit block headers, short and long runs."""
    rng = np.random.default_rng(n_runs + max_len + big)
    sym, ln = _fast_runs(rng, n_runs, max_len, big)
    idx = rb3.Index.from_runs(sym, ln)
    fn = str(tmp_path / "d.fmd")
    idx.dump_fmd(fn)
    assert rb3.get_stat("fmd_encoded_on_device") == 1
    got = open(fn, "rb").read()
    want = rb3.fmd_image(sym, ln)
    assert got == want, "device .fmd differs: %d vs %d bytes, first difference at byte %d" % (
        len(got), len(want), next((i for i in range(min(len(got), len(want))) if got[i] != want[i]), -1))
    rb3.set_param("fmd_device", 0)
    try:
        idx.dump_fmd(fn)
        assert rb3.get_stat("fmd_encoded_on_device") == 0 and open(fn, "rb").read() == want
    finally:
        rb3.set_param("fmd_device", 1)


def test_fmd_device_encoder_golden(rb3, golden, tmp_path):
    """the reference-made .fmd files, re-encoded on the device from the restored index (small lists forced onto the device)"""
    rb3.set_param("fmd_device_min_runs", 1)
    try:
        for name, key in [("merge_small", "fmd"), ("merge_div", "fmd"), ("merge_dup", "fmd"), ("long_runs", "fmd"), ("rb2", "fmd_so2")]:
            src = str(tmp_path / (name + ".fmd"))
            open(src, "wb").write(bytes(golden(name)[key]))
            idx = rb3.Index.restore(src)
            out = str(tmp_path / (name + ".out.fmd"))
            idx.dump_fmd(out)
            assert rb3.get_stat("fmd_encoded_on_device") == 1, name
            assert open(out, "rb").read() == bytes(golden(name)[key]), name
    finally:
        rb3.set_param("fmd_device_min_runs", 4096)


def test_merge_vs_oracle_seeded(rb3, oracle):
    """Fresh seeded inputs (not in the fixtures), device BWT construction included."""
    from ropebwt3_b200 import synth
    gs = synth.genomes(8, 20000, seed=101, sub=0.01, indel=0.001)
    gs.append(gs[3].copy())                       # an exact duplicate
    gs.append(np.full(3000, 1, np.uint8))         # a homopolymer
    gs.append(np.array([1, 2, 5, 5, 3, 4] * 50, np.uint8))  # Ns
    idx = sym = ln = None
    rb3.set_param("seg_len", 256)
    try:
        for i in range(0, len(gs), 2):
            text = synth.batch_text(gs[i:i + 2])
            bwt = rb3.rb3_build_sais(text)
            assert np.array_equal(bwt, oracle.build_bwt(text)), "device BWT differs from the oracle for batch %d" % i
            if idx is None:
                idx = rb3.Index.from_plain(bwt)
                sym, ln = oracle.plain2runs(bwt)
            else:
                rb, acc = idx.mg_rank_plain(bwt)
                rb0, acc0 = oracle.mg_rank_plain(sym, ln, bwt)
                assert np.array_equal(acc, acc0)
                assert np.array_equal(rb, rb0), "interleave array differs at batch %d" % i
                idx.merge_plain(bwt)
                sym, ln = oracle.merge_runs(sym, ln, rb0)
            s, l = runs_of(idx, oracle)
            assert np.array_equal(s, sym) and np.array_equal(l, ln), "merged runs differ at batch %d" % i
    finally:
        rb3.set_param("seg_len", 0)


def test_bitmap_to_rle_transition(rb3, oracle, golden):
    """An index that outgrows the 1 byte/symbol budget switches from bitmap cells to run-length cells in the merge."""
    g = golden("merge_small")
    rb3.set_param("index_kind", 0)
    rb3.set_param("bitmap_max_symbols", 3 * len(g["bwt0"]) + 10)
    try:
        idx = rb3.Index.from_plain(g["bwt0"])
        kinds = [rb3.get_stat("index_kind")]
        for b in range(1, int(g["n_batches"])):
            rb, _ = idx.mg_rank_plain(g["bwt%d" % b])
            assert np.array_equal(rb, g["rb%d" % b])
            idx.merge_plain(g["bwt%d" % b])
            kinds.append(rb3.get_stat("index_kind"))
        assert kinds[0] == 1 and kinds[-1] == 0, kinds
        s0, l0, _ = oracle.fmd_decode(bytes(g["fmd"]))
        s, l = runs_of(idx, oracle)
        assert np.array_equal(s, s0) and np.array_equal(l, l0)
    finally:
        rb3.set_param("bitmap_max_symbols", 24000000000)


@pytest.mark.parametrize("knob,value", [("fix_log", 0), ("wide_lf", 1), ("fine_len", 7), ("fine_len", 1), ("scatter_win_bits", 5), ("walk_pair", 0),
                                        ("warm_rows", 0), ("warm_rows", 40), ("mask_max_rows", 0), ("async_merge", 1), ("fix_tables", 1), ("fix_tpb", 16), ("fix_stages", 2), ("emit_staged", 0), ("piece_buf", 1), ("piece_buf", 0), ("rle_t1", 0)])
def test_optional_code_paths(rb3, oracle, golden, knob, value):
    """The tuning knobs select other kernels (multi-round generic fix-up, 64-bit LF table and rows, other mark spacing,
    the two-pass bucketed scatter of large batches, single-lane walks, no / longer warm-up before a slice, no transfer
    masks: every row of the fix-up takes the general step); every one of them must give the reference's interleave
    array and merged index."""
    g = golden("merge_dup")
    defaults = {"fix_log": 1, "wide_lf": 0, "fine_len": 0, "scatter_win_bits": 19, "walk_pair": 1, "warm_rows": 16, "mask_max_rows": 1 << 32, "async_merge": 0, "fix_tables": 0, "fix_tpb": 32, "fix_stages": 4, "emit_staged": 1, "piece_buf": -1, "rle_t1": 1}
    rb3.set_param(knob, value)
    if knob == "scatter_win_bits":
        rb3.set_param("scatter_bucket_min", 1)
    rb3.set_param("seg_len", 64)
    try:
        idx = rb3.Index.from_plain(g["bwt0"])
        for b in range(1, int(g["n_batches"])):
            rb, _ = idx.mg_rank_plain(g["bwt%d" % b])
            assert np.array_equal(rb, g["rb%d" % b]), (knob, b)
            idx.merge_plain(g["bwt%d" % b])
        s0, l0, _ = oracle.fmd_decode(bytes(g["fmd"]))
        s, l = runs_of(idx, oracle)
        assert np.array_equal(s, s0) and np.array_equal(l, l0)
    finally:
        rb3.set_param(knob, defaults[knob])
        rb3.set_param("scatter_bucket_min", 24 << 20)
        rb3.set_param("seg_len", 0)


def test_build_bwt_golden(rb3, golden):
    for name in MERGE_SETS:
        g = golden(name)
        for b in range(int(g["n_batches"])):
            assert np.array_equal(rb3.rb3_build_sais(g["text%d" % b]), g["bwt%d" % b]), (name, b)
    g = golden("reads")
    assert np.array_equal(rb3.rb3_build_sais(g["text"]), g["bwt"])
    with pytest.raises(rb3.Rb3bError):
        rb3.rb3_build_sais(np.array([1, 2, 3], np.uint8))  # no trailing sentinel (mrope.c:310)


@pytest.mark.parametrize("knob,value", [("sa_keys_only", 0), ("sa_discard", 0), ("sa_keys_x8", 0)])
def test_suffix_sorter_paths(rb3, golden, knob, value):
    """The suffix sorter's other paths -- 21-symbol key/value round 0 instead of the keys-only 15-symbol round of small
    batches, and refinement rounds that re-sort everything -- give libsais' BWT as well."""
    rb3.set_param(knob, value)
    try:
        for name in MERGE_SETS:
            g = golden(name)
            for b in range(int(g["n_batches"])):
                assert np.array_equal(rb3.rb3_build_sais(g["text%d" % b]), g["bwt%d" % b]), (knob, name, b)
        g = golden("reads")
        assert np.array_equal(rb3.rb3_build_sais(g["text"]), g["bwt"])
    finally:
        rb3.set_param(knob, 1)


def test_merge_errors(rb3, golden):
    g = golden("merge_small")
    idx = rb3.Index.from_plain(g["bwt0"])
    with pytest.raises(rb3.Rb3bError):
        idx.merge_plain(np.array([1, 2, 3, 4], np.uint8))      # no sentinel
    with pytest.raises(rb3.Rb3bError):
        idx.merge_plain(np.array([1, 0, 7], np.uint8))         # symbol out of range
    with pytest.raises(rb3.Rb3bError):
        idx.merge_plain(np.array([0, 1, 1], np.uint8))         # not a BWT: the A-cycle never reaches the sentinel
    # the index is untouched by failed merges
    assert np.array_equal(idx.acc(), g["accA0"])


def test_dump_restore_roundtrip(rb3, oracle, golden):
    g = golden("merge_div")
    with tempfile.TemporaryDirectory() as d:
        for ext, key in [("fmd", "fmd"), ("fmr", "fmr")]:
            fn = os.path.join(d, "ref." + ext)
            open(fn, "wb").write(bytes(g[key]))
            idx = rb3.Index.restore(fn)              # rb3_fmi_restore: both formats
            out = os.path.join(d, "mine.fmd")
            idx.dump_fmd(out)
            assert open(out, "rb").read() == bytes(g["fmd"])
            out = os.path.join(d, "mine.fmr")
            idx.dump_fmr(out)
            idx2 = rb3.Index.restore(out)
            assert np.array_equal(np.concatenate(idx2.export_runs()), np.concatenate(idx.export_runs()))
            out = os.path.join(d, "mine.txt")
            idx.dump_plain(out)
            s, l, _ = oracle.fmd_decode(bytes(g["fmd"]))
            assert open(out).read() == oracle.to_ascii(oracle.runs2plain(s, l)) + "\n"
        with pytest.raises(rb3.Rb3bError):
            rb3.Index.restore(os.path.join(d, "missing.fmd"))
        open(os.path.join(d, "junk"), "wb").write(b"hello world")
        with pytest.raises(rb3.Rb3bError):
            rb3.Index.restore(os.path.join(d, "junk"))


def test_incremental_append_equals_scratch(rb3, oracle, golden):
    """`-i` semantics: restore an existing index and append == building from scratch."""
    g = golden("merge_small")
    nb = int(g["n_batches"])
    with tempfile.TemporaryDirectory() as d:
        a = rb3.Index.from_plain(g["bwt0"])
        for b in range(1, 3):
            a.merge_plain(g["bwt%d" % b])
        fn = os.path.join(d, "half.fmr")
        a.dump_fmr(fn)
        r = rb3.Index.restore(fn)
        for b in range(3, nb):
            r.merge_plain(g["bwt%d" % b])
        out = os.path.join(d, "full.fmd")
        r.dump_fmd(out)
        assert open(out, "rb").read() == bytes(g["fmd"])


def test_merge_index(rb3, oracle, golden):
    """rb3_fmi_merge (BWT-vs-BWT): merging index B into A == merging B's plain BWT."""
    g = golden("merge_small")
    a = rb3.Index.from_plain(g["bwt0"])
    a.merge_plain(g["bwt1"])
    b = rb3.Index.from_plain(g["bwt2"])
    from ropebwt3_b200 import capi
    capi.check(capi.lib().rb3b_merge_index(a.h, b.h))
    c = rb3.Index.from_plain(g["bwt0"])
    c.merge_plain(g["bwt1"])
    c.merge_plain(g["bwt2"])
    assert np.array_equal(np.concatenate(a.export_runs()), np.concatenate(c.export_runs()))


@pytest.mark.parametrize("n_parts", [2, 3, 8])
def test_sharded_rank_phase(rb3, oracle, n_parts):
    """rb3b_mg_rank_part for every part (run one after the other on this GPU): MAX-combining the parts gives exactly the
    single-device interleave array; rb3b_merge_with_ka then equals rb3b_merge_plain."""
    import torch
    from ropebwt3_b200 import synth, capi
    L = capi.lib()
    gs = synth.genomes(5, 60000, seed=5, sub=0.01, indel=0.001)
    gs.append(gs[2].copy())                                   # an exact duplicate: its chains never collapse
    rb3.set_param("seg_len", 256)
    try:
        idx = rb3.Index.from_plain(rb3.rb3_build_sais(synth.batch_text(gs[:2])))
        for k, g in enumerate(gs[2:]):
            bwt = rb3.rb3_build_sais(synth.batch_text([g] if k % 2 else [g, g[:7000]]))
            n = len(bwt)
            rb, _ = idx.mg_rank_plain(bwt)
            full = (rb >> 6) - np.arange(n)
            d_bwt = torch.from_numpy(bwt).cuda()
            parts, rcs = [], []
            for p in range(n_parts):
                ka = torch.empty(n, dtype=torch.int64, device="cuda")
                torch.cuda.synchronize()
                rcs.append(capi.check(L.rb3b_mg_rank_part(idx.h, n, d_bwt.data_ptr(), p, n_parts, ka.data_ptr())))
                rb3.sync()
                parts.append(ka.cpu().numpy())
            if max(rcs) == 0:
                comb = np.max(np.stack(parts), 0)
                assert np.array_equal(comb, full), "parts=%d batch %d: %d rows differ" % (n_parts, k, int((comb != full).sum()))
                for p_ in parts:   # whatever a part did resolve is right
                    m = p_ >= 0
                    assert np.array_equal(p_[m], full[m])
            else:
                assert k == 3, "only the duplicate genome may need the fallback"
                comb = full
            d_ka = torch.from_numpy(comb).cuda()
            torch.cuda.synchronize()
            capi.check(L.rb3b_merge_with_ka(idx.h, n, d_bwt.data_ptr(), d_ka.data_ptr()))
            rb3.sync()
        s, l = runs_of(idx, oracle)
        sym, ln = None, None
        want = rb3.Index.from_plain(rb3.rb3_build_sais(synth.batch_text(gs[:2])))
        for k, g in enumerate(gs[2:]):
            want.merge_plain(rb3.rb3_build_sais(synth.batch_text([g] if k % 2 else [g, g[:7000]])))
        s0, l0 = want.export_runs()
        assert np.array_equal(s, s0) and np.array_equal(l, l0)
        # holes are refused
        bad = torch.full((n,), -1, dtype=torch.int64, device="cuda")
        with pytest.raises(rb3.Rb3bError):
            capi.check(L.rb3b_merge_with_ka(idx.h, n, d_bwt.data_ptr(), bad.data_ptr()))
    finally:
        rb3.set_param("seg_len", 0)


def test_device_pointer_entry_points(rb3, golden):
    import torch
    g = golden("merge_small")
    d0 = torch.from_numpy(g["bwt0"]).cuda()
    d1 = torch.from_numpy(g["bwt1"]).cuda()
    torch.cuda.synchronize()
    idx = rb3.Index.from_plain_dev(d0.data_ptr(), len(g["bwt0"]))
    rb = torch.empty(len(g["bwt1"]), dtype=torch.int64, device="cuda")
    acc = idx.mg_rank_plain_dev(d1.data_ptr(), len(g["bwt1"]), rb.data_ptr())
    rb3.sync()
    assert np.array_equal(rb.cpu().numpy(), g["rb1"]) and np.array_equal(acc, g["acc1"])
    idx.merge_plain_dev(d1.data_ptr(), len(g["bwt1"]))
    rb3.sync()
    assert np.array_equal(idx.acc(), g["accA1"])


def test_against_reference_binary_midsize(rb3, oracle):
    """20 x 50 kb genomes merged one per batch; the .fmd must equal the reference CLI's byte for byte."""
    from oracle import ref
    if not os.path.exists(ref.BIN):
        pytest.skip("oracle/_ref/ropebwt3 did not travel")
    from ropebwt3_b200 import synth
    gs = synth.genomes(20, 50000, seed=44)
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "g.txt")
        with open(fa, "w") as f:
            for g_ in gs:
                f.write(oracle.to_ascii(g_) + "\n")
        want = ref.run(["build", "-L", "-d", "-t4", fa])
        idx = None
        for g_ in gs:
            bwt = rb3.rb3_build_sais(synth.batch_text([g_]))
            if idx is None:
                idx = rb3.Index.from_plain(bwt)
            else:
                idx.merge_plain(bwt)
        out = os.path.join(d, "mine.fmd")
        idx.dump_fmd(out)
        got = open(out, "rb").read()
        assert got == want, "fmd differs from the reference: %d vs %d bytes" % (len(got), len(want))


# ---------------------------------------------------------------- ropebwt2 insertion order (build -2/-s/-r)

def _rb2_batches(oracle, g):
    lines = bytes(g["lines"]).decode().split()
    b = [int(x) for x in g["batch_bounds"]]
    return lines, [oracle.encode_batch(lines[b[i]:b[i + 1]]) for i in range(len(b) - 1)]


@pytest.mark.parametrize("so", [0, 1, 2])
def test_build_bwt_sorted_orders(rb3, oracle, golden, so):
    """BWT of one batch in input / RLO / RCLO order == what the reference's mr_insert_multi builds on an empty rope."""
    g = golden("rb2")
    lines, batches = _rb2_batches(oracle, g)
    assert np.array_equal(rb3.build_bwt_so(batches[0], so), g["first_so%d" % so])
    assert np.array_equal(rb3.build_bwt_so(oracle.encode_batch(lines), so), g["bwt_so%d" % so])
    toy = oracle.encode_batch(["AGG", "AGC"])
    assert oracle.to_ascii(rb3.build_bwt_so(toy, so)) == ["GTCT$$G$CGGA$ACC", "CGTT$$G$CGGA$ACC", "TTGC$$G$GCGA$ACC"][so]


@pytest.mark.parametrize("so", [0, 1, 2])
@pytest.mark.parametrize("seg_len", [16, 512])
def test_insert_multi_batches(rb3, oracle, golden, so, seg_len):
    """mr_insert_multi batch after batch (build -2/-s/-r -m ...): the index after every batch equals the reference
    algorithm's (oracle BCR restatement), the final one the reference CLI's output and .fmd."""
    g = golden("rb2")
    lines, batches = _rb2_batches(oracle, g)
    rb3.set_param("seg_len", seg_len)
    try:
        idx = rb3.Index()
        idx.set_order(so)
        ropes = [[] for _ in range(6)]
        for i, t in enumerate(batches):
            rb3.mr_insert_multi(idx, t)
            if i < 3:
                oracle.insert_multi(ropes, t, so)
                s0, l0 = oracle.plain2runs(np.array([c for r in ropes for c in r], np.uint8))
                s, l = runs_of(idx, oracle)
                assert np.array_equal(s, s0) and np.array_equal(l, l0), (so, i)
        s, l = runs_of(idx, oracle)
        assert np.array_equal(oracle.runs2plain(s, l), g["bwt_so%d" % so])
        assert rb3.fmd_image(s, l) == bytes(g["fmd_so%d" % so])
        assert idx.get_order() == so
    finally:
        rb3.set_param("seg_len", 0)


def test_insert_multi_duplicates_and_order_in_fmr(rb3, oracle, golden, tmp_path):
    """Inserting the same reads again (every new string equals an old one: the sorted-order heads run to the start of the
    string) and the sorting order surviving an .fmr round trip."""
    g = golden("rb2")
    lines, batches = _rb2_batches(oracle, g)
    for so in (1, 2):
        idx = rb3.Index()
        idx.set_order(so)
        rb3.mr_insert_multi(idx, batches[0])
        fn = str(tmp_path / ("x%d.fmr" % so))
        idx.dump_fmr(fn)
        idx2 = rb3.Index.restore(fn)
        assert idx2.get_order() == so
        rb3.mr_insert_multi(idx2, batches[0])
        rb3.mr_insert_multi(idx2, batches[1])
        want = oracle.rb2_bwt([batches[0], batches[0], batches[1]], so)
        s, l = runs_of(idx2, oracle)
        assert np.array_equal(oracle.runs2plain(s, l), want)
    assert oracle.fmr_decode(bytes(g["fmr_so2"]))[3][0] == 2  # the reference writes the order there too


# ---------------------------------------------------------------- sampled suffix array (ropebwt3 ssa)

@pytest.mark.parametrize("wide", [0, 1])
def test_ssa_golden(rb3, oracle, golden, tmp_path, wide):
    """rb3_ssa_gen + rb3_ssa_dump on the device == the reference's `ropebwt3 ssa -s SS idx.fmd`, byte for byte."""
    g = golden("ssa")
    rb3.set_param("wide_lf", wide)
    try:
        for name, key in [("merge_small", "fmd"), ("rb2", "fmd_so2")]:
            fmd = str(tmp_path / (name + ".fmd"))
            open(fmd, "wb").write(bytes(golden(name)[key]))
            idx = rb3.Index.restore(fmd)
            for ss in (0, 3, 8):
                out = str(tmp_path / ("%s.%d.ssa" % (name, ss)))
                idx.ssa_dump(out, ss)
                assert open(out, "rb").read() == bytes(g["%s_ss%d" % (name, ss)]), (name, ss)
    finally:
        rb3.set_param("wide_lf", 0)


def test_ssa_larger_against_oracle(rb3, oracle, tmp_path):
    """20 genomes of 30 kb merged on the device, then the SSA of the result against the oracle restatement."""
    from ropebwt3_b200 import synth
    gs = synth.genomes(8, 30000, seed=9)
    idx = rb3.Index.from_plain(rb3.rb3_build_sais(synth.batch_text(gs[:4])))
    idx.merge_plain(rb3.rb3_build_sais(synth.batch_text(gs[4:])))
    s, l = idx.export_runs()
    out = str(tmp_path / "x.ssa")
    idx.ssa_dump(out, 6)
    assert open(out, "rb").read() == oracle.ssa_image(oracle.runs2plain(s, l), 6)


# ---------------------------------------------------------------- BASELINE config 1 at full size

def test_config1_full_size_batching_invariance(rb3, index_kind):
    """BASELINE.json configs[1] at its full size (100 synthetic 5 Mb genomes, 10^9 symbols; bench.py's exact set): the
    reference's canonical-index property (the same collection gives the same BWT for any batching) checked
    between 99 merges of one genome and 9 merges of ten genomes, plus the symbol totals the collection must have."""
    if index_kind == "rle":
        pytest.skip("full size is run on the default (bitmap) layout only")
    import torch
    from ropebwt3_b200 import synth, capi
    gs = synth.genomes(100, 5_000_000, seed=43, sub=0.005, indel=0.0005)
    expect = np.zeros(6, np.int64)
    for g in gs:
        cnt = np.bincount(g, minlength=6)[:6]
        expect += cnt + cnt[[0, 4, 3, 2, 1, 5]]
        expect[0] += 2

    def build(per_merge):
        idx = None
        for b in range(0, len(gs), per_merge):
            text = synth.batch_text(gs[b:b + per_merge])
            d_text = torch.from_numpy(text).cuda()
            d_bwt = torch.empty_like(d_text)
            capi.check(capi.lib().rb3b_build_bwt_dev(len(text), d_text.data_ptr(), d_bwt.data_ptr()))
            if idx is None:
                idx = rb3.Index.from_plain_dev(d_bwt.data_ptr(), len(text))
                idx.reserve(int(expect.sum()))
            else:
                idx.merge_plain_dev(d_bwt.data_ptr(), len(text))
            rb3.sync()
        return idx

    one, ten = build(1), build(10)
    assert np.array_equal(np.diff(one.acc()), expect) and np.array_equal(np.diff(ten.acc()), expect)
    s1, l1 = one.export_runs()
    s10, l10 = ten.export_runs()
    assert len(s1) == len(s10) and np.array_equal(s1, s10) and np.array_equal(l1, l10)
    rng = np.random.default_rng(1)
    k = rng.integers(0, int(expect.sum()), 2000).astype(np.int64)
    ok1, r1 = one.rank1a(k)
    ok10, r10 = ten.rank1a(k)
    assert np.array_equal(ok1, ok10) and np.array_equal(r1, r10)
