"""Parity at BASELINE scale against the UNMODIFIED reference (oracle/_ref, ctypes + CLI), default knobs.

The fixtures under tests/golden are 2.5-4 kb genomes; the code paths that only exist at real size -- 384/512-row slices
with ~170-row unresolved prefixes and long fix-up cascades, the `big` regime (>= 32 Mi rows per batch: seg_len 512,
fine_len 32), the bucketed scatter (>= 24 Mi rows), the bitmap -> run-length switch on a real index -- are pinned here:

  (a) 4 x 5 Mb genomes, one genome per merge (BASELINE configs[1]'s form): the device BWT of every batch equals libsais',
      the interleave array rb[] of every merge equals rb3_mg_rank_plain's (fm-index.c:202-225) and the final .fmd equals
      rb3_enc_fmr2fmd + rld_dump's, byte for byte, on both device layouts;
  (b) one batch of 4 genomes x 5 Mb (40 M rows) merged into a 2-genome index: rb[] and .fmd, bitmap and run-length cells;
  (c) the same merge with the bitmap budget set so that the index switches to run-length cells inside that merge;
  (d) rb3b_merge_index against `ropebwt3 merge` (main.c:84-133, rb3_fmi_merge fm-index.c:251-277).

The reference needs ~1 us per LF step and thread and has 2 chains per genome, so this file costs about two minutes of
host time on the GPU box; everything is compared bit for bit.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GENOME = 5_000_000
CORES = os.cpu_count() or 1


def _need_ref():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref did not travel")
    return ref


def _set_kind(rb3, kind):
    rb3.set_param("index_kind", {"auto": 0, "rle": 1, "bitmap": 2}[kind])


@pytest.fixture(scope="module")
def genomes6():
    from ropebwt3_b200 import synth
    return synth.genomes(6, GENOME, seed=43, sub=0.005, indel=0.0005)   # the first genomes of bench.py's set


@pytest.fixture(scope="module")
def ref_one_per_merge(genomes6):
    """The reference's own build of genomes 0..3, one genome per merge: libsais BWTs, rb[] of every merge, final .fmd."""
    ref = _need_ref()
    from ropebwt3_b200 import synth
    bwts, rbs, accs = [], [], []
    rope = None
    for g in genomes6[:4]:
        bwt = ref.build_sais(synth.batch_text([g]), 2, CORES)
        bwts.append(bwt)
        if rope is None:
            rope = ref.Rope.from_plain(bwt, CORES)
        else:
            rb, acc = rope.mg_rank_plain(bwt, CORES)
            rbs.append(rb); accs.append(acc)
            rope.merge_plain(bwt, CORES)
    return {"bwt": bwts, "rb": rbs, "acc": accs, "fmd": rope.to_fmd()}


@pytest.mark.parametrize("kind", ["bitmap", "rle"])
def test_a_one_genome_per_merge_5mb(rb3, genomes6, ref_one_per_merge, kind, tmp_path):
    from ropebwt3_b200 import synth
    R = ref_one_per_merge
    _set_kind(rb3, kind)
    try:
        idx = None
        for i, g in enumerate(genomes6[:4]):
            bwt = rb3.rb3_build_sais(synth.batch_text([g]))           # device suffix sort at 10^7 symbols
            assert np.array_equal(bwt, R["bwt"][i]), "device BWT of genome %d differs from libsais'" % i
            if idx is None:
                idx = rb3.Index.from_plain(bwt)
                continue
            rb, acc = idx.mg_rank_plain(bwt)
            assert np.array_equal(acc, R["acc"][i - 1])
            bad = np.flatnonzero(rb != R["rb"][i - 1])
            assert len(bad) == 0, "merge %d: %d of %d rb[] entries differ, first at row %d" % (i, len(bad), len(rb), bad[0])
            assert rb3.get_stat("seg_len_used") == (192 if kind == "bitmap" else 384) and rb3.get_stat("fix_segments") > 1000   # default knobs, real fix-up
            idx.merge_plain(bwt)
        assert rb3.get_stat("index_kind") == (1 if kind == "bitmap" else 0)
        fn = str(tmp_path / "a.fmd")
        idx.dump_fmd(fn)
        got = open(fn, "rb").read()
        assert got == R["fmd"], ".fmd differs from the reference's: %d vs %d bytes" % (len(got), len(R["fmd"]))
    finally:
        _set_kind(rb3, "auto")


def test_a2_prepared_batches_5mb(rb3, genomes6, ref_one_per_merge, tmp_path):
    """the same four genomes through rb3b_batch_prepare + rb3b_merge_prepared (walk order from the suffix sort)"""
    from ropebwt3_b200 import synth
    R = ref_one_per_merge
    idx = rb3.Index()
    for i, g in enumerate(genomes6[:4]):
        batch = rb3.Batch.prepare(synth.batch_text([g]))
        rb3.merge_prepared(idx, batch)
        batch.close()
    fn = str(tmp_path / "a2.fmd")
    idx.dump_fmd(fn)
    assert open(fn, "rb").read() == R["fmd"]


def test_f_rle_capacity_1e11_symbols(rb3):
    """A highly repetitive index of 10^11 symbols (2x10^7 runs of ~5000): run-length cells of 65536 positions hold it in
    about 0.2 GB; rank answers against a numpy restatement on the run list; a merge into it still works."""
    rng = np.random.default_rng(11)
    n_runs = 20_000_000
    sym = (np.cumsum(rng.integers(1, 6, n_runs, dtype=np.int8), dtype=np.int64) % 6).astype(np.uint8)
    ln = rng.integers(1, 10_000, n_runs).astype(np.int64)
    rb3.set_param("index_kind", 1)
    try:
        idx = rb3.Index.from_runs(sym, ln)
        n = int(ln.sum())
        assert len(idx) == n and n > 9e10
        assert rb3.get_stat("cell_shift") == 16 and idx.nbytes() < 1 << 29, (rb3.get_stat("cell_shift"), idx.nbytes())
        starts = np.concatenate([[0], np.cumsum(ln)])
        k = np.concatenate([rng.integers(0, n, 20000), starts[:3000], starts[1:3001] - 1, [n - 1, n, n + 5]]).astype(np.int64)
        ok, ret = idx.rank1a(k)
        r = np.minimum(np.searchsorted(starts, k, side="right") - 1, n_runs - 1)
        for a in range(6):
            pre = np.concatenate([[0], np.cumsum(np.where(sym == a, ln, 0))])
            want = pre[r] + np.where(sym[r] == a, np.minimum(k, n) - starts[r], 0)
            want = np.where(k >= n, pre[-1], want)
            assert np.array_equal(ok[:, a], want), a
        assert np.array_equal(ret[k < n], sym[r][k < n].astype(np.int8)) and (ret[k >= n] == -1).all()
        s2, l2 = idx.export_runs()
        assert np.array_equal(s2, sym) and np.array_equal(l2, ln)
    finally:
        rb3.set_param("index_kind", 0)


def test_e_fmd_device_encoder_chunk_boundary(rb3, tmp_path):
    """1.1e8 short runs = more than 2^20 blocks: the device encoder crosses a 2^23-word chunk boundary (the last block of a
    chunk is one word shorter, rld0.h:81) and must still equal the host writer byte for byte"""
    rng = np.random.default_rng(7)
    n = 110_000_000
    sym = (np.cumsum(rng.integers(1, 6, n, dtype=np.int8), dtype=np.int64) % 6).astype(np.uint8)
    ln = rng.integers(1, 4, n).astype(np.int64)
    idx = rb3.Index.from_runs(sym, ln)
    fn = str(tmp_path / "big.fmd")
    idx.dump_fmd(fn)
    assert rb3.get_stat("fmd_encoded_on_device") == 1 and rb3.get_stat("fmd_device_blocks") > (1 << 20)
    got = np.fromfile(fn, np.uint8)
    want = np.frombuffer(rb3.fmd_image(sym, ln), np.uint8)
    assert len(got) == len(want) and np.array_equal(got, want)


@pytest.fixture(scope="module")
def ref_big_batch(genomes6):
    """The reference's merge of ONE batch of genomes 2..5 (40 M rows, 8 chains) into the index of genomes 0..1."""
    ref = _need_ref()
    from ropebwt3_b200 import synth
    bwt0 = ref.build_sais(synth.batch_text(genomes6[:2]), 4, CORES)
    bwt1 = ref.build_sais(synth.batch_text(genomes6[2:6]), 8, CORES)
    rope = ref.Rope.from_plain(bwt0, CORES)
    rb, acc = rope.mg_rank_plain(bwt1, CORES)
    rope.merge_plain(bwt1, CORES)
    return {"bwt0": bwt0, "bwt1": bwt1, "rb": rb, "acc": acc, "fmd": rope.to_fmd()}


@pytest.mark.parametrize("mode", ["bitmap", "rle", "switch"])
def test_b_c_big_batch_into_index(rb3, ref_big_batch, mode, tmp_path):
    """>= 32 Mi rows in one batch: the `big` regime (seg_len 512, fine_len 32) and the bucketed scatter, on both layouts;
    mode "switch": the bitmap index outgrows its budget inside this merge and comes out as run-length cells."""
    R = ref_big_batch
    n0, n1 = len(R["bwt0"]), len(R["bwt1"])
    assert n1 >= 32 << 20
    if mode == "switch":
        _set_kind(rb3, "auto")
        rb3.set_param("bitmap_max_symbols", n0 + n1 // 2)
    else:
        _set_kind(rb3, mode)
    try:
        idx = rb3.Index.from_plain(R["bwt0"])
        assert rb3.get_stat("index_kind") == (0 if mode == "rle" else 1)
        rb, acc = idx.mg_rank_plain(R["bwt1"])
        assert np.array_equal(acc, R["acc"])
        bad = np.flatnonzero(rb != R["rb"])
        assert len(bad) == 0, "%d of %d rb[] entries differ, first at row %d" % (len(bad), len(rb), bad[0])
        assert rb3.get_stat("seg_len_used") == 512
        del rb
        idx.merge_plain(R["bwt1"])
        assert rb3.get_stat("index_kind") == (1 if mode == "bitmap" else 0)
        fn = str(tmp_path / "b.fmd")
        idx.dump_fmd(fn)
        got = open(fn, "rb").read()
        assert got == R["fmd"], ".fmd differs from the reference's: %d vs %d bytes" % (len(got), len(R["fmd"]))
        if mode == "switch":   # and the switched index keeps working: append one more genome prefix
            from ropebwt3_b200 import synth
            ref = _need_ref()
            extra = ref.build_sais(synth.batch_text([synth.genomes(1, 200_000, seed=5)[0]]), 2, CORES)
            idx.merge_plain(extra)
            rope = ref.Rope.from_plain(R["bwt0"], CORES)
            rope.merge_plain(R["bwt1"], CORES)
            rope.merge_plain(extra, CORES)
            idx.dump_fmd(fn)
            assert open(fn, "rb").read() == rope.to_fmd()
    finally:
        _set_kind(rb3, "auto")
        rb3.set_param("bitmap_max_symbols", 24_000_000_000)


@pytest.mark.parametrize("kind", ["bitmap", "rle"])
def test_d_merge_index_vs_reference_merge(rb3, oracle, kind, tmp_path):
    """`ropebwt3 merge a.fmd b.fmd` (main.c:84-133: rb3_fmi_merge with B an FM-index) == rb3b_merge_index."""
    ref = _need_ref()
    from ropebwt3_b200 import synth
    gs = synth.genomes(7, 300_000, seed=77, sub=0.01, indel=0.001)

    def fa(name, part):
        fn = str(tmp_path / name)
        with open(fn, "w") as f:
            for g in part:
                f.write(oracle.to_ascii(g) + "\n")
        return fn
    a_fmd, b_fmd, b2_fmr = str(tmp_path / "a.fmd"), str(tmp_path / "b.fmd"), str(tmp_path / "b2.fmr")
    open(a_fmd, "wb").write(ref.run(["build", "-L", "-d", "-t4", fa("a.txt", gs[:4])]))
    open(b_fmd, "wb").write(ref.run(["build", "-L", "-d", "-t4", fa("b.txt", gs[4:6])]))
    open(b2_fmr, "wb").write(ref.run(["build", "-L", "-b", "-t4", fa("b2.txt", gs[6:])]))
    merged = str(tmp_path / "m.fmr")
    open(merged, "wb").write(ref.run(["merge", "-t4", a_fmd, b_fmd, b2_fmr]))
    want = ref.run(["build", "-i", merged, "-d"])
    _set_kind(rb3, kind)
    try:
        A = rb3.Index.restore(a_fmd)
        for fn in (b_fmd, b2_fmr):
            B = rb3.Index.restore(fn)
            rb3.capi.check(rb3.capi.lib().rb3b_merge_index(A.h, B.h))
            B.close()
        out = str(tmp_path / "mine.fmd")
        A.dump_fmd(out)
        assert open(out, "rb").read() == want
    finally:
        _set_kind(rb3, "auto")
