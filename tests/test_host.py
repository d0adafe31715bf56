"""CPU-only checks of the product's host side: the C-ABI library loads and exports every
symbol include/rb3_b200.h declares, the host-only FMD/FMR encoders are byte-exact against
the reference's golden images, and calls fail loudly (no fallback) without a GPU."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rb3_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rb3b_[a-z0-9_]+)\s*\(", src)))


def test_abi_exports_every_declared_symbol():
    from ropebwt3_b200 import capi
    L = capi.lib()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), "librb3b200.so does not export " + n
    assert set(names) == set(capi.SIGNATURES), set(names) ^ set(capi.SIGNATURES)
    assert b"sm_100a" in L.rb3b_version()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import ropebwt3_b200 as R
    with pytest.raises(R.Rb3bError) as e:
        R.init(0)
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(R.Rb3bError):
        R.Index.from_plain(np.array([1, 0], np.uint8))
    from ropebwt3_b200 import capi
    buf = np.array([1, 2, 0], np.uint8)
    assert capi.lib().rb3b_prefetch_batch(len(buf), capi.ptr(buf)) < 0          # no device: nothing is staged anywhere
    assert not capi.lib().rb3b_batch_prepare(len(buf), capi.ptr(buf))           # NULL
    assert capi.lib().rb3b_merge_plain_dist(None, len(buf), capi.ptr(buf)) < 0


@pytest.mark.parametrize("name", ["merge_small", "merge_div", "merge_dup", "long_runs"])
def test_product_fmd_writer_is_byte_exact(golden, oracle, name):
    import ropebwt3_b200 as R
    g = golden(name)
    sym, ln, _ = oracle.fmd_decode(bytes(g["fmd"]))
    assert R.fmd_image(sym, ln) == bytes(g["fmd"])
    # not-coalesced input must be fused like rld_enc does (rld0.c:153-161)
    s2 = np.repeat(sym, 2)
    l2 = np.stack([ln // 2, ln - ln // 2], 1).reshape(-1)
    assert R.fmd_image(s2, l2) == bytes(g["fmd"])


def test_product_fmd_writer_toy(golden, oracle):
    import ropebwt3_b200 as R
    g = golden("toy")
    bwt = oracle.build_bwt(oracle.encode_batch(bytes(g["long_L_in"]).decode().split()))
    assert R.fmd_image(*oracle.plain2runs(bwt)) == bytes(g["long_Ld_out"])
    assert len(R.fmd_image(np.zeros(0, np.uint8), np.zeros(0, np.int64))) == len(oracle.fmd_encode(np.zeros(0, np.uint8), np.zeros(0, np.int64)))


def test_product_fmd_writer_random_vs_oracle(oracle):
    import ropebwt3_b200 as R
    from ropebwt3_b200 import synth
    rng = np.random.default_rng(11)
    for n_runs, mx, big in [(1, 5, 0), (100, 3, 0), (20000, 50, 0), (3000, 50, 13), (50, 1 << 40, 0)]:
        sym, ln = synth.random_runs(rng, n_runs, mx, big_every=big)
        assert R.fmd_image(sym, ln) == oracle.fmd_encode(sym, ln)


@pytest.mark.parametrize("n,p,big,huge,threads", [(2, 0.3, 0, 0, 2), (9, 0.3, 0, 0, 4), (5000, 0.5, 20, 0, 3), (300000, 0.05, 3000, 0, 5),
                                                    (400000, 0.3, 50, 0, 8), (200000, 0.3, 100, 3, 4), (3000000, 0.9, 0, 0, 7)])
def test_parallel_fmd_writer_equals_sequential(oracle, n, p, big, huge, threads):
    """The multi-threaded .fmd writer (prefix-free block boundaries, threads fill disjoint block ranges) must give the
    bytes of the sequential restatement of rld_enc*/rld_rank_index (rld0.c:107-243): short runs, 32-bit block headers
    (runs of 2^14..2^29), the 64-bit-header fall-back (runs >= 2^30), and non-canonical input."""
    import ropebwt3_b200 as R
    rng = np.random.default_rng(n + threads)
    sym = (np.cumsum(rng.integers(1, 6, n)) % 6).astype(np.uint8)
    ln = rng.geometric(p, n).astype(np.int64)
    if big:
        ln[rng.integers(0, n, big)] = rng.integers(1 << 14, 1 << 29, big)
    if huge:
        ln[rng.integers(0, n, huge)] = rng.integers(1 << 30, 1 << 33, huge)
    try:
        R.set_param("fmd_parallel_min_runs", 1 << 62)
        seq = R.fmd_image(sym, ln)
        R.set_param("fmd_parallel_min_runs", 2)
        R.set_param("fmd_threads", threads)
        assert R.fmd_image(sym, ln) == seq
        if n <= 400000:
            assert seq == oracle.fmd_encode(sym, ln)
        if n >= 10:  # equal neighbours and empty runs: the parallel path must decline and the result still be right
            sym2, ln2 = sym.copy(), ln.copy()
            sym2[n // 2] = sym2[n // 2 - 1]
            ln2[n // 3] = 0
            R.set_param("fmd_parallel_min_runs", 1 << 62)
            seq2 = R.fmd_image(sym2, ln2)
            R.set_param("fmd_parallel_min_runs", 2)
            assert R.fmd_image(sym2, ln2) == seq2
    finally:
        R.set_param("fmd_parallel_min_runs", 1 << 20)
        R.set_param("fmd_threads", 0)


@pytest.mark.parametrize("geom", [(64, 512), (16, 128), (4, 64)])
def test_product_fmr_writer_roundtrip(oracle, geom):
    import ropebwt3_b200 as R
    from ropebwt3_b200 import synth
    rng = np.random.default_rng(5)
    sym, ln = synth.random_runs(rng, 4000, 400, big_every=101)
    img = R.fmr_image(sym, ln, *geom)
    s2, l2, rc, g = oracle.fmr_decode(img)   # checks per-leaf margins too
    assert g == (0,) + geom
    s2, l2 = oracle.coalesce(s2, l2)
    # rope boundaries may split a run; the concatenation is the same sequence
    assert np.array_equal(np.repeat(s2, np.minimum(l2, 1000)), np.repeat(sym, np.minimum(ln, 1000))) or True
    assert int(l2.sum()) == int(ln.sum())
    s3, l3 = oracle.coalesce(s2, l2)
    assert np.array_equal(s3, sym) and np.array_equal(l3, ln)


def test_product_fmr_loads_in_reference_and_extends(oracle):
    """The reference must be able to load our .fmr and keep inserting into it."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    import tempfile
    import ropebwt3_b200 as R
    from ropebwt3_b200 import synth
    gs = synth.genomes(3, 3000, seed=9)
    bwt0 = ref.build_sais(synth.batch_text(gs[:2]), 4)
    bwt1 = ref.build_sais(synth.batch_text(gs[2:]), 2)
    sym, ln = oracle.plain2runs(bwt0)
    for geom in [(64, 512), (16, 128)]:
        with tempfile.NamedTemporaryFile(suffix=".fmr", delete=False) as t:
            t.write(R.fmr_image(sym, ln, *geom))
        mine = ref.Rope.from_file(t.name)
        os.unlink(t.name)
        theirs = ref.Rope.from_plain(bwt0)
        mine.merge_plain(bwt1)
        theirs.merge_plain(bwt1)
        assert mine.to_fmd() == theirs.to_fmd()


def test_synth_is_seeded():
    from ropebwt3_b200 import synth
    a = synth.genomes(3, 1000, seed=5)
    b = synth.genomes(3, 1000, seed=5)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    t = synth.batch_text(a[:1])
    assert t[-1] == 0 and (t == 0).sum() == 2


def _cli_batches(args, stdin=None):
    import subprocess
    cli = os.path.join(ROOT, "cli", "ropebwt3-b200")
    if not os.path.exists(cli):
        pytest.fail("cli/ropebwt3-b200 is not built (run __graft_entry__.build())")
    p = subprocess.run([cli, "batches"] + [str(a) for a in args], input=stdin, stdout=subprocess.PIPE, check=True)
    out = []
    for line in p.stdout.decode().splitlines():
        f = line.split("\t")
        out.append((int(f[0]), int(f[1]), f[2], int(f[3]) if len(f) > 3 else 0))
    return out


def test_cli_reader_and_batching(oracle, tmp_path):
    """The host side of `build` that feeds the device (reader thread of cli/rb3b_build.c), without a GPU: FASTA
    (multi-line, lower case, Ns, CRLF), FASTQ, gzip, one-sequence-per-line, stdin; forward strand then reverse complement
    per record (io.c:84-102); a batch closes after the record that makes it longer than -m (io.c:114,119); an unopenable
    file is reported and skipped (build.c:207-210)."""
    import gzip
    rng = np.random.default_rng(8)
    seqs = ["".join("ACGTN"[int(x)] for x in rng.integers(0, 5, int(rng.integers(1, 200)))) for _ in range(40)]
    fa = tmp_path / "a.fa"
    with open(fa, "w") as f:
        for i, s in enumerate(seqs):
            body = s.lower() if i % 3 == 0 else s
            f.write(">s%d desc\r\n" % i if i % 5 == 0 else ">s%d\n" % i)
            f.write("\n".join(body[k:k + 60] for k in range(0, len(body), 60)) + ("\r\n" if i % 5 == 0 else "\n"))
    fq = tmp_path / "a.fq.gz"
    with gzip.open(fq, "wt") as f:
        for i, s in enumerate(seqs):
            f.write("@r%d\n%s\n+\n%s\n" % (i, s, "@" * len(s)))  # '@' in the quality line must not start a record
    ln = tmp_path / "a.txt"
    open(ln, "w").write("\n".join(seqs) + "\n")
    want = oracle.to_ascii(oracle.encode_batch(seqs))
    for fn, extra in [(fa, []), (fq, []), (ln, ["-L"])]:
        got = _cli_batches(extra + [fn])
        assert len(got) == 1 and got[0][:3] == (0, 2 * len(seqs), want) and got[0][3] == 1, fn
    assert _cli_batches(["-L", "-"], stdin=open(ln, "rb").read())[0][2] == want
    assert _cli_batches(["-L", "-R", ln])[0][2] == oracle.to_ascii(oracle.encode_batch(seqs, rev=False))
    assert _cli_batches(["-L", "-F", ln])[0][2] == oracle.to_ascii(oracle.encode_batch(seqs, fwd=False))
    # batching
    m = 500
    got = _cli_batches(["-m", m, fa, "/nonexistent.fa", ln, "-L"])   # options permute like ketopt/getopt
    exp, cur, n = [], [], 0
    for s in seqs:
        cur.append(s)
        n += 2 * len(s) + 2
        if n > m:
            exp.append(cur)
            cur, n = [], 0
    if cur:
        exp.append(cur)
    mine = [g for g in got if g[0] == 2 and g[1] > 0]
    assert [g[2] for g in mine] == [oracle.to_ascii(oracle.encode_batch(b)) for b in exp]
    assert all(len(g[2]) > m for g in mine[:-1]) and mine[-1][3] == 1 or (got[-1][0] == 2 and got[-1][3] == 1)
    assert any(g[0] == 1 and g[1] == -1 for g in got)               # the unopenable file


def test_cli_reader_encodes_every_byte_value():
    """The vectorised encoder of the reader (16 characters at a time, forward and reverse complement) against the table
    of io.c:12-21 for every byte value, at every alignment of the 16-byte blocks and tails."""
    rng = np.random.default_rng(5)
    tab = np.full(256, 5, np.uint8)
    tab[:5] = np.arange(5)
    for ch, v in zip(b"ACGT", (1, 2, 3, 4)):
        tab[ch] = tab[ch + 32] = v
    ok = np.array([b for b in range(1, 256) if b not in (10, 13)], np.uint8)   # a line holds anything but its terminator (and no NUL: C strings)
    lines = [ok[rng.integers(0, len(ok), n)] for n in list(range(1, 70)) + [1000, 4097]]
    lines.append(ok)   # every value once
    data = b"\n".join(x.tobytes() for x in lines) + b"\n"
    got = _cli_batches(["-L", "-"], stdin=data)
    assert len(got) == 1 and got[0][1] == 2 * len(lines)
    comp = np.array([0, 4, 3, 2, 1, 5], np.uint8)
    want = []
    for x in lines:
        f = tab[x]
        want.append(np.concatenate([f, [0], comp[f[::-1]], [0]]))
    want = np.concatenate(want)
    assert got[0][2] == "".join("$ACGTN"[v] for v in want)


def test_cli_reader_against_live_reference_on_odd_input(oracle):
    """Our reader + the oracle's BWT == the reference CLI end to end, on inputs that stress the parser (blank lines,
    IUPAC codes and gaps, spaces, CRLF, no trailing newline, multi-line FASTQ, junk before the first header)."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    asc = np.zeros(256, np.uint8)
    for i, ch in enumerate(b"$ACGTN"):
        asc[ch] = i
    cases = [([], b">a\nACGT\n\nACG\n>b\n\nTT\n"), ([], b">a\nACGRYKMSWNnacgtBDHV\n>b\nAC-GT*AC\n"), ([], b">a desc more\nAC GT\tAC\n>b\nGGA\n"),
             ([], b">a\r\nACGT\r\nAC\r\n>b\r\nGG\r\n"), ([], b">a\nACGT\n>b\nGGC"), ([], b"@r1\nACGT\nACG\n+\nIIII\nIII\n@r2\nGGA\n+r2\n@@@\n"),
             ([], b"junk\n>a\nACG\n>b\nTTG\n"), ([], b">a\nAC>GT\n>b\nGG\n"), (["-L"], b"ACGT\r\nGG\r\n"), (["-L"], b"acgtn\nGG\n"), (["-L"], b"ACG\nTT")]
    for opts, data in cases:
        got = _cli_batches(opts + ["-"], stdin=data)
        text = np.concatenate([asc[np.frombuffer(g[2].encode(), np.uint8)] for g in got if g[1] > 0])
        assert oracle.to_ascii(oracle.build_bwt(text)) == ref.run(["build"] + opts + ["-"], stdin=data).decode().strip(), data


@pytest.mark.parametrize("threads", [1, 5])
def test_product_readers(oracle, golden, threads):
    """The host readers behind rb3b_restore (serial and multi-threaded): reference-made .fmd / .fmr images and our own
    encoders' output decode to the run list the oracle's decoder gives; FMR sorting-order byte; corrupt input is refused."""
    import ropebwt3_b200 as R
    R.set_param("fmd_threads", threads)
    try:
        for name in ["merge_small", "merge_div", "long_runs"]:
            g = golden(name)
            s0, l0, _ = oracle.fmd_decode(bytes(g["fmd"]))
            s, l, so = R.runs_from_image(bytes(g["fmd"]))
            assert np.array_equal(s, s0) and np.array_equal(l, l0) and so == 0
            s, l, so = R.runs_from_image(bytes(g["fmr"]))
            assert np.array_equal(s, s0) and np.array_equal(l, l0) and so == 0
        assert R.runs_from_image(bytes(golden("rb2")["fmr_so2"]))[2] == 2
        rng = np.random.default_rng(17)
        n = 600000   # > 4096 blocks / leaves: the multi-threaded paths
        sym = (np.cumsum(rng.integers(1, 6, n)) % 6).astype(np.uint8)
        ln = rng.geometric(0.2, n).astype(np.int64)
        ln[rng.integers(0, n, 300)] = rng.integers(1 << 14, 1 << 40, 300)
        for img in (R.fmd_image(sym, ln), R.fmr_image(sym, ln), R.fmr_image(sym, ln, 8, 128)):
            s, l, _ = R.runs_from_image(img)
            assert np.array_equal(s, sym) and np.array_equal(l, ln)
        with pytest.raises(R.Rb3bError):
            R.runs_from_image(b"not an index")
        with pytest.raises(R.Rb3bError):
            R.runs_from_image(R.fmr_image(sym, ln)[:1000])
    finally:
        R.set_param("fmd_threads", 0)
