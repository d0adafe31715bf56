"""Pins the CPU oracle (oracle/rb3_oracle.c) to outputs of the unmodified reference:
the committed golden fixtures (tests/golden/make_golden.py) and, when oracle/_ref is
present, live differential runs.  No GPU."""
import numpy as np
import pytest

MERGE_SETS = ["merge_small", "merge_div", "merge_dup"]


def txt(a):
    return bytes(a).decode()


def test_toy_known_answers(oracle, golden):
    g = golden("toy")
    # literal answers recorded in SURVEY 4.4
    assert txt(g["agg_LR_out"]) == "GC$$GGAA\n"
    assert txt(g["agg_L_out"]) == "GTCT$$G$CGGA$ACC\n"
    assert txt(g["nn_L_out"]) == "TT$$AANNGGNNCC\n"
    assert txt(g["long_L_out"]).strip() == "GACCACGCAGCTACAATGACTTTAACATAATA$ATTATTTGTATCGAATGC$GTGTTAGCTGTA"
    assert len(g["long_Ld_out"]) == 272 and len(g["long_Lb_out"]) == 424
    # oracle BWT construction reproduces them
    for key, kw in [("agg_L", {}), ("agg_LR", {"rev": False}), ("long_L", {}), ("nn_L", {})]:
        seqs = txt(g[key + "_in"]).split()
        bwt = oracle.build_bwt(oracle.encode_batch(seqs, **kw))
        assert oracle.to_ascii(bwt) == txt(g[key + "_out"]).strip()
    # and the FMD / FMR images of the long toy
    bwt = oracle.build_bwt(oracle.encode_batch(txt(g["long_L_in"]).split()))
    sym, ln = oracle.plain2runs(bwt)
    assert oracle.fmd_encode(sym, ln) == bytes(g["long_Ld_out"])
    s2, l2, _, geom = oracle.fmr_decode(bytes(g["long_Lb_out"]))
    s2, l2 = oracle.coalesce(s2, l2)
    assert np.array_equal(s2, sym) and np.array_equal(l2, ln) and geom == (0, 64, 512)


@pytest.mark.parametrize("name", MERGE_SETS)
def test_merge_chain_against_golden(oracle, golden, name):
    g = golden(name)
    nb = int(g["n_batches"])
    sym = ln = None
    for b in range(nb):
        text, bwt = g["text%d" % b], g["bwt%d" % b]
        assert np.array_equal(oracle.build_bwt(text), bwt)  # sais-ss.c semantics
        if b == 0:
            sym, ln = oracle.plain2runs(bwt)
        else:
            rb, acc = oracle.mg_rank_plain(sym, ln, bwt)
            assert np.array_equal(rb, g["rb%d" % b])        # fm-index.c:160-225
            assert np.array_equal(acc, g["acc%d" % b])
            sym, ln = oracle.merge_runs(sym, ln, rb)        # fm-index.c:237-249
        acc_a = np.zeros(7, np.int64)
        np.add.at(acc_a, sym.astype(np.int64) + 1, ln)
        assert np.array_equal(np.cumsum(acc_a), g["accA%d" % b])
    ok, ret = oracle.rank1a(sym, ln, g["q_k"])
    assert np.array_equal(ok, g["q_ok"]) and np.array_equal(ret, g["q_ret"])
    assert oracle.fmd_encode(sym, ln) == bytes(g["fmd"])    # canonical .fmd, byte for byte
    s2, l2, mc = oracle.fmd_decode(bytes(g["fmd"]))
    assert np.array_equal(s2, sym) and np.array_equal(l2, ln)
    s3, l3, _, _ = oracle.fmr_decode(bytes(g["fmr"]))
    s3, l3 = oracle.coalesce(s3, l3)
    assert np.array_equal(s3, sym) and np.array_equal(l3, ln)


def test_reads_and_long_runs(oracle, golden):
    g = golden("reads")
    bwt = oracle.build_bwt(g["text"])
    assert np.array_equal(bwt, g["bwt"])
    assert oracle.fmd_encode(*oracle.plain2runs(bwt)) == bytes(g["fmd"])
    g = golden("long_runs")
    assert oracle.fmd_encode(g["sym"], g["len"]) == bytes(g["fmd"])  # 32-bit block headers
    ok, ret = oracle.rank1a(g["sym"], g["len"], g["q_k"])
    assert np.array_equal(ok, g["q_ok"]) and np.array_equal(ret, g["q_ret"])
    assert np.array_equal(ret, g["q_ret_fmd"])
    s, l, _ = oracle.fmd_decode(bytes(g["fmd"]))
    assert np.array_equal(s, g["sym"]) and np.array_equal(l, g["len"])


def test_edge_cases(oracle):
    # empty / single-symbol / query past the end
    s, l = oracle.plain2runs(np.zeros(0, np.uint8))
    assert len(s) == 0
    s, l = oracle.plain2runs(np.array([3], np.uint8))
    ok, ret = oracle.rank1a(s, l, [0, 1, 2])
    assert ret.tolist() == [3, -1, -1] and ok[1, 3] == 1
    with pytest.raises(ValueError):
        oracle.mg_rank_plain(s, l, np.array([0, 6], np.uint8))
    # FMR writer obeys the geometry invariants of SURVEY A.2
    rng = np.random.default_rng(3)
    from ropebwt3_b200 import synth
    sym, ln = synth.random_runs(rng, 3000, 300)
    img = oracle.fmr_encode(sym, ln, 16, 128)
    s2, l2, rc, geom = oracle.fmr_decode(img)
    assert geom == (0, 16, 128)
    assert np.array_equal(np.concatenate(oracle.coalesce(s2, l2)), np.concatenate(oracle.coalesce(sym, ln)))


def test_live_reference_differential(oracle):
    """Random inputs through the reference library itself (skipped where oracle/_ref is absent)."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    from ropebwt3_b200 import synth
    for seed in (1, 2):
        gs = synth.genomes(4, 2000, seed=seed, sub=0.01, indel=0.001)
        rope = sym = ln = None
        for i, g_ in enumerate(gs):
            text = synth.batch_text([g_])
            bwt = ref.build_sais(text, 2)
            assert np.array_equal(oracle.build_bwt(text), bwt)
            if rope is None:
                rope = ref.Rope.from_plain(bwt)
                sym, ln = oracle.plain2runs(bwt)
            else:
                rb_ref, _ = rope.mg_rank_plain(bwt)
                rb, _ = oracle.mg_rank_plain(sym, ln, bwt)
                assert np.array_equal(rb, rb_ref)
                rope.merge_plain(bwt)
                sym, ln = oracle.merge_runs(sym, ln, rb)
        # our FMR image is loadable by the reference and converts to the same canonical FMD
        import os
        import tempfile
        with tempfile.NamedTemporaryFile(suffix=".fmr", delete=False) as t:
            t.write(oracle.fmr_encode(sym, ln))
        r2 = ref.Rope.from_file(t.name)
        os.unlink(t.name)
        fmd_ref = rope.to_fmd()
        assert r2.to_fmd() == fmd_ref == oracle.fmd_encode(sym, ln)


# ---------------------------------------------------------------- ropebwt2 insertion (build -2/-s/-r)

def _rb2_batches(oracle, g):
    lines = bytes(g["lines"]).decode().split()
    b = [int(x) for x in g["batch_bounds"]]
    return lines, [oracle.encode_batch(lines[b[i]:b[i + 1]]) for i in range(len(b) - 1)]


@pytest.mark.parametrize("so", [0, 1, 2])
def test_rb2_insert_multi_against_golden(oracle, golden, so):
    """mr_insert_multi restated (BCR rounds on plain lists) == the reference CLI's `build -2/-s/-r`, for the first batch
    and for the whole multi-batch build; and == the closed form (one sort) the CUDA path uses."""
    g = golden("rb2")
    lines, batches = _rb2_batches(oracle, g)
    ropes = [[] for _ in range(6)]
    oracle.insert_multi(ropes, batches[0], so)
    assert np.array_equal(np.array([c for r in ropes for c in r], np.uint8), g["first_so%d" % so])
    assert np.array_equal(oracle.sorted_bwt(batches[0], so), g["first_so%d" % so])
    for t in batches[1:]:
        oracle.insert_multi(ropes, t, so)
    final = np.array([c for r in ropes for c in r], np.uint8)
    assert np.array_equal(final, g["bwt_so%d" % so])
    assert np.array_equal(oracle.sorted_bwt(oracle.encode_batch(lines), so), g["bwt_so%d" % so])
    sym, ln = oracle.plain2runs(final)
    assert oracle.fmd_encode(sym, ln) == bytes(g["fmd_so%d" % so])


def test_rb2_toy_known_answers(oracle, golden):
    g = golden("toy")
    seqs = txt(g["agg_Ls_in"]).split()
    for key, so in [("agg_L", 0), ("agg_Ls", 1), ("agg_Lr", 2)]:
        assert oracle.to_ascii(oracle.rb2_bwt([oracle.encode_batch(seqs)], so)) == txt(g[key + "_out"]).strip()
    assert txt(g["agg_Lr_out"]) == "TTGC$$G$GCGA$ACC\n" and txt(g["agg_Ls_out"]) == "CGTT$$G$CGGA$ACC\n"


def test_ssa_against_golden(oracle, golden):
    """ssa_gen1 / rb3_ssa_dump restated (ssa.c:17-81,198-213) == `ropebwt3 ssa -s SS` of the reference, byte for byte."""
    g = golden("ssa")
    for name, key in [("merge_small", "fmd"), ("rb2", "fmd_so2")]:
        s, l, _ = oracle.fmd_decode(bytes(golden(name)[key]))
        bwt = oracle.runs2plain(s, l)
        for ss in (0, 3, 8):
            assert oracle.ssa_image(bwt, ss) == bytes(g["%s_ss%d" % (name, ss)]), (name, ss)


def test_rb2_closed_form_fuzz_against_live_reference(oracle):
    """Random tiny read sets (small alphabets, duplicates, reads that are suffixes of others, Ns, one or both strands):
    the closed form of the RLO/RCLO order that the CUDA path sorts by == the reference CLI's `build -s/-r/-2`."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(2024)
    for it in range(40):
        n, alpha = int(rng.integers(1, 12)), int(rng.integers(1, 6))
        reads = ["".join("ACGTN"[int(x)] for x in rng.integers(0, alpha, int(rng.integers(1, 9)))) for _ in range(n)]
        if rng.random() < 0.5 and n > 1:
            reads.append(reads[0])
        if rng.random() < 0.3:
            reads.append(reads[0][1:] or "A")
        inp = ("\n".join(reads) + "\n").encode()
        both = rng.random() < 0.5
        for so, flag in [(1, "-s"), (2, "-r"), (0, "-2")]:
            want = ref.run(["build", "-L", flag] + ([] if both else ["-R"]) + ["-"], stdin=inp).decode().strip()
            assert oracle.to_ascii(oracle.sorted_bwt(oracle.encode_batch(reads, rev=both), so)) == want, (flag, both, reads)
