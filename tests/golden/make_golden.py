"""Regenerates tests/golden/*.npz with the UNMODIFIED reference built into oracle/_ref/
(make -C oracle ref).  Run from the repo root in the build container:

    python tests/golden/make_golden.py

Every array below is an output of reference code (the ropebwt3 binary or
librb3ref.so); the inputs are seeded and stored next to the outputs so the
fixtures are self-contained on the GPU box, where /root/reference is absent.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402  (text encoding helpers only)
from oracle import ref as R  # noqa: E402
from ropebwt3_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def u8(b):
    return np.frombuffer(b, np.uint8)


def toy():
    """Literal known answers, captured from the reference CLI."""
    d = {}
    cases = {
        "agg_LR": (["build", "-LR", "-"], b"AGG\nAGC\n"),
        "agg_L": (["build", "-L", "-"], b"AGG\nAGC\n"),
        "agg_LT": (["build", "-LT", "-"], b"AGG\nAGC\n"),
        "agg_Lr": (["build", "-Lr", "-"], b"AGG\nAGC\n"),
        "agg_Ls": (["build", "-Ls", "-"], b"AGG\nAGC\n"),
        "long_L": (["build", "-L", "-"], b"TGAACTCTACACAACATATTTTGTCACCAAG\n"),
        "long_Ld": (["build", "-Ld", "-"], b"TGAACTCTACACAACATATTTTGTCACCAAG\n"),
        "long_Lb": (["build", "-Lb", "-"], b"TGAACTCTACACAACATATTTTGTCACCAAG\n"),
        "nn_L": (["build", "-L", "-"], b"ACNNGT\n"),
    }
    for k, (args, inp) in cases.items():
        d[k + "_in"] = u8(inp)
        d[k + "_out"] = u8(R.run(args, stdin=inp))
    np.savez_compressed(os.path.join(OUT, "toy.npz"), **d)


def merge_set(name, n_genomes, length, seed, sub, indel, per_batch=1, n_q=300):
    """A multi-batch build: BWT of every batch (rb3_build_sais), the interleave array of
    every merge (rb3_mg_rank_plain), rank1a answers, and the final FMD/FMR images."""
    gs = synth.genomes(n_genomes, length, seed=seed, sub=sub, indel=indel)
    d = {"n_batches": np.int64((n_genomes + per_batch - 1) // per_batch)}
    rope = None
    rng = np.random.default_rng(seed + 1000)
    for b in range(int(d["n_batches"])):
        part = gs[b * per_batch:(b + 1) * per_batch]
        text = synth.batch_text(part)
        bwt = R.build_sais(text, 2 * len(part), n_threads=1)
        d["text%d" % b] = text
        d["bwt%d" % b] = bwt
        if rope is None:
            rope = R.Rope.from_plain(bwt)
        else:
            rb, acc = rope.mg_rank_plain(bwt)
            d["rb%d" % b] = rb
            d["acc%d" % b] = acc
            rope.merge_plain(bwt)
        d["accA%d" % b] = rope.acc()
    n = int(rope.acc()[6])
    k = np.concatenate([rng.integers(0, n, n_q), [0, n - 1, n, n + 5]]).astype(np.int64)
    ok, ret = rope.rank1a(k)
    d["q_k"], d["q_ok"], d["q_ret"] = k, ok, ret
    d["fmr"] = u8(rope.dump_fmr())
    d["fmd"] = u8(rope.to_fmd())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)


def reads_set():
    """C0-like: reads with errors and Ns, one batch."""
    rng = np.random.default_rng(42)
    anc = rng.integers(1, 5, 5000).astype(np.uint8)
    reads = []
    for _ in range(300):
        s = int(rng.integers(0, len(anc) - 150))
        r = anc[s:s + 150].copy()
        m = rng.random(150) < 0.005
        r[m] = rng.integers(1, 6, int(m.sum()))
        reads.append(r)
    text = synth.batch_text(reads)
    bwt = R.build_sais(text, 2 * len(reads))
    rope = R.Rope.from_plain(bwt)
    np.savez_compressed(os.path.join(OUT, "reads.npz"), text=text, bwt=bwt, fmd=u8(rope.to_fmd()))


def long_runs():
    """FMD with 32-bit block headers: a plain BWT made of long runs."""
    rng = np.random.default_rng(7)
    sym, ln = synth.random_runs(rng, 400, 60)
    idx = np.arange(0, 400, 7)
    ln[idx] = rng.integers(20000, 90000, len(idx))
    plain = np.repeat(sym, ln)
    rope = R.Rope.from_plain(plain)
    k = np.concatenate([rng.integers(0, len(plain), 200), [0, len(plain) - 1, len(plain)]]).astype(np.int64)
    ok, ret = rope.rank1a(k)
    fmr = u8(rope.dump_fmr())
    fmd = rope.to_fmd()
    f = R.Fmd(fmd)
    ok2, ret2 = f.rank1a(k)
    f.close()
    assert np.array_equal(ok, ok2)
    np.savez_compressed(os.path.join(OUT, "long_runs.npz"), sym=sym, len=ln, fmd=u8(fmd), fmr=fmr, q_k=k, q_ok=ok, q_ret=ret, q_ret_fmd=ret2)


def rb2_set():
    """build -2 / -s / -r (ropebwt2 insertion, mr_insert_multi): reads with errors, Ns, exact duplicates and reads that
    are suffixes / prefixes of other reads; one batch and several batches (-m); outputs of the reference CLI."""
    rng = np.random.default_rng(46)
    anc = rng.integers(1, 5, 3000).astype(np.uint8)
    reads = []
    for _ in range(400):
        s = int(rng.integers(0, len(anc) - 60))
        r = anc[s:s + int(rng.integers(20, 61))].copy()
        m = rng.random(len(r)) < 0.01
        r[m] = rng.integers(1, 6, int(m.sum()))
        reads.append(r)
    reads += [reads[3].copy(), reads[3][7:].copy(), reads[11][:-5].copy(), reads[3].copy(), reads[200].copy()]
    lines = [O.to_ascii(r) for r in reads]
    inp = ("\n".join(lines) + "\n").encode()
    batch_m = 6000
    d = {"lines": u8(inp), "batch_m": np.int64(batch_m)}
    # the batches rb3_seq_read forms with -m (io.c:114,119: a batch closes after the record that exceeds -m)
    bounds, n = [0], 0
    for i, l in enumerate(lines):
        n += 2 * len(l) + 2
        if n > batch_m:
            bounds.append(i + 1)
            n = 0
    if bounds[-1] != len(lines):
        bounds.append(len(lines))
    d["batch_bounds"] = np.array(bounds, np.int64)
    first = ("\n".join(lines[:bounds[1]]) + "\n").encode()
    asc = np.zeros(256, np.uint8)
    for i, ch in enumerate(b"$ACGTN"):
        asc[ch] = i
    for so, flag in [(0, "-2"), (1, "-s"), (2, "-r")]:
        one = R.run(["build", "-L", flag, "-"], stdin=inp)
        many = R.run(["build", "-L", flag, "-m", str(batch_m), "-"], stdin=inp)
        assert one == many
        d["bwt_so%d" % so] = asc[u8(one.strip())]
        d["first_so%d" % so] = asc[u8(R.run(["build", "-L", flag, "-"], stdin=first).strip())]
        d["fmd_so%d" % so] = u8(R.run(["build", "-L", "-d", flag, "-m", str(batch_m), "-"], stdin=inp))
    d["fmr_so2"] = u8(R.run(["build", "-L", "-b", "-r", "-"], stdin=inp))
    np.savez_compressed(os.path.join(OUT, "rb2.npz"), **d)


def ssa_set():
    """`ropebwt3 ssa -s SS idx.fmd` of two committed indexes (the merge_small genomes; the RCLO read set)."""
    import tempfile
    d = {}
    for name, key in [("merge_small", "fmd"), ("rb2", "fmd_so2")]:
        g = np.load(os.path.join(OUT, name + ".npz"))
        with tempfile.TemporaryDirectory() as t:
            fn = os.path.join(t, "x.fmd")
            open(fn, "wb").write(bytes(g[key]))
            for ss in (0, 3, 8):
                d["%s_ss%d" % (name, ss)] = u8(R.run(["ssa", "-s", str(ss), "-t", "2", fn]))
    np.savez_compressed(os.path.join(OUT, "ssa.npz"), **d)


if __name__ == "__main__":
    assert R.available(), "build the reference first: make -C oracle ref"
    toy()
    merge_set("merge_small", 6, 3000, 43, 0.005, 0.0005)
    merge_set("merge_div", 5, 4000, 44, 0.02, 0.002, per_batch=2)
    merge_set("merge_dup", 4, 2500, 45, 0.0, 0.0)          # exact duplicates: the walk never collapses
    reads_set()
    long_runs()
    rb2_set()
    ssa_set()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
