"""The C host CLI (cli/ropebwt3-b200 build) against the reference CLI (oracle/_ref/ropebwt3 build) on the same files:
.fmd byte-identical, plain output identical, .fmr loadable/extendable by the reference."""
import gzip
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "cli", "ropebwt3-b200")


def run(binary, args, stdin=None, check=True):
    p = subprocess.run([binary] + [str(a) for a in args], input=stdin, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    if check and p.returncode != 0:
        raise RuntimeError("%s %s failed (%d): %s" % (binary, args, p.returncode, p.stderr.decode()[-800:]))
    return p


@pytest.fixture(scope="module")
def files(tmp_path_factory, oracle):
    from oracle import ref
    if not os.path.exists(CLI):
        pytest.fail("cli/ropebwt3-b200 is not built (run __graft_entry__.build())")
    if not os.path.exists(ref.BIN):
        pytest.skip("oracle/_ref/ropebwt3 did not travel")
    from ropebwt3_b200 import synth
    d = tmp_path_factory.mktemp("cli")
    gs = synth.genomes(6, 30000, seed=77)
    gs[2][100:140] = 5  # a stretch of Ns
    fa = []
    for i, g in enumerate(gs):
        s = oracle.to_ascii(g)
        body = ">g%d some description\n" % i + "\n".join(s[k:k + 80] for k in range(0, len(s), 80)) + "\n"
        fn = str(d / ("g%d.fa" % i))
        if i == 1:
            fn += ".gz"
            with gzip.open(fn, "wb") as f:
                f.write(body.encode())
        else:
            open(fn, "w").write(body)
        fa.append(fn)
    multi = str(d / "all.fa")
    with open(multi, "w") as f:
        for i, g in enumerate(gs):
            f.write(">s%d\n%s\n" % (i, oracle.to_ascii(g).lower() if i == 3 else oracle.to_ascii(g)))
    lines = str(d / "lines.txt")
    with open(lines, "w") as f:
        for g in gs:
            f.write(oracle.to_ascii(g[:5000]) + "\n")
    fq = str(d / "reads.fq")
    rng = np.random.default_rng(3)
    with open(fq, "w") as f:
        for i in range(200):
            s0 = int(rng.integers(0, 29000))
            r = oracle.to_ascii(gs[0][s0:s0 + 150])
            f.write("@r%d\n%s\n+\n%s\n" % (i, r, "I" * len(r)))
    return {"dir": d, "fa": fa, "multi": multi, "lines": lines, "fq": fq, "ref": ref.BIN}


def test_multi_file_build_fmd(files):
    want = run(files["ref"], ["build", "-t4", "-d"] + files["fa"]).stdout
    got = run(CLI, ["build", "-t4", "-d"] + files["fa"]).stdout
    assert got == want and len(got) > 1000


def test_multi_device_build(files, tmp_path):
    """RB3B_DEVICES=0,1[,2,3]: one host thread and one replica of the index per device, every merge through
    rb3b_merge_plain_dist; the output must be the reference's, also when appending to an existing .fmr (-i)."""
    import torch
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, RB3B_DEVICES=",".join(str(i) for i in range(n)))

    def run_md(args):
        p = subprocess.run([CLI] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=600)
        assert p.returncode == 0, p.stderr.decode()[-800:]
        assert ("%d devices" % n).encode() in p.stderr
        return p.stdout

    want = run(files["ref"], ["build", "-t4", "-d"] + files["fa"]).stdout
    assert run_md(["build", "-d"] + files["fa"]) == want
    assert run_md(["build", "-d", "-m", "50k", files["multi"]]) == run(files["ref"], ["build", "-d", "-m", "50k", files["multi"]]).stdout
    first = str(tmp_path / "first.fmr")
    run(CLI, ["build", "-b", "-o", first] + files["fa"][:3])
    assert run_md(["build", "-d", "-i", first] + files["fa"][3:]) == want


@pytest.mark.parametrize("opts", [["-m", "50k"], ["-m", "7g"], ["-R"], ["-F"], ["-m", "100k", "-R"]])
def test_single_file_batches(files, opts):
    want = run(files["ref"], ["build", "-d"] + opts + [files["multi"]]).stdout
    got = run(CLI, ["build", "-d"] + opts + [files["multi"]]).stdout
    assert got == want


def test_line_input_stdin_fastq_and_plain_output(files):
    data = open(files["lines"], "rb").read()
    assert run(CLI, ["build", "-L", "-"], stdin=data).stdout == run(files["ref"], ["build", "-L", "-"], stdin=data).stdout
    assert run(CLI, ["build", "-Ld", files["lines"]]).stdout == run(files["ref"], ["build", "-Ld", files["lines"]]).stdout
    assert run(CLI, ["build", "-d", files["fq"]]).stdout == run(files["ref"], ["build", "-d", files["fq"]]).stdout
    toy = b"AGG\nAGC\n"
    assert run(CLI, ["build", "-L", "-"], stdin=toy).stdout == b"GTCT$$G$CGGA$ACC\n"
    assert run(CLI, ["build", "-LR", "-"], stdin=toy).stdout == b"GC$$GGAA\n"


def test_incremental_append_and_checkpoint(files):
    d = files["dir"]
    half, ckpt, out = str(d / "half.fmr"), str(d / "ckpt.fmr"), str(d / "full.fmd")
    run(CLI, ["build", "-b", "-o", half] + files["fa"][:3])
    run(CLI, ["build", "-d", "-i", half, "-S", ckpt, "-o", out] + files["fa"][3:])
    want = run(files["ref"], ["build", "-d"] + files["fa"]).stdout
    assert open(out, "rb").read() == want
    # the checkpoint is a valid FMR of the whole collection; the reference converts it to the same FMD
    assert run(files["ref"], ["build", "-d", "-i", ckpt]).stdout == want
    # and the reference can keep inserting into our half-way FMR
    assert run(files["ref"], ["build", "-d", "-i", half] + files["fa"][3:]).stdout == want
    # appending to a reference-made .fmd works as well (rb3_fmi_restore sniffs both formats)
    ref_half = str(d / "ref_half.fmd")
    open(ref_half, "wb").write(run(files["ref"], ["build", "-d"] + files["fa"][:3]).stdout)
    assert run(CLI, ["build", "-d", "-i", ref_half] + files["fa"][3:]).stdout == want
    # pure format conversion: no input files, only -i (build.c:169)
    assert run(CLI, ["build", "-d", "-i", half]).stdout == run(files["ref"], ["build", "-d"] + files["fa"][:3]).stdout


def test_error_behaviour(files):
    assert run(CLI, ["build", "-d", "-i", "/nonexistent.fmr", files["multi"]], check=False).returncode == 1   # build.c:175-178
    p = run(CLI, ["build", "-d", "/nonexistent.fa", files["lines"] + ".nope"], check=False)                  # build.c:207-210,243
    assert p.returncode == 1 and b"failed to open file" in p.stderr
    p = run(CLI, ["build", "-d", "/nonexistent.fa"] + files["fa"][:1], check=False)                           # bad file skipped, rest built
    assert p.returncode == 0 and p.stdout == run(files["ref"], ["build", "-d"] + files["fa"][:1]).stdout
    assert run(CLI, ["build", "-T", files["multi"]], check=False).returncode == 1                             # documented scope limit
    assert run(CLI, ["build"], check=False).returncode == 1


@pytest.mark.parametrize("flag", ["-2", "-s", "-r"])
def test_ropebwt2_insertion_orders(files, flag):
    """build -2/-s/-r (mr_insert_multi): reads in one batch and in many (-m), FASTQ, both strands, and
    appending to an index written in that order; plain output and .fmd equal the reference CLI's."""
    d = files["dir"]
    assert run(CLI, ["build", flag, files["fq"]]).stdout == run(files["ref"], ["build", flag, files["fq"]]).stdout
    want = run(files["ref"], ["build", flag, "-d", "-m", "9k", files["fq"]]).stdout
    assert run(CLI, ["build", flag, "-d", "-m", "9k", files["fq"]]).stdout == want
    assert run(CLI, ["build", flag, "-d", files["fq"]]).stdout == want
    toy = b"AGG\nAGC\n"
    assert run(CLI, ["build", "-L", flag, "-"], stdin=toy).stdout == {"-2": b"GTCT$$G$CGGA$ACC\n", "-s": b"CGTT$$G$CGGA$ACC\n", "-r": b"TTGC$$G$GCGA$ACC\n"}[flag]
    # -i keeps the order recorded in the .fmr (mr->so); the reference can continue from our file and we from its
    mine, theirs = str(d / ("rb2%s.fmr" % flag)), str(d / ("rb2%s.ref.fmr" % flag))
    run(CLI, ["build", flag, "-b", "-o", mine, files["fq"]])
    open(theirs, "wb").write(run(files["ref"], ["build", flag, "-b", files["fq"]]).stdout)
    more = files["fa"][0]
    want2 = run(files["ref"], ["build", flag, "-d", files["fq"], more]).stdout
    assert run(CLI, ["build", flag, "-d", "-i", mine, more]).stdout == want2
    assert run(CLI, ["build", flag, "-d", "-i", theirs, more]).stdout == want2
    assert run(files["ref"], ["build", flag, "-d", "-i", mine, more]).stdout == want2


def test_config0_reads_rclo(files, oracle):
    """BASELINE.json configs[0]: `build -r` of 10k synthetic 150 bp reads (0.5% errors incl. N) -- here against the
    reference binary at the full size."""
    d = files["dir"]
    rng = np.random.default_rng(42)
    anc = rng.integers(1, 5, 100000).astype(np.uint8)
    fn = str(d / "c0.txt")
    with open(fn, "w") as f:
        for _ in range(10000):
            s0 = int(rng.integers(0, len(anc) - 150))
            r = anc[s0:s0 + 150].copy()
            m = rng.random(150) < 0.005
            r[m] = rng.integers(1, 6, int(m.sum()))
            f.write(oracle.to_ascii(r) + "\n")
    want = run(files["ref"], ["build", "-r", "-L", "-t1", "-d", fn]).stdout
    assert run(CLI, ["build", "-r", "-L", "-t1", "-d", fn]).stdout == want
    assert run(CLI, ["build", "-r", "-L", "-d", "-m", "500k", fn]).stdout == want


def test_ssa_subcommand(files):
    """`ropebwt3-b200 ssa` == `ropebwt3 ssa` on an index built by either tool."""
    d = files["dir"]
    fmd = str(d / "ssa_in.fmd")
    open(fmd, "wb").write(run(CLI, ["build", "-d"] + files["fa"][:3]).stdout)
    for ss in ("8", "4"):
        assert run(CLI, ["ssa", "-s", ss, fmd]).stdout == run(files["ref"], ["ssa", "-s", ss, "-t4", fmd]).stdout
    out = str(d / "x.ssa")
    run(CLI, ["ssa", "-o", out, fmd])
    assert open(out, "rb").read() == run(files["ref"], ["ssa", fmd]).stdout


def test_merge_subcommand(files):
    """`ropebwt3 merge` (main.c:84-133): base.fmr + other indexes (.fmd and .fmr) -> FMR on stdout / -o, -S checkpoint after
    every input; the merged collection must be the reference's, compared through the canonical .fmd."""
    from oracle import ref
    d = files["dir"]
    a, b, c = str(d / "ma.fmr"), str(d / "mb.fmd"), str(d / "mc.fmr")
    open(a, "wb").write(run(ref.BIN, ["build", "-b", files["fa"][0], files["fa"][1]]).stdout)
    open(b, "wb").write(run(ref.BIN, ["build", "-d", files["fa"][2]]).stdout)
    open(c, "wb").write(run(ref.BIN, ["build", "-b", files["fa"][3], files["fa"][4]]).stdout)
    want_fmr = str(d / "want.fmr")
    open(want_fmr, "wb").write(run(ref.BIN, ["merge", a, b, c]).stdout)
    want = run(ref.BIN, ["build", "-i", want_fmr, "-d"]).stdout
    mine, ckpt = str(d / "mine.fmr"), str(d / "ckpt.fmr")
    run(CLI, ["merge", "-t", "4", "-o", mine, "-S", ckpt, a, b, c])
    assert run(ref.BIN, ["build", "-i", mine, "-d"]).stdout == want      # the reference reads our merged FMR
    assert run(CLI, ["build", "-i", ckpt, "-d"]).stdout == want          # the checkpoint after the last input is the result
    assert run(CLI, ["merge", a], check=False).returncode == 1           # usage: needs two indexes
    assert run(CLI, ["merge", str(d / "nope.fmr"), b], check=False).returncode == 1
