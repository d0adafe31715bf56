#!/usr/bin/env python
"""End-to-end CLI comparison on the GPU box: N synthetic genomes as FASTA files ->
`cli/ropebwt3-b200 build -d` vs `oracle/_ref/ropebwt3 build -t$(nproc) -d`; the two .fmd files must be identical."""
import os, subprocess, sys, tempfile, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ropebwt3_b200 import synth

def main():
    n, L = int(sys.argv[1]), int(sys.argv[2])
    gs = synth.genomes(n, L, seed=43)
    lut = np.frombuffer(b"$ACGTN", np.uint8)
    d = tempfile.mkdtemp(prefix="rb3b_cli_")
    files = []
    for i, g in enumerate(gs):
        s = lut[g]
        pad = (-len(s)) % 80
        body = np.concatenate([s, np.full(pad, ord("A"), np.uint8)]).reshape(-1, 80) if False else None
        fn = os.path.join(d, "g%03d.fa" % i)
        with open(fn, "wb") as f:
            f.write(b">g%d\n" % i)
            full = len(s) // 80 * 80
            rows = s[:full].reshape(-1, 80)
            f.write(b"\n".join(r.tobytes() for r in rows)); f.write(b"\n")
            if full < len(s): f.write(s[full:].tobytes() + b"\n")
        files.append(fn)
    out = {}
    t0 = time.time()
    p = subprocess.run([os.path.join(ROOT, "cli", "ropebwt3-b200"), "build", "-d", "-o", os.path.join(d, "mine.fmd")] + files, stderr=subprocess.PIPE)
    out["b200_cli_s"] = time.time() - t0
    assert p.returncode == 0, p.stderr.decode()[-500:]
    if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
        open(os.path.join(ROOT, "gpurun_out", "cli_e2e_stderr.txt"), "wb").write(p.stderr)
    ref = os.path.join(ROOT, "oracle", "_ref", "ropebwt3")
    t0 = time.time()
    p = subprocess.run([ref, "build", "-t%d" % (os.cpu_count() or 1), "-d", "-o", os.path.join(d, "ref.fmd")] + files, stderr=subprocess.PIPE)
    out["reference_cli_s"] = time.time() - t0
    assert p.returncode == 0
    a, b = open(os.path.join(d, "mine.fmd"), "rb").read(), open(os.path.join(d, "ref.fmd"), "rb").read()
    out.update(genomes=n, genome_len=L, fmd_bytes=len(a), identical=a == b, cores=os.cpu_count(),
               bases_per_s_b200=n * L / out["b200_cli_s"], bases_per_s_reference=n * L / out["reference_cli_s"])
    print(json.dumps(out))
    assert a == b

if __name__ == "__main__":
    main()
