#!/usr/bin/env python
"""Sampled suffix array: rb3b_ssa_gen on the device vs `ropebwt3 ssa -t$(nproc)` of the reference, on the
same index (N synthetic genomes merged on the device, written as .fmd).  Prints one JSON line."""
import json, os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import ropebwt3_b200 as R
    from ropebwt3_b200 import synth, capi
    n, L, ss = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 8
    R.init(0)
    gs = synth.genomes(n, L, seed=43)
    idx = None
    for g in gs:
        bwt = R.rb3_build_sais(synth.batch_text([g]))
        if idx is None:
            idx = R.Index.from_plain(bwt)
        else:
            idx.merge_plain(bwt)
    d = tempfile.mkdtemp(prefix="rb3b_ssa_")
    fmd = os.path.join(d, "x.fmd")
    idx.dump_fmd(fmd)
    import ctypes as C
    m, n_ssa, ms = C.c_int64(), C.c_int64(), C.c_int()
    capi.check(capi.lib().rb3b_ssa_sizes(idx.h, ss, C.byref(m), C.byref(n_ssa), C.byref(ms)))
    buf = torch.empty(m.value + n_ssa.value, dtype=torch.int64, device="cuda")
    ts = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.time()
        capi.check(capi.lib().rb3b_ssa_gen_dev(idx.h, ss, buf.data_ptr(), buf.data_ptr() + 8 * m.value))
        R.sync()
        ts.append(time.time() - t0)
    out = {"symbols": len(idx), "strings": m.value, "ssa_shift": ss, "device_ssa_gen_s": min(ts), "device_symbols_per_s": len(idx) / min(ts)}
    mine = os.path.join(d, "mine.ssa")
    t0 = time.time()
    idx.ssa_dump(mine, ss)
    out["device_gen_and_dump_s"] = time.time() - t0
    ref = os.path.join(ROOT, "oracle", "_ref", "ropebwt3")
    if os.path.exists(ref):
        t0 = time.time()
        theirs = subprocess.run([ref, "ssa", "-s", str(ss), "-t", str(os.cpu_count() or 1), fmd], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
        out["reference_ssa_s"] = time.time() - t0
        out["cores"] = os.cpu_count()
        out["identical"] = theirs == open(mine, "rb").read()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
