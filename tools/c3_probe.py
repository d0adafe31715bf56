#!/usr/bin/env python
"""C3/C4-like probe (BASELINE configs[3]/[4] scaled to what can be synthesised here): merging one 5 Mb genome per step
into a run-length index of ~10^11 symbols.  The index is synthetic (random runs of mean length 1000, the shape of a highly
repetitive collection: 0.8 GB of cells), so there is no reference to compare with -- the probe reports times and checks the
size-independent properties (symbol totals add up, the index keeps its layout); parity of the same kernels is pinned at
small and medium scale by the tests.   usage: tools/c3_probe.py [n_merges]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ropebwt3_b200 as R  # noqa: E402
from ropebwt3_b200 import capi, synth  # noqa: E402


def main():
    n_merges = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    R.init(0)
    R.set_param("index_kind", 1)   # run-length cells
    g = torch.Generator(device="cuda").manual_seed(1)
    n_runs = int(2.1e6) * 48
    sym = torch.randint(0, 6, (n_runs,), device="cuda", dtype=torch.uint8, generator=g)
    same = sym[1:] == sym[:-1]
    sym[1:][same] = (sym[1:][same] + 1) % 6
    ln = torch.randint(1, 2001, (n_runs,), device="cuda", dtype=torch.int64, generator=g)
    idx = R.Index()
    torch.cuda.synchronize()
    capi.check(capi.lib().rb3b_index_from_runs_device(idx.h, n_runs, sym.data_ptr(), ln.data_ptr()))
    R.sync()
    del sym, ln
    torch.cuda.empty_cache()
    n0, acc0 = len(idx), idx.acc()
    gs = synth.genomes(n_merges, 5_000_000, seed=43)
    out = {"index_symbols": int(n0), "index_bytes": int(idx.nbytes()), "cell_shift": int(R.get_stat("cell_shift")), "merges": []}
    expect = np.diff(acc0).astype(np.int64)
    for gi, gen in enumerate(gs):
        text = synth.batch_text([gen])
        t = torch.from_numpy(text).cuda()
        b = torch.empty_like(t)
        capi.check(capi.lib().rb3b_build_bwt_dev(len(t), t.data_ptr(), b.data_ptr()))
        R.sync()
        R.get_stat("reset")
        torch.cuda.synchronize()
        t0 = time.time()
        idx.merge_plain_dev(b.data_ptr(), len(t))
        R.sync()
        ms = (time.time() - t0) * 1e3
        cnt = np.bincount(gen, minlength=6)[:6]
        expect += cnt + cnt[[0, 4, 3, 2, 1, 5]]
        expect[0] += 2
        assert np.array_equal(np.diff(idx.acc()), expect), "symbol totals of the merged index are wrong"
        out["merges"].append({"ms": ms, "bases_per_s": 5e6 / (ms / 1e3), "index_bytes_after": int(idx.nbytes()),
                              "phases_us": {k: R.get_stat(k) for k in ["us_prep", "us_walk_first", "us_walk_fix", "us_scatter", "us_merge"]},
                              "fix_rounds": R.get_stat("fix_rounds"), "index_kind": R.get_stat("index_kind"), "cell_shift": R.get_stat("cell_shift")})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
