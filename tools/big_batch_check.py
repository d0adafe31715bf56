#!/usr/bin/env python
"""Larger-batch sanity/perf check: merge one batch of G genomes (2*G chains, G*10^7 symbols) into an index of H genomes,
for both device layouts; verifies the result against a from-scratch single-batch build of all genomes (canonical runs)."""
import sys, os, time, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ropebwt3_b200 as R
from ropebwt3_b200 import synth, capi

def main():
    H, G, L = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    R.init(0)
    gs = synth.genomes(H + G, L, seed=51)
    L_ = capi.lib()
    def bwt_of(part):
        t = synth.batch_text(part)
        d = torch.from_numpy(t).cuda(); o = torch.empty_like(d); torch.cuda.synchronize()
        t0 = time.time(); capi.check(L_.rb3b_build_bwt_dev(len(t), d.data_ptr(), o.data_ptr())); R.sync()
        return o, len(t), time.time() - t0
    for kind in (2, 1):
        R.set_param("index_kind", kind)
        b0, n0, t_b0 = bwt_of(gs[:H])
        b1, n1, t_b1 = bwt_of(gs[H:])
        idx = R.Index.from_plain_dev(b0.data_ptr(), n0)
        R.sync(); R.get_stat("reset")
        t0 = time.time(); idx.merge_plain_dev(b1.data_ptr(), n1); R.sync(); dt = time.time() - t0
        st = {k: R.get_stat(k) for k in ["us_prep", "us_walk_first", "us_walk_fix", "us_merge", "n_segments", "fix_rounds", "n_cells", "n_ovf_cells", "cell_shift", "sa_rounds"]}
        ball, nall, t_ball = bwt_of(gs)
        ref = R.Index.from_plain_dev(ball.data_ptr(), nall)
        s1, l1 = idx.export_runs(); s2, l2 = ref.export_runs()
        ok = np.array_equal(s1, s2) and np.array_equal(l1, l2)
        print(json.dumps({"kind": "bitmap" if kind == 2 else "rle", "index_genomes": H, "batch_genomes": G, "genome_len": L, "batch_symbols": n1,
                          "merge_s": dt, "batch_bases_per_s": G * L / dt, "bwt_build_s": [t_b0, t_b1, t_ball], "index_bytes": idx.nbytes(),
                          "runs": int(len(s1)), "identical_to_single_batch_build": bool(ok), "stats": st}))
        assert ok
        del idx, ref
        torch.cuda.empty_cache()

if __name__ == "__main__":
    main()
