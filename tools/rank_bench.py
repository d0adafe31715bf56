#!/usr/bin/env python
"""Micro-benchmark of the batched rank kernel on an index much larger than L2:
independent uniformly random (k, c) queries, 144 algorithmic bytes each (one 128-B index
block + 8 B position in + 8 B result out).  Prints one JSON line per variant."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ropebwt3_b200 as R  # noqa: E402
from ropebwt3_b200 import capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=float, default=16e6, help="index size in 128-B blocks")
    ap.add_argument("--queries", type=float, default=64e6)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--max-len", type=int, default=30)
    ap.add_argument("--variants", default="8,4,2,42,22,1")
    ap.add_argument("--kind", default="rle", choices=["rle", "bitmap"])
    a = ap.parse_args()
    R.init(0)
    R.set_param("index_kind", 1 if a.kind == "rle" else 2)
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    capi.check(capi.lib().rb3b_set_stream(st.cuda_stream))
    g = torch.Generator(device="cuda").manual_seed(1)
    n_runs = int(a.blocks) * 48
    sym = torch.randint(0, 6, (n_runs,), device="cuda", dtype=torch.uint8, generator=g)
    same = sym[1:] == sym[:-1]
    sym[1:][same] = (sym[1:][same] + 1) % 6   # mostly de-duplicated neighbours; leftovers are legal (not coalesced)
    ln = torch.randint(1, a.max_len + 1, (n_runs,), device="cuda", dtype=torch.int64, generator=g)
    idx = R.Index()
    torch.cuda.synchronize()
    capi.check(capi.lib().rb3b_index_from_runs_device(idx.h, n_runs, sym.data_ptr(), ln.data_ptr()))
    R.sync()
    n = len(idx)
    del sym, ln
    torch.cuda.empty_cache()
    nq = int(a.queries)
    k = torch.randint(0, n, (nq,), device="cuda", dtype=torch.int64, generator=g)
    c = torch.randint(0, 6, (nq,), device="cuda", dtype=torch.uint8, generator=g)
    out = torch.empty(nq, dtype=torch.int64, device="cuda")
    peak = 6541.8
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    ref = None
    for v in ([int(x) for x in a.variants.split(",")] if a.kind == "rle" else [0]):
        for _ in range(3):
            idx.lf_dev(nq, k.data_ptr(), c.data_ptr(), out.data_ptr(), v)
        torch.cuda.synchronize()
        ts = []
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            idx.lf_dev(nq, k.data_ptr(), c.data_ptr(), out.data_ptr(), v)
            e1.record(st)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        if ref is None:
            ref = out.clone()
        else:
            assert torch.equal(ref, out), "variants disagree"
        ms = float(np.median(ts))
        gbs = nq * 144 / (ms / 1e3) / 1e9
        print(json.dumps({"kernel": "k_lf_bm" if a.kind == "bitmap" else "k_lf_tma" if v == 1 else ("k_lf<%d,1>" % v if v < 10 else "k_lf<%d,%d>" % (v // 10, v % 10)), "variant": v, "index_bytes": idx.nbytes(), "index_symbols": n, "queries": nq,
                          "ms": ms, "gqueries_per_s": nq / ms / 1e6, "achieved_GBps": gbs, "peak_GBps": peak, "frac": gbs / peak,
                          "bytes_per_query": 144, "index_kind": a.kind}))


if __name__ == "__main__":
    main()
