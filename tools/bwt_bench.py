"""device suffix sort throughput: python tools/bwt_bench.py [genomes_per_batch ...]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, json
import ropebwt3_b200 as R
from ropebwt3_b200 import synth, capi
R.init(0)
gs = synth.genomes(20, 5_000_000, seed=43)
for G in [int(x) for x in (sys.argv[1:] or ["1", "10"])]:
    for discard, keys_only in ((1, 1), (1, 0), (0, 0)):
        R.set_param("sa_discard", discard)
        R.set_param("sa_keys_only", keys_only)
        text = synth.batch_text(gs[:G])
        t = torch.from_numpy(text).cuda()
        o = torch.empty_like(t)
        ts = []
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.time()
            capi.check(capi.lib().rb3b_build_bwt_dev(len(t), t.data_ptr(), o.data_ptr()))
            R.sync()
            ts.append(time.time() - t0)
        print(json.dumps({"genomes_per_batch": G, "symbols": len(text), "sa_discard": discard, "sa_keys_only": keys_only, "ms": min(ts) * 1e3, "bases_per_s": G * 5e6 / min(ts),
                          "rounds": R.get_stat("sa_rounds"), "ambiguous_after_round0": R.get_stat("sa_ambiguous_after_round0")}))
