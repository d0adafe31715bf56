"""per-call wall time of rb3b_merge_plain_dev with the asynchronous merge on (debug aid)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ropebwt3_b200 as R
from ropebwt3_b200 import synth, capi
R.init(0)
use_torch_stream = False
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    if k == "torch_stream":
        use_torch_stream = bool(int(v))
    else:
        R.set_param(k, int(v))
if use_torch_stream:
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    capi.check(capi.lib().rb3b_set_stream(stream.cuda_stream))
gs = synth.genomes(14, 5_000_000, seed=43)
d = []
for g in gs:
    t = torch.from_numpy(synth.batch_text([g])).cuda()
    o = torch.empty_like(t)
    capi.check(capi.lib().rb3b_build_bwt_dev(len(t), t.data_ptr(), o.data_ptr()))
    d.append(o)
R.sync()
idx = R.Index.from_plain_dev(d[0].data_ptr(), len(d[0]))
idx.reserve(sum(len(x) for x in d))
for i in range(1, 7):
    t0 = time.time()
    idx.merge_plain_dev(d[i].data_ptr(), len(d[i]))
    t1 = time.time()
    R.sync()
    t2 = time.time()
    print("merge %d: call %.3f ms, sync %.3f ms" % (i, (t1 - t0) * 1e3, (t2 - t1) * 1e3), {k: R.get_stat(k) for k in ["us_prep", "us_merge"]})
t0 = time.time()
tt = []
for i in range(7, len(d)):
    t1 = time.time()
    idx.merge_plain_dev(d[i].data_ptr(), len(d[i]))
    tt.append((time.time() - t1) * 1e3)
R.sync()
print("back to back: %.3f ms per merge" % ((time.time() - t0) * 1e3 / (len(d) - 7)), ["%.2f" % x for x in tt])
