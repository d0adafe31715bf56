python - <<'PY'
import os, sys, numpy as np
sys.path.insert(0, ".")
from ropebwt3_b200 import synth
gs = synth.genomes(6, 5000000, seed=43)
lut = np.frombuffer(b"$ACGTN", np.uint8)
os.makedirs("/tmp/fa", exist_ok=True)
for i, g in enumerate(gs):
    s = lut[g]; full = len(s)//80*80
    with open("/tmp/fa/g%d.fa" % i, "wb") as f:
        f.write(b">g%d\n" % i); f.write(b"\n".join(r.tobytes() for r in s[:full].reshape(-1,80))); f.write(b"\n" + s[full:].tobytes() + b"\n")
PY
for k in 1 2; do cli/ropebwt3-b200 build -d -o /tmp/fa/out.fmd /tmp/fa/g*.fa 2> /tmp/fa/log.txt; grep -E "read|constructed|merged|encoded|Real" /tmp/fa/log.txt | sed -n '1,6p;$p' | cut -c1-90; done
