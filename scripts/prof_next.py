"""Small fixed workloads for ncu: the BPLA kernel (600 sequences of 100 nt) and the naive stem kernel (60 C1 records)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stem_kernel_b200 import synth, bpla, nstem, api, _lib as L
ctx = api.Context(L.make_params(L.STR_SIMPLE))
rng = np.random.default_rng(2)
recs = []
for i in range(600):
    a, b = rng.uniform(0, 0.6, 100), rng.uniform(0, 0.4, 100)
    recs.append(dict(rows=["".join(rng.choice(list("acgu"), 100))], p_left=np.sqrt(a), p_right=np.sqrt(b), p_unpair=np.sqrt(np.maximum(0, 1 - a - b))))
s = bpla.BplaSet(recs)
xi, yi = np.triu_indices(len(s))
for sw in (0, 1):
    bpla.pairs(ctx, bpla.make_params(sw=sw), s, s, xi, yi)
c1 = synth.make_config(1, 60)
seqs = [r["rows"][0].lower() for r in c1]
ns = nstem.NstemSet(seqs, [nstem.dense_bp(len(q), r["bp"][0], th=0.01) for q, r in zip(seqs, c1)])
xi, yi = np.triu_indices(len(ns))
nstem.pairs(ctx, nstem.make_params(bp_mode=1, bp_bound=0.01), ns, ns, xi, yi)
print("done")
