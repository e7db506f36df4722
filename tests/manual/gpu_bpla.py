"""BPLA kernel timing: C2-like set (n random 100-nt sequences) and C1-like records; error against the oracle on a sample."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from stem_kernel_b200 import synth, bpla, api, _lib as L
from oracle import oraclebind as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
ctx = api.Context(L.make_params(L.STR_SIMPLE))
rng = np.random.default_rng(2)
recs = []
for i in range(n):
    a, b = rng.uniform(0, 0.6, 100), rng.uniform(0, 0.4, 100)
    recs.append(dict(rows=["".join(rng.choice(list("acgu"), 100))], p_left=np.sqrt(a), p_right=np.sqrt(b), p_unpair=np.sqrt(np.maximum(0, 1 - a - b))))
s = bpla.BplaSet(recs)
xi, yi = np.triu_indices(n)
for no_bp, sw in ((0, 0), (1, 0), (0, 1)):
    p = bpla.make_params(no_bp=no_bp, sw=sw)
    bpla.pairs(ctx, p, s, s, xi[:1000], yi[:1000])
    t = time.perf_counter(); v = bpla.pairs(ctx, p, s, s, xi, yi); dt = time.perf_counter() - t
    k = rng.choice(len(xi), 300, replace=False)
    w = O.bpla_pairs(p, s, s, xi[k], yi[k])
    print(f"BPLA no_bp={no_bp} sw={sw}: {len(xi)} pairs of 100x100 in {dt*1e3:.0f} ms wall (host buffers): {len(xi)/dt/1e6:.2f} M pairs/s, "
          f"{len(xi)*1e4/dt/1e9:.0f} GCUPS; max rel err vs oracle {np.max(np.abs(v[k]-w)/np.abs(w)):.2e}", flush=True)
