"""Naive stem kernel timing: n C1-like records (~75 nt) with probability tables; error against the oracle on a sample."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from stem_kernel_b200 import synth, nstem, api, _lib as L
from oracle import oraclebind as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ctx = api.Context(L.make_params(L.STR_SIMPLE))
recs = synth.make_config(1, n)
seqs = [r["rows"][0].lower() for r in recs]
for name, s, p in (("tables th=0.01", nstem.NstemSet(seqs, [nstem.dense_bp(len(q), r["bp"][0], th=0.01) for q, r in zip(seqs, recs)]),
                    nstem.make_params(bp_mode=1, bp_bound=0.01)), ("canonical + g-u", nstem.NstemSet(seqs), nstem.make_params(use_gu=True))):
    xi, yi = np.triu_indices(n)
    nstem.pairs(ctx, p, s, s, xi[:200], yi[:200])
    t = time.perf_counter(); v = nstem.pairs(ctx, p, s, s, xi, yi); dt = time.perf_counter() - t
    k = np.random.default_rng(1).choice(len(xi), 12, replace=False)
    w = O.nstem_pairs(p, s, s, xi[k], yi[k])
    cells = float(np.sum((np.array([len(q) for q in seqs])[xi] ** 2 / 2.0) * (np.array([len(q) for q in seqs])[yi] ** 2 / 2.0)))
    print(f"naive stem kernel, {name}: {len(xi)} pairs in {dt*1e3:.0f} ms: {len(xi)/dt:.0f} pairs/s, {cells/dt/1e9:.1f} G cells/s; "
          f"max rel err vs oracle {np.max(np.abs(v[k]-w)/np.abs(w)):.2e}", flush=True)
