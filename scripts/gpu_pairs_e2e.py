"""stemk_pairs with host buffers: wall time of the call against the device time of its kernels."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stem_kernel_b200 import synth, hostlib, api, sharded, _lib as L
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1200
md = hostlib.build_many(synth.make_config(3, n))
ctx = api.Context(L.make_params(L.SU_STEM)); ds = ctx.upload(md)
xi, yi = sharded.square_pairs(sharded.record_keys(ds))
for it in range(3):
    ctx.stats_reset(); t0 = time.perf_counter(); v = ctx.pairs(ds, ds, xi, yi); dt = time.perf_counter() - t0
    print(f"stemk_pairs {len(xi)} pairs: wall {1e3*dt:.0f} ms, kernels {ctx.stats()['stem_ms']:.0f} ms, checksum {v.sum():.10e}", flush=True)
