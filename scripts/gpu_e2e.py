"""Where the end-to-end time of upload + gram goes (host-buffer C ABI path), n C3 records."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stem_kernel_b200 import synth, hostlib, api, _lib as L
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
md = hostlib.build_many(synth.make_config(3, n))
ctx = api.Context(L.make_params(L.SU_STEM))
for it in range(3):
    t0 = time.perf_counter(); ds = ctx.upload(md); t1 = time.perf_counter()
    ctx.stats_reset(); G = ctx.gram(ds, normalize=True); t2 = time.perf_counter(); st = ctx.stats()
    npairs = n * (n + 1) // 2
    print(f"n={n} upload {1e3*(t1-t0):.0f} ms, gram wall {1e3*(t2-t1):.0f} ms, stem kernel {st['stem_ms']:.0f} ms, "
          f"kernel pairs/s {npairs/(st['stem_ms']*1e-3):.0f}, e2e pairs/s {npairs/(t2-t0):.0f}", flush=True)
