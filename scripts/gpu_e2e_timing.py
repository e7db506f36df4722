"""Upload + gram + set_free with the library's own breakdown (STEMK_OPT_TIMING): where an end-to-end step's time goes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stem_kernel_b200 import synth, hostlib, api, _lib as L
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
md = hostlib.build_many(synth.make_config(3, n))
flat = hostlib.SeqSet(md)
ctx = api.Context(L.make_params(L.SU_STEM)).set_option(L.OPT_TIMING, 1)
for it in range(4):
    t0 = time.perf_counter(); ds = ctx.upload(flat); t1 = time.perf_counter()
    ctx.stats_reset(); G = ctx.gram(ds, normalize=True); t2 = time.perf_counter(); st = ctx.stats()
    ds.free(); t3 = time.perf_counter()
    print(f"upload {1e3*(t1-t0):.0f} ms, gram wall {1e3*(t2-t1):.0f} ms (stem kernels {st['stem_ms']:.0f} ms), set_free {1e3*(t3-t2):.1f} ms, "
          f"step {1e3*(t3-t0):.0f} ms", flush=True)
