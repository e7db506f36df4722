import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from stem_kernel_b200 import synth, hostlib, api, _lib as L
n = 2000
md = hostlib.build_many(synth.make_config(3, n))
flat = hostlib.SeqSet(md)
ctx = api.Context(L.make_params(L.SU_STEM)).set_option(L.OPT_TIMING, 1)
for it in range(3):
    t0 = time.perf_counter(); ds = ctx.upload(flat); t1 = time.perf_counter()
    ctx.stats_reset(); G = ctx.gram(ds, normalize=True); t2 = time.perf_counter(); st = ctx.stats()
    print(f"upload {1e3*(t1-t0):.0f} ms, gram wall {1e3*(t2-t1):.0f} ms, stem kernel {st['stem_ms']:.0f} ms, gap {1e3*(t2-t1)-st['stem_ms']:.0f} ms", flush=True)
    ds.free()
