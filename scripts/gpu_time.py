"""Timing of one stem-kernel Gram matrix over n C3 records (STEMK_SO selects a tuning build of the library)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stem_kernel_b200 import synth, hostlib, api, _lib as L
n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
md = hostlib.build_many(synth.make_config(3, n))
ctx = api.Context(L.make_params(L.SU_STEM)); ds = ctx.upload(md)
ctx.gram(ds)
ctx.stats_reset(); G = ctx.gram(ds); st = ctx.stats()
npairs = n * (n + 1) // 2
print(f"{os.environ.get('STEMK_SO','default'):40s} n={n} stem_ms {st['stem_ms']:.1f} kernel pairs/s {npairs/(st['stem_ms']*1e-3):.0f} checksum {G.sum():.12e}", flush=True)
