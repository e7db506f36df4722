"""Timing of one string-kernel Gram matrix (C2: n random sequences of 100 nt)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stem_kernel_b200 import synth, hostlib, api, _lib as L
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
kind = int(sys.argv[2]) if len(sys.argv) > 2 else L.STR_SUBST
md = [hostlib.MData.seq_only(r['rows']) for r in synth.make_config(2, n)]
ctx = api.Context(L.make_params(kind)); ds = ctx.upload(md)
ctx.gram(ds)
ctx.stats_reset(); G = ctx.gram(ds); st = ctx.stats()
npairs = n * (n + 1) // 2
print(f"string kind={kind} n={n} kernel_ms {st['string_ms']:.2f} pairs/s {npairs/(st['string_ms']*1e-3):.0f} GCUPS {npairs*1e4/(st['string_ms']*1e-3)/1e9:.1f} checksum {G.sum():.12e}", flush=True)
