"""Small fixed workload for ncu: one stem-kernel Gram launch over n C3-like records."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stem_kernel_b200 import synth, hostlib, api, _lib as L
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
kind = int(sys.argv[2]) if len(sys.argv) > 2 else L.SU_STEM
cfg = int(sys.argv[3]) if len(sys.argv) > 3 else 3
recs = synth.make_config(cfg, n)
md = hostlib.build_many(recs) if cfg != 2 else [hostlib.MData.seq_only(r['rows']) for r in recs]
ctx = api.Context(L.make_params(kind))
ds = ctx.upload(md)
for _ in range(2):
    ctx.stats_reset(); G = ctx.gram(ds); print(ctx.stats())
