import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from stem_kernel_b200 import fold, synth, hostlib, api, _lib as L
rng = np.random.default_rng(3)
seqs = ["".join("acgu"[c] for c in rng.integers(0, 4, n)) for n in [0, 1, 5, 9, 33, 64, 97, 130]] + ["GGGAAATCCCNNNGGGTTTCCC"]
with fold.Folder() as f:
    r = f.bpp(seqs, cutoff=0.0, dense=True)
    r = f.bpp(seqs, cutoff=0.01)
print("fold ok", sum(len(p[0]) for p in r.pairs))
recs = [synth.ncrna_like(777, 0, 700, 760)] + synth.make_config(3, 2, offset=50) + [synth.alignment_like(5, 1, n_rows=3, L=70)]
flat = hostlib.SeqSet([hostlib.MData.from_record(r, 0.01) for r in recs])
ctx = api.Context(L.make_params(L.SU_STEM))
g = ctx.gram(ctx.upload(flat)); print("anysize ok", g[0, 0])
ctx.set_option(L.OPT_FORCE_GENERAL, True).set_option(L.OPT_FORCE_UNSTAGED, True)
small = hostlib.SeqSet([hostlib.MData.from_record(r, 0.01) for r in synth.make_config(1, 4)])
print("unstaged ok", ctx.gram(ctx.upload(small))[0, 0])
