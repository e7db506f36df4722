"""Rectangular test x train matrix through ShardedCross on one GPU (the per-rank work of config C5 at reduced size):
pairs/s, agreement with stemk_cross on a sample of rows."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from stem_kernel_b200 import synth, hostlib, api, sharded, _lib as L
nt, ns = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1000, 2000)
train = hostlib.build_many(synth.make_config(3, ns, offset=0))
test = hostlib.build_many(synth.make_config(3, nt, offset=100000))
ctx = api.Context(L.make_params(L.SU_STEM))
dtrain, dtest = ctx.upload(train), ctx.upload(test)
dev = torch.device("cuda", 0)
be = sharded.GpuCrossBackend(ctx, dtrain, dtest, dev)
sc = sharded.ShardedCross(sharded.record_keys(dtest), sharded.record_keys(dtrain), 0, 1, dev, be.compute)
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.cuda.stream(be.stream):
        m, selfv = sc.run(normalize=True)
    be.stream.synchronize(); dt = time.perf_counter() - t0
    print(f"ShardedCross {nt} x {ns}: {sc.n_pairs} pairs in {dt*1e3:.0f} ms = {sc.n_pairs/dt:.0f} pairs/s", flush=True)
rows = np.arange(0, nt, max(1, nt // 8))[:8]
sub = hostlib.SeqSet([test[i] for i in rows])
one, _ = ctx.cross(ctx.upload(sub), dtrain, normalize=True)
got = m.cpu().numpy()[rows]
print("max rel diff vs stemk_cross on", len(rows), "rows:", float(np.max(np.abs(got - one) / np.abs(one))))
