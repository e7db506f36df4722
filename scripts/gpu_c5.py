"""BASELINE config C5 at full size: rectangular 5 000 test x 10 000 train stem-kernel matrix (50 015 000 kernel
evaluations with the self terms and the train diagonals) through ShardedCross on N GPUs (torchrun, one rank per GPU),
normalised on rank 0.  Prints ONE JSON line (rank 0): pairs/s over the device-timed run, max over ranks, and the
agreement of 8 sampled rows with the single-GPU stemk_cross call."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from stem_kernel_b200 import synth, hostlib, api, sharded, _lib as L
nt, ns = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (5000, 10000)
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
t0 = time.perf_counter()
train = hostlib.build_many(synth.make_config(3, ns, offset=0))
test = hostlib.build_many(synth.make_config(3, nt, offset=100000))
t_build = time.perf_counter() - t0
ctx = api.Context(L.make_params(L.SU_STEM), device=local)
dtrain, dtest = ctx.upload(train), ctx.upload(test)
be = sharded.GpuCrossBackend(ctx, dtrain, dtest, dev)
sc = sharded.ShardedCross(sharded.record_keys(dtest), sharded.record_keys(dtrain), rank, world, dev, be.compute)
def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(be.stream):
    e0.record()
    m, selfv = sc.run(normalize=True)
    e1.record()
barrier()
t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(t[0])
    rows = np.arange(0, nt, max(1, nt // 8))[:8]
    one, _ = ctx.cross(ctx.upload(hostlib.SeqSet([test[i] for i in rows])), dtrain, normalize=True)
    got = m[torch.from_numpy(rows).to(dev)].cpu().numpy()
    print(json.dumps({"workload": f"C5: {nt} test x {ns} train, SuStemKernel, normalised", "n_gpus": world, "pairs": int(sc.n_pairs),
                      "ms": ms, "pairs_per_s": sc.n_pairs / (ms * 1e-3), "host_build_s": t_build,
                      "finite": bool(torch.isfinite(m).all()), "rows_checked": len(rows),
                      "max_rel_diff_vs_stemk_cross": float(np.max(np.abs(got - one) / np.abs(one)))}), flush=True)
if world > 1: dist.destroy_process_group()
