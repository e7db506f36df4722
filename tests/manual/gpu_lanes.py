"""GPU check of the lanes-are-rows stem kernel (STEMK_LANES=1): whole Gram matrix against the fast kernel,
sampled pairs against the oracle, and timing on C3-like records."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from stem_kernel_b200 import synth, hostlib, api, _lib as L
from oracle import oraclebind as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 3
modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["fast", "lanes"]
recs = synth.make_config(cfg, n)
md = hostlib.build_many(recs)
S = hostlib.SeqSet(md)
p = L.make_params(L.SU_STEM)
ref = None
for mode in modes:
    os.environ["STEMK_LANES"] = "0" if mode == "fast" else "1"
    os.environ["STEMK_LANES_R"] = mode[5:] if mode.startswith("lanes") and len(mode) > 5 else "0"
    ctx = api.Context(p); ds = ctx.upload(S)
    G = ctx.gram(ds)
    rng = np.random.default_rng(1)
    xi, yi = rng.integers(0, n, 24), rng.integers(0, n, 24)
    lo, hi = np.minimum(xi, yi), np.maximum(xi, yi)
    want = O.pairs(O.Params.from_buffer_copy(p), S.desc(), S.desc(), lo, hi)
    err = np.max(np.abs(G[lo, hi] - want) / np.abs(want))
    if ref is None: ref = G
    dif = np.max(np.abs(G - ref) / np.abs(ref))
    ctx.stats_reset(); t0 = time.time(); G2 = ctx.gram(ds); t1 = time.time(); st = ctx.stats()
    rep = np.max(np.abs(G2 - G))
    npairs = n * (n + 1) // 2
    print(f"{mode:8s} n={n} cfg={cfg} oracle rel err {err:.2e} vs first mode {dif:.2e} rerun diff {rep:.1e} stem_ms {st['stem_ms']:.1f} "
          f"launches {st['launches']} kernel pairs/s {npairs/(st['stem_ms']*1e-3):.0f}", flush=True)
    ctx.close()
