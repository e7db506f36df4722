"""Quick GPU check of the fast stem kernel: parity vs oracle on a few pairs + timing on C3-like records."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from stem_kernel_b200 import synth, hostlib, api, _lib as L
from oracle import oraclebind as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
recs = synth.make_config(3, n)
md = hostlib.build_many(recs)
S = hostlib.SeqSet(md)
p = L.make_params(L.SU_STEM)
for fast in ("1", "0"):
    os.environ["STEMK_FAST"] = fast
    ctx = api.Context(p); ds = ctx.upload(S)
    G = ctx.gram(ds)
    rng = np.random.default_rng(1)
    xi, yi = rng.integers(0, n, 24), rng.integers(0, n, 24)
    lo, hi = np.minimum(xi, yi), np.maximum(xi, yi)
    want = O.pairs(O.Params.from_buffer_copy(p), S.desc(), S.desc(), lo, hi)
    err = np.max(np.abs(G[lo, hi] - want) / np.abs(want))
    ctx.stats_reset(); t0 = time.time(); G = ctx.gram(ds); t1 = time.time(); st = ctx.stats()
    npairs = n * (n + 1) // 2
    print(f"fast={fast} n={n} max rel err {err:.2e} gram {t1-t0:.3f}s stem_ms {st['stem_ms']:.1f} launches {st['launches']} kernel pairs/s {npairs/(st['stem_ms']*1e-3):.0f}", flush=True)
    ctx.close()
