"""Quick GPU parity check of the fast stem kernel against the oracle (development aid; the gate is tests/test_gpu_parity.py).
Usage: python tests/manual/gpu_fast_check.py [n_c3 [n_c1]]   (STEMK_SO selects a tuning build)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from stem_kernel_b200 import synth, hostlib, api, _lib as L
from oracle import oraclebind as O

n3 = int(sys.argv[1]) if len(sys.argv) > 1 else 40
n1 = int(sys.argv[2]) if len(sys.argv) > 2 else 12
md = hostlib.build_many(synth.make_config(3, n3, offset=5100) + synth.make_config(1, n1, offset=5200))
flat = hostlib.SeqSet(md)
n = len(md)
xi, yi = np.triu_indices(n)
rng = np.random.default_rng(1)
pick = rng.choice(len(xi), min(len(xi), 150), replace=False)
bad = 0
for kind, band in ((L.SU_STEM, 10), (L.SU_STEM, 0), (L.SI_STEM, 3), (L.SU_STEM, 40)):
    p = L.make_params(kind, len_band=band)
    ctx = api.Context(p)
    ds = ctx.upload(flat)
    G = ctx.gram(ds, normalize=False)
    G2 = ctx.gram(ds, normalize=False)
    want = O.pairs(O.Params.from_buffer_copy(p), flat.desc(), flat.desc(), xi[pick].astype(np.uint32), yi[pick].astype(np.uint32))
    got = G[xi[pick], yi[pick]]
    err = np.abs(got - want) / np.abs(want)
    w = int(np.argmax(err))
    print(f"kind {kind} band {band}: max rel err {err.max():.3e} (pair {xi[pick][w]},{yi[pick][w]} got {got[w]:.6e} want {want[w]:.6e}) "
          f"repro {np.array_equal(G, G2)} finite {np.isfinite(G).all()}", flush=True)
    bad += int(err.max() > 1e-9) + int(not np.array_equal(G, G2))
    ctx.close()
print("FAST_CHECK", "OK" if bad == 0 else f"FAILED ({bad})")
sys.exit(1 if bad else 0)
