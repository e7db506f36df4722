"""First GPU contact: parity of every kernel kind against the C oracle on small sets + rough timing."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from stem_kernel_b200 import synth, hostlib, api, _lib as L
from oracle import oraclebind as O

def relerr(a, b):
    with np.errstate(divide='ignore', invalid='ignore'):
        r = np.abs(a - b) / np.abs(b)
    r[(a == b)] = 0
    return np.nanmax(r)

recs = synth.make_config(1, 12) + synth.make_config(3, 6) + [synth.alignment_like(7, i, n_rows=(i % 4) + 1) for i in range(8)]
md = [hostlib.MData.from_record(r) for r in recs]
S = hostlib.SeqSet(md); d = S.desc()
ok = True
for kind in range(9):
    for band in (10, 0):
        p = L.make_params(kind, len_band=band)
        ctx = api.Context(p)
        ds = ctx.upload(S)
        G = ctx.gram(ds, False)
        Go = O.gram(O.Params.from_buffer_copy(p), d, False)
        e = relerr(G, Go)
        Gn = ctx.gram(ds, True); Gon = O.gram(O.Params.from_buffer_copy(p), d, True)
        en = relerr(Gn, Gon)
        print(f"kind {kind} band {band}: max rel err {e:.3e} normalised {en:.3e}", flush=True)
        ok &= e < 1e-9 and en < 1e-9
        ctx.close()
# naive
seqs = [r['rows'][0] for r in synth.make_config(2, 10)]
mdn = [hostlib.MData.seq_only([s]) for s in seqs]; Sn = hostlib.SeqSet(mdn)
p = L.make_params(L.STR_NAIVE, gap=float(np.float32(0.8)))
ctx = api.Context(p); ds = ctx.upload(Sn)
e = relerr(ctx.gram(ds), O.gram(O.Params.from_buffer_copy(p), Sn.desc(), False)); print('naive', e); ok &= e < 1e-9
print('fp64 peak TF/s', ctx.fp64_peak(0.5))
ctx.close()
print('PARITY', 'OK' if ok else 'FAIL', flush=True)

# timing: C3-like
for n in (64, 256):
    recs = synth.make_config(3, n)
    t0 = time.time(); md = hostlib.build_many(recs); t1 = time.time()
    S = hostlib.SeqSet(md)
    ctx = api.Context(L.make_params(L.SU_STEM))
    t2 = time.time(); ds = ctx.upload(S); t3 = time.time()
    ctx.stats_reset()
    G = ctx.gram(ds); t4 = time.time()
    G = ctx.gram(ds); t5 = time.time()
    st = ctx.stats()
    npairs = n * (n + 1) // 2
    idx = np.triu_indices(n)
    cells, flops = ctx.pair_cost(ds, ds, idx[0], idx[1])
    print(f"n={n} build {t1-t0:.2f}s upload {t3-t2:.3f}s gram {t5-t4:.3f}s stem_ms(2 runs) {st['stem_ms']:.1f} pairs/s {npairs/(t5-t4):.0f} kernel pairs/s {2*npairs/(st['stem_ms']*1e-3):.0f} alg TF/s {2*flops.sum()/(st['stem_ms']*1e-3)/1e12:.3f}", flush=True)
    ctx.close()
for n in (512,):
    recs = synth.make_config(2, n)
    md = [hostlib.MData.seq_only(r['rows']) for r in recs]
    ctx = api.Context(L.make_params(L.STR_SUBST)); ds = ctx.upload(md)
    ctx.stats_reset(); t0 = time.time(); G = ctx.gram(ds); t1 = time.time(); st = ctx.stats()
    npairs = n * (n + 1) // 2
    print(f"string n={n}: gram {t1-t0:.3f}s kernel ms {st['string_ms']:.2f} GCUPS {npairs*1e4/(st['string_ms']*1e-3)/1e9:.2f}", flush=True)
