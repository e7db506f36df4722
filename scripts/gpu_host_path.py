"""Where the host-buffer calls spend their time (STEMK_OPT_TIMING prints the library's own breakdown on stderr):
C2 string-kernel Gram (2 000 x 100 nt), a rectangular cross call and a C3 stem Gram of n records."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
from stem_kernel_b200 import synth, hostlib, api, _lib as L
n3 = int(sys.argv[1]) if len(sys.argv) > 1 else 600
md2 = [hostlib.MData.seq_only(r["rows"]) for r in synth.make_config(2)]
ctx = api.Context(L.make_params(L.STR_SUBST)).set_option(L.OPT_TIMING, 1); ds = ctx.upload(md2)
p2 = len(md2) * (len(md2) + 1) // 2
for it in range(4):
    t0 = time.perf_counter(); G = ctx.gram(ds, normalize=True); dt = time.perf_counter() - t0
    print(f"C2 gram wall {1e3*dt:.1f} ms = {p2/dt/1e6:.1f} M pairs/s", flush=True)
test = ctx.upload(md2[:500])
for it in range(3):
    t0 = time.perf_counter(); C, s = ctx.cross(test, ds, normalize=True); dt = time.perf_counter() - t0
    print(f"C2 cross 500 x 2000 wall {1e3*dt:.1f} ms = {500*2000/dt/1e6:.1f} M pairs/s", flush=True)
ctx.close()
md3 = hostlib.build_many(synth.make_config(3, n3))
ctx = api.Context(L.make_params(L.SU_STEM)).set_option(L.OPT_TIMING, 1)
for it in range(3):
    t0 = time.perf_counter(); ds3 = ctx.upload(md3); t1 = time.perf_counter()
    ctx.stats_reset(); G = ctx.gram(ds3, normalize=True); t2 = time.perf_counter(); st = ctx.stats()
    print(f"C3 n={n3} upload {1e3*(t1-t0):.1f} ms, gram wall {1e3*(t2-t1):.1f} ms, kernels {st['stem_ms']:.1f} ms", flush=True)
    ds3.free()
