"""bench.py's end-to-end loop, piece by piece: stemk_upload / stemk_gram / stemk_set_free."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stem_kernel_b200 import synth, hostlib, api, _lib as L
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
md = hostlib.build_many(synth.make_config(3, n))
flat = hostlib.SeqSet(md); desc = flat.desc()
ctx = api.Context(L.make_params(L.SU_STEM))
for it in range(6):
    ctx.stats_reset()
    t0 = time.perf_counter(); h = C.c_void_p(); ctx._check(L.lib().stemk_upload(ctx.h, C.byref(desc), C.byref(h)))
    t1 = time.perf_counter(); out = np.empty((n, n)); t2 = time.perf_counter()
    ctx._check(L.lib().stemk_gram(ctx.h, h, 1, out.ctypes.data)); t3 = time.perf_counter()
    L.lib().stemk_set_free(ctx.h, h); t4 = time.perf_counter()
    st = ctx.stats()
    print(f"kernel {st['stem_ms']:.0f} ms | upload {1e3*(t1-t0):.0f} ms, np.empty {1e3*(t2-t1):.1f} ms, gram {1e3*(t3-t2):.0f} ms, set_free {1e3*(t4-t3):.0f} ms, total {1e3*(t4-t0):.0f} ms", flush=True)
