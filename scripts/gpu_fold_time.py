"""Timing of stemk_fold_bpp on n C3-like sequences (150-300 nt): kernel and host-buffer call (STEMK_SO selects a tuning build)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stem_kernel_b200 import fold
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
rng = np.random.default_rng(1)
seqs = ["".join("acgu"[c] for c in rng.integers(0, 4, int(rng.integers(150, 301)))) for _ in range(n)]
m = fold.default_model()
with fold.Folder() as f:
    f.bpp(seqs[:64], m, cutoff=1e-3)
    for cut in (1e-2, 1e-4):
        t = time.time(); r = f.bpp(seqs, m, cutoff=cut); dt = time.time() - t
        print(f"n={n} cutoff={cut} call {dt*1e3:.1f} ms ({n/dt:.0f} seq/s), kernel {r.kernel_ms:.1f} ms ({n/r.kernel_ms*1e3:.0f} seq/s)  pairs listed {sum(len(p[0]) for p in r.pairs)}", flush=True)
