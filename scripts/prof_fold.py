"""Small fixed workload for ncu: stemk_fold_bpp over n C3-like sequences (one fold_kernel launch per call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stem_kernel_b200 import fold
n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
rng = np.random.default_rng(1)
seqs = ["".join("acgu"[c] for c in rng.integers(0, 4, int(rng.integers(150, 301)))) for _ in range(n)]
with fold.Folder() as f:
    r = f.bpp(seqs, cutoff=1e-2)
    print(f"n={n} kernel {r.kernel_ms:.2f} ms")
