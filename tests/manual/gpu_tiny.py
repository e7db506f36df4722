import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from stem_kernel_b200 import synth, hostlib, api, _lib as L
from oracle import oraclebind as O
recs = synth.make_config(1, 4)
md = [hostlib.MData.from_record(r) for r in recs]
S = hostlib.SeqSet(md)
p = L.make_params(L.SU_STEM)
ctx = api.Context(p); ds = ctx.upload(S)
G = ctx.gram(ds)
Go = O.gram(O.Params.from_buffer_copy(p), S.desc(), False)
print(np.abs(G-Go).max()/np.abs(Go).max())
