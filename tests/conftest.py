"""Shared fixtures.  Tests marked `gpu` need a B200 and go through the C ABI (include/stemk.h); everything else
runs on CPU: the oracle against the golden vectors, the host front end, the ABI surface, the multi-GPU host logic."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
TH = 0.01


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the native pieces once (no-op when up to date; make decides)."""
    import __graft_entry__ as g
    g.build()


def load_golden_records():
    z = np.load(os.path.join(GOLDEN, "golden_pairs.npz"))
    rows = json.loads(str(z["rows_json"]))
    off, k, recs = z["bp_off"], 0, []
    for i, rr in enumerate(rows):
        bp = []
        for _ in rr:
            a, b = int(off[k]), int(off[k + 1])
            bp.append((z["bp_i"][a:b], z["bp_j"][a:b], z["bp_p"][a:b]))
            k += 1
        recs.append(dict(rows=rr, bp=bp, label=int(z["labels"][i])))
    return recs, z


@pytest.fixture(scope="session")
def golden():
    """(records, npz of reference outputs, flattened SeqSet of our own front end's MData)."""
    from stem_kernel_b200 import hostlib
    recs, z = load_golden_records()
    md = [hostlib.MData.from_record(r, TH) for r in recs]
    return dict(recs=recs, z=z, md=md, flat=hostlib.SeqSet(md))


def relerr(got, want):
    """max |got-want|/|want| with exact matches (incl. 0 == 0, NaN == NaN, inf == inf) counted as 0."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    same = (got == want) | (np.isnan(got) & np.isnan(want))
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.abs(got - want) / np.abs(want)
    r = np.where(same, 0.0, r)
    if np.isnan(r).any():
        return float("inf")
    return float(r.max()) if r.size else 0.0


def need_gpu():
    from stem_kernel_b200 import _lib as L
    if L.lib().stemk_device_count() < 1:
        pytest.fail("gpu-marked test ran without a CUDA device (there is no CPU path)")
