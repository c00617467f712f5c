"""Experiment: what does the fused epilogue of k_site cost?  Times the genotype pass with the fused site
conditions (site_flags) against the counts-only pass (site_counts: no epilogue, but 144 MB of extra stores)."""
import ctypes as C, json, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sai_b200 import _cabi
from sai_b200.encode import make_layout
from sai_b200.scoring import DeviceScorer, make_job, synth_fill

S = 6_000_000
lay = make_layout([1500, 1000, 4], [2, 2, 2], [2, 2, 2])
nbytes = int(_cabi.load().sai_packed_bytes(C.byref(lay), S))
d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
synth_fill(lay, d, S, [0, 1, 2], 20261019, 0.0)
job = make_job(0, 1, [2], True, u=dict(w=0.01, x=0.5, y_list=[("=", 1.0)]), q=dict(w=0.01, quantile=0.95, y_list=[("=", 1.0)]))
sc = DeviceScorer(lay, S, 1, 1)
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
out = {}
for v in (0, 1, 6):
    out[f"flags_v{v}"] = timed(lambda: sc.site_flags(d, [job], v))
    out[f"counts_v{v}"] = timed(lambda: sc.site_counts(d, v))
    out[f"flags+counts_v{v}"] = timed(lambda: sc.site_flags(d, [job], v, with_counts=True))
print(json.dumps({k: round(x, 4) for k, x in out.items()}))
