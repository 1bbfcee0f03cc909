"""Device time and effective HBM rate of the resampling kernels ("HBM GB/s for
select/scan/compact").  Algorithmic bytes: R3 retain 24 B/point (read |r| 4, write sum 4, read it twice for the
mean and the predicate 8, read x 8 ... the kernels re-read the 4-byte residual sum three times: 28 B/point with
d = 2), RAD probabilities 16 B/candidate (read 4, write the float64 p 8, pairwise mean re-reads 4) and one choice
round 16 B/candidate (cumsum read + write of float64) + 8 B per draw."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ai4xde_b200 import _lib
from ai4xde_b200.jet import to_soa

def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

out = {}
g = torch.Generator().manual_seed(0)
for n in (2**20, 2**24):
    res = torch.zeros((1, _lib.pad32(n)), device="cuda"); res[0, :n] = torch.randn(n, generator=g).cuda()
    x = to_soa(torch.rand((n, 2), generator=g).cuda())
    ms = timed(lambda: _lib.r3_retain(res, x, n, place_at_end=True))
    out[f"r3_retain n={n}"] = {"ms": ms, "GB/s": 28.0 * n / (ms * 1e-3) / 1e9}
for n, m in ((100000, 10000), (2**22, 2**18)):
    res = torch.zeros((1, _lib.pad32(n)), device="cuda"); res[0, :n] = torch.randn(n, generator=g).cuda()
    ms = timed(lambda: _lib.rad_probabilities(res, n, 1.0, 1.0))
    out[f"rad_probabilities n={n}"] = {"ms": ms, "GB/s": 16.0 * n / (ms * 1e-3) / 1e9}
    p = _lib.rad_probabilities(res, n, 1.0, 1.0)
    u = torch.rand(m, generator=g, dtype=torch.float64).cuda()
    found = torch.zeros(m, dtype=torch.int64, device="cuda"); nf = torch.zeros(1, dtype=torch.int64, device="cuda")
    ws, wsb = _lib.resample_ws(n, "cuda")
    def rnd():
        pp = p.clone()
        _lib.choice_round(pp, u, found, 0, nf, ws, wsb)
    ms = timed(rnd)
    out[f"choice_round n={n} draws={m}"] = {"ms": ms, "GB/s": (24.0 * n + 8.0 * m) / (ms * 1e-3) / 1e9}
for k, v in out.items():
    print(k, {a: round(b, 3) for a, b in v.items()})
