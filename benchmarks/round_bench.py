"""Wall time of one resampling ROUND as the algorithms run it (probe + selection + refill + bookkeeping + what the
next train step has to upload), on one B200:
  * R3Sampling on the cfg4 population (Helmholtz 2-[256]x6-1, 2^20 points): parity mode with a float64 host (exact
    reference arrays, NumPy refill), parity mode with a float32 host, device_rng mode (Philox refill in a kernel);
  * RAD on cfg2 (AllenCahn 2-[64]x4-1): 100 000 candidates, 10 000 selected.
Each figure is the mean of `reps` rounds after a warm-up round, host clock around resample() + the following
train step's upload (torch.cuda.synchronize() on both sides).  python benchmarks/round_bench.py"""
import json
import os
import sys
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import numpy as np
import torch

import ai4xde_b200.dde as dde
from ai4xde_b200 import algorithm, cases


def timed_rounds(solver, reps=3):
    model = solver.model
    solver.PDECase.compile(model, "adam", lr=1e-3)
    X = solver.PDECase.data.train_x
    model.train_step(X, None, None)          # stream discovery, uploads
    solver.resample()                        # warm-up round
    torch.cuda.synchronize()
    t_round, t_next = 0.0, 0.0
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        solver.resample()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        X = solver.PDECase.data.train_x
        model.device_batch(X)                # what the next train step would have to upload
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        t_round += t1 - t0
        t_next += t2 - t1
    return {"round_ms": 1e3 * t_round / reps, "next_step_upload_ms": 1e3 * t_next / reps}


def main():
    out = {}
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2**20
    for label, host_float, kw in (("r3_parity_float64_host", "float64", {}), ("r3_parity_float32_host", "float32", {}),
                                  ("r3_device_rng", "float32", {"device_rng": True})):
        dde.config.set_default_float(host_float)
        np.random.seed(0)
        torch.manual_seed(0)
        case = cases.Helmholtz(NumDomain=n, layer_size=[2] + [256] * 6 + [1], activation="tanh", honor_num_domain=True)
        solver = algorithm.R3Sampling(case, max_iter=1, **kw)
        out[label] = dict(points=n, **timed_rounds(solver))
        del solver, case
        torch.cuda.empty_cache()
    dde.config.set_default_float("float64")
    np.random.seed(0)
    torch.manual_seed(0)
    case = cases.AllenCahn(NumDomain=10000, layer_size=[2] + [64] * 4 + [1])
    solver = algorithm.RAD(case, iter=1, k=1, c=1)
    out["rad_100k_candidates"] = dict(candidates=100000, selected=10000, **timed_rounds(solver))
    np.random.seed(0)
    torch.manual_seed(0)
    case = cases.AllenCahn(NumDomain=10000, layer_size=[2] + [64] * 4 + [1])
    solver = algorithm.RAD(case, iter=1, k=1, c=1, fast=True)
    out["rad_100k_candidates_fast_scan"] = dict(candidates=100000, selected=10000, **timed_rounds(solver))
    dde.config.set_default_float("float32")
    np.random.seed(0)
    torch.manual_seed(0)
    case = cases.AllenCahn(NumDomain=10000, layer_size=[2] + [64] * 4 + [1])
    solver = algorithm.RAD(case, iter=1, k=1, c=1, device_rng=True)
    out["rad_100k_candidates_device_rng"] = dict(candidates=100000, selected=10000, **timed_rounds(solver))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
