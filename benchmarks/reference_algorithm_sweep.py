"""Container-only sweep: every algorithm class of the UNMODIFIED reference (/root/reference/AI4XDE/algorithm) is
constructed on the fused host path (`deepxde` := ai4xde_b200.dde; kernels replaced by the plain-PyTorch double,
there is no GPU here) over a small unmodified reference case and its own `closure()` is run with the iteration
counts cut down (Adam phases capped at 3 iterations, L-BFGS at 2).  Prints which run and why the others do not.
Usage: python benchmarks/reference_algorithm_sweep.py"""
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.simplefilter("ignore")
import numpy as np
import torch

import _taylor_double as td
from test_reference_dropin import load_reference

td.install()
A, dde = load_reference("fused")
dde.config.set_default_float("float32")
C = A.cases


from test_reference_dropin import ALGORITHMS, _cap_train_step

PLAN = [(name, (lambda mk: (lambda: mk(A)))(mk)) for name, mk in ALGORITHMS.items()]
capped = lambda cls: _cap_train_step(cls, dde)
ok, bad = [], []
for name, make in PLAN:
    try:
        np.random.seed(0)
        torch.manual_seed(0)
        solver = make()
        capped(type(solver))
        solver.closure()
        params = [p for p in solver.model.net.parameters()] if hasattr(solver, "model") else []
        assert all(torch.isfinite(p).all() for p in params)
        ok.append(name)
    except Exception as e:  # noqa: BLE001
        import traceback
        bad.append((name, f"{type(e).__name__}: {str(e)[:160]}", traceback.format_exc().splitlines()[-6:]))
print(f"{len(ok)} of {len(PLAN)} reference algorithm runs complete unmodified on the fused host path")
print("run:", ", ".join(ok))
for n, why, tb in bad:
    print(f"  {n}: {why}")
    if os.environ.get("SWEEP_TRACE"):
        print("\n".join("      " + t for t in tb))
