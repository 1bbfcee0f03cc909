"""Container-only sweep: every case class of the UNMODIFIED reference (/root/reference/AI4XDE/cases) is
constructed on the fused host path (`deepxde` := ai4xde_b200.dde; kernels replaced by the plain-PyTorch
double, there is no GPU here) and asked for one losses + backward evaluation.  Prints which run and why
the others do not.  Usage: python benchmarks/reference_sweep.py"""
import inspect
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
ok, bad = [], []
mods = [getattr(A.cases, m) for m in ("PDECases", "PoissonCases", "WaveCases", "ConvectionCases", "FuncCases",
                                     "InverseCases") if hasattr(A.cases, m)]
seen = set()
for mod in mods:
    for name, cls in inspect.getmembers(mod, inspect.isclass):
        if cls.__module__ != mod.__name__ or inspect.isabstract(cls) or cls.__subclasses__() or name in seen:
            continue
        seen.add(name)
        try:
            np.random.seed(0)
            torch.manual_seed(0)
            case = cls()
            model = dde.Model(case.data, case.net)
            case.compile(model, "adam", lr=1e-3)
            X, y, aux = case.data.train_next_batch()
            _, losses = model.outputs_losses_train(X, y, aux)
            case.net.zero_grad()
            losses.sum().backward()
            assert all(torch.isfinite(p.grad).all() for p in case.net.parameters() if p.grad is not None)
            ok.append(name)
        except Exception as e:  # noqa: BLE001
            bad.append((name, f"{type(e).__name__}: {str(e)[:110]}"))
print(f"{len(ok)} of {len(ok) + len(bad)} reference case classes run unmodified on the fused host path")
print("run:", ", ".join(ok))
for n, why in bad:
    print(f"  {n}: {why}")
