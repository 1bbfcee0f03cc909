#!/usr/bin/env python
"""Which of the UNMODIFIED reference case classes get their loss side traced into a residual program, and why the
others stay on autograd (container only: needs /root/reference; kernels replaced by the plain-PyTorch doubles).
For every traced case the train closure on the program is compared with the closure on autograd."""
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
import test_reference_dropin as drop


def closure(dde, model, fused):
    dde.config.set_residual_programs(fused)
    model.net.zero_grad()
    seen, step = {}, model.opt.step
    model.opt.step = lambda c: seen.__setitem__("l", c())
    try:
        model.train_step(model.data.train_x, None, getattr(model.data, "train_aux_vars", None))
    finally:
        model.opt.step = step
    g = torch.cat([p.grad.reshape(-1) for p in model.net.parameters() if p.grad is not None])
    return float(seen["l"]), g.clone()


def main():
    td.install()
    traced = declined = failed = 0
    for name, kwargs in [(c.values[0], c.values[1]) if hasattr(c, "values") else c for c in drop.CASES]:
        try:
            B, dde = drop.load_reference("fused")
            dde.config.device = lambda: torch.device("cpu")
            np.random.seed(0)
            torch.manual_seed(0)
            case = getattr(B.cases, name)(**kwargs)
            model = dde.Model(case.data, case.net)
            case.compile(model, "adam", lr=1e-3)
            l0, g0 = closure(dde, model, False)
            l1, g1 = closure(dde, model, True)
            note = model.residual_program_note()
            ok = note is not None and not note.startswith("autograd path")
            traced += ok
            declined += not ok
            agree = "" if not ok else f"  loss rel {abs(l1 - l0) / max(abs(l0), 1e-30):.1e} grad rel {float((g1 - g0).abs().max() / g0.abs().max().clamp_min(1e-30)):.1e}"
            print(f"{name:45s} {'PROGRAM ' if ok else 'autograd'} {note}{agree}", flush=True)
        except Exception as e:  # noqa: BLE001
            failed += 1
            print(f"{name:45s} ERROR    {type(e).__name__}: {str(e)[:120]}", flush=True)
        finally:
            drop._purge()
    print(f"traced {traced}, on autograd {declined}, errors {failed}")


if __name__ == "__main__":
    main()
