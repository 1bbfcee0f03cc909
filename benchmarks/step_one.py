#!/usr/bin/env python
"""A few train steps of ONE case, for ncu captures (`ncu -k regex:... python benchmarks/step_one.py --case ...`).
Cases: burgers20 (2-[20]x3-1, narrow FP32 kernels + residual program), helm_sin (reference-default Helmholtz
2-[150]x3-1 sin on the tcgen05 engine), cfg3 (NS cavity 2-[128]x5-3)."""
import argparse
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")

import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="burgers20")
    ap.add_argument("--n", type=int, default=2 ** 20)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--engine", default="auto")
    ap.add_argument("--timing", action="store_true", help="per-kernel CUDA-event times of one more step")
    args = ap.parse_args()
    import ai4xde_b200.dde as dde
    from ai4xde_b200 import cases

    dde.config.set_default_float("float32")
    dde.config.set_cuda_graphs(False)
    dde.config.set_engine(args.engine)
    np.random.seed(0)
    torch.manual_seed(0)
    lw = None
    if args.case == "burgers20":
        case = cases.Burgers(NumDomain=args.n, layer_size=[2] + [20] * 3 + [1], soft_constraints=True)
    elif args.case == "helm_sin":
        case, lw = cases.Helmholtz(NumDomain=args.n, honor_num_domain=True), [1, 100]
    elif args.case == "cfg3":
        case = cases.NS_Flow_in_LidDriven_Cavity(NumDomain=args.n, layer_size=[2] + [128] * 5 + [3])
    else:
        raise SystemExit(f"unknown case {args.case}")
    model = dde.Model(case.data, case.net)
    model.compile("adam", lr=1e-3, loss_weights=lw)
    X = case.data.train_x
    for _ in range(args.steps):
        model.train_step(X, None, None)
    torch.cuda.synchronize()
    print(args.case, len(X), "rows;", model.residual_program_note())
    if args.timing:
        import json

        from ai4xde_b200 import _lib

        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        _lib.timing_enable(True)
        e0.record()
        model.train_step(X, None, None)
        e1.record()
        rep = _lib.timing_report()
        _lib.timing_enable(False)
        print("step", round(e0.elapsed_time(e1), 3), "ms; kernels",
              json.dumps({k: [v["launches"], round(v["ms"], 3)] for k, v in sorted(rep.items())}))


if __name__ == "__main__":
    main()
