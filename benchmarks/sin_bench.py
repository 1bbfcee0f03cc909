#!/usr/bin/env python
"""sin nets of the reference on the tensor-core engine vs the FP32 engine:
Helmholtz reference-default 2-[150]x3-1 sin (cases/PDECases.py:513-515) and the hard-BC wave net 2-[100]x5-1 sin
(cases/WaveCases.py:19), one Adam train step at 2^18 points.  Usage: python benchmarks/sin_bench.py"""
import json
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")

import numpy as np
import torch


def time_steps(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    import ai4xde_b200.dde as dde
    from ai4xde_b200 import cases, _lib

    dde.config.set_default_float("float32")
    out = {}
    n = 2 ** 18
    for name, layer_size, act, engines in (
            ("Helmholtz 2-[150]x3-1 sin", [2] + [150] * 3 + [1], "sin", ("fp32", "auto")),
            ("Helmholtz 2-[150]x3-1 tanh", [2] + [150] * 3 + [1], "tanh", ("auto",)),
            ("Helmholtz 2-[100]x5-1 sin (wave-net shape)", [2] + [100] * 5 + [1], "sin", ("fp32", "auto")),
            ("Helmholtz 2-[100]x5-1 tanh", [2] + [100] * 5 + [1], "tanh", ("auto",)),
            ("Helmholtz 2-[256]x3-1 sin", [2] + [256] * 3 + [1], "sin", ("auto",))):
        for engine in engines:
            np.random.seed(0)
            torch.manual_seed(0)
            dde.config.set_engine(engine)
            case = cases.Helmholtz(NumDomain=n, layer_size=layer_size, activation=act, honor_num_domain=True)
            model = dde.Model(case.data, case.net)
            case.compile(model, "adam", lr=1e-3)
            X = case.data.train_x
            ms = time_steps(lambda: model.train_step(X, None, None), 20, 8)   # warm-up covers the CUDA-graph capture (4th call)
            dde.config.set_cuda_graphs(False)      # per-kernel event times need the eager launches
            model.train_step(X, None, None)
            _lib.timing_enable(True)
            model.train_step(X, None, None)
            rep = _lib.timing_report()
            _lib.timing_enable(False)
            dde.config.set_cuda_graphs("auto")
            out[f"{name}, 2^18 pts, engine {engine}"] = {
                "ms_per_step": round(ms, 3), "points_per_s": round(len(X) / ms * 1e3),
                "kernels_ms": {k: [v["launches"], round(v["ms"], 3)] for k, v in sorted(rep.items())}}
            del model, case
            torch.cuda.empty_cache()
    dde.config.set_engine("auto")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
