#!/usr/bin/env python
"""Train-step time with the loss side as a traced residual program (ai4xde_b200/residual.py, csrc/residual_vm.cu)
against the same step on torch.autograd over the jet leaves.  Usage: python benchmarks/residual_bench.py"""
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
    from ai4xde_b200 import cases

    dde.config.set_default_float("float32")
    mk = {
        "cfg1 Burgers 2-[20]x3-1, 2540+80+160 pts (graph replay)":
            (lambda: cases.Burgers(NumDomain=2540, layer_size=[2] + [20] * 3 + [1], soft_constraints=True), None, 200),
        "Burgers 2-[20]x3-1, 2^20 pts":
            (lambda: cases.Burgers(NumDomain=2 ** 20, layer_size=[2] + [20] * 3 + [1], soft_constraints=True), None, 20),
        "Burgers reference default (hard constraint) 2-[32]x3-1, 2^18 pts":
            (lambda: cases.Burgers(NumDomain=2 ** 18), None, 20),
        "cfg2 AllenCahn 2-[64]x4-1, 10 k pts (graph replay)":
            (lambda: cases.AllenCahn(NumDomain=10000, layer_size=[2] + [64] * 4 + [1]), None, 100),
        "cfg3 NS cavity 2-[128]x5-3, 2^18 pts":
            (lambda: cases.NS_Flow_in_LidDriven_Cavity(NumDomain=2 ** 18, layer_size=[2] + [128] * 5 + [3]), None, 20),
        "cfg4 Helmholtz 2-[256]x6-1 tanh, 2^20 pts":
            (lambda: cases.Helmholtz(NumDomain=2 ** 20, layer_size=[2] + [256] * 6 + [1], activation="tanh",
                                     honor_num_domain=True), [1, 100], 10),
        "Helmholtz reference default 2-[150]x3-1 sin, 2^18 pts":
            (lambda: cases.Helmholtz(NumDomain=2 ** 18, honor_num_domain=True), [1, 100], 20),
    }
    out = {}
    for name, (build, lw, steps) in mk.items():
        row = {}
        for fused in (False, True):
            np.random.seed(0)
            torch.manual_seed(0)
            dde.config.set_residual_programs(fused)
            case = build()
            model = dde.Model(case.data, case.net)
            model.compile("adam", lr=1e-3, loss_weights=lw)
            X = case.data.train_x
            ms = time_steps(lambda: model.train_step(X, None, None), steps, 8)
            row["program" if fused else "autograd"] = round(ms, 4)
            if fused:
                row["note"] = model.residual_program_note()
                row["points_per_s"] = round(len(X) / ms * 1e3)
            del model, case
            torch.cuda.empty_cache()
        out[name] = row
    dde.config.set_residual_programs(True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
