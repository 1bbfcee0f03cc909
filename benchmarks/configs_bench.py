#!/usr/bin/env python
"""Per-config step times of the fused path (all five BASELINE configs + the resampling rounds),
with the CPU oracle beside them on a bounded sample.  Not the driver's bench (that is /bench.py);
this fills the per-config table in DESIGN.md.  Usage: python benchmarks/configs_bench.py [--cpu]"""
import argparse
import json
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
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
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true", help="also time the oracle on the host cores")
    ap.add_argument("--engine", default="auto")
    args = ap.parse_args()
    import ai4xde_b200.dde as dde
    from ai4xde_b200 import algorithm, cases, resample

    dde.config.set_default_float("float64")
    dde.config.set_engine(args.engine)
    out = {}

    def train_ms(case, steps=50, warmup=10, lw="case"):
        model = dde.Model(case.data, case.net)
        case.compile(model, "adam", lr=1e-3)
        X = case.data.train_x
        ms = time_steps(lambda: model.train_step(X, None, None), steps, warmup)
        return ms, len(case.data.train_x_all), model

    np.random.seed(0); torch.manual_seed(0)
    c1 = cases.Burgers(NumDomain=2540, layer_size=[2] + [20] * 3 + [1], soft_constraints=True)
    ms, n, _ = train_ms(c1)
    out["cfg1 Burgers 2-[20]x3-1, 2540+80+160 pts"] = {"ms_per_step": ms, "points_per_s": n / ms * 1e3}

    np.random.seed(0); torch.manual_seed(0)
    c2 = cases.AllenCahn(NumDomain=10000, layer_size=[2] + [64] * 4 + [1])
    ms, n, m2 = train_ms(c2)
    out["cfg2 AllenCahn 2-[64]x4-1, 10k pts (output transform)"] = {"ms_per_step": ms, "points_per_s": n / ms * 1e3}
    rad = algorithm.RAD(c2, iter=1)
    c2.compile(rad.model, "adam", lr=1e-3)
    rad.resample()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        rad.resample()
    torch.cuda.synchronize()
    out["cfg2 RAD round (100k candidates -> 10k, probe + p + choice)"] = {"ms_per_round": (time.perf_counter() - t0) / 5 * 1e3}
    X = c2.geomtime.random_points(100000)
    ms = time_steps(lambda: rad.model.predict_on_device(X, c2.pde), 20, 3)
    out["cfg2 residual probe 100k candidates (forward-only jets)"] = {"ms": ms, "points_per_s": 1e5 / ms * 1e3}

    np.random.seed(0); torch.manual_seed(0)
    c3 = cases.NS_Flow_in_LidDriven_Cavity(NumDomain=2**18, layer_size=[2] + [128] * 5 + [3], train_distribution="pseudo")
    ms, n, _ = train_ms(c3, steps=20, warmup=5)
    out["cfg3 NS cavity 2-[128]x5-3, 2^18 pts, 11 loss terms"] = {"ms_per_step": ms, "points_per_s": n / ms * 1e3}

    np.random.seed(0); torch.manual_seed(0)
    c4 = cases.Helmholtz(NumDomain=2**20, layer_size=[2] + [256] * 6 + [1], activation="tanh", honor_num_domain=True)
    ms, n, _ = train_ms(c4, steps=10, warmup=3)
    out["cfg4 Helmholtz 2-[256]x6-1 tanh, 2^20 pts"] = {"ms_per_step": ms, "points_per_s": n / ms * 1e3}
    r3 = algorithm.R3Sampling(c4, max_iter=1)
    c4.compile(r3.model, "adam", lr=1e-3)
    r3.resample()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        r3.resample()
    torch.cuda.synchronize()
    out["cfg4 R3 round (2^20 population: probe + retain + refill, host bookkeeping incl.)"] = {"ms_per_round": (time.perf_counter() - t0) / 3 * 1e3}
    del c4, r3
    torch.cuda.empty_cache()

    np.random.seed(0); torch.manual_seed(0)
    c4s = cases.Helmholtz(NumDomain=2**18, layer_size=[2] + [150] * 3 + [1], honor_num_domain=True)   # reference default net: sin, width 150
    ms, n, _ = train_ms(c4s, steps=10, warmup=3)
    out["Helmholtz reference-default net 2-[150]x3-1 sin, 2^18 pts (FP32 engine)"] = {"ms_per_step": ms, "points_per_s": n / ms * 1e3}

    # cfg5: FBPINN, K = 64 sub-nets 1-[64]x3-1, all nets on every batch
    np.random.seed(0); torch.manual_seed(0)
    c5 = cases.Poisson_1D_Dirichlet(NumDomain=2**14, layer_size=[1] + [64] * 3 + [1])
    fb = algorithm.FBPINN(c5, segment=64, iter=1)
    model, pde = fb.model_list[3], fb.domains.pdeCase_list[3]
    fb.compile(pde, model, "adam", lr=1e-3)
    X = model.data.train_x
    ms = time_steps(lambda: model.train_step(X, None, None), 10, 3)
    out["cfg5 FBPINN K=64 (sub-case default nets), one subdomain step, %d pts" % len(model.data.train_x_all)] = {
        "ms_per_step": ms, "net_evals_per_s": 64 * len(model.data.train_x_all) / ms * 1e3}
    big = torch.rand(2**17, 1, device="cuda")
    with torch.no_grad():
        ms = time_steps(lambda: fb.get_predict(big, param_grad=False), 10, 3)
    out["cfg5 FBPINN K=64 batched forward, 2^17 pts (one launch for all nets)"] = {"ms": ms, "net_evals_per_s": 64 * 2**17 / ms * 1e3}

    if args.cpu:
        from oracle import cases as oc, ddelite

        ddelite.config.set_default_float("float64")
        torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
        for cfg, nd in (("cfg1", 2540), ("cfg2", 10000), ("cfg3", 2048), ("cfg4", 2048)):
            np.random.seed(0); torch.manual_seed(0)
            data, net, _ = oc.build(cfg, nd)
            model = ddelite.Model(data, net)
            model.compile("adam", lr=1e-3, loss_weights=oc.LOSS_WEIGHTS.get(cfg))
            Xc = data.train_x
            for _ in range(2):
                model.train_step(Xc, None, None)
            t0 = time.perf_counter()
            k = 5
            for _ in range(k):
                model.train_step(Xc, None, None)
            dt = (time.perf_counter() - t0) / k
            out[f"CPU oracle {cfg} ({nd} pts, float64, {torch.get_num_threads()} threads)"] = {
                "ms_per_step": dt * 1e3, "points_per_s": len(data.train_x_all) / dt}
    for k, v in out.items():
        print(json.dumps({k: v}))


if __name__ == "__main__":
    main()
