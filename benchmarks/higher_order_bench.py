#!/usr/bin/env python
"""Step times of the round's late additions on the fused path, with the same problem on the oracle
(ddelite + nested torch.autograd, float64, host cores) beside them: gPINN on the hard-constraint Burgers case
(two-direction third-order pattern; the reference's own gPINN example, algorithm/gPINN.py:77-81), a fourth-order
beam residual on a 1-[20]x3-1 net (cases/PDECases.py:428-500), and the STMsFFN wave net (cases/WaveCases.py:77).
Usage: python benchmarks/higher_order_bench.py"""
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

from configs_bench import time_steps


def beam_problem(d, n):
    geom = d.geometry.Interval(0, 1)

    def pde(x, y):
        return d.grad.hessian(d.grad.hessian(y, x), x) + 1

    left = lambda x, on: on and np.isclose(x[0], 0)
    right = lambda x, on: on and np.isclose(x[0], 1)
    bcs = [d.icbc.DirichletBC(geom, lambda x: 0, left), d.icbc.NeumannBC(geom, lambda x: 0, left),
           d.icbc.OperatorBC(geom, lambda x, y, _: d.grad.hessian(y, x), right),
           d.icbc.OperatorBC(geom, lambda x, y, _: d.grad.jacobian(d.grad.hessian(y, x), x), right)]
    data = d.data.PDE(geom, pde, bcs, num_domain=n, num_boundary=2, train_distribution="pseudo")
    return data, d.nn.FNN([1] + [20] * 3 + [1], "tanh", "Glorot uniform")


def wave_problem(d, n):
    geomtime = d.geometry.GeometryXTime(d.geometry.Interval(0, 1), d.geometry.TimeDomain(0, 1))
    sol = lambda x: np.sin(np.pi * x[:, 0:1]) * np.cos(2 * np.pi * x[:, 1:2])

    def pde(x, y):
        return d.grad.hessian(y, x, i=1, j=1) - 4 * d.grad.hessian(y, x, i=0, j=0)

    bcs = [d.icbc.DirichletBC(geomtime, sol, lambda _, on: on), d.icbc.IC(geomtime, sol, lambda _, on: on)]
    data = d.data.TimePDE(geomtime, pde, bcs, num_domain=n, num_boundary=360, num_initial=360, train_distribution="pseudo")
    net = d.nn.STMsFFN([2] + [100] * 3 + [1], "tanh", "Glorot uniform", sigmas_x=[1], sigmas_t=[1, 10])
    net.apply_feature_transform(lambda x: (x - 0.5) * 2 * np.sqrt(3))
    return data, net


def device_ms(model, data, steps=50, warmup=10):
    X = data.train_x
    aux = getattr(data, "train_aux_vars", None)
    return time_steps(lambda: model.train_step(X, None, aux), steps, warmup)


def oracle_ms(model, data, steps=3):
    X, y, aux = data.train_next_batch()
    model.train_step(X, y, aux)
    t0 = time.perf_counter()
    for _ in range(steps):
        model.train_step(X, y, aux)
    return (time.perf_counter() - t0) / steps * 1e3


BIG = int(os.environ.get("AX_BENCH_POINTS", 1 << 17))


def main():
    import ai4xde_b200.dde as dde
    from ai4xde_b200 import algorithm, cases
    from oracle import ddelite

    out = {"cpu_threads": torch.get_num_threads()}
    dde.config.set_default_float("float32")
    dde.config.set_engine("auto")
    np.random.seed(0); torch.manual_seed(0)
    for n in (2540, BIG):
        case = cases.Burgers(NumDomain=n, layer_size=[2] + [20] * 3 + [1])
        solver = algorithm.gPINN(case)
        case.compile(solver.model, "adam", lr=1e-3)
        case.data.train_next_batch()
        ms = device_ms(solver.model, case.data)
        out[f"gPINN Burgers 2-[20]x3-1 third order, {n} pts"] = {"ms_per_step": round(ms, 3), "points_per_s": round(n / ms * 1e3)}
    for name, build, n in (("beam u_xxxx 1-[20]x3-1 fourth order", beam_problem, BIG),
                           ("STMsFFN wave 2-[100]x3-1, sigmas_x [1], sigmas_t [1, 10]", wave_problem, BIG // 2)):
        data, net = build(dde, n)
        model = dde.Model(data, net)
        model.compile("adam", lr=1e-3)
        data.train_next_batch()
        ms = device_ms(model, data, steps=20, warmup=5)
        rec = {"ms_per_step": round(ms, 3), "points_per_s": round(n / ms * 1e3)}
        ns = 4096
        ddelite.config.set_default_float("float64")
        try:
            odata, onet = build(ddelite, ns)
            omodel = ddelite.Model(odata, onet)
            omodel.compile("adam", lr=1e-3)
            oms = oracle_ms(omodel, odata)
            rec["oracle_cpu_f64"] = {"ms_per_step": round(oms, 2), "points_per_s": round(ns / oms * 1e3), "sample_points": ns}
        finally:
            ddelite.config.set_default_float("float32")
            torch.set_default_dtype(torch.float32)
        out[f"{name}, {n} pts"] = rec
    print(json.dumps(out))


if __name__ == "__main__":
    main()
