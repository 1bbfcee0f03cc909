"""Narrow nets on one B200: the whole-net FP32 kernels (csrc/narrow.cuh) against the layer-by-layer FP32 kernels
(AX_NARROW=0, child process), kernel-only (forward with stash + backward) and through a full residual step
(u_t + u u_x - nu u_xx, torch residual algebra, Adam excluded), with the measured FP32-FMA peak beside them.
python benchmarks/narrow_bench.py [--points N]"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

NETS = [([2, 20, 20, 20, 1], "tanh"), ([2, 32, 32, 32, 1], "tanh"), ([2, 20, 20, 20, 1], "sin")]


def time_ms(fn, reps=5, warm=2):
    import torch

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run(n):
    import torch
    import ai4xde_b200.dde as dde
    from ai4xde_b200 import _lib
    from ai4xde_b200.jet import JetSpec, to_soa

    dde.config.set_default_float("float32")
    dde.config.set_engine("fp32")
    dev = torch.device("cuda:0")
    js = JetSpec((1, 0), ((0, 0),))     # u, u_t, u_x, u_xx  (S = 4)
    out = {}
    for widths, act in NETS:
        torch.manual_seed(0)
        net_desc = _lib.make_net_desc(widths, act)
        P = _lib.param_count(net_desc)
        params = (torch.randn((1, P)) * 0.3).to(dev)
        x = (torch.rand((n, 2)) * 2 - 1).to(dev)
        xs = to_soa(x)
        jets, stash = _lib.jet_forward(net_desc, js.desc, params, xs, n, keep_all=True, engine="fp32")
        jb = torch.randn(jets.shape, device=dev)
        f_ms = time_ms(lambda: _lib.jet_forward(net_desc, js.desc, params, xs, n, keep_all=True, engine="fp32"))
        p_ms = time_ms(lambda: _lib.jet_forward(net_desc, js.desc, params, xs, n, keep_all=False, engine="fp32"))
        b_ms = time_ms(lambda: _lib.jet_backward(net_desc, js.desc, params, xs, n, jb, stash, engine="fp32"))
        H, L = widths[1], len(widths) - 2
        f_fwd = 2.0 * (widths[0] * H + 4 * ((L - 1) * H * H + H * widths[-1]))
        net = dde.nn.FNN(widths, act, "Glorot normal")
        xr = x.clone().requires_grad_()

        def step():
            y = net(xr)
            r = dde.grad.jacobian(y, xr, i=0, j=1) + y * dde.grad.jacobian(y, xr, i=0, j=0) \
                - 0.01 * dde.grad.hessian(y, xr, i=0, j=0)
            dde.grad.clear()
            net.zero_grad()
            (r ** 2).mean().backward()

        s_ms = time_ms(step)
        out["x".join(map(str, widths)) + "_" + act] = {
            "forward_with_stash_ms": f_ms, "forward_probe_ms": p_ms, "backward_ms": b_ms,
            "kernels_points_per_s": n / ((f_ms + b_ms) * 1e-3),
            "kernels_tflops_algorithmic": 3.0 * f_fwd * n / ((f_ms + b_ms) * 1e-3) / 1e12,
            "probe_points_per_s": n / (p_ms * 1e-3),
            "residual_step_ms": s_ms, "residual_step_points_per_s": n / (s_ms * 1e-3),
            "flop_per_point_train": 3.0 * f_fwd,
        }
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=1 << 20)
    ap.add_argument("--child", action="store_true")
    a = ap.parse_args()
    if a.child:
        print(json.dumps(run(a.points)))
        return
    from ai4xde_b200 import _lib

    res = {"points": a.points, "fp32_fma_peak_tflops_measured": _lib.measure_fp32_fma_peak()}
    res["narrow_kernels"] = run(a.points)
    env = dict(os.environ, AX_NARROW="0")
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--points", str(a.points), "--child"], env=env,
                       capture_output=True, text=True)
    res["layer_by_layer_kernels"] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else r.stderr[-500:]
    for k, v in res["narrow_kernels"].items():
        v["fraction_of_fp32_peak"] = v["kernels_tflops_algorithmic"] / res["fp32_fma_peak_tflops_measured"]
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
