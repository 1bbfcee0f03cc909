"""BASELINE config 5 on one B200: FBPINN, K = 64 subdomain nets 1-[64]x3-1, windowed sum, 1-D multiscale Poisson.

  (a) the bank forward alone (value + u_x + u_xx of all K nets on n points): whole-net kernel (csrc/tc_bank.cu)
      against the layer-by-layer tensor-core kernels (AX_TC_BANK=0, child process) and the FP32 engine;
  (b) one optimiser step of sub-model k (forward of all nets, window sum, residual, loss, backward through the
      active net, Adam) through the patched closure of algorithm.FBPINN.
Prints one JSON object.  python benchmarks/bank_bench.py [--points N] [--nets K]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def time_ms(fn, reps, warm=2):
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


def forward_only(K, n, engine):
    import torch
    from ai4xde_b200 import _lib
    from ai4xde_b200.jet import JetSpec, to_soa

    dev = torch.device("cuda:0")
    widths = [1, 64, 64, 64, 1]
    js = JetSpec((0,), ((0, 0),))
    gen = torch.Generator().manual_seed(0)
    P = sum(i * o + o for i, o in zip(widths[:-1], widths[1:]))
    params = (torch.randn((K, P), generator=gen) * 0.2).to(dev)
    xs = to_soa(torch.rand((n, 1), generator=gen).to(dev))
    # the FP32 engine keeps a two-layer stash per net: bound the group size as NetBank.evaluate does
    per_net = max(_lib.stash_bytes(_lib.make_net_desc(widths, "tanh"), js.desc, xs.shape[1], False, engine), 1)
    group = max(1, min(K, (4 << 30) // per_net))

    def run():
        for k0 in range(0, K, group):
            k1 = min(K, k0 + group)
            net = _lib.make_net_desc(widths, "tanh", n_nets=k1 - k0)
            _lib.jet_forward(net, js.desc, params[k0:k1], xs, n, keep_all=False, engine=engine)

    ms = time_ms(run, 5)
    flop = K * n * 2.0 * (64 + 3 * (2 * 64 * 64 + 64))
    return {"ms": ms, "net_evals_per_s": K * n / (ms * 1e-3), "tflops": flop / (ms * 1e-3) / 1e12, "nets_per_call": group}


def fbpinn_step(K, n, engine):
    import numpy as np
    import torch
    import ai4xde_b200.dde as dde
    from ai4xde_b200 import algorithm
    from ai4xde_b200.cases.poisson_cases import multiscale_fbpinn_case

    np.random.seed(0)
    torch.manual_seed(0)
    dde.config.set_default_float("float32")
    dde.config.set_engine(engine)
    solver = algorithm.FBPINN(multiscale_fbpinn_case(NumDomain=n), segment=K, iter=1)
    k = K // 2
    model, pde = solver.model_list[k], solver.domains.pdeCase_list[k]
    solver.compile(pde, model, "adam", lr=1e-3)
    X = model.data.train_x
    ms = time_ms(lambda: model.train_step(X, None, None), 5, warm=3)
    dde.config.set_engine("auto")
    return {"ms_per_step": ms, "points_per_s": n / (ms * 1e-3), "net_evals_per_s": K * n / (ms * 1e-3),
            "loss": float(model.last_total_loss)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=2**17)
    ap.add_argument("--nets", type=int, default=64)
    ap.add_argument("--child", default=None)
    a = ap.parse_args()
    if a.child:
        print(json.dumps(forward_only(a.nets, a.points, a.child)))
        return
    out = {"nets": a.nets, "points": a.points}
    out["forward_bank_kernel_tf32"] = forward_only(a.nets, a.points, "tf32")
    env = dict(os.environ, AX_TC_BANK="0")
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--points", str(a.points), "--nets", str(a.nets),
                        "--child", "tf32"], env=env, capture_output=True, text=True)
    out["forward_layer_by_layer_tf32"] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else r.stderr[-400:]
    out["forward_fp32_engine"] = forward_only(a.nets, a.points, "fp32")
    out["fbpinn_step_tf32"] = fbpinn_step(a.nets, a.points, "tf32")
    out["fbpinn_step_fp32"] = fbpinn_step(a.nets, a.points, "fp32")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
