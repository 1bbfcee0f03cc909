"""Per-kernel time of the FP32 engine on one residual + backward evaluation (the library's CUDA-event timing
hooks): where a step of a narrow / sin / in-between-width net goes.  Usage: python benchmarks/fp32_engine_breakdown.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ai4xde_b200.dde as dde
from ai4xde_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
dde.config.set_default_float("float32")
dde.config.set_engine("fp32")
torch.manual_seed(0)
out = {"points": n}
for widths, act in (([2] + [100] * 4 + [2], "tanh"), ([2] + [150] * 3 + [1], "sin"), ([2] + [20] * 3 + [1], "tanh"),
                    ([2] + [64] * 4 + [1], "tanh")):
    net = dde.nn.FNN(widths, act, "Glorot normal")
    x = (torch.rand((n, 2)) * 2 - 1).to("cuda").requires_grad_()

    def step():
        y = net(x)
        r = dde.grad.hessian(y, x, component=0, i=0, j=0) if widths[-1] > 1 else dde.grad.hessian(y, x, i=0, j=0)
        r = r + dde.grad.jacobian(y, x, i=0, j=1)
        dde.grad.clear()
        net.zero_grad()
        (r ** 2).mean().backward()

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    _lib.timing_report()
    _lib.timing_enable(True)
    beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    beg.record()
    for _ in range(5):
        step()
    end.record()
    torch.cuda.synchronize()
    _lib.timing_enable(False)
    rep = _lib.timing_report()
    rec = {"ms_per_step_with_hooks": round(beg.elapsed_time(end) / 5, 3)}
    for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"]):
        rec[k] = {"launches_per_step": v["launches"] // 5, "ms_per_step": round(v["ms"] / 5, 3),
                  "tflops": round(v["flops"] / v["ms"] / 1e9, 2) if v["ms"] > 0 else None,
                  "gbs": round(v["bytes"] / v["ms"] / 1e6, 1) if v["ms"] > 0 else None}
    out["x".join(map(str, widths)) + "_" + act] = rec
print(json.dumps(out, indent=1))
