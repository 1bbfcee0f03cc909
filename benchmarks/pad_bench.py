"""Times one residual + backward evaluation of an FNN whose hidden width lies between the tensor-core engine's
compiled widths (the reference's Schrodinger net 2-[100]x4-2 tanh, cases/PDECases.py:1833) on the FP32 engine
and on its zero-padded 128-wide tensor-core image (engine 'auto').  Usage: python benchmarks/pad_bench.py [n_points]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ai4xde_b200.dde as dde

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
dde.config.set_default_float("float32")
torch.manual_seed(0)
out = {"points": n}
for widths in ([2] + [100] * 4 + [2], [2] + [150] * 3 + [1], [2] + [50] * 4 + [3]):
    net = dde.nn.FNN(widths, "tanh", "Glorot normal")
    x = (torch.rand((n, 2)) * 2 - 1).to("cuda").requires_grad_()
    for eng in ("fp32", "auto", "tf32"):
        dde.config.set_engine(eng)

        def step():
            y = net(x)
            r = dde.grad.hessian(y, x, component=0, i=0, j=0) if widths[-1] > 1 else dde.grad.hessian(y, x, i=0, j=0)
            r = r + dde.grad.jacobian(y, x, i=0, j=1)
            dde.grad.clear()
            net.zero_grad()
            (r ** 2).mean().backward()

        for _ in range(3):
            step()
        beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        beg.record()
        for _ in range(10):
            step()
        end.record()
        torch.cuda.synchronize()
        ms = beg.elapsed_time(end) / 10
        out[f"{'x'.join(map(str, widths))}_{eng}"] = {"ms_per_step": round(ms, 3), "points_per_s": round(n / ms * 1e3)}
dde.config.set_engine("auto")
print(json.dumps(out))
