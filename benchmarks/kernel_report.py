"""Per-kernel CUDA-event report (ax_timing_* hooks) of one BASELINE config's train step.
Usage: python benchmarks/kernel_report.py cfg3 [points]"""
import os, sys, json, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
warnings.simplefilter("ignore")
import numpy as np, torch
import ai4xde_b200.dde as dde
from ai4xde_b200 import _lib, cases

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
np.random.seed(0); torch.manual_seed(0)
dde.config.set_default_float("float64")
if cfg == "cfg3":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2**18
    case = cases.NS_Flow_in_LidDriven_Cavity(NumDomain=n, layer_size=[2] + [128] * 5 + [3], honor_num_domain=True) \
        if "honor_num_domain" in cases.NS_Flow_in_LidDriven_Cavity.__init__.__code__.co_varnames else \
        cases.NS_Flow_in_LidDriven_Cavity(NumDomain=n, layer_size=[2] + [128] * 5 + [3])
elif cfg == "cfg2":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    case = cases.AllenCahn(NumDomain=n, layer_size=[2] + [64] * 4 + [1])
else:
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2**20
    case = cases.Helmholtz(NumDomain=n, layer_size=[2] + [256] * 6 + [1], activation="tanh", honor_num_domain=True)
dde.config.set_cuda_graphs(False) if hasattr(dde.config, "set_cuda_graphs") else None
model = dde.Model(case.data, case.net)
case.compile(model, "adam", lr=1e-3)
X = case.data.train_x
for _ in range(5):
    model.train_step(X, None, None)
torch.cuda.synchronize()
_lib.timing_enable(True); _lib.timing_report()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 20
for _ in range(K):
    model.train_step(X, None, None)
e1.record(); torch.cuda.synchronize()
rep = _lib.timing_report(); _lib.timing_enable(False)
ms = e0.elapsed_time(e1) / K
print(cfg, "points", len(case.data.train_x_all), "ms/step", round(ms, 3))
tot = 0
for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"]):
    per = v["ms"] / K
    tot += per
    print(f"  {k:18s} {v['launches']//K:3d} launches/step  {per:7.3f} ms/step  {v['bytes']/(v['ms']*1e-3)/1e9 if v['ms']>0 else 0:7.0f} GB/s  {v['flops']/(v['ms']*1e-3)/1e12 if v['ms']>0 else 0:6.1f} TFLOP/s")
print("  kernels total", round(tot, 3), "ms; other", round(ms - tot, 3))
