"""Two launches of the FBPINN windowed bank forward (K = 64 nets 1-[64]x3-1, 2^17 points): the ncu target
(ncu -k regex:k_tc_bank --set full python benchmarks/bank_one.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ai4xde_b200 import _lib
from ai4xde_b200.jet import JetSpec, to_soa

K, n = 64, 1 << 17
dev = torch.device("cuda:0")
js = JetSpec((0,), ((0, 0),))
widths = [1, 64, 64, 64, 1]
net = _lib.make_net_desc(widths, "tanh", n_nets=K)
torch.manual_seed(0)
params = (torch.randn((K, _lib.param_count(_lib.make_net_desc(widths, "tanh")))) * 0.2).to(dev)
xs = to_soa(torch.rand((n, 1)).to(dev))
edges = torch.linspace(0, 1, K + 1)
lo, hi = edges[:-1].contiguous().to(dev), edges[1:].contiguous().to(dev)
lst = torch.arange(K, dtype=torch.int32, device=dev)
for _ in range(2):
    out = _lib.fbpinn_window_forward(net, js.desc, params, xs, n, lo, hi, lst, engine="tf32")
torch.cuda.synchronize()
print("ok", float(out.abs().sum()))
