"""One forward (with stash) and one backward of the narrow-net kernels on 2^20 points: the ncu target
(ncu -k regex:k_narrow --set full python benchmarks/narrow_one.py [width])."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ai4xde_b200 import _lib
from ai4xde_b200.jet import JetSpec, to_soa

H = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << 20
dev = torch.device("cuda:0")
js = JetSpec((1, 0), ((0, 0),))
widths = [2, H, H, H, 1]
net = _lib.make_net_desc(widths, "tanh")
torch.manual_seed(0)
params = (torch.randn((1, _lib.param_count(net))) * 0.3).to(dev)
xs = to_soa((torch.rand((n, 2)) * 2 - 1).to(dev))
for _ in range(2):
    jets, stash = _lib.jet_forward(net, js.desc, params, xs, n, keep_all=True, engine="fp32")
    jb = torch.ones_like(jets)
    g = _lib.jet_backward(net, js.desc, params, xs, n, jb, stash, engine="fp32")
torch.cuda.synchronize()
print("ok", float(g.abs().sum()))
