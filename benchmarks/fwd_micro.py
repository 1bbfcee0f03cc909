"""Micro-benchmark of the tensor-core forward alone (cfg4 net): keep_all (training stash) and probe mode.
Run with AX_TC_CHAIN=0/1 to compare the layer-by-layer kernels with the whole-stack kernel."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ai4xde_b200 import _lib
from ai4xde_b200.jet import JetSpec, to_soa

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2**20
dev = torch.device("cuda:0")
widths = [2] + [256] * 6 + [1]
js = JetSpec((0, 1), ((15, 15),))
net = _lib.make_net_desc(widths, "tanh")
gen = torch.Generator().manual_seed(0)
P = _lib.param_count(net)
params = (torch.randn((1, P), generator=gen) * 0.06).to(dev)
x = to_soa(torch.rand((n, 2), generator=gen).to(dev))
for keep in (True, False):
    for _ in range(3):
        out = _lib.jet_forward(net, js.desc, params, x, n, keep_all=keep, engine="tf32")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = _lib.jet_forward(net, js.desc, params, x, n, keep_all=keep, engine="tf32")
    e1.record()
    torch.cuda.synchronize()
    print("AX_TC_CHAIN=%s keep_all=%s n=%d: %.3f ms per forward" % (os.environ.get("AX_TC_CHAIN", "0"), keep, n, e0.elapsed_time(e1) / 10))
