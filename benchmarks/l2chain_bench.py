#!/usr/bin/env python
"""A/B of the forward of cfg4 (2-[256]x6-1 tanh, Laplacian pattern, layers kept) inside ONE process: one launch per
hidden layer (k_tc_gemm2) vs the L2-chained kernel (k_tc_fwd_l2) at several group sizes, with and without the L2
eviction hints, with the tile release issued at once or deferred into the next item; configurations interleaved round-robin so that clock / power-cap drift hits them alike.
Usage: python benchmarks/l2chain_bench.py [n_points [default]]"""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from ai4xde_b200 import _lib
from ai4xde_b200.jet import JetSpec, to_soa


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2 ** 20
    dev = torch.device("cuda:0")
    widths = [2] + [256] * 6 + [1]
    gen = torch.Generator().manual_seed(0)
    js = JetSpec((0, 1), ((15, 15),))
    net = _lib.make_net_desc(widths, "tanh")
    chunks = []
    for i, o in zip(widths[:-1], widths[1:]):
        chunks += [torch.randn((1, o * i), generator=gen) * (2.0 / (i + o)) ** 0.5, torch.zeros((1, o))]
    params = torch.cat(chunks, dim=1).to(dev)
    xs = to_soa(torch.rand((n, 2), generator=gen).to(dev))
    configs = [("layer launches", 0, 2, 1, 0, 0), ("l2chain G=2 hints", 1, 2, 1, 0, 0),
               ("l2chain G=3 hints release-warp", 1, 3, 1, 2, 0), ("l2chain G=3 hints release-warp wres", 1, 3, 1, 2, 1),
               ("l2chain G=2 hints release-warp wres", 1, 2, 1, 2, 1), ("l2chain G=4 hints release-warp wres", 1, 4, 1, 2, 1),
               ("l2chain G=3 release-warp wres", 1, 3, 0, 2, 1)]
    if len(sys.argv) > 2 and sys.argv[2] == "default":   # ncu captures: the shipped configuration only
        configs = [("l2chain G=3 hints release-warp wres", 1, 3, 1, 2, 1)]
    times = {c[0]: [] for c in configs}
    for keep in (True, False):
        for rnd in range(6):
            for name, on, G, hints, defer, wres in configs:
                _lib.tc_set_option("l2chain_wres", wres)
                _lib.tc_set_option("l2chain_defer", defer)
                _lib.tc_set_option("l2chain", on)
                _lib.tc_set_option("l2chain_group", G)
                _lib.tc_set_option("l2chain_hints", hints)
                _lib.jet_forward(net, js.desc, params, xs, n, keep_all=keep, engine="tf32")
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    _lib.jet_forward(net, js.desc, params, xs, n, keep_all=keep, engine="tf32")
                e1.record()
                torch.cuda.synchronize()
                if rnd > 0:
                    times[name].append(e0.elapsed_time(e1) / 3)
        print(json.dumps({"points": n, "layers_kept": keep,
                          "forward_ms_median": {k: round(statistics.median(v), 3) for k, v in times.items()},
                          "forward_ms_min": {k: round(min(v), 3) for k, v in times.items()}}))
        times = {c[0]: [] for c in configs}
    _lib.tc_set_option("l2chain", 1)
    _lib.tc_set_option("l2chain_group", 3)
    _lib.tc_set_option("l2chain_hints", 1)
    _lib.tc_set_option("l2chain_defer", 2)
    _lib.tc_set_option("l2chain_wres", 1)


if __name__ == "__main__":
    main()
