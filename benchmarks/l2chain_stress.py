#!/usr/bin/env python
"""Stress of the L2-chained forward's hand-off (release warp, resident weights): the shipped configuration is run many
times on cfg4-sized inputs and every stash layer and jet is compared bit for bit with the one-launch-per-layer path.
A lost or early hand-off reads rows that are not written yet and shows up as a mismatch.
Usage: python benchmarks/l2chain_stress.py [repetitions]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from ai4xde_b200 import _lib
from ai4xde_b200.jet import JetSpec, to_soa


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    dev = torch.device("cuda:0")
    widths = [2] + [256] * 6 + [1]
    gen = torch.Generator().manual_seed(1)
    js = JetSpec((0, 1), ((15, 15),))
    net = _lib.make_net_desc(widths, "tanh")
    chunks = []
    for i, o in zip(widths[:-1], widths[1:]):
        chunks += [torch.randn((1, o * i), generator=gen) * (2.0 / (i + o)) ** 0.5, torch.zeros((1, o))]
    params = torch.cat(chunks, dim=1).to(dev)
    report = {"repetitions": reps, "sizes": {}}
    for n in (2 ** 20 + 80, 777777, 2 ** 19 + 13):
        xs = to_soa((torch.rand((n, 2), generator=gen) * 2 - 1).to(dev))
        _lib.tc_set_option("l2chain", 0)
        jets0, stash0 = _lib.jet_forward(net, js.desc, params, xs, n, keep_all=True, engine="tf32")
        jets0, s0 = jets0.clone(), stash0.view(torch.float32).clone()
        _lib.tc_set_option("l2chain", 1)
        ldn = xs.shape[1]
        layer, rows = ldn * js.S * 256, n * js.S * 256
        bad = 0
        for rep in range(reps):
            jets1, stash1 = _lib.jet_forward(net, js.desc, params, xs, n, keep_all=True, engine="tf32")
            s1 = stash1.view(torch.float32)
            ok = torch.equal(jets1[..., :n], jets0[..., :n])
            for l in range(6):
                ok = ok and torch.equal(s1[l * layer : l * layer + rows], s0[l * layer : l * layer + rows])
            bad += not ok
        report["sizes"][str(n)] = {"mismatching_runs": bad}
        del s0, stash0, stash1, s1
        torch.cuda.empty_cache()
    print(json.dumps(report))
    sys.exit(1 if any(v["mismatching_runs"] for v in report["sizes"].values()) else 0)


if __name__ == "__main__":
    main()
