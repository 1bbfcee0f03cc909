"""Data-parallel semantics on CPU (world_size 2, gloo): each rank owns a (deliberately unequal) shard
of the collocation and boundary points, per-term losses are weighted by n_local * world / n_global,
the flat gradient is all-reduced (mean) inside dde.Model's closure, and the result equals the
single-process gradient on the concatenated point set (sums and global counts)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(points_all, rank=None, world=1):
    import warnings

    warnings.simplefilter("ignore")
    import ai4xde_b200.dde as dde
    from ai4xde_b200 import cases as pc

    dde.config.set_default_float("float64")
    torch.manual_seed(0)
    case = pc.Helmholtz(layer_size=[2, 16, 16, 1], activation="tanh")
    bdry = points_all[:16]
    dom = points_all[16:]
    if rank is not None:
        # unequal shards on purpose: the exchange must weight by global counts, not average means
        bdry = bdry[:6] if rank == 0 else bdry[6:]
        dom = dom[:40] if rank == 0 else dom[40:]
    X = np.vstack((bdry, dom))
    data = dde.data.PDE(case.geomtime, case.pde, case.data.bcs, 0, 0, anchors=X, train_distribution="pseudo")
    model = dde.Model(data, case.net)
    model.compile("sgd", lr=0.0, loss_weights=[1, 100])
    return model, data, case


def _points():
    rng = np.random.RandomState(0)
    bdry = np.stack([rng.rand(16), np.zeros(16)], 1)         # on the boundary y = 0
    return np.vstack((bdry, rng.rand(128, 2) * 0.98 + 0.01))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _taylor_double as td

    td.install()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model, data, case = _build(_points(), rank, world)
    model.train_step(data.train_x, None, None)               # closure: backward + all-reduce(mean)
    g = torch.cat([p.grad.reshape(-1) for p in case.net.parameters()])
    if rank == 0:
        torch.save(g, out)
    dist.destroy_process_group()


def test_two_rank_gradient_equals_single_process(tmp_path):
    import _taylor_double as td

    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, 29533, out), nprocs=2, join=True)
    g2 = torch.load(out)
    td.install()
    model, data, case = _build(_points())
    model.train_step(data.train_x, None, None)
    g1 = torch.cat([p.grad.reshape(-1) for p in case.net.parameters()])
    assert float((g1 - g2).abs().max() / g1.abs().max()) < 1e-5
