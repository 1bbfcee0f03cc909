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


# ------------------------------------------------------------------------------------------------
# Point-sharded resampling: probe sharded over ranks, block residuals all-gathered in rank
# order, selection replicated -> the global index set / threshold equal the single-process ones.
# The device kernels are replaced by their CPU restatements (oracle/sampling.py) for this host-logic test;
# the kernels themselves are pinned bit-exact against the same restatements in tests/test_gpu_kernels.py.
# ------------------------------------------------------------------------------------------------
def _install_cpu_resample():
    from ai4xde_b200 import _lib
    from oracle import sampling

    def rad_probabilities(res, n, k, c):
        Y = res[:, :n].double().abs().sum(dim=0).numpy() if res.shape[0] > 1 else res[0, :n].double().abs().numpy()
        return torch.as_tensor(sampling.rad_probabilities(Y.reshape(-1, 1), k, c).reshape(-1))

    def choice_round(p, u, found, n_found, n_found_out, ws, wsb):
        # one retry round of numpy's legacy choice (mtrand.pyx): zero the found entries, cdf, searchsorted
        pn = p.numpy()
        pn[found[:n_found].numpy()] = 0
        cdf = np.cumsum(pn)
        cdf /= cdf[-1]
        new = cdf.searchsorted(u.numpy(), side="right")
        _, first = np.unique(new, return_index=True)
        first.sort()
        new = new[first]
        new = new[~np.isin(new, found[:n_found].numpy())][: found.numel() - n_found]
        found[n_found:n_found + len(new)] = torch.as_tensor(new)
        n_found_out[0] = n_found + len(new)

    def r3_abs_sum(res, n):
        return res[:, :n].abs().sum(dim=0)

    def pairwise_sum(a):
        return torch.as_tensor(sampling.pairwise_sum(a.numpy()))

    def r3_retain_thr(res_sum, thr, x_soa, n, place_at_end=True):
        keep = res_sum > thr
        nk = int(keep.sum())
        out = torch.zeros_like(x_soa)
        out[:, n - nk:n] = x_soa[:, :n][:, keep]
        return out, torch.tensor([nk])

    _lib.rad_probabilities, _lib.choice_round = rad_probabilities, choice_round
    _lib.r3_abs_sum, _lib.pairwise_sum, _lib.r3_retain_thr = r3_abs_sum, pairwise_sum, r3_retain_thr
    _lib.resample_ws = lambda n, dev: (torch.zeros(1), 0)


def _resample_case():
    import warnings

    warnings.simplefilter("ignore")
    import ai4xde_b200.dde as dde
    from ai4xde_b200 import cases as pc

    dde.config.set_default_float("float64")
    torch.manual_seed(0)
    np.random.seed(5)
    case = pc.Helmholtz(NumDomain=64, layer_size=[2, 16, 16, 1], activation="tanh")
    model = dde.Model(case.data, case.net)
    model.compile("sgd", lr=0.0, loss_weights=[1, 100])
    return case, model


def _resample_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _taylor_double as td

    td.install()
    _install_cpu_resample()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ai4xde_b200 import resample

    case, model = _resample_case()
    cand = np.random.RandomState(3).rand(999, 2)            # global candidates, identical on both ranks
    ids = resample.rad_select_sharded(model, cand, case.pde, 100, k=1, c=1, rand=np.random.RandomState(9).random_sample)
    pop = np.random.RandomState(4).rand(777, 2)             # global population, sharded unevenly by block_bounds
    lo, hi = resample.block_bounds(len(pop), rank, world)
    keep, thr, x_dev, nk = resample.r3_retain_sharded(model, np.ascontiguousarray(pop[lo:hi]), case.pde)
    torch.save({"ids": ids, "keep": keep, "thr": thr, "lo": lo, "hi": hi, "nk": nk,
                "kept": x_dev[:, (hi - lo) - nk:(hi - lo)].numpy()}, f"{out}.{rank}")
    dist.destroy_process_group()


def test_sharded_resampling_equals_single_process(tmp_path):
    import _taylor_double as td
    from ai4xde_b200 import _lib, resample

    out = str(tmp_path / "rs")
    mp.spawn(_resample_worker, args=(2, 29537, out), nprocs=2, join=True)
    r0, r1 = torch.load(f"{out}.0", weights_only=False), torch.load(f"{out}.1", weights_only=False)
    saved = {k: getattr(_lib, k) for k in ("rad_probabilities", "choice_round", "r3_abs_sum", "pairwise_sum",
                                           "r3_retain_thr", "resample_ws", "jet_forward", "jet_backward")}
    try:
        td.install()
        _install_cpu_resample()
        case, model = _resample_case()
        cand = np.random.RandomState(3).rand(999, 2)
        ids = resample.rad_select_sharded(model, cand, case.pde, 100, k=1, c=1,
                                          rand=np.random.RandomState(9).random_sample)
        pop = np.random.RandomState(4).rand(777, 2)
        keep, thr, x_dev, nk = resample.r3_retain_sharded(model, pop, case.pde)
    finally:
        for k, v in saved.items():
            setattr(_lib, k, v)
    # RAD: both ranks hold the single-process index set, in draw order
    assert np.array_equal(r0["ids"], ids) and np.array_equal(r1["ids"], ids)
    assert len(set(ids.tolist())) == 100
    # R3: one global threshold; the union of the block keeps is the single-process keep mask
    assert r0["thr"] == thr and r1["thr"] == thr
    assert np.array_equal(np.concatenate([r0["keep"], r1["keep"]]), keep)
    assert r0["nk"] + r1["nk"] == nk
    kept_single = pop[keep].T.astype(np.float32)
    assert np.array_equal(np.concatenate([r0["kept"], r1["kept"]], axis=1), kept_single)
