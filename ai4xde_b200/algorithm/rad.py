"""RAD: residual-based adaptive distribution (/root/reference/AI4XDE/algorithm/RAD.py:7-42).
Candidate residuals, probabilities and the without-replacement draw run on the device
(ai4xde_b200/resample.py); the candidate points and the uniforms come from NumPy's global legacy
stream in the reference's order (candidates first, then one `random_sample` per retry round), so the
selected set equals `np.random.choice(..., replace=False, p=...)` on the same residuals."""
import numpy as np
import torch

from .. import dde, resample
from ..solver import PINNSolver


def device_round(solver, case, n_candidates, num, replace):
    """RAD / RAR_D round on the device ([ext] device_rng): None if the geometry is not a box the sampler can fill.
    `replace`: the selected points become the anchor set (RAD.py:40) or are added to it (RAR_D.py:36).  The host arrays
    of the reference API come back by ONE device-to-host copy of the selected points; the new `train_x_all` / `train_x`
    get their device mirrors registered, so the next train step uploads nothing."""
    from .r3sampling import _box_limits

    box = _box_limits(case.geomtime)
    if box is None or not torch.cuda.is_available():
        return None
    model, data = solver.model, case.data
    ids, sel, aos = resample.rad_round_on_device(model, box, case.pde, n_candidates, num, solver.k, solver.c,
                                                 solver.device_seed, solver._rng_offset)
    solver._rng_offset += n_candidates
    picked = aos.cpu().numpy()
    if replace:
        data.replace_with_anchors(picked)
        all_dev = aos
    else:
        old = model.device_batch(data.train_x_all).aos
        data.add_anchors(picked)
        all_dev = torch.cat((aos, old), dim=0)
    model.adopt_device_batch(data.train_x_all, all_dev, soa=sel if replace else None)
    n_bc = int(sum(data.num_bcs)) if data.num_bcs else 0
    if data.train_x is not data.train_x_all:
        bc_dev = torch.as_tensor(np.ascontiguousarray(data.train_x[:n_bc], dtype=np.float32)).to(all_dev.device)
        model.adopt_device_batch(data.train_x, torch.cat((bc_dev, all_dev), dim=0) if n_bc else all_dev)
    return ids.cpu().numpy()


class RAD(PINNSolver):
    def __init__(self, PDECase, iter=100, k=1, c=1, warmup_iter=1000, fast=False, device_rng=False, device_seed=0):
        # [ext] fast: parallel float64 sum / cumsum in the selection instead of NumPy's sequential order
        # [ext] device_rng: the whole round on the device -- candidates from the library's Philox box sampler, draw
        # uniforms from the device generator, selected points gathered in place (performance mode; the default keeps
        # NumPy's candidate and uniform streams, so the index sets equal the reference's bit for bit)
        self.fast = bool(fast)
        self.device_rng, self.device_seed, self._rng_offset = bool(device_rng), int(device_seed), 0
        dde.optimizers.config.set_LBFGS_options(maxiter=1000)
        super().__init__(name=f"RAD_k_{k}_c_{c}", PDECase=PDECase)
        self.iter, self.k, self.c, self.warmup_iter = iter, k, c, warmup_iter

    def resample(self):
        case = self.PDECase
        _, rank, world = resample._world()
        if world == 1 and self.device_rng:
            done = device_round(self, case, case.NumDomain * 10, case.NumDomain, replace=True)
            if done is not None:
                return done
        if world == 1:
            X = case.geomtime.random_points(case.NumDomain * 10)
            ids = resample.rad_select(self.model, X, case.pde, case.NumDomain, k=self.k, c=self.c, fast=self.fast)
            case.data.replace_with_anchors(X[ids])
            return ids
        # point-sharded run: NumDomain is this rank's share.  Every rank draws the SAME
        # global candidate set and uniforms (shared stream), probes its own block, and keeps its block of
        # the globally selected points -- the selection equals the single-process one on world * NumDomain.
        if getattr(self, "_shared", None) is None:
            self._shared = resample.shared_numpy_stream()
        with self._shared:
            X = case.geomtime.random_points(case.NumDomain * 10 * world)
            ids = resample.rad_select_sharded(self.model, X, case.pde, case.NumDomain * world, k=self.k, c=self.c)
        lo, hi = resample.block_bounds(len(ids), rank, world)
        case.data.replace_with_anchors(X[ids[lo:hi]])
        return ids

    def closure(self):
        self.train_step(iterations=self.warmup_iter)
        for _ in range(self.iter):
            self.resample()
            self.train_step(iterations=1000)
