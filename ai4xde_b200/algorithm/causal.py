"""CausalPINN: time-slab weights on the residual (/root/reference/AI4XDE/algorithm/CausalPINN.py:8-139).

The time axis is cut into `t_num` slabs; every round the residual is probed on `num_per_t` uniform points
per slab (forward-only jets on the device through `model.predict`), normalised like RAD's density
(`|r|^k / mean + c`, divided by its Python-order sum), averaged per slab, and slab i gets the weight
exp(-100 * sum of the slab errors before slab i-1) -- the reference's slice `err_per_t[: i - 1]`, kept as it is
(slabs 0 and 1 both weigh 1).  The weighted residual replaces the case's pde; the weight lookup is one
`bucketize` + gather on the device, half-open slabs [t_i, t_{i+1}) as in the reference (a point at exactly t1
weighs 0 there too)."""
import numpy as np
import torch

from .. import dde
from ..solver import PINNSolver

EPSILON = 100.0


class CausalPINN(PINNSolver):
    def __init__(self, PDECase, t_num=10, num_per_t=200, iter=10, warmup_iter=1000, k=1, c=0):
        self.t_num, self.num_per_t, self.iter, self.warmup_iter, self.k, self.c = t_num, num_per_t, iter, warmup_iter, k, c
        self.weight = None
        dde.optimizers.config.set_LBFGS_options(maxiter=1000)
        if not isinstance(PDECase.geomtime, dde.geometry.GeometryXTime):
            raise AssertionError("CausalPINN only support GeometryXTime")
        super().__init__(name="CausalPINN", PDECase=PDECase)
        self._base_pde = None

    # ---- slab grid and probe ---------------------------------------------------------------------
    def _edges(self):
        td = self.PDECase.geomtime.timedomain
        return np.linspace(td.t0, td.t1, self.t_num + 1)

    def _probe_points(self, edges):
        space = self.PDECase.geomtime.geometry
        slabs = [dde.geometry.GeometryXTime(space, dde.geometry.TimeDomain(a, b)).uniform_points(self.num_per_t)
                 for a, b in zip(edges[:-1], edges[1:])]
        return np.concatenate(slabs, axis=0)

    def get_pde_error(self, x):
        r = np.abs(self.model.predict(x, operator=self.PDECase.pde)).astype(np.float64)
        if r.shape[0] != len(x):          # several residuals come back stacked: add them up
            r = np.sum(r, axis=0).squeeze()
        return r.squeeze()

    # ---- weights ---------------------------------------------------------------------------------
    def slab_weights(self, residual):
        density = np.power(residual, self.k) / np.power(residual, self.k).mean() + self.c
        density = density / sum(density)  # Python's left-to-right float64 sum, as the reference
        per_slab = np.array([density[i * self.num_per_t:(i + 1) * self.num_per_t].mean() for i in range(self.t_num)])
        w = np.ones_like(per_slab)
        for i in range(1, self.t_num):
            w[i] = np.exp(-EPSILON * per_slab[: i - 1].sum())
        return w

    def get_weight_func(self, weights, t_list):
        inner = torch.as_tensor(np.asarray(t_list, dtype=np.float64))
        table = torch.as_tensor(np.concatenate([[0.0], np.asarray(weights, dtype=np.float64), [0.0]]))

        def weight_func(x):
            t = x[:, -1].detach()
            # compared in the points' own dtype, as the reference's tensor-vs-scalar comparisons are;
            # 0: before t0, t_num + 1: at / after t1
            slab = torch.bucketize(t, inner.to(device=t.device, dtype=t.dtype), right=True)
            return table.to(t.device)[slab].to(x.dtype).reshape(-1, 1)

        return weight_func

    def set_pde_weight(self, weight_func):
        if self._base_pde is None:
            self._base_pde = self.PDECase.pde
        pde = self.PDECase.pde            # the reference wraps the CURRENT pde: weights of successive rounds compound

        def scaled(x, res):
            w = weight_func(x)
            return [w * r for r in res] if isinstance(res, list) else w * res

        if dde.utils.get_num_args(pde) == 2:
            weighted = lambda x, y: scaled(x, pde(x, y))
        else:
            weighted = lambda x, y, ex: scaled(x, pde(x, y, ex))
        self.PDECase.set_pde(weighted)
        self.model.data = self.PDECase.data

    def update_weights(self):
        edges = self._edges()
        self.weight = self.slab_weights(self.get_pde_error(self._probe_points(edges)))
        self.set_pde_weight(self.get_weight_func(self.weight, edges))

    def closure(self):
        self.train_step(iterations=self.warmup_iter)
        for _ in range(self.iter):
            self.update_weights()
            self.train_step(iterations=1000)
