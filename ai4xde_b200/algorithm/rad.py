"""RAD: residual-based adaptive distribution (/root/reference/AI4XDE/algorithm/RAD.py:7-42).
Candidate residuals, probabilities and the without-replacement draw run on the device
(ai4xde_b200/resample.py); the candidate points and the uniforms come from NumPy's global legacy
stream in the reference's order (candidates first, then one `random_sample` per retry round), so the
selected set equals `np.random.choice(..., replace=False, p=...)` on the same residuals."""
from .. import dde, resample
from ..solver import PINNSolver


class RAD(PINNSolver):
    def __init__(self, PDECase, iter=100, k=1, c=1, warmup_iter=1000, fast=False):
        # [ext] fast: parallel float64 sum / cumsum in the selection instead of NumPy's sequential order
        self.fast = bool(fast)
        dde.optimizers.config.set_LBFGS_options(maxiter=1000)
        super().__init__(name=f"RAD_k_{k}_c_{c}", PDECase=PDECase)
        self.iter, self.k, self.c, self.warmup_iter = iter, k, c, warmup_iter

    def resample(self):
        case = self.PDECase
        _, rank, world = resample._world()
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
