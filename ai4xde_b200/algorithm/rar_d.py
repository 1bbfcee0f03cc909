"""RAR_D: residual-based adaptive refinement with distribution
(/root/reference/AI4XDE/algorithm/RAR_D.py:7-37): 100 000 candidates per round, NumDomain // 100
new anchors drawn with probability ~ |residual|^k / mean + c and ADDED to the training set."""
from .. import dde, resample
from ..solver import PINNSolver


class RAR_D(PINNSolver):
    def __init__(self, PDECase, iter=100, k=2, c=0, device_rng=False, device_seed=0):
        # [ext] device_rng: the whole round on the device (see RAD)
        self.device_rng, self.device_seed, self._rng_offset = bool(device_rng), int(device_seed), 0
        dde.optimizers.config.set_LBFGS_options(maxiter=1000)
        super().__init__(name=f"RAR_D_k_{k}_c_{c}", PDECase=PDECase)
        self.iter, self.k, self.c = iter, k, c
        self.num_candidates = 100000

    def resample(self):
        case = self.PDECase
        _, rank, world = resample._world()
        if world == 1 and self.device_rng:
            from .rad import device_round

            done = device_round(self, case, self.num_candidates, case.NumDomain // 100, replace=False)
            if done is not None:
                return done
        if world == 1:
            X = case.geomtime.random_points(self.num_candidates)
            ids = resample.rad_select(self.model, X, case.pde, case.NumDomain // 100, k=self.k, c=self.c)
            case.data.add_anchors(X[ids])
            return ids
        # point-sharded run: same global candidates / uniforms on every rank, probe sharded, each rank
        # adds its block of the globally selected anchors (see RAD.resample)
        if getattr(self, "_shared", None) is None:
            self._shared = resample.shared_numpy_stream()
        with self._shared:
            X = case.geomtime.random_points(self.num_candidates)
            ids = resample.rad_select_sharded(self.model, X, case.pde, (case.NumDomain * world) // 100,
                                              k=self.k, c=self.c)
        lo, hi = resample.block_bounds(len(ids), rank, world)
        case.data.add_anchors(X[ids[lo:hi]])
        return ids

    def closure(self):
        self.train_step()
        for _ in range(self.iter):
            self.resample()
            self.train_step(iterations=1000)
