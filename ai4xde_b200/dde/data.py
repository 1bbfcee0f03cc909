"""dde.data.PDE / TimePDE: collocation-point store and per-term losses (SURVEY.md 8a rows D1/D2,
App. B.3/B.4).

Host contract kept exactly as the algorithms use it (`train_x_all`, `train_x`, `num_bcs`,
`add_anchors`, `replace_with_anchors`, `resample_train_points`, `bcs`; e.g.
/root/reference/AI4XDE/algorithm/R3Sampling.py:89-98, RAD.py:40): `train_x_all = [anchors;
boundary; domain]` (TimePDE prepends initial points), `train_x = [rows owned by each BC...;
train_x_all]`, BC rows cached across anchor updates.  Every mutation bumps `version`, which is
what invalidates the device-resident mirror the Model keeps.
"""
import numpy as np
import torch

from . import config, utils


class Data:
    def losses(self, targets, outputs, loss_fn, inputs, model, aux=None):
        raise NotImplementedError

    def losses_train(self, targets, outputs, loss_fn, inputs, model, aux=None):
        return self.losses(targets, outputs, loss_fn, inputs, model, aux=aux)

    def losses_test(self, targets, outputs, loss_fn, inputs, model, aux=None):
        return self.losses(targets, outputs, loss_fn, inputs, model, aux=aux)


class PDE(Data):
    uses_targets = False   # the PDE / BC losses never read `targets`: dde.Model does not upload train_y per step

    def __init__(self, geometry, pde, bcs, num_domain=0, num_boundary=0, train_distribution="Hammersley",
                 anchors=None, exclusions=None, solution=None, num_test=None, auxiliary_var_function=None):
        self._train_y, self._train_y_stale = None, False
        self.geom = geometry
        self.pde = pde
        self.bcs = bcs if isinstance(bcs, (list, tuple)) else [bcs]
        self.num_domain, self.num_boundary = num_domain, num_boundary
        self.train_distribution = train_distribution
        self.anchors = None if anchors is None else anchors.astype(config.real(np))
        self.exclusions = exclusions
        self.soln = solution
        self.num_test = num_test
        self.auxiliary_var_fn = auxiliary_var_function
        self.version = 0
        self.train_x_all = self.train_x_bc = self.num_bcs = None
        self.train_x = self.train_y = self.train_aux_vars = None
        self.test_x = self.test_y = self.test_aux_vars = None
        self.train_next_batch()
        self.test()

    # ---- losses (device) -------------------------------------------------------------------
    def losses(self, targets, outputs, loss_fn, inputs, model, aux=None):
        f = []
        if self.pde is not None:
            nargs = utils.get_num_args(self.pde)
            if nargs == 2:
                f = self.pde(inputs, outputs)
            elif nargs == 3:
                f = self.pde(inputs, outputs, model.net.auxiliary_vars)
            if not isinstance(f, (list, tuple)):
                f = [f]
        n_terms = len(f) + len(self.bcs)
        if not isinstance(loss_fn, (list, tuple)):
            loss_fn = [loss_fn] * n_terms
        elif len(loss_fn) != n_terms:
            raise ValueError("There are {} errors, but only {} losses.".format(n_terms, len(loss_fn)))
        starts = [0]
        for c in self.num_bcs:
            starts.append(starts[-1] + int(c))
        out = []
        for i, fi in enumerate(f):
            err = fi[starts[-1]:]
            out.append(loss_fn[i](torch.zeros_like(err), err))
        for i, bc in enumerate(self.bcs):
            err = bc.error(self.train_x, inputs, outputs, starts[i], starts[i + 1])
            out.append(loss_fn[len(f) + i](torch.zeros_like(err), err))
        return out

    # ---- point bookkeeping (host) ------------------------------------------------------------
    @property
    def train_y(self):
        """`soln(train_x)` as DeepXDE keeps it -- evaluated on first use after the point set changed, not on every
        anchor update (an R3 round updates twice; on 2^20 points the reference solution costs ~30 ms each time and no
        loss term reads it)."""
        if self._train_y_stale:
            self._train_y = self.soln(self.train_x) if (self.soln and self.train_x is not None) else None
            self._train_y_stale = False
        return self._train_y

    @train_y.setter
    def train_y(self, value):
        self._train_y, self._train_y_stale = value, False

    def _refresh_targets(self):
        self._train_y, self._train_y_stale = None, bool(self.soln)
        if self.auxiliary_var_fn is not None:
            self.train_aux_vars = self.auxiliary_var_fn(self.train_x).astype(config.real(np))
        self.version += 1

    def _assemble(self):
        bc_rows = self.bc_points()
        self.train_x = np.vstack((bc_rows, self.train_x_all)) if self.pde is not None else bc_rows
        self._refresh_targets()

    def train_next_batch(self, batch_size=None):
        if self.train_x is None and self.train_aux_vars is None:
            self.train_x_all = self.train_points()
            self.bc_points()
            self.train_x = np.vstack((self.train_x_bc, self.train_x_all)) if self.bcs else self.train_x_all
            self._refresh_targets()
        return self.train_x, self.train_y, self.train_aux_vars

    def test(self):
        if self.test_x is None and self.test_y is None and self.test_aux_vars is None:
            self.test_x = self.train_x if self.num_test is None else self.test_points()
            self.test_y = self.soln(self.test_x) if self.soln else None
            if self.auxiliary_var_fn is not None:
                self.test_aux_vars = self.auxiliary_var_fn(self.test_x).astype(config.real(np))
        return self.test_x, self.test_y, self.test_aux_vars

    def resample_train_points(self, pde_points=True, bc_points=True):
        if pde_points:
            self.train_x_all = None
        if bc_points:
            self.train_x_bc = None
        self.train_x = self.train_y = self.train_aux_vars = None
        self.train_next_batch()

    def add_anchors(self, anchors):
        anchors = anchors.astype(config.real(np))
        self.anchors = anchors if self.anchors is None else np.vstack((anchors, self.anchors))
        self.train_x_all = np.vstack((anchors, self.train_x_all))
        self._assemble()

    def replace_with_anchors(self, anchors):
        self.anchors = anchors.astype(config.real(np))
        self.train_x_all = self.anchors
        self._assemble()

    def set_train_x_all(self, points):
        """`replace_with_anchors(kept)` followed by `add_anchors(new)` (/root/reference/AI4XDE/algorithm/
        R3Sampling.py:95-98) leaves `anchors == train_x_all == vstack(new, kept)`: this sets that state from the
        already assembled array with ONE rebuild of `train_x` instead of two."""
        self.anchors = np.ascontiguousarray(points, dtype=config.real(np))
        self.train_x_all = self.anchors
        self._assemble()

    def _drop_excluded(self, X):
        if self.exclusions is None:
            return X
        keep = [not np.any([np.allclose(x, y) for y in self.exclusions]) for x in X]
        return X[np.array(keep, dtype=bool)]

    def _sampled_points(self):
        X = np.empty((0, self.geom.dim), dtype=config.real(np))
        uniform = self.train_distribution == "uniform"
        if self.num_domain > 0:
            X = (self.geom.uniform_points(self.num_domain, boundary=False) if uniform
                 else self.geom.random_points(self.num_domain, random=self.train_distribution))
        if self.num_boundary > 0:
            b = (self.geom.uniform_boundary_points(self.num_boundary) if uniform
                 else self.geom.random_boundary_points(self.num_boundary, random=self.train_distribution))
            X = np.vstack((b, X))
        if self.anchors is not None:
            X = np.vstack((self.anchors, X))
        return self._drop_excluded(X)

    def train_points(self):
        if self.train_x_all is None:
            self.train_x_all = self._sampled_points()
        return self.train_x_all

    def bc_points(self):
        if self.train_x_bc is None:
            rows = [bc.collocation_points(self.train_x_all) for bc in self.bcs]
            self.num_bcs = [len(r) for r in rows]
            self.train_x_bc = (np.vstack(rows) if rows
                               else np.empty([0, self.train_x_all.shape[-1]], dtype=config.real(np)))
        return self.train_x_bc

    def test_points(self):
        x = self.geom.uniform_points(self.num_test, boundary=False)
        return np.vstack((self.train_x_bc, x))


class IDE(PDE):
    """dde.data.IDE: integro-differential equations by Gauss-Legendre quadrature (reference call sites:
    /root/reference/AI4XDE/cases/PDECases.py:2015-2025 and :2068-2085).  Every collocation point x brings
    `quad_deg` quadrature nodes (quad_x + 1) x / 2 of int_0^x; the batch is
    `train_x = [BC rows ; train_x_all ; quadrature nodes]`, and `pde(x, y, int_mat)` receives the
    (n_bc + n_f) x n_rows matrix whose row i holds `w_q x_i / 2 * kernel(x_i, node)` at that point's nodes, so
    `int_mat @ y` is the integral at every collocation point.  The network is evaluated once on the whole
    batch by the fused kernels; `dde.grad.jacobian` reads its first-order stream."""

    def __init__(self, geometry, ide, bcs, quad_deg, kernel=None, num_domain=0, num_boundary=0,
                 train_distribution="Hammersley", anchors=None, solution=None, num_test=None):
        self.kernel = kernel if kernel is not None else (lambda x, s: np.ones((len(x), 1)))
        self.quad_deg = quad_deg
        qx, qw = np.polynomial.legendre.leggauss(quad_deg)
        self.quad_x, self.quad_w = qx.astype(config.real(np)), qw.astype(config.real(np))
        self._int_cache = {}
        super().__init__(geometry, ide, bcs, num_domain, num_boundary, train_distribution=train_distribution,
                         anchors=anchors, solution=solution, num_test=num_test)

    def quad_points(self, X):
        return np.hstack([(self.quad_x + 1) * xi[0] / 2 for xi in X])[:, None].astype(config.real(np))

    def _assemble(self):
        bc_rows = self.bc_points()
        self.train_x = np.vstack((bc_rows, self.train_x_all, self.quad_points(self.train_x_all)))
        self._int_cache = {}
        self._refresh_targets()

    def train_next_batch(self, batch_size=None):
        if self.train_x is None and self.train_aux_vars is None:
            self.train_x_all = self.train_points()
            self._assemble()
        return self.train_x, self.train_y, self.train_aux_vars

    def test(self):
        if self.test_x is None and self.test_y is None and self.test_aux_vars is None:
            x = self.train_x_all if self.num_test is None else self.geom.uniform_points(self.num_test, True)
            self._n_test_f = len(x)
            self.test_x = np.vstack((x, self.quad_points(x)))
            self.test_y = self.soln(self.test_x) if self.soln else None
        return self.test_x, self.test_y, self.test_aux_vars

    def get_int_matrix(self, training):
        hit = self._int_cache.get(bool(training))
        if hit is not None:
            return hit
        if training:
            n_bc, X, n_f = int(sum(self.num_bcs)), self.train_x, len(self.train_x_all)
        else:
            n_bc, X, n_f = 0, self.test_x, self._n_test_f
        mat = np.zeros((n_bc + n_f, X.size), dtype=config.real(np))
        for i in range(n_f):
            x = X[i + n_bc, 0]
            beg = n_f + n_bc + self.quad_deg * i
            end = beg + self.quad_deg
            K = np.ravel(self.kernel(np.full((self.quad_deg, 1), x), X[beg:end]))
            mat[i + n_bc, beg:end] = self.quad_w * x / 2 * K
        self._int_cache[bool(training)] = mat
        return mat

    def _residual_losses(self, training, outputs, loss_fn, inputs, n_bc):
        f = self.pde(inputs, outputs, self.get_int_matrix(training))
        f = list(f) if isinstance(f, (list, tuple)) else [f]
        if not isinstance(loss_fn, (list, tuple)):
            loss_fn = [loss_fn] * (len(f) + len(self.bcs))
        out = []
        for i, fi in enumerate(f):
            err = fi[n_bc:]
            out.append(loss_fn[i](torch.zeros_like(err), err))
        return out, loss_fn, len(f)

    def losses_train(self, targets, outputs, loss_fn, inputs, model, aux=None):
        starts = [0]
        for c in self.num_bcs:
            starts.append(starts[-1] + int(c))
        out, loss_fn, nf = self._residual_losses(True, outputs, loss_fn, inputs, starts[-1])
        for i, bc in enumerate(self.bcs):
            err = bc.error(self.train_x, inputs, outputs, starts[i], starts[i + 1])
            out.append(loss_fn[nf + i](torch.zeros_like(err), err))
        return out

    def losses_test(self, targets, outputs, loss_fn, inputs, model, aux=None):
        out, _, _ = self._residual_losses(False, outputs, loss_fn, inputs, 0)
        return out + [torch.zeros((), dtype=out[0].dtype, device=out[0].device) for _ in self.bcs]

    losses = losses_train


class TimePDE(PDE):
    def __init__(self, geometryxtime, pde, ic_bcs, num_domain=0, num_boundary=0, num_initial=0,
                 train_distribution="Hammersley", anchors=None, exclusions=None, solution=None,
                 num_test=None, auxiliary_var_function=None):
        self.num_initial = num_initial
        super().__init__(geometryxtime, pde, ic_bcs, num_domain, num_boundary,
                         train_distribution=train_distribution, anchors=anchors, exclusions=exclusions,
                         solution=solution, num_test=num_test, auxiliary_var_function=auxiliary_var_function)

    def train_points(self):
        if self.train_x_all is None:
            X = self._sampled_points()
            if self.num_initial > 0:
                if self.train_distribution == "uniform":
                    t0 = self.geom.uniform_initial_points(self.num_initial)
                else:
                    t0 = self.geom.random_initial_points(self.num_initial, random=self.train_distribution)
                X = np.vstack((self._drop_excluded(t0), X))
            self.train_x_all = X
        return self.train_x_all
