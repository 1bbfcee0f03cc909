"""data.PDE / data.TimePDE point bookkeeping and losses (rows D1, D2)."""
import numpy as np
import torch

from . import config, utils
from .utils import run_if_all_none, run_if_any_none


class Data:
    def losses(self, targets, outputs, loss_fn, inputs, model, aux=None):
        raise NotImplementedError

    def losses_train(self, targets, outputs, loss_fn, inputs, model, aux=None):
        return self.losses(targets, outputs, loss_fn, inputs, model, aux=aux)

    def losses_test(self, targets, outputs, loss_fn, inputs, model, aux=None):
        return self.losses(targets, outputs, loss_fn, inputs, model, aux=aux)


class PDE(Data):
    def __init__(
        self,
        geometry,
        pde,
        bcs,
        num_domain=0,
        num_boundary=0,
        train_distribution="Hammersley",
        anchors=None,
        exclusions=None,
        solution=None,
        num_test=None,
        auxiliary_var_function=None,
    ):
        self.geom = geometry
        self.pde = pde
        self.bcs = bcs if isinstance(bcs, (list, tuple)) else [bcs]
        self.num_domain = num_domain
        self.num_boundary = num_boundary
        self.train_distribution = train_distribution
        self.anchors = None if anchors is None else anchors.astype(config.real(np))
        self.exclusions = exclusions
        self.soln = solution
        self.num_test = num_test
        self.auxiliary_var_fn = auxiliary_var_function

        self.train_x_all = None
        self.train_x_bc = None
        self.num_bcs = None
        self.train_x, self.train_y = None, None
        self.test_x, self.test_y = None, None
        self.train_aux_vars, self.test_aux_vars = None, None

        self.train_next_batch()
        self.test()

    def losses(self, targets, outputs, loss_fn, inputs, model, aux=None):
        f = []
        if self.pde is not None:
            if utils.get_num_args(self.pde) == 2:
                f = self.pde(inputs, outputs)
            elif utils.get_num_args(self.pde) == 3:
                f = self.pde(inputs, outputs, model.net.auxiliary_vars)
            if not isinstance(f, (list, tuple)):
                f = [f]
        if not isinstance(loss_fn, (list, tuple)):
            loss_fn = [loss_fn] * (len(f) + len(self.bcs))
        elif len(loss_fn) != len(f) + len(self.bcs):
            raise ValueError(
                "There are {} errors, but only {} losses.".format(len(f) + len(self.bcs), len(loss_fn))
            )
        bcs_start = np.cumsum([0] + self.num_bcs)
        bcs_start = list(map(int, bcs_start))
        error_f = [fi[bcs_start[-1] :] for fi in f]
        losses = [loss_fn[i](torch.zeros_like(error), error) for i, error in enumerate(error_f)]
        for i, bc in enumerate(self.bcs):
            beg, end = bcs_start[i], bcs_start[i + 1]
            error = bc.error(self.train_x, inputs, outputs, beg, end)
            losses.append(loss_fn[len(error_f) + i](torch.zeros_like(error), error))
        return losses

    @run_if_all_none("train_x", "train_y", "train_aux_vars")
    def train_next_batch(self, batch_size=None):
        self.train_x_all = self.train_points()
        self.bc_points()
        if self.bcs:
            self.train_x = np.vstack((self.train_x_bc, self.train_x_all))
        else:
            self.train_x = self.train_x_all
        self.train_y = self.soln(self.train_x) if self.soln else None
        if self.auxiliary_var_fn is not None:
            self.train_aux_vars = self.auxiliary_var_fn(self.train_x).astype(config.real(np))
        return self.train_x, self.train_y, self.train_aux_vars

    @run_if_all_none("test_x", "test_y", "test_aux_vars")
    def test(self):
        if self.num_test is None:
            self.test_x = self.train_x
        else:
            self.test_x = self.test_points()
        self.test_y = self.soln(self.test_x) if self.soln else None
        if self.auxiliary_var_fn is not None:
            self.test_aux_vars = self.auxiliary_var_fn(self.test_x).astype(config.real(np))
        return self.test_x, self.test_y, self.test_aux_vars

    def resample_train_points(self, pde_points=True, bc_points=True):
        if pde_points:
            self.train_x_all = None
        if bc_points:
            self.train_x_bc = None
        self.train_x, self.train_y, self.train_aux_vars = None, None, None
        self.train_next_batch()

    def add_anchors(self, anchors):
        anchors = anchors.astype(config.real(np))
        if self.anchors is None:
            self.anchors = anchors
        else:
            self.anchors = np.vstack((anchors, self.anchors))
        self.train_x_all = np.vstack((anchors, self.train_x_all))
        self.train_x = self.bc_points()
        if self.pde is not None:
            self.train_x = np.vstack((self.train_x, self.train_x_all))
        self.train_y = self.soln(self.train_x) if self.soln else None
        if self.auxiliary_var_fn is not None:
            self.train_aux_vars = self.auxiliary_var_fn(self.train_x).astype(config.real(np))

    def replace_with_anchors(self, anchors):
        self.anchors = anchors.astype(config.real(np))
        self.train_x_all = self.anchors
        self.train_x = self.bc_points()
        if self.pde is not None:
            self.train_x = np.vstack((self.train_x, self.train_x_all))
        self.train_y = self.soln(self.train_x) if self.soln else None
        if self.auxiliary_var_fn is not None:
            self.train_aux_vars = self.auxiliary_var_fn(self.train_x).astype(config.real(np))

    @run_if_all_none("train_x_all")
    def train_points(self):
        X = np.empty((0, self.geom.dim), dtype=config.real(np))
        if self.num_domain > 0:
            if self.train_distribution == "uniform":
                X = self.geom.uniform_points(self.num_domain, boundary=False)
            else:
                X = self.geom.random_points(self.num_domain, random=self.train_distribution)
        if self.num_boundary > 0:
            if self.train_distribution == "uniform":
                tmp = self.geom.uniform_boundary_points(self.num_boundary)
            else:
                tmp = self.geom.random_boundary_points(
                    self.num_boundary, random=self.train_distribution
                )
            X = np.vstack((tmp, X))
        if self.anchors is not None:
            X = np.vstack((self.anchors, X))
        if self.exclusions is not None:

            def is_not_excluded(x):
                return not np.any([np.allclose(x, y) for y in self.exclusions])

            X = np.array(list(filter(is_not_excluded, X)))
        self.train_x_all = X
        return X

    @run_if_any_none("train_x_bc")
    def bc_points(self):
        x_bcs = [bc.collocation_points(self.train_x_all) for bc in self.bcs]
        self.num_bcs = list(map(len, x_bcs))
        self.train_x_bc = (
            np.vstack(x_bcs)
            if x_bcs
            else np.empty([0, self.train_x_all.shape[-1]], dtype=config.real(np))
        )
        return self.train_x_bc

    def test_points(self):
        x = self.geom.uniform_points(self.num_test, boundary=False)
        x = np.vstack((self.train_x_bc, x))
        return x


class IDE(PDE):
    """Integro-differential equation data [DeepXDE-upstream `deepxde.data.IDE`], restated: Gauss-Legendre nodes
    (quad_x + 1) x / 2 of int_0^x for every collocation point are appended to the batch,
    train_x = [bc points ; train_x_all ; quadrature points], and the residual function receives `int_mat`
    with int_mat[i + n_bc, nodes of point i] = quad_w x_i / 2 * kernel(x_i, node).  Pinned in
    tests/test_oracle_pins.py by int_0^x sin(2 pi t) dt and by the analytic solutions of the two reference
    cases (cases/PDECases.py:1977-2092)."""

    def __init__(self, geometry, ide, bcs, quad_deg, kernel=None, num_domain=0, num_boundary=0,
                 train_distribution="Hammersley", anchors=None, solution=None, num_test=None):
        self.kernel = kernel if kernel is not None else (lambda x, s: np.ones((len(x), 1)))
        self.quad_deg = quad_deg
        self.quad_x, self.quad_w = np.polynomial.legendre.leggauss(quad_deg)
        self.quad_x = self.quad_x.astype(config.real(np))
        self.quad_w = self.quad_w.astype(config.real(np))
        super().__init__(geometry, ide, bcs, num_domain=num_domain, num_boundary=num_boundary,
                         train_distribution=train_distribution, anchors=anchors, solution=solution,
                         num_test=num_test)

    def losses_train(self, targets, outputs, loss_fn, inputs, model, aux=None):
        bcs_start = list(map(int, np.cumsum([0] + self.num_bcs)))
        f = self.pde(inputs, outputs, self.get_int_matrix(True))
        if not isinstance(f, (list, tuple)):
            f = [f]
        f = [fi[bcs_start[-1]:] for fi in f]
        losses = [loss_fn(torch.zeros_like(fi), fi) for fi in f]
        for i, bc in enumerate(self.bcs):
            beg, end = bcs_start[i], bcs_start[i + 1]
            error = bc.error(self.train_x, inputs, outputs, beg, end)
            losses.append(loss_fn(torch.zeros_like(error), error))
        return losses

    def losses_test(self, targets, outputs, loss_fn, inputs, model, aux=None):
        f = self.pde(inputs, outputs, self.get_int_matrix(False))
        if not isinstance(f, (list, tuple)):
            f = [f]
        return [loss_fn(torch.zeros_like(fi), fi) for fi in f] + [
            torch.zeros((), dtype=outputs.dtype) for _ in self.bcs]

    losses = losses_train

    @run_if_all_none("train_x", "train_y", "train_aux_vars")
    def train_next_batch(self, batch_size=None):
        self.train_x_all = self.train_points()
        x_bc = self.bc_points()
        x_quad = self.quad_points(self.train_x_all)
        self.train_x = np.vstack((x_bc, self.train_x_all, x_quad))
        self.train_y = self.soln(self.train_x) if self.soln else None
        return self.train_x, self.train_y, self.train_aux_vars

    @run_if_all_none("test_x", "test_y", "test_aux_vars")
    def test(self):
        x = self.train_x_all if self.num_test is None else self.geom.uniform_points(self.num_test, True)
        self._n_test_f = len(x)
        self.test_x = np.vstack((x, self.quad_points(x)))
        self.test_y = self.soln(self.test_x) if self.soln else None
        return self.test_x, self.test_y, self.test_aux_vars

    def replace_with_anchors(self, anchors):
        self.anchors = anchors.astype(config.real(np))
        self.train_x_all = self.anchors
        self.train_x = np.vstack((self.bc_points(), self.train_x_all, self.quad_points(self.train_x_all)))
        self.train_y = self.soln(self.train_x) if self.soln else None

    def quad_points(self, X):
        return np.hstack([(self.quad_x + 1) * xi[0] / 2 for xi in X])[:, None].astype(config.real(np))

    def get_int_matrix(self, training):
        if training:
            num_bc, X, num_f = int(sum(self.num_bcs)), self.train_x, len(self.train_x_all)
        else:
            num_bc, X, num_f = 0, self.test_x, self._n_test_f
        int_mat = np.zeros((num_bc + num_f, X.size), dtype=config.real(np))
        for i in range(num_f):
            x = X[i + num_bc, 0]
            beg = num_f + num_bc + self.quad_deg * i
            end = beg + self.quad_deg
            K = np.ravel(self.kernel(np.full((self.quad_deg, 1), x), X[beg:end]))
            int_mat[i + num_bc, beg:end] = self.quad_w * x / 2 * K
        return int_mat


class TimePDE(PDE):
    def __init__(
        self,
        geometryxtime,
        pde,
        ic_bcs,
        num_domain=0,
        num_boundary=0,
        num_initial=0,
        train_distribution="Hammersley",
        anchors=None,
        exclusions=None,
        solution=None,
        num_test=None,
        auxiliary_var_function=None,
    ):
        self.num_initial = num_initial
        super().__init__(
            geometryxtime,
            pde,
            ic_bcs,
            num_domain,
            num_boundary,
            train_distribution=train_distribution,
            anchors=anchors,
            exclusions=exclusions,
            solution=solution,
            num_test=num_test,
            auxiliary_var_function=auxiliary_var_function,
        )

    @run_if_all_none("train_x_all")
    def train_points(self):
        X = PDE.train_points.__wrapped__(self)
        if self.num_initial > 0:
            if self.train_distribution == "uniform":
                tmp = self.geom.uniform_initial_points(self.num_initial)
            else:
                tmp = self.geom.random_initial_points(
                    self.num_initial, random=self.train_distribution
                )
            if self.exclusions is not None:

                def is_not_excluded(x):
                    return not np.any([np.allclose(x, y) for y in self.exclusions])

                tmp = np.array(list(filter(is_not_excluded, tmp)))
            X = np.vstack((tmp, X))
        self.train_x_all = X
        return X
