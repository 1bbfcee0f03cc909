"""Forward PDE cases on the fused path.  Class names, constructor keywords, defaults, residuals,
geometries, point counts and BC sets follow /root/reference/AI4XDE/cases/PDECases.py (cited per
class); only what the hot path needs is kept (no plotting bodies).  Extra keyword arguments, all
defaulting to the reference behaviour, are marked `[ext]`."""
import numpy as np

from .. import dde
from ..dde import backend as bkd
from ..utils import Visualization_1D, Visualization_2D
from .base import PDECases, reference_data_path


def _xt_domain(x0, x1, t0=0, t1=1):
    return dde.geometry.GeometryXTime(dde.geometry.Interval(x0, x1), dde.geometry.TimeDomain(t0, t1))


class Burgers(PDECases):
    """u_t + u u_x = (0.01/pi) u_xx, hard IC/BC via output transform  (PDECases.py:1059-1125).
    [ext] soft_constraints=True gives DeepXDE's Burgers demo layout the BASELINE cfg1 names:
    80 boundary + 160 initial points with Dirichlet/IC loss terms and no output transform."""

    def __init__(self, NumDomain=2000, layer_size=[2] + [64] * 3 + [1], activation="tanh",
                 initializer="Glorot normal", soft_constraints=False):
        self.soft_constraints = soft_constraints
        super().__init__(name="Burgers", NumDomain=NumDomain, use_output_transform=not soft_constraints,
                         layer_size=layer_size, activation=activation, initializer=initializer,
                         visualization=Visualization_2D(x_limit=[0, 1], y_limit=[-1, 1], x_label="t", y_label="x",
                                                        feature_transform=lambda X: X[:, [1, 0]]))

    def gen_pde(self):
        nu = 0.01 / np.pi

        def pde(x, y):
            u_x = dde.grad.jacobian(y, x, i=0, j=0)
            u_t = dde.grad.jacobian(y, x, i=0, j=1)
            u_xx = dde.grad.hessian(y, x, i=0, j=0)
            return u_t + y * u_x - nu * u_xx

        return pde

    def gen_geomtime(self):
        return _xt_domain(-1, 1)

    def gen_data(self):
        if not self.soft_constraints:
            return dde.data.TimePDE(self.geomtime, self.pde, [], num_domain=self.NumDomain, num_test=10000,
                                    train_distribution="pseudo")
        bc = dde.icbc.DirichletBC(self.geomtime, lambda x: 0, lambda _, on_boundary: on_boundary)
        ic = dde.icbc.IC(self.geomtime, lambda x: -np.sin(np.pi * x[:, 0:1]), lambda _, on_initial: on_initial)
        return dde.data.TimePDE(self.geomtime, self.pde, [bc, ic], num_domain=self.NumDomain, num_boundary=80,
                                num_initial=160, train_distribution="pseudo")

    def gen_testdata(self):
        d = np.load(reference_data_path("Burgers.npz"))
        xx, tt = np.meshgrid(d["x"], d["t"])
        return np.vstack((np.ravel(xx), np.ravel(tt))).T, d["usol"].T.flatten()[:, None]

    def output_transform(self, x, y):
        xs = x[:, 0:1]
        return -bkd.sin(np.pi * xs) + (1 - xs**2) * x[:, 1:] * y


class Heat(PDECases):
    """u_t = a u_xx on [0, L] x [0, 1], Dirichlet + IC  (PDECases.py:1137-1206)."""

    def __init__(self, NumDomain=2540, layer_size=[2] + [20] * 3 + [1], activation="tanh",
                 initializer="Glorot normal"):
        self.a, self.L, self.n = 0.4, 1, 1
        super().__init__(name="Heat equation", NumDomain=NumDomain, use_output_transform=False,
                         layer_size=layer_size, activation=activation, initializer=initializer,
                         visualization=Visualization_2D(x_limit=[0, 1], y_limit=[0, self.L], x_label="t", y_label="x",
                                                        feature_transform=lambda X: X[:, [1, 0]]))

    def gen_pde(self):
        def pde(x, y):
            return dde.grad.jacobian(y, x, i=0, j=1) - self.a * dde.grad.hessian(y, x, i=0, j=0)

        return pde

    def sol(self, X):
        x, t = X[:, 0:1], X[:, 1:2]
        return np.exp(-(self.n**2 * np.pi**2 * self.a * t) / (self.L**2)) * np.sin(self.n * np.pi * x / self.L)

    def gen_geomtime(self):
        return _xt_domain(0, self.L)

    def gen_data(self):
        bc = dde.icbc.DirichletBC(self.geomtime, lambda x: 0, lambda _, on_boundary: on_boundary)
        ic = dde.icbc.IC(self.geomtime, lambda x: np.sin(self.n * np.pi * x[:, 0:1] / self.L),
                         lambda _, on_initial: on_initial)
        return dde.data.TimePDE(self.geomtime, self.pde, [bc, ic], num_domain=self.NumDomain, num_boundary=80,
                                solution=self.sol, num_test=2540, num_initial=160)


class Diffusion(PDECases):
    """u_t = u_xx + forcing with exact solution e^{-t} sin(pi x)  (PDECases.py:1225-1317)."""

    def __init__(self, NumDomain=2000, hard_constraint=True, layer_size=[2] + [32] * 3 + [1], activation="tanh",
                 initializer="Glorot uniform"):
        self.hard_constraint = hard_constraint
        super().__init__(name="Diffusion", NumDomain=NumDomain, use_output_transform=hard_constraint,
                         layer_size=layer_size, activation=activation, initializer=initializer,
                         metrics=["l2 relative error"],
                         visualization=Visualization_2D(x_limit=[0, 1], y_limit=[-1, 1], x_label="t", y_label="x",
                                                        feature_transform=lambda X: X[:, [1, 0]]))

    def gen_pde(self):
        def pde(x, y):
            dy_t = dde.grad.jacobian(y, x, j=1)
            dy_xx = dde.grad.hessian(y, x, j=0)
            xs = x[:, 0:1]
            return dy_t - dy_xx + bkd.exp(-x[:, 1:]) * (bkd.sin(np.pi * xs) - np.pi**2 * bkd.sin(np.pi * xs))

        return pde

    def sol(self, x):
        return np.sin(np.pi * x[:, 0:1]) * np.exp(-x[:, 1:])

    def gen_geomtime(self):
        return _xt_domain(-1, 1)

    def gen_data(self):
        if self.hard_constraint:
            return dde.data.TimePDE(self.geomtime, self.pde, [], num_domain=self.NumDomain, solution=self.sol,
                                    num_test=10000, train_distribution="pseudo")
        bc = dde.icbc.DirichletBC(self.geomtime, self.sol, lambda _, on_boundary: on_boundary)
        ic = dde.icbc.IC(self.geomtime, self.sol, lambda _, on_initial: on_initial)
        return dde.data.TimePDE(self.geomtime, self.pde, [bc, ic], num_domain=self.NumDomain, num_boundary=20,
                                num_initial=10, solution=self.sol, num_test=10000, train_distribution="pseudo")

    def output_transform(self, x, y):
        xs, t = x[:, 0:1], x[:, 1:]
        return bkd.sin(np.pi * xs) + (1 - xs**2) * t * y


class AllenCahn(PDECases):
    """u_t = 0.001 u_xx - 5 (u^3 - u), hard IC/BC  (PDECases.py:1419-1493)."""

    def __init__(self, NumDomain=2000, layer_size=[2] + [64] * 3 + [1], activation="tanh",
                 initializer="Glorot normal"):
        super().__init__(name="AllenCahn", NumDomain=NumDomain, use_output_transform=True, layer_size=layer_size,
                         activation=activation, initializer=initializer,
                         visualization=Visualization_2D(x_limit=[0, 1], y_limit=[-1, 1], x_label="t", y_label="x",
                                                        feature_transform=lambda X: X[:, [1, 0]]))

    def gen_pde(self):
        def pde(x, y):
            u_xx = dde.grad.hessian(y, x, i=0, j=0)
            u_t = dde.grad.jacobian(y, x, j=1)
            return u_t - 0.001 * u_xx + 5 * (y**3 - y)

        return pde

    def gen_geomtime(self):
        return _xt_domain(-1, 1)

    def gen_data(self):
        return dde.data.TimePDE(self.geomtime, self.pde, [], num_domain=self.NumDomain, num_test=10000,
                                train_distribution="pseudo")

    def gen_testdata(self):
        from scipy.io import loadmat

        d = loadmat(reference_data_path("Allen_Cahn.mat"))
        xx, tt = np.meshgrid(d["x"], d["t"])
        return np.vstack((np.ravel(xx), np.ravel(tt))).T, d["u"].flatten()[:, None]

    def output_transform(self, x, y):
        xs, t = x[:, 0:1], x[:, 1:2]
        return t * (1 + xs) * (1 - xs) * y + bkd.square(xs) * bkd.cos(np.pi * xs)


class Helmholtz(PDECases):
    """-u_xx - u_yy - k0^2 u = f on the unit square, k0 = 2 pi n  (PDECases.py:504-583).
    The reference hard-codes 400 domain + 80 boundary points whatever NumDomain says;
    [ext] honor_num_domain=True samples NumDomain pseudo-random points instead (what the
    2^20-point BASELINE cfg4 needs)."""

    def __init__(self, NumDomain=2540, hard_constraint=False, layer_size=[2] + [150] * 3 + [1], activation="sin",
                 initializer="Glorot uniform", honor_num_domain=False):
        self.n = 2
        self.k0 = 2 * np.pi * self.n
        self.hard_constraint = hard_constraint
        self.honor_num_domain = honor_num_domain
        super().__init__(name="Helmholtz equation over a 2D square domain", NumDomain=NumDomain,
                         use_output_transform=False, layer_size=layer_size, activation=activation,
                         initializer=initializer, metrics=["l2 relative error"],
                         loss_weights=None if hard_constraint else [1, 100],
                         visualization=Visualization_2D(x_limit=[0, 1], y_limit=[0, 1], x_label="x1", y_label="x2"))

    def gen_pde(self):
        k0 = self.k0

        def pde(x, y):
            u_xx = dde.grad.hessian(y, x, i=0, j=0)
            u_yy = dde.grad.hessian(y, x, i=1, j=1)
            f = k0**2 * bkd.sin(k0 * x[:, 0:1]) * bkd.sin(k0 * x[:, 1:2])
            return -u_xx - u_yy - k0**2 * y - f

        return pde

    def sol(self, x):
        return np.sin(self.k0 * x[:, 0:1]) * np.sin(self.k0 * x[:, 1:2])

    def gen_geomtime(self):
        return dde.geometry.Rectangle([0, 0], [1, 1])

    def gen_data(self):
        bc = [] if self.hard_constraint else dde.icbc.DirichletBC(self.geomtime, lambda x: 0,
                                                                   lambda _, on_boundary: on_boundary)
        nx_train = int(1 / ((1 / self.n) / 10))
        nx_test = int(1 / ((1 / self.n) / 30))
        if self.honor_num_domain:
            return dde.data.PDE(self.geomtime, self.pde, bc, num_domain=self.NumDomain, num_boundary=4 * nx_train,
                                solution=self.sol, num_test=nx_test**2, train_distribution="pseudo")
        return dde.data.PDE(self.geomtime, self.pde, bc, num_domain=nx_train**2, num_boundary=4 * nx_train,
                            solution=self.sol, num_test=nx_test**2)


class NS_Flow_in_LidDriven_Cavity(PDECases):
    """Steady incompressible Navier-Stokes in the lid-driven cavity, Re = 100: outputs (u, v, p),
    three residuals, eight Dirichlet terms  (PDECases.py:2631-2750)."""

    def __init__(self, NumDomain=2000, layer_size=[2] + [50] * 3 + [3], activation="tanh",
                 initializer="Glorot normal", train_distribution="Hammersley"):
        self.Re = 100
        self.train_distribution = train_distribution  # [ext]
        super().__init__(name="Flow in a Lid-Driven Cavity", NumDomain=NumDomain, use_output_transform=False,
                         layer_size=layer_size, activation=activation, initializer=initializer, metrics=None,
                         visualization=Visualization_2D(x_limit=[0, 1], y_limit=[0, 1], x_label="x1", y_label="x2"))

    def gen_testdata(self):
        g = np.linspace(0, 1, 100)
        X = np.array([[a, b] for a in g for b in g])
        u = np.genfromtxt(reference_data_path("Flow_in_LidDriven_Cavity_u.csv"), delimiter=",")
        v = np.genfromtxt(reference_data_path("Flow_in_LidDriven_Cavity_v.csv"), delimiter=",")
        return X, np.hstack((u.reshape(-1, 1), v.reshape(-1, 1), np.zeros((u.size, 1))))

    def gen_pde(self):
        inv_re = 1 / self.Re

        def pde(x, y):
            u, v = y[:, 0:1], y[:, 1:2]
            du = [dde.grad.jacobian(y, x, i=0, j=j) for j in (0, 1)]
            dv = [dde.grad.jacobian(y, x, i=1, j=j) for j in (0, 1)]
            dp = [dde.grad.jacobian(y, x, i=2, j=j) for j in (0, 1)]
            lap_u = dde.grad.hessian(y, x, component=0, i=0, j=0) + dde.grad.hessian(y, x, component=0, i=1, j=1)
            lap_v = dde.grad.hessian(y, x, component=1, i=0, j=0) + dde.grad.hessian(y, x, component=1, i=1, j=1)
            mom_x = u * du[0] + v * du[1] + dp[0] - inv_re * lap_u
            mom_y = u * dv[0] + v * dv[1] + dp[1] - inv_re * lap_v
            return mom_x, mom_y, du[0] + dv[1]

        return pde

    def gen_geomtime(self):
        return dde.geometry.Rectangle(xmin=[0, 0], xmax=[1, 1])

    def gen_data(self):
        def wall(axis, value):
            return lambda x, on_boundary: on_boundary and np.isclose(x[axis], value)

        walls = [wall(0, 0), wall(0, 1), wall(1, 0), wall(1, 1)]  # left, right, bottom, top
        bcs = []
        for comp in (0, 1):
            for k, w in enumerate(walls):
                lid = comp == 0 and k == 3
                bcs.append(dde.icbc.DirichletBC(self.geomtime, (lambda x: 1) if lid else (lambda x: 0), w,
                                                component=comp))
        return dde.data.PDE(self.geomtime, self.pde, bcs, self.NumDomain, 200, num_test=2000,
                            train_distribution=self.train_distribution)


class Harmonic_Oscillator_1D(PDECases):
    """m u'' + mu u' + k u = 0, under-damped, from the FBPINN paper  (PDECases.py:2764-2847)."""

    def __init__(self, NumDomain=50, hard_condition=False, layer_size=[1] + [64] * 3 + [1], activation="tanh",
                 initializer="Glorot uniform"):
        self.xi, self.w0, self.m = 2, 20, 1
        self.mu, self.k = 2 * self.xi * self.m, self.w0**2 * self.m
        self.hard_condition = hard_condition
        super().__init__(name="Harmonic_Oscillator_1D", NumDomain=NumDomain, use_output_transform=hard_condition,
                         layer_size=layer_size, activation=activation, initializer=initializer,
                         metrics=["l2 relative error"], loss_weights=None if hard_condition else [1, 1e6, 1e2],
                         visualization=Visualization_1D(x_label="t", y_label="u"))

    def gen_pde(self):
        def pde(x, y):
            return self.m * dde.grad.hessian(y, x) + self.mu * dde.grad.jacobian(y, x) + self.k * y

        return pde

    def sol(self, x):
        w = np.sqrt(self.w0**2 - self.xi**2)
        phi = np.arctan(-self.xi / w)
        A = 1 / (2 * np.cos(phi))
        return 2 * A * np.exp(-self.xi * x) * np.cos(phi + w * x)

    def gen_geomtime(self):
        return dde.geometry.TimeDomain(0, 1)

    def gen_data(self):
        if self.hard_condition:
            bcs = []
        else:
            t0 = lambda x, on_boundary: on_boundary and np.isclose(x[0], 0)
            bcs = [dde.icbc.DirichletBC(self.geomtime, lambda x: 1, t0),
                   dde.icbc.NeumannBC(self.geomtime, lambda x: 0, t0)]
        return dde.data.PDE(self.geomtime, self.pde, bcs, self.NumDomain, 2, solution=self.sol, num_test=500)

    def output_transform(self, x, y):
        return 1 + bkd.tanh(x) ** 2 * y
