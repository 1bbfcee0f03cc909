"""Reference `pde(x, y)` closures of the BASELINE configs, restated on ddelite (test
infrastructure, see oracle/__init__.py).  Each function cites the reference lines it follows;
tests/test_oracle_cases.py checks them against the UNMODIFIED reference closures when
/root/reference is present and against tests/golden/*.npz everywhere.

build(cfg) assembles the synthetic problem of SURVEY.md section 8(d) for that config:
points from np.random (seeded by the caller), Glorot-initialised dde.nn.FNN.
"""
import numpy as np
import torch

from . import ddelite as dde

bkd = dde.backend


# ---- P1: Burgers  (/root/reference/AI4XDE/cases/PDECases.py:1087-1094, transform :1124-1125)
def burgers_pde(x, y):
    dy_x = dde.grad.jacobian(y, x, i=0, j=0)
    dy_t = dde.grad.jacobian(y, x, i=0, j=1)
    dy_xx = dde.grad.hessian(y, x, i=0, j=0)
    return dy_t + y * dy_x - 0.01 / np.pi * dy_xx


def burgers_output_transform(x, y):
    return -bkd.sin(np.pi * x[:, 0:1]) + (1 - x[:, 0:1] ** 2) * (x[:, 1:]) * y


# ---- P2: Allen-Cahn  (PDECases.py:1447-1454, transform :1488-1493)
def allen_cahn_pde(x, y):
    u = y
    du_xx = dde.grad.hessian(y, x, i=0, j=0)
    du_t = dde.grad.jacobian(y, x, j=1)
    return du_t - 0.001 * du_xx + 5 * (u**3 - u)


def allen_cahn_output_transform(x, y):
    x_in = x[:, 0:1]
    t_in = x[:, 1:2]
    return t_in * (1 + x_in) * (1 - x_in) * y + bkd.square(x_in) * bkd.cos(np.pi * x_in)


# ---- P3: lid-driven cavity Navier-Stokes, Re = 100  (PDECases.py:2683-2710)
NS_RE = 100


def ns_cavity_pde(x, y):
    Re = NS_RE
    u, v = y[:, 0:1], y[:, 1:2]
    u_x = dde.grad.jacobian(y, x, i=0, j=0)
    u_y = dde.grad.jacobian(y, x, i=0, j=1)
    v_x = dde.grad.jacobian(y, x, i=1, j=0)
    v_y = dde.grad.jacobian(y, x, i=1, j=1)
    p_x = dde.grad.jacobian(y, x, i=2, j=0)
    p_y = dde.grad.jacobian(y, x, i=2, j=1)
    u_xx = dde.grad.hessian(y, x, component=0, i=0, j=0)
    u_yy = dde.grad.hessian(y, x, component=0, i=1, j=1)
    v_xx = dde.grad.hessian(y, x, component=1, i=0, j=0)
    v_yy = dde.grad.hessian(y, x, component=1, i=1, j=1)
    f_u = u * u_x + v * u_y + p_x - 1 / Re * (u_xx + u_yy)
    f_v = u * v_x + v * v_y + p_y - 1 / Re * (v_xx + v_yy)
    f_p = u_x + v_y
    return f_u, f_v, f_p


# ---- P4: Helmholtz, k0 = 2*pi*n, n = 2  (PDECases.py:538-550; sol :552-553)
HELMHOLTZ_K0 = 2 * np.pi * 2


def helmholtz_pde(x, y):
    k0 = HELMHOLTZ_K0
    dy_xx = dde.grad.hessian(y, x, i=0, j=0)
    dy_yy = dde.grad.hessian(y, x, i=1, j=1)
    f = k0**2 * bkd.sin(k0 * x[:, 0:1]) * bkd.sin(k0 * x[:, 1:2])
    return -dy_xx - dy_yy - k0**2 * y - f


def helmholtz_sol(x):
    return np.sin(HELMHOLTZ_K0 * x[:, 0:1]) * np.sin(HELMHOLTZ_K0 * x[:, 1:2])


# ---- P5: 1-D multiscale Poisson  (PoissonCases.py:48-53, forcing :469-473, sol :475-476)
POISSON_A, POISSON_B = 2, 50


def poisson1d_func(x):
    return -((np.pi * POISSON_A) ** 2) * bkd.sin(np.pi * POISSON_A * x) - 0.1 * (
        np.pi * POISSON_B
    ) ** 2 * bkd.sin(np.pi * POISSON_B * x)


def poisson1d_pde(x, y):
    dy_xx = dde.grad.hessian(y, x)
    return dy_xx - poisson1d_func(x)


def poisson1d_sol(x):
    return np.sin(np.pi * POISSON_A * x) + 0.1 * np.sin(np.pi * POISSON_B * x)


# ---- Heat (PDECases.py:1168-1174): the 2540/80/160 soft-BC layout cfg1 borrows
def heat_pde(x, y):
    a = 0.4
    dy_t = dde.grad.jacobian(y, x, i=0, j=1)
    dy_xx = dde.grad.hessian(y, x, i=0, j=0)
    return dy_t - a * dy_xx


# ---- gPINN: gradient-enhanced residual  (/root/reference/AI4XDE/algorithm/gPINN.py:31-50)
def gpinn_pde(pde):
    """[res_0 .. res_{m-1}, d res_0/d x_0, d res_0/d x_1, ..., d res_{m-1}/d x_{d-1}] for a 2-argument pde."""

    def g_pde(x, y):
        res = pde(x, y)
        res = list(res) if isinstance(res, (list, tuple)) else [res]
        out = list(res)
        for r in res:
            for j in range(x.shape[1]):
                out.append(dde.grad.jacobian(r, x, j=j))
        return out

    return g_pde


def gpinn_burgers_build(points, state_dict, layer_size):
    """The reference's own gPINN example (algorithm/gPINN.py:77-81): hard-constraint Burgers (PDECases.py:1059-1125:
    TimePDE without BCs, output transform :1124-1125) with the gradient-enhanced residual.  Returns (data, net)."""
    gt = dde.geometry.GeometryXTime(dde.geometry.Interval(-1, 1), dde.geometry.TimeDomain(0, 1))
    data = dde.data.TimePDE(gt, gpinn_pde(burgers_pde), [], num_domain=0, anchors=np.asarray(points),
                            train_distribution="pseudo")
    net = dde.nn.FNN(layer_size, "tanh", "Glorot normal")
    net.load_state_dict({k: torch.as_tensor(np.asarray(v), dtype=dde.config.real(torch)) for k, v in state_dict.items()})
    net.apply_output_transform(burgers_output_transform)
    return data, net


# ---- FBPINN window and subdomains  (/root/reference/AI4XDE/algorithm/FBPINN.py:45-79)
def fbpinn_subdomains(l, r, segment, overlap=0.1):
    x_list = np.linspace(l, r, segment + 1)
    out = []
    for i in range(segment):
        width = x_list[i + 1] - x_list[i]
        lo = max(x_list[i] - overlap * width, l)
        hi = min(x_list[i + 1] + overlap * width, r)
        out.append((lo, hi))
    return out


def fbpinn_window(x, lo, hi):
    return bkd.sigmoid(x - lo) * bkd.sigmoid(hi - x)


def fbpinn_build(limits, points, state_dicts, layer_size, activation="tanh", initializer="Glorot uniform"):
    """FBPINN on the 1-D multiscale Poisson problem (cfg5) as the reference wires it
    (/root/reference/AI4XDE/algorithm/FBPINN.py:23-43, 107-110): for every subdomain (lo, hi) a sub-case with
    geometry Interval(lo, hi), the parent's pde and Dirichlet condition on the SUBDOMAIN's ends, and a
    dde.nn.FNN whose output transform is window(x) * y (the parent's own transform returns None,
    cases/PDECases.py:93-94, so `y = y_parent` is skipped).  `points[k]` is sub-case k's train_x_all (boundary rows
    first, as dde.data.PDE lays them out) and `state_dicts[k]` its weights: the caller hands over the product's,
    so both sides see identical inputs.  Returns (datas, nets)."""
    datas, nets = [], []
    for (lo, hi), pts, sd in zip(limits, points, state_dicts):
        geom = dde.geometry.Interval(lo, hi)
        bc = dde.icbc.DirichletBC(geom, poisson1d_sol, lambda _, on_boundary: on_boundary)
        data = dde.data.PDE(geom, poisson1d_pde, bc, 0, 0, anchors=np.asarray(pts), train_distribution="pseudo",
                            solution=poisson1d_sol)
        net = dde.nn.FNN(layer_size, activation, initializer)
        net.load_state_dict({k: torch.as_tensor(np.asarray(v), dtype=dde.config.real(torch)) for k, v in sd.items()})
        net.apply_output_transform((lambda l, h: (lambda x, y: fbpinn_window(x, l, h) * y))(lo, hi))
        datas.append(data)
        nets.append(net)
    return datas, nets


def fbpinn_step_outputs(datas, nets, k, loss_weights=None):
    """One evaluation of sub-model k's patched closure (FBPINN.py:143-185): outputs = get_predict(inputs) =
    the in-order sum over ALL sub-nets (FBPINN.py:196-204) on model k's batch, losses of model k's data,
    gradient with respect to net k's parameters (the optimiser of model k holds only those, FBPINN.py:108-110)."""
    model = dde.Model(datas[k], nets[k])
    model.compile("adam", lr=1e-3, loss_weights=loss_weights)
    for net in nets:
        net.zero_grad()
    x = torch.as_tensor(datas[k].train_x).requires_grad_()
    result = None
    for net in nets:
        net.train()
        y = net(x)
        result = y if result is None else result + y
    res = datas[k].pde(x, result)
    losses = datas[k].losses_train(None, result, dde.losses.get("MSE"), x, model)
    losses = torch.stack(losses)
    if loss_weights is not None:
        losses = losses * torch.as_tensor(loss_weights)
    dde.grad.clear()
    losses.sum().backward()
    g = torch.cat([p.grad.reshape(-1) for p in nets[k].parameters()])
    return {
        "outputs": result.detach().numpy(),
        "residuals": [res.detach().numpy()],
        "losses": losses.detach().numpy(),
        "grad": g.numpy().copy(),
    }


# ---------------------------------------------------------------------------------------
CONFIGS = {
    # name: (layer_size, activation, initializer, n_domain default)
    "cfg1": ([2] + [20] * 3 + [1], "tanh", "Glorot normal", 2540),
    "cfg2": ([2] + [64] * 4 + [1], "tanh", "Glorot normal", 10000),
    "cfg3": ([2] + [128] * 5 + [3], "tanh", "Glorot normal", 2**18),
    "cfg4": ([2] + [256] * 6 + [1], "tanh", "Glorot uniform", 2**20),
    "cfg5": ([1] + [64] * 3 + [1], "tanh", "Glorot uniform", 2**20),
    # the reference's own Helmholtz defaults (cases/PDECases.py:513-515): 2-[150]x3-1, sin
    "cfg4_sin": ([2] + [150] * 3 + [1], "sin", "Glorot uniform", 2**18),
}


def build(cfg, num_domain=None, layer_size=None, activation=None):
    """Return (data, net, pde) for a BASELINE config on ddelite (current default float).
    Consumes np.random / torch RNG: seed both before calling."""
    ls, act, init, nd = CONFIGS[cfg]
    layer_size = layer_size or ls
    activation = activation or act
    nd = num_domain if num_domain is not None else nd
    if cfg == "cfg1":
        gt = dde.geometry.GeometryXTime(dde.geometry.Interval(-1, 1), dde.geometry.TimeDomain(0, 1))
        bc = dde.icbc.DirichletBC(gt, lambda x: 0, lambda _, on_boundary: on_boundary)
        ic = dde.icbc.IC(gt, lambda x: -np.sin(np.pi * x[:, 0:1]), lambda _, on_initial: on_initial)
        pde = burgers_pde
        data = dde.data.TimePDE(gt, pde, [bc, ic], num_domain=nd, num_boundary=80,
                                num_initial=160, train_distribution="pseudo")
        net = dde.nn.FNN(layer_size, activation, init)
    elif cfg == "cfg2":
        gt = dde.geometry.GeometryXTime(dde.geometry.Interval(-1, 1), dde.geometry.TimeDomain(0, 1))
        pde = allen_cahn_pde
        data = dde.data.TimePDE(gt, pde, [], num_domain=nd, train_distribution="pseudo")
        net = dde.nn.FNN(layer_size, activation, init)
        net.apply_output_transform(allen_cahn_output_transform)
    elif cfg == "cfg3":
        geom = dde.geometry.Rectangle(xmin=[0, 0], xmax=[1, 1])
        pde = ns_cavity_pde

        def side(axis, val):
            return lambda x, on_boundary: on_boundary and np.isclose(x[axis], val)

        bcs = []
        for comp in (0, 1):
            for axis, val in ((0, 0), (0, 1), (1, 0), (1, 1)):
                lid = 1 if (comp == 0 and axis == 1 and val == 1) else 0
                bcs.append(dde.icbc.DirichletBC(geom, (lambda v: (lambda x: v))(lid), side(axis, val),
                                                component=comp))
        data = dde.data.PDE(geom, pde, bcs, nd, 200, train_distribution="pseudo")
        net = dde.nn.FNN(layer_size, activation, init)
    elif cfg in ("cfg4", "cfg4_sin"):
        geom = dde.geometry.Rectangle([0, 0], [1, 1])
        pde = helmholtz_pde
        bc = dde.icbc.DirichletBC(geom, lambda x: 0, lambda _, on_boundary: on_boundary)
        data = dde.data.PDE(geom, pde, bc, num_domain=nd, num_boundary=80,
                            train_distribution="pseudo", solution=helmholtz_sol)
        net = dde.nn.FNN(layer_size, activation, init)
    elif cfg == "cfg5":
        geom = dde.geometry.Interval(0, 1)
        pde = poisson1d_pde
        bc = dde.icbc.DirichletBC(geom, poisson1d_sol, lambda _, on_boundary: on_boundary)
        data = dde.data.PDE(geom, pde, bc, nd, 2, train_distribution="pseudo", solution=poisson1d_sol)
        net = dde.nn.FNN(layer_size, activation, init)
    else:
        raise KeyError(cfg)
    return data, net, pde


LOSS_WEIGHTS = {"cfg4": [1, 100], "cfg4_sin": [1, 100]}


def train_step_outputs(data, net, loss_weights=None):
    """One reference-style evaluation: returns per-term weighted
    losses, residual list, and the flat parameter gradient of sum(losses)."""
    model = dde.Model(data, net)
    model.compile("adam", lr=1e-3, loss_weights=loss_weights)
    net.zero_grad()
    x = torch.as_tensor(data.train_x).requires_grad_()
    net.train()
    y = net(x)
    res = data.pde(x, y)
    res = list(res) if isinstance(res, (list, tuple)) else [res]
    losses = data.losses_train(None, y, dde.losses.get("MSE"), x, model)
    losses = torch.stack(losses)
    if loss_weights is not None:
        losses = losses * torch.as_tensor(loss_weights)
    dde.grad.clear()
    total = losses.sum()
    total.backward()
    g = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    return {
        "outputs": y.detach().numpy(),
        "residuals": [r.detach().numpy() for r in res],
        "losses": losses.detach().numpy(),
        "grad": g.numpy().copy(),
    }
