"""Independent known-answer tests of the boundary-condition semantics and the train_x row layout, for BOTH
restatements of DeepXDE in this repo -- oracle.ddelite (test infrastructure) and ai4xde_b200.dde (product) -- against
closed forms written out by hand here.  The two namespaces share an author; these tests share nothing with either
("oracle common-mode risk": Neumann / Robin / Periodic / Operator / PointSet / IC errors and the
`[BC rows ... ; anchors ; boundary ; domain]` layout).

Field: u0 = sin(2x) e^y, u1 = x^2 y on the rectangle [0,1] x [0,2]; every expected value below is the analytic
expression evaluated in NumPy."""
import numpy as np
import pytest
import torch


def _ns(which):
    if which == "oracle":
        from oracle import ddelite as dde
    else:
        import ai4xde_b200.dde as dde
    dde.config.set_default_float("float64")
    return dde


@pytest.fixture(params=["oracle", "product"])
def dde(request):
    ns = _ns(request.param)
    # the oracle evaluates BC targets in float64; the product keeps float64 point sets on the host and evaluates
    # targets / errors in float32 (device arithmetic)
    ns.kat_rtol = 1e-12 if request.param == "oracle" else 2e-6
    yield ns
    ns.config.set_default_float("float32")
    torch.set_default_dtype(torch.float32)


def field(x):
    return torch.cat([torch.sin(2 * x[:, 0:1]) * torch.exp(x[:, 1:2]), x[:, 0:1] ** 2 * x[:, 1:2]], dim=1)


def np_u0(X):
    return np.sin(2 * X[:, 0:1]) * np.exp(X[:, 1:2])


def np_u1(X):
    return X[:, 0:1] ** 2 * X[:, 1:2]


# boundary rows of the batch: right side (x = 1), bottom (y = 0), left (x = 0), top (y = 2); then interior rows
X_RIGHT = np.array([[1.0, 0.3], [1.0, 1.7]])
X_BOTTOM = np.array([[0.2, 0.0], [0.9, 0.0]])
X_LEFT = np.array([[0.0, 0.4], [0.0, 1.1]])
X_IN = np.array([[0.5, 0.5], [0.25, 1.5], [0.8, 0.1]])


def evaluate(dde, bc, X, beg, end):
    xt = torch.as_tensor(X, dtype=torch.float64).requires_grad_()
    y = field(xt)
    err = bc.error(X, xt, y, beg, end)
    dde.grad.clear()
    return err.detach().cpu().numpy()


def test_dirichlet_takes_the_component_and_the_rows(dde):
    geom = dde.geometry.Rectangle([0, 0], [1, 2])
    g = lambda x: 3.0 * x[:, 0:1] - x[:, 1:2]
    bc = dde.icbc.DirichletBC(geom, g, lambda x, on: on and np.isclose(x[0], 1.0), component=1)
    X = np.vstack((X_BOTTOM, X_RIGHT, X_IN))
    assert np.array_equal(bc.collocation_points(X), X_RIGHT)
    got = evaluate(dde, bc, X, 2, 4)
    np.testing.assert_allclose(got, np_u1(X_RIGHT) - g(X_RIGHT), rtol=dde.kat_rtol, atol=dde.kat_rtol)


def test_neumann_is_the_outward_normal_derivative(dde):
    geom = dde.geometry.Rectangle([0, 0], [1, 2])
    g = lambda x: 0.5 * x[:, 1:2]
    bc = dde.icbc.NeumannBC(geom, g, lambda x, on: on, component=0)
    X = np.vstack((X_RIGHT, X_BOTTOM, X_LEFT, X_IN))
    # right: n = (+1, 0) -> d u0 / dx = 2 cos(2x) e^y;  bottom: n = (0, -1) -> -sin(2x) e^y;  left: n = (-1, 0)
    want = np.vstack((2 * np.cos(2 * X_RIGHT[:, 0:1]) * np.exp(X_RIGHT[:, 1:2]),
                      -np_u0(X_BOTTOM),
                      -2 * np.cos(2 * X_LEFT[:, 0:1]) * np.exp(X_LEFT[:, 1:2]))) - g(X[:6])
    np.testing.assert_allclose(evaluate(dde, bc, X, 0, 6), want, rtol=dde.kat_rtol, atol=dde.kat_rtol)


def test_robin_uses_the_outputs_of_its_own_rows(dde):
    geom = dde.geometry.Rectangle([0, 0], [1, 2])
    bc = dde.icbc.RobinBC(geom, lambda X, y: 2.0 * y[:, 1:2] + torch.as_tensor(X[:, 1:2], dtype=y.dtype, device=y.device), lambda x, on: on and np.isclose(x[0], 1.0),
                          component=1)
    X = np.vstack((X_IN, X_RIGHT))
    # d u1 / dx at x = 1 is 2 x y = 2 y; func = 2 u1 + y = 2 y + y
    want = 2 * X_RIGHT[:, 0:1] * X_RIGHT[:, 1:2] - (2 * np_u1(X_RIGHT) + X_RIGHT[:, 1:2])
    np.testing.assert_allclose(evaluate(dde, bc, X, 3, 5), want, rtol=dde.kat_rtol, atol=dde.kat_rtol)


@pytest.mark.parametrize("order", [0, 1])
def test_periodic_pairs_each_row_with_its_image(dde, order):
    geom = dde.geometry.Rectangle([0, 0], [1, 2])
    bc = dde.icbc.PeriodicBC(geom, 0, lambda x, on: on and np.isclose(x[0], 0.0), derivative_order=order, component=0)
    rows = bc.collocation_points(np.vstack((X_LEFT, X_IN)))
    image = X_LEFT.copy()
    image[:, 0] = 1.0
    assert np.array_equal(rows, np.vstack((X_LEFT, image)))           # [points ; their periodic images]
    X = np.vstack((rows, X_IN))
    if order == 0:
        want = np_u0(X_LEFT) - np_u0(image)
    else:
        d = lambda Z: 2 * np.cos(2 * Z[:, 0:1]) * np.exp(Z[:, 1:2])
        want = d(X_LEFT) - d(image)
    np.testing.assert_allclose(evaluate(dde, bc, X, 0, 4), want, rtol=dde.kat_rtol, atol=dde.kat_rtol)


def test_operator_bc_slices_the_callers_expression(dde):
    geom = dde.geometry.Rectangle([0, 0], [1, 2])

    def op(x, y, X):
        return dde.grad.jacobian(y, x, i=1, j=1) + y[:, 0:1] - torch.as_tensor(X[:, 0:1])

    bc = dde.icbc.OperatorBC(geom, op, lambda x, on: on and np.isclose(x[1], 0.0))
    X = np.vstack((X_IN, X_BOTTOM))
    want = X_BOTTOM[:, 0:1] ** 2 + np_u0(X_BOTTOM) - X_BOTTOM[:, 0:1]        # d u1 / dy = x^2
    np.testing.assert_allclose(evaluate(dde, bc, X, 3, 5), want, rtol=dde.kat_rtol, atol=dde.kat_rtol)


def test_point_set_conditions_own_their_points(dde):
    pts = np.array([[0.3, 0.7], [0.6, 1.9]])
    vals = np.array([[0.25], [-1.0]])
    bc = dde.icbc.PointSetBC(pts, vals, component=0)
    assert np.array_equal(bc.collocation_points(X_IN), pts)                 # whatever the batch holds
    X = np.vstack((pts, X_IN))
    np.testing.assert_allclose(evaluate(dde, bc, X, 0, 2), np_u0(pts) - vals, rtol=dde.kat_rtol, atol=dde.kat_rtol)
    obc = dde.icbc.PointSetOperatorBC(pts, vals, lambda x, y, X_: dde.grad.jacobian(y, x, i=1, j=0))
    np.testing.assert_allclose(evaluate(dde, obc, X, 0, 2), 2 * pts[:, 0:1] * pts[:, 1:2] - vals, rtol=dde.kat_rtol, atol=dde.kat_rtol)


def test_initial_condition_on_geometry_x_time(dde):
    gt = dde.geometry.GeometryXTime(dde.geometry.Interval(0, 1), dde.geometry.TimeDomain(0, 2))
    ic = dde.icbc.IC(gt, lambda x: np.cos(x[:, 0:1]), lambda x, on_initial: on_initial, component=0)
    X = np.vstack((X_IN, X_BOTTOM))          # second column is time: the X_BOTTOM rows sit at t = 0
    assert np.array_equal(ic.collocation_points(X), X_BOTTOM)
    np.testing.assert_allclose(evaluate(dde, ic, X, 3, 5), np_u0(X_BOTTOM) - np.cos(X_BOTTOM[:, 0:1]), rtol=dde.kat_rtol, atol=dde.kat_rtol)


def test_train_x_row_layout(dde):
    """train_x_all = [anchors ; boundary ; domain] (TimePDE: [initial ; anchors ; boundary ; domain]) and
    train_x = [rows of BC 0 ; rows of BC 1 ; ... ; train_x_all], each BC's rows in their train_x_all order."""
    np.random.seed(5)
    geom = dde.geometry.Rectangle([0, 0], [1, 2])
    anchors = np.array([[0.5, 1.0], [1.0, 1.0]])                            # the second one lies on the right side
    right = dde.icbc.DirichletBC(geom, lambda x: 0, lambda x, on: on and np.isclose(x[0], 1.0))
    every = dde.icbc.NeumannBC(geom, lambda x: 0, lambda x, on: on)
    data = dde.data.PDE(geom, lambda x, y: y, [right, every], num_domain=7, num_boundary=6, anchors=anchors,
                        train_distribution="pseudo")
    A = data.train_x_all
    assert A.shape == (2 + 6 + 7, 2) and np.array_equal(A[:2], anchors)
    on = geom.on_boundary(A)
    assert on[2:8].all() and not on[8:].any() and on[1] and not on[0]
    rows_right = A[on & np.isclose(A[:, 0], 1.0)]
    rows_every = A[on]
    assert list(data.num_bcs) == [len(rows_right), len(rows_every)]
    assert np.array_equal(data.train_x, np.vstack((rows_right, rows_every, A)))
    # anchors added later go in front; the BC rows stay the cached ones
    data.add_anchors(np.array([[0.1, 0.1]]))
    assert np.array_equal(data.train_x_all, np.vstack(([[0.1, 0.1]], A)))
    assert np.array_equal(data.train_x, np.vstack((rows_right, rows_every, [[0.1, 0.1]], A)))
    data.replace_with_anchors(np.array([[0.2, 0.2], [0.3, 0.3]]))
    assert np.array_equal(data.train_x, np.vstack((rows_right, rows_every, [[0.2, 0.2], [0.3, 0.3]])))
    # time-dependent: initial points first
    gt = dde.geometry.GeometryXTime(dde.geometry.Interval(0, 1), dde.geometry.TimeDomain(0, 2))
    tdata = dde.data.TimePDE(gt, lambda x, y: y, [], num_domain=5, num_boundary=4, num_initial=3,
                             train_distribution="pseudo")
    T = tdata.train_x_all
    assert T.shape == (12, 2) and np.allclose(T[:3, 1], 0.0) and gt.on_boundary(T[3:7]).all() and (T[7:, 1] > 0).all()
