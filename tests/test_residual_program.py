"""Residual programs (ai4xde_b200/residual.py) on the CPU: the case's own pde / BC / loss code traced into a
register program must give the losses and parameter gradients of the ordinary autograd path.  The kernels are
replaced by the plain-PyTorch doubles (tests/_taylor_double.py: Taylor-mode network, program interpreter with
torch.autograd as its reverse pass); the CUDA interpreter is held to the same double in tests/test_gpu_residual.py."""
import numpy as np
import pytest
import torch

import _taylor_double as td


@pytest.fixture(autouse=True)
def _doubles(monkeypatch):
    import ai4xde_b200.dde as dde

    monkeypatch.setattr(dde.config, "device", lambda: torch.device("cpu"))
    td.install(monkeypatch)
    dde.config.set_default_float("float32")
    yield
    dde.config.set_residual_programs(True)


def _losses_and_grad(model, fused):
    import ai4xde_b200.dde as dde

    dde.config.set_residual_programs(fused)
    model.net.zero_grad()
    X = model.data.train_x
    seen = {}
    orig = model.opt.step

    def step(closure):            # evaluate the train closure once, without moving the weights
        seen["loss"] = closure()

    model.opt.step = step
    try:
        model.train_step(X, None, None)
    finally:
        model.opt.step = orig
    g = torch.cat([p.grad.reshape(-1) for p in model.net.parameters()]).clone()
    return float(seen["loss"]), g


def _case(name):
    from ai4xde_b200 import cases as pc

    np.random.seed(3)
    torch.manual_seed(3)
    if name == "burgers_soft":
        return pc.Burgers(NumDomain=300, layer_size=[2, 12, 12, 1], soft_constraints=True), None
    if name == "burgers_hard":
        return pc.Burgers(NumDomain=300, layer_size=[2, 12, 12, 1]), None
    if name == "allencahn":
        return pc.AllenCahn(NumDomain=300, layer_size=[2, 12, 12, 1]), None
    if name == "ns_cavity":
        return pc.NS_Flow_in_LidDriven_Cavity(NumDomain=200, layer_size=[2, 12, 12, 3]), None
    if name == "helmholtz":
        return pc.Helmholtz(NumDomain=256, layer_size=[2, 12, 12, 1], activation="sin", honor_num_domain=True), [1, 100]
    if name == "poisson1d":
        return pc.Poisson_1D_Fourier_Net(NumDomain=200, layer_size=[1, 12, 12, 1]), None
    raise KeyError(name)


@pytest.mark.parametrize("name", ["burgers_soft", "burgers_hard", "allencahn", "ns_cavity", "helmholtz", "poisson1d"])
def test_traced_program_reproduces_the_autograd_path(name):
    """BASELINE cases P1..P5 (+ the hard-constraint Burgers: output transform inside net.forward): the train
    closure on the traced program == the train closure on torch.autograd over the jet leaves."""
    import ai4xde_b200.dde as dde

    case, lw = _case(name)
    model = dde.Model(case.data, case.net)
    model.compile("adam", lr=1e-3, loss_weights=lw)
    l_ref, g_ref = _losses_and_grad(model, fused=False)
    l_fus, g_fus = _losses_and_grad(model, fused=True)
    note = model.residual_program_note()
    assert note is not None and not note.startswith("autograd path"), note
    assert abs(l_fus - l_ref) <= 2e-5 * abs(l_ref), (l_fus, l_ref)
    assert float((g_fus - g_ref).abs().max()) <= 2e-4 * float(g_ref.abs().max()), note
    print(name, note)


def test_symbolic_derivatives_match_autograd():
    """Graph.diff against torch.autograd on a transform that mixes every supported operation."""
    from ai4xde_b200 import residual as R
    from ai4xde_b200.jet import JetSpec

    spec = JetSpec.full(2)

    def loss_code(x, y):
        u = torch.sin(x[:, 0:1]) * y + torch.exp(-x[:, 1:2]) * torch.tanh(y) / (1.5 + torch.cos(x[:, 0:1] * x[:, 1:2]))
        u = u + torch.sqrt(2.0 + y * y) - abs(y) ** 3 + torch.log(3.0 + u ** 2) + 2.0 ** x[:, 1:2]
        import ai4xde_b200.dde as dde

        r = dde.grad.jacobian(u, x, i=0, j=1) - 0.3 * dde.grad.hessian(u, x, i=0, j=0) \
            + dde.grad.hessian(u, x, i=0, j=1) * u
        return [torch.mean(torch.square(r)), torch.mean(abs(u[5:20]))]

    n = 40
    prog = R.trace(loss_code, spec, n, 2, 1)
    gen = torch.Generator().manual_seed(0)
    x = torch.rand((n, 2), generator=gen, dtype=torch.float64)
    jets = torch.randn((1, spec.S, 1, n), generator=gen, dtype=torch.float64)
    got = td._run_program(prog, x.t().contiguous(), jets, n)
    # the same code on torch tensors, the "network" being a quadratic form with exactly these jets at each point
    xt = x.clone().requires_grad_()
    from ai4xde_b200 import jet as J

    streams = [jets[0, s, :, :].t().clone() for s in range(spec.S)]
    y = J.make_leaf(xt, spec, streams)
    import ai4xde_b200.dde as dde

    dde.grad.register(y, J.JetOutput(xt, spec, streams))
    want = torch.stack(loss_code(xt, y))
    dde.grad.clear()
    assert torch.allclose(got, want.detach(), rtol=1e-10, atol=1e-12), (got, want)


@pytest.mark.parametrize("what", ["detach", "row_gather", "trainable", "matmul"])
def test_untraceable_code_stays_on_autograd(what):
    """Code a program cannot express is refused by the tracer -- the step still runs, on the autograd path."""
    import ai4xde_b200.dde as dde
    from ai4xde_b200 import cases as pc

    np.random.seed(0)
    torch.manual_seed(0)
    case = pc.Burgers(NumDomain=100, layer_size=[2, 8, 8, 1], soft_constraints=True)
    base = case.data.pde
    coef = torch.tensor(0.5, requires_grad=True)

    def pde(x, y):
        r = base(x, y)
        if what == "detach":
            return r * y.detach()
        if what == "row_gather":
            return r[torch.arange(len(r))]
        if what == "trainable":
            return r * coef
        return torch.matmul(r.t(), r)

    case.data.pde = pde
    model = dde.Model(case.data, case.net)
    model.compile("adam", lr=1e-3)
    if what == "matmul":
        return            # shape-changing code: not a point-wise residual at all; nothing to run
    loss, g = _losses_and_grad(model, fused=True)
    assert np.isfinite(loss) and bool(torch.isfinite(g).all())
    note = model.residual_program_note()
    assert note is None or note.startswith("autograd path"), note


def test_program_is_retraced_when_the_point_set_changes():
    """BC target arrays are baked per point-set version: after add_anchors the program is traced again."""
    import ai4xde_b200.dde as dde
    from ai4xde_b200 import cases as pc

    np.random.seed(0)
    torch.manual_seed(0)
    case = pc.Burgers(NumDomain=100, layer_size=[2, 8, 8, 1], soft_constraints=True)
    model = dde.Model(case.data, case.net)
    model.compile("adam", lr=1e-3)
    _losses_and_grad(model, fused=True)
    first = model._residual["train"][2]
    assert first is not None
    _losses_and_grad(model, fused=True)
    assert model._residual["train"][2] is first
    case.data.add_anchors(np.random.rand(17, 2).astype(np.float32))
    l_fus, g_fus = _losses_and_grad(model, fused=True)
    assert model._residual["train"][2] is not first and model._residual["train"][2] is not None
    l_ref, g_ref = _losses_and_grad(model, fused=False)
    assert abs(l_fus - l_ref) <= 2e-5 * abs(l_ref)
    assert float((g_fus - g_ref).abs().max()) <= 2e-4 * float(g_ref.abs().max())


@pytest.mark.parametrize("name", ["burgers_hard", "allencahn", "ns_cavity", "helmholtz"])
def test_operator_programs_reproduce_predict(name):
    """model.predict(X, operator=pde) -- the residual probe of RAD / RAR_D / R3 (algorithm/RAD.py:26) -- as a traced
    program that stores the operator's columns == the operator run on tensors over the jet leaves."""
    import ai4xde_b200.dde as dde

    case, lw = _case(name)
    model = dde.Model(case.data, case.net)
    model.compile("adam", lr=1e-3, loss_weights=lw)
    X = case.geomtime.random_points(333)
    dde.config.set_residual_programs(False)
    ref = model.predict(X, operator=case.pde)
    dde.config.set_residual_programs(True)
    got = model.predict(X, operator=case.pde)
    note = model.residual_program_note("operator")
    assert note is not None and not note.startswith("autograd path"), note
    ref = ref if isinstance(ref, list) else [ref]
    got = got if isinstance(got, list) else [got]
    assert len(ref) == len(got)
    for a, b in zip(got, ref):
        assert a.shape == b.shape
        assert np.abs(a - b).max() <= 2e-5 * np.abs(b).max(), note
    # another size: its own program; the first one stays cached
    got2 = model.predict(X[:50], operator=case.pde)
    got2 = got2 if isinstance(got2, list) else [got2]
    assert np.abs(got2[0] - ref[0][:50]).max() <= 2e-5 * np.abs(ref[0]).max()


def test_trainable_variables_are_traced_as_run_time_scalars():
    """dde.Variable (inverse problems): expressions of variables formed before they meet a point-wise tensor
    (`-v / e`, `1e-3 * D`, exp(-D)) are traced too; the program returns their gradients; a step changes them."""
    import ai4xde_b200.dde as dde
    from ai4xde_b200 import cases as pc

    np.random.seed(5)
    torch.manual_seed(5)
    case = pc.Burgers(NumDomain=200, layer_size=[2, 10, 10, 1], soft_constraints=True)
    C, D = dde.Variable(2.0), dde.Variable(0.5)
    base = case.data.pde

    def pde(x, y):
        # reflected forms with the variable on the right: 1.0 - D, 3.0 / C, 2.0 ** D
        return base(x, y) * torch.exp(-D) - 1e-2 * C * dde.grad.hessian(y, x, i=0, j=0) + (-C / 7.0) * y \
            + (1.0 - D) * (3.0 / C) * y * 1e-2 + (2.0 ** D) * 1e-3 * y

    case.data.pde = pde
    model = dde.Model(case.data, case.net)
    model.compile("adam", lr=1e-2, external_trainable_variables=[C, D])
    res = {}
    for fused in (False, True):
        C.grad = D.grad = None
        l, g = _losses_and_grad(model, fused=fused)
        res[fused] = (l, g, float(C.grad), float(D.grad))
    note = model.residual_program_note()
    assert "2 trainable variables" in note, note
    assert abs(res[True][0] - res[False][0]) <= 2e-5 * abs(res[False][0])
    assert float((res[True][1] - res[False][1]).abs().max()) <= 2e-4 * float(res[False][1].abs().max())
    assert abs(res[True][2] - res[False][2]) <= 2e-4 * abs(res[False][2]) and abs(res[True][3] - res[False][3]) <= 2e-4 * abs(res[False][3])
    import ai4xde_b200.dde as dde2

    dde2.config.set_residual_programs(True)
    before = (float(C), float(D))
    model.train_step(case.data.train_x, None, None)          # the program of the NEXT step reads the moved values
    l_after, _ = _losses_and_grad(model, fused=True)
    l_ref, _ = _losses_and_grad(model, fused=False)
    assert (float(C), float(D)) != before and abs(l_after - l_ref) <= 2e-5 * abs(l_ref)
    # a variable the model does not optimise is refused
    E = dde.Variable(1.0)
    case.data.pde = lambda x, y: base(x, y) * E
    model.compile("adam", lr=1e-2, external_trainable_variables=[C, D])
    _losses_and_grad(model, fused=True)
    assert model.residual_program_note().startswith("autograd path")


def test_periodic_bc_is_traced_with_row_offsets():
    """PeriodicBC subtracts two row ranges of equal length: the second operand becomes the same expression read m rows
    further down; the lowered program carries a mirror segment for those rows (adjoints only)."""
    import ai4xde_b200.dde as dde
    from ai4xde_b200 import cases as pc, residual as R

    np.random.seed(2)
    torch.manual_seed(2)
    case = pc.Burgers(NumDomain=200, layer_size=[2, 10, 10, 1], soft_constraints=True)
    gt = case.geomtime
    on_x = lambda x, on_boundary: on_boundary and (np.isclose(x[0], -1) or np.isclose(x[0], 1))
    bcs = list(case.data.bcs) + [dde.icbc.PeriodicBC(gt, 0, on_x, derivative_order=0),
                                 dde.icbc.PeriodicBC(gt, 0, on_x, derivative_order=1)]
    data = dde.data.TimePDE(gt, case.data.pde, bcs, num_domain=200, num_boundary=40, num_initial=20,
                            train_distribution="pseudo")
    model = dde.Model(data, case.net)
    model.compile("adam", lr=1e-3)
    l_ref, g_ref = _losses_and_grad(model, fused=False)
    l_fus, g_fus = _losses_and_grad(model, fused=True)
    note = model.residual_program_note()
    assert not note.startswith("autograd path"), note
    prog = model._residual["train"][2].program
    assert any(int(op) == R.OP_SUMT_ADJ for _, _, code in prog.segments for op in code[:, 0])
    assert abs(l_fus - l_ref) <= 2e-5 * abs(l_ref)
    assert float((g_fus - g_ref).abs().max()) <= 2e-4 * float(g_ref.abs().max())


def test_feature_transform_is_chained_symbolically():
    """net.apply_feature_transform(f) (reference: Laplace_disk in polar coordinates, LotkaVolterra's Fourier features,
    cases/PDECases.py:212,405): the kernels differentiate w.r.t. the features, the program applies the chain rule through
    f symbolically; the kernels are fed f(points), the program reads the raw points."""
    import ai4xde_b200.dde as dde
    from ai4xde_b200 import cases as pc

    dde.config.set_default_float("float32")
    np.random.seed(4)
    torch.manual_seed(4)
    case = pc.Burgers(NumDomain=200, layer_size=[2, 10, 10, 1], soft_constraints=True)
    case.net.apply_feature_transform(
        lambda x: torch.cat([x[:, 0:1] * torch.sin(x[:, 1:2]), torch.exp(-x[:, 1:2]) + x[:, 0:1] ** 2], dim=1))
    model = dde.Model(case.data, case.net)
    model.compile("adam", lr=1e-3)
    l_ref, g_ref = _losses_and_grad(model, fused=False)
    l_fus, g_fus = _losses_and_grad(model, fused=True)
    note = model.residual_program_note()
    assert not note.startswith("autograd path"), note
    assert abs(float(l_fus) - float(l_ref)) <= 2e-5 * abs(float(l_ref))
    assert float((g_fus - g_ref).abs().max()) <= 2e-4 * float(g_ref.abs().max()), note
    X = case.geomtime.random_points(500)
    dde.config.set_residual_programs(False)
    r0 = model.predict(X, operator=case.pde)
    dde.config.set_residual_programs(True)
    r1 = model.predict(X, operator=case.pde)
    assert not model.residual_program_note("operator").startswith("autograd path")
    assert np.abs(r1 - r0).max() <= 2e-5 * np.abs(r0).max()


def test_a_coefficient_changed_between_steps_is_picked_up():
    """A program freezes the Python numbers its code reads; the model watches the ones it can see (closure cells and the
    numeric attributes of the objects in them) and re-traces when one moves -- a callback annealing `case.nu`."""
    import ai4xde_b200.dde as dde
    from ai4xde_b200 import cases as pc

    np.random.seed(0)
    torch.manual_seed(0)
    case = pc.Burgers(NumDomain=100, layer_size=[2, 8, 8, 1], soft_constraints=True)

    class Coef:
        nu = 0.01

    coef = Coef()
    base = case.data.pde

    def pde(x, y):
        return base(x, y) + coef.nu * dde.grad.hessian(y, x, i=0, j=0)

    case.data.pde = pde
    model = dde.Model(case.data, case.net)
    model.compile("adam", lr=1e-3)
    l0, _ = _losses_and_grad(model, fused=True)
    first = model._residual["train"][2]
    coef.nu = 0.5
    l1, g1 = _losses_and_grad(model, fused=True)
    assert model._residual["train"][2] is not first
    l_ref, g_ref = _losses_and_grad(model, fused=False)
    assert abs(l1 - l0) > 1e-6 * abs(l0) and abs(l1 - l_ref) <= 2e-5 * abs(l_ref)
    assert float((g1 - g_ref).abs().max()) <= 2e-4 * float(g_ref.abs().max())
