"""Residual programs on the GPU: the CUDA interpreter (csrc/residual_vm.cu) against the float64 torch
interpreter of the same lowered program (tests/_taylor_double.py; its reverse pass is torch.autograd), and the
train closure with traced programs against the train closure on torch.autograd over the jet leaves (which the
parity tests pin to the oracle).  FP32 bound: 1e-5 of the largest entry."""
import numpy as np
import pytest
import torch

import _taylor_double as td

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


_VARS = {}


def _all_ops_code(x, y):
    import ai4xde_b200.dde as dde

    c1, c2 = _VARS["c1"], _VARS["c2"]       # trainable scalars (dde.Variable): read at run time, gradients returned

    u = torch.sin(x[:, 0:1]) * y[:, 0:1] + torch.exp(-x[:, 1:2]) * torch.tanh(y[:, 1:2]) / (1.5 + torch.cos(x[:, 0:1] * x[:, 1:2]))
    u = u + torch.sqrt(2.0 + (y[:, 0:1] * y[:, 1:2]) ** 2) - abs(y[:, 0:1]) ** 3 + torch.log(3.0 + u ** 2) + 2.0 ** x[:, 1:2]
    r = dde.grad.jacobian(u, x, i=0, j=1) - 0.3 * dde.grad.hessian(u, x, i=0, j=0) + dde.grad.hessian(u, x, i=0, j=1) * u
    v = dde.grad.hessian(y, x, component=1, i=1, j=1) / (1.0 + y[:, 0:1] ** 2) * torch.exp(-c1) + 1e-3 * c2 * u - c1 / 3.0
    # a term that pairs rows (PeriodicBC: left boundary rows against their periodic images further down the batch)
    per = (u[10:60] - u[110:160]) * x[10:60, 0:1] + dde.grad.jacobian(y, x, i=1, j=0)[10:60] - dde.grad.jacobian(y, x, i=1, j=0)[110:160]
    return [torch.mean(torch.square(r)), torch.mean(abs(u[5:200])), torch.mean(torch.square(v[100:])), torch.sum(v * u),
            torch.mean(torch.square(per))]


@pytest.mark.parametrize("n", [1, 37, 1000, 300001])
def test_interpreter_kernels_match_the_float64_interpreter(n):
    """Every opcode, five terms over four row ranges (one of them pairing rows 50 apart: mirror segment), two trainable
    scalars, two network outputs, S = 6 streams; ragged sizes; repeated launches bit-identical (fixed-order reductions)."""
    from ai4xde_b200 import _lib, residual as R
    from ai4xde_b200.jet import JetSpec

    import ai4xde_b200.dde as dde

    dev = torch.device("cuda:0")
    spec = JetSpec.full(2)
    _VARS["c1"], _VARS["c2"] = dde.Variable(0.7), dde.Variable(-1.3)
    code = _all_ops_code if n >= 256 else (lambda x, y: [torch.mean(torch.square(
        torch.sin(x[:, 0:1]) * y[:, 0:1] - y[:, 1:2] ** 2 / (2.0 + torch.exp(x[:, 1:2]))))])
    prog = R.trace(code, spec, n, 2, 2)
    gen = torch.Generator().manual_seed(n)
    ldn = _lib.pad32(n)
    x = torch.zeros((2, ldn), dtype=torch.float64)
    x[:, :n] = torch.rand((2, n), generator=gen, dtype=torch.float64)
    jets = torch.zeros((1, spec.S, 2, ldn), dtype=torch.float64)
    jets[..., :n] = torch.randn((1, spec.S, 2, n), generator=gen, dtype=torch.float64)
    tg = torch.randn(prog.n_terms, generator=gen, dtype=torch.float64)
    want = td._run_program(prog, x, jets, n)
    want_bar, want_sg = td.residual_backward(td.ResidualTables(prog), x, jets, n, tg)
    tables = _lib.ResidualTables(prog, dev)
    xs, js = x.float().to(dev), jets.float().to(dev)
    got = _lib.residual_forward(tables, xs, js, n)
    got_bar, got_sg = _lib.residual_backward(tables, xs, js, n, tg.float().to(dev))
    assert _rel(got.double().cpu(), want) < 1e-5, (got, want)
    assert _rel(got_bar.double().cpu(), want_bar) < 1e-5
    assert bool((got_bar[..., n:] == 0).all())
    if n >= 256:
        assert prog.scalar_grad == [0, 1] and _rel(got_sg.double().cpu(), want_sg.double()) < 1e-5, (got_sg, want_sg)
    else:
        assert got_sg is None
    assert torch.equal(got, _lib.residual_forward(tables, xs, js, n))
    again_bar, again_sg = _lib.residual_backward(tables, xs, js, n, tg.float().to(dev))
    assert torch.equal(got_bar, again_bar) and (got_sg is None or torch.equal(got_sg, again_sg))
    print(f"n={n}: {prog.describe()}; losses rel {_rel(got.double().cpu(), want):.2e}, jets_bar rel {_rel(got_bar.double().cpu(), want_bar):.2e}")


def _closure_loss_and_grad(model, fused):
    import ai4xde_b200.dde as dde

    dde.config.set_residual_programs(fused)
    dde.config.set_cuda_graphs(False)
    try:
        model.net.zero_grad()
        seen = {}
        orig = model.opt.step
        model.opt.step = lambda closure: seen.__setitem__("loss", closure())
        try:
            model.train_step(model.data.train_x, None, None)
        finally:
            model.opt.step = orig
        g = torch.cat([p.grad.reshape(-1) for p in model.net.parameters()]).clone()
        return seen["loss"].detach().clone(), g
    finally:
        dde.config.set_residual_programs(True)
        dde.config.set_cuda_graphs("auto")


def _case(name):
    from ai4xde_b200 import cases as pc

    np.random.seed(3)
    torch.manual_seed(3)
    if name == "cfg1":
        return pc.Burgers(NumDomain=2540, layer_size=[2] + [20] * 3 + [1], soft_constraints=True), None, "fp32"
    if name == "burgers_hard":
        return pc.Burgers(NumDomain=2540, layer_size=[2] + [20] * 3 + [1]), None, "fp32"
    if name == "cfg2":
        return pc.AllenCahn(NumDomain=10000, layer_size=[2] + [64] * 4 + [1]), None, "fp32"
    if name == "cfg3":
        return pc.NS_Flow_in_LidDriven_Cavity(NumDomain=2 ** 14, layer_size=[2] + [128] * 5 + [3]), None, "tf32"
    if name == "cfg4":
        return pc.Helmholtz(NumDomain=2 ** 15, layer_size=[2] + [256] * 6 + [1], activation="tanh",
                            honor_num_domain=True), [1, 100], "tf32"
    if name == "cfg4_sin":
        return pc.Helmholtz(NumDomain=2 ** 14, honor_num_domain=True), [1, 100], "auto"   # 2-[150]x3-1 sin
    if name == "cfg5":
        return pc.Poisson_1D_Fourier_Net(NumDomain=2 ** 14, layer_size=[1] + [64] * 3 + [1]), None, "fp32"
    raise KeyError(name)


@pytest.mark.parametrize("name", ["cfg1", "burgers_hard", "cfg2", "cfg3", "cfg4", "cfg4_sin", "cfg5"])
def test_train_closure_on_the_traced_program_matches_autograd(name):
    """BASELINE configs (and the hard-constraint / sin variants the reference defaults to): total loss and flat
    parameter gradient of the train closure, traced program vs autograd over the jet leaves, same kernels below.
    The MLP kernels are deterministic, so the only difference is float32 summation order in the loss algebra."""
    import ai4xde_b200.dde as dde

    dde.config.set_default_float("float32")
    case, lw, engine = _case(name)
    dde.config.set_engine(engine)
    try:
        model = dde.Model(case.data, case.net)
        model.compile("adam", lr=1e-3, loss_weights=lw)
        l_ref, g_ref = _closure_loss_and_grad(model, fused=False)
        l_fus, g_fus = _closure_loss_and_grad(model, fused=True)
        note = model.residual_program_note()
        assert note is not None and not note.startswith("autograd path"), note
        l2, g2 = _closure_loss_and_grad(model, fused=True)
        assert torch.equal(l_fus, l2) and torch.equal(g_fus, g2)      # bit-reproducible
    finally:
        dde.config.set_engine("auto")
    print(f"{name}: {note}; loss rel {abs(float(l_fus) - float(l_ref)) / abs(float(l_ref)):.2e}, grad rel {_rel(g_fus, g_ref):.2e}")
    assert abs(float(l_fus) - float(l_ref)) <= 1e-5 * abs(float(l_ref))
    assert _rel(g_fus, g_ref) < 1e-5


def test_training_with_programs_follows_the_autograd_trajectory():
    """200 Adam steps of cfg1 (CUDA-graph replay active): loss history with traced programs == without, to float32
    summation-order noise."""
    import ai4xde_b200.dde as dde
    from ai4xde_b200 import cases as pc

    dde.config.set_default_float("float32")
    hist = {}
    for fused in (False, True):
        np.random.seed(0)
        torch.manual_seed(0)
        dde.config.set_residual_programs(fused)
        try:
            case = pc.Burgers(NumDomain=2540, layer_size=[2] + [20] * 3 + [1], soft_constraints=True)
            model = dde.Model(case.data, case.net)
            model.compile("adam", lr=1e-3)
            lh, _ = model.train(iterations=200, display_every=50)
            hist[fused] = np.array([np.sum(l) for l in lh.loss_train])
            if fused:
                assert not model.residual_program_note().startswith("autograd path")
        finally:
            dde.config.set_residual_programs(True)
    print(hist)
    assert np.allclose(hist[True], hist[False], rtol=2e-3)


def test_inverse_problem_variables_get_their_gradient_from_the_program():
    """dde.Variable coefficients (cases/InverseCases.py:368, :392-400: Diffusion_Inverse `dy_t - C dy_xx + ...`, and the
    derived forms `-v_e / e`, `1e-3 * D` of :314, :505): the program reads their current value at every step and returns
    d loss / d variable; compared with autograd, then trained for 50 steps both ways."""
    import ai4xde_b200.dde as dde
    from ai4xde_b200 import cases as pc

    dde.config.set_default_float("float32")
    out = {}
    for fused in (False, True):
        np.random.seed(5)
        torch.manual_seed(5)
        case = pc.Burgers(NumDomain=4000, layer_size=[2] + [20] * 3 + [1], soft_constraints=True)
        C, D = dde.Variable(2.0), dde.Variable(0.5)
        base = case.data.pde

        def make_pde(base, C, D):
            def pde(x, y):          # two arguments: deepxde counts them to decide how to call it
                return base(x, y) * torch.exp(-D) - 1e-2 * C * dde.grad.hessian(y, x, i=0, j=0) + (-C / 7.0) * y
            return pde

        case.data.pde = make_pde(base, C, D)
        model = dde.Model(case.data, case.net)
        model.compile("adam", lr=1e-3, external_trainable_variables=[C, D])
        dde.config.set_residual_programs(fused)
        try:
            l, g = _closure_loss_and_grad(model, fused)
            gv = torch.stack([C.grad.reshape(()), D.grad.reshape(())]).clone()
            dde.config.set_residual_programs(fused)
            model.train(iterations=50, display_every=50)
            out[fused] = (l, g, gv, float(C), float(D), model.residual_program_note())
        finally:
            dde.config.set_residual_programs(True)
    (l0, g0, v0, c0, d0, _), (l1, g1, v1, c1, d1, note) = out[False], out[True]
    print(note, "variable gradients", v0.tolist(), v1.tolist(), "after 50 steps", (c0, d0), (c1, d1))
    assert "2 trainable variables" in note
    assert abs(float(l1) - float(l0)) <= 1e-5 * abs(float(l0)) and _rel(g1, g0) < 1e-5 and _rel(v1, v0) < 1e-5
    assert abs(c1 - c0) < 1e-4 and abs(d1 - d0) < 1e-4 and abs(c1 - 2.0) > 1e-3


def test_periodic_bc_pairs_rows_through_mirror_segments():
    """PeriodicBC (value and first derivative: reference cases Schrodinger, Poisson_1D_Dirichlet_Periodic) -- the loss
    term subtracts row p + m from row p; the thread of row p + m gets a mirror segment for its share of the adjoint.
    Train closure with the program against autograd over the jet leaves."""
    import ai4xde_b200.dde as dde
    from ai4xde_b200 import cases as pc

    dde.config.set_default_float("float32")
    np.random.seed(2)
    torch.manual_seed(2)
    case = pc.Burgers(NumDomain=3000, layer_size=[2] + [20] * 3 + [1], soft_constraints=True)
    gt = case.geomtime
    on_x = lambda x, on_boundary: on_boundary and (np.isclose(x[0], -1) or np.isclose(x[0], 1))
    bcs = list(case.data.bcs) + [dde.icbc.PeriodicBC(gt, 0, on_x, derivative_order=0),
                                 dde.icbc.PeriodicBC(gt, 0, on_x, derivative_order=1)]
    data = dde.data.TimePDE(gt, case.data.pde, bcs, num_domain=3000, num_boundary=200, num_initial=100,
                            train_distribution="pseudo")
    model = dde.Model(data, case.net)
    model.compile("adam", lr=1e-3)
    l_ref, g_ref = _closure_loss_and_grad(model, fused=False)
    l_fus, g_fus = _closure_loss_and_grad(model, fused=True)
    note = model.residual_program_note()
    print(note, f"loss rel {abs(float(l_fus) - float(l_ref)) / abs(float(l_ref)):.2e}, grad rel {_rel(g_fus, g_ref):.2e}")
    assert not note.startswith("autograd path"), note
    assert abs(float(l_fus) - float(l_ref)) <= 1e-5 * abs(float(l_ref)) and _rel(g_fus, g_ref) < 1e-5


def test_feature_transform_is_chained_symbolically():
    """net.apply_feature_transform(f) (reference: Laplace_disk in polar coordinates, LotkaVolterra's Fourier features,
    cases/PDECases.py:212,405): the kernels differentiate w.r.t. the features, the program applies the chain rule through
    f symbolically; the kernels are fed f(points), the program reads the raw points."""
    import ai4xde_b200.dde as dde
    from ai4xde_b200 import cases as pc

    dde.config.set_default_float("float32")
    np.random.seed(4)
    torch.manual_seed(4)
    case = pc.Burgers(NumDomain=3000, layer_size=[2] + [20] * 3 + [1], soft_constraints=True)
    case.net.apply_feature_transform(
        lambda x: torch.cat([x[:, 0:1] * torch.sin(x[:, 1:2]), torch.exp(-x[:, 1:2]) + x[:, 0:1] ** 2], dim=1))
    model = dde.Model(case.data, case.net)
    model.compile("adam", lr=1e-3)
    l_ref, g_ref = _closure_loss_and_grad(model, fused=False)
    l_fus, g_fus = _closure_loss_and_grad(model, fused=True)
    note = model.residual_program_note()
    assert not note.startswith("autograd path"), note
    assert abs(float(l_fus) - float(l_ref)) <= 1e-5 * abs(float(l_ref))
    assert float((g_fus - g_ref).abs().max()) <= 1e-5 * float(g_ref.abs().max()), note
    X = case.geomtime.random_points(500)
    dde.config.set_residual_programs(False)
    r0 = model.predict(X, operator=case.pde)
    dde.config.set_residual_programs(True)
    r1 = model.predict(X, operator=case.pde)
    assert not model.residual_program_note("operator").startswith("autograd path")
    assert np.abs(r1 - r0).max() <= 2e-5 * np.abs(r0).max()


@pytest.mark.parametrize("optimizer", ["adam", "L-BFGS"])
def test_repeated_train_calls_never_replay_a_stale_capture(optimizer):
    """model.train() called again on the same compiled model re-traces the loss code (its Python constants may have
    moved); a CUDA graph captured for the previous trace points at that trace's device tables and must not be
    replayed.  Three train() calls with programs == the same three calls on autograd."""
    import ai4xde_b200.dde as dde
    from ai4xde_b200 import cases as pc

    dde.config.set_default_float("float32")
    hist = {}
    for fused in (False, True):
        np.random.seed(1)
        torch.manual_seed(1)
        dde.config.set_residual_programs(fused)
        try:
            case = pc.Heat(NumDomain=1000, layer_size=[2] + [20] * 3 + [1])
            model = dde.Model(case.data, case.net)
            if optimizer == "adam":
                model.compile("adam", lr=1e-3)
                out = []
                for _ in range(3):
                    lh, _ = model.train(iterations=40, display_every=40)
                    out.append(float(np.sum(lh.loss_train[-1])))
                    torch.cuda.empty_cache()      # a dangling table pointer would now read recycled memory
                    junk = torch.full((1 << 22,), float("nan"), device="cuda")
                    del junk
            else:
                model.compile("adam", lr=1e-3)
                model.train(iterations=200, display_every=200)
                dde.optimizers.set_LBFGS_options(maxiter=60)
                model.compile("L-BFGS")
                out = []
                for _ in range(3):
                    lh, _ = model.train(display_every=1000)
                    out.append(float(np.sum(lh.loss_train[-1])))
                    junk = torch.full((1 << 22,), float("nan"), device="cuda")
                    del junk
            hist[fused] = out
        finally:
            dde.config.set_residual_programs(True)
            dde.optimizers.set_LBFGS_options()
    print(optimizer, hist)
    assert np.all(np.isfinite(hist[True])) and np.allclose(hist[True], hist[False], rtol=5e-3)
