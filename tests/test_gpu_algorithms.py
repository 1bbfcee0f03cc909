"""GPU: the algorithm mirrors on the fused path against the oracle's NumPy selections, on the same
residuals (whatever the device computed) and the same NumPy random stream."""
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup():
    warnings.simplefilter("ignore")
    import ai4xde_b200.dde as dde

    dde.config.set_default_float("float64")
    return dde


def test_rad_round_equals_numpy_choice_on_device_residuals():
    dde = _setup()
    from ai4xde_b200 import algorithm, cases
    from oracle import sampling

    np.random.seed(0)
    torch.manual_seed(0)
    case = cases.AllenCahn(NumDomain=2000, layer_size=[2] + [64] * 4 + [1])
    solver = algorithm.RAD(case, iter=1, k=1, c=1)
    case.compile(solver.model, "adam", lr=1e-3)
    np.random.seed(21)
    state = np.random.get_state()
    ids = solver.resample()
    # oracle: same candidates, the device's residuals, NumPy's own choice
    np.random.set_state(state)
    X = case.geomtime.random_points(case.NumDomain * 10)
    res = solver.model.predict(X, operator=case.pde)
    ids_ref, picked = sampling.rad_select(X, res, case.NumDomain, k=1, c=1)
    assert np.array_equal(ids, ids_ref)
    assert np.array_equal(case.data.train_x_all, picked)
    assert len(np.unique(ids)) == case.NumDomain


@pytest.mark.parametrize("algo", ["RAD", "RAR_D"])
def test_rad_round_on_the_device_sampler(algo):
    """[ext] device_rng: candidates from the Philox box sampler, draw uniforms from the device generator, gather on
    the device.  Checked against a host replay of the same round: the candidates regenerated from (seed, offset), the
    operator on them, the oracle's NumPy-order probabilities and NumPy's own inverse-CDF draw fed the same uniforms
    (the device's parallel scan agrees with NumPy's sequential cumsum except possibly at ties: >= 99.9 % of the indices);
    the selected points are exactly candidates[ids]; the new point sets have their device mirrors registered."""
    dde = _setup()
    dde.config.set_default_float("float32")
    from ai4xde_b200 import _lib, algorithm, cases

    np.random.seed(0)
    torch.manual_seed(0)
    if algo == "RAD":
        case = cases.AllenCahn(NumDomain=2000, layer_size=[2] + [64] * 4 + [1])
        solver = algorithm.RAD(case, iter=1, k=1, c=1, device_rng=True, device_seed=7)
        n_cand, num = 20000, 2000
    else:
        case = cases.Burgers(NumDomain=1000, layer_size=[2, 20, 20, 20, 1])
        solver = algorithm.RAR_D(case, iter=1, k=2, c=0, device_rng=True, device_seed=7)
        n_cand, num = 100000, 10
    case.compile(solver.model, "adam", lr=1e-3)
    before = case.data.train_x_all.copy()
    ids = solver.resample()
    assert ids.shape == (num,) and len(np.unique(ids)) == num and ids.min() >= 0 and ids.max() < n_cand
    # replay: same candidates, same residuals, NumPy's arithmetic
    dev = torch.device("cuda:0")
    cand = torch.zeros((2, _lib.pad32(n_cand)), device=dev)
    _lib.uniform_box_points(cand, 0, n_cand, [-1.0, 0.0], [1.0, 1.0], 7, 0)
    X = cand[:, :n_cand].t().contiguous().cpu().numpy()
    assert X[:, 0].min() >= -1 and X[:, 0].max() < 1 and X[:, 1].min() >= 0 and X[:, 1].max() < 1
    picked = X[ids]
    new_all = picked if algo == "RAD" else np.vstack((picked, before))
    assert np.array_equal(case.data.train_x_all, new_all)
    res = np.abs(solver.model.predict(X, operator=case.pde)).astype(np.float64)
    err = np.power(res, solver.k) / np.power(res, solver.k).mean() + solver.c
    p = (err / sum(err))[:, 0]
    gen = torch.Generator(device=dev)
    gen.manual_seed((7 * 0x9E3779B1 + 0) & 0x7FFFFFFFFFFFFFFF)
    found, pp = [], p.copy()
    while len(found) < num:                      # numpy.random.choice(replace=False, p) with the device's uniforms
        u = torch.rand(num - len(found), dtype=torch.float64, device=dev, generator=gen).cpu().numpy()
        if found:
            pp[found] = 0
        cdf = np.cumsum(pp)
        cdf /= cdf[-1]
        new = np.searchsorted(cdf, u, side="right")
        _, first = np.unique(new, return_index=True)
        new = new[np.sort(first)]
        found += [int(i) for i in new if i not in set(found)][: num - len(found)]
    agree = np.mean(np.asarray(found[:num]) == ids)
    print(algo, "device round: agreement with the NumPy replay", agree)
    assert agree >= 0.999
    # device mirrors registered: the next step finds them, uploads nothing
    m = solver.model
    b = m.device_batch(case.data.train_x)
    assert any(b is c for c in m._batches)
    assert np.array_equal(b.aos.cpu().numpy(), case.data.train_x.astype(np.float32))
    m.train_step(case.data.train_x, None, None)
    assert np.isfinite(float(m.last_total_loss))


def test_rar_d_round_adds_anchors():
    dde = _setup()
    from ai4xde_b200 import algorithm, cases
    from oracle import sampling

    np.random.seed(1)
    torch.manual_seed(1)
    case = cases.Burgers(NumDomain=1000, layer_size=[2, 20, 20, 20, 1])
    solver = algorithm.RAR_D(case, iter=1, k=2, c=0)
    case.compile(solver.model, "adam", lr=1e-3)
    before = case.data.train_x_all.copy()
    np.random.seed(33)
    state = np.random.get_state()
    ids = solver.resample()
    np.random.set_state(state)
    X = case.geomtime.random_points(100000)
    res = solver.model.predict(X, operator=case.pde)
    ids_ref, picked = sampling.rar_d_select(X, res, 10, k=2, c=0)
    assert np.array_equal(ids, ids_ref)
    assert np.array_equal(case.data.train_x_all, np.vstack((picked, before)))


def test_r3_round_equals_numpy_selection():
    dde = _setup()
    from ai4xde_b200 import algorithm, cases, resample
    from oracle import sampling

    np.random.seed(2)
    torch.manual_seed(2)
    case = cases.Helmholtz(NumDomain=20000, layer_size=[2] + [64] * 3 + [1], activation="tanh", honor_num_domain=True)
    solver = algorithm.R3Sampling(case, max_iter=1)
    case.compile(solver.model, "adam", lr=1e-3)
    X = case.data.train_x_all.copy()
    res_dev, n = resample.residual_matrix(solver.model, X, case.pde)
    rs = sampling.r3_residual(res_dev[0, :n].cpu().numpy()[:, None])
    thr, keep_ref, Xr = sampling.r3_select(X, rs)
    np.random.seed(5)
    keep = solver.resample()
    assert np.array_equal(keep, keep_ref)
    new_all = case.data.train_x_all
    assert len(new_all) == len(X)                                   # population size is constant
    assert np.array_equal(new_all[len(X) - len(Xr):], Xr)           # [new ; retained]  (R3Sampling.py:95-98)
    # device-side compaction holds the same retained points (float32 mirror)
    keep2, thr2, x_dev, nk = resample.r3_retain(solver.model, X, case.pde)
    assert nk == int(keep_ref.sum()) and np.float32(thr) == np.float32(thr2)
    got = x_dev[:, n - nk:n].t().cpu().numpy()
    assert np.array_equal(got, Xr.astype(np.float32))


def test_fbpinn_batched_nets_on_gpu():
    dde = _setup()
    from ai4xde_b200 import algorithm, cases

    np.random.seed(0)
    torch.manual_seed(0)
    case = cases.Harmonic_Oscillator_1D(NumDomain=64, hard_condition=True, layer_size=[1, 64, 64, 64, 1])
    solver = algorithm.FBPINN(case, segment=8, iter=1)
    x = torch.rand(100, 1, device="cuda").requires_grad_()
    for engine in ("fp32", "tf32"):
        dde.config.set_engine(engine)
        y = solver.get_predict(x)
        ref = sum(m.net(x) for m in solver.model_list)
        tol = 1e-5 if engine == "fp32" else 5e-3
        assert float((y - ref).abs().max()) < tol * max(1.0, float(ref.abs().max()))
        dde.grad.clear()
    dde.config.set_engine("auto")
    # train a decomposition whose sub-cases all own boundary rows (sub-cases are built with the
    # parent's class defaults, FBPINN.py:42, so the oscillator's t=0 conditions would be empty -> NaN
    # on every subdomain but the first, exactly as in the reference)
    np.random.seed(1)
    case = cases.Poisson_1D_Dirichlet(NumDomain=64, layer_size=[1, 64, 64, 64, 1])
    solver = algorithm.FBPINN(case, segment=4, iter=1)
    dde.optimizers.set_LBFGS_options(maxiter=3)
    before = solver.bank.flat.clone()
    solver.train_step(iterations=3, eval=True)
    assert np.isfinite(solver.losshistory.loss_train[-1]).all()
    assert bool(((solver.bank.flat - before).abs().sum(dim=1) > 0).all())
    dde.optimizers.set_LBFGS_options()


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["fp32", "tf32"])
def test_fbpinn_64_subdomain_nets_match_oracle(engine):
    """BASELINE config 5 as north_star states it: FBPINN with 64 windowed subdomain nets 1-[64]x3-1 on the 1-D
    multiscale Poisson problem (reference schedule /root/reference/AI4XDE/algorithm/FBPINN.py:112-119, sum over
    nets :196-204).  get_predict on a sub-model's batch (all 64 nets: ONE launch of csrc/tc_bank.cu on the
    tensor-core engine, one batched launch sequence on the FP32 engine), the per-term losses, the residual and
    the ACTIVE net's gradient are compared with the oracle restatement (float64; pinned to the unmodified
    reference class by tests/test_oracle_pins.py), for the first, an interior and the last subdomain.
    FP32 engine: 1e-5 on the network sum; 5e-5 on residual / gradient, because the forcing 0.1 (50 pi)^2
    sin(50 pi x) is evaluated in float32 by the case's own pde code (|f| ~ 2.5e3 from an argument of ~150: 1.5e-5
    of the largest residual is float32 rounding of sin's argument, on any float32 backend).
    Tensor-core engine (TF32 operands, FP32 accumulate): stated bound 3e-3 on the sum / residual, 1e-2 on the
    gradient (measured ~1e-3)."""
    _setup()
    from parity_util import fbpinn_vs_oracle

    out, solver = fbpinn_vs_oracle(64, 500, engine=engine, check=(0, 29, 63))
    assert solver.bank.K == 64 and solver.bank.layer_sizes == [1, 64, 64, 64, 1]
    print({k: {n: (f"{v:.2e}" if isinstance(v, float) else v) for n, v in r.items()} for k, r in out.items()})
    tol_out, tol_res, tol_grad = (1e-5, 5e-5, 5e-5) if engine == "fp32" else (3e-3, 3e-3, 1e-2)
    for k, r in out.items():
        assert r["n_rows"] == 504 and r["others_untouched"], (k, r)
        assert r["out_rel"] < tol_out and r["loss_rel"] < 2 * tol_res, (k, r)
        assert r["res_rel"] < tol_res and r["grad_rel"] < tol_grad, (k, r)


def test_pinn_short_training_tracks_oracle_loss_curve():
    """1 % loss-curve agreement over the first 200 Adam steps, FP32 engine vs float64 oracle
    (cfg1: Burgers soft constraints, 2540+80+160 points)."""
    dde = _setup()
    from ai4xde_b200 import cases as pc
    from oracle import cases as oc, ddelite

    ddelite.config.set_default_float("float64")
    np.random.seed(0)
    torch.manual_seed(0)
    odata, onet, _ = oc.build("cfg1", 2540)
    omodel = ddelite.Model(odata, onet)
    omodel.compile("adam", lr=1e-3)
    state = {k: v.detach().float() for k, v in onet.state_dict().items()}
    olh, _ = omodel.train(iterations=200, display_every=20)
    ddelite.config.set_default_float("float32")
    torch.set_default_dtype(torch.float32)
    case = pc.Burgers(NumDomain=0, layer_size=[2, 20, 20, 20, 1], soft_constraints=True)
    case.data = dde.data.TimePDE(case.geomtime, case.pde, case.data.bcs, num_domain=0, anchors=odata.train_x_all,
                                 train_distribution="pseudo")
    case.net.load_state_dict(state)
    model = dde.Model(case.data, case.net)
    model.compile("adam", lr=1e-3)
    lh, _ = model.train(iterations=200, display_every=20)
    a, b = np.array(lh.loss_train).sum(1), np.array(olh.loss_train).sum(1)
    assert np.abs(a - b).max() / b.max() < 1e-2


def test_cuda_graph_step_matches_eager():
    """Small problems replay the captured Adam step; the trajectory must equal the eager one."""
    dde = _setup()
    from ai4xde_b200 import cases as pc

    def run(graphs):
        dde.config.set_cuda_graphs(graphs)
        np.random.seed(0)
        torch.manual_seed(0)
        case = pc.Burgers(NumDomain=2540, layer_size=[2, 20, 20, 20, 1], soft_constraints=True)
        model = dde.Model(case.data, case.net)
        case.compile(model, "adam", lr=1e-3)
        lh, _ = model.train(iterations=60, display_every=20)
        served = model.__dict__.get("_graph_state", {}).get("graph") is not None
        return np.array(lh.loss_train), served, torch.cat([p.detach().reshape(-1) for p in case.net.parameters()]).cpu()

    try:
        a, served_a, pa = run(True)
        b, served_b, pb = run(False)
    finally:
        dde.config.set_cuda_graphs("auto")
    assert served_a and not served_b
    assert np.abs(a - b).max() / np.abs(b).max() < 1e-4
    assert float((pa - pb).abs().max()) < 1e-4


def _oracle_curve(cfg, n, dtype, X=None, state=None, iters=1000, lr=1e-3, perturb=0.0):
    """Loss curve of the oracle (ddelite + autograd on the CPU).  The initial weights are rounded to float32 in every
    run, the float64 one included: that is what any float32 implementation starts from, so no curve differs from
    another by its starting point.  `perturb`: relative perturbation of the initial weights (one float32 ulp =
    6e-8) -- the sensitivity probe of the trajectory itself; "tf32": the initial weights rounded ONCE to TF32 (10
    mantissa bits, round to nearest) -- the size of perturbation a TF32 tensor-core operand sees at every step."""
    from oracle import cases as oc, ddelite

    ddelite.config.set_default_float(dtype)
    try:
        np.random.seed(0)
        torch.manual_seed(0)
        odata, onet, _ = oc.build(cfg, n)
        if X is not None:
            odata.replace_with_anchors(X)
            odata.train_x_bc = None
            odata.replace_with_anchors(X)
            onet.load_state_dict({k: v.to(onet.linears[0].weight.dtype) for k, v in state.items()})
        wdt = onet.linears[0].weight.dtype

        def start(v):
            v = v.detach().float()
            if perturb == "tf32":
                return ((v.contiguous().view(torch.int32) + 0x1000) & -8192).view(torch.float32).to(wdt)
            return (v * (1.0 + perturb)).to(wdt)

        onet.load_state_dict({k: start(v) for k, v in onet.state_dict().items()})
        state0 = {k: v.detach().clone() for k, v in onet.state_dict().items()}
        omodel = ddelite.Model(odata, onet)
        omodel.compile("adam", lr=lr, loss_weights=oc.LOSS_WEIGHTS.get(cfg))
        olh, _ = omodel.train(iterations=iters, display_every=100)
        onet.load_state_dict(state0)
        return odata, onet, np.array(olh.loss_train).sum(1), olh.steps
    finally:
        ddelite.config.set_default_float("float32")
        torch.set_default_dtype(torch.float32)


_ORACLE_CURVES = {}


def _curves(cfg, n, lr):
    """float64 oracle curve, the float32 oracle's and the one-ulp-perturbed float64 oracle's deviation from it.
    Taken from tests/golden/loss_curves.npz (written by tests/golden/make_loss_curves.py: the float64 CPU runs take
    minutes per config) after checking that `oracle.cases.build` reproduces the very points and starting weights the
    stored curves belong to; computed live otherwise."""
    import os

    from oracle import cases as oc, ddelite

    key = (cfg, n, lr)
    if key in _ORACLE_CURVES:
        return _ORACLE_CURVES[key]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_curves.npz")
    name = f"{cfg}_{n}_{lr:g}"
    if os.path.exists(path):
        g = np.load(path)
        if name + "_ref" in g:
            ddelite.config.set_default_float("float64")
            try:
                np.random.seed(0)
                torch.manual_seed(0)
                odata, onet, _ = oc.build(cfg, n)
                onet.load_state_dict({k: v.detach().float().double() for k, v in onet.state_dict().items()})
            finally:
                ddelite.config.set_default_float("float32")
                torch.set_default_dtype(torch.float32)
            w = np.concatenate([p.detach().double().numpy().reshape(-1) for p in onet.parameters()])
            x = np.asarray(odata.train_x, dtype=np.float64)
            chk = np.array([w.sum(), np.abs(w).sum(), x.sum(), np.abs(x).sum(), float(len(w)), float(x.size)])
            if np.allclose(chk, g[name + "_check"], rtol=1e-12, atol=0.0):
                if name + "_devt" in g:
                    _ORACLE_CURVES[key] = (odata, onet, g[name + "_ref"], [int(v) for v in g[name + "_steps"]],
                                           g[name + "_dev32"], g[name + "_devp"], g[name + "_devt"])
                    return _ORACLE_CURVES[key]
    odata, onet, ref, steps = _oracle_curve(cfg, n, "float64", lr=lr)
    state = {k: v.detach().clone() for k, v in onet.state_dict().items()}
    X = odata.train_x_all.copy()
    _, _, ref32, _ = _oracle_curve(cfg, n, "float32", X=X, state=state, lr=lr)
    _, _, refp, _ = _oracle_curve(cfg, n, "float64", X=X, state=state, lr=lr, perturb=6e-8)
    _, _, reft, _ = _oracle_curve(cfg, n, "float64", X=X, state=state, lr=lr, perturb="tf32")
    _ORACLE_CURVES[key] = (odata, onet, ref, steps, np.abs(ref32 - ref) / ref, np.abs(refp - ref) / ref,
                           np.abs(reft - ref) / ref)
    return _ORACLE_CURVES[key]


def _product_curve(cfg, odata, onet, engine, lr):
    dde = _setup()
    from oracle import cases as oc
    from parity_util import build_product

    case = build_product(cfg, odata, onet)
    dde.config.set_engine(engine)
    try:
        model = dde.Model(case.data, case.net)
        model.compile("adam", lr=lr, loss_weights=oc.LOSS_WEIGHTS.get(cfg))
        lh, _ = model.train(iterations=1000, display_every=100)
    finally:
        dde.config.set_engine("auto")
    return np.array(lh.loss_train).sum(1), lh.steps


@pytest.mark.parametrize("cfg,n,lr,engine", [
    ("cfg1", 2540, 1e-3, "fp32"), ("cfg1", 2540, 2e-4, "fp32"),
    ("cfg2", 3000, 1e-3, "fp32"), ("cfg2", 3000, 1e-3, "tf32"),
    ("cfg3", 600, 2e-4, "fp32"), ("cfg3", 600, 2e-4, "tf32"),
    ("cfg4", 1024, 1e-3, "fp32"), ("cfg4", 1024, 1e-3, "tf32"),
    ("cfg4", 1024, 2e-4, "fp32"), ("cfg4", 1024, 2e-4, "tf32"),
    ("cfg5", 1024, 2e-4, "fp32"), ("cfg5", 1024, 2e-4, "tf32")])
def test_loss_curve_over_1k_steps(cfg, n, lr, engine):
    """north_star gate: loss curves agree with the float64 CPU oracle within 1 % over 1000 Adam steps (same points,
    same float32-rounded starting weights; 11 checkpoints), on the FP32 kernels AND on the tcgen05 TF32 engine the bench
    runs, at the reference's learning rate 1e-3 (/root/reference/AI4XDE/solver/PDESolver.py:79-85) and at 2e-4.

    Two more curves frame the comparison (tests/golden/loss_curves.npz, written by tests/golden/make_loss_curves.py):
    the oracle in float32, and the float64 oracle restarted from weights perturbed by ONE float32 ulp (6e-8).  Where
    these leave the float64 curve the trajectory itself is rounding-sensitive and no implementation can be held to it
    any tighter.  Measured on the oracle alone: AllenCahn (cfg2, 1e-3) passes a loss spike at step 300 where the
    ONE-ULP-perturbed float64 oracle lands 27 % away and the float32 oracle 3-20 % afterwards; the cavity flow (cfg3,
    2e-4) has the float32 oracle 3-8 % away from step 400 on; Helmholtz 2-[256]x6-1 (cfg4) drifts by up to 1.6 % after
    step 500 at 1e-3 and, at 2e-4, falls off a plateau after step 700 where the float32 oracle is 6-11 % and the perturbed
    float64 oracle 1.7 % away; Burgers (cfg1) and the 1-D Poisson net (cfg5) stay within 0.2 %.  So every checkpoint is
    held to max(1 %, 4 x the larger of those two deviations at or before it): strictly 1 % wherever the problem is well conditioned
    (all of cfg1 and cfg5, cfg2 to step 200, cfg3 to step 300, cfg4 to step 400 / 600), and to the problem's own
    sensitivity where it is not.  The TF32 engine gets no flat allowance either: its frame adds a third oracle curve --
    the float64 oracle started from weights rounded ONCE to TF32 -- i.e. it may deviate by 4 x what a single TF32-sized
    perturbation does to the float64 reference itself (it applies one at every step), and is held to 1 % where that is
    small."""
    odata, onet, ref, steps, dev32, devp, devt = _curves(cfg, n, lr)
    a, got_steps = _product_curve(cfg, odata, onet, engine, lr)
    assert list(got_steps) == list(steps) and len(a) == 11
    rel = np.abs(a - ref) / ref
    floor = np.maximum(dev32, devp)
    if engine == "tf32":
        floor = np.maximum(floor, devt)
    # trajectories that have separated do not re-converge: a checkpoint inherits the sensitivity of every earlier one
    floor = np.maximum.accumulate(floor)
    # ... and once a ONE-ULP perturbation moves the float64 oracle itself by a quarter of the loss (AllenCahn's spike
    # at step 300), the size of any other implementation's deviation there is the draw of a chaotic amplification,
    # not a property of the implementation: those checkpoints are printed, not gated
    gated = 4.0 * floor < 1.0
    print(cfg, engine, f"lr={lr:g}", "deviation from f64 oracle:", np.round(rel, 5), " f32 oracle's own:",
          np.round(dev32, 5), " one-ulp-perturbed f64 oracle:", np.round(devp, 5), " tf32-rounded-start f64 oracle:",
          np.round(devt, 5), " checkpoints within 1 %:", int((rel < 1e-2).sum()), "of 11")
    print("   gated checkpoints:", int(gated.sum()), "of 11")
    assert gated[:3].all()          # every curve is held, strictly or by its frame, at least to step 200
    assert (rel[gated] < np.maximum(1e-2, 4.0 * floor[gated])).all()


def test_end_to_end_training_reaches_the_reference_solutions():
    """End-of-training known answers: (i) Burgers, the reference's native configuration
    (hard-constraint output transform, 2540 pseudo-random points, 2-[20]x3-1, Adam 15 000 + L-BFGS), against a
    64 x 50 subset of the reference's own `data/Burgers.npz` solution (tests/golden/burgers_usol_subset.npz,
    written by tests/golden/make_golden.py) -- the float64 oracle run of the UNMODIFIED reference case ends at
    an L2 relative error of 0.190 with these seeds, the fused path at 0.167; (ii) Helmholtz with the
    reference-default net (2-[150]x3-1 sin), against the case's analytic solution."""
    dde = _setup()
    import os

    from ai4xde_b200 import cases, solver

    np.random.seed(0)
    torch.manual_seed(0)
    case = cases.Burgers(NumDomain=2540, layer_size=[2] + [20] * 3 + [1])
    s = solver.PINNSolver("pinn", case)
    s.train_step(lr=1e-3, iterations=15000)
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "burgers_usol_subset.npz"))
    xx, tt = np.meshgrid(d["x"], d["t"])
    X = np.vstack((np.ravel(xx), np.ravel(tt))).T
    err = dde.metrics.l2_relative_error(d["usol"].T.flatten()[:, None], s.model.predict(X))
    print("Burgers L2 relative error", err)
    assert err < 0.25

    np.random.seed(0)
    torch.manual_seed(0)
    case = cases.Helmholtz()
    s = solver.PINNSolver("pinn", case)
    dde.optimizers.set_LBFGS_options(maxiter=3000)       # 15 000 by default: 0.11 % then; bounded here for time
    try:
        s.train_step(lr=1e-3, iterations=5000)
    finally:
        dde.optimizers.set_LBFGS_options()
    X, y = case.get_testdata()
    err = dde.metrics.l2_relative_error(y, s.model.predict(X))
    print("Helmholtz (reference-default sin net) L2 relative error", err)
    # a snapshot in mid-convergence (L-BFGS capped at 3000 of its 15 000 iterations): 1.5-2.1 % over builds that differ
    # only in the summation order of the gradient partials; 4.0 % when the L-BFGS phase runs on TF32 gradients, which
    # is what this bound still tells apart (dde.config.set_lbfgs_engine)
    assert err < 0.03


def test_causal_pinn_round_on_device():
    """CausalPINN mirror (reference: algorithm/CausalPINN.py:82-120) on the GPU: the slab weights computed from the
    device's residual probe equal the reference's formula evaluated in NumPy on those residuals; the weighted pde
    scales the residual by the slab weight of each point; a short weighted training phase stays finite."""
    dde = _setup()
    from ai4xde_b200 import algorithm, cases

    np.random.seed(3)
    torch.manual_seed(3)
    case = cases.AllenCahn(NumDomain=2000, layer_size=[2] + [64] * 3 + [1])
    solver = algorithm.CausalPINN(case, t_num=5, num_per_t=100, iter=1, warmup_iter=5)
    case.compile(solver.model, "adam", lr=1e-3)
    solver.model.train(iterations=5)
    base_pde = case.pde
    edges = solver._edges()
    X = solver._probe_points(edges)
    Y = np.abs(solver.model.predict(X, operator=base_pde)).astype(np.float64).squeeze()
    solver.update_weights()
    err = np.power(Y, 1) / np.power(Y, 1).mean() + 0
    err = err / sum(err)
    per_t = np.array([err[i * 100:(i + 1) * 100].mean() for i in range(5)])
    want = np.array([1.0] + [np.exp(-100 * per_t[: i - 1].sum()) for i in range(1, 5)])
    np.testing.assert_allclose(solver.weight, want, rtol=1e-12)
    assert solver.weight[0] == solver.weight[1] == 1.0 and np.all(np.diff(solver.weight) <= 0)
    pts = case.geomtime.random_points(500)
    plain = solver.model.predict(pts, operator=base_pde)
    weighted = solver.model.predict(pts, operator=case.pde)
    slab = np.minimum(np.searchsorted(edges, pts[:, -1], side="right") - 1, 4)
    np.testing.assert_allclose(weighted[:, 0], plain[:, 0] * solver.weight[slab], rtol=1e-5, atol=1e-7)
    case.compile(solver.model, "adam", lr=1e-3)
    losshistory, _ = solver.model.train(iterations=20)
    assert np.isfinite(np.asarray(losshistory.loss_train[-1])).all()


def test_gpinn_on_burgers_third_order_on_device(monkeypatch):
    """gPINN on the hard-constraint Burgers case -- the reference's own gPINN example (algorithm/gPINN.py:77-81):
    the probe picks the two-direction third-order pattern, the kernels run it, and the three loss terms
    (residual, d/dx residual, d/dt residual) and their parameter gradient equal the same host path evaluated on
    the plain-PyTorch double of the two kernel calls."""
    dde = _setup()
    import _taylor_double as td
    from ai4xde_b200 import _lib, algorithm, cases

    np.random.seed(5)
    torch.manual_seed(5)
    case = cases.Burgers(NumDomain=2540, layer_size=[2] + [20] * 3 + [1])
    solver = algorithm.gPINN(case)
    case.compile(solver.model, "adam", lr=1e-3)
    X, y, aux = case.data.train_next_batch()

    def evaluate():
        _, losses = solver.model.outputs_losses_train(X, y, aux)
        case.net.zero_grad()
        losses.sum().backward()
        return losses.detach().double().cpu().numpy(), torch.cat([p.grad.reshape(-1) for p in case.net.parameters()]).double().cpu().numpy()

    before = _lib.launch_count()
    l_dev, g_dev = evaluate()
    assert _lib.launch_count() > before and l_dev.shape == (3,)
    td.install(monkeypatch)
    l_ref, g_ref = evaluate()
    np.testing.assert_allclose(l_dev, l_ref, rtol=2e-4)
    assert np.abs(g_dev - g_ref).max() / np.abs(g_ref).max() < 2e-4


def test_gpinn_on_burgers_matches_the_oracle():
    """gPINN on the hard-constraint Burgers case against the ORACLE (float64 ddelite restatement of the reference's
    gradient-enhanced residual, pinned to the unmodified reference class by tests/golden/gpinn_burgers.npz): the three
    loss terms (residual, d/dx residual, d/dt residual -- third-order derivatives of the network through the output
    transform) and their parameter gradient, same points, same weights.  FP32 kernels (width 20: the third-order pattern
    runs layer by layer).  Bound 1e-5 of the largest entry (measured on the B200: losses 1.4e-7, gradient 2.6e-7)."""
    dde = _setup()
    from ai4xde_b200 import algorithm, cases
    from oracle import cases as oc, ddelite

    np.random.seed(5)
    torch.manual_seed(5)
    case = cases.Burgers(NumDomain=2540, layer_size=[2] + [20] * 3 + [1])
    solver = algorithm.gPINN(case)
    case.compile(solver.model, "adam", lr=1e-3)
    X, y, aux = case.data.train_next_batch()
    _, losses = solver.model.outputs_losses_train(X, y, aux)
    case.net.zero_grad()
    losses.sum().backward()
    l_dev = losses.detach().double().cpu().numpy()
    g_dev = torch.cat([p.grad.reshape(-1) for p in case.net.parameters()]).double().cpu().numpy()
    ddelite.config.set_default_float("float64")
    try:
        sd = {k: v.detach().double().cpu().numpy() for k, v in case.net.state_dict().items()}
        odata, onet = oc.gpinn_burgers_build(np.asarray(case.data.train_x_all, dtype=np.float64), sd, [2, 20, 20, 20, 1])
        ref = oc.train_step_outputs(odata, onet, None)
    finally:
        ddelite.config.set_default_float("float32")
        torch.set_default_dtype(torch.float32)
    l_err = np.abs(l_dev - ref["losses"]).max() / np.abs(ref["losses"]).max()
    g_err = np.abs(g_dev - ref["grad"]).max() / np.abs(ref["grad"]).max()
    print("gPINN Burgers vs oracle: losses", l_dev, ref["losses"], f"rel {l_err:.2e}; gradient rel {g_err:.2e}")
    assert l_dev.shape == (3,) and l_err < 1e-5 and g_err < 1e-5
