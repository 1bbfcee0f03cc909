"""Shared helper of the parity tests, smoke() and the bench: build one BASELINE config on the
oracle (CPU, float64) and on the product (fused path) with IDENTICAL points and weights, and
compare residuals, per-term losses and the flat parameter gradient."""
import numpy as np
import torch


def build_product(cfg, odata, onet, soft=True):
    """Product-side twin of oracle.cases.build(cfg): same points (as anchors), same weights."""
    import ai4xde_b200.dde as dde
    from ai4xde_b200 import cases as pc

    ls = [l.in_features for l in onet.linears] + [onet.linears[-1].out_features]
    n_all = len(odata.train_x_all)
    if cfg == "cfg1":
        case = pc.Burgers(NumDomain=0, layer_size=ls, soft_constraints=True)
        gt = case.geomtime
        data = dde.data.TimePDE(gt, case.pde, case.data.bcs, num_domain=0, anchors=odata.train_x_all,
                                train_distribution="pseudo")
    elif cfg == "cfg2":
        case = pc.AllenCahn(NumDomain=0, layer_size=ls)
        data = dde.data.TimePDE(case.geomtime, case.pde, [], num_domain=0, anchors=odata.train_x_all,
                                train_distribution="pseudo")
    elif cfg == "cfg3":
        case = pc.NS_Flow_in_LidDriven_Cavity(NumDomain=1, layer_size=ls, train_distribution="pseudo")
        data = dde.data.PDE(case.geomtime, case.pde, case.data.bcs, 0, 0, anchors=odata.train_x_all,
                            train_distribution="pseudo")
    elif cfg in ("cfg4", "cfg4_sin"):
        case = pc.Helmholtz(NumDomain=1, layer_size=ls, activation="sin" if cfg == "cfg4_sin" else "tanh",
                            honor_num_domain=True)
        data = dde.data.PDE(case.geomtime, case.pde, case.data.bcs, 0, 0, anchors=odata.train_x_all,
                            train_distribution="pseudo", solution=case.sol)
    elif cfg == "cfg5":
        case = pc.Poisson_1D_Fourier_Net(NumDomain=1, layer_size=ls)
        data = dde.data.PDE(case.geomtime, case.pde, case.data.bcs, 0, 0, anchors=odata.train_x_all,
                            train_distribution="pseudo", solution=case.sol)
    else:
        raise KeyError(cfg)
    assert len(data.train_x_all) == n_all
    case.data = data
    case.net.load_state_dict({k: v.detach().float() for k, v in onet.state_dict().items()})
    return case


def product_vs_oracle(cfg, n_domain, engine="fp32", seed=0, oracle_dtype="float64"):
    from oracle import cases as oc, ddelite
    import ai4xde_b200.dde as dde

    prev = ddelite.config.default_float()
    ddelite.config.set_default_float(oracle_dtype)
    try:
        np.random.seed(seed)
        torch.manual_seed(seed)
        odata, onet, _ = oc.build(cfg, n_domain)
        lw = oc.LOSS_WEIGHTS.get(cfg)
        ref = oc.train_step_outputs(odata, onet, lw)
    finally:
        ddelite.config.set_default_float(prev)
        torch.set_default_dtype(torch.float32)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dde.config.set_default_float("float64")  # host point sets as in the reference; device math is f32
    case = build_product(cfg, odata, onet)
    dde.config.set_engine(engine)
    try:
        model = dde.Model(case.data, case.net)
        model.compile("adam", lr=1e-3, loss_weights=lw)
        assert np.array_equal(case.data.train_x.astype(np.float64), odata.train_x.astype(np.float64))
        res = model.predict_on_device(case.data.train_x, case.pde)
        res = list(res) if isinstance(res, (list, tuple)) else [res]
        outs, losses = model.outputs_losses_train(case.data.train_x, None, None)
        case.net.zero_grad()
        losses.sum().backward()
        g = torch.cat([p.grad.reshape(-1) for p in case.net.parameters()]).double().cpu().numpy()
        # the train closure as model.train runs it: network kernels + the loss side as a traced residual program
        # (ai4xde_b200/residual.py), evaluated once without moving the weights
        case.net.zero_grad()
        seen, step = {}, model.opt.step
        model.opt.step = lambda closure: seen.__setitem__("loss", closure())
        try:
            model.train_step(case.data.train_x, None, None)
        finally:
            model.opt.step = step
        g_closure = torch.cat([p.grad.reshape(-1) for p in case.net.parameters()]).double().cpu().numpy()
        l_closure = float(seen["loss"])
    finally:
        dde.config.set_engine("auto")

    def rel(a, b):
        return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))

    def elementwise(a, b):
        """(median, 99th percentile) of |a - b| / |b| over the entries that are not tiny (|b| >= 1e-3 max|b|): the
        max-normalised figures above say nothing about small entries; these do"""
        a, b = np.asarray(a, dtype=np.float64).reshape(-1), np.asarray(b, dtype=np.float64).reshape(-1)
        keep = np.abs(b) >= 1e-3 * max(np.abs(b).max(), 1e-300)
        e = np.abs(a[keep] - b[keep]) / np.abs(b[keep])
        return (float(np.median(e)), float(np.percentile(e, 99))) if e.size else (0.0, 0.0)

    res_rel = max(rel(r.double().cpu().numpy(), rr) for r, rr in zip(res, ref["residuals"]))
    return {
        "loss_rel": rel(losses.detach().double().cpu().numpy(), ref["losses"]),
        "grad_rel": rel(g, ref["grad"]),
        "res_rel": res_rel,
        "out_rel": rel(outs.detach().double().cpu().numpy(), ref["outputs"]),
        "elementwise": {"out": elementwise(outs.detach().double().cpu().numpy(), ref["outputs"]),
                        "grad": elementwise(g, ref["grad"]),
                        "res": elementwise(np.concatenate([r.double().cpu().numpy().reshape(-1) for r in res]),
                                           np.concatenate([np.asarray(rr).reshape(-1) for rr in ref["residuals"]]))},
        "closure_loss_rel": abs(l_closure - float(np.sum(ref["losses"]))) / abs(float(np.sum(ref["losses"]))),
        "closure_grad_rel": rel(g_closure, ref["grad"]),
        "program": model.residual_program_note(),
        "n_rows": len(case.data.train_x),
        "losses": losses.detach().cpu().numpy(),
        "ref_losses": ref["losses"],
        "model": model,
        "case": case,
    }


def fbpinn_vs_oracle(segment, num_domain, engine="fp32", check=(0,), seed=0, layer_size=(1, 64, 64, 64, 1)):
    """BASELINE config 5: FBPINN with `segment` windowed subdomain nets on the 1-D multiscale Poisson problem.
    The product builds the decomposition (its own points and weights); the oracle twin
    (oracle.cases.fbpinn_build: the reference's wiring restated on ddelite, float64) is handed the same points
    and weights.  For every sub-model k in `check`: get_predict on model k's batch (the sum over ALL nets), the
    per-term losses, the residual, and the gradient of net k -- with every other net's gradient untouched."""
    from oracle import cases as oc, ddelite
    import ai4xde_b200.dde as dde
    from ai4xde_b200 import algorithm
    from ai4xde_b200.cases.poisson_cases import multiscale_fbpinn_case

    np.random.seed(seed)
    torch.manual_seed(seed)
    dde.config.set_default_float("float32")
    case = multiscale_fbpinn_case(NumDomain=num_domain, layer_size=layer_size)
    solver = algorithm.FBPINN(case, segment=segment, iter=1)
    limits = solver.domains.limits
    points = [np.asarray(m.data.train_x_all, dtype=np.float64) for m in solver.model_list]
    sds = [{k: v.detach().double().cpu().numpy() for k, v in m.net.state_dict().items()} for m in solver.model_list]
    prev = ddelite.config.default_float()
    ddelite.config.set_default_float("float64")
    try:
        datas, nets = oc.fbpinn_build(limits, points, sds, list(layer_size))
        refs = {k: oc.fbpinn_step_outputs(datas, nets, k) for k in check}
    finally:
        ddelite.config.set_default_float(prev)
        torch.set_default_dtype(torch.float32)

    def rel(a, b):
        return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))

    out = {}
    dde.config.set_engine(engine)
    try:
        for k in check:
            model, pde = solver.model_list[k], solver.domains.pdeCase_list[k]
            solver.compile(pde, model, "adam", lr=1e-3)
            assert np.array_equal(np.asarray(model.data.train_x, dtype=np.float64), datas[k].train_x)
            for m in solver.model_list:
                m.net.zero_grad()
            outs, losses = model.outputs_losses_train(model.data.train_x, None, None)
            losses.sum().backward()
            g = torch.cat([p.grad.reshape(-1) for p in model.net.parameters()]).double().cpu().numpy()
            others = [p.grad for j, m in enumerate(solver.model_list) if j != k for p in m.net.parameters()]
            x = torch.as_tensor(model.data.train_x, dtype=torch.float32, device=outs.device).requires_grad_()
            res = pde.pde(x, solver.get_predict(x, param_grad=False))
            dde.grad.clear()
            out[k] = {
                "out_rel": rel(outs.detach().double().cpu().numpy(), refs[k]["outputs"]),
                "loss_rel": rel(losses.detach().double().cpu().numpy(), refs[k]["losses"]),
                "res_rel": rel(res.detach().double().cpu().numpy(), refs[k]["residuals"][0]),
                "grad_rel": rel(g, refs[k]["grad"]),
                "others_untouched": all(t is None or float(t.abs().max()) == 0.0 for t in others),
                "n_rows": len(model.data.train_x),
            }
    finally:
        dde.config.set_engine("auto")
    return out, solver
