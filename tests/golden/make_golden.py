"""Generate tests/golden/*.npz by running the UNMODIFIED reference sources
(/root/reference/AI4XDE cases and algorithms) on top of oracle.ddelite in float64.

Run here (the container that has /root/reference):  python tests/golden/make_golden.py
What is pinned: the reference's own `pde` closures, output transforms, BC sets, loss weights and
point layouts (the part of the hot path that lives in AI4XDE itself), evaluated through ddelite's
restatement of DeepXDE (the part that does not ship with the reference).  Nets are the BASELINE
depths at width 24 so the fixtures stay small; the closures do not depend on the width.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
W = 24


def flat_state(net):
    return np.concatenate([p.detach().numpy().reshape(-1) for p in net.parameters()])


def evaluate(dde, case, loss_weights):
    data, net = case.data, case.net
    model = dde.Model(data, net)
    model.compile("adam", lr=1e-3, loss_weights=loss_weights)
    net.zero_grad()
    x = torch.as_tensor(data.train_x).requires_grad_()
    net.train()
    y = net(x)
    res = data.pde(x, y)
    res = list(res) if isinstance(res, (list, tuple)) else [res]
    losses = torch.stack(data.losses_train(None, y, dde.losses.get("MSE"), x, model))
    if loss_weights is not None:
        losses = losses * torch.as_tensor(loss_weights)
    dde.grad.clear()
    losses.sum().backward()
    grad = np.concatenate([p.grad.numpy().reshape(-1) for p in net.parameters()])
    return dict(train_x=data.train_x, train_x_all=data.train_x_all, num_bcs=np.array(data.num_bcs, dtype=np.int64),
                params=flat_state(net), outputs=y.detach().numpy(),
                residuals=np.stack([r.detach().numpy() for r in res]), losses=losses.detach().numpy(), grad=grad,
                layer_size=np.array([l.in_features for l in net.linears] + [net.linears[-1].out_features]))


def main():
    A = ref_import.load()
    import deepxde as dde  # = oracle.ddelite

    C = A.cases
    out = {}

    def seeded():
        np.random.seed(0)
        torch.manual_seed(0)

    # cfg1-shaped: the reference's 2540/80/160 soft-BC layout lives in Heat; Burgers itself is the
    # hard-constraint case.  Both are pinned.
    seeded()
    out["heat_soft"] = evaluate(dde, C.Heat(NumDomain=300, layer_size=[2] + [W] * 3 + [1]), None)
    seeded()
    out["burgers_hard"] = evaluate(dde, C.Burgers(NumDomain=300, layer_size=[2] + [W] * 3 + [1]), None)
    seeded()
    out["allencahn"] = evaluate(dde, C.AllenCahn(NumDomain=300, layer_size=[2] + [W] * 4 + [1]), None)
    seeded()
    out["ns_cavity"] = evaluate(dde, C.NS_Flow_in_LidDriven_Cavity(NumDomain=256, layer_size=[2] + [W] * 5 + [3]), None)
    seeded()
    h = C.Helmholtz(layer_size=[2] + [W] * 6 + [1], activation="tanh")
    out["helmholtz_tanh"] = evaluate(dde, h, h.loss_weights)
    seeded()
    h = C.Helmholtz(layer_size=[2] + [W] * 3 + [1])  # reference default: sin activation
    out["helmholtz_sin"] = evaluate(dde, h, h.loss_weights)

    class PoissonFNN(C.Poisson_1D_Fourier_Net):      # reference pde/forcing/BC, plain FNN instead of MsFFN
        def gen_net(self, layer_size, activation, initializer):
            return dde.nn.FNN(layer_size, activation, initializer)

        def gen_data(self):
            bc = dde.icbc.DirichletBC(self.geomtime, self.sol, lambda _, on_boundary: on_boundary)
            return dde.data.PDE(self.geomtime, self.pde, bc, self.NumDomain, 2, train_distribution="pseudo",
                                solution=self.sol, num_test=100)

    seeded()
    out["poisson1d_multiscale"] = evaluate(dde, PoissonFNN(NumDomain=256, layer_size=[1] + [W] * 3 + [1]), None)

    for name, d in out.items():
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
        print(name, d["train_x"].shape, d["losses"])

    # RAD / RAR_D / R3 selections of the reference algorithms' own NumPy code on a fixed residual field
    rng = np.random.RandomState(42)
    n = 20000
    X = rng.rand(n, 2)
    res = (rng.randn(n, 1) * np.exp(rng.randn(n, 1))).astype(np.float32)

    class FakeModel:
        def predict(self, X, operator=None):
            return res.astype(np.float64)

    class FakeData:
        def __init__(self):
            self.train_x_all = X
            self.calls = []

        def replace_with_anchors(self, a):
            self.calls.append(("replace", a.copy()))
            self.train_x_all = a

        def add_anchors(self, a):
            self.calls.append(("add", a.copy()))
            self.train_x_all = np.vstack((a, self.train_x_all))

    class FakeGeom:
        def random_points(self, m):
            return X[:m] if m == n else rng2.rand(m, 2)

    class FakeCase:
        NumDomain = n // 10
        pde = None

    sel = {}
    for algo, kw in (("RAD", dict(k=1, c=1)), ("RAR_D", dict(k=2, c=0))):
        cls = getattr(A.algorithm, algo)
        solver = cls.__new__(cls)
        solver.model, solver.PDECase = FakeModel(), FakeCase()
        solver.PDECase.data, solver.PDECase.geomtime = FakeData(), FakeGeom()
        solver.iter, solver.k, solver.c, solver.warmup_iter = 1, kw["k"], kw["c"], 0
        solver.train_step = lambda *a, **k: None
        if algo == "RAR_D":
            # RAR_D draws 100000 candidates; serve the same fixed field tiled to that size
            big = np.tile(res, (5, 1))
            Xbig = np.tile(X, (5, 1)) + 0.0
            solver.model.predict = lambda X_, operator=None: big.astype(np.float64)
            solver.PDECase.geomtime.random_points = lambda m: Xbig[:m]
            FakeCase.NumDomain = 2000
        np.random.seed(7)
        solver.closure()
        picked = solver.PDECase.data.calls[-1][1]
        sel[algo + "_picked"] = picked
    np.savez_compressed(os.path.join(OUT, "sampling.npz"), X=X, res=res, **sel)
    print("sampling", {k: v.shape for k, v in sel.items()})


def burgers_solution_subset():
    """A 64 x 50 subset of the reference's own Burgers solution (/root/reference/AI4XDE/data/Burgers.npz, the
    file `Burgers.gen_testdata` loads at cases/PDECases.py:1096-1101): the end-of-training known answer of
    tests/test_gpu_algorithms.py::test_end_to_end_training_reaches_the_reference_solutions."""
    d = np.load(os.path.join(ref_import.REFERENCE_ROOT, "AI4XDE", "data", "Burgers.npz"))
    np.savez_compressed(os.path.join(OUT, "burgers_usol_subset.npz"), x=d["x"][::4], t=d["t"][::2],
                        usol=d["usol"][::4, ::2])


if __name__ == "__main__":
    rng2 = np.random.RandomState(1)
    main()
    burgers_solution_subset()
