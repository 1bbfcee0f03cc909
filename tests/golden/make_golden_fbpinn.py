"""Generate tests/golden/fbpinn_multiscale.npz: the UNMODIFIED reference FBPINN class
(/root/reference/AI4XDE/algorithm/FBPINN.py) on oracle.ddelite in float64, on the 1-D multiscale Poisson
problem with the sigmoid window wired (BASELINE config 5 at a small size).

Run here (the container that has /root/reference):  python tests/golden/make_golden_fbpinn.py
Recorded for sub-models 0 and 2 of a 4-subdomain decomposition: the decomposition limits, every sub-case's
points and weights, and what the reference's OWN patched closure (FBPINN.py:143-185, get_predict :196-204)
returns on sub-model k's batch -- outputs, per-term losses, and the parameter gradients after
`total_loss.backward()`.  tests/test_oracle_pins.py holds oracle.cases.fbpinn_build / fbpinn_step_outputs (the
restatement the GPU parity test of config 5 compares the kernels with) to these numbers.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SEGMENT, WIDTH, NUM_DOMAIN = 4, 16, 48


def main():
    A = ref_import.load()
    import deepxde as dde  # = oracle.ddelite
    import deepxde.backend as bkd

    class Multiscale(A.cases.PoissonCases.PoissonCase1D):
        # forcing / solution / boundary condition of Poisson_1D_Fourier_Net (cases/PoissonCases.py:469-493) on the
        # BASE gen_net, so that use_output_transform=True really wires FBPINN's window (FBPINN.py:35-42)
        def __init__(self, NumDomain=NUM_DOMAIN, layer_size=[1] + [WIDTH] * 3 + [1]):
            self.A, self.B = 2, 50
            super().__init__(name="multiscale", NumDomain=NumDomain, Interval=[0, 1], use_output_transform=True,
                             layer_size=layer_size, activation="tanh", initializer="Glorot uniform")

        func = A.cases.Poisson_1D_Fourier_Net.func
        sol = A.cases.Poisson_1D_Fourier_Net.sol

        def gen_data(self):
            bc = dde.icbc.DirichletBC(self.geomtime, self.sol, lambda _, on_boundary: on_boundary)
            return dde.data.PDE(self.geomtime, self.pde, bc, self.NumDomain, 2, train_distribution="pseudo",
                                solution=self.sol, num_test=100)

    np.random.seed(0)
    torch.manual_seed(0)
    solver = A.algorithm.FBPINN(Multiscale(), segment=SEGMENT, iter=1)
    out = {"limits": np.array([[p.geomtime.l, p.geomtime.r] for p in solver.domains.pdeCase_list]),
           "layer_size": np.array([1] + [WIDTH] * 3 + [1])}
    for j, model in enumerate(solver.model_list):
        out[f"points_{j}"] = model.data.train_x_all
        for name, v in model.net.state_dict().items():
            out[f"w_{j}_{name}"] = v.detach().numpy()
    for k in (0, 2):
        pde, model = solver.domains.pdeCase_list[k], solver.model_list[k]
        solver.compile(pde, model, "adam", lr=1e-3)
        for m in solver.model_list:
            m.net.zero_grad()
        outputs, losses = model.outputs_losses_train(model.data.train_x, None, None)
        torch.sum(losses).backward()
        out[f"outputs_{k}"] = outputs.detach().numpy()
        out[f"losses_{k}"] = losses.detach().numpy()
        out[f"grad_{k}"] = np.concatenate([p.grad.numpy().reshape(-1) for p in model.net.parameters()])
        out[f"train_x_{k}"] = model.data.train_x
    np.savez_compressed(os.path.join(OUT, "fbpinn_multiscale.npz"), **out)
    print({k: v.shape for k, v in out.items() if not k.startswith("w_")})


if __name__ == "__main__":
    main()
