"""Generate tests/golden/gpinn_burgers.npz: the UNMODIFIED reference gPINN class
(/root/reference/AI4XDE/algorithm/gPINN.py) on its own example, the hard-constraint Burgers case, on oracle.ddelite in
float64: points, weights, the three loss terms (residual, d/dx residual, d/dt residual) and the parameter gradient.
tests/test_oracle_pins.py holds oracle.cases.gpinn_pde / gpinn_burgers_build to it; the GPU test compares the kernels'
third-order path with that restatement.  Run here:  python tests/golden/make_golden_gpinn.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    A = ref_import.load()
    import deepxde as dde  # = oracle.ddelite

    np.random.seed(0)
    torch.manual_seed(0)
    case = A.cases.Burgers(NumDomain=200, layer_size=[2] + [16] * 3 + [1])
    solver = A.algorithm.gPINN(case)
    model, data, net = solver.model, case.data, case.net
    case.compile(model, "adam", lr=1e-3)
    net.zero_grad()
    _, losses = model.outputs_losses_train(data.train_x, None, None)
    torch.sum(losses).backward()
    out = dict(train_x=data.train_x, losses=losses.detach().numpy(),
               grad=np.concatenate([p.grad.numpy().reshape(-1) for p in net.parameters()]),
               layer_size=np.array([2, 16, 16, 16, 1]))
    for name, v in net.state_dict().items():
        out["w_" + name] = v.detach().numpy()
    np.savez_compressed(os.path.join(OUT, "gpinn_burgers.npz"), **out)
    print({k: v.shape for k, v in out.items()}, out["losses"])


if __name__ == "__main__":
    main()
