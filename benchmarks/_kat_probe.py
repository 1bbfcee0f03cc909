import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); warnings.simplefilter("ignore")
import numpy as np, torch
import ai4xde_b200.dde as dde
from ai4xde_b200 import cases, solver
np.random.seed(0); torch.manual_seed(0)
case = cases.Burgers(NumDomain=2540, layer_size=[2] + [20] * 3 + [1])
s = solver.PINNSolver("pinn", case)
t0 = time.time()
case.compile(s.model, "adam", lr=1e-3); s.model.train(iterations=15000, display_every=5000)
torch.cuda.synchronize(); t1 = time.time()
case.compile(s.model, "L-BFGS"); s.model.train()
torch.cuda.synchronize(); t2 = time.time()
d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/burgers_usol_subset.npz"))
xx, tt = np.meshgrid(d["x"], d["t"])
X = np.vstack((np.ravel(xx), np.ravel(tt))).T
y = d["usol"].T.flatten()[:, None]
print("Burgers: Adam 15000 steps %.1f s; L-BFGS %d iterations %.1f s (graph replay: %s); L2 rel err %.4f" % (
    t1 - t0, s.model.opt.state_dict()["state"][0]["n_iter"], t2 - t1, s.model.opt._graph is not None,
    dde.metrics.l2_relative_error(y, s.model.predict(X))))
