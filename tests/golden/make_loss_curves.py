"""Generate tests/golden/loss_curves.npz: 1000-step Adam loss curves of the CPU oracle (oracle.ddelite + autograd)
for the BASELINE configs, at the reference's learning rate 1e-3 and at 2e-4:
  ref    float64 oracle (the reference's default float, /root/reference/AI4XDE/solver/PDESolver.py:6)
  dev32  |float32 oracle - ref| / ref                      (what plain float32 autograd does on the same problem)
  devp   |float64 oracle from weights perturbed by one float32 ulp - ref| / ref   (sensitivity of the trajectory)
  devt   |float64 oracle from weights rounded once to TF32 - ref| / ref           (... to a TF32-sized perturbation)
plus checksums of the points and starting weights (`oracle.cases.build` under np.random.seed(0) /
torch.manual_seed(0), weights rounded to float32), which the GPU tests verify before they compare against the
stored curves.  The float64 runs take minutes per config on the host cores; the GPU suite must not.
Run here:  python tests/golden/make_loss_curves.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = [("cfg1", 2540), ("cfg2", 3000), ("cfg3", 600), ("cfg4", 1024), ("cfg5", 1024)]
LRS = {"cfg1": (1e-3, 2e-4), "cfg2": (1e-3,), "cfg3": (2e-4,), "cfg4": (1e-3, 2e-4), "cfg5": (2e-4,)}


def checksum(odata, onet):
    w = np.concatenate([p.detach().double().numpy().reshape(-1) for p in onet.parameters()])
    x = np.asarray(odata.train_x, dtype=np.float64)
    return np.array([w.sum(), np.abs(w).sum(), x.sum(), np.abs(x).sum(), float(len(w)), float(x.size)])


def main():
    from test_gpu_algorithms import _oracle_curve

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "loss_curves.npz")
    out = dict(np.load(path)) if os.path.exists(path) else {}      # curves already there are kept
    for cfg, n in CASES:
        for lr in LRS[cfg]:
            key = f"{cfg}_{n}_{lr:g}"
            have = key + "_ref" in out
            odata, onet, ref, steps = _oracle_curve(cfg, n, "float64", lr=lr, iters=0 if have else 1000)
            state = {k: v.detach().clone() for k, v in onet.state_dict().items()}
            X = odata.train_x_all.copy()
            if have:
                assert np.allclose(checksum(odata, onet), out[key + "_check"], rtol=1e-12, atol=0.0), key
                ref = out[key + "_ref"]
            else:
                _, _, ref32, _ = _oracle_curve(cfg, n, "float32", X=X, state=state, lr=lr)
                _, _, refp, _ = _oracle_curve(cfg, n, "float64", X=X, state=state, lr=lr, perturb=6e-8)
                out[key + "_ref"] = ref
                out[key + "_steps"] = np.array(steps)
                out[key + "_dev32"] = np.abs(ref32 - ref) / ref
                out[key + "_devp"] = np.abs(refp - ref) / ref
                out[key + "_check"] = checksum(odata, onet)
            if key + "_devt" not in out:
                _, _, reft, _ = _oracle_curve(cfg, n, "float64", X=X, state=state, lr=lr, perturb="tf32")
                out[key + "_devt"] = np.abs(reft - ref) / ref
            print(key, "ref", np.round(ref, 5), "\n   dev32", np.round(out[key + "_dev32"], 5), "\n   devp",
                  np.round(out[key + "_devp"], 6), "\n   devt", np.round(out[key + "_devt"], 5), flush=True)
            np.savez_compressed(path, **out)


if __name__ == "__main__":
    main()
