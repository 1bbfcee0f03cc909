"""Device-resident residual-driven resampling.

The residual probe is the forward-only fused jet evaluation (`Model.predict_on_device`); the
selection arithmetic runs in the CUDA kernels of csrc/resample.cu in NumPy's rounding order, so that
-- given the same residuals and the same uniform stream -- the selected index sets equal the
reference's bit for bit.  Uniforms come from the caller's NumPy legacy stream (parity with
`np.random.choice`, /root/reference/AI4XDE/algorithm/RAD.py:33-38)."""
import numpy as np
import torch

from . import _lib


def residual_matrix(model, X, pde):
    """|residual| inputs of the selectors: float32 device matrix [n_res, ldr] (zero padded)."""
    res = model.predict_on_device(X, pde)
    res = list(res) if isinstance(res, (list, tuple)) else [res]
    n = res[0].shape[0]
    ldr = _lib.pad32(n)
    out = torch.zeros((len(res), ldr), dtype=torch.float32, device=res[0].device)
    for i, r in enumerate(res):
        out[i, :n] = r.reshape(-1)
    return out, n


def choice_noreplace_device(p, size, rand=None):
    """numpy.random.choice(len(p), size, replace=False, p=p) with p a float64 CUDA tensor.
    Consumes `rand(k)` exactly as NumPy's legacy RandomState does (one call per retry round)."""
    rand = rand or np.random.random_sample
    n = p.numel()
    if int((p > 0).sum().item()) < size:
        raise ValueError("Fewer non-zero entries in p than size")
    dev = p.device
    p = p.clone()
    found = torch.zeros((size,), dtype=torch.int64, device=dev)
    n_found_dev = torch.zeros((1,), dtype=torch.int64, device=dev)
    ws, wsb = _lib.resample_ws(n, dev)
    n_found = 0
    while n_found < size:
        u = torch.as_tensor(rand(size - n_found), dtype=torch.float64).to(dev)
        _lib.choice_round(p, u, found, n_found, n_found_dev, ws, wsb)
        n_found = int(n_found_dev.item())
    return found


def rad_select(model, X, pde, num, k=1, c=1, rand=None):
    """One RAD / RAR_D selection (RAD.py:26-39): indices into X (host int64 array)."""
    res, n = residual_matrix(model, X, pde)
    p = _lib.rad_probabilities(res, n, k, c)
    return choice_noreplace_device(p, num, rand).cpu().numpy()


def r3_retain(model, X, pde):
    """R3 retain step (R3Sampling.py:77-93).  Returns (keep_mask host bool [n], threshold,
    retained points on the device as SoA [d, ldn] placed at the END of the buffer, n_keep)."""
    res, n = residual_matrix(model, X, pde)
    batch = model.device_batch(X)
    x_out, n_keep, thr, res_sum = _lib.r3_retain(res, batch.soa, n, place_at_end=True)
    keep = (res_sum > thr).cpu().numpy()
    return keep, float(thr.item()), x_out, int(n_keep.item())
