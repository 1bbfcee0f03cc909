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
    """|residual| inputs of the selectors: float32 device matrix [n_res, ldr] (zero padded).  `X`: host point array,
    or a device batch (dde/model.py:_DeviceBatch) when the candidates were generated on the device."""
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
        u = rand(size - n_found)          # NumPy's stream (parity) or a float64 CUDA tensor (device sampler)
        if not torch.is_tensor(u):
            u = torch.as_tensor(u, dtype=torch.float64)
        u = u.to(device=dev, dtype=torch.float64)
        _lib.choice_round(p, u, found, n_found, n_found_dev, ws, wsb)
        n_found = int(n_found_dev.item())
    return found


def rad_select(model, X, pde, num, k=1, c=1, rand=None, fast=False):
    """One RAD / RAR_D selection (RAD.py:26-39): indices into X (host int64 array).  `fast`: the two float64
    reductions whose order NumPy fixes (builtin sum, cumsum) run as parallel scans (`ax_resample_set_fast`)."""
    res, n = residual_matrix(model, X, pde)
    prev = _lib.resample_get_fast()
    _lib.resample_set_fast(fast)
    try:
        p = _lib.rad_probabilities(res, n, k, c)
        return choice_noreplace_device(p, num, rand).cpu().numpy()
    finally:
        _lib.resample_set_fast(prev)


def rad_round_on_device(model, geom_box, pde, n_candidates, num, k, c, seed, offset):
    """A whole RAD / RAR_D selection without the host in the data path ([ext], performance mode): candidates from
    the library's Philox box sampler (`ax_uniform_box_points`), residual probe on them where they lie, probabilities
    and the without-replacement draw with the parallel float64 scans (`ax_resample_set_fast`), uniforms of the draw
    from torch's device generator, selected points gathered on the device (`ax_gather_points`).
    Returns (indices into the candidate set [num] int64 CUDA, selected points SoA [d, pad32(num)], AoS [num, d])."""
    from .dde.model import _DeviceBatch

    dev = torch.device("cuda", torch.cuda.current_device())
    d, ldc = len(geom_box[0]), _lib.pad32(n_candidates)
    cand = torch.zeros((d, ldc), dtype=torch.float32, device=dev)
    _lib.uniform_box_points(cand, 0, n_candidates, geom_box[0], geom_box[1], seed, offset)
    batch = _DeviceBatch.on_device(cand[:, :n_candidates].t().contiguous(), cand)
    res, n = residual_matrix(model, batch, pde)
    gen = torch.Generator(device=dev)
    gen.manual_seed((int(seed) * 0x9E3779B1 + int(offset)) & 0x7FFFFFFFFFFFFFFF)
    prev = _lib.resample_get_fast()
    _lib.resample_set_fast(True)
    try:
        p = _lib.rad_probabilities(res, n, k, c)
        ids = choice_noreplace_device(p, num, lambda m: torch.rand(m, dtype=torch.float64, device=dev, generator=gen))
    finally:
        _lib.resample_set_fast(prev)
    sel = _lib.gather_points(cand, ids)
    return ids, sel, sel[:, :num].t().contiguous()


def r3_retain(model, X, pde):
    """R3 retain step (R3Sampling.py:77-93).  Returns (keep_mask host bool [n], threshold,
    retained points on the device as SoA [d, ldn] placed at the END of the buffer, n_keep)."""
    res, n = residual_matrix(model, X, pde)
    batch = model.device_batch(X)
    x_out, n_keep, thr, res_sum = _lib.r3_retain(res, batch.soa, n, place_at_end=True)
    keep = (res_sum > thr).cpu().numpy()
    return keep, float(thr.item()), x_out, int(n_keep.item())


# ------------------------------------------------------------------------------------------------
# Point-sharded selection.  The expensive part -- the residual probe over the candidate
# set / the population -- is sharded: every rank evaluates its own contiguous block.  What crosses
# NVLink is only the block residuals (4 bytes per point, all-gathered in rank order); the selection
# arithmetic is then REPLICATED on every rank over the global residual vector, in the same kernels and
# therefore in the same NumPy rounding order as a single process, so the selected global index set is
# bit-identical to the single-process one on the same residuals and uniform stream, on every rank.
# ------------------------------------------------------------------------------------------------
def _world():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def block_bounds(n, rank, world):
    """Contiguous row block of rank `rank` when n rows are split over `world` ranks (np.array_split order)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allgather_rows(local, n_local, n_total):
    """[n_res, >= n_local] per-rank blocks -> [n_res, pad32(n_total)] global matrix, rank order, zero padded."""
    dist, rank, world = _world()
    n_res = local.shape[0]
    out = torch.zeros((n_res, _lib.pad32(n_total)), dtype=local.dtype, device=local.device)
    if dist is None:
        out[:, :n_local] = local[:, :n_local]
        return out
    counts = [block_bounds(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in counts)
    send = torch.zeros((n_res, width), dtype=local.dtype, device=local.device)
    send[:, :n_local] = local[:, :n_local]
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send)
    for (lo, hi), blk in zip(counts, recv):
        out[:, lo:hi] = blk[:, :hi - lo]
    return out


def rad_select_sharded(model, X, pde, num, k=1, c=1, rand=None):
    """RAD / RAR_D selection with the candidate probe sharded over ranks.  `X` is the GLOBAL candidate array
    and `rand` the uniform source; both must be identical on every rank (a shared-seed stream, see
    `shared_numpy_stream`).  Returns the global index array (identical on every rank)."""
    dist, rank, world = _world()
    n = len(X)
    lo, hi = block_bounds(n, rank, world)
    res_local, n_local = residual_matrix(model, X[lo:hi], pde)
    res = allgather_rows(res_local, n_local, n)
    p = _lib.rad_probabilities(res, n, k, c)
    return choice_noreplace_device(p, num, rand).cpu().numpy()


def r3_retain_sharded(model, X_local, pde):
    """R3 retain on this rank's block of the population against the GLOBAL threshold: mean (NumPy pairwise
    order) of sum|residual| over the rank-ordered concatenation of all blocks.  Returns (keep mask over the
    local block, threshold, retained local points on the device, n_keep)."""
    dist, rank, world = _world()
    res_local, n_local = residual_matrix(model, X_local, pde)
    res_sum = _lib.r3_abs_sum(res_local, n_local)
    if dist is None:
        n_total, glob = n_local, res_sum
    else:
        sizes = torch.zeros((world,), dtype=torch.int64, device=res_sum.device)
        sizes[rank] = n_local
        dist.all_reduce(sizes)
        sizes = [int(v) for v in sizes.tolist()]
        n_total, width = sum(sizes), max(sizes)
        send = torch.zeros((width,), dtype=torch.float32, device=res_sum.device)
        send[:n_local] = res_sum
        recv = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(recv, send)
        glob = torch.cat([blk[:m] for blk, m in zip(recv, sizes)])
    thr = _lib.pairwise_sum(glob) / torch.tensor(float(n_total), dtype=torch.float32, device=glob.device)
    batch = model.device_batch(X_local)
    x_out, n_keep = _lib.r3_retain_thr(res_sum, thr, batch.soa, n_local, place_at_end=True)
    keep = (res_sum > thr).cpu().numpy()
    return keep, float(thr.item()), x_out, int(n_keep.item())


class shared_numpy_stream:
    """Context manager: run NumPy's GLOBAL legacy generator on a stream that is identical on every rank
    (seeded once), and restore the rank's own stream afterwards.  RAD / RAR_D draw their candidates and
    uniforms inside it when the point set is sharded, so every rank sees the same global candidate array."""

    def __init__(self, seed=20240229):
        self._state = np.random.RandomState(seed).get_state()

    def __enter__(self):
        self._own = np.random.get_state()
        np.random.set_state(self._state)
        return self

    def __exit__(self, *exc):
        self._state = np.random.get_state()
        np.random.set_state(self._own)
        return False
