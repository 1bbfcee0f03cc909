"""Resampling selections restated on the CPU (test infrastructure, see oracle/__init__.py).

* ``legacy_choice_noreplace`` restates ``numpy.random.RandomState.choice(a, size,
  replace=False, p=p)`` (NumPy ``mtrand.pyx``, legacy MT19937 stream) -- the call made
  at /root/reference/AI4XDE/algorithm/RAD.py:33-38 and RAR_D.py:29-34.  Pinned against
  ``np.random.choice`` itself in tests/test_oracle_sampling.py.
* ``pairwise_sum`` restates NumPy's pairwise float summation (``pairwise_sum_@TYPE@`` in
  ``loops_utils.h.src``), i.e. the rounding order behind ``ndarray.mean()`` used at
  RAD.py:31 and R3Sampling.py:91.  Pinned against ``np.add.reduce``.
* ``rad_probabilities`` / ``rad_select`` / ``rar_d_select`` / ``r3_select`` restate
  RAD.py:26-40, RAR_D.py:24-35, R3Sampling.py:77-98 (with the single-residual crash at
  R3Sampling.py:79 restated as the evident intent).
"""
import numpy as np

PW_BLOCKSIZE = 128


def pairwise_sum(a):
    """NumPy's pairwise summation of a contiguous 1-D float array, in a's dtype."""
    a = np.ascontiguousarray(a).ravel()
    dt = a.dtype.type
    n = a.size
    if n < 8:
        res = dt(0.0) if n == 0 else None
        # numpy: res = 0.; for i: res += a[i]   (starts from -0.0-safe 0.)
        res = dt(0.0)
        for i in range(n):
            res = dt(res + a[i])
        return res
    if n <= PW_BLOCKSIZE:
        r = [a[j] for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = dt(r[j] + a[i + j])
            i += 8
        res = dt(dt(dt(r[0] + r[1]) + dt(r[2] + r[3])) + dt(dt(r[4] + r[5]) + dt(r[6] + r[7])))
        while i < n:
            res = dt(res + a[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return dt(pairwise_sum(a[:n2]) + pairwise_sum(a[n2:]))


def pairwise_leaves(n):
    """Leaf ranges [(beg, end)] of the pairwise recursion, left to right."""
    out = []

    def rec(beg, m):
        if m <= PW_BLOCKSIZE:
            out.append((beg, beg + m))
            return
        n2 = m // 2
        n2 -= n2 % 8
        rec(beg, n2)
        rec(beg + n2, m - n2)

    rec(0, n)
    return out


def legacy_choice_noreplace(n, size, p, rand=None):
    """np.random.choice(n, size, replace=False, p=p) on the global legacy RNG.

    rand(k) must return k float64 uniforms in [0,1) from the stream NumPy would use
    (default: np.random.random_sample)."""
    if rand is None:
        rand = np.random.random_sample
    p = np.array(p, dtype=np.float64, copy=True)
    if p.ndim != 1 or p.size != n:
        raise ValueError("'a' and 'p' must have same size")
    if np.count_nonzero(p > 0) < size:
        raise ValueError("Fewer non-zero entries in p than size")
    found = np.zeros(size, dtype=np.int64)
    n_uniq = 0
    while n_uniq < size:
        x = rand(size - n_uniq)
        if n_uniq > 0:
            p[found[0:n_uniq]] = 0
        cdf = np.cumsum(p)
        cdf /= cdf[-1]
        new = cdf.searchsorted(x, side="right")
        _, unique_indices = np.unique(new, return_index=True)
        unique_indices.sort()
        new = new.take(unique_indices)
        found[n_uniq : n_uniq + new.size] = new
        n_uniq += new.size
    return found


def rad_probabilities(Y, k, c):
    """RAD.py:29-32 / RAR_D.py:25-28.  Y: |residual| float64, shape (n,1) or (n_res,n,1)."""
    Y = np.asarray(Y, dtype=np.float64)
    err_eq = np.power(Y, k) / np.power(Y, k).mean() + c
    return (err_eq / sum(err_eq))[:, 0]


def rad_select(X, residual, num_domain, k=1, c=1, rand=None):
    """One RAD round (RAD.py:26-40): returns (ids, X[ids]).  `residual` is what
    model.predict(X, operator=pde) returned (array or list of arrays)."""
    sample_num = len(X)
    Y = np.abs(residual).astype(np.float64)
    if Y.shape[0] != sample_num:
        Y = np.sum(Y, axis=0)
    p = rad_probabilities(Y, k, c)
    ids = legacy_choice_noreplace(len(X), num_domain, p, rand)
    return ids, X[ids]


def rar_d_select(X, residual, num_add, k=2, c=0, rand=None):
    """One RAR_D round (RAR_D.py:24-35): returns (ids, X[ids]) to be add_anchors()-ed."""
    Y = np.abs(residual).astype(np.float64)
    p = rad_probabilities(Y, k, c)
    ids = legacy_choice_noreplace(len(X), num_add, p, rand)
    return ids, X[ids]


def r3_residual(residual):
    """R3Sampling.py:78-79 with the evident intent for one residual (sum over components of
    |r|); for >=2 residuals this equals np.sum(np.squeeze(res).T, axis=1)."""
    res = np.abs(residual)
    if isinstance(residual, (list, tuple)) or res.ndim == 3:
        return np.sum(np.squeeze(res, axis=-1).T, axis=1)
    return res.reshape(-1)


def r3_select(X, res):
    """R3Sampling.py:91-93: threshold = mean(res) in res' dtype; retain res > threshold.
    Returns (threshold, keep_mask, X_retained)."""
    threshold = np.mean(res)
    keep = res > threshold
    return threshold, keep, X[np.where(keep)]
