"""GPU parity of the C-ABI kernels (called through ai4xde_b200._lib) against the plain-PyTorch
Taylor-mode double in float64 and against NumPy for the resampling entry points."""
import numpy as np
import pytest
import torch

import _taylor_double as td
from ai4xde_b200 import _lib
from ai4xde_b200.jet import JetSpec, to_soa

pytestmark = pytest.mark.gpu

SPECS = {
    "v": ((), ()),
    "1": ((0,), ()),
    "1d": ((0,), ((0, 0),)),
    "2": ((0, 1), ()),
    "2a": ((1, 0), ((0, 0),)),
    "2d": ((0, 1), ((0, 0), (1, 1))),
    "2f": ((0, 1), ((0, 0), (0, 1), (1, 1))),
    "3d": ((0, 1, 2), ((0, 0), (1, 1), (2, 2))),
    "3f": ((0, 1, 2), ((0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2))),
    "4f": ((0, 1, 2, 3), tuple((a, b) for a in range(4) for b in range(a, 4))),
    "2t": ((0, 1), ((15, 15),)),          # Laplacian stream u_xx + u_yy
    "3t": ((0, 1, 2), ((15, 15),)),
    "1o3": ((0,), ((0, 0), (14, 14))),            # one direction to third order (gPINN on a 1-D second-order problem)
    "1o4": ((0,), ((0, 0), (14, 14), (13, 13))),  # ... to fourth order (Euler beam)
    "1o4y": ((1,), ((0, 0), (14, 14), (13, 13))), # the direction need not be input 0
    "2o3": ((0, 1), ((0, 0), (0, 1), (1, 1), (8, 8), (9, 9), (10, 10), (11, 11))),   # two directions to third order (gPINN)
    "2o3r": ((2, 0), ((0, 0), (0, 1), (1, 1), (8, 8), (9, 9), (10, 10), (11, 11))),  # ... any two inputs, any order
}

CASES = [
    # widths, activation, spec, n, n_nets
    ([2, 20, 20, 20, 1], "tanh", "2a", 3020, 1),
    ([2, 64, 64, 64, 64, 1], "tanh", "2a", 1000, 1),
    ([2, 128, 128, 3], "tanh", "2d", 777, 1),
    ([2, 256, 256, 256, 1], "tanh", "2d", 515, 1),
    ([2, 150, 150, 1], "sin", "2d", 300, 1),
    ([1, 64, 64, 64, 1], "tanh", "1d", 2048, 4),
    ([2, 50, 50, 50, 50, 3], "tanh", "2f", 333, 1),
    ([3, 33, 17, 2], "sigmoid", "3d", 100, 1),
    ([3, 40, 40, 2], "tanh", "3f", 65, 1),
    ([4, 50, 50, 4], "tanh", "4f", 130, 1),
    ([2, 32, 32, 1], "tanh", "v", 1, 1),
    ([1, 20, 20, 1], "tanh", "1", 31, 2),
    ([2, 20, 20, 1], "tanh", "2", 33, 1),
    ([2, 256, 256, 256, 1], "tanh", "2t", 515, 1),
    ([2, 150, 150, 1], "sin", "2t", 300, 1),
    ([3, 40, 40, 2], "sigmoid", "3t", 65, 2),
    ([1, 20, 20, 20, 1], "tanh", "1o4", 1003, 1),       # the reference's Euler beam net (cases/PDECases.py:428)
    ([1, 20, 20, 20, 1], "tanh", "1o3", 517, 2),
    ([1, 40, 24, 2], "sin", "1o4", 300, 1),
    ([1, 32, 32, 32, 1], "sigmoid", "1o4", 65, 3),
    ([2, 64, 64, 1], "tanh", "1o4y", 2000, 1),
    ([2, 20, 20, 20, 1], "tanh", "2o3", 2540, 1),       # gPINN on the reference's Burgers (algorithm/gPINN.py:77-81)
    ([2, 64, 64, 64, 1], "tanh", "2o3", 1001, 2),
    ([2, 48, 32, 2], "sin", "2o3", 333, 1),
    ([3, 40, 40, 1], "sigmoid", "2o3r", 130, 1),
    # narrow nets (hidden width 20 / 32): the whole-net one-thread-per-point kernels (csrc/narrow.cuh) -- every
    # chunking they are compiled with (forward OC = 20 / 4 / 32 / 16, backward KC = 4 / 8), ragged tails, several
    # nets per launch, 1..4 inputs / outputs, all three activations
    ([2, 20, 20, 20, 1], "tanh", "2t", 100003, 1),
    ([2, 20, 20, 20, 20, 3], "sin", "2d", 4099, 1),
    ([2, 20, 20, 1], "sigmoid", "2f", 1500, 2),
    ([2, 32, 32, 32, 1], "tanh", "2a", 70001, 1),       # the reference's default net (cases/PDECases.py:8-20)
    ([2, 32, 32, 32, 32, 32, 32, 1], "tanh", "2t", 3000, 1),
    ([1, 32, 32, 32, 2], "sin", "1d", 129, 3),
    ([3, 32, 32, 4], "tanh", "2", 127, 1),
    ([4, 20, 20, 20, 2], "tanh", "1", 6000, 1),
    ([1, 20, 20, 20, 1], "tanh", "1d", 128, 64),        # FBPINN-style bank of narrow nets
    ([2, 32, 32, 32, 1], "tanh", "v", 5000, 1),
    # large point sets (many point tiles per output block, ragged tails)
    ([2, 150, 150, 150, 1], "sin", "2t", 40001, 1),
    ([2, 64, 64, 3], "tanh", "2a", 38000, 2),
    ([1, 48, 48, 1], "sigmoid", "1d", 50003, 1),
]


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("widths,act,spec,n,k", CASES)
def test_jet_forward_backward_fp32(widths, act, spec, n, k):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(7)
    dirs, pairs = SPECS[spec]
    js = JetSpec(dirs, pairs)
    net = _lib.make_net_desc(widths, act, n_nets=k)
    P = _lib.param_count(net)
    params = (torch.randn((k, P), generator=g) * 0.3).to(dev)
    x = (torch.rand((n, widths[0]), generator=g) * 2 - 1).to(dev)
    x_soa = to_soa(x)
    jets, stash = _lib.jet_forward(net, js.desc, params, x_soa, n, keep_all=True, engine="fp32")
    ref, _ = td.jet_forward(net, js.desc, params.double(), x_soa.double(), n, True)
    assert _rel(jets[..., :n].double(), ref[..., :n]) < 1e-5
    # forward-only (ping-pong stash) must agree bit for bit with the stash-everything run
    jets2, _ = _lib.jet_forward(net, js.desc, params, x_soa, n, keep_all=False, engine="fp32")
    assert torch.equal(jets2[..., :n], jets[..., :n])
    jets_bar = torch.randn(jets.shape, generator=g).to(dev)
    jets_bar[..., n:] = 0
    grad = _lib.jet_backward(net, js.desc, params, x_soa, n, jets_bar, stash, engine="fp32")
    gref = td.jet_backward(net, js.desc, params.double(), x_soa.double(), n, jets_bar.double(), None)
    assert _rel(grad.double(), gref) < 2e-5
    # deterministic reduction: a second run gives the same bits
    grad2 = _lib.jet_backward(net, js.desc, params, x_soa, n, jets_bar, stash, engine="fp32")
    assert torch.equal(grad, grad2)


def test_unsupported_requests_fail_loudly():
    net = _lib.make_net_desc([2, 8, 1], "tanh")
    bad = _lib.make_jet_desc((0, 1), ((1, 1),))  # not a compiled pattern
    assert not _lib.jet_supported(bad)
    dev = torch.device("cuda:0")
    with pytest.raises(_lib.NativeLibraryError):
        _lib.jet_forward(net, bad, torch.zeros((1, 33), device=dev), torch.zeros((2, 32), device=dev), 4, True)
    with pytest.raises(_lib.NativeLibraryError):
        _lib.jet_forward(net, JetSpec((0,), ()).desc, torch.zeros((1, 33)), torch.zeros((2, 32)), 4, True)


@pytest.mark.parametrize("n", [1, 7, 8, 9, 128, 129, 1000, 4097, 100000, 2**20 + 80])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_pairwise_sum_matches_numpy(n, dtype):
    rng = np.random.RandomState(n)
    a = rng.rand(n).astype(np.float32 if dtype == torch.float32 else np.float64)
    got = _lib.pairwise_sum(torch.as_tensor(a).cuda()).item()
    want = np.add.reduce(a)
    assert got == want


@pytest.mark.parametrize("n,m,k,c,nres", [(20000, 2000, 1, 1, 1), (100000, 10000, 1, 1, 1),
                                          (100000, 20, 2, 0, 1), (1000, 999, 1, 1, 1), (5000, 500, 1, 1, 3)])
def test_rad_selection_bit_exact(n, m, k, c, nres):
    from oracle import sampling

    rng = np.random.RandomState(3)
    res = (rng.randn(nres, n, 1) * np.exp(rng.randn(nres, n, 1))).astype(np.float32)
    X = rng.rand(n, 2)
    np.random.seed(11)
    residual = res[0] if nres == 1 else [r for r in res]
    ids_ref, _ = sampling.rad_select(X, residual, m, k=k, c=c)
    p_ref = sampling.rad_probabilities(
        np.abs(res[0]).astype(np.float64) if nres == 1 else np.sum(np.abs(res).astype(np.float64), axis=0), k, c)

    dev = torch.device("cuda:0")
    ldr = _lib.pad32(n)
    r = torch.zeros((nres, ldr), dtype=torch.float32, device=dev)
    r[:, :n] = torch.as_tensor(res[:, :, 0]).to(dev)
    p = _lib.rad_probabilities(r, n, k, c)
    assert np.array_equal(p.cpu().numpy(), p_ref)
    np.random.seed(11)
    found = torch.zeros((m,), dtype=torch.int64, device=dev)
    n_found_dev = torch.zeros((1,), dtype=torch.int64, device=dev)
    ws, wsb = _lib.resample_ws(n, dev)
    n_found = 0
    while n_found < m:
        u = torch.as_tensor(np.random.random_sample(m - n_found)).to(dev)
        _lib.choice_round(p, u, found, n_found, n_found_dev, ws, wsb)
        n_found = int(n_found_dev.item())
    assert np.array_equal(found.cpu().numpy(), ids_ref)


@pytest.mark.parametrize("n,nres", [(1000, 1), (100000, 1), (2**20 + 80, 1), (4097, 3)])
def test_r3_retain_bit_exact(n, nres):
    from oracle import sampling

    rng = np.random.RandomState(5)
    res = (rng.randn(nres, n, 1) * np.exp(rng.randn(nres, n, 1))).astype(np.float32)
    X = rng.rand(n, 2).astype(np.float32)
    rs = sampling.r3_residual(res[0] if nres == 1 else res)
    thr, keep, Xr = sampling.r3_select(X, rs)

    dev = torch.device("cuda:0")
    ldn = _lib.pad32(n)
    r = torch.zeros((nres, ldn), dtype=torch.float32, device=dev)
    r[:, :n] = torch.as_tensor(res[:, :, 0]).to(dev)
    x_soa = to_soa(torch.as_tensor(X).to(dev))
    x_out, n_keep, thr_d, res_sum = _lib.r3_retain(r, x_soa, n, place_at_end=True)
    nk = int(n_keep.item())
    assert thr_d.item() == thr
    assert np.array_equal(res_sum.cpu().numpy(), rs.astype(np.float32))
    assert nk == int(keep.sum())
    got = x_out[:, n - nk : n].t().cpu().numpy()
    assert np.array_equal(got, Xr)


@pytest.mark.parametrize("n,nres,cuts", [(100000, 1, (0, 31000, 100000)), (4097, 3, (0, 5, 2000, 4097))])
def test_r3_sharded_halves_equal_fused_entry(n, nres, cuts):
    """ax_r3_abs_sum per block + ax_pairwise_sum over the rank-ordered concatenation + ax_r3_retain_thr per
    block reproduce ax_r3_retain on the whole population (threshold and retained points, bit for bit)."""
    rng = np.random.RandomState(7)
    res = (rng.randn(nres, n) * np.exp(rng.randn(nres, n))).astype(np.float32)
    X = rng.rand(n, 2).astype(np.float32)
    dev = torch.device("cuda:0")
    r = torch.zeros((nres, _lib.pad32(n)), dtype=torch.float32, device=dev)
    r[:, :n] = torch.as_tensor(res).to(dev)
    x_out, n_keep, thr_d, res_sum = _lib.r3_retain(r, to_soa(torch.as_tensor(X).to(dev)), n, place_at_end=True)
    nk = int(n_keep.item())
    whole = x_out[:, n - nk:n].cpu().numpy()
    sums = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        blk = torch.zeros((nres, _lib.pad32(hi - lo)), dtype=torch.float32, device=dev)
        blk[:, :hi - lo] = r[:, lo:hi]
        sums.append(_lib.r3_abs_sum(blk, hi - lo))
    glob = torch.cat(sums)
    assert torch.equal(glob, res_sum)
    thr = _lib.pairwise_sum(glob) / torch.tensor(float(n), dtype=torch.float32, device=dev)
    assert thr.item() == thr_d.item()
    parts = []
    for (lo, hi), sm in zip(zip(cuts[:-1], cuts[1:]), sums):
        xs = to_soa(torch.as_tensor(X[lo:hi]).to(dev))
        xo, k = _lib.r3_retain_thr(sm, thr, xs, hi - lo, place_at_end=True)
        k = int(k.item())
        parts.append(xo[:, (hi - lo) - k:(hi - lo)].cpu().numpy())
    assert np.array_equal(np.concatenate(parts, axis=1), whole)


@pytest.mark.parametrize("n,m", [(20000, 2000), (100000, 10000), (1000, 999), (2**20, 2**16)])
def test_rad_fast_mode_selects_the_same_indices(n, m):
    """`ax_resample_set_fast(1)`: the sequential float64 sum / cumsum of the RAD selection (NumPy's order) replaced by
    a parallel reduce-then-scan.  Probabilities agree to a few ulp, and on the reference fixture
    (tests/golden/sampling.npz residual field) as well as on log-normal fields of other sizes the selected index sets
    are IDENTICAL to the parity mode's (a draw would have to fall within one rounding of a cdf boundary to differ)."""
    import os

    from ai4xde_b200 import resample

    dev = torch.device("cuda:0")
    if n == 20000:
        g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sampling.npz"))
        res_h = g["res"].reshape(-1).astype(np.float32)
    else:
        rng = np.random.RandomState(n)
        res_h = (rng.randn(n) * np.exp(rng.randn(n))).astype(np.float32)
    res = torch.zeros((1, _lib.pad32(n)), device=dev)
    res[0, :n] = torch.as_tensor(res_h).to(dev)
    out = {}
    for fast in (False, True):
        _lib.resample_set_fast(fast)
        try:
            p = _lib.rad_probabilities(res, n, 1.0, 1.0)
            rs = np.random.RandomState(7)
            ids = resample.choice_noreplace_device(p, m, rand=rs.random_sample).cpu().numpy()
        finally:
            _lib.resample_set_fast(False)
        out[fast] = (p.cpu().numpy(), ids)
    assert np.allclose(out[True][0], out[False][0], rtol=1e-13, atol=0.0)
    assert np.array_equal(out[True][1], out[False][1])


@pytest.mark.parametrize("widths,act", [([2, 30, 30, 30, 30, 30, 30, 1], "tanh"), ([2, 10, 10, 2], "sigmoid"),
                                        ([1, 24, 24, 24, 1], "sin")])
def test_fnn_below_width_32_rides_the_narrow_kernels_padded(widths, act):
    """dde.nn.FNN with a uniform hidden width below 32 other than 20 / 32 (the reference's Diffusion_reaction net is
    2-[30]x6-1, cases/PDECases.py:1320) is evaluated on its zero-padded 20- / 32-wide image by the whole-net narrow
    kernels (the timing hooks name them); value, residual and the real parameters' gradients equal a float64 autograd
    evaluation of the unpadded net to 1e-5."""
    import ai4xde_b200.dde as dde

    dde.config.set_default_float("float32")
    dde.config.set_engine("fp32")
    try:
        torch.manual_seed(3)
        net = dde.nn.FNN(widths, act, "Glorot normal")
        d = widths[0]
        x = (torch.rand((4099, d)) * 2 - 1).to("cuda").requires_grad_()
        _lib.timing_report()
        _lib.timing_enable(True)
        # the streams this residual needs (what dde.Model's dry run would find): u, u_x0, u_x(d-1), u_x0x0
        spec = JetSpec((0,), ((0, 0),)) if d == 1 else JetSpec((0, d - 1), ((0, 0),))
        with net.jet_scope(spec=spec):
            y = net(x)
        r = dde.grad.hessian(y, x, component=0 if widths[-1] > 1 else None, i=0, j=0) + dde.grad.jacobian(y, x, i=0, j=d - 1) + y[:, :1] ** 2
        dde.grad.clear()
        net.zero_grad()
        (r ** 2).mean().backward()
        _lib.timing_enable(False)
        names = set(_lib.timing_report())
        assert {"narrow_fwd", "narrow_bwd"} <= names, names
        g = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).double().cpu()
    finally:
        _lib.timing_enable(False)
        dde.config.set_engine("auto")
    # float64 autograd on the plain module chain
    f = {"tanh": torch.tanh, "sin": torch.sin, "sigmoid": torch.sigmoid}[act]
    ps = [p.detach().double().cpu().requires_grad_() for p in net.parameters()]
    xd = x.detach().double().cpu().requires_grad_()
    h = xd
    for i in range(0, len(ps) - 2, 2):
        h = f(h @ ps[i].t() + ps[i + 1])
    yd = h @ ps[-2].t() + ps[-1]
    d1 = torch.autograd.grad(yd[:, 0].sum(), xd, create_graph=True)[0]
    d2 = torch.autograd.grad(d1[:, 0].sum(), xd, create_graph=True)[0]
    rd = d2[:, 0:1] + d1[:, d - 1:d] + yd[:, :1] ** 2
    gd = torch.cat([t.reshape(-1) for t in torch.autograd.grad((rd ** 2).mean(), ps)])
    assert _rel(y.detach().double().cpu(), yd.detach()) < 1e-5
    assert _rel(r.detach().double().cpu(), rd.detach()) < 1e-5
    assert _rel(g, gd) < 2e-5


_VARIANT_CHILD = r"""
import sys, torch
sys.path.insert(0, sys.argv[1])
from ai4xde_b200 import _lib
from ai4xde_b200.jet import JetSpec, to_soa
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(11)
# (a) FBPINN windowed bank forward: 8 nets 1-[64]x3-1, value / first / second derivative
K, n = 8, 5000
widths = [1, 64, 64, 64, 1]
js = JetSpec((0,), ((0, 0),))
net = _lib.make_net_desc(widths, "tanh", n_nets=K)
params = (torch.randn((K, _lib.param_count(_lib.make_net_desc(widths, "tanh"))), generator=g) * 0.2).to(dev)
xs = to_soa(torch.rand((n, 1), generator=g).to(dev))
edges = torch.linspace(0, 1, K + 1)
lst = torch.arange(K, dtype=torch.int32, device=dev)
bank = _lib.fbpinn_window_forward(net, js.desc, params, xs, n, edges[:-1].contiguous().to(dev), edges[1:].contiguous().to(dev),
                                  lst, engine="tf32")
# (b) FP32 engine gradient of a ragged-width net (edge tiles, input layer, two-unit output layer)
widths = [2, 100, 100, 100, 2]
js2 = JetSpec((0, 1), ((0, 0), (1, 1)))
net2 = _lib.make_net_desc(widths, "tanh")
p2 = (torch.randn((1, _lib.param_count(net2)), generator=g) * 0.3).to(dev)
m = 3001
x2 = to_soa((torch.rand((m, 2), generator=g) * 2 - 1).to(dev))
jets, stash = _lib.jet_forward(net2, js2.desc, p2, x2, m, keep_all=True, engine="fp32")
jb = torch.randn(jets.shape, generator=g).to(dev)
jb[..., m:] = 0
grad = _lib.jet_backward(net2, js2.desc, p2, x2, m, jb, stash, engine="fp32")
torch.save({"bank": bank[..., :n].cpu(), "grad": grad.cpu()}, sys.argv[2])
"""


def test_measurement_switches_do_not_change_a_bit(tmp_path):
    """The A/B switches of the round-2 kernels are read once per process: packed FP32 pairs
    vs scalar activation jets and one vs two block barriers per net in the FBPINN bank kernel (csrc/tc_bank.cu), the
    pipelined vs the single-buffer FP32 weight-gradient kernel (csrc/simt_kernels.cuh).  Every variant must return the
    default build's bits: same operations per element, same order of additions."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    for tag, env in (("default", {}), ("plain", {"AX_TC_BANK_PACKED": "0", "AX_TC_BANK_ONEBAR": "0", "AX_SIMT_WGRAD_PIPE": "0"})):
        path = str(tmp_path / f"{tag}.pt")
        r = subprocess.run([sys.executable, "-c", _VARIANT_CHILD, root, path], env=dict(os.environ, **env),
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[tag] = torch.load(path)
    assert torch.equal(outs["default"]["bank"], outs["plain"]["bank"])
    assert torch.equal(outs["default"]["grad"], outs["plain"]["grad"])
    assert float(outs["default"]["bank"].abs().sum()) > 0 and float(outs["default"]["grad"].abs().sum()) > 0
