"""GPU parity of the tcgen05 (TF32) engine.  Bounds are looser than the FP32 engine's 1e-5 and are
STATED here: TF32 keeps 10 mantissa bits (operands rounded to nearest, FP32 accumulate), so with
Glorot-scale weights the jets and gradients agree with the float64 Taylor-mode double to a few
1e-3 relative (of the largest entry), growing with depth.  Measured on the B200 (round 2, printed by the tests with
`-s`): jets 0.5e-3 (one GEMM layer) .. 6.5e-3 (2-[64]x4-1, three GEMM layers) .. 4.8e-3 (five), gradients 0.5..3.6e-3.
The bounds are within 3x of those: jets 3e-3 per GEMM layer (at most 1e-2), gradients 1e-2."""
import numpy as np
import pytest
import torch

import _taylor_double as td
from ai4xde_b200 import _lib
from ai4xde_b200.jet import JetSpec, to_soa
from parity_util import product_vs_oracle

pytestmark = pytest.mark.gpu

TF32_JET_TOL = 1e-2
TF32_GRAD_TOL = 1e-2


def jet_tol(widths):
    """3e-3 per tensor-core GEMM layer (hidden layers - 1), at most TF32_JET_TOL"""
    return min(TF32_JET_TOL, 3e-3 * max(1, len(widths) - 3))



def _glorot_params(widths, gen, k=1):
    chunks = []
    for i, o in zip(widths[:-1], widths[1:]):
        chunks.append(torch.randn((k, o * i), generator=gen) * (2.0 / (i + o)) ** 0.5)
        chunks.append(torch.randn((k, o), generator=gen) * 0.1)
    return torch.cat(chunks, dim=1)


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("widths,act,spec,n", [
    ([2, 128, 128, 1], "tanh", ((), ()), 300),
    ([2, 128, 128, 1], "tanh", ((0, 1), ((0, 0), (1, 1))), 300),
    ([2, 256, 256, 256, 1], "tanh", ((0, 1), ((0, 0), (1, 1))), 1000),
    ([2, 64, 64, 64, 64, 1], "tanh", ((1, 0), ((0, 0),)), 10000),
    ([1, 64, 64, 64, 1], "tanh", ((0,), ((0, 0),)), 2048),
    ([2, 128, 128, 128, 128, 128, 3], "tanh", ((0, 1), ((0, 0), (1, 1))), 5001),
    ([2, 256] + [256] * 5 + [1], "tanh", ((0, 1), ((0, 0), (1, 1))), 20000),
    ([2, 128, 128, 128, 2], "sigmoid", ((0, 1), ((0, 0), (0, 1), (1, 1))), 777),
    ([2, 64, 64, 1], "tanh", ((0, 1), ()), 95),
    ([2, 256] + [256] * 5 + [1], "tanh", ((0, 1), ((15, 15),)), 20000),     # Laplacian stream (cfg4 as run)
    ([2, 128, 128, 128, 3], "sigmoid", ((0, 1), ((15, 15),)), 3000),
    # H = 256 backward = fused dgrad + wgrad kernel (k_tc_bwd2): every tile shape it is compiled for
    ([2, 256, 256, 256, 1], "tanh", ((), ()), 333),                                  # S = 1, PT = 128
    ([1, 256, 256, 256, 1], "tanh", ((0,), ((0, 0),)), 2053),                        # S = 3, PT = 32
    ([2, 256, 256, 256, 256, 2], "tanh", ((1, 0), ((0, 0),)), 4100),                 # S = 4, PT = 32
    ([2, 256, 256, 256, 1], "sigmoid", ((0, 1), ((0, 0), (0, 1), (1, 1))), 1501),    # S = 6, PT = 16
    ([3, 256, 256, 256, 1], "tanh", ((0, 1), ((0, 0), (1, 1))), 9999),               # d_in = 3: separate first-layer bwd
    # ragged / tiny point sets: less than one tile, one point, not a multiple of the 8-point groups
    ([2, 256, 256, 256, 1], "tanh", ((0, 1), ((15, 15),)), 1),
    ([2, 256, 256, 256, 1], "tanh", ((0, 1), ((15, 15),)), 37),
    ([2, 128, 128, 128, 3], "tanh", ((0, 1), ((15, 15),)), 1),
    ([2, 128, 128, 128, 3], "tanh", ((0, 1), ((0, 0), (1, 1))), 45),
    ([1, 64, 64, 64, 1], "sigmoid", ((0,), ((0, 0),)), 3),
    # sin (the reference's default Helmholtz / wave activation, cases/PDECases.py:513-515, WaveCases.py:19): the
    # adjoint reads the pre-activation streams from the second stash
    ([2, 256, 256, 256, 1], "sin", ((0, 1), ((15, 15),)), 5000),                     # Helmholtz default, padded image
    ([2, 128, 128, 128, 128, 2], "sin", ((0, 1), ((0, 0), (0, 1), (1, 1))), 3001),   # S = 6, single-CTA backward
    ([2, 64, 64, 64, 1], "sin", ((1, 0), ((0, 0),)), 2050),
    ([3, 256, 256, 256, 1], "sin", ((0, 1), ((0, 0), (1, 1))), 999),                 # separate first-layer backward
    ([2, 256, 256, 256, 1], "sin", ((0, 1), ((15, 15),)), 37),
    ([2, 128, 128, 3], "sin", ((), ()), 100),                                        # value only, one GEMM layer
    ([2, 256] + [256] * 4 + [1], "sin", ((0, 1), ((0, 0), (1, 1))), 4099),           # wave-net depth (2-[100]x5-1 padded)
])
def test_tf32_engine_matches_double(widths, act, spec, n):
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(11)
    js = JetSpec(*spec)
    net = _lib.make_net_desc(widths, act)
    assert _lib.engine_supported(net, js.desc, "tf32")
    params = _glorot_params(widths, gen).to(dev)
    x = (torch.rand((n, widths[0]), generator=gen) * 2 - 1).to(dev)
    xs = to_soa(x)
    jets, stash = _lib.jet_forward(net, js.desc, params, xs, n, keep_all=True, engine="tf32")
    ref, _ = td.jet_forward(net, js.desc, params.double(), xs.double(), n, True)
    jet_errs = [_rel(jets[0, s, :, :n].double(), ref[0, s, :, :n]) for s in range(js.S)]
    for s in range(js.S):
        assert jet_errs[s] < jet_tol(widths), f"stream {s}"
    jets2, _ = _lib.jet_forward(net, js.desc, params, xs, n, keep_all=False, engine="tf32")
    if widths[1] == 64:
        # forward-only evaluation of width-64 nets is the whole-net kernel (csrc/tc_bank.cu): another summation
        # order and a TF32-rounded value stream, so it is held to the double, not to the layer-by-layer kernels
        for s in range(js.S):
            assert _rel(jets2[0, s, :, :n].double(), ref[0, s, :, :n]) < TF32_JET_TOL, f"bank stream {s}"
    else:
        assert torch.equal(jets2[..., :n], jets[..., :n])
    jb = torch.randn(jets.shape, generator=gen).to(dev)
    jb[..., n:] = 0
    grad = _lib.jet_backward(net, js.desc, params, xs, n, jb, stash, engine="tf32")
    gref = td.jet_backward(net, js.desc, params.double(), xs.double(), n, jb.double(), None)
    off = 0
    grad_errs = []
    for i, o in zip(widths[:-1], widths[1:]):
        for cnt in (i * o, o):
            grad_errs.append(_rel(grad[0, off:off + cnt].double(), gref[0, off:off + cnt]))
            assert grad_errs[-1] < TF32_GRAD_TOL, f"param block at {off}"
            off += cnt
    print(f"MEASURED tf32 {widths} {act} S={js.S} n={n}: jets {max(jet_errs):.2e} grads {max(grad_errs):.2e}")
    grad2 = _lib.jet_backward(net, js.desc, params, xs, n, jb, stash, engine="tf32")
    assert torch.equal(grad, grad2)  # deterministic two-level reduction


@pytest.mark.parametrize("widths,act,spec,n,k", [
    ([1, 64, 64, 64, 1], "tanh", ((0,), ((0, 0),)), 5000, 64),          # config 5: 64 subdomain nets, S = 3
    ([1, 64, 64, 64, 1], "tanh", ((0,), ((0, 0),)), 128, 3),            # exactly one tile per net
    ([1, 64, 64, 64, 1], "tanh", ((0,), ((0, 0),)), 1, 2),              # one point
    ([2, 64, 64, 64, 64, 1], "tanh", ((1, 0), ((0, 0),)), 100000, 1),   # config 2 residual probe, S = 4
    ([2, 64, 64, 1], "tanh", ((), ()), 777, 5),                          # value only, one GEMM layer
    ([2, 64, 64, 64, 3], "sigmoid", ((0, 1), ((0, 0), (1, 1))), 1037, 2),            # S = 5, three outputs
    ([2, 64, 64, 64, 64, 64, 64, 2], "tanh", ((0, 1), ((0, 0), (0, 1), (1, 1))), 300, 2),   # S = 6, six hidden layers
    ([3, 64, 64, 64, 1], "tanh", ((0, 1), ((15, 15),)), 4099, 4),       # Laplacian stream, three inputs
])
def test_bank_forward_kernel_matches_double(widths, act, spec, n, k):
    """csrc/tc_bank.cu: K nets x all layers in one launch, activations on chip.  Against the float64 double,
    net by net; bit-identical when repeated; K nets in one launch == the same nets launched one at a time."""
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(5)
    js = JetSpec(*spec)
    net = _lib.make_net_desc(widths, act, n_nets=k)
    one = _lib.make_net_desc(widths, act, n_nets=1)
    assert _lib.engine_supported(net, js.desc, "tf32")
    assert _lib.stash_bytes(net, js.desc, _lib.pad32(n), False, "tf32") == 0
    params = _glorot_params(widths, gen, k).to(dev)
    x = (torch.rand((n, widths[0]), generator=gen) * 2 - 1).to(dev)
    xs = to_soa(x)
    jets, _ = _lib.jet_forward(net, js.desc, params, xs, n, keep_all=False, engine="tf32")
    again, _ = _lib.jet_forward(net, js.desc, params, xs, n, keep_all=False, engine="tf32")
    assert torch.equal(jets[..., :n], again[..., :n])
    ref, _ = td.jet_forward(net, js.desc, params.double(), xs.double(), n, False)
    for s in range(js.S):      # relative to the stream's largest entry over the bank (as for one net)
        assert _rel(jets[:, s, :, :n].double(), ref[:, s, :, :n]) < TF32_JET_TOL, f"stream {s}"
    for j in sorted({0, k - 1}):
        single, _ = _lib.jet_forward(one, js.desc, params[j:j + 1].contiguous(), xs, n, keep_all=False, engine="tf32")
        assert torch.equal(single[0, ..., :n], jets[j, ..., :n])


@pytest.mark.parametrize("k,n,skip,spec", [
    (64, 5000, (29,), ((0,), ((0, 0),))),        # config 5: all nets but the one being trained; net chunks over grid.y
    (64, 70001, (), ((0,), ((0, 0),))),          # more tile groups than SMs: one chunk, ragged last group
    (3, 1, (0, 2), ((0,), ((0, 0),))),           # one point, one net left
    (8, 1000, (), ((), ())),                     # value only (evaluation under no_grad)
    (5, 300, (0, 1, 2, 3, 4), ((0,), ((0, 0),))),  # empty list -> zeros
])
def test_fbpinn_window_sum_kernel_matches_double(k, n, skip, spec):
    """ax_fbpinn_window_forward: sum_k sigmoid(x - lo_k) sigmoid(hi_k - x) N_k(x) with its first and second
    x-derivative, nets summed inside the launch (/root/reference/AI4XDE/algorithm/FBPINN.py:35-50, 196-204), against
    float64: per-net streams from the double, window derivatives by autograd."""
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(3)
    js = JetSpec(*spec)
    widths = [1, 64, 64, 64, 1]
    net = _lib.make_net_desc(widths, "tanh", n_nets=k)
    assert _lib.fbpinn_window_supported(net, js.desc, "tf32")
    assert not _lib.fbpinn_window_supported(net, js.desc, "fp32")
    params = _glorot_params(widths, gen, k).to(dev)
    x = torch.rand((n, 1), generator=gen).to(dev)
    xs = to_soa(x)
    edges = torch.linspace(0, 1, k + 1)
    lo = (edges[:-1] - 0.1 / k).clamp_min(0).to(dev)
    hi = (edges[1:] + 0.1 / k).clamp_max(1).to(dev)
    ids = [j for j in range(k) if j not in skip]
    lst = torch.tensor(ids, dtype=torch.int32, device=dev)
    out = _lib.fbpinn_window_forward(net, js.desc, params, xs, n, lo, hi, lst, engine="tf32")
    again = _lib.fbpinn_window_forward(net, js.desc, params, xs, n, lo, hi, lst, engine="tf32")
    assert torch.equal(out[:, :n], again[:, :n])
    ref, _ = td.jet_forward(net, js.desc, params.double(), xs.double(), n, False)      # [k, S, 1, ldn]
    xd = x.double().requires_grad_()
    w = torch.sigmoid(xd - lo.double()[None, :]) * torch.sigmoid(hi.double()[None, :] - xd)    # (n, k)
    w1 = torch.stack([torch.autograd.grad(w[:, j].sum(), xd, create_graph=True)[0][:, 0] for j in range(k)], dim=1)
    w2 = torch.stack([torch.autograd.grad(w1[:, j].sum(), xd, retain_graph=True)[0][:, 0] for j in range(k)], dim=1)
    N = [ref[:, s, 0, :n].t() for s in range(js.S)]                                      # each (n, k)
    want = [(w * N[0])[:, ids].sum(dim=1)]
    if js.S == 3:
        want.append((w1 * N[0] + w * N[1])[:, ids].sum(dim=1))
        want.append((w2 * N[0] + 2 * w1 * N[1] + w * N[2])[:, ids].sum(dim=1))
    for s in range(js.S):
        if not ids:
            assert float(out[s, :n].abs().max()) == 0.0
        else:
            assert _rel(out[s, :n].double(), want[s].detach()) < TF32_JET_TOL, f"stream {s}"


def test_tf32_engine_refuses_what_it_cannot_do():
    net = _lib.make_net_desc([2, 150, 150, 1], "tanh")
    js = JetSpec((0, 1), ((0, 0), (1, 1)))
    assert not _lib.engine_supported(net, js.desc, "tf32")
    dev = torch.device("cuda:0")
    with pytest.raises(_lib.NativeLibraryError):
        _lib.jet_forward(net, js.desc, torch.zeros((1, _lib.param_count(net)), device=dev),
                         torch.zeros((2, 32), device=dev), 8, True, engine="tf32")
    fourier = _lib.make_net_desc([2, 128, 128, 1], "tanh", first_activation="sin")
    assert not _lib.engine_supported(fourier, js.desc, "tf32")   # Fourier first layers: served by the FP32 engine
    third = JetSpec.two_direction_third(0, 1)
    assert not _lib.engine_supported(_lib.make_net_desc([2, 128, 128, 1], "sin"), third.desc, "tf32")


@pytest.mark.parametrize("H", [64, 128, 256])
def test_tc_wgrad_kernel(H):
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(0)
    for rows in (1, 31, 33, 1000, 50001):
        z = torch.randn((rows, H), generator=gen).to(dev)
        a = torch.randn((rows, H), generator=gen).to(dev)
        out = _lib.debug_tc_wgrad(z, a)
        ref = z.double().t() @ a.double()
        assert _rel(out.double(), ref) < 3e-3


# Product-level parity on the BASELINE configs whose nets the tensor-core engine serves.
# Stated TF32 bounds (relative to the largest reference entry), each within 3x of what the B200 measures (round 2:
# cfg2 loss 4e-5 / grad 5.1e-3 / res 3e-4; cfg3 9e-5 / 5e-4 / 7.3e-3; cfg4 1e-4 / 1.6e-3 / 3.6e-4; cfg5 1e-7 / 2.9e-3 / 1e-5).
TF32_CFG_TOL = {
    "cfg2": dict(loss=2e-4, grad=1.5e-2, res=1e-3),
    "cfg3": dict(loss=3e-4, grad=1.5e-3, res=2e-2),
    "cfg4": dict(loss=3e-4, grad=5e-3, res=1.1e-3),
    "cfg5": dict(loss=1e-5, grad=9e-3, res=4e-5),
}


@pytest.mark.parametrize("cfg,n", [("cfg2", 10000), ("cfg3", 4096), ("cfg4", 4096), ("cfg5", 4096)])
def test_tf32_engine_matches_oracle(cfg, n):
    r = product_vs_oracle(cfg, n, engine="tf32")
    print(cfg, {k: v for k, v in r.items() if k.endswith("_rel")},
          "elementwise (median, p99):", {k: (f"{v[0]:.1e}", f"{v[1]:.1e}") for k, v in r["elementwise"].items()})
    tol = TF32_CFG_TOL[cfg]
    # elementwise, over the entries above 1e-3 of the largest: TF32 medians 7e-5 .. 4.7e-3, 99th percentiles up to 5.9e-2
    for what, (med, p99) in r["elementwise"].items():
        assert med < 1.5e-2 and p99 < 0.18, (cfg, what, med, p99)
    assert r["loss_rel"] < tol["loss"], r
    assert r["grad_rel"] < tol["grad"], r
    assert r["res_rel"] < tol["res"], r


def test_reference_default_helmholtz_sin_net_on_the_tensor_cores():
    """The reference's own Helmholtz case (cases/PDECases.py:513-515: 2-[150]x3-1, sin) under engine `auto`: the net
    rides the tcgen05 kernels on its zero-padded 256-wide image (post-activation streams as operands, pre-activation
    streams in the second stash for the adjoint).  Against the float64 oracle on the same points and weights; the
    FP32 engine on the same case holds 1e-5.  TF32 bounds stated, to be within 3x of the printed measurement."""
    from ai4xde_b200 import _lib as lib

    before = lib.launch_count()
    r = product_vs_oracle("cfg4_sin", 8192, engine="auto")
    print("cfg4_sin auto", {k: f"{v:.2e}" for k, v in r.items() if k.endswith("_rel")})
    net = r["case"].net
    assert net._tc_pad is not None and net._tc_pad.auto_ok and lib.launch_count() > before
    assert r["loss_rel"] < 3e-4, r
    assert r["grad_rel"] < 5e-3, r
    assert r["res_rel"] < 2e-3, r
    f = product_vs_oracle("cfg4_sin", 8192, engine="fp32")
    print("cfg4_sin fp32", {k: f"{v:.2e}" for k, v in f.items() if k.endswith("_rel")})
    assert f["loss_rel"] < 1e-5 and f["grad_rel"] < 1e-5 and f["res_rel"] < 1e-5, f


L2CHAIN_DEFAULT_GROUP = 3   # csrc/tc_fwd.cu l2chain_group()


@pytest.mark.parametrize("widths,act,spec,n,keep", [
    ([2, 256] + [256] * 5 + [1], "tanh", ((0, 1), ((15, 15),)), 200003, True),      # cfg4 as trained: 5 chained layers
    ([2, 256] + [256] * 5 + [1], "tanh", ((0, 1), ((15, 15),)), 200003, False),     # residual probe: ping-pong slots
    ([2, 256, 256, 256, 1], "sin", ((0, 1), ((15, 15),)), 70001, True),             # second stash written on the way
    ([3, 256, 256, 256, 256, 2], "sigmoid", ((0, 1), ((0, 0), (1, 1))), 30011, True),   # S = 5: 48-point tiles, no fused output
    ([2, 256, 256, 256, 1], "tanh", ((), ()), 9999, True),                          # S = 1: 256-point tiles
    ([2, 256, 256, 256, 256, 1], "tanh", ((0, 1), ((15, 15),)), 100, True),         # fewer tiles than clusters
])
def test_l2_chained_forward_is_bit_identical_to_the_layer_launches(widths, act, spec, n, keep):
    """csrc/tc_fwd.cu k_tc_fwd_l2: all hidden GEMM layers in one launch, a tile's streams handed from layer to layer
    through L2 (epilogue -> release-arrive -> TMA producer).  Same MMAs, same epilogue arithmetic: jets AND every
    stash layer must equal the one-launch-per-layer path bit for bit, for every group size, five times in a row (a
    lost hand-off would read rows that are not written yet)."""
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(3)
    js = JetSpec(*spec)
    net = _lib.make_net_desc(widths, act)
    params = _glorot_params(widths, gen).to(dev)
    xs = to_soa((torch.rand((n, widths[0]), generator=gen) * 2 - 1).to(dev))
    try:
        _lib.tc_set_option("l2chain", 0)
        jets0, stash0 = _lib.jet_forward(net, js.desc, params, xs, n, keep_all=keep, engine="tf32")
        H, S, ldn = widths[1], js.S, xs.shape[1]
        rows = n * S * H                                  # written part of a stash layer: [(p, s)][unit], p < n
        layer = ldn * S * H
        L = len(widths) - 2
        n_layers = L * (2 if act == "sin" else 1) if keep else 0
        s0 = stash0.view(torch.float32)
        for wres, defer, G in [(w, d, G) for w in (1, 0) for d in (2, 1, 0) for G in (1, 2, 3, 5, 8)]:
            _lib.tc_set_option("l2chain", 1)
            _lib.tc_set_option("l2chain_wres", wres)     # weights resident per (group, layer) visit / streamed per item
            _lib.tc_set_option("l2chain_group", G)
            _lib.tc_set_option("l2chain_defer", defer)   # tile released by the release warp / during the next item / at once
            for rep in range(5 if G in (2, 3) and defer == 2 and wres else 2):
                jets1, stash1 = _lib.jet_forward(net, js.desc, params, xs, n, keep_all=keep, engine="tf32")
                assert torch.equal(jets1[..., :n], jets0[..., :n]), (wres, defer, G, rep)
                s1 = stash1.view(torch.float32)
                for l in range(n_layers):
                    assert torch.equal(s1[l * layer : l * layer + rows], s0[l * layer : l * layer + rows]), (wres, defer, G, rep, l)
    finally:
        _lib.tc_set_option("l2chain", 1)
        _lib.tc_set_option("l2chain_group", L2CHAIN_DEFAULT_GROUP)
        _lib.tc_set_option("l2chain_defer", 2)
        _lib.tc_set_option("l2chain_wres", 1)


def test_cfg4_full_size_properties():
    """BASELINE cfg4 at its full size (2^20 + 80 rows, 2-[256]x6-1 tanh, Laplacian stream pattern) through
    size-independent properties: (i) every point's jets do not depend on which tile / launch it rode in
    (full batch == concatenation of two unequal shards, bit for bit); (ii) the parameter gradient is additive
    over point shards (the data-parallel contract); (iii) the backward is linear in the jet adjoints (x2 is
    exact in binary floating point, TF32 rounding included); (iv) a random 2048-point subset agrees with the
    float64 double within the stated TF32 bound."""
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(3)
    widths = [2] + [256] * 6 + [1]
    js = JetSpec((0, 1), ((15, 15),))
    net = _lib.make_net_desc(widths, "tanh")
    params = _glorot_params(widths, gen).to(dev)
    n = 2**20 + 80
    x = (torch.rand((n, 2), generator=gen) * 2 - 1).to(dev)
    xs = to_soa(x)
    jets, stash = _lib.jet_forward(net, js.desc, params, xs, n, keep_all=True, engine="tf32")
    cut = 400_003
    ja, sa = _lib.jet_forward(net, js.desc, params, to_soa(x[:cut]), cut, keep_all=True, engine="tf32")
    jb_, sb = _lib.jet_forward(net, js.desc, params, to_soa(x[cut:]), n - cut, keep_all=True, engine="tf32")
    assert torch.equal(jets[..., :cut], ja[..., :cut])
    assert torch.equal(jets[..., cut:n], jb_[..., :n - cut])
    # (iv) subset against the float64 double
    idx = torch.randperm(n, generator=gen)[:2048].to(dev)
    ref, _ = td.jet_forward(net, js.desc, params.double(), to_soa(x[idx]).double(), 2048, True)
    for s in range(js.S):
        assert _rel(jets[0, s, :, idx].double(), ref[0, s, :, :2048]) < TF32_JET_TOL, f"stream {s}"
    # (ii) additivity of the gradient over shards, (iii) linearity in the adjoints
    jbar = torch.randn(jets.shape, generator=gen).to(dev)
    jbar[..., n:] = 0
    g_full = _lib.jet_backward(net, js.desc, params, xs, n, jbar, stash, engine="tf32")
    bar_a = torch.zeros_like(ja)
    bar_a[..., :cut] = jbar[..., :cut]
    bar_b = torch.zeros_like(jb_)
    bar_b[..., :n - cut] = jbar[..., cut:n]
    g_a = _lib.jet_backward(net, js.desc, params, to_soa(x[:cut]), cut, bar_a, sa, engine="tf32")
    g_b = _lib.jet_backward(net, js.desc, params, to_soa(x[cut:]), n - cut, bar_b, sb, engine="tf32")
    assert _rel((g_a + g_b).double(), g_full.double()) < 1e-4     # FP32 partial sums in a different order
    g2 = _lib.jet_backward(net, js.desc, params, xs, n, 2.0 * jbar, stash, engine="tf32")
    assert torch.equal(g2, 2.0 * g_full)


@pytest.mark.parametrize("flags", [{"AX_TC_CHAIN": "1"}, {"AX_TC_FUSED_BWD": "0", "AX_TC_FUSE_OUT": "0"},
                                   {"AX_TC_PAIR": "0"}])
def test_alternative_kernel_paths_stay_parity_green(flags):
    """The library reads its kernel-selection switches once per process, so the non-default paths -- the
    whole-stack forward kernel (AX_TC_CHAIN=1), separate dgrad / wgrad / output-layer kernels, single-CTA
    layer kernels for H = 256 -- are exercised in a child process over the same parity cases."""
    import os
    import subprocess
    import sys

    env = dict(os.environ, **flags)
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_gpu_tc.py"), "-x", "-q",
                        "-k", "test_tf32_engine_matches_double"], env=env, capture_output=True, text=True, cwd=here)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_fnn_between_compiled_widths_rides_the_tensor_cores_padded():
    """dde.nn.FNN([2, 100, 100, 100, 100, 2], tanh) (the reference's Schrodinger net, cases/PDECases.py:1833) under
    engine 'auto': evaluated on its zero-padded 128-wide image by the tcgen05 kernels (timing hooks name them), and
    value / residual / real-parameter gradients agree with the same net on the FP32 engine within the TF32 bound."""
    import ai4xde_b200.dde as dde

    dde.config.set_default_float("float32")
    torch.manual_seed(5)
    net = dde.nn.FNN([2] + [100] * 4 + [2], "tanh", "Glorot normal")
    x = (torch.rand((20011, 2)) * 2 - 1).to("cuda").requires_grad_()
    out = {}
    try:
        for eng in ("fp32", "auto"):
            dde.config.set_engine(eng)
            _lib.timing_report()
            _lib.timing_enable(True)
            y = net(x)
            r = dde.grad.hessian(y, x, component=0, i=0, j=0) + dde.grad.jacobian(y, x, i=1, j=1) + y[:, :1] * y[:, 1:]
            dde.grad.clear()
            net.zero_grad()
            (r ** 2).mean().backward()
            _lib.timing_enable(False)
            names = set(_lib.timing_report())
            assert any(k.startswith("tc_") for k in names) == (eng == "auto"), names
            out[eng] = (y.detach().double(), r.detach().double(), torch.cat([p.grad.reshape(-1) for p in net.parameters()]).double())
    finally:
        _lib.timing_enable(False)
        dde.config.set_engine("auto")
    (y0, r0, g0), (y1, r1, g1) = out["fp32"], out["auto"]
    assert _rel(y1, y0) < TF32_JET_TOL
    assert _rel(r1, r0) < 1e-2
    assert _rel(g1, g0) < 2e-2
