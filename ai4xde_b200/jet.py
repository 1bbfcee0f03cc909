"""Jet streams on the host side: stream-set selection, the autograd bridge to the fused
kernels, and the "function with known derivatives" leaves that let torch.autograd differentiate
through arbitrary output transforms / residual algebra down to the kernel's Taylor-mode jets.

Replaces, for the network part, DeepXDE's Jacobian/Hessian autograd passes requested by every
`pde(x, y)` closure (e.g. /root/reference/AI4XDE/cases/PDECases.py:1087-1094) -- SURVEY.md
section 8a rows G1/G2/N1.
"""
import torch
from torch.autograd import Function

from . import _lib


class JetSpec:
    """Which derivative streams the kernels carry.

    dirs  : input coordinate of first-order stream j            (stream index 1 + j)
    pairs : (a, b) indices into dirs of second-order stream q   (stream index 1 + n1 + q)
    Always one of the compiled patterns (ai4xde_b200/csrc/common.cuh); requests are widened to
    the smallest compiled superset.
    """

    TRACE = 15  # pair code of the Laplacian stream (csrc/common.cuh kTracePair)
    THIRD, FOURTH = 14, 13  # pair codes of the one-direction third / fourth derivative streams (P_1_3, P_1_4)
    TRIPLE = 8              # pair code 8 + m: third derivative with m factors along dirs[1], 3 - m along dirs[0] (P_2_3)

    def __init__(self, dirs, pairs):
        self.dirs = tuple(int(d) for d in dirs)
        self.pairs = tuple((int(a), int(b)) for a, b in pairs)
        self.n1, self.n2 = len(self.dirs), len(self.pairs)
        self.S = 1 + self.n1 + self.n2
        self.first_of_dim = {d: j for j, d in enumerate(self.dirs)}
        self.second_of_dims = {}
        self.trace = None  # index q of the Laplacian stream, if this is a trace pattern
        self.third = self.fourth = None  # index q of the third / fourth derivative stream (one direction only)
        self.triples = {}                # m -> index q of the two-direction third derivative stream
        for q, (a, b) in enumerate(self.pairs):
            if a == b and self.TRIPLE <= a < self.TRIPLE + 4:
                if self.n1 != 2 or self.pairs[:3] != ((0, 0), (0, 1), (1, 1)):
                    raise ValueError("two-direction third order streams follow the (0,0), (0,1), (1,1) streams")
                self.triples[a - self.TRIPLE] = q
                continue
            if a == self.TRACE and b == self.TRACE:
                self.trace = q
                continue
            if a == b and a in (self.THIRD, self.FOURTH):
                if self.n1 != 1 or self.pairs[0] != (0, 0):
                    raise ValueError("third / fourth order streams follow the (0, 0) stream of a one-direction pattern")
                if a == self.THIRD:
                    self.third = q
                else:
                    self.fourth = q
                continue
            da, db = self.dirs[a], self.dirs[b]
            self.second_of_dims[(da, db)] = q
            self.second_of_dims[(db, da)] = q
        self.desc = _lib.make_jet_desc(self.dirs, self.pairs)

    @staticmethod
    def laplacian(dims):
        """{value, d/dx_j for j in dims, sum_j d2/dx_j^2}: S = 2 + len(dims) streams."""
        return JetSpec(sorted(dims), [(JetSpec.TRACE, JetSpec.TRACE)])

    def key(self):
        return (self.dirs, self.pairs)

    def __repr__(self):
        return f"JetSpec(dirs={self.dirs}, pairs={self.pairs})"

    @staticmethod
    def value_only():
        return JetSpec((), ())

    @staticmethod
    def one_direction(dim, order):
        """{value, d/dx, ..., d^order/dx^order} along input `dim` (order 2..4)."""
        pairs = [(0, 0), (JetSpec.THIRD, JetSpec.THIRD), (JetSpec.FOURTH, JetSpec.FOURTH)][: order - 1]
        return JetSpec([dim], pairs)

    @staticmethod
    def two_direction_third(dim_a, dim_b):
        """everything to third order along two inputs: {u, a, b, aa, ab, bb, aaa, aab, abb, bbb}"""
        t = JetSpec.TRIPLE
        return JetSpec([dim_a, dim_b], [(0, 0), (0, 1), (1, 1)] + [(t + m, t + m) for m in range(4)])

    def children_of_second(self, q):
        """What differentiating second-order stream q once more gives: {input dim: (index of that third-order
        stream in the list of streams after the value, children of that stream)} -- empty if this pattern stops
        at second order."""
        if q == 0 and self.third is not None:
            fourth = {} if self.fourth is None else {self.dirs[0]: (self.n1 + self.fourth, {})}
            return {self.dirs[0]: (self.n1 + self.third, fourth)}
        if self.triples and q < 3:
            ia, ib = self.pairs[q]
            return {self.dirs[k]: (self.n1 + self.triples[ia + ib + k], {}) for k in (0, 1)}
        return {}

    @staticmethod
    def full(d_in):
        if d_in > _lib.MAX_FIRST:
            raise NotImplementedError(
                f"second-order jets w.r.t. {d_in} inputs exceed the compiled patterns (max {_lib.MAX_FIRST})"
            )
        dirs = list(range(d_in))
        pairs = [(a, b) for a in range(d_in) for b in range(a, d_in)]
        return JetSpec(dirs, pairs)

    @staticmethod
    def covering(need1, need2, order=2):
        """Smallest compiled pattern covering first-order dims `need1` and second-order dim
        pairs `need2` (iterables of ints / of (j, k)); `order` 3 / 4: derivatives up to that order along the
        single direction all of them share."""
        need2 = {tuple(sorted(p)) for p in need2}
        if order > 2:
            dims = sorted(set(need1) | {d for p in need2 for d in p})
            if len(dims) == 1:
                return JetSpec.one_direction(dims[0], order)
            if len(dims) == 2 and order == 3:
                return JetSpec.two_direction_third(*dims)
            raise NotImplementedError(
                "third-order derivatives are compiled for one or two input directions, fourth-order for one")
        dims = sorted(set(need1) | {d for p in need2 for d in p})
        n1 = len(dims)
        if n1 > _lib.MAX_FIRST:
            raise NotImplementedError(
                f"derivatives w.r.t. {n1} different inputs exceed the compiled patterns (max {_lib.MAX_FIRST})"
            )
        if not need2:
            return JetSpec(dims, [])
        diag_only = all(j == k for j, k in need2)
        if diag_only and n1 == 1:
            return JetSpec(dims, [(0, 0)])
        if diag_only and n1 == 2 and len(need2) == 1:
            (j, _), = need2
            other = [d for d in dims if d != j]
            return JetSpec([j] + other, [(0, 0)])
        if diag_only:
            return JetSpec(dims, [(a, a) for a in range(n1)])
        return JetSpec(dims, [(a, b) for a in range(n1) for b in range(a, n1)])


def to_soa(x):
    """(n, d) -> zero-padded [d, ldn] float32 (the kernels' point layout)."""
    n, d = x.shape
    ldn = _lib.pad32(n)
    out = torch.zeros((d, ldn), dtype=torch.float32, device=x.device)
    out[:, :n] = x.detach().t()
    return out


# ---- functions with known derivatives -------------------------------------------------------
_HIGHER_MSG = ("third- and higher-order derivatives of the network are not in the fused jet kernels beyond: "
               "third order along one or two inputs, fourth order along one")


class _LeafHigher(Function):
    """A second- or higher-order stream as a function of x.  `children` = {input dim: (index into `tensors`,
    children of that stream)} names the carried streams that are its x-derivatives; a stream without children
    is the end of what the kernels carry, differentiating it raises."""

    @staticmethod
    def forward(ctx, x, children, h, *tensors):
        ctx.children = children
        ctx.save_for_backward(x, *tensors)
        return h.clone()

    @staticmethod
    def backward(ctx, g):
        x, *tensors = ctx.saved_tensors
        gx = None
        if ctx.needs_input_grad[0]:
            if not ctx.children:
                if torch.is_grad_enabled():
                    raise NotImplementedError(_HIGHER_MSG)
            else:
                gx = torch.zeros_like(x)
                cols = []
                for dim in range(x.shape[1]):
                    hit = ctx.children.get(dim)
                    if hit is None:
                        cols.append(torch.zeros_like(g[:, 0]))
                        continue
                    t = tensors[hit[0]]
                    if torch.is_grad_enabled():
                        t = _LeafHigher.apply(x, hit[1], t, *tensors)
                    cols.append((g * t).sum(dim=1))
                gx = torch.stack(cols, dim=1)
        return (gx, None, g) + (None,) * len(tensors)


def _compact(children, after_value, keep):
    """children with indices into `after_value` -> the same tree with indices into the list `keep` collects"""
    out = {}
    for dim, (idx, sub) in children.items():
        keep.append(after_value[idx])
        pos = len(keep) - 1
        out[dim] = (pos, _compact(sub, after_value, keep))
    return out


def _second_leaf(x, spec, q, after_value):
    """Leaf over second-order stream q; `after_value` = the streams after the value (firsts, then seconds)."""
    keep = []
    children = _compact(spec.children_of_second(q), after_value, keep)
    return _LeafHigher.apply(x, children, after_value[spec.n1 + q], *keep)


class _Leaf1(Function):
    """First-order stream j as a function of x: d/dx_k = second-order stream (j, k)."""

    @staticmethod
    def forward(ctx, x, spec, j, fj, *seconds):
        ctx.spec, ctx.j = spec, j
        ctx.save_for_backward(x, *seconds)
        return fj.clone()

    @staticmethod
    def backward(ctx, gd):
        x, *seconds = ctx.saved_tensors
        spec, j = ctx.spec, ctx.j
        gx = None
        if ctx.needs_input_grad[0]:
            cols = []
            dj = spec.dirs[j]
            for dim in range(x.shape[1]):
                q = spec.second_of_dims.get((dj, dim))
                if q is None:
                    cols.append(torch.zeros_like(gd[:, 0]))
                    continue
                h = seconds[q]
                if torch.is_grad_enabled():
                    h = _second_leaf(x, spec, q, (None,) * spec.n1 + tuple(seconds))
                cols.append((gd * h).sum(dim=1))
            gx = torch.stack(cols, dim=1)
        return (gx, None, None, gd) + (None,) * len(seconds)


class _Leaf0(Function):
    """Network output as a function of x whose x-derivatives are the kernel's jets."""

    @staticmethod
    def forward(ctx, x, spec, value, *streams):
        ctx.spec = spec
        ctx.save_for_backward(x, *streams)
        return value.clone()

    @staticmethod
    def backward(ctx, gy):
        x, *streams = ctx.saved_tensors
        spec = ctx.spec
        gx = None
        if ctx.needs_input_grad[0]:
            cols = []
            for dim in range(x.shape[1]):
                j = spec.first_of_dim.get(dim)
                if j is None:
                    cols.append(torch.zeros_like(gy[:, 0]))
                    continue
                fj = streams[j]
                if torch.is_grad_enabled():
                    fj = _Leaf1.apply(x, spec, j, fj, *streams[spec.n1:])
                cols.append((gy * fj).sum(dim=1))
            gx = torch.stack(cols, dim=1)
        return (gx, None, gy) + (None,) * len(streams)


class JetOutput:
    """Record kept per network evaluation so dde.grad.jacobian/hessian can read the streams
    directly when asked about the raw network output (no output transform in between)."""

    __slots__ = ("x", "spec", "streams", "_first", "_second")

    def __init__(self, x, spec, streams):
        self.x, self.spec, self.streams = x, spec, streams
        self._first, self._second = {}, {}

    def first(self, comp, dim):
        """d y_comp / d x_dim as a tensor that is itself differentiable w.r.t. x (through the
        second-order streams), so `jacobian(jacobian(y, x), x)` keeps working."""
        j = self.spec.first_of_dim.get(dim)
        if j is None:
            return None
        if j not in self._first:
            n1 = self.spec.n1
            self._first[j] = _Leaf1.apply(self.x, self.spec, j, self.streams[1 + j], *self.streams[1 + n1:])
        return self._first[j][:, comp : comp + 1]

    def first_all(self, comp):
        """(n, d_in) Jacobian row of component `comp`, or None if a column is not carried."""
        cols = []
        for dim in range(self.x.shape[1]):
            c = self.first(comp, dim)
            if c is None:
                return None
            cols.append(c)
        return torch.cat(cols, dim=1)

    def second(self, comp, dim_i, dim_j):
        if self.spec.trace is not None:
            # Laplacian pattern (adopted only after the dry run showed the caller uses nothing but the
            # sum of the pure second derivatives): the whole sum is handed out as the first one
            if dim_i != dim_j or dim_i not in self.spec.first_of_dim:
                return None
            qt = self.spec.trace
            if qt not in self._second:
                self._second[qt] = _LeafHigher.apply(self.x, {}, self.streams[1 + self.spec.n1 + qt])
            lap = self._second[qt][:, comp : comp + 1]
            return lap if dim_i == self.spec.dirs[0] else torch.zeros_like(lap)
        q = self.spec.second_of_dims.get((dim_i, dim_j))
        if q is None:
            return None
        if q not in self._second:
            self._second[q] = _second_leaf(self.x, self.spec, q, self.streams[1:])
        return self._second[q][:, comp : comp + 1]


def make_leaf(x, spec, streams):
    """streams: S tensors (n, C). Returns y (n, C) differentiable w.r.t. x via the streams."""
    return _Leaf0.apply(x, spec, streams[0], *streams[1:])


def probe_spec(d):
    """The widest pattern a probe hands out for d inputs: everything to second order; to third for two inputs,
    to fourth for one."""
    if d == 1:
        return JetSpec.one_direction(0, 4)
    return JetSpec.two_direction_third(0, 1) if d == 2 else JetSpec.full(d)


class _ProbeRecord:
    """Stands in for JetOutput during the discovery dry run: notes every derivative the caller asks of the RAW
    network output through dde.grad.jacobian / hessian and answers None, so the request is then served by
    autograd over the probe leaf as usual.  A stream that was asked for by name is carried even if the loss
    happened not to depend on it numerically during the dry run (`dde.Variable(0.0) * u_xx`, a curriculum weight
    that starts at zero): dropping it would make that derivative silently zero for the rest of the run."""

    def __init__(self, x, d):
        self.x, self.d = x, d
        self.firsts, self.seconds = set(), set()

    def first(self, comp, dim):
        self.firsts.add(int(dim))
        return None

    def first_all(self, comp):
        self.firsts.update(range(self.d))
        return None

    def second(self, comp, dim_i, dim_j):
        self.seconds.add(tuple(sorted((int(dim_i), int(dim_j)))))
        return None


def probe_streams(x, n_out, fn, allow_trace=True, feature=None, variables=()):
    """Find which derivative streams `fn(x, y)` really uses.

    With a feature transform (`net.apply_feature_transform(f)`) the kernels differentiate w.r.t. the
    FEATURES phi = f(x): the stand-in output is a leaf over phi, `f` itself stays ordinary autograd, and the
    streams found are those of d/dphi (the chain rule through `f` usually needs all of them).

    fn receives the input tensor and a stand-in network output whose value / first / second
    derivative streams are independent random tensors, and must return the tensors that the
    training loss (or the predict operator) is built from.  A stream is needed if the result
    has a non-zero gradient w.r.t. it, or if the caller asked dde.grad.jacobian / hessian for it by name on
    the raw network output (see _ProbeRecord).  `variables` are the external trainable variables
    (`dde.Variable`s of inverse problems): entries that are exactly zero are set to 0.5 for the duration of the
    dry run, so a term they multiply is not mistaken for an absent one.
    Returns JetSpec.covering(...).  No kernel is involved.
    """
    saved = []
    for v in variables or ():
        if torch.is_tensor(v) and bool((v.detach() == 0).any()):
            saved.append((v, v.detach().clone()))
            with torch.no_grad():
                v.copy_(torch.where(v == 0, torch.full_like(v, 0.5), v))
    try:
        return _probe_streams(x, n_out, fn, allow_trace, feature)
    finally:
        with torch.no_grad():
            for v, old in saved:
                v.copy_(old)


def _probe_streams(x, n_out, fn, allow_trace, feature):
    n = x.shape[0]
    with torch.enable_grad():
        leaf_in = x if feature is None else feature(x)
    d = leaf_in.shape[1]
    if d > _lib.MAX_FIRST:
        # cannot even express the full second-order set; probe first order only
        raise NotImplementedError(f"probe for d_in={d} > {_lib.MAX_FIRST} inputs")
    if feature is not None:
        allow_trace = False
    full = probe_spec(d)
    gen = torch.Generator(device=x.device)
    gen.manual_seed(1234)
    mk = lambda: torch.randn((n, n_out), dtype=x.dtype, device=x.device, generator=gen).requires_grad_()
    streams = [mk() for _ in range(full.S)]
    from .dde import gradients as grad_mod  # local import: jet <-> dde.gradients

    asked = _ProbeRecord(leaf_in, d)
    with torch.enable_grad():
        y = make_leaf(leaf_in, full, streams)
        grad_mod.register(y, asked)
        outs = fn(x, y)
        if not isinstance(outs, (list, tuple)):
            outs = [outs]
        total = None
        for o in outs:
            if o is None or not torch.is_tensor(o) or not o.requires_grad:
                continue
            t = (o.to(x.dtype) ** 2).sum()
            total = t if total is None else total + t
        if total is None:
            return JetSpec.value_only()
        grads = torch.autograd.grad(total, streams[1:], allow_unused=True, retain_graph=True)
    need1, need2 = sorted(asked.firsts), sorted(asked.seconds)
    for j in range(full.n1):
        g = grads[j]
        if g is not None and bool((g != 0).any()) and full.dirs[j] not in need1:
            need1.append(full.dirs[j])
    order = 2
    for q, (a, b) in enumerate(full.pairs):
        g = grads[full.n1 + q]
        if g is not None and bool((g != 0).any()):
            if q == full.third or q == full.fourth:
                order = max(order, 3 if q == full.third else 4)
            elif q in full.triples.values():
                order = max(order, 3)
            elif tuple(sorted((full.dirs[a], full.dirs[b]))) not in need2:
                need2.append(tuple(sorted((full.dirs[a], full.dirs[b]))))
    if order > 2:
        return JetSpec.covering(full.dirs if full.triples else (need1 or [full.dirs[0]]), need2, order=order)
    cover = JetSpec.covering(need1, need2)
    lap = _try_laplacian(x, fn, full, streams, cover, need2, total, allow_trace)
    return lap if lap is not None else cover


def _try_laplacian(x, fn, full, streams, cover, need2, total_full, allow_trace):
    """If the caller only ever uses the SUM of the pure second derivatives over all carried
    directions (Helmholtz, steady Navier-Stokes: u_xx + u_yy), one Laplacian stream replaces them.
    Adopted only when re-running the caller's own code on the Laplacian pattern -- fed the sum of the
    same random second-order streams -- reproduces the loss and its stream gradients exactly."""
    if not allow_trace or len(need2) < 2 or any(j != k for j, k in need2):
        return None
    dims = sorted(j for j, _ in need2)
    if dims != sorted(cover.dirs) or len(dims) > 3:
        return None
    from .dde import gradients as grad  # local import: jet <-> dde.gradients

    spec = JetSpec.laplacian(dims)
    n1f = full.n1
    firsts = [streams[1 + full.first_of_dim[d]] for d in spec.dirs]
    lap = sum(streams[1 + n1f + full.second_of_dims[(d, d)]] for d in dims)
    t_streams = [streams[0]] + firsts + [lap]
    with torch.enable_grad():
        y = make_leaf(x, spec, t_streams)
        grad.register(y, JetOutput(x, spec, t_streams))
        try:
            outs = fn(x, y)
        except Exception:
            return None
        outs = outs if isinstance(outs, (list, tuple)) else [outs]
        total = None
        for o in outs:
            if o is None or not torch.is_tensor(o) or not o.requires_grad:
                continue
            t = (o.to(x.dtype) ** 2).sum()
            total = t if total is None else total + t
        if total is None:
            return None
        leaves = [streams[0]] + firsts
        g_t = torch.autograd.grad(total, leaves, allow_unused=True)
        g_f = torch.autograd.grad(total_full, leaves, allow_unused=True, retain_graph=True)
    if not torch.allclose(total, total_full, rtol=1e-5, atol=0.0):
        return None
    for a, b in zip(g_t, g_f):
        if (a is None) != (b is None):
            return None
        if a is not None and not torch.allclose(a, b, rtol=1e-4, atol=1e-6 * float(b.abs().max())):
            return None
    return spec
