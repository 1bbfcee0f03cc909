"""dde.nn.FNN of the fused path (SURVEY.md 8a row N1, 8b `net.forward` seam).

Same constructor, parameter names (`linears.{i}.weight/bias`, torch.nn.Linear layout), Glorot
initialisers, `apply_feature_transform` / `apply_output_transform`, `.train()/.eval()`,
`.auxiliary_vars`, `.regularizer` as DeepXDE's PyTorch FNN that
/root/reference/AI4XDE/cases/PDECases.py:38-42 constructs -- but `forward` runs ONE fused
Taylor-mode kernel sequence (ai4xde_b200/csrc) that returns the value and every derivative stream
the case asks for, instead of a Linear/tanh module chain to be differentiated by autograd.
All master weights live in one flat float32 device buffer; the nn.Linear parameters are views.
"""
import contextlib

import torch

from .. import _lib, jet
from . import config, gradients as grad

_INIT = {
    "Glorot normal": torch.nn.init.xavier_normal_,
    "Glorot uniform": torch.nn.init.xavier_uniform_,
    "He normal": torch.nn.init.kaiming_normal_,
    "He uniform": torch.nn.init.kaiming_uniform_,
    "zeros": torch.nn.init.zeros_,
}


class _JetScope:
    __slots__ = ("spec", "x_soa", "param_grad", "probe_streams")

    def __init__(self, spec=None, x_soa=None, param_grad=True, probe_streams=None):
        self.spec, self.x_soa, self.param_grad, self.probe_streams = spec, x_soa, param_grad, probe_streams


def _launch_forward(net, x_soa, n, spec, engine, need_backward):
    """Engine / padded-image choice and the forward kernel sequence of one FNN evaluation.
    -> (jets [1, S, C, ldn], stash, desc, pad, flat parameter image the kernels ran on)"""
    flat = net.flat_params()
    pad = net._tc_pad
    if pad is not None and not (engine == "tf32" or (engine == "auto" and pad.auto_ok)):
        pad = None
    if pad is not None and not _lib.engine_supported(pad.desc, spec.desc, "tf32"):
        pad = None                       # stream pattern outside the tensor-core engine: nothing to gain
    if pad is None and engine in ("fp32", "auto") and net._narrow_pad is not None \
            and _lib.narrow_supported(net._narrow_pad.desc, spec.desc):
        pad = net._narrow_pad
    if pad is not None:
        flat = pad.widen(flat)
    desc = net.net_desc if pad is None else pad.desc
    jets, stash = _lib.jet_forward(desc, spec.desc, flat, x_soa, n, keep_all=need_backward, engine=engine)
    return jets, stash, desc, pad, flat


def _launch_backward(ctx, jets_bar):
    """Backward kernel sequence -> the per-parameter gradient views (order of ctx.shapes)."""
    g_flat = _lib.jet_backward(ctx.desc, ctx.spec.desc, ctx.flat, ctx.x_soa, ctx.n, jets_bar, ctx.stash,
                               engine=ctx.engine)[0]
    if ctx.pad is not None:
        g_flat = ctx.pad.narrow(g_flat)
    ctx.stash = None
    outs, off = [], 0
    for shp in ctx.shapes:
        cnt = 1
        for d in shp:
            cnt *= d
        outs.append(g_flat[off : off + cnt].view(shp))
        off += cnt
    return tuple(outs)


class _FusedForward(torch.autograd.Function):
    """Routes the flat-buffer gradient of jet.JetMLP back to the per-layer nn.Parameters."""

    @staticmethod
    def forward(ctx, net, x_soa, n, spec, engine, need_backward, *params):
        jets, stash, desc, pad, flat = _launch_forward(net, x_soa, n, spec, engine, need_backward)
        ctx.net, ctx.spec, ctx.engine, ctx.n, ctx.desc, ctx.pad = net, spec, engine, n, desc, pad
        ctx.x_soa, ctx.stash, ctx.kept, ctx.flat = x_soa, stash, bool(need_backward), flat
        ctx.shapes = [p.shape for p in params]
        return tuple(jets[0, s, :, :n].t() for s in range(spec.S))

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *gouts):
        if not ctx.kept:
            raise RuntimeError("fused forward ran without a stash; parameter gradients unavailable")
        net, spec, n = ctx.net, ctx.spec, ctx.n
        ldn = ctx.x_soa.shape[1]
        c_out = net.layer_sizes[-1]
        jets_bar = torch.zeros((1, spec.S, c_out, ldn), dtype=torch.float32, device=ctx.flat.device)
        for s, g in enumerate(gouts):
            if g is not None:
                jets_bar[0, s, :, :n] = g.t()
        return (None, None, None, None, None, None) + _launch_backward(ctx, jets_bar)


class _FusedLosses(torch.autograd.Function):
    """Network AND loss side of a train step without a tensor-level graph in between: jet kernels -> residual
    program (per-term losses) ; backward: residual program in reverse (jets_bar) -> backward kernels.
    `tables`: _lib.ResidualTables of the traced loss code (ai4xde_b200/residual.py)."""

    @staticmethod
    def forward(ctx, net, tables, x_soa, n, spec, engine, need_backward, n_params, *tensors):
        """tensors = the net's parameters, then the trainable variables the program reads (inverse problems).
        x_soa: the batch's points, or (points, features) when the net has a feature transform -- the kernels
        evaluate the net on the features, the program reads the raw points."""
        params, variables = tensors[:n_params], tensors[n_params:]
        x_soa, net_soa = x_soa if isinstance(x_soa, tuple) else (x_soa, x_soa)
        jets, stash, desc, pad, flat = _launch_forward(net, net_soa, n, spec, engine, need_backward)
        ctx.spec, ctx.engine, ctx.n, ctx.desc, ctx.pad = spec, engine, n, desc, pad
        ctx.x_soa, ctx.stash, ctx.kept, ctx.flat = net_soa, stash, bool(need_backward), flat
        ctx.raw_soa = x_soa
        ctx.shapes = [p.shape for p in params]
        ctx.var_shapes = [v.shape for v in variables]
        ctx.tables, ctx.jets = tables, (jets if need_backward else None)
        return _lib.residual_forward(tables, x_soa, jets, n)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_losses):
        if not ctx.kept:
            raise RuntimeError("fused forward ran without a stash; parameter gradients unavailable")
        jets_bar, sgrad = _lib.residual_backward(ctx.tables, ctx.raw_soa, ctx.jets, ctx.n, g_losses)
        ctx.jets = None
        var_grads = tuple(sgrad[k].reshape(shp) for k, shp in zip(ctx.tables.scalar_grad, ctx.var_shapes))
        return (None, None, None, None, None, None, None, None) + _launch_backward(ctx, jets_bar) + var_grads


class _TensorCorePad:
    """Zero-padded image of an FNN whose uniform hidden width H lies strictly between the widths the tcgen05
    engine is compiled for (64 / 128 / 256): the same function on the next compiled width.  Padded units see
    zero incoming weights and bias (pre-activation 0, derivative streams 0; sin(0) = 0) and feed zero outgoing weights, so
    value, jets and the gradient of every real parameter are unchanged; the gradients of padded entries are
    dropped.  One scatter per evaluation builds the image from the flat master buffer, one gather brings the
    gradient back.  Used under engine 'auto' for widths above 64 and under an explicit 'tf32' also for widths
    33..63 (padded to 64); 'fp32' keeps the net as it is on the FP32 engine."""

    def __init__(self, layer_sizes, activation, h_pad):
        d_in, d_out, n_hidden = layer_sizes[0], layer_sizes[-1], len(layer_sizes) - 2
        padded = [d_in] + [h_pad] * n_hidden + [d_out]
        self.desc = _lib.make_net_desc(padded, activation)
        idx, off = [], 0
        for l in range(len(layer_sizes) - 1):
            o, i = layer_sizes[l + 1], layer_sizes[l]
            po, pi = padded[l + 1], padded[l]
            rows = torch.arange(o).unsqueeze(1) * pi + torch.arange(i).unsqueeze(0)
            idx.append((off + rows).reshape(-1))
            off += po * pi
            idx.append(off + torch.arange(o))
            off += po
        self.count = off
        self.index = torch.cat(idx)
        self._image = None

    @staticmethod
    def plan(layer_sizes, activation):
        hidden = layer_sizes[1:-1]
        if activation not in ("tanh", "sin", "sigmoid") or len(hidden) < 2 or len(set(hidden)) != 1:
            return None
        if layer_sizes[0] > 4 or layer_sizes[-1] > 4:
            return None
        h = hidden[0]
        if h <= 32 or h >= 256 or h in (64, 128):   # 64 / 128 / 256 need no padding
            return None
        pad = _TensorCorePad(layer_sizes, activation, 64 if h < 64 else 128 if h < 128 else 256)
        pad.auto_ok = h > 64    # narrower nets stay on FP32 FMA unless the caller asks for engine 'tf32'
        return pad

    @staticmethod
    def plan_narrow(layer_sizes, activation):
        """Uniform hidden widths below 32 other than 20 / 32 (the reference has 2-[30]x6-1, cases/PDECases.py:1320):
        the zero-padded image on 20 or 32 units runs on the whole-net narrow FP32 kernels (csrc/narrow.cuh) instead
        of layer by layer.  Same function, same real-parameter gradients (a padded unit's activation -- 0, or 0.5
        for sigmoid -- only ever meets zero outgoing weights)."""
        hidden = layer_sizes[1:-1]
        if activation not in ("tanh", "sin", "sigmoid") or len(hidden) < 1 or len(set(hidden)) != 1:
            return None
        if layer_sizes[0] > 4 or layer_sizes[-1] > 4:
            return None
        h = hidden[0]
        if h >= 32 or h == 20 or h < 4:
            return None
        return _TensorCorePad(layer_sizes, activation, 20 if h < 20 else 32)

    def widen(self, flat):
        if self._image is None or self._image.device != flat.device:
            self._image = torch.zeros((1, self.count), dtype=torch.float32, device=flat.device)
            self.index = self.index.to(flat.device)
        self._image[0, self.index] = flat.detach().reshape(-1)
        return self._image

    def narrow(self, g_padded):
        return g_padded.reshape(-1)[self.index]


class NN(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.regularizer = None
        self._auxiliary_vars = None
        self._input_transform = None
        self._output_transform = None

    @property
    def auxiliary_vars(self):
        return self._auxiliary_vars

    @auxiliary_vars.setter
    def auxiliary_vars(self, value):
        self._auxiliary_vars = value

    def apply_feature_transform(self, transform):
        self._input_transform = transform

    def apply_output_transform(self, transform):
        self._output_transform = transform

    def num_trainable_parameters(self):
        return sum(v.numel() for v in self.parameters() if v.requires_grad)


class FNN(NN):
    def __init__(self, layer_sizes, activation, kernel_initializer):
        super().__init__()
        if isinstance(activation, (list, tuple)):
            if len(set(activation)) != 1:
                raise NotImplementedError("per-layer activations are not in the fused path")
            activation = activation[0]
        if not isinstance(activation, str):
            raise NotImplementedError("callable activations are not in the fused path (tanh, sin, sigmoid are)")
        self.layer_sizes = [int(w) for w in layer_sizes]
        self.activation_name = activation
        self.net_desc = _lib.make_net_desc(self.layer_sizes, activation)
        self._tc_pad = _TensorCorePad.plan(self.layer_sizes, activation)
        self._narrow_pad = _TensorCorePad.plan_narrow(self.layer_sizes, activation)
        init = _INIT[kernel_initializer]
        self.linears = torch.nn.ModuleList()
        for i in range(1, len(self.layer_sizes)):
            lin = torch.nn.Linear(self.layer_sizes[i - 1], self.layer_sizes[i], dtype=torch.float32)
            init(lin.weight)
            torch.nn.init.zeros_(lin.bias)
            self.linears.append(lin)
        self._flat = None
        self._scope = _JetScope()
        self._pack(config.device())

    # ---- flat master buffer ----------------------------------------------------------------
    def _param_list(self):
        out = []
        for lin in self.linears:
            out += [lin.weight, lin.bias]
        return out

    def _pack(self, device):
        ps = self._param_list()
        flat = torch.empty(sum(p.numel() for p in ps), dtype=torch.float32, device=device)
        off = 0
        for p in ps:
            n = p.numel()
            flat[off : off + n] = p.detach().reshape(-1).to(device=device, dtype=torch.float32)
            p.data = flat[off : off + n].view(p.shape)
            off += n
        self._flat = flat

    def adopt_storage(self, row):
        """Move the master weights into `row` (a [P] float32 slice of a shared [K, P] bank) so that
        K structurally identical nets can be evaluated by one fused launch (NetBank)."""
        off = 0
        for p in self._param_list():
            n = p.numel()
            row[off : off + n] = p.detach().reshape(-1).to(device=row.device, dtype=torch.float32)
            p.data = row[off : off + n].view(p.shape)
            off += n
        self._flat = row

    def flat_params(self):
        """[1, P] view of the master buffer (re-packed if someone moved the parameters)."""
        ps = self._param_list()
        base, off, ok = self._flat.data_ptr(), 0, True
        for p in ps:
            if p.data_ptr() != base + 4 * off or p.dtype != torch.float32 or p.device != self._flat.device:
                ok = False
                break
            off += p.numel()
        if not ok:
            self._pack(ps[0].device if ps[0].device.type == "cuda" else config.device())
        return self._flat.view(1, -1)

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._flat = torch.empty(0)
        self._pack(self._param_list()[0].device)
        return out

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self.flat_params()
        return out

    # ---- evaluation ------------------------------------------------------------------------
    @contextlib.contextmanager
    def jet_scope(self, spec=None, x_soa=None, param_grad=True, probe_streams=None):
        prev = self._scope
        self._scope = _JetScope(spec, x_soa, param_grad, probe_streams)
        try:
            yield
        finally:
            self._scope = prev

    def forward(self, inputs):
        sc = self._scope
        x = inputs
        featured = self._input_transform is not None
        if featured:
            # net.apply_feature_transform(f) (cases/PDECases.py:212,405): the kernels differentiate w.r.t. the
            # features phi = f(x); f itself is ordinary autograd, so dde.grad.* on the raw inputs chains through it
            x = self._input_transform(inputs)
            if x.shape[1] != self.layer_sizes[0]:
                raise ValueError(f"feature transform returns {x.shape[1]} features, the net takes {self.layer_sizes[0]}")
        if sc.probe_streams is not None:
            y = jet.make_leaf(x, jet.probe_spec(x.shape[1]), sc.probe_streams)
        else:
            wants_jets = x.requires_grad and torch.is_grad_enabled()
            if not wants_jets:
                spec = jet.JetSpec.value_only()
            elif sc.spec is not None:
                spec = sc.spec
            else:
                spec = jet.JetSpec.full(x.shape[1])
            if featured or sc.x_soa is None or sc.x_soa.shape[1] < x.shape[0]:
                x_soa = jet.to_soa(x.detach().to(torch.float32))
            else:
                x_soa = sc.x_soa
            params = self._param_list()
            need_bwd = bool(sc.param_grad and torch.is_grad_enabled() and any(p.requires_grad for p in params))
            streams = _FusedForward.apply(self, x_soa, x.shape[0], spec, config.engine(), need_bwd, *params)
            if wants_jets:
                y = jet.make_leaf(x, spec, streams)
                grad.register(y, jet.JetOutput(x, spec, streams))
            else:
                # a fresh tensor, not a view of the kernels' jet buffer: callers update it in place
                # (algorithm/FBPINN.py:203 `result += y`)
                y = streams[0].clone()
        if self._output_transform is not None:
            y = self._output_transform(inputs, y)
        return y


class _FusedBankForward(torch.autograd.Function):
    """K nets, one launch.  Streams come back as (n, K*C).  FBPINN trains one subdomain net at a time
    (/root/reference/AI4XDE/algorithm/FBPINN.py:112-119), so the bank is evaluated forward-only -- on the
    tensor-core engine a single launch in which no activation leaves the SM (csrc/tc_bank.cu) -- and only the
    `active` nets are evaluated once more with their layers kept for the reverse pass: the stash holds
    len(active) nets instead of K."""

    @staticmethod
    def forward(ctx, bank, x_soa, n, spec, engine, need_backward, active, *params):
        jets = bank.evaluate(spec, x_soa, n, engine)
        stashes = []
        if need_backward:
            for k in active:
                jk, stash = _lib.jet_forward(bank.single_desc, spec.desc, bank.flat[k : k + 1], x_soa, n,
                                             keep_all=True, engine=engine)
                jets[k] = jk[0]          # the streams the reverse pass is consistent with
                stashes.append(stash)
        ctx.bank, ctx.spec, ctx.engine, ctx.n, ctx.active = bank, spec, engine, n, tuple(active)
        ctx.x_soa, ctx.stashes, ctx.kept = x_soa, stashes, bool(need_backward)
        K, S, C = jets.shape[0], jets.shape[1], jets.shape[2]
        return tuple(jets[:, s, :, :n].permute(2, 0, 1).reshape(n, K * C) for s in range(S))

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *gouts):
        if not ctx.kept:
            raise RuntimeError("fused forward ran without a stash; parameter gradients unavailable")
        bank, spec, n = ctx.bank, ctx.spec, ctx.n
        C = bank.layer_sizes[-1]
        ldn = ctx.x_soa.shape[1]
        grads = []
        for k, stash in zip(ctx.active, ctx.stashes):
            jets_bar = torch.zeros((1, spec.S, C, ldn), dtype=torch.float32, device=bank.flat.device)
            for s, g in enumerate(gouts):
                if g is not None:
                    jets_bar[0, s, :, :n] = g[:, k * C : (k + 1) * C].t()
            g_flat = _lib.jet_backward(bank.single_desc, spec.desc, bank.flat[k : k + 1], ctx.x_soa, n,
                                       jets_bar, stash, engine=ctx.engine)[0]
            off = 0
            for shp in bank.shapes:
                cnt = 1
                for d in shp:
                    cnt *= d
                grads.append(g_flat[off : off + cnt].view(shp))
                off += cnt
        ctx.stashes = None
        return (None,) * 7 + tuple(grads)


class NetBank:
    """K structurally identical FNNs sharing one [K, P] master buffer, evaluated on the same points
    by a single fused kernel sequence (`n_nets = K` in the C ABI)."""

    def __init__(self, nets):
        first = nets[0]
        for n_ in nets:
            if n_.layer_sizes != first.layer_sizes or n_.activation_name != first.activation_name:
                raise ValueError("NetBank needs structurally identical nets")
            if n_._input_transform is not None:
                raise NotImplementedError("feature transforms are not in the fused jet path yet")
        self.nets = list(nets)
        self.K = len(nets)
        self.layer_sizes = first.layer_sizes
        self.net_desc = _lib.make_net_desc(first.layer_sizes, first.activation_name, n_nets=self.K)
        self.single_desc = _lib.make_net_desc(first.layer_sizes, first.activation_name, n_nets=1)
        P = _lib.param_count(self.single_desc) if False else sum(p.numel() for p in first._param_list())
        self.flat = torch.empty((self.K, P), dtype=torch.float32, device=config.device())
        for k, n_ in enumerate(self.nets):
            n_.adopt_storage(self.flat[k])
        self.shapes = [p.shape for p in first._param_list()]

    def evaluate(self, spec, x_soa, n, engine, max_stash_bytes=4 << 30):
        """Forward-only streams [K, S, C, ldn] of all K nets.  One launch sequence when the engine keeps no
        stash (tensor-core engine, width 64) or its two-layer ping-pong stash is small; otherwise the nets go
        in groups so that the scratch stays below `max_stash_bytes`."""
        K, ldn = self.K, x_soa.shape[1]
        per_net = max(_lib.stash_bytes(self.single_desc, spec.desc, ldn, False, engine), 1)
        group = max(1, min(K, max_stash_bytes // per_net))
        if group >= K:
            return _lib.jet_forward(self.net_desc, spec.desc, self.flat, x_soa, n, keep_all=False, engine=engine)[0]
        C = self.layer_sizes[-1]
        jets = torch.empty((K, spec.S, C, ldn), dtype=torch.float32, device=x_soa.device)
        for k0 in range(0, K, group):
            k1 = min(K, k0 + group)
            desc = _lib.make_net_desc(self.layer_sizes, self.nets[0].activation_name, n_nets=k1 - k0)
            jets[k0:k1] = _lib.jet_forward(desc, spec.desc, self.flat[k0:k1], x_soa, n, keep_all=False,
                                           engine=engine)[0]
        return jets

    def _check_storage(self):
        base, P = self.flat.data_ptr(), self.flat.shape[1]
        for k, n_ in enumerate(self.nets):
            if n_._param_list()[0].data_ptr() != base + 4 * k * P:
                n_.adopt_storage(self.flat[k])

    def forward_all(self, x, x_soa=None, spec=None, active=None, param_grad=True):
        """Raw outputs of all K nets on x as one (n, K*C) tensor whose x-derivatives are the
        kernels' jet streams; parameter gradients flow to the `active` nets only."""
        self._check_storage()
        wants_jets = x.requires_grad and torch.is_grad_enabled()
        if not wants_jets:
            spec = jet.JetSpec.value_only()
        elif spec is None:
            spec = jet.JetSpec.full(x.shape[1])
        if x_soa is None or x_soa.shape[1] < x.shape[0]:
            x_soa = jet.to_soa(x)
        active = list(range(self.K)) if active is None else list(active)
        need_bwd = bool(param_grad and torch.is_grad_enabled() and active)
        params = [p for k in active for p in self.nets[k]._param_list()]
        streams = _FusedBankForward.apply(self, x_soa, x.shape[0], spec, config.engine(), need_bwd, active, *params)
        if not wants_jets:
            return streams[0]
        y = jet.make_leaf(x, spec, streams)
        grad.register(y, jet.JetOutput(x, spec, streams))
        return y


class PFNN(NN):
    """dde.nn.PFNN of the fused path (reference: cases/PoissonCases.py:625-626,
    `[1, [20, 20], [20, 20], [20, 20], 2]`): one sub-network per output.  The fully split form with
    structurally identical sub-networks -- what the reference builds -- is C independent FNNs with one
    output each, evaluated by ONE fused kernel sequence through a NetBank (`n_nets = C`); their outputs form
    a single (n, C) jet leaf, so `dde.grad.jacobian(y, x, i=c)` reads net c's streams directly.  Parameter
    names follow DeepXDE (`layers.{i}.{j}.weight`).  A shared trunk before the split, or sub-networks of
    different widths, are not in the fused path (NotImplementedError)."""

    def __init__(self, layer_sizes, activation, kernel_initializer):
        super().__init__()
        if len(layer_sizes) <= 2 or not isinstance(layer_sizes[0], int) or not isinstance(layer_sizes[-1], int):
            raise ValueError("PFNN needs integer input / output sizes and at least one hidden layer")
        n_out = int(layer_sizes[-1])
        hidden = layer_sizes[1:-1]
        for h in hidden:
            if not isinstance(h, (list, tuple)):
                raise NotImplementedError("a shared layer before the split is not in the fused PFNN path")
            if len(h) != n_out:
                raise ValueError("number of sub-layers should equal number of network outputs")
            if len(set(int(w) for w in h)) != 1:
                raise NotImplementedError("sub-networks of different widths are not in the fused PFNN path")
        self.layer_sizes = list(layer_sizes)
        sub_sizes = [int(layer_sizes[0])] + [int(h[0]) for h in hidden] + [1]
        subnets = [FNN(sub_sizes, activation, kernel_initializer) for _ in range(n_out)]
        object.__setattr__(self, "_subnets", subnets)          # parameters are registered through `layers`
        self.layers = torch.nn.ModuleList(
            torch.nn.ModuleList(sub.linears[i] for sub in subnets) for i in range(len(sub_sizes) - 1))
        object.__setattr__(self, "_bank", NetBank(subnets))
        self._scope = _JetScope()

    @contextlib.contextmanager
    def jet_scope(self, spec=None, x_soa=None, param_grad=True, probe_streams=None):
        prev = self._scope
        self._scope = _JetScope(spec, x_soa, param_grad, probe_streams)
        try:
            yield
        finally:
            self._scope = prev

    def forward(self, inputs):
        sc = self._scope
        x = inputs
        featured = self._input_transform is not None
        if featured:
            x = self._input_transform(inputs)
        if sc.probe_streams is not None:
            y = jet.make_leaf(x, jet.probe_spec(x.shape[1]), sc.probe_streams)
        else:
            wants_jets = x.requires_grad and torch.is_grad_enabled()
            x_soa = None if featured else sc.x_soa
            y = self._bank.forward_all(x, x_soa=x_soa, spec=sc.spec if wants_jets else None,
                                       param_grad=sc.param_grad)
        if self._output_transform is not None:
            y = self._output_transform(inputs, y)
        return y


class _FusedRowsForward(torch.autograd.Function):
    """K structurally identical nets whose parameter rows [K, P] are ordinary differentiable tensors (built
    from tied / sliced nn.Parameters by torch ops): one fused launch sequence evaluates all of them, the
    outputs are SUMMED over the nets, and the kernels' flat row gradients flow back through autograd."""

    @staticmethod
    def forward(ctx, desc, rows, x_soa, n, spec, engine, need_backward):
        rows_c = rows.detach().contiguous()
        jets, stash = _lib.jet_forward(desc, spec.desc, rows_c, x_soa, n, keep_all=need_backward, engine=engine)
        ctx.desc, ctx.spec, ctx.engine, ctx.n = desc, spec, engine, n
        ctx.rows, ctx.x_soa, ctx.stash, ctx.kept = rows_c, x_soa, stash, bool(need_backward)
        return tuple(jets[:, s, :, :n].sum(dim=0).t() for s in range(spec.S))

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *gouts):
        if not ctx.kept:
            raise RuntimeError("fused forward ran without a stash; parameter gradients unavailable")
        K, n, spec = ctx.rows.shape[0], ctx.n, ctx.spec
        ldn = ctx.x_soa.shape[1]
        c_out = ctx.desc.widths[ctx.desc.n_linear]
        jets_bar = torch.zeros((K, spec.S, c_out, ldn), dtype=torch.float32, device=ctx.rows.device)
        for s, g in enumerate(gouts):
            if g is not None:
                jets_bar[:, s, :, :n] = g.t().unsqueeze(0)
        g_rows = _lib.jet_backward(ctx.desc, spec.desc, ctx.rows, ctx.x_soa, n, jets_bar, ctx.stash, engine=ctx.engine)
        ctx.stash = None
        return None, g_rows, None, None, None, None, None


class MsFFN(NN):
    """dde.nn.MsFFN of the fused path (reference: cases/PoissonCases.py:478-479, sigmas [1, 10]).

    For every sigma the network is an ordinary MLP whose FIRST layer is the fixed Fourier embedding
    [cos(x b), sin(x b)] = sin(W0 x + phase) with W0 = [b^T; b^T] and phase = [pi/2 ... ; 0 ...], whose hidden
    layers are the shared `linears` and whose output layer is that sigma's column slice of `_dense` -- and the
    MsFFN output is the SUM of those K nets.  So it runs as K nets in one fused launch sequence
    (`n_nets = K`, `first_activation = sin`, FP32 engine); the parameter rows are assembled from the shared
    nn.Parameters by torch ops, which routes the kernels' row gradients back to `linears` / `_dense` (the
    embedding rows get no update: `b` are plain tensors, as in DeepXDE)."""

    def __init__(self, layer_sizes, activation, kernel_initializer, sigmas, dropout_rate=0):
        super().__init__()
        if dropout_rate:
            raise NotImplementedError("dropout is not in the fused path")
        if not isinstance(activation, str):
            raise NotImplementedError("callable activations are not in the fused path (tanh, sin, sigmoid are)")
        if len(layer_sizes) < 4:
            raise ValueError("MsFFN needs at least one hidden layer after the Fourier embedding")
        self.layer_sizes = [int(w) for w in layer_sizes]
        self.sigmas = list(sigmas)
        self.activation_name = activation
        half = self.layer_sizes[1] // 2
        if 2 * half != self.layer_sizes[2 - 1]:
            raise ValueError("layer_sizes[1] must be even (cos and sin halves of the embedding)")
        init = _INIT[kernel_initializer]
        self.b = [torch.normal(mean=0.0, std=float(sg), size=(self.layer_sizes[0], half)) for sg in self.sigmas]
        self.linears = torch.nn.ModuleList()
        for i in range(2, len(self.layer_sizes) - 1):
            lin = torch.nn.Linear(self.layer_sizes[i - 1], self.layer_sizes[i], dtype=torch.float32)
            init(lin.weight)
            torch.nn.init.zeros_(lin.bias)
            self.linears.append(lin)
        self._dense = torch.nn.Linear(self.layer_sizes[-2] * len(self.sigmas), self.layer_sizes[-1], dtype=torch.float32)
        init(self._dense.weight)
        torch.nn.init.zeros_(self._dense.bias)
        self.net_desc = _lib.make_net_desc(self.layer_sizes, activation, n_nets=len(self.sigmas), first_activation="sin")
        self._scope = _JetScope()
        self.to(config.device())

    @contextlib.contextmanager
    def jet_scope(self, spec=None, x_soa=None, param_grad=True, probe_streams=None):
        prev = self._scope
        self._scope = _JetScope(spec, x_soa, param_grad, probe_streams)
        try:
            yield
        finally:
            self._scope = prev

    def _rows(self):
        dev = self._dense.weight.device
        K, h = len(self.sigmas), self.layer_sizes[-2]
        half = self.layer_sizes[1] // 2
        phase = torch.cat([torch.full((half,), torch.pi / 2), torch.zeros(half)]).to(dev)
        shared = [t.reshape(-1) for lin in self.linears for t in (lin.weight, lin.bias)]
        rows = []
        for k, b in enumerate(self.b):
            bt = b.to(device=dev, dtype=torch.float32).t()
            w0 = torch.cat([bt, bt], dim=0).reshape(-1)                   # (2*half, d_in): cos rows, then sin rows
            w_out = self._dense.weight[:, k * h:(k + 1) * h].reshape(-1)
            rows.append(torch.cat([w0, phase] + shared + [w_out, self._dense.bias / K]))
        return torch.stack(rows)

    def forward(self, inputs):
        sc = self._scope
        x = inputs
        featured = self._input_transform is not None
        if featured:
            x = self._input_transform(inputs)
        if sc.probe_streams is not None:
            y = jet.make_leaf(x, jet.probe_spec(x.shape[1]), sc.probe_streams)
        else:
            wants_jets = x.requires_grad and torch.is_grad_enabled()
            if not wants_jets:
                spec = jet.JetSpec.value_only()
            elif sc.spec is not None:
                spec = sc.spec
            else:
                spec = jet.JetSpec.full(x.shape[1])
            if featured or sc.x_soa is None or sc.x_soa.shape[1] < x.shape[0]:
                x_soa = jet.to_soa(x.detach().to(torch.float32))
            else:
                x_soa = sc.x_soa
            rows = self._rows()
            need_bwd = bool(sc.param_grad and torch.is_grad_enabled() and rows.requires_grad)
            streams = _FusedRowsForward.apply(self.net_desc, rows, x_soa, x.shape[0], spec, config.engine(), need_bwd)
            if wants_jets:
                y = jet.make_leaf(x, spec, streams)
                grad.register(y, jet.JetOutput(x, spec, streams))
            else:
                # a fresh tensor, not a view of the kernels' jet buffer: callers update it in place
                # (algorithm/FBPINN.py:203 `result += y`)
                y = streams[0].clone()
        if self._output_transform is not None:
            y = self._output_transform(inputs, y)
        return y


class _FusedRowsEach(torch.autograd.Function):
    """As _FusedRowsForward with the K nets' outputs kept apart: S tensors (K, n, C)."""

    @staticmethod
    def forward(ctx, desc, rows, x_soa, n, spec, engine, need_backward):
        rows_c = rows.detach().contiguous()
        jets, stash = _lib.jet_forward(desc, spec.desc, rows_c, x_soa, n, keep_all=need_backward, engine=engine)
        ctx.desc, ctx.spec, ctx.engine, ctx.n = desc, spec, engine, n
        ctx.rows, ctx.x_soa, ctx.stash, ctx.kept = rows_c, x_soa, stash, bool(need_backward)
        return tuple(jets[:, s, :, :n].transpose(1, 2) for s in range(spec.S))

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *gouts):
        if not ctx.kept:
            raise RuntimeError("fused forward ran without a stash; parameter gradients unavailable")
        K, n, spec = ctx.rows.shape[0], ctx.n, ctx.spec
        ldn = ctx.x_soa.shape[1]
        c_out = ctx.desc.widths[ctx.desc.n_linear]
        jets_bar = torch.zeros((K, spec.S, c_out, ldn), dtype=torch.float32, device=ctx.rows.device)
        for s, g in enumerate(gouts):
            if g is not None:
                jets_bar[:, s, :, :n] = g.transpose(1, 2)
        g_rows = _lib.jet_backward(ctx.desc, spec.desc, ctx.rows, ctx.x_soa, n, jets_bar, ctx.stash, engine=ctx.engine)
        ctx.stash = None
        return None, g_rows, None, None, None, None, None


def _activation_jet(name, z0):
    """value, first and second derivative of the activation at the pre-activation z0 (torch ops, differentiable)."""
    if name == "tanh":
        a = torch.tanh(z0)
        s1 = 1 - a * a
        return a, s1, -2 * a * s1
    if name == "sin":
        a = torch.sin(z0)
        return a, torch.cos(z0), -a
    if name == "sigmoid":
        a = torch.sigmoid(z0)
        s1 = a * (1 - a)
        return a, s1, s1 * (1 - 2 * a)
    raise NotImplementedError(f"activation {name!r} is not in the fused path (tanh, sin, sigmoid are)")


class STMsFFN(NN):
    """dde.nn.STMsFFN of the fused path (reference: cases/WaveCases.py:125-128, sigmas_x=[1], sigmas_t=[1, 10]).

    The last input column is time.  Every space scale and every time scale is an ordinary MLP over its own
    columns -- fixed Fourier embedding as a `sin` first layer (as MsFFN above), the shared `linears` behind it --
    so the K_x space nets and the K_t time nets each run as ONE fused launch sequence whose output layer is the
    last shared Linear (pre-activation, with its value / first / second derivative streams).  What the kernels do
    not do is cheap and stays in torch on the device: the last activation applied to the jets, the point-wise
    product of every (space, time) pair (a space jet times a time jet: d/dx acts on one factor, d/dt on the other)
    and `_dense`.  The result is handed to dde.grad.* as the streams of the requested pattern, so a case's
    jacobian / hessian calls read them exactly as they do for an FNN."""

    def __init__(self, layer_sizes, activation, kernel_initializer, sigmas_x, sigmas_t, dropout_rate=0):
        super().__init__()
        if dropout_rate:
            raise NotImplementedError("dropout is not in the fused path")
        if not isinstance(activation, str):
            raise NotImplementedError("callable activations are not in the fused path (tanh, sin, sigmoid are)")
        if len(layer_sizes) < 4:
            raise ValueError("STMsFFN needs at least one hidden layer after the Fourier embedding")
        if layer_sizes[0] < 2:
            raise ValueError("STMsFFN needs at least one space column before the time column")
        self.layer_sizes = [int(w) for w in layer_sizes]
        self.sigmas_x, self.sigmas_t = list(sigmas_x), list(sigmas_t)
        self.sigmas = self.sigmas_x + self.sigmas_t
        self.activation_name = activation
        _activation_jet(activation, torch.zeros(1))
        d_in, half = self.layer_sizes[0], self.layer_sizes[1] // 2
        if 2 * half != self.layer_sizes[1]:
            raise ValueError("layer_sizes[1] must be even (cos and sin halves of the embedding)")
        init = _INIT[kernel_initializer]
        # draws in DeepXDE's order: MsFFN.__init__ (b of the joint scales, linears, its _dense), then b_x, b_t, _dense
        for sg in self.sigmas:
            torch.normal(mean=0.0, std=float(sg), size=(d_in, half))
        self.linears = torch.nn.ModuleList()
        for i in range(2, len(self.layer_sizes) - 1):
            lin = torch.nn.Linear(self.layer_sizes[i - 1], self.layer_sizes[i], dtype=torch.float32)
            init(lin.weight)
            torch.nn.init.zeros_(lin.bias)
            self.linears.append(lin)
        init(torch.empty(self.layer_sizes[-1], self.layer_sizes[-2] * len(self.sigmas)))
        self.b_x = [torch.normal(mean=0.0, std=float(sg), size=(d_in - 1, half)) for sg in self.sigmas_x]
        self.b_t = [torch.normal(mean=0.0, std=float(sg), size=(1, half)) for sg in self.sigmas_t]
        self._dense = torch.nn.Linear(self.layer_sizes[-2] * len(self.sigmas_x) * len(self.sigmas_t),
                                      self.layer_sizes[-1], dtype=torch.float32)
        init(self._dense.weight)
        torch.nn.init.zeros_(self._dense.bias)
        hidden = self.layer_sizes[1:-1]
        self._desc_x = _lib.make_net_desc([d_in - 1] + hidden, activation, n_nets=len(self.sigmas_x),
                                          first_activation="sin")
        self._desc_t = _lib.make_net_desc([1] + hidden, activation, n_nets=len(self.sigmas_t), first_activation="sin")
        self._scope = _JetScope()
        self.to(config.device())

    @contextlib.contextmanager
    def jet_scope(self, spec=None, x_soa=None, param_grad=True, probe_streams=None):
        prev = self._scope
        self._scope = _JetScope(spec, x_soa, param_grad, probe_streams)
        try:
            yield
        finally:
            self._scope = prev

    def _rows(self, bs):
        dev = self._dense.weight.device
        half = self.layer_sizes[1] // 2
        phase = torch.cat([torch.full((half,), torch.pi / 2), torch.zeros(half)]).to(dev)
        shared = [t.reshape(-1) for lin in self.linears for t in (lin.weight, lin.bias)]
        rows = []
        for b in bs:
            bt = b.to(device=dev, dtype=torch.float32).t()
            rows.append(torch.cat([torch.cat([bt, bt], dim=0).reshape(-1), phase] + shared))
        return torch.stack(rows)

    def _branch(self, desc, bs, cols, n, spec, need_bwd):
        """Activated hidden jets of one family of nets: (a, {dim: a_dim}, {(dim, dim'): a_dim_dim'}), each (K, n, H);
        dims are the family's own column indices."""
        rows = self._rows(bs)
        z = _FusedRowsEach.apply(desc, rows, jet.to_soa(cols), n, spec, config.engine(),
                                 bool(need_bwd and rows.requires_grad))
        a, s1, s2 = _activation_jet(self.activation_name, z[0])
        first = {spec.dirs[j]: s1 * z[1 + j] for j in range(spec.n1)}
        second = {}
        for q, (pa, pb) in enumerate(spec.pairs):
            da, db = spec.dirs[pa], spec.dirs[pb]
            second[(da, db)] = second[(db, da)] = s2 * z[1 + pa] * z[1 + pb] + s1 * z[1 + spec.n1 + q]
        return a, first, second

    def _streams(self, x, spec, need_bwd):
        n, d = x.shape
        t_dim = d - 1
        xs, ts = [p for p in spec.dirs if p < t_dim], [p for p in spec.dirs if p == t_dim]
        need2 = set()
        if spec.third is not None or spec.triples:
            raise NotImplementedError("third- and higher-order derivatives of an STMsFFN are not in the fused path "
                                      "(its space and time nets carry second-order jets)")
        for pa, pb in spec.pairs:
            if pa == jet.JetSpec.TRACE and pb == jet.JetSpec.TRACE:
                need2 |= {(p, p) for p in spec.dirs}
            else:
                need2.add(tuple(sorted((spec.dirs[pa], spec.dirs[pb]))))
        sx = jet.JetSpec.covering(xs, [pq for pq in need2 if pq[1] < t_dim])
        st = jet.JetSpec.covering([0] if ts else [], [(0, 0)] if (t_dim, t_dim) in need2 else [])
        xd = x.detach().to(torch.float32)
        ax, ax1, ax2 = self._branch(self._desc_x, self.b_x, xd[:, :t_dim], n, sx, need_bwd)
        at, at1, at2 = self._branch(self._desc_t, self.b_t, xd[:, t_dim:], n, st, need_bwd)
        at1 = {t_dim: v for v in at1.values()}
        at2 = {(t_dim, t_dim): v for v in at2.values()}

        def pair_product(fx, ft):   # (Kx, n, H) x (Kt, n, H) -> (n, Kx * Kt * H), space index outer
            return (fx.unsqueeze(1) * ft.unsqueeze(0)).permute(2, 0, 1, 3).reshape(n, -1)

        def second(p, q):
            if p == t_dim and q == t_dim:
                return pair_product(ax, at2[(t_dim, t_dim)])
            if p == t_dim or q == t_dim:
                return pair_product(ax1[min(p, q)], at1[t_dim])
            return pair_product(ax2[(p, q)], at)

        feats = [pair_product(ax, at)]
        for p in spec.dirs:
            feats.append(pair_product(ax, at1[t_dim]) if p == t_dim else pair_product(ax1[p], at))
        for pa, pb in spec.pairs:
            if pa == jet.JetSpec.TRACE and pb == jet.JetSpec.TRACE:
                lap = None
                for p in spec.dirs:
                    term = second(p, p)
                    lap = term if lap is None else lap + term
                feats.append(lap)
            else:
                feats.append(second(spec.dirs[pa], spec.dirs[pb]))
        w = self._dense.weight.t()
        return [feats[0] @ w + self._dense.bias] + [f @ w for f in feats[1:]]

    def forward(self, inputs):
        sc = self._scope
        x = inputs
        if self._input_transform is not None:
            x = self._input_transform(inputs)
        if x.shape[1] != self.layer_sizes[0]:
            raise ValueError(f"the net takes {self.layer_sizes[0]} columns (space..., time), got {x.shape[1]}")
        if sc.probe_streams is not None:
            y = jet.make_leaf(x, jet.probe_spec(x.shape[1]), sc.probe_streams)
        else:
            wants_jets = x.requires_grad and torch.is_grad_enabled()
            if not wants_jets:
                spec = jet.JetSpec.value_only()
            elif sc.spec is not None:
                spec = sc.spec
            else:
                spec = jet.JetSpec.full(x.shape[1])
            streams = self._streams(x, spec, sc.param_grad and torch.is_grad_enabled())
            if wants_jets:
                y = jet.make_leaf(x, spec, streams)
                grad.register(y, jet.JetOutput(x, spec, streams))
            else:
                # a fresh tensor, not a view of the kernels' jet buffer: callers update it in place
                # (algorithm/FBPINN.py:203 `result += y`)
                y = streams[0].clone()
        if self._output_transform is not None:
            y = self._output_transform(inputs, y)
        return y


def get_supported_activations():
    return sorted(_lib.ACT)
