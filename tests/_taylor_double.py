"""Plain-PyTorch Taylor-mode stand-in for the two kernel entry points (test double).

Same signatures as ai4xde_b200._lib.jet_forward / jet_backward, same tensor layouts, computed
with ordinary torch ops (any device, any float dtype).  Two uses, both in tests only:
  * `not gpu` tests monkey-patch it under the product's host code to exercise the boundary
    (stream discovery, grad interception, losses, Model loops, resamplers' host logic) on CPU;
  * `gpu` tests use it as the plain-PyTorch reference of the fused op.
It is never imported by the product package.
"""
import torch


def _act(name, z):
    """value and first four derivatives of the activation at z"""
    if name == 0:
        a = torch.tanh(z)
        s1 = 1 - a * a
        s2 = -2 * a * s1
        s3 = -2 * s1 * (1 - 3 * a * a)
        s4 = 8 * a * s1 * (2 - 3 * a * a)
    elif name == 1:
        a = torch.sin(z)
        s1 = torch.cos(z)
        s2, s3, s4 = -a, -s1, a
    else:
        a = torch.sigmoid(z)
        s1 = a * (1 - a)
        s2 = s1 * (1 - 2 * a)
        s3 = s1 * (1 + 6 * a * (a - 1))
        s4 = s2 * (1 + 12 * a * (a - 1))
    return a, s1, s2, s3, s4


def _unpack(net, flat):
    ws, off = [], 0
    for l in range(net.n_linear):
        i, o = net.widths[l], net.widths[l + 1]
        W = flat[off : off + i * o].view(o, i)
        off += i * o
        b = flat[off : off + o]
        off += o
        ws.append((W, b))
    return ws


def taylor_streams(net, jet, flat, x):
    """flat: [P] params, x: (n, d) -> list of S tensors (n, C)."""
    n1, n2 = jet.n1, jet.n2
    dirs = [jet.dir1[j] for j in range(n1)]
    pairs = [(jet.pa[q], jet.pb[q]) for q in range(n2)]
    layers = _unpack(net, flat)
    W, b = layers[0]
    z = [x @ W.t() + b]
    z += [W[:, d].unsqueeze(0).expand(x.shape[0], -1) for d in dirs]
    z += [torch.zeros_like(z[0]) for _ in pairs]
    first = getattr(net, "first_activation", 0)
    for li, (W, b) in enumerate(layers[1:]):
        a, s1, s2, s3, s4 = _act(first - 1 if (li == 0 and first > 0) else net.activation, z[0])
        A = [a] + [s1 * z[1 + j] for j in range(n1)]
        for q, (pa, pb) in enumerate(pairs):
            if pa == 15 and pb == 15:      # Laplacian stream over all first-order streams
                A.append(s2 * sum(z[1 + j] * z[1 + j] for j in range(n1)) + s1 * z[1 + n1 + q])
            elif pa == 14 and pb == 14:    # third derivative along the single direction (streams 1..3: orders 1..3 of z)
                A.append(s3 * z[1] ** 3 + 3 * s2 * z[1] * z[2] + s1 * z[3])
            elif pa == pb and 8 <= pa < 12:  # two-direction third order: m factors along direction 1 (streams 1, 2 first; 3, 4, 5 = aa, ab, bb)
                m = pa - 8
                i, j, k = int(m >= 3), int(m >= 2), int(m >= 1)
                A.append(s3 * z[1 + i] * z[1 + j] * z[1 + k]
                         + s2 * (z[3 + i + j] * z[1 + k] + z[3 + i + k] * z[1 + j] + z[3 + j + k] * z[1 + i]) + s1 * z[1 + n1 + q])
            elif pa == 13 and pb == 13:    # fourth derivative
                A.append(s4 * z[1] ** 4 + 6 * s3 * z[1] ** 2 * z[2] + 3 * s2 * z[2] ** 2 + 4 * s2 * z[1] * z[3] + s1 * z[4])
            else:
                A.append(s2 * z[1 + pa] * z[1 + pb] + s1 * z[1 + n1 + q])
        z = [A[0] @ W.t() + b] + [As @ W.t() for As in A[1:]]
    return z


def jet_forward(net, jet, params, x_soa, n, keep_all, engine="auto"):
    K = net.n_nets
    S = 1 + jet.n1 + jet.n2
    C = net.widths[net.n_linear]
    ldn = x_soa.shape[1]
    x = x_soa[:, :n].t()
    jets = torch.zeros((K, S, C, ldn), dtype=params.dtype, device=params.device)
    with torch.no_grad():
        for k in range(K):
            st = taylor_streams(net, jet, params[k], x)
            for s in range(S):
                jets[k, s, :, :n] = st[s].t()
    return jets, torch.zeros(1)


def jet_backward(net, jet, params, x_soa, n, jets_bar, stash, engine="auto"):
    K = net.n_nets
    S = 1 + jet.n1 + jet.n2
    x = x_soa[:, :n].t()
    grads = []
    for k in range(K):
        p = params[k].detach().clone().requires_grad_()
        with torch.enable_grad():
            st = taylor_streams(net, jet, p, x)
            tot = sum((st[s].t() * jets_bar[k, s, :, :n]).sum() for s in range(S))
            (g,) = torch.autograd.grad(tot, p)
        grads.append(g)
    return torch.stack(grads)


# ---- residual programs (ai4xde_b200/residual.py) interpreted in torch: tests only ------------------------------
class ResidualTables:
    """Stand-in for _lib.ResidualTables: keeps the lowered program itself."""

    def __init__(self, program, device=None):
        self.program = program
        self.n_terms, self.n_streams = program.n_terms, program.S * program.C
        self.n_scalars, self.scalar_grad = len(program.scalars), list(program.scalar_grad)


def _run_program(program, x_soa, jets, n):
    """per-term losses of a lowered program, one vectorised sweep per row segment (differentiable w.r.t. jets)"""
    from ai4xde_b200 import residual as R

    flat = jets.reshape(-1, jets.shape[-1])          # [S * C, ldn]
    sums = [torch.zeros((), dtype=jets.dtype) for _ in range(program.n_terms)]
    un = {R.OP_SIN: torch.sin, R.OP_COS: torch.cos, R.OP_EXP: torch.exp, R.OP_TANH: torch.tanh, R.OP_LOG: torch.log,
          R.OP_SQRT: torch.sqrt, R.OP_ABS: torch.abs, R.OP_SIGN: torch.sign, R.OP_NEG: torch.neg}
    for beg, end, code in program.segments:
        assert 0 <= beg < end <= n
        reg = {}
        for op, dst, a, b in code.tolist():
            if op == R.OP_LOADX:
                reg[dst] = x_soa[a, beg + b : end + b].to(jets.dtype)
            elif op == R.OP_LOADJ:
                reg[dst] = flat[a, beg + b : end + b]
            elif op == R.OP_LOADA:
                ten, first = program.aux[a]
                reg[dst] = ten.to(jets.dtype)[beg - first : end - first, b]
            elif op == R.OP_LOADS:
                reg[dst] = torch.Tensor.as_subclass(program.scalars[a], torch.Tensor).to(device=jets.device, dtype=jets.dtype).reshape(()).expand(end - beg)
            elif op == R.OP_CONST:
                reg[dst] = torch.full((end - beg,), float(program.consts[a]), dtype=jets.dtype)
            elif op == R.OP_ADD:
                reg[dst] = reg[a] + reg[b]
            elif op == R.OP_SUB:
                reg[dst] = reg[a] - reg[b]
            elif op == R.OP_MUL:
                reg[dst] = reg[a] * reg[b]
            elif op == R.OP_DIV:
                reg[dst] = reg[a] / reg[b]
            elif op == R.OP_POW:
                reg[dst] = reg[a] ** float(program.consts[b])
            elif op == R.OP_SUMT:
                sums[dst] = sums[dst] + reg[a].sum()
            elif op in (R.OP_STORE, R.OP_SUMT_ADJ):
                pass          # mirror segments only exist for the kernel's adjoint bookkeeping: autograd needs none
            else:
                reg[dst] = un[op](reg[a])
    return torch.stack(sums) * torch.as_tensor(program.scale, dtype=jets.dtype)


def residual_eval(tables, x_soa, jets, n, n_cols):
    """operator programs (OP_STORE): out [n_cols, ldn]"""
    from ai4xde_b200 import residual as R

    program = tables.program
    out = torch.zeros((n_cols, x_soa.shape[1]), dtype=torch.float32)
    flat = jets.reshape(-1, jets.shape[-1]).double()
    un = {R.OP_SIN: torch.sin, R.OP_COS: torch.cos, R.OP_EXP: torch.exp, R.OP_TANH: torch.tanh, R.OP_LOG: torch.log,
          R.OP_SQRT: torch.sqrt, R.OP_ABS: torch.abs, R.OP_SIGN: torch.sign, R.OP_NEG: torch.neg}
    for beg, end, code in program.segments:
        reg = {}
        for op, dst, a, b in code.tolist():
            if op == R.OP_LOADX:
                reg[dst] = x_soa[a, beg + b : end + b].double()
            elif op == R.OP_LOADJ:
                reg[dst] = flat[a, beg + b : end + b]
            elif op == R.OP_LOADA:
                ten, first = program.aux[a]
                reg[dst] = ten.double()[beg - first : end - first, b]
            elif op == R.OP_LOADS:
                reg[dst] = torch.Tensor.as_subclass(program.scalars[a], torch.Tensor).to(device=jets.device).double().reshape(()).expand(end - beg)
            elif op == R.OP_CONST:
                reg[dst] = torch.full((end - beg,), float(program.consts[a]), dtype=torch.float64)
            elif op == R.OP_ADD:
                reg[dst] = reg[a] + reg[b]
            elif op == R.OP_SUB:
                reg[dst] = reg[a] - reg[b]
            elif op == R.OP_MUL:
                reg[dst] = reg[a] * reg[b]
            elif op == R.OP_DIV:
                reg[dst] = reg[a] / reg[b]
            elif op == R.OP_POW:
                reg[dst] = reg[a] ** float(program.consts[b])
            elif op == R.OP_STORE:
                out[dst, beg:end] = reg[a].to(torch.float32)
            elif op in (R.OP_SUMT, R.OP_SUMT_ADJ):
                pass
            else:
                reg[dst] = un[op](reg[a])
    return out


def residual_forward(tables, x_soa, jets, n):
    with torch.no_grad():
        return _run_program(tables.program, x_soa, jets, n).to(torch.float32)


def residual_backward(tables, x_soa, jets, n, term_grads):
    """reverse pass by torch.autograd over the interpreter (the kernel's hand-written adjoint rules are held to it)"""
    j = jets.detach().clone().requires_grad_()
    prog = tables.program
    live = prog.scalars
    stand_ins = [torch.Tensor.as_subclass(t.detach(), torch.Tensor).to(jets.device).clone().requires_grad_() for t in live]
    prog.scalars = stand_ins          # plain leaves for the duration: d/d(scalar) by autograd
    try:
        with torch.enable_grad():
            tot = (_run_program(prog, x_soa, j, n) * term_grads.to(j.dtype)).sum()
            grads = torch.autograd.grad(tot, [j] + stand_ins, allow_unused=True)
    finally:
        prog.scalars = live
    g = torch.zeros_like(jets) if grads[0] is None else grads[0]
    if not tables.scalar_grad:
        return g, None
    sg = torch.stack([torch.zeros((), dtype=jets.dtype) if s is None else s.reshape(()) for s in grads[1:]])
    return g, sg.to(torch.float32)


def torch_lbfgs(params, **kw):
    """Stand-in for the device L-BFGS on CPU tensors: the optimiser it follows statement by statement."""
    return torch.optim.LBFGS(params, line_search_fn=None, **kw)


def install(monkeypatch=None):
    """Route the product's kernel calls (jet kernels, device L-BFGS) to doubles (CPU host-logic tests only)."""
    from ai4xde_b200 import _lib
    from ai4xde_b200.dde import optimizers

    if monkeypatch is not None:
        monkeypatch.setattr(_lib, "jet_forward", jet_forward)
        monkeypatch.setattr(_lib, "jet_backward", jet_backward)
        monkeypatch.setattr(_lib, "ResidualTables", ResidualTables)
        monkeypatch.setattr(_lib, "residual_forward", residual_forward)
        monkeypatch.setattr(_lib, "residual_backward", residual_backward)
        monkeypatch.setattr(_lib, "residual_eval", residual_eval)
        monkeypatch.setattr(optimizers, "lbfgs_factory", torch_lbfgs)
    else:
        _lib.jet_forward, _lib.jet_backward = jet_forward, jet_backward
        _lib.ResidualTables, _lib.residual_forward, _lib.residual_backward = ResidualTables, residual_forward, residual_backward
        _lib.residual_eval = residual_eval
        optimizers.lbfgs_factory = torch_lbfgs
