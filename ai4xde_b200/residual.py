"""Residual programs -- the case's own `pde` / BC / loss code compiled for one kernel.

DeepXDE evaluates a training loss as hundreds of small tensor kernels per step: the residual algebra of
`pde(x, y)` (e.g. /root/reference/AI4XDE/cases/PDECases.py:1087-1094 `du_t + u * du_x - nu * du_xx`), one
`bc.error` slice per boundary condition, a mean-squared reduction per term, and the same again in reverse for
the gradient.  Next to the fused jet kernels that glue is 40 % of a narrow-net step and nearly all of a
launch-bound one (cfg1).  Here the SAME Python code is run ONCE on symbolic columns (`Sym`): every torch /
operator call is recorded into an expression graph, `dde.grad.jacobian / hessian` become symbolic
differentiation that bottoms out in the kernel's jet streams, `torch.mean` closes a loss term.  The graph is
lowered to a small register program (`Program`) that `ax_residual_forward / ax_residual_backward`
(csrc/residual_vm.cu) interpret per collocation point: forward = per-term loss sums, reverse = adjoints of the
jet streams written straight into the layout `ax_jet_mlp_backward` reads.

Anything the tracer does not understand raises `Unsupported` and the caller stays on the ordinary autograd path
(same kernels underneath, same results); a program is adopted only after its losses and parameter gradient
reproduce that path on the live batch (dde/model.py).
"""
import math
import numbers
import numbers as numbers_mod
import sys

import numpy as np
import torch


class Unsupported(Exception):
    """The traced code did something a residual program cannot express (the autograd path serves it)."""


# ---- opcodes (shared with csrc/residual_vm.cu and the test double) -------------------------------------------
OP_LOADX, OP_LOADJ, OP_LOADA, OP_LOADS, OP_CONST = 0, 1, 2, 3, 4
OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG = 5, 6, 7, 8, 9
OP_SIN, OP_COS, OP_EXP, OP_TANH, OP_LOG, OP_SQRT, OP_ABS, OP_SIGN, OP_POW = 10, 11, 12, 13, 14, 15, 16, 17, 18
OP_SUMT, OP_STORE, OP_SUMT_ADJ = 19, 20, 21
_UNARY = {"sin": OP_SIN, "cos": OP_COS, "exp": OP_EXP, "tanh": OP_TANH, "log": OP_LOG, "sqrt": OP_SQRT,
          "abs": OP_ABS, "sign": OP_SIGN, "neg": OP_NEG}
_BINARY = {"add": OP_ADD, "sub": OP_SUB, "mul": OP_MUL, "div": OP_DIV}
MAX_INSTR, MAX_REG, MAX_TERMS, MAX_CONST = 4096, 380, 64, 1024


# ---- expression graph -------------------------------------------------------------------------------------------
class Node:
    __slots__ = ("op", "args", "uid")

    def __init__(self, op, args, uid):
        self.op, self.args, self.uid = op, args, uid

    def __repr__(self):
        return f"{self.op}{self.args}" if self.op in ("x", "net", "const", "aux", "scal") else \
            f"{self.op}({', '.join('#%d' % a.uid for a in self.args)})"


class Graph:
    """Hash-consed expressions over one point: inputs x_d, network streams, constants, per-row arrays, device
    scalars.  `n_in` inputs, `n_out` network outputs."""

    def __init__(self, n_in, n_out):
        self.n_in, self.n_out = n_in, n_out
        self._nodes = {}
        self._diff = {}
        self._shift = {}
        self.features = None    # net.apply_feature_transform(f): the columns of f(x); the network's streams are d/dfeature
        self.aux = []       # (tensor [rows, ncol] on the device, first row)
        self.scalars = []   # one-element device tensors read at run time (their value may change between steps)
        self.scalar_grad = set()   # indices of scalars that are trainable variables: the reverse pass returns d/d(scalar)

    def _mk(self, op, args):
        key = (op,) + tuple(a.uid if isinstance(a, Node) else a for a in args)
        hit = self._nodes.get(key)
        if hit is None:
            hit = self._nodes[key] = Node(op, args, len(self._nodes))
        return hit

    # leaves
    def x(self, d, delta=0):
        return self._mk("x", (int(d), int(delta)))

    def net(self, c, dims=(), delta=0):
        """stream of network output c differentiated along `dims`, read `delta` rows away from the point"""
        return self._mk("net", (int(c), tuple(sorted(int(d) for d in dims)), int(delta)))

    def const(self, v):
        v = float(v)
        return self._mk("const", (v if v == v else "nan",))

    def aux_col(self, k, col, delta=0):
        return self._mk("aux", (int(k), int(col), int(delta)))

    def shift(self, e, delta, _memo=None):
        """The same expression evaluated `delta` rows further down the batch (PeriodicBC pairs row p of the left
        boundary with row p + m of the right one, deepxde icbc PeriodicBC.error)."""
        if delta == 0:
            return e
        memo = self._shift.setdefault(delta, {})
        hit = memo.get(e.uid)
        if hit is not None:
            return hit
        op, a = e.op, e.args
        if op == "x":
            out = self.x(a[0], a[1] + delta)
        elif op == "net":
            out = self.net(a[0], a[1], a[2] + delta)
        elif op == "aux":
            out = self.aux_col(a[0], a[1], a[2] + delta)
        elif op in ("const", "scal"):
            out = e
        elif op == "pow":
            out = self.pow(self.shift(a[0], delta), a[1])
        elif op in ("add", "sub", "mul", "div"):
            out = getattr(self, op)(self.shift(a[0], delta), self.shift(a[1], delta))
        elif op == "neg":
            out = self.neg(self.shift(a[0], delta))
        else:
            out = self.unary(op, self.shift(a[0], delta))
        memo[e.uid] = out
        return out

    def scalar(self, k):
        return self._mk("scal", (int(k),))

    @staticmethod
    def cval(n):
        return n.args[0] if n.op == "const" and n.args[0] != "nan" else None

    # operations (with the algebraic simplifications that keep differentiated programs small)
    def add(self, a, b):
        ca, cb = self.cval(a), self.cval(b)
        if ca is not None and cb is not None:
            return self.const(ca + cb)
        if ca == 0.0:
            return b
        if cb == 0.0:
            return a
        return self._mk("add", (a, b) if a.uid <= b.uid else (b, a))

    def sub(self, a, b):
        ca, cb = self.cval(a), self.cval(b)
        if ca is not None and cb is not None:
            return self.const(ca - cb)
        if cb == 0.0:
            return a
        if ca == 0.0:
            return self.neg(b)
        if a is b:
            return self.const(0.0)
        return self._mk("sub", (a, b))

    def mul(self, a, b):
        ca, cb = self.cval(a), self.cval(b)
        if ca is not None and cb is not None:
            return self.const(ca * cb)
        if ca == 0.0 or cb == 0.0:
            return self.const(0.0)
        if ca == 1.0:
            return b
        if cb == 1.0:
            return a
        if ca == -1.0:
            return self.neg(b)
        if cb == -1.0:
            return self.neg(a)
        return self._mk("mul", (a, b) if a.uid <= b.uid else (b, a))

    def div(self, a, b):
        ca, cb = self.cval(a), self.cval(b)
        if ca is not None and cb is not None and cb != 0.0:
            return self.const(ca / cb)
        if ca == 0.0:
            return self.const(0.0)
        if cb == 1.0:
            return a
        return self._mk("div", (a, b))

    def neg(self, a):
        ca = self.cval(a)
        if ca is not None:
            return self.const(-ca)
        if a.op == "neg":
            return a.args[0]
        return self._mk("neg", (a,))

    def unary(self, name, a):
        ca = self.cval(a)
        if ca is not None:
            try:
                f = {"sin": math.sin, "cos": math.cos, "exp": math.exp, "tanh": math.tanh, "log": math.log,
                     "sqrt": math.sqrt, "abs": abs, "sign": lambda v: float((v > 0) - (v < 0))}[name]
                return self.const(f(ca))
            except (ValueError, OverflowError):
                pass
        return self._mk(name, (a,))

    def pow(self, a, e):
        e = float(e)
        if e == 0.0:
            return self.const(1.0)
        if e == 1.0:
            return a
        if e == 2.0:
            return self.mul(a, a)
        ca = self.cval(a)
        if ca is not None:
            try:
                return self.const(ca ** e)
            except (ValueError, OverflowError, ZeroDivisionError):
                pass
        return self._mk("pow", (a, e))

    # symbolic d/dx_j
    def diff(self, e, j):
        key = (e.uid, j)
        hit = self._diff.get(key)
        if hit is None:
            hit = self._diff[key] = self._diff_new(e, j)
        return hit

    def _diff_new(self, e, j):
        op, a = e.op, e.args
        D = self.diff
        if op in ("x", "net") and a[-1] != 0:
            raise Unsupported("derivative of a row-shifted expression")
        if op == "x":
            return self.const(1.0 if a[0] == j else 0.0)
        if op == "net":
            if len(a[1]) >= 4:
                raise Unsupported("fifth-order derivative of the network")
            if self.features is not None:
                # the kernels differentiate w.r.t. the FEATURES phi = f(x): chain rule through f, symbolically
                tot = self.const(0.0)
                for m, phi in enumerate(self.features):
                    dphi = D(phi, j)
                    if self.cval(dphi) != 0.0:
                        tot = self.add(tot, self.mul(self.net(a[0], a[1] + (m,)), dphi))
                return tot
            return self.net(a[0], a[1] + (j,))
        if op in ("const", "aux", "scal", "sign"):
            return self.const(0.0)
        if op == "add":
            return self.add(D(a[0], j), D(a[1], j))
        if op == "sub":
            return self.sub(D(a[0], j), D(a[1], j))
        if op == "neg":
            return self.neg(D(a[0], j))
        if op == "mul":
            return self.add(self.mul(D(a[0], j), a[1]), self.mul(a[0], D(a[1], j)))
        if op == "div":      # (a' - (a / b) b') / b
            return self.div(self.sub(D(a[0], j), self.mul(e, D(a[1], j))), a[1])
        da = D(a[0], j)
        if self.cval(da) == 0.0:
            return da
        if op == "sin":
            return self.mul(self.unary("cos", a[0]), da)
        if op == "cos":
            return self.neg(self.mul(self.unary("sin", a[0]), da))
        if op == "exp":
            return self.mul(e, da)
        if op == "tanh":
            return self.mul(self.sub(self.const(1.0), self.mul(e, e)), da)
        if op == "log":
            return self.div(da, a[0])
        if op == "sqrt":
            return self.div(da, self.mul(self.const(2.0), e))
        if op == "abs":
            return self.mul(self.unary("sign", a[0]), da)
        if op == "pow":
            return self.mul(self.mul(self.const(a[1]), self.pow(a[0], a[1] - 1.0)), da)
        raise Unsupported(f"derivative of {op}")


# ---- trainable scalars (dde.Variable of inverse problems) ----------------------------------------------------------
_TRACE = {}       # "graph": the Graph being built while a trace runs


class VariableTensor(torch.Tensor):
    """What `dde.Variable(v)` returns: an ordinary trainable leaf tensor everywhere -- except while a residual
    program is being traced, when any torch operation on it (`-self.v_e / self.e`, `1e-3 * self.D`, `torch.exp(C)`,
    cases/InverseCases.py:314,505) is recorded as an expression of a run-time scalar instead of being evaluated:
    the program reads the variable's CURRENT value from its storage at every step and returns its gradient."""

    @staticmethod
    def make(data):
        return torch.Tensor._make_subclass(VariableTensor, data, True)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        g = _TRACE.get("graph")
        if g is None:
            return super().__torch_function__(func, types, args, kwargs or {})
        name = _name_of(func)
        if name in ("__get__", "__set__", "data_ptr", "numel", "dim", "size", "is_floating_point", "__repr__",
                    "__len__", "__hash__", "__reduce_ex__", "__deepcopy__"):
            return super().__torch_function__(func, types, args, kwargs or {})      # attribute reads are not algebra
        return Sym.__torch_function__(func, types, args, kwargs)


def _scalar_sym(g, var):
    """scalar Sym over the storage of a dde.Variable (registered once per trace, by identity)"""
    if g is None:
        raise Unsupported("a dde.Variable used outside a trace")
    with torch._C.DisableTorchFunctionSubclass():
        if var.numel() != 1:
            raise Unsupported("a trainable variable with more than one element")
        if not var.is_leaf:
            raise Unsupported("a derived tensor of a variable computed before the trace")
    for k, t in enumerate(g.scalars):
        if t is var:
            break
    else:
        g.scalars.append(var)
        k = len(g.scalars) - 1
    g.scalar_grad.add(k)
    return Sym(g, [g.scalar(k)], None)


# ---- symbolic tensors ---------------------------------------------------------------------------------------------
class Term:
    """One loss term: scale * sum over the rows of `win` of `expr` (what torch.mean / torch.sum of a Sym returns)."""

    def __init__(self, graph, expr, win, scale):
        self.g, self.expr, self.win, self.scale = graph, expr, win, float(scale)

    def _scaled(self, k):
        if isinstance(k, numbers.Number):
            return Term(self.g, self.expr, self.win, self.scale * float(k))
        raise Unsupported("arithmetic on a reduced loss term")

    __mul__ = __rmul__ = _scaled

    def __truediv__(self, k):
        if isinstance(k, numbers.Number):
            return self._scaled(1.0 / float(k))
        raise Unsupported("arithmetic on a reduced loss term")


def _name_of(func):
    return getattr(func, "__name__", str(func))


class Sym:
    """(rows, k) tensor of per-point expressions on rows [win[0], win[1]) of the batch; `vec`: shape (rows,)."""

    __array_ufunc__ = None      # numpy hands `ndarray * Sym` to Sym.__rmul__ instead of broadcasting over an object
    requires_grad = True
    dtype = torch.float32
    is_input = False

    scalar = False      # True: one value for every point (an expression of trainable variables); no row range

    def __init__(self, graph, cols, win, vec=False):
        self.g, self.cols, self.vec = graph, list(cols), vec
        if win is None:
            self.win, self.scalar = (0, 1), True
        else:
            self.win = (int(win[0]), int(win[1]))

    # -- tensor-like surface --------------------------------------------------------------------------------------
    @property
    def shape(self):
        rows = self.win[1] - self.win[0]
        return torch.Size((rows,)) if self.vec else torch.Size((rows, len(self.cols)))

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self):
        return 1 if self.vec else 2

    ndim = property(dim)

    def __len__(self):
        return self.win[1] - self.win[0]

    def detach(self):
        raise Unsupported("detach() inside the traced code")

    def numel(self):
        return int(np.prod(self.shape))

    def _like(self, cols, vec=None):
        return Sym(self.g, cols, None if self.scalar else self.win, self.vec if vec is None else vec)

    def __getitem__(self, idx):
        if self.scalar:
            raise Unsupported("indexing a variable expression")
        rows, colspec = (idx if isinstance(idx, tuple) else (idx, slice(None)))
        if isinstance(idx, tuple) and len(idx) != 2:
            raise Unsupported("index with more than two axes")
        if self.vec and isinstance(idx, tuple):
            raise Unsupported("two-axis index into a one-axis tensor")
        if rows is Ellipsis:
            rows = slice(None)
        if not isinstance(rows, slice):
            raise Unsupported(f"row index {type(rows).__name__} (only slices keep the per-point structure)")
        b, e, step = rows.indices(len(self))
        if step != 1:
            raise Unsupported("strided row slice")
        win = (self.win[0] + b, self.win[0] + max(b, e))
        if isinstance(colspec, slice):
            cols = self.cols[colspec]
            if not cols:
                raise Unsupported("empty column slice")
            return Sym(self.g, cols, win, self.vec)
        if isinstance(colspec, numbers.Integral):
            return Sym(self.g, [self.cols[int(colspec)]], win, True)
        raise Unsupported(f"column index {type(colspec).__name__}")

    # -- operands -------------------------------------------------------------------------------------------------
    def _operand(self, o):
        """-> list of column expressions broadcast against self (len 1 or len(self.cols))"""
        g = self.g
        if isinstance(o, VariableTensor):
            o = _scalar_sym(g, o)
        if isinstance(o, Sym):
            if o.g is not g:
                raise Unsupported("operands of two traces")
            if o.scalar:
                return o.cols          # a variable expression: the same value at every point
            if self.scalar:
                raise Unsupported("internal: scalar operand order")
            if o.win != self.win:
                if o.win[1] - o.win[0] != self.win[1] - self.win[0] or o.vec != self.vec:
                    raise Unsupported(f"operands cover different rows {o.win} / {self.win}")
                # row i of one slice against row i of another (PeriodicBC: left boundary rows vs their periodic
                # images): the other operand is read (o.win[0] - self.win[0]) rows away
                return [g.shift(c, o.win[0] - self.win[0]) for c in o.cols]
            if o.vec != self.vec:
                raise Unsupported("one-axis with two-axis operand")
            return o.cols
        if isinstance(o, Term):
            raise Unsupported("a reduced term used point-wise")
        if isinstance(o, (numbers.Number, np.generic)):
            return [g.const(float(o))]
        if isinstance(o, np.ndarray):
            if o.size == 1:
                return [g.const(float(o.reshape(-1)[0]))]
            o = torch.as_tensor(o)
        if torch.is_tensor(o):
            if o.requires_grad:
                raise Unsupported("a trainable tensor that is not a dde.Variable inside the residual")
            if self.scalar:
                raise Unsupported("a variable expression combined with an array outside the point-wise algebra")
            if o.numel() == 1:
                if not (o.is_cuda and o.dtype == torch.float32):
                    return [g.const(float(o))]       # a host number: frozen like any Python constant
                g.scalars.append(o)                   # a device scalar: read at run time (its value may move)
                return [g.scalar(len(g.scalars) - 1)]
            rows = len(self)
            if self.vec:
                if o.dim() != 1 or o.shape[0] != rows:
                    raise Unsupported(f"array of shape {tuple(o.shape)} against ({rows},)")
                o = o.reshape(rows, 1)
            elif o.dim() != 2 or o.shape[0] != rows or o.shape[1] not in (1, len(self.cols)):
                if not (o.dim() == 2 and o.shape[0] == rows and len(self.cols) == 1):
                    raise Unsupported(f"array of shape {tuple(o.shape)} against {tuple(self.shape)}")
            g.aux.append((o, self.win[0]))
            k = len(g.aux) - 1
            return [g.aux_col(k, c) for c in range(o.shape[1])]
        raise Unsupported(f"operand of type {type(o).__name__}")

    def _binary(self, o, fn, swap=False):
        if self.scalar and isinstance(o, Sym) and not o.scalar:
            return o._binary(self, fn, swap=not swap)      # the point-wise operand decides rows and columns
        oc = self._operand(o)
        mine = self.cols
        n = max(len(mine), len(oc))
        if len(mine) not in (1, n) or len(oc) not in (1, n):
            raise Unsupported("column counts do not broadcast")
        out = []
        for i in range(n):
            a, b = mine[i if len(mine) > 1 else 0], oc[i if len(oc) > 1 else 0]
            out.append(fn(b, a) if swap else fn(a, b))
        return self._like(out)

    def __add__(self, o):
        return self._binary(o, self.g.add)

    __radd__ = __add__

    def __sub__(self, o):
        return self._binary(o, self.g.sub)

    def __rsub__(self, o):
        return self._binary(o, self.g.sub, swap=True)

    def __mul__(self, o):
        return self._binary(o, self.g.mul)

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self._binary(o, self.g.div)

    def __rtruediv__(self, o):
        return self._binary(o, self.g.div, swap=True)

    def __neg__(self):
        return self._like([self.g.neg(c) for c in self.cols])

    def __pos__(self):
        return self

    def __pow__(self, e):
        if isinstance(e, (numbers.Number, np.generic)):
            return self._like([self.g.pow(c, float(e)) for c in self.cols])
        raise Unsupported("non-constant exponent")

    def __rpow__(self, base):
        if isinstance(base, (numbers.Number, np.generic)) and float(base) > 0:
            k = math.log(float(base))
            return self._like([self.g.unary("exp", self.g.mul(self.g.const(k), c)) for c in self.cols])
        raise Unsupported("non-constant base")

    def _unary(self, name):
        return self._like([self.g.unary(name, c) for c in self.cols])

    def sin(self):
        return self._unary("sin")

    def cos(self):
        return self._unary("cos")

    def exp(self):
        return self._unary("exp")

    def tanh(self):
        return self._unary("tanh")

    def log(self):
        return self._unary("log")

    def sqrt(self):
        return self._unary("sqrt")

    def abs(self):
        return self._unary("abs")

    __abs__ = abs

    def square(self):
        return self._like([self.g.mul(c, c) for c in self.cols])

    def mean(self, *a, **k):
        if a or k:
            raise Unsupported("mean over one axis")
        return self._reduce(1.0 / max(self.numel(), 1))

    def sum(self, dim=None, keepdim=False, **k):
        dim = k.pop("axis", dim)
        if k:
            raise Unsupported(f"sum({k})")
        if dim is None:
            return self._reduce(1.0)
        if self.vec or dim not in (1, -1):
            raise Unsupported("sum over the point axis")
        tot = self.cols[0]
        for c in self.cols[1:]:
            tot = self.g.add(tot, c)
        return Sym(self.g, [tot], self.win, not keepdim)

    def _reduce(self, scale):
        if self.scalar:
            raise Unsupported("a loss term made of variables only")
        if len(self) == 0:
            raise Unsupported("reduction over an empty slice")
        tot = self.cols[0]
        for c in self.cols[1:]:
            tot = self.g.add(tot, c)
        return Term(self.g, tot, self.win, scale)

    # -- torch.* functions ------------------------------------------------------------------------------------------
    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        name = _name_of(func).strip("_")
        if name.startswith("r") and name[1:] in ("add", "sub", "mul", "truediv", "div", "pow"):
            name = name[1:]            # Tensor.__rsub__(self, other) means other - self (`1.0 - C` with C a dde.Variable)
            args = (args[1], args[0]) + tuple(args[2:])
        if any(isinstance(a, VariableTensor) for a in args):
            g0 = next((a.g for a in args if isinstance(a, Sym)), None) or _TRACE.get("graph")
            args = tuple(_scalar_sym(g0, a) if isinstance(a, VariableTensor) else a for a in args)
        first = next((a for a in args if isinstance(a, Sym)), None)
        if first is None:
            for a in args:
                if isinstance(a, (list, tuple)):
                    first = next((b for b in a if isinstance(b, Sym)), first)
        if first is None:
            raise Unsupported(f"torch.{name} without a symbolic operand")
        if name in ("add", "sub", "subtract", "mul", "multiply", "div", "divide", "true_divide", "truediv"):
            if kwargs.get("alpha", 1) != 1:
                raise Unsupported("alpha")
            a, b = args[0], args[1]
            op = {"add": "add", "sub": "sub", "subtract": "sub", "mul": "mul", "multiply": "mul"}.get(name, "div")
            fn = getattr(first.g, op)
            return a._binary(b, fn) if isinstance(a, Sym) else b._binary(a, fn, swap=True)
        if name == "pow":
            a, b = args[0], args[1] if len(args) > 1 else kwargs.get("exponent")
            return a ** b if isinstance(a, Sym) else b.__rpow__(a)
        if name in ("sin", "cos", "exp", "tanh", "log", "sqrt", "abs", "absolute"):
            return args[0]._unary("abs" if name == "absolute" else name)
        if name in ("neg", "negative"):
            return -args[0]
        if name == "square":
            return args[0].square()
        if name == "zeros_like":
            return first._like([first.g.const(0.0)] * len(first.cols))
        if name == "ones_like":
            return first._like([first.g.const(1.0)] * len(first.cols))
        if name in ("cat", "concat", "concatenate", "hstack"):
            seq = args[0]
            dim = kwargs.get("dim", kwargs.get("axis", args[1] if len(args) > 1 else (1 if name == "hstack" else 0)))
            if dim not in (1, -1) or not all(isinstance(s, Sym) and not s.vec for s in seq):
                raise Unsupported("concatenation other than of symbolic columns")
            if any(s.win != first.win for s in seq):
                raise Unsupported("concatenation of different row ranges")
            return first._like([c for s in seq for c in s.cols])
        if name == "mean":
            return args[0].mean(*args[1:], **kwargs)
        if name == "sum":
            return args[0].sum(*args[1:], **kwargs)
        raise Unsupported(f"torch.{name}")


# ---- dde.grad on symbolic tensors (dde/gradients.py forwards here) ---------------------------------------------
def jacobian(ys, xs, i=0, j=None):
    if not (isinstance(xs, Sym) and xs.is_input):
        raise Unsupported("derivative with respect to something other than the network input")
    if ys.vec or ys.win != xs.win:
        raise Unsupported("jacobian of a sliced tensor")
    dim_y, dim_x = len(ys.cols), len(xs.cols)
    if not 0 <= i < dim_y:
        raise ValueError("i={} is not valid.".format(i))
    if j is not None and not 0 <= j < dim_x:
        raise ValueError("j={} is not valid.".format(j))
    g = ys.g
    if j is None or dim_x == 1:
        return Sym(g, [g.diff(ys.cols[i], d) for d in range(dim_x)], ys.win)
    return Sym(g, [g.diff(ys.cols[i], j)], ys.win)


def hessian(y, xs, component=None, i=0, j=0, grad_y=None):
    dim_y = len(y.cols)
    if dim_y > 1:
        if component is None:
            raise ValueError("The component of y is missing.")
        if component >= dim_y:
            raise ValueError("The component of y={} cannot be larger than the dimension={}.".format(component, dim_y))
    else:
        if component is not None:
            raise ValueError("Do not use component for 1D y.")
        component = 0
    dim_x = len(xs.cols)
    if not 0 <= i < dim_x:
        raise ValueError("i={} is not valid.".format(i))
    if not 0 <= j < dim_x:
        raise ValueError("j={} is not valid.".format(j))
    g = grad_y if grad_y is not None else jacobian(y, xs, i=component, j=None)
    if not isinstance(g, Sym):
        raise Unsupported("grad_y outside the trace")
    return jacobian(g, xs, i=i, j=j)


# ---- lowering -------------------------------------------------------------------------------------------------------
def _row_offsets(expr):
    """the set of row offsets an expression's loads use"""
    seen, out, stack = set(), set(), [expr]
    while stack:
        e = stack.pop()
        if e.uid in seen:
            continue
        seen.add(e.uid)
        if e.op in ("x", "net", "aux"):
            out.add(e.args[-1])
        stack.extend(a for a in e.args if isinstance(a, Node))
    return out


def _stream_of(spec, dims):
    """index of the kernel stream that carries d^|dims| / dx_dims of the network, None for a structural zero"""
    if not dims:
        return 0
    if len(dims) == 1:
        j = spec.first_of_dim.get(dims[0])
        if j is None:
            raise Unsupported(f"first derivative along input {dims[0]} is not carried")
        return 1 + j
    base = 1 + spec.n1
    if len(dims) == 2:
        if spec.trace is not None:
            # Laplacian pattern: the sum of the pure second derivatives rides as the FIRST of them, the others
            # read zero -- exactly what jet.JetOutput.second hands the autograd path
            if dims[0] == dims[1] and dims[0] in spec.first_of_dim:
                return base + spec.trace if dims[0] == spec.dirs[0] else None
            raise Unsupported("mixed second derivative under the Laplacian pattern")
        q = spec.second_of_dims.get((dims[0], dims[1]))
        if q is None:
            raise Unsupported(f"second derivative {dims} is not carried")
        return base + q
    if len(set(dims)) == 1 and spec.n1 == 1 and dims[0] == spec.dirs[0]:
        q = spec.third if len(dims) == 3 else spec.fourth if len(dims) == 4 else None
        if q is not None:
            return base + q
    if len(dims) == 3 and spec.triples and set(dims) <= set(spec.dirs):
        return base + spec.triples[sum(1 for d in dims if d == spec.dirs[1])]
    raise Unsupported(f"derivative {dims} is not carried")


class Program:
    """A lowered residual program.  `segments`: (first row, end row, code [m, 4] int32 = op, dst, a, b); terms are
    numbered in the order the traced code produced them; `scale[t]` turns a term's row sum into its loss."""

    def __init__(self, graph, spec, terms, n_rows, outputs=None):
        """`terms`: reduced loss terms (train step); `outputs`: point-wise columns over all rows to be STORED
        (model.predict(x, operator=pde): the residual probe of RAD / RAR_D / R3) -- one of the two."""
        if outputs is None and (not terms or len(terms) > MAX_TERMS):
            raise Unsupported(f"{len(terms)} loss terms")
        self.n_outputs = 0 if outputs is None else len(outputs)
        if outputs is not None and not 1 <= len(outputs) <= MAX_TERMS:
            raise Unsupported(f"{len(outputs)} operator outputs")
        self.n_terms, self.n_rows = len(terms), int(n_rows)
        self.S, self.C, self.d_in = spec.S, graph.n_out, graph.n_in
        self.scale = np.array([t.scale for t in terms], dtype=np.float64)   # uploaded as float32
        self.aux, self.scalars = list(graph.aux), graph.scalars
        self.scalar_grad = sorted(graph.scalar_grad)     # scalars whose gradient the reverse pass accumulates
        consts, cindex = [], {}
        by_win = {}
        aux_shifted = {}
        for t, term in enumerate(terms):
            by_win.setdefault(term.win, []).append((OP_SUMT, t, term.expr))
            # a term that reads other rows (PeriodicBC: row p against row p + m): the thread of row p cannot write the
            # adjoint of row p + m's streams, so row p + m gets a MIRROR segment -- the same expression re-centred on
            # it, contributing adjoints only (OP_SUMT_ADJ) -- and every thread keeps only the adjoints of its own row
            for delta in sorted(_row_offsets(term.expr) - {0}):
                mirror = (term.win[0] + delta, term.win[1] + delta)
                by_win.setdefault(mirror, []).append((OP_SUMT_ADJ, t, graph.shift(term.expr, -delta)))
        for c, expr in enumerate(outputs or ()):
            by_win.setdefault((0, int(n_rows)), []).append((OP_STORE, c, expr))
        self.segments, n_reg = [], 0
        for win in sorted(by_win):
            if win[0] < 0 or win[1] > n_rows or win[0] >= win[1]:
                raise Unsupported(f"term over rows {win} of {n_rows}")
            code, reg = [], {}

            def emit(node):
                if node.uid in reg:
                    return reg[node.uid]
                for a in node.args:
                    if isinstance(a, Node):
                        emit(a)
                r = len(reg)
                op, a = node.op, node.args
                if op in ("x", "net", "aux") and a[-1] != 0 and (win[0] + a[-1] < 0 or win[1] + a[-1] > n_rows):
                    raise Unsupported("row offset leaves the batch")
                if op == "x":
                    code.append((OP_LOADX, r, a[0], a[1]))
                elif op == "net":
                    s = _stream_of(spec, a[1])
                    if s is None:
                        code.append((OP_CONST, r, cindex.setdefault(0.0, len(cindex)), 0))
                    else:
                        code.append((OP_LOADJ, r, s * self.C + a[0], a[2]))
                elif op == "const":
                    v = float("nan") if a[0] == "nan" else a[0]
                    code.append((OP_CONST, r, cindex.setdefault(v, len(cindex)), 0))
                elif op == "aux":
                    k = a[0]
                    if a[2] != 0:            # the same array read a[2] rows away = an entry whose first row moved
                        key = (a[0], a[2])
                        if key not in aux_shifted:
                            self.aux.append((self.aux[a[0]][0], self.aux[a[0]][1] - a[2]))
                            aux_shifted[key] = len(self.aux) - 1
                        k = aux_shifted[key]
                    ten, first = self.aux[k]
                    if first > win[0] or first + ten.shape[0] < win[1]:
                        raise Unsupported("array does not cover the rows of the term it feeds")
                    code.append((OP_LOADA, r, k, a[1]))
                elif op == "scal":
                    code.append((OP_LOADS, r, a[0], 0))
                elif op in _BINARY:
                    code.append((_BINARY[op], r, reg[a[0].uid], reg[a[1].uid]))
                elif op == "pow":
                    code.append((OP_POW, r, reg[a[0].uid], cindex.setdefault(float(a[1]), len(cindex))))
                elif op in _UNARY:
                    code.append((_UNARY[op], r, reg[a[0].uid], 0))
                else:
                    raise Unsupported(f"operation {op}")
                reg[node.uid] = r
                return r

            limit = sys.getrecursionlimit()
            sys.setrecursionlimit(max(limit, 20000))
            try:
                for sink, t, expr in by_win[win]:
                    code.append((sink, t, emit(expr), 0))
            finally:
                sys.setrecursionlimit(limit)
            n_reg = max(n_reg, len(reg))
            # loads and constants first (they read no register; every value has a register of its own): the
            # interpreter issues them eight at a time instead of one dependent global load after the other
            code = [c for c in code if c[0] <= OP_CONST] + [c for c in code if c[0] > OP_CONST]
            self.segments.append((win[0], win[1], np.array(code, dtype=np.int32).reshape(-1, 4)))
        consts = [0.0] * len(cindex)
        for v, i in cindex.items():
            consts[i] = v
        self.consts = np.array(consts or [0.0], dtype=np.float64)     # uploaded as float32
        self.n_reg = max(n_reg, 1)
        self.n_instr = sum(len(c) for _, _, c in self.segments)
        if self.n_instr > MAX_INSTR or self.n_reg > MAX_REG or len(self.consts) > MAX_CONST:
            raise Unsupported(f"program too large ({self.n_instr} instructions, {self.n_reg} registers)")
        self._device = None

    def describe(self):
        what = f"{self.n_terms} terms" if not self.n_outputs else f"{self.n_outputs} output columns"
        return (f"{what}, {len(self.segments)} row segments, {self.n_instr} instructions, "
                f"{self.n_reg} registers, {len(self.aux)} arrays")


def trace(losses_fn, spec, n_rows, n_in, n_out, output_transform=None, feature_transform=None):
    """Run `losses_fn(x, y)` -- the model's own loss code -- on symbolic columns and lower what it did.
    `losses_fn` must return the list of per-term losses (each the torch.mean / torch.sum of point-wise algebra)."""
    g = Graph(n_in, n_out)
    x = Sym(g, [g.x(d) for d in range(n_in)], (0, n_rows))
    x.is_input = True
    y = Sym(g, [g.net(c) for c in range(n_out)], (0, n_rows))
    _TRACE["graph"] = g
    try:
        if feature_transform is not None:
            phi = feature_transform(x)
            if not isinstance(phi, Sym) or phi.vec or phi.scalar or phi.win != (0, n_rows):
                raise Unsupported("feature transform does not return one row per point")
            g.features = list(phi.cols)
        if output_transform is not None:
            y = output_transform(x, y)
            if not isinstance(y, Sym) or y.vec or y.scalar or y.win != (0, n_rows):
                raise Unsupported("output transform does not return one row per point")
        terms = losses_fn(x, y)
    except Unsupported:
        raise
    except (TypeError, AttributeError, ValueError, IndexError, RuntimeError, NotImplementedError) as e:
        raise Unsupported(f"{type(e).__name__}: {e}") from e
    finally:
        _TRACE.pop("graph", None)
    if not isinstance(terms, (list, tuple)):
        terms = [terms]
    for t in terms:
        if not isinstance(t, Term):
            raise Unsupported(f"loss term of type {type(t).__name__}")
    return Program(g, spec, list(terms), n_rows)


def trace_operator(operator_fn, spec, n_rows, n_in, n_out, output_transform=None, feature_transform=None):
    """`operator_fn(x, y)` -- the operator of model.predict(x, operator=...), usually the case's pde -- on symbolic
    columns.  Returns (program that stores every output column, column counts of the tensors it returned, whether
    it returned a list)."""
    g = Graph(n_in, n_out)
    x = Sym(g, [g.x(d) for d in range(n_in)], (0, n_rows))
    x.is_input = True
    y = Sym(g, [g.net(c) for c in range(n_out)], (0, n_rows))
    _TRACE["graph"] = g
    try:
        if feature_transform is not None:
            phi = feature_transform(x)
            if not isinstance(phi, Sym) or phi.vec or phi.scalar or phi.win != (0, n_rows):
                raise Unsupported("feature transform does not return one row per point")
            g.features = list(phi.cols)
        if output_transform is not None:
            y = output_transform(x, y)
            if not isinstance(y, Sym) or y.vec or y.scalar or y.win != (0, n_rows):
                raise Unsupported("output transform does not return one row per point")
        outs = operator_fn(x, y)
    except Unsupported:
        raise
    except (TypeError, AttributeError, ValueError, IndexError, RuntimeError, NotImplementedError) as e:
        raise Unsupported(f"{type(e).__name__}: {e}") from e
    finally:
        _TRACE.pop("graph", None)
    as_list = isinstance(outs, (list, tuple))
    outs = list(outs) if as_list else [outs]
    cols, widths = [], []
    for o in outs:
        if not isinstance(o, Sym) or o.vec or o.scalar or o.win != (0, n_rows):
            raise Unsupported("operator output is not one row per point")
        cols += o.cols
        widths.append(len(o.cols))
    return Program(g, spec, [], n_rows, outputs=cols), widths, as_list


def fingerprint(*functions):
    """The Python-level numbers a traced function can see without being re-run: closure cells, defaults, and the
    numeric attributes of the objects they hold (a case instance's `self.nu`, `self.k0` ...), one level deep.  A program
    freezes such numbers at trace time; the model re-traces when this tuple changes (a callback that anneals a PDE
    coefficient), which costs a few microseconds per step to check."""
    out = []

    def numbers(obj, depth):
        if isinstance(obj, (bool, numbers_mod.Number, np.generic)):
            out.append(float(obj) if not isinstance(obj, complex) else obj)
        elif isinstance(obj, (list, tuple)) and len(obj) <= 64:
            for v in obj:
                if isinstance(v, (numbers_mod.Number, np.generic)):
                    out.append(float(v) if not isinstance(v, complex) else v)
        elif depth > 0 and hasattr(obj, "__dict__") and not callable(obj) and not isinstance(obj, (torch.nn.Module, type)):
            for k in sorted(vars(obj)):
                numbers(vars(obj)[k], depth - 1)

    for fn in functions:
        if fn is None:
            continue
        target = getattr(fn, "__func__", fn)
        if getattr(fn, "__self__", None) is not None:
            numbers(fn.__self__, 1)
        for cell in getattr(target, "__closure__", None) or ():
            try:
                numbers(cell.cell_contents, 1)
            except ValueError:       # empty cell
                pass
        for v in getattr(target, "__defaults__", None) or ():
            numbers(v, 0)
    return tuple(out)
