"""dde.Model of the fused path: compile / train / predict with the seams the reference patches
(`outputs_losses_train`, `outputs_losses_test`, `train_step`; see
/root/reference/AI4XDE/algorithm/FBPINN.py:143-194) -- SURVEY.md 8a rows M1/M2, 8b.

What changes against DeepXDE: the collocation set lives on the device (re-uploaded only when the
host arrays change), the network evaluation inside every closure is the fused jet kernel sequence,
the derivative-stream set is discovered once by a kernel-free dry run of the case's own
`pde`/BC code, and the data-parallel gradient exchange (if torch.distributed is initialised) is one
all-reduce of the flat gradient.
"""
import os

import numpy as np
import torch

from .. import _lib, jet, residual
from . import config, gradients as grad, losses as losses_module, metrics as metrics_module
from . import optimizers, utils
from .callbacks import CallbackList


class TrainState:
    def __init__(self):
        self.epoch = self.step = 0
        self.X_train = self.y_train = self.train_aux_vars = None
        self.X_test = self.y_test = self.test_aux_vars = None
        self.y_pred_train = self.loss_train = self.loss_test = None
        self.y_pred_test = self.y_std_test = self.metrics_test = None
        self.best_step = 0
        self.best_loss_train = self.best_loss_test = np.inf
        self.best_y = self.best_ystd = self.best_metrics = None

    def set_data_train(self, X_train, y_train, train_aux_vars=None):
        self.X_train, self.y_train, self.train_aux_vars = X_train, y_train, train_aux_vars

    def set_data_test(self, X_test, y_test, test_aux_vars=None):
        self.X_test, self.y_test, self.test_aux_vars = X_test, y_test, test_aux_vars

    def update_best(self):
        if self.best_loss_train > np.sum(self.loss_train):
            self.best_step = self.step
            self.best_loss_train = np.sum(self.loss_train)
            self.best_loss_test = np.sum(self.loss_test)
            self.best_y, self.best_ystd = self.y_pred_test, self.y_std_test
            self.best_metrics = self.metrics_test

    def disregard_best(self):
        self.best_loss_train = np.inf


class LossHistory:
    def __init__(self):
        self.steps, self.loss_train, self.loss_test, self.metrics_test = [], [], [], []

    def append(self, step, loss_train, loss_test, metrics_test):
        self.steps.append(step)
        self.loss_train.append(loss_train)
        self.loss_test.append(self.loss_test[-1] if loss_test is None else loss_test)
        self.metrics_test.append(self.metrics_test[-1] if metrics_test is None else metrics_test)


_UPLOAD_STREAM = {}


def _upload_stream():
    dev = torch.cuda.current_device()
    if dev not in _UPLOAD_STREAM:
        _UPLOAD_STREAM[dev] = torch.cuda.Stream(device=dev)
    return _UPLOAD_STREAM[dev]


class _DeviceBatch:
    """Device-resident mirror of one host point array: AoS tensor for the case's `pde` code,
    SoA copy for the kernels."""

    __slots__ = ("host", "aos", "soa", "version", "_feat")

    def __init__(self, host, version=None):
        self.host = host
        self.version = version
        dev = config.device()
        pinned = torch.is_tensor(host) and host.device.type == "cpu" and host.is_pinned()
        if pinned and dev.type == "cuda" and not torch.cuda.is_current_stream_capturing():
            # a pinned host buffer handed in per step (bench.py's end-to-end arm, a user's own input pipeline): copy
            # and transpose on a side stream, so the upload of step k overlaps the kernels of step k - 1 that are
            # still queued on the compute stream; the compute stream waits for the copy, not the other way round
            main, side = torch.cuda.current_stream(), _upload_stream()
            with torch.cuda.stream(side):
                t = host.to(device=dev, dtype=torch.float32, non_blocking=True)
                soa = jet.to_soa(t)
            main.wait_stream(side)
            t.record_stream(main)
            soa.record_stream(main)
            self.aos, self.soa = t, soa
            return
        if torch.is_tensor(host):
            t = host.to(device=dev, dtype=torch.float32, non_blocking=True)
        else:
            t = torch.as_tensor(np.ascontiguousarray(host, dtype=np.float32)).to(dev, non_blocking=True)
        self.aos = t
        self.soa = jet.to_soa(t)

    def soa_for(self, net):
        """what the kernels evaluate `net` on: the points, or -- net.apply_feature_transform(f) -- f(points), computed
        once per batch; with a transform the pair (points, features) for the residual-program path"""
        f = getattr(net, "_input_transform", None)
        if f is None:
            return self.soa
        hit = getattr(self, "_feat", None)
        if hit is None or hit[0] is not f:
            with torch.no_grad():
                hit = (f, jet.to_soa(f(self.aos).to(torch.float32)))
            self._feat = hit
        return (self.soa, hit[1])

    @classmethod
    def on_device(cls, aos, soa=None, host=None, version=None):
        """A batch whose points were produced on the device (resampling rounds): nothing to upload."""
        b = cls.__new__(cls)
        b.host, b.version, b.aos = host, version, aos
        b.soa = soa if soa is not None else jet.to_soa(aos)
        return b


class Model:
    verbose = int(os.environ.get("AI4XDE_B200_VERBOSE", "0"))

    def __init__(self, data, net):
        self.data, self.net = data, net
        self.opt_name = None
        self.batch_size = None
        self.callbacks = None
        self.metrics = None
        self.external_trainable_variables = []
        self.train_state = TrainState()
        self.losshistory = LossHistory()
        self.stop_training = False
        self.opt = self.lr_scheduler = None
        self.loss_weights = None
        self.outputs = self.outputs_losses_train = self.outputs_losses_test = self.train_step = None
        self._batches = []      # small LRU of _DeviceBatch
        self._specs = {}        # stream sets by use
        self._residual = {}     # traced loss programs by use: (owners, key, tables or None, note)
        self._residual_epoch = 0
        self._weights_dev = None
        self.last_total_loss = None  # device scalar of the most recent closure evaluation

    # ---- device mirror -----------------------------------------------------------------------
    def _owns(self, inputs):
        """Is `inputs` one of the point arrays this framework builds and replaces as a whole (dde/data.py makes a
        NEW array and bumps `data.version` on every mutation)?  Only those are mirrored across calls."""
        d, ts = self.data, self.train_state
        for holder, names in ((d, ("train_x", "test_x", "train_x_all")), (ts, ("X_train", "X_test"))):
            for name in names:
                if getattr(holder, name, None) is inputs:
                    return True
        return False

    def device_batch(self, inputs):
        """Device mirror of a host point array.  Framework-owned arrays are cached by identity AND data version
        (true LRU, four entries, the entry keeps the host array alive so its id cannot be recycled); anything else
        -- an array a user hands to `predict`, which may be a buffer they refill in place -- is uploaded on every
        call, as DeepXDE does, and never displaces a cached training batch."""
        if not self._owns(inputs):
            return _DeviceBatch(inputs)
        version = getattr(self.data, "version", None)
        for i, b in enumerate(self._batches):
            if b.host is inputs and b.version == version:
                self._batches.append(self._batches.pop(i))
                return b
        self._batches = [b for b in self._batches if b.host is not inputs]
        b = _DeviceBatch(inputs, version)
        self._batches.append(b)
        if len(self._batches) > 4:
            self._batches.pop(0)
        return b

    def drop_device_batches(self):
        self._batches.clear()

    def adopt_device_batch(self, host, aos, soa=None):
        """Register device tensors that ALREADY hold `host` (a framework-owned point array assembled on the device by
        a resampling round) as its mirror, so the next step does not upload it again."""
        b = _DeviceBatch.__new__(_DeviceBatch)
        b.host, b.version = host, getattr(self.data, "version", None)
        b.aos = aos
        b.soa = soa if soa is not None else jet.to_soa(aos)
        self._batches = [c for c in self._batches if c.host is not host]
        self._batches.append(b)
        if len(self._batches) > 4:
            self._batches.pop(0)
        return b

    # ---- stream-set discovery ----------------------------------------------------------------
    def _n_out(self):
        return self.net.layer_sizes[-1]

    def spec_for_losses(self, batch, targets, loss_fn, losses_fn):
        # the entry keeps data, pde and every BC alive and is matched with `is`: an id() recycled after
        # `set_pde`, or a swapped BC list of the same length, must not serve a stale stream set
        key = ("losses",)
        owners = (self.data, self.data.pde, losses_fn.__func__ if hasattr(losses_fn, "__func__") else losses_fn) \
            + tuple(self.data.bcs)
        hit = self._specs.get(key)
        spec = None
        if hit is not None and len(hit[0]) == len(owners) and all(a is b for a, b in zip(hit[0], owners)):
            spec = hit[1]
        if spec is None:
            x = batch.aos.detach().requires_grad_()

            def fn(xx, y):
                if self.net._output_transform is not None:
                    y = self.net._output_transform(xx, y)
                out = losses_fn(targets, y, loss_fn, xx, self)
                grad.clear()
                return out

            spec = jet.probe_streams(x, self._n_out(), fn, feature=getattr(self.net, "_input_transform", None),
                                     variables=self.external_trainable_variables)
            self._specs[key] = (owners, spec)
            if Model.verbose:
                print(f"[ai4xde_b200] streams for training losses: {spec}")
        return spec

    def spec_for_operator(self, batch, operator):
        key = ("operator", id(operator))
        hit = self._specs.get(key)
        if hit is not None and hit[0] is operator:
            return hit[1]
        x = batch.aos.detach().requires_grad_()

        def fn(xx, y):
            if self.net._output_transform is not None:
                y = self.net._output_transform(xx, y)
            out = self._call_operator(operator, xx, y)
            grad.clear()
            return out

        spec = jet.probe_streams(x, self._n_out(), fn, feature=getattr(self.net, "_input_transform", None),
                                 variables=self.external_trainable_variables)
        self._specs[key] = (operator, spec)
        if Model.verbose:
            print(f"[ai4xde_b200] streams for operator {getattr(operator, '__name__', operator)}: {spec}")
        return spec

    def _call_operator(self, operator, x, y):
        nargs = utils.get_num_args(operator)
        if nargs == 2:
            return operator(x, y)
        if nargs == 3:
            return operator(x, y, self.net.auxiliary_vars)
        raise ValueError("operator must take 2 or 3 arguments")

    # ---- traced loss programs (ai4xde_b200/residual.py) -------------------------------------------
    def residual_program_note(self, use="train"):
        """Why the loss side of `use` ("train", "test", or "operator": the most recent operator program) runs -- or
        does not run -- as a traced program; for logs and tests."""
        if use == "operator":
            keys = [k for k in self._residual if isinstance(k, tuple) and k[0] == "operator"]
            return self._residual[keys[-1]][3] if keys else None
        hit = self._residual.get(use)
        return None if hit is None else hit[3]

    def _residual_tables(self, use, batch, spec, loss_fn, losses_fn, autograd_losses):
        """Device tables of the loss code traced into a residual program, or None when it stays on autograd.
        Traced once per (data, pde, BCs, point-set version, stream set, output transform, compile / train call);
        adopted only if, on this very batch, its per-term losses and the parameter gradient of their sum reproduce
        `autograd_losses()` -- the ordinary path over the jet leaves."""
        net = self.net
        from .nn import FNN, _FusedLosses

        if not config.residual_programs() or type(net) is not FNN \
                or net.regularizer is not None or net.auxiliary_vars is not None:
            return None
        owners = (self.data, self.data.pde, losses_fn.__func__ if hasattr(losses_fn, "__func__") else losses_fn,
                  net._output_transform, net._input_transform,
                  loss_fn if not isinstance(loss_fn, (list, tuple)) else tuple(loss_fn)) \
            + tuple(getattr(self.data, "bcs", ()))
        n_rows = int(batch.aos.shape[0])
        key = (getattr(self.data, "version", None), n_rows, spec.key(), self._residual_epoch, batch.host is self.data.train_x,
               residual.fingerprint(self.data.pde, net._output_transform, net._input_transform))
        hit = self._residual.get(use)
        if hit is not None and hit[1] == key and len(hit[0]) == len(owners) and all(a is b for a, b in zip(hit[0], owners)):
            return hit[2]
        tables, note = None, ""
        try:
            def fn(xs, ys):
                out = losses_fn(None, ys, loss_fn, xs, self)
                grad.clear()
                return out

            prog = residual.trace(fn, spec, n_rows, int(batch.aos.shape[1]), net.layer_sizes[-1],
                                  output_transform=net._output_transform, feature_transform=net._input_transform)
            tables = _lib.ResidualTables(prog, batch.aos.device)
            note = prog.describe()
            # the trainable scalars the program reads must be exactly variables this model optimises
            tables.variables = [prog.scalars[k] for k in prog.scalar_grad]
            for v in tables.variables:
                if not any(v is w for w in self.external_trainable_variables):
                    raise residual.Unsupported("a dde.Variable that is not in external_trainable_variables")
            if tables.variables:
                note += f", {len(tables.variables)} trainable variables"
        except (residual.Unsupported, _lib.NativeLibraryError) as e:
            tables, note = None, f"autograd path: {e}"
        finally:
            grad.clear()
        if tables is not None:
            # adoption check on the live batch: same losses, same parameter (and variable) gradient as the autograd path
            params = [p for p in net._param_list() if p.requires_grad] + list(tables.variables)
            with torch.enable_grad():
                ref = autograd_losses()
                g_ref = torch.autograd.grad(ref.sum(), params, allow_unused=True)
                got = _FusedLosses.apply(net, tables, batch.soa_for(net), n_rows, spec, config.engine(), True,
                                         len(net._param_list()), *net._param_list(), *tables.variables)
                g_got = torch.autograd.grad(got.sum(), params, allow_unused=True)
            scale = float(ref.detach().abs().max().clamp_min(1e-30))
            bad = float((got.detach() - ref.detach()).abs().max()) > 1e-4 * scale or got.shape != ref.shape
            for a, b in zip(g_got, g_ref):
                if (a is None) != (b is None):
                    bad = True
                elif a is not None:
                    bad = bad or float((a - b).abs().max()) > 1e-3 * float(b.abs().max().clamp_min(1e-30))
            if bad or not bool(torch.isfinite(got).all()):
                tables, note = None, f"autograd path: traced program ({note}) did not reproduce it on the live batch"
        self._residual[use] = (owners, key, tables, note)
        if Model.verbose:
            print(f"[ai4xde_b200] loss side of {use}: {note}")
        return tables

    def _operator_program(self, batch, spec, operator, autograd_eval):
        """(tables, column counts, list?) of `operator` traced into a program that STORES its outputs, or None.
        Same eligibility, tracing and adoption check (against `autograd_eval()`, the ordinary path, on this batch)
        as the train-step programs; one entry per (operator, number of rows)."""
        net = self.net
        from .nn import FNN

        if not config.residual_programs() or type(net) is not FNN \
                or net.auxiliary_vars is not None or utils.get_num_args(operator) != 2:
            return None
        n_rows = int(batch.aos.shape[0])
        use = ("operator", id(operator), n_rows)
        owners = (operator, net._output_transform, net._input_transform)
        key = (spec.key(), self._residual_epoch, residual.fingerprint(operator, net._output_transform, net._input_transform))
        hit = self._residual.get(use)
        if hit is not None and hit[1] == key and all(a is b for a, b in zip(hit[0], owners)):
            return hit[2]
        entry, note = None, ""
        try:
            prog, widths, as_list = residual.trace_operator(
                lambda xs, ys: operator(xs, ys), spec, n_rows, int(batch.aos.shape[1]), net.layer_sizes[-1],
                output_transform=net._output_transform, feature_transform=net._input_transform)
            entry = (_lib.ResidualTables(prog, batch.aos.device), widths, as_list)
            note = prog.describe()
        except residual.Unsupported as e:
            note = f"autograd path: {e}"
        finally:
            grad.clear()
        if entry is not None:
            ref = autograd_eval()
            ref = list(ref) if isinstance(ref, (list, tuple)) else [ref]
            got = self._run_operator_program(entry, batch, spec)
            got = list(got) if isinstance(got, (list, tuple)) else [got]
            bad = len(got) != len(ref)
            for a, b in zip(got, ref):
                bad = bad or a.shape != b.shape or not bool(torch.isfinite(a).all()) \
                    or float((a - b).abs().max()) > 1e-4 * float(b.abs().max().clamp_min(1e-30))
            if bad:
                entry, note = None, f"autograd path: traced operator ({note}) did not reproduce it on the live batch"
        for k in [k for k in self._residual if isinstance(k, tuple) and k[0] == "operator"][:-7]:
            del self._residual[k]            # a handful of (operator, size) entries is all an algorithm uses
        self._residual[use] = (owners, key, entry, note)
        if Model.verbose:
            print(f"[ai4xde_b200] operator {getattr(operator, '__name__', operator)} on {n_rows} rows: {note}")
        return entry

    def _run_operator_program(self, entry, batch, spec):
        from .nn import _launch_forward

        tables, widths, as_list = entry
        n = int(batch.aos.shape[0])
        with torch.no_grad():
            soa = batch.soa_for(self.net)
            jets = _launch_forward(self.net, soa[1] if isinstance(soa, tuple) else soa, n, spec, config.engine(), False)[0]
            out = _lib.residual_eval(tables, batch.soa, jets, n, sum(widths))
        cols, res = 0, []
        for w in widths:
            res.append(out[cols : cols + w, :n].t())       # (n, w) view; w = 1 for every reference pde
            cols += w
        return res if as_list else res[0]

    def _graph_fits(self, n_rows):
        """larger point sets: is the stash of this net small enough to live in a captured graph's private pool?"""
        hit, net = self._specs.get(("losses",)), self.net
        if hit is None or not hasattr(net, "net_desc"):
            return False
        try:
            descs = [net.net_desc] + [p.desc for p in (getattr(net, "_tc_pad", None), getattr(net, "_narrow_pad", None))
                                      if p is not None]
            worst = max(_lib.stash_bytes(d, hit[1].desc, _lib.pad32(n_rows), True, "fp32") for d in descs)
        except Exception:      # noqa: BLE001 -- an estimate only: no estimate, no capture
            return False
        return 2 * worst <= config.cuda_graph_max_bytes()      # sin nets on the tensor cores keep two stashes

    # ---- CUDA-graph replay of the small-problem step -------------------------------------------
    def _graphed_step(self, inputs, targets, auxiliary_vars, closure):
        """Launch-bound regime (cfg1 is ~60 MFLOP per step): after three eager
        steps on an unchanged point set the whole closure + Adam update is captured once into a CUDA
        graph and replayed.  Returns True if this call was served by the graph."""
        import torch.distributed as dist

        mode, max_rows = config.cuda_graphs()
        st = self.__dict__.setdefault("_graph_state", {"key": None, "seen": 0, "graph": None})
        if mode is False or not torch.cuda.is_available():
            return False
        n_rows = len(inputs) if hasattr(inputs, "__len__") else 0
        eligible = (
            (mode is True or n_rows <= max_rows or self._graph_fits(n_rows))
            and not optimizers.is_external_optimizer(self.opt_name)
            and self.lr_scheduler is None
            and targets is None and auxiliary_vars is None
            and (isinstance(self.opt, (torch.optim.Adam, torch.optim.AdamW)) or getattr(self.opt, "capturable", False))
            and self.opt.defaults.get("capturable", False)
            and not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1)
            and self.train_step.__qualname__.startswith("Model.compile")
            and self.outputs_losses_train.__qualname__.startswith("Model.compile")
        )
        # a captured step bakes in the device tables of the traced loss program: a new trace (every compile / train
        # call, every new point set) must never replay an old capture
        key = (id(inputs), getattr(self.data, "version", None), id(self.opt), id(self.data.pde), self._residual_epoch)
        eligible = eligible and self._owns(inputs)   # foreign arrays are re-uploaded per call: nothing stable to capture
        if not eligible:
            st.update(key=None, seen=0, graph=None, batch=None)
            return False
        if st["key"] != key:
            # the graph bakes in the device pointers of this batch: keep the batch itself (not only the host
            # array) alive for as long as the graph may be replayed
            st.update(key=key, seen=0, graph=None, holder=inputs, batch=self.device_batch(inputs))
        if st["graph"] is not None:
            st["graph"].replay()
            return True
        st["seen"] += 1
        if st["seen"] <= 3:
            return False                      # eager warm-up (also runs the stream-set probe, BC uploads)
        stepped = False
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self.opt.step(closure)        # THIS call's step, eagerly, on a side stream (capture warm-up)
            torch.cuda.current_stream().wait_stream(side)
            stepped = True
            g = torch.cuda.CUDAGraph()
            self.opt.zero_grad(set_to_none=True)
            with torch.cuda.graph(g):         # records the next steps; nothing executes here
                self.opt.step(closure)
            st["graph"] = g
            st["program"] = self._residual.get("train")     # the tables the captured kernels point at stay alive
            return True
        except Exception as e:                # capture is an optimisation: stay on the eager fused path
            import warnings

            warnings.warn(f"ai4xde_b200: CUDA-graph capture of the train step failed ({e}); running eagerly")
            config.set_cuda_graphs(False)
            torch.cuda.synchronize()
            return stepped      # the side-stream warm-up WAS this call's step: do not step twice

    # ---- data parallel -----------------------------------------------------------------------
    def _dp_term_scale(self, losses, inputs):
        """Per-term factor n_local * world / n_global that turns the all-reduce MEAN of the ranks'
        local means into the exact global mean (sums and global counts), cached per
        point-set version.  None for a single process or for batches that are not the training set."""
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return None
        data = self.data
        if inputs is not data.train_x or getattr(data, "num_bcs", None) is None:
            return None
        key = (getattr(data, "version", None), losses.numel())
        hit = getattr(self, "_dp_scale_cache", None)
        if hit is not None and hit[0] == key:
            return hit[1]
        n_bc = [int(c) for c in data.num_bcs]
        n_res = losses.numel() - len(n_bc)
        counts = [len(data.train_x) - sum(n_bc)] * n_res + n_bc
        local = torch.tensor(counts, dtype=torch.float64, device=losses.device)
        glob = local.clone()
        dist.all_reduce(glob, op=dist.ReduceOp.SUM)
        scale = (local * dist.get_world_size() / glob.clamp_min(1)).to(torch.float32)
        self._dp_scale_cache = (key, scale)
        return scale

    # ---- compile -----------------------------------------------------------------------------
    def compile(self, optimizer, lr=None, loss="MSE", metrics=None, decay=None, loss_weights=None,
                external_trainable_variables=None):
        self.opt_name = optimizer
        self._residual_epoch += 1
        loss_fn = losses_module.get(loss)
        self.loss_weights = loss_weights
        self._weights_dev = None
        if external_trainable_variables is None:
            self.external_trainable_variables = []
        else:
            if not isinstance(external_trainable_variables, list):
                external_trainable_variables = [external_trainable_variables]
            self.external_trainable_variables = external_trainable_variables

        def outputs(training, inputs):
            self.net.train(mode=training)
            batch = self.device_batch(inputs)
            with torch.no_grad():
                with self.net.jet_scope(jet.JetSpec.value_only(), batch.soa, param_grad=False):
                    return self.net(batch.aos)

        def outputs_losses(training, inputs, targets, auxiliary_vars, losses_fn, param_grad=True, fused=False,
                           raw=False):
            self.net.auxiliary_vars = None
            if auxiliary_vars is not None:
                self.net.auxiliary_vars = torch.as_tensor(auxiliary_vars, dtype=torch.float32, device=config.device())
            self.net.train(mode=training)
            batch = self.device_batch(inputs)
            if targets is not None and not getattr(self.data, "uses_targets", True):
                targets = None      # dde.data.PDE never reads them: no per-step upload of soln(train_x)
            if targets is not None:
                targets = torch.as_tensor(targets, dtype=torch.float32, device=config.device())
            spec = self.spec_for_losses(batch, targets, loss_fn, losses_fn)
            tables = None
            if fused and param_grad and targets is None and auxiliary_vars is None and torch.is_grad_enabled():
                # the caller wants the losses only (train closure): network + loss side as two fused stages
                tables = self._residual_tables(
                    "train" if training else "test", batch, spec, loss_fn, losses_fn,
                    lambda: outputs_losses(training, inputs, targets, auxiliary_vars, losses_fn, raw=True)[1])
            if tables is not None:
                from .nn import _FusedLosses

                params = self.net._param_list()
                need_bwd = any(p.requires_grad for p in params)
                outputs_ = None
                losses = _FusedLosses.apply(self.net, tables, batch.soa_for(self.net), int(batch.aos.shape[0]), spec,
                                            config.engine(), need_bwd, len(params), *params, *tables.variables)
            else:
                x = batch.aos.detach().requires_grad_()
                with self.net.jet_scope(spec, batch.soa, param_grad=param_grad):
                    outputs_ = self.net(x)
                losses = losses_fn(targets, outputs_, loss_fn, x, self)
                if not isinstance(losses, list):
                    losses = [losses]
                losses = torch.stack(losses)
            if raw:
                grad.clear()
                return outputs_, losses
            if self.loss_weights is not None:
                if self._weights_dev is None:
                    self._weights_dev = torch.as_tensor(self.loss_weights, dtype=torch.float32, device=losses.device)
                losses = losses * self._weights_dev
            scale = self._dp_term_scale(losses, inputs)
            if scale is not None:
                losses = losses * scale
            grad.clear()
            return outputs_, losses

        def outputs_losses_train(inputs, targets, auxiliary_vars):
            return outputs_losses(True, inputs, targets, auxiliary_vars, self.data.losses_train)

        def outputs_losses_test(inputs, targets, auxiliary_vars):
            return outputs_losses(False, inputs, targets, auxiliary_vars, self.data.losses_test)

        trainable = list(self.net.parameters()) + self.external_trainable_variables
        self.opt, self.lr_scheduler = optimizers.get(trainable, self.opt_name, learning_rate=lr, decay=decay)

        flat_grad = _bind_flat_grad(trainable)   # data parallel only: every p.grad a view of ONE buffer

        def train_step(inputs, targets, auxiliary_vars):
            def closure():
                losses = outputs_losses(True, inputs, targets, auxiliary_vars, self.data.losses_train, fused=True)[1]
                total_loss = torch.sum(losses)
                if flat_grad is not None and all(p.grad is not None and p.grad.data_ptr() == q for p, q in flat_grad[1]):
                    # point-sharded ranks: autograd accumulates into the views, ONE collective on the flat buffer
                    # (no torch.cat, no copy-back); the mean over ranks of the scaled local losses is the global loss
                    flat_grad[0].zero_()
                    total_loss.backward()
                    _allreduce_flat(flat_grad[0])
                else:
                    self.opt.zero_grad()
                    total_loss.backward()
                    _allreduce_grads(trainable)
                self.last_total_loss = total_loss.detach()
                return total_loss

            if self._graphed_step(inputs, targets, auxiliary_vars, closure):
                return
            if targets is None and auxiliary_vars is None and hasattr(inputs, "__len__"):
                # lets DeviceLBFGS decide whether its per-iteration sequence may be replayed from a CUDA graph
                closure.graph_rows = len(inputs) if self._owns(inputs) else None
                closure.graph_key = (id(inputs), getattr(self.data, "version", None), id(self.data.pde),
                                     self._residual_epoch)
                closure.graph_batch = self.device_batch(inputs)   # kept alive by the optimiser's graph
            scope = None
            if optimizers.is_external_optimizer(self.opt_name) and config.engine() == "auto":
                # second-order phase: FP32-accurate gradients where they are affordable (config.set_lbfgs_engine)
                mode, max_rows = config.lbfgs_engine()
                if mode == "fp32" and hasattr(inputs, "__len__") and len(inputs) <= max_rows:
                    scope = "fp32"
            with config.engine_scope(scope):
                self.opt.step(closure)
            if self.lr_scheduler is not None:
                self.lr_scheduler.step()

        self.outputs = outputs
        self._outputs_losses_eval = lambda training, i, t, a: outputs_losses(
            training, i, t, a, self.data.losses_train if training else self.data.losses_test, param_grad=False)
        self.outputs_losses_train = outputs_losses_train
        self.outputs_losses_test = outputs_losses_test
        self.train_step = train_step
        self.metrics = [metrics_module.get(m) for m in (metrics or [])]

    # ---- training ----------------------------------------------------------------------------
    def _outputs(self, training, inputs):
        return utils.to_numpy(self.outputs(training, inputs))

    def _outputs_losses(self, training, inputs, targets, auxiliary_vars):
        # evaluation only (no parameter gradient wanted): skip the per-layer stash unless the
        # user replaced the closures (FBPINN-style patching), in which case call theirs.
        fn = self.outputs_losses_train if training else self.outputs_losses_test
        own = getattr(fn, "__qualname__", "").startswith("Model.compile")
        outs = self._outputs_losses_eval(training, inputs, targets, auxiliary_vars) if own else fn(
            inputs, targets, auxiliary_vars)
        losses = outs[1]
        dist, world = _dist_world()
        if dist is not None and torch.is_tensor(losses) and losses.is_cuda:
            # point-sharded ranks: every rank logs the GLOBAL per-term losses (mean over ranks of the count-scaled
            # local terms), not its own share
            losses = losses.detach().clone()
            dist.all_reduce(losses, op=dist.ReduceOp.SUM)
            losses /= world
        return utils.to_numpy(outs[0]), utils.to_numpy(losses)

    def _train_step(self, inputs, targets, auxiliary_vars):
        self.train_step(inputs, targets, auxiliary_vars)

    def train(self, iterations=None, batch_size=None, display_every=1000, disregard_previous_best=False,
              callbacks=None, model_restore_path=None, model_save_path=None, epochs=None):
        if iterations is None and epochs is not None:
            iterations = epochs
        self.batch_size = batch_size
        self._residual_epoch += 1     # Python-level constants of the pde / BC code are re-read once per train() call
        self.callbacks = CallbackList(callbacks=callbacks)
        self.callbacks.set_model(self)
        if disregard_previous_best:
            self.train_state.disregard_best()
        if model_restore_path is not None:
            self.restore(model_restore_path)
        self.stop_training = False
        self.train_state.set_data_train(*self.data.train_next_batch(self.batch_size))
        self.train_state.set_data_test(*self.data.test())
        self._test()
        self.callbacks.on_train_begin()
        if optimizers.is_external_optimizer(self.opt_name):
            self._train_lbfgs()
        else:
            if iterations is None:
                raise ValueError("No iterations for {}.".format(self.opt_name))
            self._train_sgd(iterations, display_every)
        self.callbacks.on_train_end()
        if model_save_path is not None:
            self.save(model_save_path)
        return self.losshistory, self.train_state

    def _train_sgd(self, iterations, display_every):
        ts = self.train_state
        for i in range(iterations):
            self.callbacks.on_epoch_begin()
            self.callbacks.on_batch_begin()
            ts.set_data_train(*self.data.train_next_batch(self.batch_size))
            self._train_step(ts.X_train, ts.y_train, ts.train_aux_vars)
            ts.epoch += 1
            ts.step += 1
            if ts.step % display_every == 0 or i + 1 == iterations:
                self._test()
            self.callbacks.on_batch_end()
            self.callbacks.on_epoch_end()
            if self.stop_training:
                break

    def _train_lbfgs(self):
        ts = self.train_state
        prev = 0
        while prev < optimizers.LBFGS_options["maxiter"]:
            self.callbacks.on_epoch_begin()
            self.callbacks.on_batch_begin()
            ts.set_data_train(*self.data.train_next_batch(self.batch_size))
            self._train_step(ts.X_train, ts.y_train, ts.train_aux_vars)
            n_iter = self.opt.state_dict()["state"][0]["n_iter"]
            if prev == n_iter:
                break
            ts.epoch += n_iter - prev
            ts.step += n_iter - prev
            prev = n_iter
            self._test()
            self.callbacks.on_batch_end()
            self.callbacks.on_epoch_end()
            if self.stop_training:
                break

    def _test(self):
        ts = self.train_state
        ts.y_pred_train, ts.loss_train = self._outputs_losses(True, ts.X_train, ts.y_train, ts.train_aux_vars)
        ts.y_pred_test, ts.loss_test = self._outputs_losses(False, ts.X_test, ts.y_test, ts.test_aux_vars)
        if isinstance(ts.y_test, (list, tuple)):
            ts.metrics_test = [m(ts.y_test[i], ts.y_pred_test[i]) for m in self.metrics for i in range(len(ts.y_test))]
        else:
            ts.metrics_test = [m(ts.y_test, ts.y_pred_test) for m in self.metrics]
        ts.update_best()
        self.losshistory.append(ts.step, ts.loss_train, ts.loss_test, ts.metrics_test)
        if np.isnan(ts.loss_train).any() or np.isnan(ts.loss_test).any():
            self.stop_training = True
        if Model.verbose:
            print(ts.step, utils.list_to_str(ts.loss_train), utils.list_to_str(ts.loss_test),
                  utils.list_to_str(ts.metrics_test), flush=True)

    # ---- inference ---------------------------------------------------------------------------
    def predict(self, x, operator=None, callbacks=None):
        if isinstance(x, tuple):
            raise NotImplementedError("tuple inputs are outside the fused path")
        x = np.asarray(x, dtype=config.real(np))
        cbs = CallbackList(callbacks=callbacks)
        cbs.set_model(self)
        cbs.on_predict_begin()
        if operator is None:
            y = self._outputs(False, x)
        else:
            y = utils.to_numpy(self.predict_on_device(x, operator))
        cbs.on_predict_end()
        return y

    def predict_on_device(self, x, operator):
        """`operator(x, net(x))` evaluated with forward-only jets; returns device tensor(s).  `x`: a host point array
        or a _DeviceBatch that is already on the device (the on-device resampling rounds).  The operator runs as a
        traced residual program (one launch after the jet kernels) whenever its code can be traced."""
        self.net.eval()
        batch = x if isinstance(x, _DeviceBatch) else self.device_batch(x)
        spec = self.spec_for_operator(batch, operator)

        def autograd_eval():
            xt = batch.aos.detach().requires_grad_()
            with self.net.jet_scope(spec, batch.soa, param_grad=False):
                outputs = self.net(xt)
            y = self._call_operator(operator, xt, outputs)
            grad.clear()
            if isinstance(y, (list, tuple)):
                return [t.detach() for t in y]
            return y.detach()

        entry = self._operator_program(batch, spec, operator, autograd_eval)
        if entry is not None:
            return self._run_operator_program(entry, batch, spec)
        return autograd_eval()

    # ---- persistence -------------------------------------------------------------------------
    def state_dict(self):
        return self.net.state_dict()

    def save(self, save_path, protocol="backend", verbose=0):
        save_path = f"{save_path}-{self.train_state.epoch}.pt"
        torch.save({"model_state_dict": self.net.state_dict(),
                    "optimizer_state_dict": self.opt.state_dict()}, save_path)
        return save_path

    def restore(self, save_path, device=None, verbose=0):
        ckpt = torch.load(save_path, map_location=device or config.device())
        self.net.load_state_dict(ckpt["model_state_dict"])
        self.opt.load_state_dict(ckpt["optimizer_state_dict"])


def _dist_world():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist, dist.get_world_size()
    return None, 1


def _bind_flat_grad(params):
    """Data parallel: make every parameter's .grad a view of ONE float32 buffer, so that the gradient exchange is a
    single all-reduce of that buffer.  Returns (buffer, [(param, data_ptr of its view)]) or None (single process,
    or parameters that are not float32 tensors of one CUDA device)."""
    dist, world = _dist_world()
    params = [p for p in params if p.requires_grad]
    if dist is None or not params:
        return None
    dev = params[0].device
    if dev.type != "cuda" or any(p.device != dev or p.dtype != torch.float32 for p in params):
        return None
    flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
    views, off = [], 0
    for p in params:
        p.grad = flat[off:off + p.numel()].view_as(p)
        views.append((p, p.grad.data_ptr()))
        off += p.numel()
    return flat, views


def _allreduce_flat(flat):
    dist, world = _dist_world()
    if dist is None:
        return
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(world)


def _allreduce_grads(params):
    """Data-parallel exchange: one all-reduce (mean) of the flattened gradient, NCCL over NVLink
    when the process group is up.  No-op for a single process."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= dist.get_world_size()
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off : off + n].view_as(g))
        off += n


def saveplot(loss_history, train_state, issave=True, isplot=True, loss_fname="loss.dat",
             train_fname="train.dat", test_fname="test.dat", output_dir=None):
    """dde.saveplot as used at /root/reference/AI4XDE/solver/PDESolver.py:31 (issave only;
    plotting is out of scope and matplotlib is not required)."""
    output_dir = output_dir or os.getcwd()
    os.makedirs(output_dir, exist_ok=True)
    if not issave:
        return
    steps = np.array(loss_history.steps)[:, None]
    cols = [steps, np.array(loss_history.loss_train).reshape(len(steps), -1),
            np.array(loss_history.loss_test).reshape(len(steps), -1)]
    m = np.array(loss_history.metrics_test).reshape(len(steps), -1)
    if m.size:
        cols.append(m)
    np.savetxt(os.path.join(output_dir, loss_fname), np.hstack(cols), header="step, loss_train, loss_test, metrics_test")
    if train_state.X_train is not None and train_state.y_pred_train is not None:
        np.savetxt(os.path.join(output_dir, train_fname),
                   np.hstack((train_state.X_train, np.asarray(train_state.y_pred_train))), header="x, y_pred")
    if train_state.X_test is not None and train_state.best_y is not None:
        np.savetxt(os.path.join(output_dir, test_fname),
                   np.hstack((train_state.X_test, np.asarray(train_state.best_y))), header="x, y_pred")
