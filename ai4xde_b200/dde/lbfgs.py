"""Device-resident L-BFGS for the second phase of every AI4XDE `train_step`
(/root/reference/AI4XDE/solver/PDESolver.py:79-85 -> `model.compile("L-BFGS"); model.train()`).

Drop-in for `torch.optim.LBFGS(lr, max_iter, max_eval, tolerance_grad, tolerance_change, history_size,
line_search_fn=None)` as DeepXDE's PyTorch backend builds it: same update rule and stopping tests
(ai4xde_b200/csrc/lbfgs.cu follows torch's `step` statement by statement), but one kernel launch performs
a whole iteration on the flat parameter vector and every decision is taken on the device.  The host reads
the 48-byte state block back only every `sync_every` iterations, and on small (launch-bound) problems the
sequence closure -> gradient gather -> iteration kernel is captured once into a CUDA graph and replayed.
`state_dict()["state"][0]["n_iter"]` keeps working -- it is what dde.Model's L-BFGS loop watches."""
import torch

from .. import _lib
from . import config


class DeviceLBFGS(torch.optim.Optimizer):
    def __init__(self, params, lr=1.0, max_iter=20, max_eval=None, tolerance_grad=1e-7, tolerance_change=1e-9,
                 history_size=100, sync_every=25):
        if max_eval is None:
            max_eval = max_iter * 5 // 4
        defaults = dict(lr=lr, max_iter=max_iter, max_eval=max_eval, tolerance_grad=tolerance_grad,
                        tolerance_change=tolerance_change, history_size=history_size)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("LBFGS doesn't support per-parameter options (parameter groups)")
        self._params = self.param_groups[0]["params"]
        dev = self._params[0].device
        if dev.type != "cuda":
            raise _lib.NativeLibraryError("DeviceLBFGS needs CUDA parameters (there is no CPU fallback)")
        self._P = sum(p.numel() for p in self._params)
        self._sync_every = int(sync_every)
        self._state_dev = torch.zeros(12, dtype=torch.int32, device=dev)       # ax_lbfgs_state
        self._ws = torch.zeros(_lib.lbfgs_ws_floats(self._P, history_size), dtype=torch.float32, device=dev)
        self._gather = torch.empty(self._P, dtype=torch.float32, device=dev)   # x when the parameters are scattered
        self._graph = None
        self._graph_key = None
        self._seen = 0
        self.state[self._params[0]].update(n_iter=0, func_evals=0)

    # ---- flat views ----------------------------------------------------------------------------------
    def _flat_alias(self):
        """One float32 tensor aliasing ALL parameters (dde.nn.FNN keeps its weights as consecutive views of one
        master buffer), or None if they are scattered."""
        ps = self._params
        base = ps[0].data_ptr()
        off = 0
        for p in ps:
            if p.dtype != torch.float32 or not p.is_contiguous() or p.data_ptr() != base + 4 * off:
                return None
            off += p.numel()
        try:
            return torch.as_strided(ps[0].detach(), (self._P,), (1,), ps[0].storage_offset())
        except RuntimeError:
            return None

    def _flat_grad(self):
        return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).to(torch.float32)
                          for p in self._params])

    def _iterate(self, closure, begin):
        import torch.distributed as dist

        g = self.param_groups[0]
        loss = closure()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            # point-sharded ranks: the closure's loss is this rank's share (terms scaled so that the MEAN over ranks
            # is the global loss, dde/model.py:_dp_term_scale).  Every stopping test of the iteration kernel that
            # looks at the loss must see the same number on every rank, or one rank stops updating alone.
            loss = loss.detach().clone()
            dist.all_reduce(loss, op=dist.ReduceOp.SUM)
            loss /= dist.get_world_size()
        flat_g = self._flat_grad()
        alias = self._flat_alias()
        x = alias
        if x is None:
            torch.cat([p.detach().reshape(-1).to(torch.float32) for p in self._params], out=self._gather)
            x = self._gather
        _lib.lbfgs_iterate(x, flat_g, loss.detach().reshape(1).to(torch.float32), self._ws, self._state_dev,
                           g["history_size"], g["lr"], g["tolerance_grad"], g["tolerance_change"], begin)
        if alias is None:
            off = 0
            for p in self._params:
                p.data.copy_(x[off:off + p.numel()].view_as(p))
                off += p.numel()
        return loss

    def _read_state(self):
        st = self._state_dev.cpu()
        rec = self.state[self._params[0]]
        rec["n_iter"], rec["func_evals"] = int(st[0]), int(st[4])
        return bool(int(st[3]))

    # ---- optional CUDA-graph replay of one iteration (launch-bound problems) ---------------------------
    def _graph_ok(self, closure):
        import torch.distributed as dist

        mode, _ = config.cuda_graphs()
        if mode is False or getattr(closure, "graph_rows", None) is None:
            return False
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return False
        return (mode is True or closure.graph_rows <= config.cuda_graphs()[1]) and self._flat_alias() is not None

    @torch.no_grad()
    def step(self, closure):
        closure = torch.enable_grad()(closure)
        g = self.param_groups[0]
        max_iter, max_eval = g["max_iter"], g["max_eval"]
        key = getattr(closure, "graph_key", None)
        use_graph = self._graph_ok(closure)
        if not use_graph or key != self._graph_key:
            self._graph, self._graph_key, self._seen = None, key, 0
        # a captured iteration replays kernels on this batch's device pointers: own it while the graph lives
        self._graph_batch = getattr(closure, "graph_batch", None) if use_graph else None
        # first evaluation of this call, eagerly; detached at once -- a live autograd graph would pin its
        # AccumulateGrad nodes to this stream and invalidate a later capture
        first = self._iterate(closure, 1).detach()
        it = 1
        done = False
        while it < max_iter and it < max_eval:
            if use_graph and self._graph is None and self._seen >= 3:
                try:
                    side = torch.cuda.Stream()
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        self._iterate(closure, 0)             # capture warm-up on a side stream (a real iteration)
                    torch.cuda.current_stream().wait_stream(side)
                    it += 1
                    graph = torch.cuda.CUDAGraph()
                    for p in self._params:
                        p.grad = None
                    with torch.cuda.graph(graph):
                        self._iterate(closure, 0)
                    self._graph = graph
                    continue
                except Exception as e:                        # capture is an optimisation only
                    import warnings

                    warnings.warn(f"ai4xde_b200: CUDA-graph capture of the L-BFGS iteration failed ({e}); running eagerly")
                    use_graph = False
                    torch.cuda.synchronize()
            if self._graph is not None:
                self._graph.replay()
            else:
                del_me = self._iterate(closure, 0)
                del del_me
                self._seen += 1
            it += 1
            if it % self._sync_every == 0:
                done = self._read_state()
                if done:
                    break
        if not done:
            self._read_state()
        return first
