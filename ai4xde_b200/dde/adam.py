"""Adam on the flat parameter buffer: one kernel launch per step (csrc/adam.cu) instead of torch's multi-tensor
update, for parameter lists that alias ONE contiguous float32 CUDA buffer -- which is how dde.nn.FNN (and every net of a
NetBank) keeps its weights.  Same update rule and state names (`step`, `exp_avg`, `exp_avg_sq`) as
torch.optim.Adam, which DeepXDE's PyTorch backend builds for the first phase of every AI4XDE train_step
(/root/reference/AI4XDE/solver/PDESolver.py:79-81).  The gradients are views of one flat buffer too (bound here), so
`zero_grad` is one memset and the data-parallel exchange one all-reduce."""
import torch

from .. import _lib


def flat_alias(params):
    """One float32 tensor aliasing ALL parameters, or None if they are not consecutive views of one buffer."""
    ps = list(params)
    if not ps or ps[0].device.type != "cuda":
        return None
    base, off = ps[0].data_ptr(), 0
    for p in ps:
        if p.dtype != torch.float32 or not p.is_contiguous() or p.data_ptr() != base + 4 * off:
            return None
        off += p.numel()
    try:
        return torch.as_strided(ps[0].detach(), (off,), (1,), ps[0].storage_offset())
    except RuntimeError:
        return None


class FlatAdam(torch.optim.Optimizer):
    capturable = True   # the step counter lives on the device: the whole step can sit in a CUDA graph

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, capturable=True)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("FlatAdam takes one parameter group")
        self._params = self.param_groups[0]["params"]
        flat = flat_alias(self._params)
        if flat is None:
            raise _lib.NativeLibraryError("FlatAdam needs parameters that are consecutive float32 views of one CUDA buffer")
        n, dev = flat.numel(), flat.device
        self._grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self._views = []
        off = 0
        for p in self._params:
            p.grad = self._grad[off:off + p.numel()].view_as(p)
            self._views.append(p.grad.data_ptr())
            off += p.numel()
        st = self.state[self._params[0]]
        st["step_dev"] = torch.zeros(2, dtype=torch.int32, device=dev)
        st["exp_avg"] = torch.zeros(n, dtype=torch.float32, device=dev)
        st["exp_avg_sq"] = torch.zeros(n, dtype=torch.float32, device=dev)

    @property
    def flat_grad(self):
        return self._grad

    def _bound(self):
        return all(p.grad is not None and p.grad.data_ptr() == q for p, q in zip(self._params, self._views))

    def zero_grad(self, set_to_none=True):
        if self._bound():
            self._grad.zero_()
        else:       # someone replaced a .grad: rebind (values are dropped, as zero_grad means)
            off = 0
            self._grad.zero_()
            for i, p in enumerate(self._params):
                p.grad = self._grad[off:off + p.numel()].view_as(p)
                self._views[i] = p.grad.data_ptr()
                off += p.numel()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        g = self.param_groups[0]
        flat = flat_alias(self._params)
        if flat is None:
            raise _lib.NativeLibraryError("FlatAdam: the parameters no longer alias one buffer")
        if not self._bound():       # gradients landed in tensors of their own: gather them once
            torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in self._params],
                      out=self._grad)
        st = self.state[self._params[0]]
        _lib.adam_step(flat, self._grad, st["exp_avg"], st["exp_avg_sq"], st["step_dev"], g["lr"], g["betas"][0],
                       g["betas"][1], g["eps"], g["weight_decay"])
        return loss
