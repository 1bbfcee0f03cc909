"""Optimiser construction as DeepXDE's PyTorch backend does it:
Adam / L-BFGS over net.parameters() + external variables; `config.set_LBFGS_options`
(/root/reference/AI4XDE/algorithm/RAD.py:13)."""
import types

import torch

LBFGS_options = {}


def set_LBFGS_options(maxcor=100, ftol=0, gtol=1e-8, maxiter=15000, maxfun=None, maxls=50):
    LBFGS_options.update(
        maxcor=maxcor, ftol=ftol, gtol=gtol, maxiter=maxiter,
        maxfun=maxfun if maxfun is not None else int(maxiter * 1.25), maxls=maxls,
    )
    LBFGS_options["iter_per_step"] = min(1000, maxiter)
    LBFGS_options["fun_per_step"] = LBFGS_options["maxfun"] * LBFGS_options["iter_per_step"] // maxiter


set_LBFGS_options()
config = types.SimpleNamespace(set_LBFGS_options=set_LBFGS_options, LBFGS_options=LBFGS_options)


def lbfgs_factory(params, **kw):
    """The L-BFGS of the fused path: one kernel launch per iteration on the flat parameter vector, no
    per-iteration host sync (dde/lbfgs.py, csrc/lbfgs.cu).  There is no PyTorch / CPU fallback: DeviceLBFGS raises
    on parameters that are not CUDA tensors (the `not gpu` host-logic tests replace this factory together with the
    kernels by their plain-PyTorch doubles)."""
    from .lbfgs import DeviceLBFGS

    return DeviceLBFGS(params, **kw)


def is_external_optimizer(name):
    return name in ("L-BFGS", "L-BFGS-B")


def get(params, optimizer, learning_rate=None, decay=None, weight_decay=0):
    if isinstance(optimizer, torch.optim.Optimizer):
        optim = optimizer
    elif is_external_optimizer(optimizer):
        if weight_decay > 0:
            raise ValueError("L-BFGS optimizer doesn't support weight_decay > 0")
        kw = dict(lr=1, max_iter=LBFGS_options["iter_per_step"], max_eval=LBFGS_options["fun_per_step"],
                  tolerance_grad=LBFGS_options["gtol"], tolerance_change=LBFGS_options["ftol"],
                  history_size=LBFGS_options["maxcor"])
        optim = lbfgs_factory(list(params), **kw)
    else:
        if learning_rate is None:
            raise ValueError("No learning rate for {}.".format(optimizer))
        table = {"sgd": torch.optim.SGD, "rmsprop": torch.optim.RMSprop, "adam": torch.optim.Adam,
                 "adamw": torch.optim.AdamW}
        if optimizer not in table:
            raise NotImplementedError(f"{optimizer} to be implemented for backend pytorch.")
        params = list(params)
        kw = {}
        optim = None
        if optimizer == "adam" and params and all(p.is_cuda for p in params):
            from .adam import FlatAdam, flat_alias

            if flat_alias(params) is not None:    # dde.nn.FNN: one kernel launch per step on the flat buffer
                optim = FlatAdam(params, lr=learning_rate, weight_decay=weight_decay)
        if optimizer in ("adam", "adamw") and params and all(p.is_cuda for p in params):
            kw["capturable"] = True   # step counters on the device: the whole step can live in a CUDA graph
        if optim is None:
            optim = table[optimizer](params, lr=learning_rate, weight_decay=weight_decay, **kw)
    sched = None
    if decay is not None:
        if decay[0] == "step":
            sched = torch.optim.lr_scheduler.StepLR(optim, step_size=decay[1], gamma=decay[2])
        elif decay[0] == "inverse time":
            sched = torch.optim.lr_scheduler.LambdaLR(optim, lambda step: 1 / (1 + decay[2] * (step / decay[1])))
        else:
            raise NotImplementedError(f"{decay[0]} learning rate decay to be implemented for backend pytorch.")
    return optim, sched
