"""deepxde.optimizers (PyTorch backend) restated: SURVEY.md App. B.5."""
import torch


class _Cfg:
    pass


LBFGS_options = {}


def set_LBFGS_options(maxcor=100, ftol=0, gtol=1e-8, maxiter=15000, maxfun=None, maxls=50):
    LBFGS_options["maxcor"] = maxcor
    LBFGS_options["ftol"] = ftol
    LBFGS_options["gtol"] = gtol
    LBFGS_options["maxiter"] = maxiter
    LBFGS_options["maxfun"] = maxfun if maxfun is not None else int(maxiter * 1.25)
    LBFGS_options["maxls"] = maxls
    LBFGS_options["iter_per_step"] = min(1000, maxiter)
    LBFGS_options["fun_per_step"] = LBFGS_options["maxfun"] * LBFGS_options["iter_per_step"] // maxiter


set_LBFGS_options()

config = _Cfg()
config.set_LBFGS_options = set_LBFGS_options
config.LBFGS_options = LBFGS_options


def is_external_optimizer(optimizer):
    return optimizer in ["L-BFGS", "L-BFGS-B"]


def get(params, optimizer, learning_rate=None, decay=None, weight_decay=0):
    if isinstance(optimizer, torch.optim.Optimizer):
        optim = optimizer
    elif optimizer in ["L-BFGS", "L-BFGS-B"]:
        if weight_decay > 0:
            raise ValueError("L-BFGS optimizer doesn't support weight_decay > 0")
        if learning_rate is not None or decay is not None:
            print("Warning: learning rate is ignored for {}".format(optimizer))
        optim = torch.optim.LBFGS(
            params,
            lr=1,
            max_iter=LBFGS_options["iter_per_step"],
            max_eval=LBFGS_options["fun_per_step"],
            tolerance_grad=LBFGS_options["gtol"],
            tolerance_change=LBFGS_options["ftol"],
            history_size=LBFGS_options["maxcor"],
            line_search_fn=None,
        )
    else:
        if learning_rate is None:
            raise ValueError("No learning rate for {}.".format(optimizer))
        if optimizer == "sgd":
            optim = torch.optim.SGD(params, lr=learning_rate, weight_decay=weight_decay)
        elif optimizer == "rmsprop":
            optim = torch.optim.RMSprop(params, lr=learning_rate, weight_decay=weight_decay)
        elif optimizer == "adam":
            optim = torch.optim.Adam(params, lr=learning_rate, weight_decay=weight_decay)
        elif optimizer == "adamw":
            optim = torch.optim.AdamW(params, lr=learning_rate, weight_decay=weight_decay)
        else:
            raise NotImplementedError(f"{optimizer} to be implemented for backend pytorch.")
    lr_scheduler = _get_lr_scheduler(optim, decay)
    return optim, lr_scheduler


def _get_lr_scheduler(optim, decay):
    if decay is None:
        return None
    if decay[0] == "step":
        return torch.optim.lr_scheduler.StepLR(optim, step_size=decay[1], gamma=decay[2])
    if decay[0] == "inverse time":
        return torch.optim.lr_scheduler.LambdaLR(
            optim, lambda step: 1 / (1 + decay[2] * (step / decay[1]))
        )
    raise NotImplementedError(f"{decay[0]} learning rate decay to be implemented for backend pytorch.")
