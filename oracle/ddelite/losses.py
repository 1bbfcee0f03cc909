"""deepxde.losses.get (used at /root/reference/AI4XDE/algorithm/FBPINN.py:138-141)."""
import torch


def mean_squared_error(y_true, y_pred):
    return torch.mean(torch.square(y_true - y_pred))


def mean_absolute_error(y_true, y_pred):
    return torch.mean(torch.abs(y_true - y_pred))


def zero(*_):
    return torch.zeros(())


LOSS_DICT = {
    "mean squared error": mean_squared_error,
    "MSE": mean_squared_error,
    "mse": mean_squared_error,
    "mean absolute error": mean_absolute_error,
    "MAE": mean_absolute_error,
    "mae": mean_absolute_error,
    "zero": zero,
}


def get(identifier):
    if isinstance(identifier, (list, tuple)):
        return list(map(get, identifier))
    if isinstance(identifier, str):
        return LOSS_DICT[identifier]
    if callable(identifier):
        return identifier
    raise ValueError("Could not interpret loss function identifier:", identifier)
