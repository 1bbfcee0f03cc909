"""deepxde.losses.get (looked up by /root/reference/AI4XDE/algorithm/FBPINN.py:138-141)."""
import torch


def mean_squared_error(y_true, y_pred):
    return torch.mean(torch.square(y_true - y_pred))


def mean_absolute_error(y_true, y_pred):
    return torch.mean(torch.abs(y_true - y_pred))


_BY_NAME = {
    "mean squared error": mean_squared_error, "MSE": mean_squared_error, "mse": mean_squared_error,
    "mean absolute error": mean_absolute_error, "MAE": mean_absolute_error, "mae": mean_absolute_error,
}


def get(identifier):
    if isinstance(identifier, (list, tuple)):
        return [get(i) for i in identifier]
    if callable(identifier):
        return identifier
    if identifier in _BY_NAME:
        return _BY_NAME[identifier]
    raise ValueError("Could not interpret loss function identifier:", identifier)
