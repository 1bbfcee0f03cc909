"""deepxde.metrics subset (l2_relative_error: /root/reference/AI4XDE/solver/PDESolver.py:75)."""
import numpy as np


def l2_relative_error(y_true, y_pred):
    return np.linalg.norm(y_true - y_pred) / np.linalg.norm(y_true)


def nanl2_relative_error(y_true, y_pred):
    err = np.nan_to_num(y_true - y_pred)
    return np.linalg.norm(err) / np.linalg.norm(np.nan_to_num(y_true))


def mean_l2_relative_error(y_true, y_pred):
    return np.mean(np.linalg.norm(y_true - y_pred, axis=1) / np.linalg.norm(y_true, axis=1))


def mean_squared_error(y_true, y_pred):
    return np.mean((y_true - y_pred) ** 2)


def accuracy(y_true, y_pred):
    return np.mean(np.equal(np.argmax(y_pred, axis=-1), np.argmax(y_true, axis=-1)))


_BY_NAME = {
    "l2 relative error": l2_relative_error, "nanl2 relative error": nanl2_relative_error,
    "mean l2 relative error": mean_l2_relative_error, "mean squared error": mean_squared_error,
    "MSE": mean_squared_error, "mse": mean_squared_error, "accuracy": accuracy,
}


def get(identifier):
    if callable(identifier):
        return identifier
    if identifier in _BY_NAME:
        return _BY_NAME[identifier]
    raise ValueError("Could not interpret metric function identifier:", identifier)
