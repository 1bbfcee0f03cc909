"""Plot helpers are outside the hot path: these holders only keep
the axis metadata the cases pass in and import matplotlib lazily if someone really plots."""


class _Viz:
    def __init__(self, **kw):
        self.options = kw

    def _plt(self):
        try:
            import matplotlib.pyplot as plt
        except ImportError as e:  # pragma: no cover
            raise ImportError("plotting needs matplotlib, which is not part of the fused path") from e
        return plt


class Visualization_1D(_Viz):
    def set_axes_1D(self, title=None):
        plt = self._plt()
        fig, axes = plt.subplots()
        axes.set_xlabel(self.options.get("x_label", "x"))
        axes.set_ylabel(self.options.get("y_label", "y"))
        if title:
            axes.set_title(title)
        return axes

    def plot_line_1D(self, case, solver, exact=True, axes=None):
        axes = axes or self.set_axes_1D(title=case.name)
        X, y = case.get_testdata()
        axes.plot(X, solver.model.predict(X), "--", label=solver.name)
        if exact:
            axes.plot(X, y, label="Exact")
        axes.legend()
        return axes


class Visualization_2D(_Viz):
    def plot_exact_predict_error_2D(self, X, y, model_y, shape, title=None, colorbar=None):
        plt = self._plt()
        fig, axes = plt.subplots(1, 3, figsize=(15, 4))
        tf = self.options.get("feature_transform")
        Xp = tf(X) if tf else X
        for ax, val, name in zip(axes, (y, model_y, abs(y - model_y)), ("Exact", "Predict", "Error")):
            im = ax.pcolormesh(Xp[:, 0].reshape(shape), Xp[:, 1].reshape(shape), val.reshape(shape))
            ax.set_title(f"{title or ''} {name}")
            fig.colorbar(im, ax=ax)
        return fig, axes


def plot_loss_curves(label, losshistory, train_cost, axes=None, train=False, use_time=False):
    """Summed train / test loss versus step (or wall-clock seconds) on a log axis; matplotlib is an
    optional dependency of the fused path and is imported only here."""
    import numpy as np

    try:
        import matplotlib.pyplot as plt
    except ImportError as e:  # pragma: no cover
        raise ImportError("plotting needs matplotlib, which is not part of the fused path") from e
    fig, axes = plt.subplots() if axes is None else (axes.get_figure(), axes)
    steps = np.asarray(losshistory.steps, dtype=float)
    if use_time:
        xs, xlabel = steps / steps[-1] * train_cost, "Time(s)"
    else:
        xs, xlabel = steps, "Steps"
    curves = {"Test loss" if train else label: np.sum(losshistory.loss_test, axis=1)}
    if train:
        curves = {"Train loss": np.sum(losshistory.loss_train, axis=1), **curves}
    for name, ys in curves.items():
        axes.semilogy(xs, ys, label=name)
    axes.set(xlabel=xlabel, ylabel="Loss", title=label)
    axes.legend()
    return fig, axes
