"""Training callbacks: the hook protocol dde.Model.train drives, and PDEPointResampler
(/root/reference/AI4XDE/algorithm/Random_R.py:14)."""

_HOOKS = ("on_epoch_begin", "on_epoch_end", "on_batch_begin", "on_batch_end", "on_train_begin",
          "on_train_end", "on_predict_begin", "on_predict_end")


class Callback:
    model = None

    def set_model(self, model):
        if model is not self.model:
            self.model = model
            self.init()

    def init(self):
        pass


for _h in _HOOKS:
    setattr(Callback, _h, lambda self: None)


class CallbackList(Callback):
    def __init__(self, callbacks=None):
        self.callbacks = list(callbacks or [])

    def set_model(self, model):
        self.model = model
        for cb in self.callbacks:
            cb.set_model(model)

    def append(self, callback):
        self.callbacks.append(callback)


def _fanout(name):
    def call(self):
        for cb in self.callbacks:
            getattr(cb, name)()

    return call


for _h in _HOOKS:
    setattr(CallbackList, _h, _fanout(_h))


class PDEPointResampler(Callback):
    """Resample the training points every `period` iterations."""

    def __init__(self, period=100, pde_points=True, bc_points=False):
        self.period, self.pde_points, self.bc_points = period, pde_points, bc_points
        self.num_bcs_initial = None
        self.epochs_since_last_resample = 0

    def on_train_begin(self):
        self.num_bcs_initial = self.model.data.num_bcs

    def on_epoch_end(self):
        self.epochs_since_last_resample += 1
        if self.epochs_since_last_resample < self.period:
            return
        self.epochs_since_last_resample = 0
        self.model.data.resample_train_points(self.pde_points, self.bc_points)
        if list(self.num_bcs_initial) != list(self.model.data.num_bcs):
            raise ValueError("`num_bcs` changed! Please update the loss function by `model.compile`.")
