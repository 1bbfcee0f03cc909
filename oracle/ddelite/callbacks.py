"""deepxde.callbacks subset: Callback, CallbackList, PDEPointResampler
(/root/reference/AI4XDE/algorithm/Random_R.py:14)."""


class Callback:
    def __init__(self):
        self.model = None

    def set_model(self, model):
        if model is not self.model:
            self.model = model
            self.init()

    def init(self):
        pass

    def on_epoch_begin(self):
        pass

    def on_epoch_end(self):
        pass

    def on_batch_begin(self):
        pass

    def on_batch_end(self):
        pass

    def on_train_begin(self):
        pass

    def on_train_end(self):
        pass

    def on_predict_begin(self):
        pass

    def on_predict_end(self):
        pass


class CallbackList(Callback):
    def __init__(self, callbacks=None):
        super().__init__()
        self.callbacks = list(callbacks or [])

    def set_model(self, model):
        self.model = model
        for c in self.callbacks:
            c.set_model(model)

    def _each(self, name):
        for c in self.callbacks:
            getattr(c, name)()

    def on_epoch_begin(self):
        self._each("on_epoch_begin")

    def on_epoch_end(self):
        self._each("on_epoch_end")

    def on_batch_begin(self):
        self._each("on_batch_begin")

    def on_batch_end(self):
        self._each("on_batch_end")

    def on_train_begin(self):
        self._each("on_train_begin")

    def on_train_end(self):
        self._each("on_train_end")

    def on_predict_begin(self):
        self._each("on_predict_begin")

    def on_predict_end(self):
        self._each("on_predict_end")

    def append(self, callback):
        self.callbacks.append(callback)


class PDEPointResampler(Callback):
    def __init__(self, period=100, pde_points=True, bc_points=False):
        super().__init__()
        self.period = period
        self.pde_points = pde_points
        self.bc_points = bc_points
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
