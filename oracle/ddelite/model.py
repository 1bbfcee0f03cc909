"""dde.Model restated for the PyTorch backend.

The closure layout (outputs_losses_train/test, train_step) follows the copy of DeepXDE's
`_compile_pytorch` visible at /root/reference/AI4XDE/algorithm/FBPINN.py:143-188."""
import os

import numpy as np
import torch

from . import config, gradients as grad, losses as losses_module, metrics as metrics_module
from . import optimizers, utils
from .callbacks import CallbackList


class TrainState:
    def __init__(self):
        self.epoch = 0
        self.step = 0
        self.X_train = self.y_train = self.train_aux_vars = None
        self.X_test = self.y_test = self.test_aux_vars = None
        self.y_pred_train = None
        self.loss_train = self.loss_test = None
        self.y_pred_test = self.y_std_test = None
        self.metrics_test = None
        self.best_step = 0
        self.best_loss_train = np.inf
        self.best_loss_test = np.inf
        self.best_y = self.best_ystd = None
        self.best_metrics = None

    def set_data_train(self, X_train, y_train, train_aux_vars=None):
        self.X_train, self.y_train, self.train_aux_vars = X_train, y_train, train_aux_vars

    def set_data_test(self, X_test, y_test, test_aux_vars=None):
        self.X_test, self.y_test, self.test_aux_vars = X_test, y_test, test_aux_vars

    def update_best(self):
        if self.best_loss_train > np.sum(self.loss_train):
            self.best_step = self.step
            self.best_loss_train = np.sum(self.loss_train)
            self.best_loss_test = np.sum(self.loss_test)
            self.best_y = self.y_pred_test
            self.best_metrics = self.metrics_test

    def disregard_best(self):
        self.best_loss_train = np.inf


class LossHistory:
    def __init__(self):
        self.steps = []
        self.loss_train = []
        self.loss_test = []
        self.metrics_test = []

    def append(self, step, loss_train, loss_test, metrics_test):
        self.steps.append(step)
        self.loss_train.append(loss_train)
        if loss_test is None:
            loss_test = self.loss_test[-1]
        if metrics_test is None:
            metrics_test = self.metrics_test[-1]
        self.loss_test.append(loss_test)
        self.metrics_test.append(metrics_test)


class Model:
    verbose = 0

    def __init__(self, data, net):
        self.data = data
        self.net = net
        self.opt_name = None
        self.batch_size = None
        self.callbacks = None
        self.metrics = None
        self.external_trainable_variables = []
        self.train_state = TrainState()
        self.losshistory = LossHistory()
        self.stop_training = False
        self.opt = None
        self.lr_scheduler = None
        self.outputs = None
        self.outputs_losses_train = None
        self.outputs_losses_test = None
        self.train_step = None

    def compile(
        self,
        optimizer,
        lr=None,
        loss="MSE",
        metrics=None,
        decay=None,
        loss_weights=None,
        external_trainable_variables=None,
    ):
        self.opt_name = optimizer
        loss_fn = losses_module.get(loss)
        self.loss_weights = loss_weights
        if external_trainable_variables is None:
            self.external_trainable_variables = []
        else:
            if not isinstance(external_trainable_variables, list):
                external_trainable_variables = [external_trainable_variables]
            self.external_trainable_variables = external_trainable_variables

        def outputs(training, inputs):
            self.net.train(mode=training)
            with torch.no_grad():
                inputs = torch.as_tensor(inputs)
                inputs.requires_grad_()
            grad.clear()
            return self.net(inputs)

        def outputs_losses(training, inputs, targets, auxiliary_vars, losses_fn):
            self.net.auxiliary_vars = None
            if auxiliary_vars is not None:
                self.net.auxiliary_vars = torch.as_tensor(auxiliary_vars)
            self.net.train(mode=training)
            inputs = torch.as_tensor(inputs)
            inputs.requires_grad_()
            outputs_ = self.net(inputs)
            if targets is not None:
                targets = torch.as_tensor(targets)
            losses = losses_fn(targets, outputs_, loss_fn, inputs, self)
            if not isinstance(losses, list):
                losses = [losses]
            losses = torch.stack(losses)
            if loss_weights is not None:
                losses *= torch.as_tensor(loss_weights)
            grad.clear()
            return outputs_, losses

        def outputs_losses_train(inputs, targets, auxiliary_vars):
            return outputs_losses(True, inputs, targets, auxiliary_vars, self.data.losses_train)

        def outputs_losses_test(inputs, targets, auxiliary_vars):
            return outputs_losses(False, inputs, targets, auxiliary_vars, self.data.losses_test)

        trainable_variables = list(self.net.parameters()) + self.external_trainable_variables
        self.opt, self.lr_scheduler = optimizers.get(
            trainable_variables, self.opt_name, learning_rate=lr, decay=decay
        )

        def train_step(inputs, targets, auxiliary_vars):
            def closure():
                losses = outputs_losses_train(inputs, targets, auxiliary_vars)[1]
                total_loss = torch.sum(losses)
                self.opt.zero_grad()
                total_loss.backward()
                return total_loss

            self.opt.step(closure)
            if self.lr_scheduler is not None:
                self.lr_scheduler.step()

        self.outputs = outputs
        self.outputs_losses_train = outputs_losses_train
        self.outputs_losses_test = outputs_losses_test
        self.train_step = train_step
        metrics = metrics or []
        self.metrics = [metrics_module.get(m) for m in metrics]

    def _outputs(self, training, inputs):
        outs = self.outputs(training, inputs)
        return utils.to_numpy(outs)

    def _outputs_losses(self, training, inputs, targets, auxiliary_vars):
        if training:
            outs = self.outputs_losses_train(inputs, targets, auxiliary_vars)
        else:
            outs = self.outputs_losses_test(inputs, targets, auxiliary_vars)
        self.net.requires_grad_(requires_grad=False)
        self.net.requires_grad_()
        return utils.to_numpy(outs[0]), utils.to_numpy(outs[1])

    def _train_step(self, inputs, targets, auxiliary_vars):
        self.train_step(inputs, targets, auxiliary_vars)

    def train(
        self,
        iterations=None,
        batch_size=None,
        display_every=1000,
        disregard_previous_best=False,
        callbacks=None,
        model_restore_path=None,
        model_save_path=None,
        epochs=None,
    ):
        if iterations is None and epochs is not None:
            iterations = epochs
        self.batch_size = batch_size
        self.callbacks = CallbackList(callbacks=callbacks)
        self.callbacks.set_model(self)
        if disregard_previous_best:
            self.train_state.disregard_best()
        self.stop_training = False
        self.train_state.set_data_train(*self.data.train_next_batch(self.batch_size))
        self.train_state.set_data_test(*self.data.test())
        self._test()
        self.callbacks.on_train_begin()
        if optimizers.is_external_optimizer(self.opt_name):
            self._train_pytorch_external_optimizer(display_every)
        else:
            if iterations is None:
                raise ValueError("No iterations for {}.".format(self.opt_name))
            self._train_sgd(iterations, display_every)
        self.callbacks.on_train_end()
        if model_save_path is not None:
            self.save(model_save_path)
        return self.losshistory, self.train_state

    def _train_sgd(self, iterations, display_every):
        for i in range(iterations):
            self.callbacks.on_epoch_begin()
            self.callbacks.on_batch_begin()
            self.train_state.set_data_train(*self.data.train_next_batch(self.batch_size))
            self._train_step(
                self.train_state.X_train, self.train_state.y_train, self.train_state.train_aux_vars
            )
            self.train_state.epoch += 1
            self.train_state.step += 1
            if self.train_state.step % display_every == 0 or i + 1 == iterations:
                self._test()
            self.callbacks.on_batch_end()
            self.callbacks.on_epoch_end()
            if self.stop_training:
                break

    def _train_pytorch_external_optimizer(self, display_every):
        prev_n_iter = 0
        while prev_n_iter < optimizers.LBFGS_options["maxiter"]:
            self.callbacks.on_epoch_begin()
            self.callbacks.on_batch_begin()
            self.train_state.set_data_train(*self.data.train_next_batch(self.batch_size))
            self._train_step(
                self.train_state.X_train, self.train_state.y_train, self.train_state.train_aux_vars
            )
            n_iter = self.opt.state_dict()["state"][0]["n_iter"]
            if prev_n_iter == n_iter:
                break
            self.train_state.epoch += n_iter - prev_n_iter
            self.train_state.step += n_iter - prev_n_iter
            prev_n_iter = n_iter
            self._test()
            self.callbacks.on_batch_end()
            self.callbacks.on_epoch_end()
            if self.stop_training:
                break

    def _test(self):
        ts = self.train_state
        ts.y_pred_train, ts.loss_train = self._outputs_losses(
            True, ts.X_train, ts.y_train, ts.train_aux_vars
        )
        ts.y_pred_test, ts.loss_test = self._outputs_losses(
            False, ts.X_test, ts.y_test, ts.test_aux_vars
        )
        if isinstance(ts.y_test, (list, tuple)):
            ts.metrics_test = [
                m(ts.y_test[i], ts.y_pred_test[i]) for m in self.metrics for i in range(len(ts.y_test))
            ]
        else:
            ts.metrics_test = [m(ts.y_test, ts.y_pred_test) for m in self.metrics]
        ts.update_best()
        self.losshistory.append(ts.step, ts.loss_train, ts.loss_test, ts.metrics_test)
        if np.isnan(ts.loss_train).any() or np.isnan(ts.loss_test).any():
            self.stop_training = True
        if Model.verbose:
            print(ts.step, utils.list_to_str(ts.loss_train), utils.list_to_str(ts.loss_test),
                  utils.list_to_str(ts.metrics_test))

    def predict(self, x, operator=None, callbacks=None):
        if isinstance(x, tuple):
            x = tuple(np.asarray(xi, dtype=config.real(np)) for xi in x)
        else:
            x = np.asarray(x, dtype=config.real(np))
        callbacks = CallbackList(callbacks=callbacks)
        callbacks.set_model(self)
        callbacks.on_predict_begin()
        if operator is None:
            y = self._outputs(False, x)
            callbacks.on_predict_end()
            return y
        self.net.eval()
        inputs = torch.as_tensor(x)
        inputs.requires_grad_()
        outputs = self.net(inputs)
        if utils.get_num_args(operator) == 2:
            y = operator(inputs, outputs)
        elif utils.get_num_args(operator) == 3:
            y = operator(inputs, outputs, None)
        grad.clear()
        y = utils.to_numpy(y)
        callbacks.on_predict_end()
        return y

    def state_dict(self):
        return self.net.state_dict()

    def save(self, save_path, protocol="backend", verbose=0):
        save_path = f"{save_path}-{self.train_state.epoch}.pt"
        checkpoint = {
            "model_state_dict": self.net.state_dict(),
            "optimizer_state_dict": self.opt.state_dict(),
        }
        torch.save(checkpoint, save_path)
        return save_path

    def restore(self, save_path, device=None, verbose=0):
        checkpoint = torch.load(save_path, map_location=device)
        self.net.load_state_dict(checkpoint["model_state_dict"])
        self.opt.load_state_dict(checkpoint["optimizer_state_dict"])


def saveplot(loss_history, train_state, issave=True, isplot=True, loss_fname="loss.dat",
             train_fname="train.dat", test_fname="test.dat", output_dir=None):
    if output_dir is None:
        output_dir = os.getcwd()
    os.makedirs(output_dir, exist_ok=True)
    if issave:
        loss = np.hstack(
            (
                np.array(loss_history.steps)[:, None],
                np.array(loss_history.loss_train),
                np.array(loss_history.loss_test),
                np.array(loss_history.metrics_test).reshape(len(loss_history.steps), -1),
            )
        )
        np.savetxt(os.path.join(output_dir, loss_fname), loss, header="step, loss_train, loss_test, metrics_test")
