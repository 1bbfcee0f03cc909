"""Case template of the fused path: the `PDECases` contract the solver and algorithms rely on
(attributes `.net .pde .geomtime .data .compile .NumDomain .loss_weights`; methods `set_pde`,
`set_loss_weights`, `get_testdata`, `output_transform`) -- mirrors
/root/reference/AI4XDE/cases/PDECases.py:8-101 on top of ai4xde_b200.dde."""
import os
from abc import ABC, abstractmethod

from .. import dde


def reference_data_path(name):
    """Reference-solution files (Burgers.npz, Allen_Cahn.mat, ...) ship with AI4XDE, not with this
    repo: point AI4XDE_DATA at an AI4XDE/data directory to evaluate L2 errors against them."""
    root = os.environ.get("AI4XDE_DATA")
    if root is None:
        raise FileNotFoundError(
            f"{name}: set AI4XDE_DATA to the AI4XDE/data directory holding the reference solutions"
        )
    return os.path.join(root, name)


class PDECases(ABC):
    def __init__(self, name, NumDomain=2000, use_output_transform=False, layer_size=[2] + [32] * 3 + [1],
                 activation="tanh", initializer="Glorot uniform", metrics=None, loss_weights=None,
                 external_trainable_variables=None, visualization=None):
        self.name = name
        self.NumDomain = NumDomain
        self.use_output_transform = use_output_transform
        self.metrics = metrics
        self.loss_weights = loss_weights
        self.external_trainable_variables = external_trainable_variables
        self.Visualization = visualization
        self.test_data = None
        self.net = self.gen_net(layer_size, activation, initializer)
        self.pde = self.gen_pde()
        self.geomtime = self.gen_geomtime()
        self.data = self.gen_data()
        self.compile = self.gen_compile()

    def gen_net(self, layer_size, activation, initializer):
        net = dde.nn.FNN(layer_size, activation, initializer)
        if self.use_output_transform:
            net.apply_output_transform(self.output_transform)
        return net

    @abstractmethod
    def gen_pde(self):
        ...

    @abstractmethod
    def gen_geomtime(self):
        ...

    @abstractmethod
    def gen_data(self):
        ...

    def gen_testdata(self):
        if callable(getattr(self, "sol", None)):
            x = self.geomtime.uniform_points(self.NumDomain)
            return x, self.sol(x)
        raise Warning("You must rewrite one of sol() and gen_testdata()")

    def get_testdata(self):
        if self.test_data is None:
            self.test_data = self.gen_testdata()
        return self.test_data

    def gen_compile(self):
        def compile(model, optimizer, lr=None, loss="MSE", decay=None):
            model.compile(optimizer, lr, loss, self.metrics, decay, self.loss_weights,
                          self.external_trainable_variables)

        return compile

    def set_loss_weights(self, loss_weights):
        self.loss_weights = loss_weights
        self.compile = self.gen_compile()

    def output_transform(self, x, y):
        pass

    def plot_result(self, solver):
        pass

    def set_pde(self, pde):
        self.pde = pde
        self.data = self.gen_data()
