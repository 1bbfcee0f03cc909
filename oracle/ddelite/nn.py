"""dde.nn.FNN restated: torch.nn.Linear stack, Glorot init,
zero bias, optional feature/output transforms applied inside forward()."""
import torch

from . import config

_ACT = {
    "tanh": torch.tanh,
    "sin": torch.sin,
    "sigmoid": torch.sigmoid,
    "relu": torch.relu,
    "elu": torch.nn.functional.elu,
    "selu": torch.nn.functional.selu,
    "silu": torch.nn.functional.silu,
    "swish": torch.nn.functional.silu,
    "gelu": torch.nn.functional.gelu,
    "linear": lambda x: x,
    None: lambda x: x,
}

_INIT = {
    "Glorot normal": torch.nn.init.xavier_normal_,
    "Glorot uniform": torch.nn.init.xavier_uniform_,
    "He normal": torch.nn.init.kaiming_normal_,
    "He uniform": torch.nn.init.kaiming_uniform_,
    "zeros": torch.nn.init.zeros_,
}


def get_activation(identifier):
    if callable(identifier):
        return identifier
    return _ACT[identifier]


def get_initializer(identifier):
    return _INIT[identifier]


class NN(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.regularizer = None
        self._auxiliary_vars = None
        self._input_transform = None
        self._output_transform = None

    @property
    def auxiliary_vars(self):
        return self._auxiliary_vars

    @auxiliary_vars.setter
    def auxiliary_vars(self, value):
        self._auxiliary_vars = value

    def apply_feature_transform(self, transform):
        self._input_transform = transform

    def apply_output_transform(self, transform):
        self._output_transform = transform

    def num_trainable_parameters(self):
        return sum(v.numel() for v in self.parameters() if v.requires_grad)


class FNN(NN):
    def __init__(self, layer_sizes, activation, kernel_initializer):
        super().__init__()
        if isinstance(activation, list):
            if not (len(layer_sizes) - 1) == len(activation):
                raise ValueError("Total number of activation functions do not match with sum of hidden layers and output layer!")
            self.activation = list(map(get_activation, activation))
        else:
            self.activation = get_activation(activation)
        initializer = get_initializer(kernel_initializer)
        initializer_zero = get_initializer("zeros")
        self.linears = torch.nn.ModuleList()
        for i in range(1, len(layer_sizes)):
            self.linears.append(
                torch.nn.Linear(layer_sizes[i - 1], layer_sizes[i], dtype=config.real(torch))
            )
            initializer(self.linears[-1].weight)
            initializer_zero(self.linears[-1].bias)

    def forward(self, inputs):
        x = inputs
        if self._input_transform is not None:
            x = self._input_transform(x)
        for j, linear in enumerate(self.linears[:-1]):
            x = (
                self.activation[j](linear(x))
                if isinstance(self.activation, list)
                else self.activation(linear(x))
            )
        x = self.linears[-1](x)
        if self._output_transform is not None:
            x = self._output_transform(inputs, x)
        return x



class PFNN(NN):
    """Parallel fully-connected network [DeepXDE-upstream `deepxde.nn.pytorch.PFNN`], restated: a layer size
    given as a list of n_output widths splits the net into one sub-network per output from that layer on (a
    plain int before the first split is a shared layer; re-joining after a split is an error).  Constructed
    by the reference at /root/reference/AI4XDE/cases/PoissonCases.py:625-626 with
    [1, [20, 20], [20, 20], [20, 20], 2].  Parameter names follow upstream: `layers.{i}.{j}.weight`."""

    def __init__(self, layer_sizes, activation, kernel_initializer):
        super().__init__()
        self.activation = get_activation(activation)
        initializer = get_initializer(kernel_initializer)
        initializer_zero = get_initializer("zeros")
        if len(layer_sizes) <= 1:
            raise ValueError("must specify input and output sizes")
        if not isinstance(layer_sizes[0], int) or not isinstance(layer_sizes[-1], int):
            raise ValueError("input and output sizes must be integers")
        n_output = layer_sizes[-1]

        def make_linear(n_in, n_out):
            lin = torch.nn.Linear(n_in, n_out, dtype=config.real(torch))
            initializer(lin.weight)
            initializer_zero(lin.bias)
            return lin

        self.layers = torch.nn.ModuleList()
        for i in range(1, len(layer_sizes) - 1):
            prev, cur = layer_sizes[i - 1], layer_sizes[i]
            if isinstance(cur, (list, tuple)):
                if len(cur) != n_output:
                    raise ValueError("number of sub-layers should equal number of network outputs")
                if isinstance(prev, (list, tuple)):
                    self.layers.append(torch.nn.ModuleList([make_linear(prev[j], cur[j]) for j in range(n_output)]))
                else:
                    self.layers.append(torch.nn.ModuleList([make_linear(prev, cur[j]) for j in range(n_output)]))
            else:
                if not isinstance(prev, int):
                    raise ValueError("cannot rejoin parallel subnetworks after splitting")
                self.layers.append(make_linear(prev, cur))
        if isinstance(layer_sizes[-2], (list, tuple)):
            self.layers.append(torch.nn.ModuleList([make_linear(layer_sizes[-2][j], 1) for j in range(n_output)]))
        else:
            self.layers.append(make_linear(layer_sizes[-2], n_output))

    def forward(self, inputs):
        x = inputs
        if self._input_transform is not None:
            x = self._input_transform(x)
        for layer in self.layers[:-1]:
            if isinstance(layer, torch.nn.ModuleList):
                if isinstance(x, list):
                    x = [self.activation(f(x_)) for f, x_ in zip(layer, x)]
                else:
                    x = [self.activation(f(x)) for f in layer]
            else:
                x = self.activation(layer(x))
        if isinstance(x, list):
            x = torch.cat([f(x_) for f, x_ in zip(self.layers[-1], x)], dim=1)
        else:
            x = self.layers[-1](x)
        if self._output_transform is not None:
            x = self._output_transform(inputs, x)
        return x



class MsFFN(NN):
    """Multi-scale Fourier feature network [DeepXDE-upstream `deepxde.nn.pytorch.MsFFN`], restated: for each
    sigma a FIXED random matrix b ~ N(0, sigma^2) of shape (d_in, layer_sizes[1] // 2) embeds the input as
    [cos(x b), sin(x b)]; every embedding runs through the SAME hidden Linear + activation stack
    (`linears`, sizes layer_sizes[2:-1]); the results are concatenated and mapped by `_dense`.  The `b` are
    plain tensors (not parameters, not in the state dict), as upstream.  Constructed by the reference at
    /root/reference/AI4XDE/cases/PoissonCases.py:478-479 with sigmas [1, 10]."""

    def __init__(self, layer_sizes, activation, kernel_initializer, sigmas, dropout_rate=0):
        super().__init__()
        self.activation = get_activation(activation)
        self.dropout_rate = dropout_rate
        self.sigmas = list(sigmas)
        initializer = get_initializer(kernel_initializer)
        initializer_zero = get_initializer("zeros")
        self.b = [torch.normal(mean=0.0, std=float(sigma), size=(layer_sizes[0], layer_sizes[1] // 2)).to(config.real(torch))
                  for sigma in self.sigmas]
        self.linears = torch.nn.ModuleList()
        for i in range(2, len(layer_sizes) - 1):
            self.linears.append(torch.nn.Linear(layer_sizes[i - 1], layer_sizes[i], dtype=config.real(torch)))
            initializer(self.linears[-1].weight)
            initializer_zero(self.linears[-1].bias)
        self._dense = torch.nn.Linear(layer_sizes[-2] * len(self.sigmas), layer_sizes[-1], dtype=config.real(torch))
        initializer(self._dense.weight)
        initializer_zero(self._dense.bias)

    def forward(self, inputs):
        x = inputs
        if self._input_transform is not None:
            x = self._input_transform(x)
        ys = []
        for b in self.b:
            xb = torch.matmul(x, b.to(x.dtype))
            y = torch.cat([torch.cos(xb), torch.sin(xb)], dim=1)
            for linear in self.linears:
                y = self.activation(linear(y))
                if self.dropout_rate > 0:
                    y = torch.nn.functional.dropout(y, p=self.dropout_rate, training=self.training)
            ys.append(y)
        y = self._dense(torch.cat(ys, dim=1))
        if self._output_transform is not None:
            y = self._output_transform(inputs, y)
        return y


class STMsFFN(MsFFN):
    """Spatio-temporal multi-scale Fourier feature network [DeepXDE-upstream `deepxde.nn.pytorch.STMsFFN`;
    the package is an absent, unpinned dependency of the reference, so this restates its published
    algorithm -- parity unpinned against upstream, pinned against the closed-form check in
    tests/test_oracle_pins.py], constructed by the reference at
    /root/reference/AI4XDE/cases/WaveCases.py:125-128 with sigmas_x=[1], sigmas_t=[1, 10].

    The last input column is time.  Space columns are embedded by fixed b_x ~ N(0, sigma_x^2) of shape
    (d_in - 1, layer_sizes[1] // 2), time by fixed b_t ~ N(0, sigma_t^2) of shape (1, layer_sizes[1] // 2); every
    embedding [cos(. b), sin(. b)] runs through the SAME hidden Linear + activation stack; the hidden outputs are
    multiplied point-wise for every (space, time) pair (space index outer), concatenated and mapped by `_dense`."""

    def __init__(self, layer_sizes, activation, kernel_initializer, sigmas_x, sigmas_t, dropout_rate=0):
        super().__init__(layer_sizes, activation, kernel_initializer, list(sigmas_x) + list(sigmas_t), dropout_rate)
        self.sigmas_x, self.sigmas_t = list(sigmas_x), list(sigmas_t)
        half = layer_sizes[1] // 2
        real = config.real(torch)
        self.b_x = [torch.normal(mean=0.0, std=float(s), size=(layer_sizes[0] - 1, half)).to(real) for s in self.sigmas_x]
        self.b_t = [torch.normal(mean=0.0, std=float(s), size=(1, half)).to(real) for s in self.sigmas_t]
        self._dense = torch.nn.Linear(layer_sizes[-2] * len(self.sigmas_x) * len(self.sigmas_t), layer_sizes[-1],
                                      dtype=real)
        get_initializer(kernel_initializer)(self._dense.weight)
        get_initializer("zeros")(self._dense.bias)

    def _branch(self, x, b):
        xb = torch.matmul(x, b.to(x.dtype))
        y = torch.cat([torch.cos(xb), torch.sin(xb)], dim=1)
        for linear in self.linears:
            y = self.activation(linear(y))
            if self.dropout_rate > 0:
                y = torch.nn.functional.dropout(y, p=self.dropout_rate, training=self.training)
        return y

    def forward(self, inputs):
        x = inputs
        if self._input_transform is not None:
            x = self._input_transform(x)
        y_x = [self._branch(x[:, :-1], b) for b in self.b_x]
        y_t = [self._branch(x[:, -1:], b) for b in self.b_t]
        y = torch.cat([torch.mul(a, c) for a in y_x for c in y_t], dim=1)
        y = self._dense(y)
        if self._output_transform is not None:
            y = self._output_transform(inputs, y)
        return y
