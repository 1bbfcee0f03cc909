"""gPINN: gradient-enhanced residuals (/root/reference/AI4XDE/algorithm/gPINN.py:4-74).
The wrapped pde differentiates the residual once more (`jacobian(res_i, x, j)`), i.e. asks for one more order
of network derivatives than the PDE itself.  The kernels carry third order along one or two inputs (patterns
P_1_3 / P_2_3, FP32 engine) and fourth along one, which covers second-order PDEs in x or (x, t) -- the
reference's own example is Burgers (gPINN.py:77-81); beyond that the jet leaves raise NotImplementedError
(loudly, at the first evaluation)."""
from .. import dde
from ..solver import PINNSolver


class gPINN(PINNSolver):
    def __init__(self, PDECase):
        self.add_gradient_enhanced_res(PDECase)
        super().__init__(name="gPINN", PDECase=PDECase)

    def add_gradient_enhanced_res(self, PDECase):
        pde = PDECase.pde
        nargs = dde.utils.get_num_args(pde)
        if PDECase.loss_weights is not None:
            n_bc = len(PDECase.data.bcs)
            x_dim = PDECase.geomtime.random_points(1).shape[1]
            res_dim = len(PDECase.loss_weights) - n_bc
            PDECase.set_loss_weights(PDECase.loss_weights[:res_dim] + [1] * (res_dim * x_dim)
                                     + PDECase.loss_weights[res_dim:])

        def enhance(x, res):
            res = list(res) if isinstance(res, list) else [res]
            out = list(res)
            for r in res:
                for j in range(x.shape[1]):
                    out.append(dde.grad.jacobian(r, x, j=j))
            return out

        if nargs == 2:
            def g_pde(x, y):
                return enhance(x, pde(x, y))
        else:
            def g_pde(x, y, ex):
                return enhance(x, pde(x, y, ex))
        PDECase.set_pde(g_pde)
