"""Algorithms available on the fused path (mirror of AI4XDE.algorithm).  HPO, FI_PINN and the
skopt-based Uniform sampler are outside the hot-path scope."""
from .basic import PINN, RAR_G, Random_R
from .rad import RAD
from .rar_d import RAR_D
from .r3sampling import R3Sampling
from .gpinn import gPINN
from .fbpinn import FBPINN
from .causal import CausalPINN

__all__ = ["PINN", "R3Sampling", "RAD", "Random_R", "RAR_D", "RAR_G", "gPINN", "FBPINN", "CausalPINN"]
