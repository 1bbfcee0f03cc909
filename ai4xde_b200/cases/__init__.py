"""Cases available on the fused path (mirror of AI4XDE.cases)."""
from .base import PDECases
from .pde_cases import (AllenCahn, Burgers, Diffusion, Harmonic_Oscillator_1D, Heat, Helmholtz,
                        NS_Flow_in_LidDriven_Cavity)
from .poisson_cases import PoissonCase1D, Poisson_1D_Dirichlet, Poisson_1D_Fourier_Net, multiscale_fbpinn_case

__all__ = ["PDECases", "Burgers", "Heat", "Diffusion", "AllenCahn", "Helmholtz", "NS_Flow_in_LidDriven_Cavity",
           "Harmonic_Oscillator_1D", "PoissonCase1D", "Poisson_1D_Dirichlet", "Poisson_1D_Fourier_Net", "multiscale_fbpinn_case"]
