from .pde_solver import PINNSolver

__all__ = ["PINNSolver"]
