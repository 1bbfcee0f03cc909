"""Side-car helpers mirroring AI4XDE.utils: seeding (used by parity tests) and lazy plotting."""
from .backend import empty_cache, setup_seed  # noqa: F401
from .visualization import Visualization_1D, Visualization_2D  # noqa: F401

__all__ = ["Visualization_1D", "Visualization_2D", "setup_seed", "empty_cache"]
