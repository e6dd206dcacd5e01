"""physics submodule: HeatSolver, MonodomainSolver (mirror of gpuSolve/physics)."""
from .._version import __version__, version  # noqa: F401
from .heatSolver import HeatSolver  # noqa: F401
from .monodomainSolver import MonodomainSolver  # noqa: F401
