"""linearsolvers submodule: ConjGrad, JacobiPrecond, AbstractPrecond (mirror of gpuSolve/linearsolvers)."""
from .._version import __version__, version  # noqa: F401
from .abstract_precond import AbstractPrecond  # noqa: F401
from .jacobi_precond import JacobiPrecond  # noqa: F401
from .conjgrad import ConjGrad  # noqa: F401
