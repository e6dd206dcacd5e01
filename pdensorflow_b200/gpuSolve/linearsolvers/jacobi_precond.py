"""JacobiPrecond - mirror of gpuSolve/linearsolvers/jacobi_precond.py:5-41: V = 1/diag(A) (zeros where
the diagonal is absent); the preconditioned residual z = V*r is fused into libpdx's PCG kernel."""
import numpy as np

from .abstract_precond import AbstractPrecond
from .. import _backend as B


class JacobiPrecond(AbstractPrecond):
    def __init__(self):
        self._V = None
        self._dim = None

    def build_preconditioner(self, I0, J0, V0, msize):
        self._dim = msize
        V = np.zeros((self._dim, 1), dtype=float)
        diag = (I0 == J0)
        V[I0[diag], 0] = np.float32(1.0) / np.asarray(V0, dtype=np.float32)[diag]
        self._V = B.as_tensor(V.astype(np.float32))

    def inverse_diagonal(self):
        return self._V

    def solve_precond_system(self, residual):
        return self._V.reshape(residual.shape) * residual
