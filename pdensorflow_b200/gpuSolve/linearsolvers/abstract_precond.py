"""AbstractPrecond - mirror of gpuSolve/linearsolvers/abstract_precond.py."""
from abc import ABC, abstractmethod


class AbstractPrecond(ABC):
    @abstractmethod
    def build_preconditioner(self, I0, J0, V0, msize):
        ...

    @abstractmethod
    def solve_precond_system(self, residual):
        ...
