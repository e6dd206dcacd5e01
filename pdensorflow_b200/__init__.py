"""pdensorflow_b200 - B200-native monodomain time stepping behind PDEnsorflow's interfaces.

* ``pdensorflow_b200.gpuSolve``  drop-in mirror of the reference's Python contracts
* ``pdensorflow_b200.fd``        fused / graph-captured / slab-distributed FD stepping
* ``pdensorflow_b200._lib``      ctypes binding of libpdx.so (include/pdx.h)
"""
from . import _lib  # noqa: F401

__all__ = ['_lib']
