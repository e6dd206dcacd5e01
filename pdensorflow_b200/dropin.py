"""``import pdensorflow_b200.dropin`` registers the mirror package as ``gpuSolve``.

After this import, ``import gpuSolve`` / ``from gpuSolve.ionic.fenton4v import Fenton4v``
resolve to pdensorflow_b200.gpuSolve, so the reference's example scripts run on the B200
kernels once their TensorFlow lines are replaced by torch.
"""
import importlib
import pkgutil
import sys

from . import gpuSolve as _pkg


def install():
    sys.modules.setdefault('gpuSolve', _pkg)
    for info in pkgutil.walk_packages(_pkg.__path__, prefix=_pkg.__name__ + '.'):
        try:
            mod = importlib.import_module(info.name)
        except Exception:      # optional sub-packages must not break the alias
            continue
        sys.modules.setdefault('gpuSolve' + info.name[len(_pkg.__name__):], mod)
    return _pkg


install()
