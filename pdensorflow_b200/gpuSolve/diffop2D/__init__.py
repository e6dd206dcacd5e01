"""diffop2D: 2-D diffusion operators (mirror of gpuSolve/diffop2D/__init__.py:20-21)."""
import torch

from .._version import __version__, version  # noqa: F401
from .. import _backend as B
from ... import _lib as L

math_mode = 'exact'


def _lap2(kind, X0, DIFF0, DX, DY):
    X0 = L.require_cuda(B.as_tensor(X0), 'X0')
    if X0.dim() != 2:
        raise L.PdxError('diffop2D operators take [W,H] tensors')
    D = None
    if DIFF0 is not None:
        D = L.require_cuda(B.as_tensor(DIFF0, like=X0), 'DIFF0')
        if D.shape != X0.shape:
            raise L.PdxError('DIFF0 must have the shape of X0')
    out = torch.empty_like(X0)
    ny, nz = X0.shape
    math = L.MATH_EXACT if math_mode == 'exact' else L.MATH_FAST
    L.check(L.lib().pdx_fd_laplace(2, 1, ny, nz, 1.0, B.as_float(DX), B.as_float(DY), kind, math,
                                   L.ptr(X0), L.ptr(D), L.ptr(out), L.stream_ptr()))
    return out


def laplace_homogeneous_isotropic_diffusion(X0, DX, DY):
    """5-point Laplacian, symmetric pad (diffop2D/laplace_homogeneous_isotropic_diffusion.py:3-36)."""
    return _lap2(L.LAP_HOMOG, X0, None, DX, DY)


def laplace_heterogeneous_isotropic_diffusion(X0, DIFF0, DX, DY):
    """div(D grad X), face-averaged D (diffop2D/laplace_heterogeneous_isotropic_diffusion.py:3-32)."""
    return _lap2(L.LAP_HETEROG, X0, DIFF0, DX, DY)


laplace_heterog = laplace_heterogeneous_isotropic_diffusion
laplace_homog = laplace_homogeneous_isotropic_diffusion
