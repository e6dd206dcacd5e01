"""diffop3D: 3-D diffusion operators (mirror of gpuSolve/diffop3D/__init__.py:25-28).

Pure functions on CUDA fp32 tensors of shape [W, H, D] returning a new tensor; each is one
launch of libpdx's Laplacian kernel (pdx_fd_laplace) instead of pad + strided slices +
elementwise TensorFlow ops.
"""
import torch

from .._version import __version__, version  # noqa: F401
from .. import _backend as B
from ... import _lib as L

#: arithmetic of the stand-alone operators ('exact' = the reference's evaluation order)
math_mode = 'exact'


def _lap3(kind, X0, DIFF0, DX, DY, DZ):
    X0 = L.require_cuda(B.as_tensor(X0), 'X0')
    if X0.dim() != 3:
        raise L.PdxError('diffop3D operators take [W,H,D] tensors')
    D = None
    if DIFF0 is not None:
        D = L.require_cuda(B.as_tensor(DIFF0, like=X0), 'DIFF0')
        if D.shape != X0.shape:
            raise L.PdxError('DIFF0 must have the shape of X0')
    out = torch.empty_like(X0)
    nx, ny, nz = X0.shape
    math = L.MATH_EXACT if math_mode == 'exact' else L.MATH_FAST
    L.check(L.lib().pdx_fd_laplace(3, nx, ny, nz, B.as_float(DX), B.as_float(DY), B.as_float(DZ), kind, math,
                                   L.ptr(X0), L.ptr(D), L.ptr(out), L.stream_ptr()))
    return out


def laplace_homogeneous_isotropic_diffusion(X0, DX, DY, DZ):
    """7-point Laplacian, symmetric pad (laplace_homogeneous_isotropic_diffusion.py:3-52)."""
    return _lap3(L.LAP_HOMOG, X0, None, DX, DY, DZ)


def laplace_convolution_homogeneous_isotropic_diffusion(X0, DX, DY, DZ):
    """Same stencil written as the reference's 3x3x3 convolution
    (laplace_convolution_homogeneous_isotropic_diffusion.py:4-47)."""
    return _lap3(L.LAP_CONV, X0, None, DX, DY, DZ)


def laplace_heterogeneous_isotropic_diffusion(X0, DIFF0, DX, DY, DZ):
    """div(D grad X) with face-averaged D (laplace_heterogeneous_isotropic_diffusion.py:3-35)."""
    return _lap3(L.LAP_HETEROG, X0, DIFF0, DX, DY, DZ)


def laplace_heterogeneous_anisotropic_diffusion(X0, DIFF0, AVEC0, DX, DY, DZ):
    """19-point fibre-tensor operator (laplace_heterogeneous_anisotropic_diffusion.py:3-95):
    DIFF0 [W,H,D,2] = (sigma_t, sigma_l - sigma_t), AVEC0 [W,H,D,6] = the fibre dyadic."""
    X0 = L.require_cuda(B.as_tensor(X0), 'X0')
    if X0.dim() != 3:
        raise L.PdxError('diffop3D operators take [W,H,D] tensors')
    D = L.require_cuda(B.as_tensor(DIFF0, like=X0), 'DIFF0')
    A = L.require_cuda(B.as_tensor(AVEC0, like=X0), 'AVEC0')
    if tuple(D.shape) != tuple(X0.shape) + (2,) or tuple(A.shape) != tuple(X0.shape) + (6,):
        raise L.PdxError('DIFF0 must be [W,H,D,2] and AVEC0 [W,H,D,6]')
    out = torch.empty_like(X0)
    nx, ny, nz = X0.shape
    L.check(L.lib().pdx_fd_laplace_aniso(nx, ny, nz, B.as_float(DX), B.as_float(DY), B.as_float(DZ), L.ptr(X0),
                                         L.ptr(D), L.ptr(A), L.ptr(out), L.stream_ptr()))
    return out


laplace_heterog = laplace_heterogeneous_isotropic_diffusion
laplace_heterog_aniso = laplace_heterogeneous_anisotropic_diffusion
laplace_homog = laplace_homogeneous_isotropic_diffusion
laplace_conv_homog = laplace_convolution_homogeneous_isotropic_diffusion
