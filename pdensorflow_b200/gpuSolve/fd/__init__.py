"""FD step helpers: what the reference's example scripts define next to the operators.

* ``enforce_boundary(X)``  - Tests/FD/Fenton/fenton.py:46-53 as one kernel
* ``fused_solve(model, U, ...)`` - the whole ``solve(U)`` of the examples
  (fenton.py:90-99, Fenton_sphere/fenton.py:109-115, HeatEquation/heat.py:86-94) as ONE
  kernel: enforce_boundary + model.differentiate + laplace + Euler update.  A drop-in
  ``solve`` body is ``return fused_solve(self, U, self.dt, self.diff, (DX, DY, DZ))``.
* ``apply_stimulus(U, stim)`` - fenton.py:154-156.
"""
import ctypes as C

import numpy as np
import torch

from .. import _backend as B
from ... import _lib as L


def enforce_boundary(X, dirichlet_value=None):
    """Symmetric pad of the interior (no-flux rule of the examples), or the constant pad
    of Tests/FD/LaplaceSolver/laplaceSolver.py:44-52 when ``dirichlet_value`` is given."""
    X = L.require_cuda(B.as_tensor(X), 'X')
    out = torch.empty_like(X)
    if X.dim() == 3:
        dims = (3,) + tuple(X.shape)
    elif X.dim() == 2:
        dims = (2, 1) + tuple(X.shape)
    else:
        raise L.PdxError('enforce_boundary takes 2-D or 3-D tensors')
    L.check(L.lib().pdx_fd_enforce_boundary(*dims, 0 if dirichlet_value is None else 1,
                                            0.0 if dirichlet_value is None else float(dirichlet_value),
                                            L.ptr(X), L.ptr(out), L.stream_ptr()))
    return out


class _Plan:
    def __init__(self, desc):
        self.handle = C.c_void_p()
        L.check(L.lib().pdx_fd_plan_create(C.byref(desc), C.byref(self.handle)))

    def __del__(self):
        try:
            L.lib().pdx_fd_plan_destroy(self.handle)
        except Exception:
            pass


_PLANS = {}
_LAPS = {'homog': L.LAP_HOMOG, 'conv': L.LAP_CONV, 'heterog': L.LAP_HETEROG, 'aniso': L.LAP_ANISO}


def fused_solve(model, U, dt, diff, spacing, lap='homog', DIFF=None, AVEC=None, math='fast', kernel='auto'):
    """One explicit step U -> U1 in a single kernel.  ``model`` is an IonicModel of this
    package (its states advance in place by ``dt``) or None for the heat equation.
    ``lap='heterog'``: DIFF [W,H,D] carries the diffusivity; ``lap='aniso'``: DIFF [W,H,D,2] and
    AVEC [W,H,D,6] as laplace_heterog_aniso takes them (both used as ``U0 + dt*dU + dt*lap``)."""
    U = L.require_cuda(U, 'U')
    ndim = U.dim()
    if ndim not in (2, 3) or len(spacing) != ndim:
        raise L.PdxError('fused_solve: U must be 2-D/3-D with matching spacing')
    model_id = L.MODEL_NONE if model is None else model._MODEL_ID
    is_lut = model_id in (L.MODEL_CRN, L.MODEL_TTP)
    params = () if (model is None or is_lut) else tuple(model.param_list())
    sp = tuple(B.as_float(s) for s in spacing)
    coef = float(np.float32(dt)) if lap in ('heterog', 'aniso') else float(np.float32(float(diff) * float(dt)))
    key = (tuple(U.shape), sp, float(np.float32(dt)), coef, lap, math, kernel, model_id, params, U.device.index,
           id(model) if is_lut else 0)
    plan = _PLANS.get(key)
    if plan is None:
        d = L.FdDesc()
        L.lib().pdx_fd_desc_init(C.byref(d))
        d.ndim = ndim
        if ndim == 3:
            d.nx, d.ny, d.nz = U.shape
            d.dx, d.dy, d.dz = sp
        else:
            d.nx, d.ny, d.nz = 1, U.shape[0], U.shape[1]
            d.dy, d.dz = sp
        d.dt, d.coef = float(np.float32(dt)), coef
        d.lap_kind, d.model = _LAPS[lap], model_id
        d.math_mode = L.MATH_EXACT if math == 'exact' else L.MATH_FAST
        d.kernel = {'auto': L.KERNEL_AUTO, 'generic': L.KERNEL_GENERIC, 'tma': L.KERNEL_TMA}[kernel]
        for i, v in enumerate(params):
            d.params[i] = v
        if len(_PLANS) > 32:
            _PLANS.clear()
        plan = _PLANS[key] = _Plan(d)
    states = []
    if model is not None:
        model.initialize_state_variables(U)
        states = model.state_tensors()
        if float(np.float32(model._dt)) != float(np.float32(dt)):
            raise L.PdxError('fused_solve: model._dt (%r) differs from dt (%r)' % (model._dt, dt))
    if is_lut:
        L.check(L.lib().pdx_fd_plan_set_lut(plan.handle, C.byref(model.lut_desc())))
    D = A = None
    if lap in ('heterog', 'aniso'):
        D = L.require_cuda(B.as_tensor(DIFF, like=U), 'DIFF')
    if lap == 'aniso':
        A = L.require_cuda(B.as_tensor(AVEC, like=U), 'AVEC')
        if tuple(D.shape) != tuple(U.shape) + (2,) or tuple(A.shape) != tuple(U.shape) + (6,):
            raise L.PdxError('aniso: DIFF must be [W,H,D,2] and AVEC [W,H,D,6]')
        L.check(L.lib().pdx_fd_plan_set_avec(plan.handle, L.ptr(A)))
    U1 = torch.empty_like(U)
    L.check(L.lib().pdx_fd_step(plan.handle, L.ptr(U), L.ptr(U1), L.ptr_array(states), L.ptr(D), 1,
                                L.stream_ptr()))
    return U1


def apply_stimulus(U, stim):
    """U = where(bool(stim), max(U, stim), U) in place on a fresh copy semantics-free path:
    returns U (modified in place), like the examples' reassignment of U."""
    U = L.require_cuda(U, 'U')
    mask = stim.get_stimregion()
    L.check(L.lib().pdx_stim_apply(L.ptr(U), L.ptr(mask), float(np.float32(stim.intensity())), U.numel(),
                                   L.stream_ptr()))
    return U
