"""IonicModel base class - mirror of gpuSolve/ionic/ionicmodel.py:34-69 on torch + libpdx.

Parameters live on ``_name`` attributes exactly like the reference (scalars as 0-d fp32
tensors, per-node arrays as [N,1] CUDA tensors); ``differentiate`` launches
``pdx_ionic_step`` instead of ~60 TensorFlow elementwise ops.
"""
import ctypes as C

import numpy as np
import torch

from .. import _backend as B
from ... import _lib as L


class IonicModel:
    """Base class gathering what the ionic models share."""

    # subclasses: libpdx model id, parameter order of include/pdx.h, state attribute names
    _MODEL_ID = L.MODEL_NONE
    _PARAM_ORDER = ()
    _STATE_ATTRS = ()
    _STATE_INIT = ()
    #: arithmetic of differentiate(): 'exact' (op-for-op fp32, the default for the drop-in
    #: contract) or 'fast' (reciprocal multiplies / FMA / ex2-based tanh)
    math_mode = 'exact'

    def __init__(self, dt=0.0, n_nodes=0):
        self._dt = dt
        self._n_nodes = n_nodes
        self._initialized = False
        self._pb_cache = None

    def initialize_state_variables(self, U):
        """Create the state tensors matching U's shape (idempotent, fenton4v.py:84-91)."""
        if not self._initialized:
            U = B.as_tensor(U)
            for attr, v in zip(self._STATE_ATTRS, self._STATE_INIT):
                setattr(self, attr, torch.full_like(U, v))
            self._initialized = True

    def differentiate(self, U):
        """dU of the model at U; advances the internal states in place by ``self._dt``."""
        if self._MODEL_ID == L.MODEL_NONE:
            raise NotImplementedError("differentiate must be implemented in subclass")
        U = L.require_cuda(U, 'U')
        states = [getattr(self, a) for a in self._STATE_ATTRS]
        if any(s is None for s in states):
            raise L.PdxError('initialize_state_variables(U) must be called before differentiate(U)')
        for s in states:
            if s.shape != U.shape:
                raise L.PdxError('state shape %s does not match U %s' % (tuple(s.shape), tuple(U.shape)))
        dU = torch.empty_like(U)
        params, parr, keep = self._param_block(U.numel())
        math = L.MATH_EXACT if (self.math_mode == 'exact' or parr is not None) else L.MATH_FAST
        L.check(L.lib().pdx_ionic_step(self._MODEL_ID, math, params, parr, U.numel(), L.ptr(U),
                                       L.ptr_array(states), L.ptr(dU), float(np.float32(self._dt)),
                                       None, None, L.stream_ptr()))
        del keep
        return dU

    # ---- parameters ------------------------------------------------------------------
    def set_parameter(self, pname, pvalue):
        """If ``pname`` exists, set it to ``pvalue`` (scalar or per-node array)."""
        internal_name = '_{}'.format(pname)
        self._pb_cache = None
        self._param_version = getattr(self, '_param_version', 0) + 1     # invalidates graph-captured steps
        if internal_name in self.__dict__.keys():
            if np.ndim(pvalue) == 0 or np.size(pvalue) == 1:
                setattr(self, internal_name, torch.tensor(float(np.float32(np.asarray(pvalue).reshape(-1)[0])),
                                                          dtype=torch.float32))
            else:
                setattr(self, internal_name, B.as_tensor(np.asarray(pvalue, dtype=np.float32)))

    def get_parameter(self, pname):
        """Value of ``pname`` if it exists, None otherwise."""
        return getattr(self, '_{}'.format(pname), None)

    def _scalar(self, name):
        v = getattr(self, name)
        return float(v) if (not isinstance(v, torch.Tensor) or v.numel() == 1) else None

    def _param_block(self, n):
        """(c_float[PDX_MAX_PARAMS], per-node pointer array or None, keep-alive list)."""
        if self._pb_cache is not None and self._pb_cache[0] == n:
            return self._pb_cache[1]
        params = (C.c_float * L.MAX_PARAMS)()
        parr = None
        keep = []
        for i, name in enumerate(self._PARAM_ORDER):
            if name is None:
                continue
            v = getattr(self, name)
            if isinstance(v, torch.Tensor) and v.numel() > 1:
                if v.numel() != n:
                    raise L.PdxError('per-node parameter %s has %d entries, U has %d' % (name, v.numel(), n))
                if parr is None:
                    parr = (C.c_void_p * L.MAX_PARAMS)()
                t = B.as_tensor(v)
                keep.append(t)
                parr[i] = t.data_ptr()
            else:
                params[i] = float(v)
        self._pb_cache = (n, (params, parr, keep))
        return params, parr, keep

    def param_list(self):
        """Scalar parameter vector in libpdx order (raises if any parameter is per-node)."""
        out = []
        for name in self._PARAM_ORDER:
            if name is None:
                out.append(0.0)
                continue
            v = getattr(self, name)
            if isinstance(v, torch.Tensor) and v.numel() > 1:
                raise L.PdxError('parameter %s is per-node; the fused FD step takes uniform parameters' % name)
            out.append(float(v))
        return out

    def state_tensors(self):
        return [getattr(self, a) for a in self._STATE_ATTRS]
