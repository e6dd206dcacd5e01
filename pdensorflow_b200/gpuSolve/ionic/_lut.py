"""Shared machinery of the two lookup-table ionic models (CourtemancheRamirezNattel,
TenTusscherPanfilov): uniform tables built on the host in NumPy fp32 exactly as the reference
builds them (construct_tables: courtemanche_ramirez_nattel.py:255-391,
ten_tusscher_panfilov.py:234-371), uploaded once, then read by libpdx's kernels
(pdx_ionic_lut_step / the fused FD step) through linear interpolation.

A table column is described by how the reference tabulates it:
  ('rl_tau', x_inf, tau)  ->  A = -x_inf*expm1(-dt/tau), B = exp(-dt/tau)       (Rush-Larsen pair)
  ('rl_ab', alpha, beta)  ->  A = (-alpha/(alpha+beta))*expm1(-dt*(alpha+beta)), B = exp(-dt*(alpha+beta))
  ('col', values)         ->  one plain column
"""
import ctypes as C

import numpy as np
import torch

from .ionicmodel import IonicModel
from .. import _backend as B
from ... import _lib as L


class HostTable:
    """One uniform lookup table: grid [mn, mx) with spacing res, columns laid out as libpdx
    expects (include/pdx.h), rows padded to an even stride."""

    def __init__(self, mn, mx, res):
        self.mn, self.mx, self.res = mn, mx, res
        self.step = 1.0 / res
        self.mx_idx = int((mx - mn) * self.step) - 1
        self.grid = np.arange(mn, mx, res).astype(np.float32)
        self.cols = []

    def add(self, *values):
        for v in values:
            self.cols.append(np.asarray(v, dtype=np.float32) * np.ones_like(self.grid))

    def add_rl_tau(self, dt, inf, tau):
        self.add((-inf) * np.expm1(-dt / tau), np.exp(-dt / tau))

    def add_rl_ab(self, dt, a, b):
        self.add(((-a) / (a + b)) * np.expm1(-dt * (a + b)), np.exp(-dt * (a + b)))

    def finish(self):
        ncols = len(self.cols)
        stride = ncols if (ncols == 1 or ncols % 2 == 0) else ncols + 1
        tab = np.zeros((self.grid.shape[0], stride), dtype=np.float32)
        for c, v in enumerate(self.cols):
            tab[:, c] = v
        # singular grid points produce inf/nan; the reference zeroes them (..._nattel.py:373)
        self.array = np.nan_to_num(tab, nan=0.0, posinf=0.0, neginf=0.0)
        self.ncols, self.stride = ncols, stride
        self.cols = None
        return self

    def to_device(self, device):
        self.tensor = torch.from_numpy(self.array).to(device)
        return self.tensor

    def fill_desc(self, d):
        d.data = self.tensor.data_ptr()
        d.nrows, d.ncols, d.stride = self.array.shape[0], self.ncols, self.stride
        d.mn, d.mx, d.res, d.step = (float(np.float32(v)) for v in (self.mn, self.mx, self.res, self.step))
        d.mx_idx = self.mx_idx


class LutIonicModel(IonicModel):
    """IonicModel whose differentiate() is pdx_ionic_lut_step.  Subclasses provide
    ``_STATE_ATTRS``/``_STATE_INIT`` (libpdx state order), ``_build_tables()`` -> 3 HostTables and
    ``_fold_constants()`` -> list of Python floats in libpdx constant order."""

    _TABLE_ATTRS = ()

    def __init__(self, dt=0.0, n_nodes=0):
        super().__init__(dt, n_nodes)
        self._desc = None
        self._tables = None

    # ---- tables / constants --------------------------------------------------------------
    def construct_tables(self):
        """Build the three lookup tables for the current ``_dt`` (host, NumPy fp32) and upload them."""
        with np.errstate(all='ignore'):
            self._tables = [t.finish() for t in self._build_tables()]
        dev = B.device()
        for attr, t in zip(self._TABLE_ATTRS, self._tables):
            setattr(self, attr, t.to_device(dev))
        self._desc = None

    def _scalar_param(self, name):
        v = getattr(self, name)
        if isinstance(v, torch.Tensor):
            if v.numel() != 1:
                raise L.PdxError('%s: per-node values are not supported for %s (scalar parameters only; the '
                                 'per-node conductances of TenTusscherPanfilov are its _GCaL/_GKr/_GKs/_Gto state arrays)'
                                 % (name, type(self).__name__))
            return float(v)
        return float(v)

    def lut_desc(self):
        """struct pdx_lut_model for the current tables, constants and dt."""
        if self._tables is None:
            raise L.PdxError('initialize_state_variables(U) must be called first (it builds the lookup tables)')
        if self._desc is None or self._desc_dt != self._dt or self._desc_math != self.math_mode:
            d = L.LutModel()
            d.model = self._MODEL_ID
            consts = self._fold_constants()
            d.nconsts = len(consts)
            d.dt = float(np.float32(self._dt))
            for i, v in enumerate(consts):
                d.consts[i] = float(np.float32(v))
            for i, t in enumerate(self._tables):
                t.fill_desc(d.tab[i])
            d.math_mode = L.MATH_FAST if self.math_mode == 'fast' else L.MATH_EXACT
            self._desc, self._desc_dt, self._desc_math = d, self._dt, self.math_mode
        return self._desc

    def set_parameter(self, pname, pvalue):
        super().set_parameter(pname, pvalue)
        self._desc = None

    # ---- IonicModel protocol ---------------------------------------------------------------
    def initialize_state_variables(self, U):
        if not self._initialized:
            self.construct_tables()
            U = B.as_tensor(U)
            for attr, v in zip(self._STATE_ATTRS, self._STATE_INIT):
                setattr(self, attr, torch.full_like(U, v))
            self._initialized = True

    def differentiate(self, U):
        U = L.require_cuda(U, 'U')
        states = self.state_tensors()
        if any(s is None for s in states):
            raise L.PdxError('initialize_state_variables(U) must be called before differentiate(U)')
        for s in states:
            if s.numel() != U.numel() or not s.is_contiguous():
                raise L.PdxError('state arrays must be contiguous and match U')
        dU = torch.empty_like(U)
        L.check(L.lib().pdx_ionic_lut_step(C.byref(self.lut_desc()), U.numel(), L.ptr(U),
                                           L.ptr_array(states, L.LUT_MAX_STATES), L.ptr(dU), None, None,
                                           L.stream_ptr()))
        return dU
