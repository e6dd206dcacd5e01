"""MonodomainSolver - mirror of gpuSolve/physics/monodomainSolver.py:26-123 on torch + libpdx.

(MASS + dt*STIFFNESS) U1 = MASS (U0 + dt (differentiate(U0) + I0)).  Per step three launches:
fused ionic update + RHS0 (pdx_ionic_step), SpMV with MASS, persistent PCG.  With the opt-in cfg key
``explicit_lumped`` the step is the north star's explicit mass-lumped update in ONE launch
(pdx_fem_explicit_step); the defaults stay drop-in.
"""
import ctypes as C

import numpy as np
import torch

from .heatSolver import HeatSolver
from ... import _lib as L

_IONIC_STATE_ATTRS = ('_H_state', '_V_state', '_W_state', '_S_state')


class MonodomainSolver(HeatSolver):
    _U_DEFAULT = -80.0

    def __init__(self, ionic_model, cfgdict=None):
        super().__init__(cfgdict)
        self._ionic_model = ionic_model
        if hasattr(self._ionic_model, '_dt'):
            self._ionic_model._dt = self._dt
        self._Ub = None

    # ---- nodal/material extensions -----------------------------------------
    def add_nodal_material_property(self, pname, ptype, prop):
        self._materials.add_nodal_property(pname, ptype, prop)

    def assign_nodal_properties(self):
        """Push nodal material properties into the ionic model (monodomainSolver.py:42-67)."""
        uniform_only = True
        names = self._materials.nodal_property_names()
        if names is not None:
            point_region_ids = self._Domain.point_region_ids()
            npt = point_region_ids.shape[0]
            for mat_prop in names:
                prtype = self._materials.nodal_property_type(mat_prop)
                refval = self._ionic_model.get_parameter(mat_prop)
                if refval is not None:
                    if prtype == 'uniform':
                        pvals = self._materials.NodalProperty(mat_prop, -1, -1)
                    else:
                        uniform_only = False
                        pvals = np.full((npt, 1), float(refval), dtype=np.float32)
                        if prtype == 'region':
                            idmap = self._materials._nodal_properties[mat_prop]['idmap']
                            for rid, val in idmap.items():
                                pvals[point_region_ids == rid] = val
                        else:
                            for pointID, regionID in enumerate(point_region_ids):
                                pvals[pointID] = self._materials.NodalProperty(mat_prop, pointID, regionID)
                    self._ionic_model.set_parameter(mat_prop, pvals)
        if uniform_only or (not self._use_renumbering):
            self._materials.remove_all_nodal_properties()

    # ---- setup overrides ---------------------------------------------------
    def set_initial_condition(self, U0=None):
        super().set_initial_condition(U0)
        self._ionic_model.initialize_state_variables(self._U)

    def finalize_for_run(self):
        if self._use_renumbering:
            perm, _ = self._perm_tensors()
            self._U = self._U.index_select(0, perm).contiguous()
            for attr in _IONIC_STATE_ATTRS:
                sv = getattr(self._ionic_model, attr, None)
                if sv is not None:
                    setattr(self._ionic_model, attr, sv.index_select(0, perm).contiguous())
            if self._StimulusDict is not None:
                for _key, stim in self._StimulusDict.items():
                    stim.apply_indices_permutation(self._renumbering['perm'])
            names = self._materials.nodal_property_names()
            if names is not None:
                for mat_prop in names:
                    prtype = self._materials.nodal_property_type(mat_prop)
                    refval = self._ionic_model.get_parameter(mat_prop)
                    if refval is not None and prtype != 'uniform' and isinstance(refval, torch.Tensor) \
                            and refval.numel() > 1:
                        self._ionic_model.set_parameter(mat_prop, refval.index_select(0, perm).cpu().numpy())
                self._materials.remove_all_nodal_properties()
        self._ready_for_run = True

    def _step_token(self):
        m = self._ionic_model
        return super()._step_token() + (id(m), getattr(m, '_param_version', 0), float(m._dt), getattr(m, 'math_mode', None),
                                        bool(self._explicit_lumped))

    # ---- per-step kernel ----------------------------------------------------
    def solve_step(self, U, I0):
        m = self._ionic_model
        states = m.state_tensors()
        dt32 = float(np.float32(self._dt))
        if m._MODEL_ID in (L.MODEL_CRN, L.MODEL_TTP):
            # lookup-table models: one pointwise kernel writes RHS0 = U + dt*(dU + I0)
            if self._explicit_lumped:
                raise L.PdxError('explicit_lumped is not available for the lookup-table ionic models')
            if float(np.float32(m._dt)) != dt32:
                raise L.PdxError('ionic model dt (%r) differs from the solver dt (%r)' % (m._dt, self._dt))
            rhs0, rhs = self._work()
            import ctypes as C
            L.check(L.lib().pdx_ionic_lut_step(C.byref(m.lut_desc()), U.numel(), L.ptr(U),
                                               L.ptr_array(states, L.LUT_MAX_STATES), None, L.ptr(I0), L.ptr(rhs0),
                                               L.stream_ptr()))
            return self._solve_from_rhs0(U, rhs0, rhs)
        params, parr, keep = m._param_block(U.numel())
        if self._explicit_lumped:
            if parr is not None:
                raise L.PdxError('explicit_lumped step takes uniform ionic parameters')
            if self._Ub is None or self._Ub[0].shape != U.shape:
                self._Ub = [torch.empty_like(U), torch.empty_like(U)]
            out = self._Ub[0] if U.data_ptr() != self._Ub[0].data_ptr() else self._Ub[1]
            K = self._STIFFNESS
            rp, ci, va = K.device_arrays()
            math = L.MATH_EXACT if m.math_mode == 'exact' else L.MATH_FAST
            L.check(L.lib().pdx_fem_explicit_step(m._MODEL_ID, math, params, K.n, L.ptr(rp), L.ptr(ci), L.ptr(va),
                                                  L.ptr(self._mlinv), L.ptr(U), L.ptr(out), L.ptr_array(states),
                                                  L.ptr(I0), dt32, L.stream_ptr()))
            return out
        rhs0, rhs = self._work()
        math = L.MATH_EXACT if (m.math_mode == 'exact' or parr is not None) else L.MATH_FAST
        # dU = differentiate(U) + I0 ; RHS0 = U + dt*dU     (monodomainSolver.py:100-103), one kernel
        L.check(L.lib().pdx_ionic_step(m._MODEL_ID, math, params, parr, U.numel(), L.ptr(U), L.ptr_array(states),
                                       None, dt32, L.ptr(I0), L.ptr(rhs0), L.stream_ptr()))
        del keep
        return self._solve_from_rhs0(U, rhs0, rhs)                         # RHS = MASS . RHS0, then PCG

    # ---- accessors ----------------------------------------------------------
    def ionic_model(self):
        return self._ionic_model

    def ionic_state(self, attr):
        sv = getattr(self._ionic_model, attr, None)
        if sv is None:
            return None
        if self._use_renumbering:
            _, iperm = self._perm_tensors()
            return sv.index_select(0, iperm)
        return sv
