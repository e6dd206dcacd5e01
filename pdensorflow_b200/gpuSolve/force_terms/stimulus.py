"""Stimulus - mirror of gpuSolve/force_terms/stimulus.py:7-165.

Timing semantics are the reference's (integer-step test for the FD examples, float-time
test for the FEM solvers); the stimulated region is a CUDA fp32 mask.
"""
import numpy as np
import torch

from .. import _backend as B


class Stimulus:
    """External / periodic pacing.  Properties: tstart, nstim, period, duration, intensity, name."""

    def __init__(self, props=None):
        self._tstart = 0.0
        self._nstim = 1
        self._period = 1.0
        self._duration = 1.0
        self._intensity = 1.0
        self._name = "s2"
        self.__active = True
        self.__stim = None
        if props:
            for attribute in ('_tstart', '_nstim', '_period', '_duration', '_intensity', '_name'):
                if attribute[1:] in props.keys():
                    setattr(self, attribute, props[attribute[1:]])
        self.__update_tend()

    def __update_tend(self):
        self.__tend = self._tstart + self._period * (self._nstim - 1) + self._duration

    def set_tstart(self, tstart):
        self._tstart = tstart
        self.__update_tend()

    def set_nstim(self, nstim):
        self._nstim = nstim
        self.__update_tend()

    def set_period(self, period):
        self._period = period
        self.__update_tend()

    def set_duration(self, duration):
        self._duration = duration
        self.__update_tend()

    def set_intensity(self, Imax):
        self._intensity = Imax

    def set_name(self, name):
        self._name = name

    def set_stimregion(self, streg):
        """Mark the stimulated points/voxels: anything non-zero counts (stimulus.py:73-80)."""
        if isinstance(streg, torch.Tensor):
            streg = streg.detach().cpu().numpy()
        region = np.squeeze(np.asarray(streg).astype(bool))
        if region.ndim == 1:
            region = region[:, np.newaxis]
        self.__stim = B.as_tensor(region.astype(np.float32))

    def get_stimregion(self):
        return self.__stim

    def apply_indices_permutation(self, perm_array):
        """Permute the region (used with reverse Cuthill-McKee renumbering)."""
        if self.__stim is not None:
            idx = torch.as_tensor(np.asarray(perm_array), dtype=torch.long, device=self.__stim.device)
            self.__stim = self.__stim.index_select(0, idx).contiguous()

    def deactivate(self):
        self.__active = False

    def activate(self):
        self.__active = True

    def is_active(self):
        return self.__active

    def stimulate_tissue_timestep(self, timestep, dt):
        """Integer-step activation test (stimulus.py:110-137)."""
        if not hasattr(self, '_int_duration'):
            self._int_duration = int(self._duration / dt)
        if not hasattr(self, '_int_period'):
            self._int_period = int(self._period / dt)
        if not hasattr(self, '_int_tstart'):
            self._int_tstart = int(self._tstart / dt)
        if not hasattr(self, '_int_tend'):
            self._int_tend = int((self._tstart + self._period * (self._nstim - 1) + self._duration) / dt)
        if self.__active:
            current_int_time = int(timestep) - self._int_tstart
            if current_int_time >= 0:
                if (current_int_time - self._int_tend) > 0:
                    self.__active = False
                elif current_int_time % self._int_period < self._int_duration:
                    return True
        return False

    def stimulate_tissue_timevalue(self, ctime_sim):
        """Float-time activation test (stimulus.py:139-155).  The FEM solver evaluates it on
        the fp32 value of the clock (heatSolver.py:141-145); pass np.float32 for that."""
        if self.__active:
            current_time = ctime_sim - float(self._tstart)
            if current_time >= 0:
                if (ctime_sim - float(self.__tend)) > 0.0:
                    self.__active = False
                elif current_time % float(self._period) <= float(self._duration):
                    return True
        return False

    def stimApp(self, ctime_sim):
        """Stimulus field at ctime_sim (zeros when inactive)."""
        if self.stimulate_tissue_timevalue(ctime_sim):
            return float(np.float32(self._intensity)) * self.__stim
        return 0.0 * self.__stim

    def intensity(self):
        return self._intensity

    def __call__(self):
        return float(np.float32(self._intensity)) * self.__stim
