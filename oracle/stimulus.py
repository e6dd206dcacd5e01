"""Oracle restatement of gpuSolve/force_terms/stimulus.py (timing logic + mask).

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  Not imported by the product.
"""
import numpy as np


class Stimulus:
    """gpuSolve/force_terms/stimulus.py:7-165."""

    def __init__(self, props=None):
        self._tstart = 0.0
        self._nstim = 1
        self._period = 1.0
        self._duration = 1.0
        self._intensity = 1.0
        self._name = "s2"
        self._active = True
        self._stim = None
        if props:
            for attribute in ('_tstart', '_nstim', '_period', '_duration', '_intensity', '_name'):
                if attribute[1:] in props:
                    setattr(self, attribute, props[attribute[1:]])
        # TF converts NumPy/Python scalars to the tensor's dtype (fp32) where they meet a
        # tensor; holding them as Python floats makes NumPy (NEP 50) do the same.
        for attribute in ('_tstart', '_period', '_duration', '_intensity'):
            setattr(self, attribute, float(getattr(self, attribute)))
        self._tend = self._tstart + self._period * (self._nstim - 1) + self._duration

    def set_stimregion(self, streg):
        """stimulus.py:73-80: anything non-zero is stimulated (astype(bool))."""
        region = np.squeeze(np.asarray(streg).astype(bool))
        if region.ndim == 1:
            region = region[:, np.newaxis]
        self._stim = region.astype(np.float32)

    def get_stimregion(self):
        return self._stim

    def apply_indices_permutation(self, perm):
        if self._stim is not None:
            self._stim = self._stim[perm]

    def is_active(self):
        return self._active

    def stimulate_tissue_timestep(self, timestep, dt):
        """stimulus.py:110-137 (integer-step test used by the FD examples)."""
        if not hasattr(self, '_int_duration'):
            self._int_duration = int(self._duration / dt)
            self._int_period = int(self._period / dt)
            self._int_tstart = int(self._tstart / dt)
            self._int_tend = int((self._tstart + self._period * (self._nstim - 1) + self._duration) / dt)
        if self._active:
            current = timestep - self._int_tstart
            if current >= 0:
                if (current - self._int_tend) > 0:
                    self._active = False
                elif current % self._int_period < self._int_duration:
                    return True
        return False

    def stimulate_tissue_timevalue(self, ctime_sim):
        """stimulus.py:139-155 (float-time test used by the FEM solvers).

        The FEM solver passes a float32 tensor (heatSolver.py:141-145); callers of the
        oracle pass ``np.float32(ctime)`` to reproduce that.
        """
        if self._active:
            current = ctime_sim - self._tstart
            if current >= 0:
                if (ctime_sim - self._tend) > 0.0:
                    self._active = False
                elif current % self._period <= self._duration:
                    return True
        return False

    def stimApp(self, ctime_sim):
        """stimulus.py:157-162."""
        if self.stimulate_tissue_timevalue(ctime_sim):
            return np.float32(self._intensity) * self._stim
        return np.float32(0.0) * self._stim

    def __call__(self):
        return np.float32(self._intensity) * self._stim
