"""MitchellSchaeffer2v - mirror of gpuSolve/ionic/ms2v.py:36-100 (Mitchell-Schaeffer 2003)."""

from .ionicmodel import IonicModel
from .fenton4v import _c, _getter
from ... import _lib as L


class MitchellSchaeffer2v(IonicModel):
    _MODEL_ID = L.MODEL_MS2V
    _PARAM_ORDER = ('_tau_in', '_tau_out', '_tau_open', '_tau_close', '_u_gate', None, '_vmin', '_DV')
    _STATE_ATTRS = ('_H_state',)
    _STATE_INIT = (1.0,)

    def __init__(self, dt=0.0, n_nodes=0):
        super().__init__(dt, n_nodes)
        self._tau_in = _c(0.3)
        self._tau_out = _c(6.0)
        self._tau_open = _c(120.0)
        self._tau_close = _c(150.0)
        self._u_gate = _c(0.13)
        self._vmin = _c(-80.0)
        self._vmax = _c(20.0)
        self._DV = self._vmax - self._vmin
        self._H_state = None

    def to_dimensionless(self, U):
        return (U - self._vmin.to(U.device)) / self._DV.to(U.device)

    def derivative_to_dimensional(self, dU):
        return self._DV.to(dU.device) * dU


for _n in ('_tau_in', '_tau_out', '_tau_open', '_tau_close', '_u_gate'):
    setattr(MitchellSchaeffer2v, _n[1:], _getter(_n))
