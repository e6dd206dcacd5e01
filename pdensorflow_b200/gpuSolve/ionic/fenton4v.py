"""Fenton4v - mirror of gpuSolve/ionic/fenton4v.py:46-185 (Cherry-Ehrlich-Nattel-Fenton 4v).

The arithmetic runs in libpdx (pdx_math.cuh: fk_exact / fk_fast); this class keeps the
reference's attribute names, defaults (fenton4v.py:56-78) and getters.
"""
import torch

from .ionicmodel import IonicModel
from ... import _lib as L


def _c(v):
    return torch.tensor(v, dtype=torch.float32)


class Fenton4v(IonicModel):
    _MODEL_ID = L.MODEL_FENTON4V
    _PARAM_ORDER = ('_tau_vp', '_tau_vn', '_tau_wp', '_tau_wn1', '_tau_wn2', '_tau_d', '_tau_si', '_tau_so',
                    '_tau_a', '_u_c', '_u_w', '_u_0', '_u_m', '_u_csi', '_u_so', '_r_sp', '_r_sn', '_k_',
                    '_a_so', '_b_so', '_c_so', '_vmin', '_DV')
    _STATE_ATTRS = ('_V_state', '_W_state', '_S_state')
    _STATE_INIT = (1.0, 1.0, 0.0)

    def __init__(self, dt=0.0, n_nodes=0):
        super().__init__(dt, n_nodes)
        self._tau_vp = _c(3.33)
        self._tau_vn = _c(19.2)
        self._tau_wp = _c(160.0)
        self._tau_wn1 = _c(75.0)
        self._tau_wn2 = _c(75.0)
        self._tau_d = _c(0.065)
        self._tau_si = _c(31.8364)
        self._tau_so = _c(31.8364)
        self._tau_a = _c(0.009)
        self._u_c = _c(0.23)
        self._u_w = _c(0.146)
        self._u_0 = _c(0.0)
        self._u_m = _c(1.0)
        self._u_csi = _c(0.8)
        self._u_so = _c(0.3)
        self._r_sp = _c(0.02)
        self._r_sn = _c(1.2)
        self._k_ = _c(3.0)
        self._a_so = _c(0.115)
        self._b_so = _c(0.84)
        self._c_so = _c(0.02)
        self._vmin = _c(-80.0)
        self._vmax = _c(20.0)
        self._DV = self._vmax - self._vmin
        self._V_state = None
        self._W_state = None
        self._S_state = None

    def to_dimensionless(self, U):
        return (U - self._vmin.to(U.device)) / self._DV.to(U.device)

    def derivative_to_dimensional(self, dU):
        return self._DV.to(dU.device) * dU


def _getter(name):
    def get(self):
        return getattr(self, name)
    get.__name__ = name[1:]
    return get


for _n in Fenton4v._PARAM_ORDER[:-2]:
    setattr(Fenton4v, _n[1:], _getter(_n))


def H(x):
    """Step function with sign(0)=0 semantics (fenton4v.py:35-38); helper for user scripts."""
    return (1.0 + torch.sign(x)) * 0.5


def G(x):
    """Complementary step (fenton4v.py:40-43)."""
    return (1.0 - torch.sign(x)) * 0.5
