"""Oracle restatement of the reference's pointwise ionic models (NumPy fp32).

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  Not imported by the product.

Semantics reproduced from TensorFlow: fp32 at every intermediate, Python scalars
rounded to fp32 where they meet a tensor, ``tf.sign(0) == 0`` (so ``H(0)=0.5``),
strict ``>`` in ``tf.where``, left-to-right evaluation as the source is written.
``tanh`` is evaluated as the correctly-rounded fp32 value (fp64 tanh rounded to
fp32); TF's own tanh (Eigen rational approximant, version dependent) differs from
it by <= 1-2 ulp, which the north-star tolerance (rel-L2 <= 1e-5) absorbs.
"""
import numpy as np

F = np.float32


def f32(x):
    return np.asarray(x, dtype=np.float32)


def tanh32(x):
    """fp32 tanh, correctly rounded (via fp64)."""
    return np.tanh(np.asarray(x, dtype=np.float64)).astype(np.float32)


def H(x):
    """Step function, gpuSolve/ionic/fenton4v.py:35-38."""
    return (F(1.0) + np.sign(x)) * F(0.5)


def G(x):
    """Complementary step, gpuSolve/ionic/fenton4v.py:40-43."""
    return (F(1.0) - np.sign(x)) * F(0.5)


class IonicModel:
    """gpuSolve/ionic/ionicmodel.py:34-69."""

    def __init__(self, dt=0.0, n_nodes=0):
        self._dt = dt
        self._n_nodes = n_nodes
        self._initialized = False

    def set_parameter(self, pname, pvalue):
        internal = '_{}'.format(pname)
        if internal in self.__dict__:
            setattr(self, internal, f32(pvalue))

    def get_parameter(self, pname):
        return getattr(self, '_{}'.format(pname), None)

    def to_dimensionless(self, U):
        return (U - self._vmin) / self._DV

    def derivative_to_dimensional(self, dU):
        return self._DV * dU


class Fenton4v(IonicModel):
    """gpuSolve/ionic/fenton4v.py:46-185 (defaults :56-78, step :168-185)."""

    PARAMS = dict(tau_vp=3.33, tau_vn=19.2, tau_wp=160.0, tau_wn1=75.0, tau_wn2=75.0,
                  tau_d=0.065, tau_si=31.8364, tau_so=31.8364, tau_a=0.009,
                  u_c=0.23, u_w=0.146, u_0=0.0, u_m=1.0, u_csi=0.8, u_so=0.3,
                  r_sp=0.02, r_sn=1.2, k_=3.0, a_so=0.115, b_so=0.84, c_so=0.02)

    def __init__(self, dt=0.0, n_nodes=0):
        super().__init__(dt, n_nodes)
        for k, v in self.PARAMS.items():
            setattr(self, '_' + k, F(v))
        self._vmin = F(-80.0)
        self._vmax = F(20.0)
        self._DV = self._vmax - self._vmin
        self._V_state = None
        self._W_state = None
        self._S_state = None

    def initialize_state_variables(self, U):
        if not self._initialized:
            self._V_state = np.ones_like(U, dtype=np.float32)
            self._W_state = np.ones_like(U, dtype=np.float32)
            self._S_state = np.zeros_like(U, dtype=np.float32)
            self._initialized = True

    def differentiate(self, U):
        """fenton4v.py:168-185: returns dU, advances V,W,S in place by self._dt."""
        one = F(1.0)
        V, W, S = self._V_state, self._W_state, self._S_state
        dt = F(self._dt)
        Uad = self.to_dimensionless(U)
        I_fi = -V * H(Uad - self._u_c) * (Uad - self._u_c) * (self._u_m - Uad) / self._tau_d
        I_si = -W * S / self._tau_si
        I_so = (F(0.5) * (self._a_so - self._tau_a) * (one + tanh32((Uad - self._b_so) / self._c_so)) +
                (Uad - self._u_0) * G(Uad - self._u_so) / self._tau_so + H(Uad - self._u_so) * self._tau_a)
        dU = -self.derivative_to_dimensional(I_fi + I_si + I_so)
        dV = np.where(Uad > self._u_c, -V / self._tau_vp, (one - V) / self._tau_vn)
        dW = np.where(Uad > self._u_c, -W / self._tau_wp,
                      np.where(Uad > self._u_w, (one - W) / self._tau_wn2, (one - W) / self._tau_wn1))
        r_s = (self._r_sp - self._r_sn) * H(Uad - self._u_c) + self._r_sn
        dS = r_s * (F(0.5) * (one + tanh32((Uad - self._u_csi) * self._k_)) - S)
        self._V_state = V + dt * dV
        self._W_state = W + dt * dW
        self._S_state = S + dt * dS
        return dU.astype(np.float32)


class ModifiedMS2v(IonicModel):
    """gpuSolve/ionic/mms2v.py:36-105 (defaults :47-55, step :95-105)."""

    PARAMS = dict(tau_in=0.1, tau_out=9.0, tau_open=100.0, tau_close=120.0, u_gate=0.13, u_crit=0.13)

    def __init__(self, dt=0.0, n_nodes=0):
        super().__init__(dt, n_nodes)
        for k, v in self.PARAMS.items():
            setattr(self, '_' + k, F(v))
        self._vmin = F(-80.0)
        self._vmax = F(20.0)
        self._DV = self._vmax - self._vmin
        self._H_state = None

    def initialize_state_variables(self, U):
        if not self._initialized:
            self._H_state = np.ones_like(U, dtype=np.float32)
            self._initialized = True

    def differentiate(self, U):
        one = F(1.0)
        Hs = self._H_state
        Uad = self.to_dimensionless(U)
        J_in = F(-1.0) * Hs * Uad * (Uad - self._u_crit) * (one - Uad) / self._tau_in
        J_out = (one - Hs) * Uad / self._tau_out
        dU = -self.derivative_to_dimensional(J_in + J_out)
        dH = np.where(Uad > self._u_gate, -Hs / self._tau_close, (one - Hs) / self._tau_open)
        self._H_state = (Hs + F(self._dt) * dH).astype(np.float32)
        return dU.astype(np.float32)


class MitchellSchaeffer2v(IonicModel):
    """gpuSolve/ionic/ms2v.py:36-100 (defaults :46-53, step :90-100)."""

    PARAMS = dict(tau_in=0.3, tau_out=6.0, tau_open=120.0, tau_close=150.0, u_gate=0.13)

    def __init__(self, dt=0.0, n_nodes=0):
        super().__init__(dt, n_nodes)
        for k, v in self.PARAMS.items():
            setattr(self, '_' + k, F(v))
        self._vmin = F(-80.0)
        self._vmax = F(20.0)
        self._DV = self._vmax - self._vmin
        self._H_state = None

    def initialize_state_variables(self, U):
        if not self._initialized:
            self._H_state = np.ones_like(U, dtype=np.float32)
            self._initialized = True

    def differentiate(self, U):
        one = F(1.0)
        Hs = self._H_state
        Uad = self.to_dimensionless(U)
        J_in = F(-1.0) * Hs * Uad * Uad * (one - Uad) / self._tau_in
        J_out = Uad / self._tau_out
        dU = -self.derivative_to_dimensional(J_in + J_out)
        dH = np.where(Uad > self._u_gate, -Hs / self._tau_close, (one - Hs) / self._tau_open)
        self._H_state = (Hs + F(self._dt) * dH).astype(np.float32)
        return dU.astype(np.float32)


MODELS = {'fenton4v': Fenton4v, 'mms2v': ModifiedMS2v, 'ms2v': MitchellSchaeffer2v}
STATE_ATTRS = {'fenton4v': ('_V_state', '_W_state', '_S_state'),
               'mms2v': ('_H_state',), 'ms2v': ('_H_state',)}
