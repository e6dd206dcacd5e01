"""Oracle restatement of the reference's two lookup-table ionic models (NumPy fp32).

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  Not imported by the product.

* ``CourtemancheRamirezNattel`` - gpuSolve/ionic/courtemanche_ramirez_nattel.py
  (constants :50-135, table geometry :186-208, interpolation :241-252, tables :255-391,
  initial state :394-420, step :423-514)
* ``TenTusscherPanfilov`` - gpuSolve/ionic/ten_tusscher_panfilov.py
  (constants :53-127, table geometry :169-188, interpolation :220-231, tables :234-371,
  initial state :374-402, step :405-523)

Both integrate concentrations with forward Euler and gates with Rush-Larsen,
``y <- A(x) + B(x) y``, where A and B are tabulated on a uniform grid for the dt the model
was built with and read by linear interpolation.  Pinned against the reference's own source
executed over tests/golden/tf_facade.py: tests/golden/ref_ionic.npz (one-call vectors and paced
trajectories) and ref_luts.npz (table samples and checksums).

Arithmetic discipline (what TF does): state arrays, tables and every tensor intermediate are
fp32; sub-expressions made only of Python scalars are folded in float64 and rounded to fp32
where they first meet a tensor; evaluation is left to right as the reference writes it.  The
tables are evaluated the way the reference evaluates them: a float64 ``arange`` cast to fp32,
then NumPy fp32 arithmetic (so exp/expm1 are NumPy's fp32 routines there), overflowed entries
zeroed by ``nan_to_num``.
"""
from math import exp, expm1, log, sqrt

import numpy as np

from .ionic import IonicModel

F32 = np.float32


def _cr(fn):
    """Correctly rounded fp32 elementwise function (fp64 evaluation, one rounding): the
    platform-independent definition the oracle uses for tf.exp / tf.math.log / tf.math.expm1
    inside differentiate (same convention as oracle/ionic.py's tanh32)."""
    def op(x):
        return fn(np.asarray(x, dtype=np.float64)).astype(np.float32)
    return op


exp32, log32, expm132 = _cr(np.exp), _cr(np.log), _cr(np.expm1)


def _grid(mn, mx, res):
    return np.arange(mn, mx, res).astype(np.float32)


def _rl_tau(inf, tau, dt):
    """Rush-Larsen pair from (x_inf, tau): A = -x_inf * expm1(-dt/tau), B = exp(-dt/tau)."""
    return (-inf) * np.expm1(-dt / tau), np.exp(-dt / tau)


def _rl_ab(a, b, dt):
    """Rush-Larsen pair from opening/closing rates: tau = 1/(a+b), x_inf = a/(a+b)."""
    return ((-a) / (a + b)) * np.expm1(-dt * (a + b)), np.exp(-dt * (a + b))


class _Lut:
    """Uniform table: ``cols`` maps a column name to its fp32 samples on the grid."""

    def __init__(self, mn, mx, res, cols):
        self.mn, self.mx, self.res = mn, mx, res
        self.step = 1.0 / res
        self.mx_idx = int((mx - mn) * self.step) - 1
        self.names = list(cols)
        self.table = np.stack([np.asarray(cols[k], dtype=np.float32) for k in self.names], axis=1)
        self.table = np.nan_to_num(self.table, nan=0.0, posinf=0.0, neginf=0.0)

    def row(self, X):
        """Linear interpolation, courtemanche_ramirez_nattel.py:241-252 /
        ten_tusscher_panfilov.py:220-231.  Returns {name: fp32 array}."""
        Xc = np.clip(X, F32(self.mn), F32(self.mx))
        idx = np.trunc((Xc - F32(self.mn)) * F32(self.step)).astype(np.int32)
        lo = np.clip(idx, 0, self.mx_idx - 1)
        pos = lo.astype(np.float32) * F32(self.res) + F32(self.mn)
        w = ((Xc - pos) / F32(self.res))[..., None]
        r = (F32(1.0) - w) * self.table[lo] + w * self.table[lo + 1]
        return {k: r[..., i] for i, k in enumerate(self.names)}


# =============================================================================================
class CourtemancheRamirezNattel(IonicModel):
    CONST = dict(C_B1c=0.00705882352941176, C_B1d=0.00537112496043221, C_B1e=11.5, C_Fn1=9.648e-13,
                 C_Fn2=2.5910306809e-13, C_dCa_rel=8.0, C_dCaup=0.0869565217391304, F=96.4867, K_Q10=3.0,
                 K_up=0.00092, KmCa=1.38, KmCmdn=0.00238, KmCsqn=0.8, KmKo=1.5, KmNa3=669921.875, KmNai=10.0,
                 KmTrpn=0.0005, Nai=11.2, R_gas=8.3143, T=310.0, Voli=13668.0, gamma=0.35, k_rel=30.0, k_sat=0.1,
                 maxCmdn=0.05, maxTrpn=0.07, tau_f_Ca=2.0, tau_tr=180.0, tau_u=8.0)
    PARAMS = dict(ACh=0.000001, Cao=1.8, GACh=0.0, GCaL=0.1238, GK1=0.09, GKr=0.0294, GKs=0.129, GNa=7.8,
                  GbCa=0.00113, GbNa=0.000674, Gto=0.1652, Ko=5.4, Nao=140.0, factorGKur=1.0, factorGrel=1.0,
                  factorGtr=1.0, factorGup=1.0, factorhGate=0.0, factormGate=0.0, factoroaGate=0.0,
                  factorxrGate=1.0, maxCaup=15.0, maxINaCa=1600.0, maxINaK=0.60, maxIpCa=0.275, maxIup=0.005)
    # state name -> initial value (:60-62,66,83-106)
    STATES = dict(Ca_rel=1.49, Ca_up=1.49, Cai=1.02e-1, Ki=139.0, d=1.37e-4, f=0.999, f_Ca=0.775, h=0.965, j=0.978,
                  m=2.91e-3, oa=3.04e-2, oi=0.999, u=0.0, ua=4.96e-3, ui=0.999, v=1.0, w=0.999, xr=3.29e-5,
                  xs=1.87e-2)
    V_init = -81.2

    def __init__(self, dt=0.0, n_nodes=0):
        super().__init__(dt, n_nodes)
        for k, v in {**self.CONST, **self.PARAMS}.items():
            setattr(self, '_' + k, v)
        self.s = None
        self.luts = None

    # ---- tables ----------------------------------------------------------------------------
    def construct_tables(self):
        p, dt = self, self._dt
        E_Na = ((p._R_gas * p._T) / p._F) * log(p._Nao / p._Nai)
        sigma = (exp(p._Nao / 67.3) - 1.0) / 7.0
        self.B_fCa = exp((-dt) / p._tau_f_Ca)
        self.B_u = exp((-dt) / p._tau_u)

        c = _grid(3e-4, 30.0, 3e-4) / 1000.0         # conCa
        ca = {}
        ca['Iup_ca'] = (p._factorGup * p._maxIup) / (1.0 + (p._K_up / c))
        ca['buf'] = (((((p._maxTrpn * p._KmTrpn) / (c + p._KmTrpn)) / (c + p._KmTrpn) +
                       ((p._maxCmdn * p._KmCmdn) / (c + p._KmCmdn)) / (c + p._KmCmdn)) + 1.0) / p._C_B1c) / 1000.0
        ca['IpCa_ca'] = (p._maxIpCa * c) / (0.0005 + c) - (((p._GbCa * p._R_gas * p._T) / 2.0) / p._F) * np.log(p._Cao / c)
        ca['conCa'] = c
        ca['fCa_A'] = (-(1.0 / (1.0 + (c / 0.00035)))) * expm1((-dt) / p._tau_f_Ca)

        V = _grid(-200.0, 200.0, 0.1)
        ex = np.exp
        v = {}
        a_h = np.where(V >= -40.0, 0.0, 0.135 * ex(((V + 80.0) - p._factorhGate) / -6.8))
        b_h = np.where(V >= -40.0, (1.0 / 0.13) / (1.0 + ex((-(V + 10.66)) / 11.1)),
                       3.56 * ex(0.079 * V) + 3.1e5 * ex(0.35 * V))
        a_j = np.where(V < -40.0, (((-127140.0 * ex(0.2444 * V)) - (3.474e-5 * ex(-0.04391 * V))) * (V + 37.78)) /
                       (1.0 + ex(0.311 * (V + 79.23))), 0.0)
        b_j = np.where(V >= -40.0, (0.3 * ex(-2.535e-7 * V)) / (1.0 + ex(-0.1 * (V + 32.0))),
                       (0.1212 * ex(-0.01052 * V)) / (1.0 + ex(-0.1378 * (V + 40.14))))
        a_m = np.where(np.abs(V + 47.13) < 1e-10, 3.2, (0.32 * (V + 47.13)) / (1.0 - ex(-0.1 * (V + 47.13))))
        b_m = 0.08 * ex((-(V - p._factormGate)) / 11.0)
        v['m_A'], v['m_B'] = _rl_ab(a_m, b_m, dt)
        v['h_A'], v['h_B'] = _rl_ab(a_h, b_h, dt)
        v['j_A'], v['j_B'] = _rl_ab(a_j, b_j, dt)

        def q10(a, b):
            return (1.0 / (a + b)) / p._K_Q10
        a_oa = 0.65 / (ex((V + 10.0) / -8.5) + ex((30.0 - V) / 59.0))
        b_oa = 0.65 / (2.5 + ex((V + 82.0) / 17.0))
        v['oa_A'], v['oa_B'] = _rl_tau(1.0 / (1.0 + ex((-((V + 20.47) - p._factoroaGate)) / 17.54)), q10(a_oa, b_oa), dt)
        a_oi = 1.0 / (18.53 + ex((V + 113.7) / 10.95))
        b_oi = 1.0 / (35.56 + ex((-(V + 1.26)) / 7.44))
        v['oi_A'], v['oi_B'] = _rl_tau(1.0 / (1.0 + ex((V + 43.1) / 5.3)), q10(a_oi, b_oi), dt)
        a_ua = 0.65 / (ex((V + 10.0) / -8.5) + ex((V - 30.0) / -59.0))
        b_ua = 0.65 / (2.5 + ex((V + 82.0) / 17.0))
        v['ua_A'], v['ua_B'] = _rl_tau(1.0 / (1.0 + ex((V + 30.3) / -9.6)), q10(a_ua, b_ua), dt)
        a_ui = 1.0 / (21.0 + ex((V - 185.0) / -28.0))
        b_ui = ex((V - 158.0) / 16.0)
        v['ui_A'], v['ui_B'] = _rl_tau(1.0 / (1.0 + ex((V - 99.45) / 27.48)), q10(a_ui, b_ui), dt)
        a_xr = p._factorxrGate * ((0.0003 * (V + 14.1)) / (1.0 - ex((V + 14.1) / -5.0)))
        b_xr = (1.0 / p._factorxrGate) * ((7.3898e-5 * (V - 3.3328)) / (ex((V - 3.3328) / 5.1237) - 1.0))
        v['xr_A'], v['xr_B'] = _rl_tau(1.0 / (1.0 + ex((V + 14.1) / -6.5)), 1.0 / (a_xr + b_xr), dt)
        a_xs = (4.0e-5 * (V - 19.9)) / (1.0 - ex((19.9 - V) / 17.0))
        b_xs = (3.5e-5 * (V - 19.9)) / (ex((V - 19.9) / 9.0) - 1.0)
        v['xs_A'], v['xs_B'] = _rl_tau(1.0 / np.sqrt(1.0 + ex((V - 19.9) / -12.7)), 0.5 / (a_xs + b_xs), dt)
        tau_d = np.where(np.abs(V + 10.0) < 1e-10, ((1.0 / 6.24) / 0.035) / 2.0,
                         ((1.0 - ex((-(V + 10.0)) / 6.24)) / 0.035 / (V + 10.0)) / (1.0 + ex((-(V + 10.0)) / 6.24)))
        v['d_A'], v['d_B'] = _rl_tau(1.0 / (1.0 + ex((-(V + 10.0)) / 8.0)), tau_d, dt)
        tau_f = 9.0 / (0.0197 * ex(-0.0337 * 0.0337 * (V + 10.0) * (V + 10.0)) + 0.02)
        v['f_A'], v['f_B'] = _rl_tau(1.0 / (1.0 + ex((V + 28.0) / 6.9)), tau_f, dt)
        den = V - 7.9
        den = np.where(np.abs(den) < 1e-10, 1e-10, den)
        tau_w = ((6.0 * (1.0 - ex((7.9 - V) / 5.0))) / (1.0 + 0.3 * ex((7.9 - V) / 5.0))) / den
        v['w_A'], v['w_B'] = _rl_tau(1.0 - 1.0 / (1.0 + ex((40.0 - V) / 17.0)), tau_w, dt)

        v['GKur'] = 0.005 + 0.05 / (1.0 + ex((15.0 - V) / 13.0))
        f_NaK = 1.0 / (1.0 + 0.1245 * ex((-0.1 * p._F * V / p._R_gas) / p._T) +
                       0.0365 * sigma * ex((-p._F * V / p._R_gas) / p._T))
        v['INaK'] = ((p._maxINaK * f_NaK / (1.0 + pow(p._KmNai / p._Nai, 1.5))) * p._Ko) / (p._Ko + p._KmKo)
        v['IbNa'] = p._GbNa * (V - E_Na)
        v['ICaL_v'] = p._GCaL * (V - 65.0)
        e1 = ex(((p._gamma - 1.0) * p._F * V / p._R_gas) / p._T)
        v['NaCa_a'] = (((((p._maxINaCa * ex((p._gamma * p._F * V / p._R_gas) / p._T)) * p._Nai * p._Nai * p._Nai * p._Cao) /
                         (p._KmNa3 + p._Nao ** 3)) / (p._KmCa + p._Cao)) / (1.0 + p._k_sat * e1))
        v['NaCa_b'] = (((((p._maxINaCa * e1) * p._Nao ** 3) / (p._KmNa3 + p._Nao ** 3)) / (p._KmCa + p._Cao)) /
                       (1.0 + p._k_sat * e1)) / 1000.0
        v['IbCa_v'] = V * p._GbCa
        v['INa_v'] = p._GNa * (V - E_Na)

        fn = _grid(-2e-11, 10.0e-11, 2e-15)
        g = {}
        tau_v = 1.91 + 2.09 / (1.0 + ex((3.4175e-13 - fn) / 13.67e-16))
        u_inf = 1.0 / (1.0 + ex((3.4175e-13 - fn) / 13.67e-16))
        v_inf = 1.0 - 1.0 / (1.0 + ex((6.835e-14 - fn) / 13.67e-16))
        g['u_A'] = (-u_inf) * expm1(-dt / p._tau_u)
        g['v_A'], g['v_B'] = _rl_tau(v_inf, tau_v, dt)
        with np.errstate(all='ignore'):
            pass
        self.luts = {'V': _Lut(-200.0, 200.0, 0.1, v), 'Cai': _Lut(3e-4, 30.0, 3e-4, ca),
                     'fn': _Lut(-2e-11, 10.0e-11, 2e-15, g)}

    def initialize_state_variables(self, U):
        if not self._initialized:
            with np.errstate(all='ignore'):
                self.construct_tables()
            self.s = {k: np.full(np.shape(U), v, dtype=np.float32) for k, v in self.STATES.items()}
            self._initialized = True

    # ---- step ------------------------------------------------------------------------------
    def differentiate(self, U):
        p, s, dt = self, self.s, self._dt
        shape = np.shape(U)
        V = np.reshape(np.asarray(U, dtype=np.float32), -1)
        y = {k: a.reshape(-1) for k, a in s.items()}
        vr = self.luts['V'].row(V)
        cr = self.luts['Cai'].row(y['Cai'])
        E_K = ((p._R_gas * p._T) / p._F) * log32(p._Ko / y['Ki'])
        ICaL = vr['ICaL_v'] * y['d'] * y['f'] * y['f_Ca']
        IKACh = (p._GACh * (10.0 / (1.0 + (9.13652 / (pow(p._ACh, 0.477811))))) *
                 (0.0517 + 0.4516 / (1.0 + exp32((V + 59.53) / 17.18)))) * (V - E_K)
        INa = vr['INa_v'] * y['m'] * y['m'] * y['m'] * y['h'] * y['j']
        INaCa = vr['NaCa_a'] - y['Cai'] * vr['NaCa_b']
        IK1 = (p._GK1 * (V - E_K)) / (1.0 + exp32(0.07 * (V + 80.0)))
        IKr = ((p._GKr * (V - E_K)) / (1.0 + exp32((V + 15.0) / 22.4))) * y['xr']
        IKs = (p._GKs * (V - E_K)) * y['xs'] * y['xs']
        IKur = ((p._factorGKur * vr['GKur']) * (V - E_K)) * y['ua'] * y['ua'] * y['ua'] * y['ui']
        IpCa = cr['IpCa_ca'] + vr['IbCa_v']
        Ito = (p._Gto * (V - E_K)) * y['oa'] * y['oa'] * y['oa'] * y['oi']
        Iion = INa + IK1 + Ito + IKur + IKr + IKs + ICaL + IpCa + INaCa + vr['IbNa'] + vr['INaK'] + IKACh

        Itr = (p._factorGtr * (y['Ca_up'] - y['Ca_rel'])) / p._tau_tr
        Irel = p._factorGrel * y['u'] * y['u'] * y['v'] * p._k_rel * y['w'] * (y['Ca_rel'] - cr['conCa'])
        dIups = cr['Iup_ca'] - (p._maxIup / p._maxCaup) * y['Ca_up']
        d_Ca_rel = (Itr - Irel) / (1.0 + (p._C_dCa_rel / (y['Ca_rel'] + p._KmCsqn)) / (y['Ca_rel'] + p._KmCsqn))
        d_Ki = (-(((Ito + IKr + IKur + IKs + IK1 + IKACh) - 2.0 * vr['INaK']))) / (p._F * p._Voli)
        d_Ca_up = dIups - Itr * p._C_dCaup
        d_Cai = ((p._C_B1d * ((INaCa + INaCa) - IpCa - ICaL)) - (p._C_B1e * dIups) + Irel) / cr['buf']
        fn = (p._C_Fn1 * Irel) - (p._C_Fn2 * (ICaL - 0.4 * INaCa))
        fr = self.luts['fn'].row(fn)

        new = {'Ca_rel': y['Ca_rel'] + d_Ca_rel * dt, 'Ca_up': y['Ca_up'] + d_Ca_up * dt,
               'Cai': y['Cai'] + d_Cai * dt, 'Ki': y['Ki'] + d_Ki * dt,
               'f_Ca': cr['fCa_A'] + self.B_fCa * y['f_Ca'], 'u': fr['u_A'] + self.B_u * y['u'],
               'v': fr['v_A'] + fr['v_B'] * y['v']}
        for gname in ('d', 'f', 'h', 'j', 'm', 'oa', 'oi', 'ua', 'ui', 'w', 'xr', 'xs'):
            new[gname] = vr[gname + '_A'] + vr[gname + '_B'] * y[gname]
        for k, a in new.items():
            s[k] = a.astype(np.float32).reshape(shape)
        return (-Iion).astype(np.float32).reshape(shape)


# =============================================================================================
class TenTusscherPanfilov(IonicModel):
    CONST = dict(EC=1.5, Vxfer=0.0038, k1_=0.15, k2_=0.045, k3=0.060, k4=0.005, maxsr=2.5, minsr=1.0)
    PARAMS = dict(Bufc=0.2, Bufsr=10.0, Bufss=0.4, CAPACITANCE=0.185, Cao=2.0, D_CaL_off=0.0, Fconst=96485.3415,
                  GK1=5.405, GNa=14.838, GbCa=0.000592, GbNa=0.00029, GpCa=0.1238, GpK=0.0146, Kbufc=0.001,
                  Kbufsr=0.3, Kbufss=0.00025, KmCa=1.38, KmK=1.0, KmNa=40.0, KmNai=87.5, Ko=5.4, KpCa=0.0005,
                  Kup=0.00025, Nao=140.0, Rconst=8314.472, T=310.0, Vc=0.016404, Vleak=0.00036, Vmaxup=0.006375,
                  Vrel=0.102, Vsr=0.001094, Vss=0.00005468, knaca=1000.0, knak=2.724, ksat=0.1, n=0.35, pKNa=0.03,
                  scl_tau_f=1.0, vHalfXs=5.0, xr2_off=0.0)
    V_init = -86.2

    def __init__(self, dt=0.0, n_nodes=0, cell_type='EPI'):
        super().__init__(dt, n_nodes)
        self._cell_type = cell_type if cell_type is not None else 'EPI'
        for k, v in {**self.CONST, **self.PARAMS}.items():
            setattr(self, '_' + k, v)
        ct = self._cell_type
        # conductances are per-node state arrays in the reference (:124-127, :380-383)
        self.STATES = dict(GCaL=0.00003980, GKr=0.153,
                           GKs=0.392 if ct == 'EPI' else 0.098 if ct == 'MCELL' else 0.392,
                           Gto=0.294 if ct == 'EPI' else 0.294 if ct == 'MCELL' else 0.073,
                           CaSR=1.3, CaSS=0.00007, Cai=0.00007, F=1.0, F2=1.0, FCaSS=1.0, H=0.75, J=0.75, Ki=138.3,
                           M=0.0, Nai=7.67, R=0.0, R_bar=1.0, S=1.0, Xr1=0.0, Xr2=1.0, Xs=0.0, D=0.0)
        self.s = None
        self.luts = None

    def construct_tables(self):
        p, dt = self, self._dt
        ex = np.exp
        Nao3 = p._Nao * p._Nao * p._Nao
        F_RT = 1.0 / ((p._Rconst * p._T) / p._Fconst)
        pmf = (p._knaca * (1.0 / (p._KmNai * p._KmNai * p._KmNai + Nao3))) * (1.0 / (p._KmCa + p._Cao))

        c = _grid(0.00001, 10.0, 0.00001)
        cs = {}
        cs['FCaSS_A'], cs['FCaSS_B'] = _rl_tau((0.6 / (1.0 + (c / 0.05) * (c / 0.05))) + 0.4,
                                               (80.0 / (1.0 + (c / 0.05) * (c / 0.05))) + 2.0, dt)

        V = _grid(-800.0, 800.0, 0.05)
        v = {}
        a_D = (1.4 / (1.0 + ex((-35.0 - V) / 13.0))) + 0.25
        b_D = 1.4 / (1.0 + ex((V + 5.0) / 5.0))
        c_D = 1.0 / (1.0 + ex((50.0 - V) / 20.0))
        v['D_A'], v['D_B'] = _rl_tau(1.0 / (1.0 + ex(((-8.0 + p._D_CaL_off) - V) / 7.5)), a_D * b_D + c_D, dt)
        a_F2 = 562.0 * ex(-(V + 27.0) * (V + 27.0) / 240.0)
        b_F2 = 31.0 / (1.0 + ex((25.0 - V) / 10.0))
        c_F2 = 80.0 / (1.0 + ex((V + 30.0) / 10.0))
        v['F2_A'], v['F2_B'] = _rl_tau((0.67 / (1.0 + ex((V + 35.0) / 7.0))) + 0.33, a_F2 + b_F2 + c_F2, dt)
        a_F = 1102.5 * ex(-(V + 27.0) * (V + 27.0) / 225.0)
        b_F = 200.0 / (1.0 + ex((13.0 - V) / 10.0))
        c_F = (180.0 / (1.0 + ex((V + 30.0) / 10.0))) + 20.0
        tF = a_F + b_F + c_F
        v['F_A'], v['F_B'] = _rl_tau(1.0 / (1.0 + ex((V + 20.0) / 7.0)), np.where(V > 0.0, tF * p._scl_tau_f, tF), dt)
        H_inf = 1.0 / ((1.0 + ex((V + 71.55) / 7.43)) * (1.0 + ex((V + 71.55) / 7.43)))
        a_H = np.where(V >= -40.0, 0.0, 0.057 * ex(-(V + 80.0) / 6.8))
        b_H = np.where(V >= -40.0, 0.77 / (0.13 * (1.0 + ex(-(V + 10.66) / 11.1))),
                       2.7 * ex(0.079 * V) + 3.1e5 * ex(0.3485 * V))
        v['H_A'], v['H_B'] = _rl_tau(H_inf, 1.0 / (a_H + b_H), dt)
        a_J = np.where(V >= -40.0, 0.0, ((-2.5428e4 * ex(0.2444 * V) - 6.948e-6 * ex(-0.04391 * V)) * (V + 37.78)) /
                       (1.0 + ex(0.311 * (V + 79.23))))
        b_J = np.where(V >= -40.0, (0.6 * ex(0.057 * V)) / (1.0 + ex(-0.1 * (V + 32.0))),
                       (0.02424 * ex(-0.01052 * V)) / (1.0 + ex(-0.1378 * (V + 40.14))))
        v['J_A'], v['J_B'] = _rl_tau(H_inf, 1.0 / (a_J + b_J), dt)
        a_M = 1.0 / (1.0 + ex((-60.0 - V) / 5.0))
        b_M = (0.1 / (1.0 + ex((V + 35.0) / 5.0))) + (0.10 / (1.0 + ex((V - 50.0) / 200.0)))
        M_inf = 1.0 / ((1.0 + ex((-56.86 - V) / 9.03)) * (1.0 + ex((-56.86 - V) / 9.03)))
        v['M_A'], v['M_B'] = _rl_tau(M_inf, a_M * b_M, dt)
        v['R_A'], v['R_B'] = _rl_tau(1.0 / (1.0 + ex((20.0 - V) / 6.0)),
                                     9.5 * ex(-(V + 40.0) * (V + 40.0) / 1800.0) + 0.8, dt)
        if self._cell_type == 'ENDO':
            tau_S = 1000.0 * ex(-(V + 67.0) * (V + 67.0) / 1000.0) + 8.0
        else:
            tau_S = 85.0 * ex(-(V + 45.0) * (V + 45.0) / 320.0) + 5.0 / (1.0 + ex((V - 20.0) / 5.0)) + 3.0
        v['S_A'], v['S_B'] = _rl_tau(1.0 / (1.0 + ex((V + 20.0) / 5.0)), tau_S, dt)
        v['Xr1_A'], v['Xr1_B'] = _rl_tau(1.0 / (1.0 + ex((-26.0 - V) / 7.0)),
                                         (450.0 / (1.0 + ex((-45.0 - V) / 10.0))) * (6.0 / (1.0 + ex((V + 30.0) / 11.5))), dt)
        v['Xr2_A'], v['Xr2_B'] = _rl_tau(1.0 / (1.0 + ex((V - (-88.0 + p._xr2_off)) / 24.0)),
                                         (3.0 / (1.0 + ex((-60.0 - V) / 20.0))) * (1.12 / (1.0 + ex((V - 60.0) / 20.0))), dt)
        v['Xs_A'], v['Xs_B'] = _rl_tau(1.0 / (1.0 + ex((-p._vHalfXs - V) / 14.0)),
                                       (1400.0 / np.sqrt(1.0 + ex((5.0 - V) / 6.0))) * (1.0 / (1.0 + ex((V - 35.0) / 15.0))) + 80.0, dt)
        v['a2'] = 0.25 * ex(2.0 * (V - 15.0) * F_RT)
        den = pmf / (1.0 + p._ksat * ex((p._n - 1.0) * V * F_RT))
        v['NaCa_A'] = (den * p._Cao) * ex(p._n * V * F_RT)
        v['NaCa_B'] = (den * ex((p._n - 1.0) * V * F_RT)) * Nao3 * 2.5
        v['rec_iNaK'] = 1.0 / (1.0 + 0.1245 * ex(-0.1 * V * F_RT) + 0.0353 * ex(-V * F_RT))
        v['rec_ipK'] = 1.0 / (1.0 + ex((25.0 - V) / 5.98))

        E = _grid(-800.0, 800.0, 0.05)
        a_K1 = 0.1 / (1.0 + ex(0.06 * (E - 200.0)))
        b_K1 = (3.0 * ex(0.0002 * (E + 100.0)) + ex(0.1 * (E - 10.0))) / (1.0 + ex(-0.5 * E))
        self.luts = {'V': _Lut(-800.0, 800.0, 0.05, v), 'CaSS': _Lut(0.00001, 10.0, 0.00001, cs),
                     'VEk': _Lut(-800.0, 800.0, 0.05, {'rec_iK1': a_K1 / (a_K1 + b_K1)})}

    def initialize_state_variables(self, U):
        if not self._initialized:
            with np.errstate(all='ignore'):
                self.construct_tables()
            self.s = {k: np.full(np.shape(U), v, dtype=np.float32) for k, v in self.STATES.items()}
            self._initialized = True

    def differentiate(self, U):
        p, s, dt = self, self.s, self._dt
        shape = np.shape(U)
        V = np.reshape(np.asarray(U, dtype=np.float32), -1)
        y = {k: a.reshape(-1) for k, a in s.items()}
        RTONF = (p._Rconst * p._T) / p._Fconst
        inverseVcF2 = 1.0 / (2.0 * p._Vc * p._Fconst)
        inverseVssF2 = 1.0 / (2.0 * p._Vss * p._Fconst)
        pmf_INaK = p._knak * (p._Ko / (p._Ko + p._KmK))
        sqrt_Ko = sqrt(p._Ko / 5.4)
        F_RT = 1.0 / RTONF
        invVcF_Cm = (1.0 / (p._Vc * p._Fconst)) * p._CAPACITANCE
        Cai, CaSS, CaSR, Nai, Ki, R_bar = y['Cai'], y['CaSS'], y['CaSR'], y['Nai'], y['Ki'], y['R_bar']

        vr = self.luts['V'].row(V)
        cr = self.luts['CaSS'].row(CaSS)
        Eca = 0.5 * RTONF * log32(p._Cao / Cai)
        Ek = RTONF * log32(p._Ko / Ki)
        Eks = RTONF * log32((p._Ko + p._pKNa * p._Nao) / (Ki + p._pKNa * Nai))
        Ena = RTONF * log32(p._Nao / Nai)
        IpCa = (p._GpCa * Cai) / (p._KpCa + Cai)
        with np.errstate(all='ignore'):
            a1 = (y['GCaL'] * p._Fconst * F_RT * 4.0) * np.where(np.abs(V - 15.0) < 1e-10, F32(0.5 * F_RT),
                                                                 (V - 15.0) / expm132(2.0 * (V - 15.0) * F_RT))
        ICaL_A = a1 * vr['a2']
        ICaL_B = a1 * p._Cao
        IKr = y['GKr'] * sqrt_Ko * y['Xr1'] * y['Xr2'] * (V - Ek)
        IKs = y['GKs'] * y['Xs'] * y['Xs'] * (V - Eks)
        INa = p._GNa * y['M'] * y['M'] * y['M'] * y['H'] * y['J'] * (V - Ena)
        INaK = pmf_INaK * (Nai / (Nai + p._KmNa)) * vr['rec_iNaK']
        IbCa = p._GbCa * (V - Eca)
        IbNa = p._GbNa * (V - Ena)
        IpK = p._GpK * vr['rec_ipK'] * (V - Ek)
        Ito = y['Gto'] * y['R'] * y['S'] * (V - Ek)
        er = self.luts['VEk'].row(V - Ek)
        ICaL = y['D'] * y['F'] * y['F2'] * y['FCaSS'] * (ICaL_A * CaSS - ICaL_B)
        INaCa = vr['NaCa_A'] * Nai * Nai * Nai - vr['NaCa_B'] * Cai
        IK1 = p._GK1 * er['rec_iK1'] * (V - Ek)
        Iion = IKr + IKs + IK1 + Ito + INa + IbNa + ICaL + IbCa + INaK + INaCa + IpCa + IpK

        Ileak = p._Vleak * (CaSR - Cai)
        Iup = p._Vmaxup / (1.0 + (p._Kup * p._Kup) / (Cai * Cai))
        Ixfer = p._Vxfer * (CaSS - Cai)
        d_Ki = (-(IK1 + Ito + IKr + IKs - 2.0 * INaK + IpK)) * invVcF_Cm
        d_Nai = (-(INa + IbNa + 3.0 * INaK + 3.0 * INaCa)) * invVcF_Cm
        kCaSR = p._maxsr - (p._maxsr - p._minsr) / (1.0 + (p._EC / CaSR) * (p._EC / CaSR))
        d_Cai = ((Ixfer - ((IbCa + IpCa - 2.0 * INaCa) * inverseVcF2 * p._CAPACITANCE)) -
                 ((Iup - Ileak) * (p._Vsr / p._Vc))) / (1.0 + (p._Bufc * p._Kbufc / (p._Kbufc + Cai)) / (p._Kbufc + Cai))
        d_R_bar = p._k4 * (1.0 - R_bar) - p._k2_ * kCaSR * CaSS * R_bar
        k1 = p._k1_ / kCaSR
        O = (k1 * CaSS * CaSS * R_bar) / (p._k3 + k1 * CaSS * CaSS)
        Irel = p._Vrel * O * (CaSR - CaSS)
        d_CaSR = ((Iup - Irel - Ileak) / (1.0 + (p._Bufsr * p._Kbufsr / (CaSR + p._Kbufsr)) / (CaSR + p._Kbufsr)))
        d_CaSS = ((-Ixfer * (p._Vc / p._Vss) + Irel * (p._Vsr / p._Vss) + (-ICaL * inverseVssF2 * p._CAPACITANCE)) /
                  (1.0 + (p._Bufss * p._Kbufss / (CaSS + p._Kbufss)) / (CaSS + p._Kbufss)))

        new = {'CaSR': CaSR + d_CaSR * dt, 'CaSS': CaSS + d_CaSS * dt, 'Cai': Cai + d_Cai * dt, 'Ki': Ki + d_Ki * dt,
               'Nai': Nai + d_Nai * dt, 'R_bar': R_bar + d_R_bar * dt,
               'FCaSS': cr['FCaSS_A'] + cr['FCaSS_B'] * y['FCaSS']}
        for gname in ('D', 'F', 'F2', 'H', 'J', 'M', 'R', 'S', 'Xr1', 'Xr2', 'Xs'):
            new[gname] = vr[gname + '_A'] + vr[gname + '_B'] * y[gname]
        for k, a in new.items():
            s[k] = a.astype(np.float32).reshape(shape)
        return (-Iion).astype(np.float32).reshape(shape)
