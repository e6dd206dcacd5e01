"""CourtemancheRamirezNattel - mirror of gpuSolve/ionic/courtemanche_ramirez_nattel.py:36-514
(Courtemanche, Ramirez, Nattel 1998 human atrial model; 19 state arrays, three lookup tables,
Rush-Larsen gates, forward-Euler concentrations).

Attribute names, defaults, initial state and the table grids are the reference's; the step itself
runs in libpdx (csrc/pdx_ionic_lut.cu: crn_update), pointwise through ``differentiate`` or fused
with the FD stencil.  Table columns are laid out for the kernel (include/pdx.h), not in the
reference's column order.
"""
from math import exp, expm1, log

import numpy as np

from ._lut import HostTable, LutIonicModel
from ... import _lib as L


class CourtemancheRamirezNattel(LutIonicModel):
    _MODEL_ID = L.MODEL_CRN
    # libpdx state order (include/pdx.h) = the reference's declaration order (:220-238)
    _STATE_ATTRS = ('_Ca_rel', '_Ca_up', '_Cai', '_Ki', '_d_state', '_f_state', '_f_Ca_state', '_h_state', '_j_state',
                    '_m_state', '_oa_state', '_oi_state', '_u_state', '_ua_state', '_ui_state', '_v_state', '_w_state',
                    '_xr_state', '_xs_state')
    _TABLE_ATTRS = ('_V_tab', '_Cai_tab', '_fn_tab')

    def __init__(self, dt=0.0, n_nodes=0):
        super().__init__(dt, n_nodes)
        # constants (:50-106)
        self._C_B1a, self._C_B1b = 3.79138232501097e-05, 0.0811764705882353
        self._C_B1c, self._C_B1d, self._C_B1e = 0.00705882352941176, 0.00537112496043221, 11.5
        self._C_Fn1, self._C_Fn2 = 9.648e-13, 2.5910306809e-13
        self._C_dCa_rel, self._C_dCaup = 8.0, 0.0869565217391304
        self._Ca_rel_init, self._Ca_up_init, self._Cai_init, self._Ki_init = 1.49, 1.49, 1.02e-1, 139.0
        self._F, self._K_Q10, self._K_up = 96.4867, 3.0, 0.00092
        self._KmCa, self._KmCmdn, self._KmCsqn, self._KmKo = 1.38, 0.00238, 0.8, 1.5
        self._KmNa, self._KmNa3, self._KmNai, self._KmTrpn = 87.5, 669921.875, 10.0, 0.0005
        self._Nai, self._R_gas, self._T, self._V_init = 11.2, 8.3143, 310.0, -81.2
        self._Volcell, self._Voli, self._Volrel, self._Volup = 20100.0, 13668.0, 96.48, 1109.52
        self._d_init, self._f_Ca_init, self._f_init, self._h_init, self._j_init = 1.37e-4, 0.775, 0.999, 0.965, 0.978
        self._gamma, self._k_rel, self._k_sat, self._m_init = 0.35, 30.0, 0.1, 2.91e-3
        self._maxCmdn, self._maxCsqn, self._maxTrpn = 0.05, 10.0, 0.07
        self._oa_init, self._oi_init = 3.04e-2, 0.999
        self._tau_f_Ca, self._tau_tr, self._tau_u = 2.0, 180.0, 8.0
        self._u_init, self._ua_init, self._ui_init, self._v_init, self._w_init = 0.0, 4.96e-3, 0.999, 1.0, 0.999
        self._xr_init, self._xs_init = 3.29e-5, 1.87e-2
        # parameters (:108-135)
        self._ACh, self._Cao, self._Cm, self._GACh, self._GCaL = 0.000001, 1.8, 100.0, 0.0, 0.1238
        self._GK1, self._GKr, self._GKs, self._GNa = 0.09, 0.0294, 0.129, 7.8
        self._GbCa, self._GbNa, self._Gto, self._Ko, self._Nao = 0.00113, 0.000674, 0.1652, 5.4, 140.0
        self._factorGKur, self._factorGrel, self._factorGtr, self._factorGup = 1.0, 1.0, 1.0, 1.0
        self._factorhGate, self._factormGate, self._factoroaGate, self._factorxrGate = 0.0, 0.0, 0.0, 1.0
        self._maxCaup, self._maxINaCa, self._maxINaK, self._maxIpCa, self._maxIup = 15.0, 1600.0, 0.60, 0.275, 0.005
        # table geometry (:186-208)
        self._Cai_T_mn, self._Cai_T_mx, self._Cai_T_res = 3e-4, 30.0, 3e-4
        self._V_T_mn, self._V_T_mx, self._V_T_res = -200.0, 200.0, 0.1
        self._fn_T_mn, self._fn_T_mx, self._fn_T_res = -2e-11, 10.0e-11, 2e-15
        self._V_tab = self._Cai_tab = self._fn_tab = None
        self._f_Ca_rush_larsen_B = self._u_rush_larsen_B = None
        for a in self._STATE_ATTRS:
            setattr(self, a, None)

    @property
    def _STATE_INIT(self):
        return (self._Ca_rel_init, self._Ca_up_init, self._Cai_init, self._Ki_init, self._d_init, self._f_init,
                self._f_Ca_init, self._h_init, self._j_init, self._m_init, self._oa_init, self._oi_init, self._u_init,
                self._ua_init, self._ui_init, self._v_init, self._w_init, self._xr_init, self._xs_init)

    # ---- tables (:255-391) -------------------------------------------------------------------
    def _build_tables(self):
        s = self._scalar_param
        dt = self._dt
        R, T, F = s('_R_gas'), s('_T'), s('_F')
        Nao, Nai, Cao, Ko = s('_Nao'), s('_Nai'), s('_Cao'), s('_Ko')
        E_Na = ((R * T) / F) * log(Nao / Nai)
        sigma = (exp(Nao / 67.3) - 1.0) / 7.0
        ex = np.exp
        # frozen at table-construction time like the reference's (:263-264)
        self._f_Ca_rush_larsen_B = exp((-dt) / s('_tau_f_Ca'))
        self._u_rush_larsen_B = exp((-dt) / s('_tau_u'))

        tv = HostTable(self._V_T_mn, self._V_T_mx, self._V_T_res)
        V = tv.grid
        fh, fm, foa, fxr = s('_factorhGate'), s('_factormGate'), s('_factoroaGate'), s('_factorxrGate')
        a_m = np.where(np.abs(V + 47.13) < 1e-10, 3.2, (0.32 * (V + 47.13)) / (1.0 - ex(-0.1 * (V + 47.13))))
        tv.add_rl_ab(dt, a_m, 0.08 * ex((-(V - fm)) / 11.0))
        a_h = np.where(V >= -40.0, 0.0, 0.135 * ex(((V + 80.0) - fh) / -6.8))
        b_h = np.where(V >= -40.0, (1.0 / 0.13) / (1.0 + ex((-(V + 10.66)) / 11.1)),
                       3.56 * ex(0.079 * V) + 3.1e5 * ex(0.35 * V))
        tv.add_rl_ab(dt, a_h, b_h)
        a_j = np.where(V < -40.0, (((-127140.0 * ex(0.2444 * V)) - (3.474e-5 * ex(-0.04391 * V))) * (V + 37.78)) /
                       (1.0 + ex(0.311 * (V + 79.23))), 0.0)
        b_j = np.where(V >= -40.0, (0.3 * ex(-2.535e-7 * V)) / (1.0 + ex(-0.1 * (V + 32.0))),
                       (0.1212 * ex(-0.01052 * V)) / (1.0 + ex(-0.1378 * (V + 40.14))))
        tv.add_rl_ab(dt, a_j, b_j)
        q10 = s('_K_Q10')
        # (x_inf, alpha, beta) of the Q10-scaled gates oa, oi, ua, ui: tau = (1/(alpha+beta))/K_Q10
        for inf, al, be in (
                (1.0 / (1.0 + ex((-((V + 20.47) - foa)) / 17.54)),
                 0.65 / (ex((V + 10.0) / -8.5) + ex((30.0 - V) / 59.0)), 0.65 / (2.5 + ex((V + 82.0) / 17.0))),
                (1.0 / (1.0 + ex((V + 43.1) / 5.3)),
                 1.0 / (18.53 + ex((V + 113.7) / 10.95)), 1.0 / (35.56 + ex((-(V + 1.26)) / 7.44))),
                (1.0 / (1.0 + ex((V + 30.3) / -9.6)),
                 0.65 / (ex((V + 10.0) / -8.5) + ex((V - 30.0) / -59.0)), 0.65 / (2.5 + ex((V + 82.0) / 17.0))),
                (1.0 / (1.0 + ex((V - 99.45) / 27.48)),
                 1.0 / (21.0 + ex((V - 185.0) / -28.0)), ex((V - 158.0) / 16.0))):
            tv.add_rl_tau(dt, inf, (1.0 / (al + be)) / q10)
        aa_xr = fxr * ((0.0003 * (V + 14.1)) / (1.0 - ex((V + 14.1) / -5.0)))
        bb_xr = (1.0 / fxr) * ((7.3898e-5 * (V - 3.3328)) / (ex((V - 3.3328) / 5.1237) - 1.0))
        tv.add_rl_tau(dt, 1.0 / (1.0 + ex((V + 14.1) / -6.5)), 1.0 / (aa_xr + bb_xr))
        aa_xs = (4.0e-5 * (V - 19.9)) / (1.0 - ex((19.9 - V) / 17.0))
        bb_xs = (3.5e-5 * (V - 19.9)) / (ex((V - 19.9) / 9.0) - 1.0)
        tv.add_rl_tau(dt, 1.0 / np.sqrt(1.0 + ex((V - 19.9) / -12.7)), 0.5 / (aa_xs + bb_xs))
        tau_d = np.where(np.abs(V + 10.0) < 1e-10, ((1.0 / 6.24) / 0.035) / 2.0,
                         ((1.0 - ex((-(V + 10.0)) / 6.24)) / 0.035 / (V + 10.0)) / (1.0 + ex((-(V + 10.0)) / 6.24)))
        tv.add_rl_tau(dt, 1.0 / (1.0 + ex((-(V + 10.0)) / 8.0)), tau_d)
        tv.add_rl_tau(dt, 1.0 / (1.0 + ex((V + 28.0) / 6.9)),
                      9.0 / (0.0197 * ex(-0.0337 * 0.0337 * (V + 10.0) * (V + 10.0)) + 0.02))
        wden = V - 7.9
        wden = np.where(np.abs(wden) < 1e-10, 1e-10, wden)
        tv.add_rl_tau(dt, 1.0 - 1.0 / (1.0 + ex((40.0 - V) / 17.0)),
                      ((6.0 * (1.0 - ex((7.9 - V) / 5.0))) / (1.0 + 0.3 * ex((7.9 - V) / 5.0))) / wden)
        gam, ksat, KmNa3, KmCa, mNaCa = s('_gamma'), s('_k_sat'), s('_KmNa3'), s('_KmCa'), s('_maxINaCa')
        f_NaK = 1.0 / (1.0 + 0.1245 * ex((-0.1 * F * V / R) / T) + 0.0365 * sigma * ex((-F * V / R) / T))
        e1 = ex(((gam - 1.0) * F * V / R) / T)
        tv.add(0.005 + 0.05 / (1.0 + ex((15.0 - V) / 13.0)),                                             # GKur
               ((s('_maxINaK') * f_NaK / (1.0 + pow(s('_KmNai') / Nai, 1.5))) * Ko) / (Ko + s('_KmKo')),  # INaK
               s('_GbNa') * (V - E_Na),                                                                  # IbNa
               s('_GCaL') * (V - 65.0),
               (((((mNaCa * ex((gam * F * V / R) / T)) * Nai * Nai * Nai * Cao) / (KmNa3 + Nao ** 3)) / (KmCa + Cao)) /
                (1.0 + ksat * e1)),
               (((((mNaCa * e1) * Nao ** 3) / (KmNa3 + Nao ** 3)) / (KmCa + Cao)) / (1.0 + ksat * e1)) / 1000.0,
               V * s('_GbCa'),
               s('_GNa') * (V - E_Na))

        tc = HostTable(self._Cai_T_mn, self._Cai_T_mx, self._Cai_T_res)
        c = tc.grid / 1000.0
        KmT, KmC = s('_KmTrpn'), s('_KmCmdn')
        tc.add((s('_factorGup') * s('_maxIup')) / (1.0 + (s('_K_up') / c)),
               (((((s('_maxTrpn') * KmT) / (c + KmT)) / (c + KmT) + ((s('_maxCmdn') * KmC) / (c + KmC)) / (c + KmC)) + 1.0) /
                s('_C_B1c')) / 1000.0,
               (s('_maxIpCa') * c) / (0.0005 + c) - (((s('_GbCa') * R * T) / 2.0) / F) * np.log(Cao / c),
               c,
               (-(1.0 / (1.0 + (c / 0.00035)))) * expm1((-dt) / s('_tau_f_Ca')))

        tf_ = HostTable(self._fn_T_mn, self._fn_T_mx, self._fn_T_res)
        fn = tf_.grid
        sig = 1.0 / (1.0 + ex((3.4175e-13 - fn) / 13.67e-16))
        tf_.add((-sig) * expm1(-dt / s('_tau_u')))
        tf_.add_rl_tau(dt, 1.0 - 1.0 / (1.0 + ex((6.835e-14 - fn) / 13.67e-16)),
                       1.91 + 2.09 / (1.0 + ex((3.4175e-13 - fn) / 13.67e-16)))
        return tv, tc, tf_

    # ---- scalar constants of the step (:460-493), Python-scalar sub-expressions folded in float64
    def _fold_constants(self):
        s = self._scalar_param
        return [(s('_R_gas') * s('_T')) / s('_F'), s('_Ko'),
                s('_GACh') * (10.0 / (1.0 + (9.13652 / (pow(s('_ACh'), 0.477811))))),
                s('_Gto'), s('_factorGKur'), s('_GKr'), s('_GKs'), s('_GK1'), s('_factorGtr'), s('_tau_tr'),
                s('_factorGrel'), s('_k_rel'), s('_maxIup') / s('_maxCaup'), s('_C_dCa_rel'), s('_KmCsqn'),
                s('_F') * s('_Voli'), s('_C_dCaup'), s('_C_B1d'), s('_C_B1e'), s('_C_Fn1'), s('_C_Fn2'),
                self._f_Ca_rush_larsen_B, self._u_rush_larsen_B]
