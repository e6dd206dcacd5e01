"""TenTusscherPanfilov - mirror of gpuSolve/ionic/ten_tusscher_panfilov.py:36-523
(ten Tusscher & Panfilov 2006 human ventricular model; 22 state arrays of which four are the
per-node conductances GCaL, GKr, GKs, Gto; three lookup tables; cell types EPI / ENDO / MCELL).

Attribute names, defaults, initial state and the table grids are the reference's; the step runs
in libpdx (csrc/pdx_ionic_lut.cu: ttp_update).  Table columns are laid out for the kernel
(include/pdx.h), not in the reference's column order.
"""
from math import sqrt

import numpy as np

from ._lut import HostTable, LutIonicModel
from ... import _lib as L


class TenTusscherPanfilov(LutIonicModel):
    _MODEL_ID = L.MODEL_TTP
    # libpdx state order (include/pdx.h) = the reference's declaration order (:196-217)
    _STATE_ATTRS = ('_GCaL', '_GKr', '_GKs', '_Gto', '_CaSR', '_CaSS', '_Cai', '_F_state', '_F2_state', '_FCaSS_state',
                    '_H_state', '_J_state', '_Ki', '_M_state', '_Nai', '_R_state', '_R_bar', '_S_state', '_Xr1_state',
                    '_Xr2_state', '_Xs_state', '_D_state')
    _TABLE_ATTRS = ('_V_tab_lut', '_CaSS_tab', '_VEk_tab')

    def __init__(self, dt=0.0, n_nodes=0, cell_type="EPI"):
        super().__init__(dt, n_nodes)
        self._cell_type = cell_type if cell_type is not None else "EPI"
        # constants (:53-81)
        self._CaSR_init, self._CaSS_init, self._Cai_init = 1.3, 0.00007, 0.00007
        self._EC, self._F2_init, self._FCaSS_init, self._F_state_init = 1.5, 1.0, 1.0, 1.0
        self._H_init, self._J_init, self._Ki_init, self._M_init, self._Nai_init = 0.75, 0.75, 138.3, 0.0, 7.67
        self._O_init, self._R_bar_init, self._R_init, self._S_init, self._V_init = 0.0, 1.0, 0.0, 1.0, -86.2
        self._Vxfer, self._Xr1_init, self._Xr2_init, self._Xs_init, self._D_init = 0.0038, 0.0, 1.0, 0.0, 0.0
        self._k1_, self._k2_, self._k3, self._k4, self._maxsr, self._minsr = 0.15, 0.045, 0.060, 0.005, 2.5, 1.0
        # parameters (:83-127)
        self._Bufc, self._Bufsr, self._Bufss, self._CAPACITANCE, self._Cao = 0.2, 10.0, 0.4, 0.185, 2.0
        self._D_CaL_off, self._Fconst, self._GK1, self._GNa = 0.0, 96485.3415, 5.405, 14.838
        self._GbCa, self._GbNa, self._GpCa, self._GpK = 0.000592, 0.00029, 0.1238, 0.0146
        self._Kbufc, self._Kbufsr, self._Kbufss = 0.001, 0.3, 0.00025
        self._KmCa, self._KmK, self._KmNa, self._KmNai, self._Ko = 1.38, 1.0, 40.0, 87.5, 5.4
        self._KpCa, self._Kup, self._Nao, self._Rconst, self._T = 0.0005, 0.00025, 140.0, 8314.472, 310.0
        self._Vc, self._Vleak, self._Vmaxup, self._Vrel = 0.016404, 0.00036, 0.006375, 0.102
        self._Vsr, self._Vss, self._knaca, self._knak, self._ksat = 0.001094, 0.00005468, 1000.0, 2.724, 0.1
        self._n, self._pKNa, self._scl_tau_f, self._vHalfXs, self._xr2_off = 0.35, 0.03, 1.0, 5.0, 0.0
        ct = self._cell_type
        self._GCaL_init, self._GKr_init = 0.00003980, 0.153
        self._GKs_init = 0.392 if ct == "EPI" else 0.098 if ct == "MCELL" else 0.392
        self._Gto_init = 0.294 if ct == "EPI" else 0.294 if ct == "MCELL" else 0.073
        # table geometry (:169-188)
        self._CaSS_T_mn, self._CaSS_T_mx, self._CaSS_T_res = 0.00001, 10.0, 0.00001
        self._V_T_mn, self._V_T_mx, self._V_T_res = -800.0, 800.0, 0.05
        self._VEk_T_mn, self._VEk_T_mx, self._VEk_T_res = -800.0, 800.0, 0.05
        self._V_tab_lut = self._CaSS_tab = self._VEk_tab = None
        for a in self._STATE_ATTRS:
            setattr(self, a, None)

    @property
    def _STATE_INIT(self):
        return (self._GCaL_init, self._GKr_init, self._GKs_init, self._Gto_init, self._CaSR_init, self._CaSS_init,
                self._Cai_init, self._F_state_init, self._F2_init, self._FCaSS_init, self._H_init, self._J_init,
                self._Ki_init, self._M_init, self._Nai_init, self._R_init, self._R_bar_init, self._S_init,
                self._Xr1_init, self._Xr2_init, self._Xs_init, self._D_init)

    # ---- tables (:234-371) -------------------------------------------------------------------
    def _build_tables(self):
        s = self._scalar_param
        dt = self._dt
        ex = np.exp
        Nao, Cao = s('_Nao'), s('_Cao')
        KmNai = s('_KmNai')
        Nao3 = Nao * Nao * Nao
        F_RT = 1.0 / ((s('_Rconst') * s('_T')) / s('_Fconst'))
        pmf_INaCa = (s('_knaca') * (1.0 / (KmNai * KmNai * KmNai + Nao3))) * (1.0 / (s('_KmCa') + Cao))
        n_, ksat = s('_n'), s('_ksat')

        tv = HostTable(self._V_T_mn, self._V_T_mx, self._V_T_res)
        V = tv.grid
        # D
        tv.add_rl_tau(dt, 1.0 / (1.0 + ex(((-8.0 + s('_D_CaL_off')) - V) / 7.5)),
                      ((1.4 / (1.0 + ex((-35.0 - V) / 13.0))) + 0.25) * (1.4 / (1.0 + ex((V + 5.0) / 5.0))) +
                      1.0 / (1.0 + ex((50.0 - V) / 20.0)))
        # F2
        tv.add_rl_tau(dt, (0.67 / (1.0 + ex((V + 35.0) / 7.0))) + 0.33,
                      562.0 * ex(-(V + 27.0) * (V + 27.0) / 240.0) + 31.0 / (1.0 + ex((25.0 - V) / 10.0)) +
                      80.0 / (1.0 + ex((V + 30.0) / 10.0)))
        # F
        tF = (1102.5 * ex(-(V + 27.0) * (V + 27.0) / 225.0) + 200.0 / (1.0 + ex((13.0 - V) / 10.0)) +
              ((180.0 / (1.0 + ex((V + 30.0) / 10.0))) + 20.0))
        tv.add_rl_tau(dt, 1.0 / (1.0 + ex((V + 20.0) / 7.0)), np.where(V > 0.0, tF * s('_scl_tau_f'), tF))
        # H, J share their steady state
        H_inf = 1.0 / ((1.0 + ex((V + 71.55) / 7.43)) * (1.0 + ex((V + 71.55) / 7.43)))
        aa_H = np.where(V >= -40.0, 0.0, 0.057 * ex(-(V + 80.0) / 6.8))
        bb_H = np.where(V >= -40.0, 0.77 / (0.13 * (1.0 + ex(-(V + 10.66) / 11.1))),
                        2.7 * ex(0.079 * V) + 3.1e5 * ex(0.3485 * V))
        tv.add_rl_tau(dt, H_inf, 1.0 / (aa_H + bb_H))
        aa_J = np.where(V >= -40.0, 0.0, ((-2.5428e4 * ex(0.2444 * V) - 6.948e-6 * ex(-0.04391 * V)) * (V + 37.78)) /
                        (1.0 + ex(0.311 * (V + 79.23))))
        bb_J = np.where(V >= -40.0, (0.6 * ex(0.057 * V)) / (1.0 + ex(-0.1 * (V + 32.0))),
                        (0.02424 * ex(-0.01052 * V)) / (1.0 + ex(-0.1378 * (V + 40.14))))
        tv.add_rl_tau(dt, H_inf, 1.0 / (aa_J + bb_J))
        # M
        tv.add_rl_tau(dt, 1.0 / ((1.0 + ex((-56.86 - V) / 9.03)) * (1.0 + ex((-56.86 - V) / 9.03))),
                      (1.0 / (1.0 + ex((-60.0 - V) / 5.0))) *
                      ((0.1 / (1.0 + ex((V + 35.0) / 5.0))) + (0.10 / (1.0 + ex((V - 50.0) / 200.0)))))
        # R, S
        tv.add_rl_tau(dt, 1.0 / (1.0 + ex((20.0 - V) / 6.0)), 9.5 * ex(-(V + 40.0) * (V + 40.0) / 1800.0) + 0.8)
        if self._cell_type == "ENDO":
            tau_S = 1000.0 * ex(-(V + 67.0) * (V + 67.0) / 1000.0) + 8.0
        else:
            tau_S = 85.0 * ex(-(V + 45.0) * (V + 45.0) / 320.0) + 5.0 / (1.0 + ex((V - 20.0) / 5.0)) + 3.0
        tv.add_rl_tau(dt, 1.0 / (1.0 + ex((V + 20.0) / 5.0)), tau_S)
        # Xr1, Xr2, Xs
        tv.add_rl_tau(dt, 1.0 / (1.0 + ex((-26.0 - V) / 7.0)),
                      (450.0 / (1.0 + ex((-45.0 - V) / 10.0))) * (6.0 / (1.0 + ex((V + 30.0) / 11.5))))
        tv.add_rl_tau(dt, 1.0 / (1.0 + ex((V - (-88.0 + s('_xr2_off'))) / 24.0)),
                      (3.0 / (1.0 + ex((-60.0 - V) / 20.0))) * (1.12 / (1.0 + ex((V - 60.0) / 20.0))))
        tv.add_rl_tau(dt, 1.0 / (1.0 + ex((-s('_vHalfXs') - V) / 14.0)),
                      (1400.0 / np.sqrt(1.0 + ex((5.0 - V) / 6.0))) * (1.0 / (1.0 + ex((V - 35.0) / 15.0))) + 80.0)
        den = pmf_INaCa / (1.0 + ksat * ex((n_ - 1.0) * V * F_RT))
        tv.add(0.25 * ex(2.0 * (V - 15.0) * F_RT),
               (den * Cao) * ex(n_ * V * F_RT),
               (den * ex((n_ - 1.0) * V * F_RT)) * Nao3 * 2.5,
               1.0 / (1.0 + 0.1245 * ex(-0.1 * V * F_RT) + 0.0353 * ex(-V * F_RT)),
               1.0 / (1.0 + ex((25.0 - V) / 5.98)))

        tc = HostTable(self._CaSS_T_mn, self._CaSS_T_mx, self._CaSS_T_res)
        c = tc.grid
        tc.add_rl_tau(dt, (0.6 / (1.0 + (c / 0.05) * (c / 0.05))) + 0.4, (80.0 / (1.0 + (c / 0.05) * (c / 0.05))) + 2.0)

        te = HostTable(self._VEk_T_mn, self._VEk_T_mx, self._VEk_T_res)
        E = te.grid
        a_K1 = 0.1 / (1.0 + ex(0.06 * (E - 200.0)))
        b_K1 = (3.0 * ex(0.0002 * (E + 100.0)) + ex(0.1 * (E - 10.0))) / (1.0 + ex(-0.5 * E))
        te.add(a_K1 / (a_K1 + b_K1))
        return tv, tc, te

    # ---- scalar constants of the step (:440-497), Python-scalar sub-expressions folded in float64
    def _fold_constants(self):
        s = self._scalar_param
        RTONF = (s('_Rconst') * s('_T')) / s('_Fconst')
        F_RT = 1.0 / RTONF
        Vc, Vss, Vsr, Fc, Cm = s('_Vc'), s('_Vss'), s('_Vsr'), s('_Fconst'), s('_CAPACITANCE')
        Ko = s('_Ko')
        return [RTONF, 0.5 * RTONF, s('_Cao'), Ko, Ko + s('_pKNa') * s('_Nao'), s('_pKNa'), s('_Nao'), s('_GpCa'),
                s('_KpCa'), Fc, F_RT, 0.5 * F_RT, sqrt(Ko / 5.4), s('_GNa'), s('_knak') * (Ko / (Ko + s('_KmK'))),
                s('_KmNa'), s('_GbCa'), s('_GbNa'), s('_GpK'), s('_GK1'), s('_Vleak'), s('_Vmaxup'),
                s('_Kup') * s('_Kup'), s('_Vxfer'), (1.0 / (Vc * Fc)) * Cm, s('_maxsr'), s('_maxsr') - s('_minsr'),
                s('_EC'), 1.0 / (2.0 * Vc * Fc), Cm, Vsr / Vc, s('_Bufc') * s('_Kbufc'), s('_Kbufc'), s('_k4'),
                s('_k2_'), s('_k1_'), s('_k3'), s('_Vrel'), s('_Bufsr') * s('_Kbufsr'), s('_Kbufsr'), Vc / Vss,
                Vsr / Vss, 1.0 / (2.0 * Vss * Fc), s('_Bufss') * s('_Kbufss'), s('_Kbufss')]
