"""BASELINE.json's configurations at their REAL sizes against the oracle (VERDICT r01 "what's weak" 1).

* config 1 - Tests/FEM/Fenton/fenton.py on the synthetic 251 x 251 stand-in of the missing mesh (63 001 nodes,
  125 000 P1 triangles, RCM on, S1 at 0 ms, S2 at 210 ms, dt 0.1, maxiter N//2, toll 1e-5) through
  ``MonodomainSolver`` (device CSR, persistent PCG kernel) vs ``oracle.fem.MonodomainOracle``.
* config 2 - 2-D Fenton-Karma on 1024 x 1024 with the S1-S2 protocol (S2 on a boolean quadrant mask) through the
  SM-resident multi-step kernel vs ``oracle.fd.fd_run``.

Tolerances (stated, as BASELINE.json asks): FD rel-L2 <= 1e-6 EXACT / 1e-5 FAST; activation times (first upward
crossing of -30 mV, Tests/CICD/unit/test_mms_1d.py:91-101) within one dt.  PCG frames: the reference stops CG at the
ABSOLUTE |r|^2 < 1e-10 checked every 5 iterations, so two correct implementations agree to 5e-5 while their
5-or-10-iteration decisions coincide (measured on two CPU variants of the oracle that differ only in the
accumulation type of the dot products: <= 2e-6 over the first 500 steps) and drift to ~1e-3 once one of them stops a
check interval earlier (steps 600-1500 of the shipped protocol); activation times stay within one dt throughout.
"""
import numpy as np
import pytest
import torch

from helpers import rel_l2
from oracle import fd as ofd
from oracle import fem as ofem
from oracle import ionic as oionic
from oracle import stimulus as ostim

pytestmark = pytest.mark.gpu
VTH = -30.0


# ------------------------------------------------------------------------------------ config 1
def _config1(tmp_path):
    """The shipped example's set-up calls (Tests/FEM/Fenton/fenton.py:60-117) on the mirror and on the oracle."""
    from pdensorflow_b200.gpuSolve.entities.synthetic import save_pkl, triangulated_square
    from pdensorflow_b200.gpuSolve.ionic import Fenton4v
    from pdensorflow_b200.gpuSolve.physics import MonodomainSolver
    P, E, F = triangulated_square()                       # 251 x 251 nodes on [0, 10]^2
    assert P.shape[0] == 63001 and E['Trias'].shape[0] == 125000
    mesh = str(tmp_path / 'triangulated_square.pkl')
    save_pkl(mesh, P, E, F)
    sig = {1: 1e-3, 2: 1e-3, 3: 1e-3, 4: 1e-3}
    dt = 0.1
    model = MonodomainSolver(Fenton4v(dt=dt), {'mesh_file_name': mesh, 'use_renumbering': True, 'dt': dt,
                                               'dt_per_plot': 10, 'Tend': 1000})
    model.add_element_material_property('sigma_l', 'region', sig)
    model.add_element_material_property('sigma_t', 'region', sig)
    model.add_material_function('mass', lambda *a: None)
    model.assign_nodal_properties()
    model.assemble_matrices()
    Lx, Ly = P[:, 0].max(), P[:, 1].max()
    S1 = P[:, 0] < 0.05 * Lx
    S2 = np.logical_and(P[:, 0] < Lx, P[:, 1] < 0.5 * Ly)
    props = [{'tstart': 0.0, 'nstim': 1, 'period': 800, 'duration': np.max([0.4, dt]), 'intensity': 60.0, 'name': 'S1'},
             {'tstart': 210.0, 'nstim': 1, 'period': 800, 'duration': np.max([0.4, dt]), 'intensity': 60.0, 'name': 'S2'}]
    model.set_initial_condition()
    model.add_stimulus(S1, props[0])
    model.add_stimulus(S2, props[1])
    model.solver().set_maxiter(P.shape[0] // 2)
    model.finalize_for_run()
    orc = ofem.MonodomainOracle(oionic.Fenton4v(), P, E, F, sig, sig, dt, use_renumbering=True, maxiter=P.shape[0] // 2)
    orc.add_stimulus(S1, dict(props[0]))
    orc.add_stimulus(S2, dict(props[1]))
    orc.finalize_for_run()
    return model, orc, P


def test_config1_full_size_shipped_protocol(cuda, tmp_path):
    """2150 semi-implicit steps = S1, the planar front crossing the sheet, S2 at 210 ms (steps 2099-2103) and the
    first 5 ms of the cross-field response: frames vs the oracle, CG iteration counts, activation times."""
    model, orc, P = _config1(tmp_path)
    n = P.shape[0]
    nsteps, dt = 2150, 0.1
    lat_g = torch.full((n,), -1, dtype=torch.int32, device='cuda')
    lat_o = np.full(n, -1, dtype=np.int32)
    prev_g = model.U().reshape(-1).clone()
    prev_o = orc.U_user()[:, 0].copy()
    ctime, errs, it_diff = 0.0, {}, 0
    for i in range(nsteps):
        ctime += dt
        model.step(ctime)
        orc.step(ctime)
        cur_g = model.U().reshape(-1)
        cur_o = orc.U_user()[:, 0]
        lat_g[(prev_g < VTH) & (cur_g >= VTH) & (lat_g < 0)] = i
        lat_o[(prev_o < VTH) & (cur_o >= VTH) & (lat_o < 0)] = i
        prev_g, prev_o = cur_g.clone(), cur_o.copy()
        it_g = model.solver().niters()
        assert it_g % 5 == 0
        it_diff += int(it_g != orc.niters[-1])
        if (i + 1) % 50 == 0:
            errs[i + 1] = rel_l2(cur_g.cpu().numpy(), cur_o)
    early = max(v for k, v in errs.items() if k <= 400)
    assert early <= 5e-5, errs
    assert max(errs.values()) <= 5e-3, errs                      # stopping-rule drift of the long run (see header)
    assert it_diff <= nsteps // 4, it_diff                       # the 5-vs-10 decision differs on a minority of the steps
    lat_g = lat_g.cpu().numpy()
    assert (lat_o >= 0).all() and np.array_equal(lat_g >= 0, lat_o >= 0)       # every node activates (S1 front or S2)
    assert np.abs(lat_g - lat_o).max() <= 1                                   # within one dt
    # the nightly test's shape assertions (Tests/CICD/nightly/test_mms_2d_regression.py:151-184) on the S1 front:
    # monotone activation along x in the upper half (never touched by S2) and a positive, uniform conduction velocity
    x, y = P[:, 0], P[:, 1]
    Lx = x.max()
    band = (y > 0.6 * y.max()) & (x >= 0.2 * Lx) & (x <= 0.8 * Lx)
    xu = np.unique(np.round(x[band], 6))
    col = np.array([lat_g[band & (np.abs(x - v) < 1e-6)].mean() for v in xu])
    assert np.all(np.diff(col) > 0.0)
    cv_g = 1.0 / float(np.polyfit(x[band], lat_g[band] * dt, 1)[0])
    cv_o = 1.0 / float(np.polyfit(x[band], lat_o[band] * dt, 1)[0])
    assert cv_g > 0 and abs(cv_g - cv_o) / cv_o < 1e-3, (cv_g, cv_o)
    Ug = model.U().cpu().numpy()
    assert Ug.min() > -81.0 and Ug.max() < 61.0


# ------------------------------------------------------------------------------------ config 2
CFG2 = dict(shape=(1024, 1024), dt=0.1, diff=0.8, spacing=(1.0, 1.0), nsteps=320, s2_step=200)


@pytest.fixture(scope='module')
def config2_oracle():
    """Oracle run of config 2 (Tests/FD/Fenton/fenton.py:119-126,150-159 transcribed to 2-D): S1 edge
    rows at +20, S2 = max(U, 60) on the boolean quadrant mask.  The shipped S2 time is step 2000; here it fires at
    step 200 so that the 320-step window holds the S1 front, the S2 response and their interaction."""
    c = CFG2
    shape = c['shape']
    U = np.full(shape, -80.0, dtype=np.float32)
    U[0:2, :] = 20.0
    mask = np.zeros(shape, dtype=np.float32)
    mask[:shape[0] // 2, :shape[1] // 2] = 1.0
    s2 = ostim.Stimulus({'tstart': c['s2_step'] * c['dt'], 'nstim': 1, 'period': 800, 'duration': c['dt'],
                         'intensity': 60.0})
    s2.set_stimregion(mask)
    m = oionic.Fenton4v()
    m._dt = c['dt']
    m.initialize_state_variables(U)
    Uo, act = ofd.fd_run(m, U.copy(), c['nsteps'], c['dt'], c['diff'], c['spacing'], s2=s2, act_threshold=VTH)
    fired = [i for i in range(c['nsteps']) if s2.stimulate_tissue_timestep(i, c['dt'])]
    return {'U0': U, 'mask': mask, 'U': Uo, 'act': act, 'fired': fired, 'gates': (m._V_state, m._W_state, m._S_state)}


@pytest.mark.parametrize('math,tol', [('exact', 1e-6), ('fast', 1e-5)])
def test_config2_resident_kernel_vs_oracle(cuda, config2_oracle, math, tol):
    """The SM-resident kernel in the launches the protocol allows (one launch up to the S2 step, one after it)."""
    from pdensorflow_b200.fd import FDStepper
    c, o = CFG2, config2_oracle
    assert o['fired'] == [c['s2_step']]
    st = FDStepper(c['shape'], c['spacing'], c['dt'], c['diff'], model='fenton4v', math=math, kernel='resident')
    assert st.kernel == 'resident'
    st.set_U(o['U0'])
    mask_d = torch.from_numpy(o['mask']).cuda()
    n0 = st.launches()
    st.step(c['s2_step'] + 1)                         # steps 0 .. s2_step, then the stimulus (fenton.py:154-156)
    st.apply_stimulus(mask_d, 60.0)
    st.step(c['nsteps'] - c['s2_step'] - 1)
    torch.cuda.synchronize()
    assert st.launches() - n0 == 2                    # two resident launches for 320 steps
    assert rel_l2(st.U().cpu().numpy(), o['U']) <= tol
    for a, b in zip(st.states, o['gates']):
        assert rel_l2(a.cpu().numpy(), b) <= 10 * tol


@pytest.mark.parametrize('math', ['exact', 'fast'])
def test_config2_activation_times_within_one_dt(cuda, config2_oracle, math):
    """Same protocol stepped in resident launches of ONE step so that the first upward crossing of -30 mV can be
    recorded per step: every node activates on the same step as the oracle or one step away."""
    from pdensorflow_b200.fd import FDStepper
    c, o = CFG2, config2_oracle
    st = FDStepper(c['shape'], c['spacing'], c['dt'], c['diff'], model='fenton4v', math=math, kernel='resident')
    st.set_U(o['U0'])
    mask_d = torch.from_numpy(o['mask']).cuda()
    act = torch.full(c['shape'], -1, dtype=torch.int32, device='cuda')
    prev = st.U().clone()
    for i in range(c['nsteps']):
        st.step(1)
        if i == c['s2_step']:
            st.apply_stimulus(mask_d, 60.0)
        cur = st.U()
        act[(prev < VTH) & (cur >= VTH) & (act < 0)] = i
        prev = cur.clone()
    act = act.cpu().numpy()
    assert (o["act"] >= 0).sum() > 0.25 * act.size
    assert np.array_equal(act >= 0, o['act'] >= 0)
    assert np.abs(act - o['act']).max() <= 1
    assert rel_l2(st.U().cpu().numpy(), o['U']) <= (1e-6 if math == 'exact' else 1e-5)
