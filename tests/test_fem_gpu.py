"""GPU parity of the FEM path (through the drop-in classes and the C ABI) against the oracle, plus the
reference's own 1-D mMS conduction-velocity test (Tests/CICD/unit/test_mms_1d.py) ported to the mirror."""
import ctypes as C
import os
import pickle

import numpy as np
import pytest
import torch

from helpers import rel_l2
from oracle import fem as ofem
from oracle import ionic as oionic

pytestmark = pytest.mark.gpu


def _square(n=31):
    from pdensorflow_b200.gpuSolve.entities.synthetic import triangulated_square
    return triangulated_square(n, n, 10.0, 10.0)


def _system(n=31, dt=0.1, sigma=1e-3):
    P, E, F = _square(n)
    A = ofem.assemble(P, E, F, {1: sigma, 2: sigma, 3: sigma, 4: sigma}, {1: sigma, 2: sigma, 3: sigma, 4: sigma})
    val = ofem.csr_axpby(A['mass'], 1.0, A['stiffness'], dt)
    return A, val


def test_spmv_matches_scipy(cuda):
    from pdensorflow_b200.gpuSolve.matrices import CSRMatrix
    A, val = _system()
    n = A['n']
    x = np.random.default_rng(0).standard_normal(n).astype(np.float32)
    M = CSRMatrix(A['rowptr'], A['col'], val, n)
    y = M.matvec(torch.from_numpy(x).cuda()).cpu().numpy()
    yo = ofem.spmv(A['rowptr'], A['col'], val, x)
    np.testing.assert_allclose(y, yo, rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize('gold_kind', ['ones', 'randn'])
def test_pcg_recovers_gold_solution(cuda, gold_kind):
    """test_conjgrad.py:117-162 on the synthetic square: A = M + K, zero guess, toll = 1e-6 ||b||."""
    from pdensorflow_b200.gpuSolve.linearsolvers import ConjGrad, JacobiPrecond
    from pdensorflow_b200.gpuSolve.matrices import CSRMatrix
    P, E, F = _square(41)
    A = ofem.assemble(P, E, F, None, None)
    val = ofem.csr_axpby(A['mass'], 1.0, A['stiffness'], 1.0)
    n = A['n']
    gold = np.ones(n, np.float32) if gold_kind == 'ones' else np.random.default_rng(0).standard_normal(n).astype(np.float32)
    b = ofem.spmv(A['rowptr'], A['col'], val, gold)
    M = CSRMatrix(A['rowptr'], A['col'], val, n)
    pc = JacobiPrecond()
    pc.build_preconditioner(*M.coo(), n)
    cg = ConjGrad({'maxiter': 1000, 'toll': 1e-6 * float(np.linalg.norm(b)), 'precond': pc})
    cg.set_matrix(M)
    cg.set_X0(torch.zeros(n, 1, device='cuda'))
    cg.set_RHS(torch.from_numpy(b[:, None]).cuda())
    cg.solve()
    x = cg.X().cpu().numpy().ravel()
    assert cg.niters() < 1000 and cg.niters() % 5 == 0
    assert np.linalg.norm(x - gold) / np.linalg.norm(gold) < 1e-3
    assert np.linalg.norm(b - ofem.spmv(A['rowptr'], A['col'], val, x)) / np.linalg.norm(b) < 1e-2
    # same stopping rule as the oracle: iteration counts agree up to one check interval
    minv = ofem.jacobi(A['rowptr'], A['col'], val, n)
    xo, nit_o, res_o = ofem.pcg(A['rowptr'], A['col'], val, minv, b, np.zeros(n, np.float32), 1000,
                                1e-6 * float(np.linalg.norm(b)))
    assert abs(cg.niters() - nit_o) <= 5
    assert np.linalg.norm(x - xo) / np.linalg.norm(xo) < 1e-4


def _build_model(n, rcm, explicit=False, math='exact'):
    from pdensorflow_b200.gpuSolve.ionic import Fenton4v
    from pdensorflow_b200.gpuSolve.physics import MonodomainSolver
    P, E, F = _square(n)
    sig = {1: 1e-3, 2: 1e-3, 3: 1e-3, 4: 1e-3}
    ionic = Fenton4v(dt=0.1)
    ionic.math_mode = math
    model = MonodomainSolver(ionic, {'use_renumbering': rcm, 'dt': 0.1, 'dt_per_plot': 10, 'Tend': 1000,
                                     'explicit_lumped': explicit})
    model.domain().set_mesh(P, E, F)
    model.add_element_material_property('sigma_l', 'region', sig)
    model.add_element_material_property('sigma_t', 'region', sig)
    model.add_material_function('mass', lambda *a: None)
    model.assign_nodal_properties()
    model.assemble_matrices()
    Lx, Ly = P[:, 0].max(), P[:, 1].max()
    S1 = P[:, 0] < 0.05 * Lx
    S2 = np.logical_and(P[:, 0] < Lx, P[:, 1] < 0.5 * Ly)
    model.set_initial_condition()
    model.add_stimulus(S1, {'tstart': 0.0, 'nstim': 1, 'period': 100, 'duration': np.max([0.4, 0.1]), 'intensity': 60.0})
    model.add_stimulus(S2, {'tstart': 3.0, 'nstim': 1, 'period': 100, 'duration': np.max([0.4, 0.1]), 'intensity': 60.0})
    model.solver().set_maxiter(P.shape[0] // 2)
    model.finalize_for_run()
    orc = ofem.MonodomainOracle(oionic.Fenton4v(), P, E, F, sig, sig, 0.1, use_renumbering=rcm, maxiter=P.shape[0] // 2)
    orc.add_stimulus(S1, {'tstart': 0.0, 'nstim': 1, 'period': 100, 'duration': np.max([0.4, 0.1]), 'intensity': 60.0})
    orc.add_stimulus(S2, {'tstart': 3.0, 'nstim': 1, 'period': 100, 'duration': np.max([0.4, 0.1]), 'intensity': 60.0})
    orc.finalize_for_run()
    return model, orc


@pytest.mark.parametrize('rcm', [False, True])
def test_monodomain_solver_steps_match_oracle(cuda, rcm):
    """Config 1 pattern (Tests/FEM/Fenton/fenton.py) on a 31x31 stand-in mesh: S1 at t=0, S2 at 3 ms,
    60 semi-implicit steps.  Both sides stop PCG at ||r||^2 < 1e-10, so U agrees to ~1e-6 relative."""
    model, orc = _build_model(31, rcm)
    ctime = 0.0
    for i in range(60):
        ctime += 0.1
        model.step(ctime)
        orc.step(ctime)
    Ug = model.U().cpu().numpy()
    Uo = orc.U_user()
    assert Ug.shape == Uo.shape
    assert Uo.max() > -30.0                           # the stimuli fired
    assert rel_l2(Ug, Uo) <= 1e-5, rel_l2(Ug, Uo)
    Vg = model.ionic_state('_V_state').cpu().numpy()
    Vo = orc.ionic._V_state[orc.ren['iperm']] if rcm else orc.ionic._V_state
    assert rel_l2(Vg, Vo) <= 1e-5
    assert model.solver().niters() % 5 == 0 and abs(model.solver().niters() - orc.niters[-1]) <= 5


@pytest.mark.parametrize('explicit', [False, True])
def test_graph_replayed_steps_are_bitwise_the_eager_ones(cuda, explicit, monkeypatch):
    """MonodomainSolver.step replays a CUDA graph from the third visit of a (buffer, forcing, parameters, settings) key on
    (HeatSolver._graph_step).  Same kernels with the same arguments: bitwise the eager loop, through two stimuli (new
    forcing fields), a set_parameter and a set_maxiter in mid-run (both must invalidate the captured steps)."""
    out = {}
    for mode in ('1', '0'):
        monkeypatch.setenv('PDX_STEP_GRAPH', mode)
        model, _ = _build_model(31, True, explicit=explicit)
        ctime, its = 0.0, []
        for i in range(70):
            ctime += 0.1
            if i == 45:
                model.ionic_model().set_parameter('tau_d', 0.07)
            if i == 55:
                model.solver().set_maxiter(200)
            model.step(ctime)
            if not explicit and i % 10 == 9:
                its.append(model.solver().niters())
        ngraphs = sum(1 for v in model._step_graphs.values() if isinstance(v, tuple))
        assert (ngraphs > 0) == (mode == '1')
        out[mode] = (model.U().cpu().numpy(), model.ionic_state('_V_state').cpu().numpy(), its)
    assert np.array_equal(out['1'][0], out['0'][0]) and np.array_equal(out['1'][1], out['0'][1])
    assert out['1'][2] == out['0'][2]
    assert out['1'][0].max() > -30.0


def test_explicit_lumped_step_matches_oracle(cuda):
    """Row E (no reference implementation): explicit mass-lumped step vs its CPU restatement."""
    model, _ = _build_model(31, False, explicit=True, math='exact')
    P, E, F = _square(31)
    sig = {1: 1e-3, 2: 1e-3, 3: 1e-3, 4: 1e-3}
    A = ofem.assemble(P, E, F, sig, sig)
    n = A['n']
    from scipy.sparse import csr_matrix
    ml = np.asarray(csr_matrix((A['mass'], A['col'], A['rowptr']), shape=(n, n)).sum(axis=1)).ravel().astype(np.float32)
    mlinv = (np.float32(1.0) / ml).astype(np.float32)
    ion = oionic.Fenton4v(); ion._dt = 0.1
    U = np.full(n, -80.0, np.float32)
    ion.initialize_state_variables(U[:, None])
    from oracle.stimulus import Stimulus as OStim
    Lx, Ly = P[:, 0].max(), P[:, 1].max()
    stims = []
    for mask, t0 in ((P[:, 0] < 0.05 * Lx, 0.0), (np.logical_and(P[:, 0] < Lx, P[:, 1] < 0.5 * Ly), 3.0)):
        st = OStim({'tstart': t0, 'nstim': 1, 'period': 100, 'duration': np.max([0.4, 0.1]), 'intensity': 60.0})
        st.set_stimregion(mask)
        stims.append(st)
    ctime = 0.0
    for i in range(40):
        ctime += 0.1
        model.step(ctime)
        I0 = np.zeros(n, np.float32)
        for st in stims:
            I0 = I0 + st.stimApp(np.float32(ctime))[:, 0]
        U = ofem.explicit_lumped_step(ion, A['rowptr'], A['col'], A['stiffness'], mlinv, U, I0, 0.1)
    Ug = model.U().cpu().numpy().ravel()
    assert U.max() > -60.0
    assert rel_l2(Ug, U) <= 1e-5, rel_l2(Ug, U)


# ---------------------------------------------------------------------------------------------
# The reference's 1-D modified Mitchell-Schaeffer cable test, ported (test_mms_1d.py:44-176)
_SIGMA, _NELEM, _LENGTH, _DT, _TEND = 0.1, 150, 2.0, 0.05, 7.0
_VMIN, _VMAX, _VTH, _CV_TOL = -80.0, 20.0, -30.0, 0.10
_XSTIM = 0.05 * _LENGTH


def _local_activation_times(Vrec, times, vth):
    lat = np.full(Vrec.shape[1], np.nan)
    for j in range(Vrec.shape[1]):
        v = Vrec[:, j]
        k = int(np.argmax((v[:-1] < vth) & (v[1:] >= vth)))
        if (v[k] < vth) and (v[k + 1] >= vth):
            lat[j] = times[k] + (vth - v[k]) / (v[k + 1] - v[k]) * (times[k + 1] - times[k])
    return lat


@pytest.fixture(scope='module')
def mms_run(tmp_path_factory):
    from pdensorflow_b200.gpuSolve.ionic.mms2v import ModifiedMS2v
    from pdensorflow_b200.gpuSolve.physics import MonodomainSolver
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    mesh = str(tmp_path_factory.mktemp('mms') / 'cable.pkl')
    npt = _NELEM + 1
    Pts = np.zeros((npt, 3)); Pts[:, 0] = np.linspace(0.0, _LENGTH, npt)
    Edges = np.zeros((_NELEM, 3), dtype=np.int32)
    Edges[:, 0] = np.arange(0, _NELEM); Edges[:, 1] = np.arange(1, _NELEM + 1); Edges[:, 2] = 1
    with open(mesh, 'wb') as fout:
        pickle.dump({'Pts': Pts, 'Elems': {'Edges': Edges}, 'Fibres': None}, fout)
    ionic = ModifiedMS2v(dt=_DT)
    model = MonodomainSolver(ionic, {'mesh_file_name': mesh, 'use_renumbering': False, 'dt': _DT, 'dt_per_plot': 1,
                                     'Tend': _TEND})
    model.add_element_material_property('sigma_l', 'region', {1: _SIGMA})
    model.add_element_material_property('sigma_t', 'region', {1: _SIGMA})
    model.add_material_function('mass', lambda *a: None)
    model.add_material_function('stiffness', lambda *a: _SIGMA * np.eye(3))
    model.assemble_matrices()
    x = model.domain().Pts()[:, 0].copy()
    U0 = np.where(x < _XSTIM, _VMAX, _VMIN).astype(np.float32)
    model.set_initial_condition(U0)
    model.solver().set_maxiter(_NELEM)
    model.finalize_for_run()
    nt = model.nt()
    Vrec = np.empty((nt + 1, x.shape[0]), dtype=np.float32)
    times = np.empty(nt + 1)
    Vrec[0] = model.U().cpu().numpy().ravel(); times[0] = 0.0
    ctime = 0.0
    for i in range(nt):
        ctime += _DT
        model.step(ctime)
        Vrec[i + 1] = model.U().cpu().numpy().ravel(); times[i + 1] = ctime
    lat = _local_activation_times(Vrec, times, _VTH)
    mask = (x >= 0.4 * _LENGTH) & (x <= 0.8 * _LENGTH) & np.isfinite(lat)
    cv_meas = 1.0 / float(np.polyfit(x[mask], lat[mask], 1)[0])
    u_crit, tau_in = float(ionic.get_parameter('u_crit')), float(ionic.get_parameter('tau_in'))
    return {'x': x, 'lat': lat, 'Vrec': Vrec, 'cv_meas': cv_meas,
            'cv_analytic': 0.5 * (1.0 - 2.0 * u_crit) * np.sqrt(2.0 * _SIGMA / tau_in)}


def test_mms_1d_front_propagates(mms_run):
    x, lat, Vrec = mms_run['x'], mms_run['lat'], mms_run['Vrec']
    assert np.all(np.isfinite(Vrec))
    interior = (x >= 0.4 * _LENGTH) & (x <= 0.8 * _LENGTH)
    assert np.all(np.isfinite(lat[interior]))
    assert np.all(np.diff(lat[interior]) > 0.0)
    assert Vrec.min() > _VMIN - 1.0 and Vrec.max() < _VMAX + 1.0


def test_mms_1d_conduction_velocity(mms_run):
    relerr = abs(mms_run['cv_meas'] - mms_run['cv_analytic']) / mms_run['cv_analytic']
    assert relerr < _CV_TOL, relerr


def test_per_node_parameters_in_the_fem_step(cuda):
    """Tests/FEM/mMS/mMS.py:100-109 installs per-node tau arrays through set_parameter."""
    from pdensorflow_b200.gpuSolve.ionic.mms2v import ModifiedMS2v
    n = 300
    rng = np.random.default_rng(5)
    U = rng.uniform(-85, 25, (n, 1)).astype(np.float32)
    tau_in = rng.uniform(0.05, 0.3, (n, 1)).astype(np.float32)
    m = ModifiedMS2v(dt=0.05)
    m.set_parameter('tau_in', tau_in)
    Ud = torch.from_numpy(U).cuda()
    m.initialize_state_variables(Ud)
    dU = m.differentiate(Ud).cpu().numpy()
    o = oionic.ModifiedMS2v(dt=0.05)
    o.set_parameter('tau_in', tau_in)
    o.initialize_state_variables(U)
    dUo = o.differentiate(U)
    assert np.array_equal(dU, dUo)
    assert np.array_equal(m._H_state.cpu().numpy(), o._H_state)
