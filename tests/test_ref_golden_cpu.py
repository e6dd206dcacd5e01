"""The oracle against fixtures produced by the REFERENCE'S OWN SOURCE.

tests/golden/ref_*.npz were generated in the build container by importing the unmodified
reference (/root/reference/PDEnsorflow) over the NumPy TF facade (tests/golden/tf_facade.py,
generator tests/golden/make_ref_golden.py).  These tests pin every oracle function to those
outputs: bit-exact wherever the arithmetic is elementwise fp32, tight tolerances only where a
reduction order (PCG dots) is involved.
"""
import os

import numpy as np
import pytest

from oracle import fd as ofd
from oracle import fem as ofem
from oracle import ionic as oi
from oracle import ionic_lut as ol
from oracle.stimulus import Stimulus

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def gold(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def assert_same(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, what
    bad = a != b
    assert not bad.any(), '%s: %d/%d entries differ, max |diff| %.3e' % (what, bad.sum(), bad.size, np.abs(a - b).max())


# ------------------------------------------------------------------------------ operators
def test_operators_bit_exact_vs_reference_source():
    g = gold('ref_fd.npz')
    dx, dy, dz = (float(v) for v in g['lap_spacing'])
    X = g['lap_X']
    assert_same(ofd.laplace_homog_3d(X, dx, dy, dz), g['lap3_homog'], 'laplace_homog 3-D')
    assert_same(ofd.laplace_conv_homog_3d(X, dx, dy, dz), g['lap3_conv'], 'laplace_conv_homog')
    assert_same(ofd.laplace_heterog_3d(X, g['lap_D'], dx, dy, dz), g['lap3_heterog'], 'laplace_heterog 3-D')
    assert_same(ofd.laplace_heterog_aniso_3d(X, g['lap_DIFF2'], g['lap_AVEC'], dx, dy, dz), g['lap3_aniso'],
                'laplace_heterog_aniso')
    assert_same(ofd.laplace_homog_2d(g['lap2_X'], dx, dy), g['lap2_homog'], 'laplace_homog 2-D')
    assert_same(ofd.laplace_heterog_2d(g['lap2_X'], g['lap2_D'], dx, dy), g['lap2_heterog'], 'laplace_heterog 2-D')
    assert_same(ofd.enforce_boundary(g['eb_in']), g['eb_out'], 'enforce_boundary')


# ------------------------------------------------------------------------------ stimulus
@pytest.mark.parametrize('n', [0, 1, 2])
def test_stimulus_step_windows(n):
    g = gold('ref_fd.npz')
    tstart, nstim, period, duration, dt, intensity = (float(v) for v in g['stim%d_cfg' % n])
    s = Stimulus({'tstart': tstart, 'nstim': int(nstim), 'period': period, 'duration': duration, 'dt': dt,
                  'intensity': intensity})
    s.set_stimregion(np.ones((4, 1), dtype=np.float32))
    fired = [i for i in range(int(g['stim%d_nsteps' % n])) if s.stimulate_tissue_timestep(i, dt)]
    assert fired == list(g['stim%d_steps' % n])


@pytest.mark.parametrize('n', [0, 1])
def test_stimulus_float_time_windows(n):
    g = gold('ref_fd.npz')
    tstart, nstim, period, duration, intensity = (float(v) for v in g['stimt%d_cfg' % n])
    s = Stimulus({'tstart': tstart, 'nstim': int(nstim), 'period': period, 'duration': duration,
                  'intensity': intensity, 'name': 'S'})
    s.set_stimregion(np.ones((4, 1), dtype=np.float32))
    ctime, act = 0.0, []
    for i in range(400):
        if float(np.max(s.stimApp(np.float32(ctime)))) != 0.0:
            act.append(i)
        ctime += 0.1
    assert act == list(g['stimt%d_steps' % n])


# ------------------------------------------------------------------------------ ionic
@pytest.mark.parametrize('tag,cls,attrs', [('fk', oi.Fenton4v, ['_V_state', '_W_state', '_S_state']),
                                           ('mms', oi.ModifiedMS2v, ['_H_state']),
                                           ('ms', oi.MitchellSchaeffer2v, ['_H_state'])])
def test_simple_models_one_call_bit_exact(tag, cls, attrs):
    g = gold('ref_ionic.npz')
    U = g[tag + '_U']
    m = cls()
    m._dt = float(g[tag + '_dt'])
    m.initialize_state_variables(U)
    for k, a in enumerate(attrs):
        getattr(m, a)[...] = g['%s_s%d_in' % (tag, k)]
    dU = m.differentiate(U)
    assert_same(dU, g[tag + '_dU'], tag + ' dU')
    for k, a in enumerate(attrs):
        assert_same(getattr(m, a), g['%s_s%d_out' % (tag, k)], tag + a)


def test_mms_per_node_parameters_bit_exact():
    g = gold('ref_ionic.npz')
    U = g['mmsnodal_U']
    m = oi.ModifiedMS2v()
    m._dt = 0.05
    m.initialize_state_variables(U)
    m._H_state[...] = g['mmsnodal_H_in']
    m.set_parameter('tau_in', g['mmsnodal_tau_in'])
    m.set_parameter('tau_open', g['mmsnodal_tau_open'])
    assert_same(m.differentiate(U), g['mmsnodal_dU'], 'mMS nodal dU')
    assert_same(m._H_state, g['mmsnodal_H_out'], 'mMS nodal H')


def pace(model, cfg):
    u0, dt, nsteps, amp, stim_ms, period, ncell, spread, every = cfg
    nsteps, ncell, every = int(nsteps), int(ncell), int(every)
    U = np.linspace(u0, u0 + spread, ncell, dtype=np.float32).reshape(ncell, 1)
    model._dt = float(dt)
    model.initialize_state_variables(U)
    trace = [U.copy()]
    for i in range(nsteps):
        t = i * dt
        dU = model.differentiate(U)
        if (t % period) < stim_ms:
            dU = dU + amp
        U = (U + dt * dU).astype(np.float32)
        if (i + 1) % every == 0:
            trace.append(U.copy())
    return np.stack(trace)[:, :, 0]


PACED = [('fk', lambda: oi.Fenton4v()), ('mms', lambda: oi.ModifiedMS2v()), ('ms', lambda: oi.MitchellSchaeffer2v()),
         ('crn', lambda: ol.CourtemancheRamirezNattel()), ('ttp', lambda: ol.TenTusscherPanfilov()),
         ('ttpendo', lambda: ol.TenTusscherPanfilov(cell_type='ENDO')),
         ('ttpm', lambda: ol.TenTusscherPanfilov(cell_type='MCELL'))]


@pytest.mark.parametrize('tag,make', PACED, ids=[p[0] for p in PACED])
def test_paced_single_cell_trajectories(tag, make):
    """Tests/DEVTESTS/ionic/ionic.py protocol (shortened), 8 cells, thousands of steps through an
    action-potential upstroke: bit-exact with the reference source."""
    g = gold('ref_ionic.npz')
    tr = pace(make(), [float(v) for v in g['pace_%s_cfg' % tag]])
    ref = g['pace_%s_trace' % tag]
    assert ref.max() > 0.0 and ref.min() < -75.0            # the fixture really contains an action potential
    assert_same(tr, ref, 'paced ' + tag)


# reference attribute name -> oracle state key
CRN_MAP = {'_Ca_rel': 'Ca_rel', '_Ca_up': 'Ca_up', '_Cai': 'Cai', '_Ki': 'Ki', '_d_state': 'd', '_f_state': 'f',
           '_f_Ca_state': 'f_Ca', '_h_state': 'h', '_j_state': 'j', '_m_state': 'm', '_oa_state': 'oa',
           '_oi_state': 'oi', '_u_state': 'u', '_ua_state': 'ua', '_ui_state': 'ui', '_v_state': 'v', '_w_state': 'w',
           '_xr_state': 'xr', '_xs_state': 'xs'}
TTP_MAP = {'_GCaL': 'GCaL', '_GKr': 'GKr', '_GKs': 'GKs', '_Gto': 'Gto', '_CaSR': 'CaSR', '_CaSS': 'CaSS', '_Cai': 'Cai',
           '_F_state': 'F', '_F2_state': 'F2', '_FCaSS_state': 'FCaSS', '_H_state': 'H', '_J_state': 'J', '_Ki': 'Ki',
           '_M_state': 'M', '_Nai': 'Nai', '_R_state': 'R', '_R_bar': 'R_bar', '_S_state': 'S', '_Xr1_state': 'Xr1',
           '_Xr2_state': 'Xr2', '_Xs_state': 'Xs', '_D_state': 'D'}


@pytest.mark.parametrize('tag,cls,amap', [('crn', ol.CourtemancheRamirezNattel, CRN_MAP),
                                          ('ttp', ol.TenTusscherPanfilov, TTP_MAP)])
def test_lut_models_one_call_all_states(tag, cls, amap):
    g = gold('ref_ionic.npz')
    names = [str(n) for n in g[tag + '1_names']]
    assert sorted(names) == sorted(amap)
    U = g[tag + '1_U']
    m = cls(dt=0.02)
    m.initialize_state_variables(U)
    for n in names:
        m.s[amap[n]] = g['%s1_%s_in' % (tag, n)].copy()
    dU = m.differentiate(U)
    assert_same(dU, g[tag + '1_dU'], tag + ' dU')
    for n in names:
        assert_same(m.s[amap[n]], g['%s1_%s_out' % (tag, n)], tag + n)


# reference column index -> oracle column name
CRN_V_COLS = ['GKur', 'INaK', 'IbNa', 'd_A', 'd_B', 'f_A', 'f_B', 'h_A', 'h_B', 'j_A', 'j_B', 'm_A', 'm_B', 'oa_A', 'oa_B',
              'oi_A', 'oi_B', 'ua_A', 'ua_B', 'ui_A', 'ui_B', 'ICaL_v', 'NaCa_a', 'NaCa_b', 'IbCa_v', 'INa_v', 'w_A', 'w_B',
              'xr_A', 'xr_B', 'xs_A', 'xs_B']
CRN_CAI_COLS = ['Iup_ca', 'buf', 'IpCa_ca', 'conCa', 'fCa_A']
CRN_FN_COLS = ['u_A', 'v_A', 'v_B']
TTP_V_COLS = ['D_A', 'D_B', 'F2_A', 'F2_B', 'F_A', 'F_B', 'H_A', 'H_B', 'NaCa_A', 'NaCa_B', 'J_A', 'J_B', 'M_A', 'M_B',
              'R_A', 'R_B', 'S_A', 'S_B', 'Xr1_A', 'Xr1_B', 'Xr2_A', 'Xr2_B', 'Xs_A', 'Xs_B', 'a2', 'rec_iNaK', 'rec_ipK']
TTP_CASS_COLS = [None, 'FCaSS_A', 'FCaSS_B']          # column 0 (CaSS2) is never filled by the reference
TTP_VEK_COLS = ['rec_iK1']


def check_table(g, key, lut, cols):
    rows = g[key + '_rows']
    stride = int(g[key + '_stride'])
    assert tuple(g[key + '_shape']) == (lut.table.shape[0], len(cols))
    mine = np.zeros_like(rows)
    for c, name in enumerate(cols):
        if name is not None:
            mine[:, c] = lut.table[::stride, lut.names.index(name)]
    # NumPy's fp32 exp/expm1 kernels are SIMD-dispatch dependent: allow a few ulp across machines
    np.testing.assert_allclose(mine, rows, rtol=2e-6, atol=1e-30, err_msg=key)
    full = np.zeros((lut.table.shape[0], len(cols)))
    for c, name in enumerate(cols):
        if name is not None:
            full[:, c] = lut.table[:, lut.names.index(name)]
    np.testing.assert_allclose(full.sum(axis=0), g[key + '_colsum'], rtol=1e-6, atol=1e-12, err_msg=key + ' checksum')
    np.testing.assert_allclose(np.abs(full).sum(axis=0), g[key + '_colabs'], rtol=1e-6, atol=1e-12, err_msg=key)


def test_lookup_tables_match_reference_construct_tables():
    g = gold('ref_luts.npz')
    with np.errstate(all='ignore'):
        c = ol.CourtemancheRamirezNattel(dt=0.02)
        c.construct_tables()
        t = ol.TenTusscherPanfilov(dt=0.02)
        t.construct_tables()
        te = ol.TenTusscherPanfilov(dt=0.02, cell_type='ENDO')
        te.construct_tables()
    check_table(g, 'crn_V_tab', c.luts['V'], CRN_V_COLS)
    check_table(g, 'crn_Cai_tab', c.luts['Cai'], CRN_CAI_COLS)
    check_table(g, 'crn_fn_tab', c.luts['fn'], CRN_FN_COLS)
    check_table(g, 'ttp_V_tab_lut', t.luts['V'], TTP_V_COLS)
    check_table(g, 'ttp_CaSS_tab', t.luts['CaSS'], TTP_CASS_COLS)
    check_table(g, 'ttp_VEk_tab', t.luts['VEk'], TTP_VEK_COLS)
    check_table(g, 'ttpendo_V_tab_lut', te.luts['V'], TTP_V_COLS)


# ------------------------------------------------------------------------------ FD example
@pytest.mark.parametrize('tag,conv', [('fdrun', False), ('fdrunconv', True)])
def test_fd_example_end_to_end_bit_exact(tag, conv):
    """Tests/FD/Fenton/fenton.py run() on a small grid (all-ones S2 quirk included)."""
    g = gold('ref_fd.npz')
    W, H, D, dx, dy, dz, dt, per_plot, diff, samples, s2_time = (float(v) for v in g[tag + '_cfg'])
    shape = (int(W), int(H), int(D))
    U = np.full(shape, -80.0, dtype=np.float32)
    U[0:2] = 20.0
    s2_init = np.full(shape, -80.0, dtype=np.float32)
    s2_init[:shape[0] // 2, :shape[1] // 2, :] = 20.0
    m = oi.Fenton4v()
    m._dt = dt
    m.initialize_state_variables(U)
    s2 = Stimulus({'tstart': s2_time, 'nstim': 1, 'period': 800, 'duration': dt, 'dt': dt, 'intensity': 60.0})
    s2.set_stimregion(s2_init)
    frames = [U.copy()]
    for i in range(int(samples)):
        U = ofd.fd_run(m, U, 1, dt, diff, (dx, dy, dz), conv=conv, s2=s2, step0=i)
        if i % int(per_plot) == 0:
            frames.append(U.copy())
    assert_same(np.stack(frames), g[tag + '_frames'], tag + ' frames')
    assert_same(m._V_state, g[tag + '_V'], 'V')
    assert_same(m._W_state, g[tag + '_W'], 'W')
    assert_same(m._S_state, g[tag + '_S'], 'S')


@pytest.mark.parametrize('tag,conv', [('heat', False), ('heatconv', True)])
def test_heat_example_end_to_end_bit_exact(tag, conv):
    """Tests/FD/HeatEquation/heat.py run(): Neumann boundary, U0 + (diff*dt)*lap(U0), S2 applied as max(U, s2())."""
    g = gold('ref_heat.npz')
    W, H, D, dx, dy, dz, dt, per_plot, diff, samples, s2_time = (float(v) for v in g[tag + '_cfg'])
    shape = (int(W), int(H), int(D))
    U = np.zeros(shape, dtype=np.float32)
    U[0:2] = 1.0
    s2_init = np.zeros(shape, dtype=np.float32)
    s2_init[:shape[0] // 2, :shape[1] // 2, :] = 1.0
    s2 = Stimulus({'tstart': s2_time, 'nstim': 1, 'period': 800, 'duration': dt, 'dt': dt, 'intensity': 1.0})
    s2.set_stimregion(s2_init)
    frames = [U.copy()]
    fired = 0
    for i in range(int(samples)):
        U = ofd.fd_step(None, U, dt, diff, (dx, dy, dz), conv=conv)
        if s2.stimulate_tissue_timestep(i, dt):
            U = ofd.apply_stimulus_max(U, s2())
            fired += 1
        if i % int(per_plot) == 0:
            frames.append(U.copy())
    assert fired == 1
    assert_same(np.stack(frames), g[tag + '_frames'], tag + ' frames')


def test_laplace_solver_step_bit_exact():
    """Tests/FD/LaplaceSolver/laplaceSolver.py: Dirichlet constant pad + heterogeneous operator, U0 + dt*lap_D(U0)."""
    g = gold('ref_heat.npz')
    assert_same(ofd.enforce_boundary_dirichlet(g['lapsolve_eb_in'], 2.5), g['lapsolve_eb_out'], 'enforce_boundary(cval)')
    dt, wall, dx, dy, dz = (float(v) for v in g['lapsolve_cfg'])
    cond = g['lapsolve_cond']
    U = ofd.enforce_boundary_dirichlet(np.zeros(cond.shape, np.float32), 10.0)
    frames = [U.copy()]
    for _ in range(30):
        U = ofd.fd_step(None, U, dt, 1.0, (dx, dy, dz), DIFF=cond, dirichlet=wall)
        frames.append(U.copy())
    assert_same(np.stack(frames)[[0, 1, 2, 10, 30]], g['lapsolve_frames'], 'LaplaceSolver frames')


@pytest.mark.parametrize('tag', ['sphere_hole', 'sphere_ball'])
def test_sphere_example_end_to_end_bit_exact(tag):
    """Tests/FD/Fenton_sphere/fenton.py run(): geometry mask built as the example builds it and held by the Domain3D
    MIRROR, DIFF = diff*mask (0 outside the tissue), heterogeneous operator, masked initial condition, masked S2,
    frames = where(mask, U, -1)."""
    from pdensorflow_b200.gpuSolve.entities import Domain3D
    g = gold('ref_heat.npz')
    W, H, D, radius, dt, per_plot, diff, samples, s2_time, hole, cyl = (float(v) for v in g[tag + '_cfg'])
    W, H, D, hole, cyl = int(W), int(H), int(D), bool(hole), bool(cyl)
    dom = Domain3D({'width': W, 'height': H, 'depth': D})
    c0 = 0.5 * np.array([dom.dx() * W, dom.dy() * H, dom.dz() * D])
    xx, yy, zz = np.meshgrid(dom.dx() * np.arange(W), dom.dy() * np.arange(H), dom.dz() * np.arange(D))
    distsq = (xx - c0[0]) ** 2 + (yy - c0[1]) ** 2 + np.float32(not cyl) * (zz - c0[2]) ** 2
    inside = distsq <= radius * radius
    vox = (np.logical_not(inside) if hole else inside).astype(np.float32)
    dom.assign_geometry(vox)
    dom.assign_conductivity(diff * vox)
    assert np.array_equal(dom.geometry(), g[tag + '_geometry'])
    DIFF = dom.conductivity().astype(np.float32)
    mask = dom.geometry() > 0.0
    U = np.full((W, H, D), -80.0, dtype=np.float32)
    s2_init = np.full((W, H, D), -80.0, dtype=np.float32)
    if hole:
        U[0:2] = 20.0
        s2_init[:W // 2, :H // 2, :] = 20.0
    else:
        U[W // 2 - 10:W // 2 + 10] = 20.0
        s2_init[:, H // 2 - 10:H // 2 + 10, :] = 20.0
    U = np.where(mask, U, np.float32(-80.0)).astype(np.float32)
    m = oi.Fenton4v()
    m._dt = dt
    m.initialize_state_variables(U)
    s2 = Stimulus({'tstart': s2_time, 'nstim': 1, 'period': 800, 'duration': dt, 'dt': dt, 'intensity': 60.0})
    s2.set_stimregion(np.where(mask, s2_init, -80.0))
    frames = [U.copy()]
    for i in range(int(samples)):
        U = ofd.fd_run(m, U, 1, dt, 1.0, (dom.dx(), dom.dy(), dom.dz()), DIFF=DIFF, s2=s2, step0=i)
        if i % int(per_plot) == 0:
            frames.append(np.where(mask, U, np.float32(-1.0)).astype(np.float32))
    assert_same(np.stack(frames), g[tag + '_frames'], tag + ' frames')
    assert_same(m._V_state, g[tag + '_V'], 'V')


def test_fd_heterogeneous_step_bit_exact():
    g = gold('ref_fd.npz')
    dx, dy, dz = (float(v) for v in g['fdhet_spacing'])
    Dm = g['fdhet_D']
    U = np.full(Dm.shape, -80.0, dtype=np.float32)
    U[0:2] = 20.0
    m = oi.Fenton4v()
    m._dt = 0.1
    m.initialize_state_variables(U)
    U = ofd.fd_run(m, U, 40, 0.1, 1.0, (dx, dy, dz), DIFF=Dm)
    assert np.isfinite(g['fdhet_U40']).all() and g['fdhet_U40'].max() > 0.0
    assert_same(U, g['fdhet_U40'], 'heterogeneous U after 40 steps')
    assert_same(m._V_state, g['fdhet_V40'], 'V')


# ------------------------------------------------------------------------------ FEM
@pytest.mark.parametrize('tag,ionic,etype,stim', [
    ('fem1d', lambda: oi.ModifiedMS2v(), 'Edges', None),
    ('fem2d', lambda: oi.Fenton4v(), 'Trias', {'frac': 0.2, 'props': {'tstart': 0.0, 'nstim': 1, 'period': 800.0,
                                                                       'duration': 0.4, 'intensity': 60.0, 'name': 'S1'}}),
    ('fem2dnr', lambda: oi.ModifiedMS2v(), 'Trias', None)])
def test_monodomain_solver_vs_reference_source(tag, ionic, etype, stim):
    """MonodomainSolver.step (assembly + RCM + stimulus + PCG) run by the reference's own source:
    same CG iteration counts, potentials equal to PCG tolerance (dot-product order differs)."""
    g = gold('ref_fem.npz')
    dt, sigma, ic_frac, nsteps, every, renum = (float(v) for v in g[tag + '_cfg'])
    pts, elems = g[tag + '_pts'], {etype: g['%s_%s' % (tag, etype)]}
    nel = elems[etype].shape[0]
    fibres = np.tile(np.array([[1.0, 0.0, 0.0]]), (nel, 1))      # sigma_l == sigma_t: Sigma = sigma*I either way
    regions = np.unique(elems[etype][:, -1])
    sig = {int(r): sigma for r in regions}
    mo = ofem.MonodomainOracle(ionic(), pts, elems, fibres, sig, sig, dt, use_renumbering=bool(renum),
                               maxiter=pts.shape[0] // 2)
    mo.U[...] = g[tag + '_U0']
    if stim is not None:
        Lx = pts[:, 0].max()
        mo.add_stimulus((pts[:, 0] < stim['frac'] * Lx).astype(np.float32).reshape(-1, 1), stim['props'])
    mo.finalize_for_run()
    frames = [mo.U_user()[:, 0].copy()]
    ctime = 0.0
    for i in range(int(nsteps)):
        ctime += dt
        mo.step(ctime)
        if (i + 1) % int(every) == 0:
            frames.append(mo.U_user()[:, 0].copy())
    ref = g[tag + '_frames']
    assert list(mo.niters) == list(g[tag + '_iters'])
    err = np.linalg.norm(np.stack(frames).astype(np.float64) - ref) / np.linalg.norm(ref.astype(np.float64))
    assert err < 1e-6, err


def test_heat_solver_vs_reference_source():
    """Tests/FEM/HeatEquation/heat.py: HeatSolver.step (implicit Euler heat equation; anisotropic tensors from the example's
    own sigmaTens callback, sigma_l = 0.02 / sigma_t = 0.005 along x; RCM; two stimuli through the float-time interface, the
    second one with three pulses) run by the reference's own source: same CG iteration counts, potentials to PCG tolerance."""
    g = gold('ref_heat.npz')
    pts, elems = g['femheat_pts'], {'Trias': g['femheat_Trias']}
    regions = np.unique(elems['Trias'][:, -1])
    mo = ofem.MonodomainOracle(None, pts, elems, g['femheat_fibres'], {int(r): 0.02 for r in regions},
                               {int(r): 0.005 for r in regions}, 0.1, use_renumbering=True, maxiter=pts.shape[0] // 2)
    for mask, p in ((g['femheat_s1'], g['femheat_p1']), (g['femheat_s2'], g['femheat_p2'])):
        mo.add_stimulus(mask, {'tstart': float(p[0]), 'nstim': int(p[1]), 'period': float(p[2]), 'duration': float(p[3]),
                               'intensity': float(p[4])})
    mo.finalize_for_run()
    frames = [mo.U_user()[:, 0].copy()]
    ctime = 0.0
    for i in range(50):
        ctime += 0.1
        mo.step(ctime)
        if (i + 1) % 5 == 0:
            frames.append(mo.U_user()[:, 0].copy())
    ref = g['femheat_frames']
    assert list(mo.niters) == list(g['femheat_iters'])
    assert ref.max() > 5.0
    err = np.linalg.norm(np.stack(frames).astype(np.float64) - ref) / np.linalg.norm(ref.astype(np.float64))
    assert err < 1e-6, err


# ------------------------------------------------------------------------------ product host code
PROD_CRN_V = ['m_A', 'm_B', 'h_A', 'h_B', 'j_A', 'j_B', 'oa_A', 'oa_B', 'oi_A', 'oi_B', 'ua_A', 'ua_B', 'ui_A', 'ui_B',
              'xr_A', 'xr_B', 'xs_A', 'xs_B', 'd_A', 'd_B', 'f_A', 'f_B', 'w_A', 'w_B', 'GKur', 'INaK', 'IbNa', 'ICaL_v',
              'NaCa_a', 'NaCa_b', 'IbCa_v', 'INa_v']
PROD_TTP_V = ['D_A', 'D_B', 'F2_A', 'F2_B', 'F_A', 'F_B', 'H_A', 'H_B', 'J_A', 'J_B', 'M_A', 'M_B', 'R_A', 'R_B', 'S_A',
              'S_B', 'Xr1_A', 'Xr1_B', 'Xr2_A', 'Xr2_B', 'Xs_A', 'Xs_B', 'a2', 'NaCa_A', 'NaCa_B', 'rec_iNaK', 'rec_ipK']


class _HostTab:
    def __init__(self, ht, names):
        self.table, self.names = ht.array[:, :ht.ncols], names


def test_product_table_builder_matches_reference_construct_tables():
    """The host-side construct_tables of the shipped CRN/TTP classes (NumPy, no GPU needed) in the
    kernel's column layout (include/pdx.h) against the reference's tables."""
    from pdensorflow_b200.gpuSolve.ionic.courtemanche_ramirez_nattel import CourtemancheRamirezNattel
    from pdensorflow_b200.gpuSolve.ionic.ten_tusscher_panfilov import TenTusscherPanfilov
    g = gold('ref_luts.npz')
    with np.errstate(all='ignore'):
        c = [t.finish() for t in CourtemancheRamirezNattel(dt=0.02)._build_tables()]
        t_ = [t.finish() for t in TenTusscherPanfilov(dt=0.02)._build_tables()]
        te = [t.finish() for t in TenTusscherPanfilov(dt=0.02, cell_type='ENDO')._build_tables()]
    check_table(g, 'crn_V_tab', _HostTab(c[0], PROD_CRN_V), CRN_V_COLS)
    check_table(g, 'crn_Cai_tab', _HostTab(c[1], ['Iup_ca', 'buf', 'IpCa_ca', 'conCa', 'fCa_A']), CRN_CAI_COLS)
    check_table(g, 'crn_fn_tab', _HostTab(c[2], ['u_A', 'v_A', 'v_B']), CRN_FN_COLS)
    check_table(g, 'ttp_V_tab_lut', _HostTab(t_[0], PROD_TTP_V), TTP_V_COLS)
    check_table(g, 'ttp_CaSS_tab', _HostTab(t_[1], ['FCaSS_A', 'FCaSS_B']), TTP_CASS_COLS)
    check_table(g, 'ttp_VEk_tab', _HostTab(t_[2], ['rec_iK1']), TTP_VEK_COLS)
    check_table(g, 'ttpendo_V_tab_lut', _HostTab(te[0], PROD_TTP_V), TTP_V_COLS)
    assert all(t.stride % 2 == 0 or t.stride == 1 for t in c + t_)
    assert len(CourtemancheRamirezNattel(dt=0.02)._fold_constants.__code__.co_varnames) >= 1
