"""The CUDA path (through libpdx's C ABI and the gpuSolve mirror) against fixtures produced by
the REFERENCE'S OWN SOURCE (tests/golden/ref_*.npz, see tests/golden/make_ref_golden.py), plus
oracle comparisons for the combinations the reference has no script for (fused FD step with the
lookup-table models, fused anisotropic step).

Tolerances: EXACT arithmetic is expected bit-identical for everything without a transcendental
function and <= 1e-6 relative otherwise; FAST arithmetic <= 1e-5 relative L2 (north star).
"""
import os

import numpy as np
import pytest
import torch

from helpers import rel_l2
from oracle import fd as ofd
from oracle import ionic_lut as ol

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def gold(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------------------ operators
def test_operators_vs_reference_source(cuda):
    from pdensorflow_b200.gpuSolve import diffop2D, diffop3D
    from pdensorflow_b200.gpuSolve.fd import enforce_boundary
    g = gold('ref_fd.npz')
    dx, dy, dz = (float(v) for v in g['lap_spacing'])
    X = dev(g['lap_X'])
    assert np.array_equal(host(diffop3D.laplace_homog(X, dx, dy, dz)), g['lap3_homog'])
    assert np.array_equal(host(diffop3D.laplace_conv_homog(X, dx, dy, dz)), g['lap3_conv'])
    assert np.array_equal(host(diffop3D.laplace_heterog(X, dev(g['lap_D']), dx, dy, dz)), g['lap3_heterog'])
    assert np.array_equal(host(diffop3D.laplace_heterog_aniso(X, dev(g['lap_DIFF2']), dev(g['lap_AVEC']), dx, dy, dz)),
                          g['lap3_aniso'])
    X2 = dev(g['lap2_X'])
    assert np.array_equal(host(diffop2D.laplace_homog(X2, dx, dy)), g['lap2_homog'])
    assert np.array_equal(host(diffop2D.laplace_heterog(X2, dev(g['lap2_D']), dx, dy)), g['lap2_heterog'])
    assert np.array_equal(host(enforce_boundary(dev(g['eb_in']))), g['eb_out'])


# ------------------------------------------------------------------------------ ionic, one call
def _simple(tag):
    from pdensorflow_b200.gpuSolve.ionic import Fenton4v, MitchellSchaeffer2v, ModifiedMS2v
    return {'fk': (Fenton4v, ['_V_state', '_W_state', '_S_state']), 'mms': (ModifiedMS2v, ['_H_state']),
            'ms': (MitchellSchaeffer2v, ['_H_state'])}[tag]


@pytest.mark.parametrize('tag', ['fk', 'mms', 'ms'])
def test_simple_models_one_call_vs_reference_source(cuda, tag):
    g = gold('ref_ionic.npz')
    cls, attrs = _simple(tag)
    m = cls(dt=float(g[tag + '_dt']))
    m.math_mode = 'exact'
    U = dev(g[tag + '_U'])
    m.initialize_state_variables(U)
    for k, a in enumerate(attrs):
        setattr(m, a, dev(g['%s_s%d_in' % (tag, k)]))
    dU = host(m.differentiate(U))
    if tag == 'fk':                       # tanh: correctly rounded on both sides
        assert rel_l2(dU, g[tag + '_dU']) <= 1e-7
    else:
        assert np.array_equal(dU, g[tag + '_dU'])
    for k, a in enumerate(attrs):
        assert rel_l2(host(getattr(m, a)), g['%s_s%d_out' % (tag, k)]) <= 1e-7


def test_mms_per_node_parameters_vs_reference_source(cuda):
    from pdensorflow_b200.gpuSolve.ionic import ModifiedMS2v
    g = gold('ref_ionic.npz')
    m = ModifiedMS2v(dt=0.05)
    U = dev(g['mmsnodal_U'])
    m.initialize_state_variables(U)
    m._H_state = dev(g['mmsnodal_H_in'])
    m.set_parameter('tau_in', g['mmsnodal_tau_in'])
    m.set_parameter('tau_open', g['mmsnodal_tau_open'])
    assert np.array_equal(host(m.differentiate(U)), g['mmsnodal_dU'])
    assert np.array_equal(host(m._H_state), g['mmsnodal_H_out'])


def _lut(tag, **kw):
    from pdensorflow_b200.gpuSolve.ionic import CourtemancheRamirezNattel, TenTusscherPanfilov
    return CourtemancheRamirezNattel(**kw) if tag == 'crn' else TenTusscherPanfilov(**kw)


@pytest.mark.parametrize('tag', ['crn', 'ttp'])
def test_lut_models_one_call_all_states_vs_reference_source(cuda, tag):
    g = gold('ref_ionic.npz')
    m = _lut(tag, dt=0.02)
    U = dev(g[tag + '1_U'])
    m.initialize_state_variables(U)
    names = [str(n) for n in g[tag + '1_names']]
    assert sorted(names) == sorted(m._STATE_ATTRS)
    for n in names:
        setattr(m, n, dev(g['%s1_%s_in' % (tag, n)]))
    dU = host(m.differentiate(U))
    ref = g[tag + '1_dU']
    assert np.isfinite(dU).all()
    np.testing.assert_allclose(dU, ref, rtol=2e-5, atol=1e-6 * np.abs(ref).max())
    for n in names:
        a, b = host(getattr(m, n)), g['%s1_%s_out' % (tag, n)]
        np.testing.assert_allclose(a, b, rtol=2e-5, atol=1e-7 * max(np.abs(b).max(), 1e-30), err_msg=n)


# ------------------------------------------------------------------------------ paced trajectories
def pace_gpu(model, cfg):
    u0, dt, nsteps, amp, stim_ms, period, ncell, spread, every = cfg
    nsteps, ncell, every = int(nsteps), int(ncell), int(every)
    U = dev(np.linspace(u0, u0 + spread, ncell, dtype=np.float32).reshape(ncell, 1))
    model._dt = float(dt)
    model.initialize_state_variables(U)
    trace = [host(U).copy()]
    for i in range(nsteps):
        t = i * dt
        dU = model.differentiate(U)
        if (t % period) < stim_ms:
            dU = dU + amp
        U = U + dt * dU
        if (i + 1) % every == 0:
            trace.append(host(U).copy())
    return np.stack(trace)[:, :, 0]


def _paced_models():
    from pdensorflow_b200.gpuSolve.ionic import (CourtemancheRamirezNattel, Fenton4v, MitchellSchaeffer2v, ModifiedMS2v,
                                                 TenTusscherPanfilov)
    return {'fk': Fenton4v, 'mms': ModifiedMS2v, 'ms': MitchellSchaeffer2v, 'crn': CourtemancheRamirezNattel,
            'ttp': TenTusscherPanfilov, 'ttpendo': lambda: TenTusscherPanfilov(cell_type='ENDO'),
            'ttpm': lambda: TenTusscherPanfilov(cell_type='MCELL')}


@pytest.mark.parametrize('tag', ['fk', 'mms', 'ms', 'crn', 'ttp', 'ttpendo', 'ttpm'])
def test_paced_single_cell_trajectories_vs_reference_source(cuda, tag):
    """Thousands of differentiate() calls through an action-potential upstroke: relative L2 of the
    potential <= 1e-5 (north star) and every sampled value within 0.05 mV."""
    g = gold('ref_ionic.npz')
    m = _paced_models()[tag]()
    m.math_mode = 'exact'
    tr = pace_gpu(m, [float(v) for v in g['pace_%s_cfg' % tag]])
    ref = g['pace_%s_trace' % tag]
    assert rel_l2(tr, ref) <= 1e-5, rel_l2(tr, ref)
    assert np.abs(tr - ref).max() < 0.05


# ------------------------------------------------------------------------------ FD example end to end
@pytest.mark.parametrize('tag,conv', [('fdrun', False), ('fdrunconv', True)])
@pytest.mark.parametrize('math', ['exact', 'fast'])
def test_fd_example_end_to_end_vs_reference_source(cuda, tag, conv, math):
    """Tests/FD/Fenton/fenton.py run() (S1 as initial condition, S2 at step 60 through the integer-step
    Stimulus logic, frames every 10 steps) on the fused kernel."""
    from pdensorflow_b200.fd import FDStepper
    from pdensorflow_b200.gpuSolve.force_terms import Stimulus
    g = gold('ref_fd.npz')
    W, H, D, dx, dy, dz, dt, per_plot, diff, samples, s2_time = (float(v) for v in g[tag + '_cfg'])
    shape = (int(W), int(H), int(D))
    U = np.full(shape, -80.0, dtype=np.float32)
    U[0:2] = 20.0
    s2_init = np.full(shape, -80.0, dtype=np.float32)
    s2_init[:shape[0] // 2, :shape[1] // 2, :] = 20.0
    st = FDStepper(shape, (dx, dy, dz), dt, diff, model='fenton4v', lap='conv' if conv else 'homog', math=math)
    st.set_U(U)
    s2 = Stimulus({'tstart': s2_time, 'nstim': 1, 'period': 800, 'duration': dt, 'dt': dt, 'intensity': 60.0})
    s2.set_stimregion(s2_init)
    frames = [host(st.U()).copy()]
    for i in range(int(samples)):
        st.step(1)
        if s2.stimulate_tissue_timestep(i, dt):
            st.apply_stimulus(s2.get_stimregion(), s2.intensity())
        if i % int(per_plot) == 0:
            frames.append(host(st.U()).copy())
    ref = g[tag + '_frames']
    tol = 1e-6 if math == 'exact' else 1e-5
    for k in range(ref.shape[0]):
        assert rel_l2(frames[k], ref[k]) <= tol, (k, rel_l2(frames[k], ref[k]))
    for s_, key in zip(st.states, ('_V', '_W', '_S')):
        assert rel_l2(host(s_), g[tag + key]) <= tol


def test_fd_heterogeneous_steps_vs_reference_source(cuda):
    from pdensorflow_b200.fd import FDStepper
    g = gold('ref_fd.npz')
    spacing = tuple(float(v) for v in g['fdhet_spacing'])
    Dm = g['fdhet_D']
    U = np.full(Dm.shape, -80.0, dtype=np.float32)
    U[0:2] = 20.0
    for math, tol in (('exact', 1e-6), ('fast', 1e-5)):
        st = FDStepper(Dm.shape, spacing, 0.1, 1.0, model='fenton4v', lap='heterog', DIFF=Dm, math=math)
        st.set_U(U)
        st.step(40)
        assert rel_l2(host(st.U()), g['fdhet_U40']) <= tol
        assert rel_l2(host(st.states[0]), g['fdhet_V40']) <= tol


@pytest.mark.parametrize('tag,conv', [('heat', False), ('heatconv', True)])
@pytest.mark.parametrize('math', ['exact', 'fast'])
def test_heat_example_end_to_end_vs_reference_source(cuda, tag, conv, math):
    """Tests/FD/HeatEquation/heat.py run() on the fused kernel without an ionic model: Neumann boundary,
    U0 + (diff*dt)*lap(U0), the S2 source applied as U = maximum(U, s2()) on every node."""
    from pdensorflow_b200.fd import FDStepper
    from pdensorflow_b200.gpuSolve.force_terms import Stimulus
    g = gold('ref_heat.npz')
    W, H, D, dx, dy, dz, dt, per_plot, diff, samples, s2_time = (float(v) for v in g[tag + '_cfg'])
    shape = (int(W), int(H), int(D))
    U = np.zeros(shape, dtype=np.float32)
    U[0:2] = 1.0
    s2_init = np.zeros(shape, dtype=np.float32)
    s2_init[:shape[0] // 2, :shape[1] // 2, :] = 1.0
    st = FDStepper(shape, (dx, dy, dz), dt, diff, model=None, lap='conv' if conv else 'homog', math=math)
    st.set_U(U)
    s2 = Stimulus({'tstart': s2_time, 'nstim': 1, 'period': 800, 'duration': dt, 'dt': dt, 'intensity': 1.0})
    s2.set_stimregion(s2_init)
    frames = [host(st.U()).copy()]
    for i in range(int(samples)):
        st.step(1)
        if s2.stimulate_tissue_timestep(i, dt):
            st.apply_stimulus(s2.get_stimregion(), s2.intensity(), everywhere=True)
        if i % int(per_plot) == 0:
            frames.append(host(st.U()).copy())
    ref = g[tag + '_frames']
    for k in range(ref.shape[0]):
        assert rel_l2(frames[k], ref[k]) <= (1e-6 if math == 'exact' else 1e-5), (k, rel_l2(frames[k], ref[k]))
    if math == 'exact':
        assert np.abs(np.stack(frames) - ref).max() <= 1e-6   # values in [0, 1]: a few ulp at most


def test_laplace_solver_step_vs_reference_source(cuda):
    """Tests/FD/LaplaceSolver/laplaceSolver.py solve(): Dirichlet constant pad (walltmp) + heterogeneous operator."""
    from pdensorflow_b200.fd import FDStepper
    from pdensorflow_b200.gpuSolve.fd import enforce_boundary
    g = gold('ref_heat.npz')
    X = torch.from_numpy(g['lapsolve_eb_in']).cuda()
    assert np.array_equal(host(enforce_boundary(X, dirichlet_value=2.5)), g['lapsolve_eb_out'])
    dt, wall, dx, dy, dz = (float(v) for v in g['lapsolve_cfg'])
    cond, ref = g['lapsolve_cond'], g['lapsolve_frames']
    st = FDStepper(cond.shape, (dx, dy, dz), dt, 1.0, model=None, lap='heterog', DIFF=cond, math='exact', dirichlet=wall)
    assert st.kernel == 'tma'          # Dirichlet pad and nz % 4 != 0 both run on the tile kernel since r02
    st.set_U(ref[0])
    got = {0: host(st.U()).copy()}
    for n in range(1, 31):
        st.step(1)
        if n in (1, 2, 10, 30):
            got[n] = host(st.U()).copy()
    for k, n in enumerate((0, 1, 2, 10, 30)):
        assert rel_l2(got[n], ref[k]) <= 1e-6, (n, rel_l2(got[n], ref[k]))
    assert ref[4].max() <= 10.0 + 1e-5 and ref[4][1:-1, 1:-1, 1:-1].max() > 0.5      # the wall temperature diffuses in


@pytest.mark.parametrize('tag', ['sphere_hole', 'sphere_ball'])
@pytest.mark.parametrize('math', ['exact', 'fast'])
def test_sphere_example_vs_reference_source(cuda, tag, math):
    """Tests/FD/Fenton_sphere/fenton.py run() on the fused heterogeneous kernel: tissue mask through the Domain3D mirror,
    DIFF = diff*mask (0 outside the tissue), masked initial condition / S2 / frames."""
    from pdensorflow_b200.fd import FDStepper
    from pdensorflow_b200.gpuSolve.entities import Domain3D
    from pdensorflow_b200.gpuSolve.force_terms import Stimulus
    g = gold('ref_heat.npz')
    W, H, D, radius, dt, per_plot, diff, samples, s2_time, hole, cyl = (float(v) for v in g[tag + '_cfg'])
    W, H, D, hole = int(W), int(H), int(D), bool(hole)
    dom = Domain3D({'width': W, 'height': H, 'depth': D})
    vox = g[tag + '_geometry']
    dom.assign_geometry(vox)
    dom.assign_conductivity(diff * vox)
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
    st = FDStepper((W, H, D), (dom.dx(), dom.dy(), dom.dz()), dt, 1.0, model='fenton4v', lap='heterog',
                   DIFF=dom.conductivity().astype(np.float32), math=math)
    st.set_U(U)
    s2 = Stimulus({'tstart': s2_time, 'nstim': 1, 'period': 800, 'duration': dt, 'dt': dt, 'intensity': 60.0})
    s2.set_stimregion(np.where(mask, s2_init, -80.0))
    frames = [U.copy()]
    for i in range(int(samples)):
        st.step(1)
        if s2.stimulate_tissue_timestep(i, dt):
            st.apply_stimulus(s2.get_stimregion(), s2.intensity())
        if i % int(per_plot) == 0:
            frames.append(np.where(mask, host(st.U()), np.float32(-1.0)).astype(np.float32))
    ref = g[tag + '_frames']
    tol = 1e-6 if math == 'exact' else 1e-5
    for k in range(ref.shape[0]):
        assert rel_l2(frames[k], ref[k]) <= tol, (k, rel_l2(frames[k], ref[k]))


# ------------------------------------------------------------------------------ fused FD + LUT models
@pytest.mark.parametrize('tag', ['crn', 'ttp'])
@pytest.mark.parametrize('lap,math', [('homog', 'exact'), ('heterog', 'exact'), ('homog', 'fast')])
def test_fused_fd_step_with_lut_models_vs_oracle(cuda, tag, lap, math):
    """The reference has no FD script for CRN/TTP; the oracle's fd_run with the (reference-pinned)
    oracle models is the arbiter.  S1 plane wave, 60 steps at dt = 0.02.  EXACT (correctly rounded
    exp/log, bit-comparable) <= 1e-6; FAST (fp32 expf/logf, the HBM-bound variant) <= 1e-5."""
    from pdensorflow_b200.fd import FDStepper
    shape, spacing, dt, diff, nsteps = (12, 10, 16), (0.25, 0.25, 0.25), 0.02, 0.1, 60
    mo = ol.CourtemancheRamirezNattel() if tag == 'crn' else ol.TenTusscherPanfilov()
    U = np.full(shape, mo.V_init, dtype=np.float32)
    U[0:3] = 20.0
    Dm = None
    if lap == 'heterog':
        rng = np.random.default_rng(2)
        Dm = (diff * (rng.uniform(size=shape) > 0.15)).astype(np.float32)
    mo._dt = dt
    mo.initialize_state_variables(U)
    Uo = ofd.fd_run(mo, U.copy(), nsteps, dt, diff, spacing, DIFF=Dm)
    mg = _lut(tag)
    st = FDStepper(shape, spacing, dt, diff, model=mg, lap=lap, DIFF=Dm, math=math)
    assert st.kernel == 'generic'
    st.set_U(U)
    st.step(nsteps)
    assert np.isfinite(host(st.U())).all()
    tol = 1e-6 if math == 'exact' else 1e-5
    assert rel_l2(host(st.U()), Uo) <= tol, rel_l2(host(st.U()), Uo)
    key = 'Cai'
    attr = '_Cai'
    assert rel_l2(host(getattr(mg, attr)), mo.s[key]) <= 10 * tol


def test_fused_anisotropic_step_vs_oracle(cuda):
    from pdensorflow_b200.fd import FDStepper
    from oracle import ionic as oi
    g = gold('ref_fd.npz')
    DF, AV = g['lap_DIFF2'], g['lap_AVEC']
    shape, spacing, dt, nsteps = DF.shape[:3], (1.0, 0.8, 1.25), 0.05, 25
    U = np.full(shape, -80.0, dtype=np.float32)
    U[0:2] = 20.0
    mo = oi.Fenton4v()
    mo._dt = dt
    mo.initialize_state_variables(U)
    Uo = U.copy()
    for _ in range(nsteps):
        U0 = ofd.enforce_boundary(Uo)
        dU = mo.differentiate(Uo)
        Uo = (U0 + np.float32(dt) * dU + np.float32(dt) * ofd.laplace_heterog_aniso_3d(U0, DF, AV, *spacing)).astype(np.float32)
    st = FDStepper(shape, spacing, dt, 1.0, model='fenton4v', lap='aniso', DIFF=DF, AVEC=AV, math='exact')
    st.set_U(U)
    st.step(nsteps)
    assert rel_l2(host(st.U()), Uo) <= 1e-6


# ------------------------------------------------------------------------------ FEM
def _fem_case(tag):
    from pdensorflow_b200.gpuSolve.ionic import Fenton4v, ModifiedMS2v
    return {'fem1d': (lambda: ModifiedMS2v(dt=0.05), 'Edges', None),
            'fem2d': (lambda: Fenton4v(dt=0.1), 'Trias', {'frac': 0.2, 'props': {'tstart': 0.0, 'nstim': 1, 'period': 800.0,
                                                                                 'duration': 0.4, 'intensity': 60.0,
                                                                                 'name': 'S1'}}),
            'fem2dnr': (lambda: ModifiedMS2v(dt=0.05), 'Trias', None)}[tag]


@pytest.mark.parametrize('tag', ['fem1d', 'fem2d', 'fem2dnr'])
def test_monodomain_solver_vs_reference_source(cuda, tag, tmp_path):
    """MonodomainSolver (host assembly -> device CSR, persistent PCG kernel) against the reference's
    MonodomainSolver run by its own source.  Each step solves (M + dt K) U1 = b only to the
    reference's stopping rule (|r|^2 < toll^2 = 1e-10 ABSOLUTE, tested every 5 iterations), and the
    iterate at the stopping point depends on the dot-product reduction order, so two correct
    implementations agree to a small multiple of that tolerance, not to 1e-5: 5e-5 relative L2 at
    every recorded frame."""
    import pickle
    from pdensorflow_b200.gpuSolve.physics import MonodomainSolver
    g = gold('ref_fem.npz')
    make_ionic, etype, stim = _fem_case(tag)
    dt, sigma, ic_frac, nsteps, every, renum = (float(v) for v in g[tag + '_cfg'])
    pts, E = g[tag + '_pts'], g['%s_%s' % (tag, etype)]
    fib = np.tile(np.array([[1.0, 0.0, 0.0]]), (E.shape[0], 1)) if etype == 'Trias' else None
    fn = str(tmp_path / 'mesh.pkl')
    with open(fn, 'wb') as f:
        pickle.dump({'Pts': pts, 'Elems': {etype: E}, 'Fibres': fib}, f)
    model = MonodomainSolver(make_ionic(), {'mesh_file_name': fn, 'use_renumbering': bool(renum), 'dt': dt,
                                            'dt_per_plot': 1, 'Tend': 100.0})
    regions = {int(r): sigma for r in np.unique(E[:, -1])}
    model.add_element_material_property('sigma_l', 'region', regions)
    model.add_element_material_property('sigma_t', 'region', regions)
    model.add_material_function('mass', lambda et, ie, dom, mp: None)
    if fib is None:
        model.add_material_function('stiffness', lambda et, ie, dom, mp: sigma * np.eye(3))
    model.assign_nodal_properties()
    model.assemble_matrices()
    model.set_initial_condition(g[tag + '_U0'])
    if stim is not None:
        Lx = pts[:, 0].max()
        model.add_stimulus((pts[:, 0] < stim['frac'] * Lx).astype(np.float32).reshape(-1, 1), stim['props'])
    model.solver().set_maxiter(pts.shape[0] // 2)
    model.finalize_for_run()
    frames = [host(model.U())[:, 0].copy()]
    ctime = 0.0
    for i in range(int(nsteps)):
        ctime += dt
        model.step(ctime)
        if (i + 1) % int(every) == 0:
            frames.append(host(model.U())[:, 0].copy())
    ref = g[tag + '_frames']
    for k in range(ref.shape[0]):
        assert rel_l2(frames[k], ref[k]) <= 5e-5, (k, rel_l2(frames[k], ref[k]))


# ------------------------------------------------------------------------------ N4: anatomical example from a PNG mosaic
@pytest.mark.parametrize('math', ['exact', 'fast'])
def test_atria_example_vs_reference_source(cuda, math):
    """Tests/FD/Fenton_atria/fenton.py run() end to end on the fused heterogeneous kernel: the 128^3 anatomy comes from
    the PNG mosaic through ImageData -> Domain3D (standard-library decoder), DIFF = 0.75 * geometry, initial excited
    slab, the all-ones S2 quirk at step 50; frames i = 0, 40, 80 against the reference run from source."""
    from test_atria_cpu import atria_setup
    from pdensorflow_b200.fd import FDStepper
    from pdensorflow_b200.gpuSolve.entities import Domain3D
    from pdensorflow_b200.gpuSolve.force_terms import Stimulus
    g = gold('ref_atria.npz')
    dom, mask, U, s2_init, dt, per_plot, samples, s2_time = atria_setup(g, Domain3D)
    st = FDStepper(U.shape, (dom.dx(), dom.dy(), dom.dz()), dt, 1.0, model='fenton4v', lap='heterog',
                   DIFF=dom.conductivity().astype(np.float32), math=math)
    assert st.kernel == 'tma'
    st.set_U(U)
    s2 = Stimulus({'tstart': s2_time, 'nstim': 1, 'period': 800, 'duration': dt, 'dt': dt, 'intensity': 60.0})
    s2.set_stimregion(s2_init)
    frames, fired = [U.copy()], []
    for i in range(samples):
        st.step(1)
        if s2.stimulate_tissue_timestep(i, dt):
            st.apply_stimulus(s2.get_stimregion(), s2.intensity())
            fired.append(i)
        if i % per_plot == 0:
            frames.append(np.where(mask, host(st.U()), np.float32(-1.0)).astype(np.float32))
    ref = g['frames']
    assert fired == [50] and len(frames) == ref.shape[0] == 4
    tol = 1e-6 if math == 'exact' else 1e-5
    for k in range(4):
        assert rel_l2(frames[k], ref[k]) <= tol, (k, rel_l2(frames[k], ref[k]))
    assert rel_l2(host(st.states[0]), g['V_end']) <= 10 * tol
