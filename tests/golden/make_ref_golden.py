"""Generate tests/golden/ref_*.npz by EXECUTING THE REFERENCE'S OWN SOURCE.

Run in the build container only (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_ref_golden.py

TensorFlow is not installable in this image, so the reference's Python files are imported
unmodified from /root/reference/PDEnsorflow with ``tests/golden/tf_facade.py`` standing in for
``tensorflow`` (each tf op -> the NumPy op with the same IEEE fp32 semantics; see that file's
header for what is and is not reproduced).  What runs here is therefore the reference's
algorithm exactly as its authors wrote it - operators, ionic models, stimulus logic, the FD
example's own ``solve()``/``run()`` loop, ConjGrad and MonodomainSolver.step - and its outputs on
seeded inputs are frozen as the fixtures the oracle (CPU tests) and the CUDA path (GPU tests)
are checked against.

Files written
  ref_ionic.npz   one differentiate() call of every ionic model on seeded fields + single-cell
                  paced trajectories (the protocol of Tests/DEVTESTS/ionic/ionic.py, shortened)
  ref_luts.npz    strided samples + fp64 checksums of the CRN / TTP lookup tables
  ref_fd.npz      the six diffop Laplacians, enforce_boundary, Stimulus step windows, and the
                  Tests/FD/Fenton/fenton.py example run end to end on a small grid
  ref_fem.npz     ConjGrad solves and MonodomainSolver stepping on small meshes
"""
import importlib.util
import os
import pickle
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_facade as tf  # noqa: E402

tf.install()
REF = tf.REF_ROOT


def load_example(relpath, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def seeded_fields(shape, seed, nstates):
    rng = np.random.default_rng(seed)
    U = rng.uniform(-85.0, 25.0, size=shape).astype(np.float32)
    # put some nodes exactly on the thresholds of the step functions (sign(0) = 0 branches)
    flat = U.reshape(-1)
    flat[0] = np.float32(-80.0) + np.float32(100.0) * np.float32(0.23)     # u == u_c (fp32 permitting)
    flat[1] = np.float32(-50.0)                                           # u == u_so = 0.3
    states = [rng.uniform(0.0, 1.0, size=shape).astype(np.float32) for _ in range(nstates)]
    return U, states


# =========================================================================== ionic
def gen_ionic(out):
    from gpuSolve.ionic.fenton4v import Fenton4v
    from gpuSolve.ionic.mms2v import ModifiedMS2v
    from gpuSolve.ionic.ms2v import MitchellSchaeffer2v
    from gpuSolve.ionic.courtemanche_ramirez_nattel import CourtemancheRamirezNattel
    from gpuSolve.ionic.ten_tusscher_panfilov import TenTusscherPanfilov

    shape = (7, 5, 12)
    # ---- one call on seeded fields -------------------------------------------------------
    for tag, cls, attrs, dt in (('fk', Fenton4v, ['_V_state', '_W_state', '_S_state'], 0.1),
                                ('mms', ModifiedMS2v, ['_H_state'], 0.05),
                                ('ms', MitchellSchaeffer2v, ['_H_state'], 0.05)):
        U, st = seeded_fields(shape, 11, len(attrs))
        m = cls(dt=dt)
        Ut = tf.Variable(U)
        m.initialize_state_variables(Ut)
        for a, s in zip(attrs, st):
            getattr(m, a).assign(s)
        dU = m.differentiate(Ut)
        out[tag + '_U'] = U
        out[tag + '_dt'] = np.float64(dt)
        out[tag + '_dU'] = dU.numpy()
        for k, (a, s) in enumerate(zip(attrs, st)):
            out['%s_s%d_in' % (tag, k)] = s
            out['%s_s%d_out' % (tag, k)] = getattr(m, a).numpy()

    # per-node parameters (Tests/FEM/mMS/mMS.py:100-109 installs [N,1] arrays via set_parameter)
    rng = np.random.default_rng(5)
    n = 64
    U = rng.uniform(-85.0, 25.0, size=(n, 1)).astype(np.float32)
    Hs = rng.uniform(0.0, 1.0, size=(n, 1)).astype(np.float32)
    tau_in = rng.uniform(0.05, 0.4, size=(n, 1)).astype(np.float32)
    tau_open = rng.uniform(80.0, 130.0, size=(n, 1)).astype(np.float32)
    m = ModifiedMS2v(dt=0.05)
    Ut = tf.Variable(U)
    m.initialize_state_variables(Ut)
    m._H_state.assign(Hs)
    m.set_parameter('tau_in', tau_in)
    m.set_parameter('tau_open', tau_open)
    out['mmsnodal_U'], out['mmsnodal_H_in'] = U, Hs
    out['mmsnodal_tau_in'], out['mmsnodal_tau_open'] = tau_in, tau_open
    out['mmsnodal_dU'] = m.differentiate(Ut).numpy()
    out['mmsnodal_H_out'] = m._H_state.numpy()

    # ---- paced single-cell trajectories (Tests/DEVTESTS/ionic/ionic.py:59-105 protocol:
    #      U += dt*(differentiate(U) [+ stim while t in the first ms of a beat])) ------------
    def pace(model, u0, dt, nsteps, stim_amp, stim_ms, period_ms, ncell, spread, every):
        U = tf.Variable(np.linspace(u0, u0 + spread, ncell, dtype=np.float32).reshape(ncell, 1))
        model._dt = dt
        model.initialize_state_variables(U)
        trace = [U.numpy().copy()]
        for i in range(nsteps):
            t = i * dt
            dU = model.differentiate(U)
            if (t % period_ms) < stim_ms:
                dU = dU + stim_amp
            U.assign(U + dt * dU)
            if (i + 1) % every == 0:
                trace.append(U.numpy().copy())
        return np.stack(trace)[:, :, 0], U

    for tag, model, u0, dt, nsteps, amp in (
            ('fk', Fenton4v(), -80.0, 0.02, 6000, 60.0),
            ('mms', ModifiedMS2v(), -80.0, 0.02, 6000, 60.0),
            ('ms', MitchellSchaeffer2v(), -80.0, 0.02, 6000, 60.0),
            ('crn', CourtemancheRamirezNattel(), -81.2, 0.02, 6000, 60.0),
            ('ttp', TenTusscherPanfilov(), -86.2, 0.02, 6000, 60.0),
            ('ttpendo', TenTusscherPanfilov(cell_type='ENDO'), -86.2, 0.02, 3000, 60.0),
            ('ttpm', TenTusscherPanfilov(cell_type='MCELL'), -86.2, 0.02, 3000, 60.0)):
        tr, _ = pace(model, u0, dt, nsteps, amp, 1.0, 1000.0, 8, 7.0, 100)
        out['pace_%s_trace' % tag] = tr
        out['pace_%s_cfg' % tag] = np.array([u0, dt, nsteps, amp, 1.0, 1000.0, 8, 7.0, 100], dtype=np.float64)
        print('paced %-8s  V range [%.2f, %.2f]' % (tag, tr.min(), tr.max()))

    # ---- one call of the LUT models on seeded fields (all states perturbed) ----------------
    for tag, model in (('crn', CourtemancheRamirezNattel(dt=0.02)), ('ttp', TenTusscherPanfilov(dt=0.02))):
        rng = np.random.default_rng(21)
        n = 96
        U = rng.uniform(-95.0, 45.0, size=(n, 1)).astype(np.float32)
        Ut = tf.Variable(U)
        model.initialize_state_variables(Ut)
        names = sorted(k for k, v in model.__dict__.items() if isinstance(v, tf.Tensor) and v.shape == U.shape)
        for k in names:
            v = getattr(model, k)
            v.assign((v.numpy() * rng.uniform(0.8, 1.2, size=U.shape)).astype(np.float32))
            out['%s1_%s_in' % (tag, k)] = v.numpy()
        out['%s1_U' % tag] = U
        out['%s1_dU' % tag] = model.differentiate(Ut).numpy()
        for k in names:
            out['%s1_%s_out' % (tag, k)] = getattr(model, k).numpy()
        out['%s1_names' % tag] = np.array(names)


def gen_luts(out):
    from gpuSolve.ionic.courtemanche_ramirez_nattel import CourtemancheRamirezNattel
    from gpuSolve.ionic.ten_tusscher_panfilov import TenTusscherPanfilov
    for tag, model, tabs in (('crn', CourtemancheRamirezNattel(dt=0.02), ['_V_tab', '_Cai_tab', '_fn_tab']),
                             ('ttp', TenTusscherPanfilov(dt=0.02), ['_V_tab_lut', '_CaSS_tab', '_VEk_tab']),
                             ('ttpendo', TenTusscherPanfilov(dt=0.02, cell_type='ENDO'), ['_V_tab_lut'])):
        model.construct_tables()
        for t in tabs:
            T = getattr(model, t).numpy()
            stride = max(1, T.shape[0] // 257)
            out['%s%s_shape' % (tag, t)] = np.array(T.shape)
            out['%s%s_stride' % (tag, t)] = np.int64(stride)
            out['%s%s_rows' % (tag, t)] = T[::stride]
            out['%s%s_colsum' % (tag, t)] = T.astype(np.float64).sum(axis=0)
            out['%s%s_colabs' % (tag, t)] = np.abs(T.astype(np.float64)).sum(axis=0)


# =========================================================================== FD
def gen_fd(out):
    from gpuSolve.diffop3D import laplace_homog, laplace_conv_homog, laplace_heterog, laplace_heterog_aniso
    from gpuSolve import diffop2D
    from gpuSolve.force_terms import Stimulus

    rng = np.random.default_rng(3)
    shp = (9, 7, 12)
    X = rng.uniform(-85.0, 25.0, size=shp).astype(np.float32)
    D = rng.uniform(0.0, 1.5, size=shp).astype(np.float32)
    D[rng.uniform(size=shp) < 0.15] = 0.0           # "outside the tissue" cells
    dx, dy, dz = 1.0, 0.5, 0.25
    out['lap_X'], out['lap_D'] = X, D
    out['lap_spacing'] = np.array([dx, dy, dz])
    DX, DY, DZ = (tf.constant(v, dtype=np.float32) for v in (dx, dy, dz))
    out['lap3_homog'] = laplace_homog(tf.constant(X), DX, DY, DZ).numpy()
    out['lap3_conv'] = laplace_conv_homog(tf.constant(X), DX, DY, DZ).numpy()
    out['lap3_heterog'] = laplace_heterog(tf.constant(X), tf.constant(D), DX, DY, DZ).numpy()
    # anisotropic: sigma_t, sigma_l - sigma_t and the fibre dyadic a(x)a
    st = rng.uniform(0.1, 0.5, size=shp).astype(np.float32)
    dl = rng.uniform(0.0, 1.0, size=shp).astype(np.float32)
    a = rng.normal(size=shp + (3,))
    a /= np.linalg.norm(a, axis=-1, keepdims=True)
    AV = np.stack([a[..., 0] * a[..., 0], a[..., 0] * a[..., 1], a[..., 0] * a[..., 2],
                   a[..., 1] * a[..., 1], a[..., 1] * a[..., 2], a[..., 2] * a[..., 2]], axis=-1).astype(np.float32)
    DF = np.stack([st, dl], axis=-1).astype(np.float32)
    out['lap_DIFF2'], out['lap_AVEC'] = DF, AV
    out['lap3_aniso'] = laplace_heterog_aniso(tf.constant(X), tf.constant(DF), tf.constant(AV), DX, DY, DZ).numpy()
    X2, D2 = X[:, :, 3].copy(), D[:, :, 5].copy()
    out['lap2_X'], out['lap2_D'] = X2, D2
    out['lap2_homog'] = diffop2D.laplace_homog(tf.constant(X2), DX, DY).numpy()
    out['lap2_heterog'] = diffop2D.laplace_heterog(tf.constant(X2), tf.constant(D2), DX, DY).numpy()

    # ---- Stimulus windows (force_terms/stimulus.py:110-165) --------------------------------
    cases = [dict(tstart=200, nstim=1, period=800, duration=0.1, dt=0.1, intensity=60.0),       # FD example's S2
             dict(tstart=0.7, nstim=3, period=2.0, duration=0.35, dt=0.1, intensity=10.0),
             dict(tstart=0.0, nstim=2, period=1.5, duration=0.4, dt=0.05, intensity=5.0)]
    for n, cfg in enumerate(cases):
        s = Stimulus(cfg)
        s.set_stimregion(np.ones((4, 1), dtype=np.float32))
        nst = 2500 if n == 0 else 120
        fired = np.array([bool(s.stimulate_tissue_timestep(i, cfg['dt'])) for i in range(nst)])
        out['stim%d_cfg' % n] = np.array([cfg[k] for k in ('tstart', 'nstim', 'period', 'duration', 'dt', 'intensity')], dtype=np.float64)
        out['stim%d_steps' % n] = np.nonzero(fired)[0]
        out['stim%d_nsteps' % n] = np.int64(nst)
    # float-time interface as HeatSolver._compute_forcing uses it (heatSolver.py:137-145):
    # running float64 clock, evaluated as a float32 constant
    for n, cfg in enumerate([dict(tstart=0.0, nstim=1, period=800.0, duration=0.4, intensity=60.0, name='S1'),
                             dict(tstart=21.0, nstim=2, period=3.0, duration=0.4, intensity=60.0, name='S2')]):
        s = Stimulus(cfg)
        s.set_stimregion(np.ones((4, 1), dtype=np.float32))
        ctime, dt, act = 0.0, 0.1, []
        for i in range(400):
            val = s.stimApp(tf.constant(ctime, dtype=np.float32))
            act.append(float(np.max(np.asarray(val))) != 0.0)
            ctime += dt
        out['stimt%d_cfg' % n] = np.array([cfg['tstart'], cfg['nstim'], cfg['period'], cfg['duration'], cfg['intensity']])
        out['stimt%d_steps' % n] = np.nonzero(np.array(act))[0]

    # ---- the FD example itself, end to end (Tests/FD/Fenton/fenton.py) ---------------------
    ex = load_example('Tests/FD/Fenton/fenton.py', 'ref_fd_fenton_example')
    out['eb_in'] = X
    out['eb_out'] = ex.enforce_boundary(tf.constant(X)).numpy()

    class Capture:
        def __init__(self):
            self.frames = []

        def imshow(self, U):
            self.frames.append(np.array(U))

        def wait(self):
            pass

    for tag, convl, shape3, spac in (('fdrun', False, (16, 12, 20), (1.0, 1.0, 1.0)),
                                     ('fdrunconv', True, (10, 12, 8), (1.0, 0.8, 1.25))):
        cfg = {'width': shape3[0], 'height': shape3[1], 'depth': shape3[2], 'dx': spac[0], 'dy': spac[1], 'dz': spac[2],
               'dt': 0.1, 'dt_per_plot': 10, 'diff': 0.8, 'samples': 120, 's2_time': 6.0, 'convl': convl}
        model = ex.Fenton4vSimple(cfg)
        cap = Capture()
        model.run(cap)
        out[tag + '_cfg'] = np.array([cfg[k] for k in ('width', 'height', 'depth', 'dx', 'dy', 'dz', 'dt', 'dt_per_plot',
                                                       'diff', 'samples', 's2_time')], dtype=np.float64)
        out[tag + '_frames'] = np.stack(cap.frames)
        out[tag + '_V'] = model._V_state.numpy()
        out[tag + '_W'] = model._W_state.numpy()
        out[tag + '_S'] = model._S_state.numpy()
        print('%s: %d frames, U range [%.2f, %.2f]' % (tag, len(cap.frames), cap.frames[-1].min(), cap.frames[-1].max()))

    # heterogeneous example step (Tests/FD/Fenton_sphere/fenton.py:109-115: U0 + dt*dU + dt*lap_D(U0))
    from gpuSolve.ionic.fenton4v import Fenton4v
    m = Fenton4v()
    m._dt = 0.1
    Uv = tf.Variable(np.where(np.arange(shp[0])[:, None, None] < 2, 20.0, -80.0).astype(np.float32) * np.ones(shp, np.float32))
    m.initialize_state_variables(Uv)
    Dm = (0.8 * (D > 0.3)).astype(np.float32)
    HX, HY, HZ = (tf.constant(v, dtype=np.float32) for v in (1.0, 0.8, 1.25))
    U = Uv
    for _ in range(40):
        U0 = ex.enforce_boundary(U)
        dU = m.differentiate(U)
        U = U0 + 0.1 * dU + 0.1 * laplace_heterog(U0, tf.constant(Dm), HX, HY, HZ)
    out['fdhet_D'] = Dm
    out['fdhet_spacing'] = np.array([1.0, 0.8, 1.25])
    out['fdhet_U40'] = U.numpy()
    out['fdhet_V40'] = m._V_state.numpy()


# =========================================================================== FEM
def cable_mesh(n_el, length):
    x = np.linspace(0.0, length, n_el + 1)
    pts = np.stack([x, np.zeros_like(x), np.zeros_like(x)], axis=1)
    edges = np.stack([np.arange(n_el), np.arange(1, n_el + 1), np.ones(n_el, dtype=int)], axis=1)
    return {'Pts': pts, 'Elems': {'Edges': edges}, 'Fibres': None}


def square_mesh(n, length, seed=0):
    g = np.linspace(0.0, length, n)
    X, Y = np.meshgrid(g, g, indexing='ij')
    pts = np.stack([X.ravel(), Y.ravel(), np.zeros(n * n)], axis=1)
    rng = np.random.default_rng(seed)
    inner = ((pts[:, 0] > 0) & (pts[:, 0] < length) & (pts[:, 1] > 0) & (pts[:, 1] < length))
    pts[inner, :2] += rng.uniform(-0.2, 0.2, size=(inner.sum(), 2)) * (length / (n - 1))
    idx = np.arange(n * n).reshape(n, n)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel(), idx[:-1, 1:].ravel()
    tr = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)])
    cen = pts[tr].mean(axis=1)
    reg = 1 + (cen[:, 0] > length / 2).astype(int) + 2 * (cen[:, 1] > length / 2).astype(int)
    fib = np.tile(np.array([[1.0, 0.0, 0.0]]), (tr.shape[0], 1))
    return {'Pts': pts, 'Elems': {'Trias': np.concatenate([tr, reg[:, None]], axis=1)}, 'Fibres': fib}


def gen_fem(out):
    from gpuSolve.physics import MonodomainSolver
    from gpuSolve.ionic.mms2v import ModifiedMS2v
    from gpuSolve.ionic.fenton4v import Fenton4v
    tmp = tempfile.mkdtemp()
    femex = load_example('Tests/FEM/Fenton/fenton.py', 'ref_fem_fenton_example')   # its dfmass / sigmaTens callbacks

    def run(tag, mesh, ionic, cfg_extra, sigma, ic_frac, nsteps, stim=None, renum=True, every=5):
        fn = os.path.join(tmp, tag + '.pkl')
        with open(fn, 'wb') as f:
            pickle.dump(mesh, f)
        cfg = {'mesh_file_name': fn, 'use_renumbering': renum, 'dt_per_plot': 1}
        cfg.update(cfg_extra)
        model = MonodomainSolver(ionic, cfg)
        regions = np.unique([int(r) for et in mesh['Elems'].values() for r in et[:, -1]])
        model.add_element_material_property('sigma_l', 'region', {int(r): sigma for r in regions})
        model.add_element_material_property('sigma_t', 'region', {int(r): sigma for r in regions})
        model.add_material_function('mass', femex.dfmass)
        if mesh['Fibres'] is not None:
            model.add_material_function('stiffness', femex.sigmaTens)
        else:   # Tests/CICD/unit/test_mms_1d.py:80-82: isotropic tensor for the fibre-less cable
            model.add_material_function('stiffness', lambda et, ie, dom, mp: sigma * np.eye(3))
        model.assign_nodal_properties()
        model.assemble_matrices()
        pts = mesh['Pts']
        Lx = pts[:, 0].max()
        U0 = np.full((pts.shape[0], 1), -80.0, dtype=np.float32)
        if ic_frac > 0:
            U0[pts[:, 0] < ic_frac * Lx] = 20.0
        model.set_initial_condition(U0)
        if stim is not None:
            mask = (pts[:, 0] < stim['frac'] * Lx).astype(np.float32).reshape(-1, 1)
            model.add_stimulus(mask, stim['props'])
        model.solver().set_maxiter(pts.shape[0] // 2)
        model.finalize_for_run()
        frames, iters = [np.array(model.U())], []
        ctime = 0.0
        for i in range(nsteps):
            ctime += model.dt()
            model.step(ctime)
            iters.append(int(model.solver()._niters))
            if (i + 1) % every == 0:
                frames.append(np.array(model.U()))
        out[tag + '_pts'] = pts
        for et, arr in mesh['Elems'].items():
            out['%s_%s' % (tag, et)] = arr
        out[tag + '_U0'] = U0
        out[tag + '_frames'] = np.stack(frames)[:, :, 0]
        out[tag + '_iters'] = np.array(iters)
        out[tag + '_cfg'] = np.array([cfg['dt'], sigma, ic_frac, nsteps, every, int(renum)], dtype=np.float64)
        print('%s: %d nodes, %d steps, CG iterations %s..%s, U range [%.2f, %.2f]' %
              (tag, pts.shape[0], nsteps, min(iters), max(iters), frames[-1].min(), frames[-1].max()))
        return model

    # 1-D cable of Tests/CICD/unit/test_mms_1d.py (shortened), mMS
    run('fem1d', cable_mesh(60, 1.2), ModifiedMS2v(dt=0.05), {'dt': 0.05, 'Tend': 3.0}, 0.1, 0.2, 60)
    # 2-D perturbed square, Fenton-Karma with an S1 stimulus through the float-time interface
    run('fem2d', square_mesh(13, 2.0), Fenton4v(dt=0.1), {'dt': 0.1, 'Tend': 10.0}, 1e-2, 0.0, 30,
        stim={'frac': 0.2, 'props': {'tstart': 0.0, 'nstim': 1, 'period': 800.0, 'duration': 0.4,
                                      'intensity': 60.0, 'name': 'S1'}})
    # same without renumbering
    run('fem2dnr', square_mesh(9, 2.0, seed=1), ModifiedMS2v(dt=0.05), {'dt': 0.05, 'Tend': 10.0}, 5e-2, 0.15, 20,
        renum=False)


def main():
    which = sys.argv[1:] or ['ionic', 'luts', 'fd', 'fem']
    for name, fn in (('ionic', gen_ionic), ('luts', gen_luts), ('fd', gen_fd), ('fem', gen_fem)):
        if name not in which:
            continue
        out = {}
        fn(out)
        path = os.path.join(HERE, 'ref_%s.npz' % name)
        np.savez_compressed(path, **out)
        print('wrote %s (%d arrays, %.1f KiB)' % (path, len(out), os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
