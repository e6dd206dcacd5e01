"""Generate tests/golden/ref_heat.npz by EXECUTING THE REFERENCE'S heat / Laplace examples (build container only).

    python tests/golden/make_heat_golden.py

Tests/FD/HeatEquation/heat.py (run() end to end: Neumann boundary, homogeneous and conv Laplacian, the S2 source applied
as U = max(U, s2())) and the step of Tests/FD/LaplaceSolver/laplaceSolver.py (Dirichlet constant pad + heterogeneous
operator, laplaceSolver.py:44-52,136-141) are imported unmodified from /root/reference over tests/golden/tf_facade.py.
LaplaceSolver.__init__ reads a PNG through imageio (absent here), so the instance is created without it and given a
seeded conductivity field; solve() is the reference's own.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_facade as tf  # noqa: E402

tf.install()
from make_ref_golden import load_example  # noqa: E402


class Capture:
    def __init__(self):
        self.frames = []

    def imshow(self, U):
        self.frames.append(np.array(U))

    def wait(self):
        pass


def main():
    out = {}
    ex = load_example('Tests/FD/HeatEquation/heat.py', 'ref_fd_heat_example')
    for tag, convl, shape3, spac in (('heat', False, (10, 12, 8), (1.0, 0.8, 1.25)), ('heatconv', True, (9, 8, 12), (1.0, 1.0, 1.0))):
        cfg = {'width': shape3[0], 'height': shape3[1], 'depth': shape3[2], 'dx': spac[0], 'dy': spac[1], 'dz': spac[2],
               'dt': 0.05, 'dt_per_plot': 10, 'diff': 0.8, 'samples': 60, 's2_time': 1.5, 'convl': convl}
        model = ex.HeatEquation(cfg)
        cap = Capture()
        model.run(cap)
        out[tag + '_cfg'] = np.array([cfg[k] for k in ('width', 'height', 'depth', 'dx', 'dy', 'dz', 'dt', 'dt_per_plot', 'diff',
                                                       'samples', 's2_time')], dtype=np.float64)
        out[tag + '_frames'] = np.stack(cap.frames)
        print(tag, len(cap.frames), cap.frames[-1].min(), cap.frames[-1].max())
    ls_mod = load_example('Tests/FD/LaplaceSolver/laplaceSolver.py', 'ref_fd_laplace_example')
    rng = np.random.default_rng(12)
    shp = (9, 11, 10)
    cond = (1.0 * (rng.uniform(size=shp) > 0.35)).astype(np.float32)           # 0 inside the "anatomy", diff outside
    ls = ls_mod.LaplaceSolver.__new__(ls_mod.LaplaceSolver)
    ls.dt, ls.walltmp = 0.1, 10
    ls.conductivity = tf.constant(cond, dtype=np.float32)
    ls.DX, ls.DY, ls.DZ = (tf.constant(v, dtype=np.float32) for v in (1.0, 1.0, 1.0))
    U = ls_mod.enforce_boundary(tf.Variable(np.zeros(shp, np.float32)))          # default cval = 10.0 (laplaceSolver.py:160)
    frames = [np.array(U)]
    for _ in range(30):
        U = ls.solve(U)
        frames.append(np.array(U))
    out['lapsolve_cond'] = cond
    out['lapsolve_frames'] = np.stack(frames)[[0, 1, 2, 10, 30]]
    out['lapsolve_cfg'] = np.array([0.1, 10.0, 1.0, 1.0, 1.0])
    rngX = rng.uniform(-3.0, 12.0, size=shp).astype(np.float32)
    out['lapsolve_eb_in'] = rngX
    out['lapsolve_eb_out'] = np.array(ls_mod.enforce_boundary(tf.constant(rngX), 2.5))
    # ---- Tests/FD/Fenton_sphere/fenton.py end to end: Domain3D geometry mask, DIFF = diff*mask (0 outside the tissue),
    # heterogeneous operator, masked initial condition / stimulus / frames ------------------------------------------
    sp = load_example('Tests/FD/Fenton_sphere/fenton.py', 'ref_fd_sphere_example')
    for tag, hole, cyl in (('sphere_hole', True, True), ('sphere_ball', False, False)):
        cfg = {'width': 24, 'height': 24, 'depth': 10, 'radius': 5 if hole else 11, 'hole': hole, 'cylindric': cyl, 'dt': 0.1,
               'dt_per_plot': 10, 'diff': 0.8, 'samples': 80, 's2_time': 4.0}
        model = sp.Fenton4vSimple(cfg)
        cap = Capture()
        model.run(cap)
        out[tag + '_cfg'] = np.array([cfg[k] for k in ('width', 'height', 'depth', 'radius', 'dt', 'dt_per_plot', 'diff', 'samples',
                                                       's2_time')] + [float(hole), float(cyl)], dtype=np.float64)
        out[tag + '_frames'] = np.stack(cap.frames)
        out[tag + '_geometry'] = np.asarray(model.domain(), dtype=np.float32)
        out[tag + '_V'] = model._V_state.numpy()
        print(tag, len(cap.frames), cap.frames[-1].min(), cap.frames[-1].max(), out[tag + '_geometry'].mean())
    # ---- Tests/FEM/HeatEquation/heat.py: HeatSolver.step (implicit Euler heat equation, two stimuli through the float-time
    # interface, RCM renumbering) on a small perturbed square, the example's own dfmass / sigmaTens callbacks -----------
    import pickle
    import tempfile
    from make_ref_golden import square_mesh
    from gpuSolve.physics import HeatSolver
    hx = load_example('Tests/FEM/HeatEquation/heat.py', 'ref_fem_heat_example')
    mesh = square_mesh(12, 2.0, seed=3)
    fn = os.path.join(tempfile.mkdtemp(), 'sq.pkl')
    with open(fn, 'wb') as f:
        pickle.dump(mesh, f)
    dt = 0.1
    model = HeatSolver({'mesh_file_name': fn, 'use_renumbering': True, 'dt': dt, 'dt_per_plot': 1, 'Tend': 20.0})
    regions = np.unique(mesh['Elems']['Trias'][:, -1])
    model.add_element_material_property('sigma_l', 'region', {int(r): 0.02 for r in regions})
    model.add_element_material_property('sigma_t', 'region', {int(r): 0.005 for r in regions})
    model.add_material_function('mass', hx.dfmass)
    model.add_material_function('stiffness', hx.sigmaTens)
    model.assemble_matrices()
    pts = mesh['Pts']
    Lx, Ly = pts[:, 0].max(), pts[:, 1].max()
    model.set_initial_condition()
    s1 = (pts[:, 0] < 0.2 * Lx).astype(np.float32).reshape(-1, 1)
    s2 = np.logical_and(pts[:, 0] < 0.6 * Lx, pts[:, 1] < 0.5 * Ly).astype(np.float32).reshape(-1, 1)
    p1 = {'tstart': 0.0, 'nstim': 1, 'period': 200, 'duration': 0.4, 'intensity': 5.0, 'name': 'S1'}
    p2 = {'tstart': 1.5, 'nstim': 3, 'period': 1.0, 'duration': 0.4, 'intensity': 5.0, 'name': 'crossstim'}
    model.add_stimulus(s1, p1)
    model.add_stimulus(s2, p2)
    model.solver().set_maxiter(pts.shape[0] // 2)
    model.finalize_for_run()
    frames, iters, ctime = [np.array(model.U())], [], 0.0
    for i in range(50):
        ctime += model.dt()
        model.step(ctime)
        iters.append(int(model.solver()._niters))
        if (i + 1) % 5 == 0:
            frames.append(np.array(model.U()))
    out['femheat_pts'], out['femheat_Trias'], out['femheat_fibres'] = pts, mesh['Elems']['Trias'], mesh['Fibres']
    out['femheat_frames'] = np.stack(frames)[:, :, 0]
    out['femheat_iters'] = np.array(iters)
    out['femheat_s1'], out['femheat_s2'] = s1, s2
    out['femheat_p1'] = np.array([p1[k] for k in ('tstart', 'nstim', 'period', 'duration', 'intensity')], dtype=np.float64)
    out['femheat_p2'] = np.array([p2[k] for k in ('tstart', 'nstim', 'period', 'duration', 'intensity')], dtype=np.float64)
    print('femheat:', pts.shape[0], 'nodes, CG iterations', min(iters), '..', max(iters), 'U range', frames[-1].min(), frames[-1].max())
    np.savez_compressed(os.path.join(HERE, 'ref_heat.npz'), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
