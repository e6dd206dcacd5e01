"""Tests/FEM/HeatEquation/heat.py replayed on the mirror's HeatSolver against the fixture frozen from the reference source
(tests/golden/ref_heat.npz femheat_*; the oracle reproduces it on CPU: tests/test_ref_golden_cpu.py::
test_heat_solver_vs_reference_source)."""
import os
import pickle

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_heat_solver_vs_reference_source(cuda, tmp_path):
    from pdensorflow_b200.gpuSolve.physics import HeatSolver
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_heat.npz'))
    pts, trias, fib = g['femheat_pts'], g['femheat_Trias'], g['femheat_fibres']
    fn = str(tmp_path / 'sq.pkl')
    with open(fn, 'wb') as f:
        pickle.dump({'Pts': pts, 'Elems': {'Trias': trias}, 'Fibres': fib}, f)
    model = HeatSolver({'mesh_file_name': fn, 'use_renumbering': True, 'dt': 0.1, 'dt_per_plot': 1, 'Tend': 20.0})
    regions = np.unique(trias[:, -1])
    model.add_element_material_property('sigma_l', 'region', {int(r): 0.02 for r in regions})
    model.add_element_material_property('sigma_t', 'region', {int(r): 0.005 for r in regions})
    model.assemble_matrices()
    model.set_initial_condition()
    for mask, p in ((g['femheat_s1'], g['femheat_p1']), (g['femheat_s2'], g['femheat_p2'])):
        model.add_stimulus(mask, {'tstart': float(p[0]), 'nstim': int(p[1]), 'period': float(p[2]), 'duration': float(p[3]),
                                  'intensity': float(p[4])})
    model.solver().set_maxiter(pts.shape[0] // 2)
    model.finalize_for_run()
    frames, iters, ctime = [model.U().cpu().numpy()[:, 0].copy()], [], 0.0
    for i in range(50):
        ctime += model.dt()
        model.step(ctime)
        iters.append(model.solver().niters())
        if (i + 1) % 5 == 0:
            frames.append(model.U().cpu().numpy()[:, 0].copy())
    ref = g['femheat_frames']
    d = np.abs(np.asarray(iters) - g['femheat_iters'])
    assert d.max() <= 5 and (d > 0).sum() <= 10
    err = np.linalg.norm(np.stack(frames).astype(np.float64) - ref) / np.linalg.norm(ref.astype(np.float64))
    assert err < 5e-5, err
