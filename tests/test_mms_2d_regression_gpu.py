"""Tests/CICD/nightly/test_mms_2d_regression.py:50-184 transcribed to the mirror (MonodomainSolver + ModifiedMS2v on a
2-D sheet: planar front, monotone activation, conduction velocity within 10 % of the analytic Nagumo speed).  The
oracle passes it on CPU (tests/test_fem_oracle_cpu.py::test_reference_mms_2d_regression_on_the_oracle); the 1-D
sibling already runs on the GPU in tests/test_fem_gpu.py."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_fem_gpu import _local_activation_times  # noqa: E402

pytestmark = pytest.mark.gpu
_SIGMA, _DT, _TEND, _VMIN, _VMAX, _VTH = 0.5, 0.05, 10.0, -80.0, 20.0, -30.0


def test_mms_2d_sheet(cuda, tmp_path):
    from pdensorflow_b200.gpuSolve.entities.synthetic import save_pkl, triangulated_square
    from pdensorflow_b200.gpuSolve.ionic.mms2v import ModifiedMS2v
    from pdensorflow_b200.gpuSolve.physics import MonodomainSolver
    mesh = str(tmp_path / 'triangulated_square.pkl')
    save_pkl(mesh, *triangulated_square())                       # 251 x 251 nodes on [0, 10]^2: the reference's spacing
    ionic = ModifiedMS2v(dt=_DT)
    model = MonodomainSolver(ionic, {'mesh_file_name': mesh, 'use_renumbering': True, 'dt': _DT, 'dt_per_plot': 1, 'Tend': _TEND})
    sig = {1: _SIGMA, 2: _SIGMA, 3: _SIGMA, 4: _SIGMA}
    model.add_element_material_property('sigma_l', 'region', sig)
    model.add_element_material_property('sigma_t', 'region', sig)
    model.add_material_function('mass', lambda *a: None)
    model.add_material_function('stiffness', lambda *a: _SIGMA * np.eye(3))
    model.assemble_matrices()
    x = model.domain().Pts()[:, 0].copy()
    Lx = float(x.max())
    model.set_initial_condition(np.where(x < 0.05 * Lx, _VMAX, _VMIN).astype(np.float32))
    model.solver().set_maxiter(x.shape[0] // 2)
    model.finalize_for_run()
    nt = model.nt()
    Vrec = np.empty((nt + 1, x.shape[0]), dtype=np.float32)
    times = np.zeros(nt + 1)
    Vrec[0] = model.U().cpu().numpy().ravel()
    t = 0.0
    for i in range(nt):
        t += _DT
        model.step(t)
        Vrec[i + 1] = model.U().cpu().numpy().ravel()
        times[i + 1] = t
    assert np.all(np.isfinite(Vrec))
    lat = _local_activation_times(Vrec, times, _VTH)
    xr = np.round(x, 6)
    xu = np.sort(np.unique(xr))
    xu = xu[(xu >= 0.4 * Lx) & (xu <= 0.8 * Lx)]
    col = np.array([np.nanmean(lat[xr == v]) for v in xu])
    assert np.all(np.isfinite(col)) and np.all(np.diff(col) > 0.0)
    assert Vrec.min() > _VMIN - 1.0 and Vrec.max() < _VMAX + 1.0
    mask = (x >= 0.4 * Lx) & (x <= 0.8 * Lx) & np.isfinite(lat)
    cv = 1.0 / float(np.polyfit(x[mask], lat[mask], 1)[0])
    cv_an = 0.5 * (1.0 - 2.0 * float(ionic.get_parameter('u_crit'))) * np.sqrt(2.0 * _SIGMA / float(ionic.get_parameter('tau_in')))
    assert abs(cv - cv_an) / cv_an < 0.10, (cv, cv_an)
