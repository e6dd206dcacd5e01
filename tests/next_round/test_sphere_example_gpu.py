"""Tests/FD/Fenton_sphere/fenton.py run() replayed on the fused heterogeneous kernel against the fixture frozen from the
reference source (tests/golden/ref_heat.npz, pinned on CPU by tests/test_ref_golden_cpu.py::test_sphere_example_*)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helpers import rel_l2  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('tag', ['sphere_hole', 'sphere_ball'])
@pytest.mark.parametrize('math', ['exact', 'fast'])
def test_sphere_example_vs_reference_source(cuda, tag, math):
    from pdensorflow_b200.fd import FDStepper
    from pdensorflow_b200.gpuSolve.entities import Domain3D
    from pdensorflow_b200.gpuSolve.force_terms import Stimulus
    g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'golden', 'ref_heat.npz'))
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
            frames.append(np.where(mask, st.U().cpu().numpy(), np.float32(-1.0)).astype(np.float32))
    ref = g[tag + '_frames']
    tol = 1e-6 if math == 'exact' else 1e-5
    for k in range(ref.shape[0]):
        assert rel_l2(frames[k], ref[k]) <= tol, (k, rel_l2(frames[k], ref[k]))
