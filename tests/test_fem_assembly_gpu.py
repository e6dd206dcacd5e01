"""Device-side P1 assembly (pdx_fem_assemble) against the oracle assembly, which is itself
pinned to the reference's NumPy assembly (tests/golden/fem_golden.npz): identical sparsity pattern (structural zeros
included), values within 2e-6 of the largest entry (the reference accumulates fp32 running sums, the device fp64
sums rounded once).  Edges, Trias with two regions + fibres (anisotropic tensors), Tetras, with and without RCM."""
import numpy as np
import pytest

from oracle import fem as ofem

pytestmark = pytest.mark.gpu


def _domain(P, E, F):
    from pdensorflow_b200.gpuSolve.entities.triangulation import Triangulation
    T = Triangulation()
    T.set_mesh(P, E, F)
    return T


def _props(sl, st):
    from pdensorflow_b200.gpuSolve.entities.materialproperties import MaterialProperties
    mp = MaterialProperties()
    if sl is not None:
        mp.add_element_property('sigma_l', 'region', sl)
        mp.add_element_property('sigma_t', 'region', st)
    return mp


def _cases():
    from pdensorflow_b200.gpuSolve.entities.synthetic import cable, tetrahedral_box, triangulated_square
    P, E, F = triangulated_square(23, 17, 3.0, 2.0)
    rng = np.random.default_rng(4)
    F = rng.standard_normal(F.shape)
    F /= np.linalg.norm(F, axis=1, keepdims=True)                      # arbitrary unit fibres: full anisotropic tensors
    yield 'trias', P, E, F, {1: 2.0, 2: 0.7, 3: 1.3, 4: 0.2}, {1: 0.5, 2: 0.7, 3: 0.1, 4: 0.05}
    P, E, F = tetrahedral_box(7, 5, 6, 1.4, 1.0, 2.0)
    yield 'tetras', P, E, F, {1: 1.5}, {1: 0.4}
    P, E, F = cable(40, 2.0)
    yield 'edges', P, E, F, None, None


@pytest.mark.parametrize('rcm', [False, True])
def test_device_assembly_matches_oracle(cuda, rcm):
    from pdensorflow_b200.gpuSolve.matrices.device_assembly import assemble_on_device
    for tag, P, E, F, sl, st in _cases():
        npt = P.shape[0]
        ren = None
        if rcm:
            ren = ofem.rcm_indexing(ofem.coo_pattern(ofem.connectivity(npt, E)))
        ref = ofem.assemble(P, E, F, sl, st, None if ren is None else ren['iperm'])
        got = assemble_on_device(_domain(P, E, F), _props(sl, st), renumbering=ren)
        for name in ('mass', 'stiffness'):
            M = got[name]
            assert np.array_equal(M.rowptr, ref['rowptr']) and np.array_equal(M.col, ref['col']), (tag, name)
            scale = np.abs(ref[name]).max()
            assert np.abs(M.val - ref[name]).max() <= 2e-6 * scale, (tag, name, np.abs(M.val - ref[name]).max() / scale)
            rp, ci, va = M.device_arrays()
            assert np.array_equal(va.cpu().numpy(), M.val) and va.is_cuda
        # the stiffness matrix annihilates constants, the mass matrix sums to the measure of the domain
        rows = np.repeat(np.arange(npt), np.diff(got['stiffness'].rowptr))
        assert np.abs(np.bincount(rows, weights=got['stiffness'].val, minlength=npt)).max() <= 1e-5 * np.abs(got['stiffness'].val).max()


def test_solver_uses_device_assembly_when_asked(cuda):
    """MonodomainSolver with cfg device_assembly=True steps like the host-assembled one (within CG tolerance)."""
    from test_fem_gpu import _square
    from pdensorflow_b200.gpuSolve.ionic import Fenton4v
    from pdensorflow_b200.gpuSolve.physics import MonodomainSolver
    P, E, F = _square(31)
    sig = {1: 1e-3, 2: 1e-3, 3: 1e-3, 4: 1e-3}
    out = []
    for dev_asm in (False, True):
        ionic = Fenton4v(dt=0.1)
        model = MonodomainSolver(ionic, {'use_renumbering': True, 'dt': 0.1, 'dt_per_plot': 10, 'Tend': 100,
                                         'device_assembly': dev_asm})
        model.domain().set_mesh(P, E, F)
        model.add_element_material_property('sigma_l', 'region', sig)
        model.add_element_material_property('sigma_t', 'region', sig)
        model.assign_nodal_properties()
        model.assemble_matrices()
        model.set_initial_condition()
        model.add_stimulus(P[:, 0] < 0.05 * P[:, 0].max(), {'tstart': 0.0, 'nstim': 1, 'period': 100, 'duration': 0.4, 'intensity': 60.0})
        model.solver().set_maxiter(P.shape[0] // 2)
        model.finalize_for_run()
        t = 0.0
        for _ in range(30):
            t += 0.1
            model.step(t)
        out.append(model.U().cpu().numpy().ravel())
    assert out[0].max() > -70.0
    assert np.linalg.norm(out[0] - out[1]) <= 5e-5 * np.linalg.norm(out[0])
