"""Host-side FEM pieces of the product (no GPU): assembly / pattern / RCM / csr_axpby / Jacobi of
pdensorflow_b200.gpuSolve.matrices must equal the oracle (which is pinned to the reference's NumPy code
by tests/test_fem_oracle_cpu.py) bit for bit, and the drop-in containers must behave like the reference's."""
import os
import pickle

import numpy as np
import pytest

from oracle import fem as ofem
from pdensorflow_b200.gpuSolve.entities import MaterialProperties, Triangulation
from pdensorflow_b200.gpuSolve.entities.synthetic import cable, tetrahedral_box, triangulated_square
from pdensorflow_b200.gpuSolve.matrices import (assemble_matrices_dict, compute_coo_pattern,
                                                compute_reverse_cuthill_mckee_indexing, csr_axpby, localMass,
                                                localStiffness)

MESHES = {'tri': lambda: triangulated_square(13, 11, 10.0, 8.0), 'tet': lambda: tetrahedral_box(5, 4, 6),
          'edge': lambda: cable(30, 2.0)}


@pytest.mark.parametrize('tag', ['tri', 'tet', 'edge'])
@pytest.mark.parametrize('rcm', [False, True])
def test_assembly_equals_oracle(tag, rcm):
    P, E, F = MESHES[tag]()
    rng = np.random.default_rng(3)
    F = rng.standard_normal(F.shape); F /= np.linalg.norm(F, axis=1, keepdims=True)
    (et, El), = E.items()
    El = El.copy(); El[:, -1] = 1 + (np.arange(El.shape[0]) % 2)
    E = {et: El}
    sl, st = {1: 0.9, 2: 0.4}, {1: 0.3, 2: 0.1}
    dom = Triangulation(); dom.set_mesh(P, E, F)
    mat = MaterialProperties()
    mat.add_element_property('sigma_l', 'region', sl)
    mat.add_element_property('sigma_t', 'region', st)
    conn = dom.mesh_connectivity(True)
    pat = compute_coo_pattern(conn)
    opat = ofem.coo_pattern(ofem.connectivity(P.shape[0], E))
    for k in ('I', 'J', 'StartIndex'):
        assert np.array_equal(pat[k], opat[k])
    ren = compute_reverse_cuthill_mckee_indexing(pat) if rcm else None
    if rcm:
        oren = ofem.rcm_indexing(opat)
        assert np.array_equal(ren['perm'], oren['perm']) and np.array_equal(ren['iperm'], oren['iperm'])
    M = assemble_matrices_dict({'mass': localMass, 'stiffness': localStiffness}, pat, dom, mat, conn, renumbering=ren)
    O = ofem.assemble(P, E, F, sl, st, None if ren is None else ren['iperm'])
    for name in ('mass', 'stiffness'):
        assert np.array_equal(M[name].rowptr, O['rowptr']) and np.array_equal(M[name].col, O['col'])
        assert np.array_equal(M[name].val, O[name])
    A = csr_axpby(M['mass'], 1.0, M['stiffness'], 0.1)
    assert np.array_equal(A.val, ofem.csr_axpby(O['mass'], 1.0, O['stiffness'], 0.1))
    # user-callback path (the examples' sigmaTens) gives the same stiffness
    def sigmaTens(elemtype, iElem, domain, matprop):
        fib = domain.Fibres()[iElem, :]
        rID = domain.Elems()[elemtype][iElem, -1]
        a, b = matprop.ElementProperty('sigma_l', elemtype, iElem, rID), matprop.ElementProperty('sigma_t', elemtype, iElem, rID)
        return b * np.eye(3) + (a - b) * np.outer(fib, fib)
    mat2 = MaterialProperties()
    mat2.add_element_property('sigma_l', 'elem', {et: [sl[int(r)] for r in El[:, -1]]})
    mat2.add_element_property('sigma_t', 'elem', {et: [st[int(r)] for r in El[:, -1]]})
    mat2.add_ud_function('stiffness', sigmaTens)
    M2 = assemble_matrices_dict({'stiffness': localStiffness}, pat, dom, mat2, conn, renumbering=ren)
    np.testing.assert_allclose(M2['stiffness'].val, M['stiffness'].val, rtol=1e-6, atol=1e-7)


def test_documented_size_of_the_square_mesh():
    # 63 001 nodes / 125 000 triangles / nnz 439 001
    P, E, F = triangulated_square()
    assert P.shape[0] == 63001 and E['Trias'].shape[0] == 125000
    dom = Triangulation(); dom.set_mesh(P, E, F)
    pat = compute_coo_pattern(dom.mesh_connectivity())
    assert pat['J'].shape[0] == 439001
    assert set(np.unique(E['Trias'][:, -1])) == {1, 2, 3, 4}


def test_triangulation_pickle_roundtrip_and_region_ids(tmp_path):
    P, E, F = triangulated_square(9, 9, 1.0, 1.0)
    fn = os.path.join(tmp_path, 'm.pkl')
    with open(fn, 'wb') as f:
        pickle.dump({'Pts': P, 'Elems': E, 'Fibres': F}, f)
    dom = Triangulation(); dom.readMesh(fn)
    assert np.array_equal(dom.Pts(), P) and dom.Elems()['Trias'].dtype == np.int32
    ids = dom.point_region_ids()
    assert ids.shape == (81,) and ids[0] == 1 and ids[-1] == 4
    with pytest.raises(NotImplementedError):
        dom.readMesh('mesh_without_extension')


def test_material_properties_lookup():
    m = MaterialProperties()
    m.add_element_property('a', 'uniform', 2.0)
    m.add_element_property('b', 'region', {1: 3.0})
    m.add_element_property('c', 'elem', {'Trias': [5.0, 6.0]})
    assert m.ElementProperty('a', 'Trias', 0, 1) == 2.0 and m.ElementProperty('b', 'Trias', 0, 1) == 3.0
    assert m.ElementProperty('c', 'Trias', 1, 1) == 6.0
    m.add_nodal_property('g', 'nodal', [1.0, 2.0])
    assert m.NodalProperty('g', 1, 0) == 2.0 and m.nodal_property_type('g') == 'nodal'
    assert m.execute_ud_func('nope') is None
    m.remove_all_element_properties()
    assert m.element_property_names() is None
