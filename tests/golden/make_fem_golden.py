"""Generate tests/golden/fem_golden.npz from the REFERENCE's own pure-NumPy code.

Run in the build container (needs /root/reference; the GPU box never runs this):
    python tests/golden/make_fem_golden.py

TensorFlow cannot be imported here, so only the TF-free pieces of the reference are executed:
  gpuSolve/matrices/localMass.py, localStiffness.py          (whole files, by path)
  gpuSolve/entities/materialproperties.py                    (whole file, by path)
  gpuSolve/matrices/globalMatrices.py lines 1-414            (everything above the
      "functions that REQUIRE TensorFlow" banner, exec'd from the source text)
The meshes are small seeded perturbations of structured triangle / tetrahedron / edge grids
with non-trivial fibres and two regions.
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = '/root/reference/PDEnsorflow/gpuSolve'
HERE = os.path.dirname(os.path.abspath(__file__))


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    for pkg in ('gpuSolve', 'gpuSolve.matrices', 'gpuSolve.entities'):
        sys.modules.setdefault(pkg, types.ModuleType(pkg))
    lm = load(os.path.join(REF, 'matrices', 'localMass.py'), 'gpuSolve.matrices.localMass')
    ls = load(os.path.join(REF, 'matrices', 'localStiffness.py'), 'gpuSolve.matrices.localStiffness')
    mp = load(os.path.join(REF, 'entities', 'materialproperties.py'), 'gpuSolve.entities.materialproperties')
    src = open(os.path.join(REF, 'matrices', 'globalMatrices.py')).read()
    cut = src.index('functions that REQUIRE TensorFlow')
    cut = src.rfind('\n####', 0, cut)
    ns = {'__name__': 'ref_globalMatrices'}
    exec(compile(src[:cut], 'globalMatrices.py[:TF banner]', 'exec'), ns)
    return lm, ls, mp, ns


class FakeDomain:
    """Just what assemble_vectmat_dict touches on a Triangulation."""

    def __init__(self, Pts, Elems, Fibres):
        self._P, self._E, self._F = Pts, Elems, Fibres

    def Pts(self):
        return self._P

    def Elems(self):
        return self._E

    def Fibres(self):
        return self._F


def tri_mesh(nx, ny, rng):
    xs, ys = np.meshgrid(np.linspace(0, 1.0, nx), np.linspace(0, 0.8, ny), indexing='ij')
    P = np.stack([xs.ravel(), ys.ravel(), np.zeros(nx * ny)], axis=1)
    P[:, :2] += rng.uniform(-0.02, 0.02, size=(nx * ny, 2))
    P[:, 2] = 0.05 * np.sin(3 * P[:, 0])                 # a curved sheet: 3-D triangles
    idx = lambda i, j: i * ny + j
    T = []
    for i in range(nx - 1):
        for j in range(ny - 1):
            reg = 1 if i < nx // 2 else 2
            T.append([idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), reg])
            T.append([idx(i, j), idx(i + 1, j + 1), idx(i, j + 1), reg])
    T = np.array(T, dtype=np.int32)
    f = rng.standard_normal((T.shape[0], 3)); f /= np.linalg.norm(f, axis=1, keepdims=True)
    return P, {'Trias': T}, f


def tet_mesh(n, rng):
    g = np.linspace(0, 1.0, n)
    X, Y, Z = np.meshgrid(g, g, g, indexing='ij')
    P = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1) + rng.uniform(-0.03, 0.03, size=(n ** 3, 3))
    idx = lambda i, j, k: (i * n + j) * n + k
    T = []
    for i in range(n - 1):
        for j in range(n - 1):
            for k in range(n - 1):
                c = [idx(i, j, k), idx(i + 1, j, k), idx(i + 1, j + 1, k), idx(i, j + 1, k),
                     idx(i, j, k + 1), idx(i + 1, j, k + 1), idx(i + 1, j + 1, k + 1), idx(i, j + 1, k + 1)]
                reg = 1 if k < (n - 1) // 2 else 2
                for a, b, d in ((1, 2, 6), (1, 6, 5), (2, 3, 6), (3, 7, 6), (5, 6, 4), (6, 7, 4)):
                    # Kuhn-type split around the diagonal c0-c6
                    T.append([c[0], c[a], c[b], c[d], reg])
    T = np.array(T, dtype=np.int32)
    f = rng.standard_normal((T.shape[0], 3)); f /= np.linalg.norm(f, axis=1, keepdims=True)
    return P, {'Tetras': T}, f


def edge_mesh(n, rng):
    x = np.sort(rng.uniform(0, 2.0, size=n))
    P = np.stack([x, 0.1 * x ** 2, np.zeros(n)], axis=1)
    E = np.stack([np.arange(n - 1), np.arange(1, n), 1 + (np.arange(n - 1) >= n // 2)], axis=1).astype(np.int32)
    f = np.tile(np.array([[1.0, 0.0, 0.0]]), (n - 1, 1))
    return P, {'Edges': E}, f


def main():
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import fem as ofem         # connectivity restatement feeds the reference's pattern code
    lm, ls, mp, ns = load_reference()
    rng = np.random.default_rng(20261019)
    out = {}
    sig_l, sig_t = {1: 0.9, 2: 0.35}, {1: 0.2, 2: 0.15}
    for tag, (P, E, f) in {'tri': tri_mesh(5, 4, rng), 'tet': tet_mesh(3, rng), 'edge': edge_mesh(9, rng)}.items():
        (et, El), = E.items()
        nodes = El[:, :-1]
        basis = ns['_batch_contrabasis_dispatch'][et](P, nodes)
        meas = basis['meas']
        Sig = np.zeros((El.shape[0], 3, 3))
        for e in range(El.shape[0]):
            r = int(El[e, -1])
            Sig[e] = sig_t[r] * np.eye(3) + (sig_l[r] - sig_t[r]) * np.outer(f[e], f[e])
        lmass_b = ns['_batch_local_mass_vectorized'](et, meas)
        lstiff_b = ns['_batch_local_stiffness_vectorized'](et, basis, meas, Sig)
        # the per-element functions of localMass.py / localStiffness.py
        lmass_e = np.stack([lm.localMass(et, {'meas': meas[e, 0]}) for e in range(El.shape[0])])
        lstiff_e = np.stack([ls.localStiffness(et, {'meas': meas[e, 0], 'v1': basis['v1'][e], 'v2': basis['v2'][e],
                                                    'v3': basis['v3'][e]}, Sig[e]) for e in range(El.shape[0])])
        conn = {i: c for i, c in enumerate(ofem.connectivity(P.shape[0], E))}
        pat = ns['compute_coo_pattern'](conn)
        ren = ns['compute_reverse_cuthill_mckee_indexing'](pat)
        props = mp.MaterialProperties()
        props.add_element_property('sigma_l', 'region', sig_l)
        props.add_element_property('sigma_t', 'region', sig_t)
        props.add_ud_function('mass', lambda *a: None)
        props.add_ud_function('stiffness', lambda *a: None)
        VM = ns['assemble_vectmat_dict']({'mass': lm.localMass, 'stiffness': ls.localStiffness}, pat,
                                         FakeDomain(P, E, f), props, conn)
        for k, v in dict(Pts=P, Elems=El, Fibres=f, v1=basis['v1'], v2=basis['v2'], v3=basis['v3'], meas=meas,
                         lmass_batch=lmass_b, lstiff_batch=lstiff_b, lmass_elem=lmass_e, lstiff_elem=lstiff_e,
                         pat_I=pat['I'], pat_J=pat['J'], pat_start=pat['StartIndex'], perm=ren['perm'],
                         iperm=ren['iperm'], glob_mass=VM['mass'], glob_stiff=VM['stiffness']).items():
            out['%s_%s' % (tag, k)] = np.asarray(v)
    out['sigma_l'] = np.array([sig_l[1], sig_l[2]]); out['sigma_t'] = np.array([sig_t[1], sig_t[2]])
    np.savez_compressed(os.path.join(HERE, 'fem_golden.npz'), **out)
    print('wrote fem_golden.npz with %d arrays' % len(out))


if __name__ == '__main__':
    main()
