"""Pin oracle/fem.py: (1) against the golden vectors produced by the reference's own NumPy code
(tests/golden/make_fem_golden.py), (2) against the reference's unit-test properties
(Tests/CICD/unit/test_matrices.py:117-160, test_csr_axpby.py:62-96, test_conjgrad.py:117-162)."""
import os

import numpy as np
import pytest
from scipy.sparse import csr_matrix

from oracle import fem

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'fem_golden.npz'))
ET = {'tri': 'Trias', 'tet': 'Tetras', 'edge': 'Edges'}


@pytest.mark.parametrize('tag', ['tri', 'tet', 'edge'])
def test_local_matrices_match_reference(tag):
    et = ET[tag]
    P, E, f = G[tag + '_Pts'], G[tag + '_Elems'], G[tag + '_Fibres']
    b = fem.contravariant_basis(et, P, E[:, :-1])
    nb = {'Edges': 1, 'Trias': 2, 'Tetras': 3}[et]
    for k in ('v1', 'v2', 'v3')[:nb]:           # only the first nb vectors enter the matrices
        np.testing.assert_allclose(b[k], G['%s_%s' % (tag, k)], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(b['meas'], G[tag + '_meas'], rtol=1e-13)
    M = fem.local_mass(et, b['meas'])
    np.testing.assert_allclose(M, G[tag + '_lmass_batch'], rtol=1e-13)
    np.testing.assert_allclose(M, G[tag + '_lmass_elem'], rtol=1e-13)
    sl = {1: G['sigma_l'][0], 2: G['sigma_l'][1]}
    st = {1: G['sigma_t'][0], 2: G['sigma_t'][1]}
    K = fem.local_stiffness(et, b, b['meas'], fem.sigma_tensors(E, f, sl, st))
    np.testing.assert_allclose(K, G[tag + '_lstiff_batch'], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(K, G[tag + '_lstiff_elem'], rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize('tag', ['tri', 'tet', 'edge'])
def test_pattern_rcm_and_global_assembly_match_reference(tag):
    et = ET[tag]
    P, E, f = G[tag + '_Pts'], G[tag + '_Elems'], G[tag + '_Fibres']
    conn = fem.connectivity(P.shape[0], {et: E})
    pat = fem.coo_pattern(conn)
    assert np.array_equal(pat['I'], G[tag + '_pat_I']) and np.array_equal(pat['J'], G[tag + '_pat_J'])
    assert np.array_equal(pat['StartIndex'], G[tag + '_pat_start'])
    ren = fem.rcm_indexing(pat)
    assert np.array_equal(ren['perm'], G[tag + '_perm']) and np.array_equal(ren['iperm'], G[tag + '_iperm'])
    sl = {1: G['sigma_l'][0], 2: G['sigma_l'][1]}
    st = {1: G['sigma_t'][0], 2: G['sigma_t'][1]}
    A = fem.assemble(P, {et: E}, f, sl, st)
    # reference values are float64 in pattern order == CSR order without renumbering
    assert np.array_equal(A['rowptr'], pat['StartIndex']) and np.array_equal(A['col'], pat['J'])
    np.testing.assert_allclose(A['mass'], G[tag + '_glob_mass'], rtol=2e-6, atol=1e-9)
    np.testing.assert_allclose(A['stiffness'], G[tag + '_glob_stiff'], rtol=2e-5, atol=2e-6)
    # with RCM: the permuted matrix is P A P^T
    B = fem.assemble(P, {et: E}, f, sl, st, iperm=ren['iperm'])
    n = P.shape[0]
    Ad = csr_matrix((A['stiffness'], A['col'], A['rowptr']), shape=(n, n)).toarray()
    Bd = csr_matrix((B['stiffness'], B['col'], B['rowptr']), shape=(n, n)).toarray()
    np.testing.assert_allclose(Bd, Ad[np.ix_(ren['perm'], ren['perm'])], rtol=1e-6, atol=1e-6)


def _cable(n_el, L):
    x = np.linspace(0.0, L, n_el + 1)
    P = np.stack([x, np.zeros_like(x), np.zeros_like(x)], axis=1)
    E = np.stack([np.arange(n_el), np.arange(1, n_el + 1), np.ones(n_el, dtype=int)], axis=1).astype(np.int32)
    f = np.tile(np.array([[1.0, 0.0, 0.0]]), (n_el, 1))
    return P, {'Edges': E}, f


def test_1d_matrices_closed_form():
    # test_matrices.py:117-160: tridiagonal closed forms, symmetry, nnz = 3n-2, sum(M) = L, K 1 = 0, PSD
    n_el, L, sigma = 20, 2.0, 0.7
    P, E, f = _cable(n_el, L)
    A = fem.assemble(P, E, f, {1: sigma}, {1: sigma})
    n = n_el + 1
    h = L / n_el
    M = csr_matrix((A['mass'], A['col'], A['rowptr']), shape=(n, n)).toarray()
    K = csr_matrix((A['stiffness'], A['col'], A['rowptr']), shape=(n, n)).toarray()
    Mx = np.zeros((n, n)); Kx = np.zeros((n, n))
    for e in range(n_el):
        Mx[e:e + 2, e:e + 2] += h * np.array([[1 / 3, 1 / 6], [1 / 6, 1 / 3]])
        Kx[e:e + 2, e:e + 2] += sigma / h * np.array([[1, -1], [-1, 1]])
    np.testing.assert_allclose(M, Mx, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(K, Kx, rtol=1e-5, atol=1e-6)
    assert len(A['col']) == 3 * n - 2
    np.testing.assert_allclose(M, M.T, atol=1e-7); np.testing.assert_allclose(K, K.T, atol=1e-6)
    np.testing.assert_allclose(M.sum(), L, rtol=1e-5)
    np.testing.assert_allclose(K @ np.ones(n), 0.0, atol=1e-4)
    assert np.linalg.eigvalsh(K.astype(np.float64)).min() > -1e-4


def test_csr_axpby_matches_dense():
    rng = np.random.default_rng(0)
    a = rng.standard_normal(22).astype(np.float32)
    b = rng.standard_normal(22).astype(np.float32)
    np.testing.assert_allclose(fem.csr_axpby(a, 2.0, b, -0.5), 2.0 * a - 0.5 * b, atol=1e-6)
    np.testing.assert_allclose(fem.csr_axpby(a, 3.0), 3.0 * a, atol=1e-6)


def test_pcg_recovers_gold_solution():
    # test_conjgrad.py:117-162 on the synthetic stand-in of the (missing) coarse square mesh, reduced
    from pdensorflow_b200.gpuSolve.entities.synthetic import triangulated_square
    P, E, f = triangulated_square(41, 41, 10.0, 10.0)
    A = fem.assemble(P, E, f, None, None)
    val = fem.csr_axpby(A['mass'], 1.0, A['stiffness'], 1.0)
    n = A['n']
    minv = fem.jacobi(A['rowptr'], A['col'], val, n)
    for gold in (np.ones(n, dtype=np.float32), np.random.default_rng(0).standard_normal(n).astype(np.float32)):
        b = fem.spmv(A['rowptr'], A['col'], val, gold)
        x, nit, res = fem.pcg(A['rowptr'], A['col'], val, minv, b, np.zeros(n, np.float32), maxiter=1000,
                              toll=1e-6 * float(np.linalg.norm(b)))
        assert nit < 1000
        assert np.linalg.norm(x - gold) / np.linalg.norm(gold) < 1e-3
        assert np.linalg.norm(b - fem.spmv(A['rowptr'], A['col'], val, x)) / np.linalg.norm(b) < 1e-2


# ------------------------------------------------------------------------------------------------
# The reference's own end-to-end acceptance tests of MonodomainSolver + assembly + PCG + mMS, run on the ORACLE:
# Tests/CICD/unit/test_mms_1d.py:44-176 (cable) and Tests/CICD/nightly/test_mms_2d_regression.py:50-184 (sheet;
# its bundled mesh is missing from the checkout, the synthetic triangulated square has the same 0.04 spacing).
# ------------------------------------------------------------------------------------------------
def _lat(Vrec, times, vth):
    lat = np.full(Vrec.shape[1], np.nan)
    for j in range(Vrec.shape[1]):
        v = Vrec[:, j]
        k = int(np.argmax((v[:-1] < vth) & (v[1:] >= vth)))
        if (v[k] < vth) and (v[k + 1] >= vth):
            lat[j] = times[k] + (vth - v[k]) / (v[k + 1] - v[k]) * (times[k + 1] - times[k])
    return lat


def _mms_front(Pts, elems, fibres, sigma, dt, tend, xfrac, rcm):
    from oracle import ionic as oi
    ionic = oi.ModifiedMS2v(dt=dt)
    sig = {r: sigma for r in (1, 2, 3, 4)}
    orc = fem.MonodomainOracle(ionic, Pts, elems, fibres, sig, sig, dt, use_renumbering=rcm, maxiter=Pts.shape[0] // 2)
    x = Pts[:, 0].copy()
    U0 = np.where(x < xfrac * x.max(), 20.0, -80.0).astype(np.float32)
    orc.U = U0[:, None].copy()
    ionic.initialize_state_variables(orc.U)
    orc.finalize_for_run()
    nt = int(tend // dt)
    Vrec = np.empty((nt + 1, x.shape[0]), dtype=np.float32)
    times = np.zeros(nt + 1)
    Vrec[0] = orc.U_user().ravel()
    t = 0.0
    for i in range(nt):
        t += dt
        orc.step(t)
        Vrec[i + 1] = orc.U_user().ravel()
        times[i + 1] = t
    lat = _lat(Vrec, times, -30.0)
    Lx = x.max()
    mask = (x >= 0.4 * Lx) & (x <= 0.8 * Lx) & np.isfinite(lat)
    cv = 1.0 / float(np.polyfit(x[mask], lat[mask], 1)[0])
    u_crit, tau_in = float(ionic.get_parameter('u_crit')), float(ionic.get_parameter('tau_in'))
    return x, lat, Vrec, cv, 0.5 * (1.0 - 2.0 * u_crit) * np.sqrt(2.0 * sigma / tau_in)


def test_reference_mms_1d_acceptance_on_the_oracle():
    from pdensorflow_b200.gpuSolve.entities.synthetic import cable
    P, E, F = cable(150, 2.0)
    x, lat, Vrec, cv, cv_an = _mms_front(P, E, F, 0.1, 0.05, 7.0, 0.05, False)
    interior = (x >= 0.4 * 2.0) & (x <= 0.8 * 2.0)
    assert np.all(np.isfinite(Vrec)) and np.all(np.isfinite(lat[interior])) and np.all(np.diff(lat[interior]) > 0.0)
    assert Vrec.min() > -81.0 and Vrec.max() < 21.0
    assert abs(cv - cv_an) / cv_an < 0.10, (cv, cv_an)


def test_reference_mms_2d_regression_on_the_oracle():
    from pdensorflow_b200.gpuSolve.entities.synthetic import triangulated_square
    P, E, F = triangulated_square(126, 126, 5.0, 5.0)          # 0.04 spacing as the reference's sheet, a quarter of its area
    x, lat, Vrec, cv, cv_an = _mms_front(P, E, F, 0.5, 0.05, 5.0, 0.05, True)
    Lx = x.max()
    assert np.all(np.isfinite(Vrec))
    xr = np.round(x, 6)
    xu = np.sort(np.unique(xr))
    xu = xu[(xu >= 0.4 * Lx) & (xu <= 0.8 * Lx)]
    col = np.array([np.nanmean(lat[xr == v]) for v in xu])
    assert np.all(np.isfinite(col)) and np.all(np.diff(col) > 0.0)
    assert Vrec.min() > -81.0 and Vrec.max() < 21.0
    assert abs(cv - cv_an) / cv_an < 0.10, (cv, cv_an)
