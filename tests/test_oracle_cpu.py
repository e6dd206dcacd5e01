"""CPU checks of the oracle itself: the reference's own unit-test properties
(Tests/CICD/unit/test_ionic.py:66-97), the stimulus windows worked out in SURVEY.md 8a row S,
and structural properties of the FD restatement."""
import numpy as np
import pytest

from oracle import fd, ionic, stimulus

_DT = 0.01


@pytest.mark.parametrize('name', ['fenton4v', 'mms2v', 'ms2v'])
def test_ionic_rest_fixed_point(name):
    # test_ionic.py: to_dimensionless(vmin)=0, (vmax)=1, dU(rest)=0 (atol 1e-6), finite, deterministic
    m = ionic.MODELS[name](dt=_DT)
    U = np.full((4, 1), -80.0, dtype=np.float32)
    m.initialize_state_variables(U)
    assert m.to_dimensionless(np.float32(-80.0)) == 0.0
    assert m.to_dimensionless(np.float32(20.0)) == 1.0
    dU = m.differentiate(U)
    assert dU.shape == U.shape and dU.dtype == np.float32
    assert np.all(np.isfinite(dU))
    assert np.max(np.abs(dU)) < 0.5
    np.testing.assert_allclose(dU, 0.0, atol=1e-6)
    m2 = ionic.MODELS[name](dt=_DT)
    m2.initialize_state_variables(U)
    assert np.array_equal(m2.differentiate(U), dU)


def test_fenton_gate_branches():
    m = ionic.Fenton4v(dt=0.1)
    U = np.array([[-80.0], [-60.0], [-50.0], [0.0], [20.0]], dtype=np.float32)   # u = 0, .2, .3, .8, 1
    m.initialize_state_variables(U)
    m.differentiate(U)
    V = m._V_state.ravel()
    # u <= u_c: dV = (1-V)/tau_vn = 0 ; u > u_c: dV = -V/tau_vp
    assert V[0] == 1.0 and V[1] == 1.0
    np.testing.assert_allclose(V[2:], 1.0 - 0.1 / 3.33, rtol=1e-6)


def test_stimulus_fd_integer_step():
    # Tests/FD/Fenton/fenton.py:134-140: tstart 200, duration dt, period 800 => fires at i == 2000 only
    s = stimulus.Stimulus({'tstart': 200, 'nstim': 1, 'period': 800, 'duration': 0.1, 'intensity': 60.0})
    fired = [i for i in range(0, 4000) if s.stimulate_tissue_timestep(i, 0.1)]
    assert fired == [2000]


def test_stimulus_fem_float_time_windows():
    # config 1 (Tests/FEM/Fenton/fenton.py:73-87,123-126): float64 running clock evaluated in fp32
    dt = 0.1
    s1 = stimulus.Stimulus({'tstart': 0.0, 'nstim': 1, 'period': 100, 'duration': np.max([0.4, dt]), 'intensity': 60.0})
    s2 = stimulus.Stimulus({'tstart': 210.0, 'nstim': 1, 'period': 100, 'duration': np.max([0.4, dt]), 'intensity': 60.0})
    f1, f2 = [], []
    ctime = 0.0
    for i in range(2200):
        ctime += dt
        t = np.float32(ctime)
        if s1.stimulate_tissue_timevalue(t):
            f1.append(i)
        if s2.stimulate_tissue_timevalue(t):
            f2.append(i)
    assert f1 == [0, 1, 2, 3]
    assert f2 == [2099, 2100, 2101, 2102, 2103]


def test_all_ones_mask_quirk():
    # F6: the shipped S2 mask is filled with -80/+20, astype(bool) makes it all ones
    s2_init = np.full((4, 4, 4), -80.0, dtype=np.float32)
    s2_init[:2, :2, :] = 20.0
    s = stimulus.Stimulus({'intensity': 60.0})
    s.set_stimregion(s2_init)
    assert np.all(s.get_stimregion() == 1.0)


def test_enforce_boundary_is_index_clamp():
    rng = np.random.default_rng(0)
    X = rng.standard_normal((6, 5, 7)).astype(np.float32)
    X0 = fd.enforce_boundary(X)
    ii = np.clip(np.arange(6), 1, 4)
    jj = np.clip(np.arange(5), 1, 3)
    kk = np.clip(np.arange(7), 1, 5)
    assert np.array_equal(X0, X[np.ix_(ii, jj, kk)])


def test_laplacians_annihilate_constants_and_agree():
    X = np.full((5, 6, 7), 3.25, dtype=np.float32)
    assert np.all(fd.laplace_homog_3d(X, 1.0, 0.5, 2.0) == 0.0)
    assert np.all(fd.laplace_heterog_3d(X, np.ones_like(X), 1.0, 0.5, 2.0) == 0.0)
    rng = np.random.default_rng(1)
    X = rng.standard_normal((8, 9, 10)).astype(np.float32)
    a = fd.laplace_homog_3d(X, 1.0, 1.0, 1.0)
    # heterogeneous operator with D == 1 is the homogeneous one up to rounding
    b = fd.laplace_heterog_3d(X, np.ones_like(X), 1.0, 1.0, 1.0)
    np.testing.assert_allclose(a, b, rtol=0, atol=2e-5)
    # the conv alias equals the slicing form when DX == DZ (the kernel as written swaps them)
    c = fd.laplace_conv_homog_3d(X, 0.5, 2.0, 0.5)
    d = fd.laplace_homog_3d(X, 0.5, 2.0, 0.5)
    np.testing.assert_allclose(c, d, rtol=0, atol=2e-4)
    # quadratic along one axis, away from the boundary: exact second difference
    q = (np.arange(8, dtype=np.float32) ** 2)[:, None, None] * np.ones((8, 9, 10), dtype=np.float32)
    np.testing.assert_allclose(fd.laplace_homog_3d(q, 1.0, 1.0, 1.0)[1:-1], 2.0, atol=1e-5)


def test_boundary_cell_rule():
    # a boundary cell's new value is its interior neighbour's old value plus its own
    # reaction term: for a field varying along x only (heat, no reaction) U1[0] == U[1] exactly,
    # because the flux through the outermost two faces is zero.
    rng = np.random.default_rng(2)
    U = (rng.uniform(0, 1, (6, 1, 1)) * np.ones((6, 5, 7))).astype(np.float32)
    U1 = fd.fd_step(None, U, 0.1, 0.8, (1.0, 1.0, 1.0))
    assert np.array_equal(U1[0], U[1]) and np.array_equal(U1[-1], U[-2])
    assert not np.array_equal(U1[1], U[1])


def test_plane_wave_stays_planar():
    m = ionic.Fenton4v()
    m._dt = 0.1
    U = np.full((24, 6, 8), -80.0, dtype=np.float32)
    U[0:2] = 20.0
    m.initialize_state_variables(U)
    U = fd.fd_run(m, U, 50, 0.1, 0.8, (1.0, 1.0, 1.0))
    assert np.all(U == U[:, :1, :1])
    assert U.max() > -30.0 and U.min() >= -81.0
