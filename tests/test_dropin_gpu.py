"""The drop-in contracts (operator functions, IonicModel protocol, Stimulus, example-style solve())
on the GPU against the oracle; mirrors the reference's unit tests (Tests/CICD/unit/test_ionic.py:66-97)."""
import numpy as np
import pytest
import torch

from helpers import random_fields, rel_l2
from oracle import fd as ofd
from oracle import ionic as oionic
from oracle import stimulus as ostim

pytestmark = pytest.mark.gpu
_DT = 0.01


def _models():
    from pdensorflow_b200.gpuSolve.ionic import Fenton4v, MitchellSchaeffer2v, ModifiedMS2v
    return {'fenton4v': Fenton4v, 'mms2v': ModifiedMS2v, 'ms2v': MitchellSchaeffer2v}


@pytest.mark.parametrize('name', ['fenton4v', 'mms2v', 'ms2v'])
def test_ionic_rest_state_like_the_reference_unit_test(cuda, name):
    m = _models()[name](dt=_DT)
    U = torch.full((4, 1), -80.0, device='cuda')
    m.initialize_state_variables(U)
    dU = m.differentiate(U)
    assert dU.shape == U.shape and bool(torch.isfinite(dU).all())
    assert float(dU.abs().max()) < 0.5
    np.testing.assert_allclose(dU.cpu().numpy(), 0.0, atol=1e-6)
    assert float(m.to_dimensionless(torch.tensor(-80.0))) == 0.0
    assert float(m.to_dimensionless(torch.tensor(20.0))) == 1.0
    m2 = _models()[name](dt=_DT)
    m2.initialize_state_variables(U)
    assert torch.equal(m2.differentiate(U), dU)                      # bit-deterministic from a fresh instance


@pytest.mark.parametrize('name', ['fenton4v', 'mms2v', 'ms2v'])
@pytest.mark.parametrize('math', ['exact', 'fast'])
def test_differentiate_matches_oracle(cuda, name, math):
    shape = (7, 9, 11)
    ns = {'fenton4v': 3, 'mms2v': 1, 'ms2v': 1}[name]
    U, states = random_fields(shape, seed=31, nstates=ns)
    m = _models()[name](dt=0.1)
    m.math_mode = math
    Ud = torch.from_numpy(U).cuda()
    m.initialize_state_variables(Ud)
    o = oionic.MODELS[name](dt=0.1)
    o.initialize_state_variables(U)
    for attr, s in zip(oionic.STATE_ATTRS[name], states):
        setattr(m, attr, torch.from_numpy(s).cuda())
        setattr(o, attr, s.copy())
    for _ in range(3):                                               # states advance in place across calls
        dU = m.differentiate(Ud)
        dUo = o.differentiate(U)
    tol = 1e-6 if math == 'exact' else 2e-5
    assert rel_l2(dU.cpu().numpy(), dUo) <= tol
    for attr in oionic.STATE_ATTRS[name]:
        assert rel_l2(getattr(m, attr).cpu().numpy(), getattr(o, attr)) <= tol
    if math == 'exact' and name != 'fenton4v':
        assert np.array_equal(dU.cpu().numpy(), dUo)                 # no transcendental: bit exact


def test_set_get_parameter_protocol(cuda):
    from pdensorflow_b200.gpuSolve.ionic import Fenton4v
    m = Fenton4v(dt=0.1)
    assert m.get_parameter('nonexistent') is None
    m.set_parameter('nonexistent', 3.0)                               # silently ignored, like the reference
    assert not hasattr(m, '_nonexistent')
    m.set_parameter('tau_d', 0.1)
    assert float(m.get_parameter('tau_d')) == pytest.approx(0.1) and float(m.tau_d()) == pytest.approx(0.1)
    U = torch.full((5, 1), 0.0, device='cuda')
    m.initialize_state_variables(U)
    o = oionic.Fenton4v(dt=0.1); o.set_parameter('tau_d', 0.1)
    o.initialize_state_variables(np.zeros((5, 1), np.float32))
    assert rel_l2(m.differentiate(U).cpu().numpy(), o.differentiate(np.zeros((5, 1), np.float32))) < 1e-6

    class Sub(Fenton4v):                                              # user scripts subclass the model
        def __init__(self):
            super().__init__()
            self.dt = 0.1
    s = Sub(); s._dt = s.dt
    s.initialize_state_variables(U)
    assert s.differentiate(U).shape == U.shape
    with pytest.raises(Exception):
        Fenton4v(dt=0.1).differentiate(U)                             # states not initialised


@pytest.mark.parametrize('shape', [(9, 8, 7), (16, 12, 24)])
def test_3d_operators_match_oracle_bitwise(cuda, shape):
    from pdensorflow_b200.gpuSolve.diffop3D import laplace_conv_homog, laplace_heterog, laplace_heterog_aniso, laplace_homog
    rng = np.random.default_rng(41)
    X = rng.standard_normal(shape).astype(np.float32)
    D = rng.uniform(0, 1, shape).astype(np.float32)
    Xd, Dd = torch.from_numpy(X).cuda(), torch.from_numpy(D).cuda()
    DX, DY, DZ = torch.tensor(0.5), 1.25, np.float32(2.0)
    assert np.array_equal(laplace_homog(Xd, DX, DY, DZ).cpu().numpy(), ofd.laplace_homog_3d(X, 0.5, 1.25, 2.0))
    assert np.array_equal(laplace_heterog(Xd, Dd, DX, DY, DZ).cpu().numpy(), ofd.laplace_heterog_3d(X, D, 0.5, 1.25, 2.0))
    assert np.array_equal(laplace_conv_homog(Xd, DX, DY, DZ).cpu().numpy(), ofd.laplace_conv_homog_3d(X, 0.5, 1.25, 2.0))
    DF = rng.uniform(0.1, 1.0, shape + (2,)).astype(np.float32)
    AV = rng.uniform(-1.0, 1.0, shape + (6,)).astype(np.float32)
    got = laplace_heterog_aniso(Xd, torch.from_numpy(DF).cuda(), torch.from_numpy(AV).cuda(), DX, DY, DZ)
    assert np.array_equal(got.cpu().numpy(), ofd.laplace_heterog_aniso_3d(X, DF, AV, 0.5, 1.25, 2.0))
    with pytest.raises(Exception):
        laplace_heterog_aniso(Xd, Dd, Dd, DX, DY, DZ)                # wrong channel counts
    with pytest.raises(Exception):
        laplace_homog(torch.zeros(4, 4), 1.0, 1.0, 1.0)              # wrong rank


def test_2d_operators_match_oracle_bitwise(cuda):
    from pdensorflow_b200.gpuSolve.diffop2D import laplace_heterog, laplace_homog
    rng = np.random.default_rng(43)
    X = rng.standard_normal((13, 22)).astype(np.float32)
    D = rng.uniform(0, 1, (13, 22)).astype(np.float32)
    Xd, Dd = torch.from_numpy(X).cuda(), torch.from_numpy(D).cuda()
    assert np.array_equal(laplace_homog(Xd, 0.7, 1.3).cpu().numpy(), ofd.laplace_homog_2d(X, 0.7, 1.3))
    assert np.array_equal(laplace_heterog(Xd, Dd, 0.7, 1.3).cpu().numpy(), ofd.laplace_heterog_2d(X, D, 0.7, 1.3))


def test_enforce_boundary_and_dirichlet(cuda):
    from pdensorflow_b200.gpuSolve.fd import enforce_boundary
    rng = np.random.default_rng(47)
    for shape in [(6, 5, 7), (8, 9)]:
        X = rng.standard_normal(shape).astype(np.float32)
        Xd = torch.from_numpy(X).cuda()
        assert np.array_equal(enforce_boundary(Xd).cpu().numpy(), ofd.enforce_boundary(X))
        assert np.array_equal(enforce_boundary(Xd, dirichlet_value=1.5).cpu().numpy(), ofd.enforce_boundary_dirichlet(X, 1.5))


def test_example_style_solve_unfused_and_fused(cuda):
    """The examples' solve() written with the drop-in pieces (Tests/FD/Fenton/fenton.py:90-99), and the
    single-kernel fused_solve, against the oracle step by step (S2 applied like fenton.py:154-156)."""
    from pdensorflow_b200.gpuSolve.diffop3D import laplace_homog as laplace
    from pdensorflow_b200.gpuSolve.fd import apply_stimulus, enforce_boundary, fused_solve
    from pdensorflow_b200.gpuSolve.force_terms import Stimulus
    from pdensorflow_b200.gpuSolve.ionic.fenton4v import Fenton4v
    shape, dt, diff = (12, 10, 16), 0.1, 0.8
    u_init = np.full(shape, -80.0, dtype=np.float32); u_init[0:2] = 20.0
    s2_init = np.zeros(shape, dtype=np.float32); s2_init[:6, :5, :] = 1.0
    results = {}
    for mode in ('unfused', 'fused'):
        model = Fenton4v(); model._dt = dt
        U = torch.from_numpy(u_init).cuda()
        model.initialize_state_variables(U)
        s2 = Stimulus({'tstart': 1.0, 'nstim': 1, 'period': 800, 'duration': dt, 'intensity': 60.0})
        s2.set_stimregion(s2_init)
        for i in range(25):
            if mode == 'unfused':
                U0 = enforce_boundary(U)
                dU = model.differentiate(U)
                U = U0 + dt * dU + diff * dt * laplace(U0, 1.0, 1.0, 1.0)
            else:
                U = fused_solve(model, U, dt, diff, (1.0, 1.0, 1.0), math='exact')
            if s2.stimulate_tissue_timestep(i, dt):
                U = apply_stimulus(U, s2)
        results[mode] = U.cpu().numpy()
    o = oionic.Fenton4v(); o._dt = dt
    o.initialize_state_variables(u_init)
    so = ostim.Stimulus({'tstart': 1.0, 'nstim': 1, 'period': 800, 'duration': dt, 'intensity': 60.0})
    so.set_stimregion(s2_init)
    Uo = ofd.fd_run(o, u_init.copy(), 25, dt, diff, (1.0, 1.0, 1.0), s2=so)
    assert Uo.max() > 0.0
    assert rel_l2(results['fused'], Uo) <= 1e-6
    assert rel_l2(results['unfused'], Uo) <= 1e-6


def test_stimulus_mirror_timing_equals_oracle(cuda):
    from pdensorflow_b200.gpuSolve.force_terms import Stimulus
    props = {'tstart': 210.0, 'nstim': 2, 'period': 100, 'duration': np.max([0.4, 0.1]), 'intensity': 60.0}
    a, b = Stimulus(props), ostim.Stimulus(props)
    mask = np.array([1, 0, 1, 0], dtype=bool)
    a.set_stimregion(mask); b.set_stimregion(mask)
    assert a.get_stimregion().shape == (4, 1)
    ctime = 0.0
    for i in range(3300):
        ctime += 0.1
        t = np.float32(ctime)
        assert a.stimulate_tissue_timevalue(t) == b.stimulate_tissue_timevalue(t), i
    c, d = Stimulus({'tstart': 200, 'period': 800, 'duration': 0.1}), ostim.Stimulus({'tstart': 200, 'period': 800, 'duration': 0.1})
    for i in range(2500):
        assert c.stimulate_tissue_timestep(i, 0.1) == d.stimulate_tissue_timestep(i, 0.1)
    assert np.array_equal(a().cpu().numpy(), b())


def test_example_scripts_run(cuda, tmp_path, monkeypatch):
    """examples/fd_fenton.py and examples/fem_fenton.py (the reference's two flagship examples on the drop-in package)
    run end to end on small sizes and leave readable result files."""
    import importlib.util
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def load(name):
        spec = importlib.util.spec_from_file_location(name, os.path.join(root, 'examples', name + '.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    monkeypatch.setattr(sys, 'argv', ['fd_fenton.py', '--size', '24', '--samples', '60', '--out', str(tmp_path / 'cube3D')])
    load('fd_fenton').main()
    cube = np.load(str(tmp_path / 'cube3D_24_24_24.npy'))
    assert cube.shape == (7, 24, 24, 24) and np.isfinite(cube).all() and cube[-1].max() > 0.0
    monkeypatch.setattr(sys, 'argv', ['fem_fenton.py', '--n', '31', '--tend', '4', '--out', str(tmp_path / 'sq.igb')])
    load('fem_fenton').main()
    from pdensorflow_b200.gpuSolve.IO.readers import IGBReader
    r = IGBReader()
    r.read(str(tmp_path / 'sq.igb'))
    assert r.nx() == 31 * 31 and r.nt() == 4 and r.ndiff() == 0 and r.data().max() > 0.0
