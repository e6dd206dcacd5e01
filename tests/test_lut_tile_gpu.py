"""Tile kernels of the lookup-table models (pdx_ionic_lut.cu: state rows through the copy engine into shared memory,
128-bit table rows, vectorised stencil) against the scalar one-thread-per-node kernels they replace.  Both evaluate the
reference's formulas op for op, so EXACT arithmetic must agree BITWISE - for every tile shape: full tiles, a short
last tile, several tiles per plane, planes smaller than a tile.  (Both kernels are pinned to the reference-source
fixtures and the oracle in tests/test_ref_golden_gpu.py.)"""
import os

import numpy as np
import pytest
import torch

from helpers import rel_l2

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_ionic.npz')


def _model(tag, dt=0.02, math='exact'):
    from pdensorflow_b200.gpuSolve.ionic import CourtemancheRamirezNattel, TenTusscherPanfilov
    m = (CourtemancheRamirezNattel if tag == 'crn' else TenTusscherPanfilov)(dt=dt)
    m.math_mode = math
    return m


def _seeded_inputs(tag, n, seed):
    """n nodes drawn (with replacement) from the reference fixture's physiological states, potentials jittered."""
    g = np.load(GOLD)
    rng = np.random.default_rng(seed)
    pick = rng.integers(0, g[tag + '1_U'].shape[0], size=n)
    U = (g[tag + '1_U'][pick, 0] + rng.uniform(-0.5, 0.5, size=n)).astype(np.float32)
    names = [str(s) for s in g[tag + '1_names']]
    return U, {s: g['%s1_%s_in' % (tag, s)][pick, 0].astype(np.float32) for s in names}


def _pointwise(tag, n, math, tile, monkeypatch, with_rhs=False):
    monkeypatch.setenv('PDX_LUT_TILE', '1' if tile else '0')
    U, st = _seeded_inputs(tag, n, seed=n)
    m = _model(tag, math=math)
    Ud = torch.from_numpy(U.reshape(n, 1)).cuda()
    m.initialize_state_variables(Ud)
    for k, v in st.items():
        setattr(m, k, torch.from_numpy(v.reshape(n, 1)).cuda())
    out = [m.differentiate(Ud).cpu().numpy()]
    out += [getattr(m, k).cpu().numpy() for k in sorted(st)]
    out.append(m.differentiate(Ud).cpu().numpy())          # a second call on the updated states
    return out


@pytest.mark.parametrize('tag', ['crn', 'ttp'])
@pytest.mark.parametrize('n', [4, 96, 508, 512, 516, 4100])
def test_pointwise_tile_kernel_equals_scalar_kernel(cuda, tag, n, monkeypatch):
    a = _pointwise(tag, n, 'exact', True, monkeypatch)
    b = _pointwise(tag, n, 'exact', False, monkeypatch)
    for x, y in zip(a, b):
        assert np.isfinite(x).all() and np.array_equal(x, y)
    fa = _pointwise(tag, n, 'fast', True, monkeypatch)
    for x, y in zip(fa, a):            # FAST: fp32 expf/logf, approximate division, FMA contraction (a concentration
        assert rel_l2(x, y) <= 1e-4    # increment is a difference of large currents: compare arrays, not elements)


@pytest.mark.parametrize('tag', ['crn', 'ttp'])
@pytest.mark.parametrize('shape,lap', [((5, 7, 132), 'homog'), ((4, 130, 8), 'homog'), ((3, 40, 64), 'heterog'),
                                       ((6, 9, 16), 'conv'), ((1, 33, 36), 'homog2d')])
def test_fused_tile_kernel_equals_scalar_kernel(cuda, tag, shape, lap, monkeypatch):
    """Planes of 924 / 1040 / 2560 / 144 / 1188 nodes = short tiles, exact multiples and ragged tails; y-z boundary
    clamps inside the 128-bit stencil loads; 2-D grids (nx = 1)."""
    from pdensorflow_b200.fd import FDStepper
    two_d = lap == 'homog2d'
    lap_ = 'homog' if two_d else lap
    shp = shape[1:] if two_d else shape
    rng = np.random.default_rng(7)
    n = int(np.prod(shp))
    U0, st = _seeded_inputs(tag, n, seed=3)
    Dm = (0.1 * (rng.uniform(size=shp) > 0.2)).astype(np.float32) if lap == 'heterog' else None
    res = []
    for tile in ('1', '0'):
        monkeypatch.setenv('PDX_LUT_TILE', tile)
        m = _model(tag)
        s = FDStepper(shp, (0.25,) * len(shp), 0.02, 0.1, model=m, lap=lap_, DIFF=Dm, math='exact')
        s.set_U(U0.reshape(shp))
        for k, v in st.items():
            getattr(m, k).copy_(torch.from_numpy(v.reshape(shp)).cuda())
        s.step(5)
        torch.cuda.synchronize()
        res.append([s.U().cpu().numpy()] + [t.cpu().numpy() for t in s.states])
    for x, y in zip(*res):
        assert np.isfinite(x).all() and np.array_equal(x, y)


@pytest.mark.parametrize('tag', ['crn', 'ttp'])
def test_fused_tile_kernel_plane_wave_vs_oracle_256_columns(cuda, tag):
    """A benchmark-shaped run (FAST, several full tiles per plane): the planar S1 wave on 24 x 64 x 64 must stay planar
    and match the oracle on a 24 x 4 x 4 grid to the north-star tolerance after 80 steps."""
    from oracle import fd as ofd
    from oracle import ionic_lut as ol
    from pdensorflow_b200.fd import FDStepper
    dt, diff, spacing, nsteps = 0.02, 0.1, (0.25, 0.25, 0.25), 80
    mo = ol.CourtemancheRamirezNattel() if tag == 'crn' else ol.TenTusscherPanfilov()
    small = (24, 4, 4)
    Uo = np.full(small, mo.V_init, dtype=np.float32)
    Uo[0:3] = 20.0
    mo._dt = dt
    mo.initialize_state_variables(Uo)
    Uo = ofd.fd_run(mo, Uo, nsteps, dt, diff, spacing)
    s = FDStepper((24, 64, 64), spacing, dt, diff, model=_model(tag, math='fast'), math='fast')
    U = s.U()
    U.fill_(float(mo.V_init))
    U[0:3] = 20.0
    s.step(nsteps)
    Ug = s.U()
    assert bool((Ug == Ug[:, :1, :1]).all())
    assert rel_l2(Ug[:, 5, 9].cpu().numpy(), Uo[:, 0, 0]) <= 1e-5
