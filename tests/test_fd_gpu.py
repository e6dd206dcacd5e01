"""GPU parity tests of the fused FD step (through the C ABI) against the NumPy oracle.

Tolerances (north star: fp32, rel-L2 of the transmembrane potential <= 1e-5 after N steps,
activation times within one dt):
  EXACT arithmetic: rel-L2 <= 1e-6 (expected ~0: same ops in the same order)
  FAST  arithmetic: rel-L2 <= 1e-5
"""
import numpy as np
import pytest
import torch

from helpers import plane_wave_ic, random_fields, rel_l2, smooth_fields
from oracle import fd as ofd
from oracle import ionic as oionic
from oracle import stimulus as ostim

pytestmark = pytest.mark.gpu

TOL = {'exact': 1e-6, 'fast': 1e-5}
NSTATES = {'fenton4v': 3, 'mms2v': 1, 'ms2v': 1, None: 0}


def make_oracle(model, dt, states):
    if model is None:
        return None
    m = oionic.MODELS[model]()
    m._dt = dt
    for attr, s in zip(oionic.STATE_ATTRS[model], states):
        setattr(m, attr, s.copy())
    m._initialized = True
    return m


def run_gpu(shape, spacing, dt, diff, model, lap, math, kernel, U, states, nsteps, DIFF=None, use_run=False):
    from pdensorflow_b200.fd import FDStepper
    st = FDStepper(shape, spacing, dt, diff, model=model, lap=lap, DIFF=DIFF, math=math, kernel=kernel)
    st.set_U(U)
    if states:
        st.set_states(states)
    if use_run:
        st.run(nsteps)
    else:
        st.step(nsteps)
    torch.cuda.synchronize()
    return st, st.U().cpu().numpy().reshape(shape), [s.cpu().numpy().reshape(shape) for s in st.states]


def run_oracle(shape, spacing, dt, diff, model, U, states, nsteps, DIFF=None, conv=False):
    m = make_oracle(model, dt, states)
    Uo = ofd.fd_run(m, U.copy(), nsteps, dt, diff, spacing, DIFF=DIFF, conv=conv)
    so = [getattr(m, a) for a in oionic.STATE_ATTRS[model]] if model else []
    return Uo, so


CASES_3D = [
    # shape,          kernel
    ((24, 20, 16), 'tma'),        # single z tile, partial y tiles
    ((10, 19, 132), 'tma'),       # two z tiles (second one partial), odd ny
    ((9, 11, 13), 'generic'),     # nz % 4 != 0 -> generic only
    ((12, 16, 20), 'generic'),
]


@pytest.mark.parametrize('shape,kernel', CASES_3D)
@pytest.mark.parametrize('model', ['fenton4v', 'mms2v', 'ms2v', None])
@pytest.mark.parametrize('math', ['exact', 'fast'])
def test_one_step_random_fields(cuda, shape, kernel, model, math):
    """One step from seeded random fields: exercises every branch of the models and every
    boundary case of the stencil (faces, edges, corners)."""
    U, states = random_fields(shape, seed=3, nstates=NSTATES[model])
    dt, diff, spacing = 0.1, 0.8, (1.0, 0.9, 1.1)
    _, Ug, sg = run_gpu(shape, spacing, dt, diff, model, 'homog', math, kernel, U, states, 1)
    Uo, so = run_oracle(shape, spacing, dt, diff, model, U, states, 1)
    assert rel_l2(Ug, Uo) <= TOL[math]
    for a, b in zip(sg, so):
        assert rel_l2(a, b) <= TOL[math]
    if math == 'exact' and model is None:
        assert np.array_equal(Ug, Uo)          # pure stencil: bit exact


@pytest.mark.parametrize('shape,kernel', [((24, 20, 16), 'tma'), ((9, 11, 13), 'generic')])
@pytest.mark.parametrize('math', ['exact', 'fast'])
def test_exact_mode_is_bitwise_on_most_nodes(cuda, shape, kernel, math):
    U, states = random_fields(shape, seed=5)
    _, Ug, sg = run_gpu(shape, (1.0, 1.0, 1.0), 0.1, 0.8, 'fenton4v', 'homog', math, kernel, U, states, 1)
    Uo, so = run_oracle(shape, (1.0, 1.0, 1.0), 0.1, 0.8, 'fenton4v', U, states, 1)
    frac = np.mean(Ug == Uo)
    if math == 'exact':
        assert frac > 0.999, frac     # only a correctly-rounded-tanh tie can differ
        for a, b in zip(sg[:2], so[:2]):
            assert np.array_equal(a, b)   # V and W involve no transcendental
    assert np.max(np.abs(Ug - Uo)) < 1e-3


@pytest.mark.parametrize('kernel,shape', [('tma', (24, 20, 16)), ('generic', (9, 11, 13))])
@pytest.mark.parametrize('model', ['fenton4v', 'mms2v', None])
@pytest.mark.parametrize('math', ['exact', 'fast'])
def test_heterogeneous_operator(cuda, kernel, shape, model, math):
    U, states = random_fields(shape, seed=7, nstates=NSTATES[model])
    rng = np.random.default_rng(11)
    DIFF = rng.uniform(0.0, 1.0, size=shape).astype(np.float32)
    DIFF[rng.uniform(size=shape) < 0.2] = 0.0                 # holes: DIFF = 0 outside the tissue
    dt, spacing = 0.1, (1.0, 1.0, 1.0)
    _, Ug, sg = run_gpu(shape, spacing, dt, 1.0, model, 'heterog', math, kernel, U, states, 1, DIFF=DIFF)
    Uo, so = run_oracle(shape, spacing, dt, 1.0, model, U, states, 1, DIFF=DIFF)
    assert rel_l2(Ug, Uo) <= TOL[math]
    if math == 'exact' and model is None:
        assert np.array_equal(Ug, Uo)


@pytest.mark.parametrize('math', ['exact', 'fast'])
def test_conv_alias(cuda, math):
    shape = (12, 16, 20)
    U, states = random_fields(shape, seed=9)
    spacing = (0.5, 2.0, 0.5)
    from pdensorflow_b200.fd import FDStepper
    st = FDStepper(shape, spacing, 0.05, 0.8, model='fenton4v', lap='conv', math=math)
    st.set_U(U); st.set_states(states); st.step(1)
    Ug = st.U().cpu().numpy()
    Uo, _ = run_oracle(shape, spacing, 0.05, 0.8, 'fenton4v', U, states, 1, conv=True)
    assert rel_l2(Ug, Uo) <= (1e-6 if math == 'exact' else 1e-5)


@pytest.mark.parametrize('shape,kernel', [((32, 24), 'tma'), ((40, 136), 'tma'), ((17, 23), 'generic')])
@pytest.mark.parametrize('model', ['fenton4v', 'mms2v', None])
@pytest.mark.parametrize('math', ['exact', 'fast'])
def test_2d(cuda, shape, kernel, model, math):
    """Config 2 is the 3-D solve() pattern transcribed to 2-D."""
    U, states = random_fields(shape, seed=13, nstates=NSTATES[model])
    dt, diff, spacing = 0.1, 0.8, (1.0, 1.2)
    _, Ug, sg = run_gpu(shape, spacing, dt, diff, model, 'homog', math, kernel, U, states, 1)
    Uo, so = run_oracle(shape, spacing, dt, diff, model, U, states, 1)
    assert rel_l2(Ug, Uo) <= TOL[math]
    rng = np.random.default_rng(17)
    DIFF = rng.uniform(0.0, 1.0, size=shape).astype(np.float32)
    _, Ug, _ = run_gpu(shape, spacing, dt, 1.0, model, 'heterog', math, kernel, U, states, 1, DIFF=DIFF)
    Uo, _ = run_oracle(shape, spacing, dt, 1.0, model, U, states, 1, DIFF=DIFF)
    assert rel_l2(Ug, Uo) <= TOL[math]


@pytest.mark.parametrize('nsteps', [10, 100])
@pytest.mark.parametrize('kernel,shape', [('tma', (96, 80, 72)), ('generic', (31, 30, 29))])
@pytest.mark.parametrize('math', ['exact', 'fast'])
def test_multi_step_plane_wave(cuda, nsteps, kernel, shape, math):
    """N steps of the examples' S1 plane-wave protocol on a non-cubic grid."""
    U = plane_wave_ic(shape)
    dt, diff, spacing = 0.1, 0.8, (1.0, 1.0, 1.0)
    _, Ug, sg = run_gpu(shape, spacing, dt, diff, 'fenton4v', 'homog', math, kernel, U, [], nsteps)
    Uo, so = run_oracle(shape, spacing, dt, diff, 'fenton4v', U,
                        [np.ones(shape, np.float32), np.ones(shape, np.float32), np.zeros(shape, np.float32)], nsteps)
    assert rel_l2(Ug, Uo) <= TOL[math], (nsteps, rel_l2(Ug, Uo))
    for a, b in zip(sg, so):
        assert rel_l2(a, b) <= 10 * TOL[math]


@pytest.mark.parametrize('math', ['exact', 'fast'])
def test_multi_step_smooth_3d_fields(cuda, math):
    """100 steps from a smooth fully 3-D field (fronts in every direction), graph replay."""
    shape = (48, 40, 36)
    U, states = smooth_fields(shape, seed=21)
    dt, diff, spacing = 0.1, 0.8, (1.0, 1.0, 1.0)
    _, Ug, sg = run_gpu(shape, spacing, dt, diff, 'fenton4v', 'homog', math, 'tma', U, states, 100, use_run=True)
    Uo, so = run_oracle(shape, spacing, dt, diff, 'fenton4v', U, states, 100)
    assert rel_l2(Ug, Uo) <= TOL[math], rel_l2(Ug, Uo)


@pytest.mark.parametrize('math', ['exact', 'fast'])
def test_1000_steps_s1s2_activation_times(cuda, math):
    """1000 steps of the S1-S2 protocol (S2 fired at step 300 on a genuine boolean mask) on 40x36x32:
    rel-L2 and activation times (first upward crossing of -30 mV) within one dt."""
    from pdensorflow_b200.fd import FDStepper
    shape = (40, 36, 32)
    dt, diff, spacing, thr = 0.1, 0.8, (1.0, 1.0, 1.0), -30.0
    U = plane_wave_ic(shape)
    mask = np.zeros(shape, dtype=np.float32)
    mask[:shape[0] // 2, :shape[1] // 2, :] = 1.0
    s2 = ostim.Stimulus({'tstart': 30.0, 'nstim': 1, 'period': 800, 'duration': dt, 'intensity': 60.0})
    s2.set_stimregion(mask)
    m = make_oracle('fenton4v', dt, [np.ones(shape, np.float32), np.ones(shape, np.float32),
                                     np.zeros(shape, np.float32)])
    Uo, act_o = ofd.fd_run(m, U.copy(), 1000, dt, diff, spacing, s2=s2, act_threshold=thr)

    st = FDStepper(shape, spacing, dt, diff, model='fenton4v', math=math)
    st.set_U(U)
    mask_d = torch.from_numpy(mask).cuda()
    act_g = torch.full(shape, -1, dtype=torch.int32, device='cuda')
    prev = st.U().clone()
    for i in range(1000):
        st.step(1)
        if i == 300:
            st.apply_stimulus(mask_d, 60.0)
        cur = st.U()
        crossed = (prev < thr) & (cur >= thr) & (act_g < 0)
        act_g[crossed] = i
        prev = cur.clone()
    Ug = st.U().cpu().numpy()
    act_g = act_g.cpu().numpy()
    assert rel_l2(Ug, Uo) <= TOL[math], rel_l2(Ug, Uo)
    assert np.array_equal(act_g >= 0, act_o >= 0)
    assert np.max(np.abs(act_g - act_o)) <= 1


def test_shipped_quirk_all_ones_mask(cuda):
    """F6: as shipped the S2 mask is all ones, so U = max(U, 60) on the whole grid (32^3, the
    shipped size), 40 steps around the stimulus."""
    from pdensorflow_b200.fd import FDStepper
    shape = (32, 32, 32)
    dt, diff, spacing = 0.1, 0.8, (1.0, 1.0, 1.0)
    U = plane_wave_ic(shape)
    s2_init = np.full(shape, -80.0, dtype=np.float32)
    s2_init[:16, :16, :] = 20.0
    s2 = ostim.Stimulus({'tstart': 2.0, 'nstim': 1, 'period': 800, 'duration': dt, 'intensity': 60.0})
    s2.set_stimregion(s2_init)
    m = make_oracle('fenton4v', dt, [np.ones(shape, np.float32), np.ones(shape, np.float32),
                                     np.zeros(shape, np.float32)])
    Uo = ofd.fd_run(m, U.copy(), 40, dt, diff, spacing, s2=s2)
    st = FDStepper(shape, spacing, dt, diff, model='fenton4v', math='exact')
    st.set_U(U)
    mask_d = torch.from_numpy(s2.get_stimregion()).cuda()
    for i in range(40):
        st.step(1)
        if i == 20:
            st.apply_stimulus(mask_d, 60.0)
    assert rel_l2(st.U().cpu().numpy(), Uo) <= 1e-6
    assert Uo.min() > -81.0


def test_tma_and_generic_kernels_agree(cuda):
    shape = (20, 24, 260)          # three z tiles
    U, states = random_fields(shape, seed=23)
    out = {}
    for kernel in ('tma', 'generic'):
        _, Ug, sg = run_gpu(shape, (1.0, 1.0, 1.0), 0.1, 0.8, 'fenton4v', 'homog', 'exact', kernel, U, states, 3)
        out[kernel] = (Ug, sg)
    assert np.array_equal(out['tma'][0], out['generic'][0])
    for a, b in zip(out['tma'][1], out['generic'][1]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize('shape', [(12, 19, 132), (20, 24, 260), (33, 260)])
@pytest.mark.parametrize('lap', ['homog', 'heterog'])
@pytest.mark.parametrize('model', [None, 'fenton4v'])
def test_dirichlet_pad_on_the_tma_kernel(cuda, shape, lap, model):
    """LaplaceSolver's constant pad (laplaceSolver.py:44-52) on the TMA tile kernel: bitwise the generic kernel in EXACT
    arithmetic, and the oracle's fd_step(dirichlet=...) within the EXACT tolerance."""
    from pdensorflow_b200.fd import FDStepper
    rng = np.random.default_rng(31)
    U, states = random_fields(shape, seed=37)
    states = states[:NSTATES[model]]
    DIFF = (0.2 + rng.random(shape)).astype(np.float32) if lap == 'heterog' else None
    spacing = (1.0, 0.9, 1.1)[:len(shape)]
    res = {}
    for kernel in ('tma', 'generic'):
        st = FDStepper(shape, spacing, 0.05, 0.8, model=model, lap=lap, DIFF=DIFF, math='exact', kernel=kernel, dirichlet=3.5)
        st.set_U(U)
        if states:
            st.set_states(states)
        st.step(3)
        res[kernel] = [st.U().cpu().numpy()] + [x.cpu().numpy() for x in st.states]
    for a, b in zip(res['tma'], res['generic']):
        assert np.array_equal(a, b)
    m = make_oracle(model, 0.05, states)
    Uo = U.copy()
    for _ in range(3):
        Uo = ofd.fd_step(m, Uo, 0.05, 0.8, spacing, DIFF=DIFF, dirichlet=3.5)
    assert rel_l2(res['tma'][0].reshape(shape), Uo) <= TOL['exact']


def _aniso_fields(shape, seed):
    rng = np.random.default_rng(seed)
    DF = np.empty(shape + (2,), np.float32)
    DF[..., 0] = 0.1 + 0.2 * rng.random(shape)
    DF[..., 1] = 0.5 * rng.random(shape)
    a = rng.standard_normal(shape + (3,))
    a /= np.linalg.norm(a, axis=-1, keepdims=True)
    AV = np.stack([a[..., 0] * a[..., 0], a[..., 0] * a[..., 1], a[..., 0] * a[..., 2], a[..., 1] * a[..., 1],
                   a[..., 1] * a[..., 2], a[..., 2] * a[..., 2]], axis=-1).astype(np.float32)
    return DF, AV


@pytest.mark.parametrize('shape,chunk', [((9, 11, 16), 0), ((21, 19, 132), 4), ((12, 24, 260), 5), ((3, 3, 4), 0)])
@pytest.mark.parametrize('model', [None, 'fenton4v', 'mms2v'])
def test_anisotropic_tile_kernel(cuda, shape, chunk, model):
    """19-point fibre-tensor operator on TMA tiles (pdx_fd_aniso.cu): EXACT arithmetic is bitwise the one-thread-per-node
    kernel (and the oracle within the EXACT tolerance) on partial tiles, several z tiles and x chunks with halo planes;
    FAST stays within 1e-5."""
    from pdensorflow_b200.fd import FDStepper
    DF, AV = _aniso_fields(shape, 41)
    U, states = random_fields(shape, seed=43)
    states = states[:NSTATES[model]]
    spacing, dt, nsteps = (1.0, 0.8, 1.25), 0.02, 3
    res = {}
    for tag, kernel, math in (('tile', 'auto', 'exact'), ('generic', 'generic', 'exact'), ('fast', 'auto', 'fast')):
        st = FDStepper(shape, spacing, dt, 1.0, model=model, lap='aniso', DIFF=DF, AVEC=AV, math=math, kernel=kernel,
                       x_chunk=chunk)
        from pdensorflow_b200 import _lib as L
        assert bool(L.lib().pdx_fd_plan_aniso_tile_capable(st._plan)) == (kernel != 'generic')
        st.set_U(U)
        if states:
            st.set_states(states)
        st.step(nsteps)
        res[tag] = [st.U().cpu().numpy().reshape(shape)] + [x.cpu().numpy().reshape(shape) for x in st.states]
    for a, b in zip(res['tile'], res['generic']):
        assert np.array_equal(a, b)
    m = make_oracle(model, dt, states)
    Uo = U.copy()
    for _ in range(nsteps):
        U0 = ofd.enforce_boundary(Uo)
        lap = ofd.laplace_heterog_aniso_3d(U0, DF, AV, *spacing)
        dU = m.differentiate(Uo) if m is not None else np.float32(0.0)
        Uo = (U0 + np.float32(dt) * dU + np.float32(dt) * lap).astype(np.float32)
    assert rel_l2(res['tile'][0], Uo) <= TOL['exact']
    assert rel_l2(res['fast'][0], Uo) <= TOL['fast']


ROW_SHAPES = [(9, 11, 13), (12, 20, 133), (7, 9, 262), (5, 6, 5), (4, 3, 3), (19, 131), (40, 13), (9, 258)]


@pytest.mark.parametrize('shape', ROW_SHAPES)
@pytest.mark.parametrize('lap', ['homog', 'heterog'])
@pytest.mark.parametrize('model', [None, 'fenton4v', 'mms2v'])
def test_rows_that_are_not_16_byte_multiples_on_the_tile_kernel(cuda, shape, lap, model):
    """nz % 4 != 0: the stepper keeps its arrays with rows padded to a multiple of four floats (pdx_fd_desc.pitch_z) and the
    tile kernel runs on them - bitwise the one-thread-per-node kernel on contiguous rows in EXACT arithmetic, x chunks,
    the Dirichlet pad and a stimulus included; FAST within 1e-5 of the oracle."""
    from pdensorflow_b200.fd import FDStepper
    rng = np.random.default_rng(53)
    U, states = random_fields(shape, seed=59)
    states = states[:NSTATES[model]]
    DIFF = (0.2 + rng.random(shape)).astype(np.float32) if lap == 'heterog' else None
    spacing = (1.0, 0.9, 1.1)[:len(shape)]
    for dirichlet in (None, 2.5):
        res = {}
        for tag, kernel, math, chunk in (('tile', 'tma', 'exact', 0), ('chunk3', 'tma', 'exact', 3), ('generic', 'generic', 'exact', 0),
                                         ('fast', 'auto', 'fast', 0)):
            st = FDStepper(shape, spacing, 0.05, 0.8, model=model, lap=lap, DIFF=DIFF, math=math, kernel=kernel,
                           dirichlet=dirichlet, x_chunk=chunk)
            assert st.kernel == ('generic' if kernel == 'generic' else 'tma')
            assert (st.pitch != shape[-1]) == (kernel != 'generic') and st.U().shape == tuple(shape)
            st.set_U(U)
            if states:
                st.set_states(states)
            st.step(2)
            mask = torch.zeros(shape, device='cuda')
            mask[..., : shape[-1] // 2] = 1.0
            st.apply_stimulus(mask, 30.0)
            st.step(1)
            res[tag] = [st.U().cpu().numpy()] + [x.cpu().numpy() for x in st.states]
        for tag in ('tile', 'chunk3'):
            for a, b in zip(res[tag], res['generic']):
                assert np.array_equal(a, b), (tag, dirichlet)
        m = make_oracle(model, 0.05, states)
        Uo = U.copy()
        for k in range(3):
            Uo = ofd.fd_step(m, Uo, 0.05, 0.8, spacing, DIFF=DIFF, dirichlet=dirichlet)
            if k == 1:
                mk = np.zeros(shape, np.float32)
                mk[..., : shape[-1] // 2] = 1.0
                Uo = np.where(mk != 0, np.maximum(Uo, np.float32(30.0) * mk), Uo).astype(np.float32)
        assert rel_l2(res['fast'][0].reshape(shape), Uo) <= TOL['fast']


def test_x_chunking_is_invisible(cuda):
    """The TMA kernel marches x in chunks with 2 halo planes: results must not depend on the chunk."""
    from pdensorflow_b200.fd import FDStepper
    shape = (37, 16, 16)
    U, states = random_fields(shape, seed=29)
    res = []
    for chunk in (0, 1, 2, 5, 37):
        st = FDStepper(shape, (1.0, 1.0, 1.0), 0.1, 0.8, model='fenton4v', math='exact', kernel='tma', x_chunk=chunk)
        st.set_U(U); st.set_states(states); st.step(2)
        res.append(st.U().cpu().numpy())
    for r in res[1:]:
        assert np.array_equal(r, res[0])


def test_full_size_plane_wave_property(cuda):
    """Config-3 size (512^3): a plane wave along x is independent of (y,z), so every column of the
    512^3 GPU run must equal the oracle run on a 512x4x4 grid (size-independent property)."""
    from pdensorflow_b200.fd import FDStepper
    n = 512
    free, _ = torch.cuda.mem_get_info()
    if free < 6 * n ** 3 * 4:
        pytest.skip('not enough device memory for 512^3')
    nsteps = 60
    small = (n, 4, 4)
    Uo, _ = run_oracle(small, (1.0, 1.0, 1.0), 0.1, 0.8, 'fenton4v', plane_wave_ic(small),
                       [np.ones(small, np.float32), np.ones(small, np.float32), np.zeros(small, np.float32)], nsteps)
    for math in ('exact', 'fast'):
        st = FDStepper((n, n, n), (1.0, 1.0, 1.0), 0.1, 0.8, model='fenton4v', math=math)
        U = st.U()
        U.fill_(-80.0)
        U[0:2] = 20.0
        st.run(nsteps)
        Ug = st.U()
        assert bool((Ug == Ug[:, :1, :1]).all())          # still planar
        col = Ug[:, n // 3, n // 5].cpu().numpy()
        assert rel_l2(col, Uo[:, 0, 0]) <= TOL[math]
        # a checksum over the whole field agrees with the column (planarity + parity)
        tot = float(Ug.double().sum())
        assert abs(tot - float(col.astype(np.float64).sum()) * n * n) <= 1e-9 * abs(tot) + 1e-3
        del st, U, Ug
        torch.cuda.empty_cache()


@pytest.mark.parametrize('kernel', ['tma', 'generic'])
def test_index_arithmetic_beyond_2_31_nodes(cuda, kernel):
    """Maximum sizes: a slab of 520 x 2048 x 2048 = 2.18e9 nodes (config 4 keeps 128-512 planes of 2048^2 per GPU)
    has element offsets beyond 2^31.  The field repeats every 8 planes along x, so - two planes away from the slab's
    ends - the solution does too: planes 512.. (offsets >= 2^31) must equal planes 16.. BITWISE, and both must equal
    the same 8-plane pattern advanced on a small 48-plane grid."""
    from pdensorflow_b200.fd import FDStepper
    nx, ny, nz, period, nsteps = 520, 2048, 2048, 8, 2
    free, _ = torch.cuda.mem_get_info()
    if free < 5.5 * nx * ny * nz * 4:
        pytest.skip('not enough device memory for 2.18e9 nodes')
    g = torch.Generator(device='cuda').manual_seed(3)
    block = torch.rand((period, ny, nz), generator=g, device='cuda') * 100.0 - 80.0

    def run(planes):
        st = FDStepper((planes, ny, nz), (1.0, 1.0, 1.0), 0.1, 0.8, model='fenton4v', math='exact', kernel=kernel)
        st.U().view(planes // period, period, ny, nz).copy_(block.unsqueeze(0).expand(planes // period, -1, -1, -1))
        st.step(nsteps)
        torch.cuda.synchronize()
        return st

    big = run(nx)
    assert big.U().numel() > 2 ** 31
    hi, lo = big.U()[512:516], big.U()[16:20]
    assert bool(torch.isfinite(hi).all()) and torch.equal(hi, lo)
    assert torch.equal(big.states[0][512:516], big.states[0][16:20])
    keep = lo.clone()
    del big, hi, lo
    torch.cuda.empty_cache()
    small = run(48)
    assert torch.equal(small.U()[16:20], keep)


# ------------------------------------------------------------------------------ round 2: rasterisation, streamer
@pytest.mark.parametrize('lap', ['homog', 'heterog'])
def test_cta_rasterisation_is_invisible(cuda, lap, monkeypatch):
    """PDX_FD_RASTER=G (tile groups of G, x-chunks next, odd chunks marched downwards) must give the bits of the
    plain layer-by-layer order for every group size and chunk length, incl. short last groups, ragged last chunks and
    the mirrored x neighbours of the heterogeneous operator."""
    from pdensorflow_b200.fd import FDStepper
    shape = (45, 24, 132)
    U, states = random_fields(shape, seed=31)
    DIFFa = np.random.default_rng(5).uniform(0.2, 1.0, size=shape).astype(np.float32) if lap == 'heterog' else None
    res = {}
    for raster in ('0', '1', '4', '16'):          # 3 y-tiles x 2 z-tiles = 6 tiles: groups of 1, 4 + 2, and one short group
        monkeypatch.setenv('PDX_FD_RASTER', raster)
        for chunk in (0, 3, 8):
            st = FDStepper(shape, (1.0, 0.9, 1.1), 0.05, 0.8, model='fenton4v', math='exact', kernel='tma', x_chunk=chunk,
                           lap=lap, DIFF=DIFFa)
            st.set_U(U); st.set_states(states); st.step(3)
            res[(raster, chunk)] = (st.U().cpu().numpy(), [s.cpu().numpy() for s in st.states])
    ref = res[('0', 0)]
    for key, (Ug, sg) in res.items():
        assert np.array_equal(Ug, ref[0]), key
        for a, b in zip(sg, ref[1]):
            assert np.array_equal(a, b), key


@pytest.mark.parametrize('steps,nchunks', [(2, 3), (6, 4), (10, 2), (10, 8)])
def test_frame_streamer_equals_whole_grid_steps(cuda, steps, nchunks):
    """fd.FrameStreamer (chunked H2D -> time-skewed range launches -> chunked D2H, single simulation living in
    pinned host memory) must reproduce, bitwise, whole-grid launches of the same number of steps, frame after frame."""
    from pdensorflow_b200.fd import FDStepper, FrameStreamer
    shape = (70, 24, 128)
    U, states = random_fields(shape, seed=37)
    ref = FDStepper(shape, (1.0, 1.0, 1.0), 0.1, 0.8, model='fenton4v', math='exact')
    ref.set_U(U); ref.set_states(states)
    st = FDStepper(shape, (1.0, 1.0, 1.0), 0.1, 0.8, model='fenton4v', math='exact')
    st.set_U(U); st.set_states(states)
    fs = FrameStreamer(st, steps, nchunks=nchunks)
    assert fs.nchunks == min(nchunks, shape[0] // (steps + 2))
    for frame in range(3):
        fs.frame()
        ref.step(steps)
        if frame == 1:
            fs.wait()              # draining between frames must not matter
    fs.wait()
    torch.cuda.synchronize()
    assert torch.equal(st.U(), ref.U())
    assert np.array_equal(fs.host.numpy(), ref.U().cpu().numpy())
    for a, b in zip(st.states, ref.states):
        assert torch.equal(a, b)


def test_frame_streamer_unskewed_slab_mode(cuda):
    """skew=False (the multi-GPU frame loop: whole-slab steps between the chunked copies) on a stand-alone grid."""
    from pdensorflow_b200.fd import FDStepper, FrameStreamer
    shape = (33, 16, 128)
    U, states = random_fields(shape, seed=41)
    ref = FDStepper(shape, (1.0, 1.0, 1.0), 0.1, 0.8, model='fenton4v', math='exact')
    ref.set_U(U); ref.set_states(states)
    st = FDStepper(shape, (1.0, 1.0, 1.0), 0.1, 0.8, model='fenton4v', math='exact')
    st.set_U(U); st.set_states(states)
    fs = FrameStreamer(st, 4, nchunks=5, skew=False)
    for _ in range(3):
        fs.frame()
        ref.step(4)
    fs.wait()
    assert torch.equal(st.U(), ref.U()) and np.array_equal(fs.host.numpy(), ref.U().cpu().numpy())


def test_raw_pointer_arguments_are_checked(cuda):
    """masks reach the kernel as raw pointers - wrong dtype / device / shape must raise, not misread."""
    from pdensorflow_b200 import _lib as L
    from pdensorflow_b200.fd import FDStepper
    shape = (8, 8, 16)
    st = FDStepper(shape, (1.0, 1.0, 1.0), 0.1, 0.8)
    st.set_U(plane_wave_ic(shape))
    good = torch.ones(shape, device='cuda')
    st.apply_stimulus(good, 60.0)
    for bad in (torch.ones(shape, device='cuda', dtype=torch.bool), torch.ones(shape, device='cuda', dtype=torch.float64),
                torch.ones(shape), torch.ones((8, 8, 8), device='cuda'), torch.ones((16, 8, 8), device='cuda').permute(2, 1, 0)):
        with pytest.raises(L.PdxError):
            st.apply_stimulus(bad, 60.0)
    assert float(st.U().min()) == 60.0


@pytest.mark.parametrize('model', ['fenton4v', 'mms2v', None])
@pytest.mark.parametrize('shape', [(70, 132), (8, 256), (200, 128), (33, 520)])
def test_2d_y_march_is_invisible(cuda, model, shape):
    """2-D grids on the TMA kernel: a CTA marches over x_chunk y-tiles of one z-strip (states staged by TMA).  Every
    chunk length, incl. ragged last chunks and tiles that touch the y boundary mid-march, must give the bits of the
    generic kernel; the heterogeneous operator too."""
    from pdensorflow_b200.fd import FDStepper
    U, states = random_fields(shape, seed=43, nstates=NSTATES[model])
    D2 = np.random.default_rng(3).uniform(0.2, 1.0, size=shape).astype(np.float32)
    for lap, DIFFa in (('homog', None), ('heterog', D2)):
        ref = FDStepper(shape, (1.0, 0.9), 0.05, 0.8, model=model, math='exact', kernel='generic', lap=lap, DIFF=DIFFa)
        ref.set_U(U); ref.set_states(states); ref.step(3)
        for chunk in (0, 1, 2, 3, 8):
            st = FDStepper(shape, (1.0, 0.9), 0.05, 0.8, model=model, math='exact', kernel='tma', x_chunk=chunk, lap=lap,
                           DIFF=DIFFa)
            st.set_U(U); st.set_states(states); st.step(3)
            assert torch.equal(st.U(), ref.U()), (lap, chunk)
            for a, b in zip(st.states, ref.states):
                assert torch.equal(a, b), (lap, chunk)
