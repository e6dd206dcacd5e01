"""N4: PNG-mosaic ingestion into Domain3D and the anatomical example Tests/FD/Fenton_atria/fenton.py,
against fixtures frozen from the reference's own reader + example run from source (tests/golden/make_atria_golden.py:
ref_atria.npz; atria_mosaic.png is the reference anatomy re-encoded by our PNG writer)."""
import os

import numpy as np
import pytest

from oracle import fd as ofd
from oracle import ionic as oionic
from oracle import stimulus as ostim

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
PNG = os.path.join(GOLD, 'atria_mosaic.png')


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(GOLD, 'ref_atria.npz'))


def atria_setup(g, domain_cls, fname=PNG):
    """The example's constructor + the head of run() (fenton.py:62-103,120-153) on a Domain3D class."""
    Mx, My, dt, per_plot, diff, samples, s2_time = (float(v) for v in g['cfg'])
    dom = domain_cls({'width': 64, 'height': 64, 'depth': 64, 'dx': 1, 'dy': 1, 'dz': 1})
    dom.load_geometry_file(fname, int(Mx), int(My), 1.e-4)
    dom.load_conductivity(fname='', cond_unif=diff)
    W, H, D = dom.width(), dom.height(), dom.depth()
    mask = dom.geometry() > 0.0
    U = np.full((W, H, D), -80.0, dtype=np.float32)
    U[:, :, (D // 2 - 10):(D // 2 + 10)] = 20.0
    s2_init = dom.geometry().astype(np.float32)
    s2_init[:, :H // 2, :] = 0.0
    s2_init = s2_init * np.float32(100.0) + np.float32(-80.0)
    U = np.where(mask, U, np.float32(-80.0)).astype(np.float32)
    return dom, mask, U, s2_init, dt, int(per_plot), int(samples), s2_time


def test_png_decoder_and_mosaic_unfolding(gold):
    from pdensorflow_b200.gpuSolve.IO.readers.imagedata import ImageData, parse_name
    img = ImageData()
    img.load_image(PNG, 16, 8)
    vol = img.get_data()
    assert vol.shape == tuple(gold['volume_shape']) and vol.dtype == np.uint8
    assert np.array_equal(np.packbits(vol > 0), gold['volume_bits']) and int(vol.astype(np.int64).sum()) == int(gold['volume_sum'])
    unit = img.get_rescaled_data('unit')
    assert unit.dtype == np.float64 and unit.min() == 0.0 and unit.max() == 1.0
    assert parse_name('a/b/structure.png')['type'] == 'png' and parse_name('x.nii.gz') == {
        'gzipped': True, 'type': 'nii', 'name': 'x', 'fname': 'x.nii.gz'}


def test_nifti_and_npy_round_trip(tmp_path):
    from pdensorflow_b200.gpuSolve.IO.readers.imagedata import ImageData
    rng = np.random.default_rng(2)
    vol = rng.integers(0, 4000, size=(5, 4, 3)).astype(np.int16)
    np.save(tmp_path / 'v.npy', vol)
    a = ImageData()
    a.load_image(str(tmp_path / 'v.npy'))
    for name in ('v.nii', 'v.nii.gz'):
        a.save_nifty(str(tmp_path / name))
        b = ImageData()
        b.load_image(str(tmp_path / name))
        assert np.array_equal(b.get_data(), vol) and b.get_fdata().dtype == np.float64
        assert np.array_equal(b.get_rescaled_data('mstd'), (vol - vol.mean()) / vol.std())


def test_domain3d_geometry_from_the_png(gold):
    from pdensorflow_b200.gpuSolve.entities import Domain3D
    dom, mask, U, s2_init, *_ = atria_setup(gold, Domain3D)
    assert (dom.width(), dom.height(), dom.depth()) == (128, 128, 128)
    assert np.array_equal(np.packbits(dom.geometry() > 0), gold['geometry_bits'])
    assert set(np.unique(dom.geometry())) == {0.0, 1.0} and np.array_equal(dom.conductivity(), 0.75 * dom.geometry())
    assert np.array_equal(U, gold['frames'][0])


def test_atria_example_first_frames_bit_exact_on_the_oracle(gold):
    """run() of the example (fenton.py:106-112,176-186) restated on the oracle: frames i = 0 and 40 (the CPU suite stops
    there; the GPU test replays all 81 steps incl. the S2 stimulus)."""
    from pdensorflow_b200.gpuSolve.entities import Domain3D
    dom, mask, U, s2_init, dt, per_plot, samples, s2_time = atria_setup(gold, Domain3D)
    DIFF = dom.conductivity().astype(np.float32)
    m = oionic.Fenton4v()
    m._dt = dt
    m.initialize_state_variables(U)
    s2 = ostim.Stimulus({'tstart': s2_time, 'nstim': 1, 'period': 800, 'duration': dt, 'intensity': 60.0})
    s2.set_stimregion(s2_init)
    assert s2.get_stimregion().min() == 1.0            # the mask is all ones (-80 is "true"): the shipped quirk
    spacing = (dom.dx(), dom.dy(), dom.dz())
    k = 1
    for i in range(41):
        U = ofd.fd_step(m, U, dt, 1.0, spacing, DIFF=DIFF)
        if s2.stimulate_tissue_timestep(i, dt):
            U = ofd.apply_stimulus(U, s2())
        if i % per_plot == 0:
            assert np.array_equal(np.where(mask, U, np.float32(-1.0)).astype(np.float32), gold['frames'][k]), i
            k += 1
    assert k == 3
