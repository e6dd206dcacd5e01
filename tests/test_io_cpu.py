"""Result writers on the stepping loop against fixtures written by the REFERENCE'S OWN
writers (tests/golden/make_io_golden.py -> ref_io.npz): IGB files byte for byte, ResultWriter arrays, chunk
counts; plus the reference's round-trip unit test (Tests/CICD/unit/test_igb_roundtrip.py) transcribed."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, 'golden', 'ref_io.npz'))


@pytest.fixture()
def in_tmp(tmp_path):
    cwd = os.getcwd()
    os.chdir(tmp_path)
    yield tmp_path
    os.chdir(cwd)


def test_igb_files_are_byte_identical_to_the_reference(in_tmp):
    from pdensorflow_b200.gpuSolve.IO.writers import IGBWriter
    fr = G['igb_a_frames']
    w = IGBWriter({'fname': 'a.igb', 'Tend': 600.0, 'nt': fr.shape[0], 'nx': fr.shape[1], 'every_N': 3})
    for f in fr:
        w.imshow(f)
    w.wait()
    assert open('a.igb', 'rb').read() == G['igb_a_bytes'].tobytes()
    fr = G['igb_b_frames']
    w = IGBWriter({'fname': 'sub/b', 'nt': fr.shape[0], 'nx': 12, 'ny': 5, 'nz': 3, 'org_t': 2.5, 'dim_t': 30.0,
                   'units_x': 'mm', 'units_y': 'mm', 'units_z': 'mm', 'units_t': 'ms', 'units': 'mV', 'max_chunk_mb': 0.001})
    for f in fr:
        w.imshow(f)
    w.wait()
    assert open('sub/b.igb', 'rb').read() == G['igb_b_bytes'].tobytes()          # no-extension name gets .igb, directory is created
    assert w.nb_chunks() == int(G['igb_b_nchunks']) and w.counter() == fr.shape[0]


def test_result_writer_dense_and_sparse_match_the_reference(in_tmp):
    from pdensorflow_b200.gpuSolve.IO.writers import ResultWriter
    fr = G['rw_frames']
    W, H, D = fr.shape[1:]
    rw = ResultWriter({'width': W, 'height': H, 'depth': D, 'prefix_name': 'cube3D', 'every_N': 2})
    for f in fr:
        rw.imshow(f)
    rw.wait()
    name = str(G['rw_dense_name'])
    got = np.load(name)
    assert got.dtype == G['rw_dense'].dtype and np.array_equal(got, G['rw_dense'])
    assert rw.nb_chunks() == int(G['rw_dense_nchunks'])
    rw = ResultWriter({'width': W, 'height': H, 'depth': D, 'prefix_name': 'sp'})
    rw.set_sparse_domain(G['rw_domain'])
    for f in fr:
        rw.imshow(f)
    rw.wait()
    d = np.load(str(G['rw_sparse_name']), allow_pickle=True).item()
    assert tuple(d['dimensions']) == tuple(G['rw_sparse_dimensions'])
    assert np.array_equal(np.stack(d['indices']), G['rw_sparse_indices'])
    assert np.array_equal(d['values'], G['rw_sparse_values'])
    assert not [f for f in os.listdir('.') if '__chunk_' in f]                    # nothing temporary left behind
    # legacy pre-allocated cube (resultwriter.py:41-47,66-72)
    rw = ResultWriter({'width': W, 'height': H, 'depth': D, 'prefix_name': 'legacy', 'samples': 4, 'dt_per_plot': 1})
    rw.initialise_cube()
    for f in fr:
        rw.imshow(f)
    rw.wait()
    assert np.array_equal(np.load('legacy_{}_{}_{}.npy'.format(H, W, D)), fr)


def test_base_writer_chunk_triggers_and_npy_aggregation(in_tmp):
    from pdensorflow_b200.gpuSolve.IO.writers import BaseWriter
    fr = G['rw_frames']
    bw = BaseWriter({'fname': 'out/base', 'max_chunk_mb': 0.001})        # 480-byte frames, 1048-byte chunks: 3 + 2
    for f in fr:
        bw.add_solution(f)
    assert bw.nb_chunks() == 1 and bw.counter() == 5
    bw.finalize()
    assert bw.nb_chunks() == 2 and np.array_equal(np.load('out/base.npy'), fr)
    assert os.listdir('out') == ['base.npy']
    empty = BaseWriter({'fname': 'none'})
    empty.finalize()
    assert np.load('none.npy').shape == (0,)
    with pytest.raises(ValueError):
        bw2 = BaseWriter({'fname': 'x'})
        bw2.add_solution(fr[0])
        bw2.add_solution(fr[0][:2])


def _reference_data(nx, nt):
    node = np.arange(nx, dtype=np.float32)
    frame = np.arange(nt, dtype=np.float32)
    return (np.sin(node)[np.newaxis, :] + 100.0 * frame[:, np.newaxis]).astype(np.float32)


@pytest.mark.parametrize('nx,nt', [(300, 5), (63001, 3)], ids=['small', 'large_nx'])
def test_igb_write_read_roundtrip(tmp_path, nx, nt):
    """Tests/CICD/unit/test_igb_roundtrip.py:58-84."""
    from pdensorflow_b200.gpuSolve.IO.readers import IGBReader
    from pdensorflow_b200.gpuSolve.IO.writers import IGBWriter
    fname = str(tmp_path / 'roundtrip.igb')
    ref = _reference_data(nx, nt)
    writer = IGBWriter({'fname': fname, 'Tend': float(nt - 1), 'nt': nt, 'nx': nx})
    for t in range(nt):
        writer.imshow(ref[t, :])
    writer.wait()
    reader = IGBReader()
    reader.read(fname)
    assert reader.nx() == nx and reader.nt() == nt and reader.ndiff() == 0
    assert reader.data().shape == (nt, nx)
    np.testing.assert_allclose(reader.data(), ref, rtol=1.0e-6, atol=1.0e-6)
    assert reader.dim_t() == float(nt - 1) and reader.org_t() == 0.0
    assert np.allclose(reader.timevalues(), np.arange(nt))
    # copy through initialise_from_IGBReader + write_data_to_file (igbwriter.py:190-213)
    copy = IGBWriter({'fname': str(tmp_path / 'copy.igb')})
    copy.initialise_from_IGBReader(reader)
    copy.write_data_to_file()
    assert open(str(tmp_path / 'copy.igb'), 'rb').read() == open(fname, 'rb').read()


def test_reader_tolerates_truncated_and_padded_files(tmp_path):
    from pdensorflow_b200.gpuSolve.IO.readers import IGBReader
    from pdensorflow_b200.gpuSolve.IO.writers.igbwriter import igb_header_bytes
    nx, nt = 10, 4
    data = _reference_data(nx, nt)
    short, extra = str(tmp_path / 'short.igb'), str(tmp_path / 'extra.igb')
    with open(short, 'wb') as f:
        f.write(igb_header_bytes(nx, 1, 1, nt, 0.0, nt - 1))
        data.reshape(-1)[:nx * 2 + 3].tofile(f)
    with open(extra, 'wb') as f:
        f.write(igb_header_bytes(nx, 1, 1, nt, 0.0, nt - 1))
        np.concatenate([data.reshape(-1), np.ones(7, np.float32)]).tofile(f)
    r = IGBReader()
    r.read(short)
    assert r.nt() == 2 and np.array_equal(r.data(), data[:2]) and r.ndiff() == nx * 2 + 3 - nx * nt
    r.read(extra)
    assert r.nt() == nt and np.array_equal(r.data(), data) and r.ndiff() == 7


def test_domain3d_matches_the_reference():
    """Domain3D (entities/domain3D.py) on the paths the FD examples use: file-less box,
    assign_geometry / assign_conductivity (isotropic, trailing singleton channel, anisotropic (sigma_t, sigma_l) ->
    (sigma_t, sigma_l - sigma_t)) - against the reference's own class run from source (ref_io.npz)."""
    from pdensorflow_b200.gpuSolve.entities import Domain3D
    d = Domain3D({'width': 3, 'height': 4, 'depth': 5, 'dx': 0.5, 'dy': 0.25, 'dz': 2.0})
    d.load_geometry_file()
    d.load_conductivity(cond_unif=0.8)
    assert np.array_equal(d.geometry(), G['dom_box_geometry']) and np.array_equal(d.conductivity(), G['dom_box_conductivity'])
    assert np.array_equal([d.width(), d.height(), d.depth(), d.dx(), d.dy(), d.dz(), float(d.anisotropic())], G['dom_box_meta'])
    geo = G['dom_geo']
    d = Domain3D({'dx': 1.0})
    with pytest.raises(SystemExit):
        d.assign_conductivity(geo)                         # no geometry yet (domain3D.py:148-149)
    d.assign_geometry(geo.copy())
    d.assign_conductivity(0.8 * geo)
    assert np.array_equal(d.conductivity(), G['dom_iso_conductivity'])
    assert np.array_equal([d.width(), d.height(), d.depth(), float(d.anisotropic())], G['dom_iso_meta'])
    d.assign_conductivity(G['dom_cond4_in'].copy())
    assert np.array_equal(d.conductivity(), G['dom_cond4']) and not d.anisotropic()
    d.assign_conductivity(G['dom_aniso_in'].copy())
    assert np.array_equal(d.conductivity(), G['dom_aniso']) and d.anisotropic() == bool(G['dom_aniso_flag'])


def test_domain3d_voxel_files(tmp_path):
    from pdensorflow_b200.gpuSolve.entities import Domain3D
    rng = np.random.default_rng(5)
    vox = rng.uniform(2.0, 9.0, size=(4, 3, 5))
    vox[0, 0, 0] = 2.0
    fib = rng.standard_normal((4, 3, 5, 3))
    np.save(tmp_path / 'geo.npy', vox)
    np.save(tmp_path / 'fib.npy', fib)
    d = Domain3D()
    d.load_geometry_file(str(tmp_path / 'geo.npy'), threshold=0.5)
    unit = (vox - vox.min()) / (vox.max() - vox.min())
    assert np.array_equal(d.geometry(), (unit > 0.5).astype(float)) and (d.width(), d.height(), d.depth()) == (4, 3, 5)
    d.load_fiber_direction(str(tmp_path / 'fib.npy'))
    T = d.fibtensor()
    assert T.shape == (4, 3, 5, 6)
    assert np.array_equal(T[..., 1], fib[..., 0] * fib[..., 1]) and np.array_equal(T[..., 5], fib[..., 2] ** 2)
    # PNG mosaic (Tests/FD/Fenton_atria/fenton.py:88-89): 6 slices of 4 x 3 voxels on a 3 x 2 grid
    from pdensorflow_b200.gpuSolve.IO.readers.imagedata import volume_to_mosaic, write_png_gray
    vol = rng.integers(0, 255, size=(4, 3, 6)).astype(np.uint8)
    vol[0, 0, 0], vol[1, 1, 1] = 0, 255
    write_png_gray(str(tmp_path / 'atria.png'), volume_to_mosaic(vol, 3, 2))
    dp = Domain3D()
    dp.load_geometry_file(str(tmp_path / 'atria.png'), 3, 2, threshold=0.25)
    assert (dp.width(), dp.height(), dp.depth()) == (4, 3, 6)
    assert np.array_equal(dp.geometry(), (vol / 255.0 > 0.25).astype(float))
    with pytest.raises(ValueError):
        Domain3D().load_geometry_file(str(tmp_path / 'atria.png'), 5, 2)            # not a 5 x 2 mosaic


def test_reference_basewriter_unit_tests(tmp_path):
    """Tests/CICD/unit/test_basewriter.py:49-148 transcribed (NumPy frames instead of TF tensors): a chunk is flushed
    exactly on multiples of every_N / as soon as max_chunk_mb is exceeded, un-flushed frames are held in `_buffer`,
    every flushed chunk is a file on disk when `_chunk_files` is looked at, the aggregate equals the input."""
    from pdensorflow_b200.gpuSolve.IO.writers import BaseWriter
    n, N, nx = 10, 3, 128
    fname = str(tmp_path / 'every_n')
    writer = BaseWriter({'fname': fname, 'every_N': N})
    rng = np.random.default_rng(0)
    frames = [rng.uniform(size=(nx,)).astype(np.float32) for _ in range(n)]
    ref = np.stack(frames, axis=0)
    prev = 0
    for i, frame in enumerate(frames, start=1):
        writer.add_solution(frame)
        expected = i // N
        assert writer.nb_chunks() == (prev + 1 if i % N == 0 else prev)
        prev = writer.nb_chunks()
        assert len(writer._buffer) == i - expected * N
        assert len(writer._chunk_files) == expected
        for cf in writer._chunk_files:
            assert os.path.exists(cf)
    assert len(writer._buffer) == n % N
    writer.finalize()
    out = np.load(fname + '.npy')
    assert out.shape == (n, nx)
    np.testing.assert_allclose(out, ref, rtol=1e-6, atol=1e-6)
    assert writer._chunk_files == []
    # memory trigger: 0.76 MB frames against a 1 MB threshold -> a dump every second frame
    nelem = 200000
    fname = str(tmp_path / 'max_mb')
    writer = BaseWriter({'fname': fname, 'max_chunk_mb': 1})
    assert writer._every_N is None
    frames = [np.ones((nelem,), dtype=np.float32) * float(k) for k in range(5)]
    writer.add_solution(frames[0])
    assert writer.nb_chunks() == 0
    writer.add_solution(frames[1])
    assert writer.nb_chunks() == 1 and len(writer._chunk_files) == 1 and os.path.exists(writer._chunk_files[0])
    assert len(writer._buffer) == 0
    writer.add_solution(frames[2])
    assert writer.nb_chunks() == 1
    writer.add_solution(frames[3])
    assert writer.nb_chunks() == 2
    writer.add_solution(frames[4])
    assert writer.nb_chunks() == 2
    writer.finalize()
    out = np.load(fname + '.npy')
    assert out.shape == (len(frames), nelem)
    np.testing.assert_allclose(out, np.stack(frames), rtol=1e-6, atol=1e-6)
