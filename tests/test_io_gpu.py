"""Writers fed with DEVICE tensors from inside a stepping loop: the snapshot is taken on
the stepping stream, the transfer runs on the copy engine into pinned memory, the file is written by a
background thread - and every frame must be exactly the field at the moment imshow() was called, although the
solver overwrites its buffers immediately afterwards."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_igb_frames_equal_synchronous_snapshots(cuda, tmp_path):
    from pdensorflow_b200.fd import FDStepper
    from pdensorflow_b200.gpuSolve.IO.readers import IGBReader
    from pdensorflow_b200.gpuSolve.IO.writers import IGBWriter
    shape = (24, 32, 40)
    n = int(np.prod(shape))
    st = FDStepper(shape, (1.0, 1.0, 1.0), 0.1, 0.8, model='fenton4v', math='fast')
    U = np.full(shape, -80.0, dtype=np.float32)
    U[0:2] = 20.0
    st.set_U(U)
    nframes = 12
    w = IGBWriter({'fname': str(tmp_path / 'fd.igb'), 'nt': nframes, 'nx': n, 'Tend': float(nframes - 1), 'every_N': 5})
    snaps = []
    for k in range(nframes):
        w.imshow(st.U())                       # no synchronisation here
        if k in (0, 5, 11):
            snaps.append((k, st.U().cpu().numpy().reshape(-1).copy()))
        st.step(7)                             # overwrites both ping-pong buffers right away
    w.wait()
    r = IGBReader()
    r.read(str(tmp_path / 'fd.igb'))
    assert r.nt() == nframes and r.nx() == n and r.ndiff() == 0
    for k, s in snaps:
        assert np.array_equal(r.data()[k], s)
    assert not np.array_equal(r.data()[0], r.data()[11])
    assert w.nb_chunks() == 3


def test_result_writer_with_device_frames_matches_host_frames(cuda, tmp_path):
    from pdensorflow_b200.gpuSolve.IO.writers import ResultWriter
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        W, H, D = 6, 10, 12
        rng = np.random.default_rng(0)
        frames = [rng.standard_normal((W, H, D)).astype(np.float32) for _ in range(9)]
        buf = torch.empty((W, H, D), device='cuda')
        rw = ResultWriter({'width': W, 'height': H, 'depth': D, 'prefix_name': 'dev', 'max_chunk_mb': 0.005})
        for f in frames:
            buf.copy_(torch.from_numpy(f))
            rw.imshow(buf)                     # same device buffer every time: the writer must have snapshotted it
        rw.wait()
        assert np.array_equal(np.load('dev_{}_{}_{}.npy'.format(H, W, D)), np.stack(frames))
    finally:
        os.chdir(cwd)


def test_monodomain_example_loop_writes_igb(cuda, tmp_path):
    """The loop of Tests/FEM/Fenton/fenton.py:113-132 (writer fed with model.U() every dt_per_plot steps)."""
    from test_fem_gpu import _build_model
    from pdensorflow_b200.gpuSolve.IO.readers import IGBReader
    from pdensorflow_b200.gpuSolve.IO.writers import IGBWriter
    model, _ = _build_model(21, True)
    npt = model.domain().Pts().shape[0]
    nt, every = 40, 10
    im = IGBWriter({'fname': str(tmp_path / 'square.igb'), 'Tend': nt * model.dt(), 'nt': 1 + nt // every, 'nx': npt})
    im.imshow(model.U())
    keep = [model.U().cpu().numpy().ravel().copy()]
    t = 0.0
    for k in range(1, nt + 1):
        t += model.dt()
        model.step(t)
        if k % every == 0:
            im.imshow(model.U())
            keep.append(model.U().cpu().numpy().ravel().copy())
    im.wait()
    r = IGBReader()
    r.read(str(tmp_path / 'square.igb'))
    assert r.nt() == 1 + nt // every and r.nx() == npt
    assert np.array_equal(r.data(), np.stack(keep))
