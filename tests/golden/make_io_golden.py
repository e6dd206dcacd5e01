"""Generate tests/golden/ref_io.npz by EXECUTING THE REFERENCE'S OWN WRITERS (build container only).

    python tests/golden/make_io_golden.py

The reference's gpuSolve/IO/writers/{basewriter,igbwriter,resultwriter}.py are imported unmodified from
/root/reference with tests/golden/tf_facade.py standing in for tensorflow (the writers only use
tf.identity / tf.convert_to_tensor / tf.stack(...).numpy()).  They write into a temporary directory; the
bytes of the .igb files and the arrays of the .npy files are frozen as fixtures for tests/test_io_cpu.py.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_facade as tf  # noqa: E402

tf.install()


def frames(nt, shape, seed=0):
    rng = np.random.default_rng(seed)
    return [rng.uniform(-85.0, 25.0, size=shape).astype(np.float32) for _ in range(nt)]


def main():
    from gpuSolve.IO.writers import IGBWriter, ResultWriter
    out = {}
    tmp = tempfile.mkdtemp()
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        # --- IGB, the configuration of Tests/FEM/Fenton/fenton.py:113-118 shrunk: nt frames of an [N,1] potential
        nx, nt = 257, 7
        fr = frames(nt, (nx, 1), seed=1)
        w = IGBWriter({'fname': 'a.igb', 'Tend': 600.0, 'nt': nt, 'nx': nx, 'every_N': 3})
        for f in fr:
            w.imshow(f)
        w.wait()
        out['igb_a_bytes'] = np.frombuffer(open('a.igb', 'rb').read(), dtype=np.uint8)
        out['igb_a_frames'] = np.stack(fr)
        # --- IGB with every optional header key (long header: line wrapping at 78 characters)
        nx, ny, nz, nt = 12, 5, 3, 4
        fr = frames(nt, (nx * ny * nz,), seed=2)
        w = IGBWriter({'fname': 'sub/b', 'nt': nt, 'nx': nx, 'ny': ny, 'nz': nz, 'org_t': 2.5, 'dim_t': 30.0,
                       'units_x': 'mm', 'units_y': 'mm', 'units_z': 'mm', 'units_t': 'ms', 'units': 'mV', 'max_chunk_mb': 0.001})
        for f in fr:
            w.imshow(f)
        w.wait()
        out['igb_b_bytes'] = np.frombuffer(open('sub/b.igb', 'rb').read(), dtype=np.uint8)
        out['igb_b_frames'] = np.stack(fr)
        out['igb_b_nchunks'] = np.int64(w.nb_chunks())
        # --- ResultWriter, dense: Tests/FD/Fenton/fenton.py:185-191 shrunk
        W, H, D, nt = 4, 5, 6, 5
        fr = frames(nt, (W, H, D), seed=3)
        rw = ResultWriter({'width': W, 'height': H, 'depth': D, 'prefix_name': 'cube3D', 'every_N': 2})
        for f in fr:
            rw.imshow(f)
        rw.wait()
        out['rw_dense_name'] = np.array('cube3D_{}_{}_{}.npy'.format(H, W, D))
        out['rw_dense'] = np.load('cube3D_{}_{}_{}.npy'.format(H, W, D))
        out['rw_dense_nchunks'] = np.int64(rw.nb_chunks())
        out['rw_frames'] = np.stack(fr)
        # --- ResultWriter, sparse domain (Tests/FD/Fenton_sphere-style compact output)
        dom = np.zeros((W, H, D), dtype=bool)
        dom[1:3, 1:4, 2:5] = True
        rw = ResultWriter({'width': W, 'height': H, 'depth': D, 'prefix_name': 'sp'})
        rw.set_sparse_domain(dom)
        for f in fr:
            rw.imshow(f)
        rw.wait()
        name = 'sp_sparse_{}_{}_{}.npy'.format(H, W, D)
        d = np.load(name, allow_pickle=True).item()
        out['rw_sparse_name'] = np.array(name)
        out['rw_sparse_dimensions'] = np.array(d['dimensions'], dtype=np.int64)
        out['rw_sparse_indices'] = np.stack([np.asarray(i) for i in d['indices']])
        out['rw_sparse_values'] = np.asarray(d['values'])
        out['rw_domain'] = dom
        out['leftovers'] = np.array(sorted(f for f in os.listdir('.') if '__chunk_' in f) or [''])
    finally:
        os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, 'ref_io.npz'), **out)
    print({k: (v.shape, str(v.dtype)) for k, v in out.items()})


if __name__ == '__main__':
    main()
