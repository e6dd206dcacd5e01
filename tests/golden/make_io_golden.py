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
    out.update(domain_golden())
    np.savez_compressed(os.path.join(HERE, 'ref_io.npz'), **out)
    print({k: (v.shape, str(v.dtype)) for k, v in out.items()})



def domain_golden():
    """Domain3D of the reference (entities/domain3D.py) on seeded arrays: the assign_* path of the FD examples
    (Tests/FD/Fenton_sphere/fenton.py:86-101) and the anisotropic conductivity / fibre-dyadic conventions."""
    # the reference's Domain3D imports its ImageData reader (imageio / nibabel): not needed on the assign path
    from gpuSolve.entities.domain3D import Domain3D
    rng = np.random.default_rng(11)
    W, H, D = 6, 5, 7
    geo = (rng.uniform(size=(W, H, D)) > 0.3).astype(np.float64)
    out = {'dom_geo': geo}
    d = Domain3D({'width': 3, 'height': 4, 'depth': 5, 'dx': 0.5, 'dy': 0.25, 'dz': 2.0})
    d.load_geometry_file()
    d.load_conductivity(cond_unif=0.8)
    out['dom_box_geometry'] = d.geometry()
    out['dom_box_conductivity'] = d.conductivity()
    out['dom_box_meta'] = np.array([d.width(), d.height(), d.depth(), d.dx(), d.dy(), d.dz(), float(d.anisotropic())])
    d = Domain3D({'dx': 1.0})
    d.assign_geometry(geo.copy())
    d.assign_conductivity(0.8 * geo)
    out['dom_iso_conductivity'] = d.conductivity()
    out['dom_iso_meta'] = np.array([d.width(), d.height(), d.depth(), float(d.anisotropic())])
    cond4 = rng.uniform(0.1, 1.0, size=(W, H, D, 1))
    d.assign_conductivity(cond4.copy())
    out['dom_cond4_in'], out['dom_cond4'] = cond4, d.conductivity()
    aniso = np.stack([rng.uniform(0.05, 0.2, size=(W, H, D)), rng.uniform(0.5, 1.0, size=(W, H, D))], axis=-1)
    d.assign_conductivity(aniso.copy())
    out['dom_aniso_in'], out['dom_aniso'] = aniso, d.conductivity()
    out['dom_aniso_flag'] = np.array(float(d.anisotropic()))
    return out


if __name__ == '__main__':
    main()
