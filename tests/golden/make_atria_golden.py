"""Generate tests/golden/ref_atria.npz + tests/golden/atria_mosaic.png by EXECUTING THE REFERENCE'S anatomical FD
example (build container only: it reads /root/reference).

    python tests/golden/make_atria_golden.py

Tests/FD/Fenton_atria/fenton.py is imported unmodified over tests/golden/tf_facade.py and run end to end on the
reference's own anatomy (Tests/data/structure.png: a 16 x 8 mosaic of 128 x 128 slices): PNG mosaic -> ImageData ->
Domain3D geometry (unit rescaling, threshold 1e-4) -> DIFF = diff * geometry -> heterogeneous Laplacian + Fenton-Karma,
initial slab of excited voxels, S2 on half of the anatomy.  The reference's readers need imageio and nibabel, which
are not in this image: ``imageio.imread`` is served by PIL (what imageio itself delegates to) and ``nibabel`` by a
20-line array holder - the mosaic unfolding, the rescaling and the thresholding are the reference's own code.

Frozen: the decoded volume (packed bits), three frames of the run (i = 0, 40, 80; fp16-exact? no: fp32, compressed),
the V gate at the end.  The anatomy is re-encoded by OUR PNG writer into atria_mosaic.png so that the GPU box (which
has no /root/reference) can run the same ingestion path from a file.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)


def _install_reader_shims():
    from PIL import Image
    imageio = types.ModuleType('imageio')
    imageio.imread = lambda f: np.asarray(Image.open(f))
    sys.modules['imageio'] = imageio

    class Nifti1Image:
        def __init__(self, data, affine):
            self._d, self.affine = np.asarray(data), affine

        def get_fdata(self):
            return self._d.astype(np.float64)

        def get_data(self):
            return self._d

    nib = types.ModuleType('nibabel')
    nib.Nifti1Image = Nifti1Image
    sys.modules['nibabel'] = nib


_install_reader_shims()
import tf_facade as tf  # noqa: E402

tf.install()
from make_ref_golden import REF, load_example  # noqa: E402

CFG = {'width': 64, 'height': 64, 'depth': 64, 'dx': 1, 'dy': 1, 'dz': 1, 'Mx': 16, 'My': 8, 'dt': 0.1, 'dt_per_plot': 40,
       'diff': 0.75, 'samples': 81, 's2_time': 5.0}


class Capture:
    def __init__(self):
        self.frames = []

    def imshow(self, U):
        self.frames.append(np.array(U))

    def wait(self):
        pass


def main():
    png = os.path.join(REF, 'Tests', 'data', 'structure.png')
    ex = load_example('Tests/FD/Fenton_atria/fenton.py', 'ref_fd_atria_example')
    # the reader on its own: reference ImageData vs our standard-library decoder
    from gpuSolve.IO.readers.imagedata import ImageData as RefImageData
    from pdensorflow_b200.gpuSolve.IO.readers import imagedata as ours
    ref_img = RefImageData()
    ref_img.load_image(png, CFG['Mx'], CFG['My'])
    vol = np.asarray(ref_img.get_data())
    mine = ours.ImageData()
    mine.load_image(png, CFG['Mx'], CFG['My'])
    assert np.array_equal(vol, mine.get_data()) and np.array_equal(ref_img.get_rescaled_data('unit'), mine.get_rescaled_data('unit'))
    print('volume', vol.shape, vol.dtype, 'tissue fraction', float((vol > 0).mean()))
    cfg = dict(CFG, fname=png)
    model = ex.Fenton4vSimple(cfg)
    cap = Capture()
    model.run(cap)
    frames = np.stack(cap.frames).astype(np.float32)
    geo = np.asarray(model.domain())
    out = {'cfg': np.array([CFG[k] for k in ('Mx', 'My', 'dt', 'dt_per_plot', 'diff', 'samples', 's2_time')], dtype=np.float64),
           'volume_shape': np.array(vol.shape), 'volume_bits': np.packbits(vol > 0), 'volume_sum': np.array(int(vol.astype(np.int64).sum())),
           'geometry_bits': np.packbits(geo > 0), 'frames': frames, 'V_end': model._V_state.numpy().astype(np.float32)}
    np.savez_compressed(os.path.join(HERE, 'ref_atria.npz'), **out)
    ours.write_png_gray(os.path.join(HERE, 'atria_mosaic.png'), ours.volume_to_mosaic(vol, CFG['Mx'], CFG['My']))
    back = ours.load_png_image(os.path.join(HERE, 'atria_mosaic.png'), CFG['Mx'], CFG['My'])
    assert np.array_equal(back, vol)
    print('frames', frames.shape, [(float(f.min()), float(f.max())) for f in frames],
          os.path.getsize(os.path.join(HERE, 'ref_atria.npz')), os.path.getsize(os.path.join(HERE, 'atria_mosaic.png')))


if __name__ == '__main__':
    main()
