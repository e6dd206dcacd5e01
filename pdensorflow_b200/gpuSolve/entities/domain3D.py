"""Domain3D - mirror of gpuSolve/entities/domain3D.py:7-232: the voxel domain of the FD examples (shape, spacings,
geometry mask, conductivity field, fibre dyadic) that feeds ``laplace_heterog`` / ``laplace_heterog_aniso``
(SURVEY.md section 8f, N4).  Host container over NumPy arrays, as in the reference.

``assign_*`` is what Tests/FD/Fenton_sphere/fenton.py:86-101 uses; ``load_*`` read PNG mosaics (Mx x My grid of
slices, Tests/FD/Fenton_atria/fenton.py:88-89), NIfTI-1 and ``.npy`` files through ``gpuSolve.IO.readers.imagedata``
(standard-library decoders; the reference needs imageio / nibabel).
"""
import sys
import time

import numpy as np


def _load_image(fname, Mx, My):
    from ..IO.readers.imagedata import ImageData
    image = ImageData()
    image.load_image(fname, Mx, My)
    return image


class Domain3D:
    def __init__(self, props=None):
        self._width = None
        self._height = None
        self._depth = None
        self._dx = 1.0
        self._dy = 1.0
        self._dz = 1.0
        self.__geometry = None
        self.__conductivity = None
        self.__fibtensor = None
        self._geometryfile = ''
        self._conductivityfile = ''
        self._fibtensorfile = ''
        self.__timecounter = 0.0
        self.__anisotropic = False
        if props:
            for attribute in list(self.__dict__.keys()):
                if attribute[1:] in props.keys():
                    setattr(self, attribute, props[attribute[1:]])

    # ---- geometry ----
    def load_geometry_file(self, fname='', Mx=0, My=0, threshold=1.e-4):
        """Geometry from a voxel file, or (no file) a full box of (width, height, depth) voxels."""
        then = time.time()
        if self._geometryfile == '':
            self._geometryfile = fname
        if len(self._geometryfile):
            print('reading the geometry from the file {0}'.format(fname))
            vox = _load_image(self._geometryfile, Mx, My).get_rescaled_data('unit')   # (x - min) / (max - min), float64
            self._width, self._height, self._depth = vox.shape
            vox[vox > threshold] = 1.0                                           # domain3D.py:56-58
            vox[vox <= threshold] = 0.0
            self.__geometry = vox
            print('Sizes: ({}X{}X{})'.format(self._width, self._height, self._depth))
        else:
            print('Cubic geometry ({}X{}X{})'.format(self._width, self._height, self._depth))
            self.__geometry = np.ones(shape=(self._width, self._height, self._depth))
        self.__timecounter += time.time() - then

    def assign_geometry(self, geometryTensor):
        then = time.time()
        self.__geometry = geometryTensor
        self._width, self._height, self._depth = geometryTensor.shape
        self.__timecounter += time.time() - then

    # ---- conductivity ----
    def __set_conductivity(self, vox):
        """Isotropic [W,H,D] (a trailing singleton channel is dropped) or anisotropic [W,H,D,2] = (sigma_t, sigma_l),
        stored as (sigma_t, sigma_l - sigma_t) - the form laplace_heterog_aniso takes (domain3D.py:111-117)."""
        if len(vox.shape) == 4 and vox.shape[-1] > 1:
            print('anisotropic conductivity (need fibres)')
            self.__anisotropic = True
            vox[:, :, :, 1] = vox[:, :, :, 1] - vox[:, :, :, 0]
        else:
            print('isotropic conductivity')
            self.__anisotropic = False
            if len(vox.shape) == 4:
                vox = vox[:, :, :, 0]
        self.__conductivity = vox

    def load_conductivity(self, fname='', Mx=0, My=0, cond_unif=1.0):
        if self._conductivityfile == '':
            self._conductivityfile = fname
        if self.__geometry is None:
            sys.exit("No geometry defined! (assign a geometry first)")
        then = time.time()
        if len(self._conductivityfile):
            print('reading the conductivity from the file {0}'.format(fname))
            self.__set_conductivity(_load_image(self._conductivityfile, Mx, My).get_fdata())
        else:
            print('homogeneous conductivity')
            self.__conductivity = cond_unif * self.__geometry
        self.__timecounter += time.time() - then

    def assign_conductivity(self, conductivityTensor):
        then = time.time()
        if self.__geometry is None:
            sys.exit("No geometry defined! (assign a geometry first)")
        self.__set_conductivity(conductivityTensor)
        self.__timecounter += time.time() - then

    # ---- fibres ----
    def load_fiber_direction(self, fname='', Mx=0, My=0):
        """Fibre unit vectors [W,H,D,3] -> dyadic [W,H,D,6] = a0a0 a0a1 a0a2 a1a1 a1a2 a2a2 (domain3D.py:171-185)."""
        if self._fibtensorfile == '':
            self._fibtensorfile = fname
        if self.__geometry is None:
            sys.exit("No geometry defined! (assign a geometry first)")
        if not len(self._fibtensorfile):
            sys.exit("Invalid fiber file!")
        then = time.time()
        print('reading the fibres from the file {0}'.format(fname))
        v = _load_image(self._fibtensorfile, Mx, My).get_fdata()
        W, H, D, _ = v.shape
        fibtens = np.zeros(shape=(W, H, D, 6))
        for c, (a, b) in enumerate(((0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2))):
            fibtens[:, :, :, c] = v[:, :, :, a] * v[:, :, :, b]
        self.__fibtensor = fibtens
        self.__timecounter += time.time() - then

    # ---- setters / getters ----
    def set_dx(self, dx):
        self._dx = dx

    def set_dy(self, dy):
        self._dy = dy

    def set_dz(self, dz):
        self._dz = dz

    def dx(self):
        return self._dx

    def dy(self):
        return self._dy

    def dz(self):
        return self._dz

    def width(self):
        return self._width

    def height(self):
        return self._height

    def depth(self):
        return self._depth

    def anisotropic(self):
        return self.__anisotropic

    def geometry(self):
        return self.__geometry

    def conductivity(self):
        return self.__conductivity

    def fibtensor(self):
        return self.__fibtensor

    def walltime(self):
        return self.__timecounter
