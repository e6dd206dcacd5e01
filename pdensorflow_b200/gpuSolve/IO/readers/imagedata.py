"""ImageData - mirror of gpuSolve/IO/readers/imagedata.py:6-156: voxel images -> 3-D NumPy arrays for ``Domain3D``
(SURVEY.md section 8f, N4).

The reference decodes PNG mosaics with imageio and NIfTI volumes with nibabel; neither package exists in this image,
and the ingestion is a one-time host step, so both formats are read here with the standard library only:

* PNG (``load_png_image``, imagedata.py:35-55): 8/16-bit greyscale, non-interlaced - zlib inflate + the five PNG
  scanline filters; the ``Mx x My`` mosaic of slices is unfolded with the reference's slice numbering
  (slice z = cy*Mx + cx, array axes (row-in-slice, column-in-slice, slice)) by one reshape/transpose instead of the
  reference's per-pixel Python loop.
* NIfTI-1 single-file ``.nii`` / ``.nii.gz`` (imagedata.py:88-89): header + raw voxels (Fortran order), ``scl_slope`` /
  ``scl_inter`` applied by ``get_fdata`` as nibabel does.
* ``.npy`` (imagedata.py:90-91).

Same class and method names; ``image()`` returns the small ``VoxelImage`` holder that stands in for nibabel's
``Nifti1Image`` (``get_fdata()``, ``get_data()``, ``shape``, ``affine``).
"""
import gzip
import struct
import sys
import zlib

import numpy as np


def parse_name(fname):
    """File type and compression from a file name (imagedata.py:6-32): {'gzipped', 'type', 'name', 'fname'}."""
    parsed = {'gzipped': False, 'type': 'unknown', 'name': '', 'fname': fname}
    parts = fname.strip().split('/')[-1].strip().split('.')
    if len(fname) == 0:
        return parsed
    if len(parts) == 1:
        parsed['name'] = parts[0]
        return parsed
    if parts[-1] == 'gz':
        parsed['gzipped'] = True
        parts = parts[:-1]
    if len(parts) == 2:
        parsed['type'], parsed['name'] = parts[-1], parts[-2]
    else:
        parsed['name'] = parts[-1]
    return parsed


# ------------------------------------------------------------------------------------------------ PNG
_PNG_SIG = b'\x89PNG\r\n\x1a\n'


def _unfilter(raw, height, stride, bpp):
    """Undo the per-scanline PNG filters (None, Sub, Up, Average, Paeth)."""
    out = np.zeros((height, stride), dtype=np.uint8)
    prev = np.zeros(stride, dtype=np.uint8)
    pos = 0
    for y in range(height):
        ft = raw[pos]
        line = np.frombuffer(raw, dtype=np.uint8, count=stride, offset=pos + 1)
        pos += stride + 1
        if ft == 0:
            cur = line.copy()
        elif ft == 2:
            cur = line + prev                                   # uint8 arithmetic wraps mod 256, as the format wants
        elif ft == 1:
            cur = (np.cumsum(line.reshape(-1, bpp).astype(np.uint32), axis=0) & 255).astype(np.uint8).reshape(-1)
        elif ft in (3, 4):
            cur = np.zeros(stride, dtype=np.uint8)
            lp, pp = line.tolist(), prev.tolist()
            c = [0] * stride
            for i in range(stride):
                a = c[i - bpp] if i >= bpp else 0
                b = pp[i]
                if ft == 3:
                    pred = (a + b) >> 1
                else:
                    d = pp[i - bpp] if i >= bpp else 0
                    p = a + b - d
                    pa, pb, pc = abs(p - a), abs(p - b), abs(p - d)
                    pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else d)
                c[i] = (lp[i] + pred) & 255
            cur[:] = c
        else:
            raise ValueError('corrupt PNG: unknown filter type %d' % ft)
        out[y] = cur
        prev = cur
    return out


def read_png_gray(fname):
    """Decode a non-interlaced 8- or 16-bit greyscale PNG into a 2-D array [height, width] (uint8 / uint16),
    what ``imageio.imread`` returns for the reference's mosaics."""
    with open(fname, 'rb') as f:
        data = f.read()
    if data[:8] != _PNG_SIG:
        raise ValueError('%s is not a PNG file' % fname)
    pos, idat, hdr = 8, [], None
    while pos + 8 <= len(data):
        length, ctype = struct.unpack('>I4s', data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + length]
        if ctype == b'IHDR':
            hdr = struct.unpack('>IIBBBBB', body)
        elif ctype == b'IDAT':
            idat.append(body)
        elif ctype == b'IEND':
            break
        pos += 12 + length
    if hdr is None:
        raise ValueError('corrupt PNG: no IHDR chunk')
    width, height, depth, ctype, _comp, _filt, interlace = hdr
    if ctype != 0 or depth not in (8, 16) or interlace != 0:
        raise NotImplementedError('PNG mosaics must be non-interlaced 8/16-bit greyscale (colour type %d, depth %d, '
                                  'interlace %d)' % (ctype, depth, interlace))
    bpp = depth // 8
    stride = width * bpp
    raw = zlib.decompress(b''.join(idat))
    if len(raw) != height * (stride + 1):
        raise ValueError('corrupt PNG: %d bytes of image data, expected %d' % (len(raw), height * (stride + 1)))
    px = _unfilter(raw, height, stride, bpp)
    if depth == 16:
        return px.reshape(height, width, 2).astype(np.uint16) @ np.array([256, 1], dtype=np.uint16)
    return px


def write_png_gray(fname, img):
    """Encode a 2-D uint8 / uint16 array as a greyscale PNG (filter type Up on every scanline).  Used to produce
    mosaics for tests and examples; the reference has no writer."""
    img = np.ascontiguousarray(img)
    if img.ndim != 2 or img.dtype not in (np.uint8, np.uint16):
        raise ValueError('write_png_gray takes a 2-D uint8 / uint16 array')
    h, w = img.shape
    depth = 8 * img.dtype.itemsize
    rows = img.astype('>u2').view(np.uint8).reshape(h, 2 * w) if depth == 16 else img
    up = rows.copy()
    up[1:] = rows[1:] - rows[:-1]
    raw = np.concatenate([np.full((h, 1), 2, dtype=np.uint8), up], axis=1).tobytes()

    def chunk(ctype, body):
        return struct.pack('>I', len(body)) + ctype + body + struct.pack('>I', zlib.crc32(ctype + body) & 0xffffffff)
    with open(fname, 'wb') as f:
        f.write(_PNG_SIG + chunk(b'IHDR', struct.pack('>IIBBBBB', w, h, depth, 0, 0, 0, 0)) +
                chunk(b'IDAT', zlib.compress(raw, 6)) + chunk(b'IEND', b''))


def mosaic_to_volume(im, mx, my):
    """Unfold an ``mx x my`` mosaic [H, L] into [h, l, mx*my] with slice z = cy*mx + cx (imagedata.py:42-54)."""
    H, L = im.shape
    h, l = int(H / my), int(L / mx)
    if h * my != H or l * mx != L:
        raise ValueError('a %dx%d image is not a %dx%d mosaic' % (H, L, mx, my))
    return np.ascontiguousarray(im.reshape(my, h, mx, l).transpose(1, 3, 0, 2).reshape(h, l, my * mx))


def volume_to_mosaic(vol, mx, my):
    """Inverse of ``mosaic_to_volume`` (for writing test / example inputs)."""
    h, l, d = vol.shape
    if d != mx * my:
        raise ValueError('%d slices do not fill a %dx%d mosaic' % (d, mx, my))
    return np.ascontiguousarray(vol.reshape(h, l, my, mx).transpose(2, 0, 3, 1).reshape(my * h, mx * l))


def load_png_image(imgfile, mx, my):
    """A domain stored as a PNG mosaic of slices on an ``mx x my`` grid -> [h, l, mx*my] (imagedata.py:35-55)."""
    return mosaic_to_volume(read_png_gray(imgfile), mx, my)


# ---------------------------------------------------------------------------------------------- NIfTI-1
_NII_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16,
               768: np.uint32, 1024: np.int64, 1280: np.uint64}


class VoxelImage:
    """What the reference gets from nibabel: the voxel array + an affine (``Nifti1Image(data, affine)``)."""

    def __init__(self, data, affine=None, slope=1.0, inter=0.0):
        self._data = np.asarray(data)
        self.affine = np.eye(4) if affine is None else np.asarray(affine, dtype=np.float64)
        self._slope, self._inter = slope, inter

    @property
    def shape(self):
        return self._data.shape

    def get_data(self):
        return self._data

    def get_fdata(self):
        out = self._data.astype(np.float64)
        if self._slope not in (0.0, 1.0) or self._inter != 0.0:
            if self._slope != 0.0 and np.isfinite(self._slope):
                out = out * self._slope + (self._inter if np.isfinite(self._inter) else 0.0)
        return out


def read_nifti(fname):
    """Single-file NIfTI-1 (.nii / .nii.gz)."""
    opener = gzip.open if fname.endswith('.gz') else open
    with opener(fname, 'rb') as f:
        raw = f.read()
    end = '<' if struct.unpack('<i', raw[:4])[0] == 348 else '>'
    if struct.unpack(end + 'i', raw[:4])[0] != 348:
        raise ValueError('%s is not a NIfTI-1 file' % fname)
    dim = struct.unpack(end + '8h', raw[40:56])
    datatype, bitpix = struct.unpack(end + 'hh', raw[70:74])
    vox_offset = int(struct.unpack(end + 'f', raw[108:112])[0])
    slope, inter = struct.unpack(end + 'ff', raw[112:120])
    if raw[344:348] not in (b'n+1\x00',):
        raise NotImplementedError('only single-file NIfTI-1 (magic n+1) is read')
    if datatype not in _NII_DTYPES:
        raise NotImplementedError('NIfTI datatype %d' % datatype)
    shape = tuple(int(v) for v in dim[1:1 + dim[0]])
    dt = np.dtype(_NII_DTYPES[datatype]).newbyteorder(end)
    n = int(np.prod(shape))
    data = np.frombuffer(raw, dtype=dt, count=n, offset=max(vox_offset, 352)).reshape(shape, order='F')
    srow = np.array(struct.unpack(end + '12f', raw[280:328]), dtype=np.float64).reshape(3, 4)
    affine = np.vstack([srow, [0.0, 0.0, 0.0, 1.0]])
    return VoxelImage(np.ascontiguousarray(data.astype(dt.newbyteorder('='))), affine, float(slope), float(inter))


def write_nifti(fname, img):
    """Minimal single-file NIfTI-1 writer (``ImageData.save_nifty``)."""
    data = np.asarray(img.get_data())
    code = [k for k, v in _NII_DTYPES.items() if np.dtype(v) == data.dtype]
    if not code:
        data, code = data.astype(np.float32), [16]
    hdr = bytearray(352)
    struct.pack_into('<i', hdr, 0, 348)
    dim = [data.ndim] + list(data.shape) + [1] * (7 - data.ndim)
    struct.pack_into('<8h', hdr, 40, *dim)
    struct.pack_into('<hh', hdr, 70, code[0], 8 * data.dtype.itemsize)
    struct.pack_into('<8f', hdr, 76, 1.0, *([1.0] * 7))
    struct.pack_into('<f', hdr, 108, 352.0)
    struct.pack_into('<ff', hdr, 112, 1.0, 0.0)
    struct.pack_into('<hh', hdr, 252, 0, 2)                 # qform 0, sform 2 (aligned)
    struct.pack_into('<12f', hdr, 280, *np.asarray(img.affine, dtype=np.float64)[:3].ravel())
    hdr[344:348] = b'n+1\x00'
    opener = gzip.open if fname.endswith('.gz') else open
    with opener(fname, 'wb') as f:
        f.write(bytes(hdr) + data.astype(data.dtype.newbyteorder('<')).tobytes(order='F'))


# ------------------------------------------------------------------------------------------------ class
class ImageData:
    """Reads NIfTI .nii(.gz), .npy, or a .png mosaic into a 3-D array (imagedata.py:61-156)."""

    def __init__(self):
        self.Mx = 1
        self.My = 1
        self._img = None
        self._imgfile = None

    def load_image(self, fname, mx=1, my=1):
        """For png images the mosaic grid (mx, my) must be given."""
        print('load file {0}'.format(fname), flush=True)
        self._imgfile = parse_name(fname)
        kind = self._imgfile['type']
        if kind == 'png':
            self.Mx, self.My = mx, my
            self._img = VoxelImage(load_png_image(fname, mx, my), np.eye(4))
        elif kind == 'nii':
            self._img = read_nifti(fname)
        elif kind == 'npy':
            self._img = VoxelImage(np.load(fname), np.eye(4))
        else:
            sys.exit('file type {0} not wnown'.format(kind))

    def save_nifty(self, fout):
        if self._img:
            print('saving NIfTI image in {0}'.format(fout), flush=True)
            write_nifti(fout, self._img)

    def get_data(self):
        return self._img.get_data() if self._img else None

    def get_fdata(self):
        return self._img.get_fdata() if self._img else None

    def get_rescaled_data(self, scaling_type='unit'):
        """'unit': (x - min) / (max - min); 'mstd': (x - mean) / std (imagedata.py:126-148)."""
        if not self._img:
            return None
        x = self._img.get_fdata()
        if scaling_type == 'unit':
            mval = np.nanmin(x)
            delta = np.nanmax(x) - mval
        elif scaling_type == 'mstd':
            mval, delta = np.nanmean(x), np.nanstd(x)
        else:
            sys.exit('unknown scaling method')
        return (x - mval) / delta

    def image(self):
        return self._img
