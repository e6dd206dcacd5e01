"""ResultWriter - mirror of gpuSolve/IO/writers/resultwriter.py:5-145 on the pinned-memory pipeline of BaseWriter.

The FD examples hand every ``dt_per_plot``-th potential field to ``imshow`` and call ``wait()`` at the end
(Tests/FD/Fenton/fenton.py:149-164,191); the result is ``{prefix_name}_{height}_{width}_{depth}.npy`` holding
the frames stacked along axis 0 (the reference's naming puts height before width).  The reference dumps
chunk files and re-reads them through a memmap to build that file; here the background thread APPENDS each
chunk to the final .npy directly - the header is written with room for the frame count and patched when the
run ends - so every byte goes to disk once.  The sparse form (``set_sparse_domain``) keeps only the values
inside the domain mask, chunk by chunk, and is saved as the reference's pickled dict.
"""
import os

import numpy as np

from .basewriter import BaseWriter

_NPY_HEADER_BYTES = 256        # fixed-size header so that the shape can be rewritten in place


def _npy_header(shape, dtype):
    d = "{{'descr': {!r}, 'fortran_order': False, 'shape': {!r}, }}".format(np.lib.format.dtype_to_descr(np.dtype(dtype)),
                                                                            tuple(int(v) for v in shape))
    body = d.encode('latin1')
    pad = _NPY_HEADER_BYTES - 10 - len(body) - 1
    if pad < 0:
        raise ValueError('npy header too long')
    hdr = body + b' ' * pad + b'\n'
    return b'\x93NUMPY\x01\x00' + np.uint16(len(hdr)).astype('<u2').tobytes() + hdr


class ResultWriter(BaseWriter):
    def __init__(self, config=None):
        super().__init__(config)
        self.width = 1
        self.height = 1
        self.depth = 1
        self.samples = 1
        self.dt_per_plot = 1
        self.not_saved = True
        self.prefix_name = 'cube3D'
        self._sparse = False
        self.__save_on_exit = True
        self.initval = 0.0
        self.__cube = None
        self.__cube_exists = False
        self.__counter = None
        self.__use_cube = False
        self.__fobj = None
        self.__nframes = 0
        self.__frame = None
        self.__dtype = None
        self.__sparse_values = []
        if config:
            for attribute in list(self.__dict__.keys()):
                if attribute in config.keys():
                    setattr(self, attribute, config[attribute])
        self._sparse = False

    # ---- legacy pre-allocated host cube (resultwriter.py:41-47) ----
    def initialise_cube(self):
        n = 1 + int(self.samples // self.dt_per_plot)
        self.__cube = np.full(shape=(n, self.width, self.height, self.depth), fill_value=self.initval, dtype=np.float32)
        self.__counter = 0
        self.__cube_exists = True
        self.__use_cube = True

    def disable_save_on_exit(self):
        self.__save_on_exit = False

    def enable_save_on_exit(self):
        self.__save_on_exit = True

    def set_sparse_domain(self, domain):
        self._domain = np.asarray(domain).astype(bool)
        self._sparse = True

    def imshow(self, VolData):
        if self.__use_cube:
            if not self.__cube_exists:
                self.initialise_cube()
            v = VolData.detach().cpu().numpy() if hasattr(VolData, 'detach') else np.asarray(VolData)
            self.__cube[self.__counter, :, :, :] = v
            self.__counter += 1
        else:
            self.add_solution(VolData)

    def wait(self):
        if self.not_saved:
            self.save()
        self.not_saved = False
        self._bw_shutdown()

    def save(self):
        if self.__use_cube:
            self.__save_array(self.__cube)
        else:
            self.finalize()

    def __save_array(self, cube):
        if self._sparse:
            cube_sparse = {'dimensions': cube.shape, 'indices': np.where(self._domain), 'values': cube[:, self._domain]}
            fname = '{0}_sparse_{1}_{2}_{3}'.format(self.prefix_name, self.height, self.width, self.depth)
            print('saving file {0}'.format(fname))
            np.save(fname, cube_sparse)
        else:
            fname = '{0}_{1}_{2}_{3}'.format(self.prefix_name, self.height, self.width, self.depth)
            print('saving file {0}'.format(fname))
            np.save(fname, cube)

    def _tmp_prefix(self):
        return '{0}_{1}_{2}_{3}'.format(self.prefix_name, self.height, self.width, self.depth)

    def _final_path(self):
        return '{0}_{1}_{2}_{3}.npy'.format(self.prefix_name, self.height, self.width, self.depth)

    # ---- streaming output (writer thread) ----
    def _write_chunk(self, chunk, idx):
        if self.__frame is None:
            self.__frame, self.__dtype = tuple(chunk.shape[1:]), chunk.dtype
        if self._sparse:
            self.__sparse_values.append(np.array(chunk[:, self._domain]))
        else:
            if self.__fobj is None:
                target = self._final_path()
                fdir = os.path.split(target)[0]
                if fdir and not os.path.exists(fdir):
                    os.makedirs(fdir)
                self.__fobj = open(target, 'wb')
                self.__fobj.write(_npy_header((0,) + self.__frame, self.__dtype))
            np.ascontiguousarray(chunk).tofile(self.__fobj)
        self.__nframes += chunk.shape[0]
        return None

    def _aggregate(self, chunk_files):
        if self._sparse:
            if self.__nframes == 0:
                dims, vals = (0,), np.empty((0,), dtype=np.float32)
            else:
                dims, vals = (self.__nframes,) + self.__frame, np.concatenate(self.__sparse_values, axis=0)
            fname = '{0}_sparse_{1}_{2}_{3}'.format(self.prefix_name, self.height, self.width, self.depth)
            print('saving file {0}'.format(fname))
            np.save(fname, {'dimensions': dims, 'indices': np.where(self._domain), 'values': vals})
            self.__sparse_values = []
            return
        target = self._final_path()
        if self.__fobj is None:                               # no frame received
            np.save(target, np.empty((0,), dtype=np.float32))
            return
        self.__fobj.seek(0)
        self.__fobj.write(_npy_header((self.__nframes,) + self.__frame, self.__dtype))
        self.__fobj.close()
        self.__fobj = None
        print('saving file {0}'.format(target))

    def __del__(self):
        try:
            if self.not_saved and self.__save_on_exit and (self.__use_cube or self.counter() > 0):
                self.save()
            self._bw_shutdown()
        except Exception:
            pass
