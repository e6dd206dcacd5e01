"""IGBWriter - mirror of gpuSolve/IO/writers/igbwriter.py:7-315 on the pinned-memory pipeline of BaseWriter.

IGB (meshalyzer / openCARP) layout: a 1024-byte text header - ``key:value`` tokens separated by blanks, CRLF
lines of at most 80 bytes, padded with blank CRLF lines and closed by a form feed (igbwriter.py:232-278) -
followed by the frames as little-endian fp32.  The frames are appended chunk by chunk by BaseWriter's
background thread, so there are no temporary files and nothing to aggregate.
"""
import os

import numpy as np

from .basewriter import BaseWriter

_HEADER_BYTES = 1024


def igb_header_bytes(nx, ny, nz, nt, org_t, dim_t, units_x=None, units_y=None, units_z=None, units_t=None, units=None):
    """The 1024 header bytes of an IGB file of fp32 little-endian frames."""
    keyvals = ['x:{}'.format(nx), 'y:{}'.format(ny), 'z:{}'.format(nz), 't:{}'.format(nt), 'type:float',
               'systeme:little_endian', 'org_t:{}'.format(org_t), 'dim_t:{}'.format(dim_t)]
    if units_x is not None and units_y is not None and units_z is not None:
        keyvals += ['unites_x:{}'.format(units_x), 'unites_y:{}'.format(units_y), 'unites_z:{}'.format(units_z)]
    if units_t is not None:
        keyvals.append('unites_t:{}'.format(units_t))
    if units is not None:
        keyvals.append('unites:{}'.format(units))
    lines, line = [], ''
    for kv in keyvals:                                   # greedy fill of 78-character lines
        if line and len(line) + 1 + len(kv) > 78:
            lines.append(line)
            line = kv
        else:
            line = kv if not line else line + ' ' + kv
    if line:
        lines.append(line)
    header = bytearray(('\r\n'.join(lines) + '\r\n').encode('utf8'))
    if len(header) >= _HEADER_BYTES:
        raise ValueError('IGB header does not fit 1024 bytes')
    while _HEADER_BYTES - len(header) > 80:
        header += b' ' * 78 + b'\r\n'
    header += b' ' * (_HEADER_BYTES - len(header) - 1) + b'\x0c'
    return bytes(header)


class IGBWriter(BaseWriter):
    def __init__(self, config=None):
        super().__init__(config)
        self._fname = 'out.igb'
        self._nt = None
        self._nx = None
        self._ny = None
        self._nz = None
        self._org_t = 0.0
        self._Tend = None
        self._dim_t = None
        self._units_x = None
        self._units_y = None
        self._units_z = None
        self._units_t = None
        self._units = None
        self.__fobj = None
        self.__hwrt = True
        self.__data = None
        if config is not None:
            for attribute in list(self.__dict__.keys()):
                if attribute[1:] in config.keys() and not attribute.startswith('_bw_') and '__' not in attribute:
                    setattr(self, attribute, config[attribute[1:]])
        if self._dim_t is None and self._Tend is not None:
            self._dim_t = self._Tend - self._org_t
        elif self._Tend is None and self._dim_t is not None:
            self._Tend = self._org_t + self._dim_t

    # ---- configuration ----
    def set_fname(self, fname):
        self._fname = fname

    def set_space_units(self, space_units):
        self._units_x = self._units_y = self._units_z = space_units

    def set_time_unit(self, time_unit):
        self._units_t = time_unit

    def set_variable_unit(self, variable_unit):
        self._units = variable_unit

    def set_nt(self, nt):
        self._nt = nt

    def set_nx(self, nx):
        self._nx = nx

    def set_ny(self, ny=1):
        self._ny = ny

    def set_nz(self, nz=1):
        self._nz = nz

    def set_org_t(self, t0):
        self._org_t = t0
        if self._Tend is not None:
            self._dim_t = self._Tend - self._org_t
        elif self._dim_t is not None:
            self._Tend = self._org_t + self._dim_t

    def set_dim_t(self, dim_t):
        self._dim_t = dim_t
        self._Tend = self._org_t + self._dim_t

    def set_data(self, data):
        self.__data = data

    def nx(self):
        return self._nx

    def ny(self):
        return self._ny

    def nz(self):
        return self._nz

    def nt(self):
        return self._nt

    def org_t(self):
        return self._org_t

    def dim_t(self):
        return self._dim_t

    def units_x(self):
        return self._units_x

    def units_y(self):
        return self._units_y

    def units_z(self):
        return self._units_z

    def units_t(self):
        return self._units_t

    def units(self):
        return self._units

    def data(self):
        return self.__data

    # ---- writing ----
    def imshow(self, data):
        """Receive one solution (device tensor or array); it reaches the file in the background."""
        self.add_solution(data)

    def _write_chunk(self, chunk, idx):
        if self.__hwrt:
            self.__write_header()
        np.ascontiguousarray(chunk, dtype='<f4').tofile(self.__fobj)
        return None

    def _aggregate(self, chunk_files):
        pass

    def wait(self):
        self.finalize()
        if self.__fobj is not None:
            self.__fobj.close()
            self.__fobj = None
        self._bw_shutdown()

    def write_data_to_file(self):
        """Stand-alone use: write the header and the (nt, nx) array installed with set_data / initialise_*."""
        if self.__hwrt:
            self.__write_header()
        np.ascontiguousarray(self.__data[:self._nt], dtype='<f4').tofile(self.__fobj)
        self.__fobj.close()
        self.__fobj = None

    def initialise_from_data(self, header, data):
        self.__parse_header(header)
        self.__data = data

    def initialise_from_IGBReader(self, reader):
        self.__parse_header(reader.header())
        self.__data = reader.data()

    def __del__(self):
        try:
            if self.__fobj is not None:
                self.__fobj.close()
            self._bw_shutdown()
        except Exception:
            pass

    def __write_header(self):
        self.__openfile()
        if self._dim_t is None:
            self._dim_t = self._nt - 1
        if self._ny is None:
            self._ny = 1
        if self._nz is None:
            self._nz = 1
        self.__fobj.write(igb_header_bytes(self._nx, self._ny, self._nz, self._nt, self._org_t, self._dim_t, self._units_x,
                                           self._units_y, self._units_z, self._units_t, self._units))
        self.__hwrt = False

    def __openfile(self):
        if self.__fobj is None:
            fdir, fname = os.path.split(self._fname)
            if fname.rfind('.') == -1:
                self._fname = '{}.igb'.format(self._fname)
            if fdir and not os.path.exists(fdir):
                os.makedirs(fdir)
            self.__fobj = open(self._fname, 'wb')

    def __parse_header(self, header):
        for attribute, key in zip(['_nx', '_ny', '_nz', '_nt'], ['x', 'y', 'z', 't']):
            if key in header.keys():
                setattr(self, attribute, header[key])
        for attribute in list(self.__dict__.keys()):
            if attribute[1:] in header.keys() and not attribute.startswith('_bw_') and '__' not in attribute:
                setattr(self, attribute, header[attribute[1:]])
        if self._dim_t is None and self._Tend is not None:
            self._dim_t = self._Tend - self._org_t
        elif self._Tend is None and self._dim_t is not None:
            self._Tend = self._org_t + self._dim_t
