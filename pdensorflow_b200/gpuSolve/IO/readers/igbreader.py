"""IGBReader - mirror of gpuSolve/IO/readers/igbreader.py:7-176.

IGB layout: a 1024-byte text header of ``key:value`` tokens, then nt frames of nx*ny*nz little-endian
fp32 values.  The reference exposes the data as an (nt, x) array, truncating or dropping trailing values
when the file length disagrees with the header (igbreader.py:33-56); so does this reader.
"""
import numpy as np

_HEADER_BYTES = 1024
_INT_KEYS = ('x', 'y', 'z', 't')
_DEFAULTS = (('y', 1), ('z', 1), ('org_t', 0.0), ('dim_t', None), ('unites_x', 'unk'), ('unites_y', 'unk'),
             ('unites_z', 'unk'), ('unites_t', 'unk'), ('unites', 'unk'), ('facteur', 1), ('zero', 0))


def parse_igb_header(raw):
    """dict of the header tokens: x/y/z/t as int, numbers as float, the rest as str; missing optional
    keys get the reference's defaults (igbreader.py:128-172)."""
    out = {}
    for tok in raw.decode('utf-8', errors='replace').split():
        kv = tok.split(':', 1)
        if len(kv) != 2:
            continue
        key, val = kv
        if key in _INT_KEYS:
            try:
                out[key] = int(val)
            except ValueError:
                continue
        else:
            try:
                out[key] = float(val)
            except ValueError:
                out[key] = val
    for key, dflt in _DEFAULTS:
        if key not in out:
            out[key] = float(out['t'] - 1) if key == 'dim_t' else dflt
    return out


class IGBReader:
    def __init__(self):
        self.__header = None
        self.__data = None
        self.__filename = None
        self.__ndiff = 0

    def read(self, igbfname):
        self.__filename = igbfname
        with open(igbfname, 'rb') as f:
            hdr = parse_igb_header(f.read(_HEADER_BYTES))
            y = np.fromfile(f, dtype='<f4')
        nt, nx = hdr['t'], hdr['x']
        nentries, ntot = y.shape[0], nt * nx
        self.__ndiff = nentries - ntot
        if nentries > ntot:
            print('Warning: problem with the igb file', flush=True)
            print('Discarding the last {} elements'.format(self.__ndiff))
            y = y[:ntot]
        elif nentries < ntot:
            nt = nentries // nx
            if nt == 0:
                raise SystemExit('ERROR: {}: {} values, expected {} (problems with the igb file)'.format(igbfname, nentries, ntot))
            print('Warning: problem with the igb file', flush=True)
            print('Missing {} elements to reach {}'.format(nentries % nx, hdr['t']))
            print('Reshaping to {} time steps'.format(nt))
            y = y[:nt * nx]
        hdr['t'] = nt
        self.__header = hdr
        self.__data = np.array(y.reshape(nt, nx), dtype=np.float32)

    def header(self):
        return self.__header

    def data(self):
        return self.__data

    def filename(self):
        return self.__filename

    def ndiff(self):
        return self.__ndiff

    def org_t(self):
        return self.__header['org_t']

    def nt(self):
        return self.__header['t']

    def nx(self):
        return self.__header['x']

    def ny(self):
        return self.__header['y']

    def nz(self):
        return self.__header['z']

    def dim_t(self):
        return self.__header['dim_t']

    def dt(self):
        return self.dim_t() / (self.nt() - 1) if self.nt() > 1 else np.nan

    def timevalues(self, shifted=False):
        tline = self.dt() * np.arange(self.nt())
        return tline if shifted else self.org_t() + tline
