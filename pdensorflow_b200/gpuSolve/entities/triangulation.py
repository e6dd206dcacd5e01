"""Triangulation - mesh container, mirror of gpuSolve/entities/triangulation.py:143-430 (host side).

Connectivity and point-region ids are computed with vectorised NumPy instead of the reference's
per-node Python loops (triangulation.py:361-383,410-430): same results, usable at 2e7 nodes.
"""
import pickle

import numpy as np


class Triangulation:
    def __init__(self):
        self._Pts = None
        self._Elems = None
        self._Fibres = None
        self._connectivity = None
        self._pointRegIDs = None

    def Pts(self):
        return self._Pts

    def Elems(self):
        return self._Elems

    def Fibres(self):
        return self._Fibres

    def set_mesh(self, Pts, Elems, Fibres=None):
        """Install an in-memory mesh (same schema as the pickle)."""
        self._Pts = np.asarray(Pts, dtype=np.float64)
        self._Elems = {k: np.asarray(v).astype(np.int32) for k, v in Elems.items() if v is not None}
        self._Fibres = None if Fibres is None else np.asarray(Fibres, dtype=np.float64)
        self._connectivity = None
        self._pointRegIDs = None

    def readMesh(self, filename):
        """'.pkl' pickle {'Pts','Elems','Fibres'} (triangulation.py:326-342).  CARP text meshes need the
        reference's CarpMeshReader, which is out of scope here."""
        if not filename.endswith('.pkl'):
            raise NotImplementedError('only .pkl meshes are supported by pdensorflow_b200 (CARP reader out of scope)')
        with open(filename, 'rb') as f:
            mesh0 = pickle.load(f)
        self.set_mesh(mesh0['Pts'], mesh0['Elems'], mesh0['Fibres'])

    def saveMesh(self, foutName):
        with open('{}.pkl'.format(foutName), 'wb') as f:
            pickle.dump({'Pts': self._Pts, 'Elems': self._Elems, 'Fibres': self._Fibres}, f)

    def exportCarpFormat(self, foutPrefix):
        """Write .pts/.elem/.lon text files (Tests/FEM/Fenton/fenton.py:111 calls this)."""
        np.savetxt(foutPrefix + '.pts', self._Pts, header=str(self._Pts.shape[0]), comments='')
        names = {'Trias': 'Tr', 'Tetras': 'Tt', 'Edges': 'Ln'}
        nel = sum(v.shape[0] for v in self._Elems.values())
        with open(foutPrefix + '.elem', 'w') as f:
            f.write('%d\n' % nel)
            for k, E in self._Elems.items():
                for row in E:
                    f.write(names.get(k, k) + ' ' + ' '.join(str(int(x)) for x in row) + '\n')
        if self._Fibres is not None:
            np.savetxt(foutPrefix + '.lon', self._Fibres, header='1', comments='')

    def mesh_connectivity(self, storeConn=False):
        """dict node -> sorted unique neighbour ids (itself included)."""
        if self._connectivity is not None:
            return self._connectivity
        npt = self._Pts.shape[0]
        rows, cols = [np.arange(npt, dtype=np.int64)], [np.arange(npt, dtype=np.int64)]
        for E in self._Elems.values():
            nodes = E[:, :-1].astype(np.int64)
            nv = nodes.shape[1]
            for a in range(nv):
                for b in range(nv):
                    if a != b:
                        rows.append(nodes[:, a]); cols.append(nodes[:, b])
        key = np.unique(np.concatenate(rows) * npt + np.concatenate(cols))
        r, c = key // npt, (key % npt).astype(np.int32)
        start = np.searchsorted(r, np.arange(npt + 1))
        conn = _CSRConnectivity(start, c)
        if storeConn in (True, 'True'):
            self._connectivity = conn
        return conn

    def point_region_ids(self, storeIDs=False):
        """Most frequent region id among the elements touching each point (triangulation.py:410-430)."""
        if self._pointRegIDs is not None:
            return self._pointRegIDs
        npt = self._Pts.shape[0]
        pts, regs = [], []
        for E in self._Elems.values():
            nv = E.shape[1] - 1
            pts.append(E[:, :-1].ravel()); regs.append(np.repeat(E[:, -1], nv))
        pts, regs = np.concatenate(pts).astype(np.int64), np.concatenate(regs).astype(np.int64)
        nreg = int(regs.max()) + 1
        counts = np.bincount(pts * nreg + regs, minlength=npt * nreg).reshape(npt, nreg)
        ids = np.argmax(counts, axis=1).astype(np.float64)
        if storeIDs:
            self._pointRegIDs = ids
        return ids

    def release_connectivity(self):
        self._connectivity = None

    def release_point_region_ids(self):
        self._pointRegIDs = None

    def release_contravariant_basis(self):
        pass


class _CSRConnectivity:
    """Read-only dict-like view (node -> neighbour array) over a CSR adjacency."""

    def __init__(self, start, col):
        self.start, self.col = start, col

    def __len__(self):
        return len(self.start) - 1

    def __getitem__(self, i):
        return self.col[self.start[i]:self.start[i + 1]]

    def items(self):
        for i in range(len(self)):
            yield i, self[i]

    def keys(self):
        return range(len(self))
