"""Global FEM matrices - mirror of gpuSolve/matrices/globalMatrices.py on NumPy (host assembly)
+ device CSR (int32 indices, fp32 values) consumed by libpdx's SpMV / PCG kernels.

``CSRMatrix`` plays the role of TF's CSRSparseMatrix wrapper.  Assembly follows the reference
(globalMatrices.py:612-675): triplets over all local (i,j) pairs, node ids mapped through
``iperm`` when renumbering is on, values cast to fp32 and then summed over duplicates, sorted
row-major; the pattern keeps the explicit structural zeros.
"""
from time import time

import numpy as np
import torch
from scipy.sparse import csr_matrix
from scipy.sparse.csgraph import reverse_cuthill_mckee

from .. import _backend as B
from ... import _lib as L
from .localMass import localMass, batch_local_mass
from .localStiffness import localStiffness, batch_local_stiffness


class CSRMatrix:
    """Square CSR matrix: host arrays (rowptr, col int32; val fp32) + lazily created device copies."""

    def __init__(self, rowptr, col, val, n):
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        self.col = np.ascontiguousarray(col, dtype=np.int32)
        self.val = np.ascontiguousarray(val, dtype=np.float32)
        self.n = int(n)
        self.dtype = np.float32
        self._dev = None

    @property
    def nnz(self):
        return int(self.col.shape[0])

    def device_arrays(self):
        if self._dev is None:
            dev = B.device()
            self._dev = (torch.from_numpy(self.rowptr).to(dev), torch.from_numpy(self.col).to(dev),
                         torch.from_numpy(self.val).to(dev))
            # lanes-per-row of the CSR kernels come from nnz / n: registered once here, never read back from the device
            L.check(L.lib().pdx_csr_hint(L.ptr(self._dev[0]), self.n, self.nnz))
        return self._dev

    def to_scipy(self):
        return csr_matrix((self.val, self.col, self.rowptr), shape=(self.n, self.n))

    def coo(self):
        """(I, J, V) in row-major order (what JacobiPrecond.build_preconditioner takes)."""
        I = np.repeat(np.arange(self.n, dtype=np.int32), np.diff(self.rowptr))
        return I, self.col, self.val

    def diagonal_rowsum(self):
        return np.asarray(self.to_scipy().sum(axis=1)).ravel().astype(np.float32)

    def matvec(self, x, out=None):
        """y = A x on the device (pdx_csr_spmv); x is an [N] or [N,1] CUDA fp32 tensor."""
        x = L.require_cuda(x, 'x')
        rp, ci, va = self.device_arrays()
        y = torch.empty_like(x) if out is None else out
        L.check(L.lib().pdx_csr_spmv(self.n, L.ptr(rp), L.ptr(ci), L.ptr(va), L.ptr(x), L.ptr(y), L.stream_ptr()))
        return y


def compute_coo_pattern(connectivity):
    """Sparsity pattern (I, J, StartIndex) from the mesh connectivity (globalMatrices.py:15-52)."""
    if hasattr(connectivity, 'start'):
        start, J = connectivity.start.astype(np.int32), connectivity.col.astype(np.int32)
    else:
        npt = len(connectivity)
        counts = np.array([len(connectivity[i]) for i in range(npt)])
        start = np.zeros(npt + 1, dtype=np.int32)
        start[1:] = np.cumsum(counts)
        J = np.concatenate([np.asarray(connectivity[i]) for i in range(npt)]).astype(np.int32)
    I = np.repeat(np.arange(len(start) - 1, dtype=np.int32), np.diff(start))
    return {'I': I, 'J': J, 'StartIndex': start}


def compute_reverse_cuthill_mckee_indexing(matrix_pattern, sym_mat=True):
    """perm / iperm concentrating the entries around the diagonal (globalMatrices.py:54-74)."""
    indptr, indices = matrix_pattern['StartIndex'], matrix_pattern['J']
    npt = indptr.shape[0] - 1
    A0 = csr_matrix((np.ones(indices.shape, dtype=np.float32), indices, indptr), shape=(npt, npt))
    perm = reverse_cuthill_mckee(A0, symmetric_mode=sym_mat).astype(np.int32)
    iperm = np.zeros(npt, dtype=np.int32)
    iperm[perm] = np.arange(npt, dtype=np.int32)
    return {'perm': perm, 'iperm': iperm}


def _contravariant_basis(elemtype, Pts, nodes):
    V0 = Pts[nodes[:, 0]]
    v10 = Pts[nodes[:, 1]] - V0
    if elemtype == 'Edges':
        ne = nodes.shape[0]
        meas = np.linalg.norm(v10, axis=1, keepdims=True)
        out = {k: np.zeros((ne, 3)) for k in ('v1', 'v2', 'v3')}
        for e in range(ne):
            u, _, _ = np.linalg.svd(v10[e, :, None])
            con = np.linalg.inv(np.stack([v10[e], u[:, 1], u[:, 2]], axis=0))
            out['v1'][e], out['v2'][e], out['v3'][e] = con[:, 0], con[:, 1], con[:, 2]
        out['meas'] = meas
        return out
    v20 = Pts[nodes[:, 2]] - V0
    if elemtype == 'Trias':
        nrm = np.cross(v10, v20)
        a2 = np.linalg.norm(nrm, axis=1, keepdims=True)
        third, meas = nrm / a2, 0.5 * a2
    elif elemtype == 'Tetras':
        third = Pts[nodes[:, 3]] - V0
        meas = np.abs(np.sum(np.cross(v10, v20) * third, axis=1, keepdims=True)) / 6.0
    else:
        raise NotImplementedError('element type %s' % elemtype)
    con = np.linalg.inv(np.stack([v10, v20, third], axis=1))
    return {'v1': con[:, :, 0], 'v2': con[:, :, 1], 'v3': con[:, :, 2], 'meas': meas}


def _sigma_batch(matr_name, elemtype, Elements, domain, matprops):
    """Diffusion tensors of all elements: vectorised for region-based sigma_l/sigma_t + fibres
    (globalMatrices.py:245-287), otherwise through the user callback, identity when it returns None."""
    ne = Elements.shape[0]
    names = matprops.element_property_names()
    fibs = domain.Fibres()
    if (names is not None and 'sigma_l' in names and 'sigma_t' in names and fibs is not None
            and matprops.element_property_type('sigma_l') == 'region'
            and matprops.element_property_type('sigma_t') == 'region'):
        reg = Elements[:, -1]
        lmap = matprops._element_properties['sigma_l']['idmap']
        tmap = matprops._element_properties['sigma_t']['idmap']
        sl = np.array([lmap[int(r)] for r in reg], dtype=np.float64)
        st = np.array([tmap[int(r)] for r in reg], dtype=np.float64)
        f = fibs[:ne]
        return st[:, None, None] * np.eye(3)[None] + (sl - st)[:, None, None] * (f[:, :, None] * f[:, None, :])
    Sig = np.zeros((ne, 3, 3))
    for e in range(ne):
        s = matprops.execute_ud_func(matr_name, elemtype, e, domain, matprops)
        Sig[e] = np.identity(3) if s is None else s
    return Sig


def assemble_matrices_dict(local_matrices_dict, matrix_pattern, domain, matprops, connectivity=None,
                           renumbering=None):
    """dict name -> CSRMatrix for every local-matrix function in ``local_matrices_dict``
    (localMass / localStiffness are vectorised; other callables run per element)."""
    t0 = time()
    npt = domain.Pts().shape[0]
    rows, cols = [], []
    vals = {name: [] for name in local_matrices_dict}
    for elemtype, Elements in domain.Elems().items():
        nodes = Elements[:, :-1]
        basis = _contravariant_basis(elemtype, domain.Pts(), nodes)
        meas = basis['meas']
        lm = {}
        for name, fn in local_matrices_dict.items():
            if fn is localMass:
                lm[name] = batch_local_mass(elemtype, meas)
            elif fn is localStiffness:
                lm[name] = batch_local_stiffness(elemtype, basis, meas,
                                                 _sigma_batch(name, elemtype, Elements, domain, matprops))
            else:
                ne, nv = nodes.shape
                out = np.zeros((ne, nv, nv))
                for e in range(ne):
                    ed = {k: basis[k][e] for k in ('v1', 'v2', 'v3')}
                    ed['meas'] = meas[e, 0]
                    out[e] = fn(elemtype, ed, matprops.execute_ud_func(name, elemtype, e, domain, matprops))
                lm[name] = out
        nv = nodes.shape[1]
        for a in range(nv):
            for b in range(nv):
                rows.append(nodes[:, a]); cols.append(nodes[:, b])
                for name in local_matrices_dict:
                    vals[name].append(lm[name][:, a, b])
    r = np.concatenate(rows).astype(np.int64)
    c = np.concatenate(cols).astype(np.int64)
    if renumbering is not None:
        iperm = renumbering['iperm']
        r, c = iperm[r].astype(np.int64), iperm[c].astype(np.int64)
    uniq, inv = np.unique(r * npt + c, return_inverse=True)
    ur, uc = (uniq // npt).astype(np.int32), (uniq % npt).astype(np.int32)
    rowptr = np.searchsorted(ur, np.arange(npt + 1)).astype(np.int32)
    MATRICES = {}
    for name in local_matrices_dict:
        v32 = np.concatenate(vals[name]).astype(np.float32)
        acc = np.zeros(uniq.shape[0], dtype=np.float32)
        np.add.at(acc, inv, v32)
        MATRICES[name] = CSRMatrix(rowptr, uc, acc, npt)
    print('Assembly of {} done in {:3.2f} s'.format(list(local_matrices_dict.keys()), time() - t0), flush=True)
    return MATRICES


def assemble_mass_matrix(matrix_pattern, domain, connectivity=None, renumbering=None):
    from ..entities.materialproperties import MaterialProperties
    return assemble_matrices_dict({'mass': localMass}, matrix_pattern, domain, MaterialProperties(),
                                  connectivity, renumbering)['mass']


def assemble_stiffness_matrix(matrix_pattern, domain, matprops, stif_pname='Sigma', connectivity=None,
                              renumbering=None):
    return assemble_matrices_dict({stif_pname: localStiffness}, matrix_pattern, domain, matprops,
                                  connectivity, renumbering)[stif_pname]


def csr_axpby(A, alpha, B=None, beta=None):
    """alpha*A + beta*B (or alpha*A) as a CSRMatrix (globalMatrices.py:465-492).  A and B must
    share their pattern, which holds for matrices assembled from the same mesh."""
    a = np.float32(alpha)
    if B is None:
        return CSRMatrix(A.rowptr, A.col, A.val * a, A.n)
    if A.nnz != B.nnz or not (np.array_equal(A.rowptr, B.rowptr) and np.array_equal(A.col, B.col)):
        C = (A.to_scipy() * a + B.to_scipy() * np.float32(beta)).tocsr()
        C.sort_indices()
        return CSRMatrix(C.indptr, C.indices, C.data.astype(np.float32), A.n)
    return CSRMatrix(A.rowptr, A.col, A.val * a + B.val * np.float32(beta), A.n)


def csr_to_sparse_tensor(A):
    return A.coo()
