"""Device-side assembly of the P1 mass and stiffness matrices.

Mirror of ``assemble_matrices_dict`` (gpuSolve/matrices/globalMatrices.py:612-675) for the case every
reference example uses - ``localMass`` + ``localStiffness`` with region-based ``sigma_l`` / ``sigma_t`` and
element fibres (or no conductivities at all: identity tensor): the sparsity pattern is built from the element
vertex pairs (sort + unique of 64-bit keys, torch on the device), the local matrices are evaluated and
accumulated by ``pdx_fem_assemble`` (one thread per element, fp64 atomics into CSR order) and rounded to fp32
once.  The host NumPy path stays the default for small meshes (it is the one pinned bit for bit to the
reference's NumPy assembly); ``HeatSolver`` switches to this one above ``DEVICE_ASSEMBLY_MIN_ELEMENTS``
elements or when cfg ``device_assembly`` is True.
"""
from time import time

import numpy as np
import torch

from .. import _backend as B
from ... import _lib as L

DEVICE_ASSEMBLY_MIN_ELEMENTS = 200000
_NV = {'Edges': 2, 'Trias': 3, 'Tetras': 4}


def supported(domain, matprops):
    """True when the conductivities are region maps (+ fibres) or absent, i.e. no per-element Python callback."""
    if any(k not in _NV for k in domain.Elems().keys()):
        return False
    names = matprops.element_property_names()
    if names is None or not ('sigma_l' in names or 'sigma_t' in names):
        return True
    return ('sigma_l' in names and 'sigma_t' in names and domain.Fibres() is not None
            and matprops.element_property_type('sigma_l') == 'region'
            and matprops.element_property_type('sigma_t') == 'region')


def _region_values(idmap, regions, dev):
    keys = np.array(sorted(idmap.keys()), dtype=np.int64)
    lut = torch.zeros(int(keys.max()) + 1, dtype=torch.float64, device=dev)
    lut[torch.from_numpy(keys).to(dev)] = torch.tensor([float(idmap[int(k)]) for k in keys], dtype=torch.float64, device=dev)
    return lut[regions]


def _sigma_device(Elements_d, domain, matprops, dev):
    """[ne][6] fp64 (xx xy xz yy yz zz) of sigma_t*I + (sigma_l - sigma_t) f (x) f, or None.  Like the reference
    (globalMatrices.py:262-286) the fibres of element e are row e of domain.Fibres() for every element type."""
    names = matprops.element_property_names()
    if names is None or 'sigma_l' not in names:
        return None
    reg = Elements_d[:, -1].long()
    sl = _region_values(matprops._element_properties['sigma_l']['idmap'], reg, dev)
    st = _region_values(matprops._element_properties['sigma_t']['idmap'], reg, dev)
    ne = Elements_d.shape[0]
    f = torch.from_numpy(np.ascontiguousarray(domain.Fibres()[:ne], dtype=np.float64)).to(dev)
    d = (sl - st)
    out = torch.empty((ne, 6), dtype=torch.float64, device=dev)
    out[:, 0] = st + d * f[:, 0] * f[:, 0]
    out[:, 1] = d * f[:, 0] * f[:, 1]
    out[:, 2] = d * f[:, 0] * f[:, 2]
    out[:, 3] = st + d * f[:, 1] * f[:, 1]
    out[:, 4] = d * f[:, 1] * f[:, 2]
    out[:, 5] = st + d * f[:, 2] * f[:, 2]
    return out


def assemble_on_device(domain, matprops, renumbering=None, names=('mass', 'stiffness'), device=None):
    """dict name -> CSRMatrix (host copies + device arrays already in place) of the P1 mass / stiffness matrices."""
    from .globalMatrices import CSRMatrix
    if not supported(domain, matprops):
        raise L.PdxError('device assembly takes region-based sigma_l / sigma_t with fibres (or none); use the host path')
    t0 = time()
    dev = B.device() if device is None else torch.device(device)
    npt = domain.Pts().shape[0]
    pts = torch.from_numpy(np.ascontiguousarray(domain.Pts(), dtype=np.float64)).to(dev)
    rowid = None
    if renumbering is not None:
        rowid = torch.from_numpy(np.ascontiguousarray(renumbering['iperm'], dtype=np.int32)).to(dev)
    blocks = []
    keys = []
    for elemtype, Elements in domain.Elems().items():
        E = torch.from_numpy(np.ascontiguousarray(Elements, dtype=np.int32)).to(dev)
        nv = _NV[elemtype]
        ids = E[:, :nv].long()
        if rowid is not None:
            ids = rowid.long()[ids]
        keys.append((ids.unsqueeze(2) * npt + ids.unsqueeze(1)).reshape(-1))       # all nv^2 (row, col) pairs
        blocks.append((nv, E, _sigma_device(E, domain, matprops, dev)))
    uniq = torch.unique(torch.cat(keys), sorted=True)
    del keys
    nnz = int(uniq.numel())
    col = (uniq % npt).to(torch.int32)
    rows = uniq // npt
    rowptr = torch.searchsorted(rows, torch.arange(npt + 1, device=dev, dtype=torch.int64)).to(torch.int32)
    del uniq, rows
    want_m, want_k = 'mass' in names, any(n != 'mass' for n in names)
    macc = torch.zeros(nnz, dtype=torch.float64, device=dev) if want_m else None
    kacc = torch.zeros(nnz, dtype=torch.float64, device=dev) if want_k else None
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    for nv, E, sig in blocks:
        L.check(L.lib().pdx_fem_assemble(nv, E.shape[0], L.ptr(E), E.shape[1], L.ptr(pts), L.ptr(rowid), L.ptr(sig),
                                         L.ptr(rowptr), L.ptr(col), L.ptr(macc), L.ptr(kacc), L.ptr(status), L.stream_ptr()))
    out = {}
    rp_h, col_h = rowptr.cpu().numpy(), col.cpu().numpy()
    if int(status.item()) != 0:
        raise L.PdxError('device assembly: an element entry is missing from the sparsity pattern')
    for name in names:
        acc = macc if name == 'mass' else kacc
        val = torch.empty(nnz, dtype=torch.float32, device=dev)
        L.check(L.lib().pdx_round_to_f32(L.ptr(acc), L.ptr(val), nnz, L.stream_ptr()))
        M = CSRMatrix(rp_h, col_h, val.cpu().numpy(), npt)
        M._dev = (rowptr, col, val)               # already resident: no re-upload
        out[name] = M
    print('Device assembly of {} done in {:3.2f} s'.format(list(names), time() - t0), flush=True)
    return out
