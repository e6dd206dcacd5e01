"""Deterministic synthetic meshes (the reference's bundled meshes are missing from the checkout:
/root/reference/.MISSING_LARGE_BLOBS).  Schema of the reference's mesh pickle
(entities/triangulation.py:326-342): Pts float64 [N,3]; Elems {'Trias'|'Tetras'|'Edges': int32
[nE, nV+1]} (last column = region id); Fibres float64 [nE,3]."""
import pickle

import numpy as np


def triangulated_square(nx=251, ny=251, Lx=10.0, Ly=10.0):
    """Stand-in for triangulated_square.pkl: nx*ny nodes on [0,Lx]x[0,Ly] (z=0), every cell split
    into 2 P1 triangles, region ids 1-4 by quadrant, unit fibres along x.
    251x251 gives the documented 63 001 nodes / 125 000 triangles / nnz 439 001."""
    xs, ys = np.meshgrid(np.linspace(0.0, Lx, nx), np.linspace(0.0, Ly, ny), indexing='ij')
    P = np.stack([xs.ravel(), ys.ravel(), np.zeros(nx * ny)], axis=1)
    i, j = np.meshgrid(np.arange(nx - 1), np.arange(ny - 1), indexing='ij')
    i, j = i.ravel(), j.ravel()
    n00, n10, n11, n01 = i * ny + j, (i + 1) * ny + j, (i + 1) * ny + j + 1, i * ny + j + 1
    reg = 1 + (i >= (nx - 1) // 2).astype(np.int64) + 2 * (j >= (ny - 1) // 2).astype(np.int64)
    T = np.concatenate([np.stack([n00, n10, n11, reg], axis=1), np.stack([n00, n11, n01, reg], axis=1)])
    F = np.tile(np.array([[1.0, 0.0, 0.0]]), (T.shape[0], 1))
    return P, {'Trias': T.astype(np.int32)}, F


def cable(n_el=150, L=2.0):
    """1-D cable of the reference's unit test (Tests/CICD/unit/test_mms_1d.py:67-78)."""
    x = np.linspace(0.0, L, n_el + 1)
    P = np.stack([x, np.zeros_like(x), np.zeros_like(x)], axis=1)
    E = np.stack([np.arange(n_el), np.arange(1, n_el + 1), np.ones(n_el, dtype=np.int64)], axis=1)
    F = np.tile(np.array([[1.0, 0.0, 0.0]]), (n_el, 1))
    return P, {'Edges': E.astype(np.int32)}, F


def tetrahedral_box(nx, ny, nz, Lx=1.0, Ly=1.0, Lz=1.0):
    """Structured box, each hexahedral cell split into 6 tetrahedra around the main diagonal
    (config 5: 272^3 nodes ~ 2.0e7)."""
    xs, ys, zs = np.meshgrid(np.linspace(0, Lx, nx), np.linspace(0, Ly, ny), np.linspace(0, Lz, nz), indexing='ij')
    P = np.stack([xs.ravel(), ys.ravel(), zs.ravel()], axis=1)
    i, j, k = np.meshgrid(np.arange(nx - 1), np.arange(ny - 1), np.arange(nz - 1), indexing='ij')
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    idx = lambda a, b, c: ((a * ny) + b) * nz + c
    c = [idx(i, j, k), idx(i + 1, j, k), idx(i + 1, j + 1, k), idx(i, j + 1, k),
         idx(i, j, k + 1), idx(i + 1, j, k + 1), idx(i + 1, j + 1, k + 1), idx(i, j + 1, k + 1)]
    reg = np.ones_like(i)
    tets = [np.stack([c[0], c[a], c[b], c[d], reg], axis=1)
            for a, b, d in ((1, 2, 6), (1, 6, 5), (2, 3, 6), (3, 7, 6), (5, 6, 4), (6, 7, 4))]
    T = np.concatenate(tets)
    F = np.tile(np.array([[1.0, 0.0, 0.0]]), (T.shape[0], 1))
    return P, {'Tetras': T.astype(np.int32)}, F


def save_pkl(fname, P, E, F):
    with open(fname, 'wb') as f:
        pickle.dump({'Pts': P, 'Elems': E, 'Fibres': F}, f)


# corner offsets of the hexahedral cell in tetrahedral_box's corner order c[0..7], and its 6 tetrahedra
_CELL_CORNERS = np.array([(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)])
_CELL_TETS = ((0, 1, 2, 6), (0, 1, 6, 5), (0, 2, 3, 6), (0, 3, 7, 6), (0, 5, 6, 4), (0, 6, 7, 4))


def _cell_matrices(h, sigma):
    """P1 stiffness (8x8), lumped mass (8) and consistent mass (8x8) of ONE hexahedral cell of
    tetrahedral_box with edge lengths h = (hx, hy, hz) and isotropic conductivity sigma (float64)."""
    X = _CELL_CORNERS * np.asarray(h, dtype=np.float64)
    K = np.zeros((8, 8))
    M = np.zeros((8, 8))
    m = np.zeros(8)
    for tet in _CELL_TETS:
        P = X[list(tet)]
        A = np.hstack([np.ones((4, 1)), P])
        vol = abs(np.linalg.det(A)) / 6.0
        G = np.linalg.inv(A)[1:, :]                 # gradients of the four hat functions (3 x 4)
        Kt = vol * sigma * (G.T @ G)
        for a in range(4):
            m[tet[a]] += vol / 4.0                  # row sum of the consistent P1 mass matrix
            for b in range(4):
                K[tet[a], tet[b]] += Kt[a, b]
                M[tet[a], tet[b]] += vol / 20.0 * (2.0 if a == b else 1.0)   # P1 tetrahedron mass matrix
    return K, m, M


def _pair_in_cell_tet(a, b):
    return any(a in t and b in t for t in _CELL_TETS)


def structured_tet_operator(nx, ny, nz, h=(1.0, 1.0, 1.0), sigma=1.0, row_lo=0, row_hi=None, consistent_mass=False):
    """Rows [row_lo, row_hi) of the P1 stiffness matrix K (CSR, global column indices, fp32) and of the
    lumped mass M_L of tetrahedral_box(nx, ny, nz) WITHOUT materialising the mesh: the mesh is a
    translation of one 6-tetrahedron cell, so K(row, row + d) = sum over the cell corners (a, b) with
    offset(b) - offset(a) = d of Kcell[a, b] * [the cell based at node - offset(a) exists].
    This is what lets config 5 (272^3 = 2.0e7 nodes, 1.2e8 tetrahedra) be set up per rank in seconds;
    tests/test_fem_partition_cpu.py checks it against the general element-by-element assembly.
    Returns rowptr int32 [n+1], col int32 [nnz], kval fp32 [nnz], mlump fp32 [n] (n = row_hi - row_lo);
    with consistent_mass=True also mval fp32 [nnz], the consistent P1 mass matrix on the same pattern
    (the MASS of the semi-implicit step, fem_dist.PartitionedSemiImplicitFEM)."""
    n_global = nx * ny * nz
    row_hi = n_global if row_hi is None else row_hi
    Kc, mc, Mc = _cell_matrices(h, sigma)
    rows = np.arange(row_lo, row_hi, dtype=np.int64)
    i, rem = np.divmod(rows, ny * nz)
    j, k = np.divmod(rem, nz)

    def cell_exists(off):       # is the cell whose corner `off` is this node inside the grid?
        bi, bj, bk = i - off[0], j - off[1], k - off[2]
        return (bi >= 0) & (bi <= nx - 2) & (bj >= 0) & (bj <= ny - 2) & (bk >= 0) & (bk <= nz - 2)

    valid = [cell_exists(_CELL_CORNERS[a]) for a in range(8)]
    mlump = np.zeros(rows.shape[0])
    for a in range(8):
        mlump += mc[a] * valid[a]
    # group the 64 corner pairs by node offset d = off(b) - off(a)
    # every pair of corners that shares a tetrahedron is a structural entry, also where the stiffness
    # sums to zero (the reference's assembled pattern keeps such explicit zeros: nnz = N + 2*edges)
    connected = {(t[a], t[b]) for t in _CELL_TETS for a in range(4) for b in range(4)}
    groups = {}
    for a, b in sorted(connected):
        d = tuple(int(v) for v in (_CELL_CORNERS[b] - _CELL_CORNERS[a]))
        groups.setdefault(d, []).append((a, b))
    offs = sorted(groups, key=lambda d: (d[0] * ny + d[1]) * nz + d[2])
    vals = np.zeros((rows.shape[0], len(offs)))
    mvals = np.zeros((rows.shape[0], len(offs))) if consistent_mass else None
    mask = np.zeros((rows.shape[0], len(offs)), dtype=bool)
    cols = np.zeros((rows.shape[0], len(offs)), dtype=np.int64)
    for c, d in enumerate(offs):
        cols[:, c] = rows + (d[0] * ny + d[1]) * nz + d[2]
        for a, b in groups[d]:
            vals[:, c] += Kc[a, b] * valid[a]
            if consistent_mass:
                mvals[:, c] += Mc[a, b] * valid[a]
            mask[:, c] |= valid[a] & _pair_in_cell_tet(a, b)
    rowptr = np.zeros(rows.shape[0] + 1, dtype=np.int64)
    np.cumsum(mask.sum(axis=1), out=rowptr[1:])
    out = (rowptr.astype(np.int32), cols[mask].astype(np.int32), vals[mask].astype(np.float32),
           mlump.astype(np.float32))
    return out + (mvals[mask].astype(np.float32),) if consistent_mass else out
