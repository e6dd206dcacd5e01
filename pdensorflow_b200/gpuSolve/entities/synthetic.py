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
