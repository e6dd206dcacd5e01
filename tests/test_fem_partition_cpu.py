"""Host logic of the row-block partitioned FEM path (no GPU): the stencil-based generator of the
structured tetrahedral operator against the general element-by-element assembly (oracle, which is
itself pinned to the reference's NumPy assembly), the balanced row split, and the ghost-exchange plan."""
import numpy as np
import pytest

from oracle import fem as ofem
from pdensorflow_b200.fem_dist import plan_exchange, row_blocks
from pdensorflow_b200.gpuSolve.entities.synthetic import structured_tet_operator, tetrahedral_box


@pytest.mark.parametrize('dims,h,sigma', [((5, 4, 6), (0.5, 0.25, 1.0), 0.7), ((3, 3, 3), (1.0, 1.0, 1.0), 1.0),
                                          ((2, 7, 4), (0.1, 0.2, 0.3), 1e-3)])
def test_structured_operator_equals_assembled_matrices(dims, h, sigma):
    nx, ny, nz = dims
    P, E, F = tetrahedral_box(nx, ny, nz, Lx=(nx - 1) * h[0], Ly=(ny - 1) * h[1], Lz=(nz - 1) * h[2])
    m = ofem.assemble(P, E, F, {1: sigma}, {1: sigma})
    rp, ci, kv, ml = structured_tet_operator(nx, ny, nz, h, sigma)
    assert np.array_equal(rp, m['rowptr']) and np.array_equal(ci, m['col'])      # same pattern, structural zeros included
    np.testing.assert_allclose(kv, m['stiffness'], rtol=0, atol=2e-6 * np.abs(m['stiffness']).max())
    n = nx * ny * nz
    mass_rowsum = np.add.reduceat(m['mass'].astype(np.float64), m['rowptr'][:-1])
    np.testing.assert_allclose(ml, mass_rowsum, rtol=1e-6)
    np.testing.assert_allclose(ml.sum(dtype=np.float64), np.prod([(d - 1) * s for d, s in zip(dims, h)]), rtol=1e-6)
    # K annihilates constants
    rows = np.repeat(np.arange(n), np.diff(rp))
    assert np.abs(np.bincount(rows, weights=kv, minlength=n)).max() < 1e-5 * np.abs(kv).max()
    # a row block is a slice of the full operator
    lo, hi = n // 3, 2 * n // 3 + 1
    rp2, ci2, kv2, ml2 = structured_tet_operator(nx, ny, nz, h, sigma, lo, hi)
    assert np.array_equal(ci2, ci[rp[lo]:rp[hi]]) and np.array_equal(kv2, kv[rp[lo]:rp[hi]])
    assert np.array_equal(rp2, rp[lo:hi + 1] - rp[lo]) and np.array_equal(ml2, ml[lo:hi])


def test_row_blocks_are_balanced_and_cover():
    for n, w in ((10, 3), (63001, 8), (7, 7), (20123648, 8)):
        b = row_blocks(n, w)
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1


def test_exchange_plan_on_a_banded_operator():
    nx, ny, nz, world = 9, 4, 5, 3
    n = nx * ny * nz
    blocks = row_blocks(n, world)
    ranges = []
    for lo, hi in blocks:
        _, ci, _, _ = structured_tet_operator(nx, ny, nz, row_lo=lo, row_hi=hi)
        ranges.append((int(ci.min()), int(ci.max())))
    band = ny * nz + nz + 1
    for r in range(world):
        send_lo, send_hi = plan_exchange(blocks, ranges, r)
        assert (0 < send_lo <= band if r > 0 else send_lo == 0) and (0 < send_hi <= band if r < world - 1 else send_hi == 0)
    # every ghost column of every rank is covered by what its neighbours send
    for r, (lo, hi) in enumerate(blocks):
        _, ci, _, _ = structured_tet_operator(nx, ny, nz, row_lo=lo, row_hi=hi)
        ghosts = np.unique(ci[(ci < lo) | (ci >= hi)])
        for g in ghosts:
            owner = next(k for k, (a, b) in enumerate(blocks) if a <= g < b)
            assert abs(owner - r) == 1
            s_lo, s_hi = plan_exchange(blocks, ranges, owner)
            a, b = blocks[owner]
            assert (g < a + s_lo) if owner > r else (g >= b - s_hi)
    with pytest.raises(ValueError):                       # ghosts two blocks away: not banded enough
        plan_exchange(row_blocks(n, 9), [(max(lo - 3 * band, 0), min(hi + 3 * band, n) - 1) for lo, hi in row_blocks(n, 9)], 4)


def test_window_plan_of_the_semi_implicit_partition():
    from pdensorflow_b200.fem_dist import plan_windows
    nx, ny, nz, world = 9, 4, 5, 3
    n = nx * ny * nz
    blocks = row_blocks(n, world)
    ranges, cols = [], []
    for lo, hi in blocks:
        _, ci, _, _ = structured_tet_operator(nx, ny, nz, row_lo=lo, row_hi=hi)
        ranges.append((int(ci.min()), int(ci.max())))
        cols.append(ci)
    plans = plan_windows(blocks, ranges)
    for r, (p, (lo, hi)) in enumerate(zip(plans, blocks)):
        assert p['window'] == p['ghost_lo'] + (hi - lo) + p['ghost_hi']
        assert (p['ghost_lo'] == 0) == (r == 0) and (p['ghost_hi'] == 0) == (r == world - 1)
        local = cols[r].astype(np.int64) - (lo - p['ghost_lo'])         # window-local columns stay inside the window
        assert local.min() >= 0 and local.max() < p['window']
        if r > 0:                                                       # my row i sits at lower_window_index + i below
            q = plans[r - 1]
            assert p['send_lo'] == q['ghost_hi'] and p['lower_window_index'] == q['ghost_lo'] + q['n_local']
        if r < world - 1:                                               # my row n - send_hi + j sits at upper_window_index + j above
            q = plans[r + 1]
            assert p['send_hi'] == q['ghost_lo']
            g_first = hi - p['send_hi']                                 # global index of the first row sent up
            assert g_first - (q['lo'] - q['ghost_lo']) == p['upper_window_index']
    with pytest.raises(ValueError):
        plan_windows(row_blocks(n, 9), [(max(lo - 200, 0), min(hi + 200, n) - 1) for lo, hi in row_blocks(n, 9)])


def test_structured_consistent_mass_matches_assembly():
    nx, ny, nz, h = 4, 5, 3, (0.3, 0.2, 0.5)
    P, E, F = tetrahedral_box(nx, ny, nz, Lx=(nx - 1) * h[0], Ly=(ny - 1) * h[1], Lz=(nz - 1) * h[2])
    m = ofem.assemble(P, E, F, {1: 1.0}, {1: 1.0})
    rp, ci, kv, ml, mv = structured_tet_operator(nx, ny, nz, h, 1.0, consistent_mass=True)
    assert np.array_equal(ci, m['col'])
    np.testing.assert_allclose(mv, m['mass'], rtol=0, atol=2e-6 * np.abs(m['mass']).max())
    np.testing.assert_allclose(np.add.reduceat(mv.astype(np.float64), rp[:-1]), ml, rtol=1e-5)


@pytest.mark.parametrize('row_lo', [0, 7])
def test_sell32_conversion_is_a_faithful_relayout(row_lo):
    """csr_to_sell32: slice s holds rows 32s..32s+31 column-major (entry k of lane l at slice_ptr[s] + 32k + l), short
    rows padded with (col = the row's own index + row_lo, val = 0); y = A x must be unchanged and the padding small."""
    from pdensorflow_b200.fem_dist import csr_to_sell32
    rng = np.random.default_rng(2)
    n, ncols = 150, 400
    lens = rng.integers(1, 12, size=n)
    rowptr = np.zeros(n + 1, dtype=np.int32)
    rowptr[1:] = np.cumsum(lens)
    col = np.concatenate([np.sort(rng.choice(ncols, size=int(k), replace=False)) for k in lens]).astype(np.int32)
    val = rng.standard_normal(col.shape[0]).astype(np.float32)
    sp, scol, sval = csr_to_sell32(rowptr, col, val, row_lo)
    nsl = (n + 31) // 32
    assert sp.shape[0] == nsl + 1 and sp[0] == 0 and (np.diff(sp) % 32 == 0).all() and sp[-1] == scol.shape[0] == sval.shape[0]
    x = rng.standard_normal(ncols).astype(np.float64)
    y = np.zeros(n)
    seen = 0
    for s in range(nsl):
        width = (sp[s + 1] - sp[s]) // 32
        assert width == lens[32 * s:32 * s + 32].max()
        for lane in range(32):
            row = 32 * s + lane
            idx = sp[s] + np.arange(width) * 32 + lane
            if row < n:
                k = lens[row]
                assert np.array_equal(scol[idx[:k]], col[rowptr[row]:rowptr[row + 1]])          # row order kept
                assert np.array_equal(sval[idx[:k]], val[rowptr[row]:rowptr[row + 1]])
                assert (sval[idx[k:]] == 0).all() and (scol[idx[k:]] == row + row_lo).all()     # padding
                y[row] = (sval[idx].astype(np.float64) * x[scol[idx]]).sum()
                seen += k
            else:
                assert (sval[idx] == 0).all() and (scol[idx] == n - 1 + row_lo).all()
    rows = np.repeat(np.arange(n), lens)
    assert seen == col.shape[0]
    assert np.allclose(y, np.bincount(rows, weights=val.astype(np.float64) * x[col], minlength=n))


# ------------------------------------------------------------------------------ N3: general ghost lists, SELL-C-sigma order
def unstructured_tet_mesh(npts=700, seed=4):
    """A genuinely unstructured tetrahedral mesh: Delaunay tetrahedra of random points in the unit cube (slivers of
    negligible volume dropped).  Returns Pts [N,3], {'Tetras': [nE,5]}, fibres."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    P = rng.uniform(0.0, 1.0, size=(npts, 3))
    T = Delaunay(P).simplices.astype(np.int32)
    v = P[T[:, 1:]] - P[T[:, :1]]
    vol = np.abs(np.linalg.det(v)) / 6.0
    T = T[vol > 1e-7]
    assert np.unique(T).shape[0] == npts
    E = np.concatenate([T, np.ones((T.shape[0], 1), dtype=np.int32)], axis=1)
    return P, {'Tetras': E}, np.tile(np.array([[1.0, 0.0, 0.0]]), (E.shape[0], 1))


def test_ghost_lists_cover_every_off_block_column_of_an_unstructured_mesh():
    from pdensorflow_b200.fem_dist import ghost_columns, plan_exchange, plan_ghost_lists
    P, E, F = unstructured_tet_mesh()
    n = P.shape[0]
    ren = ofem.rcm_indexing(ofem.coo_pattern(ofem.connectivity(n, E)))
    m = ofem.assemble(P, E, F, {1: 1.0}, {1: 1.0}, ren['iperm'])
    world = 9
    blocks = row_blocks(n, world)
    cols = [m['col'][m['rowptr'][lo]:m['rowptr'][hi]] for lo, hi in blocks]
    ranges = [(int(c.min()), int(c.max())) for c in cols]
    with pytest.raises(ValueError):                       # 3-D RCM band > block: the two-neighbour plan refuses it
        plan_exchange(blocks, ranges, 4)
    ghosts = [ghost_columns(c, lo, hi) for c, (lo, hi) in zip(cols, blocks)]
    plans = plan_ghost_lists(blocks, ghosts)
    received = [set() for _ in range(world)]
    for r, (rows, peers) in enumerate(plans):
        lo, hi = blocks[r]
        assert rows.shape == peers.shape and np.all((rows >= lo) & (rows < hi)) and np.all(peers != r)
        assert len(set(zip(rows.tolist(), peers.tolist()))) == rows.shape[0]          # nothing is sent twice
        for row, peer in zip(rows, peers):
            received[peer].add(int(row))
    for r in range(world):
        assert received[r] == set(int(g) for g in ghosts[r])
    assert max(len(set(p[1].tolist())) for p in plans) > 2                            # someone talks to > 2 ranks


def test_sigma_window_order_is_partition_aware_and_reduces_padding():
    from pdensorflow_b200.fem_dist import csr_to_sell32, permute_csr, sell_padding, sigma_window_order
    P, E, F = unstructured_tet_mesh(1500, seed=9)
    n = P.shape[0]
    ren = ofem.rcm_indexing(ofem.coo_pattern(ofem.connectivity(n, E)))
    m = ofem.assemble(P, E, F, {1: 1.0}, {1: 1.0}, ren['iperm'])
    lens = np.diff(m['rowptr'])
    blocks = row_blocks(n, 4)
    perm = sigma_window_order(lens, sigma=128, blocks=blocks)
    assert np.array_equal(np.sort(perm), np.arange(n))
    for lo, hi in blocks:                                  # rows never leave their block, nor their window
        assert np.array_equal(np.sort(perm[lo:hi]), np.arange(lo, hi))
        for w0 in range(lo, hi, 128):
            w1 = min(w0 + 128, hi)
            assert np.array_equal(np.sort(perm[w0:w1]), np.arange(w0, w1))
            assert np.all(np.diff(lens[perm[w0:w1]]) <= 0)
    before, after = sell_padding(lens), sell_padding(lens, perm)
    assert after < 0.85 * before, (before, after)            # 1.52 -> 1.21 here; sigma = 512 reaches 1.08
    assert sell_padding(lens, sigma_window_order(lens, sigma=512, blocks=blocks)) < 1.10
    # P A P^T: same operator in the new numbering, and the host SELL conversion stores exactly `after` x nnz entries
    rp2, ci2, kv2 = permute_csr(m['rowptr'], m['col'], m['stiffness'], perm)
    x = np.random.default_rng(0).standard_normal(n).astype(np.float32)
    y = ofem.spmv(m['rowptr'], m['col'], m['stiffness'], x)
    y2 = ofem.spmv(rp2, ci2, kv2, x[perm])
    np.testing.assert_allclose(y2, y[perm], rtol=0, atol=1e-4 * np.abs(y).max())
    sp, scol, sval = csr_to_sell32(rp2, ci2, kv2)
    assert scol.shape[0] == round(after * m['col'].shape[0])
    with pytest.raises(ValueError):
        sigma_window_order(lens, sigma=100)
