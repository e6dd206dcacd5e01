"""Host-side logic of the slab decomposition, world_size 2 and 3 over gloo on CPU.

Each rank advances its slab with the ORACLE step applied to the local array (ghost planes
included) using the partition's clamp range - exactly what the CUDA kernel does with
x_begin/x_end/x_clamp - and exchanges one plane of U per neighbour per step through
pdensorflow_b200.fd.exchange_halos.  The gathered result must equal the single-domain oracle
bit for bit.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fd as ofd
from oracle import ionic as oionic
from pdensorflow_b200.fd import SlabPartition, exchange_halos


def test_partition_covers_the_axis():
    for NX, world in [(12, 1), (12, 2), (13, 3), (1024, 8), (25, 8)]:
        parts = [SlabPartition(NX, r, world) for r in range(world)]
        assert parts[0].x0 == 0 and parts[-1].x1 == NX
        for a, b in zip(parts[:-1], parts[1:]):
            assert a.x1 == b.x0
        assert abs(max(p.nloc for p in parts) - min(p.nloc for p in parts)) <= 1
        # clamp range is the image of global [1, NX-2]
        for p in parts:
            g_lo = p.x_clamp_lo + p.x0 - 1
            g_hi = p.x_clamp_hi + p.x0 - 1
            assert g_lo == max(1, p.x0 - 1) and g_hi == min(NX - 2, p.x1)
    with pytest.raises(ValueError):
        SlabPartition(5, 0, 2)


def _local_oracle_step(model, U, part, dt, diff, spacing):
    """Oracle step of the owned planes of a local array whose plane indices are clamped to
    [x_clamp_lo, x_clamp_hi] (y,z clamped to [1, n-2] as usual)."""
    nxa = U.shape[0]
    ii = np.clip(np.arange(nxa), part.x_clamp_lo, part.x_clamp_hi)
    jj = np.clip(np.arange(U.shape[1]), 1, U.shape[1] - 2)
    kk = np.clip(np.arange(U.shape[2]), 1, U.shape[2] - 2)
    U0 = U[np.ix_(ii, jj, kk)]
    # symmetric pad of U0 == clamp of the neighbour index to the array, then the same clamps
    im = np.clip(np.arange(nxa) - 1, 0, nxa - 1)
    ip = np.clip(np.arange(nxa) + 1, 0, nxa - 1)
    F = np.float32
    c = U0
    two = F(2.0)
    pad = np.pad(U0, ((0, 0), (1, 1), (1, 1)), mode='symmetric')
    lap = ((U0[im] - two * c + U0[ip]) / (F(spacing[0]) * F(spacing[0]))
           + (pad[:, 0:-2, 1:-1] - two * c + pad[:, 2:, 1:-1]) / (F(spacing[1]) * F(spacing[1]))
           + (pad[:, 1:-1, 0:-2] - two * c + pad[:, 1:-1, 2:]) / (F(spacing[2]) * F(spacing[2])))
    dU = model.differentiate(U)
    U1 = U0 + F(dt) * dU + F(diff * dt) * lap
    out = U.copy()
    out[part.x_begin:part.x_end] = U1[part.x_begin:part.x_end]
    return out.astype(np.float32)


def _worker(rank, world, port, NX, nsteps, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        shape = (NX, 7, 6)
        dt, diff, spacing = 0.1, 0.8, (1.0, 1.0, 1.0)
        rng = np.random.default_rng(0)
        G = rng.uniform(-85, 25, size=shape).astype(np.float32)
        part = SlabPartition(NX, rank, world)
        U = part.scatter(G)
        m = oionic.Fenton4v()
        m._dt = dt
        m.initialize_state_variables(U)
        for _ in range(nsteps):
            U = _local_oracle_step(m, U, part, dt, diff, spacing)
            t = torch.from_numpy(U)
            exchange_halos(t, part)
            U = t.numpy()
        q.put((rank, part.x0, part.owned(U).copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,NX', [(2, 11), (3, 13)])
def test_slab_exchange_matches_single_domain(world, NX):
    nsteps = 4
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, NX, nsteps, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shape = (NX, 7, 6)
    rng = np.random.default_rng(0)
    G = rng.uniform(-85, 25, size=shape).astype(np.float32)
    m = oionic.Fenton4v()
    m._dt = 0.1
    m.initialize_state_variables(G)
    ref = ofd.fd_run(m, G, nsteps, 0.1, 0.8, (1.0, 1.0, 1.0))
    out = np.zeros_like(ref)
    for rank, x0, owned in got:
        out[x0:x0 + owned.shape[0]] = owned
    assert np.array_equal(out, ref)


# ------------------------------------------------------------------------------------------------
# Row-block partitioned PCG (pdx_pcg_part_solve): the kernel's protocol restated in NumPy over gloo
# ------------------------------------------------------------------------------------------------
def _pcg_part_worker(rank, world, port, dims, q):
    """What pcg_part_kernel does, with gloo standing in for NVLink peer memory: windows [ghost_lo | owned | ghost_hi],
    ghost x and p RECOMPUTED locally from the neighbours' z = Minv*r of the edge rows, dot products summed in rank
    order (all_gather), two exchanges per iteration, convergence test before the p update."""
    from pdensorflow_b200.fem_dist import plan_windows, row_blocks
    from pdensorflow_b200.gpuSolve.entities.synthetic import structured_tet_operator
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        F = np.float32
        nx, ny, nz = dims
        n = nx * ny * nz
        blocks = row_blocks(n, world)
        lo, hi = blocks[rank]
        rp, ci, kv, ml, mv = structured_tet_operator(nx, ny, nz, (0.25, 0.25, 0.25), 1.0, lo, hi, consistent_mass=True)
        mine = torch.tensor([int(ci.min()), int(ci.max())])
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        plan = plan_windows(blocks, [tuple(int(v) for v in t) for t in allr])[rank]
        glo, ghi, nloc, W = plan['ghost_lo'], plan['ghost_hi'], hi - lo, plan['window']
        from scipy.sparse import csr_matrix
        A = csr_matrix(((mv + F(0.1) * kv).astype(F), ci.astype(np.int64) - (lo - glo), rp), shape=(nloc, W))
        rows = np.repeat(np.arange(lo, hi), np.diff(rp))
        minv = np.zeros(nloc, F)
        minv[rows[rows == ci] - lo] = F(1.0) / A.data[rows == ci]
        rng = np.random.default_rng(7)
        bg = rng.standard_normal(n).astype(F)
        b = bg[lo:hi]
        own = slice(glo, glo + nloc)
        x, p, zg = np.zeros(W, F), np.zeros(W, F), np.zeros(W, F)

        def push_edges(z, dst):
            """z of my first send_lo / last send_hi rows -> the neighbours' ghost slots of `dst`."""
            ops, bufs = [], {}
            if plan['send_lo']:
                ops.append(dist.P2POp(dist.isend, torch.from_numpy(z[:plan['send_lo']].copy()), rank - 1))
            if plan['send_hi']:
                ops.append(dist.P2POp(dist.isend, torch.from_numpy(z[nloc - plan['send_hi']:].copy()), rank + 1))
            if glo:
                bufs['lo'] = torch.zeros(glo)
                ops.append(dist.P2POp(dist.irecv, bufs['lo'], rank - 1))
            if ghi:
                bufs['hi'] = torch.zeros(ghi)
                ops.append(dist.P2POp(dist.irecv, bufs['hi'], rank + 1))
            for r_ in (dist.batch_isend_irecv(ops) if ops else []):
                r_.wait()
            if glo:
                dst[:glo] = bufs['lo'].numpy()
            if ghi:
                dst[glo + nloc:] = bufs['hi'].numpy()

        def all_sum(*vals):
            t = torch.tensor([float(v) for v in vals], dtype=torch.float32)
            got = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(got, t)
            out = np.zeros(len(vals), F)
            for g in got:                                  # rank order: same bits on every rank
                out = (out + g.numpy()).astype(F)
            return out

        r = (b - A @ x).astype(F)
        z = (minv * r).astype(F)
        p[own] = z
        push_edges(z, p)
        rz, res = all_sum(np.sum(r * z, dtype=F), np.sum(r * r, dtype=F))
        its = 0
        for k in range(1, 201):
            its = k
            qv = (A @ p).astype(F)
            pq, = all_sum(np.sum(p[own] * qv, dtype=F))
            alpha = F(rz / pq)
            x = (x + alpha * p).astype(F)                  # owned rows AND ghosts
            r = (r - alpha * qv).astype(F)
            z = (minv * r).astype(F)
            push_edges(z, zg)
            rznew, res = all_sum(np.sum(r * z, dtype=F), np.sum(r * r, dtype=F))
            if k % 5 == 0 and (not np.isfinite(res) or float(res) < 1e-12):
                break
            beta = F(rznew / rz)
            zw = zg.copy()
            zw[own] = z
            p = (zw + beta * p).astype(F)
            rz = rznew
        q.put((rank, lo, glo, ghi, x.copy(), its, bg))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_partitioned_pcg_protocol_over_gloo(world):
    from oracle import fem as ofem
    from pdensorflow_b200.gpuSolve.entities.synthetic import structured_tet_operator
    dims = (9, 4, 5)
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_pcg_part_worker, args=(r, world, port, dims, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(world)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rp, ci, kv, ml, mv = structured_tet_operator(*dims, (0.25, 0.25, 0.25), 1.0, consistent_mass=True)
    n = len(ml)
    val = ofem.csr_axpby(mv, 1.0, kv, 0.1)
    bg = got[0][6]
    xo, nit, _ = ofem.pcg(rp, ci, val, ofem.jacobi(rp, ci, val, n), bg, np.zeros(n, np.float32), 200, 1e-6)
    assert len({g[5] for g in got}) == 1 and abs(got[0][5] - nit) <= 5       # same decision on every rank
    full = np.zeros(n, np.float32)
    for rank, lo, glo, ghi, x, its, _ in got:
        full[lo:lo + len(x) - glo - ghi] = x[glo:len(x) - ghi]
    assert np.linalg.norm(full - xo) <= 1e-5 * np.linalg.norm(xo)
    for (_, lo, glo, ghi, x, _, _) in got:                                    # ghosts were never exchanged, yet equal their owners bitwise
        assert np.array_equal(x[:glo], full[lo - glo:lo])
        assert np.array_equal(x[len(x) - ghi:], full[lo + len(x) - glo - ghi:lo + len(x) - glo])
