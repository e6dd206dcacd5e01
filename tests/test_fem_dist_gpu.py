"""Row-block partitioned explicit lumped FEM step (config 5 path): one rank against the oracle,
two/three ranks (NVLink symmetric memory, ghost stores fused into the step kernel) bitwise against one rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from helpers import rel_l2

pytestmark = pytest.mark.gpu
DIMS, H, SIGMA, DT = (13, 6, 8), (0.25, 0.25, 0.25), 1e-3, 0.1


def _initial(n, dims):
    U = np.full(n, -80.0, dtype=np.float32)
    U[: 2 * dims[1] * dims[2]] = 20.0              # S1: the first two node planes
    return U


def test_single_rank_matches_oracle(cuda):
    from oracle import fem as ofem
    from oracle import ionic as oi
    from pdensorflow_b200.fem_dist import PartitionedExplicitFEM
    from pdensorflow_b200.gpuSolve.entities.synthetic import structured_tet_operator
    nx, ny, nz = DIMS
    n = nx * ny * nz
    rp, ci, kv, ml = structured_tet_operator(nx, ny, nz, H, SIGMA)
    for math, tol, layout in (('exact', 1e-6, 'sell'), ('fast', 1e-5, 'sell'), ('exact', 1e-6, 'csr'), ('fast', 1e-5, 'csr')):
        fem = PartitionedExplicitFEM(n, rp, ci, kv, ml, DT, model='fenton4v', math=math, layout=layout)
        U0 = _initial(n, DIMS)
        fem.set_global_U(U0)
        fem.step(50)
        mo = oi.Fenton4v()
        mo._dt = DT
        mo.initialize_state_variables(U0[:, None])
        mlinv = (1.0 / ml.astype(np.float64)).astype(np.float32)
        Uo = U0.copy()
        for _ in range(50):
            Uo = ofem.explicit_lumped_step(mo, rp, ci, kv, mlinv, Uo, None, DT)
        got = fem.current_U().cpu().numpy()
        assert np.isfinite(got).all() and got.max() > 0.0
        assert rel_l2(got, Uo) <= tol, (math, rel_l2(got, Uo))


def fem_gather(fem, dist, world):
    """The global potential assembled from every rank's owned rows (host array)."""
    torch.cuda.synchronize()
    parts = [None] * world
    dist.all_gather_object(parts, (fem.lo, fem.owned_U().cpu().numpy()))
    return np.concatenate([p for _, p in sorted(parts, key=lambda t: t[0])])


def _worker(rank, world, port, nsteps, q):
    import torch.distributed as dist
    from pdensorflow_b200.fem_dist import PartitionedExplicitFEM, row_blocks
    from pdensorflow_b200.gpuSolve.entities.synthetic import structured_tet_operator
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        nx, ny, nz = DIMS
        n = nx * ny * nz
        lo, hi = row_blocks(n, world)[rank]
        rp, ci, kv, ml = structured_tet_operator(nx, ny, nz, H, SIGMA, lo, hi)
        fem = PartitionedExplicitFEM(n, rp, ci, kv, ml, DT, model='fenton4v', math='exact', rank=rank, world=world,
                                     sync='flags' if rank >= 0 and world % 2 == 0 else 'barrier')
        assert (fem.send_lo > 0) == (rank > 0) and (fem.send_hi > 0) == (rank < world - 1)
        assert fem.sync == ('flags' if world % 2 == 0 else 'barrier')     # both orderings are exercised (2 / 3 ranks)
        fem.set_global_U(_initial(n, DIMS))
        fem.step(nsteps // 2)
        fem.set_global_U(fem_gather(fem, dist, world))      # re-install mid-run: flags keep counting across it
        fem.run(nsteps - nsteps // 2, block=6)              # graph replays of 6 steps + a remainder of 2
        torch.cuda.synchronize()
        q.put((rank, lo, fem.owned_U().cpu().numpy(), fem.states[0].cpu().numpy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [1, 3])
def test_copy_engine_staged_explicit_step_is_bitwise_the_plain_one(cuda, world, monkeypatch):
    """The persistent kernel that streams the SELL slices through shared memory (the form blocks of 10^6 rows and more
    run; forced here) against the one-thread-per-row kernel: same accumulation order per row, bitwise equal - alone and
    with three row blocks sharing one GPU (ghost stores + neighbour flags from the persistent grid)."""
    from pdensorflow_b200.fem_dist import PartitionedExplicitFEM, row_blocks
    from pdensorflow_b200.gpuSolve.entities.synthetic import structured_tet_operator
    dims = (31, 17, 19)
    n = dims[0] * dims[1] * dims[2]
    out = {}
    for stage in ('0', '1'):
        monkeypatch.setenv('PDX_EXPL_STAGE', stage)
        ranks = []
        for r, (lo, hi) in enumerate(row_blocks(n, world)):
            rp, ci, kv, ml = structured_tet_operator(*dims, H, SIGMA, lo, hi)
            ranks.append(PartitionedExplicitFEM(n, rp, ci, kv, ml, DT, model='fenton4v', math='exact', rank=r, world=world,
                                                connect=world == 1, sync='flags' if world > 1 else 'auto'))
        if world > 1:
            PartitionedExplicitFEM.connect_local(ranks)
        for f in ranks:
            f.set_global_U(_initial(n, dims))
        PartitionedExplicitFEM.step_local(ranks, 30)
        torch.cuda.synchronize()
        out[stage] = np.concatenate([f.owned_U().cpu().numpy() for f in ranks])
    assert np.array_equal(out['0'], out['1']) and out['0'].max() > 0.0


def test_graph_replayed_blocks_equal_single_steps(cuda):
    from pdensorflow_b200.fem_dist import PartitionedExplicitFEM
    from pdensorflow_b200.gpuSolve.entities.synthetic import structured_tet_operator
    nx, ny, nz = DIMS
    n = nx * ny * nz
    rp, ci, kv, ml = structured_tet_operator(nx, ny, nz, H, SIGMA)
    out = []
    for how in ('step', 'run'):
        fem = PartitionedExplicitFEM(n, rp, ci, kv, ml, DT, model='fenton4v', math='exact')
        fem.set_global_U(_initial(n, DIMS))
        if how == 'step':
            fem.step(27)
        else:
            fem.run(27, block=6)
        torch.cuda.synchronize()
        assert fem.nsteps_done == 27
        out.append((fem.current_U().cpu().numpy(), fem.states[2].cpu().numpy()))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


@pytest.mark.parametrize('world', [2, 3])
def test_partitioned_ranks_match_single_rank_bitwise(cuda, world):
    if torch.cuda.device_count() < world:
        pytest.skip('needs %d GPUs' % world)
    from pdensorflow_b200.fem_dist import PartitionedExplicitFEM
    from pdensorflow_b200.gpuSolve.entities.synthetic import structured_tet_operator
    nsteps = 40
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nsteps, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        got = [q.get(timeout=90) for _ in range(world)]
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:
        for p in procs:                 # never leave a rank behind on the GPU
            if p.is_alive():
                p.kill()
    nx, ny, nz = DIMS
    n = nx * ny * nz
    rp, ci, kv, ml = structured_tet_operator(nx, ny, nz, H, SIGMA)
    ref = PartitionedExplicitFEM(n, rp, ci, kv, ml, DT, model='fenton4v', math='exact')
    ref.set_global_U(_initial(n, DIMS))
    ref.step(nsteps)
    Uref, Vref = ref.current_U().cpu().numpy(), ref.states[0].cpu().numpy()
    for rank, lo, U, V in got:
        assert np.array_equal(U, Uref[lo:lo + U.shape[0]])
        assert np.array_equal(V, Vref[lo:lo + V.shape[0]])


@pytest.mark.parametrize('sync', ['flags', 'barrier'])
@pytest.mark.parametrize('world,layout', [(2, 'sell'), (3, 'sell'), (3, 'csr')])
def test_ranks_sharing_one_gpu_match_single_rank_bitwise(cuda, world, layout, sync):
    """The ghost stores fused into the explicit step kernel on a ONE-GPU box: the row blocks live in one process and
    store their edge rows into each other's buffers through plain device pointers (PartitionedExplicitFEM.connect_local)."""
    from pdensorflow_b200.fem_dist import PartitionedExplicitFEM, row_blocks
    from pdensorflow_b200.gpuSolve.entities.synthetic import structured_tet_operator
    nx, ny, nz = DIMS
    n = nx * ny * nz
    nsteps = 40
    ranks = []
    for r, (lo, hi) in enumerate(row_blocks(n, world)):
        rp, ci, kv, ml = structured_tet_operator(nx, ny, nz, H, SIGMA, lo, hi)
        ranks.append(PartitionedExplicitFEM(n, rp, ci, kv, ml, DT, model='fenton4v', math='exact', rank=r, world=world,
                                            layout=layout, connect=False, sync=sync))
    PartitionedExplicitFEM.connect_local(ranks)
    for f in ranks:
        assert f.sync == sync
        assert (f.send_lo > 0) == (f.rank > 0) and (f.send_hi > 0) == (f.rank < world - 1)
        f.set_global_U(_initial(n, DIMS))
    PartitionedExplicitFEM.step_local(ranks, nsteps)
    torch.cuda.synchronize()
    rp, ci, kv, ml = structured_tet_operator(nx, ny, nz, H, SIGMA)
    ref = PartitionedExplicitFEM(n, rp, ci, kv, ml, DT, model='fenton4v', math='exact', layout=layout)
    ref.set_global_U(_initial(n, DIMS))
    ref.step(nsteps)
    Uref, Vref = ref.current_U().cpu().numpy(), ref.states[0].cpu().numpy()
    assert Uref.max() > 0.0
    for f in ranks:
        assert np.array_equal(f.owned_U().cpu().numpy(), Uref[f.lo:f.hi])
        assert np.array_equal(f.states[0].cpu().numpy(), Vref[f.lo:f.hi])


# ------------------------------------------------------------------------------ N3: unstructured meshes
def _unstructured_operator(npts=2500, seed=6):
    """Stiffness + lumped mass of a Delaunay tetrahedral mesh in RCM order, conductivity scaled so that the explicit
    step is stable (dt * Gershgorin bound = 0.5)."""
    from oracle import fem as ofem
    from test_fem_partition_cpu import unstructured_tet_mesh
    P, E, F = unstructured_tet_mesh(npts, seed)
    n = P.shape[0]
    ren = ofem.rcm_indexing(ofem.coo_pattern(ofem.connectivity(n, E)))
    m = ofem.assemble(P, E, F, {1: 1.0}, {1: 1.0}, ren['iperm'])
    rp, ci = m['rowptr'], m['col']
    ml = np.add.reduceat(m['mass'].astype(np.float64), rp[:-1])
    bound = (np.add.reduceat(np.abs(m['stiffness']).astype(np.float64), rp[:-1]) / ml).max()
    kv = (m['stiffness'] * np.float32(0.5 / (DT * bound))).astype(np.float32)
    # S1 region + a rough sub-threshold field everywhere, so that every row's K.U term matters in the bits
    U0 = np.where(P[ren['perm'], 0] < 0.2, 20.0, -80.0 + 25.0 * np.random.default_rng(seed).uniform(size=n)).astype(np.float32)
    return n, rp, ci, kv, ml.astype(np.float32), U0


def test_device_sell_conversion_equals_host_conversion(cuda):
    from pdensorflow_b200.fem_dist import csr_to_sell32, csr_to_sell32_device
    n, rp, ci, kv, ml, _ = _unstructured_operator(1100)
    for lo, hi in ((0, n), (137, 1001), (500, 532), (40, 41)):
        rpl = (rp[lo:hi + 1] - rp[lo]).astype(np.int32)
        cl, vl = ci[rp[lo]:rp[hi]], kv[rp[lo]:rp[hi]]
        sp, sc, sv = csr_to_sell32(rpl, cl, vl, lo)
        dsp, dsc, dsv = csr_to_sell32_device(torch.from_numpy(rpl).cuda(), torch.from_numpy(cl.astype(np.int32)).cuda(),
                                             torch.from_numpy(vl).cuda(), lo)
        assert np.array_equal(dsp.cpu().numpy(), sp) and np.array_equal(dsc.cpu().numpy(), sc)
        assert np.array_equal(dsv.cpu().numpy(), sv)


@pytest.mark.parametrize('world,layout', [(11, 'sell'), (6, 'csr')])
def test_general_ghost_lists_on_an_unstructured_mesh_bitwise(cuda, world, layout):
    """Row blocks of an unstructured (Delaunay) tetrahedral mesh: the RCM band is wider than a block, so ghosts reach
    past the two neighbours and travel through index lists + pdx_push_rows.  Ranks share one GPU (plain device pointers,
    one stream); every rank's rows must equal the single-rank run bitwise."""
    from pdensorflow_b200.fem_dist import PartitionedExplicitFEM, row_blocks
    n, rp, ci, kv, ml, U0 = _unstructured_operator()
    nsteps = 30
    ranks = []
    for r, (lo, hi) in enumerate(row_blocks(n, world)):
        sl = slice(rp[lo], rp[hi])
        ranks.append(PartitionedExplicitFEM(n, rp[lo:hi + 1] - rp[lo], ci[sl], kv[sl], ml[lo:hi], DT, model='fenton4v',
                                            math='exact', rank=r, world=world, layout=layout, connect=False))
    PartitionedExplicitFEM.connect_local(ranks)
    assert all(f.exchange == 'lists' for f in ranks)
    assert max(len(set(f._push[1].cpu().tolist())) for f in ranks) > 2          # some rank feeds more than two others
    for f in ranks:
        f.set_global_U(U0)
    PartitionedExplicitFEM.step_local(ranks, nsteps)
    torch.cuda.synchronize()
    ref = PartitionedExplicitFEM(n, rp, ci, kv, ml, DT, model='fenton4v', math='exact', layout=layout)
    ref.set_global_U(U0)
    ref.step(nsteps)
    Uref, Vref = ref.current_U().cpu().numpy(), ref.states[0].cpu().numpy()
    assert np.isfinite(Uref).all() and np.unique(Uref).shape[0] > 0.9 * n
    for f in ranks:
        assert np.array_equal(f.owned_U().cpu().numpy(), Uref[f.lo:f.hi]), f.rank
        assert np.array_equal(f.states[0].cpu().numpy(), Vref[f.lo:f.hi])
    # forcing the band plan on this partition is refused
    with pytest.raises(ValueError):
        bad = [PartitionedExplicitFEM(n, rp[lo:hi + 1] - rp[lo], ci[rp[lo]:rp[hi]], kv[rp[lo]:rp[hi]], ml[lo:hi], DT,
                                      rank=r, world=world, layout=layout, connect=False, exchange='band')
               for r, (lo, hi) in enumerate(row_blocks(n, world))]
        PartitionedExplicitFEM.connect_local(bad)
