"""Row-block partitioned SEMI-IMPLICIT FEM step (FEM): the persistent cooperative PCG
kernel whose ranks talk through peer memory (pdx_pcg_part_solve).  One rank against the single-GPU PCG and the
oracle; 2/3 ranks sharing ONE GPU (plain device pointers as peers, one stream per rank) and 2 ranks on 2 GPUs
(NVLink symmetric memory) against one rank; a missing peer is reported, not waited for."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from helpers import rel_l2

pytestmark = pytest.mark.gpu
DIMS, H, SIGMA, DT = (13, 6, 8), (0.25, 0.25, 0.25), 1e-3, 0.1
TOL_CG = 5e-5            # two correct CG implementations agree to a small multiple of the stopping tolerance


def _same_iterations(a, b):
    """Iteration counts of two correct implementations: equal, except that a residual sitting on the stopping
    threshold may tip the batched test (every 5 iterations) the other way on a few steps."""
    d = np.abs(np.asarray(a) - np.asarray(b))
    return len(a) == len(b) and d.max() <= 5 and (d > 0).sum() <= max(2, len(a) // 5)


def _operator(lo=0, hi=None):
    from pdensorflow_b200.gpuSolve.entities.synthetic import structured_tet_operator
    return structured_tet_operator(*DIMS, H, SIGMA, lo, hi, consistent_mass=True)


def _initial():
    n = DIMS[0] * DIMS[1] * DIMS[2]
    U = np.full(n, -80.0, dtype=np.float32)
    U[: 2 * DIMS[1] * DIMS[2]] = 20.0
    return U


def _make(rank=0, world=1, **kw):
    from pdensorflow_b200.fem_dist import PartitionedSemiImplicitFEM, row_blocks
    n = DIMS[0] * DIMS[1] * DIMS[2]
    lo, hi = row_blocks(n, world)[rank]
    rp, ci, kv, ml, mv = _operator(lo, hi)
    return PartitionedSemiImplicitFEM(n, rp, ci, mv, kv, DT, rank=rank, world=world, **kw)


def test_one_rank_solve_matches_single_gpu_pcg(cuda):
    from pdensorflow_b200 import _lib as L
    n = DIMS[0] * DIMS[1] * DIMS[2]
    fem = _make(model='none', maxiter=400, toll=1e-6, layout='csr')
    sell = _make(model='none', maxiter=400, toll=1e-6, layout='sell')
    rng = np.random.default_rng(3)
    xs = rng.standard_normal(n).astype(np.float32)
    b = torch.from_numpy(xs).cuda()
    # single-GPU persistent PCG on the same matrix
    h = C.c_void_p()
    L.check(L.lib().pdx_pcg_create(n, L.ptr(fem.rowptr), L.ptr(fem.col), L.ptr(fem.aval), L.ptr(fem.minv), C.byref(h)))
    x1 = torch.zeros(n, device='cuda')
    info = torch.zeros(4, dtype=torch.int32, device='cuda')
    L.check(L.lib().pdx_pcg_solve(h, L.ptr(b), L.ptr(x1), 400, 1e-6, 5, L.ptr(info), L.stream_ptr()))
    fem.U.zero_()
    fem.solve(b)
    torch.cuda.synchronize()
    it, res, flags = fem.last_solve()
    L.lib().pdx_pcg_destroy(h)
    assert flags == 0 and abs(it - int(info[0].item())) <= 5 and res < 1e-12
    assert rel_l2(fem.owned_U().cpu().numpy(), x1.cpu().numpy()) <= 1e-5
    sell.U.zero_()
    sell.solve(b)                                   # same system, SELL-32 storage
    assert abs(sell.last_solve()[0] - it) <= 5 and sell.padding < 1.5     # tiny mesh: boundary rows are short
    assert rel_l2(sell.owned_U().cpu().numpy(), x1.cpu().numpy()) <= 1e-5
    # and it solves the system
    import scipy.sparse as sp
    A = sp.csr_matrix((fem.aval.cpu().numpy().astype(np.float64), fem.col.cpu().numpy(), fem.rowptr.cpu().numpy()), shape=(n, n))
    assert np.linalg.norm(A @ fem.owned_U().cpu().numpy().astype(np.float64) - xs) <= 1e-5 * np.linalg.norm(xs)


def _oracle_steps(nsteps):
    from oracle import fem as ofem
    from oracle import ionic as oi
    rp, ci, kv, ml, mv = _operator()
    n = len(ml)
    A = ofem.csr_axpby(mv, 1.0, kv, DT)
    minv = ofem.jacobi(rp, ci, A, n)
    mo = oi.Fenton4v()
    mo._dt = DT
    U = _initial()
    mo.initialize_state_variables(U[:, None])
    its = []
    for _ in range(nsteps):
        rhs0 = U + np.float32(DT) * mo.differentiate(U[:, None])[:, 0]
        rhs = ofem.spmv(rp, ci, mv, rhs0)
        U, nit, _ = ofem.pcg(rp, ci, A, minv, rhs, U, 100, 1e-5)
        its.append(nit)
    return U, its


def test_copy_engine_staged_sell_product_matches_the_plain_one(cuda, monkeypatch):
    """q = A p with the SELL slices staged through shared memory by cp.async.bulk (rows_dot_staged, the form large systems
    run) against the per-thread loads.  The row products have the same accumulation order; the staged form runs other
    block sizes, so the block partials of the dot products are added in another order: same iteration decisions, iterates
    equal to the CG tolerance."""
    from pdensorflow_b200.fem_dist import PartitionedSemiImplicitFEM
    from pdensorflow_b200.gpuSolve.entities.synthetic import structured_tet_operator
    dims = (37, 29, 23)
    n = dims[0] * dims[1] * dims[2]
    rp, ci, kv, ml, mv = structured_tet_operator(*dims, H, SIGMA, 0, None, consistent_mass=True)
    U0 = np.full(n, -80.0, dtype=np.float32)
    U0[: 2 * dims[1] * dims[2]] = 20.0
    out = {}
    for stage in ('0', '1'):
        monkeypatch.setenv('PDX_PCG_STAGE', stage)
        fem = PartitionedSemiImplicitFEM(n, rp, ci, mv, kv, DT, math='exact', layout='sell')
        fem.set_global_U(U0)
        recs = []
        for _ in range(12):
            fem.step(1)
            recs.append(fem.last_solve())
        out[stage] = (fem.owned_U().cpu().numpy(), recs)
    assert rel_l2(out['1'][0], out['0'][0]) <= TOL_CG
    assert _same_iterations([r[0] for r in out['1'][1]], [r[0] for r in out['0'][1]])
    assert all(r[2] == 0 for r in out['1'][1])
    assert out['0'][0].max() > 0.0


@pytest.mark.parametrize('layout,stage', [('sell', '0'), ('sell', '1'), ('csr', '0')])
def test_one_rank_steps_match_oracle(cuda, layout, stage, monkeypatch):
    monkeypatch.setenv('PDX_PCG_STAGE', stage)
    nsteps = 25
    Uo, its = _oracle_steps(nsteps)
    fem = _make(math='exact', layout=layout)
    fem.set_global_U(_initial())
    got_its = []
    for _ in range(nsteps):
        fem.step(1)
        got_its.append(fem.last_solve()[0])
    got = fem.owned_U().cpu().numpy()
    assert np.isfinite(got).all() and got.max() > 0.0
    assert rel_l2(got, Uo) <= TOL_CG, rel_l2(got, Uo)
    assert _same_iterations(got_its, its), (got_its, its)


@pytest.mark.parametrize('world,layout,stage', [(2, 'sell', '0'), (3, 'sell', '0'), (3, 'sell', '1'), (2, 'csr', '0')])
def test_ranks_sharing_one_gpu_match_one_rank(cuda, world, layout, stage, monkeypatch):
    from pdensorflow_b200.fem_dist import PartitionedSemiImplicitFEM
    monkeypatch.setenv('PDX_PCG_STAGE', stage)
    nsteps = 20
    ref = _make(math='exact', layout=layout)
    ref.set_global_U(_initial())
    ref_its = []
    for _ in range(nsteps):
        ref.step(1)
        ref_its.append(ref.last_solve()[0])
    Uref, Vref = ref.owned_U().cpu().numpy(), ref.owned_state(0).cpu().numpy()
    ranks = [_make(r, world, math='exact', connect=False, max_blocks=16, timeout_ms=5000, stream=torch.cuda.Stream(),
                   layout=layout) for r in range(world)]
    PartitionedSemiImplicitFEM.connect_local(ranks)
    for f in ranks:
        assert (f.glo > 0) == (f.rank > 0) and (f.ghi > 0) == (f.rank < world - 1)
        f.set_global_U(_initial())
    its = [[] for _ in ranks]
    for _ in range(nsteps):
        for f in ranks:                  # every rank's two launches go to its own stream; the kernels meet on the device
            f.step(1)
        for k, f in enumerate(ranks):
            its[k].append(f.last_solve()[0])
    torch.cuda.synchronize()
    assert all(i == its[0] for i in its)                # same convergence decision on every rank
    assert _same_iterations(its[0], ref_its), (its[0], ref_its)
    for f in ranks:
        assert rel_l2(f.owned_U().cpu().numpy(), Uref[f.lo:f.hi]) <= TOL_CG
        assert np.array_equal(f.owned_state(0).cpu().numpy(), Vref[f.lo:f.hi]) or \
            rel_l2(f.owned_state(0).cpu().numpy(), Vref[f.lo:f.hi]) <= TOL_CG
    # ghosts are recomputed, never exchanged: they must equal the owner's values BITWISE
    for a, b in zip(ranks[:-1], ranks[1:]):
        assert np.array_equal(a.U[a.glo + a.nloc:].cpu().numpy(), b.owned_U()[:a.ghi].cpu().numpy())
        assert np.array_equal(b.U[:b.glo].cpu().numpy(), a.owned_U()[a.nloc - b.glo:].cpu().numpy())
        assert np.array_equal(a.states[0][a.glo + a.nloc:].cpu().numpy(), b.owned_state(0)[:a.ghi].cpu().numpy())


def test_missing_peer_is_reported_not_waited_for(cuda):
    from pdensorflow_b200 import _lib as L
    from pdensorflow_b200.fem_dist import PartitionedSemiImplicitFEM
    ranks = [_make(r, 2, connect=False, max_blocks=16, timeout_ms=200) for r in range(2)]
    PartitionedSemiImplicitFEM.connect_local(ranks)
    ranks[0].set_global_U(_initial())
    ranks[0].step(1)                                    # rank 1 never launches
    torch.cuda.synchronize()
    with pytest.raises(L.PdxError):
        ranks[0].last_solve()


def _worker(rank, world, port, nsteps, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        fem = _make(rank, world, math='exact', timeout_ms=20000)
        fem.set_global_U(_initial())
        its = []
        for _ in range(nsteps):
            fem.step(1)
            its.append(fem.last_solve()[0])
        torch.cuda.synchronize()
        q.put((rank, fem.lo, fem.owned_U().cpu().numpy(), its))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_gpus_match_one_rank(cuda):
    world = 2
    if torch.cuda.device_count() < world:
        pytest.skip('needs %d GPUs' % world)
    nsteps = 20
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nsteps, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        got = [q.get(timeout=120) for _ in range(world)]
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:
        for p in procs:
            if p.is_alive():
                p.kill()
    ref = _make(math='exact')
    ref.set_global_U(_initial())
    ref_its = []
    for _ in range(nsteps):
        ref.step(1)
        ref_its.append(ref.last_solve()[0])
    Uref = ref.owned_U().cpu().numpy()
    for rank, lo, U, its in got:
        assert _same_iterations(its, ref_its), (its, ref_its)
        assert rel_l2(U, Uref[lo:lo + U.shape[0]]) <= TOL_CG


def test_conjgrad_takes_the_sell_path_for_large_systems(cuda):
    """gpuSolve.linearsolvers.ConjGrad above SELL_MIN_ROWS rows = the single-rank form of the partitioned kernel on
    SELL-32 storage: test_conjgrad.py:117-162 (seeded gold solution, zero guess) on a 48^3 tetrahedral operator."""
    from oracle import fem as ofem
    from pdensorflow_b200.gpuSolve.entities.synthetic import structured_tet_operator
    from pdensorflow_b200.gpuSolve.linearsolvers import ConjGrad, JacobiPrecond
    from pdensorflow_b200.gpuSolve.linearsolvers.conjgrad import SELL_MIN_ROWS
    from pdensorflow_b200.gpuSolve.matrices import CSRMatrix
    m = 48
    n = m ** 3
    assert n >= SELL_MIN_ROWS
    rp, ci, kv, ml, mv = structured_tet_operator(m, m, m, H, 1.0, consistent_mass=True)     # sigma = 1: a real Laplacian part
    val = ofem.csr_axpby(mv, 1.0, kv, DT)
    gold = np.random.default_rng(0).standard_normal(n).astype(np.float32)
    b = ofem.spmv(rp, ci, val, gold)
    M = CSRMatrix(rp, ci, val, n)
    pc = JacobiPrecond()
    pc.build_preconditioner(*M.coo(), n)
    toll = 1e-6 * float(np.linalg.norm(b))
    cg = ConjGrad({'maxiter': 2000, 'toll': toll, 'precond': pc})
    cg.set_matrix(M)
    cg.set_X0(torch.zeros(n, 1, device='cuda'))
    cg.set_RHS(torch.from_numpy(b[:, None]).cuda())
    cg.solve()
    assert cg._part is not None and cg._handle is None
    x = cg.X().cpu().numpy().ravel()
    assert cg.niters() < 2000 and cg.niters() % 5 == 0
    assert np.linalg.norm(b - ofem.spmv(rp, ci, val, x)) / np.linalg.norm(b) < 1e-4
    minv = ofem.jacobi(rp, ci, val, n)
    xo, nit_o, _ = ofem.pcg(rp, ci, val, minv, b, np.zeros(n, np.float32), 2000, toll)
    assert abs(cg.niters() - nit_o) <= 5
    assert np.linalg.norm(x - xo) / np.linalg.norm(xo) < 1e-4


def test_monodomain_solver_large_mesh_fuses_the_mass_product(cuda, monkeypatch):
    """Above SELL_MIN_ROWS rows MonodomainSolver.solve_step = ionic kernel + ONE solve kernel that forms MASS*RHS0
    itself (SELL-32); it must step like the small-system path (MASS SpMV + CSR PCG) on the same mesh."""
    from test_fem_gpu import _square
    from pdensorflow_b200.gpuSolve.ionic import Fenton4v
    from pdensorflow_b200.gpuSolve.linearsolvers import conjgrad as cgmod
    from pdensorflow_b200.gpuSolve.physics import MonodomainSolver
    P, E, F = _square(320)                                   # 102 400 nodes
    sig = {1: 1e-3, 2: 1e-3, 3: 1e-3, 4: 1e-3}
    out, fused = [], []
    for min_rows in (10 ** 9, cgmod.SELL_MIN_ROWS):
        monkeypatch.setattr(cgmod, 'SELL_MIN_ROWS', min_rows)
        model = MonodomainSolver(Fenton4v(dt=0.1), {'use_renumbering': True, 'dt': 0.1, 'dt_per_plot': 10, 'Tend': 100})
        model.domain().set_mesh(P, E, F)
        model.add_element_material_property('sigma_l', 'region', sig)
        model.add_element_material_property('sigma_t', 'region', sig)
        model.assign_nodal_properties()
        model.assemble_matrices()
        model.set_initial_condition()
        model.add_stimulus(P[:, 0] < 0.05 * P[:, 0].max(), {'tstart': 0.0, 'nstim': 1, 'period': 100, 'duration': 0.4, 'intensity': 60.0})
        model.solver().set_maxiter(200)
        model.finalize_for_run()
        t = 0.0
        for _ in range(20):
            t += 0.1
            model.step(t)
        out.append(model.U().cpu().numpy().ravel())
        fused.append(model.solver()._part is not None)
    assert fused == [False, True] and out[0].max() > -70.0
    assert np.linalg.norm(out[0] - out[1]) <= TOL_CG * np.linalg.norm(out[0])
