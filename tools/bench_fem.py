"""FEM throughput (not the contract bench): config 5 (272^3-node structured tetrahedral mesh, explicit
mass-lumped Fenton-Karma step, row-block partitioned over the ranks) and, on rank 0 of a 1-rank run,
config 1 (63 001-node triangulated square, the reference's semi-implicit MonodomainSolver step).

    python tools/bench_fem.py [--n 272]          |  torchrun --nproc-per-node N tools/bench_fem.py
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', '--size', dest='n', type=int, default=272)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--skip-config1', action='store_true')
    ap.add_argument('--layout', default='sell', choices=['sell', 'csr'])
    ap.add_argument('--sync', default='auto', choices=['auto', 'flags', 'barrier'])
    ap.add_argument('--ab-stage', action='store_true', help='A/B in ONE process: the one-thread-per-row kernel and the copy-engine staged kernel, alternating')
    ap.add_argument('--no-graph', action='store_true', help='one Python launch per step instead of graph replays of 10 steps')
    args = ap.parse_args()
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from pdensorflow_b200.fem_dist import PartitionedExplicitFEM, row_blocks
    from pdensorflow_b200.gpuSolve.entities.synthetic import structured_tet_operator
    nx = ny = nz = args.n
    n = nx * ny * nz
    lo, hi = row_blocks(n, world)[rank]
    t0 = time.time()
    # h = 0.25 mm, sigma = 1e-3 (Tests/FEM/Fenton/fenton.py:64-65), dt = 0.1: dt*lambda_max ~ 0.5
    rp, ci, kv, ml = structured_tet_operator(nx, ny, nz, (0.25, 0.25, 0.25), 1e-3, lo, hi)
    setup = time.time() - t0
    fem = PartitionedExplicitFEM(n, rp, ci, kv, ml, 0.1, model='fenton4v', math='fast', rank=rank, world=world,
                                 layout=args.layout, sync=args.sync)
    U = fem.current_U()
    U.fill_(-80.0)
    U[: 2 * ny * nz] = 20.0
    fem.step(20)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.ab_stage and world == 1:
        res = {'0': [], '2': []}
        for rep in range(4):
            for mode in ('0', '2'):
                os.environ['PDX_EXPL_STAGE'] = mode
                fem.step(10)
                torch.cuda.synchronize()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                fem.step(args.steps)
                a1.record()
                torch.cuda.synchronize()
                res[mode].append(round(a0.elapsed_time(a1) / args.steps, 4))
        print(json.dumps({'ab_ms_per_step': {'thread_per_row': res['0'], 'copy_engine_staged': res['2']}, 'nodes': n}), flush=True)
        return
    advance = fem.step if args.no_graph else fem.run
    advance(20)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0.record()
    advance(args.steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    rbar = len(kv) / (hi - lo)
    bpn = 40 + 8 * rbar
    ok = bool(torch.isfinite(fem.owned_U()).all())
    if rank == 0:
        print(json.dumps({'workload': 'config 5: explicit lumped FK, structured tet mesh %d^3 = %d nodes, row-block x%d' % (args.n, n, world),
                          'n_gpus': world, 'ms_per_step': round(ms, 4), 'gnode_steps_per_s': round(n / ms / 1e6, 2),
                          'nnz_per_row': round(rbar, 2), 'algorithmic_bytes_per_node_step': round(bpn, 1),
                          'algorithmic_gbs_per_gpu': round(n / world * bpn / ms / 1e6, 1), 'setup_s_per_rank': round(setup, 1),
                          'ghost_rows_sent': [fem.send_lo, fem.send_hi], 'sync': fem.sync, 'graph_blocks': not args.no_graph, 'finite': ok, 'layout': args.layout,
                          'parity': 'UNPINNED: the explicit lumped step has no reference implementation (SURVEY 8a row E); '
                                    'checked against the oracle restatement only (tests/test_fem_dist_gpu.py)',
                          'sell_padding': round(getattr(fem, 'padding', 1.0), 4)}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
        return
    if args.skip_config1:
        return
    # ---- config 1: Tests/FEM/Fenton/fenton.py on the synthetic stand-in mesh -------------------
    import tempfile
    from pdensorflow_b200.gpuSolve.entities.synthetic import save_pkl, triangulated_square
    from pdensorflow_b200.gpuSolve.ionic import Fenton4v
    from pdensorflow_b200.gpuSolve.physics import MonodomainSolver
    fn = os.path.join(tempfile.mkdtemp(), 'sq.pkl')
    save_pkl(fn, *triangulated_square())
    for explicit in (False, True):
        cfg = {'mesh_file_name': fn, 'use_renumbering': True, 'dt': 0.1, 'dt_per_plot': 10, 'Tend': 1000}
        if explicit:
            cfg['explicit_lumped'] = True
        model = MonodomainSolver(Fenton4v(dt=0.1), cfg)
        s = {1: 0.001, 2: 0.001, 3: 0.001, 4: 0.001}
        model.add_element_material_property('sigma_l', 'region', s)
        model.add_element_material_property('sigma_t', 'region', s)
        model.assign_nodal_properties()
        model.assemble_matrices()
        pts = model.domain().Pts()
        Lx = pts[:, 0].max()
        model.set_initial_condition()
        model.add_stimulus(pts[:, 0] < 0.05 * Lx, {'tstart': 0.0, 'nstim': 1, 'period': 100, 'duration': 0.4,
                                                    'intensity': 60.0, 'name': 'S1'})
        model.solver().set_maxiter(pts.shape[0] // 2)
        model.finalize_for_run()
        ctime, its = 0.0, []
        for i in range(50):
            ctime += 0.1
            model.step(ctime)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nst = 300
        for i in range(nst):
            ctime += 0.1
            model.step(ctime)
        enq = (time.perf_counter() - t0) / nst * 1e3        # host time to ENQUEUE a step (Python + ctypes + launches)
        torch.cuda.synchronize()
        ms1 = (time.perf_counter() - t0) / nst * 1e3
        for i in range(6 if not explicit else 0):           # iteration counts: reading them synchronises - not timed
            ctime += 0.1
            model.step(ctime)
            its.append(model.solver().niters())
        N = pts.shape[0]
        print(json.dumps({'workload': 'config 1: Tests/FEM/Fenton/fenton.py on the synthetic 251x251 triangulated square (%d nodes), %s'
                          % (N, 'explicit lumped (opt-in)' if explicit else 'semi-implicit Jacobi-PCG as shipped'),
                          'ms_per_step': round(ms1, 4), 'mnode_steps_per_s': round(N / ms1 / 1e3, 2),
                          'host_enqueue_ms_per_step': round(enq, 4), 'step_graphs': sum(1 for v in model._step_graphs.values() if isinstance(v, tuple)), 'cg_iterations_per_step_sampled': its}), flush=True)


if __name__ == '__main__':
    main()
