"""Semi-implicit FEM step across GPUs (not the contract bench): structured tetrahedral mesh n^3 nodes,
Fenton-Karma, (MASS + dt*STIFFNESS) U1 = MASS (U0 + dt*(dU + I0)) solved by the row-block partitioned
persistent PCG kernel (pdx_pcg_part_solve); reports time per time step and per CG iteration.

    python tools/bench_fem_semi.py [--n 160]      |  torchrun --nproc-per-node N tools/bench_fem_semi.py
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
    ap.add_argument('--n', '--size', dest='n', type=int, default=160)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--profile', action='store_true', help='print the phase durations of one solve (device %globaltimer stamps)')
    args = ap.parse_args()
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from pdensorflow_b200.fem_dist import PartitionedSemiImplicitFEM, row_blocks
    from pdensorflow_b200.gpuSolve.entities.synthetic import structured_tet_operator
    nx = ny = nz = args.n
    n = nx * ny * nz
    lo, hi = row_blocks(n, world)[rank]
    t0 = time.time()
    rp, ci, kv, ml, mv = structured_tet_operator(nx, ny, nz, (0.25, 0.25, 0.25), 1e-3, lo, hi, consistent_mass=True)
    setup = time.time() - t0
    fem = PartitionedSemiImplicitFEM(n, rp, ci, mv, kv, 0.1, model='fenton4v', math='fast', rank=rank, world=world,
                                     maxiter=200, toll=1e-5, timeout_ms=30000)
    U = np.full(n, -80.0, dtype=np.float32)
    U[: 2 * ny * nz] = 20.0
    fem.set_global_U(U)
    fem.step(5)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    its = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        fem.step(1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    # iteration counts of a second, instrumented pass (reading them synchronises, so not inside the timed loop)
    for _ in range(5):
        fem.step(1)
        its.append(fem.last_solve()[0])
    if world > 1:
        t = torch.tensor([ms], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    rbar = len(kv) / (hi - lo)
    if args.profile and rank == 0:
        from pdensorflow_b200 import _lib as L
        stamps = torch.zeros(512, dtype=torch.int64, device='cuda')
        L.check(L.lib().pdx_pcg_part_set_profile(fem._h, L.ptr(stamps), 512))
        fem.step(1)
        torch.cuda.synchronize()
        L.check(L.lib().pdx_pcg_part_set_profile(fem._h, None, 0))
        t = stamps.cpu().numpy()
        t = t[t > 0]
        d = np.diff(t) / 1e3
        names = ['setup all-reduce'] + ['spmv', 'all-reduce 1', 'x/r update', 'all-reduce 2', 'p update', 'barrier'] * 64
        per = {}
        for nm, v in zip(names, d):
            per.setdefault(nm, []).append(float(v))
        print(json.dumps({'phase_us_median': {k: round(float(np.median(v)), 2) for k, v in per.items()},
                          'solve_us': round(float(t[-1] - t[0]) / 1e3, 1), 'stamps': int(len(t))}), flush=True)
    it = float(np.mean(its))
    bpi = 44 + 8 * rbar                    # algorithmic bytes per node-iteration
    bps = it * bpi + 32 + 2 * (8 * rbar + 12)   # per time step: + the ionic pass and the two set-up products (same table)
    if rank == 0:
        print(json.dumps({'workload': 'semi-implicit FK, structured tet mesh %d^3 = %d nodes, row-block x%d' % (args.n, n, world),
                          'n_gpus': world, 'ms_per_step': round(ms, 4), 'cg_iterations_per_step': it,
                          'us_per_cg_iteration': round(1e3 * ms / max(it, 1), 2),
                          'gnode_iterations_per_s': round(n * it / ms / 1e6, 2), 'gnode_steps_per_s': round(n / ms / 1e6, 3),
                          'nnz_per_row': round(rbar, 2), 'algorithmic_bytes_per_node_iteration': round(bpi, 1),
                          'algorithmic_gbs_per_gpu_iterations_only': round(n / world * it * bpi / ms / 1e6, 1),
                          'algorithmic_bytes_per_node_step': round(bps, 1),
                          'algorithmic_gbs_per_gpu': round(n / world * bps / ms / 1e6, 1),
                          'cooperative_blocks': fem.blocks_per_grid, 'threads_per_block': fem.threads_per_block,
                          'sell_slices_staged_by_copy_engine': fem.stage_slot_bytes > 0, 'ghost_rows': [fem.glo, fem.ghi],
                          'setup_s_per_rank': round(setup, 1), 'finite': bool(torch.isfinite(fem.owned_U()).all())}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
