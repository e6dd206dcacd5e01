"""CPU legs of BASELINE.json's configurations 1, 2 and 5 (records for bench.py's other_configs and BASELINE.md
section 4): the reference's unfused op sequence restated in torch-CPU (oracle/torch_cpu.py - TensorFlow cannot be
installed in this image), all host threads, on bounded samples.  No GPU is touched.

    python tools/cpu_baselines.py [--which 1,2,5]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def config2(threads):
    from oracle import torch_cpu
    n, steps = 1024, 100
    sec, threads, _ = torch_cpu.time_fd_2d((n, n), steps, threads=threads)
    return {'workload': 'config 2 CPU leg: 2-D FD Fenton-Karma %dx%d, %d steps, unfused torch-CPU restatement of the TF path'
                        % (n, n, steps), 'cores': threads, 'kind': 'port', 'ms_per_step': round(sec / steps * 1e3, 3),
            'gnode_steps_per_s': round(n * n * steps / sec / 1e9, 5)}


def config1(threads):
    from oracle import fem as ofem
    from oracle import torch_cpu
    from pdensorflow_b200.gpuSolve.entities.synthetic import triangulated_square
    torch.set_num_threads(threads)
    P, E, F = triangulated_square()
    sig = {1: 1e-3, 2: 1e-3, 3: 1e-3, 4: 1e-3}
    n, dt = P.shape[0], 0.1
    ren = ofem.rcm_indexing(ofem.coo_pattern(ofem.connectivity(n, E)))
    m = ofem.assemble(P, E, F, sig, sig, ren['iperm'])
    Aval = ofem.csr_axpby(m['mass'], 1.0, m['stiffness'], dt)
    minv = torch.from_numpy(ofem.jacobi(m['rowptr'], m['col'], Aval, n))[:, None]
    MASS = torch_cpu.sparse_csr(m['rowptr'], m['col'], m['mass'], n)
    A = torch_cpu.sparse_csr(m['rowptr'], m['col'], Aval, n)
    U = torch.full((n, 1), -80.0)
    stim = torch.from_numpy((P[:, 0] < 0.05 * P[:, 0].max())[ren['perm']].astype(np.float32))[:, None]
    ion = torch_cpu.Fenton4v(dt)
    ion.initialize_state_variables(U)
    its, steps = [], 200
    for i in range(20 + steps):
        if i == 20:
            t0 = time.perf_counter()
        I0 = (60.0 if i < 4 else 0.0) * stim
        U, nit = torch_cpu.semi_implicit_step(ion, U, I0, dt, MASS, A, minv, n // 2)
        its.append(nit)
    sec = time.perf_counter() - t0
    return {'workload': 'config 1 CPU leg: Tests/FEM/Fenton/fenton.py step (ionic + MASS SpMV + Jacobi-PCG, eager op sequence) '
                        'on the 63 001-node stand-in mesh, %d steps, torch-CPU sparse CSR' % steps,
            'cores': threads, 'kind': 'port', 'ms_per_step': round(sec / steps * 1e3, 3),
            'mnode_steps_per_s': round(n * steps / sec / 1e6, 3), 'cg_iterations_per_step_mean': float(np.mean(its[20:]))}


def config5(threads):
    from oracle import torch_cpu
    from pdensorflow_b200.gpuSolve.entities.synthetic import structured_tet_operator
    torch.set_num_threads(threads)
    g, steps, dt = 96, 20, 0.1
    n = g ** 3
    rp, ci, kv, ml = structured_tet_operator(g, g, g, (0.25, 0.25, 0.25), 1e-3, 0, n)
    K = torch_cpu.sparse_csr(rp, ci, kv, n)
    mlinv = torch.from_numpy((1.0 / ml.astype(np.float64)).astype(np.float32))[:, None]
    U = torch.full((n, 1), -80.0)
    U[: 2 * g * g] = 20.0
    ion = torch_cpu.Fenton4v(dt)
    ion.initialize_state_variables(U)
    I0 = torch.zeros_like(U)
    for i in range(2 + steps):
        if i == 2:
            t0 = time.perf_counter()
        U = torch_cpu.explicit_lumped_step(ion, U, I0, dt, K, mlinv)
    sec = time.perf_counter() - t0
    return {'workload': 'config 5 CPU leg: explicit lumped FK step on a %d^3-node sample (%d nodes, %.1f nnz/row) of the '
                        'structured tetrahedral mesh, %d steps, torch-CPU sparse CSR; NO reference implementation exists '
                        'for this step (SURVEY 8a row E): parity unpinned, the leg is a port of the oracle'
                        % (g, n, len(kv) / n, steps),
            'cores': threads, 'kind': 'port', 'ms_per_step': round(sec / steps * 1e3, 2),
            'gnode_steps_per_s': round(n * steps / sec / 1e9, 5), 'finite': bool(torch.isfinite(U).all())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--which', default='1,2,5')
    args = ap.parse_args()
    threads = os.cpu_count() or 1
    for k in args.which.split(','):
        fn = {'1': config1, '2': config2, '5': config5}[k.strip()]
        try:
            print(json.dumps(fn(threads)), flush=True)
        except Exception as e:
            print(json.dumps({'workload': 'config %s CPU leg' % k, 'error': repr(e)[:200]}), flush=True)


if __name__ == '__main__':
    main()
