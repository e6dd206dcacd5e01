"""Config 2 throughput (not the contract bench): 2-D Fenton-Karma 1024 x 1024, fp32, one GPU.
Per-step launches replayed as a CUDA graph (TMA / generic kernel) against the SM-resident kernel
(one launch for the whole block of steps).

    python tools/bench_config2.py [--n 1024] [--steps 2000]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=1024)
    ap.add_argument('--steps', type=int, default=2000)
    args = ap.parse_args()
    from pdensorflow_b200.fd import FDStepper
    shape = (args.n, args.n)
    U = np.full(shape, -80.0, dtype=np.float32)
    U[0:2, :] = 20.0
    for kernel in ('tma', 'resident'):
        try:
            st = FDStepper(shape, (1.0, 1.0), 0.1, 0.8, model='fenton4v', math='fast', kernel=kernel)
        except Exception as e:           # grid too large for the resident kernel
            print(json.dumps({'kernel': kernel, 'error': str(e)}), flush=True)
            continue
        st.set_U(U)
        (st.run if kernel == 'tma' else st.step)(200)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if kernel == 'tma':
                st.run(args.steps, block=100)
            else:
                st.step(args.steps)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / args.steps)
        n = args.n * args.n
        print(json.dumps({'workload': 'config 2: 2-D FD Fenton-Karma %dx%d, fp32' % shape, 'kernel': kernel,
                          'us_per_step': round(best * 1e3, 3), 'gnode_steps_per_s': round(n / best / 1e6, 2),
                          'algorithmic_gbs': round(32 * n / best / 1e6, 1), 'steps_timed': args.steps,
                          'finite': bool(torch.isfinite(st.U()).all()), 'umax': float(st.U().max())}), flush=True)


if __name__ == '__main__':
    main()
