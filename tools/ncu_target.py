"""Short, deterministic launch sequence of one fused FD kernel for ncu (profiles/): ``--steps`` single-step launches of
the FAST Fenton-Karma step on ``--shape`` after the S1 wave has developed a little.

    python tools/ncu_target.py --shape 128,2048,2048 [--steps 12] [--model fenton4v] [--lap homog] [--kernel auto]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--shape', default='512,512,512')
    ap.add_argument('--steps', type=int, default=12)
    ap.add_argument('--model', default='fenton4v')
    ap.add_argument('--lap', default='homog')
    ap.add_argument('--kernel', default='auto')
    ap.add_argument('--math', default='fast')
    ap.add_argument('--x-chunk', type=int, default=0)
    args = ap.parse_args()
    from pdensorflow_b200.fd import FDStepper
    shape = tuple(int(v) for v in args.shape.split(','))
    model = None if args.model in ('none', 'heat') else args.model
    if model in ('crn', 'ttp'):
        from pdensorflow_b200.gpuSolve.ionic import CourtemancheRamirezNattel, TenTusscherPanfilov
        model = (CourtemancheRamirezNattel if model == 'crn' else TenTusscherPanfilov)(dt=0.02)
        dt = 0.02
    else:
        dt = 0.1
    DIFF = torch.full(shape, 0.8, device='cuda') if args.lap == 'heterog' else None
    AVEC = None
    if args.lap == 'aniso':
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from bench_variants import aniso_fields
        DIFF, AVEC = aniso_fields(shape, 0.8)
    st = FDStepper(shape, (1.0,) * len(shape), dt, 0.8, model=model, math=args.math, kernel=args.kernel, lap=args.lap,
                   DIFF=DIFF, AVEC=AVEC, x_chunk=args.x_chunk)
    U = st.U(); U.fill_(-80.0); U[0:2] = 20.0
    for _ in range(args.steps):
        st.step(1)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(st.U()).all())
    print('ok', shape, st.kernel, st.launches())


if __name__ == '__main__':
    main()
