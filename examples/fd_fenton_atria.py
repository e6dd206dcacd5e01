#!/usr/bin/env python
"""Tests/FD/Fenton_atria/fenton.py on the B200 path: anatomy from a PNG mosaic (ImageData -> Domain3D, standard-library
decoder), heterogeneous Laplacian fused with Fenton-Karma, S2 through the Stimulus class, frames to a ResultWriter.

    python examples/fd_fenton_atria.py [--png tests/golden/atria_mosaic.png --Mx 16 --My 8] [--samples 2000]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--png', default=os.path.join(ROOT, 'tests', 'golden', 'atria_mosaic.png'))
    ap.add_argument('--Mx', type=int, default=16)
    ap.add_argument('--My', type=int, default=8)
    ap.add_argument('--samples', type=int, default=2000)
    ap.add_argument('--s2-time', type=float, default=190.0)
    ap.add_argument('--out', default='')
    args = ap.parse_args()
    import torch
    import pdensorflow_b200.dropin  # noqa: F401  (registers the mirror as `gpuSolve`)
    from gpuSolve.entities.domain3D import Domain3D
    from gpuSolve.force_terms import Stimulus
    from pdensorflow_b200.fd import FDStepper
    dt, diff, per_plot = 0.1, 0.75, 10
    dom = Domain3D({'dx': 1, 'dy': 1, 'dz': 1})
    dom.load_geometry_file(args.png, args.Mx, args.My, 1.e-4)
    dom.load_conductivity(fname='', cond_unif=diff)
    W, H, D = dom.width(), dom.height(), dom.depth()
    mask = dom.geometry() > 0.0
    U = np.full((W, H, D), -80.0, dtype=np.float32)
    U[:, :, (D // 2 - 10):(D // 2 + 10)] = 20.0
    U = np.where(mask, U, np.float32(-80.0)).astype(np.float32)
    s2_init = dom.geometry().astype(np.float32)
    s2_init[:, :H // 2, :] = 0.0
    s2_init = s2_init * np.float32(100.0) + np.float32(-80.0)
    st = FDStepper((W, H, D), (dom.dx(), dom.dy(), dom.dz()), dt, 1.0, model='fenton4v', lap='heterog',
                   DIFF=dom.conductivity().astype(np.float32), math='fast')
    st.set_U(U)
    s2 = Stimulus({'tstart': args.s2_time, 'nstim': 1, 'period': 800, 'duration': dt, 'dt': dt, 'intensity': 60.0})
    s2.set_stimregion(s2_init)
    writer = None
    if args.out:
        from gpuSolve.IO.writers import ResultWriter
        writer = ResultWriter({'prefix_name': args.out, 'width': W, 'height': H, 'depth': D, 'samples': args.samples,
                               'dt_per_plot': per_plot})
    mask_d = torch.from_numpy(mask).cuda()
    t0 = time.time()
    for i in range(args.samples):
        st.step(1)
        if s2.stimulate_tissue_timestep(i, dt):
            st.apply_stimulus(s2.get_stimregion(), s2.intensity())
        if writer is not None and i % per_plot == 0:
            writer.imshow(torch.where(mask_d, st.U(), torch.tensor(-1.0, device='cuda')))
    torch.cuda.synchronize()
    el = time.time() - t0
    if writer is not None:
        writer.wait()
    print('%dx%dx%d voxels (%.1f %% tissue), %d steps in %.2f s = %.2f Gnode-steps/s; U in [%.2f, %.2f]' % (
        W, H, D, 100.0 * mask.mean(), args.samples, el, W * H * D * args.samples / el / 1e9, float(st.U().min()), float(st.U().max())))


if __name__ == '__main__':
    main()
