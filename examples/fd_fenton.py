"""The reference's FD example (Tests/FD/Fenton/fenton.py: Fenton-Karma monodomain on a box, S1 plane + S2 quadrant,
frames to a ResultWriter) written against the drop-in package.

    python examples/fd_fenton.py [--size 64] [--samples 1000] [--out /tmp/cube3D]

What changes for a PDEnsorflow user: ``import pdensorflow_b200.dropin`` first, tensors are CUDA torch tensors, and the
body of ``solve`` + the per-step Python loop collapse into ``FDStepper`` (one fused kernel per time step, CUDA-graph
blocks between frames).  Everything else - the ionic-model class, Stimulus, the writer - keeps the reference's API.
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pdensorflow_b200.dropin  # noqa: F401  (registers the mirror as `gpuSolve`)
from gpuSolve.force_terms import Stimulus
from gpuSolve.IO.writers import ResultWriter
from pdensorflow_b200.fd import FDStepper


def run(config, writer=None):
    W, H, D = config['width'], config['height'], config['depth']
    dt, per_plot = config['dt'], config['dt_per_plot']
    stepper = FDStepper((W, H, D), (config['dx'], config['dy'], config['dz']), dt, config['diff'], model='fenton4v',
                        lap='conv' if config.get('convl') else 'homog')
    U = np.full((W, H, D), -80.0, dtype=np.float32)
    U[0:2] = 20.0                                               # S1: the first two planes (fenton.py:119-121)
    stepper.set_U(U)
    s2_init = np.full((W, H, D), -80.0, dtype=np.float32)
    s2_init[:W // 2, :H // 2, :] = 20.0
    s2 = Stimulus({'tstart': config['s2_time'], 'nstim': 1, 'period': 800, 'duration': dt, 'dt': dt, 'intensity': 60.0})
    s2.set_stimregion(s2_init > 0.0)                            # the cross-field the example intends
    if writer is not None:
        writer.imshow(stepper.U())
    i = 0
    t0 = time.time()
    while i < config['samples']:
        nxt = min(config['samples'], (i // per_plot + 1) * per_plot)             # next frame
        fires = [s2.stimulate_tissue_timestep(k, dt) for k in range(i, nxt)]      # host-side timing logic, as the reference
        if not any(fires):
            stepper.run(nxt - i)                                # the whole frame interval: one CUDA-graph replay
        else:
            for f in fires:
                stepper.step(1)
                if f:
                    stepper.apply_stimulus(s2.get_stimregion(), s2.intensity())
        i = nxt
        if writer is not None:
            writer.imshow(stepper.U())                          # returns at once: copy engine + writer thread
    torch.cuda.synchronize()
    print('solution: %d steps of %d nodes in %.3f s' % (config['samples'], W * H * D, time.time() - t0))
    if writer is not None:
        writer.wait()
    return stepper


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=64)
    ap.add_argument('--samples', type=int, default=1000)
    ap.add_argument('--out', default='/tmp/cube3D')
    a = ap.parse_args()
    config = {'width': a.size, 'height': a.size, 'depth': a.size, 'dx': 1, 'dy': 1, 'dz': 1, 'dt': 0.1, 'dt_per_plot': 10,
              'diff': 0.8, 'samples': a.samples, 's2_time': 200, 'convl': False, 'prefix_name': a.out}
    im = ResultWriter(config)
    im.width, im.height, im.depth = config['width'], config['height'], config['depth']
    run(config, im)


if __name__ == '__main__':
    main()
