"""The reference's FEM example (Tests/FEM/Fenton/fenton.py: Fenton-Karma monodomain on a triangulated square,
semi-implicit Jacobi-PCG, S1-S2, frames to an IGB file) on the drop-in package - the script is the reference's,
minus TensorFlow, on a synthetic stand-in for the bundled mesh.

    python examples/fem_fenton.py [--n 251] [--tend 50] [--out /tmp/square.igb]
"""
import argparse
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pdensorflow_b200.dropin  # noqa: F401
from gpuSolve.ionic.fenton4v import Fenton4v
from gpuSolve.IO.writers import IGBWriter
from gpuSolve.physics import MonodomainSolver
from pdensorflow_b200.gpuSolve.entities.synthetic import save_pkl, triangulated_square


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=251)
    ap.add_argument('--tend', type=float, default=50.0)
    ap.add_argument('--out', default='/tmp/square.igb')
    a = ap.parse_args()
    mesh = os.path.join(tempfile.mkdtemp(), 'triangulated_square.pkl')
    save_pkl(mesh, *triangulated_square(a.n, a.n))
    dt = 0.1
    cfg = {'mesh_file_name': mesh, 'use_renumbering': True, 'dt': dt, 'dt_per_plot': int(1.0 / dt), 'Tend': a.tend}
    model = MonodomainSolver(Fenton4v(), cfg)
    sigma = {1: 0.001, 2: 0.001, 3: 0.001, 4: 0.001}
    model.add_element_material_property('sigma_l', 'region', sigma)
    model.add_element_material_property('sigma_t', 'region', sigma)
    model.assign_nodal_properties()
    model.assemble_matrices()
    pts = model.domain().Pts()
    Lx, Ly = pts[:, 0].max(), pts[:, 1].max()
    model.set_initial_condition()
    model.add_stimulus(pts[:, 0] < 0.05 * Lx, {'tstart': 0.0, 'nstim': 1, 'period': 800, 'duration': max(0.4, dt),
                                               'intensity': 60.0, 'name': 'S1'})
    model.add_stimulus(np.logical_and(pts[:, 0] < Lx, pts[:, 1] < 0.5 * Ly), {'tstart': 210.0, 'nstim': 1, 'period': 800,
                                                                              'duration': max(0.4, dt), 'intensity': 60.0,
                                                                              'name': 'S2'})
    model.solver().set_maxiter(pts.shape[0] // 2)
    model.finalize_for_run()
    im = IGBWriter({'fname': a.out, 'Tend': model.Tend(), 'nt': 1 + model.nt() // model.dt_per_plot(), 'nx': pts.shape[0]})
    im.imshow(model.U())
    ctime, t0 = 0.0, time.time()
    for i in range(model.nt()):
        ctime += model.dt()
        model.step(ctime)
        if (i + 1) % model.dt_per_plot() == 0:
            im.imshow(model.U())
    torch.cuda.synchronize()
    print('%d steps of %d nodes in %.3f s (last solve: %d CG iterations)' % (model.nt(), pts.shape[0], time.time() - t0,
                                                                            model.solver().niters()))
    im.wait()


if __name__ == '__main__':
    main()
