"""Assembly time, host NumPy path against the device path (pdx_fem_assemble), on a tetrahedral box of m^3 nodes.
Not the contract bench.     python tools/bench_assembly.py [--size 64]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=64)
    args = ap.parse_args()
    from pdensorflow_b200.gpuSolve.entities.materialproperties import MaterialProperties
    from pdensorflow_b200.gpuSolve.entities.synthetic import tetrahedral_box
    from pdensorflow_b200.gpuSolve.entities.triangulation import Triangulation
    from pdensorflow_b200.gpuSolve.matrices import globalMatrices as gm
    from pdensorflow_b200.gpuSolve.matrices.device_assembly import assemble_on_device
    from pdensorflow_b200.gpuSolve.matrices.localMass import localMass
    from pdensorflow_b200.gpuSolve.matrices.localStiffness import localStiffness
    m = args.size
    P, E, F = tetrahedral_box(m, m, m)
    T = Triangulation()
    T.set_mesh(P, E, F)
    mp = MaterialProperties()
    mp.add_element_property('sigma_l', 'region', {1: 1.0})
    mp.add_element_property('sigma_t', 'region', {1: 0.3})
    torch.zeros(1, device='cuda')
    t0 = time.time()
    conn = T.mesh_connectivity('True')
    pat = gm.compute_coo_pattern(conn)
    t_conn = time.time() - t0
    t0 = time.time()
    H = gm.assemble_matrices_dict({'mass': localMass, 'stiffness': localStiffness}, pat, T, mp, conn)
    t_host = time.time() - t0
    assemble_on_device(T, mp)                      # warm-up (context, allocator)
    torch.cuda.synchronize()
    t0 = time.time()
    D = assemble_on_device(T, mp)
    torch.cuda.synchronize()
    t_dev = time.time() - t0
    err = max(float(np.abs(D[k].val - H[k].val).max() / np.abs(H[k].val).max()) for k in ('mass', 'stiffness'))
    print(json.dumps({'mesh': 'tetrahedral box %d^3' % m, 'nodes': int(P.shape[0]), 'elements': int(E['Tetras'].shape[0]),
                      'nnz': int(H['mass'].nnz), 'host_connectivity_s': round(t_conn, 2), 'host_assembly_s': round(t_host, 2),
                      'device_assembly_s_incl_pattern_and_d2h': round(t_dev, 3),
                      'speedup_vs_host_total': round((t_conn + t_host) / t_dev, 1), 'max_rel_diff': err,
                      'same_pattern': bool(np.array_equal(D['mass'].col, H['mass'].col))}), flush=True)


if __name__ == '__main__':
    main()
