"""CPU oracle for PDEnsorflow's monodomain time-stepping hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pdensorflow_b200/`` may import this
package: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the reported CPU baseline, never as the product.

What this is
------------
An op-for-op NumPy fp32 restatement of the reference algorithm (TensorFlow ops
replaced by the NumPy op with the same IEEE semantics, evaluated in the order
the reference source writes them).  Every function cites the reference
``file:line`` it follows (paths relative to ``/root/reference/PDEnsorflow/``).

Pinning status (see DESIGN.md "Oracle")
---------------------------------------
TensorFlow cannot be imported in this image, so the reference's TF *kernels* cannot be run
here.  The reference's own *source* can: ``tests/golden/make_ref_golden.py`` imports the
unmodified reference from /root/reference over ``tests/golden/tf_facade.py`` (every tf op ->
the NumPy op with the same IEEE fp32 semantics) and freezes its outputs on seeded inputs in
``tests/golden/ref_*.npz``.  ``tests/test_ref_golden_cpu.py`` holds this oracle to those
fixtures:

* bit-exact: all six diffop Laplacians (2-D/3-D homogeneous, conv, heterogeneous,
  anisotropic), ``enforce_boundary``, Stimulus step/float-time windows, one ``differentiate``
  call of all five ionic models (per-node parameters included), paced single-cell
  trajectories of all five models (three TTP cell types) through an action potential, the
  ``Tests/FD/Fenton/fenton.py`` example run end to end (homogeneous, conv and heterogeneous
  operators, the all-ones S2 quirk included);
* CRN/TTP lookup tables: samples within 2e-6 relative (NumPy's fp32 exp is SIMD-dispatch
  dependent) and fp64 column checksums;
* ``MonodomainSolver.step`` (assembly + RCM + stimulus + Jacobi-PCG) on 1-D and 2-D meshes:
  identical CG iteration counts every step, potentials within 1e-6 relative L2;
* FEM local/global assembly also against the reference's pure-NumPy code imported by file
  path (``tests/golden/fem_golden.npz``, ``make_fem_golden.py``), plus the reference's unit-test
  properties (rest fixed point, closed-form 1-D matrices, csr_axpby, CG gold recovery, CV +-10 %).

What the facade cannot pin: ulp-level differences between TF's Eigen/XLA transcendental
kernels (tanh, exp, log) and the correctly rounded fp32 values used here, and TF's
unspecified reduction order in ``reduce_sum``; both are far inside the north-star tolerance
(rel-L2 <= 1e-5).  Only the explicit mass-lumped FEM step has no reference
implementation at all and is therefore PARITY UNPINNED.
"""
from . import fd, ionic, ionic_lut, stimulus  # noqa: F401
