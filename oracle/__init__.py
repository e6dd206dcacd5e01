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
TensorFlow cannot be imported in this image, so the reference's own kernels
cannot be run here.  What IS pinned:

* FEM local/global assembly - checked against the reference's own pure-NumPy
  code (``gpuSolve/matrices/localMass.py``, ``localStiffness.py`` and the
  TF-free half of ``globalMatrices.py``), imported by file path in this
  container; golden vectors under ``tests/golden/`` with the generating script.
* Ionic models - the reference's unit-test properties (rest fixed point,
  determinism, |dU| bound: ``Tests/CICD/unit/test_ionic.py:66-97``).
* 1-D mass/stiffness closed forms (``test_matrices.py:117-160``), csr_axpby
  (``test_csr_axpby.py:62-96``), PCG gold-solution recovery
  (``test_conjgrad.py:117-162``), 1-D/2-D mMS conduction velocity +-10 %
  (``test_mms_1d.py:140-176``, ``nightly/test_mms_2d_regression.py``).
* Stimulus timing - the step windows worked out in SURVEY.md section 8a row S.

PARITY UNPINNED by any reference golden vector: the FD Laplacians,
``enforce_boundary`` and Fenton-4v trajectories away from rest (the reference
holds no test or fixture for them).  For those the restatement below is the
only arbiter; it is frozen into ``tests/golden/*.npz`` with seeds.
"""
from . import fd, ionic, stimulus  # noqa: F401
