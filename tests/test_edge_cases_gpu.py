"""Degenerate inputs through the C ABI (the reference's own tests cover rest states and tiny meshes; SURVEY.md
section 4): empty arrays, zero steps, the smallest legal grid, an empty stimulus mask, maxiter = 0, a non-finite
right-hand side.  Nothing may crash, hang or touch memory it was not given."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import rel_l2, smooth_fields
from test_fd_gpu import run_gpu, run_oracle

pytestmark = pytest.mark.gpu


def test_empty_and_zero_step_calls(cuda):
    from pdensorflow_b200 import _lib as L
    from pdensorflow_b200.fd import FDStepper
    lib = L.lib()
    params = (C.c_float * L.MAX_PARAMS)(*L.default_params(L.MODEL_FENTON4V))
    z = torch.zeros(1, device='cuda')
    states = L.ptr_array([z, z, z])
    L.check(lib.pdx_ionic_step(L.MODEL_FENTON4V, L.MATH_FAST, params, None, 0, L.ptr(z), states, L.ptr(z), 0.1, None, None,
                               L.stream_ptr()))
    assert lib.pdx_csr_spmv(0, L.ptr(z), L.ptr(z), L.ptr(z), L.ptr(z), L.ptr(z), L.stream_ptr()) == 0
    L.check(lib.pdx_stim_apply(L.ptr(z), L.ptr(z), 1.0, 0, L.stream_ptr()))
    torch.cuda.synchronize()
    assert float(z[0]) == 0.0
    U, states = smooth_fields((6, 7, 8), seed=1)
    st = FDStepper((6, 7, 8), (1.0, 1.0, 1.0), 0.1, 0.8, model='fenton4v', math='exact')
    st.set_U(U)
    st.step(0)
    assert np.array_equal(st.U().cpu().numpy(), U)
    mask = torch.zeros((6, 7, 8), device='cuda')
    st.apply_stimulus(mask, 60.0)                       # nobody is stimulated
    assert np.array_equal(st.U().cpu().numpy(), U)


@pytest.mark.parametrize('shape', [(3, 3, 3), (3, 3), (3, 4, 1024), (4, 3)])
def test_smallest_grids(cuda, shape):
    U, states = smooth_fields(shape, seed=2)
    spacing = (1.0,) * len(shape)
    _, Ug, sg = run_gpu(shape, spacing, 0.1, 0.8, 'fenton4v', 'homog', 'exact', 'auto', U, states, 5)
    Uo, so = run_oracle(shape, spacing, 0.1, 0.8, 'fenton4v', U, states, 5)
    assert rel_l2(Ug, Uo) <= 1e-6
    _, Ug9, _ = run_gpu(shape, spacing, 0.1, 0.8, 'fenton4v', 'homog', 'exact', 'auto', U, states, 9)   # >= 8: resident path in 2-D
    Uo9, _ = run_oracle(shape, spacing, 0.1, 0.8, 'fenton4v', U, states, 9)
    assert rel_l2(Ug9, Uo9) <= 1e-6


def test_pcg_zero_iterations_and_non_finite_guard(cuda):
    from pdensorflow_b200 import _lib as L
    from pdensorflow_b200.gpuSolve.entities.synthetic import structured_tet_operator
    lib = L.lib()
    rp, ci, kv, ml, mv = structured_tet_operator(5, 4, 6, consistent_mass=True)
    n = len(ml)
    val = (mv + np.float32(0.1) * kv).astype(np.float32)
    d = [torch.from_numpy(a).cuda() for a in (rp, ci, val)]
    h = C.c_void_p()
    L.check(lib.pdx_pcg_create(n, L.ptr(d[0]), L.ptr(d[1]), L.ptr(d[2]), None, C.byref(h)))
    try:
        b = torch.ones(n, device='cuda')
        x = torch.full((n,), 0.5, device='cuda')
        info = torch.zeros(4, dtype=torch.int32, device='cuda')
        L.check(lib.pdx_pcg_solve(h, L.ptr(b), L.ptr(x), 0, 1e-6, 5, L.ptr(info), L.stream_ptr()))
        torch.cuda.synchronize()
        assert int(info[0]) == 0 and bool((x == 0.5).all())            # maxiter = 0: the guess is returned
        b[3] = float('nan')
        L.check(lib.pdx_pcg_solve(h, L.ptr(b), L.ptr(x), 200, 1e-6, 5, L.ptr(info), L.stream_ptr()))
        torch.cuda.synchronize()
        assert int(info[2]) & 2 and int(info[0]) <= 5                  # stopped at the first check: NaN/Inf guard (conjgrad.py:293-300)
    finally:
        lib.pdx_pcg_destroy(h)
