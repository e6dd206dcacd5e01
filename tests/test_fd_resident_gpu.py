"""SM-resident multi-step kernel for 2-D grids (config 2 path, pdx_fd_resident.cu): ONE launch advances
nsteps steps with U and the gates held in shared memory, CTAs exchanging edge rows through flags.
EXACT arithmetic must be bit-identical to the one-launch-per-step generic kernel; both against the oracle."""
import numpy as np
import pytest
import torch

from helpers import rel_l2, smooth_fields
from test_fd_gpu import NSTATES, TOL, run_gpu, run_oracle

pytestmark = pytest.mark.gpu
DT, DIFF, SPACING = 0.1, 0.8, (1.0, 1.2)


@pytest.mark.parametrize('shape', [(64, 96), (37, 50), (3, 5), (300, 1000), (1024, 1024), (149, 33)])
@pytest.mark.parametrize('model', ['fenton4v', 'mms2v', None])
def test_exact_is_bitwise_equal_to_the_per_step_kernel(cuda, shape, model):
    U, states = smooth_fields(shape, seed=5, nstates=NSTATES[model])
    for nsteps in (1, 2, 9, 24):
        st, Ur, sr = run_gpu(shape, SPACING, DT, DIFF, model, 'homog', 'exact', 'resident', U, states, nsteps)
        assert st.kernel == 'resident' and st.launches() == 1
        _, Ug, sg = run_gpu(shape, SPACING, DT, DIFF, model, 'homog', 'exact', 'generic', U, states, nsteps)
        assert np.array_equal(Ur, Ug), (shape, model, nsteps, rel_l2(Ur, Ug))
        for a, b in zip(sr, sg):
            assert np.array_equal(a, b)


@pytest.mark.parametrize('math', ['exact', 'fast'])
def test_against_oracle_and_chained_calls(cuda, math):
    from pdensorflow_b200.fd import FDStepper
    shape = (70, 45)
    U, states = smooth_fields(shape, seed=8)
    st = FDStepper(shape, SPACING, DT, DIFF, model='fenton4v', math=math, kernel='resident')
    st.set_U(U)
    st.set_states(states)
    for n in (7, 8, 1, 14):                 # odd / even calls: buffer parity and the flag base move along
        st.step(n)
    torch.cuda.synchronize()
    Uo, so = run_oracle(shape, SPACING, DT, DIFF, 'fenton4v', U, states, 30)
    assert rel_l2(st.U().cpu().numpy(), Uo) <= TOL[math]
    for a, b in zip(st.states, so):
        assert rel_l2(a.cpu().numpy(), b) <= 10 * TOL[math]


def test_auto_uses_it_for_long_calls_only(cuda):
    from pdensorflow_b200.fd import FDStepper
    shape = (128, 256)
    U, states = smooth_fields(shape, seed=9)
    ref = FDStepper(shape, SPACING, DT, DIFF, model='fenton4v', math='exact', kernel='generic')
    ref.set_U(U); ref.set_states(states)
    ref.step(43)
    st = FDStepper(shape, SPACING, DT, DIFF, model='fenton4v', math='exact')
    assert st.resident
    st.set_U(U); st.set_states(states)
    st.step(3)                              # short call: per-step kernel
    n0 = st.launches()
    st.run(40)                              # long call: one resident launch, no graph
    assert n0 == 3 and st.launches() == 4
    torch.cuda.synchronize()
    assert np.array_equal(st.U().cpu().numpy(), ref.U().cpu().numpy())
    big = FDStepper((2048, 2048), SPACING, DT, DIFF, model='fenton4v')       # does not fit on chip
    assert not big.resident
    with pytest.raises(Exception):
        FDStepper((2048, 2048), SPACING, DT, DIFF, model='fenton4v', kernel='resident')


def test_config2_protocol_short(cuda):
    """Config 2 (1024^2, S1 edge stimulus, fp32) for 200 steps: resident FAST vs per-step FAST."""
    from pdensorflow_b200.fd import FDStepper
    shape = (1024, 1024)
    U = np.full(shape, -80.0, dtype=np.float32)
    U[0:2, :] = 20.0
    out = []
    for kernel in ('resident', 'tma'):
        st = FDStepper(shape, (1.0, 1.0), 0.1, 0.8, model='fenton4v', math='fast', kernel=kernel)
        st.set_U(U)
        st.step(200)
        torch.cuda.synchronize()
        out.append(st.U().cpu().numpy())
    assert out[0].max() > 0 and rel_l2(out[0], out[1]) <= 1e-5
