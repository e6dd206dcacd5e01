"""Slab-decomposed stepping on >= 2 GPUs must equal the single-GPU run bit for bit in EXACT
arithmetic (same per-node operations), for both halo transports: 'peer' (halo stores fused into the
step kernel over NVLink symmetric memory + one barrier per step) and 'nccl' (send/recv, overlapped
or not).  Skipped on a 1-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, shape, nsteps, overlap, halo, q):
    import torch.distributed as dist
    from pdensorflow_b200.fd import DistributedFDStepper
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        rng = np.random.default_rng(0)
        G = rng.uniform(-85, 25, size=shape).astype(np.float32)
        S = [rng.uniform(0, 1, size=shape).astype(np.float32) for _ in range(3)]
        ds = DistributedFDStepper(shape, (1.0, 1.0, 1.0), 0.1, 0.8, rank=rank, world=world, overlap=overlap, halo=halo,
                                  model='fenton4v', math='exact', device=torch.device('cuda', rank))
        assert ds.halo == halo
        ds.set_global(G, S)
        ds.step(nsteps)
        torch.cuda.synchronize()
        q.put((rank, ds.part.x0, ds.owned_U().cpu().numpy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('halo,overlap,world', [('peer', True, 2), ('nccl', True, 2), ('nccl', False, 2), ('peer', True, 3),
                                                ('peer', True, 8)])
def test_multi_gpu_slabs_match_single_gpu(cuda, halo, overlap, world):
    if torch.cuda.device_count() < world:
        pytest.skip('needs %d GPUs' % world)
    from pdensorflow_b200.fd import FDStepper
    shape, nsteps = (22 if world == 2 else 37, 24, 132), 7
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, shape, nsteps, overlap, halo, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        got = [q.get(timeout=90) for _ in range(world)]
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:
        for p in procs:                 # never leave a rank behind on the GPU
            if p.is_alive():
                p.kill()
    rng = np.random.default_rng(0)
    G = rng.uniform(-85, 25, size=shape).astype(np.float32)
    S = [rng.uniform(0, 1, size=shape).astype(np.float32) for _ in range(3)]
    st = FDStepper(shape, (1.0, 1.0, 1.0), 0.1, 0.8, model='fenton4v', math='exact')
    st.set_U(G); st.set_states(S); st.step(nsteps)
    ref = st.U().cpu().numpy()
    out = np.zeros_like(ref)
    for rank, x0, owned in got:
        out[x0:x0 + owned.shape[0]] = owned
    assert np.array_equal(out, ref)


@pytest.mark.parametrize('world,kernel,lap', [(2, 'tma', 'homog'), (3, 'tma', 'homog'), (4, 'generic', 'homog'), (3, 'tma', 'heterog')])
def test_slabs_sharing_one_gpu_match_the_whole_grid(cuda, world, kernel, lap):
    """The fused halo store of pdx_fd_step_mirror on a ONE-GPU box: the slabs of a decomposition live in one process,
    each step kernel stores its edge planes straight into the neighbouring slabs' ghost planes (plain device pointers
    instead of NVLink mappings; stream order instead of the barrier).  Unequal slabs; bitwise equal to the whole grid."""
    import ctypes as C
    from pdensorflow_b200 import _lib as L
    from pdensorflow_b200.fd import FDStepper, SlabPartition
    shape, nsteps = (23, 24, 132) if kernel == 'tma' else (23, 12, 13), 9        # plane bytes are a multiple of 16
    rng = np.random.default_rng(1)
    G = rng.uniform(-85, 25, size=shape).astype(np.float32)
    S = [rng.uniform(0, 1, size=shape).astype(np.float32) for _ in range(3)]
    D = (0.8 * (rng.uniform(size=shape) > 0.1)).astype(np.float32) if lap == 'heterog' else None
    whole = FDStepper(shape, (1.0, 1.0, 1.0), 0.1, 0.8, model='fenton4v', math='exact', kernel=kernel, lap=lap, DIFF=D)
    whole.set_U(G); whole.set_states(S); whole.step(nsteps)
    parts = [SlabPartition(shape[0], r, world) for r in range(world)]
    slabs = []
    for p in parts:
        st = FDStepper(shape, (1.0, 1.0, 1.0), 0.1, 0.8, model='fenton4v', math='exact', kernel=kernel, lap=lap, part=p,
                       DIFF=None if D is None else p.scatter(D))
        st.set_U(p.scatter(G)); st.set_states([p.scatter(s) for s in S])
        slabs.append(st)
    plane = shape[1] * shape[2] * 4
    for _ in range(nsteps):
        for r, st in enumerate(slabs):
            lo = slabs[r - 1].other_U().data_ptr() + (parts[r - 1].nloc + 1) * plane if r > 0 else None
            hi = slabs[r + 1].other_U().data_ptr() if r < world - 1 else None
            L.check(L.lib().pdx_fd_step_mirror(st._plan, L.ptr(st.current_U()), L.ptr(st.other_U()), st._states_arr,
                                               L.ptr(st.DIFF), C.c_void_p(lo), C.c_void_p(hi), L.stream_ptr()))
        for st in slabs:
            st.swap()
    torch.cuda.synchronize()
    ref = whole.U().cpu().numpy()
    for p, st in zip(parts, slabs):
        assert np.array_equal(p.owned(st.U().cpu().numpy()), ref[p.x0:p.x1])
    assert len({p.nloc for p in parts}) > 1 or shape[0] % world == 0
