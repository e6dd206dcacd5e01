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
