"""Scratch: where does the multi-GPU step time go?  torchrun --nproc-per-node 2 tools/dist_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from pdensorflow_b200.fd import DistributedFDStepper, FDStepper

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
G = (128 * world, 2048, 2048)

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

if world == 1:
    for chunk in (0, 16, 32, 64, 128):
        st = FDStepper(G, (1., 1., 1.), 0.1, 0.8, x_chunk=chunk)
        st.U().fill_(-80.); st.U()[0:2] = 20.
        print('single 128x2048x2048 chunk=%d: %.3f ms/step' % (chunk, timeit(lambda: st.step(1))), flush=True)
        del st
else:
    for overlap in (True, False):
        ds = DistributedFDStepper(G, (1., 1., 1.), 0.1, 0.8, rank=rank, world=world, overlap=overlap, halo="nccl" if not overlap else "peer")
        ds.fd.U().fill_(-80.)
        t = timeit(lambda: ds.step(1))
        fd, p = ds.fd, ds.part
        t_int = timeit(lambda: fd.step_range(p.x_begin + 1, p.x_end - 1))
        t_b = timeit(lambda: (fd.step_range(p.x_begin, p.x_begin + 1), fd.step_range(p.x_end - 1, p.x_end)))
        t_x = timeit(lambda: ds.exchange(fd.other_U()))
        if rank == 0:
            print('world %d overlap=%s: step %.3f ms | interior only %.3f | 2 boundary planes %.3f | exchange only %.3f'
                  % (world, overlap, t, t_int, t_b, t_x), flush=True)
        del ds
    dist.destroy_process_group()
