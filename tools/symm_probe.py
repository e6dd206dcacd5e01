import os, sys, torch, torch.distributed as dist
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
import torch.distributed._symmetric_memory as symm
print('symm_mem import ok; backend:', getattr(symm, 'get_backend', lambda d: '?')(torch.device('cuda', local)), flush=True)
t = symm.empty((4, 1024), dtype=torch.float32, device='cuda')
h = symm.rendezvous(t, dist.group.WORLD)
print(rank, 'rendezvous ok', type(h), [m for m in dir(h) if not m.startswith('_')], flush=True)
t.fill_(float(rank))
h.barrier()
peer = h.get_buffer((rank + 1) % world, (4, 1024), torch.float32)
print(rank, 'peer value', float(peer[0, 0]), 'ptrs', [hex(p) for p in h.buffer_ptrs], flush=True)
peer[1].fill_(100.0 + rank)          # remote write
h.barrier()
print(rank, 'my row1 after peer write', float(t[1, 0]), flush=True)
# timing of barrier
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100): h.barrier()
e1.record(); torch.cuda.synchronize()
print(rank, 'barrier us', e0.elapsed_time(e1) * 10, flush=True)
dist.destroy_process_group()
