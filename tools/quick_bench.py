"""Scratch timing of the fused FD kernel (CUDA events); not the contract bench."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pdensorflow_b200.fd import FDStepper

def bench(shape, model='fenton4v', math='fast', kernel='auto', lap='homog', nsteps=40, x_chunk=0, reps=3):
    DIFF = None
    if lap == 'heterog':
        DIFF = torch.full(shape, 0.8, device='cuda')
    st = FDStepper(shape, (1.0,) * len(shape), 0.1, 0.8, model=model, math=math, kernel=kernel, lap=lap, DIFF=DIFF, x_chunk=x_chunk)
    U = st.U(); U.fill_(-80.0); U[0:2] = 20.0
    st.run(200)  # let the wave develop a bit + warm-up
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); st.run(nsteps); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / nsteps)
    n = 1
    for s in shape: n *= s
    nb = {'fenton4v': 32, 'mms2v': 16, 'ms2v': 16, None: 8}[model] + (4 if lap == 'heterog' else 0)
    print('%-14s %-9s %-5s %-7s %-7s chunk=%-3d  %.3f ms/step  %.1f Gnode-steps/s  %.0f GB/s algorithmic (%.1f%% of 6536)' % (
        'x'.join(map(str, shape)), model, math, kernel, lap, x_chunk, best, n / best / 1e6, n * nb / best / 1e6, n * nb / best / 1e6 / 65.36), flush=True)

if __name__ == '__main__':
    torch.cuda.set_device(0)
    S = (512, 512, 512)
    for math in ('fast', 'exact'):
        bench(S, math=math)
    for chunk in (8, 16, 32, 64, 128):
        bench(S, x_chunk=chunk)
    bench(S, kernel='generic')
    bench(S, model='mms2v'); bench(S, model=None)
    bench(S, lap='heterog')
    bench((1024, 1024), nsteps=400)
    bench((256, 256, 256))
