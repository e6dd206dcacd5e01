"""x-chunk length and CTA rasterisation sweep of the fused FD kernel (scratch timing, CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pdensorflow_b200.fd import FDStepper
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
shapes = ((512, 512, 512), (128, 2048, 2048))
chunks = tuple(int(c) for c in (sys.argv[1].split(',') if len(sys.argv) > 1 else '6,8,12'.split(',')))
for G in shapes:
    for raster in ('0', '4', '16', '64'):
        os.environ['PDX_FD_RASTER'] = raster
        for chunk in chunks:
            st = FDStepper(G, (1., 1., 1.), 0.1, 0.8, x_chunk=chunk)
            st.U().fill_(-80.); st.U()[0:2] = 20.
            st.step(60)
            t = timeit(lambda: st.step(1))
            n = G[0] * G[1] * G[2]
            print('%s raster=%s chunk=%2d: %.4f ms/step %.1f Gn/s' % (G, raster, chunk, t, n / t / 1e6), flush=True)
            del st
