"""nchunks sweep of the single-GPU frame streamer (the e2e leg of bench.py): python tools/streamer_sweep.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pdensorflow_b200.fd import FDStepper, FrameStreamer

S = (512, 512, 512)
for nch in (4, 8, 16, 32):
    st = FDStepper(S, (1.0, 1.0, 1.0), 0.1, 0.8, model='fenton4v', math='fast')
    U = st.U(); U.fill_(-80.0); U[0:2] = 20.0
    fs = FrameStreamer(st, 10, nchunks=nch, skew=True)
    fs.host.copy_(st.U().cpu())
    for _ in range(3):
        fs.frame()
    fs.wait(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    k = 20
    for _ in range(k):
        fs.frame()
    fs.wait(); torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / k
    print(json.dumps({'grid': list(S), 'nchunks': nch, 'ms_per_frame': round(ms, 3), 'e2e_gnode_steps_per_s': round(512 ** 3 * 10 / ms / 1e6, 1),
                      'host_link_gbs_per_direction': round(fs.frame_bytes / ms / 1e6, 1)}), flush=True)
    del fs, st
