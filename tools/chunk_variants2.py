"""x_chunk sweeps (2): y-tiles marched per CTA on 2-D grids, planes per CTA of the anisotropic tile kernel.
    python tools/chunk_variants2.py  ->  profiles/r02_chunk_variants2.jsonl"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from pdensorflow_b200.fd import FDStepper
from bench_variants import aniso_fields


def run(shape, nsteps, **kw):
    st = FDStepper(shape, (1.0,) * len(shape), 0.1, 0.8, math='fast', **kw)
    U = st.U(); U.fill_(-80.0); U[0:2] = 20.0
    st.step(50); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); st.step(nsteps); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / nsteps)
    n = 1
    for v in shape:
        n *= v
    return round(best, 4), round(n / best / 1e6, 1)


for shape in ((4096, 4096), (2048, 2048)):
    for chunk in (0, 4, 8, 16, 32, 64):
        ms, g = run(shape, 200, model='fenton4v', x_chunk=chunk, kernel='tma')
        print(json.dumps({'grid': list(shape), 'model': 'fenton4v', 'x_chunk_y_tiles': chunk, 'ms': ms, 'gnode_steps_per_s': g}), flush=True)
A = (256, 512, 512)
DIFF, AVEC = aniso_fields(A, 0.8)
for model in ('fenton4v', None):
    for chunk in (0, 8, 16, 32, 64, 128):
        ms, g = run(A, 10, model=model, lap='aniso', DIFF=DIFF, AVEC=AVEC, x_chunk=chunk)
        print(json.dumps({'grid': list(A), 'model': model or 'heat', 'lap': 'aniso', 'x_chunk': chunk, 'ms': ms, 'gnode_steps_per_s': g}), flush=True)
