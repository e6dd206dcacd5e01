"""Throughput of every fused-step variant (CUDA events, graph-free), one JSON line each ->
the table in DESIGN.md / profiles/.  Not the contract bench (bench.py is)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pdensorflow_b200.fd import FDStepper
from pdensorflow_b200.gpuSolve.ionic import CourtemancheRamirezNattel, TenTusscherPanfilov

PEAK = 6536.4
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                             'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    pass
BYTES = {'fenton4v': 32, 'mms2v': 16, 'ms2v': 16, None: 8, 'crn': 8 + 19 * 8, 'ttp': 8 + 18 * 8 + 4 * 4}


def aniso_fields(shape, diff):
    """sigma_t = diff/4, sigma_l = diff, fibres rotating in the (x,y) plane with z: a full tensor at every node."""
    DIFF = torch.empty(tuple(shape) + (2,), device='cuda')
    DIFF[..., 0] = 0.25 * diff
    DIFF[..., 1] = 0.75 * diff
    ang = torch.linspace(0.0, 3.0, shape[2], device='cuda').view(1, 1, -1).expand(*shape)
    a = torch.stack([torch.cos(ang), torch.sin(ang), 0.3 * torch.ones_like(ang)], dim=-1)
    a = a / a.norm(dim=-1, keepdim=True)
    AVEC = torch.stack([a[..., 0] * a[..., 0], a[..., 0] * a[..., 1], a[..., 0] * a[..., 2], a[..., 1] * a[..., 1],
                        a[..., 1] * a[..., 2], a[..., 2] * a[..., 2]], dim=-1).contiguous()
    return DIFF, AVEC


def bench(shape, model='fenton4v', math='fast', kernel='auto', lap='homog', nsteps=40, reps=3, dt=0.1, diff=0.8):
    name = model
    mobj = model
    u_rest = -80.0
    if model == 'crn':
        mobj, u_rest, dt, diff = CourtemancheRamirezNattel(), -81.2, 0.02, 0.1
    elif model == 'ttp':
        mobj, u_rest, dt, diff = TenTusscherPanfilov(), -86.2, 0.02, 0.1
    DIFF = torch.full(shape, diff, device='cuda') if lap == 'heterog' else None
    AVEC = None
    if lap == 'aniso':
        DIFF, AVEC = aniso_fields(shape, diff)
    st = FDStepper(shape, (1.0,) * len(shape), dt, diff, model=mobj, math=math, kernel=kernel, lap=lap, DIFF=DIFF, AVEC=AVEC)
    U = st.U()
    U.fill_(u_rest)
    U[0:2] = 20.0
    st.step(20 if name in ('crn', 'ttp') or lap == 'aniso' else 200)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st.step(nsteps)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / nsteps)
    n = int(np.prod(shape))
    # aniso: the tile kernel reads the six tensor components the plan formed once (24 B); the generic kernel reads DIFF
    # [.,2] + AVEC [.,6] as the reference passes them (32 B)
    tile = lap == 'aniso' and kernel != 'generic' and os.environ.get('PDX_ANISO_TILE', '1') != '0'
    nb = BYTES[name] + (4 if lap == 'heterog' else 0) + ((24 if tile else 32) if lap == 'aniso' else 0)
    assert bool(torch.isfinite(st.U()).all())
    print(json.dumps({'grid': list(shape), 'model': name or 'heat', 'lap': lap, 'math': math, 'kernel': ('aniso-tile' if tile else st.kernel),
                      'ms_per_step': round(best, 4), 'gnode_steps_per_s': round(n / best / 1e6, 2),
                      'bytes_per_node_step': nb, 'algorithmic_gbs': round(n * nb / best / 1e6, 1),
                      'frac_of_measured_hbm_peak': round(n * nb / best / 1e6 / PEAK, 3)}), flush=True)


if __name__ == '__main__':
    torch.cuda.set_device(0)
    S = (512, 512, 512)
    if '--lut' in sys.argv:            # only the lookup-table models, both arithmetic policies
        for m in ('crn', 'ttp'):
            for mm in ('fast', 'exact'):
                bench((256, 256, 256), model=m, nsteps=10, math=mm)
        sys.exit(0)
    if '--aniso' in sys.argv:          # 19-point fibre-tensor operator: tile kernel vs the one-thread-per-node kernel
        A = (256, 512, 512)
        bench(A, lap='aniso', nsteps=10)
        bench(A, lap='aniso', nsteps=10, math='exact')
        bench(A, lap='aniso', nsteps=10, model=None)
        bench(A, lap='aniso', nsteps=10, model='mms2v')
        bench(A, lap='aniso', nsteps=4, kernel='generic')
        bench(A, lap='aniso', nsteps=4, kernel='generic', math='exact')
        sys.exit(0)
    if '--odd' in sys.argv:            # nz % 4 != 0: row-wise TMA variant of the tile kernel vs the one-thread-per-node kernel
        O = (512, 512, 510)
        bench(O, kernel='tma')
        bench(O, kernel='generic')
        bench(O, kernel='tma', lap='heterog')
        bench(O, kernel='tma', model='mms2v')
        bench((4096, 4090), nsteps=100, kernel='tma')
        bench(S)                        # the bench kernel itself, same run
        sys.exit(0)
    if '--2d' in sys.argv:             # 2-D grids on the per-step TMA kernel (y-march) and the generic kernel
        for n2 in (4096, 2048, 1024):
            bench((n2, n2), nsteps=200, kernel='tma')
        bench((4096, 4096), nsteps=100, kernel='tma', lap='heterog')
        bench((4096, 4096), nsteps=100, kernel='tma', model='mms2v')
        bench((4096, 4096), nsteps=100, kernel='generic')
        sys.exit(0)
    bench(S)
    bench(S, math='exact')
    bench(S, kernel='generic')
    bench(S, lap='heterog')
    bench(S, model='mms2v')
    bench(S, model='ms2v')
    bench(S, model=None)
    bench((1024, 1024), nsteps=400)
    bench((4096, 4096), nsteps=100)
    bench((256, 256, 256), model='crn', nsteps=10)
    bench((256, 256, 256), model='ttp', nsteps=10)
    bench((256, 256, 256), model='crn', nsteps=10, math='exact')
    bench((256, 256, 256), model='ttp', nsteps=10, math='exact')
