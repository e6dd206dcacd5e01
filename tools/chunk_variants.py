"""x_chunk sweep of the TMA kernel for the variants whose planes are cheap (heat), carry two staged fields (heterogeneous)
or one gate (mMS): python tools/chunk_variants.py  ->  profiles/r02_chunk_variants.jsonl"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pdensorflow_b200.fd import FDStepper
S=(512,512,512)
for model, lap in ((None,'homog'), ('fenton4v','heterog'), ('mms2v','homog')):
    for chunk in (0, 8, 12, 16, 24, 32):
        DIFF = torch.full(S, 0.8, device='cuda') if lap=='heterog' else None
        st = FDStepper(S, (1.0,1.0,1.0), 0.1, 0.8, model=model, math='fast', lap=lap, DIFF=DIFF, x_chunk=chunk)
        U = st.U(); U.fill_(-80.0); U[0:2] = 20.0
        st.step(100); torch.cuda.synchronize()
        best=1e9
        for _ in range(3):
            e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
            e0.record(); st.step(40); e1.record(); torch.cuda.synchronize()
            best=min(best,e0.elapsed_time(e1)/40)
        print(json.dumps({'model': model or 'heat', 'lap': lap, 'x_chunk': chunk, 'ms': round(best,4), 'gnode_steps_per_s': round(512**3/best/1e6,1)}), flush=True)
        del st, DIFF
