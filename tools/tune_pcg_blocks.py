"""Grid-size sweep of the small-system persistent PCG kernel on config 1 (63 001 nodes, L2 resident, barrier bound).
    python tools/tune_pcg_blocks.py"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
for nb in (0, 148, 222, 296, 444, 592):
    env = dict(os.environ)
    if nb:
        env['PDX_PCG_BLOCKS'] = str(nb)
    r = subprocess.run([sys.executable, os.path.join(HERE, 'bench_fem.py'), '--size', '32'], env=env, capture_output=True, text=True)
    for line in r.stdout.splitlines():
        if 'semi-implicit' in line:
            d = json.loads(line)
            print(json.dumps({'pcg_blocks': nb or 'auto', 'ms_per_step': d['ms_per_step'], 'its': d['cg_iterations_per_step_sampled']}), flush=True)
