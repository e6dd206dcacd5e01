#!/usr/bin/env python
"""bench.py - Gnode-steps/s of the fused 3-D Fenton-Karma finite-difference monodomain step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A bench "step" is one frame interval of the reference examples: ``dt_per_plot`` = 10 explicit
time steps (Tests/FD/Fenton/fenton.py:176,158-159), replayed as one CUDA graph of 10 fused
kernel launches.  Workload (BASELINE.json): N=1 -> config 3, 512^3 fp32 grid; N>1 -> config 4
weak scaling, 128 planes of 2048x2048 per GPU (N=8 is the full 1024x2048x2048 grid), slab
decomposed along the slowest axis with one-plane NCCL halo exchange per time step.

Printed JSON (rank 0): value = whole-job node-steps/s with state resident in HBM; e2e = same
metric with the potential uploaded from pinned host memory before and read back after every
frame; roofline = algorithmic bytes (32 B per node-step) / kernel time vs the measured HBM
peak; cpu_baseline = the reference path's op-for-op restatement timed on this box's cores;
other_configs (N=1 only, after the contract measurement, child processes, --skip-extras to omit) = records of
BASELINE.json's other single-GPU configurations from the tools/ scripts (config 2, config 5 on one GPU, config 1,
large-system semi-implicit step) - they are not the metric.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

STEPS_PER_FRAME = 10          # dt_per_plot of the reference FD examples
BYTES_PER_NODE_STEP = 32      # read U,V,W,S + write U',V,W,S (fp32): SURVEY.md 8(d)
DT, DIFF, SPACING = 0.1, 0.8, (1.0, 1.0, 1.0)


def hbm_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for s in self.samples:
            parts = [p.strip() for p in s.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(pw) if pw else None, 'samples': len(sm), 'reasons': sorted(reasons)}


def workload(n_gpus):
    if n_gpus == 1:
        return {'name': 'config 3: 3-D FD Fenton-Karma 512^3, fp32, 1 GPU', 'global': (512, 512, 512)}
    return {'name': 'config 4 (weak scaling): 3-D FD Fenton-Karma, 128 planes of 2048x2048 per GPU, z-slab halo '
                    'exchange; N=8 is the full 2048x2048x1024 grid stored [1024,2048,2048]',
            'global': (128 * n_gpus, 2048, 2048)}


# ------------------------------------------------------------------------------------ ours
def run_ours(args):
    import torch
    import torch.distributed as dist
    from pdensorflow_b200 import _lib as L
    from pdensorflow_b200.fd import DistributedFDStepper, FDStepper

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        raise SystemExit('--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)' % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    wl = workload(world)
    G = wl['global']

    if world == 1:
        st = FDStepper(G, SPACING, DT, DIFF, model='fenton4v', math='fast')
        fd = st
        U = st.U(); U.fill_(-80.0); U[0:2] = 20.0

        def frame():
            st.run(STEPS_PER_FRAME, block=STEPS_PER_FRAME)
        nodes_local = G[0] * G[1] * G[2]
    else:
        ds = DistributedFDStepper(G, SPACING, DT, DIFF, rank=rank, world=world, model='fenton4v', math='fast')
        fd = ds.fd
        U = fd.U(); U.fill_(-80.0)
        if rank == 0:
            U[1:3] = 20.0           # global planes 0,1 (array planes 1,2 on rank 0)
        for s_, v in zip(fd.states, (1.0, 1.0, 0.0)):
            s_.fill_(v)

        def frame():
            ds.step(STEPS_PER_FRAME)
        nodes_local = ds.part.nloc * G[1] * G[2]
    nodes_global = G[0] * G[1] * G[2]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # let the S1 wave develop so that both branches of every tf.where are exercised
    for _ in range(max(args.warmup, 3)):
        frame()
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = fd.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        frame()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = int(fd.launches() - l0)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    node_steps = nodes_global * STEPS_PER_FRAME * args.steps
    value = node_steps / (ms * 1e-3) / 1e9

    # ---- e2e: host buffers in and out, every frame ------------------------------------------
    # N=1: three independent simulations of the same workload pipelined over the H2D / compute /
    # D2H engines (FramePipeline); every frame of every simulation uploads its potential from
    # pinned host memory and downloads it back, dependencies through CUDA events only.
    # N>1: one slab per rank, upload -> 10 steps -> download, serial (the slab loop is not graphed).
    ke = max(2, min(args.steps, 5))
    if world == 1:
        from pdensorflow_b200.fd import FramePipeline
        sims = [st]
        for _ in range(2):
            s2_ = FDStepper(G, SPACING, DT, DIFF, model='fenton4v', math='fast')
            s2_.set_U(st.U())
            s2_.set_states(st.states)
            sims.append(s2_)
        pipe = FramePipeline(sims, STEPS_PER_FRAME)
        pipe.frame()
        pipe.wait()
        barrier()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(ke):
            pipe.frame()
        pipe.wait()
        e1.record()
        barrier()
        ms_e = max(e0.elapsed_time(e1), 0.0)
        ms_wall = (time.perf_counter() - t0) * 1e3
        ms_e = max(ms_e, ms_wall)            # streams other than the current one: wall clock bounds it
        nsim = len(sims)
        e2e_value = nsim * nodes_global * STEPS_PER_FRAME * ke / (ms_e * 1e-3) / 1e9
        frame_bytes = pipe.frame_bytes * nsim
        e2e_what = ('per bench step: %d independent simulations x (pinned-host U -> device, 10 fused steps as one CUDA '
                    'graph, device U -> pinned host), H2D / compute / D2H streams overlapped across simulations' % nsim)
    else:
        hostU = torch.empty(fd.array_shape, dtype=torch.float32, pin_memory=True)
        hostU.copy_(fd.U())
        barrier()

        def frame_e2e():
            fd.current_U().copy_(hostU, non_blocking=True)       # H2D of this frame's input potential
            frame()
            hostU.copy_(fd.current_U(), non_blocking=True)        # D2H of the frame (the writer's snapshot)
        frame_e2e()
        barrier()
        e0.record()
        for _ in range(ke):
            frame_e2e()
        e1.record()
        barrier()
        ms_e = e0.elapsed_time(e1)
        t = torch.tensor([ms_e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e = float(t.item())
        e2e_value = nodes_global * STEPS_PER_FRAME * ke / (ms_e * 1e-3) / 1e9
        frame_bytes = hostU.numel() * 4 * world
        e2e_what = 'per frame (10 time steps) and rank: pinned-host U slab -> device, 10 fused steps, device -> pinned host'

    out = None
    if rank == 0:
        peak, peak_src = hbm_peak()
        # dominant kernel: fd_tma_kernel, one launch per time step per rank
        launches_per_rank = STEPS_PER_FRAME * args.steps * (1 if world == 1 else ds.launches_per_step())
        step_ms = ms / (STEPS_PER_FRAME * args.steps)
        achieved = nodes_local * BYTES_PER_NODE_STEP / (step_ms * 1e-3) / 1e9
        tr = ncu_traffic() if world == 1 else None      # the committed capture is the 512^3 launch
        out = {
            'metric': 'Gnode-steps/s for 3D Fenton-Karma monodomain (fused FD step)',
            'value': value, 'unit': 'Gnode-steps/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': wl['name'], 'global_grid': list(G), 'nodes': nodes_global,
                       'time_steps_per_bench_step': STEPS_PER_FRAME, 'ionic_model': 'fenton4v', 'dt': DT, 'diff': DIFF,
                       'arithmetic': 'FAST (rel-L2 <= 1e-5 vs oracle, tests/test_fd_gpu.py)', 'kernel': fd.kernel,
                       'cache': 'inputs larger than L2: %.2f GB of state per GPU vs 126 MB L2' %
                                (nodes_local * 20 / 1e9),
                       'parallelism': 'single GPU' if world == 1 else
                       ('x-slab x%d, halo planes stored into the neighbours\' ghost planes by the step kernel over NVLink '
                        'peer memory, one symmetric-memory barrier per step' % world if ds.halo == 'peer' else
                        'x-slab x%d, 1-plane NCCL halo per step' % world)},
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': 'Gnode-steps/s', 'h2d_bytes_per_step': frame_bytes,
                    'd2h_bytes_per_step': frame_bytes, 'steps': ke,
                    'what': e2e_what},
            'gpu_launches': launches,
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'frac_of_nominal_8000_gbs': achieved / 8000.0,
                         'traffic': tr.get('dram_bytes_per_launch') if tr else None,
                         'kernel': 'fd_tma_kernel<FENTON4V,homog,FAST,3D,TMA-staged states>', 'peak_source': peak_src,
                         'algorithmic_bytes_per_launch': nodes_local * BYTES_PER_NODE_STEP,
                         'launches_timed_per_rank': launches_per_rank, 'avg_step_ms': step_ms},
        }
        if world == 1 and not args.skip_cpu:
            out['cpu_baseline'] = cpu_baseline(sample_n=256, nsteps=3)
        if world == 1 and not args.skip_extras:
            try:
                del st, fd
                torch.cuda.empty_cache()
                out['other_configs'] = other_configs()
            except Exception as e:          # records only: never let them break the contract line
                out['other_configs'] = [{'error': repr(e)[:200]}]
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


# ------------------------------------------------------------------------------ other configs
def other_configs():
    """Throughput of BASELINE.json's other single-GPU configurations, measured by the tools/ scripts in child
    processes AFTER the contract measurement (they are records, not the metric): config 2 (2-D 1024^2, per-step graph
    vs SM-resident kernel), config 5 on one GPU (2e7-node tetrahedral mesh, SELL-32 explicit lumped step), config 1
    (63 001-node semi-implicit example) and the large-system semi-implicit step (partitioned PCG, world = 1).
    A failing or slow child only leaves its entry out."""
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    runs = [('tools/bench_config2.py', ['--steps', '1000'], 60), ('tools/bench_fem.py', [], 110),
            ('tools/bench_fem_semi.py', ['--size', '160'], 60)]      # typical: 10 + 35 + 15 s
    got = []
    for script, extra, limit in runs:
        try:
            r = subprocess.run([sys.executable, os.path.join(here, script)] + extra, capture_output=True, text=True,
                               timeout=limit, cwd=here)
            for line in r.stdout.splitlines():
                if line.startswith('{'):
                    d = json.loads(line)
                    d['tool'] = script
                    got.append(d)
            if r.returncode != 0:
                got.append({'tool': script, 'error': (r.stderr or '').strip().splitlines()[-1:] or ['exit %d' % r.returncode]})
        except Exception as e:      # timeout, missing file, bad JSON: a record less, never a failed bench
            got.append({'tool': script, 'error': repr(e)[:200]})
    return got


# ------------------------------------------------------------------------------ CPU baseline
def cpu_baseline(sample_n=256, nsteps=3):
    """The reference path (TF CPU) restated op for op in torch-CPU (oracle/torch_cpu.py), all
    host threads, on a bounded sample of the same workload."""
    from oracle import torch_cpu
    shape = (sample_n, sample_n, sample_n)
    sec, threads, _ = torch_cpu.time_fd(shape, nsteps, DT, DIFF, SPACING)
    v = sample_n ** 3 * nsteps / sec / 1e9
    return {'value': v, 'unit': 'Gnode-steps/s', 'cores': threads, 'kind': 'port',
            'sample': '%d^3 sub-grid of the 512^3 workload, %d time steps of the S1 plane-wave problem, unfused '
                      'op-for-op restatement of the TF path in torch-CPU (TensorFlow is not installable in this '
                      'image)' % (sample_n, nsteps),
            'seconds': sec}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (its restatement:
    the reference is Python-on-TensorFlow and TF cannot be installed here), rank 0 only."""
    if int(os.environ.get('RANK', '0')) != 0:
        return None
    from oracle import torch_cpu
    import torch
    n = 256
    shape = (n, n, n)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    U = torch.full(shape, -80.0, dtype=torch.float32)
    U[0:2] = 20.0
    m = torch_cpu.Fenton4v(DT)
    m.initialize_state_variables(U)
    for _ in range(max(1, min(args.warmup, 3))):
        U = torch_cpu.fd_step(m, U, DT, DIFF, SPACING)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        U = torch_cpu.fd_step(m, U, DT, DIFF, SPACING)
    sec = time.perf_counter() - t0
    v = n ** 3 * args.steps / sec / 1e9
    wl = workload(args.gpus)
    sample = ('each bench step = ONE explicit time step on a %d^3 sub-grid of the workload (bounded sample); unfused '
              'op-for-op torch-CPU restatement of Tests/FD/Fenton/fenton.py solve() (TensorFlow not installable here)' % n)
    return {'impl': 'reference', 'metric': 'Gnode-steps/s for 3D Fenton-Karma monodomain (fused FD step)', 'value': v,
            'unit': 'Gnode-steps/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': sec / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': wl['name'], 'global_grid': list(wl['global']), 'sample_grid': list(shape),
                       'time_steps_per_bench_step': 1, 'ionic_model': 'fenton4v', 'dt': DT, 'diff': DIFF},
            'cpu_baseline': {'value': v, 'unit': 'Gnode-steps/s', 'cores': threads, 'kind': 'port', 'sample': sample},
            'e2e': {'value': v, 'unit': 'Gnode-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--skip-cpu', action='store_true', help='omit the cpu_baseline leg (profiling runs)')
    ap.add_argument('--skip-extras', action='store_true', help='omit the other_configs records (tools/ child processes)')
    args = ap.parse_args()
    out = run_reference(args) if args.impl == 'reference' else run_ours(args)
    if out is not None:
        print(json.dumps(out), flush=True)


if __name__ == '__main__':
    main()
