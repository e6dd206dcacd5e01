#!/usr/bin/env python
"""bench.py - Gnode-steps/s of the fused 3-D Fenton-Karma finite-difference monodomain step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload fd|fem5|fem1|fd2d]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A bench "step" is one frame interval of the reference examples: ``dt_per_plot`` = 10 explicit
time steps (Tests/FD/Fenton/fenton.py:176,158-159), replayed as one CUDA graph of 10 fused
kernel launches.  Workload (BASELINE.json): N=1 -> config 3, 512^3 fp32 grid; N>1 -> config 4
weak scaling, 128 planes of 2048x2048 per GPU (N=8 is the full 1024x2048x2048 grid), slab
decomposed along the slowest axis; halo planes are stored by the step kernel straight into the
neighbours' ghost planes over NVLink peer memory, one symmetric-memory barrier per time step.

Printed JSON (rank 0):
  value        whole-job node-steps/s over EXACTLY K bench steps with state resident in HBM
  sustained    the same loop continued until >= 2 s have been timed (clock samples span both)
  e2e          same metric for ONE simulation whose potential lives in pinned HOST memory between frames:
               chunked H2D -> time-skewed x-march -> chunked D2H on three streams (fd.FrameStreamer)
  roofline     algorithmic bytes (32 B per node-step) / kernel time vs the measured HBM peak; traffic = ncu
               dram bytes per launch of the same kernel on the same per-GPU shape (profiles/traffic.json)
  parity       in-run check at EVERY N: the same stepper (same kernel, same halo path) restarted from a planar
               multi-front state and compared with the NumPy oracle on an NX x 4 x 4 grid; N>1 also compares a
               small slab-decomposed run bitwise with the same grid advanced on one GPU
  same_shape_n1 (N=1 only) the per-GPU shape of the multi-GPU runs (128 x 2048 x 2048) timed on one GPU: the
               honest denominator of the weak-scaling efficiency
  cpu_baseline the reference path's op-for-op restatement timed on this box's cores
  other_configs records of BASELINE.json's other configurations (child processes of tools/, --skip-extras omits)
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
SLAB = (128, 2048, 2048)      # per-GPU shape of config 4 (weak scaling)
PARITY_STEPS = 200
PARITY_TOL = 1e-5             # north star: rel-L2 of the potential vs the reference path (FAST arithmetic)
SUSTAIN_SECONDS = 2.0


def hbm_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic(shape):
    """DRAM bytes per launch of the dominant kernel on this per-GPU shape from the committed ncu capture."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            d = json.load(f)
        return d.get('shapes', {}).get('x'.join(str(v) for v in shape))
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


def workload(n_gpus, strong=False):
    if strong:
        return {'name': 'config 4 (STRONG scaling): 3-D FD Fenton-Karma, the full 2048x2048x1024 grid stored [1024,2048,2048] '
                        '(86 GB of state) split into %d x-slab(s)' % n_gpus, 'global': (1024, SLAB[1], SLAB[2])}
    if n_gpus == 1:
        return {'name': 'config 3: 3-D FD Fenton-Karma 512^3, fp32, 1 GPU', 'global': (512, 512, 512)}
    return {'name': 'config 4 (weak scaling): 3-D FD Fenton-Karma, 128 planes of 2048x2048 per GPU, z-slab halo '
                    'exchange; N=8 is the full 2048x2048x1024 grid stored [1024,2048,2048]',
            'global': (SLAB[0] * n_gpus, SLAB[1], SLAB[2])}


# ------------------------------------------------------------------------------------ parity
def parity_fronts(nx, world):
    """Planes excited in the parity run: the S1 planes of the examples plus six planes straddling every slab
    boundary, so that steep fronts cross each halo from the first step on."""
    planes = [0, 1]
    base, rem = divmod(nx, world)
    for r in range(1, world):
        b = r * base + min(r, rem)
        planes += list(range(b - 3, b + 3))
    return sorted(set(p for p in planes if 0 <= p < nx))


def parity_check(fd, ds, G, rank, world, dev):
    """Restart the timed stepper from a planar multi-front state, advance PARITY_STEPS steps through the very
    launch path that was timed, and compare the x-column with the oracle on an NX x 4 x 4 grid (a field that
    depends on x only stays planar, so every column of the big grid must equal the small one)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    fronts = parity_fronts(G[0], world)
    U = fd.current_U()
    U.fill_(-80.0)
    off = 0 if ds is None else 1 - ds.part.x0           # array plane of global plane g is g + off
    for g in fronts:
        a = g + off
        if 0 <= a < U.shape[0]:
            U[a] = 20.0
    for s_, v in zip(fd.states, (1.0, 1.0, 0.0)):
        s_.fill_(v)
    if ds is not None:
        ds._primed = False
        ds.step(PARITY_STEPS)
        own = ds.owned_U()
    else:
        fd.run(PARITY_STEPS, block=STEPS_PER_FRAME)
        own = fd.U()
    torch.cuda.synchronize()
    planar = bool((own == own[:, :1, :1]).all())
    col = own[:, G[1] // 3, G[2] // 5].contiguous()
    if world > 1:
        nmax = -(-G[0] // world)
        pad = torch.zeros(nmax, device=dev)
        pad[:col.numel()] = col
        cols = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(cols, pad)
        flags = torch.tensor([int(planar)], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        planar = bool(flags.item())
        from pdensorflow_b200.fd import SlabPartition
        col = torch.cat([c[:SlabPartition(G[0], r, world).nloc] for r, c in enumerate(cols)])
    out = None
    if rank == 0:
        from oracle import fd as ofd
        from oracle import ionic as oionic
        small = (G[0], 4, 4)
        Uo = np.full(small, -80.0, dtype=np.float32)
        Uo[fronts] = 20.0
        m = oionic.Fenton4v()
        m._dt = DT
        m.initialize_state_variables(Uo)
        Uo = ofd.fd_run(m, Uo, PARITY_STEPS, DT, DIFF, SPACING)
        ref = Uo[:, 0, 0].astype(np.float64)
        got = col.cpu().numpy().astype(np.float64)
        err = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
        out = {'rel_l2': err, 'n_steps': PARITY_STEPS, 'tol': PARITY_TOL, 'ok': bool(err <= PARITY_TOL and planar),
               'planar': planar, 'excited_planes': fronts, 'umax': float(got.max()),
               'against': 'oracle.fd.fd_run (NumPy fp32 restatement pinned to the reference source) on a %dx4x4 grid, '
                          'same initial state; column (y,z)=(%d,%d) of the %s grid advanced by the timed launch path'
                          % (G[0], G[1] // 3, G[2] // 5, 'x'.join(map(str, G)))}
    return out


def dist_bitwise_check(rank, world, dev):
    """N>1: a small random-field grid advanced by the slab-decomposed stepper (peer-memory halos) must equal,
    BITWISE, the same grid advanced on one GPU (EXACT arithmetic); rank 0 reports."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from pdensorflow_b200.fd import DistributedFDStepper, FDStepper
    shape, nsteps = (9 * world + 5, 40, 136), 24
    rng = np.random.default_rng(11)
    U = rng.uniform(-85.0, 25.0, size=shape).astype(np.float32)
    states = [rng.uniform(0.0, 1.0, size=shape).astype(np.float32) for _ in range(3)]
    ds = DistributedFDStepper(shape, SPACING, DT, DIFF, rank=rank, world=world, model='fenton4v', math='exact')
    ds.set_global(U, states)
    ds.step(nsteps)
    torch.cuda.synchronize()
    own = ds.owned_U().contiguous()
    nmax = -(-shape[0] // world)
    pad = torch.zeros((nmax,) + shape[1:], device=dev)
    pad[:own.shape[0]] = own
    parts = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    out = None
    if rank == 0:
        from pdensorflow_b200.fd import SlabPartition
        full = torch.cat([p[:SlabPartition(shape[0], r, world).nloc] for r, p in enumerate(parts)])
        one = FDStepper(shape, SPACING, DT, DIFF, model='fenton4v', math='exact')
        one.set_U(U)
        one.set_states(states)
        one.step(nsteps)
        torch.cuda.synchronize()
        out = {'grid': list(shape), 'n_steps': nsteps, 'halo': ds.halo, 'bitwise_equal_to_one_gpu': bool(torch.equal(full, one.U())),
               'max_abs_diff': float((full - one.U()).abs().max())}
    del ds
    return out


# ------------------------------------------------------------------------------------ ours
def time_frames(frame, nframes, barrier):
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(nframes):
        frame()
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from pdensorflow_b200.fd import DistributedFDStepper, FDStepper, FrameStreamer

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
    wl = workload(world, args.strong)
    G = wl['global']

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ds = None
    if world == 1:
        st = FDStepper(G, SPACING, DT, DIFF, model='fenton4v', math='fast')
        fd = st
        U = st.U(); U.fill_(-80.0); U[0:2] = 20.0

        def frame():
            st.run(STEPS_PER_FRAME, block=STEPS_PER_FRAME)
        nodes_local = G[0] * G[1] * G[2]
        local_shape = G
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
        local_shape = (ds.part.nloc, G[1], G[2])
    nodes_global = G[0] * G[1] * G[2]

    if args.verify:            # correctness only (cheap multi-GPU lease): no timing
        out = {'verify': True, 'n_gpus': world, 'parity': parity_check(fd, ds, G, rank, world, dev)}
        if world > 1:
            out['distributed'] = dist_bitwise_check(rank, world, dev)
            dist.barrier()
            dist.destroy_process_group()
        return out if rank == 0 else None

    # let the S1 wave develop so that both branches of every tf.where are exercised
    for _ in range(max(args.warmup, 3)):
        frame()
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = fd.launches()
    ms = allmax(time_frames(frame, args.steps, barrier))
    launches = int(fd.launches() - l0)
    # the same loop continued: >= SUSTAIN_SECONDS of timed device work in total, clock samples span both regions
    extra = max(0, int(SUSTAIN_SECONDS * 1e3 / max(ms / args.steps, 1e-3)) + 1 - args.steps)
    extra = min(extra, 2000)
    ms_s = allmax(time_frames(frame, extra, barrier)) if extra else 0.0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    node_steps = nodes_global * STEPS_PER_FRAME * args.steps
    value = node_steps / (ms * 1e-3) / 1e9
    sustained = {'frames': args.steps + extra, 'seconds': (ms + ms_s) * 1e-3,
                 'value': nodes_global * STEPS_PER_FRAME * (args.steps + extra) / ((ms + ms_s) * 1e-3) / 1e9,
                 'unit': 'Gnode-steps/s'}

    # ---- e2e: ONE simulation, potential in pinned host memory between frames ------------------------------
    ke = max(3, min(args.steps, 10))
    e2e_value = frame_bytes = e2e_launches = None
    e2e_what = 'skipped (--strong: the 17 GB frames are not part of the strong-scaling record)'
    if args.strong:
        pass
    elif world == 1:
        streamer = FrameStreamer(st, STEPS_PER_FRAME, nchunks=8, skew=True)
        e2e_what = ('ONE simulation; per frame (10 time steps): pinned-host U -> device in 8 plane chunks, time-skewed '
                    'x-march (80 range launches of the fused kernel) behind the arriving chunks, finished planes -> '
                    'pinned host; H2D / compute / D2H on three streams, a chunk is re-uploaded only after its download')
    else:
        streamer = FrameStreamer(fd, STEPS_PER_FRAME, nchunks=8, skew=False, advance=ds.step,
                                 planes=(ds.part.x_begin, ds.part.x_end))
        e2e_what = ('ONE simulation; per frame (10 time steps) and rank: pinned-host slab -> device in 8 plane chunks, '
                    '10 whole-slab fused steps (peer-memory halos), slab -> pinned host in 8 chunks; the next frame\'s '
                    'upload of a chunk starts when its download is done (PCIe duplex), NUMA-local pinned buffers')
    ms_e = None
    if not args.strong:
        for _ in range(2):
            streamer.frame()
        streamer.wait()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            streamer.frame()
        streamer.wait()
        torch.cuda.synchronize()
        ms_e = allmax((time.perf_counter() - t0) * 1e3)      # three streams: the host clock around a full drain bounds it
        barrier()
        e2e_value = nodes_global * STEPS_PER_FRAME * ke / (ms_e * 1e-3) / 1e9
        frame_bytes = streamer.frame_bytes
        if world > 1:
            fb = torch.tensor([frame_bytes], device=dev, dtype=torch.int64)
            dist.all_reduce(fb, op=dist.ReduceOp.SUM)
            frame_bytes = int(fb.item())
        e2e_launches = streamer.launches_per_frame
        del streamer

    # ---- parity of the timed launch path (every N) -----------------------------------------------------------
    parity = parity_check(fd, ds, G, rank, world, dev)
    dist_check = dist_bitwise_check(rank, world, dev) if world > 1 else None

    out = None
    if rank == 0:
        peak, peak_src = hbm_peak()
        # dominant kernel: fd_tma_kernel, one launch per time step per rank
        launches_per_rank = STEPS_PER_FRAME * args.steps * (1 if world == 1 else ds.launches_per_step())
        step_ms = ms / (STEPS_PER_FRAME * args.steps)
        achieved = nodes_local * BYTES_PER_NODE_STEP / (step_ms * 1e-3) / 1e9
        tr = ncu_traffic(local_shape)
        out = {
            'metric': 'Gnode-steps/s for 3D Fenton-Karma monodomain (fused FD step)',
            'value': value, 'unit': 'Gnode-steps/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'strong' if args.strong else 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': wl['name'], 'global_grid': list(G), 'nodes': nodes_global,
                       'time_steps_per_bench_step': STEPS_PER_FRAME, 'ionic_model': 'fenton4v', 'dt': DT, 'diff': DIFF,
                       'arithmetic': 'FAST (rel-L2 <= 1e-5 vs oracle: tests/test_fd_gpu.py and the parity key of this line)',
                       'kernel': fd.kernel,
                       'cache': 'inputs larger than L2: %.2f GB of state per GPU vs 126 MB L2' %
                                (nodes_local * 20 / 1e9),
                       'parallelism': 'single GPU' if world == 1 else
                       ('x-slab x%d, halo planes stored into the neighbours\' ghost planes by the step kernel over NVLink '
                        'peer memory, one symmetric-memory barrier per step' % world if ds.halo == 'peer' else
                        'x-slab x%d, 1-plane NCCL halo per step' % world)},
            'clocks': clocks,
            'sustained': sustained,
            'e2e': {'value': e2e_value, 'unit': 'Gnode-steps/s', 'h2d_bytes_per_step': frame_bytes,
                    'd2h_bytes_per_step': frame_bytes, 'steps': ke, 'simulations': 1,
                    'host_link_gbs_per_direction_all_gpus': (frame_bytes * ke / (ms_e * 1e-3) / 1e9) if ms_e else None,
                    'kernel_launches_per_step_per_rank': e2e_launches, 'what': e2e_what},
            'gpu_launches': launches,
            'parity': parity,
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'frac_of_nominal_8000_gbs': achieved / 8000.0,
                         'traffic': tr.get('dram_bytes_per_launch') if tr else None,
                         'traffic_source': tr.get('source') if tr else None,
                         'kernel': 'fd_tma_kernel<FENTON4V,homog,FAST,3D,TMA-staged states>', 'peak_source': peak_src,
                         'per_gpu_shape': list(local_shape),
                         'algorithmic_bytes_per_launch': nodes_local * BYTES_PER_NODE_STEP,
                         'launches_timed_per_rank': launches_per_rank, 'avg_step_ms': step_ms},
        }
        if dist_check is not None:
            out['parity']['distributed'] = dist_check
    if world == 1 and not args.strong:
        try:
            out['same_shape_n1'] = same_shape_n1(args)
        except Exception as e:
            out['same_shape_n1'] = {'error': repr(e)[:200]}
        torch.cuda.empty_cache()
        if not args.skip_cpu:
            out['cpu_baseline'] = cpu_baseline(sample_n=256, nsteps=30)
        if not args.skip_extras:
            try:
                out['other_configs'] = other_configs(world)
            except Exception as e:          # records only: never let them break the contract line
                out['other_configs'] = [{'error': repr(e)[:200]}]
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
        if rank == 0 and not args.skip_extras and not args.strong:
            # config 5 (row-block partitioned explicit FEM step) on the same N GPUs, after the process group of the
            # contract measurement is gone: its own torchrun, a record only
            try:
                out['other_configs'] = other_configs(world)
            except Exception as e:
                out['other_configs'] = [{'error': repr(e)[:200]}]
    return out


def same_shape_n1(args):
    """The per-GPU shape of the multi-GPU runs on ONE GPU: 128 planes of 2048 x 2048 as a stand-alone grid."""
    import torch
    from pdensorflow_b200.fd import FDStepper
    free, _ = torch.cuda.mem_get_info()
    n = SLAB[0] * SLAB[1] * SLAB[2]
    if free < 5.5 * n * 4:
        return {'skipped': 'not enough free device memory'}
    st = FDStepper(SLAB, SPACING, DT, DIFF, model='fenton4v', math='fast')
    U = st.U(); U.fill_(-80.0); U[0:2] = 20.0

    def frame():
        st.run(STEPS_PER_FRAME, block=STEPS_PER_FRAME)
    for _ in range(3):
        frame()
    k = max(5, min(args.steps, 20))
    ms = time_frames(frame, k, torch.cuda.synchronize)
    step_ms = ms / (k * STEPS_PER_FRAME)
    peak, _ = hbm_peak()
    ach = n * BYTES_PER_NODE_STEP / (step_ms * 1e-3) / 1e9
    tr = ncu_traffic(SLAB)
    return {'grid': list(SLAB), 'value': n / (step_ms * 1e-3) / 1e9, 'unit': 'Gnode-steps/s', 'avg_step_ms': step_ms,
            'frames_timed': k, 'roofline_frac': ach / peak, 'traffic': tr.get('dram_bytes_per_launch') if tr else None,
            'algorithmic_bytes_per_launch': n * BYTES_PER_NODE_STEP,
            'note': 'divide the per-GPU value of an N-GPU line by this one for the cost of the slab decomposition itself'}


# ------------------------------------------------------------------------------ other configs
def other_configs(world=1):
    """Throughput of BASELINE.json's other configurations, measured by the tools/ scripts in child processes AFTER
    the contract measurement (they are records, not the metric).  N=1: config 2 (2-D 1024^2), config 5 on one GPU
    (2e7-node tetrahedral mesh, SELL-32 explicit lumped step), config 1 (63 001-node semi-implicit example), the
    large-system semi-implicit step, and the CPU legs of configs 1, 2, 5.  N>1: config 5 row-block partitioned over
    the same N GPUs (its own torchrun).  A failing or slow child only leaves its entry out."""
    here = ROOT
    if world == 1:
        runs = [([sys.executable, 'tools/bench_config2.py', '--steps', '1000'], 60),
                ([sys.executable, 'tools/bench_fem.py'], 110),
                ([sys.executable, 'tools/bench_fem_semi.py', '--size', '160'], 60),
                ([sys.executable, 'tools/cpu_baselines.py'], 120)]
    else:
        tr = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
              '--master-addr', '127.0.0.1', '--master-port', str(29610 + world)]
        runs = [(tr + ['tools/bench_fem.py'], 150)]
    got = []
    env = {k: v for k, v in os.environ.items()
           if not k.startswith('TORCHELASTIC') and k not in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE', 'MASTER_PORT', 'GROUP_RANK',
                                                             'ROLE_RANK', 'LOCAL_WORLD_SIZE', 'ROLE_WORLD_SIZE',
                                                             'GROUP_WORLD_SIZE', 'ROLE_NAME')}
    for cmd, limit in runs:
        tool = [c for c in cmd if c.startswith('tools/')][0]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=limit, cwd=here, env=env)
            for line in r.stdout.splitlines():
                if line.startswith('{'):
                    d = json.loads(line)
                    d['tool'] = tool
                    got.append(d)
            if r.returncode != 0:
                got.append({'tool': tool, 'error': (r.stderr or '').strip().splitlines()[-1:] or ['exit %d' % r.returncode]})
        except Exception as e:      # timeout, missing file, bad JSON: a record less, never a failed bench
            got.append({'tool': tool, 'error': repr(e)[:200]})
    return got


# ------------------------------------------------------------------------------ CPU baseline
def cpu_baseline(sample_n=256, nsteps=30):
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
    """--impl reference: the reference's CPU implementation of the path (its restatement: the reference is
    Python-on-TensorFlow and TF cannot be installed here), rank 0 only, on the FULL grid of the N=1 workload
    (512^3): each bench step is ONE explicit time step of it (ours: ten) - the metric is per node-step."""
    if int(os.environ.get('RANK', '0')) != 0:
        return None
    from oracle import torch_cpu
    import torch
    wl = workload(1)
    shape = wl['global']
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
    nodes = shape[0] * shape[1] * shape[2]
    v = nodes * args.steps / sec / 1e9
    sample = ('each bench step = ONE explicit time step of the full %dx%dx%d grid (the N=1 workload; bounded in time, '
              'not in space); unfused op-for-op torch-CPU restatement of Tests/FD/Fenton/fenton.py solve() '
              '(TensorFlow not installable here)' % shape)
    return {'impl': 'reference', 'metric': 'Gnode-steps/s for 3D Fenton-Karma monodomain (fused FD step)', 'value': v,
            'unit': 'Gnode-steps/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': sec / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': wl['name'], 'global_grid': list(shape), 'nodes': nodes,
                       'time_steps_per_bench_step': 1, 'ionic_model': 'fenton4v', 'dt': DT, 'diff': DIFF},
            'cpu_baseline': {'value': v, 'unit': 'Gnode-steps/s', 'cores': threads, 'kind': 'port', 'sample': sample},
            'e2e': {'value': v, 'unit': 'Gnode-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}


def run_tool_workload(args):
    """--workload fem5 | fem1 | fd2d: the tools/ scripts of BASELINE.json's other configurations under the bench's own
    launch convention (torchrun included), so that a scaling driver can exercise them.  Their JSON lines are printed
    as they are (records; the contract line is the default --workload fd)."""
    rank = int(os.environ.get('RANK', '0'))
    tool = {'fem5': ['tools/bench_fem.py', '--skip-config1'], 'fem1': ['tools/bench_fem.py', '--n', '64'],
            'fd2d': ['tools/bench_config2.py']}[args.workload]
    sys.argv = [tool[0]] + tool[1:]
    import runpy
    runpy.run_path(os.path.join(ROOT, tool[0]), run_name='__main__')
    return None if rank else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default=os.environ.get('PDX_BENCH_WORKLOAD', 'fd'), choices=['fd', 'fem5', 'fem1', 'fd2d'])
    ap.add_argument('--strong', action='store_true', help='strong scaling of config 4: the full 1024x2048x2048 grid at every N (no e2e / extras)')
    ap.add_argument('--verify', action='store_true', help='correctness checks only (parity + distributed bitwise), no timing')
    ap.add_argument('--skip-cpu', action='store_true', help='omit the cpu_baseline leg (profiling runs)')
    ap.add_argument('--skip-extras', action='store_true', help='omit the other_configs records (tools/ child processes)')
    args = ap.parse_args()
    if args.impl == 'reference':
        out = run_reference(args)
    elif args.workload != 'fd':
        out = run_tool_workload(args)
    else:
        out = run_ours(args)
    if out is not None:
        print(json.dumps(out), flush=True)


if __name__ == '__main__':
    main()
