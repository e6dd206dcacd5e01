"""Fused explicit finite-difference monodomain stepping on B200 (host side).

``FDStepper`` owns the device buffers of one grid (or one slab of a distributed grid) and
drives ``pdx_fd_step``: one kernel launch per time step that fuses what the reference
examples' ``solve()`` spreads over ~150 TensorFlow ops (Tests/FD/Fenton/fenton.py:90-99:
enforce_boundary + differentiate + laplace + Euler update).  ``run()`` replays a
CUDA-graph-captured block of steps instead of the reference's Python per-step loop
(fenton.py:150-159).

``SlabPartition``/``DistributedFDStepper`` split the slowest axis (reference axis 0) over
the ranks of a torch.distributed group.  Default halo path: the step kernel itself stores each
slab's edge planes into the neighbours' ghost planes over NVLink peer memory (torch symmetric
memory), one symmetric-memory barrier per step orders the ranks; ``halo='nccl'`` (batched
send/recv on a side stream, overlapped with the interior planes) is the fallback.
``FrameStreamer`` is the end-to-end frame loop over HOST buffers (chunked H2D / skewed x-march /
chunked D2H, all three engines busy on ONE simulation).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L

_MODEL_IDS = {None: L.MODEL_NONE, 'none': L.MODEL_NONE, 'heat': L.MODEL_NONE,
              'fenton4v': L.MODEL_FENTON4V, 'mms2v': L.MODEL_MMS2V, 'ms2v': L.MODEL_MS2V}
_LAP_IDS = {'homog': L.LAP_HOMOG, 'conv': L.LAP_CONV, 'heterog': L.LAP_HETEROG, 'aniso': L.LAP_ANISO}
_MATH_IDS = {'exact': L.MATH_EXACT, 'fast': L.MATH_FAST}
_KERNEL_IDS = {'auto': L.KERNEL_AUTO, 'generic': L.KERNEL_GENERIC, 'tma': L.KERNEL_TMA, 'resident': L.KERNEL_RESIDENT}
STATE_INIT = {L.MODEL_FENTON4V: (1.0, 1.0, 0.0), L.MODEL_MMS2V: (1.0,), L.MODEL_MS2V: (1.0,), L.MODEL_NONE: ()}


class SlabPartition:
    """Contiguous split of the slowest axis (NX planes) over ``world`` ranks.

    Local arrays carry one ghost plane per side (array plane 0 and nloc+1); global plane g
    lives at array plane g - x0 + 1.  Stencil plane indices are clamped to the image of the
    global range [1, NX-2] (enforce_boundary + symmetric pad), which is a no-op on interior
    ranks whose ghost planes hold the neighbours' data.
    """

    def __init__(self, NX, rank=0, world=1):
        if world < 1 or not (0 <= rank < world):
            raise ValueError('bad rank/world')
        if NX < 3 * world:
            raise ValueError('need at least 3 planes per rank')
        base, rem = divmod(NX, world)
        self.NX, self.rank, self.world = NX, rank, world
        self.x0 = rank * base + min(rank, rem)
        self.nloc = base + (1 if rank < rem else 0)
        self.x1 = self.x0 + self.nloc
        self.nx_array = self.nloc + 2
        self.x_begin, self.x_end = 1, self.nloc + 1
        self.x_clamp_lo = max(1 - self.x0 + 1, 0)
        self.x_clamp_hi = min(NX - 2 - self.x0 + 1, self.nloc + 1)
        self.lo_rank = rank - 1 if rank > 0 else None
        self.hi_rank = rank + 1 if rank < world - 1 else None

    def scatter(self, G):
        """Local array (with ghost planes filled from the global field) of a global array."""
        loc = np.zeros((self.nx_array,) + tuple(G.shape[1:]), dtype=G.dtype)
        lo, hi = max(self.x0 - 1, 0), min(self.x1 + 1, self.NX)
        loc[lo - self.x0 + 1: hi - self.x0 + 1] = G[lo:hi]
        return loc

    def owned(self, loc):
        return loc[1:self.nloc + 1]


def exchange_halos(U, part, group=None):
    """Fill the ghost planes of local array ``U`` from the neighbouring ranks.

    Array plane 1 goes to the lower neighbour's ghost plane nloc+1, array plane nloc to the
    upper neighbour's ghost plane 0.  Works on CUDA (NCCL) and CPU (gloo) tensors.
    Returns the list of outstanding work handles (already waited for gloo).
    """
    import torch.distributed as dist
    ops = []
    if part.lo_rank is not None:
        ops.append(dist.P2POp(dist.isend, U[1], part.lo_rank, group))
        ops.append(dist.P2POp(dist.irecv, U[0], part.lo_rank, group))
    if part.hi_rank is not None:
        ops.append(dist.P2POp(dist.isend, U[part.nloc], part.hi_rank, group))
        ops.append(dist.P2POp(dist.irecv, U[part.nloc + 1], part.hi_rank, group))
    if not ops:
        return []
    reqs = dist.batch_isend_irecv(ops)
    for r in reqs:
        r.wait()          # NCCL: stream-orders the current stream after the transfer
    return reqs


class FDStepper:
    """One grid (or slab) advanced by the fused kernel.

    shape:    (W, H, D) or (W, H): the reference's array layout (last axis contiguous)
    spacing:  (DX, DY, DZ) or (DX, DY)
    dt, diff: Python floats as in the examples' config (fenton.py:170-183); the
              coefficient f32(diff*dt) is formed in float64 first, like the reference does
    model:    'fenton4v' | 'mms2v' | 'ms2v' | None (heat equation), or an IonicModel instance of
              pdensorflow_b200.gpuSolve.ionic - required for the lookup-table models
              (CourtemancheRamirezNattel, TenTusscherPanfilov), whose tables/constants/state
              arrays the instance owns
    lap:      'homog' | 'conv' | 'heterog' (DIFF [W,H,D] carries the diffusivity) |
              'aniso' (DIFF [W,H,D,2] and AVEC [W,H,D,6] as laplace_heterog_aniso takes them)
    dirichlet: None = the examples' no-flux boundary (symmetric pad); a float = the constant boundary value of
              Tests/FD/LaplaceSolver/laplaceSolver.py:44-52 (constant pad of the interior)
    """

    def __init__(self, shape, spacing, dt, diff=1.0, model='fenton4v', params=None, lap='homog', DIFF=None,
                 math='fast', kernel='auto', part=None, device=None, x_chunk=0, AVEC=None, buffers=None, dirichlet=None):
        self.device = torch.device('cuda' if device is None else device)
        if self.device.type != 'cuda':
            raise L.PdxError('FDStepper needs a CUDA device: there is no CPU path')
        if self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        # plans, tensor maps and kernel attributes belong to the CURRENT device: make the stepper's device
        # current for construction and for every call (a process may drive several GPUs)
        with torch.cuda.device(self.device):
            self._init(shape, spacing, dt, diff, model, params, lap, DIFF, math, kernel, part, x_chunk, AVEC, buffers,
                       dirichlet)

    def _stream(self, stream=None):
        return L.stream_ptr(torch.cuda.current_stream(self.device) if stream is None else stream)

    def _check_field(self, t, name, extra=()):
        """Pointers handed to the library are raw: refuse anything that is not a contiguous fp32 tensor of the
        stepper's array shape on the stepper's device (a bool / float64 / CPU / global-shape mask would be misread)."""
        L.require_cuda(t, name)
        if t.device != self.device:
            raise L.PdxError('%s lives on %s, the stepper on %s' % (name, t.device, self.device))
        if tuple(t.shape) != tuple(self.array_shape) + tuple(extra):
            raise L.PdxError('%s has shape %s, expected %s' % (name, tuple(t.shape), tuple(self.array_shape) + tuple(extra)))
        return t

    def _init(self, shape, spacing, dt, diff, model, params, lap, DIFF, math, kernel, part, x_chunk, AVEC, buffers,
              dirichlet):
        self.ndim = len(shape)
        if self.ndim not in (2, 3) or len(spacing) != self.ndim:
            raise ValueError('shape/spacing must both be 2-D or 3-D')
        self.part = part
        self.ionic = None
        if model is not None and not isinstance(model, str):
            self.ionic = model
            self.model_id = model._MODEL_ID
            if params is None and self.model_id not in (L.MODEL_CRN, L.MODEL_TTP):
                params = model.param_list()
        else:
            self.model_id = _MODEL_IDS[model]
        self.is_lut = self.model_id in (L.MODEL_CRN, L.MODEL_TTP)
        if self.is_lut and self.ionic is None:
            raise ValueError('lookup-table models are passed as IonicModel instances')
        self.lap_id = _LAP_IDS[lap]
        d = L.FdDesc()
        L.lib().pdx_fd_desc_init(C.byref(d))
        d.ndim = self.ndim
        if self.ndim == 3:
            nx = part.nx_array if part is not None else shape[0]
            d.nx, d.ny, d.nz = nx, shape[1], shape[2]
            d.dx, d.dy, d.dz = (float(np.float32(s)) for s in spacing)
            if part is not None:
                d.x_begin, d.x_end = part.x_begin, part.x_end
                d.x_clamp_lo, d.x_clamp_hi = part.x_clamp_lo, part.x_clamp_hi
        else:
            if part is not None:
                raise ValueError('slab partition is 3-D only')
            d.nx, d.ny, d.nz = 1, shape[0], shape[1]
            d.dy, d.dz = (float(np.float32(s)) for s in spacing)
        self.array_shape = (d.nx, d.ny, d.nz) if self.ndim == 3 else (d.ny, d.nz)
        # Rows that are not 16-byte multiples (nz % 4 != 0) cannot be tiled by tensor maps: the stepper then keeps ITS
        # arrays (U pair, gates, DIFF) with rows padded to a multiple of four floats (pdx_fd_desc.pitch_z) and hands out
        # [..., :nz] views, so the tile kernel runs at full speed; caller-owned buffers, the lookup-table models and
        # the anisotropic operator keep contiguous rows (one-thread-per-node kernel).
        self.pitch = d.nz
        if (d.nz % 4 != 0 and buffers is None and part is None and not self.is_lut and lap != 'aniso'
                and kernel not in ('generic', 'resident')):
            self.pitch = (d.nz + 3) // 4 * 4
            d.pitch_z = self.pitch
        d.dt = float(np.float32(dt))
        d.coef = float(np.float32(dt)) if lap in ('heterog', 'aniso') else float(np.float32(float(diff) * float(dt)))
        d.lap_kind, d.model = self.lap_id, self.model_id
        d.math_mode, d.kernel = _MATH_IDS[math], _KERNEL_IDS[kernel]
        d.x_chunk = int(x_chunk)
        if dirichlet is not None:
            d.dirichlet, d.dirichlet_value = 1, float(np.float32(dirichlet))
        if self.model_id != L.MODEL_NONE and not self.is_lut:
            p = L.default_params(self.model_id) if params is None else list(params)
            for i, v in enumerate(p):
                d.params[i] = float(v)
        self.desc = d
        handle = C.c_void_p()
        L.check(L.lib().pdx_fd_plan_create(C.byref(d), C.byref(handle)))
        self._plan = handle
        self.kernel = {L.KERNEL_GENERIC: 'generic', L.KERNEL_TMA: 'tma',
                       L.KERNEL_RESIDENT: 'resident'}[L.lib().pdx_fd_plan_kernel(handle)]
        # 2-D grids that fit in the SMs' shared memory: a multi-step call is ONE launch (pdx_fd_resident.cu)
        self.resident = bool(L.lib().pdx_fd_plan_resident_capable(handle))
        self.dt, self.diff = dt, diff
        self.nstates = L.lib().pdx_model_num_states(self.model_id)
        # buffers
        if buffers is not None:      # caller-owned ping-pong pair (e.g. NVLink symmetric memory)
            self.Ua, self.Ub = buffers
            for b in (self.Ua, self.Ub):
                if tuple(b.shape) != tuple(self.array_shape) or b.dtype != torch.float32 or not b.is_contiguous():
                    raise ValueError('buffers must be contiguous fp32 tensors of the array shape')
        else:
            self.Ua = self._new_field(0.0)
            self.Ub = self._new_field(0.0)
        if self.is_lut:
            # the model instance owns tables, constants and the 19/22 state arrays
            self.ionic._dt = dt
            self.ionic.initialize_state_variables(self.Ua)
            self.states = self.ionic.state_tensors()
            L.check(L.lib().pdx_fd_plan_set_lut(self._plan, C.byref(self.ionic.lut_desc())))
        else:
            self.states = [self._new_field(v) for v in STATE_INIT[self.model_id]]
        self.DIFF = self.AVEC = None
        if lap == 'heterog':
            if DIFF is None:
                raise ValueError("lap='heterog' needs DIFF")
            self.DIFF = self._to_device(DIFF)
            if self.pitch != self.array_shape[-1]:
                t = self._new_field(0.0)
                t.copy_(self.DIFF)
                self.DIFF = t
        elif lap == 'aniso':
            if DIFF is None or AVEC is None:
                raise ValueError("lap='aniso' needs DIFF [W,H,D,2] and AVEC [W,H,D,6]")
            self.DIFF = self._to_device(DIFF, extra=(2,))
            self.AVEC = self._to_device(AVEC, extra=(6,))
            L.check(L.lib().pdx_fd_plan_set_avec(self._plan, L.ptr(self.AVEC)))
            self.refresh_aniso_tensor()
        self._cur = 0            # 0: current U is Ua
        self._states_arr = L.ptr_array(self.states)
        self._graphs = {}
        self.nsteps_done = 0
        self._captured_launches = 0
        self.graph_kernel_launches = 0   # kernels executed through CUDA-graph replays

    def __del__(self):
        try:
            if getattr(self, '_plan', None):
                L.lib().pdx_fd_plan_destroy(self._plan)
                self._plan = None
        except Exception:
            pass

    # ---- data in/out -----------------------------------------------------------------
    def refresh_aniso_tensor(self):
        """lap='aniso': (re)form the six tensor components from DIFF / AVEC for the tile kernel (pdx_fd_aniso.cu).  Called
        at construction; call it again after changing DIFF or AVEC in place.  Shapes the tile kernel does not take
        (nz % 4 != 0, Dirichlet pad, kernel='generic') keep the one-thread-per-node kernel, which reads DIFF / AVEC itself."""
        if self.AVEC is None:
            return False
        if not L.lib().pdx_fd_plan_aniso_tile_capable(self._plan):
            return False
        with torch.cuda.device(self.device):
            L.check(L.lib().pdx_fd_plan_prepare_aniso(self._plan, L.ptr(self.DIFF), L.ptr(self.AVEC), L.stream_ptr()))
        return True

    def _new_field(self, value):
        """A field of the stepper's array shape; pitched steppers: a [..., :nz] view of rows padded to ``pitch`` floats
        (its data_ptr is the padded array's, which is what the library addresses)."""
        nz = self.array_shape[-1]
        if self.pitch == nz:
            return torch.full(self.array_shape, value, dtype=torch.float32, device=self.device)
        full = torch.full(tuple(self.array_shape[:-1]) + (self.pitch,), value, dtype=torch.float32, device=self.device)
        return full[..., :nz]

    def _to_device(self, A, extra=()):
        if isinstance(A, torch.Tensor):
            t = A.to(device=self.device, dtype=torch.float32)
        else:
            t = torch.from_numpy(np.ascontiguousarray(A, dtype=np.float32)).to(self.device)
        shape = tuple(self.array_shape) + tuple(extra)
        if self.part is not None and tuple(t.shape) != shape:
            raise ValueError('distributed steppers take local arrays (SlabPartition.scatter)')
        return t.reshape(shape).contiguous()

    def set_U(self, U):
        self.current_U().copy_(self._to_device(U))

    def set_states(self, states):
        for dst, src in zip(self.states, states):
            dst.copy_(self._to_device(src))

    def current_U(self):
        return self.Ua if self._cur == 0 else self.Ub

    def other_U(self):
        return self.Ub if self._cur == 0 else self.Ua

    def U(self):
        """Current transmembrane potential (device tensor, array shape)."""
        return self.current_U()

    # ---- stepping --------------------------------------------------------------------
    def step(self, nsteps=1, stream=None):
        """nsteps fused steps, one kernel launch each, on the current stream."""
        a, b = self.current_U(), self.other_U()
        with torch.cuda.device(self.device):
            L.check(L.lib().pdx_fd_step(self._plan, L.ptr(a), L.ptr(b), self._states_arr, L.ptr(self.DIFF),
                                        int(nsteps), self._stream(stream)))
        if nsteps & 1:
            self._cur ^= 1
        self.nsteps_done += nsteps

    def step_range(self, xb, xe, stream=None, reverse=False):
        """Advance only array planes [xb, xe) from current_U into other_U (no buffer swap); ``reverse`` reads
        other_U and writes current_U (the odd steps of a skewed multi-step march, see FrameStreamer)."""
        a, b = (self.other_U(), self.current_U()) if reverse else (self.current_U(), self.other_U())
        with torch.cuda.device(self.device):
            L.check(L.lib().pdx_fd_step_range(self._plan, L.ptr(a), L.ptr(b), self._states_arr, L.ptr(self.DIFF),
                                              int(xb), int(xe), self._stream(stream)))

    def swap(self):
        self._cur ^= 1
        self.nsteps_done += 1

    def capture(self, nsteps):
        """CUDA graph of ``nsteps`` steps (nsteps even keeps the buffer parity)."""
        if nsteps % 2:
            raise ValueError('captured blocks use an even number of steps (ping-pong parity)')
        key = (nsteps, self._cur)
        if key not in self._graphs:
            with torch.cuda.device(self.device):
                g = torch.cuda.CUDAGraph()
                s = torch.cuda.Stream(device=self.device)
                s.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(s):
                    a, b = self.current_U(), self.other_U()
                    with torch.cuda.graph(g, stream=s):
                        L.check(L.lib().pdx_fd_step(self._plan, L.ptr(a), L.ptr(b), self._states_arr,
                                                    L.ptr(self.DIFF), int(nsteps), L.stream_ptr(s)))
                torch.cuda.current_stream(self.device).wait_stream(s)
            self._graphs[key] = g
            self._captured_launches += nsteps
        return self._graphs[key]

    def run(self, nsteps, block=10):
        """Advance nsteps using graph-captured blocks of ``block`` steps plus a remainder.  SM-resident
        plans need neither blocks nor graphs: the whole run is one launch."""
        if self.resident and nsteps >= 8:
            self.step(nsteps)
            return
        block = max(2, block - block % 2)
        nblk, rem = divmod(nsteps, block)
        if nblk:
            g = self.capture(block)
            with torch.cuda.device(self.device):
                for _ in range(nblk):
                    g.replay()
            self.nsteps_done += nblk * block
            self.graph_kernel_launches += nblk * block     # one kernel node per captured step
        if rem:
            self.step(rem)

    def apply_stimulus(self, mask, intensity, stream=None, everywhere=False):
        """U = where(mask, max(U, intensity*mask), U): Tests/FD/Fenton/fenton.py:154-156; ``everywhere=True`` is the
        heat examples' U = maximum(U, intensity*mask) on every node (Tests/FD/HeatEquation/heat.py:143-144).
        ``mask`` is a contiguous fp32 CUDA tensor of the stepper's array shape (ghost planes included for slabs)."""
        U = self.current_U()
        self._check_field(mask, 'stimulus mask')
        if self.pitch != self.array_shape[-1]:
            # pitched rows: the flat stimulus kernel does not apply; the same arithmetic as tensor operations (rare call)
            with torch.cuda.device(self.device), torch.cuda.stream(torch.cuda.current_stream(self.device) if stream is None else stream):
                v = mask * float(np.float32(intensity))
                U.copy_(torch.maximum(U, v) if everywhere else torch.where(v != 0, torch.maximum(U, v), U))
            return
        fn = L.lib().pdx_stim_apply_max if everywhere else L.lib().pdx_stim_apply
        with torch.cuda.device(self.device):
            L.check(fn(L.ptr(U), L.ptr(mask), float(np.float32(intensity)), U.numel(), self._stream(stream)))

    def launches(self):
        """Kernel launches issued for this stepper: direct ones plus graph-replayed nodes
        (launches recorded while capturing a graph execute only when the graph is replayed)."""
        return int(L.lib().pdx_fd_plan_launches(self._plan)) - self._captured_launches + self.graph_kernel_launches


class DistributedFDStepper:
    """Slab-decomposed 3-D stepping: one process per GPU.

    halo='peer' (default on CUDA when torch symmetric memory is available): the two U buffers live
    in NVLink symmetric memory; ONE kernel per step advances the whole slab and, fused into it,
    stores its first/last owned planes straight into the neighbours' ghost planes over NVLink
    (``pdx_fd_step_mirror``); one symmetric-memory barrier (~6 us) per step orders the ranks.  No
    NCCL kernel, no separate boundary launches, nothing to overlap.
    Why one barrier suffices: step k reads buffer A and writes B (+ the neighbours' B ghosts).  The
    barrier after step k guarantees (i) every B ghost is complete before any rank's step k+1 reads
    it, and (ii) every rank has finished reading A before a neighbour's step k+1 overwrites its A ghost.

    halo='nccl': per step the two slab-boundary planes are advanced first, their exchange with the
    neighbours (batch_isend_irecv) runs on a side stream while the interior planes are advanced.
    """

    def __init__(self, global_shape, spacing, dt, diff=1.0, rank=0, world=1, group=None, overlap=True, halo='auto',
                 **kw):
        self.part = SlabPartition(global_shape[0], rank, world)
        self.group = group
        self.overlap = overlap and world > 1
        self.world, self.rank = world, rank
        self.halo = 'nccl'
        buffers = None
        if world > 1 and halo in ('auto', 'peer'):
            try:
                buffers = self._alloc_symmetric(global_shape, kw.get('device'))
                self.halo = 'peer'
            except Exception:
                if halo == 'peer':
                    raise
        self.fd = FDStepper(global_shape, spacing, dt, diff, part=self.part, buffers=buffers, **kw)
        self.comm_stream = torch.cuda.Stream(device=self.fd.device, priority=-1) if world > 1 else None
        self._ev = torch.cuda.Event()
        self._primed = False
        if self.halo == 'peer':
            p = self.part
            plane_bytes = self.fd.array_shape[1] * self.fd.array_shape[2] * 4
            nloc_lo = SlabPartition(global_shape[0], rank - 1, world).nloc if p.lo_rank is not None else 0
            # [buffer index] -> device address of the neighbour's ghost plane that mirrors my edge plane
            self._mirror_lo = [None if p.lo_rank is None else h.buffer_ptrs[p.lo_rank] + (nloc_lo + 1) * plane_bytes
                               for h in self._hdl]
            self._mirror_hi = [None if p.hi_rank is None else h.buffer_ptrs[p.hi_rank] for h in self._hdl]

    def _alloc_symmetric(self, global_shape, device):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        # symmetric allocations must have the same size on every rank: size them for the largest slab
        nmax = -(-global_shape[0] // self.part.world) + 2
        full = [symm.empty((nmax, global_shape[1], global_shape[2]), dtype=torch.float32, device=dev) for _ in range(2)]
        self._hdl = [symm.rendezvous(b, self.group if self.group is not None else dist.group.WORLD) for b in full]
        self._symm_full = full
        return [b[:self.part.nx_array] for b in full]

    def set_global(self, U, states=None, DIFF=None):
        """Install this rank's slab of global NumPy fields (ghost planes included)."""
        self.fd.set_U(self.part.scatter(U))
        if states is not None:
            self.fd.set_states([self.part.scatter(s) for s in states])
        if DIFF is not None:
            self.fd.DIFF.copy_(self.fd._to_device(self.part.scatter(DIFF)))
        self._primed = False

    def exchange(self, U):
        return exchange_halos(U, self.part, self.group)

    def _prime(self):
        """Make the current buffer's ghost planes valid and line the ranks up before the first step."""
        self.exchange(self.fd.current_U())
        if self.halo == 'peer':
            self._hdl[0].barrier(channel=0)
        self._primed = True

    def step(self, nsteps=1):
        fd, p = self.fd, self.part
        if p.world > 1 and not self._primed:
            self._prime()
        for _ in range(nsteps):
            if p.world == 1:
                fd.step(1)
                continue
            out = fd.other_U()
            main = torch.cuda.current_stream(fd.device)
            if self.halo == 'peer':
                oi = 0 if out is fd.Ua else 1
                L.check(L.lib().pdx_fd_step_mirror(fd._plan, L.ptr(fd.current_U()), L.ptr(out), fd._states_arr,
                                                   L.ptr(fd.DIFF), C.c_void_p(self._mirror_lo[oi]),
                                                   C.c_void_p(self._mirror_hi[oi]), L.stream_ptr(main)))
                self._hdl[0].barrier(channel=0)
            elif self.overlap and p.nloc >= 3:
                fd.step_range(p.x_begin, p.x_begin + 1)
                fd.step_range(p.x_end - 1, p.x_end)
                self._ev.record(main)
                # the interior is enqueued BEFORE the (CPU-expensive) NCCL calls so that the GPU never
                # idles between the boundary planes and the interior while the host issues the exchange
                fd.step_range(p.x_begin + 1, p.x_end - 1)
                with torch.cuda.stream(self.comm_stream):
                    self.comm_stream.wait_event(self._ev)
                    self.exchange(out)
                main.wait_stream(self.comm_stream)
            else:
                fd.step_range(p.x_begin, p.x_end)
                self.exchange(out)
            fd.swap()

    def launches_per_step(self):
        if self.part.world == 1 or self.halo == 'peer':
            return 1
        return 3 if (self.overlap and self.part.nloc >= 3) else 1

    def owned_U(self):
        return self.part.owned(self.fd.current_U())


def numa_local_pinned(shape, device, dtype=torch.float32):
    """Pinned host tensor whose pages sit on the NUMA node the GPU hangs off (Linux first-touch policy: the
    allocating thread is bound to that node's CPUs while cudaHostAlloc populates the pages).  On a box with a single
    node, or without the sysfs entries, this is a plain pinned allocation."""
    import os
    old = None
    try:
        idx = torch.device(device).index
        idx = torch.cuda.current_device() if idx is None else idx
        bus = torch.cuda.get_device_properties(idx).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(idx), 'pci_domain_id', 0)
        dev_id = getattr(torch.cuda.get_device_properties(idx), 'pci_device_id', 0)
        path = '/sys/bus/pci/devices/%04x:%02x:%02x.0' % (dom, bus, dev_id)
        with open(path + '/numa_node') as f:
            node = int(f.read().strip())
        if node >= 0:
            with open('/sys/devices/system/node/node%d/cpulist' % node) as f:
                cpus = set()
                for part in f.read().strip().split(','):
                    a, _, b = part.partition('-')
                    cpus.update(range(int(a), int(b or a) + 1))
            allowed = os.sched_getaffinity(0)
            if cpus & allowed:
                old = allowed
                os.sched_setaffinity(0, cpus & allowed)
    except Exception:
        old = None
    try:
        t = torch.empty(shape, dtype=dtype, pin_memory=True)
        t.zero_()                      # touch every page while bound
    finally:
        if old is not None:
            try:
                os.sched_setaffinity(0, old)
            except Exception:
                pass
    return t


class FrameStreamer:
    """End-to-end frame loop of ONE simulation over HOST buffers with all three engines busy.

    A frame is what the reference examples do between two writer calls (``Tests/FD/Fenton/fenton.py:150-159``:
    ``dt_per_plot`` steps, then the potential goes to the writer) with the potential living in pinned HOST memory
    between frames: upload U, advance ``steps`` fused steps, download U.  The x axis (slowest, contiguous planes) is
    cut into ``nchunks`` chunks and the three phases are pipelined chunk by chunk:

    * H2D stream: chunk k of the frame's input goes up as soon as the previous frame's download has released
      the host planes it reads (chunk k+1 of that frame);
    * compute stream, ``skew=True``: a time-skewed march - once planes [0, B_k) are on the device, step s may advance
      planes below B_k - s (step s at plane x needs step s-1 at x-1..x+1), so every arriving chunk triggers ``steps``
      range launches ``pdx_fd_step_range`` on shifted plane windows, ping-ponging between the two U buffers in
      place; the last chunk runs every step to the end of the grid.  ``skew=False``: all steps after the last chunk
      (``advance`` callback - the slab-decomposed stepper, whose halo protocol wants whole-slab launches);
    * D2H stream: the planes a chunk's last step completed go down at once.

    Results are bitwise those of ``steps`` whole-grid launches (tests/test_fd_gpu.py).  Frame time -> the PCIe time
    of one direction instead of H2D + compute + D2H.
    """

    def __init__(self, stepper, steps=10, nchunks=8, skew=True, advance=None, planes=None):
        if steps % 2:
            raise ValueError('frames use an even number of steps (ping-pong parity)')
        self.fd, self.steps, self.skew, self.advance = stepper, int(steps), bool(skew), advance
        if getattr(stepper, 'pitch', stepper.array_shape[-1]) != stepper.array_shape[-1]:
            raise L.PdxError('FrameStreamer moves whole contiguous planes: grids with nz % 4 != 0 (pitched rows) are not streamed')
        dev = stepper.device
        # planes [x0, x1) of the array are the ones that travel (a slab leaves its ghost planes out)
        self.x0, self.x1 = planes if planes is not None else (0, stepper.array_shape[0])
        n = self.x1 - self.x0
        if stepper.ndim != 3:
            raise ValueError('FrameStreamer streams 3-D grids plane by plane')
        if skew:
            if stepper.part is not None:
                raise ValueError('the time-skewed march needs a stand-alone grid (use skew=False for slabs)')
            nchunks = max(1, min(int(nchunks), n // (self.steps + 2)))
        else:
            nchunks = max(1, min(int(nchunks), n))
        self.bounds = [self.x0 + (n * k) // nchunks for k in range(nchunks + 1)]
        self.nchunks = nchunks
        self.h2d, self.comp, self.d2h = (torch.cuda.Stream(device=dev) for _ in range(3))
        self.host = numa_local_pinned(stepper.array_shape, dev)
        self.host.copy_(stepper.U())
        torch.cuda.synchronize(dev)
        self._up = [torch.cuda.Event() for _ in range(nchunks)]
        self._fin = [torch.cuda.Event() for _ in range(nchunks)]
        self._down = [torch.cuda.Event() for _ in range(nchunks)]
        self._first = True
        self.frame_bytes = n * stepper.array_shape[1] * stepper.array_shape[2] * 4
        self.launches_per_frame = self.steps * (nchunks if skew else 1)

    def _limit(self, s, k):
        """Planes below this index have their step-s value once chunk k is on the device."""
        return self.x1 if k == self.nchunks - 1 else self.bounds[k + 1] - s

    def frame(self):
        """One frame (enqueue only; ``wait()`` or an event orders the host)."""
        fd, S, C, b = self.fd, self.steps, self.nchunks, self.bounds
        A = fd.current_U()
        for k in range(C):
            with torch.cuda.stream(self.h2d):
                if not self._first:      # the host planes hold the previous frame once its download got there
                    self.h2d.wait_event(self._down[min(k + 1, C - 1)])
                A[b[k]:b[k + 1]].copy_(self.host[b[k]:b[k + 1]], non_blocking=True)
                self._up[k].record(self.h2d)
            if self.skew:
                with torch.cuda.stream(self.comp):
                    self.comp.wait_event(self._up[k])
                    for s in range(1, S + 1):
                        lo = self.x0 if k == 0 else self._limit(s, k - 1)
                        hi = self._limit(s, k)
                        if hi > lo:
                            fd.step_range(lo, hi, stream=self.comp, reverse=(s % 2 == 0))
                    self._fin[k].record(self.comp)
                self._download(k, A)
        if not self.skew:
            with torch.cuda.stream(self.comp):
                self.comp.wait_event(self._up[C - 1])
                if self.advance is not None:
                    self.advance(S)
                else:
                    fd.step(S, stream=self.comp)
                A = fd.current_U()
                self._fin[0].record(self.comp)
            for k in range(C):
                self._download(k, A, fin=self._fin[0])
        else:
            fd.nsteps_done += S
        self._first = False

    def _download(self, k, A, fin=None):
        S, C, b = self.steps, self.nchunks, self.bounds
        if self.skew:
            lo = self.x0 if k == 0 else self._limit(S, k - 1)
            hi = self._limit(S, k)
        else:
            lo, hi = b[k], b[k + 1]
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(self._fin[k] if fin is None else fin)
            if hi > lo:
                self.host[lo:hi].copy_(A[lo:hi], non_blocking=True)
            self._down[k].record(self.d2h)

    def wait(self):
        for st in (self.h2d, self.comp, self.d2h):
            st.synchronize()


class FramePipeline:
    """End-to-end frame loop over host buffers, pipelined across independent simulations.

    Each simulation keeps the reference's per-frame dataflow (``Tests/FD/Fenton/fenton.py:150-159``:
    advance ``dt_per_plot`` steps, hand the potential to the writer) but through HOST memory: its
    potential is uploaded from a pinned host buffer, advanced ``steps_per_frame`` fused steps (one
    CUDA graph), and downloaded back into the same pinned buffer.  With several simulations (an
    ensemble / parameter sweep) in flight the three engines overlap: while one simulation
    computes, another one's frame is on its way down and a third one's input on its way up
    (H2D stream, compute stream, D2H stream; dependencies are CUDA events, no host sync inside).
    """

    def __init__(self, steppers, steps_per_frame=10):
        self.sims = list(steppers)
        self.steps = int(steps_per_frame)
        dev = self.sims[0].device
        self.h2d, self.comp, self.d2h = (torch.cuda.Stream(device=dev) for _ in range(3))
        self.host = [torch.empty(s.array_shape, dtype=torch.float32, pin_memory=True) for s in self.sims]
        for h, s in zip(self.host, self.sims):
            h.copy_(s.U())
            s.capture(self.steps + self.steps % 2)
        torch.cuda.synchronize(dev)
        self._up = [torch.cuda.Event() for _ in self.sims]
        self._done = [torch.cuda.Event() for _ in self.sims]
        self._down = [torch.cuda.Event() for _ in self.sims]
        self._first = True
        self.frame_bytes = self.host[0].numel() * 4

    def frame(self):
        """One frame of every simulation (enqueue only)."""
        for k, s in enumerate(self.sims):
            with torch.cuda.stream(self.h2d):
                if not self._first:
                    self.h2d.wait_event(self._down[k])       # the host buffer holds the previous frame
                s.current_U().copy_(self.host[k], non_blocking=True)
                self._up[k].record(self.h2d)
            with torch.cuda.stream(self.comp):
                self.comp.wait_event(self._up[k])
                s.run(self.steps, block=self.steps)
                self._done[k].record(self.comp)
            with torch.cuda.stream(self.d2h):
                self.d2h.wait_event(self._done[k])
                self.host[k].copy_(s.current_U(), non_blocking=True)
                self._down[k].record(self.d2h)
        self._first = False

    def wait(self):
        for st in (self.h2d, self.comp, self.d2h):
            st.synchronize()
