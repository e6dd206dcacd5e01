"""Row-block partitioned explicit mass-lumped FEM stepping (BASELINE.json config 5).

    U1 = U + dt*(dU(U) + I0) - dt * M_L^-1 * (K U)          

The mesh nodes are ordered so that K is banded (structured / RCM order); rank r owns the contiguous
rows [lo_r, hi_r) and needs the potentials of the columns its rows touch outside that range - its
ghost nodes, which for a banded matrix sit in the two neighbouring blocks.  Every rank keeps
full-length U buffers (8e7 bytes for 2e7 nodes) in NVLink symmetric memory; the step kernel
(``pdx_fem_explicit_step_part``) writes the new potential of its own rows and, fused into the same
kernel, stores the rows its neighbours hold as ghosts directly into THEIR buffers over NVLink.
One symmetric-memory barrier per step orders the ranks (same argument as the FD slabs,
``fd.DistributedFDStepper``).  No collective, no packing kernels.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from .fd import STATE_INIT, _MATH_IDS, _MODEL_IDS


def row_blocks(n, world):
    """[(lo, hi)] of a balanced contiguous split of n rows over ``world`` ranks."""
    base, rem = divmod(n, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def plan_exchange(blocks, col_ranges, rank):
    """How many of my first / last rows the lower / upper neighbour holds as ghosts.

    blocks:     [(lo, hi)] per rank;  col_ranges: [(min_col, max_col)] of each rank's rows.
    Returns (send_lo, send_hi).  Raises when a rank's ghosts reach beyond its two neighbours
    (the matrix is not banded enough for this partition - reorder the mesh first).
    """
    world = len(blocks)
    for r, ((lo, hi), (cmin, cmax)) in enumerate(zip(blocks, col_ranges)):
        lo_limit = blocks[r - 1][0] if r > 0 else lo
        hi_limit = blocks[r + 1][1] if r < world - 1 else hi
        if cmin < lo_limit or cmax >= hi_limit:
            raise ValueError('rank %d needs columns [%d, %d] outside its neighbours\' rows [%d, %d): the row-block '
                             'partition needs a banded ordering (RCM)' % (r, cmin, cmax, lo_limit, hi_limit))
    lo, hi = blocks[rank]
    send_lo = send_hi = 0
    if rank > 0:                       # the lower neighbour's rows reach up to its max column
        send_lo = max(0, col_ranges[rank - 1][1] + 1 - lo)
    if rank < world - 1:               # the upper neighbour's rows reach down to its min column
        send_hi = max(0, hi - col_ranges[rank + 1][0])
    return send_lo, send_hi


def csr_to_sell32(rowptr, col, val, row_lo=0):
    """CSR (local rows, global columns) -> SELL-32: (slice_ptr int32, col int32, val fp32).  Rows of a
    32-row slice are padded to the slice's longest row with (col = the row's own global index, val = 0)
    and stored column-major inside the slice (include/pdx.h, pdx_fem_explicit_step_sell)."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    n = rowptr.shape[0] - 1
    lens = np.diff(rowptr)
    nsl = (n + 31) // 32
    lens_p = np.zeros(nsl * 32, dtype=np.int64)
    lens_p[:n] = lens
    width = lens_p.reshape(nsl, 32).max(axis=1)
    slice_ptr = np.zeros(nsl + 1, dtype=np.int64)
    np.cumsum(width * 32, out=slice_ptr[1:])
    if slice_ptr[-1] >= 2 ** 31:
        raise ValueError('SELL storage exceeds int32 offsets: partition the matrix over more ranks')
    own = np.minimum(np.arange(nsl * 32, dtype=np.int64), n - 1) + row_lo
    scol = np.repeat(own.reshape(nsl, 32), width, axis=0).reshape(-1)      # [sum(width), 32] rows -> padded with own index
    sval = np.zeros(scol.shape[0], dtype=np.float32)
    rows = np.repeat(np.arange(n, dtype=np.int64), lens)
    k = np.arange(rowptr[-1], dtype=np.int64) - rowptr[rows]              # position of each entry inside its row
    dst = slice_ptr[rows // 32] + k * 32 + (rows % 32)
    scol[dst] = np.asarray(col, dtype=np.int64)
    sval[dst] = np.asarray(val, dtype=np.float32)
    return slice_ptr.astype(np.int32), scol.astype(np.int32), sval


def csr_to_sell32_device(rowptr, col, val, row_lo=0):
    """``csr_to_sell32`` on the device (pdx_sell_slice_widths / pdx_sell_fill): CUDA tensors in (rowptr int32 [n+1],
    col int32, val fp32), CUDA tensors out (slice_ptr int32, col int32, val fp32), identical to the host conversion.
    Sits next to the device assembly (gpuSolve.matrices.device_assembly): a mesh assembled on the GPU never visits the
    host on its way to the SELL kernels."""
    for t, dt in ((rowptr, torch.int32), (col, torch.int32), (val, torch.float32)):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dt and t.is_contiguous()):
            raise L.PdxError('csr_to_sell32_device takes contiguous CUDA tensors (int32 rowptr/col, fp32 val)')
    n = rowptr.numel() - 1
    nsl = (n + 31) // 32
    lib = L.lib()
    with torch.cuda.device(rowptr.device):
        widths = torch.empty(max(nsl, 1), dtype=torch.int32, device=rowptr.device)
        L.check(lib.pdx_sell_slice_widths(n, L.ptr(rowptr), L.ptr(widths), L.stream_ptr()))
        sp = torch.zeros(nsl + 1, dtype=torch.int64, device=rowptr.device)
        torch.cumsum(widths[:nsl].to(torch.int64) * 32, 0, out=sp[1:])
        total = int(sp[-1])                                     # one host read: the size of the SELL arrays
        if total >= 2 ** 31:
            raise ValueError('SELL storage exceeds int32 offsets: partition the matrix over more ranks')
        sp32 = sp.to(torch.int32)
        scol = torch.empty(total, dtype=torch.int32, device=rowptr.device)
        sval = torch.empty(total, dtype=torch.float32, device=rowptr.device)
        L.check(lib.pdx_sell_fill(n, int(row_lo), L.ptr(rowptr), L.ptr(col), L.ptr(val), L.ptr(sp32), L.ptr(scol),
                                  L.ptr(sval), L.stream_ptr()))
        # widest slice: lets the explicit step stream large matrices through shared memory (pdx_sell_hint)
        L.check(lib.pdx_sell_hint(L.ptr(sp32), int(widths[:nsl].max()) if nsl else 0))
    return sp32, scol, sval


def sell_padding(row_lengths, order=None):
    """Stored / useful entries of SELL-32 for rows of these lengths (optionally visited in ``order``)."""
    lens = np.asarray(row_lengths, dtype=np.int64)
    if order is not None:
        lens = lens[np.asarray(order)]
    n = lens.shape[0]
    pad = np.zeros((n + 31) // 32 * 32, dtype=np.int64)
    pad[:n] = lens
    return float(pad.reshape(-1, 32).max(axis=1).sum() * 32) / max(int(lens.sum()), 1)


def sigma_window_order(row_lengths, sigma=256, blocks=None):
    """SELL-C-sigma row order (C = 32): inside windows of ``sigma`` consecutive rows the rows are sorted by decreasing
    length, so that the 32 rows of a slice have similar lengths and little padding; windows never straddle a
    partition block (``blocks`` = [(lo, hi)] of the row-block partition), so the ordering is partition aware: the
    rows a rank owns stay the rows it owns and the band structure moves by less than sigma.
    Returns ``perm`` (new position -> old row), to be composed with the mesh renumbering like the reference's RCM
    ``perm`` (gpuSolve/matrices/globalMatrices.py:54-74) - i.e. applied to rows AND columns before assembly /
    with ``permute_csr``."""
    lens = np.asarray(row_lengths, dtype=np.int64)
    n = lens.shape[0]
    if sigma % 32 or sigma <= 0:
        raise ValueError('sigma must be a positive multiple of the slice height 32')
    blocks = [(0, n)] if blocks is None else list(blocks)
    win = np.empty(n, dtype=np.int64)
    base = 0
    for lo, hi in blocks:
        win[lo:hi] = base + (np.arange(hi - lo) // sigma)
        base += -(-(hi - lo) // sigma)
    return np.lexsort((np.arange(n), -lens, win)).astype(np.int64)      # stable: ties keep their order


def permute_csr(rowptr, col, val, perm):
    """P A P^T for the renumbering new -> old ``perm``: row i of the result is old row perm[i] with its columns
    relabelled through the inverse permutation; the entries of a row keep their order."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    perm = np.asarray(perm, dtype=np.int64)
    n = perm.shape[0]
    iperm = np.empty(n, dtype=np.int64)
    iperm[perm] = np.arange(n)
    lens = np.diff(rowptr)[perm]
    new_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=new_ptr[1:])
    src = np.repeat(rowptr[perm], lens) + (np.arange(new_ptr[-1]) - np.repeat(new_ptr[:-1], lens))
    return new_ptr.astype(np.int32), iperm[np.asarray(col, dtype=np.int64)[src]].astype(np.int32), np.asarray(val)[src]


def ghost_columns(col, lo, hi):
    """Sorted unique global columns a block of rows [lo, hi) touches outside itself."""
    c = np.asarray(col, dtype=np.int64)
    return np.unique(c[(c < lo) | (c >= hi)])


def plan_ghost_lists(blocks, ghosts):
    """General ghost exchange of a row-block partition (no band assumption).

    blocks: [(lo, hi)] per rank; ghosts: per rank the sorted global columns it needs from other ranks
    (``ghost_columns``).  Returns per rank a pair (rows, peers): after each step rank r stores row rows[k] of its
    new potential into rank peers[k]'s buffer (pdx_push_rows)."""
    world = len(blocks)
    bounds = np.array([b[0] for b in blocks] + [blocks[-1][1]], dtype=np.int64)
    send = [([], []) for _ in range(world)]
    for r, g in enumerate(ghosts):
        g = np.asarray(g, dtype=np.int64)
        if g.size == 0:
            continue
        if g.min() < bounds[0] or g.max() >= bounds[-1]:
            raise ValueError('rank %d needs columns outside the matrix' % r)
        owner = np.searchsorted(bounds, g, side='right') - 1
        for o in np.unique(owner):
            if o == r:
                raise ValueError('rank %d lists one of its own rows as a ghost' % r)
            rows = g[owner == o]
            send[o][0].append(rows)
            send[o][1].append(np.full(rows.shape[0], r, dtype=np.int64))
    out = []
    for rows, peers in send:
        if rows:
            out.append((np.concatenate(rows).astype(np.int32), np.concatenate(peers).astype(np.int32)))
        else:
            out.append((np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32)))
    return out


class PartitionedExplicitFEM:
    """One rank's block of the explicit lumped monodomain step.

    n_global:            number of mesh nodes
    rowptr, col, kval:   CSR of K for the rows [lo, hi) of this rank (rowptr local, col GLOBAL)
    mlump:               lumped mass of the same rows
    """

    def __init__(self, n_global, rowptr, col, kval, mlump, dt, model='fenton4v', params=None, rank=0, world=1,
                 group=None, math='fast', device=None, layout='sell', connect=True, exchange='auto', sync='auto'):
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != 'cuda':
            raise L.PdxError('PartitionedExplicitFEM needs a CUDA device: there is no CPU path')
        self.n, self.rank, self.world, self.group = int(n_global), rank, world, group
        self.blocks = row_blocks(self.n, world)
        self.lo, self.hi = self.blocks[rank]
        self.nloc = self.hi - self.lo
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        col = np.ascontiguousarray(col, dtype=np.int32)
        if rowptr.shape[0] != self.nloc + 1:
            raise ValueError('rowptr must describe the %d rows of this rank' % self.nloc)
        self.col_range = (int(col.min()), int(col.max()))
        self.ghosts = ghost_columns(col, self.lo, self.hi)
        self.send_lo = self.send_hi = 0
        # exchange: 'band' = the neighbours' edge rows, stored by the step kernel itself (banded orderings);
        # 'lists' = index lists + pdx_push_rows (any partition); 'auto' = band when the partition allows it
        if exchange not in ('auto', 'band', 'lists'):
            raise ValueError("exchange must be 'auto', 'band' or 'lists'")
        self.exchange = exchange
        # sync: 'barrier' = one symmetric-memory barrier per step (default, 'auto'); 'flags' = the step kernels order
        # themselves through neighbour flags in peer memory (pdx_fem_explicit_step_sync, band exchange only; steps are
        # then graph replayable).  Measured on config 5: level on 2 GPUs (0.0756 vs 0.0766 ms per step at 2.5 M rows
        # each), SLOWER on 8 (0.083-0.090 vs 0.076-0.077: the system-scope fences and flag words of 2 x 2312 edge slices
        # per step cost more than the ~6 us barrier) - hence opt-in
        if sync not in ('auto', 'flags', 'barrier'):
            raise ValueError("sync must be 'auto', 'flags' or 'barrier'")
        self.sync = sync
        self._ctrl = self._ctrl_lo = self._ctrl_hi = None
        self._push = None
        self._local = world > 1 and not connect      # ranks sharing one GPU in one process: see connect_local
        if world > 1 and connect:
            import torch.distributed as dist
            mine = torch.tensor(self.col_range, dtype=torch.int64, device=self.device)
            allr = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allr, mine, group=group)
            ranges = [tuple(int(v) for v in t.tolist()) for t in allr]
            self._plan(ranges, lambda: self._gather_ghosts(group))
        dev = self.device
        self.layout = layout
        self.nnz = int(rowptr[-1])
        if layout not in ('sell', 'csr'):
            raise ValueError("layout must be 'sell' or 'csr'")
        self.rowptr = torch.from_numpy(rowptr).to(dev)
        self.col = torch.from_numpy(col).to(dev)
        self.kval = torch.from_numpy(np.ascontiguousarray(kval, dtype=np.float32)).to(dev)
        if layout == 'sell':       # CSR -> SELL-32 on the device (pdx_sell_fill); the CSR copy is dropped
            self.rowptr, self.col, self.kval = csr_to_sell32_device(self.rowptr, self.col, self.kval, self.lo)
            self.padding = self.col.numel() / max(self.nnz, 1)
        if layout == 'csr':
            L.check(L.lib().pdx_csr_hint(L.ptr(self.rowptr), self.nloc, self.nnz))     # lanes per row of the CSR kernel
        self.mlinv = torch.from_numpy((1.0 / np.asarray(mlump, dtype=np.float64)).astype(np.float32)).to(dev)
        self.model_id = _MODEL_IDS[model]
        self.math_id = _MATH_IDS[math]
        p = (L.default_params(self.model_id) if params is None else list(params)) if self.model_id != L.MODEL_NONE else []
        self._params = (C.c_float * L.MAX_PARAMS)(*p)
        self.dt = dt
        self.states = [torch.full((self.nloc,), v, dtype=torch.float32, device=dev) for v in STATE_INIT[self.model_id]]
        self._states_arr = L.ptr_array(self.states)
        self._hdl = None
        if world > 1 and connect:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm
            self.U = [symm.empty((self.n,), dtype=torch.float32, device=dev) for _ in range(2)]
            self._hdl = [symm.rendezvous(b, group if group is not None else dist.group.WORLD) for b in self.U]
            self._peer_lo = [h.buffer_ptrs[rank - 1] if rank > 0 else None for h in self._hdl]
            self._peer_hi = [h.buffer_ptrs[rank + 1] if rank < world - 1 else None for h in self._hdl]
        else:
            self.U = [torch.empty((self.n,), dtype=torch.float32, device=dev) for _ in range(2)]
            self._peer_lo = self._peer_hi = [None, None]
        for b in self.U:
            b.fill_(-80.0 if self.model_id != L.MODEL_NONE else 0.0)
        if self._hdl is not None and self.exchange == 'band' and self.sync == 'flags':
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm
            self._ctrl = symm.empty((32,), dtype=torch.int32, device=dev)
            self._ctrl.zero_()
            h = symm.rendezvous(self._ctrl, group if group is not None else dist.group.WORLD)
            self._ctrl_hdl = h
            self._ctrl_lo = h.buffer_ptrs[rank - 1] if rank > 0 else None
            self._ctrl_hi = h.buffer_ptrs[rank + 1] if rank < world - 1 else None
            self.sync = 'flags'
        elif self.sync == 'flags' and not self._local:
            if world > 1:
                raise ValueError("sync='flags' needs the band exchange")
            self.sync = 'barrier'
        elif self.sync == 'auto' and not self._local:
            self.sync = 'barrier'
        if self._hdl is not None and self.exchange == 'lists':
            self._install_push([[int(q) for q in h.buffer_ptrs] for h in self._hdl])
        self._cur = 0
        self._graphs = {}
        self._primed = world == 1 or self._local
        self.nsteps_done = 0

    def __del__(self):
        try:        # the slice-width hint is keyed by the slice_ptr address: forget it with the matrix
            if getattr(self, 'layout', None) == 'sell' and getattr(self, 'rowptr', None) is not None:
                L.lib().pdx_sell_hint(L.ptr(self.rowptr), 0)
        except Exception:
            pass

    # ---- exchange planning --------------------------------------------------------------------
    def _gather_ghosts(self, group):
        import torch.distributed as dist
        allg = [None] * self.world
        dist.all_gather_object(allg, self.ghosts, group=group)
        return allg

    def _plan(self, ranges, all_ghosts):
        """Band exchange when every rank's ghosts sit in its two neighbours, index lists otherwise."""
        if self.exchange in ('auto', 'band'):
            try:
                self.send_lo, self.send_hi = plan_exchange(self.blocks, ranges, self.rank)
                self.exchange = 'band'
                return
            except ValueError:
                if self.exchange == 'band':
                    raise
        self.exchange = 'lists'
        self.send_lo = self.send_hi = 0
        self._push_plan = plan_ghost_lists(self.blocks, all_ghosts())[self.rank]

    def _install_push(self, peer_ptrs):
        """peer_ptrs[b][r] = device address of rank r's U buffer b."""
        rows, peers = self._push_plan
        dev = self.device
        self._push = (torch.from_numpy(rows).to(dev), torch.from_numpy(peers).to(dev),
                      [torch.tensor(p, dtype=torch.int64, device=dev) for p in peer_ptrs], int(rows.shape[0]))

    @staticmethod
    def connect_local(ranks):
        """Wire instances (constructed with connect=False) that live in ONE process on ONE GPU: the neighbours'
        buffers are plain device pointers and stream order replaces the barrier - the ghost stores fused into
        the step kernel are exercised on a single-GPU box.  Step the ranks with ``step_local``."""
        ranks = sorted(ranks, key=lambda f: f.rank)
        ranges = [f.col_range for f in ranks]
        ghosts = [f.ghosts for f in ranks]
        for f in ranks:
            f._plan(ranges, lambda: ghosts)
            f._peer_lo = [ranks[f.rank - 1].U[b].data_ptr() if f.rank > 0 else None for b in range(2)]
            f._peer_hi = [ranks[f.rank + 1].U[b].data_ptr() if f.rank < f.world - 1 else None for b in range(2)]
            if f.exchange == 'lists':
                f._peer_lo = f._peer_hi = [None, None]
                f._install_push([[g.U[b].data_ptr() for g in ranks] for b in range(2)])
        if all(f.exchange == 'band' and f.sync == 'flags' for f in ranks):
            # the neighbour-flag protocol on one GPU: the ranks share a stream, so a waiting kernel's neighbours have
            # always finished already - the flags are exercised, never contended
            for f in ranks:
                f._ctrl = torch.zeros(32, dtype=torch.int32, device=f.device)
            for f in ranks:
                f.sync = 'flags'
                f._ctrl_lo = ranks[f.rank - 1]._ctrl.data_ptr() if f.rank > 0 else None
                f._ctrl_hi = ranks[f.rank + 1]._ctrl.data_ptr() if f.rank < f.world - 1 else None
        else:
            for f in ranks:
                f.sync = 'barrier'

    @staticmethod
    def step_local(ranks, nsteps=1):
        for _ in range(nsteps):
            for f in ranks:             # same stream: every rank's step k completes before any rank's step k+1
                f.step(1)

    # ---- data -------------------------------------------------------------------------------
    def set_global_U(self, U):
        """Install a global potential (host array): own rows and ghost columns become valid."""
        t = torch.from_numpy(np.ascontiguousarray(U, dtype=np.float32).reshape(-1)).to(self.device)
        if self._ctrl is not None and not self._local and self._primed:
            # neighbour flags: a neighbour's last step may still be storing ghost rows into this rank's buffers
            self._hdl[0].barrier(channel=0)
        self.U[self._cur].copy_(t)
        self._primed = self.world == 1 or self._local

    def current_U(self):
        return self.U[self._cur]

    def owned_U(self):
        return self.U[self._cur][self.lo:self.hi]

    # ---- stepping ---------------------------------------------------------------------------
    def _prime(self):
        self._hdl[0].barrier(channel=0)        # every rank has installed its buffers
        self._primed = True

    def step(self, nsteps=1, I0=None):
        if not self._primed:
            self._prime()
        lib = L.lib()
        dt32 = float(np.float32(self.dt))
        fn = lib.pdx_fem_explicit_step_sell if self.layout == 'sell' else lib.pdx_fem_explicit_step_part
        flags = self._ctrl is not None
        for _ in range(nsteps):
            i, o = self._cur, self._cur ^ 1
            if flags:
                L.check(lib.pdx_fem_explicit_step_sync(
                    1 if self.layout == 'sell' else 0, self.model_id, self.math_id, self._params, self.nloc, self.lo,
                    L.ptr(self.rowptr), L.ptr(self.col), L.ptr(self.kval), L.ptr(self.mlinv), L.ptr(self.U[i]),
                    L.ptr(self.U[o]), self._states_arr, L.ptr(I0), dt32, self.send_lo, C.c_void_p(self._peer_lo[o]),
                    self.send_hi, C.c_void_p(self._peer_hi[o]), L.ptr(self._ctrl), C.c_void_p(self._ctrl_lo),
                    C.c_void_p(self._ctrl_hi), L.stream_ptr()))
                self._cur = o
                continue
            L.check(fn(
                self.model_id, self.math_id, self._params, self.nloc, self.lo, L.ptr(self.rowptr), L.ptr(self.col),
                L.ptr(self.kval), L.ptr(self.mlinv), L.ptr(self.U[i]), L.ptr(self.U[o]), self._states_arr, L.ptr(I0),
                dt32, self.send_lo, C.c_void_p(self._peer_lo[o]), self.send_hi, C.c_void_p(self._peer_hi[o]),
                L.stream_ptr()))
            if self._push is not None and self._push[3]:      # general partitions: rows other ranks hold as ghosts
                rows, peers, bufs, count = self._push
                L.check(lib.pdx_push_rows(L.ptr(self.U[o]), count, L.ptr(rows), L.ptr(peers), L.ptr(bufs[o]),
                                          L.stream_ptr()))
            if self.world > 1 and not self._local:
                self._hdl[0].barrier(channel=0)
            self._cur = o
        self.nsteps_done += nsteps

    def run(self, nsteps, block=10, I0=None):
        """nsteps steps as CUDA-graph replays of ``block`` steps (+ a remainder through step()).  The neighbour-flag
        protocol lives on the device (the kernels count their own steps), so a captured block replays correctly;
        with sync='barrier' on several GPUs the per-step barrier keeps the steps un-captured."""
        if self.world > 1 and not self._local and self._ctrl is None:
            return self.step(nsteps, I0)
        if self._push is not None and self._push[3] and self.world > 1 and not self._local:
            return self.step(nsteps, I0)
        if not self._primed:
            self._prime()
        block = max(2, block - block % 2)
        nblk, rem = divmod(nsteps, block)
        if nblk:
            key = (block, self._cur, None if I0 is None else I0.data_ptr())
            graphs = self._graphs
            if key not in graphs:
                g = torch.cuda.CUDAGraph()
                st = torch.cuda.Stream(device=self.device)
                st.wait_stream(torch.cuda.current_stream(self.device))
                done0, cur0 = self.nsteps_done, self._cur
                with torch.cuda.stream(st):
                    with torch.cuda.graph(g, stream=st):
                        self.step(block, I0)
                torch.cuda.current_stream(self.device).wait_stream(st)
                self.nsteps_done, self._cur = done0, cur0          # capturing executes nothing
                graphs[key] = g
            for _ in range(nblk):
                graphs[key].replay()
            self.nsteps_done += nblk * block
        if rem:
            self.step(rem, I0)


# ------------------------------------------------------------------------------------------------
# Semi-implicit step (MonodomainSolver.solve_step) across GPUs
# ------------------------------------------------------------------------------------------------
def plan_windows(blocks, col_ranges):
    """Ghost windows of a row-block partition of a banded matrix.

    blocks: [(lo, hi)] per rank; col_ranges: [(min_col, max_col)] of each rank's rows (global indices).
    Returns one dict per rank: ghost_lo / ghost_hi (rows of the lower / upper neighbour this rank's rows
    reach), send_lo / send_hi (how many of its first / last rows the neighbours hold as ghosts) and the
    index of those rows in the neighbours' windows (include/pdx.h, pdx_pcg_part_desc).  Raises when a
    rank's columns reach beyond its two neighbours (reorder the mesh: RCM)."""
    world = len(blocks)
    plans = []
    for r, ((lo, hi), (cmin, cmax)) in enumerate(zip(blocks, col_ranges)):
        lo_limit = blocks[r - 1][0] if r > 0 else lo
        hi_limit = blocks[r + 1][1] if r < world - 1 else hi
        if cmin < lo_limit or cmax >= hi_limit:
            raise ValueError('rank %d needs columns [%d, %d] outside its neighbours\' rows [%d, %d): the row-block '
                             'partition needs a banded ordering (RCM)' % (r, cmin, cmax, lo_limit, hi_limit))
        plans.append({'lo': lo, 'hi': hi, 'n_local': hi - lo, 'ghost_lo': max(0, lo - cmin), 'ghost_hi': max(0, cmax + 1 - hi)})
    for r, p in enumerate(plans):
        p['send_lo'] = plans[r - 1]['ghost_hi'] if r > 0 else 0
        p['send_hi'] = plans[r + 1]['ghost_lo'] if r < world - 1 else 0
        if p['send_lo'] > p['n_local'] or p['send_hi'] > p['n_local']:
            raise ValueError('rank %d: a neighbour needs more ghost rows than the block has' % r)
        p['lower_window_index'] = plans[r - 1]['ghost_lo'] + plans[r - 1]['n_local'] if r > 0 else -1
        p['upper_window_index'] = plans[r + 1]['ghost_lo'] - p['send_hi'] if r < world - 1 else -1
        p['window'] = p['ghost_lo'] + p['n_local'] + p['ghost_hi']
    return plans


class PartitionedSemiImplicitFEM:
    """One rank's block of the reference's semi-implicit monodomain step

        (MASS + dt*STIFFNESS) U1 = MASS (U0 + dt (differentiate(U0) + I0))
        (gpuSolve/physics/monodomainSolver.py:98-108; heatSolver.py:127-135 when model='none')

    with the nodes split into contiguous row blocks (banded / RCM order).  Two launches per time step and
    per rank: the ionic update over the rank's WINDOW (owned + ghost nodes: ghost gates are advanced
    redundantly, bit-identically to their owner, so RHS0 needs no exchange) and one persistent cooperative
    PCG kernel that talks to the other ranks through NVLink peer memory (pdx_pcg_part_solve).

    rowptr, col, mval, kval: CSR rows [lo, hi) of MASS and STIFFNESS on one pattern (rowptr local, col GLOBAL).

    Ranks are either the processes of a torch.distributed group (workspaces in symmetric memory), or -
    ``connect=False`` + ``PartitionedSemiImplicitFEM.connect_local([...])`` - several instances inside one
    process on one GPU, each on its own stream with ``max_blocks`` capped so that the cooperative grids fit
    side by side (how the protocol is tested on a single GPU).
    """

    def __init__(self, n_global, rowptr, col, mval, kval, dt, model='fenton4v', params=None, rank=0, world=1,
                 group=None, math='fast', device=None, maxiter=100, toll=1e-5, check_every=5, max_blocks=0,
                 timeout_ms=0, stream=None, connect=True, layout='sell'):
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != 'cuda':
            raise L.PdxError('PartitionedSemiImplicitFEM needs a CUDA device: there is no CPU path')
        if world > L.PART_MAX_WORLD:
            raise ValueError('at most %d ranks' % L.PART_MAX_WORLD)
        self.n, self.rank, self.world, self.group = int(n_global), rank, world, group
        self.blocks = row_blocks(self.n, world)
        self.lo, self.hi = self.blocks[rank]
        self.nloc = self.hi - self.lo
        self._rowptr_h = np.ascontiguousarray(rowptr, dtype=np.int32)
        self._col_h = np.ascontiguousarray(col, dtype=np.int64)
        if self._rowptr_h.shape[0] != self.nloc + 1:
            raise ValueError('rowptr must describe the %d rows of this rank' % self.nloc)
        mval = np.ascontiguousarray(mval, dtype=np.float32)
        kval = np.ascontiguousarray(kval, dtype=np.float32)
        # A = MASS + dt*STIFFNESS value-wise on the shared pattern, fp32 (csr_axpby, globalMatrices.py:465-492)
        self._aval_h = mval * np.float32(1.0) + kval * np.float32(dt)
        self._mval_h = mval
        rows = np.repeat(np.arange(self.lo, self.hi, dtype=np.int64), np.diff(self._rowptr_h))
        diag = rows == self._col_h
        minv = np.zeros(self.nloc, dtype=np.float32)          # JacobiPrecond: zeros where the diagonal is absent
        minv[rows[diag] - self.lo] = np.float32(1.0) / self._aval_h[diag]
        self._minv_h = minv
        self.col_range = (int(self._col_h.min()), int(self._col_h.max()))
        self.dt, self.maxiter, self.toll, self.check_every = dt, int(maxiter), float(toll), int(check_every)
        self.max_blocks, self.timeout_ms = int(max_blocks), int(timeout_ms)
        self.model_id = _MODEL_IDS[model]
        self.math_id = _MATH_IDS[math]
        p = (L.default_params(self.model_id) if params is None else list(params)) if self.model_id != L.MODEL_NONE else []
        self._params = (C.c_float * L.MAX_PARAMS)(*p)
        self.stream = stream
        if layout not in ('sell', 'csr'):
            raise ValueError("layout must be 'sell' or 'csr'")
        self.layout = layout
        self._h = None
        self._symm = None
        self.nsteps_done = 0
        if not connect:
            return
        if world == 1:
            self._plan_and_allocate([self.col_range])
            self._attach([self.ws.data_ptr()])
            return
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        mine = torch.tensor(self.col_range, dtype=torch.int64, device=self.device)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine, group=group)
        self._plan_and_allocate([tuple(int(v) for v in t.tolist()) for t in allr], symmetric=True)
        self._symm = symm.rendezvous(self.ws, group if group is not None else dist.group.WORLD)
        self._attach([int(q) for q in self._symm.buffer_ptrs])
        self._symm.barrier(channel=0)                       # every rank has cleared its block before any solve
        torch.cuda.synchronize(self.device)

    @staticmethod
    def connect_local(ranks):
        """Wire instances that live in ONE process on ONE GPU (plain device pointers as peers)."""
        ranks = sorted(ranks, key=lambda f: f.rank)
        ranges = [f.col_range for f in ranks]
        for f in ranks:
            f._plan_and_allocate(ranges)
        ptrs = [f.ws.data_ptr() for f in ranks]
        for f in ranks:
            f._attach(ptrs)
        torch.cuda.synchronize()

    # ---- setup ------------------------------------------------------------------------------
    def _plan_and_allocate(self, ranges, symmetric=False):
        plans = plan_windows(self.blocks, ranges)
        self.plan = plans[self.rank]
        self.max_window = max(p['window'] for p in plans)
        self.glo, self.ghi, self.window = self.plan['ghost_lo'], self.plan['ghost_hi'], self.plan['window']
        nbytes = int(L.lib().pdx_pcg_part_workspace_bytes(self.max_window))
        if symmetric:
            import torch.distributed._symmetric_memory as symm
            self.ws = symm.empty((nbytes,), dtype=torch.uint8, device=self.device)
        else:
            self.ws = torch.empty((nbytes,), dtype=torch.uint8, device=self.device)
        self.ws.zero_()

    def _attach(self, peer_ptrs):
        dev = self.device
        rowptr, col = self._rowptr_h, (self._col_h - (self.lo - self.glo)).astype(np.int32)   # window-local columns
        aval, mval = self._aval_h, self._mval_h
        self.rowptr = torch.from_numpy(rowptr).to(dev)
        self.col = torch.from_numpy(col).to(dev)
        self.aval = torch.from_numpy(np.ascontiguousarray(aval)).to(dev)
        self.mval = torch.from_numpy(np.ascontiguousarray(mval)).to(dev)
        if self.layout == 'sell':       # SELL-32: thread per row, coalesced matrix loads (as the explicit step), built on the device
            nnz = int(rowptr[-1])
            sp, scol, self.aval = csr_to_sell32_device(self.rowptr, self.col, self.aval, self.glo)
            _, _, self.mval = csr_to_sell32_device(self.rowptr, self.col, self.mval, self.glo)
            self.rowptr, self.col = sp, scol
            self.padding = scol.numel() / max(nnz, 1)
        self.minv = torch.from_numpy(self._minv_h).to(dev)
        d = L.PcgPartDesc()
        d.n_local, d.ghost_lo, d.ghost_hi = self.nloc, self.glo, self.ghi
        d.rank, d.world = self.rank, self.world
        d.send_lo, d.send_hi = self.plan['send_lo'], self.plan['send_hi']
        d.lower_window_index, d.upper_window_index = self.plan['lower_window_index'], self.plan['upper_window_index']
        d.max_blocks, d.timeout_ms = self.max_blocks, self.timeout_ms
        d.sell = 1 if self.layout == 'sell' else 0
        d.rowptr, d.col, d.aval = self.rowptr.data_ptr(), self.col.data_ptr(), self.aval.data_ptr()
        d.mval, d.minv = self.mval.data_ptr(), self.minv.data_ptr()
        arr = (C.c_void_p * self.world)(*peer_ptrs)
        h = C.c_void_p()
        L.check(L.lib().pdx_pcg_part_create(C.byref(d), arr, self.max_window, C.byref(h)))
        self._h = h
        # potential window (guess in, solution out) and RHS0 window: ordinary device arrays, no peer touches them
        self.U = torch.full((self.window,), -80.0 if self.model_id != L.MODEL_NONE else 0.0, dtype=torch.float32, device=dev)
        self.rhs0 = torch.empty_like(self.U)
        self.states = [torch.full((self.window,), v, dtype=torch.float32, device=dev) for v in STATE_INIT[self.model_id]]
        self._states_arr = L.ptr_array(self.states)
        self.info = torch.zeros(4, dtype=torch.int32, device=dev)
        self.blocks_per_grid = int(L.lib().pdx_pcg_part_blocks(h))
        self.threads_per_block = int(L.lib().pdx_pcg_part_threads(h))
        self.stage_slot_bytes = int(L.lib().pdx_pcg_part_stage_bytes(h))

    def __del__(self):
        try:
            if self._h is not None:
                L.lib().pdx_pcg_part_destroy(self._h)
        except Exception:
            pass
        self._h = None

    # ---- data -------------------------------------------------------------------------------
    def window_of(self, global_array):
        """This rank's window (ghosts included) of a global nodal array, as a device tensor."""
        g = np.ascontiguousarray(global_array, dtype=np.float32).reshape(-1)
        return torch.from_numpy(g[self.lo - self.glo:self.hi + self.ghi].copy()).to(self.device)

    def set_global_U(self, U):
        self.U.copy_(self.window_of(U))
        torch.cuda.current_stream().synchronize()           # the solves may run on another stream

    def owned_U(self):
        return self.U[self.glo:self.glo + self.nloc]

    def owned_state(self, i):
        return self.states[i][self.glo:self.glo + self.nloc]

    def last_solve(self):
        """(iterations, ||r||^2, flags) of the last solve; synchronises.  Raises when a peer timed out."""
        if self.stream is not None:
            self.stream.synchronize()
        h = self.info.cpu().numpy()
        flags = int(h[2])
        if flags & 4:
            raise L.PdxError('partitioned PCG: a peer rank did not answer within the timeout')
        return int(h[0]), float(h[1:2].view(np.float32)[0]), flags

    # ---- stepping ---------------------------------------------------------------------------
    def solve(self, b=None):
        """A x = b on the potential window (b: owned rows; None: b = MASS * rhs0 window)."""
        L.check(L.lib().pdx_pcg_part_solve(self._h, L.ptr(b), L.ptr(self.rhs0), L.ptr(self.U), self.maxiter, self.toll,
                                           self.check_every, L.ptr(self.info), L.stream_ptr(self.stream)))

    def step(self, nsteps=1, I0=None):
        """nsteps semi-implicit time steps; I0: window-length forcing (window_of(global I0)) or None."""
        lib = L.lib()
        dt32 = float(np.float32(self.dt))
        sp = L.stream_ptr(self.stream)
        for _ in range(nsteps):
            if self.model_id == L.MODEL_NONE:
                with torch.cuda.stream(self.stream or torch.cuda.current_stream()):
                    if I0 is None:
                        self.rhs0.copy_(self.U)
                    else:
                        torch.add(self.U, I0, alpha=dt32, out=self.rhs0)
            else:
                # RHS0 = U + dt*(differentiate(U) + I0) over the whole window   (monodomainSolver.py:100-103)
                L.check(lib.pdx_ionic_step(self.model_id, self.math_id, self._params, None, self.window, L.ptr(self.U),
                                           self._states_arr, None, dt32, L.ptr(I0), L.ptr(self.rhs0), sp))
            L.check(lib.pdx_pcg_part_solve(self._h, None, L.ptr(self.rhs0), L.ptr(self.U), self.maxiter, self.toll,
                                           self.check_every, L.ptr(self.info), sp))
        self.nsteps_done += nsteps
