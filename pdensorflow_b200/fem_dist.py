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


class PartitionedExplicitFEM:
    """One rank's block of the explicit lumped monodomain step.

    n_global:            number of mesh nodes
    rowptr, col, kval:   CSR of K for the rows [lo, hi) of this rank (rowptr local, col GLOBAL)
    mlump:               lumped mass of the same rows
    """

    def __init__(self, n_global, rowptr, col, kval, mlump, dt, model='fenton4v', params=None, rank=0, world=1,
                 group=None, math='fast', device=None, layout='sell'):
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
        self.send_lo = self.send_hi = 0
        if world > 1:
            import torch.distributed as dist
            mine = torch.tensor(self.col_range, dtype=torch.int64, device=self.device)
            allr = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allr, mine, group=group)
            ranges = [tuple(int(v) for v in t.tolist()) for t in allr]
            self.send_lo, self.send_hi = plan_exchange(self.blocks, ranges, rank)
        dev = self.device
        self.layout = layout
        self.nnz = int(rowptr[-1])
        if layout == 'sell':
            sp, scol, sval = csr_to_sell32(rowptr, col, kval, self.lo)
            self.padding = scol.shape[0] / max(self.nnz, 1)
            rowptr, col, kval = sp, scol, sval
        elif layout != 'csr':
            raise ValueError("layout must be 'sell' or 'csr'")
        self.rowptr = torch.from_numpy(rowptr).to(dev)
        self.col = torch.from_numpy(col).to(dev)
        self.kval = torch.from_numpy(np.ascontiguousarray(kval, dtype=np.float32)).to(dev)
        self.mlinv = torch.from_numpy((1.0 / np.asarray(mlump, dtype=np.float64)).astype(np.float32)).to(dev)
        self.model_id = _MODEL_IDS[model]
        self.math_id = _MATH_IDS[math]
        p = (L.default_params(self.model_id) if params is None else list(params)) if self.model_id != L.MODEL_NONE else []
        self._params = (C.c_float * L.MAX_PARAMS)(*p)
        self.dt = dt
        self.states = [torch.full((self.nloc,), v, dtype=torch.float32, device=dev) for v in STATE_INIT[self.model_id]]
        self._states_arr = L.ptr_array(self.states)
        self._hdl = None
        if world > 1:
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
        self._cur = 0
        self._primed = world == 1
        self.nsteps_done = 0

    # ---- data -------------------------------------------------------------------------------
    def set_global_U(self, U):
        """Install a global potential (host array): own rows and ghost columns become valid."""
        t = torch.from_numpy(np.ascontiguousarray(U, dtype=np.float32).reshape(-1)).to(self.device)
        self.U[self._cur].copy_(t)
        self._primed = self.world == 1

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
        for _ in range(nsteps):
            i, o = self._cur, self._cur ^ 1
            L.check(fn(
                self.model_id, self.math_id, self._params, self.nloc, self.lo, L.ptr(self.rowptr), L.ptr(self.col),
                L.ptr(self.kval), L.ptr(self.mlinv), L.ptr(self.U[i]), L.ptr(self.U[o]), self._states_arr, L.ptr(I0),
                dt32, self.send_lo, C.c_void_p(self._peer_lo[o]), self.send_hi, C.c_void_p(self._peer_hi[o]),
                L.stream_ptr()))
            if self.world > 1:
                self._hdl[0].barrier(channel=0)
            self._cur = o
        self.nsteps_done += nsteps
