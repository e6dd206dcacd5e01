"""ConjGrad - mirror of gpuSolve/linearsolvers/conjgrad.py:7-312 on libpdx.

The reference runs ~12 eager TensorFlow ops per CG iteration and synchronises with the host every
``check_every`` iterations (conjgrad.py:287-301).  Here the whole solve - r = b - A x, the Jacobi
scaling, every SpMV / dot / axpy and the batched convergence test with its NaN/Inf guard - is ONE
persistent cooperative kernel (pdx_pcg_solve); the iteration count and ||r||^2 come back through a
4-word device record that is only read when somebody asks for it.  Small systems (config 1: 63 001 rows,
L2 resident, latency bound) use the CSR kernel with a few lanes per row; above ``SELL_MIN_ROWS`` rows the solve
is bandwidth bound and uses SELL-32 storage (thread per row, coalesced matrix loads) through the single-rank
form of the partitioned kernel (pdx_pcg_part_solve, world = 1).
"""
import ctypes as C

import numpy as np
import torch

from .. import _backend as B
from ... import _lib as L
from ..matrices.globalMatrices import CSRMatrix


SELL_MIN_ROWS = 100000


class ConjGrad:
    def __init__(self, config=None):
        self._maxiter = 100
        self._toll = 1.e-5
        self._verbose = False
        self._A = None
        self._RHS = None
        self._X = None
        self._Precond = None
        self._check_every = 5
        self._use_graph_loop = False      # accepted for compatibility: the solve is always device resident
        if config is not None:
            for attribute in list(self.__dict__.keys()):
                name = attribute[1:]
                if name in config.keys():
                    setattr(self, attribute, config[name])
            if 'precond' in config.keys():
                self._Precond = config['precond']
        self._handle = None
        self._handle_key = None
        self._part = None                 # (handle, workspace, sell arrays) of the large-system path
        self._Mass = None                 # MASS on A's pattern: lets the large-system kernel form b = MASS*rhs0 itself
        self._info = None
        self._info_host = None

    def __del__(self):
        self._release()

    def _release(self):
        try:
            if self._handle is not None:
                L.lib().pdx_pcg_destroy(self._handle)
            if self._part is not None:
                L.lib().pdx_pcg_part_destroy(self._part[0])
        except Exception:
            pass
        self._handle = None
        self._part = None

    # ---- configuration ----
    def set_precond(self, prcnd):
        self._Precond = prcnd
        self._release()

    def set_maxiter(self, maxit):
        self._maxiter = maxit

    def set_toll(self, toll):
        self._toll = toll

    def set_matrix(self, Amat):
        if not isinstance(Amat, CSRMatrix):
            raise L.PdxError('ConjGrad.set_matrix takes a gpuSolve.matrices.CSRMatrix')
        self._A = Amat
        self._release()

    def fuse_mass(self, Mmat):
        """(not in the reference) Give the solver the MASS matrix of (MASS + dt*STIFFNESS) U1 = MASS*RHS0:
        ``solve_from_rhs0`` then forms the right-hand side inside the solve kernel - one launch and one vector
        round trip less per time step."""
        if self._A is not None and (Mmat.nnz != self._A.nnz or not np.array_equal(Mmat.col, self._A.col)):
            raise L.PdxError('ConjGrad.fuse_mass: MASS must share the pattern of the system matrix')
        self._Mass = Mmat
        self._release()

    def fused_mass_ready(self):
        return self._Mass is not None and self._A is not None

    def solve_from_rhs0(self, rhs0, x):
        """A x = MASS*rhs0, x holds the guess and receives the solution."""
        self._ensure_handle()
        if self._Mass is None:
            raise L.PdxError('ConjGrad.solve_from_rhs0 needs fuse_mass()')
        if self._part is not None:
            L.check(L.lib().pdx_pcg_part_solve(self._part[0], None, L.ptr(rhs0), L.ptr(x), int(self._maxiter),
                                               float(self._toll), int(self._check_every), L.ptr(self._info), L.stream_ptr()))
        else:
            mval = self._Mass.device_arrays()[2]
            L.check(L.lib().pdx_pcg_solve_rhs0(self._handle, L.ptr(mval), L.ptr(rhs0), L.ptr(x), int(self._maxiter),
                                               float(self._toll), int(self._check_every), L.ptr(self._info), L.stream_ptr()))
        self._X = x

    def set_RHS(self, RHS):
        RHS = L.require_cuda(B.as_tensor(RHS), 'RHS')
        if self._RHS is not None and self._RHS.shape == RHS.shape:
            self._RHS.copy_(RHS)
        else:
            self._RHS = RHS.clone()

    def set_X0(self, X0):
        X0 = L.require_cuda(B.as_tensor(X0), 'X0')
        if self._X is not None and self._X.shape == X0.shape:
            if self._X.data_ptr() != X0.data_ptr():
                self._X.copy_(X0)
        else:
            self._X = X0.clone()

    def Precond(self):
        return self._Precond

    def maxiter(self):
        return self._maxiter

    def toll(self):
        return self._toll

    def matrix(self):
        return self._A

    def RHS(self):
        return self._RHS

    def X(self):
        return self._X

    def verbose(self):
        return self._verbose

    # ---- results (lazy: reading them synchronises with the device) ----
    def _read_info(self):
        if self._info is None:
            return 0, 1.e32, 0
        h = self._info.cpu().numpy()
        return int(h[0]), float(h[1:2].view(np.float32)[0]), int(h[2])

    @property
    def _niters(self):
        return self._read_info()[0]

    @property
    def _residual(self):
        return self._read_info()[1]

    def niters(self):
        return self._niters

    def residual(self):
        return self._residual

    def summary(self):
        n, r, _ = self._read_info()
        if n < self._maxiter:
            print('CG converged in {} iterations (residual: {:4.3f})'.format(n, r))
        else:
            print('WARNING: max nb of iteration reached (residual: {:4.3f})'.format(r))

    def _ensure_handle(self):
        if self._A is None:
            raise L.PdxError('ConjGrad: set_matrix() must be called before solve()')
        minv = None
        if self._Precond is not None:
            if not hasattr(self._Precond, 'inverse_diagonal'):
                raise L.PdxError('ConjGrad: only diagonal (Jacobi) preconditioners are supported on the device')
            minv = self._Precond.inverse_diagonal()
        key = (id(self._A), None if minv is None else minv.data_ptr())
        if (self._handle is None and self._part is None) or self._handle_key != key:
            self._release()
            rp, ci, va = self._A.device_arrays()
            h = C.c_void_p()
            if self._A.n >= SELL_MIN_ROWS:
                from ...fem_dist import csr_to_sell32
                sp, scol, sval = csr_to_sell32(self._A.rowptr, self._A.col, self._A.val)
                dev = rp.device
                arrs = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (sp, scol, sval)]
                if self._Mass is not None:
                    arrs.append(torch.from_numpy(csr_to_sell32(self._Mass.rowptr, self._Mass.col, self._Mass.val)[2]).to(dev))
                ws = torch.zeros(int(L.lib().pdx_pcg_part_workspace_bytes(self._A.n)), dtype=torch.uint8, device=dev)
                d = L.PcgPartDesc()
                d.n_local, d.rank, d.world, d.sell = self._A.n, 0, 1, 1
                d.lower_window_index = d.upper_window_index = -1
                d.rowptr, d.col, d.aval = (a.data_ptr() for a in arrs[:3])
                d.mval = arrs[3].data_ptr() if self._Mass is not None else None
                d.minv = None if minv is None else minv.data_ptr()
                peers = (C.c_void_p * 1)(ws.data_ptr())
                L.check(L.lib().pdx_pcg_part_create(C.byref(d), peers, self._A.n, C.byref(h)))
                self._part = (h, ws, arrs)
            else:
                L.check(L.lib().pdx_pcg_create(self._A.n, L.ptr(rp), L.ptr(ci), L.ptr(va), L.ptr(minv), C.byref(h)))
                self._handle = h
            self._handle_key, self._minv = key, minv
            self._info = torch.zeros(4, dtype=torch.int32, device=rp.device)

    def _launch(self, b, x):
        if self._part is not None:
            L.check(L.lib().pdx_pcg_part_solve(self._part[0], L.ptr(b), None, L.ptr(x), int(self._maxiter), float(self._toll),
                                               int(self._check_every), L.ptr(self._info), L.stream_ptr()))
        else:
            L.check(L.lib().pdx_pcg_solve(self._handle, L.ptr(b), L.ptr(x), int(self._maxiter), float(self._toll),
                                          int(self._check_every), L.ptr(self._info), L.stream_ptr()))

    def solve(self):
        """Solve A X = RHS from the initial guess held in X (in place)."""
        self._ensure_handle()
        if self._X is None or self._RHS is None:
            raise L.PdxError('ConjGrad: set_X0() and set_RHS() must be called before solve()')
        self._launch(self._RHS, self._X)
        if self._verbose:
            self.summary()

    def solve_inplace(self, b, x):
        """Zero-copy variant used by the solvers: A x = b, x holds the guess and receives the solution."""
        self._ensure_handle()
        self._launch(b, x)
        self._X = x
