"""Multi-threaded CPU twin of the oracle's FD step (torch CPU tensors, fp32).

TEST / BASELINE INFRASTRUCTURE - see ``oracle/__init__.py``.  Used only by ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs (TensorFlow, hence the reference itself, cannot
be installed in this image) and cross-checked against ``oracle.fd`` in the CPU tests.

It mirrors the reference's unfused op chain one op per line: symmetric pad -> strided
slices -> elementwise arithmetic -> in-place gate assignment (Tests/FD/Fenton/fenton.py:90-99,
gpuSolve/diffop3D/laplace_homogeneous_isotropic_diffusion.py:24-31,
gpuSolve/ionic/fenton4v.py:168-185), each op a full pass over the array executed by torch's
intra-op thread pool, like TF's Eigen CPU kernels.
"""
import torch
import torch.nn.functional as F


def _sympad3(X):
    # pad by one cell repeating the edge value == tf.pad(mode='symmetric') for pad width 1
    return F.pad(X[None, None], (1, 1, 1, 1, 1, 1), mode='replicate')[0, 0]


def enforce_boundary(X):
    return _sympad3(X[1:-1, 1:-1, 1:-1])


def laplace_homog_3d(X0, DX, DY, DZ):
    X = _sympad3(X0)
    dxsq, dysq, dzsq = DX * DX, DY * DY, DZ * DZ
    c = X[1:-1, 1:-1, 1:-1]
    return ((X[0:-2, 1:-1, 1:-1] - 2.0 * c + X[2:, 1:-1, 1:-1]) / dxsq
            + (X[1:-1, 0:-2, 1:-1] - 2.0 * c + X[1:-1, 2:, 1:-1]) / dysq
            + (X[1:-1, 1:-1, 0:-2] - 2.0 * c + X[1:-1, 1:-1, 2:]) / dzsq)


def _H(x):
    return (1.0 + torch.sign(x)) * 0.5


def _G(x):
    return (1.0 - torch.sign(x)) * 0.5


class Fenton4v:
    """Same defaults and update as oracle.ionic.Fenton4v, on torch CPU tensors."""

    def __init__(self, dt):
        self._dt = dt
        self.tau_vp, self.tau_vn, self.tau_wp, self.tau_wn1, self.tau_wn2 = 3.33, 19.2, 160.0, 75.0, 75.0
        self.tau_d, self.tau_si, self.tau_so, self.tau_a = 0.065, 31.8364, 31.8364, 0.009
        self.u_c, self.u_w, self.u_0, self.u_m, self.u_csi, self.u_so = 0.23, 0.146, 0.0, 1.0, 0.8, 0.3
        self.r_sp, self.r_sn, self.k_, self.a_so, self.b_so, self.c_so = 0.02, 1.2, 3.0, 0.115, 0.84, 0.02
        self.vmin, self.DV = -80.0, 100.0
        self.V = self.W = self.S = None

    def initialize_state_variables(self, U):
        self.V = torch.ones_like(U)
        self.W = torch.ones_like(U)
        self.S = torch.zeros_like(U)

    def differentiate(self, U):
        V, W, S = self.V, self.W, self.S
        Uad = (U - self.vmin) / self.DV
        I_fi = -V * _H(Uad - self.u_c) * (Uad - self.u_c) * (self.u_m - Uad) / self.tau_d
        I_si = -W * S / self.tau_si
        I_so = (0.5 * (self.a_so - self.tau_a) * (1 + torch.tanh((Uad - self.b_so) / self.c_so)) +
                (Uad - self.u_0) * _G(Uad - self.u_so) / self.tau_so + _H(Uad - self.u_so) * self.tau_a)
        dU = -(self.DV * (I_fi + I_si + I_so))
        dV = torch.where(Uad > self.u_c, -V / self.tau_vp, (1 - V) / self.tau_vn)
        dW = torch.where(Uad > self.u_c, -W / self.tau_wp,
                         torch.where(Uad > self.u_w, (1 - W) / self.tau_wn2, (1 - W) / self.tau_wn1))
        r_s = (self.r_sp - self.r_sn) * _H(Uad - self.u_c) + self.r_sn
        dS = r_s * (0.5 * (1 + torch.tanh((Uad - self.u_csi) * self.k_)) - S)
        self.V = V + self._dt * dV
        self.W = W + self._dt * dW
        self.S = S + self._dt * dS
        return dU


def fd_step(model, U, dt, diff, spacing):
    """U1 = U0 + dt*dU(U) + (diff*dt)*laplace(U0): Tests/FD/Fenton/fenton.py:90-99."""
    U0 = enforce_boundary(U)
    dU = model.differentiate(U)
    return U0 + dt * dU + (diff * dt) * laplace_homog_3d(U0, *spacing)


def time_fd(shape, nsteps, dt=0.1, diff=0.8, spacing=(1.0, 1.0, 1.0), threads=None):
    """Seconds for ``nsteps`` steps of the S1 plane-wave problem on ``shape`` (after one
    untimed step), and the thread count used."""
    import os
    import time
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    U = torch.full(shape, -80.0, dtype=torch.float32)
    U[0:2] = 20.0
    m = Fenton4v(dt)
    m.initialize_state_variables(U)
    U = fd_step(m, U, dt, diff, spacing)
    t0 = time.perf_counter()
    for _ in range(nsteps):
        U = fd_step(m, U, dt, diff, spacing)
    return time.perf_counter() - t0, threads, U


# ---------------------------------------------------------------------------------------------------------
# 2-D step (config 2) and the FEM steps (configs 1 and 5), same op-for-op style
def _sympad2(X):
    return F.pad(X[None, None], (1, 1, 1, 1), mode='replicate')[0, 0]


def fd_step_2d(model, U, dt, diff, spacing):
    """The examples' solve() transcribed to 2-D with gpuSolve/diffop2D/laplace_homogeneous_isotropic_diffusion.py:3-36."""
    U0 = _sympad2(U[1:-1, 1:-1])
    X = _sympad2(U0)
    DX, DY = spacing
    c = X[1:-1, 1:-1]
    lap = ((X[0:-2, 1:-1] - 2.0 * c + X[2:, 1:-1]) / (DX * DX) + (X[1:-1, 0:-2] - 2.0 * c + X[1:-1, 2:]) / (DY * DY))
    return U0 + dt * model.differentiate(U) + (diff * dt) * lap


def time_fd_2d(shape, nsteps, dt=0.1, diff=0.8, spacing=(1.0, 1.0), threads=None):
    import os
    import time
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    U = torch.full(shape, -80.0, dtype=torch.float32)
    U[0:2] = 20.0
    m = Fenton4v(dt)
    m.initialize_state_variables(U)
    U = fd_step_2d(m, U, dt, diff, spacing)
    t0 = time.perf_counter()
    for _ in range(nsteps):
        U = fd_step_2d(m, U, dt, diff, spacing)
    return time.perf_counter() - t0, threads, U


def pcg(A, minv, b, x0, maxiter, toll=1e-5, check_every=5):
    """gpuSolve/linearsolvers/conjgrad.py:172-203,287-301 (eager path) on torch-CPU: one op per line, a host
    read of the residual every ``check_every`` iterations like the reference's ``.numpy()``."""
    x = x0
    r = b - A @ x
    z = minv * r
    p = z
    rz = torch.sum(r * z)
    toll_sq = toll * toll
    n = 0
    for n in range(1, 1 + maxiter):
        q = A @ p
        alpha = rz / torch.sum(p * q)
        x = x + alpha * p
        r = r - alpha * q
        res = torch.sum(r * r)
        z = minv * r
        rznew = torch.sum(r * z)
        beta = rznew / rz
        p = z + beta * p
        if n % check_every == 0:
            rv = float(res)
            if rv != rv or rv in (float('inf'), float('-inf')) or rv < toll_sq:
                break
        rz = rznew
    return x, n


def semi_implicit_step(model, U, I0, dt, MASS, A, minv, maxiter):
    """MonodomainSolver.solve_step (gpuSolve/physics/monodomainSolver.py:98-108): ionic update, RHS = MASS (U + dt (dU
    + I0)), Jacobi-PCG solve of (MASS + dt STIFFNESS) U1 = RHS from the guess U.  MASS / A are torch sparse CSR."""
    dU = model.differentiate(U) + I0
    rhs = MASS @ (U + dt * dU)
    return pcg(A, minv, rhs, U, maxiter)


def explicit_lumped_step(model, U, I0, dt, K, mlinv):
    """Row E (no reference implementation): U1 = U + dt (dU + I0) - dt M_L^-1 (K U), unfused, torch sparse CSR."""
    dU = model.differentiate(U) + I0
    return U + dt * dU - dt * (mlinv * (K @ U))


def sparse_csr(rowptr, col, val, n):
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')          # "sparse CSR support is in beta"
        return _sparse_csr(rowptr, col, val, n)


def _sparse_csr(rowptr, col, val, n):
    return torch.sparse_csr_tensor(torch.as_tensor(rowptr, dtype=torch.int64), torch.as_tensor(col, dtype=torch.int64),
                                   torch.as_tensor(val, dtype=torch.float32), size=(n, n))
