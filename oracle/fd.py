"""Oracle restatement of the finite-difference operators and the explicit FD step.

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  Not imported by the product.
NumPy fp32, op-for-op in the order the reference source writes the expressions.
PINNED: the reference ships no golden vectors for these operators, so its own source is executed over a NumPy
TensorFlow facade (tests/golden/make_ref_golden.py, make_heat_golden.py, make_atria_golden.py) and the outputs are
frozen in tests/golden/ref_fd.npz / ref_heat.npz / ref_atria.npz; tests/test_ref_golden_cpu.py and
tests/test_atria_cpu.py check that this restatement reproduces them bit for bit.
"""
import numpy as np

F = np.float32


def _sympad(X):
    """tf.pad(mode='symmetric') with one cell per side: the edge value is repeated."""
    return np.pad(X, 1, mode='symmetric')


def enforce_boundary(X):
    """Tests/FD/Fenton/fenton.py:46-53 (3-D) and its 2-D transcription.

    Symmetric pad of the interior: every index is clamped to [1, n-2] on every axis.
    """
    if X.ndim == 3:
        return _sympad(X[1:-1, 1:-1, 1:-1])
    return _sympad(X[1:-1, 1:-1])


def enforce_boundary_dirichlet(X, value):
    """Tests/FD/LaplaceSolver/laplaceSolver.py:44-52 style: constant pad of the interior."""
    inner = X[1:-1, 1:-1, 1:-1] if X.ndim == 3 else X[1:-1, 1:-1]
    return np.pad(inner, 1, mode='constant', constant_values=F(value))


def laplace_homog_3d(X0, DX, DY, DZ):
    """gpuSolve/diffop3D/laplace_homogeneous_isotropic_diffusion.py:3-52 (expression :29-31)."""
    X = _sympad(X0)
    DX, DY, DZ = F(DX), F(DY), F(DZ)
    dxsq, dysq, dzsq = DX * DX, DY * DY, DZ * DZ
    two = F(2.0)
    c = X[1:-1, 1:-1, 1:-1]
    lapla = ((X[0:-2, 1:-1, 1:-1] - two * c + X[2:, 1:-1, 1:-1]) / dxsq
             + (X[1:-1, 0:-2, 1:-1] - two * c + X[1:-1, 2:, 1:-1]) / dysq
             + (X[1:-1, 1:-1, 0:-2] - two * c + X[1:-1, 1:-1, 2:]) / dzsq)
    return lapla.astype(np.float32)


def laplace_conv_homog_3d(X0, DX, DY, DZ):
    """gpuSolve/diffop3D/laplace_convolution_homogeneous_isotropic_diffusion.py:4-47.

    conv3d VALID of the symmetric-padded field with the 3x3x3 kernel built at :30-40
    (float64 NumPy entries from fp32 DX.. products, then cast to fp32).  Note the kernel
    as written puts 1/dzsq on the axis-0 neighbours and 1/dxsq on the axis-2 neighbours
    (:31-39); it equals laplace_homog only for DX == DZ, and that is reproduced here.
    The accumulation order inside conv3d is unspecified; taps are summed in kernel
    memory order, so the result equals laplace_homog up to rounding order.
    """
    X = _sympad(X0)
    DX, DY, DZ = F(DX), F(DY), F(DZ)
    dxsq, dysq, dzsq = DX * DX, DY * DY, DZ * DZ
    dssq = 1.0 / dxsq + 1.0 / dysq + 1.0 / dzsq            # weak python 1.0 -> fp32
    k_ax0 = F(1.0 / dzsq)
    k_ax1 = F(1.0 / dysq)
    k_ax2 = F(1.0 / dxsq)
    k_c = F(-2.0 * dssq)
    c = X[1:-1, 1:-1, 1:-1]
    out = k_ax0 * X[0:-2, 1:-1, 1:-1]
    out = out + k_ax1 * X[1:-1, 0:-2, 1:-1]
    out = out + k_ax2 * X[1:-1, 1:-1, 0:-2]
    out = out + k_c * c
    out = out + k_ax2 * X[1:-1, 1:-1, 2:]
    out = out + k_ax1 * X[1:-1, 2:, 1:-1]
    out = out + k_ax0 * X[2:, 1:-1, 1:-1]
    return out.astype(np.float32)


def laplace_heterog_3d(X0, DIFF0, DX, DY, DZ):
    """gpuSolve/diffop3D/laplace_heterogeneous_isotropic_diffusion.py:3-35."""
    X = _sympad(X0)
    D = _sympad(DIFF0)
    DX, DY, DZ = F(DX), F(DY), F(DZ)
    h, mh = F(0.5), F(-0.5)
    c = X[1:-1, 1:-1, 1:-1]
    dc = D[1:-1, 1:-1, 1:-1]
    e = h * (D[2:, 1:-1, 1:-1] + dc) * (X[2:, 1:-1, 1:-1] - c)
    w = mh * (D[0:-2, 1:-1, 1:-1] + dc) * (X[0:-2, 1:-1, 1:-1] - c)
    n = h * (D[1:-1, 2:, 1:-1] + dc) * (X[1:-1, 2:, 1:-1] - c)
    s = mh * (D[1:-1, 0:-2, 1:-1] + dc) * (X[1:-1, 0:-2, 1:-1] - c)
    u = h * (D[1:-1, 1:-1, 2:] + dc) * (X[1:-1, 1:-1, 2:] - c)
    d = mh * (D[1:-1, 1:-1, 0:-2] + dc) * (X[1:-1, 1:-1, 0:-2] - c)
    dxsq, dysq, dzsq = DX * DX, DY * DY, DZ * DZ
    return ((e - w) / dxsq + (n - s) / dysq + (u - d) / dzsq).astype(np.float32)


def laplace_heterog_aniso_3d(X0, DIFF0, AVEC0, DX, DY, DZ):
    """gpuSolve/diffop3D/laplace_heterogeneous_anisotropic_diffusion.py:3-95.

    DIFF0[...,0] = sigma_t, DIFF0[...,1] = sigma_l - sigma_t; AVEC0[...,0:6] = the fibre dyadic
    (a1a1, a1a2, a1a3, a2a2, a2a3, a3a3).  C = sigma_t*I + (sigma_l - sigma_t) a(x)a on the
    symmetric-padded arrays (:40-45).  For each axis ``a`` and side ``sg`` the face flux is
        sg*0.5*( sum over b = x,y,z of T_b ),
        T_a = (C_aa[sg e_a] + C_aa[0]) * (X[sg e_a] - X[0]) / D_a                       (normal part)
        T_b = 0.5*(C_ab[sg e_a + e_b] + C_ab[sg e_a - e_b]) * (X[sg e_a + e_b] - X[sg e_a - e_b]) / D_b
    (:50-90; centred transverse differences taken AT the neighbour cell), and
        lap = (e - w)/DX + (n - s)/DY + (u - d)/DZ                                      (:93).
    Note the result is divided by the spacing once here and once inside each flux.
    """
    X = _sympad(X0)
    pad4 = ((1, 1), (1, 1), (1, 1), (0, 0))
    D = np.pad(DIFF0, pad4, mode='symmetric')
    A = np.pad(AVEC0, pad4, mode='symmetric')
    st, dl = D[..., 0], D[..., 1]
    # symmetric tensor components indexed by the (unordered) axis pair
    C = {(0, 0): st + A[..., 0] * dl, (0, 1): A[..., 1] * dl, (0, 2): A[..., 2] * dl,
         (1, 1): st + A[..., 3] * dl, (1, 2): A[..., 4] * dl, (2, 2): st + A[..., 5] * dl}
    h = (F(DX), F(DY), F(DZ))
    n = X0.shape
    half = F(0.5)

    def at(arr, off):
        return arr[1 + off[0]:1 + off[0] + n[0], 1 + off[1]:1 + off[1] + n[1], 1 + off[2]:1 + off[2] + n[2]]

    def e(axis, s=1):
        v = [0, 0, 0]
        v[axis] = s
        return np.array(v)

    zero = np.zeros(3, dtype=int)
    flux = {}
    for a in range(3):
        for sg in (1, -1):
            total = None
            for b in range(3):
                if b == a:
                    Caa = C[(a, a)]
                    t = (at(Caa, sg * e(a)) + at(Caa, zero)) * (at(X, sg * e(a)) - at(X, zero)) / h[a]
                else:
                    Cab = C[(min(a, b), max(a, b))]
                    p_, m_ = sg * e(a) + e(b), sg * e(a) - e(b)
                    t = half * (at(Cab, p_) + at(Cab, m_)) * (at(X, p_) - at(X, m_)) / h[b]
                total = t if total is None else total + t
            flux[(a, sg)] = (half if sg > 0 else F(-0.5)) * total
    lap = ((flux[(0, 1)] - flux[(0, -1)]) / h[0] + (flux[(1, 1)] - flux[(1, -1)]) / h[1]
           + (flux[(2, 1)] - flux[(2, -1)]) / h[2])
    return lap.astype(np.float32)


def laplace_homog_2d(X0, DX, DY):
    """gpuSolve/diffop2D/laplace_homogeneous_isotropic_diffusion.py:3-36."""
    X = _sympad(X0)
    DX, DY = F(DX), F(DY)
    dxsq, dysq = DX * DX, DY * DY
    two = F(2.0)
    c = X[1:-1, 1:-1]
    lapla = ((X[0:-2, 1:-1] - two * c + X[2:, 1:-1]) / dxsq
             + (X[1:-1, 0:-2] - two * c + X[1:-1, 2:]) / dysq)
    return lapla.astype(np.float32)


def laplace_heterog_2d(X0, DIFF0, DX, DY):
    """gpuSolve/diffop2D/laplace_heterogeneous_isotropic_diffusion.py:3-32."""
    X = _sympad(X0)
    D = _sympad(DIFF0)
    DX, DY = F(DX), F(DY)
    h, mh = F(0.5), F(-0.5)
    c = X[1:-1, 1:-1]
    dc = D[1:-1, 1:-1]
    e = h * (D[2:, 1:-1] + dc) * (X[2:, 1:-1] - c)
    w = mh * (D[0:-2, 1:-1] + dc) * (X[0:-2, 1:-1] - c)
    n = h * (D[1:-1, 2:] + dc) * (X[1:-1, 2:] - c)
    s = mh * (D[1:-1, 0:-2] + dc) * (X[1:-1, 0:-2] - c)
    dxsq, dysq = DX * DX, DY * DY
    return ((e - w) / dxsq + (n - s) / dysq).astype(np.float32)


def fd_step(model, U, dt, diff, spacing, DIFF=None, conv=False, dirichlet=None):
    """One explicit step as the FD examples' ``solve`` writes it.

    Homogeneous (Tests/FD/Fenton/fenton.py:90-99):
        U1 = U0 + dt*dU(U) + (diff*dt)*laplace(U0, DX, DY, DZ)
    Heterogeneous (Tests/FD/Fenton_sphere/fenton.py:109-115, DIFF carries the diffusivity):
        U1 = U0 + dt*dU(U) + dt*laplace(U0, DIFF, DX, DY, DZ)
    ``model`` is an oracle ionic model (or None for the heat equation,
    Tests/FD/HeatEquation/heat.py:86-94).  ``dt`` and ``diff`` are Python floats: their
    product is formed in float64 and rounded to fp32 where it meets the tensor.
    ``dirichlet``: constant boundary value of Tests/FD/LaplaceSolver/laplaceSolver.py:44-52,136-141 (U0 = constant pad
    of the interior instead of the symmetric one).
    """
    U0 = enforce_boundary(U) if dirichlet is None else enforce_boundary_dirichlet(U, dirichlet)
    three_d = U.ndim == 3
    if DIFF is None:
        if three_d:
            lap = (laplace_conv_homog_3d if conv else laplace_homog_3d)(U0, *spacing)
        else:
            lap = laplace_homog_2d(U0, *spacing)
        coef = F(float(diff) * float(dt))
    else:
        lap = (laplace_heterog_3d if three_d else laplace_heterog_2d)(U0, DIFF, *spacing)
        coef = F(dt)
    if model is not None:
        dU = model.differentiate(U)
        U1 = U0 + F(dt) * dU + coef * lap
    else:
        U1 = U0 + coef * lap
    return U1.astype(np.float32)


def apply_stimulus(U, stim):
    """Tests/FD/Fenton/fenton.py:154-156: U = where(bool(stim), max(U, stim), U)."""
    return np.where(stim.astype(bool), np.maximum(U, stim), U).astype(np.float32)


def apply_stimulus_max(U, stim):
    """Tests/FD/HeatEquation/heat.py:143-144: U = maximum(U, s2()) on every node."""
    return np.maximum(U, stim).astype(np.float32)


def fd_run(model, U, nsteps, dt, diff, spacing, DIFF=None, conv=False, s2=None, step0=0,
           act_threshold=None):
    """The examples' run() loop (Tests/FD/Fenton/fenton.py:150-159) without I/O.

    ``s2`` is an oracle Stimulus (integer-step activation).  When ``act_threshold`` is
    given, also returns the first step index at which each node crosses it upwards
    (activation time in steps, -1 if never), as Tests/CICD/unit/test_mms_1d.py:91-101.
    """
    act = None
    if act_threshold is not None:
        act = np.full(U.shape, -1, dtype=np.int32)
    for i in range(step0, step0 + nsteps):
        U1 = fd_step(model, U, dt, diff, spacing, DIFF=DIFF, conv=conv)
        if s2 is not None and s2.stimulate_tissue_timestep(i, dt):
            U1 = apply_stimulus(U1, s2())
        if act is not None:
            crossed = (U < F(act_threshold)) & (U1 >= F(act_threshold)) & (act < 0)
            act[crossed] = i
        U = U1
    return (U, act) if act is not None else U
