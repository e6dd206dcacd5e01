"""HeatSolver - mirror of gpuSolve/physics/heatSolver.py:28-193 on torch + libpdx.

One implicit Euler step of the heat equation, (MASS + dt*STIFFNESS) U1 = MASS (U0 + dt*I0): the caller
owns the time loop and passes t to step().  Per step: one SpMV launch + one PCG launch.
"""
import os
import numpy as np
import torch

from .. import _backend as B
from ... import _lib as L
from ..entities.triangulation import Triangulation
from ..entities.materialproperties import MaterialProperties
from ..matrices.globalMatrices import (assemble_matrices_dict, compute_coo_pattern,
                                       compute_reverse_cuthill_mckee_indexing, csr_axpby)
from ..matrices.localMass import localMass
from ..matrices.localStiffness import localStiffness
from ..linearsolvers.conjgrad import ConjGrad
from ..linearsolvers.jacobi_precond import JacobiPrecond
from ..force_terms import Stimulus


class HeatSolver:
    _U_DEFAULT = 0.0

    def __init__(self, cfgdict=None):
        self._mesh_file_name = None
        self._dt = 0.1
        self._dt_per_plot = 2
        self._Tend = 10
        self._use_renumbering = False
        self._explicit_lumped = False     # opt-in (not in the reference): explicit mass-lumped step
        self._graph_steps = True          # (not in the reference) replay a time step as a CUDA graph once its shape is known
        self._device_assembly = None      # None: on the device above DEVICE_ASSEMBLY_MIN_ELEMENTS elements; True / False
        if cfgdict is not None:
            for attribute in list(self.__dict__.keys()):
                if attribute[1:] in cfgdict.keys():
                    setattr(self, attribute, cfgdict[attribute[1:]])
        self._Domain = Triangulation()
        self._materials = MaterialProperties()
        self._Solver = ConjGrad()
        self._Precond = JacobiPrecond()
        self._MASS = None
        self._STIFFNESS = None
        self._U = None
        self._ready_for_run = False
        self._ctime = 0.0
        self._nbstim = 0
        self._renumbering = None
        self._StimulusDict = None
        self._nt = int(self._Tend // self._dt)
        self._I0_cache = {}
        self._step_graphs = {}
        self._rhs0 = None
        self._rhs = None
        self._perm_d = None
        self._iperm_d = None
        if self._mesh_file_name is not None:
            self._Domain.readMesh('{}'.format(self._mesh_file_name))

    # ---- setup --------------------------------------------------------------
    def loadMesh(self, fname):
        self._mesh_file_name = fname
        self._Domain.readMesh('{}'.format(self._mesh_file_name))

    def add_element_material_property(self, pname, ptype, prop):
        self._materials.add_element_property(pname, ptype, prop)

    def add_material_function(self, fname, fsign):
        self._materials.add_ud_function(fname, fsign)

    def assemble_matrices(self):
        from ..matrices import device_assembly as DA
        connectivity = self._Domain.mesh_connectivity('True')
        pattern = compute_coo_pattern(connectivity)
        if self._use_renumbering:
            self._renumbering = compute_reverse_cuthill_mckee_indexing(pattern)
        lmatr = {'mass': localMass, 'stiffness': localStiffness}
        nelem = sum(E.shape[0] for E in self._Domain.Elems().values())
        on_device = self._device_assembly
        if on_device is None:
            on_device = nelem >= DA.DEVICE_ASSEMBLY_MIN_ELEMENTS and DA.supported(self._Domain, self._materials)
        if on_device:      # one thread per element, fp64 atomics into CSR order (pdx_fem_assemble)
            MATRICES = DA.assemble_on_device(self._Domain, self._materials, renumbering=self._renumbering)
        else:
            MATRICES = assemble_matrices_dict(lmatr, pattern, self._Domain, self._materials, connectivity,
                                              renumbering=self._renumbering)
        self._MASS = MATRICES['mass']
        self._STIFFNESS = MATRICES['stiffness']
        A = csr_axpby(self._MASS, 1.0, self._STIFFNESS, self._dt)
        self._Domain.release_connectivity()
        self._materials.remove_all_element_properties()
        self._Solver.set_matrix(A)
        self._Solver.fuse_mass(self._MASS)          # large systems: RHS = MASS*RHS0 is formed inside the solve kernel
        I, J, V = A.coo()
        self._Precond.build_preconditioner(I, J, V, A.n)
        self._Solver.set_precond(self._Precond)
        if self._explicit_lumped:
            ml = self._MASS.diagonal_rowsum()
            self._mlinv = B.as_tensor((np.float32(1.0) / ml).astype(np.float32))

    def set_initial_condition(self, U0=None):
        npt = self._Domain.Pts().shape[0]
        if U0 is not None:
            U0 = np.asarray(U0, dtype=np.float32)
            self._U = B.as_tensor(U0[:, np.newaxis] if U0.ndim == 1 else U0)
        else:
            self._U = torch.full((npt, 1), float(self._U_DEFAULT), dtype=torch.float32, device=B.device())

    def add_stimulus(self, stimreg, stimprops):
        self._nbstim += 1
        if self._StimulusDict is None:
            self._StimulusDict = {}
        self._StimulusDict[self._nbstim] = Stimulus(stimprops)
        self._StimulusDict[self._nbstim].set_stimregion(stimreg)

    def _perm_tensors(self):
        if self._perm_d is None:
            dev = self._U.device
            self._perm_d = torch.as_tensor(self._renumbering['perm'].astype(np.int64), device=dev)
            self._iperm_d = torch.as_tensor(self._renumbering['iperm'].astype(np.int64), device=dev)
        return self._perm_d, self._iperm_d

    def finalize_for_run(self):
        if self._use_renumbering:
            perm, _ = self._perm_tensors()
            self._U = self._U.index_select(0, perm).contiguous()
            if self._StimulusDict is not None:
                for _key, stim in self._StimulusDict.items():
                    stim.apply_indices_permutation(self._renumbering['perm'])
        self._ready_for_run = True

    # ---- per-step kernel ----------------------------------------------------
    def _work(self):
        if self._rhs0 is None or self._rhs0.shape != self._U.shape:
            self._rhs0 = torch.empty_like(self._U)
            self._rhs = torch.empty_like(self._U)
        return self._rhs0, self._rhs

    def solve_step(self, U, I0):
        """Implicit Euler solve for the heat equation (heatSolver.py:127-135)."""
        rhs0, rhs = self._work()
        torch.add(U, I0, alpha=float(np.float32(self._dt)), out=rhs0)
        return self._solve_from_rhs0(U, rhs0, rhs)

    def _solve_from_rhs0(self, U, rhs0, rhs):
        """(MASS + dt*STIFFNESS) U1 = MASS*RHS0 from the guess U."""
        self._Solver.set_X0(U)
        if self._Solver.fused_mass_ready():
            self._Solver.solve_from_rhs0(rhs0, self._Solver.X())
        else:
            self._MASS.matvec(rhs0, out=rhs)
            self._Solver.solve_inplace(rhs, self._Solver.X())
        return self._Solver.X()

    def _compute_forcing(self, ctime):
        """Sum of the active stimuli at time ctime, evaluated on the fp32 value of the clock like the
        reference (heatSolver.py:137-145).  Fields are cached per set of active stimuli, so a step
        does no allocation / upload."""
        active = []
        if self._StimulusDict is not None:
            t32 = np.float32(ctime)
            for key, stim in self._StimulusDict.items():
                if stim.stimulate_tissue_timevalue(t32):
                    active.append(key)
        # the reference recomputes intensity*mask every step: key the cached field on everything it depends on
        # (set_intensity / set_stimregion / apply_indices_permutation replace the value or the mask tensor)
        k = tuple((key, float(self._StimulusDict[key].intensity()), self._StimulusDict[key].get_stimregion().data_ptr())
                  for key in active)
        I0 = self._I0_cache.get(k)
        if I0 is None:
            I0 = torch.zeros_like(self._U)
            for key in active:
                I0 = I0 + self._StimulusDict[key]()
            self._I0_cache[k] = I0
        return I0

    # ---- CUDA-graph replay of a time step ------------------------------------------------------
    # The reference's loop calls step(t) once per time step (Tests/FEM/Fenton/fenton.py:119-132); on a 63 001-node mesh
    # the two kernels of a step take ~0.06-0.08 ms on the device, about what Python needs to prepare and enqueue them.
    # A step is fully described by (potential buffer, forcing field, ionic parameters, solver settings): the first step
    # with a given description runs eagerly (allocations, caches), the second is captured into a CUDA graph, every later
    # one is a single graph launch.  cfg key ``graph_steps`` (default True) / PDX_STEP_GRAPH=0 turn it off.
    def _step_token(self):
        # everything a captured step froze besides the buffers: solver settings and the matrices behind the handles
        return (int(self._Solver.maxiter()), float(self._Solver.toll()), float(self._dt), id(self._MASS), id(self._STIFFNESS),
                id(getattr(self._Solver, '_A', None)), id(getattr(self._Solver, '_handle', None)), id(getattr(self._Solver, '_part', None)))

    def _graph_step(self, I0):
        if os.environ.get('PDX_STEP_GRAPH', '1') == '0' or torch.cuda.is_current_stream_capturing():
            return self.solve_step(self._U, I0)
        key = (self._U.data_ptr(), I0.data_ptr(), self._step_token())
        ent = self._step_graphs.get(key)
        if ent is None:                                   # first visit: eager
            if len(self._step_graphs) > 16:
                self._step_graphs.clear()
            self._step_graphs[key] = 'seen'
            return self.solve_step(self._U, I0)
        if ent == 'eager':
            return self.solve_step(self._U, I0)
        if ent == 'seen':                                 # second visit: capture (executes nothing), then replay
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            g = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.stream(side):
                    with torch.cuda.graph(g, stream=side):
                        out = self.solve_step(self._U, I0)
            except Exception:
                torch.cuda.synchronize()
                self._step_graphs[key] = 'eager'          # e.g. a launch kind the driver cannot capture
                return self.solve_step(self._U, I0)
            cur.wait_stream(side)
            ent = (g, out, self._U, I0)                   # the graph reads / writes these tensors: keep them alive
            self._step_graphs[key] = ent
        ent[0].replay()
        return ent[1]

    def step(self, t):
        """Advance one time step; t is the absolute simulation time at the end of the step."""
        if not self._ready_for_run:
            raise Exception("model not initialised for run!")
        I0 = self._compute_forcing(t)
        U1 = self._graph_step(I0) if self._graph_steps else self.solve_step(self._U, I0)
        self._U = U1
        self._ctime = t
        return self._U

    # ---- accessors ----------------------------------------------------------
    def domain(self):
        return self._Domain

    def solver(self):
        return self._Solver

    def precond(self):
        return self._Precond

    def stimulus(self):
        return self._StimulusDict

    def U(self):
        if self._use_renumbering:
            _, iperm = self._perm_tensors()
            return self._U.index_select(0, iperm)
        return self._U

    def renumbering(self):
        return self._renumbering

    def nt(self):
        return self._nt

    def dt(self):
        return self._dt

    def dt_per_plot(self):
        return self._dt_per_plot

    def ctime(self):
        return self._ctime

    def Tend(self):
        return self._Tend
