"""No-GPU checks of the boundary: libpdx.so loads, exports every symbol include/pdx.h declares, the
ctypes signatures cover them, argument validation fails loudly, and nothing in the product imports the
oracle or offers a CPU fallback."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'pdx.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(pdx_[a-z0-9_]+)\s*\(', src)))


@pytest.fixture(scope='module')
def lib():
    from pdensorflow_b200.csrc import build
    build.build()
    from pdensorflow_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol(lib):
    handle = ctypes.CDLL(lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(handle, name), 'libpdx.so does not export %s' % name
    assert set(lib.EXPORTED_SYMBOLS) == set(declared), set(lib.EXPORTED_SYMBOLS) ^ set(declared)


def test_host_only_entry_points(lib):
    h = lib.lib()
    assert h.pdx_version() >= 110
    assert h.pdx_model_num_states(lib.MODEL_FENTON4V) == 3 and h.pdx_model_num_states(lib.MODEL_MMS2V) == 1
    assert h.pdx_model_num_params(lib.MODEL_FENTON4V) == 23 and h.pdx_model_num_states(99) == -1
    p = lib.default_params(lib.MODEL_FENTON4V)
    assert abs(p[0] - 3.33) < 1e-6 and p[21] == -80.0 and p[22] == 100.0
    assert lib.default_params(lib.MODEL_MS2V)[0] == pytest.approx(0.3)
    assert h.pdx_model_num_states(lib.MODEL_CRN) == 19 and h.pdx_model_num_states(lib.MODEL_TTP) == 22
    assert h.pdx_model_num_params(lib.MODEL_CRN) == 23 and h.pdx_model_num_params(lib.MODEL_TTP) == 45


def test_argument_validation_fails_loudly(lib):
    d = lib.FdDesc()
    lib.lib().pdx_fd_desc_init(ctypes.byref(d))
    d.ndim, d.nx, d.ny, d.nz = 3, 2, 8, 8           # fewer than 3 planes
    h = ctypes.c_void_p()
    assert lib.lib().pdx_fd_plan_create(ctypes.byref(d), ctypes.byref(h)) != 0
    assert b'at least 3' in lib.lib().pdx_last_error()
    with pytest.raises(lib.PdxError):
        lib.check(lib.lib().pdx_stim_apply(None, None, 1.0, 10, None))
    with pytest.raises(lib.PdxError):
        lib.check(lib.lib().pdx_csr_spmv(5, None, None, None, None, None, None))


def test_no_cpu_fallback_and_no_oracle_in_product():
    import torch
    from pdensorflow_b200 import _lib
    with pytest.raises(_lib.PdxError):
        _lib.require_cuda(torch.zeros(4), 'x')            # CPU tensors are refused
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'pdensorflow_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                if re.search(r'^\s*(from|import)\s+oracle\b', txt, flags=re.M) or 'import tensorflow' in txt \
                        or 'import triton' in txt:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
    if not torch.cuda.is_available():
        from pdensorflow_b200.fd import FDStepper
        with pytest.raises(Exception):
            FDStepper((8, 8, 8), (1.0, 1.0, 1.0), 0.1, 0.8)


def test_sass_has_tma_and_no_legacy_tensor_path():
    """The fused kernel stages U through TMA (UTMALDG in SASS) and arrives on mbarriers (SYNCS)."""
    so = os.path.join(ROOT, 'pdensorflow_b200', 'libpdx.so')
    try:
        sass = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
    except FileNotFoundError:
        sass = ''
    if not sass:
        pytest.skip('cuobjdump unavailable')
    assert 'UTMALDG' in sass and 'SYNCS' in sass
    assert 'sm_100a' in subprocess.run(['cuobjdump', '-lelf', so], capture_output=True, text=True).stdout
    # partitioned PCG: 8-byte (epoch, value) words published / acquired at SYSTEM scope (peer memory over NVLink)
    assert 'STG.E.64.STRONG.SYS' in sass and 'LDG.E.64.STRONG.SYS' in sass
    # resident 2-D kernel (two nodes per thread): 64-bit shared-memory words, GPU-scope flag release / acquire
    assert 'LDS.64' in sass and 'STS.64' in sass and 'STG.E.STRONG.GPU' in sass and 'LDG.E.STRONG.GPU' in sass


def test_dropin_alias_registers_gpusolve():
    code = ("import pdensorflow_b200.dropin, sys; import gpuSolve; from gpuSolve.diffop3D import laplace_homog, "
            "laplace_conv_homog, laplace_heterog, laplace_heterog_aniso; from gpuSolve.diffop2D import laplace_homog as l2;"
            "from gpuSolve.ionic.fenton4v import *; from gpuSolve.ionic.mms2v import ModifiedMS2v; "
            "from gpuSolve.ionic.ms2v import MitchellSchaeffer2v; from gpuSolve.force_terms import Stimulus; "
            "from gpuSolve.physics import MonodomainSolver, HeatSolver; from gpuSolve.linearsolvers import ConjGrad, JacobiPrecond;"
            "from gpuSolve.matrices import csr_axpby, localMass; assert Fenton4v().tau_vp() is not None; print('ok')")
    r = subprocess.run([sys.executable, '-c', code], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0 and 'ok' in r.stdout, r.stderr


def test_package_imports_and_reports_version():
    """Tests/CICD/unit/test_import.py: the package imports (through the drop-in alias) and version() is a dotted numeric
    string; the IO / entities sub-packages the examples import resolve too."""
    code = ("import pdensorflow_b200.dropin, gpuSolve; v = gpuSolve.version(); parts = v.split('.'); "
            "assert isinstance(v, str) and len(parts) >= 2 and all(p.isdigit() for p in parts), v; "
            "import gpuSolve.IO.writers as w, gpuSolve.IO.readers as r, gpuSolve.entities as e; "
            "assert w.IGBWriter and w.ResultWriter and w.BaseWriter and r.IGBReader and e.Domain3D and e.Triangulation; "
            "print('ok')")
    r = subprocess.run([sys.executable, '-c', code], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0 and 'ok' in r.stdout, r.stderr
