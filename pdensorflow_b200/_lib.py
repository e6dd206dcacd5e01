"""ctypes binding of libpdx.so (the C ABI in include/pdx.h).

There is no CPU fallback: if the library is missing, or a compute entry point fails,
the caller gets an exception.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libpdx.so')

MODEL_NONE, MODEL_FENTON4V, MODEL_MMS2V, MODEL_MS2V, MODEL_CRN, MODEL_TTP = 0, 1, 2, 3, 4, 5
LAP_HOMOG, LAP_CONV, LAP_HETEROG, LAP_ANISO = 0, 1, 2, 3
MATH_EXACT, MATH_FAST = 0, 1
KERNEL_AUTO, KERNEL_GENERIC, KERNEL_TMA, KERNEL_RESIDENT = 0, 1, 2, 3
MAX_PARAMS, MAX_STATES = 24, 4
LUT_MAX_STATES, LUT_MAX_CONSTS = 24, 64


class PdxError(RuntimeError):
    pass


class FdDesc(C.Structure):
    """struct pdx_fd_desc (include/pdx.h)."""
    _fields_ = [
        ('ndim', C.c_int32), ('nx', C.c_int32), ('ny', C.c_int32), ('nz', C.c_int32),
        ('dx', C.c_float), ('dy', C.c_float), ('dz', C.c_float),
        ('dt', C.c_float), ('coef', C.c_float),
        ('lap_kind', C.c_int32), ('model', C.c_int32), ('math_mode', C.c_int32), ('kernel', C.c_int32),
        ('x_begin', C.c_int32), ('x_end', C.c_int32), ('x_clamp_lo', C.c_int32), ('x_clamp_hi', C.c_int32),
        ('dirichlet', C.c_int32), ('dirichlet_value', C.c_float),
        ('x_chunk', C.c_int32),
        ('params', C.c_float * MAX_PARAMS),
        ('pitch_z', C.c_int32),
    ]


class LutTable(C.Structure):
    """struct pdx_lut_table (include/pdx.h)."""
    _fields_ = [('data', C.c_void_p), ('nrows', C.c_int32), ('ncols', C.c_int32), ('stride', C.c_int32),
                ('mn', C.c_float), ('mx', C.c_float), ('res', C.c_float), ('step', C.c_float), ('mx_idx', C.c_int32)]


class LutModel(C.Structure):
    """struct pdx_lut_model (include/pdx.h)."""
    _fields_ = [('model', C.c_int32), ('nconsts', C.c_int32), ('dt', C.c_float), ('tab', LutTable * 3),
                ('consts', C.c_float * LUT_MAX_CONSTS), ('math_mode', C.c_int32)]


class PcgPartDesc(C.Structure):
    """struct pdx_pcg_part_desc (include/pdx.h)."""
    _fields_ = [('n_local', C.c_int32), ('ghost_lo', C.c_int32), ('ghost_hi', C.c_int32),
                ('rank', C.c_int32), ('world', C.c_int32), ('send_lo', C.c_int32), ('send_hi', C.c_int32),
                ('lower_window_index', C.c_int32), ('upper_window_index', C.c_int32),
                ('max_blocks', C.c_int32), ('timeout_ms', C.c_int32), ('sell', C.c_int32),
                ('rowptr', C.c_void_p), ('col', C.c_void_p), ('aval', C.c_void_p), ('mval', C.c_void_p),
                ('minv', C.c_void_p)]


PART_MAX_WORLD = 16
_lib = None

_P = C.c_void_p
_SIGNATURES = {
    'pdx_version': (C.c_int, []),
    'pdx_last_error': (C.c_char_p, []),
    'pdx_launch_count': (C.c_int64, []),
    'pdx_model_num_params': (C.c_int, [C.c_int]),
    'pdx_model_num_states': (C.c_int, [C.c_int]),
    'pdx_model_default_params': (C.c_int, [C.c_int, C.POINTER(C.c_float)]),
    'pdx_fd_desc_init': (None, [C.POINTER(FdDesc)]),
    'pdx_fd_plan_create': (C.c_int, [C.POINTER(FdDesc), C.POINTER(_P)]),
    'pdx_fd_plan_destroy': (None, [_P]),
    'pdx_fd_plan_kernel': (C.c_int, [_P]),
    'pdx_fd_plan_resident_capable': (C.c_int, [_P]),
    'pdx_fd_plan_launches': (C.c_int64, [_P]),
    'pdx_fd_step': (C.c_int, [_P, _P, _P, C.POINTER(_P), _P, C.c_int, _P]),
    'pdx_fd_step_mirror': (C.c_int, [_P, _P, _P, C.POINTER(_P), _P, _P, _P, _P]),
    'pdx_fd_step_range': (C.c_int, [_P, _P, _P, C.POINTER(_P), _P, C.c_int, C.c_int, _P]),
    'pdx_fd_laplace': (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                 C.c_int, C.c_int, _P, _P, _P, _P]),
    'pdx_fd_enforce_boundary': (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P, _P]),
    'pdx_ionic_step': (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(_P), C.c_int64, _P,
                                 C.POINTER(_P), _P, C.c_float, _P, _P, _P]),
    'pdx_ionic_lut_step': (C.c_int, [C.POINTER(LutModel), C.c_int64, _P, C.POINTER(_P), _P, _P, _P, _P]),
    'pdx_fd_plan_set_lut': (C.c_int, [_P, C.POINTER(LutModel)]),
    'pdx_fd_plan_set_avec': (C.c_int, [_P, _P]),
    'pdx_fd_plan_aniso_tile_capable': (C.c_int, [_P]),
    'pdx_fd_plan_prepare_aniso': (C.c_int, [_P, _P, _P, _P]),
    'pdx_fd_laplace_aniso': (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, _P, _P, _P, _P, _P]),
    'pdx_stim_apply': (C.c_int, [_P, _P, C.c_float, C.c_int64, _P]),
    'pdx_stim_apply_max': (C.c_int, [_P, _P, C.c_float, C.c_int64, _P]),
    'pdx_csr_spmv': (C.c_int, [C.c_int32, _P, _P, _P, _P, _P, _P]),
    'pdx_pcg_create': (C.c_int, [C.c_int32, _P, _P, _P, _P, C.POINTER(_P)]),
    'pdx_pcg_destroy': (None, [_P]),
    'pdx_pcg_solve': (C.c_int, [_P, _P, _P, C.c_int, C.c_double, C.c_int, _P, _P]),
    'pdx_pcg_solve_rhs0': (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_double, C.c_int, _P, _P]),
    'pdx_csr_hint': (C.c_int, [_P, C.c_int32, C.c_int64]),
    'pdx_sell_hint': (C.c_int, [_P, C.c_int32]),
    'pdx_fem_explicit_step': (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int32, _P, _P, _P, _P, _P, _P,
                                        C.POINTER(_P), _P, C.c_float, _P]),
    'pdx_fem_explicit_step_part': (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int32, C.c_int32, _P, _P, _P, _P, _P,
                                             _P, C.POINTER(_P), _P, C.c_float, C.c_int32, _P, C.c_int32, _P, _P]),
    'pdx_fem_explicit_step_sell': (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int32, C.c_int32, _P, _P, _P, _P, _P,
                                             _P, C.POINTER(_P), _P, C.c_float, C.c_int32, _P, C.c_int32, _P, _P]),
    'pdx_fem_explicit_step_sync': (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int32, C.c_int32, _P, _P, _P, _P, _P,
                                             _P, C.POINTER(_P), _P, C.c_float, C.c_int32, _P, C.c_int32, _P, _P, _P, _P, _P]),
    'pdx_fem_assemble': (C.c_int, [C.c_int, C.c_int64, _P, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'pdx_round_to_f32': (C.c_int, [_P, _P, C.c_int64, _P]),
    'pdx_sell_slice_widths': (C.c_int, [C.c_int32, _P, _P, _P]),
    'pdx_sell_fill': (C.c_int, [C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, _P]),
    'pdx_push_rows': (C.c_int, [_P, C.c_int32, _P, _P, _P, _P]),
    'pdx_pcg_part_workspace_bytes': (C.c_int64, [C.c_int32]),
    'pdx_pcg_part_create': (C.c_int, [C.POINTER(PcgPartDesc), C.POINTER(_P), C.c_int32, C.POINTER(_P)]),
    'pdx_pcg_part_destroy': (None, [_P]),
    'pdx_pcg_part_blocks': (C.c_int32, [_P]),
    'pdx_pcg_part_threads': (C.c_int32, [_P]),
    'pdx_pcg_part_stage_bytes': (C.c_int32, [_P]),
    'pdx_pcg_part_set_profile': (C.c_int, [_P, _P, C.c_int32]),
    'pdx_pcg_part_solve': (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_double, C.c_int, _P, _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib():
    """Load libpdx.so once.  Raises PdxError (never falls back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PdxError('libpdx.so not built: run `python pdensorflow_b200/csrc/build.py` '
                           '(or __graft_entry__.build()); there is no CPU fallback')
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise PdxError(lib().pdx_last_error().decode('utf-8', 'replace') or 'libpdx call failed (%d)' % rc)


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def ptr_array(tensors, n=None):
    tensors = list(tensors)
    if n is None:
        n = MAX_STATES if len(tensors) <= MAX_STATES else LUT_MAX_STATES
    arr = (C.c_void_p * n)()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def stream_ptr(stream=None):
    import torch
    s = torch.cuda.current_stream() if stream is None else stream
    return C.c_void_p(s.cuda_stream)


def default_params(model):
    buf = (C.c_float * MAX_PARAMS)()
    check(lib().pdx_model_default_params(model, buf))
    return list(buf)[:lib().pdx_model_num_params(model)]


def require_cuda(t, name='tensor'):
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise PdxError('%s must be a CUDA torch.Tensor: this library has no CPU path' % name)
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise PdxError('%s must be a contiguous float32 tensor' % name)
    return t
