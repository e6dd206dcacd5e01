"""A NumPy-backed stand-in for the slice of the TensorFlow API that PDEnsorflow's hot path uses.

TEST INFRASTRUCTURE - used only by the golden-vector generators in this directory, in the
build container where /root/reference exists.  TensorFlow is not installable in this image
(no network, no wheel), so the reference's kernels cannot be run; what CAN be run is the
reference's own Python source, unmodified, with every ``tf.*`` call it makes mapped onto the
NumPy operation that has the same IEEE-754 fp32 semantics.  ``install()`` registers this
module as ``tensorflow`` (plus the CSR wrapper module the reference imports) and makes
``import gpuSolve`` resolve to /root/reference/PDEnsorflow/gpuSolve without executing its
``__init__`` (which re-execs the interpreter to patch LD_LIBRARY_PATH).

Semantics reproduced
  * dtype discipline: tensors are fp32 unless created otherwise; a Python scalar meeting a
    tensor is converted to the tensor's dtype (TF's convert_to_tensor rule, NumPy 2 weak
    scalars); NumPy float64 scalars are demoted the same way.
  * tf.pad SYMMETRIC = np.pad(mode='symmetric'); tf.sign(0) = 0; tf.where is strict select;
    tf.cast(float -> int32) truncates toward zero; tf.gather is axis-0 take;
    tf.clip_by_value = np.clip.
  * elementwise transcendental functions (tanh, exp, log, expm1) return the correctly rounded
    fp32 value (fp64 evaluation rounded once): platform independent, and within 1-2 ulp of
    TF's Eigen/XLA kernels, which are not bit-reproducible across TF versions either.
  * reductions (tf.reduce_sum) use NumPy's pairwise fp32 summation; TF's order is unspecified.
  * CSR matrices are SciPy csr_matrix fp32; SparseMatrixMatMul = csr @ dense in fp32.
Everything is eager; tf.function is the identity decorator.
"""
import sys
import types

import numpy as np
import scipy.sparse as sp

REF_ROOT = '/root/reference/PDEnsorflow'


# --------------------------------------------------------------------------- dtypes
class DType:
    def __init__(self, name, np_dtype):
        self.name, self.as_numpy_dtype = name, np_dtype

    def __repr__(self):
        return 'tf.' + self.name

    def __eq__(self, other):
        return isinstance(other, DType) and other.name == self.name

    def __hash__(self):
        return hash(self.name)


float32 = DType('float32', np.float32)
float64 = DType('float64', np.float64)
int32 = DType('int32', np.int32)
int64 = DType('int64', np.int64)
bool = DType('bool', np.bool_)  # noqa: A001 - mirrors tf.bool
_BY_ENUM = {1: float32, 2: float64, 3: int32, 9: int64, 10: bool}   # DataType enum values (tf.constant(dtype=1))


def _np_dtype(dtype):
    if dtype is None:
        return None
    if isinstance(dtype, DType):
        return dtype.as_numpy_dtype
    if isinstance(dtype, int):
        return _BY_ENUM[dtype].as_numpy_dtype
    return np.dtype(dtype).type


# --------------------------------------------------------------------------- tensor
class Tensor(np.ndarray):
    """ndarray subclass carrying the few tf.Tensor/tf.Variable methods the reference calls."""

    name = 'Tensor:0'

    def __new__(cls, value):
        return np.asarray(value).view(cls)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        # TF converts non-tensor operands to the tensor's dtype: demote NumPy float64 scalars
        ins = []
        for x in inputs:
            if isinstance(x, Tensor):
                x = x.view(np.ndarray)
            elif isinstance(x, np.floating):
                x = float(x)
            elif isinstance(x, np.integer):
                x = int(x)
            ins.append(x)
        if 'out' in kwargs:
            kwargs['out'] = tuple(o.view(np.ndarray) if isinstance(o, Tensor) else o for o in kwargs['out'])
        res = getattr(ufunc, method)(*ins, **kwargs)
        if isinstance(res, tuple):
            return tuple(_wrap(r) for r in res)
        return _wrap(res)

    def numpy(self):
        return np.array(self.view(np.ndarray))

    def assign(self, value):
        v = np.asarray(value)
        if v.shape != self.shape:
            v = v.reshape(self.shape)
        self.view(np.ndarray)[...] = v.astype(self.dtype, copy=False)
        return self

    def assign_add(self, value):
        return self.assign(self + value)

    def assign_sub(self, value):
        return self.assign(self - value)

    def get_shape(self):
        return self.shape

    def __bool__(self):
        return builtins_bool(self.view(np.ndarray))

    def __float__(self):
        return float(self.view(np.ndarray).reshape(()))

    def __int__(self):
        return int(self.view(np.ndarray).reshape(()))

    def __hash__(self):
        return id(self)


import builtins as _b  # noqa: E402
builtins_bool = _b.bool


def _wrap(x):
    if isinstance(x, np.ndarray) and not isinstance(x, Tensor):
        return x.view(Tensor)
    if isinstance(x, np.generic):
        return np.asarray(x).view(Tensor)
    return x


def _raw(x):
    return x.view(np.ndarray) if isinstance(x, Tensor) else x


def convert_to_tensor(value, dtype=None, name=None):
    npd = _np_dtype(dtype)
    if isinstance(value, Tensor) and (npd is None or value.dtype == npd):
        return value
    a = np.asarray(_raw(value))
    if npd is None:
        if a.dtype == np.float64 and not isinstance(value, np.ndarray):
            npd = np.float32            # Python floats / lists of floats become tf.float32
        elif a.dtype == np.int64 and not isinstance(value, np.ndarray):
            npd = np.int32
    if npd is not None:
        a = a.astype(npd)
    return np.array(a).view(Tensor)


def constant(value, dtype=None, shape=None, name=None):
    t = convert_to_tensor(value, dtype)
    if shape is not None:
        t = np.broadcast_to(_raw(t), shape).copy().view(Tensor)
    return t


def Variable(initial_value, trainable=None, name=None, dtype=None, **kw):   # noqa: N802
    t = convert_to_tensor(initial_value, dtype)
    v = np.array(_raw(t)).view(Tensor)              # own storage (assign() mutates it)
    v.name = (name or 'Variable') + ':0'
    return v


tensor = constant
identity = lambda x, name=None: np.array(_raw(convert_to_tensor(x))).view(Tensor)  # noqa: E731


def function(func=None, **kw):
    if func is None:
        return lambda f: f
    return func


# --------------------------------------------------------------------------- ops
def _f(x):
    """operand -> ndarray, Python scalars left alone (weak)."""
    return _raw(x)


def shape(x, out_type=None):
    return tuple(np.shape(_raw(x)))


def reshape(x, shp, name=None):
    shp = tuple(int(s) for s in np.atleast_1d(_raw(shp))) if not isinstance(shp, tuple) else shp
    return np.reshape(_raw(convert_to_tensor(x)), shp).view(Tensor)


def fill(dims, value, name=None):
    v = convert_to_tensor(value)
    return np.full(tuple(int(d) for d in dims), _raw(v), dtype=v.dtype).view(Tensor)


def ones_like(x, dtype=None):
    return np.ones_like(_raw(x), dtype=_np_dtype(dtype)).view(Tensor)


def zeros_like(x, dtype=None):
    return np.zeros_like(_raw(x), dtype=_np_dtype(dtype)).view(Tensor)


def zeros(shp, dtype=float32):
    return np.zeros(shp, dtype=_np_dtype(dtype)).view(Tensor)


def ones(shp, dtype=float32):
    return np.ones(shp, dtype=_np_dtype(dtype)).view(Tensor)


def cast(x, dtype, name=None):
    a = _raw(convert_to_tensor(x))
    npd = _np_dtype(dtype)
    if np.issubdtype(npd, np.integer) and np.issubdtype(a.dtype, np.floating):
        a = np.trunc(a)
    return a.astype(npd).view(Tensor)


def expand_dims(x, axis, name=None):
    return np.expand_dims(_raw(x), axis).view(Tensor)


def squeeze(x, axis=None, name=None):
    return np.squeeze(_raw(x), axis=axis).view(Tensor)


def stack(values, axis=0, name=None):
    return np.stack([_raw(v) for v in values], axis=axis).view(Tensor)


def concat(values, axis, name=None):
    return np.concatenate([_raw(v) for v in values], axis=axis).view(Tensor)


def range(*a, dtype=None, **kw):  # noqa: A001
    r = np.arange(*[_raw(x) for x in a])
    npd = _np_dtype(dtype)
    return r.astype(npd if npd is not None else (np.int32 if np.issubdtype(r.dtype, np.integer) else np.float32)).view(Tensor)


def gather(params, indices, axis=0, name=None, **kw):
    return np.take(_raw(params), _raw(indices), axis=axis).view(Tensor)


def pad(x, paddings, mode='CONSTANT', constant_values=0, name=None):
    pw = [tuple(int(v) for v in row) for row in np.asarray(_raw(paddings))]
    m = mode.upper()
    a = _raw(x)
    if m == 'SYMMETRIC':
        return np.pad(a, pw, mode='symmetric').view(Tensor)
    if m == 'REFLECT':
        return np.pad(a, pw, mode='reflect').view(Tensor)
    cv = np.asarray(_raw(constant_values)).astype(a.dtype)
    return np.pad(a, pw, mode='constant', constant_values=cv).view(Tensor)


def where(cond, x=None, y=None, name=None):
    c = _raw(cond)
    if x is None:
        return np.argwhere(c).view(Tensor)
    xa, ya = _raw(x), _raw(y)
    # a Python scalar branch takes the other branch's dtype
    dt = None
    for v in (xa, ya):
        if isinstance(v, np.ndarray):
            dt = v.dtype
            break
    if dt is None:
        dt = np.float32
    return np.where(c, np.asarray(xa, dtype=dt), np.asarray(ya, dtype=dt)).view(Tensor)


def _binary(np_fn):
    def op(a, b, name=None):
        return _wrap(np_fn(convert_to_tensor(a) if not isinstance(a, (int, float)) else a,
                           convert_to_tensor(b) if not isinstance(b, (int, float)) else b))
    return op


add = _binary(np.add)
subtract = _binary(np.subtract)
multiply = _binary(np.multiply)
divide = _binary(np.divide)
maximum = _binary(np.maximum)
minimum = _binary(np.minimum)
logical_and = _binary(np.logical_and)
logical_or = _binary(np.logical_or)


def _unary(np_fn):
    def op(x, name=None):
        return _wrap(np_fn(convert_to_tensor(x)))
    return op


def _cr(np_fn):
    """Correctly rounded fp32 transcendental: evaluate in fp64, round once.  Platform
    independent (NumPy's fp32 SIMD kernels are not); TF's own kernels are within 1-2 ulp."""
    def op(x, name=None):
        a = _raw(convert_to_tensor(x))
        if a.dtype == np.float32:
            return np_fn(a.astype(np.float64)).astype(np.float32).view(Tensor)
        return _wrap(np_fn(a))
    return op


sign = _unary(np.sign)
tanh = _cr(np.tanh)
exp = _cr(np.exp)
abs = _unary(np.abs)  # noqa: A001
sqrt = _unary(np.sqrt)
square = _unary(np.square)
negative = _unary(np.negative)
logical_not = _unary(np.logical_not)


def clip_by_value(x, lo, hi, name=None):
    a = convert_to_tensor(x)
    return np.clip(_raw(a), np.asarray(_raw(lo), dtype=a.dtype), np.asarray(_raw(hi), dtype=a.dtype)).view(Tensor)


def reduce_sum(x, axis=None, keepdims=False, name=None):
    a = _raw(convert_to_tensor(x))
    if axis is None and not keepdims:
        # full reduction: flatten first so NumPy's pairwise summation applies whatever the shape
        # ([N,1] column vectors would otherwise be summed sequentially); TF/Eigen reduce with
        # vectorised partial sums + a tree, i.e. pairwise-like accuracy.
        a = np.ascontiguousarray(a).reshape(-1)
    return _wrap(np.sum(a, axis=axis, keepdims=keepdims, dtype=a.dtype))


def reduce_max(x, axis=None, keepdims=False, name=None):
    return _wrap(np.max(_raw(x), axis=axis, keepdims=keepdims))


def reduce_min(x, axis=None, keepdims=False, name=None):
    return _wrap(np.min(_raw(x), axis=axis, keepdims=keepdims))


def unique(x, out_idx=int32):
    """tf.unique: unique values in order of first occurrence + index of each input in them."""
    a = _raw(x)
    vals, first, inv = np.unique(a, return_index=True, return_inverse=True)
    order = np.argsort(first, kind='stable')
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    return vals[order].view(Tensor), rank[inv].astype(_np_dtype(out_idx)).view(Tensor)


def print(*a, **kw):  # noqa: A001
    pass


def while_loop(cond, body, loop_vars, **kw):
    lv = tuple(loop_vars)
    while builtins_bool(cond(*lv)):
        lv = tuple(body(*lv))
    return lv


def conv3d_valid(x, k):
    """tf.nn.conv3d, NDHWC, one channel, stride 1, VALID: correlation, taps accumulated in
    kernel memory order (the order is an implementation detail of TF's backend)."""
    a = _raw(x)[0, :, :, :, 0]
    w = _raw(k)[:, :, :, 0, 0].astype(a.dtype)
    kd, kh, kw_ = w.shape
    od, oh, ow = a.shape[0] - kd + 1, a.shape[1] - kh + 1, a.shape[2] - kw_ + 1
    out = np.zeros((od, oh, ow), dtype=a.dtype)
    for i in np.arange(kd):
        for j in np.arange(kh):
            for l in np.arange(kw_):
                if w[i, j, l] != 0:
                    out = out + w[i, j, l] * a[i:i + od, j:j + oh, l:l + ow]
    return out[None, :, :, :, None].view(Tensor)


nn = types.SimpleNamespace(conv3d=lambda x, filters, strides, padding, **kw: conv3d_valid(x, filters))


def _segment_sum(data, ids, num_segments):
    d = _raw(data)
    out = np.zeros((int(num_segments),) + d.shape[1:], dtype=d.dtype)
    np.add.at(out, _raw(ids), d)          # sequential fp32 accumulation in input order
    return out.view(Tensor)


math = types.SimpleNamespace(
    log=_cr(np.log), exp=exp, expm1=_cr(np.expm1), tanh=tanh, sign=sign, abs=abs, sqrt=sqrt,
    is_finite=_unary(np.isfinite), is_nan=_unary(np.isnan), maximum=maximum, minimum=minimum,
    scalar_mul=lambda s, x, name=None: _wrap(convert_to_tensor(s) * convert_to_tensor(x)),
    unsorted_segment_sum=_segment_sum, reduce_sum=reduce_sum, multiply=multiply, add=add,
    logical_and=logical_and, logical_not=logical_not)

config = types.SimpleNamespace(run_functions_eagerly=lambda flag: None,
                               list_physical_devices=lambda kind=None: [],
                               functions_run_eagerly=lambda: True)
errors = types.SimpleNamespace(NotFoundError=type('NotFoundError', (Exception,), {}),
                               InvalidArgumentError=type('InvalidArgumentError', (Exception,), {}))
__version__ = '2.x-numpy-facade'


# --------------------------------------------------------------------------- sparse
class SparseTensor:
    def __init__(self, indices, values, dense_shape):
        self.indices = np.asarray(_raw(indices)).astype(np.int64).reshape(-1, 2).view(Tensor)
        self.values = _raw(convert_to_tensor(values)).view(Tensor)
        self.shape = tuple(int(s) for s in np.asarray(_raw(dense_shape)).reshape(-1))
        self.dense_shape = np.asarray(self.shape, dtype=np.int64).view(Tensor)
        self.dtype = self.values.dtype


def _sp_reorder(s):
    idx = _raw(s.indices)
    order = np.lexsort((idx[:, 1], idx[:, 0]))
    return SparseTensor(idx[order], _raw(s.values)[order], s.dense_shape)


def _to_scipy(s):
    idx = _raw(s.indices)
    m = sp.csr_matrix((_raw(s.values), (idx[:, 0], idx[:, 1])), shape=s.shape)
    m.sort_indices()
    return m


def _from_scipy(m):
    c = m.tocoo()
    order = np.lexsort((c.col, c.row))
    return SparseTensor(np.stack([c.row[order], c.col[order]], axis=1), c.data[order], m.shape)


def _sp_add(a, b, threshold=0):
    """tf.sparse.add: union of the two patterns, fp32 sums (structural zeros kept)."""
    A, B = _to_scipy(a), _to_scipy(b)
    # keep explicit zeros: add patterns via ones-matrices to fix the union pattern first
    pat = (abs_pattern(A) + abs_pattern(B)).tocsr()
    pat.sort_indices()
    out = pat.copy()
    out.data = np.zeros_like(out.data, dtype=A.dtype)
    for M in (A, B):
        Mc = M.tocoo()
        pos = _positions(pat, Mc.row, Mc.col)
        np.add.at(out.data, pos, Mc.data)
    return _from_scipy_keep(out)


def abs_pattern(M):
    P = M.copy()
    P.data = np.ones_like(P.data)
    return P


def _positions(csr, rows, cols):
    pos = np.empty(rows.size, dtype=np.int64)
    for n, (r, c) in enumerate(zip(rows, cols)):
        lo, hi = csr.indptr[r], csr.indptr[r + 1]
        pos[n] = lo + np.searchsorted(csr.indices[lo:hi], c)
    return pos


def _from_scipy_keep(m):
    rows = np.repeat(np.arange(m.shape[0]), np.diff(m.indptr))
    return SparseTensor(np.stack([rows, m.indices], axis=1), m.data, m.shape)


def _sp_dense_matmul(s, dense):
    return _wrap((_to_scipy(s) @ _raw(dense)).astype(_raw(s.values).dtype))


sparse = types.SimpleNamespace(
    SparseTensor=SparseTensor, reorder=_sp_reorder, add=_sp_add,
    map_values=lambda op, s, *args: SparseTensor(s.indices, op(s.values, *args), s.dense_shape),
    sparse_dense_matmul=_sp_dense_matmul,
    to_dense=lambda s: _wrap(_to_scipy(s).toarray()))


class _CsrHandle:
    """What CSRSparseMatrix(...)._matrix is in TF: an opaque variant tensor."""

    def __init__(self, m):
        self.m = m


class CSRSparseMatrix:
    def __init__(self, value, indices=None, name=None):
        if isinstance(value, SparseTensor):
            st = _sp_reorder(value)
            idx = _raw(st.indices)
            m = sp.csr_matrix((_raw(st.values), (idx[:, 0], idx[:, 1])), shape=st.shape)
            # csr_matrix((data, ij)) SUMS duplicates; TF requires unique indices, so nothing is lost
        elif isinstance(value, _CsrHandle):
            m = value.m
        else:
            m = sp.csr_matrix(_raw(value))
        m.sort_indices()
        self._matrix = _CsrHandle(m)
        self.shape = m.shape
        self.dtype = m.dtype

    def to_sparse_tensor(self, type=None):  # noqa: A002
        return _from_scipy_keep(self._matrix.m)

    def to_dense(self):
        return _wrap(self._matrix.m.toarray())


def _sparse_matrix_matmul(a, b, **kw):
    m = a.m if isinstance(a, _CsrHandle) else a._matrix.m
    x = _raw(b)
    return _wrap((m @ x).astype(m.dtype))


def _sparse_matrix_add(a, b, alpha, beta, **kw):
    raise errors.NotFoundError('SparseMatrixAdd has no CPU kernel (the reference falls back to COO add)')


raw_ops = types.SimpleNamespace(SparseMatrixMatMul=_sparse_matrix_matmul, SparseMatrixAdd=_sparse_matrix_add)


# --------------------------------------------------------------------------- install
def install():
    """Register the facade as ``tensorflow`` and expose the reference's ``gpuSolve``."""
    me = sys.modules[__name__]
    sys.modules['tensorflow'] = me
    chain = ['tensorflow.python', 'tensorflow.python.ops', 'tensorflow.python.ops.linalg',
             'tensorflow.python.ops.linalg.sparse', 'tensorflow.python.ops.linalg.sparse.sparse_csr_matrix_ops']
    for name in chain:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules[chain[-1]].CSRSparseMatrix = CSRSparseMatrix
    for name in ('nibabel', 'imageio', 'imageio.v2', 'vedo'):       # optional readers' deps
        if name not in sys.modules:
            stub = types.ModuleType(name)
            stub.__getattr__ = lambda attr: type(attr, (), {})      # annotations only; never called
            sys.modules[name] = stub
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    if 'gpuSolve' not in sys.modules:
        pkg = types.ModuleType('gpuSolve')
        pkg.__path__ = [REF_ROOT + '/gpuSolve']
        sys.modules['gpuSolve'] = pkg
        import importlib
        ver = importlib.import_module('gpuSolve._version')
        pkg.__version__ = ver.__version__
    return me
