"""Shared helpers of the gpuSolve mirror: device selection and tensor conversion."""
import numpy as np
import torch

from .. import _lib as L  # noqa: F401

_DEVICE = None


def device():
    """CUDA device the mirror allocates on (current device unless set_device was called)."""
    global _DEVICE
    if _DEVICE is None:
        if not torch.cuda.is_available():
            raise L.PdxError('no CUDA device: pdensorflow_b200 has no CPU path')
        _DEVICE = torch.device('cuda', torch.cuda.current_device())
    return _DEVICE


def set_device(dev):
    global _DEVICE
    _DEVICE = torch.device(dev)


def as_tensor(x, like=None):
    """fp32 contiguous CUDA tensor from a tensor / ndarray / scalar."""
    if isinstance(x, torch.Tensor):
        t = x
        if not t.is_cuda:
            t = t.to(like.device if like is not None else device())
        if t.dtype != torch.float32:
            t = t.to(torch.float32)
        return t.contiguous()
    a = np.ascontiguousarray(x, dtype=np.float32)
    return torch.from_numpy(a).to(like.device if like is not None else device())


def as_float(x):
    """Python float holding the fp32 value of a scalar / 0-d tensor / ndarray."""
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return float(np.float32(np.asarray(x).reshape(-1)[0] if np.ndim(x) else x))
