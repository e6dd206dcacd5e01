"""Drop-in mirror of the reference's ``gpuSolve`` interfaces for the monodomain hot path.

Same import paths, class / function names, cfg keys and error behaviour as
/root/reference/PDEnsorflow/gpuSolve, with ``tf.Tensor`` replaced by CUDA ``torch.Tensor``
(fp32, contiguous) and every TensorFlow op chain replaced by a call into libpdx.so
(hand-written sm_100a kernels, see include/pdx.h).  There is no TensorFlow, no Triton and
no CPU fallback: without the library or a CUDA device the calls raise.

Use ``import pdensorflow_b200.dropin`` to register this package under the name
``gpuSolve`` so unmodified user scripts (``from gpuSolve.diffop3D import laplace_homog``)
pick it up.
"""
from ._version import __version__, version  # noqa: F401
