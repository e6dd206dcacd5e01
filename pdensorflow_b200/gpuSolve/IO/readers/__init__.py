"""readers: IGBReader (gpuSolve/IO/readers/__init__.py:24).  ImageData / CarpMeshReader are set-up-time
readers outside the time-stepping path and are not mirrored."""
from .igbreader import IGBReader  # noqa: F401
