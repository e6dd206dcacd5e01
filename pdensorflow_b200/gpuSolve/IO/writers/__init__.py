"""writers: BaseWriter, ResultWriter, IGBWriter (gpuSolve/IO/writers/__init__.py:26-30).  CarpMeshWriter and
VedoPlotter are outside the time-stepping path and are not mirrored."""
from .basewriter import BaseWriter  # noqa: F401
from .resultwriter import ResultWriter  # noqa: F401
from .igbwriter import IGBWriter  # noqa: F401
