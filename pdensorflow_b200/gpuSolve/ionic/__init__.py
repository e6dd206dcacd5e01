"""ionic submodule: IonicModel, Fenton4v, ModifiedMS2v, MitchellSchaeffer2v
(mirror of gpuSolve/ionic/__init__.py; the 20/23-variable Courtemanche and ten Tusscher
models are not built yet - see DESIGN.md "Out of scope this round")."""
from .._version import __version__, version  # noqa: F401
from .ionicmodel import IonicModel  # noqa: F401
from .fenton4v import Fenton4v  # noqa: F401
from .mms2v import ModifiedMS2v  # noqa: F401
from .ms2v import MitchellSchaeffer2v  # noqa: F401
