"""ionic submodule: IonicModel, Fenton4v, ModifiedMS2v, MitchellSchaeffer2v,
CourtemancheRamirezNattel, TenTusscherPanfilov (mirror of gpuSolve/ionic/__init__.py)."""
from .._version import __version__, version  # noqa: F401
from .ionicmodel import IonicModel  # noqa: F401
from .fenton4v import Fenton4v  # noqa: F401
from .mms2v import ModifiedMS2v  # noqa: F401
from .ms2v import MitchellSchaeffer2v  # noqa: F401
from .courtemanche_ramirez_nattel import CourtemancheRamirezNattel  # noqa: F401
from .ten_tusscher_panfilov import TenTusscherPanfilov  # noqa: F401
