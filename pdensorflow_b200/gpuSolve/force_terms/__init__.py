"""force_terms submodule: Stimulus (mirror of gpuSolve/force_terms/__init__.py)."""
from .._version import __version__, version  # noqa: F401
from .stimulus import Stimulus  # noqa: F401
