"""entities submodule: Triangulation, MaterialProperties, synthetic mesh generators.
(Domain3D / image ingestion of the reference are out of scope: SURVEY.md section 8f row N4.)"""
from .._version import __version__, version  # noqa: F401
from .triangulation import Triangulation  # noqa: F401
from .materialproperties import MaterialProperties  # noqa: F401
