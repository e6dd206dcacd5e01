"""entities submodule: Domain3D (voxel domain of the FD examples), Triangulation, MaterialProperties, synthetic mesh
generators.  Image ingestion (PNG mosaic / NIfTI) of the reference is out of scope: SURVEY.md section 8f row N4."""
from .._version import __version__, version  # noqa: F401
from .domain3D import Domain3D  # noqa: F401
from .triangulation import Triangulation  # noqa: F401
from .materialproperties import MaterialProperties  # noqa: F401
