"""matrices submodule: local P1 matrices and global assembly (mirror of gpuSolve/matrices)."""
from .._version import __version__, version  # noqa: F401
from .localMass import localMass  # noqa: F401
from .localStiffness import localStiffness  # noqa: F401
from .globalMatrices import (CSRMatrix, assemble_mass_matrix, assemble_matrices_dict,  # noqa: F401
                             assemble_stiffness_matrix, compute_coo_pattern,
                             compute_reverse_cuthill_mckee_indexing, csr_axpby, csr_to_sparse_tensor)
