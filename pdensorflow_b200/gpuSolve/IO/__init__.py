"""IO sub-package of the gpuSolve mirror: the result writers that sit ON the stepping loop of every reference
example and the IGB reader that round-trips them.  Mesh / image readers and the
vedo plotter are outside the hot path and are not mirrored."""
