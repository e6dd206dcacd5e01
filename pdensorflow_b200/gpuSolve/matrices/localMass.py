"""P1 local mass matrices - mirror of gpuSolve/matrices/localMass.py (Edges, Trias, Tetras)."""
import numpy as np

_TEMPLATES = {
    'Edges': (1.0, np.array([[1.0 / 3.0, 1.0 / 6.0], [1.0 / 6.0, 1.0 / 3.0]])),
    'Trias': (2.0, np.full((3, 3), 1.0 / 24.0) + np.eye(3) * (1.0 / 12.0 - 1.0 / 24.0)),
    'Tetras': (6.0, np.full((4, 4), 1.0 / 120.0) + np.eye(4) * (1.0 / 60.0 - 1.0 / 120.0)),
}


def localMass(elemtype, elemData, props=None):
    """Local mass of one element (elemData['meas'] = length / area / volume)."""
    if elemtype not in _TEMPLATES:
        raise NotImplementedError('localMass: element type %s' % elemtype)
    scale, T = _TEMPLATES[elemtype]
    return scale * float(np.squeeze(elemData['meas'])) * T


def batch_local_mass(elemtype, measures):
    """(nE,1) measures -> (nE,nV,nV)."""
    scale, T = _TEMPLATES[elemtype]
    return scale * measures[:, :, None] * T[None]
