"""P1 local stiffness matrices - mirror of gpuSolve/matrices/localStiffness.py:
meas * G^T Sigma G with G = (contravariant basis) x (reference gradients)."""
import numpy as np

_LGRAD = {
    'Edges': np.array([[-1.0, 1.0]]),
    'Trias': np.array([[-1.0, 1.0, 0.0], [-1.0, 0.0, 1.0]]),
    'Tetras': np.array([[-1.0, 1.0, 0.0, 0.0], [-1.0, 0.0, 1.0, 0.0], [-1.0, 0.0, 0.0, 1.0]]),
}


def localStiffness(elemtype, elemData, Sigma=None):
    if elemtype not in _LGRAD:
        raise NotImplementedError('localStiffness: element type %s' % elemtype)
    Sigma = np.identity(3) if Sigma is None else Sigma
    Lg = _LGRAD[elemtype]
    C = np.stack([np.asarray(elemData[k]) for k in ('v1', 'v2', 'v3')[:Lg.shape[0]]], axis=1)
    G = C @ Lg
    return float(np.squeeze(elemData['meas'])) * (G.T @ (Sigma @ G))


def batch_local_stiffness(elemtype, basis, measures, Sigma):
    Lg = _LGRAD[elemtype]
    C = np.stack([basis[k] for k in ('v1', 'v2', 'v3')[:Lg.shape[0]]], axis=2)
    G = np.einsum('eij,jk->eik', C, Lg)
    return measures[:, :, None] * np.einsum('eji,ejk->eik', G, np.einsum('eij,ejk->eik', Sigma, G))
