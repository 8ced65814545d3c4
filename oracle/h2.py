"""ORACLE (test infrastructure; never imported by the product path).

numpy restatement of the H2 far-field kernel blocks of the reference,
assembleFarFieldInteractions (nl/PyNucleus_nl/clusterMethodCy.pyx:2153-2238):
kernelInterpolant[i, j] = -2 gamma(xi_i, xi_j) at tensor Chebyshev nodes of the two cluster boxes, nodes
enumerated with the last dimension fastest (productIterator, :2120-2151).
Pinned by tests/golden/h2_*.npz (far_blocks, produced by the reference's getH2).
"""
from itertools import product

import numpy as np

from .tables import fractional_scaling


def chebyshev_points(box, m):
    """box: (dim, 2).  Returns (m^dim, dim) points."""
    dim = box.shape[0]
    eta = np.cos((2.0*np.arange(m, 0, -1)-1.0)/(2.0*m)*np.pi)
    pts = np.empty((m**dim, dim))
    for k, idx in enumerate(product(range(m), repeat=dim)):
        for j in range(dim):
            eta_p = eta[idx[j]]+1.0
            pts[k, j] = (box[j, 1]-box[j, 0])*0.5*eta_p+box[j, 0]
    return pts


def farfield_block(dim, s, box1, box2, m1, m2):
    C = fractional_scaling(dim, s)
    x = chebyshev_points(np.asarray(box1), int(m1))
    y = chebyshev_points(np.asarray(box2), int(m2))
    d2 = np.zeros((x.shape[0], y.shape[0]))
    for j in range(dim):
        d2 += (x[:, None, j]-y[None, :, j])*(x[:, None, j]-y[None, :, j])
    return -2.0*(C*np.power(d2, -0.5*dim-s))
