"""Regular triangle quadrature family of the product.

The reference obtains its 2D regular-pair rules from the un-vendored modepy
package (fem/PyNucleus_fem/quadrature.pyx:13-14, 521-545); those tables are not
available offline.  This family has the same polynomial exactness per order:
closed-form symmetric rules for orders 1, 2, 4, 5 and the reference's own
conical Gauss-Jacobi product (quadrature.pyx:481-518) otherwise.  Weights sum
to one, all nodes are interior, all weights positive.
"""
from math import sqrt

import numpy as np
from scipy.special import roots_sh_jacobi


def gauss_jacobi_01(order, alpha, beta):
    """nodes/weights on [0,1] for x^alpha (1-x)^beta, exact to `order`
    (node count rule of quadrature.pyx:458-466)."""
    k = (int(order)+1)//2
    if 2*k-1 != int(order):
        k += 1
    x, w = roots_sh_jacobi(k, beta+alpha+1, alpha+1)
    return np.asarray(x, dtype=np.float64), np.asarray(w, dtype=np.float64)


def _sym3(a):
    b = 1.-2.*a
    return np.array([[b, a, a], [a, b, a], [a, a, b]]).T


def triangle_rule(order):
    order = int(order)
    if order == 1:
        return np.full((3, 1), 1./3.), np.ones(1)
    if order == 2:
        return np.ascontiguousarray(_sym3(1./6.)), np.full(3, 1./3.)
    if order == 4:
        r = sqrt(38.-44.*sqrt(2./5.))
        q = sqrt(213125.-53320.*sqrt(10.))
        bary = np.hstack((_sym3((8.-sqrt(10.)+r)/18.), _sym3((8.-sqrt(10.)-r)/18.)))
        w = np.repeat([(620.+q)/3720., (620.-q)/3720.], 3)
        return np.ascontiguousarray(bary), w
    if order == 5:
        bary = np.hstack((np.full((3, 1), 1./3.), _sym3((6.-sqrt(15.))/21.), _sym3((6.+sqrt(15.))/21.)))
        w = np.concatenate(([9./40.], np.repeat([(155.-sqrt(15.))/1200., (155.+sqrt(15.))/1200.], 3)))
        return np.ascontiguousarray(bary), w
    # conical product: axis 0 of order+1 with weight (1-x), axis 1 of `order`
    x0, w0 = gauss_jacobi_01(order+1, 0, 1)
    x1, w1 = gauss_jacobi_01(order, 0, 0)
    X0, X1 = np.meshgrid(x0, x1, indexing='ij')
    W0, W1 = np.meshgrid(w0, w1, indexing='ij')
    l2 = (X1*(1.-X0)).ravel()
    l1 = X0.ravel()
    l0 = (1.-l1)-l2
    w = ((1.0*W0)*W1).ravel()*2.
    return np.ascontiguousarray(np.vstack((l0, l1, l2))), w
