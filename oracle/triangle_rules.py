"""ORACLE (test infrastructure, never imported by the product path).

Regular (non-touching) triangle quadrature family used in place of the
un-vendored third-party tables.

The reference takes its 2D regular-pair rules from ``modepy``'s
``XiaoGimbutasSimplexQuadrature`` (fem/PyNucleus_fem/quadrature.pyx:13-14,
521-545).  modepy is not vendored in /root/reference and is not installed
here, so those tables are unavailable.  The oracle, the
stub-built reference used to generate goldens (oracle/refbuild) and the CUDA
product all use the SAME constructible family of equal polynomial exactness,
defined here as ``family(order)``:

  order 1           centroid rule (1 node)
  order 2           3 interior nodes (1/6, 1/6, 2/3), closed form
  order 4           6-node symmetric rule, closed form
  order 5           7-node Radon rule, closed form
  any other order   conical Gauss-Jacobi product == the reference's own
                    ``simplexDuffyTransformation(order, 2, 2)``
                    (fem/PyNucleus_fem/quadrature.pyx:481-518)

All rules have positive weights, nodes strictly inside the triangle, weights
summing to 1 (the reference scales modepy's weights by 0.5 so they sum to 1,
quadrature.pyx:535-538) and are exact for polynomials of total degree
``order`` (checked by ``check_exactness``).

2D regular-panel parity against the *true* reference tables is therefore
UNPINNED; everything else on the path does not depend on this choice.
"""
from math import factorial, sqrt

import numpy as np


def gauss_jacobi_unit(order, alpha, beta):
    """1D rule on [0,1] for weight x^alpha (1-x)^beta, exact to degree `order`.

    Follows the node-count rule and the scipy call of the reference's
    ``GaussJacobi`` (quadrature.pyx:451-466): k=(order+1)//2, one more if
    2k-1 != order, nodes from ``js_roots(k, beta+alpha+1, alpha+1)``.
    """
    from scipy.special import roots_sh_jacobi
    k = (order + 1) // 2
    if 2 * k - 1 != order:
        k += 1
    a1 = alpha + 1
    b1 = beta + a1
    x, w = roots_sh_jacobi(k, b1, a1)
    return np.asarray(x, dtype=np.float64), np.asarray(w, dtype=np.float64)


def conical_rule(order):
    """quadrature.pyx:481-518 for dim = manifold_dim = 2.

    Axis d uses (order + 1 - d) with weight (1-x)^(1-d); axis 0 is the slow
    index of the tensor product (itertools.product, quadrature.pyx:472-476).
    Barycentric nodes: l2 = x1 (1-x0), l1 = x0, l0 = 1 - l1 - l2.
    Weights are doubled so they sum to 1.
    """
    x0, w0 = gauss_jacobi_unit(order + 1, 0, 1)
    x1, w1 = gauss_jacobi_unit(order, 0, 0)
    n = x0.shape[0] * x1.shape[0]
    bary = np.empty((3, n))
    w = np.empty(n)
    k = 0
    for i in range(x0.shape[0]):
        for j in range(x1.shape[0]):
            l2 = x1[j]
            l2 *= (1. - x0[i])
            l1 = x0[i]
            l0 = 1.
            l0 -= l1
            l0 -= l2
            bary[0, k] = l0
            bary[1, k] = l1
            bary[2, k] = l2
            wk = 1.0
            wk *= w0[i]
            wk *= w1[j]
            w[k] = wk * 2.
            k += 1
    return bary, w


def _orbit3(a):
    """Nodes (a,a,1-2a) and permutations."""
    b = 1. - 2. * a
    return [(b, a, a), (a, b, a), (a, a, b)]


def _closed_form(order):
    if order == 1:
        return np.array([[1. / 3.], [1. / 3.], [1. / 3.]]), np.array([1.])
    if order == 2:
        pts = _orbit3(1. / 6.)
        return np.array(pts).T.copy(), np.full(3, 1. / 3.)
    if order == 4:
        r = sqrt(38. - 44. * sqrt(2. / 5.))
        a1 = (8. - sqrt(10.) + r) / 18.
        a2 = (8. - sqrt(10.) - r) / 18.
        q = sqrt(213125. - 53320. * sqrt(10.))
        w1 = (620. + q) / 3720.
        w2 = (620. - q) / 3720.
        pts = _orbit3(a1) + _orbit3(a2)
        w = [w1] * 3 + [w2] * 3
        return np.array(pts).T.copy(), np.array(w)
    if order == 5:
        a1 = (6. - sqrt(15.)) / 21.
        a2 = (6. + sqrt(15.)) / 21.
        w1 = (155. - sqrt(15.)) / 1200.
        w2 = (155. + sqrt(15.)) / 1200.
        pts = [(1. / 3., 1. / 3., 1. / 3.)] + _orbit3(a1) + _orbit3(a2)
        w = [9. / 40.] + [w1] * 3 + [w2] * 3
        return np.array(pts).T.copy(), np.array(w)
    return None


CLOSED_FORM_ORDERS = (1, 2, 4, 5)


def family(order):
    """(bary[3, n], w[n]) of the adopted family for integer `order` >= 1."""
    order = int(order)
    cf = _closed_form(order)
    if cf is not None:
        return np.ascontiguousarray(cf[0]), np.ascontiguousarray(cf[1])
    return conical_rule(order)


def check_exactness(bary, w, degree):
    """max abs error over all monomials l0^a l1^b l2^c with a+b+c <= degree.

    int_T l0^a l1^b l2^c / |T| = 2 a! b! c! / (a+b+c+2)!
    """
    err = 0.
    for a in range(degree + 1):
        for b in range(degree + 1 - a):
            for c in range(degree + 1 - a - b):
                exact = 2. * factorial(a) * factorial(b) * factorial(c) / factorial(a + b + c + 2)
                val = np.sum(w * bary[0] ** a * bary[1] ** b * bary[2] ** c)
                err = max(err, abs(val - exact))
    return err


if __name__ == '__main__':
    for p in range(1, 31):
        b, w = family(p)
        print(p, b.shape[1], check_exactness(b, w, p), w.min() > 0, b.min() > 0)
