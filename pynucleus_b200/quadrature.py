"""Host-side quadrature tables of the nonlocal assembly path (built once per
problem with scipy, uploaded through pnb_rules_t).

Mirrors, in vectorised form, the rule constructors of the reference:
  GaussJacobi                               fem/PyNucleus_fem/quadrature.pyx:451-478
  singularityCancelationQuadRule1D[_boundary]   nl/PyNucleus_nl/fractionalLaplacian1D.pyx:35-179
  singularityCancelationQuadRule2D[_boundary]   nl/PyNucleus_nl/fractionalLaplacian2D.pyx:36-563
  simplexDuffyTransformation / simplexXiaoGimbutas   quadrature.pyx:481-545
Tables hold barycentric coordinates (x point rows first, then y point rows);
P1 shape functions are the barycentric coordinates, so PSI/PHI tables
(fractionalLaplacian2D.pyx:644-813) are formed on the device.
"""
from functools import lru_cache
from math import ceil, log

import numpy as np

from .triangle_rules import gauss_jacobi_01, triangle_rule

COMMON_VERTEX, COMMON_EDGE, COMMON_FACE = -1, -2, -3


def tensor_gauss_jacobi(specs):
    """eta[d, n], w[n]; first axis slowest (itertools.product order)."""
    xs, ws = zip(*[gauss_jacobi_01(o, a, b) for (o, a, b) in specs])
    X = np.meshgrid(*xs, indexing='ij')
    W = np.meshgrid(*ws, indexing='ij')
    w = np.ones(X[0].shape)
    for Wm in W:
        w = w*Wm
    return np.vstack([x.ravel() for x in X]), w.ravel()


def _t(x1, x2):
    return np.vstack((1-x1, x1-x2, x2))


def _s(x):
    return np.vstack((1-x, x))


def _join(parts, weights):
    return np.ascontiguousarray(np.hstack([np.vstack(p) for p in parts])), np.concatenate(weights)


@lru_cache(maxsize=256)
def singular2d(panel, sg, qod, qodV):
    if panel == COMMON_FACE:
        (a, b, c, d), w = tensor_gauss_jacobi(((1, 3+sg, 0), (1, 2+sg, 0), (1, 1+sg, 0), (qod, 0, 0)))
        f = 2.0*w*(a*b*c)**(-sg)
        return _join([(_t(a, a*b*(1-c+c*d)), _t(a*(1-b*c), a*b*(1-c))),
                      (_t(a, a*b), _t(a*(1-b*c*d), a*b*(1-c))),
                      (_t(a, a*b*(1-c)), _t(a*(1-b*c*d), a*b*(1-c*d)))], [f, f, f])
    if panel == COMMON_EDGE:
        (a, b, c, d), w = tensor_gauss_jacobi(((1, 3+sg, 0), (1, 2+sg, 0), (qod, 0, 0), (qod, 0, 0)))
        f0 = w*(a*b)**(-sg)
        parts = [(_t(a*(1-b*c), a*b*(1-c)), _t(a, a*b*d)),
                 (_t(a, a*b*d), _t(a*(1-b*c), a*b*(1-c)))]
        (a, b, c, d), w = tensor_gauss_jacobi(((1, 3+sg, 0), (1, 2+sg, 0), (qod, 1, 0), (qod, 0, 0)))
        f1 = w*(a*b)**(-sg)
        parts += [(_t(a*(1-b*c*d), a*b*c*(1-d)), _t(a, a*b)),
                  (_t(a, a*b), _t(a*(1-b*c*d), a*b*c*(1-d)))]
        return _join(parts, [f0, f0, f1, f1])
    if panel == COMMON_VERTEX:
        (a, b, c, d), w = tensor_gauss_jacobi(((1, 3+sg, 0), (qodV, 0, 0), (qodV, 1, 0), (qodV, 0, 0)))
        f = w*a**(-sg)
        return _join([(_t(a, a*b), _t(a*c, a*c*d)), (_t(a*c, a*c*d), _t(a, a*b))], [f, f])
    raise NotImplementedError(panel)


@lru_cache(maxsize=256)
def singular2d_boundary(panel, sg, qod):
    if panel == COMMON_EDGE:
        (a, b, c), w = tensor_gauss_jacobi(((qod, 1.+sg, 1.), (qod, 0., 0.), (qod, 0., 0.)))
        f = w*a**(-sg)
        return _join([(np.vstack((1-a-(1-a)*c, a+(1-a)*c-a*b, a*b)), np.vstack((1-c*(1-a), c*(1-a)))),
                      (np.vstack((1-a-c+a*c, c-a*c, a)), np.vstack((1-c+a*c+a*b-a, c-a*c-a*b+a))),
                      (np.vstack((1-c+a*c-a*b, c-a*c, a*b)), np.vstack((1-c+a*c-a, c-a*c+a)))], [f, f, f])
    if panel == COMMON_VERTEX:
        (a, b, c), w0 = tensor_gauss_jacobi(((qod, 2.0+sg, 0), (qod, 0, 0), (qod, 0, 0)))
        p0 = (np.vstack((1-a, a*(1-b), a*b)), np.vstack((1-a*c, a*c)))
        f0 = w0*a**(-sg)
        (a, b, c), w1 = tensor_gauss_jacobi(((qod, 2.0+sg, 0), (qod, 1, 0), (qod, 0, 0)))
        p1 = (np.vstack((1-a*b, a*b*(1-c), a*b*c)), np.vstack((1-a, a)))
        return _join([p0, p1], [f0, w1*a**(-sg)])
    raise NotImplementedError(panel)


@lru_cache(maxsize=256)
def singular1d(panel, sg, qod, qor):
    if panel == COMMON_EDGE:
        (a, b), w = tensor_gauss_jacobi(((qor, 1+sg, 0), (qor, 0+sg, 0)))
        return _join([(_s(a*(1-b)), _s(a))], [2.0*w*(a*b)**(-sg)])
    if panel == COMMON_VERTEX:
        (a, b), w = tensor_gauss_jacobi(((qor, 1+sg, 0), (qod, 0, 0)))
        f = w*a**(-sg)
        return _join([(_s(a*b), _s(a)), (_s(a), _s(a*b))], [f, f])
    raise NotImplementedError(panel)


@lru_cache(maxsize=256)
def singular1d_boundary(sg, qod):
    (a, ), w = tensor_gauss_jacobi(((qod, sg, 0), ))
    return np.ascontiguousarray(np.vstack((1-a, a, np.ones_like(a)))), w*a**(-sg)


@lru_cache(maxsize=256)
def regular(order, manifold_dim):
    """rule of integer `order` on a simplex of dimension `manifold_dim`"""
    if manifold_dim == 0:
        return np.ones((1, 1)), np.ones(1)
    if manifold_dim == 1:
        x, w = gauss_jacobi_01(order, 0, 0)
        return np.ascontiguousarray(np.vstack((1.-x, x))), w.copy()
    return triangle_rule(order)


class localMatrixOrders:
    """target order and singular quadrature orders of the interior and boundary
    local matrices (setKernel of fractionalLaplacian{1,2}D[_boundary])."""

    def __init__(self, dim, singularity, bsingularity, hmin, H0, num_dofs, target_order=None, polynomialOrder=1,
                 min_singularity=None, min_bsingularity=None):
        # variable orders: `singularity` is kernel.max_singularity (from s.max), `min_singularity` kernel.min_singularity
        # (from s.min); they only differ in the 1D default target order
        lg = abs(log(hmin/H0))
        min_singularity = singularity if min_singularity is None else min_singularity
        min_bsingularity = bsingularity if min_bsingularity is None else min_bsingularity
        if dim == 2:
            # fractionalLaplacian2D.pyx:600-615, 1210-1220
            to = 0.5 if target_order is None else target_order
            smax = max(-0.5*(singularity+2), 0.)
            self.target_order = to
            self.quad_order_diagonal = int(max(ceil((to+1.+smax)/0.43*lg), 4))
            self.quad_order_diagonalV = int(max(ceil((to+1.+smax)/0.7*lg), 4))
            smaxb = max(0.5*(-bsingularity-1.), 0.)
            self.btarget_order = to
            self.bquad_order_diagonal = int(max(ceil((to+0.5+smaxb)/0.35*lg), 2))
        else:
            # fractionalLaplacian1D.pyx:218-228, 629-639
            smax = max(-0.5*(singularity+1), 0.)
            smin = max(-0.5*(min_singularity+1), 0.)
            to = polynomialOrder+1-smin if target_order is None else target_order
            self.target_order = to
            self.quad_order_diagonal = int(max(ceil(((to+2.)*log(num_dofs*H0)+(2.*smax-1.)*lg)/0.8), 2))
            self.quad_order_diagonalV = self.quad_order_diagonal
            smaxb = max(0.5*(-bsingularity), 0.)
            sminb = max(0.5*(-min_bsingularity), 0.)
            tob = polynomialOrder+1-sminb if target_order is None else target_order
            self.btarget_order = tob
            self.bquad_order_diagonal = int(max(ceil(((tob+1.)*log(num_dofs*H0)+(2.*smaxb-1.)*lg)/0.8), 2))


def singular_tables(dim, singularity, bsingularity, orders, polynomialOrder=1):
    """dict name -> (bary, w) for the five singular tables of pnb_rules_t"""
    out = {}
    # cancellation orders (fractionalLaplacian2D.pyx:591-600, fractionalLaplacian1D.pyx:209-216): the integrand cancels two
    # orders of the singularity within an element, and two across elements for continuous elements (none for P0)
    sg = 2.+singularity
    sga = singularity if polynomialOrder == 0 else 2.+singularity
    if dim == 2:
        out['identical'] = singular2d(COMMON_FACE, sg, orders.quad_order_diagonal, orders.quad_order_diagonalV)
        out['edge'] = singular2d(COMMON_EDGE, sga, orders.quad_order_diagonal, orders.quad_order_diagonalV)
        out['vertex'] = singular2d(COMMON_VERTEX, sga, orders.quad_order_diagonal, orders.quad_order_diagonalV)
        sgb = bsingularity if bsingularity > -2.+1e-3 else 2.+bsingularity   # fractionalLaplacian2D.pyx:1271-1274
        out['bedge'] = singular2d_boundary(COMMON_EDGE, sgb, orders.bquad_order_diagonal)
        out['bvertex'] = singular2d_boundary(COMMON_VERTEX, bsingularity, orders.bquad_order_diagonal)
    else:
        qor = 2*max(polynomialOrder, 1)
        out['identical'] = singular1d(COMMON_EDGE, sg, orders.quad_order_diagonal, qor)
        out['vertex'] = singular1d(COMMON_VERTEX, sga, orders.quad_order_diagonal, qor)
        sgb = bsingularity if bsingularity > -1.+1e-3 else 2.+bsingularity   # fractionalLaplacian1D.pyx:688-691
        out['bvertex'] = singular1d_boundary(sgb, orders.bquad_order_diagonal)
    return out
