"""ORACLE (test infrastructure; never imported by the product path).

CPU restatement of the reference's quadrature-table construction for the
nonlocal assembly path:

* Gauss-Jacobi tensor rules     fem/PyNucleus_fem/quadrature.pyx:451-478
* 1D singular rules             nl/PyNucleus_nl/fractionalLaplacian1D.pyx:35-179
* 2D singular rules             nl/PyNucleus_nl/fractionalLaplacian2D.pyx:36-563
* diagonal order heuristics     fractionalLaplacian2D.pyx:587-620, 1207-1224,
                                fractionalLaplacian1D.pyx:203-232, 626-642
* regular rules                 quadrature.pyx:481-545 (2D: adopted family,
                                oracle/triangle_rules.py)

Pinned against tables dumped from the reference itself
(tests/golden/*.npz: qrId/qrEdge/qrVertex/bqrEdge/bqrVertex nodes+weights).

A rule is returned as (bary, w): bary[r, n] holds the barycentric coordinates
of the x point (rows 0..dim) followed by those of the y point.
P1 shape functions are the barycentric coordinates themselves
(fem/PyNucleus_fem/DoFMaps.pyx:1854), so no PSI tables are stored.
"""
from itertools import product
from math import ceil, log, sqrt

import numpy as np

from .triangle_rules import family, gauss_jacobi_unit

COMMON_VERTEX = -1
COMMON_EDGE = -2
COMMON_FACE = -3


def gauss_jacobi_tensor(specs):
    """quadrature.pyx:451-478. specs = ((order, alpha, beta), ...).
    Returns eta[len(specs), n], w[n]; axis 0 is the slowest index."""
    one_d = [gauss_jacobi_unit(o, a, b) for (o, a, b) in specs]
    sizes = [x.shape[0] for x, _ in one_d]
    n = int(np.prod(sizes))
    eta = np.empty((len(specs), n))
    w = np.ones(n)
    for k, idx in enumerate(product(*[range(m) for m in sizes])):
        for m, i in enumerate(idx):
            eta[m, k] = one_d[m][0][i]
            w[k] *= one_d[m][1][i]
    return eta, w


def _tri(x1, x2):
    return np.vstack((1-x1, x1-x2, x2))


def _seg(x):
    return np.vstack((1-x, x))


# --------------------------------------------------------------------------
# 2D, element x element                    fractionalLaplacian2D.pyx:36-399
# --------------------------------------------------------------------------

def singular_rule_2d(panel, sg, qod, qodV):
    """`sg` = cancellation order (2) + kernel singularity, as passed by
    getNearQuadRule (fractionalLaplacian2D.pyx:667-671, 698-702, 771-775)."""
    if panel == COMMON_FACE:
        e, w = gauss_jacobi_tensor(((1, 3+sg, 0), (1, 2+sg, 0), (1, 1+sg, 0), (qod, 0, 0)))
        e0, e1, e2, e3 = e
        fac = 2.0*w*(e0*e1*e2)**(-sg)
        parts = [
            (_tri(e0, e0*e1*(1-e2+e2*e3)), _tri(e0*(1-e1*e2), e0*e1*(1-e2))),
            (_tri(e0, e0*e1), _tri(e0*(1-e1*e2*e3), e0*e1*(1-e2))),
            (_tri(e0, e0*e1*(1-e2)), _tri(e0*(1-e1*e2*e3), e0*e1*(1-e2*e3))),
        ]
        weights = [fac, fac, fac]
    elif panel == COMMON_EDGE:
        ea, wa = gauss_jacobi_tensor(((1, 3+sg, 0), (1, 2+sg, 0), (qod, 0, 0), (qod, 0, 0)))
        eb, wb = gauss_jacobi_tensor(((1, 3+sg, 0), (1, 2+sg, 0), (qod, 1, 0), (qod, 0, 0)))
        e0, e1, e2, e3 = ea
        fa = wa*(e0*e1)**(-sg)
        parts = [
            (_tri(e0*(1-e1*e2), e0*e1*(1-e2)), _tri(e0, e0*e1*e3)),
            (_tri(e0, e0*e1*e3), _tri(e0*(1-e1*e2), e0*e1*(1-e2))),
        ]
        e0, e1, e2, e3 = eb
        fb = wb*(e0*e1)**(-sg)
        parts += [
            (_tri(e0*(1-e1*e2*e3), e0*e1*e2*(1-e3)), _tri(e0, e0*e1)),
            (_tri(e0, e0*e1), _tri(e0*(1-e1*e2*e3), e0*e1*e2*(1-e3))),
        ]
        weights = [fa, fa, fb, fb]
    elif panel == COMMON_VERTEX:
        e, w = gauss_jacobi_tensor(((1, 3+sg, 0), (qodV, 0, 0), (qodV, 1, 0), (qodV, 0, 0)))
        e0, e1, e2, e3 = e
        f = w*e0**(-sg)
        parts = [
            (_tri(e0, e0*e1), _tri(e0*e2, e0*e2*e3)),
            (_tri(e0*e2, e0*e2*e3), _tri(e0, e0*e1)),
        ]
        weights = [f, f]
    else:
        raise NotImplementedError(panel)
    bary = np.hstack([np.vstack(p) for p in parts])
    return np.ascontiguousarray(bary), np.concatenate(weights)


# --------------------------------------------------------------------------
# 2D, element x boundary edge             fractionalLaplacian2D.pyx:402-563
# --------------------------------------------------------------------------

def singular_rule_2d_boundary(panel, sg, qod):
    qor = qod   # getNearQuadRule passes quad_order_diagonal twice (:1272-1274, :1297)
    if panel == COMMON_EDGE:
        e, w = gauss_jacobi_tensor(((qor, 1.+sg, 1.), (qod, 0., 0.), (qod, 0., 0.)))
        e0, e1, e2 = e
        f = w*e0**(-sg)
        parts = [
            (np.vstack((1-e0-(1-e0)*e2, e0+(1-e0)*e2-e0*e1, e0*e1)),
             np.vstack((1-e2*(1-e0), e2*(1-e0)))),
            (np.vstack((1-e0-e2+e0*e2, e2-e0*e2, e0)),
             np.vstack((1-e2+e0*e2+e0*e1-e0, e2-e0*e2-e0*e1+e0))),
            (np.vstack((1-e2+e0*e2-e0*e1, e2-e0*e2, e0*e1)),
             np.vstack((1-e2+e0*e2-e0, e2-e0*e2+e0))),
        ]
        weights = [f, f, f]
    elif panel == COMMON_VERTEX:
        ea, wa = gauss_jacobi_tensor(((qor, 2.0+sg, 0), (qod, 0, 0), (qod, 0, 0)))
        eb, wb = gauss_jacobi_tensor(((qor, 2.0+sg, 0), (qod, 1, 0), (qod, 0, 0)))
        e0, e1, e2 = ea
        parts = [(np.vstack((1-e0, e0*(1-e1), e0*e1)), np.vstack((1-e0*e2, e0*e2)))]
        weights = [wa*e0**(-sg)]
        e0, e1, e2 = eb
        parts += [(np.vstack((1-e0*e1, e0*e1*(1-e2), e0*e1*e2)), np.vstack((1-e0, e0)))]
        weights += [wb*e0**(-sg)]
    else:
        raise NotImplementedError(panel)
    bary = np.hstack([np.vstack(p) for p in parts])
    return np.ascontiguousarray(bary), np.concatenate(weights)


# --------------------------------------------------------------------------
# 1D                                        fractionalLaplacian1D.pyx:35-179
# --------------------------------------------------------------------------

def singular_rule_1d(panel, sg, qod, qor):
    if panel == COMMON_EDGE:      # identical cells in 1D
        e, w = gauss_jacobi_tensor(((qor, 1+sg, 0), (qor, 0+sg, 0)))
        e0, e1 = e
        parts = [(_seg(e0*(1-e1)), _seg(e0))]
        weights = [2.0*w*(e0*e1)**(-sg)]
    elif panel == COMMON_VERTEX:
        e, w = gauss_jacobi_tensor(((qor, 1+sg, 0), (qod, 0, 0)))
        e0, e1 = e
        f = w*e0**(-sg)
        parts = [(_seg(e0*e1), _seg(e0)), (_seg(e0), _seg(e0*e1))]
        weights = [f, f]
    else:
        raise NotImplementedError(panel)
    bary = np.hstack([np.vstack(p) for p in parts])
    return np.ascontiguousarray(bary), np.concatenate(weights)


def singular_rule_1d_boundary(sg, qod):
    e, w = gauss_jacobi_tensor(((qod, sg, 0), ))
    eta = e[0]
    bary = np.vstack((1-eta, eta, np.ones_like(eta)))
    return np.ascontiguousarray(bary), w*eta**(-sg)


# --------------------------------------------------------------------------
# regular rules                                   quadrature.pyx:481-545
# --------------------------------------------------------------------------

def regular_rule(order, manifold_dim):
    """simplexXiaoGimbutas(order, dim, manifold_dim): Duffy/Gauss-Jacobi for
    points and segments, the adopted triangle family for triangles."""
    if manifold_dim == 0:
        return np.ones((1, 1)), np.ones(1)
    if manifold_dim == 1:
        x, w = gauss_jacobi_unit(order, 0, 0)
        bary = np.empty((2, x.shape[0]))
        bary[1] = x
        bary[0] = 1.
        bary[0] -= bary[1]
        return bary, w.copy()
    if manifold_dim == 2:
        return family(order)
    raise NotImplementedError()


# --------------------------------------------------------------------------
# order heuristics
# --------------------------------------------------------------------------

def diag_orders(dim, singularity, boundary_singularity, hmin, H0, num_dofs, target_order=None, poly_order=1,
                min_singularity=None, min_boundary_singularity=None):
    """Returns dict with target orders and the quadrature orders of the
    singular rules for the interior and the boundary local matrices."""
    out = {}
    lg = abs(log(hmin/H0))
    # variable orders (fractionalLaplacian1D.pyx:218-222, 629-633): smax from kernel.max_singularity, smin from
    # kernel.min_singularity
    if min_singularity is None:
        min_singularity = singularity
    if min_boundary_singularity is None:
        min_boundary_singularity = boundary_singularity
    if dim == 2:
        to = 0.5 if target_order is None else target_order
        smax = max(-0.5*(singularity+2), 0.)
        out['target_order'] = to
        out['qod'] = int(max(ceil((to+1.+smax)/0.43*lg), 4))
        out['qodV'] = int(max(ceil((to+1.+smax)/0.7*lg), 4))
        smaxb = max(0.5*(-boundary_singularity-1.), 0.)
        out['b_target_order'] = to
        out['b_qod'] = int(max(ceil((to+0.5+smaxb)/0.35*lg), 2))
    elif dim == 1:
        smax = max(-0.5*(singularity+1), 0.)
        smin = max(-0.5*(min_singularity+1), 0.)
        to = poly_order+1-smin if target_order is None else target_order
        out['target_order'] = to
        out['qod'] = int(max(ceil(((to+2.)*log(num_dofs*H0)+(2.*smax-1.)*lg)/0.8), 2))
        smaxb = max(0.5*(-boundary_singularity), 0.)
        sminb = max(0.5*(-min_boundary_singularity), 0.)
        tob = poly_order+1-sminb if target_order is None else target_order
        out['b_target_order'] = tob
        out['b_qod'] = int(max(ceil(((tob+1.)*log(num_dofs*H0)+(2.*smaxb-1.)*lg)/0.8), 2))
    else:
        raise NotImplementedError()
    return out


def near_rules(dim, singularity, boundary_singularity, orders, poly_order=1):
    """All singular tables for one (constant) kernel singularity.

    interior: getNearQuadRule  (2D :644-813, 1D :255-339)
    boundary: getNearQuadRule  (2D :1255-1314, 1D :671-709)
    Returns {('interior'|'boundary', panel): (bary, w)}."""
    rules = {}
    # cancellation across elements: two orders for continuous elements, none for P0 (fractionalLaplacian2D.pyx:591-600,
    # fractionalLaplacian1D.pyx:209-216)
    sga = singularity if poly_order == 0 else 2.+singularity
    if dim == 2:
        sg = 2.+singularity
        for p in (COMMON_FACE, COMMON_EDGE, COMMON_VERTEX):
            rules[('interior', p)] = singular_rule_2d(p, sg if p == COMMON_FACE else sga, orders['qod'], orders['qodV'])
        if boundary_singularity > -2.+1e-3:
            sgb = boundary_singularity
        else:
            sgb = 2.+boundary_singularity
        rules[('boundary', COMMON_EDGE)] = singular_rule_2d_boundary(COMMON_EDGE, sgb, orders['b_qod'])
        rules[('boundary', COMMON_VERTEX)] = singular_rule_2d_boundary(COMMON_VERTEX, boundary_singularity, orders['b_qod'])
    else:
        sg = 2.+singularity
        qor = 2*max(poly_order, 1)
        rules[('interior', COMMON_EDGE)] = singular_rule_1d(COMMON_EDGE, sg, orders['qod'], qor)
        rules[('interior', COMMON_VERTEX)] = singular_rule_1d(COMMON_VERTEX, sga, orders['qod'], qor)
        if boundary_singularity > -1.+1e-3:
            sgb = boundary_singularity
        else:
            sgb = 2.+boundary_singularity
        rules[('boundary', COMMON_VERTEX)] = singular_rule_1d_boundary(sgb, orders['b_qod'])
    return rules


def fractional_scaling(dim, s, horizon=np.inf, tempered=0.):
    """kernelNormalization.pyx:70-89"""
    from scipy.special import gamma
    from math import pi
    if horizon < np.inf:
        return (2.-2*s) * pow(horizon, 2*s-2.) * dim * gamma(0.5*dim)/pow(pi, 0.5*dim) * 0.5
    if tempered != 0. and s != 0.5:
        return gamma(0.5*dim) / abs(gamma(-2*s))/pow(pi, 0.5*dim) * 0.5 * 0.5
    return 2.0**(2.0*s) * s * gamma(s+0.5*dim)/pow(pi, 0.5*dim)/gamma(1.0-s) * 0.5


H0_FACTOR = 1./sqrt(8.)    # nonlocalOperator_{SCALAR}.pxi:435
