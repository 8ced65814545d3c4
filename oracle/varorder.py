"""ORACLE -- test infrastructure only (never imported by the product).

CPU restatement (numpy, loops over cell pairs: small cases only) of the reference's dense assembly for fractional orders
that VARY INSIDE A CELL: s(x, y) = sFun(x) (singleVariableUnsymmetricFractionalOrder, fractionalOrders.pyx:153-183; the
driver's `twoDomainNonSym`), kernel.piecewise == False.  What the reference does on that path:

  * the unsymmetric local matrices fractionalLaplacian{1,2}D_nonsym (fractionalLaplacian1D.pyx:410-603,
    fractionalLaplacian2D.pyx:894-1184; distant pairs eval_distant_nonsym, nonlocalOperator_{SCALAR}.pxi:849-911):
        a(I, J) = vol * sum_q w_q [gamma(x_q, y_q) phi_I(x_q) - gamma(y_q, x_q) phi_I(y_q)] [phi_J(x_q) - phi_J(y_q)]
    for BOTH orientations of every cell pair (nonlocalAssembly_{SCALAR}.pxi:1412-1428);
  * order, scaling constant and kernel re-evaluated at every quadrature node (updateAndEvalFractional,
    kernelsCy.pyx:596-622; variableFractionalLaplacianScaling.evalPtr, kernelNormalization.pyx:421-440):
        gamma(x, y) = C(s(x)) |x-y|^(-d-2 s(x)),   C(s) = 2^(2s) s Gamma(s+d/2) / (pi^(d/2) Gamma(1-s)) / 2;
  * per cell pair the singularity -d - 2 max(s) over the two centres and all vertices of the two cells
    (evalParamsOnSimplices, kernelsCy.pyx:1826-1850) selects the regular order (getQuadOrder) and a singular rule of its
    own, cached per value (getNearQuadRule, fractionalLaplacian1D.pyx:452-547);
  * surface terms with the boundary kernel gamma_b(x, y) = C(s(x)) / s(x) |x-y|^(1-d-2 s(x)) (kernels.py:151-160) and
    the symmetric boundary local matrices.

Pinned against fixtures produced by the reference (oracle/refbuild/make_golden_varorder.py, tests/golden/varorder_*.npz)
in tests/test_oracle_golden.py.
"""
from math import pi, log, ceil, sqrt

import numpy as np
from scipy.special import gamma

from . import tables


class smoothStep:
    """fractionalOrders.pyx:389-416"""

    def __init__(self, sl, sr, r=0.1, interface=0.):
        self.sl, self.sr, self.r, self.slope, self.interface = sl, sr, r, 0.5/r, interface
        self.min, self.max = min(sl, sr), max(sl, sr)

    def __call__(self, x):
        x0 = np.asarray(x, dtype=np.float64)[..., 0]
        t = (x0-self.interface)*self.slope+0.5
        v = self.sl+(self.sr-self.sl)*(3.0*t**2-2.0*t**3)
        return np.where(x0 < self.interface-self.r, self.sl, np.where(x0 > self.interface+self.r, self.sr, v))


class linearStep(smoothStep):
    """fractionalOrders.pyx:447-470"""

    def __init__(self, sl, sr, r=0.1, interface=0.):
        smoothStep.__init__(self, sl, sr, r, interface)
        self.slope = 0.5*(sr-sl)/r

    def __call__(self, x):
        x0 = np.asarray(x, dtype=np.float64)[..., 0]
        v = self.sl+self.slope*(x0-self.interface+self.r)
        return np.where(x0 < self.interface-self.r, self.sl, np.where(x0 > self.interface+self.r, self.sr, v))


class feOrder:
    """feFractionalOrder (fractionalOrders.pyx:660-668) on the mesh of the operator: P1 function given by its vertex values;
    at a point of cell c with barycentric coordinates lam the order is sum_k lam_k u[vertex k of c]"""

    def __init__(self, vertex_values, smin=None, smax=None):
        self.vertex_values = np.asarray(vertex_values, dtype=np.float64)
        self.min = float(self.vertex_values.min() if smin is None else smin)
        self.max = float(self.vertex_values.max() if smax is None else smax)


def scaling(dim, s):
    """variableFractionalLaplacianScaling, normalized, infinite horizon (kernelNormalization.pyx:438-439)"""
    return 2.0**(2.0*s)*s*gamma(s+0.5*dim)*pi**(-0.5*dim)/gamma(1.0-s)*0.5


def kernel_value(dim, sFun, x, y, boundary=False, s=None):
    """gamma(x, y) (fracKernelInfinite*, kernelsCy.pyx:159-183; boundary form with phi = 1/s); s: the order at x if known"""
    if s is None:
        s = sFun(x)
    d2 = ((x-y)**2).sum(axis=-1)
    if boundary:
        return scaling(dim, s)/s*d2**(0.5*(1-dim)-s)
    return scaling(dim, s)*d2**(-0.5*dim-s)


def _proto(v1, v2, identical):
    """getProtoPanelType (nonlocalOperator_{SCALAR}.pxi:280-378): shared vertices, first-match permutations"""
    n1, n2 = len(v1), len(v2)
    if identical:
        return -n1, list(range(n1)), list(range(n2))
    p1, p2 = [], []
    for a in range(n1):
        for b in range(n2):
            if b in p2:
                continue
            if v1[a] == v2[b]:
                p1.append(a)
                p2.append(b)
                break
    common = len(p1)
    p1 += [a for a in range(n1) if a not in p1]
    p2 += [b for b in range(n2) if b not in p2]
    return -common, p1, p2


def _h_simplex(X):
    """get_h_simplex / get_h_surface_simplex: the longest edge"""
    n = X.shape[0]
    if n == 1:
        return 1.
    return max(sqrt(((X[i]-X[j])**2).sum()) for i in range(n) for j in range(i+1, n))


def shape_functions(dim, dpe, lam):
    """local shape functions at barycentric coordinates lam (nvc x nq) in the reference's local order (DoFMaps.pyx:1776-2005):
    P0 (one dof per cell), P1 (the vertices), P2 (vertices, then the edges (0,1), (1,2), (0,2); 1D: vertices, then the cell)"""
    nvc = dim+1
    if dpe == 1:
        return np.ones((1, lam.shape[1]))
    if dpe == nvc:
        return lam
    assert dpe == (3 if dim == 1 else 6)
    out = [lam[k]*(2.*lam[k]-1.) for k in range(nvc)]
    out.append(4.*lam[0]*lam[1])
    if dim == 2:
        out += [4.*lam[1]*lam[2], 4.*lam[0]*lam[2]]
    return np.vstack(out)


def dense(vertices, cells, dofs, num_dofs, sFun, bfacets, zero_exterior=True, target_order=None, hmin=None, diam=None):
    vertices = np.asarray(vertices, dtype=np.float64)
    cells = np.asarray(cells)
    dofs = np.asarray(dofs)
    dim = vertices.shape[1]
    nc, nvc = cells.shape
    dpe = dofs.shape[1]
    poly_order = {1: 0, nvc: 1}.get(dpe, 2)
    T = vertices[cells]                      # nc x nvc x dim
    centers = np.zeros((nc, dim))
    for k in range(nvc):                     # precomputeSimplices: running sum, then the mean
        centers += T[:, k]
    centers /= nvc
    if dim == 1:
        vol = np.abs(T[:, 1, 0]-T[:, 0, 0])
    else:
        e1, e2 = T[:, 1]-T[:, 0], T[:, 2]-T[:, 0]
        vol = 0.5*np.abs(e1[:, 0]*e2[:, 1]-e1[:, 1]*e2[:, 0])
    hc = np.array([_h_simplex(T[c]) for c in range(nc)])
    hmin = float(hc.min() if hmin is None else hmin)
    if diam is None:
        diam = sqrt(((vertices.max(axis=0)-vertices.min(axis=0))**2).sum())
    H0 = diam/sqrt(8.)
    # the unsymmetric local matrices never see params['target_order'] (fractionalLaplacian2D.pyx:911)
    del target_order
    orders = tables.diag_orders(dim, -dim-2*sFun.max, 1.-dim-2*sFun.max, hmin, H0, num_dofs, None, poly_order=poly_order,
                                min_singularity=-dim-2*sFun.min, min_boundary_singularity=1.-dim-2*sFun.min)
    to, tob = orders['target_order'], orders['b_target_order']
    fe = hasattr(sFun, 'vertex_values')
    if fe:
        smax_cell = sFun.vertex_values[cells].max(axis=1)
    else:
        smax_cell = np.maximum(sFun(centers), sFun(T).max(axis=1))

    def order_at(c, lam, x):
        """the order at the points x of cell c (barycentric coordinates lam, nvc x nq, in the cell's own vertex order)"""
        return lam.T.dot(sFun.vertex_values[cells[c]]) if fe else sFun(x)
    near_cache = {}

    def near(smax):
        if smax not in near_cache:
            near_cache[smax] = tables.near_rules(dim, -dim-2*smax, 1.-dim-2*smax, orders, poly_order=poly_order)
        return near_cache[smax]

    def quad_order(h1, h2, d, smax):
        ld1, ld2 = log(d/h1), log(d/h2)
        if dim == 2:
            c = (0.5*to+0.5)*log(num_dofs*H0**2)
            a1, a2 = abs(log(h1/H0)), abs(log(h2/H0))
            am = max(a1, a2)
            s = max(smax, 0.)
            p1 = max(ceil((c+(s-1.)*a2+am-s*ld2)/(max(ld1, 0)+0.4)), 2)
            p2 = max(ceil((c+(s-1.)*a1+am-s*ld1)/(max(ld2, 0)+0.4)), 2)
        else:
            c = (to+2.)*log(num_dofs*H0)
            s = max(smax, 0.)
            p1 = max(ceil((c+(2.*s-1.)*abs(log(h2/H0))-2.*s*ld2)/(max(ld1, 0)+0.8)), 2)
            p2 = max(ceil((c+(2.*s-1.)*abs(log(h1/H0))-2.*s*ld1)/(max(ld2, 0)+0.8)), 2)
        return int(max(p1, p2))

    def bquad_order(h1, h2, d, smax):
        ld1, ld2 = max(log(d/h1), 0.), max(log(d/h2), 0.)
        s = max(smax, 0.)      # 0.5 (-bsing - 1) with d = 2, 0.5 (-bsing - 1) + ... : see below
        if dim == 2:
            c = (0.5*tob+0.25)*log(num_dofs*H0**2)
            a1, a2 = abs(log(h1/H0)), abs(log(h2/H0))
            am = max(a1, a2)
            p1 = max(ceil((c+am+(s-1.)*a2-s*ld2)/(ld1+0.35)), 2)
            p2 = max(ceil((c+am+(s-1.)*a1-s*ld1)/(ld2+0.35)), 2)
        else:
            # fractionalLaplacian1D.pyx:650: s = max(0.5 (-singularity - 1), 0) with singularity = -2 smax
            s = max(smax-0.5, 0.)
            c = (tob+1.)*log(num_dofs*H0)
            p1 = max(ceil((c+(2.*s-1.)*abs(log(h2/H0))-2.*s*log(d/h2))/(ld1+0.8)), 2)
            p2 = max(ceil((c+(2.*s-1.)*abs(log(h1/H0))-2.*s*log(d/h1))/(ld2+0.8)), 2)
        return int(max(p1, p2))

    def shape(lam):
        return shape_functions(dim, dpe, lam)

    A = np.zeros((num_dofs, num_dofs))

    def scatter(dA, dB, M):
        idx = np.concatenate((dA, dB))
        for i in range(2*dpe):
            if idx[i] < 0:
                continue
            for j in range(2*dpe):
                if idx[j] >= 0:
                    A[idx[i], idx[j]] += M[i, j]

    def local(cA, cB, lamx, lamy, w, x, y):
        gxy = kernel_value(dim, sFun, x, y, s=order_at(cA, lamx, x))
        gyx = kernel_value(dim, sFun, y, x, s=order_at(cB, lamy, y))
        px, py = shape(lamx), shape(lamy)             # nvc x nq
        rowI = np.vstack((px*(w*gxy), -py*(w*gyx)))    # temp PHI[I,0] - temp2 PHI[I,1]
        colJ = np.vstack((px, -py))
        return rowI.dot(colJ.T)

    for c1 in range(nc):
        for c2 in range(c1, nc):
            smax = max(smax_cell[c1], smax_cell[c2])
            orientations = ((c1, c2), ) if c1 == c2 else ((c1, c2), (c2, c1))
            for (cA, cB) in orientations:
                panel, p1, p2 = _proto(cells[cA], cells[cB], cA == cB)
                if panel == 0:
                    d = sqrt(((centers[cA]-centers[cB])**2).sum())
                    o = quad_order(hc[cA], hc[cB], d, smax)
                    b, w1 = tables.regular_rule(o, dim)
                    n = b.shape[1]
                    lamx = np.repeat(b, n, axis=1)
                    lamy = np.tile(b, (1, n))
                    w = np.repeat(w1, n)*np.tile(w1, n)
                    x = lamx.T.dot(T[cA])
                    y = lamy.T.dot(T[cB])
                    M = local(cA, cB, lamx, lamy, w, x, y)*(vol[cA]*vol[cB])
                else:
                    bary, w = near(smax)[('interior', panel)]
                    SA, SB = T[cA][p1], T[cB][p2]
                    x = bary[:nvc].T.dot(SA)
                    y = bary[nvc:].T.dot(SB)
                    lamx = np.zeros((nvc, bary.shape[1]))
                    lamy = np.zeros((nvc, bary.shape[1]))
                    lamx[p1] = bary[:nvc]
                    lamy[p2] = bary[nvc:]
                    M = local(cA, cB, lamx, lamy, w, x, y)*((4. if dim == 2 else 1.)*vol[cA]*vol[cB])
                scatter(dofs[cA], dofs[cB], M)
    if zero_exterior:
        bfacets = np.asarray(bfacets).reshape(-1, dim)
        nvf = dim
        for c1 in range(nc):
            for f in range(bfacets.shape[0]):
                F = vertices[bfacets[f]]
                fc = F.mean(axis=0) if nvf > 1 else F[0]
                if fe:
                    smax = max(smax_cell[c1], float(sFun.vertex_values[bfacets[f]].max()))
                else:
                    smax = max(smax_cell[c1], float(sFun(fc[None])[0]), float(sFun(F).max()))
                panel, p1, p2 = _proto(cells[c1], bfacets[f], False)
                if dim == 2:
                    nrm = np.array([F[1, 1]-F[0, 1], F[0, 0]-F[1, 0]])
                    bvol = sqrt((nrm**2).sum())
                    nrm = nrm/bvol
                else:
                    bvol = 1.
                if panel == 0:
                    d = sqrt(((centers[c1]-fc)**2).sum())
                    o = bquad_order(hc[c1], _h_simplex(F), d, smax)
                    b, w1 = tables.regular_rule(o, dim)
                    bf, wf = tables.regular_rule(o, dim-1)
                    n, m = b.shape[1], bf.shape[1]
                    lamx = np.repeat(b, m, axis=1)
                    x = lamx.T.dot(T[c1])
                    y = np.tile(bf, (1, n)).T.dot(F)
                    w = np.repeat(w1, m)*np.tile(wf, n)
                    g = kernel_value(dim, sFun, x, y, boundary=True, s=order_at(c1, lamx, x))
                    if dim == 2:
                        # gamma_b n.(y-x)/|y-x| (eval_distant_boundary, nonlocalOperator_{SCALAR}.pxi:1069-1108)
                        g = g*((y-x).dot(nrm))/np.sqrt(((x-y)**2).sum(axis=1))
                    px = shape(lamx)
                    M = (px*(w*g)).dot(px.T)*(vol[c1]*bvol)
                else:
                    bary, w = near(smax)[('boundary', panel)]
                    SA, SF = T[c1][p1], F[p2[:nvf]]
                    x = bary[:nvc].T.dot(SA)
                    y = bary[nvc:nvc+nvf].T.dot(SF)
                    lamx = np.zeros((nvc, bary.shape[1]))
                    lamx[p1] = bary[:nvc]
                    g = kernel_value(dim, sFun, x, y, boundary=True, s=order_at(c1, lamx, x))
                    if dim == 2:
                        # fractionalLaplacian2D.pyx:1356-1407: n.(x-y)/|x-y| and the factor -2 vol1 vol2
                        g = g*((x-y).dot(nrm))/np.sqrt(((x-y)**2).sum(axis=1))
                        px = shape(lamx)
                        M = (px*(w*g)).dot(px.T)*(-2.*vol[c1]*bvol)
                    else:
                        px = shape(lamx)
                        M = (px*(w*g)).dot(px.T)*vol[c1]
                idx = dofs[c1]
                for i in range(dpe):
                    for j in range(dpe):
                        if idx[i] >= 0 and idx[j] >= 0:
                            A[idx[i], idx[j]] += M[i, j]
    return A
