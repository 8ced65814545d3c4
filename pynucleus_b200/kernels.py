"""Kernel objects of the nonlocal operators (user-facing drop-in).

Mirrors the factories of nl/PyNucleus_nl/kernels.py:109-231 and the kernel
classes of nl/PyNucleus_nl/kernelsCy.pyx:625-868 (Kernel), :1564-2027
(FractionalKernel), the fractional orders of fractionalOrders.pyx:72-93 and the
scaling constants of kernelNormalization.pyx:70-104.  The objects only carry
parameters; the kernel VALUES on the hot path are evaluated on the GPU from
the parameter block built by ``nonlocalBuilder`` (pnb_kernel_t).
"""
from math import pi, gamma

import numpy as np

FRACTIONAL = 0
INDICATOR = 1
PERIDYNAMIC = 2


class constFractionalOrder:
    """s(x,y) = const (fractionalOrders.pyx:72-93)"""
    symmetric = True
    numParameters = 1

    def __init__(self, s):
        self.value = float(s)
        self.min = self.max = float(s)

    def __call__(self, x, y):
        return self.value

    def __repr__(self):
        return '{}'.format(self.value)


class variableConstFractionalOrder(constFractionalOrder):
    """s(x,y) = const declared as a *variable* order (fractionalOrders.pyx:203-220).  The reference then runs
    its variable-kernel code path (per-pair evalParams, lazily built singular rules, NO.pxi:509-513, 534-535),
    whose result equals the constant-order one to rounding (4.5e-16, tests/golden/disc_varconst*.npz).  On the
    device the order is a constant, so the kernel is assembled by the same kernels; only the host-side flags
    (`kernel.variable`, `kernel.variableOrder`) differ."""
    pass


class _blockFractionalOrder:
    """piecewise constant order s(x,y) = sVals[block(x), block(y)].  The reference evaluates such an order at the cell
    centres, once per ORDERED cell pair (kernel.evalParams, nonlocalOperator_{SCALAR}.pxi:509-513), and visits both
    orientations of a pair when sVals is not symmetric (nonlocalAssembly_{SCALAR}.pxi:1412-1428).

    `labels(points)` (block of every point, uint8 < 4) and `blockOrders()` (the matrix sVals) describe the order to the
    device path, see nonlocalBuilder.setKernel."""
    numParameters = 1

    def _finish(self, sVals):
        self.sVals = np.array(sVals, dtype=np.float64)
        self.min = float(self.sVals.min())
        self.max = float(self.sVals.max())

    def blockOrders(self):
        return self.sVals

    def classes(self):
        """symmetric orders: (distinct orders, pair_class[4][4]) -- pass k takes the cell pairs of class k"""
        assert self.symmetric
        vals = []
        for v in self.sVals.ravel().tolist():
            if v not in vals:
                vals.append(v)
        pc = np.zeros((4, 4), dtype=np.uint8)
        n = self.sVals.shape[0]
        for i in range(n):
            for j in range(n):
                pc[i, j] = vals.index(self.sVals[i, j])
        return vals, pc

    def __call__(self, x, y):
        lx, ly = self.labels(np.atleast_2d(np.asarray(x, dtype=float)))[0], self.labels(np.atleast_2d(np.asarray(y, dtype=float)))[0]
        return float(self.sVals[lx, ly])


class leftRightFractionalOrder(_blockFractionalOrder):
    """interface x_0 = const (fractionalOrders.pyx:285-335): sll / srr for both points left / right of it, slr / srl across
    (x left and y right / x right and y left); symmetric iff slr == srl"""

    def __init__(self, sll, srr, slr=np.nan, srl=np.nan, interface=0.):
        if not np.isfinite(slr):
            slr = 0.5*(sll+srr)
        if not np.isfinite(srl):
            srl = 0.5*(sll+srr)
        self.sll, self.srr, self.slr, self.srl, self.interface = float(sll), float(srr), float(slr), float(srl), float(interface)
        self.symmetric = self.slr == self.srl
        self._finish([[self.sll, self.slr], [self.srl, self.srr]])

    def labels(self, points):
        return (np.asarray(points)[:, 0] >= self.interface).astype(np.uint8)

    def __repr__(self):
        return 'leftRightFractionalOrder(ll={},rr={},lr={},rl={},interface={},sym={})'.format(self.sll, self.srr, self.slr, self.srl,
                                                                                            self.interface, int(self.symmetric))


class constantNonSymFractionalOrder(_blockFractionalOrder):
    """s(x,y) = const declared as an UNSYMMETRIC order (fractionalOrders.pyx:631-638): the reference then runs its
    unsymmetric code path -- fractionalLaplacian*_nonsym, both orientations of every cell pair, order and scaling evaluated
    per quadrature node (the kernel cannot be piecewise, kernels.py:147-149) -- on a constant.  One block, two half-weight
    passes over the two orientations; the per-node scaling equals the constant one (same formula,
    kernelNormalization.pyx:438-439)."""
    symmetric = False

    def __init__(self, s):
        self.value_ = float(s)
        self._finish([[float(s)]])

    def labels(self, points):
        return np.zeros(np.asarray(points).shape[0], dtype=np.uint8)

    def __repr__(self):
        return 'constantNonSymFractionalOrder({})'.format(self.value_)


class piecewiseConstantFractionalOrder(_blockFractionalOrder):
    """blocks given by an indicator function x -> block number (fractionalOrders.pyx:218-283); off-diagonal orders that
    are not finite default to the mean of the two diagonal ones; symmetric iff |sVals - sVals^T| < 1e-10"""

    def __init__(self, dim, blockIndicator, sVals):
        sVals = np.array(sVals, dtype=np.float64)
        assert sVals.shape[0] == sVals.shape[1]
        n = sVals.shape[0]
        for i in range(n):
            for j in range(n):
                if i == j:
                    assert np.isfinite(sVals[i, j])
                elif not np.isfinite(sVals[i, j]):
                    sVals[i, j] = 0.5*(sVals[i, i]+sVals[j, j])
        self.dim = dim
        self.blockIndicator = blockIndicator
        self.symmetric = bool(np.absolute(sVals-sVals.T).max() < 1e-10)
        self._finish(sVals)

    @property
    def numBlocks(self):
        return self.sVals.shape[0]

    def labels(self, points):
        lab = np.array([int(self.blockIndicator(np.asarray(q, dtype=float))) for q in np.asarray(points)], dtype=np.int64)
        if lab.size and (lab.min() < 0 or lab.max() >= self.numBlocks):
            raise ValueError('block indicator outside [0, numBlocks)')
        return lab.astype(np.uint8)

    def __repr__(self):
        return 'piecewiseConstantFractionalOrder(numBlocks={},sym={})'.format(self.numBlocks, self.symmetric)


class layersFractionalOrder(_blockFractionalOrder):
    """layers along the last coordinate (fractionalOrders.pyx:822-893): layer i = [b_i, b_{i+1}], the first match wins,
    points outside fall into the first / last layer"""

    def __init__(self, dim, layerBoundaries, layerOrders):
        self.dim = dim
        self.layerBoundaries = np.array(layerBoundaries, dtype=np.float64)
        layerOrders = np.array(layerOrders, dtype=np.float64)
        n = self.layerBoundaries.shape[0]-1
        assert layerOrders.shape == (n, n)
        self.layerOrders = layerOrders
        self.symmetric = bool((layerOrders == layerOrders.T).all())
        self._finish(layerOrders)

    def labels(self, points):
        c = np.asarray(points)[:, self.dim-1]
        b = self.layerBoundaries
        n = b.shape[0]-1
        # first i with b_i <= c <= b_{i+1}
        lab = np.searchsorted(b[1:], c, side='left')
        lab = np.where(c <= b[0], 0, np.where(c >= b[n], n-1, np.minimum(lab, n-1)))
        return lab.astype(np.uint8)

    def __repr__(self):
        return 'layersFractionalOrder(numLayers={})'.format(self.layerOrders.shape[0])


class innerOuterFractionalOrder(_blockFractionalOrder):
    """ball |x-center| < r and its complement (fractionalOrders.pyx:675-733)"""

    def __init__(self, dim, sii, soo, r, center, sio=np.nan, soi=np.nan):
        if not np.isfinite(sio):
            sio = 0.5*(sii+soo)
        if not np.isfinite(soi):
            soi = 0.5*(sii+soo)
        self.dim, self.r2, self.center = dim, float(r)*float(r), np.array(center, dtype=np.float64)
        self.sii, self.soo, self.sio, self.soi = float(sii), float(soo), float(sio), float(soi)
        self.symmetric = self.sio == self.soi
        self._finish([[self.sii, self.sio], [self.soi, self.soo]])

    def labels(self, points):
        p = np.asarray(points)
        r2 = np.zeros(p.shape[0])
        for i in range(self.dim):
            r2 = r2+(p[:, i]-self.center[i])**2
        return (r2 >= self.r2).astype(np.uint8)

    def __repr__(self):
        return 'innerOuterFractionalOrder(ii={},oo={},io={},oi={},r={},sym={})'.format(self.sii, self.soo, self.sio, self.soi,
                                                                                     np.sqrt(self.r2), self.symmetric)


class islandsFractionalOrder(_blockFractionalOrder):
    """islands r <= |x_i| <= r2 in every coordinate (fractionalOrders.pyx:737-819, 2D)"""

    def __init__(self, sii, soo, r, r2, sio=np.nan, soi=np.nan):
        if not np.isfinite(sio):
            sio = 0.5*(sii+soo)
        if not np.isfinite(soi):
            soi = 0.5*(sii+soo)
        self.r, self.r2 = float(r), float(r2)
        self.sii, self.soo, self.sio, self.soi = float(sii), float(soo), float(sio), float(soi)
        self.symmetric = self.sio == self.soi
        self._finish([[self.sii, self.sio], [self.soi, self.soo]])

    def labels(self, points):
        p = np.absolute(np.asarray(points)[:, :2])
        inside = ((p >= self.r) & (p <= self.r2)).all(axis=1)
        return (~inside).astype(np.uint8)

    def __repr__(self):
        return 'islandsFractionalOrder(ii={},oo={},io={},oi={},r={},r2={},sym={})'.format(self.sii, self.soo, self.sio, self.soi,
                                                                                        self.r, self.r2, self.symmetric)


class singleVariableUnsymmetricFractionalOrder:
    """s(x, y) = sFun(x): an order that varies INSIDE the cells (fractionalOrders.pyx:153-183).  The kernel cannot be
    piecewise (kernels.py:147-149): the reference evaluates order, scaling constant and kernel at every quadrature node
    (updateAndEvalFractional, kernelsCy.pyx:596-622) inside the unsymmetric local matrices.  Device path:
    pnb_dense_assemble_varorder; `orderFunction()` describes sFun to it (PNB_ORDERFUN_*)."""
    symmetric = False
    numParameters = 2
    ORDERFUN_CONST, ORDERFUN_SMOOTHSTEP, ORDERFUN_LINEARSTEP, ORDERFUN_SMOOTHSTEP_RADIAL = 0, 1, 2, 3

    def orderFunction(self):
        return self.fun, self.sl, self.sr, self.r, self.slope, self.interface

    def evaluate(self, points):
        """s at an array of points (..., dim), the reference's formulas (fractionalOrders.pyx:389-416, 447-470)"""
        t = np.asarray(points, dtype=np.float64)[..., 0]
        if self.fun == self.ORDERFUN_LINEARSTEP:
            v = self.sl+self.slope*(t-self.interface+self.r)
        else:
            u = (t-self.interface)*self.slope+0.5
            v = self.sl+(self.sr-self.sl)*(3.0*u**2-2.0*u**3)
        return np.where(t < self.interface-self.r, self.sl, np.where(t > self.interface+self.r, self.sr, v))

    def __call__(self, x, y):
        return float(self.evaluate(np.atleast_2d(np.asarray(x, dtype=float)))[0])


class smoothedLeftRightFractionalOrder(singleVariableUnsymmetricFractionalOrder):
    """sl left of interface - r, sr right of interface + r, a cubic step in between (fractionalOrders.pyx:641-645 with
    smoothStep :389-416); the order of the driver flag `--s twoDomainNonSym(sl,sr)` (nonlocalProblems.py:95)"""
    fun = singleVariableUnsymmetricFractionalOrder.ORDERFUN_SMOOTHSTEP

    def __init__(self, sl, sr, r=0.1, slope=200., interface=0.):
        self.sl, self.sr, self.r, self.interface = float(sl), float(sr), float(r), float(interface)
        self.slope = 0.5/self.r
        self.min, self.max = min(self.sl, self.sr), max(self.sl, self.sr)

    def __repr__(self):
        return 'smoothedLeftRightFractionalOrder(sl={},sr={},r={},interface={})'.format(self.sl, self.sr, self.r, self.interface)


class linearLeftRightFractionalOrder(singleVariableUnsymmetricFractionalOrder):
    """the same with a linear ramp (fractionalOrders.pyx:648-651 with linearStep :447-470)"""
    fun = singleVariableUnsymmetricFractionalOrder.ORDERFUN_LINEARSTEP

    def __init__(self, sl, sr, r=0.1, interface=0.):
        self.sl, self.sr, self.r, self.interface = float(sl), float(sr), float(r), float(interface)
        self.slope = 0.5*(self.sr-self.sl)/self.r
        self.min, self.max = min(self.sl, self.sr), max(self.sl, self.sr)

    def __repr__(self):
        return 'linearLeftRightFractionalOrder(sl={},sr={},r={},interface={})'.format(self.sl, self.sr, self.r, self.interface)


class feFractionalOrder(singleVariableUnsymmetricFractionalOrder):
    """s(x, y) = s_h(x) for a P1 finite element function s_h = sum_i u_i phi_i (fractionalOrders.pyx:660-668; dofs of the
    map that are boundary dofs contribute nothing, lookupExtended.evalPtr :573-586).  The reference takes an fe_vector and looks
    the cell of every point up; here the function lives on the ASSEMBLY mesh (`dm.mesh` must be the mesh of the operator's
    DoFMap), so that the cell of a quadrature node is known and the order is `sum_k lambda_k(x) u[dof_k]` on the device
    (PNB_ORDERFUN_FE).  Any order that varies smoothly over the mesh can be supplied this way through its nodal values."""
    fun = 4

    def __init__(self, dm, u, smin=None, smax=None):
        if dm.polynomialOrder != 1:
            raise NotImplementedError('feFractionalOrder: P1 functions')
        self.dm = dm
        self.u = np.ascontiguousarray(u, dtype=np.float64)
        assert self.u.shape == (dm.num_dofs, )
        self.min = float(self.u.min() if smin is None else smin)
        self.max = float(self.u.max() if smax is None else smax)
        # the bounds double as the two values whose powers come from tables on the device
        self.sl, self.sr, self.r, self.slope, self.interface = self.min, self.max, 0., 0., 0.

    def vertexValues(self, mesh):
        """the order at the vertices of `mesh` (which must be the function's own mesh)"""
        m = self.dm.mesh
        if m is not mesh and not (m.vertices.shape == mesh.vertices.shape and np.array_equal(m.vertices, mesh.vertices)
                                  and np.array_equal(m.cells, mesh.cells)):
            raise NotImplementedError('feFractionalOrder: the order must be given on the mesh of the operator')
        v = np.zeros(mesh.num_vertices)
        for k in range(self.dm.dofs.shape[1]):
            ok = self.dm.dofs[:, k] >= 0
            v[mesh.cells[ok, k]] = self.u[self.dm.dofs[ok, k]]
        return v

    def evaluate(self, points):
        raise NotImplementedError('feFractionalOrder is evaluated cell by cell on the device (no point lookup on the host)')

    def __repr__(self):
        return 'feFractionalOrder({} dofs, min={}, max={})'.format(self.u.shape[0], self.min, self.max)


def variableFractionalLaplacianScaling(dim, s):
    """C(d, s(x,y)) / 2 of a normalised kernel with infinite horizon (kernelNormalization.pyx:438-439); s may be an array"""
    from scipy.special import gamma as gamma_
    s = np.asarray(s, dtype=np.float64)
    return 2.0**(2.0*s)*s*gamma_(s+0.5*dim)*pi**(-0.5*dim)/gamma_(1.0-s)*0.5


class constant:
    """constant function, used for the horizon (fem functions.pyx)"""

    def __init__(self, value):
        self.value = float(value)

    def __call__(self, x):
        return self.value


def constantFractionalLaplacianScaling(dim, s, horizon, tempered=0.):
    """C(d,s,delta)/2 (kernelNormalization.pyx:70-89)"""
    if 1. < s < 2.:
        s = s-1.
    if horizon <= 0. or s <= 0. or s >= 1.:
        return np.nan
    if horizon < np.inf:
        return (2.-2*s)*pow(horizon, 2*s-2.)*dim*gamma(0.5*dim)/pow(pi, 0.5*dim)*0.5
    if tempered == 0. or s == 0.5:
        return 2.0**(2.0*s)*s*gamma(s+0.5*dim)/pow(pi, 0.5*dim)/gamma(1.0-s)*0.5
    return gamma(0.5*dim)/abs(gamma(-2*s))/pow(pi, 0.5*dim)*0.5*0.5


class FractionalKernel:
    """gamma(x,y) = C(d,s) |x-y|^{-d-2s}  (boundary form: |x-y|^{-(d-1)-2s})

    Attribute names follow kernelsCy.pyx:1564-1640."""
    kernelType = FRACTIONAL
    valueSize = 1

    def __init__(self, dim, s, horizon, scaling, boundary=False, phi=None, piecewise=True, tempered=0.):
        self.dim = int(dim)
        self.s = s
        self.horizon = horizon
        # tempered kernel C |x-y|^(-d-2s) exp(-tempered |x-y|) (temperedFracKernelInfinite*, kernelsCy.pyx:186-213)
        self.tempered = float(tempered)
        self.boundary = boundary
        self.piecewise = piecewise
        self.phi = phi
        self.scalingPrePhi = scaling
        self.scalingValue = scaling if phi is None else phi*scaling
        self.variableOrder = isinstance(s, variableConstFractionalOrder) or not isinstance(s, constFractionalOrder)
        self.variableHorizon = False
        self.variableScaling = False
        self.variable = self.variableOrder
        self.symmetric = bool(s.symmetric)
        self.horizonValue = horizon.value
        self.horizonValue2 = horizon.value**2
        self.finiteHorizon = horizon.value != np.inf
        self.complement = False
        off = 1. if boundary else 0.
        if hasattr(s, 'value'):
            self.sValue = s.value
            self.singularityValue = off-self.dim-2*self.sValue
        else:
            # piecewise variable order: the current values are set per cell pair (evalParams); they start out as nan
            # (kernelsCy.pyx:1606-1611)
            self.sValue = np.nan
            self.singularityValue = np.nan
            self.variableScaling = True
        self.min_singularity = off-self.dim-2*s.min
        self.max_singularity = off-self.dim-2*s.max

    def getModifiedKernel(self, s=None, horizon=None, scaling=None):
        s = self.s if s is None else s
        horizon = self.horizon if horizon is None else horizon
        return getFractionalKernel(self.dim, s, horizon, scaling=scaling, piecewise=self.piecewise, boundary=self.boundary)

    def getBoundaryKernel(self):
        """kernel of the Gauss-theorem surface term, scaled by 1/s (kernelsCy.pyx:1982-2027)"""
        phi = 1./self.s.value if hasattr(self.s, 'value') else None
        # `tempered` is not handed on (kernelsCy.pyx:2011-2020): the surface terms of a tempered kernel are the plain power
        # law with the tempered scaling constant
        return FractionalKernel(self.dim, self.s, self.horizon, self.scalingPrePhi, boundary=True,
                                phi=phi, piecewise=self.piecewise)

    def __call__(self, x, y):
        x = np.atleast_1d(np.asarray(x, dtype=float))
        y = np.atleast_1d(np.asarray(y, dtype=float))
        d2 = float(((x-y)**2).sum())
        if self.finiteHorizon and d2 > self.horizonValue2:
            return 0.
        if isinstance(self.s, singleVariableUnsymmetricFractionalOrder):
            # order and scaling per point (updateAndEvalFractional, kernelsCy.pyx:596-622); boundary form: phi = 1/s
            sv = self.s(x, y)
            C = float(variableFractionalLaplacianScaling(self.dim, sv))
            if not self.boundary:
                return C*pow(d2, -0.5*self.dim-sv)
            return C/sv*pow(d2, -0.5*(self.dim-1)-sv)
        if not self.boundary:
            if self.tempered != 0.:
                return self.scalingValue*pow(d2, -0.5*self.dim-self.sValue)*np.exp(-self.tempered*np.sqrt(d2))
            return self.scalingValue*pow(d2, -0.5*self.dim-self.sValue)
        return self.scalingValue*pow(d2, -0.5*(self.dim-1)-self.sValue)

    def __repr__(self):
        return 'kernel({}fractional, s={}, horizon={}, scaling={}{}{})'.format('tempered-' if self.tempered != 0. else '', self.s,
                                                                             self.horizonValue, self.scalingValue,
                                                                             ', tempered={}'.format(self.tempered) if self.tempered != 0. else '',
                                                                             ', boundary' if self.boundary else '')


INDICATOR, PERIDYNAMIC = 1, 2      # kernel_params.pxi:88-90
GAUSSIAN, EXPONENTIAL = 3, 8
# smooth factors of pnb_dense_assemble_element_smooth (include/pnb200.h)
SMOOTH_NONE, SMOOTH_EXP_R, SMOOTH_EXP_R2, SMOOTH_ERFC_R, SMOOTH_EXP_R2_OVER_R = 0, 1, 2, 3, 4


def getKernelEnum(kernelTypeString):
    """kernelsCy.pyx:49-59"""
    k = kernelTypeString.upper()
    if k == 'FRACTIONAL':
        return FRACTIONAL
    if k in ('INDICATOR', 'CONSTANT'):
        return INDICATOR
    if k in ('INVERSEDISTANCE', 'INVERSEOFDISTANCE', 'PERIDYNAMIC'):
        return PERIDYNAMIC
    if k == 'GAUSSIAN':
        return GAUSSIAN
    if k == 'EXPONENTIAL':
        return EXPONENTIAL
    raise NotImplementedError(kernelTypeString)


def constantIntegrableScaling(kType, dim, horizon, gaussian_variance=1., exponentialRate=1.):
    """normalisation of the integrable kernels on the l2 ball, and of the Gaussian / exponential kernels on the full space
    (kernelNormalization.pyx:225-283)"""
    if horizon <= 0.:
        return np.nan
    if kType == GAUSSIAN and horizon == np.inf:
        if dim == 1:
            return 1.0/np.sqrt(2.0*pi*gaussian_variance)/2.
        if dim == 2:
            return 1.0/(2.0*pi*gaussian_variance)/2.
    if kType == EXPONENTIAL and horizon == np.inf and dim == 1:
        return exponentialRate**3/2.0/2.
    if kType == INDICATOR:
        if dim == 1:
            return 3./horizon**3/2.
        if dim == 2:
            return 8./pi/horizon**4/2.
    elif kType == PERIDYNAMIC:
        if dim == 1:
            return 2./horizon**2/2.
        if dim == 2:
            return 6./pi/horizon**3/2.
    raise NotImplementedError()


class Kernel:
    """integrable kernels gamma(x,y) = C |x-y|^singularity chi(|x-y| <= delta): 'constant' / 'indicator'
    (singularity 0, kernelsCy.pyx:273-295) and 'inverseDistance' / 'peridynamic' (singularity -1, :321-359) on the
    l2 ball (ball2_retriangulation).  Attribute names follow kernelsCy.pyx:620-700."""
    valueSize = 1
    s = None
    sValue = 0.

    def __init__(self, dim, kType, horizon, scaling, boundary=False, phi=None, piecewise=True, variance=1., exponentialRate=1.):
        if kType not in (INDICATOR, PERIDYNAMIC, GAUSSIAN, EXPONENTIAL):
            raise NotImplementedError('kernel type {} is not supported yet'.format(kType))
        self.dim = int(dim)
        self.kernelType = kType
        self.horizon = horizon
        self.boundary = boundary
        self.piecewise = piecewise
        self.phi = phi
        self.scalingPrePhi = scaling
        self.scalingValue = scaling if phi is None else phi*scaling
        self.variableOrder = self.variableHorizon = self.variableScaling = self.variable = False
        self.symmetric = True
        self.horizonValue = horizon.value
        self.horizonValue2 = horizon.value**2
        self.finiteHorizon = horizon.value != np.inf
        if kType in (GAUSSIAN, EXPONENTIAL):
            # Gaussian C exp(-|x-y|^2 / (2 variance^d)) and exponential C exp(-rate |x-y|) on the full space
            # (gaussianKernel*, exponentialKernel: kernelsCy.pyx:388-477; fEXPONENTINVERSE: Kernel.__init__ :690-697), the
            # kernels of the reference's driver tests `--interaction fullSpace --horizon inf`
            if self.finiteHorizon:
                raise NotImplementedError('Gaussian / exponential kernels: infinite horizon (interaction fullSpace)')
            if kType == EXPONENTIAL and self.dim != 1:
                raise NotImplementedError('exponential kernel: 1D (the reference has no normalisation for 2D)')
            self.variance, self.exponentialRate = float(variance), float(exponentialRate)
            self.exponentInverse = 0.5/self.variance**self.dim if kType == GAUSSIAN else self.exponentialRate
        elif not self.finiteHorizon:
            raise NotImplementedError('integrable kernels need a finite horizon')
        self.complement = False
        self.singularityValue = -1. if kType == PERIDYNAMIC else 0.
        self.min_singularity = self.max_singularity = self.singularityValue

    def getModifiedKernel(self, horizon=None, scaling=None):
        horizon = self.horizon if horizon is None else horizon
        return getIntegrableKernel(self.dim, self.kernelType, horizon, scaling=scaling, piecewise=self.piecewise,
                                   variance=getattr(self, 'variance', 1.), exponentialRate=getattr(self, 'exponentialRate', 1.))

    def getBoundaryKernel(self):
        """The surface forms of the integrable kernels (kernelsCy.pyx:297-318, 361-386) only enter operators with a zero
        exterior, which a finite horizon rules out (nonlocalAssembly_{SCALAR}.pxi:918-921); this object only carries the
        singularity the boundary tables are built for."""
        bk = Kernel.__new__(Kernel)
        bk.__dict__.update(self.__dict__)
        bk.boundary = True
        if self.kernelType in (GAUSSIAN, EXPONENTIAL):
            # surface forms gaussianKernel{1,2}Dboundary / exponentialKernelBoundary (kernelsCy.pyx:418-445, 463-477): same
            # scaling and singularity entries as the interior kernel
            return bk
        bk.scalingValue = 0.
        bk.singularityValue = self.singularityValue+1.
        bk.min_singularity = bk.max_singularity = bk.singularityValue
        return bk

    def smoothFactors(self):
        """(mode, a, boundary mode, boundary a, constant of the boundary power law) for pnb_dense_assemble_element_smooth"""
        C, a = self.scalingValue, self.exponentInverse
        if self.kernelType == EXPONENTIAL:
            return SMOOTH_EXP_R, a, SMOOTH_EXP_R, a, 2.0*C/a
        if self.dim == 1:
            return SMOOTH_EXP_R2, a, SMOOTH_ERFC_R, a, C*np.sqrt(pi/a)
        return SMOOTH_EXP_R2, a, SMOOTH_EXP_R2_OVER_R, a, C/a

    def __call__(self, x, y):
        x = np.atleast_1d(np.asarray(x, dtype=float))
        y = np.atleast_1d(np.asarray(y, dtype=float))
        d2 = float(((x-y)**2).sum())
        if d2 > self.horizonValue2:
            return 0.
        if self.kernelType in (GAUSSIAN, EXPONENTIAL):
            from scipy.special import erfc
            C, a = self.scalingValue, self.exponentInverse
            if self.kernelType == EXPONENTIAL:
                return (2.0*C/a if self.boundary else C)*np.exp(-a*np.sqrt(d2))
            if not self.boundary:
                return C*np.exp(-d2*a)
            if self.dim == 1:
                return C*np.sqrt(pi/a)*erfc(np.sqrt(d2*a))
            return C/(d2*a)*np.exp(-d2*a)*np.sqrt(d2)
        return self.scalingValue*pow(d2, 0.5*self.singularityValue)

    def __repr__(self):
        name = {INDICATOR: 'indicator', PERIDYNAMIC: 'peridynamic', GAUSSIAN: 'Gaussian', EXPONENTIAL: 'exponential'}[self.kernelType]
        return 'kernel({}, horizon={}, scaling={})'.format(name, self.horizonValue, self.scalingValue)


def getIntegrableKernel(dim, kernel, horizon, scaling=None, interaction=None, normalized=True, piecewise=True, phi=None,
                        boundary=False, variance=1., exponentialRate=1., **kwargs):
    """kernels.py:172-202"""
    dim = getattr(dim, 'dim', dim)
    kType = getKernelEnum(kernel) if isinstance(kernel, str) else int(kernel)
    horizonFun = _getHorizon(horizon)
    if interaction is not None and interaction not in ('ball2', 'fullSpace'):
        raise NotImplementedError('interaction domains: the l2 ball, the full space for an infinite horizon')
    if interaction == 'fullSpace' and horizonFun.value != np.inf:
        raise NotImplementedError('the full space needs an infinite horizon')
    if scaling is None:
        scaling = (constantIntegrableScaling(kType, dim, horizonFun.value, gaussian_variance=variance, exponentialRate=exponentialRate)
                   if normalized else 0.5)
    return Kernel(dim, kType, horizonFun, scaling, boundary=boundary, phi=phi, piecewise=piecewise, variance=variance,
                  exponentialRate=exponentialRate)


def _getFractionalOrder(s):
    if isinstance(s, (int, float)):
        return constFractionalOrder(s)
    return s


def _getHorizon(horizon):
    if horizon is None:
        return constant(np.inf)
    if isinstance(horizon, (int, float)):
        return constant(horizon)
    return horizon


def getFractionalKernel(dim, s, horizon=None, interaction=None, scaling=None, normalized=True, piecewise=True,
                        phi=None, boundary=False, derivative=0, tempered=0., max_horizon=np.nan, manifold=False):
    """kernels.py:109-165"""
    dim = getattr(dim, 'dim', dim)
    sFun = _getFractionalOrder(s)
    horizonFun = _getHorizon(horizon)
    if derivative != 0 or manifold:
        raise NotImplementedError('derivative / manifold kernels are outside the accelerated path')
    if tempered != 0. and (not isinstance(sFun, constFractionalOrder) or isinstance(sFun, variableConstFractionalOrder)
                           or horizonFun.value != np.inf or boundary):
        raise NotImplementedError('tempered kernels: constant order, infinite horizon')
    if isinstance(sFun, _blockFractionalOrder):
        if horizonFun.value != np.inf or not normalized or scaling is not None:
            raise NotImplementedError('piecewise orders: infinite horizon, normalised kernels only')
        if isinstance(sFun, constantNonSymFractionalOrder):
            piecewise = False       # kernels.py:147-149: single-variable unsymmetric orders cannot be piecewise
        elif not piecewise:
            raise NotImplementedError('orders evaluated per quadrature node (piecewise=False) are outside the accelerated path')
        # the scaling is a function of s(x,y) (variableFractionalLaplacianScaling, kernelNormalization.pyx:421-499):
        # evaluated per class by the builder
        return FractionalKernel(dim, sFun, horizonFun, np.nan, boundary=boundary, phi=phi, piecewise=piecewise)
    if isinstance(sFun, singleVariableUnsymmetricFractionalOrder):
        if horizonFun.value != np.inf or not normalized or scaling is not None or phi is not None:
            raise NotImplementedError('orders varying inside a cell: infinite horizon, normalised kernels only')
        # kernels.py:147-149: "Variable s kernels cannot be piecewise. Switching to piecewise == False."
        return FractionalKernel(dim, sFun, horizonFun, np.nan, boundary=boundary, phi=None, piecewise=False)
    if not isinstance(sFun, constFractionalOrder):
        raise NotImplementedError('this variable fractional order is not supported yet')
    if scaling is None:
        scaling = constantFractionalLaplacianScaling(dim, sFun.value, horizonFun.value, tempered) if normalized else 0.5
    if boundary and phi is None:
        phi = 1./sFun.value
    return FractionalKernel(dim, sFun, horizonFun, scaling, boundary=boundary, phi=phi, piecewise=piecewise, tempered=tempered)


def getKernel(dim, s=None, horizon=None, scaling=None, interaction=None, normalized=True, piecewise=True, phi=None,
              kernel=FRACTIONAL, boundary=False, **kwargs):
    """kernels.py:213-231"""
    kType = getKernelEnum(kernel) if isinstance(kernel, str) else int(kernel)
    if kType == FRACTIONAL:
        return getFractionalKernel(dim, s, horizon, interaction, scaling, normalized, piecewise, phi, boundary)
    return getIntegrableKernel(dim, kType, horizon, scaling=scaling, interaction=interaction, normalized=normalized,
                               piecewise=piecewise, phi=phi, **{k: v for k, v in kwargs.items() if k in ('variance', 'exponentialRate')})
