"""ORACLE INFRASTRUCTURE: fixtures for fractional orders that VARY INSIDE A CELL (kernel.piecewise == False), produced by
running the REFERENCE ITSELF (stub-built copy in oracle/_ref, see build_reference.sh):

    PYTHONPATH=oracle/_ref python oracle/refbuild/make_golden_varorder.py

smoothedLeftRightFractionalOrder is the order behind the driver flag `--s twoDomainNonSym(sl,sr)`
(nonlocalProblems.py:95); the reference evaluates s, the scaling constant and the kernel per quadrature node
(updateAndEvalFractional, kernelsCy.pyx:596-622) inside fractionalLaplacian{1,2}D_nonsym and builds a singular rule per
distinct pair singularity (evalParamsOnSimplices, kernelsCy.pyx:1826-1850).  Every array is an output of reference code.
"""
import os
import sys
import warnings
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..', '..'))
OUT = os.path.join(ROOT, 'tests', 'golden')

from PyNucleus_fem.mesh import simpleInterval, uniform_disc  # noqa: E402
from PyNucleus_fem.DoFMaps import P0_DoFMap, P1_DoFMap, P2_DoFMap  # noqa: E402
from PyNucleus_nl.kernels import getFractionalKernel  # noqa: E402
from PyNucleus_nl.nonlocalAssembly import nonlocalBuilder  # noqa: E402
from PyNucleus_nl.fractionalOrders import smoothedLeftRightFractionalOrder, linearLeftRightFractionalOrder  # noqa: E402
from make_golden_nonsym import mesh_arrays  # noqa: E402


def case(dim, noRef, sFun, name, params, extra, element='P1'):
    mesh = uniform_disc() if dim == 2 else simpleInterval(-1, 1)
    for _ in range(noRef):
        mesh = mesh.refine()
    dm = {'P0': P0_DoFMap, 'P1': P1_DoFMap, 'P2': P2_DoFMap}[element](mesh)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        kernel = getFractionalKernel(dim, sFun, np.inf)
    assert kernel.variable and not kernel.piecewise and not kernel.symmetric
    out = mesh_arrays(mesh, dm)
    for ze, key in ((True, 'A'), (False, 'A_interior')):
        b = nonlocalBuilder(dm, kernel, dict(params), zeroExterior=ze)
        out[key] = np.array(b.getDense().data)
    # the order and the kernel at a few points, straight from the reference's objects
    rng = np.random.default_rng(5)
    X = rng.uniform(-1, 1, size=(64, dim))*0.7
    Y = rng.uniform(-1, 1, size=(64, dim))*0.7
    out['points_x'], out['points_y'] = X, Y
    out['s_values'] = np.array([kernel.s(X[i], Y[i]) for i in range(64)])
    out['kernel_values'] = np.array([kernel(X[i], Y[i]) for i in range(64)])
    kb = kernel.getBoundaryKernel()
    out['bkernel_values'] = np.array([kb(X[i], Y[i]) for i in range(64)])
    out.update(symmetric=0, element=element, local_matrix=type(b.local_matrix).__name__, target_order_used=b.local_matrix.target_order,
               quad_order_diagonal=b.local_matrix.quad_order_diagonal,
               btarget_order_used=b.local_matrix_zeroExterior.target_order,
               bquad_order_diagonal=b.local_matrix_zeroExterior.quad_order_diagonal, **extra)
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, out['A'].shape, type(b.local_matrix).__name__, 'asym of A: %.3e' % np.abs(out['A']-out['A'].T).max(), flush=True)


if __name__ == '__main__':
    which = sys.argv[1:] or ['all']
    if 'all' in which or '1d' in which:
        case(1, 5, smoothedLeftRightFractionalOrder(0.25, 0.75), 'varorder_interval_smoothed_r5', {},
             dict(kind='smoothedLeftRight', sl=0.25, sr=0.75, r=0.1, interface=0.))
        case(1, 6, smoothedLeftRightFractionalOrder(0.75, 0.25, r=0.3, interface=0.2), 'varorder_interval_smoothed_r6', {},
             dict(kind='smoothedLeftRight', sl=0.75, sr=0.25, r=0.3, interface=0.2))
        case(1, 5, linearLeftRightFractionalOrder(0.3, 0.6, r=0.25), 'varorder_interval_linear_r5', {},
             dict(kind='linearLeftRight', sl=0.3, sr=0.6, r=0.25, interface=0.))
    if 'all' in which or 'elements' in which:
        case(1, 4, smoothedLeftRightFractionalOrder(0.25, 0.75, r=0.3), 'varorder_p2_interval_smoothed_r4', {},
             dict(kind='smoothedLeftRight', sl=0.25, sr=0.75, r=0.3, interface=0.), element='P2')
    if 'all' in which or '2d' in which:
        case(2, 2, smoothedLeftRightFractionalOrder(0.25, 0.75), 'varorder_disc_smoothed_r2', {'target_order': 0.5},
             dict(kind='smoothedLeftRight', sl=0.25, sr=0.75, r=0.1, interface=0.))
        case(2, 3, smoothedLeftRightFractionalOrder(0.75, 0.25, r=0.3), 'varorder_disc_smoothed_r3', {'target_order': 0.5},
             dict(kind='smoothedLeftRight', sl=0.75, sr=0.25, r=0.3, interface=0.))
        case(2, 1, smoothedLeftRightFractionalOrder(0.25, 0.75, r=0.3), 'varorder_p2_disc_smoothed_r1', {'target_order': 0.5},
             dict(kind='smoothedLeftRight', sl=0.25, sr=0.75, r=0.3, interface=0.), element='P2')
        case(2, 2, smoothedLeftRightFractionalOrder(0.2, 0.4, r=0.3), 'varorder_p0_disc_smoothed_r2', {'target_order': 0.5},
             dict(kind='smoothedLeftRight', sl=0.2, sr=0.4, r=0.3, interface=0.), element='P0')
        # smoothedInnerOuterFractionalOrder cannot be constructed in the reference (fractionalOrders.pyx:657 passes
        # numParameters = 0, which fractionalOrderBase.__init__ :51 rejects): no fixture


def case_fe(dim, noRef, fun, smin, smax, name, params):
    """feFractionalOrder (fractionalOrders.pyx:660-668): the order is a P1 finite element function on the mesh of the operator"""
    from PyNucleus_fem import NO_BOUNDARY
    from PyNucleus_fem.functions import Lambda
    from PyNucleus_nl.fractionalOrders import feFractionalOrder
    mesh = uniform_disc() if dim == 2 else simpleInterval(-1, 1)
    for _ in range(noRef):
        mesh = mesh.refine()
    dm = P1_DoFMap(mesh)
    dms = P1_DoFMap(mesh, NO_BOUNDARY)
    vec = dms.interpolate(Lambda(fun))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        kernel = getFractionalKernel(dim, feFractionalOrder(vec, smin, smax), np.inf)
    assert kernel.variable and not kernel.piecewise and not kernel.symmetric
    out = mesh_arrays(mesh, dm)
    for ze, key in ((True, 'A'), (False, 'A_interior')):
        b = nonlocalBuilder(dm, kernel, dict(params), zeroExterior=ze)
        out[key] = np.array(b.getDense().data)
    out.update(symmetric=0, element='P1', kind='fe', order_dofs=np.array(dms.dofs), order_values=np.array(vec), smin=smin, smax=smax,
               target_order_used=b.local_matrix.target_order, quad_order_diagonal=b.local_matrix.quad_order_diagonal,
               btarget_order_used=b.local_matrix_zeroExterior.target_order,
               bquad_order_diagonal=b.local_matrix_zeroExterior.quad_order_diagonal)
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, out['A'].shape, type(b.local_matrix).__name__, 'asym of A: %.3e' % np.abs(out['A']-out['A'].T).max(), flush=True)


if __name__ == '__main__' and ('all' in (sys.argv[1:] or ['all']) or 'fe' in sys.argv[1:]):
    case_fe(1, 5, lambda x: 0.5+0.2*np.sin(2*x[0]), 0.3, 0.7, 'varorder_fe_interval_r5', {})
    case_fe(2, 2, lambda x: 0.5+0.2*np.sin(2*x[0])*np.cos(x[1]), 0.3, 0.7, 'varorder_fe_disc_r2', {'target_order': 0.5})
    case_fe(2, 3, lambda x: 0.45+0.3*x[0]*x[1], 0.25, 0.65, 'varorder_fe_disc_r3', {'target_order': 0.5})
