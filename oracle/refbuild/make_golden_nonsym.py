"""ORACLE INFRASTRUCTURE: fixtures for UNSYMMETRIC piecewise constant fractional orders, produced by running the
REFERENCE ITSELF (stub-built copy in oracle/_ref, see build_reference.sh):

    PYTHONPATH=oracle/_ref python oracle/refbuild/make_golden_nonsym.py

The reference then uses fractionalLaplacian{1,2}D_nonsym (fractionalLaplacian2D.pyx:894-1184) and visits both
orientations of every cell pair (nonlocalAssembly_{SCALAR}.pxi:1412-1428).  Every array is an output of reference code.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..', '..'))
OUT = os.path.join(ROOT, 'tests', 'golden')

from PyNucleus_fem.mesh import simpleInterval, uniform_disc  # noqa: E402
from PyNucleus_fem.DoFMaps import P1_DoFMap  # noqa: E402
from PyNucleus_nl.kernels import getFractionalKernel  # noqa: E402
from PyNucleus_nl.nonlocalAssembly import nonlocalBuilder  # noqa: E402
from PyNucleus_nl.fractionalOrders import (leftRightFractionalOrder, layersFractionalOrder, innerOuterFractionalOrder,  # noqa: E402
                                           constFractionalOrder)


def mesh_arrays(mesh, dm):
    out = dict(vertices=np.array(mesh.vertices), cells=np.array(mesh.cells), dofs=np.array(dm.dofs), num_dofs=dm.num_dofs,
               hVector=np.array(mesh.hVector), volVector=np.array(mesh.volVector), h=mesh.h, hmin=mesh.hmin, diam=mesh.diam)
    if mesh.dim == 2:
        out['boundaryEdges'] = np.array(mesh.boundaryEdges)
    else:
        out['boundaryVertices'] = np.array(mesh.boundaryVertices)
    return out


def case(dim, noRef, sFun, name, params, extra):
    mesh = uniform_disc() if dim == 2 else simpleInterval(-1, 1)
    for _ in range(noRef):
        mesh = mesh.refine()
    dm = P1_DoFMap(mesh)
    kernel = getFractionalKernel(dim, sFun, np.inf)
    assert kernel.variable and kernel.piecewise
    out = mesh_arrays(mesh, dm)
    for ze, key in ((True, 'A'), (False, 'A_interior')):
        b = nonlocalBuilder(dm, kernel, dict(params), zeroExterior=ze)
        out[key] = np.array(b.getDense().data)
    out.update(symmetric=int(kernel.symmetric), local_matrix=type(b.local_matrix).__name__,
               target_order_used=b.local_matrix.target_order, quad_order_diagonal=b.local_matrix.quad_order_diagonal,
               **extra)
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, out['A'].shape, type(b.local_matrix).__name__, 'sym' if kernel.symmetric else 'nonsym',
          'asym of A: %.3e' % np.abs(out['A']-out['A'].T).max())


if __name__ == '__main__':
    which = sys.argv[1:] or ['all']
    if 'all' in which or 'leftright' in which:
        case(2, 2, leftRightFractionalOrder(0.25, 0.75, 0.6, 0.4, 0.), 'nonsym_disc_leftright_r2', {'target_order': 0.5},
             dict(kind='leftRight', sll=0.25, srr=0.75, slr=0.6, srl=0.4, interface=0.))
        case(2, 3, leftRightFractionalOrder(0.4, 0.8, 0.3, 0.7, 0.1), 'nonsym_disc_leftright_r3', {'target_order': 0.5},
             dict(kind='leftRight', sll=0.4, srr=0.8, slr=0.3, srl=0.7, interface=0.1))
        case(1, 5, leftRightFractionalOrder(0.25, 0.75, 0.6, 0.4, 0.), 'nonsym_interval_leftright_r5', {},
             dict(kind='leftRight', sll=0.25, srr=0.75, slr=0.6, srl=0.4, interface=0.))
    if 'all' in which or 'layers' in which:
        bnd = np.array([-1., -0.3, 0.35, 1.])
        orders = np.array([[0.3, 0.45, 0.5], [0.35, 0.6, 0.7], [0.5, 0.65, 0.8]])
        case(2, 3, layersFractionalOrder(2, bnd, orders), 'nonsym_disc_layers_r3', {'target_order': 0.5},
             dict(kind='layers', layerBoundaries=bnd, layerOrders=orders))
        orders_sym = 0.5*(orders+orders.T)
        case(2, 2, layersFractionalOrder(2, bnd, orders_sym), 'disc_layers_sym_r2', {'target_order': 0.5},
             dict(kind='layers', layerBoundaries=bnd, layerOrders=orders_sym))
    if 'all' in which or 'innerouter' in which:
        # 1D only: the reference stores center[0] alone (fractionalOrders.pyx:716) and reads center[1] from an
        # uninitialised parameter slot, so its 2D innerOuter order is not reproducible
        case(1, 5, innerOuterFractionalOrder(1, 0.3, 0.7, 0.5, np.array([0.1]), 0.55, 0.35), 'nonsym_interval_innerouter_r5',
             {}, dict(kind='innerOuter', sii=0.3, soo=0.7, r=0.5, center=np.array([0.1]), sio=0.55, soi=0.35))


def case_constant_nonsym(dim, noRef, s, name, params):
    """constantNonSymFractionalOrder: s(x,y) = const flagged as unsymmetric (fractionalOrders.pyx:631-638); the kernel cannot
    be piecewise (kernels.py:147-149), so the reference evaluates order and scaling per quadrature node
    (updateAndEvalFractional, kernelsCy.pyx:596-622) inside the unsymmetric local matrices"""
    from PyNucleus_nl.fractionalOrders import constantNonSymFractionalOrder
    import warnings
    mesh = uniform_disc() if dim == 2 else simpleInterval(-1, 1)
    for _ in range(noRef):
        mesh = mesh.refine()
    dm = P1_DoFMap(mesh)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        kernel = getFractionalKernel(dim, constantNonSymFractionalOrder(s), np.inf)
    assert kernel.variable and not kernel.piecewise and not kernel.symmetric
    out = mesh_arrays(mesh, dm)
    for ze, key in ((True, 'A'), (False, 'A_interior')):
        b = nonlocalBuilder(dm, kernel, dict(params), zeroExterior=ze)
        out[key] = np.array(b.getDense().data)
    out.update(symmetric=0, local_matrix=type(b.local_matrix).__name__, target_order_used=b.local_matrix.target_order,
               quad_order_diagonal=b.local_matrix.quad_order_diagonal, kind='constantNonSym', s=s)
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, out['A'].shape, type(b.local_matrix).__name__, 'asym of A: %.3e' % np.abs(out['A']-out['A'].T).max())


if __name__ == '__main__' and ('all' in (sys.argv[1:] or ['all']) or 'constant' in sys.argv[1:]):
    case_constant_nonsym(2, 2, 0.75, 'nonsym_disc_constant0.75_r2', {'target_order': 0.5})
    case_constant_nonsym(2, 3, 0.25, 'nonsym_disc_constant0.25_r3', {'target_order': 0.5})
    case_constant_nonsym(1, 5, 0.75, 'nonsym_interval_constant0.75_r5', {})
