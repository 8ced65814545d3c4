"""ORACLE INFRASTRUCTURE: fixtures for P2 elements, produced by running the REFERENCE ITSELF (stub-built copy in
oracle/_ref, see build_reference.sh):

    PYTHONPATH=oracle/_ref python oracle/refbuild/make_golden_p2.py

P2_DoFMap (fem/PyNucleus_fem/DoFMaps.pyx:1978-2031) with the reference's nonlocalBuilder.getDense
(nonlocalAssembly_{SCALAR}.pxi:1262-1473; 78 local entries per triangle pair).  Every array is an output of reference code.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..', '..'))
OUT = os.path.join(ROOT, 'tests', 'golden')

from PyNucleus_fem.mesh import simpleInterval, uniform_disc  # noqa: E402
from PyNucleus_fem.DoFMaps import P2_DoFMap, P0_DoFMap, P3_DoFMap  # noqa: E402
from PyNucleus_nl.kernels import getFractionalKernel  # noqa: E402
from PyNucleus_nl.nonlocalAssembly import nonlocalBuilder  # noqa: E402
from PyNucleus_nl.fractionalOrders import constFractionalOrder  # noqa: E402


def case(dim, noRef, s, name, params, DoFMap=None):
    mesh = uniform_disc() if dim == 2 else simpleInterval(-1, 1)
    for _ in range(noRef):
        mesh = mesh.refine()
    dm = (P2_DoFMap if DoFMap is None else DoFMap)(mesh)
    kernel = getFractionalKernel(dim, constFractionalOrder(s), np.inf)
    out = dict(vertices=np.array(mesh.vertices), cells=np.array(mesh.cells), dofs=np.array(dm.dofs), num_dofs=dm.num_dofs,
               num_boundary_dofs=dm.num_boundary_dofs, hVector=np.array(mesh.hVector), volVector=np.array(mesh.volVector),
               h=mesh.h, hmin=mesh.hmin, diam=mesh.diam, s=s)
    if dim == 2:
        out['boundaryEdges'] = np.array(mesh.boundaryEdges)
    out['boundaryVertices'] = np.array(mesh.boundaryVertices)
    for ze, key in ((True, 'A'), (False, 'A_interior')):
        b = nonlocalBuilder(dm, kernel, dict(params), zeroExterior=ze)
        out[key] = np.array(b.getDense().data)
    out.update(target_order_used=b.local_matrix.target_order, quad_order_diagonal=b.local_matrix.quad_order_diagonal,
               boundary_target_order=b.local_matrix_zeroExterior.target_order,
               boundary_quad_order_diagonal=b.local_matrix_zeroExterior.quad_order_diagonal)
    shape = out['A'].shape
    print(name, shape, 'asym %.2e' % np.abs(out['A']-out['A'].T).max())
    if shape[0] > 400:
        # every 8th row and the diagonal only (fixture size)
        rows = np.arange(0, shape[0], 8)
        out.update(rows=rows, A_rows=out['A'][rows], A_interior_rows=out['A_interior'][rows], A_diagonal=np.diag(out['A']).copy())
        del out['A'], out['A_interior']
    np.savez_compressed(os.path.join(OUT, name), **out)
    return
    print(name, out['A'].shape, 'asym %.2e' % np.abs(out['A']-out['A'].T).max())


if __name__ == '__main__':
    which = sys.argv[1:] or ['all']
    if 'all' in which or 'interval' in which:
        case(1, 4, 0.25, 'p2_interval_s0.25_r4', {})
        case(1, 5, 0.75, 'p2_interval_s0.75_r5', {})
    if 'all' in which or 'disc' in which:
        case(2, 1, 0.75, 'p2_disc_s0.75_r1', {'target_order': 0.5})
        case(2, 2, 0.75, 'p2_disc_s0.75_r2', {'target_order': 0.5})
        case(2, 2, 0.25, 'p2_disc_s0.25_r2', {'target_order': 0.5})
        case(2, 3, 0.75, 'p2_disc_s0.75_r3', {'target_order': 0.5})
    if 'all' in which or 'p0' in which:
        # piecewise constants: s < 1/2 only (fractionalLaplacian2D.pyx:596-598)
        case(1, 5, 0.25, 'p0_interval_s0.25_r5', {}, DoFMap=P0_DoFMap)
        case(2, 2, 0.25, 'p0_disc_s0.25_r2', {'target_order': 0.5}, DoFMap=P0_DoFMap)
        case(2, 3, 0.4, 'p0_disc_s0.4_r3', {'target_order': 0.5}, DoFMap=P0_DoFMap)
    if 'all' in which or 'p3' in which:
        case(1, 4, 0.25, 'p3_interval_s0.25_r4', {}, DoFMap=P3_DoFMap)
        case(1, 4, 0.75, 'p3_interval_s0.75_r4', {}, DoFMap=P3_DoFMap)
    if 'all' in which or 'tiny' in which:
        # smallest meshes: one hexagon (6 cells), two intervals
        case(2, 0, 0.75, 'p2_disc_s0.75_r0', {'target_order': 0.5})
        case(2, 0, 0.25, 'p0_disc_s0.25_r0', {'target_order': 0.5}, DoFMap=P0_DoFMap)
        case(1, 1, 0.75, 'p2_interval_s0.75_r1', {})
        case(1, 1, 0.25, 'p3_interval_s0.25_r1', {}, DoFMap=P3_DoFMap)
