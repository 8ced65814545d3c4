"""ORACLE INFRASTRUCTURE: fixtures for the integrable kernels with INFINITE horizon that the reference's driver tests run
(tests/test_drivers_intFracLapl.py:42-43: `--kernelType gaussian --gaussianVariance 0.1 --interaction fullSpace` and
`--kernelType exponential --exponentialRate 8`): gaussianKernel* / exponentialKernel (kernelsCy.pyx:388-477) with
constantIntegrableScaling (kernelNormalization.pyx:257-283), produced by running the REFERENCE ITSELF (stub-built copy in
oracle/_ref, see build_reference.sh):

    PYTHONPATH=oracle/_ref python oracle/refbuild/make_golden_smooth.py

Every array is an output of reference code.
"""
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..', '..'))
OUT = os.path.join(ROOT, 'tests', 'golden')

from PyNucleus_fem.mesh import simpleInterval, uniform_disc  # noqa: E402
from PyNucleus_fem.DoFMaps import P1_DoFMap, P2_DoFMap  # noqa: E402
from PyNucleus_nl.kernels import getIntegrableKernel  # noqa: E402
from PyNucleus_nl.nonlocalAssembly import nonlocalBuilder  # noqa: E402
from make_golden_nonsym import mesh_arrays  # noqa: E402


def case(dim, noRef, ktype, kw, name, params, element='P1'):
    mesh = uniform_disc() if dim == 2 else simpleInterval(-1, 1)
    for _ in range(noRef):
        mesh = mesh.refine()
    dm = (P1_DoFMap if element == 'P1' else P2_DoFMap)(mesh)
    kernel = getIntegrableKernel(dim, ktype, np.inf, **kw)
    assert kernel.symmetric and not kernel.variable and not kernel.finiteHorizon
    out = mesh_arrays(mesh, dm)
    for ze, key in ((True, 'A'), (False, 'A_interior')):
        b = nonlocalBuilder(dm, kernel, dict(params), zeroExterior=ze)
        out[key] = np.array(b.getDense().data)
    rng = np.random.default_rng(9)
    X = rng.uniform(-1, 1, size=(32, dim))*0.7
    Y = rng.uniform(-1, 1, size=(32, dim))*0.7
    kb = kernel.getBoundaryKernel()
    out.update(points_x=X, points_y=Y, kernel_values=np.array([kernel(X[i], Y[i]) for i in range(32)]),
               bkernel_values=np.array([kb(X[i], Y[i]) for i in range(32)]), scaling=kernel.scalingValue,
               bscaling=kb.scalingValue, kernelType=ktype, element=element, **kw,
               target_order_used=b.local_matrix.target_order, quad_order_diagonal=b.local_matrix.quad_order_diagonal,
               btarget_order_used=b.local_matrix_zeroExterior.target_order,
               bquad_order_diagonal=b.local_matrix_zeroExterior.quad_order_diagonal)
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, out['A'].shape, 'asym of A: %.3e' % np.abs(out['A']-out['A'].T).max(), flush=True)


if __name__ == '__main__':
    case(1, 5, 'gaussian', dict(variance=0.1), 'gaussian_interval_v0.1_r5', {})
    case(1, 6, 'gaussian', dict(variance=0.02), 'gaussian_interval_v0.02_r6', {})
    case(1, 5, 'exponential', dict(exponentialRate=8.0), 'exponential_interval_a8_r5', {})
    case(1, 6, 'exponential', dict(exponentialRate=2.5), 'exponential_interval_a2.5_r6', {})
    case(2, 2, 'gaussian', dict(variance=0.1), 'gaussian_disc_v0.1_r2', {'target_order': 0.5})
    case(2, 3, 'gaussian', dict(variance=0.3), 'gaussian_disc_v0.3_r3', {'target_order': 0.5})
    case(1, 4, 'gaussian', dict(variance=0.1), 'gaussian_p2_interval_v0.1_r4', {}, element='P2')
