"""ORACLE INFRASTRUCTURE: fixtures for TEMPERED fractional kernels gamma(x,y) = C |x-y|^(-d-2s) exp(-lambda |x-y|)
(temperedFracKernelInfinite*, kernelsCy.pyx:186-213; scaling kernelNormalization.pyx:84-88), produced by running the
REFERENCE ITSELF (stub-built copy in oracle/_ref, see build_reference.sh):

    PYTHONPATH=oracle/_ref python oracle/refbuild/make_golden_tempered.py

Note that FractionalKernel.getBoundaryKernel (kernelsCy.pyx:1982-2027) does not hand `tempered` on: the surface terms of the
zero-exterior operator use the UNtempered power law with the tempered scaling constant (times 1/s).  Every array is an
output of reference code.
"""
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..', '..'))
OUT = os.path.join(ROOT, 'tests', 'golden')

from PyNucleus_fem.mesh import simpleInterval, uniform_disc  # noqa: E402
from PyNucleus_fem.DoFMaps import P1_DoFMap, P2_DoFMap  # noqa: E402
from PyNucleus_nl.kernels import getFractionalKernel  # noqa: E402
from PyNucleus_nl.nonlocalAssembly import nonlocalBuilder  # noqa: E402
from make_golden_nonsym import mesh_arrays  # noqa: E402


def case(dim, noRef, s, lam, name, params, element='P1'):
    mesh = uniform_disc() if dim == 2 else simpleInterval(-1, 1)
    for _ in range(noRef):
        mesh = mesh.refine()
    dm = (P1_DoFMap if element == 'P1' else P2_DoFMap)(mesh)
    kernel = getFractionalKernel(dim, s, np.inf, tempered=lam)
    assert kernel.symmetric and not kernel.variable
    out = mesh_arrays(mesh, dm)
    for ze, key in ((True, 'A'), (False, 'A_interior')):
        b = nonlocalBuilder(dm, kernel, dict(params), zeroExterior=ze)
        out[key] = np.array(b.getDense().data)
    rng = np.random.default_rng(7)
    X = rng.uniform(-1, 1, size=(32, dim))*0.7
    Y = rng.uniform(-1, 1, size=(32, dim))*0.7
    kb = kernel.getBoundaryKernel()
    out.update(points_x=X, points_y=Y, kernel_values=np.array([kernel(X[i], Y[i]) for i in range(32)]),
               bkernel_values=np.array([kb(X[i], Y[i]) for i in range(32)]), scaling=kernel.scalingValue,
               bscaling=kb.scalingValue, s=s, tempered=lam, element=element, local_matrix=type(b.local_matrix).__name__,
               target_order_used=b.local_matrix.target_order, quad_order_diagonal=b.local_matrix.quad_order_diagonal)
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, out['A'].shape, type(b.local_matrix).__name__, 'asym of A: %.3e' % np.abs(out['A']-out['A'].T).max(), flush=True)


if __name__ == '__main__':
    case(1, 5, 0.75, 2.0, 'tempered_interval_s0.75_l2_r5', {})
    case(1, 6, 0.25, 0.5, 'tempered_interval_s0.25_l0.5_r6', {})
    case(2, 2, 0.75, 2.0, 'tempered_disc_s0.75_l2_r2', {'target_order': 0.5})
    case(2, 3, 0.25, 1.0, 'tempered_disc_s0.25_l1_r3', {'target_order': 0.5})
    case(1, 4, 0.75, 1.5, 'tempered_p2_interval_s0.75_l1.5_r4', {}, element='P2')
    case(2, 1, 0.75, 1.5, 'tempered_p2_disc_s0.75_l1.5_r1', {'target_order': 0.5}, element='P2')
