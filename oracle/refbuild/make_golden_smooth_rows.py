"""ORACLE INFRASTRUCTURE: sampled rows of 2 977-DoF operators (disc, 5 refinements) for a tempered fractional kernel and a Gaussian
kernel on the full space, produced by running the REFERENCE ITSELF (stub-built copy in oracle/_ref):

    PYTHONPATH=oracle/_ref python oracle/refbuild/make_golden_smooth_rows.py

Rows, diagonal and the product A x of the reference's operator.  Every array is an output of reference code.
"""
import os
import sys
import time
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
OUT = os.path.join(HERE, '..', '..', 'tests', 'golden')

from PyNucleus_fem.mesh import uniform_disc  # noqa: E402
from PyNucleus_fem.DoFMaps import P1_DoFMap  # noqa: E402
from PyNucleus_nl.kernels import getFractionalKernel, getIntegrableKernel  # noqa: E402
from PyNucleus_nl.nonlocalAssembly import nonlocalBuilder  # noqa: E402
from make_golden_nonsym import mesh_arrays  # noqa: E402

mesh = uniform_disc()
for _ in range(5):
    mesh = mesh.refine()
dm = P1_DoFMap(mesh)
for name, kernel, extra in (('tempered_disc_s0.75_l2_r5_rows', getFractionalKernel(2, 0.75, np.inf, tempered=2.0), dict(s=0.75, tempered=2.0)),
                            ('gaussian_disc_v0.05_r5_rows', getIntegrableKernel(2, 'gaussian', np.inf, variance=0.05), dict(variance=0.05))):
    t = time.time()
    b = nonlocalBuilder(dm, kernel, {'target_order': 0.5}, zeroExterior=True)
    A = np.array(b.getDense().data)
    out = mesh_arrays(mesh, dm)
    rows = np.unique(np.concatenate((np.arange(0, dm.num_dofs, 97), [dm.num_dofs-1])))
    x = np.linspace(0., 1., dm.num_dofs)
    out.update(rows=rows, A_rows=A[rows], diag=np.diag(A).copy(), Ax=A.dot(x), x=x, scaling=kernel.scalingValue,
               quad_order_diagonal=b.local_matrix.quad_order_diagonal,
               bquad_order_diagonal=b.local_matrix_zeroExterior.quad_order_diagonal, **extra)
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, dm.num_dofs, '%.1f s' % (time.time()-t), flush=True)
