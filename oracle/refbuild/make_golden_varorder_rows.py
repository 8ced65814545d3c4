"""ORACLE INFRASTRUCTURE: sampled rows of larger 2D operators for an order that varies inside the cells
(smoothedLeftRightFractionalOrder(0.25, 0.75)), produced by running the REFERENCE ITSELF (stub-built copy in oracle/_ref):

    PYTHONPATH=oracle/_ref python oracle/refbuild/make_golden_varorder_rows.py 4     # 721 DoFs, 37 s
    PYTHONPATH=oracle/_ref python oracle/refbuild/make_golden_varorder_rows.py 5     # 2 977 DoFs, ~10 min

Rows, diagonal and products A x, A^T x of the reference's operator.  Every array is an output of reference code.
"""
import numpy as np, time, warnings, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from PyNucleus_fem.mesh import uniform_disc
from PyNucleus_fem.DoFMaps import P1_DoFMap
from PyNucleus_nl.kernels import getFractionalKernel
from PyNucleus_nl.nonlocalAssembly import nonlocalBuilder
from PyNucleus_nl.fractionalOrders import smoothedLeftRightFractionalOrder
from make_golden_nonsym import mesh_arrays
mesh = uniform_disc()
NR = int(sys.argv[1])
for _ in range(NR): mesh = mesh.refine()
dm = P1_DoFMap(mesh)
with warnings.catch_warnings():
    warnings.simplefilter('ignore')
    kernel = getFractionalKernel(2, smoothedLeftRightFractionalOrder(0.25, 0.75), np.inf)
t=time.time()
b = nonlocalBuilder(dm, kernel, {'target_order': 0.5}, zeroExterior=True)
A = np.array(b.getDense().data)
print(dm.num_dofs, time.time()-t, flush=True)
out = mesh_arrays(mesh, dm)
rows = np.unique(np.concatenate((np.arange(0, dm.num_dofs, 23 if NR == 4 else 97), [dm.num_dofs-1])))
x = np.linspace(0., 1., dm.num_dofs)
out.update(rows=rows, A_rows=A[rows], diag=np.diag(A).copy(), Ax=A.dot(x), ATx=A.T.dot(x), x=x, kind='smoothedLeftRight', sl=0.25, sr=0.75, r=0.1, interface=0.,
           quad_order_diagonal=b.local_matrix.quad_order_diagonal, bquad_order_diagonal=b.local_matrix_zeroExterior.quad_order_diagonal)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', 'tests', 'golden', 'varorder_disc_smoothed_r%d_rows' % NR), **out)
print('saved')
