"""ORACLE INFRASTRUCTURE: BASELINE configs[2] (drivers/runNonlocal.py --domain square --kernelType constant|fractional
--problem poly-Dirichlet --solver cg-mg --matrixFormat dense|H2) run through the REFERENCE's own driver classes
(nonlocalPoissonProblem, discretizedNonlocalProblem; the stub-built copy in oracle/_ref).

    python oracle/refbuild/make_golden_nonlocal_driver.py constant

One deviation, forced by the image: the reference meshes the square and its interaction collar with meshpy (Triangle),
which is not installed; its own structured variant of the same constructor (squareWithInteractions(..., uniform=True),
fem/PyNucleus_fem/mesh.py:441-460) is selected instead.  Everything else is the unmodified driver path.  Output:
tests/golden/nonlocal_square_<kernel>.npz with the finest mesh, both DoFMaps, sampled rows / products of the operator and of
the Dirichlet coupling block, right-hand side, solution and the errors the driver reports.
"""
import os
import sys
import types
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..', '..'))
OUT = os.path.join(ROOT, 'tests', 'golden')
m = types.ModuleType('PyNucleus')
m.subpackages = {}
sys.modules['PyNucleus'] = m
sys.path.insert(0, os.path.join(ROOT, 'oracle', '_ref'))
from PyNucleus_base import driver  # noqa: E402
from PyNucleus_nl.nonlocalProblems import nonlocalPoissonProblem  # noqa: E402
from PyNucleus_nl import nonlocalProblems as NP  # noqa: E402
from PyNucleus_nl.discretizedProblems import discretizedNonlocalProblem  # noqa: E402
import PyNucleus_base.utilsFem as _uf  # noqa: E402
_uf.getSystemInfo = lambda *a, **k: None

f = NP.nonlocalMeshFactory.overlappingMeshFactory
nm, ct, prm = f.classes['square']
prm = dict(prm)
prm['uniform'] = True
f.classes['square'] = (nm, ct, prm)

KT = sys.argv[1] if len(sys.argv) > 1 else 'constant'
MF = sys.argv[2] if len(sys.argv) > 2 else 'dense'
d = driver()
p = nonlocalPoissonProblem(d)
dp = discretizedNonlocalProblem(d, p)
d.process(override={'domain': 'square', 'kernelType': KT, 'problem': 'poly-Dirichlet', 'solver': 'cg-mg', 'matrixFormat': MF})
sol = dp.modelSolution
H = dp.hierarchy
A = H[-1]['A']
print(type(A), A.shape, 'levels', len(H))
Ad = np.array(A.toarray())
ABC = np.array(dp.A_BC.toarray())
mesh = dp.finalMesh
dmI = dp.dmInterior
dmBC = dp.dmBC
rng = np.random.default_rng(3)
rows = np.sort(rng.choice(Ad.shape[0], 16, replace=False))
x = rng.standard_normal(Ad.shape[1])
xb = rng.standard_normal(ABC.shape[1])
kernel = p.kernel
out = dict(vertices=np.array(mesh.vertices), cells=np.array(mesh.cells), boundaryEdges=np.array(mesh.boundaryEdges),
           dofs=np.array(dmI.dofs), num_dofs=dmI.num_dofs, dofsBC=np.array(dmBC.dofs), num_dofsBC=dmBC.num_dofs,
           coarse_vertices=np.array(H[0]['mesh'].vertices) if 'mesh' in H[0] else np.zeros((0, 2)),
           coarse_cells=np.array(H[0]['mesh'].cells) if 'mesh' in H[0] else np.zeros((0, 3), dtype=np.int32),
           level_sizes=np.array([h['A'].shape[0] for h in H]),
           kernel_type=KT, matrixFormat=MF, operator_type=type(A).__name__,
           horizon=kernel.horizonValue, scaling=kernel.scalingValue, singularity=kernel.singularityValue,
           s=getattr(kernel, 'sValue', 0.), target_order=p.target_order, eta=p.eta,
           rows=rows, A_rows=Ad[rows], diagonal=np.diag(Ad).copy(), x=x, Ax=Ad.dot(x), frobenius=np.linalg.norm(Ad),
           xb=xb, ABCxb=ABC.dot(xb), ABC_rows=ABC[rows], ABC_frobenius=np.linalg.norm(ABC),
           b=np.array(sol.b) if hasattr(sol, 'b') else np.array(dp.b), u=np.array(sol.u),
           uD=np.array(dmBC.interpolate(p.dirichletData)),
           u_interp=np.array(sol.u_interp) if sol.u_interp is not None else np.zeros(0),
           L2_error=sol.L2_error if sol.L2_error is not None else np.nan,
           rel_L2_error=sol.rel_L2_error if sol.rel_L2_error is not None else np.nan,
           uI=np.array(sol.uRestricted), iterations=getattr(sol, 'iterations', -1),
           residuals=np.array(getattr(dp.solver, 'residuals', [])), tol=getattr(dp.solver, 'tolerance', np.nan), hmin=mesh.hmin, h=mesh.h, diam=mesh.diam,
           hVector=np.array(mesh.hVector), volVector=np.array(mesh.volVector))
name = 'nonlocal_square_{}{}.npz'.format(KT, '' if MF == 'dense' else '_'+MF)
np.savez_compressed(os.path.join(OUT, name), **out)
print(name, 'N', dmI.num_dofs, 'NBC', dmBC.num_dofs, 'L2', sol.L2_error, 'rel', sol.rel_L2_error, 'its', out['iterations'],
      'kernel', kernel)
