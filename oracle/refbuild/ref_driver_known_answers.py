import sys, os, types, numpy as np
SOLVER, TOL = os.environ.get('SOLVER','lu'), float(os.environ.get('TOL','1e-6'))
m = types.ModuleType('PyNucleus'); m.subpackages = {}; sys.modules['PyNucleus'] = m
sys.path.insert(0, 'oracle/_ref')
from PyNucleus_base import driver, solverFactory
from PyNucleus_nl.nonlocalProblems import fractionalLaplacianProblem
from PyNucleus_nl.discretizedProblems import discretizedNonlocalProblem
import PyNucleus_base.utilsFem as _uf
_uf.getSystemInfo = lambda *a, **k: None
d = driver()
p = fractionalLaplacianProblem(d, False)
dp = discretizedNonlocalProblem(d, p)
d.process(override={'domain':'disc','kernel':'fractional','s':'const(0.75)','problem':'constant','element':'P1','solver':SOLVER,'matrixFormat':'dense','noRef':5, 'tol': TOL})
sol = dp.modelSolution
print('N', dp.finalMesh.num_vertices, sol.u.shape)
print('Hs', sol.Hs_error, 'L2', sol.L2_error)
u = sol.u
dmm = u.dm
M = dp.mass if u.dm == dp.dm else dp.massInterior
z = dmm.assembleRHS(p.analyticSolution)
print('exactL2Squared %.17g zu %.17g uMu %.17g' % (p.exactL2Squared, z.inner(u), u.inner(M*u)))
print('bu %.17g' % sol.b.inner(sol.uRestricted) if hasattr(sol,'b') else '')
np.savez('/tmp/ref_driver_out.npz', u=np.array(u), z=np.array(z), verts=np.array(dp.finalMesh.vertices_as_array), cells=np.array(dp.finalMesh.cells_as_array), dofs=np.array(dmm.dofs))
H = dp.hierarchy
print(type(H), len(H), list(H[-1].keys()))
A = H[-1]['A']
print(type(A), A.shape)
np.save('/tmp/ref_driver_A.npy', np.array(A.toarray()))
print('params', p.target_order, p.eta, dp.zeroExterior if hasattr(dp,'zeroExterior') else None)
