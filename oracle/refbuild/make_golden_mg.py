"""ORACLE INFRASTRUCTURE: multigrid goldens from the REFERENCE ITSELF (stub-built copy in oracle/_ref).

    python oracle/refbuild/make_golden_mg.py

Runs the reference's driver path of BASELINE config 3 (runFractional: disc, s = varconst(0.75), P1, dense, solver
cg-mg; nl/PyNucleus_nl/discretizedProblems.py:560-650) at small refinement counts and stores the hierarchy's
restriction operators, the right-hand side, the solution, and the iteration counts / residual histories of cg-mg
and of the multigrid iteration alone (multilevelSolver/PyNucleus_multilevelSolver/multigrid_{SCALAR}.pxi).
"""
import os
import sys
import types
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..', '..'))
sys.path.insert(0, os.path.join(ROOT, 'oracle', '_ref'))
m = types.ModuleType('PyNucleus')
m.subpackages = {}
sys.modules['PyNucleus'] = m
from PyNucleus_base import driver, solverFactory  # noqa: E402
from PyNucleus_nl.nonlocalProblems import fractionalLaplacianProblem  # noqa: E402
from PyNucleus_nl.discretizedProblems import discretizedNonlocalProblem  # noqa: E402
import PyNucleus_base.utilsFem as _uf  # noqa: E402
_uf.getSystemInfo = lambda *a, **k: None


def case(noRef, name):
    d = driver()
    p = fractionalLaplacianProblem(d, False)
    dp = discretizedNonlocalProblem(d, p)
    d.process(override={'domain': 'disc', 'kernel': 'fractional', 's': 'varconst(0.75)', 'problem': 'constant',
                        'element': 'P1', 'solver': 'cg-mg', 'matrixFormat': 'dense', 'noRef': noRef, 'tol': 1e-8})
    sol = dp.modelSolution
    H = dp.hierarchy
    out = dict(noRef=noRef, num_levels=len(H), s=0.75, tol=float(sol.tol), b=np.array(sol.b), u=np.array(sol.uInterior),
               cgmg_iterations=int(sol.iterations), cgmg_residuals=np.array(sol.residuals),
               Hs_error=float(sol.Hs_error), L2_error=float(sol.L2_error),
               level_num_dofs=np.array([lvl['A'].shape[0] for lvl in H]))
    for k, lvl in enumerate(H):
        out['diag%d' % k] = np.array(lvl['A'].diagonal)
        if 'R' in lvl:
            R = lvl['R']
            out['R%d_indptr' % k] = np.array(R.indptr)
            out['R%d_indices' % k] = np.array(R.indices)
            out['R%d_data' % k] = np.array(R.data)
    # the multigrid iteration alone and preconditioned GMRES on the same hierarchy
    b = np.array(sol.b)
    mg = solverFactory.build('mg', hierarchy=H, setup=True)
    mg.tolerance = 1e-8
    mg.maxIter = 60
    x = np.zeros_like(b)
    its = mg(b, x)
    out.update(mg_iterations=its, mg_residuals=np.array(mg.residuals), mg_x=x.copy())
    cg = solverFactory.build('cg-mg', hierarchy=H, setup=True)
    cg.tolerance = 1e-8
    cg.maxIter = 100
    x = np.zeros_like(b)
    its = cg(b, x)
    out.update(cg2_iterations=its, cg2_residuals=np.array(cg.residuals), cg2_x=x.copy())
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', name), **out)
    print(name, out['level_num_dofs'], 'cg-mg', sol.iterations, len(sol.residuals), 'mg', out['mg_iterations'],
          'cg2', its, 'Hs', sol.Hs_error)


if __name__ == '__main__':
    case(3, 'mg_disc_varconst0.75_r3')
    case(4, 'mg_disc_varconst0.75_r4')
