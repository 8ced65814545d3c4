"""driver-level known answers (reference tests/cache_runFractional.py--*): Hs / L2 errors through this package"""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from scipy.special import gamma
import pynucleus_b200 as pb
from test_gpu_parity import _p1_load_vector, _p1_mass

def run(domain, s, fmt, noRef, params, variable=False, solver='cg'):
    dim = 1 if domain == 'interval' else 2
    mesh = pb.refined(pb.simpleInterval(-1, 1) if dim == 1 else pb.uniform_disc(), noRef)
    dm = pb.P1_DoFMap(mesh)
    order = pb.variableConstFractionalOrder(s) if variable else s
    b_ = pb.nonlocalBuilder(dm, pb.getFractionalKernel(dim, order), params)
    A = b_.getH2() if fmt == 'H2' else b_.getDense()
    b = _p1_load_vector(mesh, dm)
    u, its, res = pb.cg(A, b, tol=1e-12, maxiter=3000)
    C = 2.**(-2.*s)*gamma(dim/2.)/gamma(dim/2.+s)/gamma(1.+s)
    if dim == 1:
        int_u = C*np.sqrt(np.pi)*gamma(s+1)/gamma(s+1.5)
        L2_ex2 = C**2*np.sqrt(np.pi)*gamma(2*s+1)/gamma(2*s+1.5)
        # Gauss1D(order=3): 2 Gauss-Legendre nodes (fem/PyNucleus_fem/femCy.pyx:2640, quadrature.pyx:303-316)
        t, w = np.polynomial.legendre.leggauss(2)
        rule = (np.stack(((t+1)/2, 1-(t+1)/2)), w/2)
    else:
        int_u = C*np.pi/(s+1)
        L2_ex2 = C**2*np.pi/(1+2*s)
        rule = (np.array([[0.5, 0.0, 0.5], [0.5, 0.5, 0.0], [0.0, 0.5, 0.5]]), np.full(3, 1./3.))
    Hs = np.sqrt(abs(b.dot(u)-int_u))
    def u_exact(x):
        return C*np.maximum(1.-(x**2).sum(axis=-1), 0.)**s
    out = [Hs]
    for r in [rule]:
        z = _p1_load_vector(mesh, dm, u_exact, rule=r)
        M = _p1_mass(mesh, dm)
        out.append(np.sqrt(abs(L2_ex2-2*z.dot(u)+u.dot(M.dot(u)))))
    return dm.num_dofs, its, out

for case, ref in [(('interval', 0.25, 'dense', 7, {}), (0.09611243700804001, 0.026655318974538753)),
                  (('interval', 0.75, 'dense', 7, {}), (0.04184296289342096, 0.0014584869810690354)),
                  (('interval', 0.25, 'H2', 7, {}), (0.0961124909768421, 0.026655322403497637)),
                  (('interval', 0.75, 'H2', 7, {}), (0.041849732677658555, 0.001458788789368659)),
                                    ]:
    print(case, run(*case), 'ref', ref, flush=True)
print(('interval varconst', run('interval', 0.75, 'dense', 7, {}, variable=True)), 'ref', (0.041842962898268554, 0.0014584869817160686))
