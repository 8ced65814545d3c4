import numpy as np, sys, time, os
sys.path.insert(0,'.')
from scipy.special import gamma
import oracle
import pynucleus_b200 as pb
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
from test_gpu_parity import _p1_load_vector, _p1_mass
s=0.75
mesh = pb.refined(pb.uniform_disc(), 5)
dm = pb.P1_DoFMap(mesh)
print(dm.num_dofs)
t=time.time()
P = oracle.disc_problem(5, s)
A = P.dense(True)
print('oracle', time.time()-t, A.shape)
b=_p1_load_vector(mesh, dm)
u=np.linalg.solve(A,b)
C = 2.**(-2.*s)*gamma(1.)/gamma(1.+s)/gamma(1.+s)
exactHs2 = C*np.pi/(s+1)
print('Hs', np.sqrt(abs(b.dot(u)-exactHs2)), 0.060319591944560894)
def u_exact(x):
    return C*np.maximum(1.-(x**2).sum(axis=-1), 0.)**s
midpoints = (np.array([[0.5, 0.0, 0.5], [0.5, 0.5, 0.0], [0.0, 0.5, 0.5]]), np.full(3, 1./3.))
for o in (2,):
    z=_p1_load_vector(mesh, dm, u_exact, rule=midpoints)
    M=_p1_mass(mesh, dm)
    print(o, np.sqrt(abs(C**2*np.pi/(1+2*s)-2*z.dot(u)+u.dot(M.dot(u)))), 0.002256341047519089)
print('L2ex2 %.17g zu %.17g uMu %.17g bu %.17g'%(C**2*np.pi/(1+2*s), z.dot(u), u.dot(M.dot(u)), b.dot(u)))
R=np.load('/tmp/ref_driver_out.npz')
print(np.abs(R['verts']-mesh.vertices).max(), (R['cells']==mesh.cells).all())
# map: reference u on full dm (3169) ; ours interior only
rd=R['dofs']; 
uu=np.zeros(dm.num_dofs); zz=np.zeros(dm.num_dofs)
m=dm.dofs>=0
uu[dm.dofs[m]]=R['u'][rd[m]]; zz[dm.dofs[m]]=R['z'][rd[m]]
print('u diff', np.abs(uu-u).max(), 'z diff', np.abs(zz-z).max(), np.abs(z).max())
r=A.dot(uu)-b
print('resid of ref u in oracle A', np.abs(r).max(), np.abs(b).max(), 'argmax', np.argmax(np.abs(uu-u)), np.sort(np.abs(uu-u))[-5:])
Ar=np.load('/tmp/ref_driver_A.npy')
print(Ar.shape)
# dof map of reference interior dm? assume same numbering
if Ar.shape==A.shape:
    print('A diff', np.abs(Ar-A).max(), np.abs(A).max(), np.unravel_index(np.argmax(np.abs(Ar-A)),A.shape))
    ur=np.linalg.solve(Ar,b); print('u diff with ref A', np.abs(ur-u).max(), np.abs(ur-uu).max())
