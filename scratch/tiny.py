import sys, time
sys.path.insert(0,'.')
import numpy as np, torch
import pynucleus_b200 as pb
from pynucleus_b200 import _lib
from pynucleus_b200.assembly import _Problem
mesh = pb.refined(pb.uniform_disc(), 2); dm = pb.P1_DoFMap(mesh)
b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2,0.75), {'target_order':0.5})
b.getDense()
N = dm.num_dofs
A = torch.empty((N,N), dtype=torch.float64, device='cuda')
t0=time.perf_counter()
for k in range(100):
    prob = _Problem(dm, b.kernel, b.kernelBoundary, b.orders, 0, b.problem.max_order)
t1=time.perf_counter()
for k in range(100):
    _lib.check(_lib.lib().pnb_dense_assemble(prob.handle, 1, 0, N, A.data_ptr(), N, 1))
torch.cuda.synchronize(); t2=time.perf_counter()
probs=[]
for k in range(100):
    prob = _Problem(dm, b.kernel, b.kernelBoundary, b.orders, 0, b.problem.max_order)
    _lib.check(_lib.lib().pnb_dense_assemble(prob.handle, 1, 0, N, A.data_ptr(), N, 1))
torch.cuda.synchronize(); t3=time.perf_counter()
print('N', N, 'create %.2f ms, assemble (schedule cached) %.2f ms, create+first assemble %.2f ms' % ((t1-t0)*10, (t2-t1)*10, (t3-t2)*10))
