import sys, time
sys.path.insert(0,'.')
import numpy as np, torch
import pynucleus_b200 as pb
r = int(sys.argv[1])
mesh = pb.refined(pb.uniform_disc(), r); dm = pb.P1_DoFMap(mesh)
b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2,0.75), {'target_order':0.5})
t0=time.time(); A = b.getDense(); torch.cuda.synchronize(); t1=time.time()
import cProfile, pstats
pr=cProfile.Profile(); pr.enable()
H, Pnear = b.getH2(returnNearField=True); torch.cuda.synchronize()
pr.disable(); t2=time.time()
print('N', dm.num_dofs, 'dense %.2f s, H2 %.2f s'%(t1-t0,t2-t1), H, 'near pairs', len(Pnear), 'near nnz frac %.3f'%(H.Anear.nnz/dm.num_dofs**2))
x = torch.randn(dm.num_dofs, dtype=torch.float64, device='cuda')
y=H.matvec_device(x); torch.cuda.synchronize(); t=time.time(); y=H.matvec_device(x); torch.cuda.synchronize(); print('H2 matvec (2nd call) %.2f ms'%((time.time()-t)*1e3), 'rel err vs dense %.2e'%float((y-A.matvec_device(x)).abs().max()/A.matvec_device(x).abs().max()))
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
