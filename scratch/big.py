import sys, os, time
sys.path.insert(0,'.')
import numpy as np, torch
import pynucleus_b200 as pb
sides, r = int(sys.argv[1]), int(sys.argv[2])
t0=time.time()
mesh = pb.refined(pb.polygon_disc(sides), r); dm = pb.P1_DoFMap(mesh)
N = dm.num_dofs
print('mesh', mesh.num_cells, N, 'host s %.1f'%(time.time()-t0), flush=True)
A = torch.empty((N, N), dtype=torch.float64, device='cuda')
b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2,0.75), {'target_order':0.5})
for rep in range(2):
    t0=time.time(); b.getDense(out=A); torch.cuda.synchronize(); print('wall %.2f s'%(time.time()-t0), b.getStats(), flush=True)
print('mem GB', torch.cuda.max_memory_allocated()/1e9, 'free/total', [x/1e9 for x in torch.cuda.mem_get_info()])
# properties: symmetry on a sample, row sums positive, matvec
idx = torch.randint(0, N, (2000,), device='cuda')
sub = A[idx][:, idx]
print('sym sample', bool(torch.equal(sub, sub.T)), 'diag>0', bool((torch.diagonal(A) > 0).all()))
x = torch.ones(N, dtype=torch.float64, device='cuda')
y = pb.Dense_LinearOperator(A).matvec_device(x)
print('min rowsum', float(y.min()), 'max', float(y.max()))
