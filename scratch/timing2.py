import sys, os, time
sys.path.insert(0,'.')
import numpy as np, torch
import pynucleus_b200._lib as L
if len(sys.argv) > 1:
    L.LIB_PATH = sys.argv[1]
import pynucleus_b200 as pb
mesh = pb.refined(pb.polygon_disc(10), 6); dm = pb.P1_DoFMap(mesh)
b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2,0.75), {'target_order':0.5})
A = torch.empty((dm.num_dofs, dm.num_dofs), dtype=torch.float64, device='cuda')
for rep in range(2):
    b.getDense(out=A); torch.cuda.synchronize()
st = b.getStats()
print(os.path.basename(L.LIB_PATH), os.environ.get('PNB_DEBUG'), 'tiles %.1f ms total %.1f'%(st['ms_tiles'], st['ms_total']), float(A[0,0]), float(A[100,5000]))
