import sys, os, time
sys.path.insert(0,'.')
import numpy as np, torch
import pynucleus_b200 as pb
sides, r = (10,6) if len(sys.argv)<2 else (int(sys.argv[1]), int(sys.argv[2]))
mesh = pb.refined(pb.polygon_disc(sides), r); dm = pb.P1_DoFMap(mesh)
N = dm.num_dofs
A = torch.empty((N, N), dtype=torch.float64, device='cuda')
b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2,0.75), {'target_order':0.5})
for rep in range(3):
    b.getDense(out=A); torch.cuda.synchronize()
st = b.getStats()
print('tiles %.1f ms  f2 %.1f  near %.1f  mix %.1f  sym %.1f  pairs %d' % (st['ms_tiles'], st['ms_f2'], st['ms_near'], st['ms_mix'], st['ms_symmetrize'], st['evaluated_pairs']))
