import sys, os
sys.path.insert(0,'.')
import numpy as np, torch
import pynucleus_b200 as pb
sides, r = (10,6) if len(sys.argv)<2 else (int(sys.argv[1]), int(sys.argv[2]))
mesh = pb.refined(pb.polygon_disc(sides), r); dm = pb.P1_DoFMap(mesh)
N = dm.num_dofs
A = torch.empty((N, N), dtype=torch.float64, device='cuda')
b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2,0.75), {'target_order':0.5})
b.getDense(out=A); torch.cuda.synchronize()
print(b.getStats())
