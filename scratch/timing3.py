import sys, os, time
sys.path.insert(0,'.')
import numpy as np, torch
import pynucleus_b200 as pb
sides, r = (10,6) if len(sys.argv)<2 else (int(sys.argv[1]), int(sys.argv[2]))
mesh = pb.refined(pb.polygon_disc(sides), r); dm = pb.P1_DoFMap(mesh)
N = dm.num_dofs
A = torch.empty((N, N), dtype=torch.float64, device='cuda')
for name, dbg in (('all', 0), ('only near', 0x1200), ('only mix', 0x1100), ('only f2', 0x300), ('none', 0x1300)):
    os.environ['PNB_DEBUG'] = str(dbg)
    b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2,0.75), {'target_order':0.5})
    for rep in range(2):
        b.getDense(out=A); torch.cuda.synchronize()
    st = b.getStats()
    print('%-10s tiles %.1f ms  pairs %d' % (name, st['ms_tiles'], st['evaluated_pairs']))
