import sys, os
sys.path.insert(0,'.')
import numpy as np, torch
import pynucleus_b200 as pb
mesh = pb.refined(pb.polygon_disc(10), 6); dm = pb.P1_DoFMap(mesh)
N = dm.num_dofs
A = torch.empty((N, N), dtype=torch.float64, device='cuda')
for kern, dbg in (('f2', 0x300), ('mix', 0x1100)):
    for ab in (0, 1, 2, 3, 4, 7):
        os.environ['PNB_DEBUG'] = str(dbg); os.environ['PNB_ABLATE'] = str(ab)
        b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2,0.75), {'target_order':0.5})
        for rep in range(2):
            b.getDense(out=A); torch.cuda.synchronize()
        print('%s ablate %d: %.1f ms' % (kern, ab, b.getStats()['ms_tiles']))
