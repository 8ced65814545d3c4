"""scratch: device time of the row-owner kernel for orders that vary inside a cell (gpurun)"""
import os
import sys
import time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import pynucleus_b200 as pb

for dim, noRef in ((1, 10), (2, 5), (2, 6)):
    mesh = pb.refined(pb.simpleInterval(-1, 1) if dim == 1 else pb.uniform_disc(), noRef)
    dm = pb.P1_DoFMap(mesh)
    b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(dim, pb.smoothedLeftRightFractionalOrder(0.25, 0.75)),
                           {'target_order': 0.5} if dim == 2 else {})
    t0 = time.time()
    b.getDense()
    torch.cuda.synchronize()
    t1 = time.time()
    b.getDense()
    torch.cuda.synchronize()
    t2 = time.time()
    print('varorder dim', dim, 'dofs', dm.num_dofs, 'values', b._varorder['values'].shape[0], 'first %.3f s' % (t1-t0), 'second %.3f s' % (t2-t1), flush=True)
