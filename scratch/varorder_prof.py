"""scratch: one assembly with an order that varies inside the cells, for ncu (gpurun)"""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import pynucleus_b200 as pb

mesh = pb.refined(pb.uniform_disc(), int(sys.argv[1]) if len(sys.argv) > 1 else 4)
dm = pb.P1_DoFMap(mesh)
b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, pb.smoothedLeftRightFractionalOrder(0.25, 0.75)), {'target_order': 0.5})
b.getDense()
torch.cuda.synchronize()
print('dofs', dm.num_dofs)
