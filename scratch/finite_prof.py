import sys
sys.path.insert(0, '.')
import torch
import pynucleus_b200 as pb
mesh = pb.refined(pb.uniform_disc(), 6); dm = pb.P1_DoFMap(mesh)
b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, 0.4, 0.15), {'target_order': 0.5})
b.getDense(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
b.getDense(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(b.getStats())
print(b.getPanelHistogram())
