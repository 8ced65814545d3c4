import sys, cProfile, pstats
sys.path.insert(0, '.')
import torch
import pynucleus_b200 as pb
mesh = pb.refined(pb.uniform_disc(), 6); dm = pb.P1_DoFMap(mesh)
b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, 0.75), {'target_order': 0.5})
b.getDense(); b.getH2(); torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
H = b.getH2(); torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
