import sys
sys.path.insert(0, '.')
import torch
import pynucleus_b200 as pb
from pynucleus_b200.cluster_tree import admissible_clusters
mesh = pb.refined(pb.uniform_disc(), 6); dm = pb.P1_DoFMap(mesh)
b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, 0.75), {'target_order': 0.5})
b.getDense(); torch.cuda.synchronize()
root = b.getTree()
Pnear, Pfar = admissible_clusters(root)
torch.cuda.cudart().cudaProfilerStart()
near = b.assembleClusters(Pnear); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
