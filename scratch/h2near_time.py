import sys, time
sys.path.insert(0,'.')
import numpy as np, torch
import pynucleus_b200 as pb
from pynucleus_b200 import h2
from pynucleus_b200.cluster_tree import admissible_clusters
r = int(sys.argv[1]) if len(sys.argv) > 1 else 6
mesh = pb.refined(pb.uniform_disc(), r); dm = pb.P1_DoFMap(mesh)
b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2,0.75), {'target_order':0.5, 'near_field_batched': len(sys.argv) < 3})
b.getDense(); torch.cuda.synchronize()
root = b.getTree()
Pnear, Pfar = admissible_clusters(root)
import os
os.environ['PNB_BENCH_VERBOSE'] = '1'
for rep in range(2):
    t0 = time.time(); near = b.assembleClusters(Pnear); torch.cuda.synchronize(); t1 = time.time()
    print('assembleClusters %.3f s  pairs %d  nnz %d' % (t1-t0, len(Pnear), near.nnz))
print(b.getStats() if False else '')
