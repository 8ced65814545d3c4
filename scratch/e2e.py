import sys, os, time
sys.path.insert(0,'.')
import numpy as np, torch
import pynucleus_b200 as pb
sides, r = (10,6) if len(sys.argv)<2 else (int(sys.argv[1]), int(sys.argv[2]))
mesh = pb.refined(pb.polygon_disc(sides), r); dm = pb.P1_DoFMap(mesh)
N = dm.num_dofs
host = torch.empty((N, N), dtype=torch.float64).pin_memory(); hA = host.numpy()
kernel = pb.getFractionalKernel(2,0.75)
for k in range(4):
    torch.cuda.synchronize(); t0=time.perf_counter()
    b = pb.nonlocalBuilder(dm, kernel, {'target_order':0.5}); b.problem
    t1=time.perf_counter()
    b.getDenseHost(out=hA)
    t2=time.perf_counter()
    del b
    t3=time.perf_counter()
    print('e2e step %d: setup %.1f ms, assemble+copy %.1f ms, destroy %.1f ms' % (k,(t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3), file=sys.stderr)
