import sys, time
sys.path.insert(0,'.')
import numpy as np, torch
import pynucleus_b200 as pb
for r, delta in ((5, 0.3), (6, 0.15), (6, 0.08)):
    mesh = pb.refined(pb.uniform_disc(), r); dm = pb.P1_DoFMap(mesh)
    for k in (pb.getFractionalKernel(2, 0.4, delta), pb.getIntegrableKernel(2, 'constant', delta)):
        b = pb.nonlocalBuilder(dm, k, {'target_order': 0.5})
        b.getDense(); torch.cuda.synchronize()
        t0 = time.perf_counter(); A = b.getDense(); torch.cuda.synchronize(); t1 = time.perf_counter()
        d = A.device_data
        nnz = int((d != 0).sum())
        print('r=%d N=%d h=%.3f delta=%.2f %-40s %.1f ms  nonzeros %.1f%%  sym %s' % (r, dm.num_dofs, mesh.h, delta, repr(k)[:40], (t1-t0)*1e3, 100.*nnz/d.numel(), bool(torch.equal(d, d.t()))))
