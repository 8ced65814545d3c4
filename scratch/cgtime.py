"""CG iteration time: fused library kernels vs the torch step-by-step loop (scratch; gpurun)"""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import pynucleus_b200 as pb
for r in (5, 6):
    mesh = pb.refined(pb.uniform_disc(), r); dm = pb.P1_DoFMap(mesh)
    b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, 0.75), {'target_order': 0.5})
    A = b.getDense(); H = b.getH2()
    rhs = torch.ones(dm.num_dofs, dtype=torch.float64, device='cuda')
    for name, op in (('dense', A), ('H2', H)):
        for fused in (True, False):
            pb.cg(op, rhs, tol=1e-10, maxiter=500, fused=fused)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            x, its, res = pb.cg(op, rhs, tol=1e-10, maxiter=500, fused=fused)
            torch.cuda.synchronize(); dt = time.perf_counter()-t0
            print('N %d %s fused=%s: %d iterations, %.1f ms, %.1f us per iteration' % (dm.num_dofs, name, fused, its, dt*1e3, dt*1e6/max(its, 1)))
