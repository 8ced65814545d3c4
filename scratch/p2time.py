"""P2 dense assembly timings (row-owner kernel) next to the P1 production path on the same meshes (scratch; gpurun)."""
import sys
import time
import torch
sys.path.insert(0, '.')
import pynucleus_b200 as pb

for r in [int(a) for a in sys.argv[1:]] or [3, 4]:
    mesh = pb.refined(pb.uniform_disc(), r)
    for name, DM in (('P1', pb.P1_DoFMap), ('P2', pb.P2_DoFMap)):
        t0 = time.time()
        dm = DM(mesh)
        t1 = time.time()
        b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, 0.75), {'target_order': 0.5})
        A = b.getDense()
        torch.cuda.synchronize()
        t2 = time.time()
        A = b.getDense()
        torch.cuda.synchronize()
        t3 = time.time()
        print('r %d %s cells %d dofs %d: DoFMap %.2f s, first getDense %.3f s, second %.3f s (%.2e entries/s)' % (
            r, name, mesh.num_cells, dm.num_dofs, t1-t0, t2-t1, t3-t2, dm.num_dofs**2/(t3-t2)))
