"""H2 operator at N = 12 097 and 48 769: assembly time, matvec time of the library's own kernels vs the sparse-product
formulation vs the dense matvec (scratch; run with gpurun)."""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, '.')
import pynucleus_b200 as pb

for r in [int(a) for a in sys.argv[1:]] or [6]:
    mesh = pb.refined(pb.uniform_disc(), r)
    dm = pb.P1_DoFMap(mesh)
    b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, 0.75), {'target_order': 0.5})
    b.problem
    torch.cuda.synchronize()
    t0 = time.time()
    H = b.getH2()
    torch.cuda.synchronize()
    t1 = time.time()
    x = torch.as_tensor(np.sin(np.arange(dm.num_dofs)*0.37)+0.1).cuda()
    y = torch.empty_like(x)

    def timeit(f, n=50):
        for _ in range(5):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            f()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)/n
    t_engine = timeit(lambda: H.matvec_device(x, y))
    ye = H.matvec_device(x).clone()
    eng = H._engine
    H._engine = None
    H.compile()
    t_spmv = timeit(lambda: H.matvec_device(x))
    ys = H.matvec_device(x)
    H._engine = eng
    print('N', dm.num_dofs, 'getH2 %.2f s' % (t1-t0), 'nnz near %.3e (%.1f %% of N^2)' % (H.Anear.nnz, 100.*H.Anear.nnz/dm.num_dofs**2),
          'far pairs', sum(len(v) for v in H.Pfar.values()),
          'matvec ms: own kernels %.3f, sparse products %.3f' % (t_engine, t_spmv), 'diff %.2e' % float((ye-ys).abs().max()/ys.abs().max()))
    if dm.num_dofs <= 50000:
        A = b.getDense()
        t_dense = timeit(lambda: A.matvec_device(x, y))
        yd = A.matvec_device(x)
        print('   dense matvec ms %.3f, H2 vs dense %.2e' % (t_dense, float((ye-yd).abs().max()/yd.abs().max())))
