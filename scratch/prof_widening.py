"""one product of the H2 operator (N = 48 769) and one P2 assembly (2 977 dofs) inside a profiler range (scratch; ncu)"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import pynucleus_b200 as pb
r = int(sys.argv[1]) if len(sys.argv) > 1 else 7
mesh = pb.refined(pb.uniform_disc(), r)
dm = pb.P1_DoFMap(mesh)
b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, 0.75), {'target_order': 0.5})
H = b.getH2()
x = torch.as_tensor(np.sin(np.arange(dm.num_dofs)*0.37)+0.1).cuda()
y = torch.empty_like(x)
for _ in range(3):
    H.matvec_device(x, y)
mesh2 = pb.refined(pb.uniform_disc(), 4)
b2 = pb.nonlocalBuilder(pb.P2_DoFMap(mesh2), pb.getFractionalKernel(2, 0.75), {'target_order': 0.5})
A2 = b2.getDense()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
H.matvec_device(x, y)
b2.getDense(out=A2.device_data)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('N', dm.num_dofs, 'near nnz', H.Anear.nnz, 'far pairs', sum(len(v) for v in H.Pfar.values()))
