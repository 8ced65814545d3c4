import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import pynucleus_b200 as pb
mesh = pb.refined(pb.uniform_disc(), 6)
dm = pb.P1_DoFMap(mesh)
b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, 0.75), {'target_order': 0.5})
H = b.getH2()
x = torch.as_tensor(np.sin(np.arange(dm.num_dofs)*0.37)+0.1).cuda()
y = torch.empty_like(x)
for _ in range(3):
    H.matvec_device(x, y)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
H.matvec_device(x, y)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
nodes = list(H.tree.get_tree_nodes())
print('nodes', len(nodes), 'leaves', sum(n.isLeaf for n in nodes), 'levels', max(n.levelNo for n in nodes)+1,
      'orders', sorted(set(n.interpolation_order for n in nodes)),
      'far MB %.1f' % (sum(cp.kernelInterpolant.size for v in H.Pfar.values() for cp in v)*8e-6),
      'transfer MB %.1f' % (sum(n.transferOperator.size for n in nodes if n.parent is not None)*8e-6),
      'leaf MB %.1f' % (sum(n.value.size for n in nodes if n.isLeaf)*8e-6), 'near MB %.1f' % (H.Anear.nnz*12e-6))
