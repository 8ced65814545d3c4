import sys
sys.path.insert(0,'.')
import numpy as np, torch
import pynucleus_b200 as pb
mesh = pb.refined(pb.polygon_disc(10), 4)
dm = pb.P1_DoFMap(mesh)
b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, 0.75), {'target_order': 0.5})
D = b.getDense().data
host = torch.empty((dm.num_dofs, dm.num_dofs), dtype=torch.float64).pin_memory()
host.fill_(float('nan'))
H = b.getDenseHost(out=host.numpy())
print('nan', np.isnan(H).sum())
d = np.abs(H-D)
print('max diff', np.nanmax(d), 'ndiff', (d>0).sum(), 'of', d.size)
idx = np.argwhere(d>0)
print(idx[:20])
print('rows with diff', np.unique(idx[:,0])[:30], len(np.unique(idx[:,0])))
i,j = idx[0]
print(H[i,j], D[i,j], H[j,i], D[j,i])
D2 = b.getDense().data
print('device twice equal', np.array_equal(D, D2))
H2 = b.getDenseHost(out=host.numpy()).copy()
print('host twice equal', np.array_equal(H, H2), 'sym', np.array_equal(H, H.T))
