import sys, os, time
sys.path.insert(0,'.')
import numpy as np, torch
import pynucleus_b200 as pb
sides, r = (10,4) if len(sys.argv)<2 else (int(sys.argv[1]), int(sys.argv[2]))
mesh = pb.refined(pb.polygon_disc(sides), r); dm = pb.P1_DoFMap(mesh)
N = dm.num_dofs
out = {}
for name, dbg in (('tile', '2048'), ('group', '0')):
    os.environ['PNB_DEBUG'] = dbg
    b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2,0.75), {'target_order':0.5})
    A = torch.empty((N, N), dtype=torch.float64, device='cuda')
    for rep in range(2):
        b.getDense(out=A); torch.cuda.synchronize()
    print(name, b.getStats())
    out[name] = A.cpu().numpy()
    A2 = torch.empty((N, N), dtype=torch.float64, device='cuda'); b.getDense(out=A2)
    print(name, 'deterministic', torch.equal(A, A2), 'symmetric', torch.equal(A, A.T))
A, B = out['tile'], out['group']
d = np.sqrt(np.abs(np.diag(A)))
scale = np.maximum(np.abs(A), 1e-2*np.outer(d, d))
err = np.abs(A-B)/scale
print('N', N, 'max entry err', err.max(), 'at', np.unravel_index(err.argmax(), err.shape), 'max abs', np.abs(A-B).max())
