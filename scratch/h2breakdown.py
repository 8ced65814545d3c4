"""where getH2 spends its time at N = 12 097 (scratch; gpurun)"""
import sys, time, os
sys.path.insert(0, '.')
import numpy as np, torch
import pynucleus_b200 as pb
from pynucleus_b200 import h2
from pynucleus_b200.cluster_tree import admissible_clusters
r = int(sys.argv[1]) if len(sys.argv) > 1 else 6
mesh = pb.refined(pb.uniform_disc(), r); dm = pb.P1_DoFMap(mesh)
b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, 0.75), {'target_order': 0.5})
b.getDense(); torch.cuda.synchronize()
for rep in range(2):
    T = [time.time()]
    root = b.getTree(); T.append(time.time())
    Pnear, Pfar_nodes = admissible_clusters(root); T.append(time.time())
    for n in root.get_tree_nodes():
        if n.parent is not None:
            n.transferOperator = h2.transfer_operator(n.parent, n)
    T.append(time.time())
    pairs = [(lvl, a, c) for lvl in sorted(Pfar_nodes) for a, c in Pfar_nodes[lvl]]
    blocks = b.getFarFieldBlocks(np.array([a.box for _, a, _ in pairs]), np.array([c.box for _, _, c in pairs]),
                                 [a.interpolation_order for _, a, _ in pairs], [c.interpolation_order for _, _, c in pairs])
    T.append(time.time())
    os.environ['PNB_BENCH_VERBOSE'] = '1'
    near = b.assembleClusters(Pnear); torch.cuda.synchronize(); T.append(time.time())
    os.environ.pop('PNB_BENCH_VERBOSE')
    near.compile(); torch.cuda.synchronize(); T.append(time.time())
    Pfar = {}
    for (lvl, a, c), K in zip(pairs, blocks):
        Pfar.setdefault(lvl, []).append(h2.farFieldClusterPair(a, c, K))
    H = h2.H2Matrix(root, Pfar, near, dm.num_dofs, torch.device('cuda', 0))
    H.build_engine(mesh, dm); torch.cuda.synchronize(); T.append(time.time())
    names = ['tree', 'admissible', 'transfer ops', 'far blocks', 'assembleClusters', 'near compile', 'engine (+leaf moments)']
    print(' | '.join('%s %.3f' % (n, T[i+1]-T[i]) for i, n in enumerate(names)), '| total %.3f s' % (T[-1]-T[0]))
