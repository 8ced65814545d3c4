import numpy as np, sys
sys.path.insert(0,'.')
import pynucleus_b200 as pb, oracle
for noRef, s in ((4,0.75),(4,0.3),(5,0.75)):
    mesh = pb.refined(pb.uniform_disc(), noRef); dm = pb.P1_DoFMap(mesh)
    for ze in (False, True):
        b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, s), {'target_order': 0.5}, zeroExterior=ze)
        A = b.getDense().data
        P = oracle.Problem(mesh.vertices, mesh.cells, dm.dofs, dm.num_dofs, s, bfacets=mesh.boundaryFacets, target_order=0.5)
        Aref = P.dense(ze)
        E = np.abs(A-Aref)/np.abs(Aref).max()
        rel = np.abs(A-Aref)/np.maximum(np.abs(Aref),1e-300)
        bad = np.argwhere(rel > 1e-12)
        print(noRef, s, ze, 'maxE', E.max(), 'maxrel', rel.max(), 'nbad', len(bad), 'maxorder', b.problem.max_order, P.P.max_order)
        for (i,j) in bad[:12]:
            print('   ', i, j, i//64, j//64, A[i,j], Aref[i,j], rel[i,j])
