import numpy as np, sys
sys.path.insert(0,'.')
import pynucleus_b200 as pb, oracle
for noRef, s in ((8,0.25),(7,0.25),(8,0.75)):
    mesh = pb.refined(pb.simpleInterval(-1.,1.), noRef); dm = pb.P1_DoFMap(mesh)
    b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(1, s), {})
    A = b.getDense().data
    P = oracle.Problem(mesh.vertices, mesh.cells, dm.dofs, dm.num_dofs, s, bfacets=mesh.boundaryFacets)
    Aref = P.dense(True)
    d = np.sqrt(np.abs(np.diag(Aref))); scale = np.maximum(np.abs(Aref), 1e-2*np.outer(d,d))
    E = np.abs(A-Aref)/scale
    bad = np.argwhere(E > 1e-12)
    print(noRef, s, 'N', dm.num_dofs, 'max', E.max(), 'nbad', len(bad), 'maxorder', b.problem.max_order, P.P.max_order)
    for (i,j) in bad[:10]: print('   ', i, j, A[i,j], Aref[i,j], E[i,j])
