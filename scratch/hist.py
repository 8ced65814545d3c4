import sys, time, numpy as np
sys.path.insert(0,'.')
import oracle, pynucleus_b200 as pb
sides, r = int(sys.argv[1]), int(sys.argv[2])
mesh = pb.refined(pb.polygon_disc(sides), r); dm = pb.P1_DoFMap(mesh)
P = oracle.Problem(mesh.vertices, mesh.cells, dm.dofs, dm.num_dofs, 0.75, bfacets=mesh.boundaryFacets, target_order=0.5)
t=time.time(); h = P.histogram(); print(time.time()-t)
nn = {1:1,2:3,3:6,4:6,5:7}
tot=0; rows=[]
for k,v in sorted(h.items()):
    if k>0:
        n = nn.get(k, ((k+2)//2)**2 if k>5 else 1)   # conical GJ: ceil((k+1)/2)^2
        rows.append((k,v,n,v*n*n)); tot+=v*n*n
for k,v,n,w in rows: print(k, v, n, '%.3e'%w, '%.1f%%'%(100*w/tot))
print({k:v for k,v in h.items() if k<=0}, 'total node pairs %.3e'%tot)
