import sys, numpy as np
sys.path.insert(0,'.')
import pynucleus_b200 as pb
def hilbert(x, y):
    d = np.zeros_like(x, dtype=np.uint64); x=x.copy(); y=y.copy()
    s = 32768
    while s>0:
        rx = (x & s)>0; ry=(y & s)>0
        d += np.uint64(s)*np.uint64(s)*((3*rx.astype(np.uint64))^ry.astype(np.uint64))
        m = ~ry
        fl = m & rx
        x[fl] = 65535-x[fl]; y[fl]=65535-y[fl]
        xs = x[m].copy(); x[m]=y[m]; y[m]=xs
        s//=2
    return d
sides, r, GC = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
mesh = pb.refined(pb.polygon_disc(sides), r); dm = pb.P1_DoFMap(mesh)
ctr = mesh.vertices[mesh.cells].mean(axis=1)
lo = ctr.min(0); hi = ctr.max(0)
q = ((ctr-lo)/(hi-lo)*65535).astype(np.int64)
order = np.argsort(hilbert(q[:,0],q[:,1]), kind='stable')
nc = mesh.num_cells
ng = (nc+GC-1)//GC
nld=[]; nb=[]
grp = np.empty(nc, dtype=int)
for g in range(ng):
    cs = order[g*GC:(g+1)*GC]; grp[cs]=g
    d = dm.dofs[cs]; nld.append(len(np.unique(d[d>=0])))
    # greedy batches
    batches=[]; 
    for c in cs:
        v=set(mesh.cells[c])
        for b in batches:
            if len(b[0])<16 and not (b[1]&v):
                b[0].append(c); b[1]|=v; break
        else: batches.append([[c],set(v)])
    nb.append(len(batches))
nld=np.array(nld); nb=np.array(nb)
print('groups',ng,'nld mean %.1f max %d'%(nld.mean(),nld.max()),'batches mean %.2f max %d'%(nb.mean(),nb.max()), 'N',dm.num_dofs, 'sum nld^2/N^2 %.2f'%( (nld.sum()**2)/dm.num_dofs**2))
# adjacency / colors
from collections import defaultdict
vg = defaultdict(set)
for c in range(nc):
    for v in mesh.cells[c]: vg[v].add(grp[c])
adj=[set() for _ in range(ng)]
for v,s in vg.items():
    for a in s:
        adj[a]|=s
col=-np.ones(ng,dtype=int)
for g in range(ng):
    used={col[a] for a in adj[g] if a!=g and col[a]>=0}
    c=0
    while c in used: c+=1
    col[g]=c
print('colors',col.max()+1,'mean adj',np.mean([len(a)-1 for a in adj]))
def balanced(cs, B0):
    batches=[[[],set()] for _ in range(B0)]
    for c in cs:
        v=set(mesh.cells[c])
        best=None
        for b in batches:
            if len(b[0])<16 and not (b[1]&v):
                if best is None or len(b[0])<len(best[0]): best=b
        if best is None:
            best=[[],set()]; batches.append(best)
        best[0].append(c); best[1]|=v
    return len(batches)
for extra in (0,1,2):
    nb2=[balanced(order[g*GC:(g+1)*GC], (min(GC,nc-g*GC)+15)//16+extra) for g in range(ng)]
    print('balanced extra',extra,'mean %.2f max %d'%(np.mean(nb2),np.max(nb2)))
def roundrobin(cs, B0):
    batches=[[[],set()] for _ in range(B0)]
    for i,c in enumerate(cs):
        v=set(mesh.cells[c]); done=False
        for t in range(len(batches)):
            b=batches[(i+t)%len(batches)]
            if len(b[0])<16 and not (b[1]&v):
                b[0].append(c); b[1]|=v; done=True; break
        if not done:
            batches.append([[c],set(v)])
    return len(batches)
nb3=[roundrobin(order[g*GC:(g+1)*GC], (min(GC,nc-g*GC)+15)//16) for g in range(ng)]
print('roundrobin mean %.2f max %d'%(np.mean(nb3),np.max(nb3)), np.bincount(nb3))
