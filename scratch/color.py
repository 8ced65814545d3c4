import sys, numpy as np, random
sys.path.insert(0,'.')
import pynucleus_b200 as pb
exec(open('scratch/groups.py').read().split("sides, r, GC")[0])
sides, r, GC = 10, 6, int(sys.argv[1])
mesh = pb.refined(pb.polygon_disc(sides), r); dm = pb.P1_DoFMap(mesh)
ctr = mesh.vertices[mesh.cells].mean(axis=1)
lo = ctr.min(0); hi = ctr.max(0)
q = ((ctr-lo)/(hi-lo)*65535).astype(np.int64)
order = np.argsort(hilbert(q[:,0],q[:,1]), kind='stable')
nc = mesh.num_cells; ng=(nc+GC-1)//GC
cells = mesh.cells
def color(cs, B, seed, mode):
    rnd = random.Random(seed)
    cs = list(cs)
    if mode=='rand': rnd.shuffle(cs)
    verts = [set() for _ in range(B)]; cnt=[0]*B; assign={}
    # DSATUR-ish: process cells in order of most constrained (number of feasible batches), ties random
    remaining=set(cs)
    vsets={c:set(cells[c]) for c in cs}
    while remaining:
        if mode=='dsat':
            best=None;bf=None
            for c in remaining:
                f=[b for b in range(len(cnt)) if cnt[b]<16 and not (verts[b]&vsets[c])]
                key=(len(f), rnd.random())
                if best is None or key<bk: best=c;bk=key;bf=f
            c=best;f=bf
        else:
            c=cs[len(assign)]
            f=[b for b in range(len(cnt)) if cnt[b]<16 and not (verts[b]&vsets[c])]
        if not f:
            verts.append(set());cnt.append(0);f=[len(cnt)-1]
        b=min(f,key=lambda b:(cnt[b],rnd.random()))
        assign[c]=b;cnt[b]+=1;verts[b]|=vsets[c];remaining.discard(c)
    return len(cnt)
res={'seq':[], 'rand5':[], 'dsat':[]}
for g in range(0,ng,8):
    cs=order[g*GC:(g+1)*GC]
    B=(len(cs)+15)//16
    res['seq'].append(color(cs,B,0,'seq'))
    res['rand5'].append(min(color(cs,B,s,'rand') for s in range(5)))
    res['dsat'].append(min(color(cs,B,s,'dsat') for s in range(2)))
for k,v in res.items(): print(k, np.mean(v), np.bincount(v))
