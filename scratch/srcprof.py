import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
cur=None; hdr=None
agg = collections.defaultdict(lambda:[0,0,''])
files = collections.defaultdict(lambda:[0,0])
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if r[0]=='Line No': hdr=r; iI=hdr.index('Instructions Executed'); iS=hdr.index('# Samples'); continue
    if r[0]=='Function Name' or hdr is None or len(r)<len(hdr): continue
    # source-level rows have line number and empty address
    if r[0]=='' : continue
    try: n=int(r[iI] or 0); sm=int(r[iS] or 0)
    except: continue
    if r[2] != '-': continue   # SASS rows have an address
    agg[(cur,int(r[0]))][0]+=n; agg[(cur,int(r[0]))][1]+=sm; agg[(cur,int(r[0]))][2]=r[1][:100]
    files[cur][0]+=n; files[cur][1]+=sm
tot=sum(v[0] for v in files.values()); tots=sum(v[1] for v in files.values())
print('total inst %.3e samples %d'%(tot,tots))
for f,v in files.items(): print('  %-28s inst %5.1f%% samples %5.1f%%'%(f,100*v[0]/tot,100*v[1]/tots))
top = int(sys.argv[2]) if len(sys.argv)>2 else 40
print('--- by samples')
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:top]:
    print('%-16s %5d inst %5.1f%% samp %5.1f%%  %s'%(k[0][:16],k[1],100*v[0]/tot,100*v[1]/tots,v[2]))
