import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]
want=['Kernel Name','gpu__time_duration.sum','launch__grid_size','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','dram__bytes_read.sum','dram__bytes_write.sum']
for r in rows[2:]:
    d=dict(zip(hdr,r))
    for w in want:
        if w in d: print(w, d[w])
    st=sorted(((float(d[h]) if d[h] not in ('','n/a') else 0,h) for h in hdr if 'warps_issue_stalled' in h and 'ratio' in h), reverse=True)[:7]
    for v,h in st: print('   %.2f %s'%(v,h.replace('smsp__average_warps_issue_stalled_','').replace('smsp__average_warp_latency_issue_stalled_','')))
