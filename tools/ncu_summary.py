import csv,sys,collections,subprocess
rep=sys.argv[1]; units=float(sys.argv[2]) if len(sys.argv)>2 else 1
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
for r in rows[2:]:
    print(' | '.join(f"{w.split('.')[0].replace('smsp__','').replace('sm__','')}={r[hdr.index(w)]}" for w in want if w in hdr))
    st={hh.replace('smsp__pcsamp_warps_issue_stalled_',''):float(r[i]) for i,hh in enumerate(hdr) if 'pcsamp_warps_issue_stalled' in hh and 'not_issued' not in hh and r[i] not in ('','n/a')}
    tot=sum(st.values())
    print('   stalls:', ', '.join(f'{k}={v/tot*100:.0f}%' for k,v in sorted(st.items(),key=lambda kv:-kv[1])[:8]))
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
his=[i for i,r in enumerate(rows) if r and r[0]=='Address']
hi=his[0]; hdr=rows[hi]
ie=hdr.index('Instructions Executed'); sc=hdr.index('Source')
agg=collections.Counter(); tot=0
end=his[1] if len(his)>1 else len(rows)
for r in rows[hi+1:end]:
    if len(r)<=ie or not r[0].startswith('0x'): continue
    op=r[sc].split(); o=op[1] if op[0].startswith('@') else op[0]; o=o.split('.')[0]
    n=int(r[ie]); agg[o]+=n; tot+=n
print('instr total',tot,'per unit',tot/units)
print('  '.join(f'{o}={n/units:.1f}' for o,n in agg.most_common(22)))
