import csv,sys,subprocess
rep=sys.argv[1]; kern=sys.argv[2]; skip=sys.argv[3] if len(sys.argv)>3 else '0'
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv','--kernel-name','regex:'+kern,'--launch-skip',skip,'--launch-count','1'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; d=dict(zip(hdr,rows[2]))
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','launch__grid_size','sm__cycles_elapsed.avg','launch__shared_mem_per_block_dynamic']
for k in keys:
    if k in d: print(f'{k} = {d[k]} {rows[1][hdr.index(k)]}')
print('stalls per issue:', {k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''):round(float(v),2) for k,v in d.items() if 'issue_stalled' in k and 'per_issue_active' in k and 'pcsamp' not in k and float(v)>0.05})
src=subprocess.run(['ncu','-i',rep,'--page','source','--print-source','cuda,sass','--csv','--kernel-name','regex:'+kern,'--launch-skip',skip,'--launch-count','1'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
cur=None; agg={}; hdr=None
for r in rows:
    if len(r)==2 and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if len(r)>7 and r[0]=='Line No': hdr=r; ie=hdr.index('Instructions Executed'); smp=hdr.index('# Samples'); continue
    if hdr and len(r)>ie and r[0].isdigit() and r[2]=='-':
        try: agg[(cur,int(r[0]))]=(int(r[ie]), int(r[smp]), r[1][:95])
        except: pass
tot=sum(v[0] for v in agg.values()); tots=sum(v[1] for v in agg.values())
print('total warp inst',tot,'samples',tots)
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:int(sys.argv[4]) if len(sys.argv)>4 else 30]:
    print(f"{v[0]/tot*100:5.1f}% inst {v[1]/tots*100:5.1f}% smp  {k[0]}:{k[1]:<4} {v[2]}")
