import csv, subprocess, sys
from collections import Counter
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__warps_eligible.avg.per_cycle_active','lts__t_sector_hit_rate.pct','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__inst_executed_op_global_red.sum','lts__t_sectors_srcunit_tex_op_red.sum','sm__cycles_active.avg','launch__waves_per_multiprocessor','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print('----', r[hdr.index('Kernel Name')][:50])
    for w in want:
        if w in hdr: print(f"  {w:64s} {r[hdr.index(w)]}")
for kn in sys.argv[2:]:
    src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--kernel-name','regex:'+kn],capture_output=True,text=True).stdout
    rows=list(csv.reader(src.splitlines()))
    hdr=rows[1]; data=rows[2:]
    ie=hdr.index('Instructions Executed'); isrc=hdr.index('Source'); isamp=hdr.index('# Samples')
    data=[r for r in data if len(r)>max(ie,isrc,isamp)]; tot=sum(int(r[ie]) for r in data if r[ie].isdigit())
    c=Counter()
    for r in data:
        if not r[ie].isdigit(): continue
        t=r[isrc].split(); op=(t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        c[op]+=int(r[ie])
    print('==', kn, 'total warp-inst', tot/1e6, 'M')
    print('  ', ' '.join(f"{o}:{v/1e6:.1f}" for o,v in c.most_common(16)))
    stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    st=Counter()
    for r in data:
        for h in stalls:
            v=r[hdr.index(h)]
            if v.isdigit(): st[h]+=int(v)
    tot_s=sum(st.values())
    print('  stalls:', ' '.join(f"{k[6:]}:{100*v/tot_s:.0f}%" for k,v in st.most_common(8)))
