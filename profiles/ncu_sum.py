import csv,sys,subprocess
out=subprocess.run(['ncu','-i',sys.argv[1],'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__grid_size','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','lts__t_sector_hit_rate.pct']
want+= [h for h in hdr if 'warps_issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h]
for r in rows[2:]:
    print('==',r[hdr.index('Kernel Name')][:60])
    for w in want:
        if w in hdr:
            v=r[hdr.index(w)]
            try:
                if float(v.replace(',',''))==0: continue
            except: pass
            print('   ',w.replace('smsp__average_warps_issue_stalled_','stall_').replace('_per_issue_active.ratio',''),v, rows[1][hdr.index(w)])
