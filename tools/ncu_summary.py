import csv,sys,subprocess
f=sys.argv[1]
out=subprocess.run(['ncu','-i',f,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]; data=rows[2:]
want=['gpu__time_duration.sum','sm__cycles_elapsed.max','sm__cycles_active.avg','sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__m_xbar2l1tex_read_bytes.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpc__cycles_elapsed.avg.per_second','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_active','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed_pipe_xu.sum','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','smsp__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__warps_active.avg.per_cycle_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
for w in want:
    for i,h in enumerate(hdr):
        if h==w: print(f"{h} [{units[i]}] = {[r[i] for r in data]}")
