import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
H=rows[0]
want=['Kernel Name','gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','sm__warps_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__thread_inst_executed.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed_op_local_ld.sum','smsp__inst_executed_op_local_st.sum','smsp__inst_executed_op_shared_ld.sum','smsp__inst_executed_op_global_ld.sum','smsp__inst_executed_op_global_st.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed_pipe_fp64.sum','sm__inst_executed_pipe_xu.sum','smsp__inst_executed_pipe_xu.sum','smsp__inst_executed_pipe_fma.sum','smsp__inst_executed_pipe_alu.sum','smsp__inst_executed_pipe_lsu.sum']
for w in want:
    if w in H:
        i=H.index(w); print(f"{w:70s}", [r[i][:60] for r in rows[2:]])
stall=[h for h in H if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h]
vals=[(float((rows[2][H.index(h)] or '0').replace(',','')),h) for h in stall]
for v,h in sorted(vals,reverse=True)[:8]: print(f"   stall (warps per issue slot) {h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''):28s} {v:.2f}")
