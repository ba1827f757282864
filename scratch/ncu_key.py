import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum','launch__block_size','launch__grid_size','launch__registers_per_thread ','shared_mem_per_block_dynamic','smsp__inst_executed.sum ','smsp__issue_active.avg.pct','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_fma.avg.pct','stalled_barrier_per','stalled_short_sc','stalled_long_sc','stalled_not_sel','stalled_wait_per','stalled_no_inst','stalled_branch','stalled_dispatch','stalled_math','stalled_mio','stalled_lg','inst_executed_op_local','bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed_pipe_lsu','sm__inst_executed_pipe_fma.sum ','sm__inst_executed_pipe_alu.sum ','sm__inst_executed_pipe_fmaheavy.sum ', 'sm__inst_executed_pipe_fmalite', 'smsp__inst_executed_pipe','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ','sm__warps_active.avg.pct', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ']
for vals in rows[2:]:
    for h,u,v in zip(hdr,units,vals):
        if any(w in h+' ' for w in want) and 'pcsamp' not in h and '_not_issued' not in h:
            print(h, u, v)
    print('---')
