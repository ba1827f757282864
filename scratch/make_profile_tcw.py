"""profiles/r01_tcw_ncu_full.md from an `ncu --set full` capture of cc_tcw_bwd_kernel (scratch).
usage: ncu -i gpurun_out/prof_tcw_final.ncu-rep --page raw --csv > /tmp/tcw_raw.csv ; python scratch/make_profile_tcw.py /tmp/tcw_raw.csv B T "<command>" """
import csv, sys
rows = list(csv.reader(open(sys.argv[1]))); B, T, cmd = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
hdr, units, r = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.max.per_second"]
md = ["# Round 1 — `cc_tcw_bwd_kernel` (ModelTrainer reverse pass with the tensor-core weight gradient), `ncu --set full --clock-control none --import-source on` (1×B200)", "",
      "Command: `%s` — %d trajectories × %d steps, compact model S = 50 (BASELINE configs[3] architecture). Numbers under ncu are not bench values." % (cmd, B, T), "",
      "| metric | unit | value |", "|---|---|---|"]
v = {}
for w in want:
    if w in hdr:
        i = hdr.index(w); md.append("| `%s` | %s | %s |" % (w, units[i], r[i])); v[w] = r[i]
def gb(name):
    i = hdr.index(name); x = float(r[i]); u = units[i].lower()
    return x * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}[u]
rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
cyc = float(v["sm__cycles_active.avg"]); tiles = (B + 127) // 128; rounds = -(-tiles // 148)
md += ["", "Reading.",
       "* SASS: `UTCHMMA` with a tensor-memory A operand (adjoint and recompute chains) and with two shared-memory descriptors (weight gradient), `USETMAXREG` (56 / 224 registers).",
       "* DRAM: %.1f MB read + %.1f MB written = **%.1f B per trajectory-step** (algorithmic: tape 200 B + u 4 B + ybar 4 B; the stage-input slots and the flushed" % (rd / 1e6, wr / 1e6, (rd + wr) / (B * T)),
       "  gradient sums are written and re-read by the same thread and stay in L2: hit rate %s %%)." % v.get("lts__t_sector_hit_rate.pct", "?"),
       "* %.0f active cycles per step and round (%d tiles on 148 SMs = %d round(s)) = %.1f k cycles per window (4 windows per step); tensor pipe active %s %% of the" % (cyc / T / rounds, tiles, rounds, cyc / T / rounds / 4 / 1e3, v.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "?")),
       "  active cycles: the kernel is bound by the per-window dependent chain of the ONE tile an SM holds (barrier → 42 TS-mode MMAs → tcgen05.ld → epilogue → tcgen05.st / operand stores → barrier),",
       "  not by a pipe.  Ablations (`CC_B200_TCW_DEBUG`, B = 16,384 × 1000, at the 16.9 ms state): without the weight-gradient MMAs 16.0 ms; without the operand stores 13.9 ms; without both 12.9 ms.",
       "* History (reverse pass of 16,384 × 1000): first version (288 threads, 168 registers, spills) 23.2 ms → setmaxnreg 56/224 with lean descriptor arithmetic in the issuing thread 25.1 ms (40-register issuer with",
       "  spilled descriptors: 46 ms) → pointer-walk tape/slot addressing (20 → 3 instructions per row; `no_instruction` stalls gone) 17.0 ms → ybar/u fetched one step ahead 16.7 ms →",
       "  g^T d folded into the adjoint MMA (the readout cotangent rides in k̄_1's spare columns; 224 shared-memory reads per thread and step less) **15.1 ms**; 65,536 × 1000: 416 ms (FFMA kernel) → **59.4 ms**.", ""]
open('profiles/r01_tcw_ncu_full.md', 'w').write("\n".join(md))
print("\n".join(md[-9:]))
