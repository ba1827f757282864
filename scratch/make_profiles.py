"""regenerate profiles/r01_tc_* from gpurun_out (launch list, full capture exported with `ncu --page raw --csv`, bench JSON, sweep)."""
import csv, collections, json, sys
bench = json.loads(open('gpurun_out/bench_r01_tc.json').read().strip().split('\n')[-1])
cfg2 = json.loads(open('gpurun_out/bench_r01_cfg2.json').read().strip().split('\n')[-1])
cfg1 = json.loads(open('gpurun_out/bench_r01_cfg1.json').read().strip().split('\n')[-1])
rf = bench['roofline']
rows = list(csv.reader(open('gpurun_out/launches_r01_tc.csv')))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    name = r[idx['Kernel Name']]; v = float(r[idx['Metric Value']]); u = r[idx['Metric Unit']]
    ms = v / 1e6 if u in ('ns', 'nsecond') else (v / 1e3 if u in ('us', 'usecond') else v)
    agg[name][0] += 1; agg[name][1] += ms
tot = sum(v[1] for v in agg.values())
out = ["# Round 1 (tcgen05 kernels, final state) — ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (1×B200)", "",
       "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` (cold-cache, serialised: compare SHARES). Raw CSV: `profiles/r01_tc_launches.csv`.", "",
       "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
share = {}
for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    out.append('| `%s` | %d | %.2f | %.1f%% |' % (n[:72], c, ms, 100 * ms / tot)); share[n] = 100 * ms / tot
fwd_ms, bwd_ms, step_ms = rf['fwd_kernel']['launch_ms'], rf['launch_ms'], bench['ms_per_step']
sb = [v for k, v in share.items() if 'tc_bwd' in k][0]; sf = [v for k, v in share.items() if 'tc_fwd' in k][0]
out += ["", "The two tcgen05 rollout kernels are %.1f %% of device time (`cc_tc_bwd_kernel` %.1f %%, `cc_tc_fwd_kernel` %.1f %%); the rest are the" % (sb + sf, sb, sf),
        "FP32 probes `bench.py` runs for the roofline yardstick, the fused MSE and the gradient reduction.",
        "The live CUDA-event timing of `bench.py` (fwd %.2f ms, bwd %.2f ms of a %.1f ms step → %.1f %% / %.1f %%) agrees with these shares." % (fwd_ms, bwd_ms, step_ms, 100 * fwd_ms / step_ms, 100 * bwd_ms / step_ms),
        "`cc_tc_bwd_kernel` is the dominant kernel and the one `roofline` in the bench line describes.", ""]
open('profiles/r01_tc_launches_summary.md', 'w').write("\n".join(out))
rows = list(csv.reader(open('/tmp/tcb_raw.csv')))
hdr = rows[0]; units = rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum", "sm__cycles_elapsed.max.per_second"]
vals = {}
md = ["# Round 1 (tcgen05 kernels, final state) — `ncu --set full --clock-control none --import-source on -k regex:cc_tc -s 6 -c 2` inside `python bench.py --steps 2 --warmup 3` (1×B200)", "",
      "Workload cfg2x: closed-loop ControllerTrainer step, 65,536 trajectories × 1,001 steps, compact controller (S_c=5) + frozen compact model (S_m=50).",
      "Numbers under ncu are not bench values; they explain the bench line. Report: `gpurun_out/prof_r01_tc_bench.ncu-rep` (scratch).", "",
      "| metric | unit | launch 0 | launch 1 |", "|---|---|---|---|"]
for w in want:
    if w in hdr:
        i = hdr.index(w); md.append("| `%s` | %s | %s | %s |" % (w, units[i], rows[2][i], rows[3][i])); vals[w] = (rows[2][i], rows[3][i])
tp = vals["TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"]
ins = vals["smsp__inst_executed.sum"]
rd = float(vals["dram__bytes_read.sum"][1]); wr = float(vals["dram__bytes_write.sum"][1])
md += ["", "Reading.",
       "* Both kernels issue `tcgen05.mma` (SASS `UTCHMMA`, kind::tf32, A operand in tensor memory): the tensor pipe is active %.0f %% (forward) / %.0f %% (reverse) of elapsed cycles." % (float(tp[0]), float(tp[1])),
       "  DRAM traffic equals the algorithmic bytes: forward writes the 6-row tape (x_c, y_prev: 1.57 GB) + yhat (0.26 GB) and reads the references (0.26 GB);",
       "  the reverse pass reads tape + references + cotangent = %.2f GB = **%.1f B per trajectory-step** (no wasted re-reads; < 3 %% of HBM bandwidth)." % (rd + wr, (rd + wr) * 1e9 / (65536 * 1001)),
       "* Neither pipe is saturated: a tile of 128 trajectories is a 1001 × 4-stage dependent chain `MMA batch → tcgen05.ld → RK4 bookkeeping + hi/lo split → tcgen05.st → barrier`.",
       "  Tensor memory (512 columns: D 64 + A_hi 56 + A_lo 56 per tile) and the register file (the trajectory's state and stage sum live in registers: 253–255 per thread) allow two tiles",
       "  per SM; they run in lock step (alternating them under a token was measured slower: `tcgen05.ld/st` traffic slows concurrent MMAs, `gpurun_out/tc_probe_lock.log`). Per stage",
       "  (timeline `profiles/r01_tc_timeline.log`): ≈ 1400–1600 cycles from the barrier to the completion of the two tiles' 42 MMAs (the N = 64 rate of 32 cycles per MMA), ≈ 500–600 cycles of",
       "  epilogue, ≈ 75 cycles to the next barrier. Stall profile: `wait` (fixed-latency dependencies of the single resident warp per scheduler and tile) and `no_instruction` lead.",
       "* Instructions per warp-step: forward %.1f k, reverse %.1f k. The reverse pass's extra is the tiny controller: recomputation (spread over the MMA windows), reverse pass and" % (float(ins[0]) / (2048 * 1001) / 1e3, float(ins[1]) / (2048 * 1001) / 1e3),
       "  per-thread gradient accumulation (shared-memory read-modify-writes), the last two still sequential after the model's final stage.",
       "* History of the round on this workload (bench step, ms): FFMA kernels 180 → first tcgen05 version 86 → register-resident controller + run-time stage loops (instruction cache) 52 →",
       "  unrolled MMA issue 49.6 → rotating issuer + controller stages behind the publishes 48.3 → compile-time controller shape (rows of a stage = one basic block) 34.3 → packed fp32x2",
       "  epilogues 33.7 → recomputation spread over the windows 32.4 → plan read from the constant bank + activation dispatch folded **%.1f**." % step_ms, ""]
open('profiles/r01_tc_ncu_full.md', 'w').write("\n".join(md))
# sweeps
sw = [json.loads(l) for l in open('gpurun_out/sweep_r01_tc.jsonl') if l.startswith('{')]
json.dump(sw, open('profiles/r01_tc_sweep_cfg3_cfg5.json', 'w'), indent=1)
m2 = ["# Round 1 (final state) — BASELINE.json configs[2] (forward-only sweep), configs[4] (T = 10,000 closed-loop BPTT) and the other bench workloads, 1×B200", "",
      "`scratch/sweep_configs.py`; CUDA-event timing, best of 2 after a warm-up; algorithmic FLOPs per SURVEY §8d; measured FP32-FMA peak 71 TFLOP/s.",
      "S50_d0 (depth-0 model): latency kernels (one warp per trajectory) at B ≤ 1024, tcgen05 kernels above; the deep MLPs run on the FFMA kernels.", "",
      "## configs[2]: forward rollout, T = 1000", "", "| arch (S, W, depth) | kernels | B | ms | traj-steps/s | alg. TFLOP/s | × FP32-FMA peak |", "|---|---|---:|---:|---:|---:|---:|"]
for r in sw:
    if r['config'] != 'cfg3': continue
    k = ('latency' if r['B'] <= 1024 else 'tcgen05') if r['arch'] == 'S50_d0' else 'FFMA'
    if 'error' in r: m2.append("| %s | %s | %d | unsupported | %s | | |" % (r['arch'], k, r['B'], r['error'][:90])); continue
    m2.append("| %s | %s | %d | %.1f | %.3g | %.1f | %.2f |" % (r['arch'], k, r['B'], r['ms'], r['traj_steps_per_s'], r['alg_tflops'], r['alg_tflops'] / 71.0))
m2 += ["", "## configs[4]: closed-loop BPTT at T_r = 10,001 (compact controller S_c=5 + frozen compact model S_m=50, stable parameters, tape at every step)", "",
       "| B | kernels | ms per fwd+BPTT | traj-steps/s | alg. TFLOP/s | tape MB | loss | grad finite | FFMA kernels (first half of the round) |", "|---:|---|---:|---:|---:|---:|---:|---|---:|"]
old = {64: 688, 4096: 703}
for r in sw:
    if r['config'] != 'cfg5': continue
    m2.append("| %d | %s | %.0f | %.3g | %.2f | %.0f | %.4f | %s | %d ms |" % (r['B'], 'latency' if r['B'] <= 1024 else 'tcgen05', r['ms'], r['traj_steps_per_s'], r['alg_tflops'], r['tape_MB'], r['loss'], r['grad_finite'], old[r['B']]))
m2 += ["", "Both sizes are bound by the latency of the 10,001-step chain.", "",
       "## other bench workloads (`bench.py --workload ...`)", "", "| workload | ms per step | traj-steps/s | note |", "|---|---:|---:|---|",
       "| cfg2 = configs[1] literally: 64 trajectories × 1001 steps, closed-loop fwd + BPTT + adam | %.2f | %.3g | latency kernels (one warp per trajectory): 3.0 µs per step and direction; the tcgen05 kernels need 12.6 ms, the first FFMA kernels needed 69 ms |" % (cfg2['ms_per_step'], cfg2['value']),
       "| cfg1 = configs[0] literally: ModelTrainer, 8 trajectories × 1000 steps, fwd + BPTT (weight gradients) + adam | %.1f | %.2g | latency kernels (forward 1.6 ms, reverse 6.5 ms); with the FFMA reverse kernel 98.7 ms |" % (cfg1['ms_per_step'], cfg1['value']),
       "| cfg4 = configs[3]: ModelTrainer, 65,536 × 1000, S=50 depth 0 | 435.6 | 1.5e8 | forward 19 ms on tcgen05; the reverse pass with the model's weight gradient (416 ms, 6.5 TFLOP/s alg.) is still the FFMA kernel — next round |",
       "| cfg2x (default), 2 GPUs (torchrun, NCCL allreduce of 47 floats) | = 1 GPU | 2.00× | weak scaling measured earlier in the round (48.9 ms per step on 1 and on 2 GPUs) |", "",
       "Both arms on the reference's own small configurations, same box (`bench.py --workload cfgN` vs `bench.py --impl reference --workload cfgN`, 16 host threads):", "",
       "| workload | CPU restatement (C/OpenMP) | B200 | ratio |", "|---|---:|---:|---:|",
       "| cfg1: ModelTrainer, 8 × 1000 | 47.3 ms per step | 8.2 ms per step | 5.8× |",
       "| cfg2: ControllerTrainer, 64 × 1001 | 30.5 ms per step | 6.06 ms per step | 5.0× |",
       "| cfg2x: ControllerTrainer, 65,536 × 1001 (1,024-trajectory CPU sample) | 1.35e7 traj-steps/s | 2.25e9 traj-steps/s (e2e 2.17e9) | 166× (e2e 161×) |", ""]
open('profiles/r01_tc_sweep_cfg3_cfg5.md', 'w').write("\n".join(m2))
print('ok', step_ms, fwd_ms, bwd_ms)
