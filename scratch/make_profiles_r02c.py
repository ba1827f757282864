"""regenerate profiles/r02c_* from gpurun_out (fourth session of round 2: family 6 = deep-MLP reverse pass, two-segment data
source, final bench lines).  Usage: python scratch/make_profiles_r02c.py [tag]   (tag = prefix of the gpurun_out files, default r02g)"""
import collections, csv, json, os, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
P, G = "profiles", "gpurun_out"
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02g"

def last_json(path):
    return json.loads(open(path).read().strip().splitlines()[-1])

for src, dst in ((f"{TAG}_bench_n1.json", "r02c_bench_n1.json"), (f"{TAG}_bench_ref_n1.json", "r02c_bench_ref_n1.json"), (f"{TAG}_launches.csv", "r02c_launches.csv"),
                 (f"{TAG}_mlpb_raw.csv", "r02c_mlpb_raw.csv"), (f"{TAG}_env_raw.csv", "r02c_env_raw.csv"), (f"{TAG}_env_time.log", "r02c_env_time.log")):
    if os.path.exists(f"{G}/{src}"):
        shutil.copy(f"{G}/{src}", f"{P}/{dst}")
full = last_json(f"{P}/r02c_bench_n1.json")

# ---------------- launch list ----------------
rows = list(csv.reader(open(f"{P}/r02c_launches.csv")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    v = float(r[idx["Metric Value"]]); u = r[idx["Metric Unit"]]
    ms = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
    agg[r[idx["Kernel Name"]]][0] += 1; agg[r[idx["Kernel Name"]]][1] += ms
tot = sum(v[1] for v in agg.values())
sm = full["config"]["step_ms"]
out = ["# Round 2 (final, fourth session) — ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-others` (1xB200)", "",
       "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` (cold-cache, serialised: compare SHARES). Raw CSV: `profiles/r02c_launches.csv`.", "",
       "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    out.append("| `%s` | %d | %.2f | %.1f%% |" % (n[:90], c, ms, 100 * ms / tot))
out += ["", "Live CUDA-event timing of the same step in `bench.py` (profiles/r02c_bench_n1.json): forward %.2f ms, MSE %.2f ms, reverse pass %.2f ms of a %.2f ms step (%.0f %% / %.0f %% / %.0f %%)." %
        (sm["fwd"], sm["mse"], sm["bwd"], full["ms_per_step"], 100 * sm["fwd"] / full["ms_per_step"], 100 * sm["mse"] / full["ms_per_step"], 100 * sm["bwd"] / full["ms_per_step"]), ""]
open(f"{P}/r02c_launches_summary.md", "w").write("\n".join(out))

# ---------------- ncu --set full: family 6 and the data-source kernel ----------------
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]
md = ["# Round 2 (fourth session) — `ncu --set full --clock-control none --import-source on` of the deep-MLP reverse kernel (family 6) and the two-segment data-source kernel", "",
      "`cc_mlpb_bwd_kernel<3, 2>`: (64,64,2) model, 18,944 trajectories (148 tiles = one wave) x 16 steps (scratch/prof_mlpb.py); `cc_env_two_segments_kernel`: 65,536 environments x 64 control steps",
      "(640 RK4 physics steps each; scratch/prof_env.py).  Numbers under ncu are not bench values.  Raw pages: `profiles/r02c_mlpb_raw.csv`, `profiles/r02c_env_raw.csv`.", ""]
cols = []
for f in ("r02c_mlpb_raw.csv", "r02c_env_raw.csv"):
    rows = list(csv.reader(open(f"{P}/{f}")))
    cols.append((rows[0], rows[1], rows[2]))
md += ["| metric | " + " | ".join("`%s`" % c[2][c[0].index("Kernel Name")].split("(")[0].replace("void ", "") for c in cols) + " |", "|---|" + "---:|" * len(cols)]
for w in want:
    if all(w in c[0] for c in cols):
        md.append("| %s [%s] | " % (w, cols[0][1][cols[0][0].index(w)]) + " | ".join(c[2][c[0].index(w)] for c in cols) + " |")
for h in cols[0][0]:
    if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio") and all(h in c[0] for c in cols):
        try:
            vals = [float(c[2][c[0].index(h)].replace(",", "")) for c in cols]
        except ValueError:
            continue
        if max(vals) > 0.3:
            md.append("| stall %s (per issue) | " % h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "") + " | ".join("%.2f" % v for v in vals) + " |")
def val(c, k):
    return float(c[2][c[0].index(k)].replace(",", ""))
env = cols[1]
md += ["", "Reading.",
       "* `cc_mlpb_bwd_kernel`: one tile per SM (202 KB of shared memory: two streamed weight slabs, the K = trajectory operands of the weight-gradient product, staging), tensor pipe ~30 % active;",
       "  the rest of a round is the dependent chain `tcgen05.commit -> mbarrier -> tcgen05.ld -> activation' / hi-lo split -> tcgen05.st + operand stores -> arrive` (long_scoreboard + barrier stalls)",
       "  that a single resident tile cannot overlap with the next round's MMAs (29 rounds per step for L = 3).  DRAM traffic is the tape and the per-CTA gradient sums; L2 serves the weight slabs.",
       "* `cc_env_two_segments_kernel`: bound by the fp64 pipe (`sm__pipe_fp64_cycles_active` %.0f %% with %.0f %% of the warp slots filled: 122 registers -> 4 CTAs of 128 threads per SM); no local memory;" % (val(env, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"), val(env, "sm__warps_active.avg.pct_of_peak_sustained_active")),
       "  global loads %.1f and stores %.1f sectors per request (actions / observations move as 128-byte rows through the per-warp shared-memory tiles; a per-thread access at stride T would touch 32)." %
       (val(env, "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum") / val(env, "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"), val(env, "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum") / val(env, "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum")), ""]
if os.path.exists(f"{P}/r02c_env_time.log"):
    md += ["Timed without ncu (CUDA events, `scratch/prof_env.py`, two_segments_v1, 1000 control steps = 10,000 RK4 physics steps per episode):", "", "```"] + \
          [l.rstrip() for l in open(f"{P}/r02c_env_time.log") if l.startswith("B=")] + ["```", ""]
ds = full["config"].get("other_workloads", {}).get("data_source")
if ds and "env_steps_per_s" in ds:
    md += ["`bench.py` (`other_workloads.data_source`): %d episodes x %d control steps in %.1f ms = %.3g environment-steps/s; NumPy oracle on one host thread: %.3g." %
           (ds["B"], ds["T"], ds["ms_per_rollout"], ds["env_steps_per_s"], ds.get("cpu_env_steps_per_s", float("nan"))), ""]
open(f"{P}/r02c_mlpb_env_ncu_full.md", "w").write("\n".join(md))

# ---------------- bench summary table ----------------
ow = full["config"].get("other_workloads", {})
t = ["# Round 2 (final) — `python bench.py` on 1xB200 (profiles/r02c_bench_n1.json; reference arm profiles/r02c_bench_ref_n1.json)", "",
     "| workload | ms per step | trajectory-steps/s | forward ms | reverse ms | CPU port ms per step |", "|---|---:|---:|---:|---:|---:|",
     "| bench (cfg2x: closed loop 65,536 x 1001) | %.2f | %.3g | %.2f | %.2f | — |" % (full["ms_per_step"], full["value"], sm["fwd"], sm["bwd"])]
for k in ("cfg1", "cfg2", "cfg4", "cfg5"):
    if k in ow and "ms_per_step" in ow[k]:
        e = ow[k]
        t.append("| %s: %d x %d%s | %.2f | %.3g | %.2f | %.2f | %s |" % (k, e["B"], e["T"], ", ckpt %d" % e["ckpt_every"] if e["ckpt_every"] > 1 else "", e["ms_per_step"], e["value"], e["fwd_ms"], e["bwd_ms"],
                                                                     "%.0f" % e["cpu_ms_per_step"] if "cpu_ms_per_step" in e else "—"))
t += ["", "e2e (host buffers, H2D of %.0f MB per step inside the timed region): %.2f ms per step = %.3g trajectory-steps/s (host -> device copy rate of this box %.1f GB/s; per-step host wall clock min / median / max %s)." %
      (full["e2e"]["h2d_bytes_per_step"] / 1e6, full["e2e"]["ms_per_step"], full["e2e"]["value"], full["e2e"]["h2d_gbs_this_box"],
       " / ".join("%.2f" % full["e2e"]["host_step_ms"][k] for k in ("min", "median", "max")) if "host_step_ms" in full["e2e"] else "n/a"), ""]
c3 = ow.get("cfg3", {}).get("points", [])
if c3:
    t += ["configs[2] points (T = 1000): forward and forward + BPTT", "", "| architecture | B | fwd family | fwd ms | fwd alg. TFLOP/s | x FP32-FMA peak | BPTT family | fwd+tape ms | reverse ms | fwd+BPTT alg. TFLOP/s | x FP32-FMA peak |", "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
    for p_ in c3:
        b = p_.get("bptt", {})
        t.append("| %s | %d | %s | %.1f | %.1f | %.2f | %s |" % (p_["arch"], p_["B"], p_["kernel_family"], p_["ms"], p_["alg_tflops"], p_["frac_fp32_fma_peak"],
                 " | ".join([str(b.get("kernel_family", "—")), "%.1f" % b["ms_fwd_with_tape"], "%.1f" % b["ms_bwd"], "%.1f" % b["alg_tflops"], "%.2f" % b["frac_fp32_fma_peak"]]) if "ms_bwd" in b else "unsupported | — | — | — | —"))
    t.append("")
open(f"{P}/r02c_bench_summary.md", "w").write("\n".join(t))
print("wrote profiles/r02c_*")
