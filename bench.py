#!/usr/bin/env python
"""bench.py — throughput of the B200 rollout + BPTT path (trajectory-steps/s) on the configuration
BASELINE.json's metric is quoted on.

    python bench.py --gpus N --steps K --warmup W          # N > 1: launched by torchrun (one rank per GPU)
    python bench.py --impl reference ...                    # the CPU path (oracle port) on the host cores

One "step" = one pass of the hot path over one batch of synthetic trajectories:
closed-loop forward rollout -> MSE value + cotangent -> BPTT -> gradient reduction
(-> NCCL allreduce of the P+1 gradient/loss vector when N > 1).  Default workload (`cfg2x`): the
closed-loop ControllerTrainer step of BASELINE.json configs[1] (compact controller S_c=5 + frozen
compact model S_m=50, RK4, dt=0.01, T_r=1001) at the batch of configs[3] (65,536 trajectories per GPU,
weak scaling).  `--workload cfg2` is configs[1] literally (64 trajectories, latency bound);
`--workload cfg4` is the open-loop ModelTrainer step of configs[3].
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

DT = 0.01


def specs():
    from chain_control_b200 import arch as A

    model = A.make_node_spec(50, 1, 1, DT, f_depth=0, readout_pre_step=True, u_transform=A.UT_ARCTAN)
    ctrl = A.make_node_spec(5, 2, 1, DT, f_depth=0, readout_pre_step=True, out_scale=0.5)
    return ctrl, model


def init_theta(spec, seed, stable):
    """eqx.nn.Linear-style U(+-1/sqrt(fan_in)); stable: subtract I from f's state block (contractive)."""
    rng = np.random.default_rng(seed)
    out = []
    for mlp in (spec.f, spec.g):
        for i, o in zip(mlp.sizes[:-1], mlp.sizes[1:]):
            lim = 1.0 / np.sqrt(i)
            W = rng.uniform(-lim, lim, size=(o, i))
            if stable and mlp is spec.f and mlp.n_layers == 1:
                W[:, : spec.state_dim] -= np.eye(spec.state_dim)
            out.append(W.ravel())
            if mlp.use_bias:
                out.append(rng.uniform(-lim, lim, size=(o,)))
    return np.concatenate(out).astype(np.float32)


def make_refs(B, Tr, seed):
    """random_steps_source-style references: constant per trajectory, sign * U(0, 3) (collect/circus.py:22-33)."""
    rng = np.random.default_rng(seed)
    amp = np.sign(rng.normal(size=(B, 1, 1))) * rng.uniform(0, 3, size=(B, 1, 1))
    return np.broadcast_to(amp, (B, Tr, 1)).astype(np.float32).copy()


WORKLOADS = {
    # name: (closed_loop, B per GPU, T)
    "cfg2x": (True, 65536, 1001),
    "cfg2": (True, 64, 1001),
    "cfg4": (False, 65536, 1000),
    "cfg1": (False, 8, 1000),
}
# algorithmic FLOPs per trajectory-step: F = 2 (n_st M_f + M_g)
F_C = 2 * (4 * (5 + 2) * 5 + 5)       # 290
F_M = 2 * (4 * (50 + 1) * 50 + 50)    # 20,500


# dram read+write bytes per trajectory-step of cc_tc_bwd_kernel from the ncu --set full capture (profiles/r01_tc_ncu_full.md);
# None until a capture of the current kernel exists
TRAFFIC_TC_BWD_BYTES_PER_TS = 32.4  # (2.120 GB read + 0.008 GB written) / (65,536 x 1,001)
# same for cc_tcw_bwd_kernel (ModelTrainer reverse pass): (13.658 GB read + 0.064 GB written) / (65,536 x 1,000), profiles/r01_tcw_ncu_full.md
TRAFFIC_TCW_BWD_BYTES_PER_TS = 209.4


def alg_flops(closed):
    fwd = F_C + F_M if closed else F_M
    tot = 3 * F_C + 2 * F_M if closed else 3 * F_M  # frozen model: forward + input-gradient only
    return fwd, tot


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (rows are time-stamped; only the rows that fall
    inside [mark_start, mark_end] count - the sampler itself is started before the warm-up so that it is already running)."""

    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()
        ok = [(t, r) for t, r in self.rows if len(r) >= 10 and r[2].replace(".", "").isdigit()]
        inside = [r for t, r in ok if self.t0 is not None and self.t0 <= t <= (self.t1 or 1e30) + 0.02]
        region = "timed region"
        if not inside:  # timed region shorter than nvidia-smi's sampling period: fall back to everything seen under load
            inside, region = [r for _, r in ok], "warm-up + timed region (timed region shorter than the sampling period)"
        sm = [float(r[2]) for r in inside]
        mx = [float(r[3]) for r in inside if r[3].replace(".", "").isdigit()]
        reasons = set()
        for r in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[6:10]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": region}


def cpu_baseline(closed, T, march="native", target_s=12.0):
    """the oracle's C port (OpenMP, all host threads) on a bounded sample of the same workload."""
    from oracle import c_oracle as CO

    ctrl, model = specs()
    thc, thm = init_theta(ctrl, 2, False), init_theta(model, 1, True)
    cores = CO.max_threads()
    B = 16 * cores
    run = (lambda refs: CO.closedloop(ctrl, model, thc, thm, refs, mse_loss=True, march=march)) if closed else \
          (lambda u: CO.rollout(model, thm, u, targets=np.zeros((u.shape[0], T + 1, 1), np.float32), march=march))
    x = make_refs(B, T, 7)
    run(x[: 16])
    t0 = time.time(); run(x); dt = time.time() - t0
    reps = max(1, int(target_s / max(dt, 1e-3)))
    reps = min(reps, 50)
    t0 = time.time()
    for _ in range(reps):
        run(x)
    dt = (time.time() - t0) / reps
    return {"value": B * T / dt, "unit": "traj-steps/s", "cores": cores, "kind": "port",
            "sample": f"{B} trajectories x {T} steps, fwd+BPTT, {reps} repetitions, C/OpenMP oracle port (-march={march}); "
                      "the reference's JAX path cannot be installed here"}


def run_reference(args):
    """--impl reference: the CPU implementation of the path (oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    closed, Bg, T = WORKLOADS[args.workload]
    from oracle import c_oracle as CO

    ctrl, model = specs()
    thc, thm = init_theta(ctrl, 2, False), init_theta(model, 1, True)
    cores = CO.max_threads()
    B = min(Bg, 16 * cores * 4)  # bounded sample of the workload per step
    x = make_refs(B, T, 7)
    tg = np.zeros((B, T + 1, 1), np.float32)
    step = (lambda: CO.closedloop(ctrl, model, thc, thm, x, mse_loss=True, march="native")) if closed else \
           (lambda: CO.rollout(model, thm, x, targets=tg, march="native"))
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        step()
    dt = time.time() - t0
    val = B * T * args.steps / dt
    sample = f"{B} trajectories x {T} steps per step (bounded sample of {Bg}/GPU), C/OpenMP oracle port, -march=native"
    print(json.dumps({
        "impl": "reference", "metric": "closed-loop neural-ODE traj-steps/s (fwd+BPTT)" if closed else "neural-ODE traj-steps/s (fwd+BPTT)",
        "value": val, "unit": "traj-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "closed_loop": closed, "B_sample": B, "T": T, "note": "CPU restatement, not JAX"},
        "cpu_baseline": {"value": val, "unit": "traj-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "traj-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2x", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from chain_control_b200 import _lib, ops, parallel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    closed, B, T = WORKLOADS[args.workload]
    ctrl, model = specs()
    theta_c = torch.tensor(init_theta(ctrl, 2, False), device=dev)
    theta_m = torch.tensor(init_theta(model, 1, True), device=dev)
    # synthetic trajectories of this rank's shard (weak scaling: B per GPU fixed)
    host_in = torch.from_numpy(make_refs(B, T, 100 + rank)).pin_memory()
    x = host_in.to(dev)
    targets = torch.zeros(B, T + 1, 1, device=dev) if not closed else None
    n_total = world * B * (T if closed else T + 1)  # global element count of the mean
    P = ctrl.n_params() if closed else model.n_params()
    L = _lib.lib()
    # Adam state of the trainable vector (optax.adam defaults) - the update is part of the step
    theta = (theta_c if closed else theta_m).clone()
    m1, m2 = torch.zeros_like(theta), torch.zeros_like(theta)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    kt = {"fwd": 0.0, "bwd": 0.0}
    step_no = [0]

    def step(xdev, timed=False):
        e = [ev() for _ in range(4)] if timed else None
        if timed: e[0].record()
        if closed:
            yhat, _, tape = ops.closedloop_fwd(ctrl, model, theta, theta_m, xdev, want_tape=True)
            if timed: e[1].record()
            loss, ybar = ops.mse_value_and_cotangent(xdev, yhat, n_total)
            if timed: e[2].record()
            g = ops.closedloop_bwd(ctrl, model, theta, theta_m, xdev, tape, ybar)
        else:
            y, _, tape = ops.rollout_fwd(model, theta, xdev, want_tape=True)
            if timed: e[1].record()
            loss, ybar = ops.mse_value_and_cotangent(targets, y, n_total)
            if timed: e[2].record()
            g, _ = ops.rollout_bwd(model, theta, xdev, tape, ybar)
        if timed: e[3].record()
        gsum, lsum = parallel.allreduce_grad_and_loss(g, loss)  # N > 1: one NCCL allreduce of P+1 floats over NVLink
        gl = torch.cat([gsum, lsum.reshape(1)])
        step_no[0] += 1
        t = step_no[0]
        gg = gl[:P]
        m1.mul_(0.9).add_(gg, alpha=0.1)
        m2.mul_(0.999).addcmul_(gg, gg, value=0.001)
        theta.sub_(1e-3 * (m1 / (1 - 0.9 ** t)) / ((m2 / (1 - 0.999 ** t)).sqrt() + 1e-8))
        return gl, e

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ---------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step(x)
    barrier()
    sampler.mark_start()
    l0 = L.cc_launch_count()
    t_ev0, t_ev1 = ev(), ev()
    t_ev0.record()
    evs = []
    for _ in range(args.steps):
        gl, e = step(x, timed=True)
        evs.append(e)
    t_ev1.record()
    barrier()
    sampler.mark_end()
    launches = L.cc_launch_count() - l0
    ms = t_ev0.elapsed_time(t_ev1)
    for e in evs:
        kt["fwd"] += e[0].elapsed_time(e[1])
        kt["bwd"] += e[2].elapsed_time(e[3])
    clocks = sampler.stop() if rank == 0 else None
    tms = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    loss_val = float(gl[-1].item())

    # ---- end to end through the public API with HOST buffers ---------------------------------------
    # every step copies its own inputs host -> device (pinned memory) and reads its result back; the copy of step i+1
    # runs on a side stream while the kernels of step i execute (chain_control_b200.parallel.HostFeed, double buffered)
    hres = torch.empty(P + 1).pin_memory()
    for _ in range(2):
        step(host_in.to(dev, non_blocking=True))
    barrier()
    feed = parallel.HostFeed(dev)
    e0, e1 = ev(), ev()
    e0.record()
    feed.push(host_in)                               # H2D of step 0's inputs: inside the timed region
    for i in range(args.steps):
        xd = feed.pop()
        if i + 1 < args.steps:
            feed.push(host_in)                       # H2D of the next step's inputs, overlapped with this step's kernels
        gl2, _ = step(xd)
        feed.release()
        hres.copy_(gl2, non_blocking=True)           # D2H of the step's result (gradient + loss)
        torch.cuda.current_stream().synchronize()
    e1.record()
    barrier()
    ms_e2e = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    ms_e2e = float(ms_e2e.item())

    # ---- FP32 peak (live FFMA probe) ---------------------------------------------------------------
    sink = torch.zeros(4, device=dev)
    iters = 2048
    L.check(L.cc_fp32_peak_probe(sink.data_ptr(), iters, torch.cuda.current_stream().cuda_stream), "probe")
    torch.cuda.synchronize()
    p0, p1 = ev(), ev()
    p0.record()
    L.check(L.cc_fp32_peak_probe(sink.data_ptr(), iters, torch.cuda.current_stream().cuda_stream), "probe")
    p1.record()
    torch.cuda.synchronize()
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    peak_tflops = sms * 4 * 512 * iters * 16 * 8 * 2 / (p0.elapsed_time(p1) * 1e-3) / 1e12
    zin = torch.zeros(1024, device=dev)
    for _ in range(2):
        p0.record()
        L.check(L.cc_fp32_tile_probe(sink.data_ptr(), zin.data_ptr(), iters, torch.cuda.current_stream().cuda_stream), "tile probe")
        p1.record()
        torch.cuda.synchronize()
    regtile_tflops = sms * 4 * 512 * iters * 4 * 32 * 2 / (p0.elapsed_time(p1) * 1e-3) / 1e12

    if rank == 0:
        fwd_f, tot_f = alg_flops(closed)
        units = B * T  # trajectory-steps per launch
        bwd_ms = kt["bwd"] / args.steps
        fwd_ms = kt["fwd"] / args.steps
        achieved = units * (tot_f - fwd_f) / (bwd_ms * 1e-3) / 1e12
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            peaks = {}
        hbm = peaks.get("hbm_gbs", 6650.0)
        # algorithmic HBM bytes per trajectory-step of the whole step: inputs/outputs/cotangent + tape write+read
        tape_rows = (5 + 1) if closed else 50  # closed loop: x_c and y_prev (the frozen linear model needs no state)
        bytes_ts = 4 * (1 + 1 + 1 + 1 + 1) + 2 * 4 * tape_rows  # refs, yhat (w), yhat+refs (mse r), ybar (w), ybar+refs (bwd r) + tape w+r
        # ---- roofline of the dominant kernel -------------------------------------------------------
        # closed loop (default): both kernels run on the tensor pipe (tcgen05, kind::tf32, 3xTF32 split for fp32 accuracy).
        #   achieved = ALGORITHMIC FLOPs (2(n_st M_f + M_g) per trajectory-step, recomputation / splits / padding not
        #   counted) / launch time; peak = dense TF32 rate = half the measured dense bf16 rate of MEASURED_PEAKS.json
        #   (sustained figure: the kernel runs inside a long step).  `executed` counts what the pipe really does:
        #   3 passes (lo.hi, hi.lo, hi.hi) over the padded 128 x 64 x 56 tiles, 4 stages per trajectory-step.
        # open loop (cfg4): the reverse pass (weight gradients) is the FFMA kernel -> bound by the FP32 FMA pipe.
        import ctypes as _C
        cm, mm = ctrl.to_c(), model.to_c()
        fam_f = L.cc_kernel_family_at(_C.byref(cm) if closed else None, _C.byref(mm), 0, B)
        fam_b = L.cc_kernel_family_at(_C.byref(cm) if closed else None, _C.byref(mm), 1, B)
        fam_name = {0: "ffma", 1: "tcgen05", 2: "latency", 3: "tcgen05", 4: "blocked-linear tcgen05"}
        tensor = fam_b in (1, 3, 4)
        bf16_peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 2250.0 * 0.62))
        tf32_peak = bf16_peak / 2.0
        exec_flops_ts = 3 * 2 * 64 * 56 * 4  # per trajectory-step of one rollout kernel (forward or reverse): 3 x 2 x NP x KP x stages
        exec_note = "3xTF32 split (fp32-accurate products) over 128 x 64 x 56 padded tiles; 4.2x the algorithmic count"
        bwd_kernel, bwd_traffic = "cc_tc_bwd_kernel", TRAFFIC_TC_BWD_BYTES_PER_TS
        if fam_b == 3:  # ModelTrainer reverse pass: adjoint (4 stages) + recomputation (3 stages) + weight gradient (2 x M128 x N64 per trajectory and stage)
            exec_flops_ts = 3 * 2 * 64 * 56 * (4 + 3) + 2 * 2 * 128 * 64 * 4
            exec_note = "3xTF32: adjoint + recomputed stage inputs (128 x 64 x 56 tiles) + SS-mode weight-gradient product (hi/lo rows stacked: 2 x M128 x N64 per trajectory and stage)"
            bwd_kernel, bwd_traffic = "cc_tcw_bwd_kernel", TRAFFIC_TCW_BWD_BYTES_PER_TS
        if fam_b == 4:
            exec_flops_ts = 3 * 2 * 64 * 64 / 13.0
            exec_note = "blocked linear kernels: one 3xTF32 128 x 64 x 64 MMA batch per 13 steps"
            bwd_kernel, bwd_traffic = ("cc_lin_cl_bwd_kernel" if closed else "cc_linw_bwd_kernel"), None
        if tensor:
            roofline = {"bound": "tensor", "kernel": bwd_kernel, "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s",
                        "frac": achieved / tf32_peak,
                        "traffic": bwd_traffic * units if bwd_traffic else None,
                        "traffic_source": "ncu --set full capture of the same kernel (profiles/r01_tc_ncu_full.md, r01_tcw_ncu_full.md), dram read+write per trajectory-step x units of this launch",
                        "peak_source": "MEASURED_PEAKS.json dense bf16 (sustained) / 2 = dense TF32; the kernel issues tcgen05.mma kind::tf32",
                        "alg_flops_per_traj_step": tot_f - fwd_f, "launch_ms": bwd_ms,
                        "executed": {"tflops": units * exec_flops_ts / (bwd_ms * 1e-3) / 1e12, "frac": units * exec_flops_ts / (bwd_ms * 1e-3) / 1e12 / tf32_peak,
                                     "flops_per_traj_step": exec_flops_ts,
                                     "note": exec_note},
                        "fp32_fma": {"peak": peak_tflops, "frac": achieved / peak_tflops, "peak_register_tile": regtile_tflops,
                                     "note": "BASELINE.json's yardstick: algorithmic TFLOP/s over the FP32 FMA peak (FFMA-chain probe measured live; nominal 74.4)"},
                        "fwd_kernel": {"kernel": "cc_tc_fwd_kernel", "achieved": units * fwd_f / (fwd_ms * 1e-3) / 1e12, "launch_ms": fwd_ms, "alg_flops_per_traj_step": fwd_f,
                                       "frac": units * fwd_f / (fwd_ms * 1e-3) / 1e12 / tf32_peak, "frac_of_fp32_fma_peak": units * fwd_f / (fwd_ms * 1e-3) / 1e12 / peak_tflops},
                        "step": {"achieved": units * tot_f / ((fwd_ms + bwd_ms) * 1e-3) / 1e12, "alg_flops_per_traj_step": tot_f,
                                 "frac_of_fp32_fma_peak": units * tot_f / ((fwd_ms + bwd_ms) * 1e-3) / 1e12 / peak_tflops},
                        "hbm": {"achieved_gbs": units * bytes_ts / ((fwd_ms + bwd_ms) * 1e-3) / 1e9, "peak_gbs": hbm, "alg_bytes_per_traj_step": bytes_ts}}
        else:
            roofline = {"bound": "fp32", "kernel": {0: "cc_rollout_bwd_kernel", 2: ("cc_wpt_bwd_kernel" if closed else "cc_wpt_obwd_kernel")}.get(fam_b, "cc_rollout_bwd_kernel"), "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                        "frac": achieved / peak_tflops, "traffic": None,
                        "peak_source": "FFMA-chain probe measured live in this run (MEASURED_PEAKS.json carries no FP32 figure; nominal 74.4)",
                        "peak_register_tile": regtile_tflops,
                        "alg_flops_per_traj_step": tot_f - fwd_f, "launch_ms": bwd_ms,
                        "fwd_kernel": {"kernel": {0: "cc_rollout_fwd_kernel", 1: "cc_tc_fwd_kernel", 2: "cc_wpt_fwd_kernel"}[fam_f], "achieved": units * fwd_f / (fwd_ms * 1e-3) / 1e12, "launch_ms": fwd_ms, "alg_flops_per_traj_step": fwd_f},
                        "step": {"achieved": units * tot_f / ((fwd_ms + bwd_ms) * 1e-3) / 1e12, "alg_flops_per_traj_step": tot_f},
                        "hbm": {"achieved_gbs": units * bytes_ts / ((fwd_ms + bwd_ms) * 1e-3) / 1e9, "peak_gbs": hbm, "alg_bytes_per_traj_step": bytes_ts}}
        out = {
            "metric": ("closed-loop " if closed else "") + "neural-ODE traj-steps/s (fwd+BPTT)",
            "value": world * units * args.steps / (ms * 1e-3), "unit": "traj-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": {"cfg2x": "closed-loop ControllerTrainer step (configs[1] architecture: compact controller S_c=5 + frozen compact model S_m=50, RK4, dt=0.01, T_r=1001) at configs[3]'s batch, 65536 trajectories per GPU",
                                    "cfg2": "configs[1]: closed-loop ControllerTrainer BPTT, 64 trajectories x 1001 steps",
                                    "cfg4": "configs[3]: ModelTrainer step, 65536 trajectories per GPU x 1000 steps, S=50 depth 0",
                                    "cfg1": "configs[0]: ModelTrainer step, 8 trajectories x 1000 steps, S=50 depth 0"}[args.workload],
                       "B_per_gpu": B, "T": T, "closed_loop": closed, "integrator": "RK4", "ckpt_every": 1,
                       "kernels": {"forward": fam_name[fam_f], "reverse": fam_name[fam_b],
                                   "note": "tcgen05 = 3xTF32 tensor-core kernels (A in tensor memory); latency = one warp per trajectory (small batches); ffma = register-tiled FP32"},
                       "step": "fwd rollout + mse cotangent + BPTT + grad reduce" + (" + NCCL allreduce(P+1)" if world > 1 else "") + " + adam",
                       "l2": "inputs+tape (>= 0.26 GB + %.1f GB) exceed the 126 MB L2" % (2 * 4 * tape_rows * units / 2 / 1e9),
                       "loss": loss_val},
            "e2e": {"value": world * units * args.steps / (ms_e2e * 1e-3), "unit": "traj-steps/s",
                    "h2d_bytes_per_step": int(host_in.numel() * 4), "d2h_bytes_per_step": int((P + 1) * 4)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
        }
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(closed, T)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
