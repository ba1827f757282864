#!/usr/bin/env python
"""bench.py — throughput of the B200 rollout + BPTT path (trajectory-steps/s) on the configuration
BASELINE.json's metric is quoted on.

    python bench.py --gpus N --steps K --warmup W          # N > 1: launched by torchrun (one rank per GPU)
    python bench.py --impl reference ...                    # the CPU path (oracle port) on the host cores

One "step" = one pass of the hot path over one batch of synthetic trajectories:
closed-loop forward rollout -> MSE value + cotangent -> BPTT -> gradient reduction
(-> NCCL allreduce of the P+1 gradient/loss vector when N > 1) -> fused clip + Adam update (cc_opt_step).
Default workload (`cfg2x`): the closed-loop ControllerTrainer step of BASELINE.json configs[1] (compact controller
S_c=5 + frozen compact model S_m=50, RK4, dt=0.01, T_r=1001) at the batch of configs[3] (65,536 trajectories per GPU,
weak scaling); the same line carries a `strong` block (65,536 trajectories in total, sharded B/N) and, at N = 1,
`config.other_workloads` with configs[0], [1], [3] literally and configs[4] (T_r = 10,001, ckpt_every = 100).
Prints ONE JSON line on rank 0.
"""
import argparse
import csv
import glob
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
    # name: closed loop?, B per GPU, T, ckpt_every, description
    "cfg2x": (True, 65536, 1001, 1, "closed-loop ControllerTrainer step (configs[1] architecture: compact controller S_c=5 + frozen compact model "
                                    "S_m=50, RK4, dt=0.01, T_r=1001) at configs[3]'s batch, 65536 trajectories per GPU"),
    "cfg2": (True, 64, 1001, 1, "configs[1]: closed-loop ControllerTrainer BPTT, 64 trajectories x 1001 steps"),
    "cfg4": (False, 65536, 1000, 1, "configs[3]: ModelTrainer step, 65536 trajectories x 1000 steps, S=50 depth 0"),
    "cfg1": (False, 8, 1000, 1, "configs[0]: two_segments env (dt = 0.01, 1001 steps): ModelTrainer step on 8 trajectories recorded from the "
                                "two-segment data source (4 GP draws + 4 cosines: cc/utils/high_level/defaults.py:21-23 seeds), S=50 depth 0"),
    "cfg5": (True, 4096, 10001, 100, "configs[4]: closed-loop BPTT, T_r = 10001, 4096 trajectories, checkpointed recomputation (ckpt_every = 100)"),
}
# algorithmic FLOPs per trajectory-step: F = 2 (n_st M_f + M_g)
F_C = 2 * (4 * (5 + 2) * 5 + 5)       # 290
F_M = 2 * (4 * (50 + 1) * 50 + 50)    # 20,500


def alg_flops(closed):
    fwd = F_C + F_M if closed else F_M
    tot = 3 * F_C + 2 * F_M if closed else 3 * F_M  # frozen model: forward + input-gradient only
    return fwd, tot


def alg_bytes(closed, m_block=13):
    """algorithmic HBM bytes per trajectory-step of the forward / reverse kernel (4(I+O) = 8 B of inputs and
    outputs) plus the tape: closed loop (x_c, y_prev) = 24 B per step, written forward and read in reverse; open loop the
    state at the block starts, 4 S / 13 B per step"""
    tape = 24.0 if closed else 4.0 * 50 / m_block
    return 8.0 + tape, 8.0 + tape


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


def host_threads():
    """all the host cores, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1 to its workers)"""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_run(closed, B, T, threads, reps=1):
    """oracle C port (OpenMP, `threads` threads), fwd + BPTT of B trajectories x T steps: seconds per pass"""
    from oracle import c_oracle as CO

    ctrl, model = specs()
    thc, thm = init_theta(ctrl, 2, False), init_theta(model, 1, True)
    x = make_refs(B, T, 7)
    tg = np.zeros((B, T + 1, 1), np.float32)
    run = (lambda: CO.closedloop(ctrl, model, thc, thm, x, mse_loss=True, march="native", nthreads=threads)) if closed else \
          (lambda: CO.rollout(model, thm, x, targets=tg, march="native", nthreads=threads))
    run()
    t0 = time.time()
    for _ in range(reps):
        run()
    return (time.time() - t0) / reps


def cpu_baseline(closed, T, target_s=12.0):
    """the oracle's C port (OpenMP, all host threads) on a bounded sample of the same workload."""
    cores = host_threads()
    B = 16 * cores
    dt = cpu_run(closed, B, T, cores)
    reps = min(max(1, int(target_s / max(dt, 1e-3))), 50)
    dt = cpu_run(closed, B, T, cores, reps)
    return {"value": B * T / dt, "unit": "traj-steps/s", "cores": cores, "kind": "port",
            "sample": f"{B} trajectories x {T} steps, fwd+BPTT, {reps} repetitions, C/OpenMP oracle port (-march=native), {cores} threads; "
                      "the reference's JAX path cannot be installed here"}


def run_reference(args):
    """--impl reference: the CPU implementation of the path (oracle port) on the host cores: the same bounded sample at every N,
    all host threads whatever OMP_NUM_THREADS says."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    closed, Bg, T, _, _ = WORKLOADS[args.workload]
    cores = host_threads()
    B = min(Bg, 1024)  # bounded sample of the workload per step (the same at every N)
    cpu_run(closed, min(B, 64), T, cores)
    for _ in range(max(args.warmup, 1) - 1):
        cpu_run(closed, B, T, cores)
    dt = cpu_run(closed, B, T, cores, args.steps)
    val = B * T / dt
    sample = f"{B} trajectories x {T} steps per step (bounded sample of {Bg}/GPU), C/OpenMP oracle port, -march=native, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": "closed-loop neural-ODE traj-steps/s (fwd+BPTT)" if closed else "neural-ODE traj-steps/s (fwd+BPTT)",
        "value": val, "unit": "traj-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "closed_loop": closed, "B_sample": B, "T": T, "note": "CPU restatement, not JAX"},
        "cpu_baseline": {"value": val, "unit": "traj-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "traj-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cfg3_points(torch, dev, fp32_peak_tflops, cpu_threads):
    """BASELINE.json configs[2] (sweep, MLP width 32..128): three (architecture, batch) points, T = 1000, forward and forward + BPTT"""
    import ctypes as _C

    from chain_control_b200 import _lib, arch as A, ops
    from oracle import c_oracle as CO

    out = {"workload": "configs[2]: forward-only batched rollout, deep-MLP model (S = W, f depth 2, tanh final activation), T = 1000; "
                       "alg_tflops = 2 (4 M_f + M_g) FLOP per trajectory-step / time", "points": []}
    T = 1000
    for S, B in ((32, 131072), (64, 65536), (128, 16384)):
        spec = A.make_node_spec(S, 1, 1, DT, f_width=S, f_depth=2, f_act_final=A.ACT_TANH)
        rng = np.random.default_rng(S)
        parts = []
        for mlp in (spec.f, spec.g):
            for i, o in zip(mlp.sizes[:-1], mlp.sizes[1:]):
                parts += [0.5 * rng.uniform(-1, 1, size=o * i) / np.sqrt(i), rng.uniform(-1, 1, size=o) / np.sqrt(i)]
        th_h = np.concatenate(parts).astype(np.float32)
        th = torch.tensor(th_h, device=dev)
        u = torch.rand(B, T, 1, device=dev) * 2 - 1
        fam = _lib.lib().cc_kernel_family_at(None, _C.byref(spec.to_c()), 0, B)
        ops.rollout_fwd(spec, th, u)
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.rollout_fwd(spec, th, u)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = min(ts)
        mf = sum(i * o for i, o in zip(spec.f.sizes[:-1], spec.f.sizes[1:]))
        mg = sum(i * o for i, o in zip(spec.g.sizes[:-1], spec.g.sizes[1:]))
        fl = 2 * (4 * mf + mg)
        pt = {"arch": f"S{S}_W{S}_d2", "B": B, "T": T, "kernel_family": int(fam), "ms": ms, "value": B * T / (ms * 1e-3),
              "alg_tflops": B * T * fl / (ms * 1e-3) / 1e12, "frac_fp32_fma_peak": B * T * fl / (ms * 1e-3) / 1e12 / fp32_peak_tflops}
        # forward + BPTT (ModelTrainer's value_and_grad on the same architecture): algorithmic FLOPs = 3 x forward
        fam_b = _lib.lib().cc_kernel_family_at(None, _C.byref(spec.to_c()), 1, B)
        if fam_b < 0:
            pt["bptt"] = {"unsupported": _lib.lib().last_error()}
        else:
            ybar = torch.randn(B, T + 1, 1, device=dev) / (B * (T + 1))
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            for it in range(2):  # (first pass = warm-up)
                e[0].record()
                _, _, tape = ops.rollout_fwd(spec, th, u, want_tape=True)
                e[1].record()
                tb, _ = ops.rollout_bwd(spec, th, u, tape, ybar)
                e[2].record()
                torch.cuda.synchronize()
            msf, msb = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
            pt["bptt"] = {"kernel_family": int(fam_b), "ms_fwd_with_tape": msf, "ms_bwd": msb, "value": B * T / ((msf + msb) * 1e-3),
                          "alg_tflops": 3 * B * T * fl / ((msf + msb) * 1e-3) / 1e12,
                          "frac_fp32_fma_peak": 3 * B * T * fl / ((msf + msb) * 1e-3) / 1e12 / fp32_peak_tflops,
                          "grad_finite": bool(torch.isfinite(tb).all())}
            del tape, ybar
        if cpu_threads:
            Bc, Tc = 256, 100
            uc = u[:Bc, :Tc].cpu().numpy()
            CO.rollout(spec, th_h, uc, march="native", nthreads=cpu_threads)
            t0 = time.time()
            CO.rollout(spec, th_h, uc, march="native", nthreads=cpu_threads)
            dtc = time.time() - t0
            pt["cpu_value"] = Bc * Tc / dtc
            pt["cpu_sample"] = f"{Bc} trajectories x {Tc} steps forward on {cpu_threads} threads (C/OpenMP oracle port)"
        out["points"].append(pt)
        del u
    return out


def ncu_traffic(kernel_substr):
    """dram read + write bytes of one launch of the kernel from the newest committed `ncu --set full` raw page
    (profiles/r*_raw.csv, written from the .ncu-rep with `ncu --page raw --csv`): (bytes, file) or None"""
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r0*_raw.csv"))):
        try:
            rows = list(csv.reader(open(path)))
            hdr = rows[0]
            ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            for r in rows[2:]:
                if kernel_substr in r[ik]:
                    best = (float(r[ir]) * scale.get(rows[1][ir], 1.0) + float(r[iw]) * scale.get(rows[1][iw], 1.0), os.path.basename(path))
        except Exception:
            continue
    return best


class Workload:
    """one train step of a workload on this rank's shard"""

    def __init__(self, name, B, world, rank, dev, torch):
        from chain_control_b200 import ops

        self.torch, self.ops, self.name = torch, ops, name
        self.closed, _, self.T, self.ckpt, self.desc = WORKLOADS[name]
        self.B, self.world = B, world
        self.ctrl, self.model = specs()
        self.theta_m = torch.tensor(init_theta(self.model, 1, True), device=dev)
        theta0 = torch.tensor(init_theta(self.ctrl, 2, False), device=dev) if self.closed else self.theta_m.clone()
        self.P = theta0.numel()
        self.opt = ops.FusedAdam(theta0, 1e-3, clip_global_norm=1.0)  # chain(clip_by_global_norm(1.0), adam(1e-3)), test_trainer.py:58
        self.theta = self.opt.theta
        if name == "cfg1":
            # configs[0] literally: inputs and targets RECORDED from the two-segment environment (cc_env_two_segments_kernel),
            # sample_feedforward_and_collect(env, train_gp[:4], train_cos[:4]) of cc/utils/high_level/defaults.py:21-23
            from chain_control_b200.env import make_env, sample_feedforward_and_collect

            sample = sample_feedforward_and_collect(make_env("two_segments_v1", device=dev), seeds_gp=[0, 1, 2, 3], seeds_cos=[4, 8, 12, 2])
            self.host_in = torch.from_numpy(np.ascontiguousarray(sample.action[:B], np.float32)).pin_memory()
            self.targets = torch.from_numpy(np.ascontiguousarray(sample.obs["xpos_of_segment_end"][:B], np.float32)).to(dev)
        else:
            self.host_in = torch.from_numpy(make_refs(B, self.T, 100 + rank)).pin_memory()
            self.targets = None if self.closed else torch.zeros(B, self.T + 1, 1, device=dev)
        self.x = self.host_in.to(dev)
        self.n_total = world * B * (self.T if self.closed else self.T + 1)
        self.units = B * self.T

    def grads(self, x, ev=None):
        """rollout -> loss + cotangent -> BPTT: [theta_bar ; loss] of this shard"""
        ops, torch = self.ops, self.torch
        if ev: ev[0].record()
        if self.closed:
            yhat, _, tape = ops.closedloop_fwd(self.ctrl, self.model, self.theta, self.theta_m, x, want_tape=True, ckpt_every=self.ckpt)
            if ev: ev[1].record()
            loss, ybar = ops.mse_value_and_cotangent(x, yhat, self.n_total)
            if ev: ev[2].record()
            g = ops.closedloop_bwd(self.ctrl, self.model, self.theta, self.theta_m, x, tape, ybar, ckpt_every=self.ckpt)
        else:
            y, _, tape = ops.rollout_fwd(self.model, self.theta, x, want_tape=True)
            if ev: ev[1].record()
            loss, ybar = ops.mse_value_and_cotangent(self.targets, y, self.n_total)
            if ev: ev[2].record()
            g, _ = ops.rollout_bwd(self.model, self.theta, x, tape, ybar)
        if ev: ev[3].record()
        return torch.cat([g, loss.reshape(1)])

    def step(self, x, ev=None):
        import torch.distributed as dist

        gl = self.grads(x, ev)
        if self.world > 1:
            dist.all_reduce(gl, op=dist.ReduceOp.SUM)  # ONE NCCL allreduce of P+1 floats over NVLink
        self.opt.step(gl[: self.P])
        return gl


def env_point(torch, dev, cores):
    """two_segments_v1, B episodes x 1000 control steps (10 RK4 physics steps each) in one launch; CPU: the NumPy oracle
    (vectorised over a 256-episode sample x 100 steps, one thread) scaled to the same work"""
    from chain_control_b200.env import make_env

    B, T = 16384, 1000
    env = make_env("two_segments_v1", n_envs=B, device=dev)
    acts = (torch.rand(B, T, device=dev) * 2 - 1)
    for _ in range(2):
        env.rollout(acts)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        obs = env.rollout(acts)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    entry = {"workload": "two_segments_v1 episodes (cc/env data source, MuJoCo-free CUDA kernel, fp64 RK4 at 1 ms)", "B": B, "T": T,
             "ms_per_rollout": ms, "env_steps_per_s": B * T / (ms * 1e-3), "physics_steps_per_s": 10 * B * T / (ms * 1e-3),
             "obs_absmax": float(obs["xpos_of_segment_end"].abs().max())}
    if cores is not None:
        from oracle import two_segments_np as TS

        a = acts[:256, :100].cpu().numpy()
        t0 = time.perf_counter()
        TS.rollout("two_segments_v1", a)
        dt = time.perf_counter() - t0
        entry["cpu_env_steps_per_s"] = 256 * 100 / dt
        entry["cpu_sample"] = "NumPy oracle, 256 episodes x 100 steps vectorised on 1 thread"
    return entry


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2x", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip config.other_workloads")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from chain_control_b200 import _lib, parallel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    # N > 1: one process per GPU, pinned to the cores of its GPU's NUMA node before any pinned host buffer exists (first-touch
    # placement).  N = 1 keeps every core (the cpu_baseline leg uses them all).
    numa = parallel.bind_host_to_gpu_numa(local) if world > 1 else None
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    warm = max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxrank(ms):
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(w, steps, with_events=False):
        """W warm-up steps, then exactly `steps` timed steps between barriers; device time, max over ranks"""
        for _ in range(warm):
            w.step(w.x)
        barrier()
        l0 = L.cc_launch_count()
        e0, e1 = ev(), ev()
        evs = []
        e0.record()
        for _ in range(steps):
            e = [ev() for _ in range(4)] if with_events else None
            gl = w.step(w.x, e)
            evs.append(e)
        e1.record()
        barrier()
        launches = L.cc_launch_count() - l0
        ms = maxrank(e0.elapsed_time(e1))
        kt = {"fwd": 0.0, "mse": 0.0, "bwd": 0.0}
        if with_events:
            for e in evs:
                kt["fwd"] += e[0].elapsed_time(e[1]) / steps
                kt["mse"] += e[1].elapsed_time(e[2]) / steps
                kt["bwd"] += e[2].elapsed_time(e[3]) / steps
        return ms / steps, kt, launches, float(gl[-1].item())

    closed, B, T, ckpt, desc = WORKLOADS[args.workload]
    W = Workload(args.workload, B, world, rank, dev, torch)

    # ---- device-resident throughput (weak scaling: B per GPU fixed) -------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(warm):
        W.step(W.x)
    barrier()
    sampler.mark_start()
    ms_step, kt, launches, loss_val = timed(W, args.steps, with_events=True)
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None

    # ---- the same step replayed from a CUDA graph (launch gaps removed); reported beside the eager number -------------
    graph_ms = None
    if os.environ.get("BENCH_GRAPH", "1") != "0" and world == 1:
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(2):
                    W.step(W.x)
            torch.cuda.current_stream(dev).wait_stream(side)
            barrier()
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph):
                W.step(W.x)
            for _ in range(3):
                gph.replay()
            barrier()
            g0, g1 = ev(), ev()
            g0.record()
            for _ in range(args.steps):
                gph.replay()
            g1.record()
            barrier()
            graph_ms = g0.elapsed_time(g1) / args.steps
            del gph
        except Exception as exc:  # capture is an optimisation of the launch path only
            graph_ms = None
            print(f"[bench] CUDA graph capture skipped: {type(exc).__name__}: {exc}", file=sys.stderr)
            torch.cuda.synchronize()

    # ---- strong scaling: configs[3]'s 65,536 trajectories IN TOTAL, sharded B/N per GPU -----------------------------
    Bs = 65536 // world
    Ws = W if Bs == B else Workload(args.workload, Bs, world, rank, dev, torch)
    ms_s, _, _, _ = timed(Ws, args.steps)
    strong = {"B_total": Bs * world, "B_per_gpu": Bs, "ms_per_step": ms_s, "value": Bs * world * T / (ms_s * 1e-3), "unit": "traj-steps/s",
              "note": "total work fixed (configs[3]: 65,536 trajectories sharded B/N); one 128-trajectory tile per CTA: 65,536 / N trajectories fill "
                      "512 / N of the 296 CTA slots, and one tile's 1001-step chain is the floor"}
    if Ws is not W:
        del Ws

    # ---- end to end through the public API with HOST buffers ---------------------------------------
    # every step copies its own inputs host -> device (pinned memory) and reads its result back; the copy of step i+1
    # runs on a side stream while the kernels of step i execute (chain_control_b200.parallel.HostFeed, double buffered)
    P = W.P
    hres = torch.empty(P + 1).pin_memory()
    feed = parallel.HostFeed(dev)
    # warm-up THROUGH the feed (side stream, both device slots and the pinned result buffer exist before the timed region)
    feed.push(W.host_in)
    for i in range(3):
        xd = feed.pop()
        feed.push(W.host_in)
        gl2 = W.step(xd)
        feed.release()
        hres.copy_(gl2, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    feed.pop()
    barrier()
    e0, e1 = ev(), ev()
    step_t = []
    e0.record()
    feed.push(W.host_in)                               # H2D of step 0's inputs: inside the timed region
    for i in range(args.steps):
        th0 = time.perf_counter()
        xd = feed.pop()
        if i + 1 < args.steps:
            feed.push(W.host_in)                       # H2D of the next step's inputs, overlapped with this step's kernels
        gl2 = W.step(xd)
        feed.release()
        hres.copy_(gl2, non_blocking=True)             # D2H of the step's result (gradient + loss)
        torch.cuda.current_stream().synchronize()
        step_t.append(1e3 * (time.perf_counter() - th0))
    e1.record()
    barrier()
    ms_e2e = maxrank(e0.elapsed_time(e1)) / args.steps
    # host -> device bandwidth of this box for the same buffer (explains e2e once the kernels are faster than the copy)
    c0, c1 = ev(), ev()
    dst = torch.empty_like(W.x)
    dst.copy_(W.host_in, non_blocking=True)
    torch.cuda.synchronize()
    c0.record()
    dst.copy_(W.host_in, non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    h2d_gbs = W.host_in.numel() * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9
    del dst

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

    # ---- the other configurations of BASELINE.json, literally (N = 1 only) ---------------------------
    others = None
    if world == 1 and not args.no_others:
        others = {}
        cores = host_threads()
        for name in ("cfg1", "cfg2", "cfg4", "cfg5"):
            if name == args.workload:
                continue
            c2, B2, T2, ck2, d2 = WORKLOADS[name]
            try:
                Wo = Workload(name, B2, 1, 0, dev, torch)
                ms_o, kt_o, _, _ = timed(Wo, max(3, min(args.steps, 5)), with_events=True)
                entry = {"workload": d2, "B": B2, "T": T2, "ckpt_every": ck2, "ms_per_step": ms_o, "value": B2 * T2 / (ms_o * 1e-3),
                         "fwd_ms": kt_o["fwd"], "bwd_ms": kt_o["bwd"]}
                if not args.no_cpu_baseline:
                    Bc = min(B2, 256)  # CPU: the literal batch where it is small, else a 256-trajectory sample scaled to the batch
                    dtc = cpu_run(c2, Bc, T2, cores, 2 if B2 <= 64 else 1)
                    entry["cpu_ms_per_step"] = 1e3 * dtc * (B2 / Bc)
                    entry["cpu_sample"] = f"{Bc} trajectories x {T2} steps on {cores} threads" + ("" if Bc == B2 else f", scaled x{B2 // Bc}")
                others[name] = entry
                del Wo
            except Exception as exc:
                others[name] = {"error": f"{type(exc).__name__}: {exc}"}
            torch.cuda.empty_cache()
        # configs[2]: forward-only batched rollout of deep-MLP models, widths 32..128 (deep-MLP tcgen05 family, one MMA batch per
        # layer).  Three sweep points; the full B = 1K..1M sweep is profiles/r02_cfg3_sweep.md (scratch/sweep_cfg3_r02.py)
        try:
            others["cfg3"] = cfg3_points(torch, dev, peak_tflops, None if args.no_cpu_baseline else cores)
        except Exception as exc:
            others["cfg3"] = {"error": f"{type(exc).__name__}: {exc}"}
        torch.cuda.empty_cache()
        # the two-segment data source (cc_env_two_segments_kernel, fp64): 16,384 episodes of 1000 control steps
        try:
            others["data_source"] = env_point(torch, dev, None if args.no_cpu_baseline else cores)
        except Exception as exc:
            others["data_source"] = {"error": f"{type(exc).__name__}: {exc}"}

    if rank == 0:
        import ctypes as _C

        fwd_f, tot_f = alg_flops(closed)
        units = B * T  # trajectory-steps per launch
        fwd_ms, bwd_ms = kt["fwd"], kt["bwd"]
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            peaks = {}
        hbm = peaks.get("hbm_gbs", 6650.0)
        cm, mm = W.ctrl.to_c(), W.model.to_c()
        fam_f = L.cc_kernel_family_at(_C.byref(cm) if closed else None, _C.byref(mm), 0, B)
        fam_b = L.cc_kernel_family_at(_C.byref(cm) if closed else None, _C.byref(mm), 1, B)
        fam_name = {0: "ffma", 1: "tcgen05", 2: "latency", 3: "tcgen05", 4: "blocked-linear tcgen05"}
        kern_b = {4: "cc_lin_cl_bwd_kernel" if closed else "cc_linw_bwd_kernel", 1: "cc_tc_bwd_kernel", 3: "cc_tcw_bwd_kernel",
                  2: "cc_wpt_bwd_kernel", 0: "cc_rollout_bwd_kernel"}[fam_b]
        kern_f = {4: "cc_lin_cl_fwd_kernel" if closed else "cc_lin_ol_fwd_kernel", 1: "cc_tc_fwd_kernel", 2: "cc_wpt_fwd_kernel",
                  0: "cc_rollout_fwd_kernel"}[fam_f]
        bf16_peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 2250.0 * 0.62))
        tf32_peak = bf16_peak / 2.0
        fb, rb = alg_bytes(closed)
        # ---- roofline of the dominant kernel (the reverse pass) ----------------------------------------------------
        # The blocked linear kernels advance the linear model 13 steps per MMA batch, so the tensor pipe executes ~1.9 kFLOP
        # per trajectory-step for 21 kFLOP algorithmic: neither FLOP count bounds them any more.  What is left per step is a
        # scalar recurrence per trajectory (controller, input transform, Markov sums) fed from and writing to HBM: the next
        # roofline above the kernel is HBM, so `achieved` = ALGORITHMIC bytes / launch time against the measured copy bandwidth.
        traffic = ncu_traffic(kern_b)
        ach_b = units * rb / (bwd_ms * 1e-3) / 1e9
        roofline = {
            "bound": "hbm", "kernel": kern_b, "achieved": ach_b, "peak": hbm, "unit": "GB/s", "frac": ach_b / hbm,
            "traffic": traffic[0] if traffic else None,
            "traffic_source": (f"profiles/{traffic[1]}: dram__bytes_read.sum + dram__bytes_write.sum of one launch ({units} trajectory-steps)" if traffic else None),
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy bandwidth measured on this pool's B200s)",
            "alg_bytes_per_traj_step": rb, "launch_ms": bwd_ms,
            "fwd_kernel": {"kernel": kern_f, "achieved": units * fb / (fwd_ms * 1e-3) / 1e9, "frac": units * fb / (fwd_ms * 1e-3) / 1e9 / hbm,
                           "alg_bytes_per_traj_step": fb, "launch_ms": fwd_ms},
            "tensor": {"note": "the same launches against the FLOP yardsticks: algorithmic FLOPs (SURVEY 8d: 2(n_st M_f + M_g) per trajectory-step; "
                               "collapsing RK4 and blocking 13 steps means far fewer are executed) over dense TF32 = MEASURED_PEAKS bf16 sustained / 2 "
                               "and over the FP32-FMA peak measured live (BASELINE.json's yardstick)",
                       "alg_flops_per_traj_step": {"fwd": fwd_f, "bwd": tot_f - fwd_f},
                       "bwd_alg_tflops": units * (tot_f - fwd_f) / (bwd_ms * 1e-3) / 1e12, "fwd_alg_tflops": units * fwd_f / (fwd_ms * 1e-3) / 1e12,
                       "tf32_peak": tf32_peak, "bwd_frac_tf32": units * (tot_f - fwd_f) / (bwd_ms * 1e-3) / 1e12 / tf32_peak,
                       "fp32_fma_peak": peak_tflops, "step_frac_fp32_fma": units * tot_f / ((fwd_ms + bwd_ms) * 1e-3) / 1e12 / peak_tflops,
                       "executed_flops_per_traj_step": 3 * 2 * 64 * 64 / 13.0 if fam_b == 4 else None},
            "step": {"alg_bytes_per_traj_step": fb + rb + 12.0, "achieved_gbs": units * (fb + rb + 12.0) / (ms_step * 1e-3) / 1e9,
                     "note": "whole step incl. the MSE kernel (yhat + refs read, ybar written: 12 B)"},
        }
        out = {
            "metric": ("closed-loop " if closed else "") + "neural-ODE traj-steps/s (fwd+BPTT)",
            "value": world * units / (ms_step * 1e-3), "unit": "traj-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "B_per_gpu": B, "T": T, "closed_loop": closed, "integrator": "RK4", "ckpt_every": ckpt,
                       "kernels": {"forward": fam_name[fam_f], "reverse": fam_name[fam_b],
                                   "note": "blocked-linear = 13 integrator steps of the linear model per 3xTF32 tcgen05.mma batch (csrc/cc_lin.h); "
                                           "tcgen05 = one batch per RK4 stage; latency = one warp per trajectory; ffma = register-tiled FP32"},
                       "step": "fwd rollout + mse cotangent + BPTT + grad reduce" + (" + NCCL allreduce(P+1)" if world > 1 else "") + " + fused clip + adam (cc_opt_step)",
                       "step_ms": {"fwd": fwd_ms, "mse": kt["mse"], "bwd": bwd_ms},
                       "cuda_graph_ms_per_step": graph_ms,
                       "l2": "inputs + tape (0.26 GB + 1.6 GB) exceed the 126 MB L2",
                       "loss": loss_val},
            "e2e": {"value": world * units / (ms_e2e * 1e-3), "unit": "traj-steps/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(W.host_in.numel() * 4), "d2h_bytes_per_step": int((P + 1) * 4),
                    "h2d_gbs_this_box": h2d_gbs, "host_numa_binding": numa,
                    "host_step_ms": {"min": min(step_t), "median": statistics.median(step_t), "max": max(step_t),
                                     "note": "host wall clock of each timed step on rank 0 (diagnostic: a one-off stall shows as max >> median; "
                                             "value is total device time / steps as the contract says)"},
                    "note": "the input copy of step i+1 overlaps the kernels of step i; once the step is shorter than the copy, e2e is the PCIe rate"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "strong": strong,
        }
        if others is not None:
            out["config"]["other_workloads"] = others
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(closed, T)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
