"""Data-parallel sharding of the hot path across the GPUs of one box.

Trajectories never interact in the forward pass (eqx.filter_vmap, cc/train/step_fn.py:136,257); the only
coupling is the sum over the batch inside the mean loss / gradient (cc/utils/utils.py:53-55).  Each rank takes
a contiguous slice of the batch, scales its cotangent by the GLOBAL element count (so no communication before
the backward pass), and one allreduce(sum) of the P+1 vector [gradient ; loss partial] per step makes every
rank apply the identical optimiser update (no parameter broadcast).  One process per GPU, NCCL over
NVLink/NVSwitch (gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def bind_host_to_gpu_numa(device_index: int):
    """One process per GPU: pin this process to the CPU cores of the NUMA node its GPU hangs off, BEFORE pinned host buffers
    are allocated (first-touch places them on that node), so that host -> device copies do not cross the socket interconnect.
    Reads /sys/bus/pci/devices/<bus id>/numa_node and /sys/devices/system/node/node<N>/cpulist; a no-op (returns None) where
    that information is missing (single-socket boxes, containers without sysfs).  CC_B200_NUMA_BIND=0 disables it."""
    import os

    if os.environ.get("CC_B200_NUMA_BIND", "1") == "0":
        return None
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus": len(allowed), "pci": bus}
    except Exception:  # pragma: no cover - topology files differ between boxes
        return None


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(B: int, rank: int, world_size: int):
    """contiguous split of the batch axis; the first (B % world) ranks get one extra trajectory"""
    base, rem = divmod(B, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tree, rank: int, world_size: int):
    from .utils import tree_map, tree_shape

    lo, hi = shard_bounds(tree_shape(tree), rank, world_size)
    return tree_map(lambda a: a[lo:hi], tree)


def allreduce_grad_and_loss(grad: torch.Tensor, loss_partial: torch.Tensor):
    """sum over ranks of [theta_bar ; loss partial] in ONE collective of P+1 floats"""
    buf = torch.cat([grad.reshape(-1), loss_partial.reshape(-1)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return buf[:-1].reshape(grad.shape), buf[-1]


def dp_shard(tree):
    """this rank's contiguous slice of a batched pytree and its share B_rank / B of the batch (1.0 without a process group)"""
    from .utils import tree_shape

    rank, ws = world()
    if ws == 1:
        return tree, 1.0
    B = tree_shape(tree)
    lo, hi = shard_bounds(B, rank, ws)
    return shard_batch(tree, rank, ws), (hi - lo) / B


def dp_combine(grad: torch.Tensor, logs: dict, frac: float):
    """data-parallel train step of the trainers (cc/train/step_fn.py:150-179,287-322 on a sharded minibatch): every rank holds
    the gradient and the scalar logs of the MEAN loss over its shard; the batch mean is sum_r (B_r / B) * local, done as ONE
    allreduce of [gradient ; scalar logs].  Terms that do not depend on the data (regularisers) are identical on all ranks
    and survive the weighted sum unchanged."""
    rank, ws = world()
    if ws == 1:
        return grad, logs
    keys = sorted(logs)
    buf = torch.cat([grad.reshape(-1)] + [torch.as_tensor(logs[k], dtype=grad.dtype, device=grad.device).reshape(1) for k in keys]) * frac
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    n = grad.numel()
    return buf[:n].reshape(grad.shape), {k: buf[n + i] for i, k in enumerate(keys)}


class HostFeed:
    """Double-buffered host -> device feed for the train step's inputs: the copy of step i+1 (pinned host memory, side
    stream) overlaps the rollout kernels of step i.  Every step still moves its own inputs host -> device; only the
    waiting is hidden.  Usage:
        feed = HostFeed(device); feed.push(host_batch)            # before the loop
        for ...:  x = feed.pop(); feed.push(next_host_batch); step(x)
    The device buffers are reused (two slots), so `pop()`ed tensors are valid until the second-next `push`.
    Slot reuse is ordered automatically: `push` makes its copy wait for everything queued on the compute stream so far,
    i.e. for the step that read the slot two pushes ago (in the loop above that step is already queued when `push` runs, and
    the step that overlaps the copy -- step(x) -- is queued after it)."""

    def __init__(self, device):
        import torch

        self.torch = torch
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.slots = [None, None]
        self.ready = [None, None]
        self.free = [None, None]   # event recorded on the compute stream when the slot's previous consumer was queued
        self.head = 0              # next slot to fill
        self.queue = []

    def push(self, host_tensor):
        torch = self.torch
        s = self.head
        self.head ^= 1
        if self.slots[s] is None or self.slots[s].shape != host_tensor.shape or self.slots[s].dtype != host_tensor.dtype:
            self.slots[s] = torch.empty(host_tensor.shape, dtype=host_tensor.dtype, device=self.device)
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(self.device))  # every consumer of this slot queued so far
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(done)
            if self.free[s] is not None:
                self.stream.wait_event(self.free[s])
            self.slots[s].copy_(host_tensor, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.ready[s] = ev
        self.queue.append(s)

    def pop(self):
        s = self.queue.pop(0)
        cur = self.torch.cuda.current_stream(self.device)
        cur.wait_event(self.ready[s])
        self._last = s
        return self.slots[s]

    def release(self):
        """optional (push orders slot reuse by itself): marks the last `pop()`ed slot free once the work queued on the
        current stream so far has finished, for consumers running on OTHER streams than the one `push` is called from"""
        ev = self.torch.cuda.Event()
        ev.record(self.torch.cuda.current_stream(self.device))
        self.free[self._last] = ev
