"""Data-parallel ModelControllerTrainer on two GPUs (one process per GPU, NCCL): the sharded step must reproduce the
single-process step (same parameters after the update, same logs).  Needs 2 CUDA devices: skipped on a one-GPU box."""
import os
import sys
from collections import OrderedDict

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(rank, world, port, out):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, os.path.dirname(HERE))
    torch.cuda.set_device(rank)
    if world > 1:
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("nccl", rank=rank, world_size=world)
    from chain_control_b200.examples.neural_ode_controller_compact_example import make_neural_ode_controller
    from chain_control_b200.examples.neural_ode_model_compact_example import make_neural_ode_model
    from chain_control_b200.train import (ModelControllerTrainer, SupervisedDataset, TrainingOptionsController, TrainingOptionsModel,
                                          UnsupervisedDataset, make_dataloader, optim)

    dev = f"cuda:{rank}"
    rng = np.random.default_rng(0)
    B, T = 11, 120  # uneven shards: 6 + 5
    u = rng.uniform(-1, 1, size=(B, T, 1)).astype(np.float32)
    teacher = make_neural_ode_model(1, OrderedDict(y=1), 0.01, 4, f_depth=0, key=7, u_transform=np.arctan, device=dev)
    with torch.no_grad():
        y = teacher.unroll_batched(u)
    model = make_neural_ode_model(1, OrderedDict(y=1), 0.01, 3, f_depth=0, key=3, u_transform=np.arctan, device=dev)
    mo = TrainingOptionsModel(make_dataloader(SupervisedDataset(u, y), key=2, n_minibatches=1),
                              optim.chain(optim.clip_by_global_norm(1.0), optim.adam(3e-3)))
    tr = ModelControllerTrainer(model, model_train_options=mo)
    logs_m = [tr.step() for _ in range(2)]
    theta_m = tr.model.flat_params().detach().cpu().numpy() if hasattr(tr, "model") else None
    refs = OrderedDict(y=(rng.uniform(-1, 1, size=(B, 1, 1)) * np.ones((B, T, 1))).astype(np.float32))
    ctrl = make_neural_ode_controller(2, 1, 0.01, 5, f_depth=0, key=5, device=dev)
    co = TrainingOptionsController(make_dataloader(UnsupervisedDataset(refs), key=1, n_minibatches=1),
                                   optim.chain(optim.clip_by_global_norm(1.0), optim.adam(1e-3)))
    tc = ModelControllerTrainer(teacher, ctrl, controller_train_options=co)
    logs_c = [tc.step() for _ in range(2)]
    if rank == 0:
        out.put((world, {k: float(v) for k, v in logs_m[-1].items()}, {k: float(v) for k, v in logs_c[-1].items()}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def test_two_gpu_trainer_matches_single_process():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    res = {}
    for world in (1, 2):
        port = 29700 + (os.getpid() % 2000) + world
        procs = [ctx.Process(target=_run, args=(r, world, port, out)) for r in range(world)]
        for p in procs:
            p.start()
        w, lm, lc = out.get(timeout=600)
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
        res[w] = (lm, lc)
    for a, b in zip(res[1], res[2]):
        for k in a:
            assert abs(a[k] - b[k]) <= 2e-5 * max(1.0, abs(a[k])), (k, a[k], b[k])
