"""N > 1 path on CPU: two gloo ranks each run the (emulated) kernels on their shard of the batch, allreduce the
P+1 gradient/loss vector, and must reproduce the single-process full-batch result of the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, out):
    sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import emu_lib as E
    from cchelpers import conv
    from ccspecs import controller_variants, model_variants
    from chain_control_b200 import parallel
    from oracle import np_oracle as O

    mspec, cspec = model_variants()["compact_s50_d0"], controller_variants()["compact_s5_d0"]
    thc, thm = O.init_params(cspec, 2).astype(np.float32), O.init_params(mspec, 1, stable=True).astype(np.float32)
    B, Tr = 10, 12  # uneven shards: 5 + 5 at world 2, exercised with B = 11 below too
    for B in (10, 11):
        refs = np.random.default_rng(0).uniform(-2, 2, size=(B, Tr, 1)).astype(np.float32)
        lo, hi = parallel.shard_bounds(B, rank, world)
        mine = refs[lo:hi]
        n_total = B * Tr * 1
        yhat, _, _ = E.np_closedloop_fwd(E.emu(), conv(cspec), conv(mspec), thc, thm, mine)
        ybar = (2.0 / n_total) * (yhat - mine)  # local cotangent with the GLOBAL count
        _, g = E.np_closedloop_bwd(E.emu(), conv(cspec), conv(mspec), thc, thm, mine, ybar)
        loss_part = np.float32(((yhat - mine) ** 2).sum() / n_total)
        g_all, loss = parallel.allreduce_grad_and_loss(torch.tensor(g), torch.tensor([loss_part]))
        if rank == 0:
            y64, _ = O.closedloop_fwd(cspec, mspec, thc, thm, refs)
            ref_g = O.closedloop_bwd(cspec, mspec, thc, thm, refs, O.mse_cotangent(refs, y64))
            out.put((B, float(np.linalg.norm(g_all.numpy() - ref_g) / np.linalg.norm(ref_g)), abs(float(loss) - O.mse(refs, y64))))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_gradient_matches_full_batch():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=600) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for B, gerr, lerr in res:
        assert gerr < 1e-4, (B, gerr)
        assert lerr < 1e-6, (B, lerr)


def test_shard_bounds_cover_the_batch():
    from chain_control_b200.parallel import shard_bounds

    for B in (1, 7, 64, 65536, 65537):
        for W in (1, 2, 4, 8):
            spans = [shard_bounds(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
