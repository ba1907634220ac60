"""world_size-2 run of the sharded path on CPU (gloo): each rank piles its -J shard, runs the (emulated) kernel,
and rank 0 checks that the concatenation of the shard outputs equals the unsharded run (no data-path collective;
the only exchange is the one-time broadcast of the packed read database)."""
import os
import sys
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, q):
    sys.path.insert(0, HERE); sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from common import default_params, run_emu
    from daccord_b200.host import Dataset
    from bench import j_shard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ds = Dataset.simulate(8000, read_len=1500, coverage=14, seed=6)
    pi, pd, cor = ds.profile()
    p = default_params(p_i=pi, p_d=pd, est_cor=cor)
    # one-time broadcast of the packed DB from rank 0 (NCCL on the GPU box, gloo here)
    packed = torch.from_numpy(np.array(ds.packed(), copy=True)) if rank == 0 else torch.zeros(len(ds.packed()), dtype=torch.uint8)
    dist.broadcast(packed, src=0)
    assert (packed.numpy() == ds.packed()).all()
    lo, hi = j_shard(ds.nreads, rank, world)
    b = ds.pile(lo, hi, nthreads=2)
    r = run_emu(p, packed.numpy(), b.win.copy(), b.sl.copy(), 1)
    fa, n = b.vote(r[0], r[1], r[2])
    cnt = torch.tensor([len(b.win), int((r[0]["status"] == 1).sum())], dtype=torch.int64)
    dist.all_reduce(cnt)
    gathered = [None] * world
    dist.all_gather_object(gathered, fa)
    if rank == 0:
        ball = ds.pile(nthreads=2)
        rall = run_emu(p, ds.packed(), ball.win.copy(), ball.sl.copy(), 1)
        fall, _ = ball.vote(rall[0], rall[1], rall[2])
        strip = lambda t: [l if not l.startswith(b">") else l.split(b"/")[0] + b"/" + l.split(b"/", 2)[2] for l in t.split(b"\n")]
        ok = strip(b"".join(gathered)) == strip(fall) and int(cnt[0]) == len(ball.win) and int(cnt[1]) == int((rall[0]["status"] == 1).sum())
        q.put(ok)
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(300)
        assert pr.exitcode == 0
    assert q.get(timeout=5) is True
