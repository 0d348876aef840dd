"""CPU, world_size 2 over gloo: the multi-GPU path is GOP sharding with no data-path collective; the only
communication is the max-over-ranks of the step time.  This exercises that plumbing without a GPU."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import h264_b200_pkg  # noqa: F401
    from h264_b200 import synth
    # every rank owns a different GOP: different clip seed -> different pictures, same job layout
    frames = synth.luma_frames(176, 144, 2, seed=1234 + 17 * rank)
    jobs = synth.synthetic_jobs(176, 144, 1, (1, 1, 1), seed=21 + rank)
    t = torch.tensor([10.0 + rank], dtype=torch.float64)          # pretend step time of this rank
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    csum = torch.tensor([float(frames[0].sum() % 1000003)], dtype=torch.float64)
    gathered = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, csum)
    q.put((rank, float(t.item()), [float(g.item()) for g in gathered], len(jobs)))
    dist.barrier()
    dist.destroy_process_group()


def test_gop_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == res[1][1] == 11.0                      # max over ranks
    assert res[0][2] == res[1][2] and res[0][2][0] != res[0][2][1]   # every rank saw both, and the GOPs differ
    assert res[0][3] == res[1][3] == 99
