"""CPU: the multi-GPU path is GOP sharding with no data-path collective.  Covered here without a GPU:
  * the assignment rule of h.264_b200/gop_shard.py (rank -> GOPs, gather back into GOP order);
  * world_size 2 over gloo: every rank searches the job arrays of ITS GOPs (the oracle stands in for the device), the per-rank
    results are gathered on rank 0 and must equal the single-process results of all GOPs in GOP order;
  * the real-encoder driver with the plain reference binary: one process per GOP on StartFrame ranges (configfile.h:49,57),
    per-GOP bitstreams concatenated in GOP order whatever order the processes finish in."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, has_ref_binaries

W, H, N_GOPS, JOBS_PER_GOP = 176, 144, 5, 6


def _gop_inputs(g):
    """pictures and a few real job arrays of GOP g (its own clip: different seed)"""
    import h264_b200_pkg  # noqa: F401
    from h264_b200 import synth
    frames = synth.luma_frames(W, H, 2, seed=1234 + 17 * g)
    jobs = synth.synthetic_jobs(W, H, 1, (334459,) * 3, seed=21 + g)[7 * g:7 * g + JOBS_PER_GOP]
    return frames, jobs


def _search_gop(g):
    from oracle import binding as ob
    frames, jobs = _gop_inputs(g)
    ref = ob.Ref(frames[0])
    out, out_int = ob.mb_search(ob.make_cfg(), np.ascontiguousarray(frames[1]), ref, jobs)
    ref.close()
    return out.tobytes() + out_int.tobytes()


def test_assignment_rule_round_trips():
    import h264_b200_pkg  # noqa: F401
    from h264_b200 import gop_shard as gs
    for world in (1, 2, 3, 8):
        for n in (0, 1, 5, 8, 17):
            seen = sorted(i for r in range(world) for i in gs.units_of_rank(n, r, world))
            assert seen == list(range(n))
            units = ["gop%d" % i for i in range(n)]
            per_rank = [gs.shard_units(units, r, world) for r in range(world)]
            assert max(map(len, per_rank)) - min(map(len, per_rank)) <= 1          # balanced
            assert gs.gather_units(per_rank, n) == units
    with pytest.raises(ValueError):
        gs.gather_units([["a"], []], 3)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import h264_b200_pkg  # noqa: F401
    from h264_b200 import gop_shard as gs
    mine = [_search_gop(g) for g in gs.units_of_rank(N_GOPS, rank, world)]     # no collective on the data path
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)                                  # results gathered on the host of rank 0
    if rank == 0:
        q.put(gs.gather_units(gathered, N_GOPS))
    dist.barrier()
    dist.destroy_process_group()


def test_gop_sharding_world2_equals_single_process(oracle_lib):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    sharded = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = [_search_gop(g) for g in range(N_GOPS)]
    assert sharded == single
    assert len(set(single)) == N_GOPS                                           # the GOPs really differ


@pytest.mark.skipif(not has_ref_binaries(), reason="oracle/_ref not built (needs the reference tree at build time)")
def test_encoder_gop_driver_with_the_reference_binary(tmp_path):
    """run_encoder_gops() with exe = lencod: three closed GOPs of two pictures on two lanes; every per-GOP stream equals a direct run
    of the encoder on that StartFrame range and the gathered stream is their concatenation in GOP order"""
    import subprocess
    import h264_b200_pkg  # noqa: F401
    from h264_b200 import gop_shard as gs, synth
    yuv = str(tmp_path / "clip.yuv")
    synth.write_yuv420(yuv, W, H, 6)
    extra = ["-p", "SearchRange=16", "-p", "NumberReferenceFrames=2", "-p", "SearchMode=0"]
    res = gs.run_encoder_gops(yuv, W, H, 3, 2, 2, extra=extra, exe="lencod", workdir=str(tmp_path / "w"), max_parallel=2)
    assert [g["gop"] for g in res["gops"]] == [0, 1, 2] and all(g["me_s"] is not None for g in res["gops"])
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    cat = b""
    for g in range(3):
        d = tmp_path / ("direct%d" % g)
        d.mkdir()
        for f in os.listdir(os.path.join(ref_dir, "fixtures")):
            if f.endswith(".cfg"):
                (d / f).write_bytes(open(os.path.join(ref_dir, "fixtures", f), "rb").read())
        r = subprocess.run([os.path.join(ref_dir, "lencod")] + gs.encoder_args("encoder_main.cfg", yuv, W, H, 2 * g, 2, extra), cwd=str(d),
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-800:]
        one = (d / "gop.264").read_bytes()
        assert one == open(res["gops"][g]["path"], "rb").read()
        cat += one
    assert cat == open(res["stream"], "rb").read()
    assert len({g["md5"] for g in res["gops"]}) == 3                              # different pictures, different streams
