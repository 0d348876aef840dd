"""GOP sharding of the motion-estimation path across the GPUs of one box (BASELINE.json config 5).

The path has no exchange step: closed GOPs are independent units.  Two users of the same assignment rule:

  * bench.py / the frame batch: rank r of W ranks owns the GOPs r, r + W, r + 2W, ...; results are gathered on the host
    (shard_units / gather_units below, used with torch.distributed only for the gather of the per-rank arrays).
  * the real encoder: one `lencod_cuda` process per GOP (the reference's own StartFrame / FramesToBeEncoded parameters,
    lencod/inc/configfile.h:49,57, cut the input clip into closed GOPs), at most one process per GPU at a time, device chosen
    with ME_CUDA_DEVICE; the per-GOP bitstreams are concatenated in GOP order on the host.  run_encoder_gops() does that and can
    check every GOP against the plain reference encoder.

    python -m h264_b200.gop_shard is not available (the package directory is not an identifier); run
    python h.264_b200/gop_shard.py --gpus 8 --gops 8 --gop-frames 4 --width 3840 --height 2160 --search-range 64 --compare
"""
import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


# ---------------------------------------------------------------------------------------------- assignment rule
def units_of_rank(n_units, rank, world):
    """indices of the units (GOPs) rank `rank` of `world` owns: round robin, so that any prefix of the GOPs is balanced"""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("rank %d of %d" % (rank, world))
    return list(range(rank, n_units, world))


def shard_units(units, rank, world):
    """the sub-list of `units` (one entry per GOP, e.g. its job array) this rank works on"""
    return [units[i] for i in units_of_rank(len(units), rank, world)]


def gather_units(per_rank, n_units):
    """inverse of shard_units: per_rank[r] = results of rank r in its own order -> results in GOP order"""
    world = len(per_rank)
    out = [None] * n_units
    for r in range(world):
        idx = units_of_rank(n_units, r, world)
        if len(idx) != len(per_rank[r]):
            raise ValueError("rank %d returned %d results for %d GOPs" % (r, len(per_rank[r]), len(idx)))
        for i, res in zip(idx, per_rank[r]):
            out[i] = res
    return out


# ---------------------------------------------------------------------------------------------- the real encoder
def _md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def _me_seconds(stdout):
    for ln in stdout.splitlines():
        if "Total ME time for sequence" in ln:
            return float(ln.split(":")[1].split("sec")[0])
    return None


def encoder_args(cfg, yuv, width, height, start_frame, frames, extra):
    return ["-d", cfg, "-p", "InputFile=" + yuv, "-p", "SourceWidth=%d" % width, "-p", "SourceHeight=%d" % height,
            "-p", "StartFrame=%d" % start_frame, "-p", "FramesToBeEncoded=%d" % frames, "-p", "FrameSkip=0",
            "-p", "NumberBFrames=0", "-p", "OutputFile=gop.264", "-p", "ReconFile=gop_rec.yuv"] + list(extra)


def run_encoder_gops(yuv, width, height, n_gops, gop_frames, n_gpus, cfg="encoder_main.cfg", extra=(), exe="lencod_cuda",
                     workdir=None, ref_dir=None, lib=None, max_parallel=None):
    """Encodes GOP g = frames [g * gop_frames, (g + 1) * gop_frames) of `yuv` with one encoder process per GOP.  exe = "lencod_cuda":
    GPU g % n_gpus (ME_CUDA_DEVICE), at most one process per GPU at a time; exe = "lencod": the plain reference on the host cores.
    Returns dict(gops=[{md5, me_s, wall_s, device, path}], wall_s, stream=path of the concatenated bitstream)."""
    ref_dir = ref_dir or os.path.join(ROOT, "oracle", "_ref")
    binary = os.path.join(ref_dir, exe)
    if not os.path.exists(binary):
        raise RuntimeError("%s not built (make -C oracle ref, where the reference tree is mounted)" % binary)
    workdir = workdir or tempfile.mkdtemp(prefix="gops_")
    env = dict(os.environ)
    env["ME_CUDA_LIB"] = lib or os.path.join(HERE, "libme_cuda.so")
    for k in ("ME_CUDA_DISABLE", "ME_CUDA_VERIFY"):
        env.pop(k, None)
    lanes = n_gpus if exe == "lencod_cuda" else (max_parallel or os.cpu_count() or 1)
    pending = list(range(n_gops))
    running = {}               # lane -> (gop, Popen, t0, dir)
    done = [None] * n_gops
    t_start = time.perf_counter()
    while pending or running:
        for lane in range(lanes):
            if lane not in running and pending:
                # a GPU lane takes the next GOP it owns first (the static round-robin rule), any other when it has none left
                mine = [g for g in pending if g % lanes == lane]
                g = (mine or pending)[0]
                pending.remove(g)
                d = os.path.join(workdir, "%s_gop%d" % (exe, g))
                os.makedirs(d, exist_ok=True)
                for f in os.listdir(os.path.join(ref_dir, "fixtures")):
                    if f.endswith(".cfg"):
                        shutil.copy(os.path.join(ref_dir, "fixtures", f), d)
                e = dict(env)
                if exe == "lencod_cuda":
                    e["ME_CUDA_DEVICE"] = str(lane)
                p = subprocess.Popen([binary] + encoder_args(cfg, os.path.abspath(yuv), width, height, g * gop_frames, gop_frames, extra),
                                     cwd=d, env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
                running[lane] = (g, p, time.perf_counter(), d)
        for lane, (g, p, t0, d) in list(running.items()):
            if p.poll() is None:
                continue
            out, err = p.communicate()
            if p.returncode != 0:
                for _, q, _, _ in running.values():
                    if q.poll() is None:
                        q.kill()
                raise RuntimeError("%s failed on GOP %d: %s" % (exe, g, (out[-800:] + err[-800:])))
            done[g] = dict(gop=g, device=lane if exe == "lencod_cuda" else None, wall_s=time.perf_counter() - t0, me_s=_me_seconds(out),
                           md5=_md5(os.path.join(d, "gop.264")), rec_md5=_md5(os.path.join(d, "gop_rec.yuv")),
                           path=os.path.join(d, "gop.264"), served=[ln for ln in err.splitlines() if "me_cuda shim" in ln])
            del running[lane]
        time.sleep(0.005)
    wall = time.perf_counter() - t_start
    stream = os.path.join(workdir, "%s_all.264" % exe)
    with open(stream, "wb") as f:                   # gathered on the host in GOP order
        for g in done:
            f.write(open(g["path"], "rb").read())
    return dict(gops=done, wall_s=wall, stream=stream)


def main(argv=None):
    ap = argparse.ArgumentParser(description="one lencod_cuda process per GOP and GPU; bitstreams gathered on the host")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--gops", type=int, default=8)
    ap.add_argument("--gop-frames", type=int, default=3)
    ap.add_argument("--width", type=int, default=352)
    ap.add_argument("--height", type=int, default=288)
    ap.add_argument("--search-range", type=int, default=32)
    ap.add_argument("--refs", type=int, default=5)
    ap.add_argument("--input", default=None, help="4:2:0 clip (default: the synthetic clip of h264_b200.synth)")
    ap.add_argument("--compare", action="store_true", help="also run the plain reference encoder per GOP and compare byte for byte")
    a = ap.parse_args(argv)
    sys.path.insert(0, ROOT)
    import h264_b200_pkg  # noqa: F401
    from h264_b200 import synth
    work = tempfile.mkdtemp(prefix="gopshard_")
    yuv = a.input
    if not yuv:
        yuv = os.path.join(work, "clip.yuv")
        synth.write_yuv420(yuv, a.width, a.height, a.gops * a.gop_frames)
    extra = ["-p", "SearchRange=%d" % a.search_range, "-p", "NumberReferenceFrames=%d" % a.refs, "-p", "LevelIDC=51", "-p", "SearchMode=0"]
    gpu = run_encoder_gops(yuv, a.width, a.height, a.gops, a.gop_frames, a.gpus, extra=extra, workdir=work)
    coded_h = (a.height + 15) // 16 * 16
    units = (a.width // 16) * (coded_h // 16) * sum(min(k, a.refs) for k in range(1, a.gop_frames)) * a.gops     # (macroblock, ref) searches
    line = {"what": "GOP-sharded lencod_cuda", "gpus": a.gpus, "gops": a.gops, "gop_frames": a.gop_frames, "size": [a.width, a.height],
            "search_range": a.search_range, "refs": a.refs, "wall_s": gpu["wall_s"],
            "me_s_per_gop": [g["me_s"] for g in gpu["gops"]], "devices": [g["device"] for g in gpu["gops"]],
            "encoder_me_mb_ref_per_s": units / max(1e-9, max(sum(g["me_s"] for g in gpu["gops"] if g["device"] == d) for d in range(a.gpus))),
            "stream_md5": _md5(gpu["stream"]), "service": gpu["gops"][0]["served"]}
    if a.compare:
        cpu = run_encoder_gops(yuv, a.width, a.height, a.gops, a.gop_frames, a.gpus, extra=extra, exe="lencod", workdir=work)
        line["identical_per_gop"] = [g["md5"] == c["md5"] and g["rec_md5"] == c["rec_md5"] for g, c in zip(gpu["gops"], cpu["gops"])]
        line["identical_stream"] = _md5(gpu["stream"]) == _md5(cpu["stream"])
        line["reference_wall_s"] = cpu["wall_s"]
        line["reference_me_s_per_gop"] = [c["me_s"] for c in cpu["gops"]]
    print(json.dumps(line))
    shutil.rmtree(work, ignore_errors=True)
    return 0 if (not a.compare or all(line["identical_per_gop"])) else 1


if __name__ == "__main__":
    sys.exit(main())
