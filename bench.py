#!/usr/bin/env python
"""bench.py -- throughput of the motion-estimation hot path (BASELINE.json metric: 1080p FS-ME + QPel macroblocks/s,
search range +-32, 5 references, all 41 partitions, SAD integer-pel + Hadamard-SATD sub-pel).

One "step" = the motion estimation of one 1920x1088 P picture: the quarter-pel planes of the newest reference picture
are generated (1 plane-generation launch pair) and every macroblock is searched in all 5 references
(8160 x 5 = 40800 (macroblock, ref) jobs: integer-pel fast full search + half/quarter-pel refinement).

  python bench.py --gpus N --steps K --warmup W            our arm  (one process per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...   the unmodified reference encoder on the host cores

value : macroblocks/s, whole job (N GPUs), inputs resident in HBM, CUDA-event timed per step, max over ranks.
e2e   : same metric through the public host API (me_cuda_ref_update_u8 / me_cuda_set_current_u8 / me_cuda_frame_search)
        with pinned HOST buffers: H2D of the picture, the new reference and the job array and D2H of the results are
        inside the timed region.
Multi-GPU: the path shards by independent GOPs: every rank runs its own clip, no collective on the data
path; value = N * 8160 / (max over ranks of the step time)  ("scaling": "weak").
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, N_REFS, SEARCH_RANGE = 1920, 1088, 5, 32   # the headline workload (BASELINE.json configs[2])
if os.environ.get("ME_BENCH_WORKLOAD", "") == "4k":   # side measurement of configs[4] (not the bench contract): 3840x2176, SR 64
    W, H, N_REFS, SEARCH_RANGE = 3840, 2176, 5, 64
MBS = (W // 16) * (H // 16)                      # 8160
LAMBDA = (334459, 334459, 334459)               # QP 28 P-slice lambda factors of the reference run (trace of bin/encoder.cfg)
P_POS = (2 * SEARCH_RANGE + 1) ** 2              # 4225 search positions
LANE_OPS_PER_JOB = 212 * P_POS                   # 64 VABSDIFF4 + 25 adds + 3*41 cost ops per position
SAD_OPS_PER_JOB = 64 * P_POS
METRIC = "1080p FS-ME+QPel macroblocks/s (SR32, 5 refs)"
WORKLOAD = "1080p_sr32_5ref_allpartitions_qpel" if SEARCH_RANGE == 32 else "4k_sr64_5ref_allpartitions_qpel (side measurement)"
UNIT = "MB/s"


def bench_config(world, sub_metric=2):
    """the `config` object of the JSON line: the same for both arms (the reference arm runs this workload on the host cores)"""
    return {"workload": WORKLOAD, "width": W, "height": H, "search_range": SEARCH_RANGE, "refs": N_REFS,
            "macroblocks_per_step": MBS, "jobs_per_step": MBS * N_REFS, "partitions": 41,
            "metrics": "SAD fpel / %s hpel,qpel" % ("SAD", "SSE", "SATD")[sub_metric],
            "parallelism": "gop-sharded x%d, no collective" % world,
            "l2": "256 MB write between timed steps (L2 flush) and ~250 MB working set per step",
            "predictors": "synthetic per-partition predictors (pan + jitter), seed 21",
            "host_transfer": "pinned job / result arrays in place" if host_transfer_mode(world) != "0"
                             else "pinned job / result arrays through four pipelined copy chunks (8 processes share the host)"}


def host_transfer_mode(world):
    """How me_cuda_frame_search moves the pinned caller arrays (library switch ME_CUDA_ZEROCOPY; the library does not choose by
    itself): "1" read and written in place by the kernels, "0" four pipelined copy chunks.  Measured on this pool's boxes, ms per
    step end to end at 1 / 2 / 4 / 8 processes per host (profiles/r02p_e2e_variants.txt): in place 2.60 / 2.65 / 2.76 / 3.14-3.18,
    chunked 2.94 / 2.95 / 2.92 / 2.98 -- kernel-issued PCIe accesses win until eight processes share the host's root complex."""
    return os.environ.get("ME_CUDA_ZEROCOPY") or ("0" if world >= 8 else "1")


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ---------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.lines, self.stamps, self.proc = gpu_index, [], [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())
            self.stamps.append(time.perf_counter())

    def stop(self, timed=None):
        """timed = (t0, t1) perf_counter bounds of the timed steps: the record says how many samples fall inside them"""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power, inside = [], [], set(), [], 0
        for ln, ts in zip(self.lines, self.stamps):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            if timed and timed[0] <= ts <= timed[1]:
                inside += 1
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "samples_in_timed_steps": inside,
                "sampled_over": "the warm-up loop (same step, same load, >= 2.5 s) and the timed steps; nvidia-smi every 100 ms",
                "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------- reference arm
def reference_cfg_args(yuv, frames, refs=N_REFS, sr=SEARCH_RANGE, w=W, h=H):
    return ["-d", "encoder_main.cfg", "-p", "InputFile=" + yuv, "-p", "SourceWidth=%d" % w, "-p", "SourceHeight=%d" % h,
            "-p", "FramesToBeEncoded=%d" % frames, "-p", "FrameSkip=0", "-p", "NumberBFrames=0", "-p", "SearchMode=0",
            "-p", "SearchRange=%d" % sr, "-p", "NumberReferenceFrames=%d" % refs, "-p", "RestrictSearchRange=2",
            "-p", "RDOptimization=1", "-p", "MEDistortionFPel=0", "-p", "MEDistortionHPel=2", "-p", "MEDistortionQPel=2",
            "-p", "LevelIDC=51", "-p", "QPISlice=28", "-p", "QPPSlice=28", "-p", "OutputFile=out.264", "-p", "ReconFile=rec.yuv"]


def run_lencod(workdir, frames, yuv="clip.yuv", h=H):
    """runs the unmodified reference encoder; returns (ME seconds, (macroblock, ref) units searched)"""
    exe = os.path.join(ROOT, "oracle", "_ref", "lencod")
    r = subprocess.run([exe] + reference_cfg_args(yuv, frames, h=h), cwd=workdir, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("lencod failed: " + r.stdout[-1500:] + r.stderr[-500:])
    me_s = None
    for ln in r.stdout.splitlines():
        if "Total ME time for sequence" in ln:
            me_s = float(ln.split(":")[1].split("sec")[0])
    # IPPP: P picture k (k = 1..frames-1) searches min(k, N_REFS) references
    units = (W // 16) * (h // 16) * sum(min(k, N_REFS) for k in range(1, frames))
    return me_s, units


def prepare_reference_dir(frames, h=H):
    import shutil
    import h264_b200_pkg  # noqa: F401
    from h264_b200 import synth
    d = tempfile.mkdtemp(prefix="h264ref_")
    fx = os.path.join(ROOT, "oracle", "_ref", "fixtures")
    for f in os.listdir(fx):
        if f.endswith(".cfg"):
            shutil.copy(os.path.join(fx, f), d)
    synth.write_yuv420(os.path.join(d, "clip.yuv"), W, h, frames)
    return d


def have_reference():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "lencod")) and \
        os.path.exists(os.path.join(ROOT, "oracle", "_ref", "fixtures", "encoder_main.cfg"))


def cpu_baseline_sample():
    """bounded sample on one host core: reference encoder on I + 2 P pictures (1 + 2 = 3 (frame, ref) units, ~20 s),
    normalised to the 5-reference workload (ME cost is linear in the number of references)."""
    if have_reference():
        d = prepare_reference_dir(3)
        me_s, units = run_lencod(d, 3)
        return {"value": units / N_REFS / me_s, "unit": UNIT, "cores": 1, "kind": "reference",
                "sample": "oracle/_ref/lencod (unmodified JM 12.4, gcc -O3), 1920x1088 synthetic IPP, 3 pictures = %d (MB,ref) searches, "
                          "ME time %.2f s from the encoder's own 'Total ME time' counter; MB/s = (MB,ref)/s / 5 refs" % (units, me_s)}
    # the oracle port: time the C restatement on a sample of jobs
    from oracle import binding as ob
    import h264_b200_pkg  # noqa: F401
    from h264_b200 import synth
    frames = synth.luma_frames(W, H, 2)
    jobs = synth.synthetic_jobs(W, H, 1, LAMBDA, seed=21)[:1500]
    cfg = ob.make_cfg()
    ref = ob.Ref(frames[0])
    t0 = time.perf_counter()
    ob.mb_search(cfg, np.ascontiguousarray(frames[1]), ref, jobs)
    dt = time.perf_counter() - t0
    return {"value": len(jobs) / N_REFS / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "oracle/me_oracle.c ora_mb_search on %d (MB,ref) jobs of the 1080p workload, %.2f s" % (len(jobs), dt)}


def reference_arm(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    if not have_reference():
        # the oracle port, one process per core
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/lencod not built on this box"}))
        return 0
    # Bounded sample: per step every core encodes I + P of a full-width strip of the 1080p clip.  A calibration step on an
    # 8-row strip measures this box; the strip then gets as many macroblock rows (8 .. 68 = the whole picture) as K + W steps
    # can afford in about 150 s.  The per-macroblock work does not depend on the picture height.
    import shutil
    exe = os.path.join(ROOT, "oracle", "_ref", "lencod")

    def make_workdirs(hs):
        d = prepare_reference_dir(2, h=hs)
        wds = []
        for c in range(cores):
            wd = os.path.join(d, "p%d" % c)
            os.makedirs(wd)
            for f in os.listdir(d):
                if f.endswith(".cfg"):
                    shutil.copy(os.path.join(d, f), wd)
            os.symlink(os.path.join(d, "clip.yuv"), os.path.join(wd, "clip.yuv"))
            wds.append(wd)
        return wds

    def run_step(wds, hs):
        t0 = time.perf_counter()
        procs = [subprocess.Popen([exe] + reference_cfg_args("clip.yuv", 2, h=hs), cwd=wd, stdout=subprocess.PIPE,
                                  stderr=subprocess.DEVNULL, text=True) for wd in wds]
        me = []
        for p in procs:
            out = p.communicate()[0]
            for ln in out.splitlines():
                if "Total ME time for sequence" in ln:
                    me.append(float(ln.split(":")[1].split("sec")[0]))
        return time.perf_counter() - t0, me

    # Beside the stepped sample: ONE encoder process on the stated configuration itself -- the whole 1920x1088 picture, 8 pictures
    # IPPP, so that pictures 5..7 search all five references -- started now on one core and collected at the
    # end; the stepped strips use the remaining cores.  It pins the per-core figure the strips extrapolate.
    full = None
    strip_cores = cores
    if cores >= 4 and os.environ.get("ME_BENCH_REF_FULL", "1") != "0":
        full_dir = prepare_reference_dir(8)
        full = (subprocess.Popen([exe] + reference_cfg_args("clip.yuv", 8), cwd=full_dir, stdout=subprocess.PIPE,
                                 stderr=subprocess.DEVNULL, text=True), time.perf_counter())
        strip_cores = cores - 1
    cores_all, cores = cores, strip_cores
    cal_s, _ = run_step(make_workdirs(16 * 8), 16 * 8)
    rows = int(8 * 150.0 / (max(1, args.steps + args.warmup) * cal_s))
    rows = max(8, min(H // 16, rows))
    hs = 16 * rows
    mbs = (W // 16) * rows
    workdirs = make_workdirs(hs)

    def one_step():
        return run_step(workdirs, hs)

    for _ in range(args.warmup):
        one_step()
    walls, mes = [], []
    for _ in range(args.steps):
        wall, me = one_step()
        walls.append(wall); mes.append(float(np.mean(me)))
    # throughput of the hot path: every process searched MBS (MB,ref) units in its ME seconds, all processes concurrently
    me_mean = float(np.mean(mes))
    per_core = mbs / N_REFS / me_mean
    full_note = None
    if full:
        out = full[0].communicate()[0]
        me_full = None
        for ln in out.splitlines():
            if "Total ME time for sequence" in ln:
                me_full = float(ln.split(":")[1].split("sec")[0])
        units_full = MBS * sum(min(k, N_REFS) for k in range(1, 8))           # 25 (picture, ref) searches of 8160 macroblocks
        if full[0].returncode == 0 and me_full:
            full_note = {"what": "one lencod process, 1920x1088, 8 pictures IPPP, SR 32, 5 references: pictures 5..7 search five references",
                         "me_seconds": me_full, "mb_ref_searches": units_full, "mb_per_s_one_core": units_full / N_REFS / me_full,
                         "strip_sample_mb_per_s_one_core": per_core, "strip_over_full": per_core / (units_full / N_REFS / me_full),
                         "wall_s": time.perf_counter() - full[1], "ran": "concurrently with the stepped strips, on its own core"}
            # the stated configuration, measured directly, is the per-core figure; the strips confirm it step by step
            per_core = units_full / N_REFS / me_full
    cores = cores_all
    value = cores * per_core
    line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(walls)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": bench_config(args.gpus),
            "reference_sample": {"stepped": "per step %d host cores each run the unmodified lencod on I+P (1 reference) of a %dx%d strip of the clip "
                                            "(%d macroblocks); per-core MB/s = %d (MB,ref) / 5 refs / mean ME seconds (encoder's own ME timer) = %.1f"
                                            % (strip_cores, W, hs, mbs, mbs, mbs / N_REFS / me_mean),
                                 "full_config": full_note,
                                 "value": "host cores (%d) x per-core MB/s of the %s" % (cores, "full-configuration run" if full_note else "strips")},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                             "sample": "%d concurrent lencod processes per step on I+P strips of %dx%d%s" %
                                       (strip_cores, W, hs, "; one process on the full 1920x1088 x 8-picture, 5-reference configuration" if full_note else "")},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    import h264_b200_pkg  # noqa: F401
    from h264_b200 import MeCuda, synth, JOB_DTYPE, RESULT_DTYPE

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the motion-estimation path has no CPU fallback")
    os.environ["ME_CUDA_ZEROCOPY"] = host_transfer_mode(world)      # read by the library at its first batch
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL writes its version banner to STDOUT when the communicator comes up; stdout carries the one JSON line only, so
        # file descriptor 1 points at stderr while the group is created and used for the first time
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    dev = torch.device("cuda", local)

    # ---- workload: this rank's GOP (its own clip)
    n_frames = N_REFS + 2
    frames = synth.luma_frames(W, H, n_frames, seed=1234 + 17 * rank)
    jobs_np = synth.synthetic_jobs(W, H, N_REFS, LAMBDA, seed=21 + rank)
    n_jobs = len(jobs_np)
    # ME_BENCH_T8=1: side measurement with Transform8x8Mode (High profile: 8x8 Hadamard for block types 1..4)
    # ME_BENCH_SUBPEL_METRIC=0|1: side measurement with SAD / SSE as the sub-pel metric (MEDistortionHPel / QPel)
    sub_metric = env_int("ME_BENCH_SUBPEL_METRIC", 2)
    me = MeCuda(W, H, search_range=SEARCH_RANGE, max_num_references=N_REFS, device=local,
                me_metric=(0, sub_metric, sub_metric),
                transform8x8=1 if os.environ.get("ME_BENCH_T8", "") == "1" else 0)
    stream = torch.cuda.ExternalStream(me.stream(), device=dev)

    # device-resident inputs
    d_frames = [torch.from_numpy(f).to(dev) for f in frames]
    d_jobs = torch.from_numpy(jobs_np.view(np.uint8).reshape(n_jobs, JOB_DTYPE.itemsize)).to(dev)
    d_out = torch.empty((n_jobs, RESULT_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    d_out_int = torch.empty((n_jobs, RESULT_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    # pinned host buffers for the end-to-end path
    h_frames = [torch.from_numpy(f).pin_memory() for f in frames]
    h_jobs = torch.from_numpy(jobs_np.view(np.uint8).reshape(n_jobs, JOB_DTYPE.itemsize)).pin_memory()
    h_out = torch.empty((n_jobs, RESULT_DTYPE.itemsize), dtype=torch.uint8).pin_memory()

    handles = [me.ref_create(frames[N_REFS - 1 - r]) for r in range(N_REFS)]   # ref r = r+1 pictures back
    me.set_ref_list(0, handles)
    me.sync()

    def rotate(step):
        # the oldest frame store receives the newest reconstructed picture and becomes reference 0
        h_old = handles.pop()
        handles.insert(0, h_old)
        return h_old, (N_REFS + step) % n_frames

    def device_step(step):
        h_new, fi = rotate(step)
        me.ref_update_device(h_new, d_frames[fi].data_ptr(), W)
        me.set_ref_list(0, handles)
        me.set_current_device(d_frames[(fi + 1) % n_frames].data_ptr(), W)
        me.frame_search_device(d_jobs.data_ptr(), n_jobs, d_out.data_ptr(), d_out_int.data_ptr())

    def host_step(step):
        h_new, fi = rotate(step)
        me.ref_update(h_new, (h_frames[fi].data_ptr(), W))
        me.set_ref_list(0, handles)
        cur = h_frames[(fi + 1) % n_frames]
        me._check(me.L.me_cuda_set_current_u8(cur.data_ptr(), W, W, H))
        me._check(me.L.me_cuda_frame_search(h_jobs.data_ptr(), n_jobs, h_out.data_ptr(), None))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing.  The clock sampler (nvidia-smi every 100 ms) runs from the first warm-up step to the end of the
    # timed region: same load throughout, and a timed region of a few hundred ms alone would yield too few samples.
    sampler = ClockSampler(local)
    sampler.start()
    t_warm = time.perf_counter()
    s = 0
    while s < args.warmup or time.perf_counter() - t_warm < 2.5:     # at least W steps and 2.5 s of load (>= 20 clock samples)
        device_step(s)
        s += 1
        if s % 8 == 0:
            me.sync()
    n_warm = s
    me.sync()
    barrier()
    launches0 = me.launch_count()
    me.set_timing(True)
    step_ms = []
    t_wall0 = time.perf_counter()
    for s in range(args.steps):
        with torch.cuda.stream(stream):
            flush.fill_(s & 255)                                      # evict L2 between timed iterations (not timed)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            device_step(n_warm + s)
            e1.record(stream)
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
    barrier()
    wall_s = time.perf_counter() - t_wall0
    launches = me.launch_count() - launches0
    ffs_ms, ffs_n = me.kernel_time_ms(1)
    sub_ms, sub_n = me.kernel_time_ms(2)
    pl_ms, pl_n = me.kernel_time_ms(0)
    me.set_timing(False)
    clocks = sampler.stop((t_wall0, t_wall0 + wall_s))
    # share of the partitions the sub-pel kernel reads from its TMA-staged region (the rest: clamped reads from global memory); one
    # more step outside the timed region with the counters on
    me.subpel_region_stats(True)
    device_step(n_warm + args.steps)
    me.sync()
    region_served, region_total = me.subpel_region_stats(False)
    total_ms = float(sum(step_ms))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * MBS / (ms_per_step * 1e-3)

    # ---- end-to-end timing (host buffers in, host buffers out)
    for s in range(3):
        host_step(s)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        host_step(3 + s)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * MBS * args.steps / e2e_s
    h2d = 2 * W * H + n_jobs * JOB_DTYPE.itemsize
    d2h = n_jobs * RESULT_DTYPE.itemsize

    # sanity: the end-to-end results equal the device-resident ones for the same inputs (same code path)
    checksum = int(torch.from_numpy(h_out.numpy().view(np.int32)).sum().item())

    # ---- roofline of the dominant kernel (integer-pel search).  The bound is integer issue, not HBM and not tensor:
    # denominator = the two 16-lane integer pipes of every SM sub-partition, each measured alone on this GPU
    # (VABSDIFF4 on the ALU pipe + IMAD on the FMA pipe; MEASURED_PEAKS.json carries no integer figure).
    peaks = {}
    if rank == 0:
        names = {0: "vabsdiff4", 2: "imad", 3: "viaddmnmx", 5: "shf", 9: "vabsdiff4+imad", 10: "vabsdiff4+viaddmnmx",
                 15: "vabsdiff4+lds", 18: "sweep_mix_32vabsdiff4_21addmin_33imad"}
        for k, nm in names.items():
            peaks[nm] = me.measure_int_peak(k) / 1e9
    ffs_avg_ms = ffs_ms / max(1, ffs_n)
    achieved = LANE_OPS_PER_JOB * n_jobs / (ffs_avg_ms * 1e-3) / 1e9 if ffs_n else None
    peak = (peaks["vabsdiff4"] + peaks["imad"]) if peaks else None
    mp = {}
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    alg_bytes = n_jobs * ((2 * SEARCH_RANGE + 16) ** 2 + 256 + JOB_DTYPE.itemsize + RESULT_DTYPE.itemsize)
    traffic = None          # DRAM bytes per launch of the search kernel, from the committed ncu capture (not measurable in-process)
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["me_ffs_tuned_kernel"]
        traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:
        pass
    roofline = {"bound": "int_alu", "kernel": "integer-pel fast full search (me_ffs_*)", "achieved": achieved, "peak": peak,
                "unit": "Glaneop/s", "frac": (achieved / peak) if (achieved and peak) else None,
                "peak_source": "measured on this GPU by me_cuda_measure_int_peak: VABSDIFF4 alone (ALU pipe) + IMAD alone (FMA pipe), "
                               "the two integer pipes; MEASURED_PEAKS.json has no integer figure",
                "algorithmic_laneops_per_launch": LANE_OPS_PER_JOB * n_jobs,
                "algorithmic_laneops_per_job": "212 * (2R+1)^2 = 64 SAD + 25 larger-block adds + 3*41 cost ops per position (SURVEY.md 8d)",
                "kernel_ms": ffs_avg_ms,
                "frac_of_alu_pipe_alone": (achieved / peaks["vabsdiff4"]) if (achieved and peaks) else None,
                # what a dependency-free stream of the sweep's own arithmetic (32 VABSDIFF4 : 21 add+min : 33 IMAD per window row) issues on
                # this GPU: the two pipes do not overlap fully (register-file reads, profiles/r02_ffs_ceiling.md) -- the practical ceiling
                "sweep_mix_issue_ceiling_glaneops": peaks.get("sweep_mix_32vabsdiff4_21addmin_33imad") if peaks else None,
                "frac_of_sweep_mix_ceiling": (achieved / peaks["sweep_mix_32vabsdiff4_21addmin_33imad"]) if (achieved and peaks) else None,
                "sad_only_frac_of_alu_pipe": (SAD_OPS_PER_JOB * n_jobs / (ffs_avg_ms * 1e-3) / 1e9 / peaks["vabsdiff4"]) if (ffs_n and peaks) else None,
                "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (ffs_avg_ms * 1e-3) / 1e9 if ffs_n else None,
                        "peak_gbs": mp.get("hbm_gbs", 6650.0), "peak_source": "MEASURED_PEAKS.json" if mp else "fallback"},
                "traffic": traffic, "traffic_source": "profiles/ncu_traffic.json (ncu --set full, same workload); windows are re-read from L2, "
                                                      "DRAM sees each reference plane once",
                "int_peaks_glaneops": peaks,
                "share_of_step": {"planes_ms": pl_ms / max(1, args.steps), "ffs_ms": ffs_ms / max(1, args.steps),
                                  "subpel_ms": sub_ms / max(1, args.steps)},
                "subpel_region_share": (region_served / region_total) if region_total else None}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline_sample()
        except Exception as e:  # the baseline is informative; never lose the bench line over it
            cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": "failed: %s" % str(e)[:200]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic",
                "config": bench_config(world, sub_metric),
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": 1e3 * e2e_s / args.steps, "result_checksum": checksum},
                "roofline": roofline, "wall_s_timed_region": wall_s}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    me.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
