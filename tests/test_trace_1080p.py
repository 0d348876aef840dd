"""GPU: the headline picture size on the reference's OWN predictors, every job checked.

The unmodified reference encoder (oracle/_ref/lencod_trace) codes four 1920x1088 pictures I P P P of the synthetic clip (SR 32,
three references in use for the last one) and traces every FastFullPelBlockMotionSearch / SubPelBlockMotionSearch call of the
last picture: 8160 macroblocks x 3 references x 41 partitions.  The traced calls of one (macroblock, list, ref) become one
me_mb_job (tests/trace_jobs.py) and go through me_cuda_frame_search -- the two throughput kernels; every traced integer-pel
vector + cost and final vector + cost must come back, picture-border macroblocks included.  The run also records what the
kernels cost on these predictors (not the synthetic ones of bench.py) and which share of the partitions the sub-pel kernel
served from its staged region: gpurun_out/trace_1080p.json when that directory exists."""
import json
import os
import subprocess
import time

import numpy as np
import pytest

from conftest import REF_DIR, ROOT, has_ref_binaries

pytestmark = pytest.mark.gpu
W, H, PICS = 1920, 1088, 4


@pytest.mark.skipif(not has_ref_binaries(), reason="oracle/_ref not built (needs the reference tree at build time)")
def test_1080p_reference_trace_through_the_frame_batch(tmp_path, me_lib_path):
    import h264_b200_pkg  # noqa: F401
    from h264_b200 import MeCuda, synth
    from h264_b200.me_cuda import JOB_DTYPE
    from trace_jobs import compare_fast, jobs_of_picture_fast
    from trace_util import Trace
    d = str(tmp_path)
    for f in os.listdir(os.path.join(REF_DIR, "fixtures")):
        if f.endswith(".cfg"):
            open(os.path.join(d, f), "wb").write(open(os.path.join(REF_DIR, "fixtures", f), "rb").read())
    synth.write_yuv420(os.path.join(d, "syn.yuv"), W, H, PICS)
    args = ["-d", "encoder_main.cfg", "-p", "InputFile=syn.yuv", "-p", "SourceWidth=%d" % W, "-p", "SourceHeight=%d" % H,
            "-p", "FramesToBeEncoded=%d" % PICS, "-p", "FrameSkip=0", "-p", "NumberBFrames=0", "-p", "SearchRange=32",
            "-p", "NumberReferenceFrames=5", "-p", "LevelIDC=51"]
    t0 = time.perf_counter()
    r = subprocess.run([os.path.join(REF_DIR, "lencod_trace")] + args, cwd=d, capture_output=True, text=True, timeout=1500,
                       env=dict(os.environ, ME_TRACE_DIR=d, ME_TRACE_PLANES="0", ME_TRACE_FROM_PIC=str(PICS - 1)))
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    cpu_s = time.perf_counter() - t0
    tr = Trace(d)
    pic = PICS - 1
    jobs, exp = jobs_of_picture_fast(tr.recs, pic, JOB_DTYPE)
    n_refs = 1 + max(r for (l, r) in exp["ref_ids"])
    assert len(jobs) == (W // 16) * (H // 16) * n_refs and n_refs == 3 and exp["mask"].all()
    with MeCuda(W, H, search_range=32, max_num_references=5, rdopt=1, full_search=2, me_metric=(0, 2, 2), transform8x8=0,
                mv_limits=(-511, 511)) as me:
        handles = [me.ref_create(tr.ref(exp["ref_ids"][(0, i)])) for i in range(n_refs)]
        me.set_ref_list(0, handles)
        me.set_current(tr.cur(pic))
        out, out_int = me.frame_search(jobs, want_int=True)           # warm-up and the checked run
        n, bad = compare_fast(exp, out, out_int)
        assert n == len(jobs) * 41
        assert not len(bad), (len(bad), [(int(j), int(p), out_int["mv"][j][p].tolist(), int(out_int["cost"][j][p]), exp["int_mv"][j][p].tolist(),
                                          int(exp["int_cost"][j][p]), out["mv"][j][p].tolist(), int(out["cost"][j][p]), exp["mv"][j][p].tolist(),
                                          int(exp["cost"][j][p])) for j, p in bad[:4]])
        # what the kernels cost on the reference's predictors
        me.subpel_region_stats(True)
        me.set_timing(True)
        for _ in range(5):
            me.frame_search(jobs)
        ffs_ms, ffs_n = me.kernel_time_ms(1)
        sub_ms, sub_n = me.kernel_time_ms(2)
        me.set_timing(False)
        served, total = me.subpel_region_stats(False)
    assert total == 5 * len(jobs) * 41 and 0 < served <= total
    # distance of the traced partition predictors from the traced 16x16 predictor: what the synthetic jobs of bench.py imitate
    dpm = np.abs(jobs["pmv"].astype(np.int64) - jobs["pmv16"].astype(np.int64)[:, None, :]).max(axis=2)
    line = {"what": "1080p, SR 32, picture %d of I P P P coded by the unmodified reference: %d (macroblock, ref) jobs x 41 partitions, every traced "
                    "result compared" % (pic, len(jobs)), "mismatches": int(len(bad)), "refs": n_refs,
            "ffs_ms_per_batch": ffs_ms / 5.0, "subpel_ms_per_batch": sub_ms / 5.0, "launches_per_batch": (ffs_n + sub_n) / 5.0,
            "jobs_per_batch": int(len(jobs)), "subpel_region_share": served / total,
            "pmv_minus_pmv16_qpel": {"median": float(np.median(dpm)), "p99": float(np.percentile(dpm, 99)), "max": int(dpm.max())},
            "reference_encoder_wall_s": cpu_s}
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "trace_1080p.json"), "w") as fh:
            fh.write(json.dumps(line) + "\n")
