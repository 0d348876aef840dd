"""CPU: the job builder of tests/trace_jobs.py checked on the oracle -- frame-batch jobs made from the reference's own call
traces must give the traced integer-pel and final results when they go through the oracle's batch form (ora_mb_search).
The GPU twin (tests/test_gpu_parity.py::test_golden_traces_through_the_frame_batch) sends the same jobs through
me_cuda_frame_search, i.e. through the two throughput kernels."""
import numpy as np
import pytest

from conftest import golden_trace
from replay import picture_lists
from trace_jobs import compare, jobs_of_picture
from trace_util import REC_CURPIC


@pytest.mark.parametrize("case", ["kat", "baseline16", "wp"])
def test_oracle_batch_reproduces_the_traced_results(oracle_lib, case):
    ob = oracle_lib
    import h264_b200_pkg  # noqa: F401
    from h264_b200.me_cuda import JOB_DTYPE
    tr, cfgkw = golden_trace(case)
    cfg = ob.make_cfg(**{k: v for k, v in cfgkw.items() if k != "chroma_me"})
    total = 0
    for pic in tr.cur_ids():
        jobs, expect, weights = jobs_of_picture(tr, pic, JOB_DTYPE)
        if not len(jobs):
            continue
        lists = picture_lists(tr, pic)
        cur = np.ascontiguousarray(tr.cur(pic))
        res, res_int = None, None
        for (lst, ref) in sorted(set(zip(jobs["list"].tolist(), jobs["ref"].tolist()))):
            sel = np.nonzero((jobs["list"] == lst) & (jobs["ref"] == ref))[0]
            r = ob.Ref(tr.ref(lists[lst][ref]))
            r.set_weights(*weights.get((lst, ref), (0, 32, 0, 5)))
            o, oi = ob.mb_search(cfg, cur, r, jobs[sel])
            r.close()
            if res is None:
                res, res_int = np.zeros(len(jobs), o.dtype), np.zeros(len(jobs), oi.dtype)
            res[sel], res_int[sel] = o, oi
        n, bad = compare(jobs, expect, res, res_int)
        assert not bad, (case, pic, n, len(bad), bad[:4])
        total += n
    assert total > 1000, total


@pytest.mark.skipif(not __import__("conftest").has_ref_binaries(), reason="oracle/_ref not built (needs the reference tree at build time)")
def test_fast_builder_equals_the_general_one_on_a_live_trace(tmp_path, oracle_lib):
    """jobs_of_picture_fast (vectorised, for full-size traces; tests/test_trace_1080p.py) against jobs_of_picture on a live QCIF
    trace of the unmodified reference encoder written with ME_TRACE_FROM_PIC, and both against the oracle's batch form"""
    import os
    import subprocess
    import h264_b200_pkg  # noqa: F401
    from h264_b200 import synth
    from h264_b200.me_cuda import JOB_DTYPE
    from conftest import REF_DIR
    from trace_jobs import compare_fast, jobs_of_picture_fast
    from trace_util import Trace, REC_FFS
    ob = oracle_lib
    d = str(tmp_path)
    for f in os.listdir(os.path.join(REF_DIR, "fixtures")):
        if f.endswith(".cfg"):
            open(os.path.join(d, f), "wb").write(open(os.path.join(REF_DIR, "fixtures", f), "rb").read())
    synth.write_yuv420(os.path.join(d, "syn.yuv"), 176, 144, 3)
    args = ["-d", "encoder_main.cfg", "-p", "InputFile=syn.yuv", "-p", "SourceWidth=176", "-p", "SourceHeight=144", "-p", "FramesToBeEncoded=3",
            "-p", "FrameSkip=0", "-p", "NumberBFrames=0", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2"]
    r = subprocess.run([os.path.join(REF_DIR, "lencod_trace")] + args, cwd=d, capture_output=True, text=True,
                       env=dict(os.environ, ME_TRACE_DIR=d, ME_TRACE_PLANES="0", ME_TRACE_FROM_PIC="2"))
    assert r.returncode == 0, r.stdout[-1000:]
    tr = Trace(d)
    assert set(tr.of_type(REC_FFS)[:, 1].tolist()) == {2}                    # only the last picture's searches were kept
    jobs, exp = jobs_of_picture_fast(tr.recs, 2, JOB_DTYPE)
    slow, slow_exp, _ = jobs_of_picture(tr, 2, JOB_DTYPE)
    assert len(jobs) == len(slow) == 99 * 2
    assert jobs.tobytes() == slow.tobytes()
    for k in (0, 57, 197):
        for p, e in slow_exp[k].items():
            assert e == (exp["int_mv"][k][p][0], exp["int_mv"][k][p][1], exp["int_cost"][k][p], exp["mv"][k][p][0], exp["mv"][k][p][1], exp["cost"][k][p])
    cfg = ob.make_cfg(search_range=16, rdopt=1, full_search=2, metric=(0, 2, 2), transform8x8=0, mv_limit_min_y=-511, mv_limit_max_y=511)
    cur = np.ascontiguousarray(tr.cur(2))
    res, res_int = None, None
    for ref in (0, 1):
        sel = np.nonzero(jobs["ref"] == ref)[0]
        rr = ob.Ref(tr.ref(exp["ref_ids"][(0, ref)]))
        o, oi = ob.mb_search(cfg, cur, rr, jobs[sel])
        rr.close()
        if res is None:
            res, res_int = np.zeros(len(jobs), o.dtype), np.zeros(len(jobs), oi.dtype)
        res[sel], res_int[sel] = o, oi
    n, bad = compare_fast(exp, res, res_int)
    assert n == len(jobs) * 41 and not len(bad), (n, bad[:4])
