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
