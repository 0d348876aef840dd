"""Frame-batch jobs (me_mb_job) built from the reference's own call traces.

A trace of the unmodified reference encoder (oracle/trace_wrap.c) holds, per (picture, macroblock, list, ref), the calls
PartitionMotionSearch made (mv-search.c:1378-1454): one FastFullPelBlockMotionSearch record and one SubPelBlockMotionSearch
record per searched block, each with the predictor, the lambda factors and the result.  The frame batch of include/me_cuda.h
takes exactly these inputs per partition (pmv16 = the predictor SetupFastFullPelSearch used for the search centre,
me_fullfast.c:550; 41 partition predictors; lambda), so the traced calls of one (macroblock, list, ref) become one job, and
the traced results are what the batch must return for the traced partitions.

A partition can be searched more than once per (macroblock, list, ref) -- with Transform8x8Mode the 8x8 blocks are searched
in both sub-macroblock loops (md_high.c:195-253), with different predictors when the neighbours' decisions changed in
between.  The n-th search of a partition goes to the n-th job ("layer") of that (macroblock, list, ref).
"""
import numpy as np

from trace_util import REC_FFS, REC_SUBPEL

PART_TYPE = [1, 2, 2, 3, 3, 4, 4, 4, 4] + [5] * 8 + [6] * 8 + [7] * 16
PART_BLK = [0, 0, 8, 0, 2, 0, 2, 8, 10, 0, 2, 4, 6, 8, 10, 12, 14, 0, 1, 2, 3, 8, 9, 10, 11] + list(range(16))
PART_OF = {(t, b): p for p, (t, b) in enumerate(zip(PART_TYPE, PART_BLK))}


def jobs_of_picture(trace, pic, job_dtype):
    """-> (jobs, expect): jobs of `job_dtype` (h264_b200.me_cuda.JOB_DTYPE) for coded picture `pic`, and per job a dict
    partition -> (int mv x, y, int cost, final mv x, y, final cost) from the trace.  Also returns the per-(list, ref) weights."""
    recs = trace.recs[(trace.recs[:, 1] == pic) & ((trace.recs[:, 0] == REC_FFS) | (trace.recs[:, 0] == REC_SUBPEL))]
    groups = {}            # (ox, oy, list, ref) -> list of layers; layer = dict(pmv16, lam, parts {p: [ffs rec, sub rec]})
    order = []
    weights = {}
    pending = {}           # (key, p) -> (layer index) of the FFS record still waiting for its sub-pel record
    for r in recs:
        key = (int(r[3]), int(r[4]), int(r[5]), int(r[6]))
        blk = (int(r[9]) - int(r[4])) + ((int(r[8]) - int(r[3])) >> 2)
        p = PART_OF[(int(r[7]), blk)]
        if key not in groups:
            groups[key] = []
            order.append(key)
        layers = groups[key]
        weights[(key[2], key[3])] = (int(r[28]), int(r[29]), int(r[30]), int(r[31]))
        if int(r[0]) == REC_FFS:
            # search centre: the predictor the set-up call saw, or (later calls) any vector that truncates to the same centre
            pmv16 = (int(r[23]), int(r[24])) if int(r[22]) else (4 * (int(r[15]) - int(r[3])), 4 * (int(r[16]) - int(r[4])))
            lam_f = int(r[13])
            li = 0
            while li < len(layers) and (p in layers[li]["parts"] or layers[li]["lam_f"] != lam_f):
                li += 1
            if li == len(layers):
                layers.append(dict(pmv16=pmv16, lam_f=lam_f, lam=None, parts={}, slice_type=int(r[2]) & 255))
            lay = layers[li]
            lay["parts"][p] = [r, None]
            pending[(key, p)] = li
        else:
            li = pending.pop((key, p), None)
            if li is None:
                continue                                   # a sub-pel call without its integer-pel call in the kept subset
            ffs = layers[li]["parts"][p][0]
            # the sub-pel call BlockMotionSearch issues right after (mv-search.c:770-826): starts at the integer result << 2
            if (int(r[12]), int(r[13])) != (4 * int(ffs[18]), 4 * int(ffs[19])):
                del layers[li]["parts"][p]
                continue
            lam3 = (int(r[15]), int(r[16]), int(r[17]))
            if layers[li]["lam"] is None:
                layers[li]["lam"] = lam3
            if layers[li]["lam"] != lam3 or lam3[0] != layers[li]["lam_f"]:
                del layers[li]["parts"][p]
                continue
            layers[li]["parts"][p][1] = r
    jobs, expect = [], []
    for key in order:
        # the centre of the whole group comes from its set-up record
        centre16 = None
        for lay in groups[key]:
            for p, (f, s) in lay["parts"].items():
                if int(f[22]):
                    centre16 = (int(f[23]), int(f[24]))
        for lay in groups[key]:
            parts = {p: fs for p, fs in lay["parts"].items() if fs[1] is not None}
            if not parts or lay["lam"] is None:
                continue
            j = np.zeros((), job_dtype)
            j["opix_x"], j["opix_y"], j["list"], j["ref"] = key
            j["slice_type"] = 1 if lay["slice_type"] == 1 else 0
            j["pmv16"] = centre16 if centre16 is not None else lay["pmv16"]
            j["lambda"] = lay["lam"]
            exp = {}
            for p, (f, s) in parts.items():
                j["pmv"][p] = (int(f[10]), int(f[11]))
                assert (int(s[10]), int(s[11])) == (int(f[10]), int(f[11]))
                exp[p] = (int(f[18]), int(f[19]), int(f[20]), int(s[18]), int(s[19]), int(s[20]))
            jobs.append(j)
            expect.append(exp)
    return (np.array(jobs, job_dtype) if jobs else np.zeros(0, job_dtype)), expect, weights


def compare(jobs, expect, out, out_int):
    """-> list of mismatches (job, partition, got, expected)"""
    bad = []
    n = 0
    for k, exp in enumerate(expect):
        for p, e in exp.items():
            got = (int(out_int["mv"][k][p][0]), int(out_int["mv"][k][p][1]), int(out_int["cost"][k][p]),
                   int(out["mv"][k][p][0]), int(out["mv"][k][p][1]), int(out["cost"][k][p]))
            n += 1
            if got != e:
                bad.append((k, p, got, e))
    return n, bad


def jobs_of_picture_fast(recs, pic, job_dtype):
    """Vectorised jobs_of_picture for a picture in which every partition of a (macroblock, list, ref) is searched exactly once, each
    FastFullPelBlockMotionSearch call directly followed by its SubPelBlockMotionSearch call (P pictures, Transform8x8Mode off,
    mv-search.c:745-826) -- full-size traces hold millions of records.  -> (jobs, exp) with exp = dict(mask [n, 41],
    int_mv [n, 41, 2], int_cost [n, 41], mv [n, 41, 2], cost [n, 41], ref_ids {(list, ref): picture id})."""
    f = recs[(recs[:, 0] == REC_FFS) & (recs[:, 1] == pic)]
    s = recs[(recs[:, 0] == REC_SUBPEL) & (recs[:, 1] == pic)]
    assert len(f) and len(f) == len(s), (len(f), len(s))
    for c in (3, 4, 5, 6, 7, 8, 9, 10, 11):                        # same block, same predictor, call by call
        assert np.array_equal(f[:, c], s[:, c]), c
    assert np.array_equal(4 * f[:, 18], s[:, 12]) and np.array_equal(4 * f[:, 19], s[:, 13])     # sub-pel starts at the integer result << 2
    assert np.array_equal(f[:, 13], s[:, 15])                                                    # lambda[F_PEL]
    key = f[:, 3].astype(np.int64) | (f[:, 4].astype(np.int64) << 16) | (f[:, 5].astype(np.int64) << 32) | (f[:, 6].astype(np.int64) << 40)
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first)                                      # jobs in the order the encoder reached them
    rank = np.empty(len(uniq), np.int64)
    rank[order] = np.arange(len(uniq))
    ji = rank[inv]
    lut = np.full(8 * 16, -1, np.int64)
    for p, (t, b) in enumerate(zip(PART_TYPE, PART_BLK)):
        lut[t * 16 + b] = p
    blk = (f[:, 9] - f[:, 4]) + ((f[:, 8] - f[:, 3]) >> 2)
    p = lut[f[:, 7] * 16 + blk]
    assert (p >= 0).all()
    n = len(uniq)
    mask = np.zeros((n, 41), bool)
    assert not mask[ji, p].any()
    mask[ji, p] = True
    assert len(np.unique(ji * 41 + p)) == len(f), "a partition searched twice: use jobs_of_picture"
    jobs = np.zeros(n, job_dtype)
    jobs["opix_x"][ji] = f[:, 3]; jobs["opix_y"][ji] = f[:, 4]; jobs["list"][ji] = f[:, 5]; jobs["ref"][ji] = f[:, 6]
    jobs["slice_type"][ji] = np.where((f[:, 2] & 255) == 1, 1, 0)
    jobs["pmv"][ji, p, 0] = f[:, 10]; jobs["pmv"][ji, p, 1] = f[:, 11]
    su = f[f[:, 22] != 0]                                           # the call that ran SetupFastFullPelSearch: its predictor is the centre's
    sji = rank[np.searchsorted(uniq, su[:, 3].astype(np.int64) | (su[:, 4].astype(np.int64) << 16) | (su[:, 5].astype(np.int64) << 32) | (su[:, 6].astype(np.int64) << 40))]
    got = np.zeros(n, bool); got[sji] = True
    assert got.all(), "a (macroblock, list, ref) without its set-up call"
    jobs["pmv16"][sji, 0] = su[:, 23]; jobs["pmv16"][sji, 1] = su[:, 24]
    jobs["lambda"][ji, 0] = s[:, 15]; jobs["lambda"][ji, 1] = s[:, 16]; jobs["lambda"][ji, 2] = s[:, 17]
    exp = dict(mask=mask, int_mv=np.zeros((n, 41, 2), np.int64), int_cost=np.zeros((n, 41), np.int64),
               mv=np.zeros((n, 41, 2), np.int64), cost=np.zeros((n, 41), np.int64), ref_ids={})
    exp["int_mv"][ji, p, 0] = f[:, 18]; exp["int_mv"][ji, p, 1] = f[:, 19]; exp["int_cost"][ji, p] = f[:, 20]
    exp["mv"][ji, p, 0] = s[:, 18]; exp["mv"][ji, p, 1] = s[:, 19]; exp["cost"][ji, p] = s[:, 20]
    for l, r, i in set(map(tuple, f[:, [5, 6, 21]].tolist())):
        exp["ref_ids"][(int(l), int(r))] = int(i)
    return jobs, exp


def compare_fast(exp, out, out_int):
    """-> (number of traced partition results, indices [job, partition] of the mismatches)"""
    m = exp["mask"]
    bad = m & ((out_int["mv"].astype(np.int64) != exp["int_mv"]).any(axis=2) | (out_int["cost"].astype(np.int64) != exp["int_cost"]) |
               (out["mv"].astype(np.int64) != exp["mv"]).any(axis=2) | (out["cost"].astype(np.int64) != exp["cost"]))
    return int(m.sum()), np.argwhere(bad)
