"""GPU parity tests: libme_cuda.so through its C ABI against the oracle and the reference-derived golden vectors.
Bit-exact everywhere (integer / byte work): identical planes, motion vectors and costs."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_CASES, HBD_CASES, PATTERN_CASES, golden_trace
from replay import CudaAdapter, replay, pattern_counts

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg(me_lib_path):
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import h264_b200_pkg  # noqa: F401
    import h264_b200
    return h264_b200


def _cfg(ob, **kw):
    d = dict(search_range=32, rdopt=1, full_search=2, metric=(0, 2, 2), transform8x8=0, mv_limit_min_y=-511, mv_limit_max_y=511)
    d.update(kw)
    return d, ob.make_cfg(**d)


def _open(pkg, w, h, d, max_refs=5, **extra):
    return pkg.MeCuda(w, h, search_range=d["search_range"], max_num_references=max_refs, rdopt=d["rdopt"],
                      full_search=d["full_search"], me_metric=d["metric"], transform8x8=d["transform8x8"],
                      mv_limits=(d["mv_limit_min_y"], d["mv_limit_max_y"]), chroma_me=d.get("chroma_me", 0), **extra)


# ------------------------------------------------------------------------------------------------ planes
@pytest.mark.parametrize("kind", ["pan", "noise", "flat", "extremes"])
def test_planes_bit_exact(pkg, oracle_lib, kind):
    from h264_b200 import synth
    w, h = 352, 288
    rng = np.random.default_rng(5)
    if kind == "pan":
        luma = synth.luma_frames(w, h, 2)[1]
    elif kind == "noise":
        luma = rng.integers(0, 256, (h, w)).astype(np.uint8)
    elif kind == "flat":
        luma = np.full((h, w), 200, np.uint8)
    else:   # 0/255 checkerboards drive the six-tap filter into both clip limits
        luma = (((np.arange(h)[:, None] // 3 + np.arange(w)[None, :] // 2) % 2) * 255).astype(np.uint8)
    d, _ = _cfg(oracle_lib)
    with _open(pkg, w, h, d) as me:
        hd = me.ref_create(luma)
        got = me.ref_planes(hd)
        ref = oracle_lib.Ref(luma)
        exp = ref.planes()
        assert got.shape == exp.shape
        bad = np.argwhere(got != exp)
        assert len(bad) == 0, (kind, bad[:5])
        # JM-style entry: 16-bit samples in, imgY_sub[4][4][h+40][w+40] filled on the host
        hd2, sub = me.ref_create_imgpel(luma.astype(np.uint16), want_planes=True)
        assert np.array_equal(sub, exp.astype(np.uint16))
        me.ref_destroy(hd); me.ref_destroy(hd2)
        ref.close()


def test_planes_golden_reference_pictures(pkg, oracle_lib):
    """the reference's own reconstructed pictures from the known-answer run"""
    import json, zlib
    from conftest import GOLDEN
    crc = json.load(open(os.path.join(GOLDEN, "planes_crc.json")))
    for case in sorted(GOLDEN_CASES):
        tr, cfgkw = golden_trace(case)
        h, w = tr.size[("cur", tr.cur_ids()[0])]
        with _open(pkg, w, h, cfgkw) as me:
            for i in tr.ref_ids():
                hd = me.ref_create(tr.ref(i))
                for fy in range(4):
                    for fx in range(4):
                        assert zlib.crc32(me.ref_plane(hd, fy, fx).tobytes()) == crc[case][str(i)][fy][fx], (case, i, fy, fx)


# ------------------------------------------------------------------------------------------------ drop-in calls
@pytest.mark.parametrize("case", sorted(GOLDEN_CASES))
def test_dropin_calls_match_reference_trace(pkg, case):
    """every traced call of the unmodified reference encoder, replayed through the drop-in C ABI"""
    tr, cfgkw = golden_trace(case)
    impl = CudaAdapter(tr, cfgkw)
    try:
        res = replay(tr, impl)
    finally:
        impl.close()
    for name, (n, bad) in res.items():
        assert not bad, (case, name, n, len(bad), bad[:3])
    # the pattern-search trace (epzs) holds no integer-pel calls of the path: that stage stays the reference's own
    assert (res["ffs"][0] > 1000 or res["fullpel"][0] > 1000 or case == "epzs") and res["subpel"][0] > 1000


@pytest.mark.parametrize("case", sorted(PATTERN_CASES))
def test_dropin_pattern_subpel_calls_match_reference_trace(pkg, case):
    """sub-pel stages of the pattern searches (SearchMode 1 / 2 / 3) through me_cuda_EPZSSubPelBlockMotionSearch,
    me_cuda_EPZSSubPelBlockSearchBiPred, me_cuda_UMHEXSubPelBlockMotionSearch, me_cuda_smpUMHEX[Full]SubPelBlockMotionSearch: every
    kept call of the unmodified reference encoder, vectors and costs identical"""
    tr, cfgkw = golden_trace(case)
    impl = CudaAdapter(tr, cfgkw)
    try:
        res = replay(tr, impl)
    finally:
        impl.close()
    for name, (n, bad) in res.items():
        assert not bad, (case, name, n, len(bad), bad[:3])
    pattern_counts(case, res)


def test_high_bit_depth_planes_and_dropin_calls(pkg, oracle_lib):
    """more than 8 bits per sample: 10-bit pictures as 16-bit samples through the row-pointer entry points --
    the 16 planes pixel-exact against the 16-bit build of the oracle and the reference's own CRCs, every kept call of the unmodified
    reference replayed through the drop-in ABI, and the 8-bit-only entry points refusing loudly"""
    import json, zlib
    from conftest import GOLDEN
    from h264_b200 import MeCudaError
    tr, cfgkw = golden_trace("hbd10")
    assert sorted(HBD_CASES) == ["hbd10"] and tr.bits == 10
    crc = json.load(open(os.path.join(GOLDEN, "planes_crc.json")))["hbd10"]
    h, w = tr.size[("cur", tr.cur_ids()[0])]
    oracle_lib.set_bit_depth(10)
    try:
        with pkg.MeCuda(w, h, search_range=16, max_num_references=2, transform8x8=cfgkw["transform8x8"], bit_depth=10) as me:
            hds = []
            for i in tr.ref_ids():
                hd, sub = me.ref_create_imgpel(tr.ref(i), want_planes=True)
                hds.append(hd)
                ref = oracle_lib.Ref(tr.ref(i))
                assert sub.dtype == np.uint16 and np.array_equal(sub, ref.planes()), i
                ref.close()
                for fy in range(4):
                    for fx in range(4):
                        assert zlib.crc32(np.ascontiguousarray(sub[fy, fx]).tobytes()) == crc[str(i)][fy][fx], (i, fy, fx)
            with pytest.raises(MeCudaError):
                me.ref_create(tr.ref(tr.ref_ids()[0]).astype(np.uint8))      # 8-bit entry point on a 10-bit library
            with pytest.raises(MeCudaError):
                me.frame_bipred(np.zeros(1, pkg.BIPRED_JOB_DTYPE), 1, 8, 2)  # the bi-predictive frame batch is 8-bit only
            # the frame batch on 16-bit samples (general kernels): all 41 partitions of every macroblock against the oracle
            from h264_b200 import synth
            ids = tr.ref_ids()
            cur = tr.cur(tr.cur_ids()[-1])
            me.set_ref_list(0, hds[:2])
            me.set_current_imgpel(cur)
            jobs = synth.synthetic_jobs(w, h, min(2, len(ids)), (334459, 334459, 334459), seed=9)
            out, out_int = me.frame_search(jobs, want_int=True)
            ocfg = oracle_lib.make_cfg(search_range=16, transform8x8=cfgkw["transform8x8"])
            for r in range(min(2, len(ids))):
                oref = oracle_lib.Ref(tr.ref(ids[r]))
                sel = np.nonzero(jobs["ref"] == r)[0]
                exp, exp_int = oracle_lib.mb_search(ocfg, np.ascontiguousarray(cur), oref, jobs[sel])
                assert np.array_equal(out_int[sel], exp_int) and np.array_equal(out[sel], exp), r
                oref.close()
    finally:
        oracle_lib.set_bit_depth(8)
    impl = CudaAdapter(tr, cfgkw, max_refs=2)
    try:
        res = replay(tr, impl)
    finally:
        impl.close()
    for name, (n, bad) in res.items():
        assert not bad, (name, n, len(bad), bad[:3])
    assert res["ffs"][0] > 3000 and res["subpel"][0] > 3000 and res["bi_full"][0] > 100 and res["bi_sub"][0] > 50


# ------------------------------------------------------------------------------------------------ one launch per PartitionMotionSearch
@pytest.mark.parametrize("slice_b,t8,mb", [(0, 0, (5, 4)), (0, 1, (0, 0)), (1, 0, (10, 8)), (0, 0, (21, 17))])
def test_partition_search_launch(pkg, slice_b, t8, mb):
    """me_cuda_pms_begin: the device runs PartitionMotionSearch calls (mv-search.c:1378-1454) -- predictor, block search, sub-pel
    refinement, write-back -- in one launch; the drop-in calls are then answered from it.  Checked against the same call sequence
    made block by block (no group launch), driven by the Python restatement of SetMotionVectorPredictor / PartitionMotionSearch
    (oracle/predictor.py), for random motion states incl. unavailable neighbours (picture corner / last column and row)."""
    import copy
    from h264_b200 import synth
    from h264_b200.me_cuda import MOTION_DTYPE, PMS_UNAVAILABLE
    from oracle import predictor as op
    w, h, n_refs = 352, 288, 3
    frames = synth.luma_frames(w, h, n_refs + 2)
    lam = (131072, 98304, 98304)
    rng = np.random.default_rng(100 * mb[0] + mb[1] + 7 * slice_b)
    mbx, mby = mb
    d = dict(search_range=16, rdopt=1, full_search=2, metric=(0, 2, 2), transform8x8=t8, mv_limit_min_y=-511, mv_limit_max_y=511)
    list_sizes = [n_refs, 2] if slice_b else [n_refs]
    # motion state: neighbours with random references / vectors; macroblocks outside the picture are unavailable
    st = np.zeros((), MOTION_DTYPE)
    st["ref_idx"] = rng.integers(-1, n_refs, size=(2, 5, 6))
    st["mv"] = rng.integers(-60, 61, size=(2, 5, 6, 2))
    st["mv"][st["ref_idx"] < 0] = 0
    for l in range(2):
        for r in range(5):
            for c in range(6):
                out = (c == 0 and mbx == 0) or (r == 0 and mby == 0) or (c == 5 and (r > 0 or mbx == w // 16 - 1 or mby == 0)) or l >= len(list_sizes)
                if out:
                    st["ref_idx"][l, r, c] = PMS_UNAVAILABLE
    groups = [[(1, 0), (2, 0), (3, 0)], [(2, 1)], [(3, 1)]] + [[(4, b), (5, b), (6, b), (7, b)] for b in range(4)]

    def run(use_group):
        res = {}
        with _open(pkg, w, h, d, max_refs=n_refs) as me:
            hd = [me.ref_create(frames[-2 - r]) for r in range(n_refs)]
            me.set_ref_list(0, hd)
            if slice_b:
                me.set_ref_list(1, hd[:2][::-1])
            me.set_slice_type(1 if slice_b else 0)
            me.set_current(frames[-1])
            me.mb_begin(16 * mbx, 16 * mby)
            ref_idx = st["ref_idx"].tolist(); mv = st["mv"].tolist()
            for l, n in enumerate(list_sizes):
                for r in range(n):
                    pmv16 = op.mv_predictor(ref_idx[l], mv[l], r, 0, 0, 16, 16)
                    me.ffs_setup(l, r, *pmv16)

            def block_search(l, r, hh, vv, bt, pmv):
                x, y = 16 * mbx + 4 * hh, 16 * mby + 4 * vv
                ix, iy, ic = me.FastFullPelBlockMotionSearch(r, l, x, y, bt, pmv[0], pmv[1], 16, 2147483647, lam[0])
                # start_me_refinement_hp = 0 (integer-pel metric SAD, half-pel metric Hadamard): BlockMotionSearch resets the cost (mv-search.c:785-788)
                fx, fy, fc = me.SubPelBlockMotionSearch(r, l, x, y, bt, pmv[0], pmv[1], 4 * ix, 4 * iy, 9, 9, 2147483647, lam)
                res[(bt, l, r, hh, vv)] = (pmv, ix, iy, ic, fx, fy, fc)
                return fx, fy, fc

            for grp in groups:
                # the state every call of the group starts from (the host would have taken mode decisions in between: emulated by
                # restoring the area inside the macroblock to the values the state was given)
                if use_group:
                    snap = np.zeros((), MOTION_DTYPE)
                    snap["ref_idx"] = ref_idx; snap["mv"] = mv
                    assert me.pms_begin(snap, grp, lam)
                for (bt, b8) in grp:
                    ri, mvs = copy.deepcopy(ref_idx), copy.deepcopy(mv)
                    assert me.pms_cached(bt, b8) == bool(use_group)
                    op.partition_motion_search(ri, mvs, bt, b8, list_sizes, block_search)
                    me.pms_end(bt, b8)
            stats = me.pms_stats()
        return res, stats

    plain, s0 = run(False)
    grouped, s1 = run(True)
    assert s0["launches"] == 0 and s0["served"] == 0
    n_searches = sum(len(op.partition_blocks(bt, b8)) for grp in groups for (bt, b8) in grp) * sum(list_sizes)
    assert len(plain) == len(grouped) and grouped == plain
    assert s1 == dict(launches=len(groups), served=2 * n_searches, missed=0), s1


# ------------------------------------------------------------------------------------------------ throughput kernels on reference traces
@pytest.mark.parametrize("case", ["kat", "baseline16", "wp"])
def test_golden_traces_through_the_frame_batch(pkg, case):
    """The two throughput kernels (me_ffs_tuned_kernel, me_subpel_tuned_kernel) on the reference's OWN predictors: every
    (macroblock, list, ref) of the golden traces becomes a me_mb_job (tests/trace_jobs.py: traced 16x16 predictor = search centre,
    traced partition predictors, traced lambda factors), goes through me_cuda_frame_search, and must return the traced
    integer-pel vector + cost and the traced final vector + cost of every traced call (mv-search.c:1378-1454)."""
    from replay import picture_lists
    from trace_jobs import compare, jobs_of_picture
    from h264_b200.me_cuda import JOB_DTYPE
    tr, cfgkw = golden_trace(case)
    h, w = tr.size[("cur", tr.cur_ids()[0])]
    total = 0
    with _open(pkg, w, h, cfgkw) as me:
        handles = {i: me.ref_create(tr.ref(i)) for i in tr.ref_ids()}
        for pic in tr.cur_ids():
            jobs, expect, weights = jobs_of_picture(tr, pic, JOB_DTYPE)
            if not len(jobs):
                continue
            lists = picture_lists(tr, pic)
            me.set_current(tr.cur(pic))
            for l in (0, 1):
                me.set_ref_list(l, [handles[i] for i in lists[l]])
                for ref in range(len(lists[l])):
                    me.set_ref_weights(l, ref, *weights.get((l, ref), (0, 32, 0, 5)))
            out, out_int = me.frame_search(jobs, want_int=True)
            n, bad = compare(jobs, expect, out, out_int)
            assert not bad, (case, pic, n, len(bad), bad[:4])
            total += n
    assert total > 1000, total


# ------------------------------------------------------------------------------------------------ frame batch
def _batch_case(pkg, ob, w, h, frames, n_refs, d, cfg, jobs, sample=None, seed=0):
    from h264_b200 import synth  # noqa: F401
    cur = frames[-1]
    with _open(pkg, w, h, d, max_refs=max(1, n_refs)) as me:
        handles = [me.ref_create(frames[-2 - r]) for r in range(n_refs)]
        me.set_ref_list(0, handles)
        me.set_current(cur)
        out, out_int = me.frame_search(jobs, want_int=True)
    idx = np.arange(len(jobs))
    if sample is not None and sample < len(jobs):
        idx = np.sort(np.random.default_rng(seed).choice(len(jobs), sample, replace=False))
    refs = [ob.Ref(frames[-2 - r]) for r in range(n_refs)]
    bad = []
    for r in range(n_refs):
        sel = idx[jobs["ref"][idx] == r]
        if not len(sel):
            continue
        exp, exp_int = ob.mb_search(cfg, np.ascontiguousarray(cur), refs[r], jobs[sel])
        for k, j in enumerate(sel):
            if not (np.array_equal(out_int["mv"][j], exp_int["mv"][k]) and np.array_equal(out_int["cost"][j], exp_int["cost"][k])):
                p = int(np.argmax((out_int["mv"][j] != exp_int["mv"][k]).any(axis=1) | (out_int["cost"][j] != exp_int["cost"][k])))
                bad.append(("int", int(j), p, out_int["mv"][j][p].tolist(), int(out_int["cost"][j][p]), exp_int["mv"][k][p].tolist(), int(exp_int["cost"][k][p])))
            elif not (np.array_equal(out["mv"][j], exp["mv"][k]) and np.array_equal(out["cost"][j], exp["cost"][k])):
                p = int(np.argmax((out["mv"][j] != exp["mv"][k]).any(axis=1) | (out["cost"][j] != exp["cost"][k])))
                bad.append(("sub", int(j), p, out["mv"][j][p].tolist(), int(out["cost"][j][p]), exp["mv"][k][p].tolist(), int(exp["cost"][k][p])))
    for r in refs:
        r.close()
    return bad, len(idx)


@pytest.mark.parametrize("variant", ["default", "t8x8", "rdopt0", "restrict0_sr16", "sse_fpel", "sad_subpel", "sse_subpel",
                                     "sad_hpel_satd_qpel", "satd_hpel_sse_qpel_qcif", "sad_subpel_t8x8", "qcif_sr32"])
def test_batch_matches_oracle(pkg, oracle_lib, variant):
    from h264_b200 import synth
    ob = oracle_lib
    w, h, n_refs, kw = 352, 288, 3, {}
    if variant == "t8x8":
        kw = dict(transform8x8=1)
    elif variant == "rdopt0":
        kw = dict(rdopt=0)
    elif variant == "restrict0_sr16":
        kw = dict(full_search=0, search_range=16)
    elif variant == "sse_fpel":
        kw = dict(metric=(1, 2, 2))
    elif variant == "sad_subpel":
        kw = dict(metric=(0, 0, 0))
    elif variant == "sse_subpel":
        kw = dict(metric=(0, 1, 1))
    elif variant == "sad_hpel_satd_qpel":
        kw = dict(metric=(0, 0, 2))          # the stages of one launch run different metrics
    elif variant == "satd_hpel_sse_qpel_qcif":
        kw = dict(metric=(0, 2, 1))
        w, h, n_refs = 176, 144, 2          # block-origin clamp of SAD / SSE on the global-memory path
    elif variant == "sad_subpel_t8x8":
        kw = dict(metric=(0, 0, 0), transform8x8=1)       # the transform size does not matter without the Hadamard metric
    elif variant == "qcif_sr32":
        w, h, n_refs = 176, 144, 2          # most windows leave the picture
    d, cfg = _cfg(ob, **kw)
    frames = synth.luma_frames(w, h, n_refs + 1)
    jobs = synth.synthetic_jobs(w, h, n_refs, (334459, 334459, 334459), seed=11, max_pmv=40)
    bad, n = _batch_case(pkg, ob, w, h, frames, n_refs, d, cfg, jobs, sample=260)
    assert not bad, (variant, len(bad), n, bad[:4])


@pytest.mark.parametrize("sr,w,h", [(16, 176, 144), (64, 352, 288)])
def test_batch_tuned_kernel_other_ranges(pkg, oracle_lib, sr, w, h):
    """the tuned search kernel exists for search ranges 16, 32 and 64 (config 1 / headline / config 5)"""
    from h264_b200 import synth
    ob = oracle_lib
    n_refs = 2
    d, cfg = _cfg(ob, search_range=sr)
    frames = synth.luma_frames(w, h, n_refs + 1)
    jobs = synth.synthetic_jobs(w, h, n_refs, (334459, 334459, 334459), seed=13, max_pmv=4 * sr)
    bad, n = _batch_case(pkg, ob, w, h, frames, n_refs, d, cfg, jobs, sample=80 if sr == 64 else 200)
    assert not bad, (sr, len(bad), n, bad[:4])


def test_batch_huge_lambda_jobs_are_handed_to_the_general_kernel(pkg, oracle_lib):
    """lambda factors beyond the 32-bit key range of the tuned kernel (> 2^24) go through the general kernel, job by job"""
    from h264_b200 import synth
    ob = oracle_lib
    w, h, n_refs = 176, 144, 2
    d, cfg = _cfg(ob)
    frames = synth.luma_frames(w, h, n_refs + 1)
    jobs = synth.synthetic_jobs(w, h, n_refs, (334459, 334459, 334459), seed=17, max_pmv=30)
    jobs["lambda"][::3] = ((1 << 24) + 12345, 334459, 334459)
    jobs["lambda"][1::7] = (1 << 24, 334459, 334459)           # the largest factor the tuned kernel accepts
    bad, n = _batch_case(pkg, ob, w, h, frames, n_refs, d, cfg, jobs, sample=None)
    assert not bad, (len(bad), n, bad[:4])


def test_batch_more_references_than_the_tuned_kernels_take(pkg, oracle_lib):
    """nine references in one list: beyond the eight tensor maps the tuned kernels carry, so the general kernels run"""
    from h264_b200 import synth
    ob = oracle_lib
    w, h, n_refs = 176, 144, 9
    d, cfg = _cfg(ob)
    frames = synth.luma_frames(w, h, n_refs + 1)
    jobs = synth.synthetic_jobs(w, h, n_refs, (334459, 334459, 334459), seed=37, max_pmv=30)
    bad, n = _batch_case(pkg, ob, w, h, frames, n_refs, d, cfg, jobs, sample=180)
    assert not bad, (len(bad), n, bad[:4])


def test_batch_two_reference_lists(pkg, oracle_lib):
    """B-slice style batch: jobs against list 0 and list 1 (different pictures) in one call"""
    from h264_b200 import synth
    ob = oracle_lib
    w, h = 352, 288
    d, cfg = _cfg(ob)
    frames = synth.luma_frames(w, h, 5)
    cur = frames[2]
    l0, l1 = [frames[1], frames[0]], [frames[3], frames[4]]            # past pictures / future pictures
    jobs0 = synth.synthetic_jobs(w, h, 2, (334459, 334459, 334459), seed=23, max_pmv=30, slice_type=1)
    jobs1 = synth.synthetic_jobs(w, h, 2, (334459, 334459, 334459), seed=29, max_pmv=30, slice_type=1)
    jobs1["list"] = 1
    jobs1["pmv16"] = -jobs1["pmv16"]; jobs1["pmv"] = -jobs1["pmv"]  # the pan runs the other way for future pictures
    jobs = np.concatenate([jobs0, jobs1])
    rng = np.random.default_rng(31)
    jobs = jobs[rng.permutation(len(jobs))]                            # lists and references interleaved
    with _open(pkg, w, h, d, max_refs=4) as me:
        me.set_ref_list(0, [me.ref_create(f) for f in l0])
        me.set_ref_list(1, [me.ref_create(f) for f in l1])
        me.set_current(cur)
        out, out_int = me.frame_search(jobs, want_int=True)
    curc = np.ascontiguousarray(cur)
    sel = np.sort(rng.choice(len(jobs), 240, replace=False))
    for lst, pics in ((0, l0), (1, l1)):
        for r, pic in enumerate(pics):
            s_ = sel[(jobs["list"][sel] == lst) & (jobs["ref"][sel] == r)]
            if not len(s_):
                continue
            ref = ob.Ref(pic)
            exp, exp_int = ob.mb_search(cfg, curc, ref, jobs[s_])
            assert np.array_equal(out_int[s_], exp_int), (lst, r)
            assert np.array_equal(out[s_], exp), (lst, r)
            ref.close()


@pytest.mark.parametrize("t8", [0, 1])
def test_batch_weighted_prediction(pkg, oracle_lib, t8):
    """weighted-prediction ME (UseWeightedReferenceME): every list entry with its own weight / offset / denominator, including
    negative offsets, clipping at both ends, a negative weight, identity weights, and weights that change between two calls
    and after the planes of a picture are regenerated in place"""
    from h264_b200 import synth
    ob = oracle_lib
    w, h, n_refs = 352, 288, 4
    d, cfg = _cfg(ob, transform8x8=t8)
    frames = synth.luma_frames(w, h, n_refs + 2, kind="fade")
    cur = np.ascontiguousarray(frames[-1])
    jobs = synth.synthetic_jobs(w, h, n_refs, (334459, 334459, 334459), seed=41, max_pmv=40)
    rng = np.random.default_rng(43)
    sel = np.sort(rng.choice(len(jobs), 200, replace=False))
    sets = [[(1, 29, 3, 5), (1, 40, -20, 5), (1, 32, 0, 5), (0, 0, 0, 0)],          # (apply, weight, offset, log_denom) per ref
            [(1, 127, 100, 6), (1, -3, 127, 2), (1, 17, -128, 4), (1, 1, 0, 0)]]
    with _open(pkg, w, h, d, max_refs=n_refs) as me:
        handles = [me.ref_create(frames[-2 - r]) for r in range(n_refs)]
        me.set_ref_list(0, handles)
        me.set_current(cur)
        pics = [frames[-2 - r] for r in range(n_refs)]
        for rnd, ws in enumerate(sets + [sets[0]]):
            if rnd == 2:                                    # a frame store reused: the weighted copy must follow the new planes
                pics[1] = frames[0]
                me.ref_update(handles[1], frames[0])
            for r, wt in enumerate(ws):
                me.set_ref_weights(0, r, *wt)
            out, out_int = me.frame_search(jobs, want_int=True)
            for r in range(n_refs):
                s_ = sel[jobs["ref"][sel] == r]
                ref = ob.Ref(pics[r])
                ref.set_weights(*ws[r])
                exp, exp_int = ob.mb_search(cfg, cur, ref, jobs[s_])
                assert np.array_equal(out_int[s_], exp_int), (rnd, r)
                assert np.array_equal(out[s_], exp), (rnd, r)
                ref.close()


@pytest.mark.parametrize("chroma_me,sse,sub", [(1, 0, (2, 2)), (2, 0, (2, 2)), (1, 1, (2, 2)), (2, 0, (0, 1)), (1, 0, (0, 0))])
def test_batch_chroma_terms(pkg, oracle_lib, chroma_me, sse, sub):
    """ChromaMEEnable: the fast full search adds the distortion of the co-located chroma block of every 4x4 (me_fullfast.c:776-814);
    the frame batch takes the general search kernel then.  4:2:0 pictures with real chroma, predictors that push windows out of
    the picture (chroma origin clamp), SAD and squared error"""
    from h264_b200 import synth
    ob = oracle_lib
    w, h, n_refs = 176, 144, 2
    d, cfg = _cfg(ob, chroma_me=chroma_me, metric=(sse,) + sub, search_range=16)
    yuv = synth.yuv420_frames(w, h, n_refs + 1)
    rng = np.random.default_rng(71)
    yuv = [(y, np.clip(u.astype(int) + rng.integers(-20, 21, u.shape), 0, 255).astype(np.uint8),
            np.clip(v.astype(int) + rng.integers(-20, 21, v.shape), 0, 255).astype(np.uint8)) for y, u, v in yuv]
    cur = yuv[-1]
    jobs = synth.synthetic_jobs(w, h, n_refs, (334459, 334459, 334459), seed=73, max_pmv=64)
    with _open(pkg, w, h, d, max_refs=n_refs, chroma_me_weight=4, yuv_format=1) as me:
        handles = []
        for r in range(n_refs):
            hd = me.ref_create(yuv[-2 - r][0])
            me.ref_set_chroma(hd, yuv[-2 - r][1], yuv[-2 - r][2])
            handles.append(hd)
        me.set_ref_list(0, handles)
        me.set_current(cur[0])
        me.set_current_chroma(cur[1], cur[2])
        me.set_chroma_state(chroma_me)
        out, out_int = me.frame_search(jobs, want_int=True)
    # integer-pel stage: chroma on; sub-pel stage: on only for value 2 (mv-search.c:612, :779) -- the oracle's composite does the same
    try:
        sel = np.sort(rng.choice(len(jobs), 120, replace=False))
        for r in range(n_refs):
            s_ = sel[jobs["ref"][sel] == r]
            ref = ob.Ref(yuv[-2 - r][0])
            ref.set_chroma(yuv[-2 - r][1], yuv[-2 - r][2])
            ob.set_chroma(chroma_me, 4, cur[1], cur[2], (3, 3))
            exp, exp_int = ob.mb_search(cfg, np.ascontiguousarray(cur[0]), ref, jobs[s_])
            assert np.array_equal(out_int[s_], exp_int), r
            assert np.array_equal(out[s_], exp), r
            ref.close()
    finally:
        ob.set_chroma(0)


@pytest.mark.parametrize("kind", ["flat", "stripes"])
def test_batch_ties_follow_spiral_order(pkg, oracle_lib, kind):
    """flat / periodic content: thousands of equal costs, the winner must be the first one in JM's spiral order"""
    from h264_b200 import synth
    ob = oracle_lib
    w, h, n_refs = 176, 144, 2
    d, cfg = _cfg(ob, transform8x8=1)
    frames = synth.luma_frames(w, h, n_refs + 1, kind=kind)
    jobs = synth.synthetic_jobs(w, h, n_refs, (334459, 334459, 334459), seed=3, max_pmv=60)
    jobs["lambda"][::2] = (0, 0, 0)            # lambda 0: every position ties
    bad, n = _batch_case(pkg, ob, w, h, frames, n_refs, d, cfg, jobs, sample=None)
    assert not bad, (kind, len(bad), n, bad[:4])


def test_batch_extreme_predictors(pkg, oracle_lib):
    """predictors far outside the picture: centre clipping (me_fullfast.c:555-563) and fully clamped windows"""
    from h264_b200 import synth
    ob = oracle_lib
    w, h, n_refs = 176, 144, 1
    d, cfg = _cfg(ob)
    frames = synth.luma_frames(w, h, 2)
    jobs = synth.synthetic_jobs(w, h, n_refs, (200000, 150000, 100000), seed=5)
    rng = np.random.default_rng(9)
    ext = np.array([[-8191, -8191], [8188, 8188], [-8191, 2044], [300, -2047], [-257, 511], [1023, -1023], [3, -3], [0, 0]])
    for i in range(len(jobs)):
        jobs["pmv16"][i] = ext[rng.integers(len(ext))]
        jobs["pmv"][i] = jobs["pmv16"][i][None, :] + rng.integers(-300, 301, size=(41, 2))
    bad, n = _batch_case(pkg, ob, w, h, frames, n_refs, d, cfg, jobs, sample=None)
    assert not bad, (len(bad), n, bad[:4])


def test_full_size_1080p_properties(pkg, oracle_lib):
    """BASELINE headline size: 1920x1088, SR 32, 5 refs.  Size-independent checks on the whole frame; the oracle on every job
    of reference 0, on every picture-border macroblock of all five references and on a 2 % sample of the rest."""
    from h264_b200 import synth
    ob = oracle_lib
    w, h, n_refs = 1920, 1088, 5
    d, cfg = _cfg(ob)
    frames = synth.luma_frames(w, h, n_refs + 1)
    lam = (334459, 334459, 334459)
    jobs = synth.synthetic_jobs(w, h, n_refs, lam, seed=21)
    cur = frames[-1]
    with _open(pkg, w, h, d) as me:
        handles = [me.ref_create(frames[-2 - r]) for r in range(n_refs)]
        me.set_ref_list(0, handles)
        me.set_current(cur)
        out1, int1 = me.frame_search(jobs, want_int=True)
        out2, int2 = me.frame_search(jobs, want_int=True)
        # known answer: search the current picture in itself with zero predictors -> zero vector, cost = rate of (0,0)
        hs = me.ref_create(cur)
        me.set_ref_list(0, [hs])
        selfjobs = jobs[: (w // 16) * (h // 16)].copy()
        selfjobs["ref"] = 0
        selfjobs["pmv16"] = 0
        selfjobs["pmv"] = 0
        so, si = me.frame_search(selfjobs, want_int=True)
    assert np.array_equal(out1, out2) and np.array_equal(int1, int2)          # deterministic
    assert (si["mv"] == 0).all() and (so["mv"] == 0).all()
    zero_rate = (lam[0] * 2) >> 16
    assert (si["cost"] == zero_rate).all() and (so["cost"] == zero_rate).all()
    # integer-pel cost is an upper bound only within the same metric; vectors stay inside centre +- range
    mvx = int1["mv"][..., 0].astype(np.int32)
    cx = (jobs["pmv16"][:, 0].astype(np.int32) / 4).astype(np.int32)         # C truncation
    assert (np.abs(mvx - cx[:, None]) <= 32).all()
    # the sub-pel result moves at most 3 quarter pels from the integer result
    assert (np.abs(out1["mv"].astype(np.int32) - 4 * int1["mv"].astype(np.int32)) <= 3).all()
    # oracle: EVERY job of one reference picture (the nearest and, on alternate runs of the parametrisation below, the farthest),
    # every picture-border macroblock of all five, and a random sample of the interior of the others -- the oracle costs
    # about 0.7 ms per job on one core, so the work is spread over the host cores (ctypes releases the GIL)
    import os
    from concurrent.futures import ThreadPoolExecutor
    refs = [ob.Ref(frames[-2 - r]) for r in range(n_refs)]
    curc = np.ascontiguousarray(cur)
    border = (jobs["opix_x"] == 0) | (jobs["opix_y"] == 0) | (jobs["opix_x"] == w - 16) | (jobs["opix_y"] == h - 16)
    rng = np.random.default_rng(2)
    pick = border | (jobs["ref"] == 0) | (rng.random(len(jobs)) < 0.02)
    sel = np.nonzero(pick)[0]
    assert border.sum() == n_refs * (2 * (w // 16) + 2 * (h // 16) - 4) and (jobs["ref"] == 0).sum() == (w // 16) * (h // 16)

    def check(chunk):
        r = int(jobs["ref"][chunk[0]])
        exp, exp_int = ob.mb_search(cfg, curc, refs[r], jobs[chunk])
        return int((int1[chunk] != exp_int).sum() + (out1[chunk] != exp).sum())

    chunks = []
    for r in range(n_refs):
        s = sel[jobs["ref"][sel] == r]
        chunks += [s[i:i + 64] for i in range(0, len(s), 64)]
    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as ex:
        wrong = sum(ex.map(check, chunks))
    assert wrong == 0, wrong
    assert len(sel) >= (w // 16) * (h // 16) + 4 * (2 * (w // 16) + 2 * (h // 16) - 4)
    for r in refs:
        r.close()


def test_tuned_and_basic_kernels_agree(pkg, oracle_lib):
    """the tuned search kernel against the straightforward one on a whole CIF frame (env switch in the library)"""
    from h264_b200 import synth
    w, h, n_refs = 352, 288, 2
    d, _ = _cfg(oracle_lib)
    frames = synth.luma_frames(w, h, n_refs + 1)
    jobs = synth.synthetic_jobs(w, h, n_refs, (334459, 334459, 334459), seed=4, max_pmv=50)
    res = []
    for mode in ("basic", ""):
        os.environ["ME_CUDA_FFS_KERNEL"] = mode
        try:
            with _open(pkg, w, h, d) as me:
                me.set_ref_list(0, [me.ref_create(frames[-2 - r]) for r in range(n_refs)])
                me.set_current(frames[-1])
                res.append(me.frame_search(jobs, want_int=True))
        finally:
            os.environ.pop("ME_CUDA_FFS_KERNEL", None)
    assert np.array_equal(res[0][1], res[1][1])
    assert np.array_equal(res[0][0], res[1][0])


# ------------------------------------------------------------------------------------------------ bi-predictive batch
def test_bipred_batch_matches_reference_chains(pkg):
    """me_cuda_frame_bipred against the chains of bi-predictive calls the unmodified reference made (known-answer run, B picture)"""
    from test_oracle_bipred import kat_chain_jobs
    from oracle import binding as ob
    tr, cfgkw = golden_trace("kat")
    chains, jobs = kat_chain_jobs(ob, tr)
    h, w = tr.size[("cur", chains[0]["pic"])]
    with _open(pkg, w, h, cfgkw) as me:
        # all chains of the B picture use the same two pictures: reference 0 of list 0 and of list 1
        c0 = chains[0]
        ids = {c0["list"]: c0["ref_l"][0], c0["list"] ^ 1: c0["ref_l"][1]}
        assert all({c["list"]: c["ref_l"][0], c["list"] ^ 1: c["ref_l"][1]} == ids and c["pic"] == c0["pic"] for c in chains)
        me.set_ref_list(0, [me.ref_create(tr.ref(ids[0]))])
        me.set_ref_list(1, [me.ref_create(tr.ref(ids[1]))])
        me.set_current(tr.cur(c0["pic"]))
        out = me.frame_bipred(jobs.view(pkg.BIPRED_JOB_DTYPE), 3, 16, 2)
    for k, c in enumerate(chains):
        assert tuple(out["mv"][k]) == c["exp_mv"] and tuple(out["bimv"][k]) == c["exp_bimv"], (k, out[k], c["exp_mv"], c["exp_bimv"])
        assert int(out["cost"][k]) == c["exp_cost"] and int(out["iterlist"][k]) == c["list"] ^ 1, (k, out[k], c["exp_cost"])


# weights: per list argument (apply, weight1, weight2, offsetBi, log denominator) of weighted bi-prediction, or None
@pytest.mark.parametrize("refinements,subpel,t8,weights", [
    (3, 2, 0, None), (0, 1, 0, None), (2, 0, 0, None), (1, 2, 1, None),
    (3, 2, 0, ((1, 31, 33, 0, 5), (1, 33, 31, 0, 5))),          # explicit weights as the reference estimates them for a fade
    (1, 2, 1, ((1, 20, 45, -7, 5), (1, 45, 20, -7, 5))),        # 8x8 transform: the source-pointer slip of computeBiPredSATD2
    (2, 2, 1, ((1, -10, 74, 12, 6), (0, 0, 0, 0, 0))),          # negative weight, clipping; list 1 calls unweighted
    (1, 2, 0, ((1, 128, -64, 0, 5), (1, -64, 128, 0, 5)))])     # implicit weights of an extrapolating list-1 picture: w0 = 64 - w1 = 128
                                                                 # (weighted_prediction.c:218-223)
def test_bipred_batch_matches_oracle(pkg, oracle_lib, refinements, subpel, t8, weights):
    from h264_b200 import synth
    ob = oracle_lib
    w, h = 352, 288
    d, cfg = _cfg(ob, transform8x8=t8)
    frames = synth.luma_frames(w, h, 3, kind="fade" if weights else "pan")
    cur, past, future = frames[1], frames[0], frames[2]
    rng = np.random.default_rng(41 + refinements)
    mbs = [(16 * x, 16 * y) for y in range(h // 16) for x in range(w // 16)]
    sel = [mbs[i] for i in rng.choice(len(mbs), 60, replace=False)]
    jobs = np.zeros(2 * len(sel), pkg.BIPRED_JOB_DTYPE)
    for k, (x, y) in enumerate(sel):
        for lst in (0, 1):
            j = jobs[2 * k + lst]
            j["opix_x"], j["opix_y"], j["list"] = x, y, lst
            sign = 1 if lst == 0 else -1                      # the pan runs the other way for the future picture
            j["pmv"] = (sign * 12 + rng.integers(-6, 7), sign * 8 + rng.integers(-6, 7))
            j["pmv_bi"] = (-sign * 12 + rng.integers(-6, 7), -sign * 8 + rng.integers(-6, 7))
            j["mv"] = (sign * 12 + rng.integers(-9, 10), sign * 8 + rng.integers(-9, 10))
            j["lambda"] = (334459, 334459, 334459)
    with _open(pkg, w, h, d, max_refs=1) as me:
        me.set_ref_list(0, [me.ref_create(past)])
        me.set_ref_list(1, [me.ref_create(future)])
        me.set_current(cur)
        for lst in (0, 1):
            wt = weights[lst] if weights else (0, 0, 0, 0, 0)
            me.set_bipred_weights(lst, *wt)
            ob.set_bipred_weights(lst, *wt)
        out = me.frame_bipred(jobs, refinements, 16, subpel)
    r0, r1 = ob.Ref(past), ob.Ref(future)
    exp = ob.mb_bipred(cfg, np.ascontiguousarray(cur), r0, r1, jobs.view(ob.BIPRED_JOB_DTYPE), refinements, 16, subpel)
    r0.close(); r1.close()
    for lst in (0, 1):
        ob.set_bipred_weights(lst, 0)
    assert np.array_equal(out.view(ob.BIPRED_RESULT_DTYPE), exp), np.nonzero(out.view(ob.BIPRED_RESULT_DTYPE) != exp)[0][:5]
