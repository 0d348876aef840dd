"""GPU: interlace -- reference pictures and current pictures that are FIELDS of the initialised frame size
(dpb_split_field, mbuffer.c:2971; field_picture, image.c:1015), and the per-macroblock field views of a frame that stays on the
device (MBAFF field macroblocks, macroblock.c:165-190).  The oracle is simply run on the field arrays (even / odd rows)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LAM = (334459, 334459, 334459)


@pytest.fixture(scope="module")
def pkg(me_lib_path):
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import h264_b200_pkg  # noqa: F401
    import h264_b200
    return h264_b200


def _fields(frame):
    return np.ascontiguousarray(frame[0::2]), np.ascontiguousarray(frame[1::2])


def test_field_reference_pictures_have_their_own_size(pkg, oracle_lib):
    """plane sets of half-height pictures next to full-height ones, pixel exact against the oracle"""
    from h264_b200 import MeCuda, MeCudaError, synth
    w, h = 176, 160
    frames = synth.luma_frames(w, h, 2)
    with MeCuda(w, h, search_range=16, max_num_references=4) as me:
        hf = me.ref_create(frames[0])
        top, bot = _fields(frames[0])
        ht, hb = me.ref_create(top), me.ref_create(bot)
        for hd, pic in ((hf, frames[0]), (ht, top), (hb, bot)):
            ref = oracle_lib.Ref(pic)
            assert np.array_equal(me.ref_planes(hd, height=pic.shape[0]), ref.planes()), pic.shape
            ref.close()
        with pytest.raises(MeCudaError):
            me.ref_create(frames[0][:48])                        # neither a frame nor a field of the initialised size
        with pytest.raises(MeCudaError):
            me.ref_update(ht, frames[1])                         # a frame into the slot of a field
        me.ref_update(ht, _fields(frames[1])[0])                 # a field into the slot of a field: fine
        ref = oracle_lib.Ref(_fields(frames[1])[0])
        assert np.array_equal(me.ref_planes(ht, height=h // 2), ref.planes())
        ref.close()
    with MeCuda(176, 144, search_range=16) as me:                # 72 rows are not whole macroblocks: no fields of this frame size
        with pytest.raises(MeCudaError):
            me.ref_create(frames[0][:72])


@pytest.mark.parametrize("t8", [0, 1])
def test_field_views_of_the_current_frame(pkg, oracle_lib, t8):
    """the frame is uploaded once; macroblocks then search as frame macroblocks (view 0, frame references) or as field macroblocks
    (view 1 / 2, field references with list changes in between): block searches and sub-pel refinements equal the oracle's on the
    corresponding arrays"""
    from h264_b200 import MeCuda, synth
    w, h = 176, 160
    frames = synth.luma_frames(w, h, 4)
    cur = frames[3]
    cfgd = dict(search_range=16, rdopt=1, full_search=2, metric=(0, 2, 2), transform8x8=t8, mv_limit_min_y=-511, mv_limit_max_y=511)
    cfg = oracle_lib.make_cfg(**cfgd)
    rng = np.random.default_rng(7 + t8)
    with MeCuda(w, h, search_range=16, max_num_references=4, transform8x8=t8) as me:
        frame_refs = [frames[2], frames[1]]
        field_refs = [f for fr in frame_refs for f in _fields(fr)]              # top, bottom of the newest, then of the older picture
        h_frame = [me.ref_create(p) for p in frame_refs]
        h_field = [me.ref_create(p) for p in field_refs]
        o_frame = [oracle_lib.Ref(p) for p in frame_refs]
        o_field = [oracle_lib.Ref(p) for p in field_refs]
        me.set_current(cur)
        me.set_slice_type(0)
        views = {0: (cur, h_frame, o_frame), 1: (_fields(cur)[0], h_field, o_field), 2: (_fields(cur)[1], h_field, o_field)}
        n_checked = 0
        for it in range(18):
            view = it % 3
            pic, handles, orefs = views[view]
            ph = pic.shape[0]
            me.select_current_view(view)
            me.set_ref_list(0, handles)
            ox, oy = 16 * int(rng.integers(0, w // 16)), 16 * int(rng.integers(0, ph // 16))
            if it < 6:                                           # picture corners and the last macroblock row of the field
                ox, oy = [(0, 0), (w - 16, 0), (0, ph - 16), (w - 16, ph - 16), (80, ph - 16), (0, 32)][it]
            me.mb_begin(ox, oy)
            for ref in range(len(handles)):
                pmv16 = (int(rng.integers(-24, 25)), int(rng.integers(-24, 25)))
                me.ffs_setup(0, ref, *pmv16)
                cx, cy = oracle_lib.search_centre(cfg, ref, ox, oy, *pmv16)
                srange = oracle_lib.lib().ora_setup_range(cfg, ref)
                surf = oracle_lib.ffs_surfaces(cfg, pic, orefs[ref], ox, oy, cx, cy, srange)
                for bt, blk in ((1, 0), (2, 8), (3, 2), (4, 10), (5, 6), (6, 9), (7, 15)):
                    ppx, ppy = ox + 4 * (blk % 4), oy + 4 * (blk // 4)
                    pmv = (int(rng.integers(-30, 31)), int(rng.integers(-30, 31)))
                    got = me.FastFullPelBlockMotionSearch(ref, 0, ppx, ppy, bt, pmv[0], pmv[1], srange, 2147483647, LAM[0])
                    exp = oracle_lib.ffs_block(cfg, surf[oracle_lib.part_index(bt, blk)], srange, ox, oy, cx, cy, pmv[0], pmv[1], srange,
                                               2147483647, LAM[0])
                    assert tuple(got) == tuple(exp), (view, ox, oy, ref, bt, got, exp)
                    got2 = me.SubPelBlockMotionSearch(ref, 0, ppx, ppy, bt, pmv[0], pmv[1], 4 * got[0], 4 * got[1], 9, 9, 2147483647, LAM)
                    exp2 = oracle_lib.subpel(cfg, pic, orefs[ref], ppx, ppy, bt, ref, 0, pmv[0], pmv[1], 4 * exp[0], 4 * exp[1],
                                             2147483647, LAM)
                    assert tuple(got2) == tuple(exp2), (view, ox, oy, ref, bt, got2, exp2)
                    n_checked += 1
        assert n_checked == 6 * 7 * len(h_frame) + 12 * 7 * len(h_field)          # six visits per view, seven block types per reference
        for r in o_frame + o_field:
            r.close()


def test_frame_batch_on_an_uploaded_field_and_refusal_under_a_view(pkg, oracle_lib):
    """the frame batch on a field picture: the field itself is uploaded (half the rows) with field references; under a field VIEW of
    an uploaded frame the batch is refused loudly"""
    from h264_b200 import MeCuda, MeCudaError, synth
    w, h, n_refs = 176, 160, 2
    frames = synth.luma_frames(w, h, n_refs + 1)
    cfg = oracle_lib.make_cfg(search_range=16)
    with MeCuda(w, h, search_range=16, max_num_references=n_refs) as me:
        for parity in (0, 1):
            cur = _fields(frames[-1])[parity]
            refs = [_fields(frames[-2 - r])[parity ^ (r & 1)] for r in range(n_refs)]
            handles = [me.ref_create(p) for p in refs]
            me.set_ref_list(0, handles)
            me.set_current(cur)
            jobs = synth.synthetic_jobs(w, h // 2, n_refs, LAM, seed=3 + parity)
            out, out_int = me.frame_search(jobs, want_int=True)
            for r in range(n_refs):
                oref = oracle_lib.Ref(refs[r])
                sel = np.nonzero(jobs["ref"] == r)[0]
                exp, exp_int = oracle_lib.mb_search(cfg, cur, oref, jobs[sel])
                assert np.array_equal(out_int[sel], exp_int) and np.array_equal(out[sel], exp), (parity, r)
                oref.close()
            bad = jobs.copy(); bad["opix_y"][2] = h // 2                     # below the last row of the FIELD
            with pytest.raises(MeCudaError):
                me.frame_search(bad)
            for hd in handles:
                me.ref_destroy(hd)
        hd = me.ref_create(frames[0])
        me.set_ref_list(0, [hd])
        me.set_current(frames[-1])
        me.select_current_view(1)
        with pytest.raises(MeCudaError):
            me.frame_search(synth.synthetic_jobs(w, h, 1, LAM, seed=1))
        me.select_current_view(0)
        assert len(me.frame_search(synth.synthetic_jobs(w, h, 1, LAM, seed=1))) == (w // 16) * (h // 16)
        with pytest.raises(MeCudaError):
            me.select_current_view(3)
