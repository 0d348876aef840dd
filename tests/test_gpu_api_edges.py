"""GPU: error behaviour and edge cases of the C ABI (the reference's failure mode is error(text, code): loud, never silent)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg(me_lib_path):
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import h264_b200_pkg  # noqa: F401
    import h264_b200
    return h264_b200


def test_init_rejects_unsupported_configurations(pkg):
    from h264_b200 import MeCuda, MeCudaError
    for kw in (dict(width=180, height=144), dict(width=176, height=150), dict(width=0, height=144)):
        with pytest.raises(MeCudaError):
            MeCuda(kw["width"], kw["height"])
    with pytest.raises(MeCudaError):
        MeCuda(176, 144, search_range=65)
    with pytest.raises(MeCudaError):
        MeCuda(176, 144, max_num_references=33)
    with pytest.raises(MeCudaError):
        MeCuda(176, 144, device=99)


def test_calls_before_init_and_bad_arguments_fail_loudly(pkg):
    from h264_b200 import MeCuda, MeCudaError, synth
    w, h = 176, 144
    frames = synth.luma_frames(w, h, 2)
    with MeCuda(w, h, max_num_references=2) as me:
        hd = me.ref_create(frames[0])
        with pytest.raises(MeCudaError):
            me.ref_create(frames[0][:, :160])                    # picture size differs from the initialised one
        with pytest.raises(MeCudaError):
            me.set_ref_list(0, [hd, 12345])                      # unknown handle
        me.set_ref_list(0, [hd])
        me.set_current(frames[1])
        jobs = synth.synthetic_jobs(w, h, 1, (334459,) * 3, seed=1)
        bad = jobs.copy(); bad["ref"][5] = 1                     # reference index beyond the list
        with pytest.raises(MeCudaError):
            me.frame_search(bad)
        bad = jobs.copy(); bad["opix_x"][3] = 8                  # not a macroblock origin
        with pytest.raises(MeCudaError):
            me.frame_search(bad)
        bad = jobs.copy(); bad["opix_y"][7] = h                  # outside the picture
        with pytest.raises(MeCudaError):
            me.frame_search(bad)
        out = me.frame_search(jobs[:0])                          # empty batch: nothing to do, no error
        assert len(out) == 0
        me.ref_destroy(hd)
        with pytest.raises(MeCudaError):
            me.ref_destroy(hd)                                   # double free
        with pytest.raises(MeCudaError):
            me.frame_search(jobs)                                # the list now holds a destroyed picture
    me.close()                                                   # idempotent
    with pytest.raises(MeCudaError):
        me.set_current(frames[1])                                # after shutdown


def test_single_job_and_ragged_batches(pkg, oracle_lib):
    """batches of 1, 2, 3, 127, 129 jobs (fewer jobs than CTAs, odd tails of the persistent kernels)"""
    from h264_b200 import MeCuda, synth
    ob = oracle_lib
    w, h = 352, 288
    frames = synth.luma_frames(w, h, 2)
    cfg = ob.make_cfg()
    jobs = synth.synthetic_jobs(w, h, 1, (334459,) * 3, seed=9, max_pmv=30)
    ref = ob.Ref(frames[0])
    cur = np.ascontiguousarray(frames[1])
    with MeCuda(w, h, max_num_references=1) as me:
        me.set_ref_list(0, [me.ref_create(frames[0])])
        me.set_current(frames[1])
        for n in (1, 2, 3, 127, 129):
            sel = jobs[200:200 + n]
            out, out_int = me.frame_search(sel, want_int=True)
            exp, exp_int = ob.mb_search(cfg, cur, ref, sel)
            assert np.array_equal(out_int, exp_int), n
            assert np.array_equal(out, exp), n
    ref.close()


def test_reference_slot_reuse_and_in_place_update(pkg, oracle_lib):
    """a decoded-picture-buffer style life cycle: create, destroy, re-create in the freed slot, update in place"""
    from h264_b200 import MeCuda, synth
    ob = oracle_lib
    w, h = 176, 144
    frames = synth.luma_frames(w, h, 4)
    with MeCuda(w, h, max_num_references=2) as me:
        a = me.ref_create(frames[0]); b = me.ref_create(frames[1])
        me.ref_destroy(a)
        c = me.ref_create(frames[2])
        assert c == a                                            # the freed slot is reused
        r2 = ob.Ref(frames[2])
        assert np.array_equal(me.ref_planes(c), r2.planes())
        me.ref_update(b, frames[3])
        r3 = ob.Ref(frames[3])
        assert np.array_equal(me.ref_planes(b), r3.planes())
        r2.close(); r3.close()


def test_weight_arguments_are_checked(pkg):
    """me_cuda_set_ref_weights / me_cuda_set_bipred_weights: ranges of the slice header syntax (weights / offsets -128..127,
    log denominator 0..7; bi-predictive weights -128..128, the range implicit weighting produces, weighted_prediction.c:218-223); the bi-predictive batch reads the pictures' own planes whatever the list entries' weights are"""
    from h264_b200 import MeCuda, MeCudaError, synth
    from h264_b200.me_cuda import BIPRED_JOB_DTYPE
    w, h = 176, 144
    frames = synth.luma_frames(w, h, 3)
    with MeCuda(w, h, max_num_references=2) as me:
        me.set_ref_list(0, [me.ref_create(frames[0])])
        me.set_ref_list(1, [me.ref_create(frames[2])])
        me.set_current(frames[1])
        for bad in ((0, 0, 1, 128, 0, 5), (0, 0, 1, 32, -129, 5), (0, 0, 1, 32, 0, 8), (2, 0, 1, 32, 0, 5), (0, 32, 1, 32, 0, 5)):
            with pytest.raises(MeCudaError):
                me.set_ref_weights(*bad)
        jobs = np.zeros(1, BIPRED_JOB_DTYPE)
        jobs["lambda"] = 334459
        for bad in ((2, 1, 32, 32, 0, 5), (0, 1, 129, 32, 0, 5), (0, 1, 32, -129, 0, 5), (1, 1, 32, 32, 128, 5), (1, 1, 32, 32, 0, 8)):
            with pytest.raises(MeCudaError):
                me.set_bipred_weights(*bad)
        me.set_bipred_weights(0, 1, 128, -64, 0, 5)          # w0 = 128, w1 = -64: an extrapolating implicit pair is accepted
        me.set_bipred_weights(0, 0, 32, 32, 0, 5)
        plain = me.frame_bipred(jobs, 3, 16, 2)
        me.set_ref_weights(1, 0, 1, 30, 2, 5)                 # uni-predictive weights of the same list entry: no effect here
        assert np.array_equal(me.frame_bipred(jobs, 3, 16, 2), plain)


def test_device_entry_points_on_a_callers_stream(pkg):
    """The *_device calls take a caller's stream; the library orders it against its own stream with events on both sides of
    every such call (enter_stream / leave_stream in me_cuda_host.cu), so program order = device order whichever streams the calls
    name.  A long-running kernel on the caller's stream in front of every producer makes a missing dependency visible: the
    consumers on the library stream would read planes / the current picture before they are written."""
    import torch
    from h264_b200 import MeCuda, synth
    from h264_b200.me_cuda import JOB_DTYPE, RESULT_DTYPE
    w, h, n_refs = 352, 288, 2
    frames = synth.luma_frames(w, h, n_refs + 2)
    jobs = synth.synthetic_jobs(w, h, n_refs, (334459,) * 3, seed=3)
    with MeCuda(w, h, search_range=16, max_num_references=n_refs) as me:      # the library stream only
        hd = [me.ref_create(frames[-2 - r]) for r in range(n_refs)]
        me.set_ref_list(0, hd)
        me.set_current(frames[-1])
        want, want_int = me.frame_search(jobs, want_int=True)
        me.ref_update(hd[0], frames[0])                                        # frame store reused
        want2 = me.frame_search(jobs)
    dev = torch.device("cuda:0")
    side = torch.cuda.Stream(device=dev)
    burn = torch.empty(1 << 26, dtype=torch.float32, device=dev)

    def delay():                                                               # ~ a millisecond of work queued on the side stream
        with torch.cuda.stream(side):
            for _ in range(8):
                burn.mul_(1.0001)

    d_frames = [torch.from_numpy(np.ascontiguousarray(f)).to(dev) for f in frames]
    d_jobs = torch.from_numpy(jobs.view(np.uint8).reshape(len(jobs), JOB_DTYPE.itemsize).copy()).to(dev)
    d_out = torch.zeros(len(jobs) * RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_int = torch.zeros_like(d_out)
    torch.cuda.synchronize()
    with MeCuda(w, h, search_range=16, max_num_references=n_refs) as me:
        sp = side.cuda_stream
        hd = []
        for r in range(n_refs):
            delay()
            hd.append(me.ref_create_device(d_frames[-2 - r].data_ptr(), w, stream=sp))      # planes written on the side stream ...
        me.set_ref_list(0, hd)
        delay()
        me.set_current_device(d_frames[-1].data_ptr(), w, stream=sp)
        got, got_int = me.frame_search(jobs, want_int=True)                                 # ... read on the library stream
        assert np.array_equal(got, want) and np.array_equal(got_int, want_int)
        # producer on the library stream, consumer on the side stream, scratch shared with the library stream
        me.ref_update(hd[0], frames[0])
        me.frame_search_device(d_jobs.data_ptr(), len(jobs), d_out.data_ptr(), d_int.data_ptr(), stream=sp)
        delay()
        me.ref_update_device(hd[0], d_frames[-2].data_ptr(), w, stream=sp)                  # must wait for the search on the side stream
        side.synchronize()
        got2 = d_out.cpu().numpy().view(RESULT_DTYPE)
        assert np.array_equal(got2, want2)
        got3 = me.frame_search(jobs)                                                        # sees the planes of the last update
        assert np.array_equal(got3, want)


def test_pinned_buffers_are_searched_in_place(pkg, oracle_lib):
    """me_cuda_frame_search with page-locked job / result arrays: the kernels read and write them in place (no copies, the job array is
    validated by a kernel in front of the searches).  Same results as with pageable arrays; an invalid job is refused with the same
    message and nothing is searched."""
    import torch
    from h264_b200 import MeCuda, MeCudaError, synth
    from h264_b200.me_cuda import JOB_DTYPE, RESULT_DTYPE
    w, h = 352, 288
    frames = synth.luma_frames(w, h, 3)
    jobs = synth.synthetic_jobs(w, h, 2, (334459,) * 3, seed=5)
    with MeCuda(w, h, search_range=32, max_num_references=2) as me:
        me.set_ref_list(0, [me.ref_create(frames[1]), me.ref_create(frames[0])])
        me.set_current(frames[2])
        ref_out, ref_int = me.frame_search(jobs, want_int=True)                      # pageable numpy arrays: the copying path
        pj = torch.from_numpy(jobs.view(np.uint8).reshape(len(jobs), JOB_DTYPE.itemsize).copy()).pin_memory()
        po = torch.zeros((len(jobs), RESULT_DTYPE.itemsize), dtype=torch.uint8).pin_memory()
        pi = torch.zeros((len(jobs), RESULT_DTYPE.itemsize), dtype=torch.uint8).pin_memory()
        launches = me.launch_count()
        me._check(me.L.me_cuda_frame_search(pj.data_ptr(), len(jobs), po.data_ptr(), pi.data_ptr()))
        assert me.launch_count() - launches <= 4                                       # check + search (+ hand-back) + sub-pel: one batch, no chunks
        assert np.array_equal(po.numpy().reshape(-1).view(RESULT_DTYPE), ref_out)
        assert np.array_equal(pi.numpy().reshape(-1).view(RESULT_DTYPE), ref_int)
        # an invalid reference index in the middle of the array
        bad = jobs.copy()
        bad["ref"][len(bad) // 2] = 7
        pb = torch.from_numpy(bad.view(np.uint8).reshape(len(bad), JOB_DTYPE.itemsize).copy()).pin_memory()
        po.zero_()
        with pytest.raises(MeCudaError, match="job %d" % (len(bad) // 2)):
            me._check(me.L.me_cuda_frame_search(pb.data_ptr(), len(bad), po.data_ptr(), None))
        assert not po.numpy().any()                                                    # nothing was written
        # and the context still works
        me._check(me.L.me_cuda_frame_search(pj.data_ptr(), len(jobs), po.data_ptr(), None))
        assert np.array_equal(po.numpy().reshape(-1).view(RESULT_DTYPE), ref_out)
