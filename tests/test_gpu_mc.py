"""GPU: motion-compensated prediction and the mode-decision costs built on it against the oracle's
restatement of LumaPrediction (macroblock.c:832), GetSkipCostMB / GetDirectCost8x8 / BIDPartitionCost (mv-search.c:1136, :1273,
:1052): same seeded inputs, bit-exact.  Vectors reach far outside the picture (origin clamp of UMVLine4X), both lists, weighted
forms, all three mode-decision metrics, with and without the 8x8 transform."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
W, H = 176, 144


@pytest.fixture(scope="module")
def pkg():
    import h264_b200_pkg  # noqa: F401
    import h264_b200
    return h264_b200


from test_mc_host import make_blocks as _blocks  # noqa: E402


def _setup(pkg, ob):
    from h264_b200 import synth
    frames = synth.luma_frames(W, H, 5)
    me = pkg.MeCuda(W, H, search_range=16, max_num_references=2)
    hs = [me.ref_create(frames[i]) for i in range(4)]
    me.set_ref_list(0, hs[:2])
    me.set_ref_list(1, hs[2:])
    me.set_current(frames[4])
    refs = [ob.Ref(frames[i]) for i in range(4)]
    return me, frames, refs


@pytest.mark.parametrize("size", [(4, 4), (8, 8), (16, 16), (16, 8), (4, 8)])
@pytest.mark.parametrize("weighted", [0, 1])
def test_luma_prediction_matches_oracle(pkg, oracle_lib, size, weighted):
    ob = oracle_lib
    me, frames, refs = _setup(pkg, ob)
    try:
        rng = np.random.default_rng(100 * size[0] + size[1] + weighted)
        b = _blocks(rng, 300, max(size), weighted)
        got = me.luma_prediction(b, size[0], size[1])
        for k in range(len(b)):
            exp = ob.luma_prediction(refs[:2], refs[2:], b[k], size[0], size[1])
            assert np.array_equal(got[k], exp), (k, b[k])
        # a batch beyond the mapped staging area takes the device-buffer route: same result
        big = np.tile(b, 12)
        assert np.array_equal(me.luma_prediction(big, size[0], size[1])[:len(b)], got)
    finally:
        me.close()
        for r in refs:
            r.close()


@pytest.mark.parametrize("metric", [0, 1, 2])
@pytest.mark.parametrize("rdopt,t8", [(1, 0), (0, 1), (1, 1), (0, 0)])
def test_mode_decision_costs_match_oracle(pkg, oracle_lib, metric, rdopt, t8):
    from h264_b200 import mc
    ob = oracle_lib
    me, frames, refs = _setup(pkg, ob)
    cur = frames[4]
    try:
        rng = np.random.default_rng(7 + 10 * metric + 3 * rdopt + t8)
        for trial in range(24):
            ox, oy = 16 * int(rng.integers(0, W // 16)), 16 * int(rng.integers(0, H // 16))
            blocks = _blocks(rng, 16, 4, weighted=trial % 2, allow_none=False)
            for by in range(4):
                for bx in range(4):
                    blocks["pix_x"][4 * by + bx], blocks["pix_y"][4 * by + bx] = ox + 4 * bx, oy + 4 * by
            # GetSkipCostMB: list 0, reference 0, one vector for the macroblock
            skip = blocks.copy()
            skip["p_dir"], skip["ref"], skip["mv"] = 0, 0, blocks["mv"][0]
            assert mc.GetSkipCostMB(me, skip, rdopt, t8, metric) == ob.get_skip_cost_mb(cur, refs[:2], refs[2:], skip, rdopt, t8, metric)
            # GetDirectCost8x8: per-block direction / references / vectors; now and then a block without a direct mode
            direct = blocks.copy()
            if trial % 5 == 4:
                direct["p_dir"][int(rng.integers(0, 16))] = -1
            for block in range(4):
                exp = ob.get_direct_cost8x8(cur, refs[:2], refs[2:], direct, block, rdopt, t8, metric, cost8x8=11)
                assert mc.GetDirectCost8x8(me, direct, block, rdopt, t8, metric, cost8x8=11) == exp
            # BIDPartitionCost: both lists, the partition's vectors, every block type and 8x8 block
            bid = blocks.copy()
            bid["p_dir"] = 2
            pmv = rng.integers(-60, 61, (16, 2, 2)).astype(np.int16)
            for blocktype in range(1, 8):
                for block8x8 in range(4 if blocktype >= 4 else 2 if blocktype >= 2 else 1):
                    got = mc.BIDPartitionCost(me, bid, pmv, blocktype, block8x8, 334459, t8, metric)
                    exp = ob.bid_partition_cost(cur, refs[:2], refs[2:], bid, pmv, blocktype, block8x8, 334459, t8, metric)
                    assert got == exp, (trial, blocktype, block8x8, got, exp)
    finally:
        me.close()
        for r in refs:
            r.close()


def test_plane_regions_equal_the_planes(pkg, oracle_lib):
    """me_cuda_ref_download_region: what the binding pages into StorablePicture::imgY_sub on demand"""
    ob = oracle_lib
    me, frames, refs = _setup(pkg, ob)
    try:
        rng = np.random.default_rng(3)
        hs = [1, 2, 3, 4]
        planes = refs[0].planes()
        h0 = me.ref_create(frames[0])
        for _ in range(40):
            fy, fx = int(rng.integers(0, 4)), int(rng.integers(0, 4))
            w, h = int(rng.integers(1, 97)), int(rng.integers(1, 65))
            x0, y0 = int(rng.integers(0, W + 40 - w + 1)), int(rng.integers(0, H + 40 - h + 1))
            got = me.ref_region(h0, fy, fx, x0, y0, w, h)
            assert np.array_equal(got, planes[fy][fx][y0:y0 + h, x0:x0 + w].astype(np.uint16))
        big = me.ref_region(h0, 2, 1, 0, 0, W + 40, H + 40)            # beyond the mapped staging area
        assert np.array_equal(big, planes[2][1].astype(np.uint16))
        with pytest.raises(pkg.MeCudaError):
            me.ref_region(h0, 0, 0, W, H, 64, 64)
        del hs
    finally:
        me.close()
        for r in refs:
            r.close()
