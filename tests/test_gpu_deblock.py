"""GPU: the deblocking filter against the oracle's restatement of DeblockFrame / EdgeLoopLumaNormal /
EdgeLoopChromaNormal (loopFilter.c:87-253, :529-672, :814-921) on the same seeded pictures and boundary strengths: bit-exact,
4:2:0 and 4:0:0, strengths of every kind (strong intra edges, inner edges, unfiltered edges, 8x8-transform macroblocks), the whole
quantiser range with filter offsets, and the plane set built from the filtered picture that stayed on the device."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import h264_b200_pkg  # noqa: F401
    import h264_b200
    return h264_b200


def make_case(w, h, seed, blocky=True):
    from h264_b200.me_cuda import DEBLOCK_MB_DTYPE
    from h264_b200 import synth
    rng = np.random.default_rng(seed)
    mbw, mbh = w // 16, h // 16
    y = synth.luma_frames(w, h, 1, seed=seed)[0].astype(np.int32)
    if blocky:      # coding-like steps at the 4x4 grid: the filter has something to decide on
        y += np.kron(rng.integers(-6, 7, (h // 4, w // 4)), np.ones((4, 4), np.int32))
    y = np.clip(y, 0, 255).astype(np.uint8)
    u = np.clip(128 + np.kron(rng.integers(-8, 9, (h // 8, w // 8)), np.ones((4, 4), np.int32)) + rng.integers(-2, 3, (h // 2, w // 2)), 0, 255).astype(np.uint8)
    v = np.clip(128 + np.kron(rng.integers(-8, 9, (h // 8, w // 8)), np.ones((4, 4), np.int32)) + rng.integers(-2, 3, (h // 2, w // 2)), 0, 255).astype(np.uint8)
    mbs = np.zeros(mbw * mbh, DEBLOCK_MB_DTYPE)
    mbs["qp"] = rng.integers(0, 52, mbw * mbh)
    mbs["qpc"] = np.clip(mbs["qp"][:, None] + rng.integers(-6, 1, (mbw * mbh, 2)), 0, 51)
    mbs["alpha_off"] = rng.integers(-6, 7) * 2 if seed % 3 == 0 else 0
    mbs["beta_off"] = rng.integers(-6, 7) * 2 if seed % 3 == 0 else 0
    for a in range(mbw * mbh):
        mx, my = a % mbw, a // mbw
        kind = rng.integers(0, 4)                       # 0 intra, 1 inter with residual, 2 inter skip-like, 3 8x8 transform
        for d in range(2):
            for e in range(4):
                if e == 0 and (mx == 0 if d == 0 else my == 0):
                    continue                              # picture boundary: not filtered
                if kind == 0:
                    s = np.full(16, 4 if e == 0 else 3)
                elif kind == 2:
                    s = np.repeat(rng.integers(0, 2, 4), 4)
                else:
                    s = np.repeat(rng.integers(0, 3, 4), 4)
                mbs["strength"][a, d, e] = s
        if kind == 3:
            mbs["luma_skip"][a] = 0b10101010           # edges 1 and 3 of both directions
    return y, u, v, mbs


@pytest.mark.parametrize("w,h,seed,chroma", [(176, 144, 1, True), (352, 288, 2, True), (176, 144, 3, False), (1920, 1088, 4, True), (64, 48, 6, True)])
def test_deblock_matches_oracle(pkg, oracle_lib, w, h, seed, chroma):
    ob = oracle_lib
    y, u, v, mbs = make_case(w, h, seed)
    ey, eu, ev = ob.deblock_frame(y, u if chroma else None, v if chroma else None, mbs)
    assert (ey != y).mean() > (0.02 if w >= 176 else 0.0)   # the filter did change a good share of the samples
    with pkg.MeCuda(w, h, search_range=16, max_num_references=1) as me:
        gy, gu, gv = y.astype(np.uint16), u.astype(np.uint16), v.astype(np.uint16)
        me.deblock_frame(gy, gu if chroma else None, gv if chroma else None, mbs)
        assert np.array_equal(gy, ey)
        if chroma:
            assert np.array_equal(gu, eu) and np.array_equal(gv, ev)
        else:
            assert np.array_equal(gu, u) and np.array_equal(gv, v)
        # the reference picture built from the filtered luma that stayed on the device
        hd = me.ref_create_deblocked()
        ref = ob.Ref(ey)
        assert np.array_equal(me.ref_planes(hd), ref.planes())
        ref.close()
        # a second picture through the same context (buffers reused), all edges off: nothing changes
        mbs0 = mbs.copy()
        mbs0["strength"] = 0
        g2 = y.astype(np.uint16)
        me.deblock_frame(g2, None, None, mbs0)
        assert np.array_equal(g2, y)


def test_deblock_argument_checks(pkg):
    from h264_b200.me_cuda import DEBLOCK_MB_DTYPE
    with pkg.MeCuda(176, 144, search_range=16, max_num_references=1) as me:
        with pytest.raises(pkg.MeCudaError):
            me.ref_create_deblocked()                     # nothing deblocked yet
        y = np.zeros((144, 176), np.uint16)
        u = np.zeros((72, 176), np.uint16)                # 4:2:2 chroma: refused
        with pytest.raises(pkg.MeCudaError):
            me.deblock_frame(y, u, u.copy(), np.zeros(99, DEBLOCK_MB_DTYPE))
