"""CPU: the host-side mirror of GetSkipCostMB / GetDirectCost8x8 / BIDPartitionCost (h.264_b200/mc.py) composes the three
reference functions from per-8x8-region costs.  Here the region primitive is played by the oracle (LumaPrediction +
distortion4x4 / distortion8x8, macroblock.c:832, me_distortion.c:76, :110), so that the composition itself -- which regions,
4x4 or 8x8 transform, motion-vector rate, the 'direct mode not allowed' case -- is checked against the oracle's restatement of
the whole functions (mv-search.c:1136, :1273, :1052) without a GPU.  tests/test_gpu_mc.py runs the same cases on the device."""
import ctypes as C

import numpy as np
import pytest

W, H = 176, 144


class OracleRegions:
    """stands in for MeCuda.pred_cost8x8"""

    def __init__(self, ob, cur, l0, l1):
        self.ob, self.cur, self.l0, self.l1 = ob, cur, l0, l1
        self.L = ob.lib()
        self.L.ora_distortion4x4.argtypes = [C.c_int, C.c_void_p]
        self.L.ora_distortion8x8.argtypes = [C.c_int, C.c_void_p]

    def pred_cost8x8(self, blocks, metric):
        n = len(blocks) // 4
        c4, c8 = np.zeros(n, np.int32), np.zeros(n, np.int32)
        for r in range(n):
            d = np.zeros((8, 8), np.int32)
            s4, allowed = 0, True
            for k in range(4):
                b = blocks[4 * r + k]
                if b["p_dir"] < 0:
                    allowed = False
                    break
                p = self.ob.luma_prediction(self.l0, self.l1, b, 4, 4).astype(np.int32)
                dd = self.cur[b["pix_y"]:b["pix_y"] + 4, b["pix_x"]:b["pix_x"] + 4].astype(np.int32) - p
                d[4 * (k >> 1):4 * (k >> 1) + 4, 4 * (k & 1):4 * (k & 1) + 4] = dd
                dd = np.ascontiguousarray(dd.reshape(16))
                s4 += self.L.ora_distortion4x4(metric, dd.ctypes.data)
            if not allowed:
                c4[r] = c8[r] = 2147483647
                continue
            d64 = np.ascontiguousarray(d.reshape(64))
            c4[r], c8[r] = s4, self.L.ora_distortion8x8(metric, d64.ctypes.data)
        return c4, c8


def make_blocks(rng, n, size, weighted, allow_none=False):
    import h264_b200_pkg  # noqa: F401
    from h264_b200.mc import PRED_BLOCK_DTYPE
    b = np.zeros(n, PRED_BLOCK_DTYPE)
    b["pix_x"] = rng.integers(0, (W - size) // 4 + 1, n) * 4
    b["pix_y"] = rng.integers(0, (H - size) // 4 + 1, n) * 4
    b["p_dir"] = rng.integers(-1 if allow_none else 0, 3, n)
    b["ref"] = rng.integers(0, 2, (n, 2))
    far = rng.random(n) < 0.25
    mv = rng.integers(-40, 41, (n, 2, 2))
    mv[far] = rng.integers(-900, 901, (int(far.sum()), 2, 2))               # far outside: the clamp of UMVLine4X decides
    b["mv"] = mv
    if weighted:
        b["apply_weights"] = rng.integers(0, 2, n)
        b["weight"] = rng.integers(-20, 90, (n, 2))
        b["offset"] = rng.integers(-20, 21, n)
        b["log_denom"] = rng.integers(1, 8, n)
    return b


@pytest.mark.parametrize("metric", [0, 1, 2])
def test_cost_functions_compose_from_region_costs(oracle_lib, metric):
    import h264_b200_pkg  # noqa: F401
    from h264_b200 import mc, synth
    ob = oracle_lib
    frames = synth.luma_frames(W, H, 5)
    refs = [ob.Ref(frames[i]) for i in range(4)]
    cur = frames[4]
    me = OracleRegions(ob, cur, refs[:2], refs[2:])
    rng = np.random.default_rng(1 + metric)
    for rdopt, t8 in ((1, 0), (0, 1), (1, 1)):
        for trial in range(4):
            ox, oy = 16 * int(rng.integers(0, W // 16)), 16 * int(rng.integers(0, H // 16))
            blocks = make_blocks(rng, 16, 4, trial % 2)
            for by in range(4):
                for bx in range(4):
                    blocks["pix_x"][4 * by + bx], blocks["pix_y"][4 * by + bx] = ox + 4 * bx, oy + 4 * by
            skip = blocks.copy()
            skip["p_dir"], skip["ref"], skip["mv"] = 0, 0, blocks["mv"][0]
            assert mc.GetSkipCostMB(me, skip, rdopt, t8, metric) == ob.get_skip_cost_mb(cur, refs[:2], refs[2:], skip, rdopt, t8, metric)
            direct = blocks.copy()
            if trial == 3:
                direct["p_dir"][5] = -1
            for block in range(4):
                exp = ob.get_direct_cost8x8(cur, refs[:2], refs[2:], direct, block, rdopt, t8, metric, cost8x8=11)
                assert mc.GetDirectCost8x8(me, direct, block, rdopt, t8, metric, cost8x8=11) == exp
            bid = blocks.copy()
            bid["p_dir"] = 2
            pmv = rng.integers(-60, 61, (16, 2, 2)).astype(np.int16)
            for blocktype in range(1, 8):
                for block8x8 in range(4 if blocktype >= 4 else 2 if blocktype >= 2 else 1):
                    got = mc.BIDPartitionCost(me, bid, pmv, blocktype, block8x8, 334459, t8, metric)
                    assert got == ob.bid_partition_cost(cur, refs[:2], refs[2:], bid, pmv, blocktype, block8x8, 334459, t8, metric)
    for r in refs:
        r.close()
