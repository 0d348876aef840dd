"""CPU: the oracle (oracle/me_oracle.c) against the committed golden vectors made from the unmodified reference."""
import json
import os
import zlib

import numpy as np
import pytest

from conftest import GOLDEN, GOLDEN_CASES, HBD_CASES, PATTERN_CASES, golden_trace
from replay import OracleAdapter, replay, pattern_counts


@pytest.mark.parametrize("case", sorted(GOLDEN_CASES) + sorted(PATTERN_CASES) + sorted(HBD_CASES))
def test_planes_match_reference_crc(oracle_lib, case):
    tr, _ = golden_trace(case)
    crc = json.load(open(os.path.join(GOLDEN, "planes_crc.json")))[case]
    oracle_lib.set_bit_depth(tr.bits)              # hbd10: the 16-bit build, clips at 1023
    try:
        for i in tr.ref_ids():
            ref = oracle_lib.Ref(tr.ref(i))
            for fy in range(4):
                for fx in range(4):
                    assert zlib.crc32(ref.plane(fy, fx).tobytes()) == crc[str(i)][fy][fx], (case, i, fy, fx)
            ref.close()
    finally:
        oracle_lib.set_bit_depth(8)


@pytest.mark.parametrize("case", sorted(GOLDEN_CASES))
def test_calls_match_reference_trace(oracle_lib, case):
    tr, cfgkw = golden_trace(case)
    impl = OracleAdapter(tr, cfgkw)
    res = replay(tr, impl)
    impl.close()
    for name, (n, bad) in res.items():
        assert not bad, (case, name, n, bad[:3])
    # the pattern-search trace (epzs) holds no integer-pel calls of the path: that stage stays the reference's own
    assert (res["ffs"][0] > 1000 or res["fullpel"][0] > 1000 or case == "epzs") and res["subpel"][0] > 1000
    if case == "kat":
        assert res["bi_full"][0] > 100 and res["bi_sub"][0] > 50


def test_high_bit_depth_calls_match_reference_trace(oracle_lib):
    """10-bit pictures (imgpel u16): the 16-bit build of the oracle against every kept call of the unmodified reference"""
    tr, cfgkw = golden_trace("hbd10")
    assert tr.bits == 10 and tr.ref(tr.ref_ids()[0]).dtype == np.uint16 and int(tr.ref(tr.ref_ids()[0]).max()) > 255
    impl = OracleAdapter(tr, cfgkw)
    res = replay(tr, impl)
    impl.close()
    for name, (n, bad) in res.items():
        assert not bad, (name, n, bad[:3])
    assert res["ffs"][0] > 3000 and res["subpel"][0] > 3000 and res["bi_full"][0] > 100 and res["bi_sub"][0] > 50


@pytest.mark.parametrize("case", sorted(PATTERN_CASES))
def test_pattern_subpel_calls_match_reference_trace(oracle_lib, case):
    """sub-pel stages of SearchMode 1 / 2 / 3 (EPZSSubPelBlockMotionSearch, EPZSSubPelBlockSearchBiPred, UMHEXSubPelBlockMotionSearch,
    smpUMHEXSubPelBlockMotionSearch, smpUMHEXFullSubPelBlockMotionSearch): every kept call of the unmodified reference"""
    tr, cfgkw = golden_trace(case)
    impl = OracleAdapter(tr, cfgkw)
    res = replay(tr, impl)
    impl.close()
    for name, (n, bad) in res.items():
        assert not bad, (case, name, n, bad[:3])
    pattern_counts(case, res)


def test_spiral_and_mvbits(oracle_lib):
    sx, sy = oracle_lib.spiral(3)
    # mv-search.c:366-393: centre, then per ring top/bottom rows (alternating), then left/right columns
    assert list(zip(sx[:9], sy[:9])) == [(0, 0), (0, -1), (0, 1), (-1, -1), (1, -1), (-1, 0), (1, 0), (-1, 1), (1, 1)]
    assert len(set(zip(sx.tolist(), sy.tolist()))) == 49
    L = oracle_lib.lib()
    # se(v) code lengths: 1, 3, 5, 5, 7 ...
    assert [L.ora_mvbits(d) for d in (0, 1, -1, 2, 3, 4, 7, 8, -1023)] == [1, 3, 3, 5, 5, 7, 7, 9, 21]


def test_hadamard_known_answers(oracle_lib):
    L = oracle_lib.lib()
    d = np.zeros(16, np.int32); d[0] = 16
    assert L.ora_hadamard4x4(d.ctypes.data) == (16 * 16 + 1) >> 1        # a delta spreads to all 16 coefficients
    d8 = np.full(64, 3, np.int32)
    assert L.ora_hadamard8x8(d8.ctypes.data) == (3 * 64 + 2) >> 2         # DC only
    rng = np.random.default_rng(1)
    for _ in range(20):                                                  # against a matrix Hadamard
        d = rng.integers(-255, 256, 16).astype(np.int32)
        H = np.array([[1, 1, 1, 1], [1, 1, -1, -1], [1, -1, -1, 1], [1, -1, 1, -1]])
        exp = (np.abs(H @ d.reshape(4, 4) @ H.T).sum() + 1) >> 1
        assert L.ora_hadamard4x4(d.ctypes.data) == exp
