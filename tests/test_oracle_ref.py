"""CPU: the oracle against LIVE traces of the unmodified reference encoder (oracle/_ref/lencod_trace), when that binary
is present (it is built from /root/reference by `make -C oracle ref`; it travels to the GPU box but cannot be rebuilt there)."""
import hashlib
import os
import shutil
import subprocess

import pytest

from conftest import REF_DIR, has_ref_binaries
from replay import OracleAdapter, replay
from trace_util import Trace, cfg_from_trace

pytestmark = pytest.mark.skipif(not has_ref_binaries(), reason="oracle/_ref not built")


def _trace(tmp, args, yuv=None):
    fx = os.path.join(REF_DIR, "fixtures")
    for f in os.listdir(fx):
        if f.endswith(".cfg") or f.endswith(".yuv"):
            shutil.copy(os.path.join(fx, f), tmp)
    env = dict(os.environ, ME_TRACE_DIR=tmp, ME_TRACE_PLANES="1")
    r = subprocess.run([os.path.join(REF_DIR, "lencod_trace")] + args, cwd=tmp, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-1000:]
    return Trace(tmp)


def test_tracer_does_not_change_the_bitstream(tmp_path):
    _trace(str(tmp_path), ["-d", "encoder.cfg"])
    got = hashlib.md5(open(tmp_path / "test.264", "rb").read()).hexdigest()
    exp = hashlib.md5(open(os.path.join(REF_DIR, "fixtures", "golden_test.264"), "rb").read()).hexdigest()
    assert got == exp == "92402ba469c35bc09f97017a3903bdf9"


def test_oracle_reproduces_every_call_of_the_known_answer_run(tmp_path, oracle_lib):
    tr = _trace(str(tmp_path), ["-d", "encoder.cfg"])
    for i in tr.ref_ids():                       # all 16 planes of both reference pictures, pixel exact
        ref = oracle_lib.Ref(tr.ref(i))
        assert (ref.planes() == tr.ref_sub(i)).all()
        ref.close()
    # every 3rd record group keeps the CPU suite short; the golden test covers a different subset
    cfgkw = cfg_from_trace(tr, search_range=32, full_search=2)
    impl = OracleAdapter(tr, cfgkw)
    res = replay(tr, impl, max_calls=12000)
    impl.close()
    for name, (n, bad) in res.items():
        assert not bad, (name, n, bad[:3])


def test_oracle_reproduces_chroma_me_in_444(tmp_path, oracle_lib):
    """ChromaMEEnable = 2 in 4:4:4 common mode: the chroma sub-images are plane sets made by the luma interpolator; every traced call
    (fast full search and bi-predictive integer-pel search carry chroma terms)"""
    tr = _trace(str(tmp_path), ["-d", "encoder.cfg", "-p", "InputFile=foreman_part_qcif_444.yuv", "-p", "YUVFormat=3", "-p", "ProfileIDC=244",
                                "-p", "NumberReferenceFrames=2", "-p", "SearchRange=16", "-p", "ChromaMEEnable=2"])
    assert tr.chroma_info() == dict(enable=2, weight=4, yuv_format=3, cw=176, ch=144)
    impl = OracleAdapter(tr, cfg_from_trace(tr, search_range=16, full_search=2))
    res = replay(tr, impl, max_calls=9000)
    impl.close()
    for name, (n, bad) in res.items():
        assert not bad, (name, n, bad[:3])
    assert res["ffs"][0] > 3000


@pytest.mark.parametrize("name,args,metric,clip", [
    ("epzs_weighted", ["-d", "encoder_main.cfg", "-p", "InputFile=syn.yuv", "-p", "SourceWidth=176", "-p", "SourceHeight=144", "-p", "FramesToBeEncoded=4",
                       "-p", "FrameSkip=0", "-p", "NumberBFrames=0", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=3", "-p", "WeightedPrediction=1",
                       "-p", "UseWeightedReferenceME=1", "-p", "SearchMode=3"], (0, 2, 2), 4),
    ("epzs_integer_grid", ["-d", "encoder.cfg", "-p", "SearchMode=3", "-p", "EPZSSubPelGrid=0", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2"], (0, 2, 2), 0),
    ("epzs_weighted_bipred", ["-d", "encoder.cfg", "-p", "InputFile=syn.yuv", "-p", "FramesToBeEncoded=3", "-p", "WeightedBiprediction=1", "-p", "WeightedPrediction=1",
                              "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2", "-p", "SearchMode=3"], (0, 2, 2), 5),
    ("umhex_chroma_me_1", ["-d", "encoder.cfg", "-p", "SearchMode=1", "-p", "ChromaMEEnable=1", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2"], (0, 2, 2), 0),
    ("umhex_simple_chroma_sse", ["-d", "encoder.cfg", "-p", "SearchMode=2", "-p", "ChromaMEEnable=2", "-p", "MEDistortionHPel=1", "-p", "MEDistortionQPel=1",
                                 "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2"], (0, 1, 1), 0),
    ("umhex_sad", ["-d", "encoder_baseline.cfg", "-p", "SearchMode=1", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2", "-p", "MEDistortionHPel=0",
                   "-p", "MEDistortionQPel=0", "-p", "RDOptimization=0"], (0, 0, 0), 0)])
def test_oracle_reproduces_the_pattern_searches_subpel_stages(tmp_path, oracle_lib, name, args, metric, clip):
    """SearchMode 1 / 2 / 3 of the live reference: every call of EPZSSubPelBlockMotionSearch, EPZSSubPelBlockSearchBiPred,
    UMHEXSubPelBlockMotionSearch and smpUMHEX[Full]SubPelBlockMotionSearch (all macroblocks), with weighted reference samples, chroma
    terms and the three metrics -- beyond what the committed fixtures (tests/golden: *_sub.npz) keep"""
    import h264_b200_pkg  # noqa: F401
    from h264_b200 import synth
    if clip:
        synth.write_yuv420(str(tmp_path / "syn.yuv"), 176, 144, clip, kind="fade")
    tr = _trace(str(tmp_path), args)
    impl = OracleAdapter(tr, cfg_from_trace(tr, search_range=16, full_search=2, metric=metric))
    res = replay(tr, impl)
    impl.close()
    for key, (n, bad) in res.items():
        assert not bad, (name, key, n, bad[:3])
    assert res["subpel_pat"][0] > 10000 and (res["bi_sub_pat"][0] > 300 or "bipred" not in name)


@pytest.mark.parametrize("bits,profile,extra,metric", [
    (12, 122, [], (0, 2, 2)),
    (10, 110, ["-p", "MEDistortionHPel=0", "-p", "MEDistortionQPel=1", "-p", "RDOptimization=0", "-p", "Transform8x8Mode=0"], (0, 0, 1))])
def test_oracle_reproduces_high_bit_depth_runs(tmp_path, oracle_lib, bits, profile, extra, metric):
    """more than 8 bits per sample: planes of the live reference pixel-exact in the 16-bit build of the oracle (clips at
    max_imgpel_value) and every traced call, for 12 bit (High 4:2:2) and for SAD / squared-error refinements without RDO at 10 bit"""
    import importlib.util
    import numpy as np
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    mg.write_hbd_clip(str(tmp_path / "hbd.yuv"), bits)
    tr = _trace(str(tmp_path), ["-d", "encoder.cfg", "-p", "InputFile=hbd.yuv", "-p", "BitDepthLuma=%d" % bits, "-p", "BitDepthChroma=%d" % bits,
                                "-p", "ProfileIDC=%d" % profile, "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2"] + extra)
    assert tr.bits == bits
    oracle_lib.set_bit_depth(bits)
    try:
        for i in tr.ref_ids():
            ref = oracle_lib.Ref(tr.ref(i))
            assert ref.planes().dtype == np.uint16 and (ref.planes() == tr.ref_sub(i)).all()
            ref.close()
    finally:
        oracle_lib.set_bit_depth(8)
    impl = OracleAdapter(tr, cfg_from_trace(tr, search_range=16, full_search=2, metric=metric))
    res = replay(tr, impl)
    impl.close()
    for key, (n, bad) in res.items():
        assert not bad, (bits, key, n, bad[:3])
    assert res["ffs"][0] > 15000 and res["bi_full"][0] > 500


@pytest.mark.parametrize("name,args", [
    ("weighted_pred", ["-d", "encoder_main.cfg", "-p", "SourceWidth=176", "-p", "SourceHeight=144", "-p", "FramesToBeEncoded=4", "-p", "FrameSkip=0",
                       "-p", "NumberBFrames=0", "-p", "NumberReferenceFrames=3", "-p", "WeightedPrediction=1", "-p", "UseWeightedReferenceME=1"]),
    ("weighted_bipred", ["-d", "encoder.cfg", "-p", "NumberReferenceFrames=2", "-p", "FramesToBeEncoded=3", "-p", "WeightedBiprediction=1",
                         "-p", "WeightedPrediction=1"])])
def test_oracle_reproduces_weighted_high_bit_depth_runs(tmp_path, oracle_lib, name, args):
    """10-bit fading clip: weighted reference samples / weighted bi-prediction with the clip at 1023 in the 16-bit build of the oracle"""
    import numpy as np
    import h264_b200_pkg  # noqa: F401
    from h264_b200 import synth
    from trace_util import REC_FFS, REC_BI_FULL
    synth.write_yuv420(str(tmp_path / "fade8.yuv"), 176, 144, 7, kind="fade")
    src = np.fromfile(str(tmp_path / "fade8.yuv"), np.uint8).astype(np.uint16)
    ((src << 2) + np.random.default_rng(5).integers(0, 4, src.shape).astype(np.uint16)).astype("<u2").tofile(str(tmp_path / "hbd.yuv"))
    tr = _trace(str(tmp_path), args + ["-p", "InputFile=hbd.yuv", "-p", "BitDepthLuma=10", "-p", "BitDepthChroma=10", "-p", "ProfileIDC=110",
                                      "-p", "SearchRange=16"])
    assert tr.bits == 10 and set(tr.of_type(REC_FFS)[:, 29].tolist()) - {32}, "the fading clip must produce non-default weights"
    if name == "weighted_bipred":
        assert (tr.of_type(REC_BI_FULL)[:, 29] > 0).all()
    impl = OracleAdapter(tr, cfg_from_trace(tr, search_range=16, full_search=2))
    res = replay(tr, impl)
    impl.close()
    for key, (n, bad) in res.items():
        assert not bad, (name, key, n, bad[:3])
    assert res["ffs"][0] > 20000


@pytest.mark.parametrize("case", ["epzs_sub", "umhex_sub", "smpumhex_sub", "epzs_chroma_sad", "hbd10"])
def test_new_tracer_wrappers_do_not_change_the_bitstream(tmp_path, case):
    """the fixtures of the pattern searches and of the 10-bit run were traced with lencod_trace (wrappers around the five pattern
    sub-pel functions, 16-bit dumps): the PLAIN reference encoder must write the bitstream whose md5 was committed with them"""
    import importlib.util
    import json
    from conftest import GOLDEN
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    fx = os.path.join(REF_DIR, "fixtures")
    for f in os.listdir(fx):
        if f.endswith(".cfg") or f.endswith(".yuv"):
            shutil.copy(os.path.join(fx, f), str(tmp_path))
    if case == "hbd10":
        mg.write_hbd_clip(str(tmp_path / "hbd.yuv"), 10)
        args = ["-d", "encoder.cfg", "-p", "InputFile=hbd.yuv", "-p", "BitDepthLuma=10", "-p", "BitDepthChroma=10", "-p", "ProfileIDC=110",
                "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2"]
    else:
        args = mg.PATTERN_CASES[case][0]
    r = subprocess.run([os.path.join(REF_DIR, "lencod")] + args, cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-800:]
    want = json.load(open(os.path.join(GOLDEN, "bitstream_md5.json")))[case]["test.264"]
    assert hashlib.md5(open(tmp_path / "test.264", "rb").read()).hexdigest() == want
