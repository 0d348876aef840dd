"""GPU: the unmodified reference encoder with its hot path routed through libme_cuda.so (oracle/_ref/lencod_cuda, the
--wrap shim of h.264_b200/shim/jm_shim.c) must write byte-identical bitstreams and reconstructions."""
import hashlib
import json
import os
import shutil
import subprocess

import pytest

from conftest import GOLDEN, REF_DIR, ROOT

pytestmark = pytest.mark.gpu


def _md5(p):
    return hashlib.md5(open(p, "rb").read()).hexdigest()


def _have():
    """oracle/_ref (the reference encoder binaries, built where the reference tree is mounted) travels to the GPU box with the
    repository snapshot.  Without it these tests cannot run: on a box WITH a GPU that is a failure, not a skip -- a green run
    must have executed them."""
    ok = os.path.exists(os.path.join(REF_DIR, "lencod_cuda")) and os.path.exists(os.path.join(REF_DIR, "fixtures", "encoder.cfg"))
    if not ok:
        import torch
        if torch.cuda.is_available() and not os.environ.get("ME_ALLOW_MISSING_REF"):
            raise RuntimeError("oracle/_ref/lencod_cuda is missing on a GPU box: build it with `make -C oracle ref` where /root/reference "
                               "is mounted (it is git-ignored but not gpurun-ignored)")
    return ok


def _prep(tmp):
    fx = os.path.join(REF_DIR, "fixtures")
    for f in os.listdir(fx):
        if f.endswith(".cfg") or f.endswith(".yuv"):
            shutil.copy(os.path.join(fx, f), tmp)


def _run(exe, args, cwd, env=None):
    e = dict(os.environ)
    e["ME_CUDA_LIB"] = os.path.join(ROOT, "h.264_b200", "libme_cuda.so")
    for k in ("ME_CUDA_DISABLE", "ME_CUDA_VERIFY", "ME_CUDA_PMS", "ME_CUDA_PMS_MBAFF", "ME_CUDA_MC", "ME_CUDA_HOST_PLANES", "ME_CUDA_POISON_PLANES", "ME_CUDA_DEBLOCK"):       # a stray switch must not turn the comparison into reference-vs-reference
        e.pop(k, None)
    if env:
        e.update(env)
    r = subprocess.run([os.path.join(REF_DIR, exe)] + args, cwd=cwd, env=e, capture_output=True, text=True, timeout=3000)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    return r.stdout, r.stderr


def _served(stderr, want=("plane sets",)):
    """the shim's exit report: how many calls of each wrapped function the device answered.  Returns the counts; asserts that
    the run was a GPU run (plane sets generated on the device, and every kind in `want` served at least once)."""
    import re
    m = re.search(r"served on the device: (\d+) plane sets, (\d+) ffs, (\d+) fullpel, (\d+) subpel, (\d+) bi_full, (\d+) bi_sub calls", stderr)
    assert m, "no service report from the shim (pass-through run?): " + stderr[-800:]
    c = dict(zip(["plane sets", "ffs", "fullpel", "subpel", "bi_full", "bi_sub"], map(int, m.groups())))
    p = re.search(r"partition-search launches (\d+), block searches answered from them (\d+), argument mismatches (\d+)", stderr)
    assert p, stderr[-800:]
    c["pms_launches"], c["pms_served"], c["pms_missed"] = map(int, p.groups())
    q = re.search(r"mode-decision costs on the device: (\d+) BIDPartitionCost, (\d+) GetDirectCost8x8 calls; plane tiles paged to the host: (\d+) \((\d+) bytes", stderr)
    assert q, stderr[-800:]
    c["bid_cost"], c["direct_cost"], c["paged_tiles"], c["paged_bytes"] = map(int, q.groups())
    d = re.search(r"pictures deblocked on the device: (\d+), reference plane sets built from them without an upload: (\d+)", stderr)
    assert d, stderr[-800:]
    c["deblocked"], c["deblocked_reused"] = map(int, d.groups())
    for k in want:
        assert c[k] > 0, (k, c)
    return c


@pytest.mark.skipif(not _have(), reason="oracle/_ref/lencod_cuda not built (needs the reference tree at build time)")
def test_known_answer_bitstream_is_byte_identical(tmp_path, me_lib_path):
    """bin/encoder.cfg: the run behind the reference's committed bin/test.264 / bin/test_rec.yuv"""
    _prep(str(tmp_path))
    _, err = _run("lencod_cuda", ["-d", "encoder.cfg"], str(tmp_path))
    c = _served(err, ("plane sets", "ffs", "subpel", "bi_full", "bi_sub", "bid_cost", "paged_tiles", "deblocked", "deblocked_reused"))
    assert c["pms_launches"] > 0 and c["pms_served"] > 0, c
    # the planes stayed on the device: the host paged in what it read, tile by tile, never more than the 16 x 216 x 184 samples x 2 bytes
    # per reference picture a copy-back would have moved (on this QCIF clip most tiles are touched; at 1080p a small share is)
    assert 0 < c["paged_bytes"] <= c["plane sets"] * 16 * 216 * 184 * 2, c
    md5 = json.load(open(os.path.join(GOLDEN, "bitstream_md5.json")))["kat"]
    assert _md5(str(tmp_path / "test.264")) == md5["test.264"] == _md5(os.path.join(REF_DIR, "fixtures", "golden_test.264"))
    assert _md5(str(tmp_path / "test_rec.yuv")) == md5["test_rec.yuv"] == _md5(os.path.join(REF_DIR, "fixtures", "golden_test_rec.yuv"))


@pytest.mark.skipif(not _have(), reason="oracle/_ref/lencod_cuda not built")
@pytest.mark.parametrize("args,want", [
    (["-d", "encoder.cfg"], "bid_cost"),                                                       # B pictures, Transform8x8Mode = 1, RDO
    (["-d", "encoder.cfg", "-p", "RDOptimization=0", "-p", "SearchRange=16"], "direct_cost"),    # rdopt = 0: GetDirectCost8x8 with the 8x8 costs
    (["-d", "encoder_main.cfg", "-p", "RDOptimization=0", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2", "-p", "WeightedBiprediction=1"],
     "direct_cost")])                                                                            # explicit weights in the predictions
def test_mode_decision_costs_and_paged_planes(tmp_path, me_lib_path, args, want):
    """SURVEY.md 8(f) row 2.  BIDPartitionCost / GetDirectCost8x8 (mv-search.c:1052, :1273) served by the library, the planes paged into
    imgY_sub on demand instead of copied back: (1) byte-identical output with the host plane sets poisoned first -- a host read
    that bypassed the paging would change the bitstream; (2) verify mode: the reference function runs next to every device call of the two
    cost functions (and of the searches) and every result is compared."""
    import re
    a, b = tmp_path / "gpu", tmp_path / "cpu"
    for d in (a, b):
        d.mkdir()
        _prep(str(d))
    _, err = _run("lencod_cuda", args, str(a), env={"ME_CUDA_POISON_PLANES": "1"})
    _run("lencod", args, str(b))
    assert _md5(str(a / "test.264")) == _md5(str(b / "test.264"))
    assert _md5(str(a / "test_rec.yuv")) == _md5(str(b / "test_rec.yuv"))
    c = _served(err, ("plane sets", "paged_tiles", "deblocked", want))
    _, err = _run("lencod_cuda", args, str(a), env={"ME_CUDA_VERIFY": "1"})
    m = re.search(r"verify mode: (\d+) calls checked, (\d+) mismatches", err)
    assert m and int(m.group(1)) > c["bid_cost"] + c["direct_cost"] and int(m.group(2)) == 0, err[-1500:]
    # the same run with the two switches off: reference host code for the costs, all planes copied back -- same output
    _, err = _run("lencod_cuda", args, str(a), env={"ME_CUDA_MC": "0", "ME_CUDA_HOST_PLANES": "1", "ME_CUDA_DEBLOCK": "0"})
    assert _md5(str(a / "test.264")) == _md5(str(b / "test.264"))
    c0 = _served(err, ("plane sets",))
    assert c0["bid_cost"] == 0 and c0["direct_cost"] == 0 and c0["paged_tiles"] == 0 and c0["deblocked"] == 0, c0


@pytest.mark.skipif(not _have(), reason="oracle/_ref/lencod_cuda not built")
@pytest.mark.parametrize("case", ["baseline_sr16_1ref", "main_cif_5ref", "cif_5ref_8frames", "rdopt0_restrict0", "yuv444_3ref_bframes", "sr64_2ref",
                                  "hd1080_sr32_ip", "weighted_pred_fade", "weighted_bipred_fade", "full_search_mode_rdopt0",
                                  "full_search_mode_satd", "yuv444_independent_planes", "chroma_me_2_bframes",
                                  "chroma_me_1_full_search_rdopt0", "chroma_me_1_yuv422", "chroma_me_2_sad_sse_subpel",
                                  "chroma_me_1_sad_subpel", "chroma_me_2_yuv444", "hadamard_fpel_bframes",
                                  "hadamard_fpel_chroma_me_1", "weighted_implicit_bipred_chroma_me_2",
                                  "weighted_pred_chroma_me_1", "umhex_plain_subpel", "umhex_simple_bipred_subpel",
                                  "epzs_plain_subpel", "epzs_subpel_grid_plain_subpel",
                                  "epzs_pattern_subpel", "epzs_pattern_subpel_weighted", "umhex_pattern_subpel_chroma_sad",
                                  "umhex_simple_pattern_subpel_rdopt0",
                                  "interlace_field_pictures", "interlace_picture_adaptive", "interlace_mbaff_field_macroblocks",
                                  "interlace_mbaff_adaptive", "interlace_picture_and_mbaff_adaptive_chroma_me", "interlace_field_pictures_epzs",
                                  "high_bit_depth_10_bframes", "high_bit_depth_12_bframes", "high_bit_depth_10_sad_sse_rdopt0",
                                  "high_bit_depth_10_field_pictures", "high_bit_depth_10_epzs", "high_bit_depth_10_umhex",
                                  "high_bit_depth_10_full_search_mode_rdopt0",
                                  "high_bit_depth_10_weighted_pred_fade", "high_bit_depth_10_weighted_bipred_fade"])
def test_configs_match_reference_encoder(tmp_path, me_lib_path, case):
    """BASELINE configs (reduced frame counts): lencod_cuda against the plain reference binary on the same input"""
    import sys
    sys.path.insert(0, ROOT)
    import h264_b200_pkg  # noqa: F401
    from h264_b200 import synth
    a, b = tmp_path / "gpu", tmp_path / "cpu"
    a.mkdir(); b.mkdir()
    if case == "baseline_sr16_1ref":
        args = ["-d", "encoder_baseline.cfg", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=1"]
    elif case == "main_cif_5ref":
        args = ["-d", "encoder_main.cfg", "-p", "InputFile=syn.yuv", "-p", "SourceWidth=352", "-p", "SourceHeight=288",
                "-p", "FramesToBeEncoded=4", "-p", "FrameSkip=0", "-p", "NumberBFrames=0", "-p", "SearchRange=32",
                "-p", "NumberReferenceFrames=5", "-p", "LevelIDC=51"]
    elif case == "cif_5ref_8frames":
        # BASELINE config 2 with ALL FIVE references in use: pictures 5..7 of eight search five reference pictures each
        args = ["-d", "encoder_main.cfg", "-p", "InputFile=syn.yuv", "-p", "SourceWidth=352", "-p", "SourceHeight=288",
                "-p", "FramesToBeEncoded=8", "-p", "FrameSkip=0", "-p", "NumberBFrames=0", "-p", "SearchRange=32",
                "-p", "NumberReferenceFrames=5", "-p", "LevelIDC=51"]
    elif case == "yuv444_3ref_bframes":
        # BASELINE config 4 on the reference's own 4:4:4 fixture: High 4:4:4, I-B-P with bi-predictive ME, 3 references
        args = ["-d", "encoder.cfg", "-p", "InputFile=foreman_part_qcif_444.yuv", "-p", "YUVFormat=3", "-p", "ProfileIDC=244",
                "-p", "NumberReferenceFrames=3"]
    elif case == "yuv444_independent_planes":
        # 4:4:4 independent mode (SeparateColourPlane = 1): three colour planes coded as luma pictures, each with its own
        # quarter-pel plane set per reference picture (UnifiedOneForthPix_JV, image.c:1672) and its own original plane
        args = ["-d", "encoder.cfg", "-p", "InputFile=foreman_part_qcif_444.yuv", "-p", "YUVFormat=3", "-p", "ProfileIDC=244",
                "-p", "SeparateColourPlane=1", "-p", "NumberReferenceFrames=2", "-p", "SearchRange=16"]
    elif case == "chroma_me_2_bframes":
        # SURVEY.md 8(f) row 1, chroma half: chroma terms in the fast full search and the bi-predictive integer-pel search
        args = ["-d", "encoder.cfg", "-p", "ChromaMEEnable=2", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2"]
    elif case == "chroma_me_1_full_search_rdopt0":
        args = ["-d", "encoder_main.cfg", "-p", "ChromaMEEnable=1", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2",
                "-p", "SearchMode=-1", "-p", "RDOptimization=0"]
    elif case == "chroma_me_2_sad_sse_subpel":
        # chroma terms also in the sub-pel stages (SAD half-pel, SSE quarter-pel) and in the bi-predictive sub-pel refinement
        args = ["-d", "encoder.cfg", "-p", "ChromaMEEnable=2", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2",
                "-p", "MEDistortionHPel=0", "-p", "MEDistortionQPel=1"]
    elif case == "chroma_me_1_sad_subpel":
        # value 1: the sub-pel stages run WITHOUT chroma although their metric would carry the term
        args = ["-d", "encoder_main.cfg", "-p", "ChromaMEEnable=1", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2",
                "-p", "MEDistortionHPel=0", "-p", "MEDistortionQPel=0"]
    elif case == "chroma_me_2_yuv444":
        # 4:4:4 common mode: the chroma sub-images are the quarter-pel plane sets of U and V (luma interpolator, image.c:1647-1657)
        args = ["-d", "encoder.cfg", "-p", "InputFile=foreman_part_qcif_444.yuv", "-p", "YUVFormat=3", "-p", "ProfileIDC=244",
                "-p", "NumberReferenceFrames=2", "-p", "SearchRange=16", "-p", "ChromaMEEnable=2"]
    elif case in ("hadamard_fpel_bframes", "hadamard_fpel_chroma_me_1"):
        # MEDistortionFPel = 2: squared error in the fast full search (me_fullfast.c:512), Hadamard SAD (8x8 transform) in the
        # bi-predictive integer-pel search (computeBiPredSATD1, me_distortion.c:824)
        args = ["-d", "encoder.cfg", "-p", "MEDistortionFPel=2", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2"]
        if case.endswith("chroma_me_1"):
            args += ["-p", "ChromaMEEnable=1"]
    elif case == "weighted_implicit_bipred_chroma_me_2":
        # two B pictures between references: implicit weights 43 / 21 for luma AND chroma in the bi-predictive searches
        args = ["-d", "encoder.cfg", "-p", "InputFile=syn.yuv", "-p", "SourceWidth=176", "-p", "SourceHeight=144",
                "-p", "FramesToBeEncoded=3", "-p", "FrameSkip=2", "-p", "NumberBFrames=2", "-p", "WeightedBiprediction=2",
                "-p", "WeightedPrediction=1", "-p", "ChromaMEEnable=2", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2"]
    elif case == "weighted_pred_chroma_me_1":
        args = ["-d", "encoder_main.cfg", "-p", "InputFile=syn.yuv", "-p", "SourceWidth=176", "-p", "SourceHeight=144",
                "-p", "FramesToBeEncoded=4", "-p", "FrameSkip=0", "-p", "NumberBFrames=0", "-p", "SearchRange=16",
                "-p", "NumberReferenceFrames=3", "-p", "WeightedPrediction=1", "-p", "ChromaMEEnable=1"]
    elif case == "chroma_me_1_yuv422":
        args = ["-d", "encoder_yuv422.cfg", "-p", "ChromaMEEnable=1", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2"]
    elif case == "hd1080_sr32_ip":
        # the headline picture size and search range (BASELINE config 3), one I and one P picture
        args = ["-d", "encoder_main.cfg", "-p", "InputFile=syn.yuv", "-p", "SourceWidth=1920", "-p", "SourceHeight=1088",
                "-p", "FramesToBeEncoded=2", "-p", "FrameSkip=0", "-p", "NumberBFrames=0", "-p", "SearchRange=32",
                "-p", "NumberReferenceFrames=5", "-p", "LevelIDC=51"]
    elif case == "weighted_pred_fade":
        # explicit weighted prediction with weighted-reference ME on a clip that fades
        args = ["-d", "encoder_main.cfg", "-p", "InputFile=syn.yuv", "-p", "SourceWidth=176", "-p", "SourceHeight=144",
                "-p", "FramesToBeEncoded=4", "-p", "FrameSkip=0", "-p", "NumberBFrames=0", "-p", "SearchRange=16",
                "-p", "NumberReferenceFrames=3", "-p", "WeightedPrediction=1", "-p", "UseWeightedReferenceME=1"]
    elif case == "weighted_bipred_fade":
        # I-B-P with explicit weighted bi-prediction and bi-predictive ME (computeBiPred*2), High profile with the 8x8 transform
        args = ["-d", "encoder.cfg", "-p", "InputFile=syn.yuv", "-p", "SourceWidth=176", "-p", "SourceHeight=144",
                "-p", "FramesToBeEncoded=3", "-p", "WeightedBiprediction=1", "-p", "WeightedPrediction=1", "-p", "SearchRange=16",
                "-p", "NumberReferenceFrames=2"]
    elif case in ("full_search_mode_rdopt0", "full_search_mode_satd"):
        # SearchMode = -1: FullPelBlockMotionSearch (me_fullsearch.c:47) for every block; without RDO the zero-vector bonus and the
        # early-exit wrap at picture origin are in play; with MEDistortionFPel = 2 the integer-pel metric is the Hadamard SAD
        args = ["-d", "encoder_baseline.cfg", "-p", "InputFile=syn.yuv", "-p", "SourceWidth=176", "-p", "SourceHeight=144",
                "-p", "FramesToBeEncoded=3", "-p", "FrameSkip=0", "-p", "SearchMode=-1", "-p", "SearchRange=8",
                "-p", "NumberReferenceFrames=2"]
        args += ["-p", "RDOptimization=0"] if case.endswith("rdopt0") else ["-p", "MEDistortionFPel=2"]
    elif case in ("umhex_plain_subpel", "umhex_simple_bipred_subpel", "epzs_plain_subpel", "epzs_subpel_grid_plain_subpel"):
        # SURVEY.md 8(f) row 4, first half: the pattern searches keep their integer-pel stage (and their own sub-pel patterns) on the
        # host; the plain sub-pel routines their dispatch still calls (mv-search.c:799, :823, :1018) and the plane sets are the
        # library's.  With EPZSSubPelGrid the integer stage hands over vectors that are not multiples of four.
        args = ["-d", "encoder.cfg", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2"]
        args += {"umhex_plain_subpel": ["-p", "SearchMode=1"],
                 "umhex_simple_bipred_subpel": ["-p", "SearchMode=2"],
                 "epzs_plain_subpel": ["-p", "SearchMode=3", "-p", "EPZSSubPelME=0", "-p", "EPZSSubPelMEBiPred=0", "-p", "EPZSSubPelGrid=0"],
                 "epzs_subpel_grid_plain_subpel": ["-p", "SearchMode=3", "-p", "EPZSSubPelME=0", "-p", "EPZSSubPelMEBiPred=0",
                                                   "-p", "EPZSSubPelGrid=1"]}[case]
    elif case.startswith("interlace_"):
        # field pictures (PicInterlace 1: the two fields of every frame coded as pictures of half the rows with
        # field references), picture-adaptive frame / field coding (2), MBAFF frames (MbInterlace 1: every pair as field macroblocks,
        # 2: adaptive) whose field macroblocks search the top / bottom field of the frame with list offsets 2 / 4.  The reference
        # refuses ReferenceReorder / PocMemoryManagement with interlace (configfile.c:1032-1043); 144 rows are padded to 160.
        args = ["-d", "encoder.cfg", "-p", "ReferenceReorder=0", "-p", "PocMemoryManagement=0", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2"]
        args += {"interlace_field_pictures": ["-p", "PicInterlace=1"],
                 "interlace_picture_adaptive": ["-p", "PicInterlace=2"],
                 "interlace_mbaff_field_macroblocks": ["-p", "MbInterlace=1"],
                 "interlace_mbaff_adaptive": ["-p", "MbInterlace=2"],
                 "interlace_picture_and_mbaff_adaptive_chroma_me": ["-p", "PicInterlace=2", "-p", "MbInterlace=2", "-p", "ChromaMEEnable=2"],
                 "interlace_field_pictures_epzs": ["-p", "PicInterlace=1", "-p", "SearchMode=3"]}[case]
    elif case.startswith("high_bit_depth_"):
        # SURVEY.md 8(f) row 4, last third: more than 8 bits per sample (imgpel u16): 16-bit plane sets, block searches, plain and
        # bi-predictive refinements and the launch per PartitionMotionSearch on 16-bit samples.  Input: the reference's QCIF clip
        # shifted up with random low bits (hbd.yuv, 16-bit little-endian samples).
        bits = 12 if "_12_" in case else 10
        args = ["-d", "encoder.cfg", "-p", "InputFile=hbd.yuv", "-p", "BitDepthLuma=%d" % bits, "-p", "BitDepthChroma=%d" % bits,
                "-p", "ProfileIDC=%d" % (122 if bits == 12 else 110), "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2"]
        if "sad_sse_rdopt0" in case:
            args += ["-p", "MEDistortionHPel=0", "-p", "MEDistortionQPel=1", "-p", "RDOptimization=0", "-p", "Transform8x8Mode=0"]
        if "field_pictures" in case:
            args += ["-p", "PicInterlace=1", "-p", "ReferenceReorder=0", "-p", "PocMemoryManagement=0"]
        if case.endswith("_epzs"):
            args += ["-p", "SearchMode=3"]
        if case.endswith("_umhex"):
            args += ["-p", "SearchMode=1"]
        if "weighted_pred_fade" in case:                   # explicit weights, searches on weighted 16-bit reference samples
            args[1] = "encoder_main.cfg"
            args += ["-p", "SourceWidth=176", "-p", "SourceHeight=144", "-p", "FramesToBeEncoded=4", "-p", "FrameSkip=0", "-p", "NumberBFrames=0",
                     "-p", "NumberReferenceFrames=3", "-p", "WeightedPrediction=1", "-p", "UseWeightedReferenceME=1"]
        if "weighted_bipred_fade" in case:                 # weighted bi-predictive searches (computeBiPred*2), clip at 1023
            args += ["-p", "FramesToBeEncoded=3", "-p", "WeightedBiprediction=1", "-p", "WeightedPrediction=1"]
        if "full_search_mode" in case:
            args[1] = "encoder_baseline.cfg"               # no bi-predictive ME: its range must not exceed SearchRange
            args += ["-p", "SearchMode=-1", "-p", "SearchRange=8", "-p", "RDOptimization=0"]
    elif case == "epzs_pattern_subpel":
        # SURVEY.md 8(f) row 4, second half: EPZS as shipped (EPZSSubPelME = EPZSSubPelMEBiPred = EPZSSubPelGrid = 1):
        # EPZSSubPelBlockMotionSearch / EPZSSubPelBlockSearchBiPred (me_epzs.c:2390, :2728) on the device
        args = ["-d", "encoder.cfg", "-p", "SearchMode=3", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2"]
    elif case == "epzs_pattern_subpel_weighted":
        # ... on weighted reference samples (explicit weights, fading clip) and with weighted bi-prediction in the B pictures
        args = ["-d", "encoder.cfg", "-p", "InputFile=syn.yuv", "-p", "SourceWidth=176", "-p", "SourceHeight=144", "-p", "FramesToBeEncoded=3",
                "-p", "WeightedBiprediction=1", "-p", "WeightedPrediction=1", "-p", "UseWeightedReferenceME=1", "-p", "SearchRange=16",
                "-p", "NumberReferenceFrames=2", "-p", "SearchMode=3"]
    elif case == "umhex_pattern_subpel_chroma_sad":
        # UMHEXSubPelBlockMotionSearch (me_umhex.c:562) with SAD at the sub-pel levels and ChromaMEEnable = 2: chroma terms in the walk
        args = ["-d", "encoder.cfg", "-p", "SearchMode=1", "-p", "ChromaMEEnable=2", "-p", "MEDistortionHPel=0", "-p", "MEDistortionQPel=0",
                "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2"]
    elif case == "umhex_simple_pattern_subpel_rdopt0":
        # smpUMHEXSubPelBlockMotionSearch / smpUMHEXFullSubPelBlockMotionSearch (me_umhexsmp.c:616, :422) without RDO (check_position0)
        args = ["-d", "encoder_baseline.cfg", "-p", "SearchMode=2", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2", "-p", "RDOptimization=0"]
    elif case == "sr64_2ref":
        # the search range of BASELINE config 5 (4K), on a QCIF clip so that the CPU side finishes in seconds
        args = ["-d", "encoder_main.cfg", "-p", "InputFile=syn.yuv", "-p", "SourceWidth=176", "-p", "SourceHeight=144",
                "-p", "FramesToBeEncoded=3", "-p", "FrameSkip=0", "-p", "NumberBFrames=0", "-p", "SearchRange=64",
                "-p", "NumberReferenceFrames=2", "-p", "LevelIDC=51"]
    else:
        args = ["-d", "encoder_baseline.cfg", "-p", "InputFile=syn.yuv", "-p", "SourceWidth=176", "-p", "SourceHeight=144",
                "-p", "FramesToBeEncoded=3", "-p", "FrameSkip=0", "-p", "RDOptimization=0", "-p", "RestrictSearchRange=0",
                "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2"]
    for d in (a, b):
        _prep(str(d))
        if "InputFile=hbd.yuv" in args:
            import numpy as np
            if "fade" in case:
                synth.write_yuv420(str(d / "fade8.yuv"), 176, 144, 7, kind="fade")
            src = np.fromfile(str(d / "fade8.yuv") if "fade" in case else os.path.join(REF_DIR, "fixtures", "foreman_part_qcif.yuv"), np.uint8).astype(np.uint16)
            sh = bits - 8
            ((src << sh) + np.random.default_rng(5).integers(0, 1 << sh, src.shape).astype(np.uint16)).astype("<u2").tofile(str(d / "hbd.yuv"))
        if "InputFile=syn.yuv" in args:
            wdt = int([a for a in args if a.startswith("SourceWidth=")][0].split("=")[1])
            hgt = int([a for a in args if a.startswith("SourceHeight=")][0].split("=")[1])
            synth.write_yuv420(str(d / "syn.yuv"), wdt, hgt, 2 if wdt == 1920 else 8 if case == "cif_5ref_8frames" else 7,
                               kind="fade" if ("fade" in case or "weighted" in case) else "pan")
    _, err = _run("lencod_cuda", args, str(a))
    out_cpu, _ = _run("lencod", args, str(b))
    assert _md5(str(a / "test.264")) == _md5(str(b / "test.264"))
    assert _md5(str(a / "test_rec.yuv")) == _md5(str(b / "test_rec.yuv"))
    # the device did the work: plane sets, and the searches of the mode under test
    pattern = "plain_subpel" in case or "bipred_subpel" in case or "pattern_subpel" in case or case.endswith("_epzs") or case.endswith("_umhex")
    c = _served(err, ("plane sets", "bi_sub") if "bipred_subpel" in case else ("plane sets",) if "pattern_subpel" in case or case.endswith("_epzs") or case.endswith("_umhex") else ("plane sets", "subpel") if pattern else ("plane sets", "fullpel", "subpel") if "full_search_mode" in case or "full_search_rdopt0" in case
                else ("plane sets", "ffs", "subpel"))
    if case.startswith("interlace_"):
        import re
        m = re.search(r"interlace: (\d+) macroblock visits searched on a field view of an MBAFF frame, (\d+) on the fields of field pictures", err)
        assert m, err[-800:]
        n_mbaff, n_field_pic = map(int, m.groups())
        assert (n_mbaff > 100) == ("mbaff" in case) and (n_field_pic > 100) == ("picture" in case), (case, m.group(0))
    if not pattern and "full_search" not in case:
        # fast full search: one launch per PartitionMotionSearch call, and the device's own motion vector predictors agreed with the
        # reference's SetMotionVectorPredictor on every block search they were used for
        assert c["pms_launches"] > 0 and c["pms_served"] > 0 and c["pms_missed"] == 0, c
    if case == "cif_5ref_8frames":
        # all five references were searched: the reference's own report lists the fifth reference in use for the last pictures
        assert c["ffs"] >= 41 * 396 * (1 + 2 + 3 + 4 + 5 + 5 + 5), c
    if pattern:
        # the modes' own sub-pel stages ran on the device (shim report: EPZS, EPZS bi-pred, UMHEX, smpUMHEX, smpUMHEX full)
        import re
        m = re.search(r"sub-pel stages of the pattern searches served on the device: (\d+) EPZS, (\d+) EPZS bi-pred, (\d+) UMHEX, (\d+) smpUMHEX, (\d+) smpUMHEX full", err)
        assert m, err[-800:]
        n_epzs, n_epzs_bi, n_umhex, n_smp, n_smp_full = map(int, m.groups())
        mode = int([a for a in args if a.startswith("SearchMode=")][0].split("=")[1])
        if mode == 1:
            assert n_umhex > 1000 and c["subpel"] > 0, m.group(0)
        elif mode == 2:
            assert n_smp > 1000 and n_smp_full > 100, m.group(0)
        elif "EPZSSubPelME=0" not in args:
            assert n_epzs > 1000 and (n_epzs_bi > 0 or "baseline" in args[1]), m.group(0)
    if pattern or case.startswith("interlace_") or case.startswith("high_bit_depth_"):
        # the device did serve these modes: verify mode counts the SubPelBlockMotionSearch calls it answered (and compares each)
        import re
        r = subprocess.run([os.path.join(REF_DIR, "lencod_cuda")] + args, cwd=str(a), capture_output=True, text=True, timeout=1500,
                           env=dict(os.environ, ME_CUDA_LIB=os.path.join(ROOT, "h.264_b200", "libme_cuda.so"), ME_CUDA_VERIFY="1"))
        m = re.search(r"verify mode: (\d+) calls checked, (\d+) mismatches", r.stderr)
        assert r.returncode == 0 and m, r.stderr[-1500:]
        assert int(m.group(1)) > 0 and int(m.group(2)) == 0, m.group(0)


@pytest.mark.skipif(not _have(), reason="oracle/_ref/lencod_cuda not built")
@pytest.mark.parametrize("cfg", ["encoder_baseline.cfg", "encoder_main.cfg", "encoder_extended.cfg", "encoder_yuv422.cfg"])
def test_shipped_configurations_as_they_are(tmp_path, me_lib_path, cfg):
    """the reference's own cfg files, unchanged, on its own QCIF fixtures (Baseline / Main / Extended with B pictures / High 4:2:2):
    SR 32, 5 references"""
    if cfg == "encoder_yuv422.cfg" and not os.path.exists(os.path.join(REF_DIR, "fixtures", "foreman_part_qcif_422.yuv")):
        pytest.skip("4:2:2 fixture not in oracle/_ref/fixtures (rebuild with make -C oracle ref)")
    a, b = tmp_path / "gpu", tmp_path / "cpu"
    a.mkdir(); b.mkdir()
    for d in (a, b):
        _prep(str(d))
    _, err = _run("lencod_cuda", ["-d", cfg], str(a))
    _run("lencod", ["-d", cfg], str(b))
    assert _md5(str(a / "test.264")) == _md5(str(b / "test.264"))
    assert _md5(str(a / "test_rec.yuv")) == _md5(str(b / "test_rec.yuv"))
    c = _served(err, ("plane sets", "ffs", "subpel"))
    assert c["pms_launches"] > 0 and c["pms_missed"] == 0, c


@pytest.mark.skipif(not _have(), reason="oracle/_ref/lencod_cuda not built")
@pytest.mark.parametrize("case", ["hd720_444_ibp_3ref", "uhd4k_sr64_ip"])
def test_stated_size_configs_match_committed_md5(tmp_path, me_lib_path, case):
    """BASELINE configs 4 and 5 at their stated picture sizes (1280x720 4:4:4 I-B-P with 3 references; 3840x2160 SR 64 I + P):
    lencod_cuda against the md5 of the plain reference encoder's bitstream and reconstruction, committed by
    tests/golden/make_golden.py (`big`) -- the CPU encoder needs minutes for these and is not re-run on the GPU box."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    md5 = json.load(open(os.path.join(GOLDEN, "bitstream_md5.json")))[case]
    _prep(str(tmp_path))
    mg.write_big_input(case, str(tmp_path / "syn.yuv"))
    _, err = _run("lencod_cuda", mg.BIG_CASES[case]["args"], str(tmp_path))
    assert _md5(str(tmp_path / "test.264")) == md5["test.264"]
    assert _md5(str(tmp_path / "test_rec.yuv")) == md5["test_rec.yuv"]
    c = _served(err, ("plane sets", "ffs", "subpel"))
    assert c["pms_launches"] > 0, c


@pytest.mark.skipif(not _have(), reason="oracle/_ref binaries not built")
@pytest.mark.parametrize("case", ["kat", "baseline16"])
def test_live_full_trace_replays_through_the_c_abi(tmp_path, me_lib_path, case):
    """every call of a fresh run of the traced reference encoder (all macroblocks), replayed through the drop-in ABI"""
    from replay import CudaAdapter, replay
    from trace_util import Trace, cfg_from_trace
    _prep(str(tmp_path))
    args = ["-d", "encoder.cfg"] if case == "kat" else ["-d", "encoder_baseline.cfg", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=1"]
    _run("lencod_trace", args, str(tmp_path), env={"ME_TRACE_DIR": str(tmp_path), "ME_TRACE_PLANES": "0"})
    tr = Trace(str(tmp_path))
    cfgkw = cfg_from_trace(tr, search_range=32 if case == "kat" else 16, full_search=2)
    impl = CudaAdapter(tr, cfgkw)
    try:
        res = replay(tr, impl)
    finally:
        impl.close()
    for name, (n, bad) in res.items():
        assert not bad, (case, name, n, len(bad), bad[:4])


@pytest.mark.skipif(not _have(), reason="oracle/_ref/lencod_cuda not built")
def test_gop_sharded_encoder_matches_reference_per_gop(tmp_path, me_lib_path):
    """one lencod_cuda process per GOP (StartFrame ranges), GPU chosen with ME_CUDA_DEVICE, per-GOP bitstreams gathered
    on the host -- every GOP byte-identical to the plain reference encoder on the same range, and the gathered streams equal."""
    import torch
    import h264_b200_pkg  # noqa: F401
    from h264_b200 import gop_shard as gs, synth
    yuv = str(tmp_path / "clip.yuv")
    w, h, gops, frames = 176, 144, 4, 3
    synth.write_yuv420(yuv, w, h, gops * frames)
    extra = ["-p", "SearchRange=16", "-p", "NumberReferenceFrames=2", "-p", "SearchMode=0"]
    n_gpus = max(1, min(2, torch.cuda.device_count()))
    gpu = gs.run_encoder_gops(yuv, w, h, gops, frames, n_gpus, extra=extra, workdir=str(tmp_path / "g"))
    cpu = gs.run_encoder_gops(yuv, w, h, gops, frames, n_gpus, extra=extra, exe="lencod", workdir=str(tmp_path / "c"))
    for g, c in zip(gpu["gops"], cpu["gops"]):
        assert g["md5"] == c["md5"] and g["rec_md5"] == c["rec_md5"], g["gop"]
        assert any("served on the device" in ln for ln in g["served"]), g
        assert g["device"] == g["gop"] % n_gpus
    assert _md5(gpu["stream"]) == _md5(cpu["stream"])
