#!/usr/bin/env python
"""Generates the committed golden fixtures from the UNMODIFIED reference encoder.

Runs oracle/_ref/lencod_trace (reference objects + the --wrap call tracer oracle/trace_wrap.c; built by
`make -C oracle ref` in a container that has /root/reference) and stores, per case, a compact .npz:
  recs            traced calls of the hot-path functions (arguments and results; layout in oracle/trace_wrap.c)
  cur_<n>, ref_<i> original / reconstructed luma pictures the calls operated on
plus planes_crc.json: CRC32 of each of the 16 quarter-pel planes the reference built for every reference picture.

Cases
  kat        bin/encoder.cfg as shipped (the run behind the committed bin/test.264): QCIF I-B-P, SR 32, 5 refs, RDO on,
             8x8 transform, bi-pred ME.  Every 4th macroblock is kept.
  baseline16 encoder_baseline.cfg with SearchRange=16, NumberReferenceFrames=1 (BASELINE.json config 1), every 3rd macroblock.
  rdopt0     encoder_baseline.cfg, RDOptimization=0, RestrictSearchRange=0, SR 16, 2 refs, synthetic QCIF IPPP --
             pins the !rdopt zero-vector rules and the reduced search ranges.  Every 5th macroblock is kept.
  wp         encoder_main.cfg, WeightedPrediction=1 + UseWeightedReferenceME=1 on a synthetic QCIF clip that fades
             (h264_b200.synth kind="fade"), IPPP, SR 16, 3 references: the encoder estimates weights 28..31 / 32 and the
             searches run on weighted reference samples (computeSADWP / SATDWP).  Every 4th macroblock is kept.
  fullsearch encoder_baseline.cfg with SearchMode=-1 (FullPelBlockMotionSearch, me_fullsearch.c:47, for every block), SR 8,
             2 references, RDOptimization=0 on the synthetic QCIF clip: includes the blocks at picture origin whose first
             candidate is accepted with a one-row SAD (early-exit threshold wrap, me_fullsearch.c:129-138).  Every 3rd MB.
  chroma     encoder.cfg with ChromaMEEnable=2, SR 16, 2 references (I-B-P, 8x8 transform): chroma terms in the fast full search
             and in the bi-predictive integer-pel search.  chroma_fs: encoder_main.cfg with ChromaMEEnable=1, SearchMode=-1,
             RDOptimization=0: chroma terms of computeSAD in the per-block full search.  Every 6th macroblock; the fixtures
             carry the original and reconstructed chroma planes.
  wpbc       encoder.cfg with two B pictures between references, WeightedBiprediction=2 (implicit weights 43 / 21),
             WeightedPrediction=1 and ChromaMEEnable=2 on the fading clip: weighted chroma samples in the bi-predictive
             integer-pel search, explicit luma weights + chroma terms in the P pictures.  Every 12th macroblock.
  wpb        encoder.cfg (High profile, I-B-P, bi-predictive ME, 8x8 transform) with WeightedBiprediction=1 and
             WeightedPrediction=1 on the fading clip, SR 16, 2 references: weighted bi-predictive searches
             (computeBiPredSAD2 / SATD2, including the 8x8 branch's source-pointer slip).  Every 6th macroblock is kept.
Usage: python tests/golden/make_golden.py            (rewrites tests/golden/*.npz, planes_crc.json, bitstream_md5.json)
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from trace_util import Trace, REC_FFS, REC_SUBPEL, REC_BI_FULL, REC_BI_SUB, REC_FULLPEL, REC_SUBPEL_PAT, REC_BI_SUB_PAT  # noqa: E402
import h264_b200_pkg  # noqa: E402,F401
from h264_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")


def run_trace(workdir, args, planes=True, binary="lencod_trace"):
    for f in os.listdir(os.path.join(REF, "fixtures")):
        if f.endswith(".cfg"):
            shutil.copy(os.path.join(REF, "fixtures", f), workdir)
    env = dict(os.environ, ME_TRACE_DIR=workdir, ME_TRACE_PLANES="1" if planes else "0")
    r = subprocess.run([os.path.join(REF, binary)] + args, cwd=workdir, env=env, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stdout[-2000:] + r.stderr[-2000:])
    return r.stdout


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def keep_mask(tr, step):
    recs = tr.recs
    mb_w = tr.size[("cur", tr.cur_ids()[0])][1] // 16
    mb = (recs[:, 4] // 16) * mb_w + recs[:, 3] // 16
    is_call = np.isin(recs[:, 0], [REC_FFS, REC_SUBPEL, REC_BI_FULL, REC_BI_SUB, REC_FULLPEL, REC_SUBPEL_PAT, REC_BI_SUB_PAT])
    return ~is_call | (mb % step == 0)


def planes_crc(tr, workdir):
    out = {}
    for i in tr.ref_ids():
        sub = tr.ref_sub(i)
        out[str(i)] = [[int(zlib.crc32(sub[fy, fx].tobytes())) for fx in range(4)] for fy in range(4)]
    return out


def main():
    crc, md5s = {}, {}
    # ---- kat
    with tempfile.TemporaryDirectory() as d:
        shutil.copy(os.path.join(REF, "fixtures", "foreman_part_qcif.yuv"), d)
        run_trace(d, ["-d", "encoder.cfg"])
        md5s["kat"] = {"test.264": md5(os.path.join(d, "test.264")), "test_rec.yuv": md5(os.path.join(d, "test_rec.yuv"))}
        assert md5s["kat"]["test.264"] == md5(os.path.join(REF, "fixtures", "golden_test.264"))
        tr = Trace(d)
        tr.save_npz(os.path.join(HERE, "kat.npz"), keep_mask(tr, 4))
        crc["kat"] = planes_crc(tr, d)
    # ---- rdopt0
    with tempfile.TemporaryDirectory() as d:
        synth.write_yuv420(os.path.join(d, "syn_qcif.yuv"), 176, 144, 4)
        args = ["-d", "encoder_baseline.cfg", "-p", "InputFile=syn_qcif.yuv", "-p", "FramesToBeEncoded=4",
                "-p", "RDOptimization=0", "-p", "RestrictSearchRange=0", "-p", "SearchRange=16",
                "-p", "NumberReferenceFrames=2", "-p", "FrameSkip=0"]
        run_trace(d, args)
        md5s["rdopt0"] = {"test.264": md5(os.path.join(d, "test.264"))}
        tr = Trace(d)
        tr.save_npz(os.path.join(HERE, "rdopt0.npz"), keep_mask(tr, 5))
        crc["rdopt0"] = planes_crc(tr, d)
    # ---- baseline16: BASELINE.json config 1 (encoder_baseline.cfg, SR 16, 1 reference, real QCIF clip, IPP)
    with tempfile.TemporaryDirectory() as d:
        shutil.copy(os.path.join(REF, "fixtures", "foreman_part_qcif.yuv"), d)
        run_trace(d, ["-d", "encoder_baseline.cfg", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=1"])
        md5s["baseline16"] = {"test.264": md5(os.path.join(d, "test.264")), "test_rec.yuv": md5(os.path.join(d, "test_rec.yuv"))}
        tr = Trace(d)
        tr.save_npz(os.path.join(HERE, "baseline16.npz"), keep_mask(tr, 3))
        crc["baseline16"] = planes_crc(tr, d)
    # ---- wp: weighted-prediction ME (records carry apply / weight / offset / denominator in fields 28..31)
    with tempfile.TemporaryDirectory() as d:
        synth.write_yuv420(os.path.join(d, "syn.yuv"), 176, 144, 4, kind="fade")
        args = ["-d", "encoder_main.cfg", "-p", "InputFile=syn.yuv", "-p", "SourceWidth=176", "-p", "SourceHeight=144",
                "-p", "FramesToBeEncoded=4", "-p", "FrameSkip=0", "-p", "NumberBFrames=0", "-p", "SearchRange=16",
                "-p", "NumberReferenceFrames=3", "-p", "WeightedPrediction=1", "-p", "UseWeightedReferenceME=1"]
        run_trace(d, args)
        md5s["wp"] = {"test.264": md5(os.path.join(d, "test.264"))}
        tr = Trace(d)
        assert set(np.unique(tr.of_type(REC_FFS)[:, 29])) - {32}, "the fade clip must produce non-default weights"
        tr.save_npz(os.path.join(HERE, "wp.npz"), keep_mask(tr, 4))
        crc["wp"] = planes_crc(tr, d)
    # ---- wpb: weighted bi-prediction (bi-predictive records carry the weights in fields 29..31), High profile, 8x8 transform
    with tempfile.TemporaryDirectory() as d:
        synth.write_yuv420(os.path.join(d, "syn.yuv"), 176, 144, 5, kind="fade")
        args = ["-d", "encoder.cfg", "-p", "InputFile=syn.yuv", "-p", "FramesToBeEncoded=3", "-p", "WeightedBiprediction=1",
                "-p", "WeightedPrediction=1", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2"]
        run_trace(d, args)
        md5s["wpb"] = {"test.264": md5(os.path.join(d, "test.264"))}
        tr = Trace(d)
        assert (tr.of_type(REC_BI_SUB)[:, 29] > 0).all() and (tr.of_type(REC_BI_SUB)[:, 26] == 1).all()
        tr.save_npz(os.path.join(HERE, "wpb.npz"), keep_mask(tr, 6))
        crc["wpb"] = planes_crc(tr, d)
    # ---- fullsearch: SearchMode = -1, FullPelBlockMotionSearch for every block (records of type 7)
    with tempfile.TemporaryDirectory() as d:
        synth.write_yuv420(os.path.join(d, "syn.yuv"), 176, 144, 3)
        args = ["-d", "encoder_baseline.cfg", "-p", "InputFile=syn.yuv", "-p", "SourceWidth=176", "-p", "SourceHeight=144",
                "-p", "FramesToBeEncoded=3", "-p", "FrameSkip=0", "-p", "SearchMode=-1", "-p", "SearchRange=8",
                "-p", "NumberReferenceFrames=2", "-p", "RDOptimization=0"]
        run_trace(d, args)
        md5s["fullsearch"] = {"test.264": md5(os.path.join(d, "test.264"))}
        tr = Trace(d)
        assert len(tr.of_type(REC_FULLPEL)) > 10000 and not len(tr.of_type(REC_FFS))
        tr.save_npz(os.path.join(HERE, "fullsearch.npz"), keep_mask(tr, 3))
        crc["fullsearch"] = planes_crc(tr, d)
    # ---- chroma: ChromaMEEnable = 2 (chroma terms in the fast full search and in the bi-predictive integer-pel search); the
    # fixtures also hold the original / reconstructed chroma planes
    with tempfile.TemporaryDirectory() as d:
        shutil.copy(os.path.join(REF, "fixtures", "foreman_part_qcif.yuv"), d)
        run_trace(d, ["-d", "encoder.cfg", "-p", "ChromaMEEnable=2", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2"])
        md5s["chroma"] = {"test.264": md5(os.path.join(d, "test.264"))}
        tr = Trace(d)
        assert tr.chroma_info()["enable"] == 2 and (tr.of_type(REC_BI_FULL)[:, 26] == 1).all()
        tr.save_npz(os.path.join(HERE, "chroma.npz"), keep_mask(tr, 6))
        crc["chroma"] = planes_crc(tr, d)
    # ---- chroma_fs: ChromaMEEnable = 1 with SearchMode = -1 (chroma terms of computeSAD in the per-block full search), no RDO
    with tempfile.TemporaryDirectory() as d:
        shutil.copy(os.path.join(REF, "fixtures", "foreman_part_qcif.yuv"), d)
        run_trace(d, ["-d", "encoder_main.cfg", "-p", "ChromaMEEnable=1", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2",
                      "-p", "SearchMode=-1", "-p", "RDOptimization=0"])
        md5s["chroma_fs"] = {"test.264": md5(os.path.join(d, "test.264"))}
        tr = Trace(d)
        assert tr.chroma_info()["enable"] == 1 and len(tr.of_type(REC_FULLPEL)) > 10000
        tr.save_npz(os.path.join(HERE, "chroma_fs.npz"), keep_mask(tr, 6))
        crc["chroma_fs"] = planes_crc(tr, d)
    # ---- wpbc: implicit weighted bi-prediction (two B pictures between references: weights 43 / 21) with ChromaMEEnable = 2 --
    # weighted chroma samples in the bi-predictive integer-pel search; explicit weights + chroma terms in the P pictures
    with tempfile.TemporaryDirectory() as d:
        synth.write_yuv420(os.path.join(d, "syn.yuv"), 176, 144, 7, kind="fade")
        run_trace(d, ["-d", "encoder.cfg", "-p", "InputFile=syn.yuv", "-p", "FramesToBeEncoded=3", "-p", "FrameSkip=2", "-p", "NumberBFrames=2",
                      "-p", "WeightedBiprediction=2", "-p", "WeightedPrediction=1", "-p", "ChromaMEEnable=2", "-p", "SearchRange=16",
                      "-p", "NumberReferenceFrames=2"])
        md5s["wpbc"] = {"test.264": md5(os.path.join(d, "test.264"))}
        tr = Trace(d)
        bf = tr.of_type(REC_BI_FULL)
        assert (bf[:, 29] > 0).all() and ((bf[:, 27] >> 24) & 1).all() and len(set(bf[:, 30].tolist())) > 1
        tr.save_npz(os.path.join(HERE, "wpbc.npz"), keep_mask(tr, 12))
        crc["wpbc"] = planes_crc(tr, d)
    epzs(crc, md5s)
    patterns(crc, md5s)
    hbd(crc, md5s)
    big(md5s)
    json.dump(crc, open(os.path.join(HERE, "planes_crc.json"), "w"), indent=1)
    json.dump(md5s, open(os.path.join(HERE, "bitstream_md5.json"), "w"), indent=1)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


# BASELINE.json configs 4 and 5 at their stated picture sizes: bitstream + reconstruction md5 of the plain reference encoder
# (oracle/_ref/lencod), which tests/test_jm_dropin.py::test_stated_size_configs_match_committed_md5 checks lencod_cuda against
# on the GPU box without re-running the CPU encoder there (minutes of CPU time).
BIG_CASES = {
    # config 4: synthetic 720p 4:4:4, High 4:4:4 profile, I-B-P with bi-predictive ME, SR 32, 3 references
    "hd720_444_ibp_3ref": dict(writer="444", w=1280, h=720, frames=3,
                               args=["-d", "encoder.cfg", "-p", "InputFile=syn.yuv", "-p", "SourceWidth=1280", "-p", "SourceHeight=720",
                                     "-p", "YUVFormat=3", "-p", "ProfileIDC=244", "-p", "LevelIDC=51", "-p", "NumberReferenceFrames=3",
                                     "-p", "SearchRange=32", "-p", "FramesToBeEncoded=2"]),
    # config 5: synthetic 4K (3840x2160, coded 3840x2176), SR 64, one I and one P picture of a GOP
    "uhd4k_sr64_ip": dict(writer="420", w=3840, h=2160, frames=2,
                          args=["-d", "encoder_main.cfg", "-p", "InputFile=syn.yuv", "-p", "SourceWidth=3840", "-p", "SourceHeight=2160",
                                "-p", "FramesToBeEncoded=2", "-p", "FrameSkip=0", "-p", "NumberBFrames=0", "-p", "SearchRange=64",
                                "-p", "NumberReferenceFrames=5", "-p", "LevelIDC=51"]),
}


def write_big_input(case, path):
    c = BIG_CASES[case]
    (synth.write_yuv444 if c["writer"] == "444" else synth.write_yuv420)(path, c["w"], c["h"], c["frames"])


def big(md5s, only=None):
    for case, c in BIG_CASES.items():
        if only and case not in only:
            continue
        with tempfile.TemporaryDirectory() as d:
            write_big_input(case, os.path.join(d, "syn.yuv"))
            run_trace(d, c["args"], planes=False, binary="lencod")
            md5s[case] = {"test.264": md5(os.path.join(d, "test.264")), "test_rec.yuv": md5(os.path.join(d, "test_rec.yuv"))}
            print(case, md5s[case], flush=True)


def epzs(crc, md5s):
    """epzs: encoder.cfg with SearchMode=3 (EPZS), EPZSSubPelGrid=1, EPZSSubPelME=0, EPZSSubPelMEBiPred=0 -- the integer-pel stage is
    the reference's own pattern search (not traced, not on the path); what reaches the traced functions are the plain sub-pel
    routines it still calls (mv-search.c:823, :1018), with start vectors that are not multiples of four.  Every 3rd macroblock."""
    with tempfile.TemporaryDirectory() as d:
        shutil.copy(os.path.join(REF, "fixtures", "foreman_part_qcif.yuv"), d)
        run_trace(d, ["-d", "encoder.cfg", "-p", "SearchMode=3", "-p", "EPZSSubPelME=0", "-p", "EPZSSubPelMEBiPred=0",
                      "-p", "EPZSSubPelGrid=1", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2"])
        md5s["epzs"] = {"test.264": md5(os.path.join(d, "test.264"))}
        tr = Trace(d)
        sp = tr.of_type(REC_SUBPEL)
        assert len(sp) > 1000 and not len(tr.of_type(REC_FFS)) and ((sp[:, 12] & 3) != 0).any() and len(tr.of_type(REC_BI_SUB))
        tr.save_npz(os.path.join(HERE, "epzs.npz"), keep_mask(tr, 3))
        crc["epzs"] = planes_crc(tr, d)


# Sub-pel stages of the pattern searches (records of type 8 / 9, oracle/trace_wrap.c)
PATTERN_CASES = {
    # EPZS as shipped in encoder.cfg (EPZSSubPelME = EPZSSubPelMEBiPred = EPZSSubPelGrid = 1): EPZSSubPelBlockMotionSearch and
    # EPZSSubPelBlockSearchBiPred, SATD, 8x8 transform, I-B-P, start vectors off the integer grid
    "epzs_sub": (["-d", "encoder.cfg", "-p", "SearchMode=3", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2"], 4),
    # UMHexagonS: UMHEXSubPelBlockMotionSearch for block types 4..7, the plain refinement for 1..3
    "umhex_sub": (["-d", "encoder.cfg", "-p", "SearchMode=1", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2"], 4),
    # simplified UMHexagonS without RDO: smpUMHEXSubPelBlockMotionSearch / smpUMHEXFullSubPelBlockMotionSearch with the
    # check_position0 bonus and the SubPelThreshold exits
    "smpumhex_sub": (["-d", "encoder_baseline.cfg", "-p", "SearchMode=2", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2",
                      "-p", "RDOptimization=0"], 3),
    # EPZS with SAD at the sub-pel levels and ChromaMEEnable = 2: chroma terms in the pattern's candidates (uni- and bi-predictive)
    "epzs_chroma_sad": (["-d", "encoder.cfg", "-p", "SearchMode=3", "-p", "ChromaMEEnable=2", "-p", "MEDistortionHPel=0",
                         "-p", "MEDistortionQPel=0", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2"], 8),
}


def patterns(crc, md5s, only=None):
    for case, (args, step) in PATTERN_CASES.items():
        if only and case not in only:
            continue
        with tempfile.TemporaryDirectory() as d:
            shutil.copy(os.path.join(REF, "fixtures", "foreman_part_qcif.yuv"), d)
            run_trace(d, args)
            md5s[case] = {"test.264": md5(os.path.join(d, "test.264"))}
            tr = Trace(d)
            pat = tr.of_type(REC_SUBPEL_PAT)
            assert len(pat) > 5000 and not len(tr.of_type(REC_FFS))
            if case.startswith("epzs"):
                assert len(tr.of_type(REC_BI_SUB_PAT)) and ((pat[:, 22] >> 4) == 1).all() and ((pat[:, 12] & 3) != 0).any()
            if case == "umhex_sub":
                assert ((pat[:, 22] >> 4) == 2).all() and len(tr.of_type(REC_SUBPEL))
            if case == "smpumhex_sub":
                assert set((pat[:, 22] >> 4).tolist()) == {3, 4} and not ((pat[:, 22] >> 3) & 1).any()
            tr.save_npz(os.path.join(HERE, case + ".npz"), keep_mask(tr, step))
            crc[case] = planes_crc(tr, d)


def write_hbd_clip(path, bits, seed=5):
    """the reference's QCIF clip shifted up to `bits` bits with random low bits, 16-bit little-endian samples"""
    src = np.fromfile(os.path.join(REF, "fixtures", "foreman_part_qcif.yuv"), np.uint8).astype(np.uint16)
    sh = bits - 8
    ((src << sh) + np.random.default_rng(seed).integers(0, 1 << sh, src.shape).astype(np.uint16)).astype("<u2").tofile(path)


def hbd(crc, md5s):
    """hbd10: encoder.cfg in High 10 with BitDepthLuma = BitDepthChroma = 10 (I-B-P, bi-predictive ME, 8x8 transform, SR 16, 2 references):
    16-bit pictures and plane sets (last third).  Every 4th macroblock is kept; pictures are stored as uint16."""
    with tempfile.TemporaryDirectory() as d:
        write_hbd_clip(os.path.join(d, "hbd.yuv"), 10)
        run_trace(d, ["-d", "encoder.cfg", "-p", "InputFile=hbd.yuv", "-p", "BitDepthLuma=10", "-p", "BitDepthChroma=10", "-p", "ProfileIDC=110",
                      "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2"])
        md5s["hbd10"] = {"test.264": md5(os.path.join(d, "test.264"))}
        tr = Trace(d)
        assert tr.bits == 10 and tr.ref(tr.ref_ids()[0]).dtype == np.uint16 and int(tr.ref(tr.ref_ids()[0]).max()) > 255
        assert len(tr.of_type(REC_FFS)) > 10000 and len(tr.of_type(REC_BI_FULL)) > 500
        tr.save_npz(os.path.join(HERE, "hbd10.npz"), keep_mask(tr, 4))
        crc["hbd10"] = planes_crc(tr, d)


if __name__ == "__main__":
    if sys.argv[1:2] == ["hbd"]:                      # add this case without rewriting the other fixtures
        crc = json.load(open(os.path.join(HERE, "planes_crc.json")))
        md5s = json.load(open(os.path.join(HERE, "bitstream_md5.json")))
        hbd(crc, md5s)
        json.dump(crc, open(os.path.join(HERE, "planes_crc.json"), "w"), indent=1)
        json.dump(md5s, open(os.path.join(HERE, "bitstream_md5.json"), "w"), indent=1)
    elif sys.argv[1:2] == ["patterns"]:               # add these cases without rewriting the other fixtures
        crc = json.load(open(os.path.join(HERE, "planes_crc.json")))
        md5s = json.load(open(os.path.join(HERE, "bitstream_md5.json")))
        patterns(crc, md5s, sys.argv[2:] or None)
        json.dump(crc, open(os.path.join(HERE, "planes_crc.json"), "w"), indent=1)
        json.dump(md5s, open(os.path.join(HERE, "bitstream_md5.json"), "w"), indent=1)
    elif sys.argv[1:] == ["epzs"]:                  # add this case without rewriting the other fixtures
        crc = json.load(open(os.path.join(HERE, "planes_crc.json")))
        md5s = json.load(open(os.path.join(HERE, "bitstream_md5.json")))
        epzs(crc, md5s)
        json.dump(crc, open(os.path.join(HERE, "planes_crc.json"), "w"), indent=1)
        json.dump(md5s, open(os.path.join(HERE, "bitstream_md5.json"), "w"), indent=1)
    elif sys.argv[1:2] == ["big"]:                # the stated-size cases only (minutes of CPU time)
        md5s = json.load(open(os.path.join(HERE, "bitstream_md5.json")))
        big(md5s, sys.argv[2:] or None)
        json.dump(md5s, open(os.path.join(HERE, "bitstream_md5.json"), "w"), indent=1)
    else:
        main()
