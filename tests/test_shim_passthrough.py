"""CPU: the reference-side binding (h.264_b200/shim/jm_shim.c linked into oracle/_ref/lencod_cuda with --wrap) must be inert when it is
switched off: with ME_CUDA_DISABLE=1 every wrapper forwards to the reference function, no library is loaded, and the encoder writes
the reference's own bitstream -- for the known-answer run (the reference's committed bin/test.264) and for the configurations whose
wrappers were added in round 2 (pattern searches, interlace, more than 8 bits per sample).  Also: without a CUDA device and without
the switch the binding fails loudly instead of falling back."""
import hashlib
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import REF_DIR, ROOT, has_ref_binaries

pytestmark = pytest.mark.skipif(not (has_ref_binaries() and os.path.exists(os.path.join(REF_DIR, "lencod_cuda"))), reason="oracle/_ref not built")


def _md5(p):
    return hashlib.md5(open(p, "rb").read()).hexdigest()


def _prep(d):
    fx = os.path.join(REF_DIR, "fixtures")
    for f in os.listdir(fx):
        if f.endswith(".cfg") or f.endswith(".yuv"):
            shutil.copy(os.path.join(fx, f), d)


def _run(exe, args, cwd, **env):
    e = dict(os.environ, ME_CUDA_LIB=os.path.join(ROOT, "h.264_b200", "libme_cuda.so"), **env)
    return subprocess.run([os.path.join(REF_DIR, exe)] + args, cwd=cwd, env=e, capture_output=True, text=True, timeout=600)


def test_disabled_binding_reproduces_the_known_answer_bitstream(tmp_path):
    _prep(str(tmp_path))
    r = _run("lencod_cuda", ["-d", "encoder.cfg"], str(tmp_path), ME_CUDA_DISABLE="1")
    assert r.returncode == 0, r.stderr[-800:]
    assert "pass-through (ME_CUDA_DISABLE): nothing served on the device" in r.stderr
    assert _md5(str(tmp_path / "test.264")) == _md5(os.path.join(REF_DIR, "fixtures", "golden_test.264")) == "92402ba469c35bc09f97017a3903bdf9"
    assert _md5(str(tmp_path / "test_rec.yuv")) == _md5(os.path.join(REF_DIR, "fixtures", "golden_test_rec.yuv"))


@pytest.mark.parametrize("name,args", [
    ("epzs", ["-p", "SearchMode=3"]),
    ("umhex", ["-p", "SearchMode=1"]),
    ("umhex_simple", ["-p", "SearchMode=2"]),
    ("mbaff_and_field_pictures", ["-p", "PicInterlace=2", "-p", "MbInterlace=2", "-p", "ReferenceReorder=0", "-p", "PocMemoryManagement=0"]),
    ("high_bit_depth", ["-p", "InputFile=hbd.yuv", "-p", "BitDepthLuma=10", "-p", "BitDepthChroma=10", "-p", "ProfileIDC=110"])])
def test_disabled_binding_equals_the_plain_encoder(tmp_path, name, args):
    a, b = tmp_path / "shim", tmp_path / "plain"
    args = ["-d", "encoder.cfg", "-p", "SearchRange=16", "-p", "NumberReferenceFrames=2"] + args
    for d in (a, b):
        d.mkdir()
        _prep(str(d))
        if name == "high_bit_depth":
            src = np.fromfile(os.path.join(REF_DIR, "fixtures", "foreman_part_qcif.yuv"), np.uint8).astype(np.uint16)
            ((src << 2) + np.random.default_rng(5).integers(0, 4, src.shape).astype(np.uint16)).astype("<u2").tofile(str(d / "hbd.yuv"))
    r = _run("lencod_cuda", args, str(a), ME_CUDA_DISABLE="1")
    assert r.returncode == 0, r.stderr[-800:]
    p = _run("lencod", args, str(b))
    assert p.returncode == 0, p.stderr[-800:]
    assert _md5(str(a / "test.264")) == _md5(str(b / "test.264")) and _md5(str(a / "test_rec.yuv")) == _md5(str(b / "test_rec.yuv"))


def test_binding_without_a_device_fails_loudly(tmp_path):
    """no CPU fallback: on a machine without a GPU the active binding stops with the library's message"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _prep(str(tmp_path))
    e = {k: v for k, v in os.environ.items() if k != "ME_CUDA_DISABLE"}
    e["ME_CUDA_LIB"] = os.path.join(ROOT, "h.264_b200", "libme_cuda.so")
    r = subprocess.run([os.path.join(REF_DIR, "lencod_cuda"), "-d", "encoder.cfg"], cwd=str(tmp_path), env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode != 0
    assert "no CUDA device" in r.stderr or "CUDA" in r.stderr, r.stderr[-500:]
    assert not os.path.exists(tmp_path / "test.264") or os.path.getsize(tmp_path / "test.264") < 1000
