"""CPU: libme_cuda.so builds (nvcc cross-compiles sm_100a without a GPU), loads, exports every symbol the header declares,
and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "me_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(me_cuda_[A-Za-z0-9_]+)\s*\(", text)) - {"me_cuda_error_fn"})


def test_header_and_binding_agree():
    import h264_b200_pkg  # noqa: F401
    from h264_b200 import me_cuda
    assert header_symbols() == sorted(me_cuda.ABI_SYMBOLS)
    text = open(os.path.join(ROOT, "include", "me_cuda.h")).read()
    assert int(re.search(r"#define ME_CUDA_ABI_VERSION (\d+)", text).group(1)) == me_cuda.ABI_VERSION
    assert me_cuda.MOTION_DTYPE.itemsize == 344                    # me_mb_motion


def test_library_exports_every_declared_symbol(me_lib_path):
    lib = C.CDLL(me_lib_path)
    for name in header_symbols():
        assert hasattr(lib, name), name


def test_library_is_sm100a_only(me_lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", me_lib_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_struct_sizes_match_abi():
    import h264_b200_pkg  # noqa: F401
    from h264_b200 import me_cuda
    from oracle import binding
    assert me_cuda.JOB_DTYPE.itemsize == 188 and me_cuda.RESULT_DTYPE.itemsize == 328
    assert me_cuda.JOB_DTYPE == binding.JOB_DTYPE and me_cuda.RESULT_DTYPE == binding.RESULT_DTYPE
    assert me_cuda.PART_TYPE == binding.PART_TYPE and me_cuda.PART_BLK == binding.PART_BLK


def test_no_cpu_fallback(me_lib_path):
    """Without a CUDA device init must fail loudly; with one this test is skipped (the gpu tests cover it)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import h264_b200_pkg  # noqa: F401
    from h264_b200 import MeCuda, MeCudaError
    with pytest.raises(MeCudaError) as e:
        MeCuda(176, 144)
    assert "no CPU path" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_touches_the_oracle():
    """the package and the library sources must not reference oracle/ (the oracle is a checker, not a fallback)"""
    pkg = os.path.join(ROOT, "h.264_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in text and "me_oracle" not in text and "oracle.binding" not in text, f
