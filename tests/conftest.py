import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def has_ref_binaries():
    return os.path.exists(os.path.join(REF_DIR, "lencod_trace")) and os.path.exists(os.path.join(REF_DIR, "fixtures", "encoder.cfg"))


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import binding
    binding.build()
    return binding


@pytest.fixture(scope="session")
def me_lib_path():
    """libme_cuda.so, built in-tree (nvcc cross-compiles without a GPU)"""
    import h264_b200_pkg  # noqa: F401
    from h264_b200 import build
    return build.build()


# case name -> encoder-wide parameters of the traced run (tests/golden/make_golden.py)
GOLDEN_CASES = {
    "kat": dict(search_range=32, full_search=2, metric=(0, 2, 2), mv_limits=(-511, 511)),
    "baseline16": dict(search_range=16, full_search=2, metric=(0, 2, 2), mv_limits=(-511, 511)),
    "rdopt0": dict(search_range=16, full_search=0, metric=(0, 2, 2), mv_limits=(-511, 511)),
    "wp": dict(search_range=16, full_search=2, metric=(0, 2, 2), mv_limits=(-511, 511)),     # weighted-prediction ME
    "fullsearch": dict(search_range=8, full_search=2, metric=(0, 2, 2), mv_limits=(-511, 511)),   # SearchMode = -1
    "chroma": dict(search_range=16, full_search=2, metric=(0, 2, 2), mv_limits=(-511, 511)),       # ChromaMEEnable = 2
    "chroma_fs": dict(search_range=16, full_search=2, metric=(0, 2, 2), mv_limits=(-511, 511)),    # ChromaMEEnable = 1, SearchMode = -1
    "wpbc": dict(search_range=16, full_search=2, metric=(0, 2, 2), mv_limits=(-511, 511)),      # implicit weighted bi-prediction + chroma
    "wpb": dict(search_range=16, full_search=2, metric=(0, 2, 2), mv_limits=(-511, 511)),    # + weighted bi-prediction
    "epzs": dict(search_range=16, full_search=2, metric=(0, 2, 2), mv_limits=(-511, 511)),   # plain sub-pel calls under SearchMode = 3
}


# sub-pel stages of the pattern searches (SearchMode 1, 2, 3; records of type 8 / 9): EPZSSubPelBlockMotionSearch / ...BiPred,
# UMHEXSubPelBlockMotionSearch, smpUMHEX[Full]SubPelBlockMotionSearch
PATTERN_CASES = {
    "epzs_sub": dict(search_range=16, full_search=2, metric=(0, 2, 2), mv_limits=(-511, 511)),
    "umhex_sub": dict(search_range=16, full_search=2, metric=(0, 2, 2), mv_limits=(-511, 511)),
    "smpumhex_sub": dict(search_range=16, full_search=2, metric=(0, 2, 2), mv_limits=(-511, 511)),        # RDOptimization = 0
    "epzs_chroma_sad": dict(search_range=16, full_search=2, metric=(0, 0, 0), mv_limits=(-511, 511)),     # ChromaMEEnable = 2, SAD sub-pel
}


# more than 8 bits per sample: 16-bit pictures and plane sets (High 10, BitDepthLuma = 10)
HBD_CASES = {
    "hbd10": dict(search_range=16, full_search=2, metric=(0, 2, 2), mv_limits=(-511, 511)),
}


def golden_trace(name):
    from trace_util import Trace, cfg_from_trace
    tr = Trace.load_npz(os.path.join(GOLDEN, name + ".npz"))
    return tr, cfg_from_trace(tr, **(GOLDEN_CASES.get(name) or PATTERN_CASES.get(name) or HBD_CASES[name]))
