"""CPU: pieces of bench.py's contract that do not need a GPU -- the config object both arms print, the host-transfer policy, and
the reference arm's command line (the unmodified encoder on the stated configuration)."""
import importlib.util
import os

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_config_names_the_headline_workload(bench):
    c = bench.bench_config(1)
    assert c["workload"] == "1080p_sr32_5ref_allpartitions_qpel" and (c["width"], c["height"], c["search_range"], c["refs"]) == (1920, 1088, 32, 5)
    assert c["jobs_per_step"] == 8160 * 5 and c["partitions"] == 41
    assert "model" not in c and "l2" in c                        # a workload description, not a model; says how L2 is kept cold
    assert bench.METRIC.startswith("1080p FS-ME+QPel macroblocks/s") and bench.UNIT == "MB/s"
    # the algorithmic work the roofline is computed from: 212 lane-operations per position
    assert bench.P_POS == 65 * 65 and bench.SAD_OPS_PER_JOB == 64 * 65 * 65


def test_both_arms_print_the_same_config(bench):
    for n in (1, 2, 4, 8):
        assert bench.bench_config(n) == bench.bench_config(n)
        assert bench.bench_config(n)["parallelism"] == "gop-sharded x%d, no collective" % n


def test_host_transfer_policy(bench, monkeypatch):
    monkeypatch.delenv("ME_CUDA_ZEROCOPY", raising=False)
    assert [bench.host_transfer_mode(n) for n in (1, 2, 4, 8)] == ["1", "1", "1", "0"]     # measured: profiles/r02p_e2e_variants.txt
    assert "in place" in bench.bench_config(4)["host_transfer"] and "chunks" in bench.bench_config(8)["host_transfer"]
    monkeypatch.setenv("ME_CUDA_ZEROCOPY", "0")
    assert bench.host_transfer_mode(1) == "0"                     # an explicit choice wins


def test_reference_arm_runs_the_stated_configuration(bench):
    args = bench.reference_cfg_args("clip.yuv", 8)
    s = " ".join(args)
    for want in ("SearchMode=0", "SearchRange=32", "NumberReferenceFrames=5", "SourceWidth=1920", "SourceHeight=1088", "FramesToBeEncoded=8"):
        assert want in s, (want, s)
