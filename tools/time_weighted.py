#!/usr/bin/env python
"""Side measurement (not part of the bench contract): weighted-prediction ME at the headline size.  One 1080p P picture, 5
references with explicit weights (UseWeightedReferenceME): cost of building the weighted copies of the reference planes
(me_weight_planes_kernel, once per picture and weight set) and the search on them, next to the unweighted search; and the
bi-predictive batch with and without weighted bi-prediction."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import h264_b200_pkg  # noqa: E402,F401
from h264_b200 import MeCuda, synth, BIPRED_JOB_DTYPE  # noqa: E402

W, H, NREF = 1920, 1088, 5
frames = synth.luma_frames(W, H, NREF + 1, kind="fade")
jobs = synth.synthetic_jobs(W, H, NREF, (334459, 334459, 334459), seed=21)


def timed(fn, n=10):
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e3


with MeCuda(W, H, max_num_references=NREF) as me:
    handles = [me.ref_create(frames[NREF - 1 - r]) for r in range(NREF)]
    me.set_ref_list(0, handles)
    me.set_current(frames[NREF])
    plain = timed(lambda: me.frame_search(jobs))
    for r in range(NREF):
        me.set_ref_weights(0, r, 1, 31 - r, r - 2, 5)
    me.set_timing(True)
    t0 = time.perf_counter()
    me.frame_search(jobs)                      # builds the five weighted copies, then searches them
    first = (time.perf_counter() - t0) * 1e3
    lut_ms, lut_n = me.kernel_time_ms(0)
    me.set_timing(False)
    steady = timed(lambda: me.frame_search(jobs))
    print("1080p, 5 refs, 40800 jobs (host buffers in/out): unweighted %.2f ms; weighted, first call %.2f ms "
          "(%d table passes over planes + search plane: %.3f ms), later calls %.2f ms" % (plain, first, lut_n, lut_ms, steady))

frames = synth.luma_frames(W, H, 3, kind="fade")
rng = np.random.default_rng(5)
mbs = [(16 * x, 16 * y) for y in range(H // 16) for x in range(W // 16)]
bj = np.zeros(2 * len(mbs), BIPRED_JOB_DTYPE)
for k, (x, y) in enumerate(mbs):
    for lst in (0, 1):
        j = bj[2 * k + lst]
        sign = 1 if lst == 0 else -1
        j["opix_x"], j["opix_y"], j["list"] = x, y, lst
        j["pmv"] = (sign * 12 + rng.integers(-6, 7), sign * 8 + rng.integers(-6, 7))
        j["pmv_bi"] = (-sign * 12 + rng.integers(-6, 7), -sign * 8 + rng.integers(-6, 7))
        j["mv"] = (sign * 12 + rng.integers(-9, 10), sign * 8 + rng.integers(-9, 10))
        j["lambda"] = (334459, 334459, 334459)
with MeCuda(W, H, max_num_references=1) as me:
    me.set_ref_list(0, [me.ref_create(frames[0])])
    me.set_ref_list(1, [me.ref_create(frames[2])])
    me.set_current(frames[1])
    a = timed(lambda: me.frame_bipred(bj, 3, 16, 2))
    me.set_bipred_weights(0, 1, 31, 33, 0, 5)
    me.set_bipred_weights(1, 1, 33, 31, 0, 5)
    b = timed(lambda: me.frame_bipred(bj, 3, 16, 2))
    print("bi-predictive batch, %d jobs: %.2f ms unweighted, %.2f ms with weighted bi-prediction" % (len(bj), a, b))
