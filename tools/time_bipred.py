#!/usr/bin/env python
"""Side measurement (not part of the bench contract): me_cuda_frame_bipred on one 1080p B picture, every macroblock, both lists
(16320 jobs), BiPredMERefinements 3, BiPredMESearchRange 16, BiPredMESubPel 2 -- the settings of the reference's cfg files."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import h264_b200_pkg  # noqa: E402,F401
from h264_b200 import MeCuda, synth, BIPRED_JOB_DTYPE  # noqa: E402

W, H = 1920, 1088
frames = synth.luma_frames(W, H, 3)
rng = np.random.default_rng(5)
mbs = [(16 * x, 16 * y) for y in range(H // 16) for x in range(W // 16)]
jobs = np.zeros(2 * len(mbs), BIPRED_JOB_DTYPE)
for k, (x, y) in enumerate(mbs):
    for lst in (0, 1):
        j = jobs[2 * k + lst]
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
    for _ in range(3):
        me.frame_bipred(jobs, 3, 16, 2)
    t0 = time.perf_counter()
    n = 10
    for _ in range(n):
        out = me.frame_bipred(jobs, 3, 16, 2)
    dt = (time.perf_counter() - t0) / n
print("bi-pred batch, %d jobs: %.2f ms per picture (host buffers in/out), %.0f macroblocks/s" % (len(jobs), dt * 1e3, len(mbs) / dt))
