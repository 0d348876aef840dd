#!/usr/bin/env python
"""Side measurement: latency of the drop-in calls (one launch + one host wake-up each) on a CIF picture:
me_cuda_ffs_setup, me_cuda_FastFullPelBlockMotionSearch (with the speculated sub-pel refinement), me_cuda_SubPelBlockMotionSearch
answered from the speculation and answered by its own launch."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import h264_b200_pkg  # noqa: E402,F401
from h264_b200 import MeCuda, synth  # noqa: E402

W, H = 352, 288
frames = synth.luma_frames(W, H, 3)
LAM = [334459] * 3
with MeCuda(W, H, search_range=32, max_num_references=2) as me:
    me.set_ref_list(0, [me.ref_create(frames[1]), me.ref_create(frames[0])])
    me.set_current(frames[2])
    me.mb_begin(160, 128)
    n = 2000
    t0 = time.perf_counter()
    for _ in range(n):
        me.mb_begin(160, 128)
        me.ffs_setup(0, 0, 12, 8)
    t_setup = (time.perf_counter() - t0) / n
    mx, my, c = me.FastFullPelBlockMotionSearch(0, 0, 160, 128, 1, 12, 8, 32, 2147483647, LAM[0])
    me.SubPelBlockMotionSearch(0, 0, 160, 128, 1, 12, 8, mx * 4, my * 4, 9, 9, 2147483647, LAM)     # shows the library the sub-pel parameters
    t0 = time.perf_counter()
    for _ in range(n):
        mx, my, c = me.FastFullPelBlockMotionSearch(0, 0, 160, 128, 1, 12, 8, 32, 2147483647, LAM[0])
    t_ffs = (time.perf_counter() - t0) / n
    t0 = time.perf_counter()
    for _ in range(n):
        mx, my, c = me.FastFullPelBlockMotionSearch(0, 0, 160, 128, 1, 12, 8, 32, 2147483647, LAM[0])
        me.SubPelBlockMotionSearch(0, 0, 160, 128, 1, 12, 8, mx * 4, my * 4, 9, 9, 2147483647, LAM)
    t_pair = (time.perf_counter() - t0) / n
    t0 = time.perf_counter()
    for _ in range(n):
        me.SubPelBlockMotionSearch(0, 0, 160, 128, 7, 12, 8, mx * 4, my * 4, 9, 9, 2147483647, LAM)    # not predicted: own launch
    t_sub = (time.perf_counter() - t0) / n
print("setup (surfaces kernel, asynchronous) %.1f us; block search + speculated sub-pel %.1f us; pair with the sub-pel call served "
      "from the speculation %.1f us; sub-pel call with its own launch %.1f us" % (t_setup * 1e6, t_ffs * 1e6, t_pair * 1e6, t_sub * 1e6))
