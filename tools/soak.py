#!/usr/bin/env python
"""Randomised parity soak (not part of the test suite): frame batches on random clips, sizes, ranges, predictors, lambda factors
and weights against the oracle on a sample of the jobs.  usage: soak.py [rounds] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import h264_b200_pkg  # noqa: E402,F401
from h264_b200 import MeCuda, synth  # noqa: E402
from oracle import binding as ob  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 12345)
ob.build()
bad = 0
for it in range(rounds):
    w, h = [(176, 144), (352, 288), (320, 240), (640, 368)][rng.integers(4)]
    sr = int([16, 32, 64, 24][rng.integers(4)])
    n_refs = int(rng.integers(1, 4))
    t8 = int(rng.integers(2))
    rdopt = int(rng.integers(4) > 0)
    metric = [(0, 2, 2), (0, 2, 2), (1, 2, 2), (0, 0, 0), (0, 1, 2)][rng.integers(5)]
    lam = int([334459, 65536, 1200000, 20000000, 3][rng.integers(5)])
    kind = ["pan", "fade", "stripes", "pan"][rng.integers(4)]
    d = dict(search_range=sr, rdopt=rdopt, full_search=2, metric=metric, transform8x8=t8, mv_limit_min_y=-511, mv_limit_max_y=511)
    cfg = ob.make_cfg(**d)
    frames = synth.luma_frames(w, h, n_refs + 1, seed=int(rng.integers(1 << 30)), kind=kind)
    if kind == "stripes":
        frames = [np.roll(f, int(rng.integers(8)), axis=1) for f in frames]
    cur = np.ascontiguousarray(frames[-1])
    jobs = synth.synthetic_jobs(w, h, n_refs, (lam, lam, lam), seed=int(rng.integers(1 << 30)), max_pmv=int(rng.integers(4, 4 * sr + 200)), rdopt=rdopt)
    weights = [(int(rng.integers(2)), int(rng.integers(-20, 90)), int(rng.integers(-40, 40)), int(rng.integers(0, 7))) for _ in range(n_refs)]
    with MeCuda(w, h, search_range=sr, max_num_references=n_refs, rdopt=rdopt, me_metric=metric, transform8x8=t8) as me:
        me.set_ref_list(0, [me.ref_create(frames[-2 - r]) for r in range(n_refs)])
        me.set_current(cur)
        for r, wt in enumerate(weights):
            me.set_ref_weights(0, r, *wt)
        out, out_int = me.frame_search(jobs, want_int=True)
    sel = np.sort(rng.choice(len(jobs), min(len(jobs), 60), replace=False))
    ok = True
    for r in range(n_refs):
        s_ = sel[jobs["ref"][sel] == r]
        if not len(s_):
            continue
        ref = ob.Ref(frames[-2 - r])
        ref.set_weights(*weights[r])
        exp, exp_int = ob.mb_search(cfg, cur, ref, jobs[s_])
        ok = ok and np.array_equal(out_int[s_], exp_int) and np.array_equal(out[s_], exp)
        ref.close()
    print("round %2d %s: %dx%d sr %d refs %d t8 %d rdopt %d metric %s lambda %d %s weights %s" %
          (it, "ok " if ok else "MISMATCH", w, h, sr, n_refs, t8, rdopt, metric, lam, kind, weights), flush=True)
    bad += not ok
# ---- bi-predictive batch: random predictors, vectors, refinements, ranges, sub-pel levels, weights
from h264_b200 import BIPRED_JOB_DTYPE  # noqa: E402
for it in range(max(1, rounds // 3)):
    w, h = [(176, 144), (352, 288)][rng.integers(2)]
    t8 = int(rng.integers(2))
    metric = [(0, 2, 2), (1, 2, 2), (2, 2, 2), (0, 0, 1)][rng.integers(4)]
    refinements, srange, subpel = int(rng.integers(0, 4)), int([4, 8, 16, 32][rng.integers(4)]), int(rng.integers(0, 3))
    d = dict(search_range=32, rdopt=1, full_search=2, metric=metric, transform8x8=t8, mv_limit_min_y=-511, mv_limit_max_y=511)
    cfg = ob.make_cfg(**d)
    frames = synth.luma_frames(w, h, 3, seed=int(rng.integers(1 << 30)), kind=["pan", "fade"][rng.integers(2)])
    cur, past, future = np.ascontiguousarray(frames[1]), frames[0], frames[2]
    mbs = [(16 * x, 16 * y) for y in range(h // 16) for x in range(w // 16)]
    pick = [mbs[i] for i in rng.choice(len(mbs), 24, replace=False)]
    jobs = np.zeros(2 * len(pick), BIPRED_JOB_DTYPE)
    for k, (x, y) in enumerate(pick):
        for lst in (0, 1):
            j = jobs[2 * k + lst]
            j["opix_x"], j["opix_y"], j["list"] = x, y, lst
            j["pmv"] = rng.integers(-60, 61, 2); j["pmv_bi"] = rng.integers(-60, 61, 2); j["mv"] = rng.integers(-80, 81, 2)
            j["lambda"] = int([334459, 65536, 3000000][rng.integers(3)])
    bw = [(int(rng.integers(2)), int(rng.integers(-10, 70)), int(rng.integers(-10, 70)), int(rng.integers(-20, 20)), int(rng.integers(3, 7))) for _ in (0, 1)]
    with MeCuda(w, h, search_range=32, max_num_references=1, me_metric=metric, transform8x8=t8) as me:
        me.set_ref_list(0, [me.ref_create(past)])
        me.set_ref_list(1, [me.ref_create(future)])
        me.set_current(cur)
        for lst in (0, 1):
            me.set_bipred_weights(lst, *bw[lst])
            ob.set_bipred_weights(lst, *bw[lst])
        out = me.frame_bipred(jobs, refinements, srange, subpel)
    r0, r1 = ob.Ref(past), ob.Ref(future)
    exp = ob.mb_bipred(cfg, cur, r0, r1, jobs.view(ob.BIPRED_JOB_DTYPE), refinements, srange, subpel)
    r0.close(); r1.close()
    for lst in (0, 1):
        ob.set_bipred_weights(lst, 0)
    ok = np.array_equal(out.view(ob.BIPRED_RESULT_DTYPE), exp)
    print("bipred %2d %s: %dx%d t8 %d metric %s refinements %d range %d subpel %d weights %s" %
          (it, "ok " if ok else "MISMATCH", w, h, t8, metric, refinements, srange, subpel, bw), flush=True)
    bad += not ok
print("soak: %d rounds, %d mismatching" % (rounds, bad))
sys.exit(1 if bad else 0)
