"""Replay of reference call traces through an implementation under test.

An adapter exposes the reference's call sequence (mb_begin / ffs_setup / FastFullPelBlockMotionSearch / ...).
OracleAdapter drives oracle/liboracle_me.so (CPU restatement); CudaAdapter drives libme_cuda.so through its C ABI.
replay() walks a Trace in call order and returns the mismatching records per function.
"""
import numpy as np

from trace_util import REC_FFS, REC_SUBPEL, REC_BI_FULL, REC_BI_SUB, REC_CURPIC, REC_REFPIC, REC_FULLPEL, REC_SUBPEL_PAT, REC_BI_SUB_PAT


def picture_lists(trace, pic):
    """reference lists of coded picture `pic` as {list: [ref picture id by index]} recovered from the records"""
    lists = {0: {}, 1: {}}
    recs = trace.recs[(trace.recs[:, 1] == pic)]
    for r in recs:
        t = int(r[0])
        if t in (REC_FFS, REC_SUBPEL, REC_FULLPEL, REC_SUBPEL_PAT):
            lists[int(r[5])][int(r[6])] = int(r[21])
        elif t in (REC_BI_FULL, REC_BI_SUB, REC_BI_SUB_PAT):
            lists[int(r[5])][int(r[6])] = int(r[21])
            lists[int(r[5]) ^ 1][0] = int(r[22])
    out = {}
    for l in (0, 1):
        if lists[l]:
            n = max(lists[l]) + 1
            # indices never used by a traced call get any valid picture
            anyid = next(iter(lists[l].values()))
            out[l] = [lists[l].get(i, anyid) for i in range(n)]
        else:
            out[l] = []
    return out


class OracleAdapter:
    def __init__(self, trace, cfgkw):
        from oracle import binding as ob
        self.ob = ob
        ob.set_bit_depth(trace.bits)          # more than 8 bits per sample: the 16-bit build of the oracle (reset in close())
        self.cfg = ob.make_cfg(**cfgkw)
        self.refs = {i: ob.Ref(trace.ref(i)) for i in trace.ref_ids()}
        self.surf = {}
        self.wp = {}
        self.ci = trace.chroma_info()
        self.crefs = []
        if self.ci:
            for i, r in self.refs.items():
                u, v = trace.chroma("ref", i)
                if self.ci["yuv_format"] == 3:            # 4:4:4: chroma sub-images = quarter-pel plane sets of U and V
                    ru, rv = ob.Ref(u), ob.Ref(v)
                    self.crefs += [ru, rv]
                    r.set_chroma444(ru, rv)
                else:
                    r.set_chroma(u, v)
        self.trace = trace

    def set_chroma_state(self, pic, enable):
        """the global ChromaMEEnable the next call sees (0 / 1 / 2: any non-zero value switches the chroma terms on)"""
        if not self.ci:
            return
        u, v = self.trace.chroma("cur", pic)
        shift = {1: (3, 3), 2: (3, 2), 3: (2, 2)}[self.ci["yuv_format"]]
        self.ob.set_chroma(enable, self.ci["weight"], u, v, shift)

    def set_weights(self, lst, ref, apply, weight, offset, denom):
        self.wp[(lst, ref)] = (apply, weight, offset, denom)

    def _ref(self, lst, ref):
        """the picture at (list, ref) carrying that entry's weights (uni-predictive searches)"""
        r = self.refs[self.lists[lst][ref]]
        r.set_weights(*self.wp.get((lst, ref), (0, 32, 0, 5)))
        return r

    def begin_picture(self, cur, slice_type, lists, pic=None):
        self.cur, self.slice_type, self.lists = np.ascontiguousarray(cur), slice_type, lists
        self.wp = {}

    def mb_begin(self, ox, oy):
        self.ox, self.oy = ox, oy
        self.surf = {}

    def ffs_setup(self, lst, ref, pmvx, pmvy):
        ob = self.ob
        cx, cy = ob.search_centre(self.cfg, ref, self.ox, self.oy, pmvx, pmvy)
        rng = ob.lib().ora_setup_range(self.cfg, ref)
        s = ob.ffs_surfaces(self.cfg, self.cur, self._ref(lst, ref), self.ox, self.oy, cx, cy, rng)
        self.surf[(lst, ref)] = (s, cx, cy, rng)
        return cx, cy, rng

    def ffs_block(self, ref, lst, ppx, ppy, bt, pmvx, pmvy, search_range, min_mcost, lam):
        s, cx, cy, rng = self.surf[(lst, ref)]
        p = self.ob.part_index(bt, (ppy - self.oy) + ((ppx - self.ox) >> 2))
        return self.ob.ffs_block(self.cfg, s[p], rng, self.ox, self.oy, cx, cy, pmvx, pmvy, search_range, min_mcost, lam)

    def fullpel(self, ref, lst, ppx, ppy, bt, pmvx, pmvy, mvx, mvy, search_range, min_mcost, lam):
        return self.ob.fullpel_block(self.cfg, self.cur, self._ref(lst, ref), ppx, ppy, bt, ref, self.slice_type, pmvx, pmvy,
                                     mvx, mvy, search_range, min_mcost, lam)

    def subpel(self, ref, lst, ppx, ppy, bt, pmvx, pmvy, mvx, mvy, sp2, sp4, min_mcost, lam3):
        return self.ob.subpel(self.cfg, self.cur, self._ref(lst, ref), ppx, ppy, bt, ref, self.slice_type,
                              pmvx, pmvy, mvx, mvy, min_mcost, lam3, sp2, sp4)

    def subpel_pattern(self, pattern, ref, lst, ppx, ppy, bt, pmv, mv, sp2, sp4, min_mcost, lam3, subthres, uplayer, thresholds):
        return self.ob.subpel_pattern(self.cfg, self.cur, self._ref(lst, ref), pattern, ppx, ppy, bt, ref, self.slice_type, pmv, mv,
                                      min_mcost, lam3, sp2, sp4, subthres, uplayer, thresholds)

    def bipred_sub_pattern(self, ref, lst, ppx, ppy, bt, pmv1, pmv2, mv, smv, sp2, sp4, min_mcost, lam3):
        r1, r2 = self.refs[self.lists[lst][ref]], self.refs[self.lists[lst ^ 1][0]]
        self.ob.select_bipred_list(lst)
        return self.ob.epzs_subpel_bipred(self.cfg, self.cur, r1, r2, ppx, ppy, bt, pmv1, pmv2, mv, smv, min_mcost, lam3, sp2, sp4)

    def set_bipred_weights(self, lst, apply, w1, w2, obi, denom):
        self.ob.set_bipred_weights(lst, apply, w1, w2, obi, denom)

    def set_bipred_chroma_weights(self, lst, k, w1, w2, obi):
        self.ob.set_bipred_chroma_weights(lst, k, w1, w2, obi)

    def bipred_full(self, ref, lst, ppx, ppy, bt, pmv1, pmv2, mv, smv, sr, min_mcost, lam):
        r1, r2 = self.refs[self.lists[lst][ref]], self.refs[self.lists[lst ^ 1][0]]
        self.ob.select_bipred_list(lst)
        return self.ob.bipred_fullpel(self.cfg, self.cur, r1, r2, ppx, ppy, bt, pmv1, pmv2, mv, smv, sr, min_mcost, lam)

    def bipred_sub(self, ref, lst, ppx, ppy, bt, pmv, mv, smv, sp2, sp4, min_mcost, lam3):
        r1, r2 = self.refs[self.lists[lst][ref]], self.refs[self.lists[lst ^ 1][0]]
        self.ob.select_bipred_list(lst)
        return self.ob.bipred_subpel(self.cfg, self.cur, r1, r2, ppx, ppy, bt, pmv, mv, smv, min_mcost, lam3, sp2, sp4)

    def close(self):
        for r in list(self.refs.values()) + self.crefs:
            r.close()
        self.ob.set_bit_depth(8)


class CudaAdapter:
    """the drop-in call sequence of include/me_cuda.h"""

    def __init__(self, trace, cfgkw, max_refs=5):
        import h264_b200_pkg  # noqa: F401
        from h264_b200 import MeCuda
        h, w = trace.size[("cur", trace.cur_ids()[0])]
        self.ci = trace.chroma_info()
        self.trace = trace
        ckw = dict(chroma_me=self.ci["enable"], chroma_me_weight=self.ci["weight"], yuv_format=self.ci["yuv_format"]) if self.ci else {}
        self.me = MeCuda(w, h, search_range=cfgkw["search_range"], max_num_references=max_refs, rdopt=cfgkw["rdopt"],
                         full_search=cfgkw["full_search"], me_metric=cfgkw["metric"], transform8x8=cfgkw["transform8x8"],
                         mv_limits=(cfgkw["mv_limit_min_y"], cfgkw["mv_limit_max_y"]), bit_depth=trace.bits, **ckw)
        self.hbd = trace.bits > 8
        self.handles = {i: (self.me.ref_create_imgpel(trace.ref(i))[0] if self.hbd else self.me.ref_create(trace.ref(i))) for i in trace.ref_ids()}
        if self.ci:
            for i, hd in list(self.handles.items()):
                u, v = trace.chroma("ref", i)
                if self.ci["yuv_format"] == 3:
                    self.me.ref_set_chroma_planes(hd, self.me.ref_create(u), self.me.ref_create(v))
                else:
                    self.me.ref_set_chroma(hd, u, v)

    def begin_picture(self, cur, slice_type, lists, pic=None):
        if self.hbd:
            self.me.set_current_imgpel(cur)
        else:
            self.me.set_current(cur)
        if self.ci:
            self.me.set_current_chroma(*self.trace.chroma("cur", pic))
        self.me.set_slice_type(slice_type)
        for l in (0, 1):
            self.me.set_ref_list(l, [self.handles[i] for i in lists[l]])

    def set_weights(self, lst, ref, apply, weight, offset, denom):
        self.me.set_ref_weights(lst, ref, apply, weight, offset, denom)

    def set_chroma_state(self, pic, enable):
        if self.ci:
            self.me.set_chroma_state(enable)

    def set_bipred_weights(self, lst, apply, w1, w2, obi, denom):
        self.me.set_bipred_weights(lst, apply, w1, w2, obi, denom)

    def set_bipred_chroma_weights(self, lst, k, w1, w2, obi):
        self.me.set_bipred_chroma_weights(lst, k, w1, w2, obi)

    def mb_begin(self, ox, oy):
        self.me.mb_begin(ox, oy)

    def ffs_setup(self, lst, ref, pmvx, pmvy):
        self.me.ffs_setup(lst, ref, pmvx, pmvy)
        return None

    def ffs_block(self, ref, lst, ppx, ppy, bt, pmvx, pmvy, search_range, min_mcost, lam):
        return self.me.FastFullPelBlockMotionSearch(ref, lst, ppx, ppy, bt, pmvx, pmvy, search_range, min_mcost, lam)

    def fullpel(self, ref, lst, ppx, ppy, bt, pmvx, pmvy, mvx, mvy, search_range, min_mcost, lam):
        return self.me.FullPelBlockMotionSearch(ref, lst, ppx, ppy, bt, pmvx, pmvy, mvx, mvy, search_range, min_mcost, lam)

    def subpel(self, ref, lst, ppx, ppy, bt, pmvx, pmvy, mvx, mvy, sp2, sp4, min_mcost, lam3):
        return self.me.SubPelBlockMotionSearch(ref, lst, ppx, ppy, bt, pmvx, pmvy, mvx, mvy, sp2, sp4, min_mcost, lam3)

    def bipred_full(self, ref, lst, ppx, ppy, bt, pmv1, pmv2, mv, smv, sr, min_mcost, lam):
        return self.me.FullPelBlockMotionBiPred(ref, lst, ppx, ppy, bt, pmv1, pmv2, mv, smv, sr, min_mcost, lam)

    def bipred_sub(self, ref, lst, ppx, ppy, bt, pmv, mv, smv, sp2, sp4, min_mcost, lam3):
        return self.me.SubPelBlockSearchBiPred(ref, lst, ppx, ppy, bt, pmv, mv, smv, sp2, sp4, min_mcost, lam3)

    def subpel_pattern(self, pattern, ref, lst, ppx, ppy, bt, pmv, mv, sp2, sp4, min_mcost, lam3, subthres, uplayer, thresholds):
        return self.me.SubPelPattern(pattern, ref, lst, ppx, ppy, bt, pmv, mv, sp2, sp4, min_mcost, lam3, subthres, uplayer, thresholds)

    def bipred_sub_pattern(self, ref, lst, ppx, ppy, bt, pmv1, pmv2, mv, smv, sp2, sp4, min_mcost, lam3):
        return self.me.EPZSSubPelBlockSearchBiPred(ref, lst, ppx, ppy, bt, pmv1, pmv2, mv, smv, sp2, sp4, min_mcost, lam3)

    def close(self):
        self.me.close()


def replay(trace, impl, max_calls=None):
    """returns {name: (n_calls, [mismatch descriptions])}"""
    res = {"ffs": [0, []], "centre": [0, []], "subpel": [0, []], "bi_full": [0, []], "bi_sub": [0, []], "fullpel": [0, []],
           "subpel_pat": [0, []], "bi_sub_pat": [0, []]}
    cur_pic, cur_mb = None, None
    bi_chroma = {}
    slice_of = {int(r[1]): int(r[2]) for r in trace.of_type(REC_CURPIC)}
    n = 0
    for r in trace.recs:
        t = int(r[0])
        if t in (REC_REFPIC, REC_CURPIC):
            continue
        pic, ox, oy = int(r[1]), int(r[3]), int(r[4])
        if pic != cur_pic:
            impl.begin_picture(trace.cur(pic), slice_of.get(pic, int(r[2]) & 255), picture_lists(trace, pic), pic)
            cur_pic, cur_mb = pic, None
        if (ox, oy) != cur_mb:
            impl.mb_begin(ox, oy)
            cur_mb = (ox, oy)
        lst, ref, bt, ppx, ppy = int(r[5]), int(r[6]), int(r[7]), int(r[8]), int(r[9])
        if t in (REC_FFS, REC_SUBPEL, REC_FULLPEL, REC_SUBPEL_PAT):
            # weighted-prediction ME: the weights the reference derived for this (list, ref) (0 = off in older fixtures)
            impl.set_weights(lst, ref, int(r[28]), int(r[29]), int(r[30]), int(r[31]))
        # ChromaMEEnable as the call sees it (the reference resets the global between the stages of BlockMotionSearch)
        impl.set_chroma_state(pic, int(r[27]) if t == REC_FFS else int(r[26]) if t in (REC_FULLPEL, REC_BI_FULL) else int(r[2]) >> 8)
        if t in (REC_BI_FULL, REC_BI_SUB, REC_BI_SUB_PAT):
            # weighted bi-prediction: 1 + log denominator (0 = off), weight1 | weight2 << 16, offsetBi
            w = int(r[30])
            impl.set_bipred_weights(lst, int(r[29]) > 0, (w << 16 >> 16) if w & 0x8000 else w & 0xffff, w >> 16, int(r[31]), max(0, int(r[29]) - 1))
            # chroma weights: recorded with the integer-pel search (8 bits each), valid for the sub-pel calls of the same list
            if t == REC_BI_FULL and int(r[29]) > 0 and (int(r[27]) >> 24) & 1:
                s8 = lambda v: (v & 0xff) - 256 if v & 0x80 else v & 0xff   # noqa: E731
                bi_chroma[(pic, lst)] = [(s8(int(r[27 + k])), s8(int(r[27 + k]) >> 8), s8(int(r[27 + k]) >> 16)) for k in (0, 1)]
            if int(r[29]) > 0 and (pic, lst) in bi_chroma:
                for k, (c1, c2, ob_) in enumerate(bi_chroma[(pic, lst)]):
                    impl.set_bipred_chroma_weights(lst, k, c1, c2, ob_)
        if t == REC_FFS:
            if r[22]:
                c = impl.ffs_setup(lst, ref, int(r[23]), int(r[24]))
                if c is not None:
                    res["centre"][0] += 1
                    if c != (int(r[15]), int(r[16]), int(r[17])):
                        res["centre"][1].append((r[:26].tolist(), c))
            got = impl.ffs_block(ref, lst, ppx, ppy, bt, int(r[10]), int(r[11]), int(r[12]), int(r[14]), int(r[13]))
            exp = (int(r[18]), int(r[19]), int(r[20]))
            key = "ffs"
        elif t == REC_FULLPEL:
            got = impl.fullpel(ref, lst, ppx, ppy, bt, int(r[10]), int(r[11]), int(r[15]), int(r[16]), int(r[12]), int(r[14]), int(r[13]))
            exp = (int(r[18]), int(r[19]), int(r[20]))
            key = "fullpel"
        elif t == REC_SUBPEL:
            got = impl.subpel(ref, lst, ppx, ppy, bt, int(r[10]), int(r[11]), int(r[12]), int(r[13]), int(r[26]), int(r[27]),
                              int(r[14]), [int(r[15]), int(r[16]), int(r[17])])
            exp = (int(r[18]), int(r[19]), int(r[20]))
            key = "subpel"
        elif t == REC_SUBPEL_PAT:
            s16 = lambda v: ((v & 0xffff) ^ 0x8000) - 0x8000   # noqa: E731
            got = impl.subpel_pattern(int(r[22]) >> 4, ref, lst, ppx, ppy, bt, (int(r[10]), int(r[11])), (int(r[12]), int(r[13])),
                                      int(r[26]), int(r[27]), int(r[14]), [int(r[15]), int(r[16]), int(r[17])], int(r[24]),
                                      (s16(int(r[25])), int(r[25]) >> 16), (int(r[23]) & 0xffff, int(r[23]) >> 16))
            exp = (int(r[18]), int(r[19]), int(r[20]))
            key = "subpel_pat"
        elif t == REC_BI_SUB_PAT:
            s16 = lambda v: ((v & 0xffff) ^ 0x8000) - 0x8000   # noqa: E731
            got = impl.bipred_sub_pattern(ref, lst, ppx, ppy, bt, (int(r[10]), int(r[11])), (s16(int(r[20])), int(r[20]) >> 16),
                                          (int(r[12]), int(r[13])), (int(r[14]), int(r[15])), 9, 9, int(r[16]),
                                          [int(r[17]), int(r[18]), int(r[19])])
            exp = (int(r[23]), int(r[24]), int(r[25]))
            key = "bi_sub_pat"
        elif t == REC_BI_FULL:
            got = impl.bipred_full(ref, lst, ppx, ppy, bt, (int(r[10]), int(r[11])), (int(r[12]), int(r[13])),
                                   (int(r[17]), int(r[18])), (int(r[19]), int(r[20])), int(r[16]), int(r[14]), int(r[15]))
            exp = (int(r[23]), int(r[24]), int(r[25]))
            key = "bi_full"
        else:
            got = impl.bipred_sub(ref, lst, ppx, ppy, bt, (int(r[10]), int(r[11])), (int(r[12]), int(r[13])),
                                  (int(r[14]), int(r[15])), 9, 9, int(r[16]),
                                  [int(r[17]), int(r[18]), int(r[19])])
            exp = (int(r[23]), int(r[24]), int(r[25]))
            key = "bi_sub"
        res[key][0] += 1
        if tuple(got) != exp:
            res[key][1].append((r[:29].tolist(), tuple(got)))
        n += 1
        if max_calls and n >= max_calls:
            break
    return {k: (v[0], v[1]) for k, v in res.items()}


def pattern_counts(case, res):
    """what each pattern fixture must exercise (tests/golden/make_golden.py: PATTERN_CASES)"""
    assert res["subpel_pat"][0] > 1500 and not res["ffs"][0]
    if case.startswith("epzs"):
        assert res["bi_sub_pat"][0] > 20
    if case == "umhex_sub":
        assert res["subpel"][0] > 300 and res["bi_sub"][0] > 20
