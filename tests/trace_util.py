"""Reader for the call traces written by oracle/trace_wrap.c (the --wrap tracer around the unmodified reference
encoder) and replay helpers that push every traced call through an implementation under test.

Record layouts (32 x int32) are defined in oracle/trace_wrap.c.
"""
import os

import numpy as np

REC_FFS, REC_SUBPEL, REC_BI_FULL, REC_BI_SUB, REC_REFPIC, REC_CURPIC, REC_FULLPEL = 1, 2, 3, 4, 5, 6, 7
REC_SUBPEL_PAT, REC_BI_SUB_PAT = 8, 9      # sub-pel stages of the pattern searches (SearchMode 1, 2, 3)


class Trace:
    def __init__(self, directory=None, recs=None, cur=None, ref=None, chroma=None):
        self.dir = directory
        if recs is None:
            recs = np.fromfile(os.path.join(directory, "trace.bin"), dtype="<i4").reshape(-1, 32)
        self.recs = recs
        self._cur = dict(cur or {})
        self._ref = dict(ref or {})
        self._chroma = dict(chroma or {})
        self.size = {}
        for r in recs[recs[:, 0] == REC_CURPIC]:
            self.size[("cur", int(r[1]))] = (int(r[4]), int(r[3]))
        for r in recs[recs[:, 0] == REC_REFPIC]:
            self.size[("ref", int(r[1]))] = (int(r[3]), int(r[2]))

    def of_type(self, t):
        return self.recs[self.recs[:, 0] == t]

    @property
    def bits(self):
        """bit depth of the traced run (REC_CURPIC field 12; 0 in fixtures made before the field existed = 8)"""
        r = self.of_type(REC_CURPIC)
        return int(r[0][12]) if len(r) and int(r[0][12]) else 8

    @property
    def dtype(self):
        return np.uint16 if self.bits > 8 else np.uint8

    def cur(self, n):
        if n not in self._cur:
            h, w = self.size[("cur", n)]
            self._cur[n] = np.fromfile(os.path.join(self.dir, "cur_%d.y" % n), self.dtype).reshape(h, w)
        return self._cur[n]

    def ref(self, i):
        if i not in self._ref:
            h, w = self.size[("ref", i)]
            self._ref[i] = np.fromfile(os.path.join(self.dir, "ref_%d.y" % i), self.dtype).reshape(h, w)
        return self._ref[i]

    # ChromaMEEnable runs: original / reconstructed chroma planes (REC_CURPIC fields 7..11: ChromaMEEnable, ChromaMEWeight,
    # yuv_format, chroma width, chroma height)
    def chroma_info(self):
        r = self.of_type(REC_CURPIC)
        if not len(r) or not int(r[0][7]):
            return None
        return dict(enable=int(r[0][7]), weight=int(r[0][8]), yuv_format=int(r[0][9]), cw=int(r[0][10]), ch=int(r[0][11]))

    def chroma(self, kind, n):
        """(U, V) of 'cur' picture n or 'ref' picture n"""
        key = (kind, n)
        if key not in self._chroma:
            ci = self.chroma_info()
            self._chroma[key] = tuple(np.fromfile(os.path.join(self.dir, "%s_%d.%s" % (kind, n, c)), np.uint8).reshape(ci["ch"], ci["cw"])
                                      for c in "uv")
        return self._chroma[key]

    def ref_sub(self, i):
        h, w = self.size[("ref", i)]
        return np.fromfile(os.path.join(self.dir, "ref_%d.sub" % i), self.dtype).reshape(4, 4, h + 40, w + 40)

    def cur_ids(self):
        return sorted(n for (k, n) in self.size if k == "cur")

    def ref_ids(self):
        return sorted(n for (k, n) in self.size if k == "ref")

    # ------------------------------------------------------------------ compact, committable form
    def save_npz(self, path, keep=None):
        recs = self.recs if keep is None else self.recs[keep]
        d = {"recs": recs.astype(np.int32)}
        for n in self.cur_ids():
            d["cur_%d" % n] = self.cur(n)
        for i in self.ref_ids():
            d["ref_%d" % i] = self.ref(i)
        if self.chroma_info():
            for kind, ids in (("cur", self.cur_ids()), ("ref", self.ref_ids())):
                for n in ids:
                    u, v = self.chroma(kind, n)
                    d["%su_%d" % (kind, n)] = u
                    d["%sv_%d" % (kind, n)] = v
        np.savez_compressed(path, **d)

    @staticmethod
    def load_npz(path):
        z = np.load(path)
        cur = {int(k[4:]): z[k] for k in z.files if k.startswith("cur_")}
        ref = {int(k[4:]): z[k] for k in z.files if k.startswith("ref_")}
        chroma = {}
        for k in z.files:
            if k.startswith(("curu_", "refu_")):
                kind, n = k[:3], int(k[5:])
                chroma[(kind, n)] = (z[k], z[kind + "v_%d" % n])
        t = Trace(None, z["recs"], cur, ref, chroma)
        for n, a in cur.items():
            t.size[("cur", n)] = a.shape
        for i, a in ref.items():
            t.size[("ref", i)] = a.shape
        return t


def cfg_from_trace(trace, search_range=32, full_search=2, metric=(0, 2, 2), mv_limits=(-511, 511)):
    """Encoder-wide parameters recoverable from the records themselves."""
    sub = trace.of_type(REC_SUBPEL)
    if not len(sub):
        sub = trace.of_type(REC_SUBPEL_PAT)[:, :23].copy()
        sub[:, 22] &= 1
    ffs = trace.of_type(REC_FFS)
    if not len(ffs):
        ffs = trace.of_type(REC_FULLPEL)
    pat = trace.of_type(REC_SUBPEL_PAT)
    rdopt = int(ffs[0][25]) if len(ffs) else (int(pat[0][22]) >> 3) & 1 if len(pat) else 1
    t8 = int(sub[:, 22].max()) if len(sub) else 0
    ci = trace.chroma_info()
    return dict(search_range=search_range, rdopt=rdopt, full_search=full_search, metric=metric,
                transform8x8=t8, mv_limit_min_y=mv_limits[0], mv_limit_max_y=mv_limits[1], chroma_me=ci["enable"] if ci else 0)


def bipred_chains(trace, refinements=3, subpel=2):
    """Groups the bi-predictive records of a trace into the per-(macroblock, list) chains BlockMotionSearch runs
    (mv-search.c:864-1034): 1 + refinements REC_BI_FULL records followed by `subpel` REC_BI_SUB records.
    Returns a list of dicts: job fields recovered from the first record, the records themselves, and the expected result."""
    recs = trace.recs[(trace.recs[:, 0] == REC_BI_FULL) | (trace.recs[:, 0] == REC_BI_SUB)]
    n = 1 + refinements + subpel
    chains = []
    i = 0
    while i + n <= len(recs):
        grp = recs[i:i + n]
        assert (grp[:1 + refinements, 0] == REC_BI_FULL).all() and (grp[1 + refinements:, 0] == REC_BI_SUB).all(), i
        assert len(set(map(tuple, grp[:, [1, 3, 4]]))) == 1, i         # same picture, same macroblock
        f0 = grp[0]
        last_sub = grp[-1]
        first_sub = grp[1 + refinements]
        chains.append(dict(
            pic=int(f0[1]), opix=(int(f0[3]), int(f0[4])), list=int(f0[5]),
            pmv=(int(f0[10]), int(f0[11])), pmv_bi=(int(f0[12]), int(f0[13])),
            mv_int=(int(f0[19]), int(f0[20])),                       # the fixed vector of pass 0 = the uni-predictive result
            lam=(int(last_sub[17]), int(last_sub[18]), int(last_sub[19])), search_range=int(f0[16]),
            ref_l=(int(f0[21]), int(f0[22])),                        # picture ids: reference 0 of `list`, of the other list
            exp_mv=(int(last_sub[23]), int(last_sub[24])) if subpel == 2 else (int(first_sub[14]), int(first_sub[15])),
            exp_bimv=(int(first_sub[23]), int(first_sub[24])), exp_cost=int(last_sub[25]), recs=grp))
        i += n
    return chains
