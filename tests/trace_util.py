"""Reader for the call traces written by oracle/trace_wrap.c (the --wrap tracer around the unmodified reference
encoder) and replay helpers that push every traced call through an implementation under test.

Record layouts (32 x int32) are defined in oracle/trace_wrap.c.
"""
import os

import numpy as np

REC_FFS, REC_SUBPEL, REC_BI_FULL, REC_BI_SUB, REC_REFPIC, REC_CURPIC = 1, 2, 3, 4, 5, 6


class Trace:
    def __init__(self, directory=None, recs=None, cur=None, ref=None):
        self.dir = directory
        if recs is None:
            recs = np.fromfile(os.path.join(directory, "trace.bin"), dtype="<i4").reshape(-1, 32)
        self.recs = recs
        self._cur = dict(cur or {})
        self._ref = dict(ref or {})
        self.size = {}
        for r in recs[recs[:, 0] == REC_CURPIC]:
            self.size[("cur", int(r[1]))] = (int(r[4]), int(r[3]))
        for r in recs[recs[:, 0] == REC_REFPIC]:
            self.size[("ref", int(r[1]))] = (int(r[3]), int(r[2]))

    def of_type(self, t):
        return self.recs[self.recs[:, 0] == t]

    def cur(self, n):
        if n not in self._cur:
            h, w = self.size[("cur", n)]
            self._cur[n] = np.fromfile(os.path.join(self.dir, "cur_%d.y" % n), np.uint8).reshape(h, w)
        return self._cur[n]

    def ref(self, i):
        if i not in self._ref:
            h, w = self.size[("ref", i)]
            self._ref[i] = np.fromfile(os.path.join(self.dir, "ref_%d.y" % i), np.uint8).reshape(h, w)
        return self._ref[i]

    def ref_sub(self, i):
        h, w = self.size[("ref", i)]
        return np.fromfile(os.path.join(self.dir, "ref_%d.sub" % i), np.uint8).reshape(4, 4, h + 40, w + 40)

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
        np.savez_compressed(path, **d)

    @staticmethod
    def load_npz(path):
        z = np.load(path)
        cur = {int(k[4:]): z[k] for k in z.files if k.startswith("cur_")}
        ref = {int(k[4:]): z[k] for k in z.files if k.startswith("ref_")}
        t = Trace(None, z["recs"], cur, ref)
        for n, a in cur.items():
            t.size[("cur", n)] = a.shape
        for i, a in ref.items():
            t.size[("ref", i)] = a.shape
        return t


def cfg_from_trace(trace, search_range=32, full_search=2, metric=(0, 2, 2), mv_limits=(-511, 511)):
    """Encoder-wide parameters recoverable from the records themselves."""
    sub = trace.of_type(REC_SUBPEL)
    ffs = trace.of_type(REC_FFS)
    rdopt = int(ffs[0][25]) if len(ffs) else 1
    t8 = int(sub[:, 22].max()) if len(sub) else 0
    return dict(search_range=search_range, rdopt=rdopt, full_search=full_search, metric=metric,
                transform8x8=t8, mv_limit_min_y=mv_limits[0], mv_limit_max_y=mv_limits[1])
