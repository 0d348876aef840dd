"""ctypes binding of oracle/liboracle_me.so -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
The product path (h.264_b200/, libme_cuda.so) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle_me.so")
LIB16 = os.path.join(HERE, "liboracle_me16.so")     # the same source compiled with 16-bit samples (more than 8 bits per sample)

NUM_PART = 41
PART_TYPE = [1, 2, 2, 3, 3, 4, 4, 4, 4] + [5] * 8 + [6] * 8 + [7] * 16
PART_BLK = [0, 0, 8, 0, 2, 0, 2, 8, 10, 0, 2, 4, 6, 8, 10, 12, 14, 0, 1, 2, 3, 8, 9, 10, 11] + list(range(16))
BLOCK_SIZE = {1: (16, 16), 2: (16, 8), 3: (8, 16), 4: (8, 8), 5: (8, 4), 6: (4, 8), 7: (4, 4)}
INT_MAX = 2147483647


def part_index(blocktype, blk):
    """partition number p of (block type, top-left 4x4 block index)"""
    for p in range(NUM_PART):
        if PART_TYPE[p] == blocktype and PART_BLK[p] == blk:
            return p
    raise KeyError((blocktype, blk))


class OraCfg(C.Structure):
    _fields_ = [("search_range", C.c_int), ("rdopt", C.c_int), ("full_search", C.c_int),
                ("metric", C.c_int * 3), ("transform8x8", C.c_int),
                ("mv_limit_min_y", C.c_int), ("mv_limit_max_y", C.c_int), ("chroma_me", C.c_int)]


class MbJob(C.Structure):
    _pack_ = 1
    _fields_ = [("opix_x", C.c_int16), ("opix_y", C.c_int16), ("list", C.c_int8), ("ref", C.c_int8),
                ("slice_type", C.c_int8), ("reserved", C.c_int8), ("pmv16", C.c_int16 * 2),
                ("pmv", (C.c_int16 * 2) * NUM_PART), ("lambda_", C.c_int32 * 3)]


class MbResult(C.Structure):
    _fields_ = [("mv", (C.c_int16 * 2) * NUM_PART), ("cost", C.c_int32 * NUM_PART)]


# numpy views of the same layouts (used for bulk job arrays)
JOB_DTYPE = np.dtype([("opix_x", "<i2"), ("opix_y", "<i2"), ("list", "i1"), ("ref", "i1"), ("slice_type", "i1"),
                      ("reserved", "i1"), ("pmv16", "<i2", (2,)), ("pmv", "<i2", (NUM_PART, 2)),
                      ("lambda", "<i4", (3,))])
RESULT_DTYPE = np.dtype([("mv", "<i2", (NUM_PART, 2)), ("cost", "<i4", (NUM_PART,))])
assert JOB_DTYPE.itemsize == 188 and C.sizeof(MbJob) == 188
assert RESULT_DTYPE.itemsize == 328 and C.sizeof(MbResult) == 328


def build(force=False):
    src = os.path.getmtime(os.path.join(HERE, "me_oracle.c"))
    if force or any(not os.path.exists(f) or os.path.getmtime(f) < src for f in (LIB, LIB16)):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    return LIB


_libs = {}
_bits = 8


def set_bit_depth(bits=8):
    """8: the 8-bit oracle (default).  9..14: the 16-bit build, clips at (1 << bits) - 1; pictures are uint16 arrays then."""
    global _bits
    _bits = int(bits)
    if _bits > 8:
        lib().ora_set_max_pel((1 << _bits) - 1)


def pel_dtype():
    return np.uint16 if _bits > 8 else np.uint8


def _stride(a):
    """row stride of a picture in samples (the C side indexes sample pointers)"""
    return a.strides[0] // a.itemsize


def lib():
    path = LIB16 if _bits > 8 else LIB
    _lib = _libs.get(path)
    if _lib is None:
        build()
        L = C.CDLL(path)
        vp, ip, sp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_short)
        L.ora_ref_create.restype = vp
        L.ora_ref_create.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        L.ora_ref_destroy.argtypes = [vp]
        L.ora_ref_plane.restype = vp
        L.ora_ref_plane.argtypes = [vp, C.c_int, C.c_int]
        L.ora_ref_set_weights.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ora_ref_set_chroma.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int]
        L.ora_ref_set_chroma444.argtypes = [vp, vp, vp]
        L.ora_set_chroma.argtypes = [C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int]
        L.ora_set_bipred_weights.argtypes = [C.c_int] * 6
        L.ora_select_bipred_list.argtypes = [C.c_int]
        L.ora_set_bipred_chroma_weights.argtypes = [C.c_int] * 5
        L.ora_set_max_pel.argtypes = [C.c_int]
        L.ora_spiral.argtypes = [C.c_int, vp, vp]
        L.ora_mvbits.argtypes = [C.c_int]
        L.ora_mv_cost.argtypes = [C.c_int] * 5
        L.ora_setup_range.argtypes = [C.POINTER(OraCfg), C.c_int]
        L.ora_block_search_range.argtypes = [C.POINTER(OraCfg), C.c_int, C.c_int]
        L.ora_search_centre.argtypes = [C.POINTER(OraCfg)] + [C.c_int] * 5 + [ip, ip]
        L.ora_ffs_surfaces.argtypes = [C.POINTER(OraCfg), vp, C.c_int, vp] + [C.c_int] * 5 + [vp]
        L.ora_ffs_block.argtypes = [C.POINTER(OraCfg), vp] + [C.c_int] * 10 + [sp, sp]
        L.ora_sad_block.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int]
        L.ora_satd_block.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int]
        L.ora_hadamard4x4.argtypes = [vp]
        L.ora_hadamard8x8.argtypes = [vp]
        L.ora_subpel.argtypes = [C.POINTER(OraCfg), vp, C.c_int, vp] + [C.c_int] * 7 + [sp, sp, C.c_int, C.c_int, C.c_int, ip]
        L.ora_fullpel_block.argtypes = [C.POINTER(OraCfg), vp, C.c_int, vp] + [C.c_int] * 7 + [sp, sp, C.c_int, C.c_int, C.c_int]
        L.ora_bipred_fullpel.argtypes = ([C.POINTER(OraCfg), vp, C.c_int, vp, vp] + [C.c_int] * 7
                                         + [sp, sp, sp, sp, C.c_int, C.c_int, C.c_int])
        L.ora_bipred_subpel.argtypes = ([C.POINTER(OraCfg), vp, C.c_int, vp, vp] + [C.c_int] * 5
                                        + [sp, sp, sp, sp, C.c_int, C.c_int, C.c_int, ip])
        L.ora_epzs_subpel.argtypes = [C.POINTER(OraCfg), vp, C.c_int, vp] + [C.c_int] * 5 + [sp, sp, C.c_int, C.c_int, C.c_int, ip, C.c_int]
        L.ora_epzs_subpel_bipred.argtypes = ([C.POINTER(OraCfg), vp, C.c_int, vp, vp] + [C.c_int] * 7
                                             + [sp, sp, sp, sp, C.c_int, C.c_int, C.c_int, ip])
        L.ora_umhex_subpel.argtypes = [C.POINTER(OraCfg), vp, C.c_int, vp] + [C.c_int] * 5 + [sp, sp, C.c_int, C.c_int]
        L.ora_smpumhex_subpel.argtypes = [C.POINTER(OraCfg), vp, C.c_int, vp] + [C.c_int] * 5 + [sp, sp] + [C.c_int] * 6
        L.ora_smpumhex_full_subpel.argtypes = [C.POINTER(OraCfg), vp, C.c_int, vp] + [C.c_int] * 7 + [sp, sp] + [C.c_int] * 6
        L.ora_mb_search.argtypes = [C.POINTER(OraCfg), vp, C.c_int, vp, vp, vp, vp]
        L.ora_mb_bipred.argtypes = [C.POINTER(OraCfg), vp, C.c_int, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]
        if path == LIB:
            L.ora_luma_prediction.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp]
        if path == LIB:
            L.ora_deblock_frame.argtypes = [vp, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, vp]
        if path == LIB:
            L.ora_get_skip_cost_mb.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, C.c_int, C.c_int]
        if path == LIB:
            L.ora_get_direct_cost8x8.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, ip]
        if path == LIB:
            L.ora_bid_partition_cost.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        _lib = _libs[path] = L
    return _lib


def make_cfg(search_range=32, rdopt=1, full_search=2, metric=(0, 2, 2), transform8x8=0,
             mv_limit_min_y=-511, mv_limit_max_y=511, chroma_me=0):
    c = OraCfg()
    c.chroma_me = chroma_me
    c.search_range, c.rdopt, c.full_search = search_range, rdopt, full_search
    c.metric[0], c.metric[1], c.metric[2] = metric
    c.transform8x8 = transform8x8
    c.mv_limit_min_y, c.mv_limit_max_y = mv_limit_min_y, mv_limit_max_y
    return c


def spiral(R):
    n = (2 * R + 1) ** 2
    sx = np.zeros(n, np.int16)
    sy = np.zeros(n, np.int16)
    lib().ora_spiral(R, sx.ctypes.data, sy.ctypes.data)
    return sx, sy


class Ref:
    """16 quarter-pel planes of one reconstructed luma picture (uint8 HxW)."""

    def __init__(self, luma):
        luma = np.ascontiguousarray(luma, pel_dtype())
        self.h, self.w = luma.shape
        self.dtype = luma.dtype
        self.hp, self.wp = self.h + 40, self.w + 40
        self._luma = luma
        self.ptr = lib().ora_ref_create(luma.ctypes.data, _stride(luma), self.w, self.h)

    def plane(self, fy, fx):
        p = lib().ora_ref_plane(self.ptr, fy, fx)
        buf = ((C.c_uint16 if self.dtype == np.uint16 else C.c_uint8) * (self.hp * self.wp)).from_address(p)
        return np.frombuffer(buf, self.dtype).reshape(self.hp, self.wp)

    def planes(self):
        return np.stack([np.stack([self.plane(fy, fx) for fx in range(4)]) for fy in range(4)])

    def set_chroma(self, u, v):
        """ChromaMEEnable: the picture's reconstructed chroma planes (uint8, same shape)"""
        u, v = np.ascontiguousarray(u, np.uint8), np.ascontiguousarray(v, np.uint8)
        lib().ora_ref_set_chroma(self.ptr, u.ctypes.data, v.ctypes.data, _stride(u), u.shape[1], u.shape[0])

    def set_chroma444(self, ref_u, ref_v):
        """4:4:4: the plane sets (Ref objects, kept alive by the caller) built from the picture's U and V planes"""
        self._c444 = (ref_u, ref_v)
        lib().ora_ref_set_chroma444(self.ptr, ref_u.ptr, ref_v.ptr)

    def set_weights(self, apply, weight=32, offset=0, log_denom=5):
        """weighted-prediction ME: the weights of the (list, ref) entry this picture is about to be searched as"""
        lib().ora_ref_set_weights(self.ptr, int(apply), int(weight), int(offset), int(log_denom))

    def close(self):
        if self.ptr:
            lib().ora_ref_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fullpel_block(cfg, cur, ref, ppx, ppy, bt, ref_idx, slice_type, pmvx, pmvy, mvx, mvy, search_range, min_mcost, lam):
    """FullPelBlockMotionSearch (SearchMode = -1): returns (mv_x, mv_y, cost), vectors in integer pel"""
    mx, my = C.c_short(mvx), C.c_short(mvy)
    cost = lib().ora_fullpel_block(C.byref(cfg), cur.ctypes.data, _stride(cur), ref.ptr, ppx, ppy, bt, ref_idx, slice_type,
                                   pmvx, pmvy, C.byref(mx), C.byref(my), search_range, min_mcost, lam)
    return mx.value, my.value, cost


_chroma_keep = []


def set_chroma(enable, weight=0, cur_u=None, cur_v=None, shift=(3, 3)):
    """chroma state of the next search calls: the global ChromaMEEnable they see, ChromaMEWeight, original chroma planes,
    (chroma_shift_x, chroma_shift_y)"""
    if not enable or cur_u is None:
        lib().ora_set_chroma(0, 0, None, None, 0, shift[0], shift[1])
        return
    cur_u, cur_v = np.ascontiguousarray(cur_u, np.uint8), np.ascontiguousarray(cur_v, np.uint8)
    _chroma_keep[:] = [cur_u, cur_v]
    lib().ora_set_chroma(int(enable), int(weight), cur_u.ctypes.data, cur_v.ctypes.data, _stride(cur_u), shift[0], shift[1])


def set_bipred_weights(lst, apply, w1=32, w2=32, offset_bi=0, log_denom=5):
    """weight set of bi-predictive searches called with list argument `lst` (weighted bi-prediction, computeBiPred*2)"""
    lib().ora_set_bipred_weights(lst, int(apply), int(w1), int(w2), int(offset_bi), int(log_denom))


def set_bipred_chroma_weights(lst, component, w1, w2, offset_bi=0):
    lib().ora_set_bipred_chroma_weights(lst, component, int(w1), int(w2), int(offset_bi))


def select_bipred_list(lst):
    lib().ora_select_bipred_list(lst)


def _shorts(*vals):
    return [C.c_short(v) for v in vals]


def ffs_surfaces(cfg, cur, ref, opix_x, opix_y, cx, cy, rng):
    P = (2 * rng + 1) ** 2
    out = np.zeros((NUM_PART, P), np.int32)
    lib().ora_ffs_surfaces(C.byref(cfg), cur.ctypes.data, _stride(cur), ref.ptr, opix_x, opix_y, cx, cy, rng,
                           out.ctypes.data)
    return out


def ffs_block(cfg, surf_p, setup_range, opix_x, opix_y, cx, cy, pmv_x, pmv_y, search_range, min_mcost, lam):
    mx, my = _shorts(0, 0)
    surf_p = np.ascontiguousarray(surf_p, np.int32)
    cost = lib().ora_ffs_block(C.byref(cfg), surf_p.ctypes.data, setup_range, opix_x, opix_y, cx, cy,
                               pmv_x, pmv_y, search_range, min_mcost, lam, C.byref(mx), C.byref(my))
    return mx.value, my.value, cost


def search_centre(cfg, ref_idx, opix_x, opix_y, pmv16_x, pmv16_y):
    cx, cy = C.c_int(0), C.c_int(0)
    lib().ora_search_centre(C.byref(cfg), ref_idx, opix_x, opix_y, pmv16_x, pmv16_y, C.byref(cx), C.byref(cy))
    return cx.value, cy.value


def subpel(cfg, cur, ref, pic_pix_x, pic_pix_y, blocktype, ref_idx, slice_type, pmv_x, pmv_y, mv_x, mv_y,
           min_mcost, lam, search_pos2=9, search_pos4=9):
    mx, my = _shorts(mv_x, mv_y)
    lam3 = (C.c_int * 3)(*lam)
    cost = lib().ora_subpel(C.byref(cfg), cur.ctypes.data, _stride(cur), ref.ptr, pic_pix_x, pic_pix_y, blocktype,
                            ref_idx, slice_type, pmv_x, pmv_y, C.byref(mx), C.byref(my), search_pos2, search_pos4,
                            min_mcost, lam3)
    return mx.value, my.value, cost


def bipred_fullpel(cfg, cur, ref1, ref2, pic_pix_x, pic_pix_y, blocktype, pmv1, pmv2, mv, s_mv, search_range,
                   min_mcost, lam):
    mx, my, sx, sy = _shorts(mv[0], mv[1], s_mv[0], s_mv[1])
    cost = lib().ora_bipred_fullpel(C.byref(cfg), cur.ctypes.data, _stride(cur), ref1.ptr, ref2.ptr,
                                    pic_pix_x, pic_pix_y, blocktype, pmv1[0], pmv1[1], pmv2[0], pmv2[1],
                                    C.byref(mx), C.byref(my), C.byref(sx), C.byref(sy), search_range, min_mcost, lam)
    return mx.value, my.value, cost


def bipred_subpel(cfg, cur, ref1, ref2, pic_pix_x, pic_pix_y, blocktype, pmv, mv, s_mv, min_mcost, lam,
                  search_pos2=9, search_pos4=9):
    mx, my, sx, sy = _shorts(mv[0], mv[1], s_mv[0], s_mv[1])
    lam3 = (C.c_int * 3)(*lam)
    cost = lib().ora_bipred_subpel(C.byref(cfg), cur.ctypes.data, _stride(cur), ref1.ptr, ref2.ptr,
                                   pic_pix_x, pic_pix_y, blocktype, pmv[0], pmv[1],
                                   C.byref(mx), C.byref(my), C.byref(sx), C.byref(sy), search_pos2, search_pos4,
                                   min_mcost, lam3)
    return mx.value, my.value, cost


# sub-pel stages of the pattern searches: pattern 1 EPZSSubPelBlockMotionSearch, 2 UMHEXSubPelBlockMotionSearch,
# 3 smpUMHEXSubPelBlockMotionSearch, 4 smpUMHEXFullSubPelBlockMotionSearch (lam: three factors; the UMHexagonS ones use lam[2])
def subpel_pattern(cfg, cur, ref, pattern, pic_pix_x, pic_pix_y, blocktype, ref_idx, slice_type, pmv, mv, min_mcost, lam,
                   search_pos2=9, search_pos4=9, subthres=0, uplayer=(0, 0), thresholds=(1000, 400)):
    mx, my = _shorts(mv[0], mv[1])
    L, a = lib(), (C.byref(cfg), cur.ctypes.data, _stride(cur), ref.ptr, pic_pix_x, pic_pix_y, blocktype)
    if pattern == 1:
        cost = L.ora_epzs_subpel(*a, pmv[0], pmv[1], C.byref(mx), C.byref(my), search_pos2, search_pos4, min_mcost, (C.c_int * 3)(*lam), subthres)
    elif pattern == 2:
        cost = L.ora_umhex_subpel(*a, pmv[0], pmv[1], C.byref(mx), C.byref(my), min_mcost, lam[2])
    elif pattern == 3:
        cost = L.ora_smpumhex_subpel(*a, pmv[0], pmv[1], C.byref(mx), C.byref(my), min_mcost, lam[2], uplayer[0], uplayer[1], *thresholds)
    elif pattern == 4:
        cost = L.ora_smpumhex_full_subpel(*a, ref_idx, slice_type, pmv[0], pmv[1], C.byref(mx), C.byref(my), search_pos2, search_pos4,
                                          min_mcost, lam[2], *thresholds)
    else:
        raise ValueError(pattern)
    return mx.value, my.value, cost


def epzs_subpel_bipred(cfg, cur, ref1, ref2, pic_pix_x, pic_pix_y, blocktype, pmv1, pmv2, mv, s_mv, min_mcost, lam,
                       search_pos2=9, search_pos4=9):
    mx, my, sx, sy = _shorts(mv[0], mv[1], s_mv[0], s_mv[1])
    cost = lib().ora_epzs_subpel_bipred(C.byref(cfg), cur.ctypes.data, _stride(cur), ref1.ptr, ref2.ptr, pic_pix_x, pic_pix_y, blocktype,
                                        pmv1[0], pmv1[1], pmv2[0], pmv2[1], C.byref(mx), C.byref(my), C.byref(sx), C.byref(sy),
                                        search_pos2, search_pos4, min_mcost, (C.c_int * 3)(*lam))
    return mx.value, my.value, cost


BIPRED_JOB_DTYPE = np.dtype([("opix_x", "<i2"), ("opix_y", "<i2"), ("list", "i1"), ("reserved", "i1", (3,)), ("pmv", "<i2", (2,)),
                             ("pmv_bi", "<i2", (2,)), ("mv", "<i2", (2,)), ("lambda", "<i4", (3,))])
BIPRED_RESULT_DTYPE = np.dtype([("mv", "<i2", (2,)), ("bimv", "<i2", (2,)), ("cost", "<i4"), ("iterlist", "<i4")])
assert BIPRED_JOB_DTYPE.itemsize == 32 and BIPRED_RESULT_DTYPE.itemsize == 16


def mb_bipred(cfg, cur, ref_l0, ref_l1, jobs, refinements, search_range, subpel):
    """jobs: numpy array of BIPRED_JOB_DTYPE; ref_l0 / ref_l1 = reference 0 of the two lists."""
    jobs = np.ascontiguousarray(jobs)
    out = np.zeros(len(jobs), BIPRED_RESULT_DTYPE)
    L = lib()
    for i in range(len(jobs)):
        L.ora_mb_bipred(C.byref(cfg), cur.ctypes.data, _stride(cur), ref_l0.ptr, ref_l1.ptr,
                        jobs.ctypes.data + i * BIPRED_JOB_DTYPE.itemsize, refinements, search_range, subpel,
                        out.ctypes.data + i * BIPRED_RESULT_DTYPE.itemsize)
    return out


def mb_search(cfg, cur, ref, jobs):
    """jobs: numpy array of JOB_DTYPE (all against the same reference picture).  Returns (final, integer-pel)."""
    jobs = np.ascontiguousarray(jobs)
    out = np.zeros(len(jobs), RESULT_DTYPE)
    out_int = np.zeros(len(jobs), RESULT_DTYPE)
    L = lib()
    for i in range(len(jobs)):
        L.ora_mb_search(C.byref(cfg), cur.ctypes.data, _stride(cur), ref.ptr,
                        jobs.ctypes.data + i * JOB_DTYPE.itemsize,
                        out.ctypes.data + i * RESULT_DTYPE.itemsize,
                        out_int.ctypes.data + i * RESULT_DTYPE.itemsize)
    return out, out_int


# ---- motion-compensated prediction / mode-decision costs (me_oracle.c: ora_luma_prediction, ora_get_skip_cost_mb, ...)
PRED_BLOCK_DTYPE = np.dtype([("pix_x", "<i2"), ("pix_y", "<i2"), ("p_dir", "i1"), ("ref", "i1", (2,)), ("apply_weights", "i1"),
                             ("mv", "<i2", (2, 2)), ("weight", "<i2", (2,)), ("offset", "<i2"), ("log_denom", "<i2")])
assert PRED_BLOCK_DTYPE.itemsize == 24


def _ref_array(refs):
    arr = (C.c_void_p * max(1, len(refs)))()
    for i, r in enumerate(refs):
        arr[i] = r.ptr
    return arr


def luma_prediction(list0, list1, block, bsx, bsy):
    """LumaPrediction (macroblock.c:832) of one block (a PRED_BLOCK_DTYPE record) -> uint8 [bsy, bsx]"""
    b = np.ascontiguousarray(np.array(block, PRED_BLOCK_DTYPE).reshape(1))
    out = np.zeros((bsy, bsx), np.uint8)
    a0, a1 = _ref_array(list0), _ref_array(list1)
    lib().ora_luma_prediction(a0, a1, b.ctypes.data, bsx, bsy, out.ctypes.data)
    return out


def get_skip_cost_mb(cur, list0, list1, blocks16, rdopt, transform8x8, metric=2):
    cur = np.ascontiguousarray(cur, np.uint8)
    b = np.ascontiguousarray(blocks16, PRED_BLOCK_DTYPE)
    a0, a1 = _ref_array(list0), _ref_array(list1)
    return lib().ora_get_skip_cost_mb(cur.ctypes.data, _stride(cur), a0, a1, b.ctypes.data, rdopt, transform8x8, metric)


def get_direct_cost8x8(cur, list0, list1, blocks16, block, rdopt, transform8x8, metric=2, cost8x8=0):
    cur = np.ascontiguousarray(cur, np.uint8)
    b = np.ascontiguousarray(blocks16, PRED_BLOCK_DTYPE)
    a0, a1 = _ref_array(list0), _ref_array(list1)
    c8 = C.c_int(cost8x8)
    c = lib().ora_get_direct_cost8x8(cur.ctypes.data, _stride(cur), a0, a1, b.ctypes.data, block, rdopt, transform8x8, metric, C.byref(c8))
    return c, c8.value


def bid_partition_cost(cur, list0, list1, blocks16, pmv16, blocktype, block8x8, lambda_factor, transform8x8, metric=2):
    cur = np.ascontiguousarray(cur, np.uint8)
    b = np.ascontiguousarray(blocks16, PRED_BLOCK_DTYPE)
    pm = np.ascontiguousarray(pmv16, np.int16)
    assert pm.shape == (16, 2, 2)
    a0, a1 = _ref_array(list0), _ref_array(list1)
    return lib().ora_bid_partition_cost(cur.ctypes.data, _stride(cur), a0, a1, b.ctypes.data, pm.ctypes.data, blocktype, block8x8,
                                        lambda_factor, transform8x8, metric)


# ---- deblocking (me_oracle.c: ora_deblock_frame)
DEBLOCK_MB_DTYPE = np.dtype([("strength", "u1", (2, 4, 16)), ("qp", "i1"), ("qpc", "i1", (2,)), ("alpha_off", "i1"), ("beta_off", "i1"),
                             ("luma_skip", "u1"), ("reserved", "u1", (10,))])
assert DEBLOCK_MB_DTYPE.itemsize == 144


def deblock_frame(y, u, v, mbs):
    """DeblockFrame (loopFilter.c:87) on copies of the planes (uint8); u, v None for 4:0:0 -> filtered (y, u, v)"""
    y = np.array(y, np.uint8, order="C")
    h, w = y.shape
    mbs = np.ascontiguousarray(mbs, DEBLOCK_MB_DTYPE)
    assert len(mbs) == (w // 16) * (h // 16)
    if u is None:
        lib().ora_deblock_frame(y.ctypes.data, _stride(y), None, None, 0, w, h, mbs.ctypes.data)
        return y, None, None
    u, v = np.array(u, np.uint8, order="C"), np.array(v, np.uint8, order="C")
    lib().ora_deblock_frame(y.ctypes.data, _stride(y), u.ctypes.data, v.ctypes.data, _stride(u), w, h, mbs.ctypes.data)
    return y, u, v
