"""ctypes binding of libme_cuda.so (include/me_cuda.h).  Host-side mirror of the reference interface: method names
follow the reference functions they replace (FastFullPelBlockMotionSearch, SubPelBlockMotionSearch, ...), same
argument meaning, same return value; failures raise MeCudaError.  No CPU fallback: loading fails loudly when the
library is missing, and init fails when there is no CUDA device.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libme_cuda.so")

NUM_PART = 41
PART_TYPE = [1, 2, 2, 3, 3, 4, 4, 4, 4] + [5] * 8 + [6] * 8 + [7] * 16
PART_BLK = [0, 0, 8, 0, 2, 0, 2, 8, 10, 0, 2, 4, 6, 8, 10, 12, 14, 0, 1, 2, 3, 8, 9, 10, 11] + list(range(16))
P_SLICE, B_SLICE = 0, 1

JOB_DTYPE = np.dtype([("opix_x", "<i2"), ("opix_y", "<i2"), ("list", "i1"), ("ref", "i1"), ("slice_type", "i1"),
                      ("reserved", "i1"), ("pmv16", "<i2", (2,)), ("pmv", "<i2", (NUM_PART, 2)),
                      ("lambda", "<i4", (3,))])
RESULT_DTYPE = np.dtype([("mv", "<i2", (NUM_PART, 2)), ("cost", "<i4", (NUM_PART,))])
BIPRED_JOB_DTYPE = np.dtype([("opix_x", "<i2"), ("opix_y", "<i2"), ("list", "i1"), ("reserved", "i1", (3,)), ("pmv", "<i2", (2,)),
                             ("pmv_bi", "<i2", (2,)), ("mv", "<i2", (2,)), ("lambda", "<i4", (3,))])
BIPRED_RESULT_DTYPE = np.dtype([("mv", "<i2", (2,)), ("bimv", "<i2", (2,)), ("cost", "<i4"), ("iterlist", "<i4")])
assert JOB_DTYPE.itemsize == 188 and RESULT_DTYPE.itemsize == 328
# me_mb_motion: motion state around a macroblock (ref_idx -128 = 4x4 block not available), me_pms_item
MOTION_DTYPE = np.dtype([("ref_idx", "i1", (2, 5, 6)), ("mv", "<i2", (2, 5, 6, 2)), ("mbaff", "i1"), ("reserved", "i1", (3,)),
                         ("d_left_ref_idx", "i1", (2, 4)), ("d_left_mv", "<i2", (2, 4, 2))])
PMS_ITEM_DTYPE = np.dtype([("blocktype", "i1"), ("block8x8", "i1")])
PMS_UNAVAILABLE = -128
ABI_VERSION = 2            # ME_CUDA_ABI_VERSION of include/me_cuda.h
DEBLOCK_MB_DTYPE = np.dtype([("strength", "u1", (2, 4, 16)), ("qp", "i1"), ("qpc", "i1", (2,)), ("alpha_off", "i1"), ("beta_off", "i1"),
                             ("luma_skip", "u1"), ("reserved", "u1", (10,))])
assert DEBLOCK_MB_DTYPE.itemsize == 144
assert BIPRED_JOB_DTYPE.itemsize == 32 and BIPRED_RESULT_DTYPE.itemsize == 16
assert MOTION_DTYPE.itemsize == 344 and PMS_ITEM_DTYPE.itemsize == 2

# every symbol include/me_cuda.h declares (tests check that the library exports all of them)
ABI_SYMBOLS = [
    "me_cuda_set_error_handler", "me_cuda_last_error", "me_cuda_init", "me_cuda_shutdown",
    "me_cuda_ref_create", "me_cuda_ref_create_u8", "me_cuda_ref_destroy", "me_cuda_ref_download_plane",
    "me_cuda_set_current", "me_cuda_set_current_u8", "me_cuda_set_ref_list", "me_cuda_set_slice_type",
    "me_cuda_select_current_view", "me_cuda_mb_begin", "me_cuda_ffs_setup", "me_cuda_ffs_is_set_up",
    "me_cuda_FastFullPelBlockMotionSearch", "me_cuda_FullPelBlockMotionSearch", "me_cuda_SubPelBlockMotionSearch",
    "me_cuda_FullPelBlockMotionBiPred", "me_cuda_SubPelBlockSearchBiPred",
    "me_cuda_EPZSSubPelBlockMotionSearch", "me_cuda_EPZSSubPelBlockSearchBiPred", "me_cuda_UMHEXSubPelBlockMotionSearch",
    "me_cuda_smpUMHEXSubPelBlockMotionSearch", "me_cuda_smpUMHEXFullSubPelBlockMotionSearch",
    "me_cuda_frame_search", "me_cuda_frame_search_device", "me_cuda_frame_bipred", "me_cuda_ref_create_device",
    "me_cuda_set_current_device", "me_cuda_stream", "me_cuda_sync", "me_cuda_launch_count", "me_cuda_subpel_region_stats",
    "me_cuda_luma_prediction", "me_cuda_pred_cost8x8", "me_cuda_ref_download_region", "me_cuda_deblock_frame", "me_cuda_ref_create_deblocked",
    "me_cuda_measure_int_peak", "me_cuda_ref_update_u8", "me_cuda_ref_update_device", "me_cuda_set_timing",
    "me_cuda_kernel_time_ms", "me_cuda_set_ref_weights", "me_cuda_set_bipred_weights", "me_cuda_set_bipred_chroma_weights",
    "me_cuda_set_current_chroma", "me_cuda_set_current_chroma_u8", "me_cuda_ref_set_chroma", "me_cuda_ref_set_chroma_u8",
    "me_cuda_set_chroma_state", "me_cuda_ref_set_chroma_planes",
    "me_cuda_pms_begin", "me_cuda_pms_cached", "me_cuda_pms_end", "me_cuda_pms_stats",
]


class MeCudaError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("bit_depth", C.c_int32),
                ("search_range", C.c_int32), ("max_num_references", C.c_int32), ("rdopt", C.c_int32),
                ("full_search", C.c_int32), ("me_metric", C.c_int32 * 3), ("transform8x8", C.c_int32),
                ("mv_limit_min_y", C.c_int32), ("mv_limit_max_y", C.c_int32), ("chroma_me_enable", C.c_int32),
                ("weighted_me", C.c_int32), ("device", C.c_int32), ("chroma_me_weight", C.c_int32), ("yuv_format", C.c_int32),
                ("reserved", C.c_int32 * 6)]


def load_library(path=LIB_PATH):
    if not os.path.exists(path):
        raise MeCudaError("libme_cuda.so not built (%s): run __graft_entry__.build(); there is no CPU fallback" % path)
    L = C.CDLL(path)
    vp, sp, ip = C.c_void_p, C.POINTER(C.c_short), C.POINTER(C.c_int)
    L.me_cuda_last_error.restype = C.c_char_p
    L.me_cuda_init.argtypes = [C.POINTER(Params)]
    L.me_cuda_ref_create_u8.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32)]
    L.me_cuda_ref_create.argtypes = [vp, C.c_int, C.c_int, vp, C.POINTER(C.c_int32)]
    L.me_cuda_ref_create_device.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.POINTER(C.c_int32)]
    L.me_cuda_ref_destroy.argtypes = [C.c_int32]
    L.me_cuda_ref_download_plane.argtypes = [C.c_int32, C.c_int, C.c_int, vp, C.c_int]
    L.me_cuda_set_current_u8.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.me_cuda_set_current.argtypes = [vp, C.c_int, C.c_int]
    L.me_cuda_set_current_device.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
    L.me_cuda_set_ref_list.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int32)]
    L.me_cuda_set_slice_type.argtypes = [C.c_int]
    L.me_cuda_set_ref_weights.argtypes = [C.c_int] * 6
    L.me_cuda_set_bipred_chroma_weights.argtypes = [C.c_int] * 5
    L.me_cuda_set_current_chroma_u8.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
    L.me_cuda_ref_set_chroma_u8.argtypes = [C.c_int32, vp, vp, C.c_int, C.c_int, C.c_int]
    L.me_cuda_set_chroma_state.argtypes = [C.c_int]
    L.me_cuda_ref_set_chroma_planes.argtypes = [C.c_int32] * 3
    L.me_cuda_set_bipred_weights.argtypes = [C.c_int] * 6
    L.me_cuda_mb_begin.argtypes = [C.c_int, C.c_int]
    L.me_cuda_select_current_view.argtypes = [C.c_int]
    L.me_cuda_ffs_setup.argtypes = [C.c_int] * 4
    L.me_cuda_ffs_is_set_up.argtypes = [C.c_int] * 2
    L.me_cuda_FastFullPelBlockMotionSearch.argtypes = ([C.c_short, C.c_int, C.c_int, C.c_int, C.c_int, C.c_short, C.c_short,
                                                        sp, sp, C.c_int, C.c_int, C.c_int])
    L.me_cuda_FullPelBlockMotionSearch.argtypes = L.me_cuda_FastFullPelBlockMotionSearch.argtypes
    L.me_cuda_SubPelBlockMotionSearch.argtypes = ([C.c_short, C.c_int, C.c_int, C.c_int, C.c_int, C.c_short, C.c_short,
                                                   sp, sp, C.c_int, C.c_int, C.c_int, ip])
    L.me_cuda_FullPelBlockMotionBiPred.argtypes = ([C.c_short, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_short] * 4
                                                   + [sp, sp, sp, sp, C.c_int, C.c_int, C.c_int])
    L.me_cuda_SubPelBlockSearchBiPred.argtypes = ([C.c_short, C.c_int, C.c_int, C.c_int, C.c_int, C.c_short, C.c_short,
                                                   sp, sp, sp, sp, C.c_int, C.c_int, C.c_int, ip])
    s2 = C.POINTER(C.c_short * 2)
    L.me_cuda_EPZSSubPelBlockMotionSearch.argtypes = [C.c_short, C.c_int, C.c_int, C.c_int, C.c_int, s2, s2, C.c_int, C.c_int, C.c_int, ip, C.c_int]
    L.me_cuda_EPZSSubPelBlockSearchBiPred.argtypes = [C.c_short, C.c_int, C.c_int, C.c_int, C.c_int, s2, s2, s2, s2, C.c_int, C.c_int, C.c_int, ip]
    pat = [C.c_short, C.c_int, C.c_int, C.c_int, C.c_int, C.c_short, C.c_short, sp, sp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.me_cuda_UMHEXSubPelBlockMotionSearch.argtypes = pat
    L.me_cuda_smpUMHEXSubPelBlockMotionSearch.argtypes = pat + [C.c_int] * 4
    L.me_cuda_smpUMHEXFullSubPelBlockMotionSearch.argtypes = pat + [C.c_int] * 2
    L.me_cuda_frame_search.argtypes = [vp, C.c_int, vp, vp]
    L.me_cuda_frame_search_device.argtypes = [vp, C.c_int, vp, vp, vp]
    L.me_cuda_frame_bipred.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]
    L.me_cuda_stream.restype = vp
    L.me_cuda_launch_count.restype = C.c_int64
    L.me_cuda_measure_int_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    L.me_cuda_ref_update_u8.argtypes = [C.c_int32, vp, C.c_int, C.c_int, C.c_int]
    L.me_cuda_ref_update_device.argtypes = [C.c_int32, vp, C.c_int, C.c_int, C.c_int, vp]
    L.me_cuda_set_timing.argtypes = [C.c_int]
    L.me_cuda_kernel_time_ms.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.me_cuda_pms_begin.argtypes = [vp, C.c_int, vp, ip]
    L.me_cuda_pms_cached.argtypes = [C.c_int, C.c_int]
    L.me_cuda_pms_end.argtypes = [C.c_int, C.c_int]
    L.me_cuda_pms_stats.argtypes = [C.POINTER(C.c_int64)] * 3
    L.me_cuda_luma_prediction.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
    L.me_cuda_pred_cost8x8.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    L.me_cuda_ref_download_region.argtypes = [C.c_int32] + [C.c_int] * 6 + [vp, C.c_int]
    L.me_cuda_deblock_frame.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]
    L.me_cuda_ref_create_deblocked.argtypes = [C.c_int, C.c_int, vp, C.POINTER(C.c_int32)]
    return L


class MeCuda:
    """One initialised libme_cuda context (the library keeps global state like the reference does)."""

    def __init__(self, width, height, search_range=32, max_num_references=5, rdopt=1, full_search=2,
                 me_metric=(0, 2, 2), transform8x8=0, mv_limits=(-511, 511), device=-1, lib_path=LIB_PATH,
                 chroma_me=0, chroma_me_weight=1, yuv_format=1, bit_depth=8):
        """bit_depth > 8: pictures are uint16 arrays and go through the row-pointer entry points (ref_create_imgpel,
        set_current_imgpel); the fast-full-search drop-in path serves them, the 8-bit-only entry points refuse"""
        self.L = load_library(lib_path)
        p = Params()
        p.abi_version = ABI_VERSION
        p.width, p.height, p.bit_depth = width, height, bit_depth
        p.search_range, p.max_num_references, p.rdopt, p.full_search = search_range, max_num_references, rdopt, full_search
        for i in range(3):
            p.me_metric[i] = me_metric[i]
        p.transform8x8 = transform8x8
        p.mv_limit_min_y, p.mv_limit_max_y = mv_limits
        p.device = device
        p.chroma_me_enable, p.chroma_me_weight, p.yuv_format = chroma_me, chroma_me_weight, yuv_format
        self.width, self.height = width, height
        self._check(self.L.me_cuda_init(C.byref(p)))
        self.open = True

    def _check(self, rc):
        if rc != 0:
            raise MeCudaError(self.L.me_cuda_last_error().decode())

    def close(self):
        if getattr(self, "open", False):
            self.L.me_cuda_shutdown()
            self.open = False

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- pictures
    def ref_create(self, luma):
        luma = np.ascontiguousarray(luma, np.uint8)
        h = C.c_int32(0)
        self._check(self.L.me_cuda_ref_create_u8(luma.ctypes.data, luma.strides[0], luma.shape[1], luma.shape[0], C.byref(h)))
        return h.value

    def ref_create_device(self, ptr, stride, stream=None):
        h = C.c_int32(0)
        self._check(self.L.me_cuda_ref_create_device(ptr, stride, self.width, self.height, stream, C.byref(h)))
        return h.value

    def ref_create_imgpel(self, luma_u16, want_planes=False):
        """JM-style entry: 16-bit rows in, optionally the 16 planes back as imgY_sub[4][4][h+40][w+40] (u16)."""
        luma_u16 = np.ascontiguousarray(luma_u16, np.uint16)
        hgt, wid = luma_u16.shape
        rows = (C.c_void_p * hgt)(*[luma_u16.ctypes.data + y * luma_u16.strides[0] for y in range(hgt)])
        h = C.c_int32(0)
        sub = None
        if want_planes:
            hp, wp = hgt + 40, wid + 40
            sub = np.zeros((4, 4, hp, wp), np.uint16)
            # imgpel**** : [4] -> [4] -> [hp] row pointers
            self._rowptr = [[(C.c_void_p * hp)(*[sub[fy, fx].ctypes.data + y * sub.strides[2] for y in range(hp)])
                             for fx in range(4)] for fy in range(4)]
            self._lvl2 = [(C.c_void_p * 4)(*[C.cast(self._rowptr[fy][fx], C.c_void_p).value for fx in range(4)]) for fy in range(4)]
            self._lvl1 = (C.c_void_p * 4)(*[C.cast(self._lvl2[fy], C.c_void_p).value for fy in range(4)])
            self._check(self.L.me_cuda_ref_create(rows, wid, hgt, self._lvl1, C.byref(h)))
        else:
            self._check(self.L.me_cuda_ref_create(rows, wid, hgt, None, C.byref(h)))
        return h.value, sub

    def ref_update(self, h, luma):
        """luma: uint8 ndarray (pinned memory makes the upload asynchronous) or (ptr, stride) of host memory"""
        height = self.height
        if isinstance(luma, tuple):
            ptr, stride = luma
        else:
            luma = np.ascontiguousarray(luma, np.uint8)
            ptr, stride, height = luma.ctypes.data, luma.strides[0], luma.shape[0]      # a frame, or a field of the frame size
        self._check(self.L.me_cuda_ref_update_u8(h, ptr, stride, self.width, height))

    def ref_update_device(self, h, ptr, stride, stream=None):
        self._check(self.L.me_cuda_ref_update_device(h, ptr, stride, self.width, self.height, stream))

    def set_timing(self, on):
        self._check(self.L.me_cuda_set_timing(1 if on else 0))

    def kernel_time_ms(self, which):
        ms, n = C.c_double(0), C.c_int(0)
        self._check(self.L.me_cuda_kernel_time_ms(which, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def ref_destroy(self, h):
        self._check(self.L.me_cuda_ref_destroy(h))

    def ref_plane(self, h, fy, fx, height=None):
        """height: rows of the picture behind the handle when it is a field (default: the initialised frame height)"""
        out = np.zeros(((height or self.height) + 40, self.width + 40), np.uint8)
        self._check(self.L.me_cuda_ref_download_plane(h, fy, fx, out.ctypes.data, out.strides[0]))
        return out

    def ref_planes(self, h, height=None):
        return np.stack([np.stack([self.ref_plane(h, fy, fx, height) for fx in range(4)]) for fy in range(4)])

    def set_current(self, luma):
        luma = np.ascontiguousarray(luma, np.uint8)
        self._check(self.L.me_cuda_set_current_u8(luma.ctypes.data, luma.strides[0], luma.shape[1], luma.shape[0]))

    def set_current_imgpel(self, luma_u16):
        """JM-style entry (me_cuda_set_current): rows of 16-bit samples, as the reference holds imgY_org"""
        luma_u16 = np.ascontiguousarray(luma_u16, np.uint16)
        hgt, wid = luma_u16.shape
        rows = (C.c_void_p * hgt)(*[luma_u16.ctypes.data + y * luma_u16.strides[0] for y in range(hgt)])
        self._check(self.L.me_cuda_set_current(rows, wid, hgt))

    def set_ref_list(self, lst, handles):
        arr = (C.c_int32 * max(1, len(handles)))(*handles)
        self._check(self.L.me_cuda_set_ref_list(lst, len(handles), arr))

    def set_slice_type(self, t):
        self._check(self.L.me_cuda_set_slice_type(t))

    def set_ref_weights(self, lst, ref, apply, weight=32, offset=0, log_denom=5):
        """weighted-prediction ME: wp_weight / wp_offset [list][ref][0] and luma_log_weight_denom of a list entry"""
        self._check(self.L.me_cuda_set_ref_weights(lst, ref, int(apply), int(weight), int(offset), int(log_denom)))

    def set_bipred_chroma_weights(self, lst, component, w1, w2, offset_bi=0):
        self._check(self.L.me_cuda_set_bipred_chroma_weights(lst, component, int(w1), int(w2), int(offset_bi)))

    # ---- chroma terms of the integer-pel searches (ChromaMEEnable)
    def set_current_chroma(self, u, v):
        u, v = np.ascontiguousarray(u, np.uint8), np.ascontiguousarray(v, np.uint8)
        self._check(self.L.me_cuda_set_current_chroma_u8(u.ctypes.data, v.ctypes.data, u.strides[0], u.shape[1], u.shape[0]))

    def ref_set_chroma(self, h, u, v):
        u, v = np.ascontiguousarray(u, np.uint8), np.ascontiguousarray(v, np.uint8)
        self._check(self.L.me_cuda_ref_set_chroma_u8(h, u.ctypes.data, v.ctypes.data, u.strides[0], u.shape[1], u.shape[0]))

    def ref_set_chroma_planes(self, h, h_u, h_v):
        """4:4:4: attach the pictures made from the U and V planes (their quarter-pel plane sets are the chroma sub-images)"""
        self._check(self.L.me_cuda_ref_set_chroma_planes(h, h_u, h_v))

    def set_chroma_state(self, enable):
        """the global ChromaMEEnable as the next search call sees it (mv-search.c:612, :779)"""
        self._check(self.L.me_cuda_set_chroma_state(int(enable)))

    def set_bipred_weights(self, lst, apply, w1=32, w2=32, offset_bi=0, log_denom=5):
        """weighted bi-prediction for bi-predictive searches called with list argument `lst` (me_fullsearch.c:202-227)"""
        self._check(self.L.me_cuda_set_bipred_weights(lst, int(apply), int(w1), int(w2), int(offset_bi), int(log_denom)))

    # ---- drop-in sequence
    def select_current_view(self, view):
        """0: the current picture as uploaded; 1 / 2: its top / bottom field (interlace: field pictures, MBAFF field macroblocks)"""
        self._check(self.L.me_cuda_select_current_view(view))

    def mb_begin(self, opix_x, opix_y):
        self._check(self.L.me_cuda_mb_begin(opix_x, opix_y))

    def ffs_setup(self, lst, ref, pmv16_x, pmv16_y):
        self._check(self.L.me_cuda_ffs_setup(lst, ref, pmv16_x, pmv16_y))

    def FastFullPelBlockMotionSearch(self, ref, lst, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y,
                                     search_range, min_mcost, lambda_factor):
        mx, my = C.c_short(0), C.c_short(0)
        cost = self.L.me_cuda_FastFullPelBlockMotionSearch(ref, lst, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y,
                                                           C.byref(mx), C.byref(my), search_range, min_mcost, lambda_factor)
        return mx.value, my.value, cost

    def FullPelBlockMotionSearch(self, ref, lst, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y, mv_x, mv_y,
                                 search_range, min_mcost, lambda_factor):
        """SearchMode = -1: mv_x / mv_y in = search centre (integer pel); returns (mv_x, mv_y, cost)"""
        mx, my = C.c_short(mv_x), C.c_short(mv_y)
        cost = self.L.me_cuda_FullPelBlockMotionSearch(ref, lst, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y,
                                                       C.byref(mx), C.byref(my), search_range, min_mcost, lambda_factor)
        return mx.value, my.value, cost

    def SubPelBlockMotionSearch(self, ref, lst, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y, mv_x, mv_y,
                                search_pos2, search_pos4, min_mcost, lambda_factor):
        mx, my = C.c_short(mv_x), C.c_short(mv_y)
        lam = (C.c_int * 3)(*lambda_factor)
        cost = self.L.me_cuda_SubPelBlockMotionSearch(ref, lst, pic_pix_x, pic_pix_y, blocktype, pred_mv_x, pred_mv_y,
                                                      C.byref(mx), C.byref(my), search_pos2, search_pos4, min_mcost, lam)
        return mx.value, my.value, cost

    def FullPelBlockMotionBiPred(self, ref, lst, pic_pix_x, pic_pix_y, blocktype, pmv1, pmv2, mv, s_mv, search_range,
                                 min_mcost, lambda_factor):
        mx, my, sx, sy = C.c_short(mv[0]), C.c_short(mv[1]), C.c_short(s_mv[0]), C.c_short(s_mv[1])
        cost = self.L.me_cuda_FullPelBlockMotionBiPred(ref, lst, pic_pix_x, pic_pix_y, blocktype, pmv1[0], pmv1[1], pmv2[0], pmv2[1],
                                                       C.byref(mx), C.byref(my), C.byref(sx), C.byref(sy), search_range,
                                                       min_mcost, lambda_factor)
        return mx.value, my.value, cost

    def SubPelBlockSearchBiPred(self, ref, lst, pic_pix_x, pic_pix_y, blocktype, pmv, mv, s_mv, search_pos2, search_pos4,
                                min_mcost, lambda_factor):
        mx, my, sx, sy = C.c_short(mv[0]), C.c_short(mv[1]), C.c_short(s_mv[0]), C.c_short(s_mv[1])
        lam = (C.c_int * 3)(*lambda_factor)
        cost = self.L.me_cuda_SubPelBlockSearchBiPred(ref, lst, pic_pix_x, pic_pix_y, blocktype, pmv[0], pmv[1],
                                                      C.byref(mx), C.byref(my), C.byref(sx), C.byref(sy),
                                                      search_pos2, search_pos4, min_mcost, lam)
        return mx.value, my.value, cost

    # ---- sub-pel stages of the pattern searches (include/me_cuda.h): pattern 1 EPZSSubPelBlockMotionSearch, 2 UMHEXSubPelBlockMotionSearch,
    # 3 smpUMHEXSubPelBlockMotionSearch, 4 smpUMHEXFullSubPelBlockMotionSearch (the UMHexagonS ones take the quarter-pel factor lam[2])
    def SubPelPattern(self, pattern, ref, lst, pic_pix_x, pic_pix_y, blocktype, pmv, mv, search_pos2, search_pos4, min_mcost, lambda_factor,
                      subthres=0, uplayer=(0, 0), thresholds=(1000, 400)):
        lam = (C.c_int * 3)(*lambda_factor)
        if pattern == 1:
            pm, m = (C.c_short * 2)(*pmv), (C.c_short * 2)(*mv)
            cost = self.L.me_cuda_EPZSSubPelBlockMotionSearch(ref, lst, pic_pix_x, pic_pix_y, blocktype, pm, m, search_pos2, search_pos4,
                                                              min_mcost, lam, subthres)
            return m[0], m[1], cost
        mx, my = C.c_short(mv[0]), C.c_short(mv[1])
        a = (ref, lst, pic_pix_x, pic_pix_y, blocktype, pmv[0], pmv[1], C.byref(mx), C.byref(my), search_pos2, search_pos4, min_mcost, lam[2])
        if pattern == 2:
            cost = self.L.me_cuda_UMHEXSubPelBlockMotionSearch(*a)
        elif pattern == 3:
            cost = self.L.me_cuda_smpUMHEXSubPelBlockMotionSearch(*a, uplayer[0], uplayer[1], thresholds[0], thresholds[1])
        elif pattern == 4:
            cost = self.L.me_cuda_smpUMHEXFullSubPelBlockMotionSearch(*a, thresholds[0], thresholds[1])
        else:
            raise ValueError(pattern)
        return mx.value, my.value, cost

    def EPZSSubPelBlockSearchBiPred(self, ref, lst, pic_pix_x, pic_pix_y, blocktype, pmv1, pmv2, mv, s_mv, search_pos2, search_pos4,
                                    min_mcost, lambda_factor):
        p1, p2, m, sm = (C.c_short * 2)(*pmv1), (C.c_short * 2)(*pmv2), (C.c_short * 2)(*mv), (C.c_short * 2)(*s_mv)
        cost = self.L.me_cuda_EPZSSubPelBlockSearchBiPred(ref, lst, pic_pix_x, pic_pix_y, blocktype, p1, p2, m, sm, search_pos2, search_pos4,
                                                          min_mcost, (C.c_int * 3)(*lambda_factor))
        return m[0], m[1], cost

    # ---- one launch per PartitionMotionSearch call
    def pms_begin(self, state, items, lambda_factor):
        """state: MOTION_DTYPE scalar; items: [(blocktype, block8x8), ...].  Returns False when the configuration is not covered."""
        st = np.ascontiguousarray(state, MOTION_DTYPE)
        it = np.array(items, PMS_ITEM_DTYPE)
        lam = (C.c_int * 3)(*lambda_factor)
        rc = self.L.me_cuda_pms_begin(st.ctypes.data, len(it), it.ctypes.data, lam)
        if rc > 0:
            self._check(rc)
        return rc == 0

    def pms_cached(self, blocktype, block8x8):
        return bool(self.L.me_cuda_pms_cached(blocktype, block8x8))

    def pms_end(self, blocktype, block8x8):
        self._check(self.L.me_cuda_pms_end(blocktype, block8x8))

    def pms_stats(self):
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        self._check(self.L.me_cuda_pms_stats(C.byref(a), C.byref(b), C.byref(c)))
        return dict(launches=a.value, served=b.value, missed=c.value)

    # ---- frame batch
    def frame_search(self, jobs, want_int=False):
        jobs = np.ascontiguousarray(jobs)
        assert jobs.dtype == JOB_DTYPE
        out = np.zeros(len(jobs), RESULT_DTYPE)
        out_int = np.zeros(len(jobs), RESULT_DTYPE) if want_int else None
        self._check(self.L.me_cuda_frame_search(jobs.ctypes.data, len(jobs), out.ctypes.data,
                                                out_int.ctypes.data if want_int else None))
        return (out, out_int) if want_int else out

    def frame_bipred(self, jobs, refinements, search_range, subpel):
        """bi-predictive refinement of B-slice macroblocks (mv-search.c:864-1034): jobs of BIPRED_JOB_DTYPE"""
        jobs = np.ascontiguousarray(jobs)
        assert jobs.dtype == BIPRED_JOB_DTYPE
        out = np.zeros(len(jobs), BIPRED_RESULT_DTYPE)
        self._check(self.L.me_cuda_frame_bipred(jobs.ctypes.data, len(jobs), refinements, search_range, subpel, out.ctypes.data))
        return out

    def luma_prediction(self, blocks, bsx, bsy):
        """LumaPrediction (macroblock.c:832) of blocks (mc.PRED_BLOCK_DTYPE) of bsx x bsy samples -> uint8 [n, bsy, bsx]"""
        from .mc import PRED_BLOCK_DTYPE
        blocks = np.ascontiguousarray(blocks)
        assert blocks.dtype == PRED_BLOCK_DTYPE
        out = np.zeros((len(blocks), bsy, bsx), np.uint8)
        self._check(self.L.me_cuda_luma_prediction(blocks.ctypes.data, len(blocks), bsx, bsy, out.ctypes.data))
        return out

    def pred_cost8x8(self, blocks, metric=2):
        """per 8x8 region (four PRED_BLOCK records, raster order): (sum of distortion4x4, distortion8x8) of original - prediction"""
        from .mc import PRED_BLOCK_DTYPE
        blocks = np.ascontiguousarray(blocks)
        assert blocks.dtype == PRED_BLOCK_DTYPE and len(blocks) % 4 == 0
        n = len(blocks) // 4
        c4, c8 = np.zeros(n, np.int32), np.zeros(n, np.int32)
        self._check(self.L.me_cuda_pred_cost8x8(blocks.ctypes.data, n, metric, c4.ctypes.data, c8.ctypes.data))
        return c4, c8

    def deblock_frame(self, y, u, v, mbs):
        """DeblockFrame (loopFilter.c:87): y, u, v uint16 (imgpel) planes filtered in place (u, v None: 4:0:0); mbs: DEBLOCK_MB_DTYPE"""
        assert y.dtype == np.uint16 and y.flags.c_contiguous and mbs.dtype == DEBLOCK_MB_DTYPE and mbs.flags.c_contiguous

        def rows(a):
            arr = (C.c_void_p * a.shape[0])()
            for i in range(a.shape[0]):
                arr[i] = a.ctypes.data + i * a.strides[0]
            return arr
        ry = rows(y)
        ru, rv = (rows(u), rows(v)) if u is not None else (None, None)
        cw, ch = (u.shape[1], u.shape[0]) if u is not None else (0, 0)
        self._check(self.L.me_cuda_deblock_frame(ry, ru, rv, y.shape[1], y.shape[0], cw, ch, mbs.ctypes.data))

    def ref_create_deblocked(self):
        """the plane set of the picture me_cuda_deblock_frame has just filtered, built from the copy that stayed on the device"""
        h = C.c_int32(0)
        self._check(self.L.me_cuda_ref_create_deblocked(self.width, self.height, None, C.byref(h)))
        return h.value

    def ref_region(self, h, fy, fx, x0, y0, w, hgt):
        """a rectangle of plane Pic[fy][fx] as imgpel (uint16) rows"""
        out = np.zeros((hgt, w), np.uint16)
        self._check(self.L.me_cuda_ref_download_region(h, fy, fx, x0, y0, w, hgt, out.ctypes.data, w))
        return out

    def frame_search_device(self, d_jobs, n, d_out, d_out_int=None, stream=None):
        self._check(self.L.me_cuda_frame_search_device(d_jobs, n, d_out, d_out_int, stream))

    def set_current_device(self, ptr, stride, stream=None):
        self._check(self.L.me_cuda_set_current_device(ptr, stride, self.width, self.height, stream))

    def stream(self):
        return self.L.me_cuda_stream()

    def sync(self):
        self._check(self.L.me_cuda_sync())

    def launch_count(self):
        return int(self.L.me_cuda_launch_count())

    def subpel_region_stats(self, enable):
        """(partitions served from the staged region, all partitions) of the throughput sub-pel kernel since the count was enabled;
        enable=True restarts the count, False stops it"""
        a, b = C.c_int64(0), C.c_int64(0)
        self._check(self.L.me_cuda_subpel_region_stats(1 if enable else 0, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def measure_int_peak(self, kind):
        v = C.c_double(0)
        rc = self.L.me_cuda_measure_int_peak(kind, C.byref(v))
        if rc:
            raise MeCudaError("me_cuda_measure_int_peak(%d) failed: %d" % (kind, rc))
        return v.value
