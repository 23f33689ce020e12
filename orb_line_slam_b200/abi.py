"""ctypes mirror of include/olf_abi.h: POD layouts shared by the product wrapper and the tests."""
from __future__ import annotations
import ctypes as C
import numpy as np

OLF_OK, OLF_ERR_ARG, OLF_ERR_CUDA, OLF_ERR_CAPACITY, OLF_ERR_NO_DEVICE, OLF_ERR_INTERNAL = 0, -1, -2, -3, -4, -5

# olf_keypoint / olf_keyline as numpy structured dtypes (C layout, no padding: all 4-byte fields)
KEYPOINT = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4")])
KEYLINE = np.dtype([("angle", "<f4"), ("class_id", "<i4"), ("octave", "<i4"), ("pt_x", "<f4"), ("pt_y", "<f4"),
                    ("response", "<f4"), ("size", "<f4"),
                    ("startPointX", "<f4"), ("startPointY", "<f4"), ("endPointX", "<f4"), ("endPointY", "<f4"),
                    ("sPointInOctaveX", "<f4"), ("sPointInOctaveY", "<f4"), ("ePointInOctaveX", "<f4"), ("ePointInOctaveY", "<f4"),
                    ("lineLength", "<f4"), ("numOfPixels", "<i4")])
assert KEYPOINT.itemsize == 24 and KEYLINE.itemsize == 68


class LineParams(C.Structure):
    """olf_line_params; defaults = Examples/PL/*.yaml (lsd_* keys) of the reference."""
    _fields_ = [("lsd_nfeatures", C.c_int), ("min_line_length", C.c_double), ("lsd_refine", C.c_int),
                ("lsd_scale", C.c_double), ("lsd_sigma_scale", C.c_double), ("lsd_quant", C.c_double),
                ("lsd_ang_th", C.c_double), ("lsd_log_eps", C.c_double), ("lsd_density_th", C.c_double),
                ("lsd_n_bins", C.c_int)]

    def __init__(self, lsd_nfeatures=500, min_line_length=0.025, lsd_refine=0, lsd_scale=1.2, lsd_sigma_scale=0.6,
                 lsd_quant=2.0, lsd_ang_th=22.5, lsd_log_eps=1.0, lsd_density_th=0.6, lsd_n_bins=1024):
        super().__init__(lsd_nfeatures, min_line_length, lsd_refine, lsd_scale, lsd_sigma_scale, lsd_quant,
                         lsd_ang_th, lsd_log_eps, lsd_density_th, lsd_n_bins)


class LineMatchParams(C.Structure):
    """olf_line_match_params; defaults = src/Config.cpp:42-87 of the reference."""
    _fields_ = [("best_lr_matches", C.c_int), ("min_ratio_12_l", C.c_double), ("line_sim_th", C.c_double),
                ("matching_s_ws", C.c_int), ("min_disp", C.c_double), ("line_horiz_th", C.c_double),
                ("stereo_overlap_th", C.c_double), ("ls_min_disp_ratio", C.c_double)]

    def __init__(self, best_lr_matches=1, min_ratio_12_l=0.9, line_sim_th=0.75, matching_s_ws=10, min_disp=1.0,
                 line_horiz_th=0.1, stereo_overlap_th=0.75, ls_min_disp_ratio=0.7):
        super().__init__(best_lr_matches, min_ratio_12_l, line_sim_th, matching_s_ws, min_disp, line_horiz_th,
                         stereo_overlap_th, ls_min_disp_ratio)


class Camera(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("bf", C.c_float),
                ("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float), ("max_y", C.c_float)]


_P = C.c_void_p


class FrontendParams(C.Structure):
    """olf_frontend_params"""
    _fields_ = [("nfeatures", C.c_int), ("scale_factor", C.c_float), ("nlevels", C.c_int), ("ini_th_fast", C.c_int),
                ("min_th_fast", C.c_int), ("has_lines", C.c_int), ("line", LineParams), ("line_match", LineMatchParams),
                ("cam", Camera), ("cap_points", C.c_int), ("cap_lines", C.c_int)]


class FrameHeader(C.Structure):
    _fields_ = [("n_l", C.c_int), ("n_r", C.c_int), ("m_l", C.c_int), ("m_r", C.c_int), ("cap_points", C.c_int),
                ("cap_lines", C.c_int), ("status", C.c_int), ("reserved", C.c_int)]


class FrameOffsets(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("kps_l", "desc_l", "kps_r", "desc_r", "u_right", "depth", "kls_l", "ldesc_l",
                                          "kls_r", "ldesc_r", "lmatch", "ldisp", "lle", "total")]


class VocabDesc(C.Structure):
    """olf_vocab_desc: a DBoW2 vocabulary tree flattened (include/olf_abi.h)."""
    _fields_ = [("k", C.c_int), ("L", C.c_int), ("n_nodes", C.c_int), ("node_desc", _P), ("child_begin", _P), ("child_count", _P),
                ("children", _P), ("word_id", _P), ("weight", _P)]


class BowMatchArgs(C.Structure):
    _fields_ = [("kf_desc", _P), ("kf_kps_un", _P), ("n_kf", C.c_int), ("kf_has_point", _P),
                ("kf_fv_node", _P), ("kf_fv_begin", _P), ("kf_fv_index", _P), ("kf_n_nodes", C.c_int),
                ("f_desc", _P), ("f_kps", _P), ("n_f", C.c_int),
                ("f_fv_node", _P), ("f_fv_begin", _P), ("f_fv_index", _P), ("f_n_nodes", C.c_int),
                ("nn_ratio", C.c_float), ("check_orientation", C.c_int)]


class SbpLastArgs(C.Structure):
    _fields_ = [("cur_kps", _P), ("cur_desc", _P), ("cur_u_right", _P), ("n_cur", C.c_int),
                ("cam", Camera), ("scale_factors", _P), ("nlevels", C.c_int),
                ("Rcw", C.c_float * 9), ("tcw", C.c_float * 3), ("Rlw", C.c_float * 9), ("tlw", C.c_float * 3),
                ("last_kps", _P), ("n_last", C.c_int), ("last_has_point", _P), ("last_point_observed", _P),
                ("last_world_pos", _P), ("last_point_desc", _P),
                ("th", C.c_float), ("mono", C.c_int), ("check_orientation", C.c_int), ("cur_occupied", _P)]


class SbpMapArgs(C.Structure):
    _fields_ = [("cur_kps", _P), ("cur_desc", _P), ("cur_u_right", _P), ("n_cur", C.c_int), ("cur_occupied", _P),
                ("cam", Camera), ("scale_factors", _P), ("nlevels", C.c_int),
                ("n_points", C.c_int), ("proj_x", _P), ("proj_y", _P), ("proj_xr", _P), ("pred_level", _P),
                ("view_cos", _P), ("point_observed", _P), ("point_desc", _P), ("th", C.c_float), ("nn_ratio", C.c_float)]


class WindowSearchArgs(C.Structure):
    """olf_window_search_args (include/olf_abi.h): the grid-window search shared by Fuse / SearchBySim3 / SearchByProjection(KF|reloc)."""
    _fields_ = [("kps", _P), ("desc", _P), ("n", C.c_int), ("cam", Camera), ("u_right", _P), ("inv_level_sigma2", _P), ("nlevels", C.c_int),
                ("n_queries", C.c_int), ("u", _P), ("v", _P), ("radius", _P), ("ur", _P), ("min_level", _P), ("max_level", _P), ("qdesc", _P),
                ("blocked", _P), ("sequential_blocking", C.c_int), ("chi2_check", C.c_int), ("max_dist", C.c_int)]


class TriangulationArgs(C.Structure):
    """olf_triangulation_args (include/olf_abi.h): ORBmatcher::SearchForTriangulation."""
    _fields_ = [("kps1", _P), ("desc1", _P), ("n1", C.c_int), ("skip1", _P), ("u_right1", _P),
                ("fv1_node", _P), ("fv1_begin", _P), ("fv1_index", _P), ("fv1_n_nodes", C.c_int),
                ("kps2", _P), ("desc2", _P), ("n2", C.c_int), ("skip2", _P), ("u_right2", _P),
                ("fv2_node", _P), ("fv2_begin", _P), ("fv2_index", _P), ("fv2_n_nodes", C.c_int),
                ("scale_factors2", _P), ("level_sigma2_2", _P), ("nlevels", C.c_int),
                ("F12", C.c_float * 9), ("ex", C.c_float), ("ey", C.c_float), ("only_stereo", C.c_int), ("check_orientation", C.c_int)]


def ptr(a):
    """void* of a contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def as_u8_image(img):
    img = np.ascontiguousarray(img)
    if img.dtype != np.uint8 or img.ndim != 2:
        raise TypeError("image must be a 2-D uint8 array (CV_8UC1)")
    return img


class FrontEndApi:
    """Binds the olf_* (product) or orc_* (oracle) symbol family of a loaded shared library.

    The product library takes a trailing `device` argument on handle constructors and stateless matchers;
    the oracle does not.  Everything else is identical, which is what lets the parity tests call both alike.
    """

    def __init__(self, lib: C.CDLL, prefix: str, device: int | None):
        self.lib, self.prefix, self.device = lib, prefix, device
        self._dev = () if device is None else (C.c_int(device),)

    def fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def check(self, rc, what):
        if rc != OLF_OK:
            msg = ""
            if self.prefix == "olf_":
                self.lib.olf_last_error.restype = C.c_char_p
                msg = (self.lib.olf_last_error() or b"").decode()
            raise RuntimeError(f"{self.prefix}{what} failed with code {rc} {msg}")

    # ---- ORB ----
    def orb_create(self, nfeatures=2000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        f = self.fn("orb_create"); f.restype = C.c_void_p
        h = f(C.c_int(nfeatures), C.c_float(scale_factor), C.c_int(nlevels), C.c_int(ini_th), C.c_int(min_th), *self._dev)
        if not h:
            raise RuntimeError(self.prefix + "orb_create failed")
        return C.c_void_p(h)

    def orb_destroy(self, h):
        f = self.fn("orb_destroy"); f.restype = None; f(h)

    def orb_extract(self, h, img, cap=8192):
        img = as_u8_image(img)
        kps = np.zeros(cap, dtype=KEYPOINT); desc = np.zeros((cap, 32), dtype=np.uint8); n = C.c_int(0)
        rc = self.fn("orb_extract")(h, ptr(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), C.c_int(img.strides[0]),
                                    ptr(kps), ptr(desc), C.c_int(cap), C.byref(n))
        self.check(rc, "orb_extract")
        return kps[:n.value].copy(), desc[:n.value].copy()

    def orb_level(self, h, level):
        w, hh = C.c_int(0), C.c_int(0)
        self.check(self.fn("orb_level_size")(h, C.c_int(level), C.byref(w), C.byref(hh)), "orb_level_size")
        out = np.zeros((hh.value, w.value), dtype=np.uint8)
        self.check(self.fn("orb_get_level")(h, C.c_int(level), ptr(out), C.c_int(w.value)), "orb_get_level")
        return out

    def orb_scale_factors(self, h, nlevels=8):
        s = np.zeros(nlevels, np.float32); i = np.zeros(nlevels, np.float32)
        s2 = np.zeros(nlevels, np.float32); i2 = np.zeros(nlevels, np.float32)
        self.check(self.fn("orb_scale_factors")(h, ptr(s), ptr(i), ptr(s2), ptr(i2)), "orb_scale_factors")
        return s, i, s2, i2

    def orb_features_per_level(self, h, nlevels=8):
        o = np.zeros(nlevels, np.int32)
        self.check(self.fn("orb_features_per_level")(h, ptr(o)), "orb_features_per_level")
        return o

    def orb_last_candidates(self, h, cap=400000):
        o = np.zeros((cap, 4), np.int32); n = C.c_int(0)
        self.check(self.fn("orb_last_candidates")(h, ptr(o), C.c_int(cap), C.byref(n)), "orb_last_candidates")
        return o[:n.value].copy()

    # ---- lines ----
    def line_create(self, params: LineParams | None = None):
        params = params or LineParams()
        f = self.fn("line_create"); f.restype = C.c_void_p
        h = f(C.byref(params), *self._dev)
        if not h:
            raise RuntimeError(self.prefix + "line_create failed")
        return C.c_void_p(h)

    def line_destroy(self, h):
        f = self.fn("line_destroy"); f.restype = None; f(h)

    def lsd_detect(self, h, img, cap=65536):
        img = as_u8_image(img)
        segs = np.zeros((cap, 4), np.float32); n = C.c_int(0)
        rc = self.fn("lsd_detect")(h, ptr(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), C.c_int(img.strides[0]),
                                   ptr(segs), C.c_int(cap), C.byref(n))
        self.check(rc, "lsd_detect")
        return segs[:n.value].copy()

    def line_extract(self, h, img, cap=65536):
        img = as_u8_image(img)
        kls = np.zeros(cap, dtype=KEYLINE); desc = np.zeros((cap, 32), np.uint8); n = C.c_int(0)
        rc = self.fn("line_extract")(h, ptr(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), C.c_int(img.strides[0]),
                                     ptr(kls), ptr(desc), C.c_int(cap), C.byref(n))
        self.check(rc, "line_extract")
        return kls[:n.value].copy(), desc[:n.value].copy()

    def lbd_compute(self, h, img, kls):
        img = as_u8_image(img); kls = np.ascontiguousarray(kls, dtype=KEYLINE)
        desc = np.zeros((len(kls), 32), np.uint8)
        rc = self.fn("lbd_compute")(h, ptr(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), C.c_int(img.strides[0]),
                                    ptr(kls), C.c_int(len(kls)), ptr(desc))
        self.check(rc, "lbd_compute")
        return desc

    # ---- matchers ----
    def knn2_hamming(self, d1, d2):
        d1 = np.ascontiguousarray(d1, np.uint8); d2 = np.ascontiguousarray(d2, np.uint8)
        n1, n2 = len(d1), len(d2)
        o = [np.zeros(n1, np.int32) for _ in range(4)]
        rc = self.fn("knn2_hamming")(ptr(d1), C.c_int(n1), ptr(d2), C.c_int(n2), ptr(o[0]), ptr(o[1]), ptr(o[2]), ptr(o[3]), *self._dev)
        self.check(rc, "knn2_hamming")
        return tuple(o)      # idx0, dist0, idx1, dist1

    def match_nnr(self, d1, d2, nnr):
        d1 = np.ascontiguousarray(d1, np.uint8); d2 = np.ascontiguousarray(d2, np.uint8)
        m = np.zeros(len(d1), np.int32); n = C.c_int(0)
        rc = self.fn("match_nnr")(ptr(d1), C.c_int(len(d1)), ptr(d2), C.c_int(len(d2)), C.c_float(nnr), ptr(m), C.byref(n), *self._dev)
        self.check(rc, "match_nnr")
        return m, n.value

    def match_lines(self, d1, d2, nnr, best_lr=True):
        d1 = np.ascontiguousarray(d1, np.uint8); d2 = np.ascontiguousarray(d2, np.uint8)
        m = np.zeros(len(d1), np.int32); n = C.c_int(0)
        rc = self.fn("match_lines")(ptr(d1), C.c_int(len(d1)), ptr(d2), C.c_int(len(d2)), C.c_float(nnr), C.c_int(int(best_lr)),
                                    ptr(m), C.byref(n), *self._dev)
        self.check(rc, "match_lines")
        return m, n.value

    def stereo_points(self, hl, hr, kl, dl, kr, dr, bf, fx):
        kl = np.ascontiguousarray(kl, KEYPOINT); kr = np.ascontiguousarray(kr, KEYPOINT)
        dl = np.ascontiguousarray(dl, np.uint8); dr = np.ascontiguousarray(dr, np.uint8)
        u = np.zeros(len(kl), np.float32); d = np.zeros(len(kl), np.float32)
        rc = self.fn("stereo_points")(hl, hr, ptr(kl), ptr(dl), C.c_int(len(kl)), ptr(kr), ptr(dr), C.c_int(len(kr)),
                                      C.c_float(bf), C.c_float(fx), ptr(u), ptr(d))
        self.check(rc, "stereo_points")
        return u, d

    def stereo_lines(self, kl, dl, kr, dr, w, h, params: LineMatchParams | None = None):
        params = params or LineMatchParams()
        kl = np.ascontiguousarray(kl, KEYLINE); kr = np.ascontiguousarray(kr, KEYLINE)
        dl = np.ascontiguousarray(dl, np.uint8); dr = np.ascontiguousarray(dr, np.uint8)
        m = np.zeros(len(kl), np.int32); disp = np.zeros((len(kl), 2), np.float32); le = np.zeros((len(kl), 3), np.float64)
        rc = self.fn("stereo_lines")(ptr(kl), ptr(dl), C.c_int(len(kl)), ptr(kr), ptr(dr), C.c_int(len(kr)), C.c_int(w), C.c_int(h),
                                     C.byref(params), ptr(m), ptr(disp), ptr(le), *self._dev)
        self.check(rc, "stereo_lines")
        return m, disp, le

    # ---- whole frame (product library only) ----
    def frame_layout(self, cap_points, cap_lines) -> "FrameOffsets":
        o = FrameOffsets()
        self.check(self.fn("frame_layout")(C.c_int(cap_points), C.c_int(cap_lines), C.byref(o)), "frame_layout")
        return o

    def frontend_create(self, params: "FrontendParams", max_frames: int = 1):
        f = self.fn("frontend_create_batch"); f.restype = C.c_void_p
        h = f(C.byref(params), *self._dev, C.c_int(max_frames))
        if not h:
            self.check(OLF_ERR_INTERNAL, "frontend_create")
        return C.c_void_p(h)

    def frontend_destroy(self, h):
        f = self.fn("frontend_destroy"); f.restype = None; f(h)

    def frontend_process(self, h, img_l, img_r, width, height, stride, on_device, result: np.ndarray):
        """img_l / img_r: numpy uint8 images (host) or integer device addresses (on_device=True)."""
        if on_device:
            pl, pr = C.c_void_p(int(img_l)), C.c_void_p(int(img_r))
        else:
            pl, pr = ptr(img_l), ptr(img_r)
        rc = self.fn("frontend_process")(h, pl, pr, C.c_int(width), C.c_int(height), C.c_int(stride), C.c_int(int(on_device)), ptr(result))
        self.check(rc, "frontend_process")

    def frontend_process_batch(self, h, imgs_l, imgs_r, width, height, stride, on_device, results):
        """imgs_l / imgs_r: lists of numpy uint8 images (host) or of integer device addresses; results: list of result blocks."""
        n = len(imgs_l)
        P = C.c_void_p * n
        if on_device:
            pl, pr = P(*[int(a) for a in imgs_l]), P(*[int(a) for a in imgs_r])
        else:
            pl, pr = P(*[a.ctypes.data for a in imgs_l]), P(*[a.ctypes.data for a in imgs_r])
        pres = P(*[r.ctypes.data for r in results])
        rc = self.fn("frontend_process_batch")(h, pl, pr, C.c_int(n), C.c_int(width), C.c_int(height), C.c_int(stride), C.c_int(int(on_device)), pres)
        self.check(rc, "frontend_process_batch")

    def search_by_projection_last(self, args: SbpLastArgs, keep):
        n_last, n_cur = args.n_last, args.n_cur
        a = np.zeros(n_last, np.int32); c = np.zeros(n_cur, np.int32); n = C.c_int(0)
        rc = self.fn("search_by_projection_last")(C.byref(args), ptr(a), ptr(c), C.byref(n), *self._dev)
        self.check(rc, "search_by_projection_last")
        return a, c, n.value

    # ---- ComputeDistinctiveDescriptors ----
    def distinctive_descriptors(self, desc, group_begin):
        desc = np.ascontiguousarray(desc, np.uint8); group_begin = np.ascontiguousarray(group_begin, np.int32)
        best = np.zeros(len(group_begin) - 1, np.int32)
        rc = self.fn("distinctive_descriptors")(ptr(desc), ptr(group_begin), C.c_int(len(best)), ptr(best), *self._dev)
        self.check(rc, "distinctive_descriptors")
        return best

    # ---- bag of words ----
    def vocab_create(self, tree: dict):
        """tree: dict(k, L, node_desc (n,32) u8, child_begin, child_count, children, word_id (int32), weight (float64))."""
        keep = {k: np.ascontiguousarray(tree[k], dt) for k, dt in (("node_desc", np.uint8), ("child_begin", np.int32), ("child_count", np.int32),
                                                                   ("children", np.int32), ("word_id", np.int32), ("weight", np.float64))}
        d = VocabDesc(int(tree["k"]), int(tree["L"]), len(keep["child_begin"]), ptr(keep["node_desc"]), ptr(keep["child_begin"]),
                      ptr(keep["child_count"]), ptr(keep["children"]), ptr(keep["word_id"]), ptr(keep["weight"]))
        f = self.fn("vocab_create"); f.restype = C.c_void_p
        h = f(C.byref(d), *self._dev)
        if not h:
            self.check(OLF_ERR_INTERNAL, "vocab_create")
        return C.c_void_p(h)

    def vocab_destroy(self, h):
        f = self.fn("vocab_destroy"); f.restype = None; f(h)

    def bow_transform(self, vocab, desc, levelsup=4):
        """Per feature: (word id, word weight, node id `levelsup` levels above the leaves)."""
        desc = np.ascontiguousarray(desc, np.uint8)
        n = len(desc)
        w = np.zeros(n, np.int32); v = np.zeros(n, np.float64); nd = np.zeros(n, np.int32)
        self.check(self.fn("bow_transform")(vocab, ptr(desc), C.c_int(n), C.c_int(levelsup), ptr(w), ptr(v), ptr(nd)), "bow_transform")
        return w, v, nd

    def bow_assemble(self, word_id, weight, node_id):
        """BowVector (words ascending, L1-normalised values) and FeatureVector (CSR over node ids ascending)."""
        n = len(word_id)
        bw = np.zeros(max(n, 1), np.int32); bv = np.zeros(max(n, 1), np.float64); nw = C.c_int(0)
        fn_ = np.zeros(max(n, 1), np.int32); fb = np.zeros(n + 1, np.int32); fi = np.zeros(max(n, 1), np.int32); nn = C.c_int(0)
        rc = self.fn("bow_assemble")(ptr(np.ascontiguousarray(word_id, np.int32)), ptr(np.ascontiguousarray(weight, np.float64)),
                                     ptr(np.ascontiguousarray(node_id, np.int32)), C.c_int(n), ptr(bw), ptr(bv), C.byref(nw), ptr(fn_), ptr(fb), ptr(fi), C.byref(nn))
        self.check(rc, "bow_assemble")
        return bw[:nw.value].copy(), bv[:nw.value].copy(), fn_[:nn.value].copy(), fb[:nn.value + 1].copy(), fi[:fb[nn.value]].copy()

    def search_by_bow(self, kf_desc, kf_kps, kf_has_point, kf_fv, f_desc, f_kps, f_fv, nn_ratio=0.7, check_orientation=True):
        """ORBmatcher::SearchByBoW(KeyFrame*, Frame&): kf_fv / f_fv = (node, begin, index) from bow_assemble; returns (match_f, n)."""
        keep = [np.ascontiguousarray(kf_desc, np.uint8), np.ascontiguousarray(kf_kps), np.ascontiguousarray(kf_has_point, np.uint8),
                np.ascontiguousarray(kf_fv[0], np.int32), np.ascontiguousarray(kf_fv[1], np.int32), np.ascontiguousarray(kf_fv[2], np.int32),
                np.ascontiguousarray(f_desc, np.uint8), np.ascontiguousarray(f_kps),
                np.ascontiguousarray(f_fv[0], np.int32), np.ascontiguousarray(f_fv[1], np.int32), np.ascontiguousarray(f_fv[2], np.int32)]
        a = BowMatchArgs(ptr(keep[0]), ptr(keep[1]), len(keep[0]), ptr(keep[2]), ptr(keep[3]), ptr(keep[4]), ptr(keep[5]), len(keep[3]),
                         ptr(keep[6]), ptr(keep[7]), len(keep[6]), ptr(keep[8]), ptr(keep[9]), ptr(keep[10]), len(keep[8]),
                         nn_ratio, int(check_orientation))
        m = np.zeros(max(len(keep[6]), 1), np.int32); n = C.c_int(0)
        self.check(self.fn("search_by_bow")(C.byref(a), ptr(m), C.byref(n), *self._dev), "search_by_bow")
        return m[:len(keep[6])], n.value

    # ---- the remaining ORBmatcher overloads ----
    def search_by_bow_kf(self, d1, kps1, has1, fv1, d2, kps2, has2, fv2, nn_ratio=0.75, check_orientation=True):
        """ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*): returns (matches12 over KF1 features, n)."""
        keep = [np.ascontiguousarray(d1, np.uint8), np.ascontiguousarray(kps1), np.ascontiguousarray(has1, np.uint8),
                np.ascontiguousarray(fv1[0], np.int32), np.ascontiguousarray(fv1[1], np.int32), np.ascontiguousarray(fv1[2], np.int32),
                np.ascontiguousarray(d2, np.uint8), np.ascontiguousarray(kps2),
                np.ascontiguousarray(fv2[0], np.int32), np.ascontiguousarray(fv2[1], np.int32), np.ascontiguousarray(fv2[2], np.int32),
                np.ascontiguousarray(has2, np.uint8)]
        a = BowMatchArgs(ptr(keep[0]), ptr(keep[1]), len(keep[0]), ptr(keep[2]), ptr(keep[3]), ptr(keep[4]), ptr(keep[5]), len(keep[3]),
                         ptr(keep[6]), ptr(keep[7]), len(keep[6]), ptr(keep[8]), ptr(keep[9]), ptr(keep[10]), len(keep[8]),
                         nn_ratio, int(check_orientation))
        m = np.zeros(max(len(keep[0]), 1), np.int32); n = C.c_int(0)
        self.check(self.fn("search_by_bow_kf")(C.byref(a), ptr(keep[11]), ptr(m), C.byref(n), *self._dev), "search_by_bow_kf")
        return m[:len(keep[0])], n.value

    def window_search(self, kps, desc, cam: Camera, u, v, radius, min_level, max_level, qdesc, max_dist, blocked=None, sequential=False,
                      chi2=None):
        """olf_window_search; chi2 = (u_right, inv_level_sigma2, ur) switches the Fuse reprojection gate on.  Returns (best_idx, best_dist)."""
        f32 = lambda x: np.ascontiguousarray(x, np.float32)
        keep = [np.ascontiguousarray(kps), np.ascontiguousarray(desc, np.uint8), f32(u), f32(v), f32(radius),
                np.ascontiguousarray(min_level, np.int32), np.ascontiguousarray(max_level, np.int32), np.ascontiguousarray(qdesc, np.uint8),
                None if blocked is None else np.ascontiguousarray(blocked, np.uint8)]
        ck = [None, None, None] if chi2 is None else [f32(chi2[0]), f32(chi2[1]), f32(chi2[2])]
        nq = len(keep[2])
        a = WindowSearchArgs(ptr(keep[0]), ptr(keep[1]), len(keep[0]), cam, ptr(ck[0]), ptr(ck[1]), 0 if chi2 is None else len(ck[1]),
                             nq, ptr(keep[2]), ptr(keep[3]), ptr(keep[4]), ptr(ck[2]), ptr(keep[5]), ptr(keep[6]), ptr(keep[7]),
                             ptr(keep[8]), int(sequential), int(chi2 is not None), int(max_dist))
        bi = np.zeros(max(nq, 1), np.int32); bd = np.zeros(max(nq, 1), np.int32)
        self.check(self.fn("window_search")(C.byref(a), ptr(bi), ptr(bd), *self._dev), "window_search")
        return bi[:nq], bd[:nq]

    def search_for_triangulation(self, kf1: dict, kf2: dict, F12, ex, ey, only_stereo=False, check_orientation=False):
        """kf = dict(kps, desc, skip, u_right, fv=(node, begin, index)); kf2 also scale_factors, level_sigma2.  Returns (matches12, n)."""
        f32 = lambda x: np.ascontiguousarray(x, np.float32)
        i32 = lambda x: np.ascontiguousarray(x, np.int32)
        k = [np.ascontiguousarray(kf1["kps"]), np.ascontiguousarray(kf1["desc"], np.uint8), np.ascontiguousarray(kf1["skip"], np.uint8), f32(kf1["u_right"]),
             i32(kf1["fv"][0]), i32(kf1["fv"][1]), i32(kf1["fv"][2]),
             np.ascontiguousarray(kf2["kps"]), np.ascontiguousarray(kf2["desc"], np.uint8), np.ascontiguousarray(kf2["skip"], np.uint8), f32(kf2["u_right"]),
             i32(kf2["fv"][0]), i32(kf2["fv"][1]), i32(kf2["fv"][2]), f32(kf2["scale_factors"]), f32(kf2["level_sigma2"])]
        a = TriangulationArgs(ptr(k[0]), ptr(k[1]), len(k[0]), ptr(k[2]), ptr(k[3]), ptr(k[4]), ptr(k[5]), ptr(k[6]), len(k[4]),
                              ptr(k[7]), ptr(k[8]), len(k[7]), ptr(k[9]), ptr(k[10]), ptr(k[11]), ptr(k[12]), ptr(k[13]), len(k[11]),
                              ptr(k[14]), ptr(k[15]), len(k[14]), (C.c_float * 9)(*[float(x) for x in np.asarray(F12, np.float32).ravel()]),
                              float(np.float32(ex)), float(np.float32(ey)), int(only_stereo), int(check_orientation))
        m = np.zeros(max(len(k[0]), 1), np.int32); n = C.c_int(0)
        self.check(self.fn("search_for_triangulation")(C.byref(a), ptr(m), C.byref(n), *self._dev), "search_for_triangulation")
        return m[:len(k[0])], n.value

    def search_for_initialization(self, kps1, desc1, kps2, desc2, cam: Camera, prev_matched, window_size=100, nn_ratio=0.9, check_orientation=True):
        """ORBmatcher::SearchForInitialization: returns (matches12, n, updated prev_matched [n1, 2])."""
        k1, d1, k2, d2 = np.ascontiguousarray(kps1), np.ascontiguousarray(desc1, np.uint8), np.ascontiguousarray(kps2), np.ascontiguousarray(desc2, np.uint8)
        pm = np.array(prev_matched, np.float32, copy=True).reshape(-1, 2) if len(k1) else np.zeros((0, 2), np.float32)
        pm = np.ascontiguousarray(pm)
        m = np.zeros(max(len(k1), 1), np.int32); n = C.c_int(0)
        rc = self.fn("search_for_initialization")(ptr(k1), ptr(d1), C.c_int(len(k1)), ptr(k2), ptr(d2), C.c_int(len(k2)), C.byref(cam), ptr(pm) if len(k1) else None,
                                                  C.c_int(int(window_size)), C.c_float(nn_ratio), C.c_int(int(check_orientation)), ptr(m), C.byref(n), *self._dev)
        self.check(rc, "search_for_initialization")
        return m[:len(k1)], n.value, pm

    def search_by_projection_map(self, args: SbpMapArgs, keep):
        a = np.zeros(args.n_points, np.int32); n = C.c_int(0)
        rc = self.fn("search_by_projection_map")(C.byref(args), ptr(a), C.byref(n), *self._dev)
        self.check(rc, "search_by_projection_map")
        return a, n.value
