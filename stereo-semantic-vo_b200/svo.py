"""ctypes binding of libsvo_b200.so (include/svo_b200.h) — the host-side mirror used by the
tests and bench.py.  Every call goes through the C ABI; there is no Python/CPU fallback:
loading fails loudly when the library is missing, and every call raises SvoError when the
CUDA path is unavailable.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# SVO_B200_LIB: another build of the same library (A/B measurements of two builds in one run); default: the in-tree one
LIB_PATH = os.environ.get("SVO_B200_LIB") or os.path.join(HERE, "libsvo_b200.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4")])

OK, E_INVALID, E_CUDA, E_CAPACITY, E_NOMEM = 0, -1, -2, -3, -4
DIST_RETAIN_BEST, DIST_OCTREE = 0, 1     # svo_config.distribution
OUT_COMPACT, OUT_NO_RIGHT, OUT_POSE_INPUTS = 1, 2, 4         # svo_set_outputs
TAP_LEVEL, TAP_BLUR, TAP_FAST, TAP_SELECT1, TAP_SELECT2 = 0, 1, 2, 3, 4
PASS1, PASS2 = 0, 1
STAGES = ("total", "h2d", "pyramid", "fast", "select1", "harris", "select2", "blur", "describe",
          "stereo", "match", "d2h", "k_pairs", "k_shortlist2")


class Config(C.Structure):
    _fields_ = [("device", C.c_int), ("width", C.c_int), ("height", C.c_int), ("nfeatures", C.c_int),
                ("nlevels", C.c_int), ("scale_factor", C.c_float), ("fast_threshold", C.c_int),
                ("max_batch", C.c_int), ("lanes", C.c_int), ("max_rows", C.c_int), ("stream", C.c_void_p),
                ("max_channels", C.c_int), ("distribution", C.c_int), ("skip_match_score", C.c_int)]


class Veto(C.Structure):
    _fields_ = [("boxes", C.c_void_p), ("n_boxes", C.c_int), ("F", C.c_void_p),
                ("row_xy", C.c_void_p), ("cur_xy", C.c_void_p)]


class FrameIn(C.Structure):
    _fields_ = [("left", C.c_void_p), ("right", C.c_void_p), ("stride", C.c_int),
                ("bf", C.c_float), ("baseline", C.c_float),
                ("prev_desc", C.c_void_p), ("n_prev", C.c_int), ("prev_live", C.c_void_p),
                ("map_desc", C.c_void_p), ("n_map", C.c_int), ("map_prev_row", C.c_void_p), ("channels", C.c_int),
                ("map_win_uvr", C.c_void_p),
                ("boxes", C.c_void_p), ("n_boxes", C.c_int), ("F", C.c_void_p), ("prev_xy", C.c_void_p),
                ("map_xyz", C.c_void_p), ("map_octave", C.c_void_p), ("Tcw_pred", C.c_void_p),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("proj_th", C.c_float),
                ("track_seq", C.c_int), ("frame_id", C.c_int)]


class FrameOut(C.Structure):
    _fields_ = [("status", C.c_int32), ("n_left", C.c_int32), ("n_right", C.c_int32), ("n_stereo", C.c_int32),
                ("kp_left", C.c_void_p), ("kp_right", C.c_void_p), ("desc_left", C.c_void_p), ("desc_right", C.c_void_p),
                ("u_right", C.c_void_p), ("depth", C.c_void_p),
                ("bf_idx", C.c_void_p), ("bf_dist", C.c_void_p), ("bf_keep", C.c_void_p),
                ("p1_best_idx", C.c_void_p), ("p1_best", C.c_void_p), ("p1_second", C.c_void_p),
                ("p1_row_claimed", C.c_void_p), ("p1_row_bad", C.c_void_p), ("p2_row_claimed", C.c_void_p),
                ("claim_row", C.c_void_p), ("n_prev", C.c_int32), ("n_map", C.c_int32),
                ("mp_create", C.c_void_p), ("mp_xyz", C.c_void_p)]


class TrackView(C.Structure):
    _fields_ = [("n_prev", C.c_int32), ("n_map", C.c_int32), ("last_desc", C.c_void_p), ("prev_desc", C.c_void_p),
                ("prev_live", C.c_void_p), ("prev_map_row", C.c_void_p), ("prev_create", C.c_void_p), ("prev_xyz", C.c_void_p),
                ("prev_xy", C.c_void_p), ("map_desc", C.c_void_p), ("map_create", C.c_void_p), ("map_link", C.c_void_p),
                ("map_xyz", C.c_void_p), ("previous", C.c_int32)]


class PoseProblem(C.Structure):
    _fields_ = [("pts3d", C.c_void_p), ("pts2d", C.c_void_p), ("n", C.c_int),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("Tcw", C.c_float * 16)]


class PnpResult(C.Structure):
    _fields_ = [("R", C.c_double * 9), ("t", C.c_double * 3), ("n_inliers", C.c_int32),
                ("best_iteration", C.c_int32), ("best_solution", C.c_int32), ("n_hypotheses", C.c_int32)]


class SvoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("svo_b200 error %d: %s" % (code, msg))
        self.code = code


EXPORTS = ["svo_default_config", "svo_version", "svo_create", "svo_destroy", "svo_last_error", "svo_get_geometry",
           "svo_extract", "svo_extract_bgr", "svo_stereo_sparse", "svo_match_bf", "svo_match_greedy", "svo_disp2depth",
           "svo_batch_submit", "svo_batch_wait", "svo_batch_result", "svo_alloc_pinned", "svo_free_pinned",
           "svo_alloc_device", "svo_free_device", "svo_copy_to_device", "svo_launch_count", "svo_batch_stage_ms",
           "svo_set_profiling", "svo_lane_stream", "svo_debug_tap", "svo_debug_retain_best",
           "svo_pnp_ransac", "svo_pose_optimize", "svo_debug_hamming_matrix", "svo_debug_tc_profile", "svo_project_map",
           "svo_track_create", "svo_track_reset", "svo_track_state", "svo_track_kp_capacity", "svo_set_outputs",
           "svo_png_info", "svo_png_decode"]

_lib = None


def load():
    """dlopen libsvo_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError("%s not found: build it with `python __graft_entry__.py` or `make -C %s/csrc`" % (LIB_PATH, HERE))
    L = C.CDLL(LIB_PATH)
    L.svo_version.restype = C.c_char_p
    L.svo_last_error.restype = C.c_char_p
    L.svo_last_error.argtypes = [C.c_void_p]
    L.svo_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
    L.svo_destroy.argtypes = [C.c_void_p]
    L.svo_destroy.restype = None
    L.svo_get_geometry.argtypes = [C.c_void_p] + [C.c_void_p] * 4
    L.svo_extract.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    L.svo_extract_bgr.argtypes = L.svo_extract.argtypes
    L.svo_stereo_sparse.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.svo_match_bf.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.svo_match_greedy.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.POINTER(Veto), C.c_void_p]
    L.svo_disp2depth.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_float]
    L.svo_batch_submit.argtypes = [C.c_void_p, C.c_int, C.POINTER(FrameIn), C.c_int]
    L.svo_batch_wait.argtypes = [C.c_void_p, C.c_int]
    L.svo_batch_result.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(FrameOut)]
    L.svo_alloc_pinned.argtypes = [C.c_void_p, C.c_size_t]
    L.svo_alloc_pinned.restype = C.c_void_p
    L.svo_free_pinned.argtypes = [C.c_void_p, C.c_void_p]
    L.svo_free_pinned.restype = None
    L.svo_alloc_device.argtypes = [C.c_void_p, C.c_size_t]
    L.svo_alloc_device.restype = C.c_void_p
    L.svo_free_device.argtypes = [C.c_void_p, C.c_void_p]
    L.svo_free_device.restype = None
    L.svo_copy_to_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    L.svo_launch_count.argtypes = [C.c_void_p]
    L.svo_launch_count.restype = C.c_longlong
    L.svo_batch_stage_ms.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.svo_set_profiling.argtypes = [C.c_void_p, C.c_int]
    L.svo_set_outputs.argtypes = [C.c_void_p, C.c_int]
    L.svo_png_info.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.svo_png_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t]
    L.svo_lane_stream.argtypes = [C.c_void_p, C.c_int]
    L.svo_lane_stream.restype = C.c_void_p
    L.svo_debug_tap.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    L.svo_debug_tap.restype = C.c_longlong
    L.svo_debug_retain_best.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.svo_pnp_ransac.argtypes = [C.c_void_p, C.POINTER(PoseProblem), C.c_int, C.c_int, C.c_float, C.c_uint32, C.c_int,
                                 C.POINTER(PnpResult), C.c_void_p]
    L.svo_pose_optimize.argtypes = [C.c_void_p, C.POINTER(PoseProblem), C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.svo_project_map.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]
    L.svo_debug_tc_profile.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.svo_debug_hamming_matrix.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    L.svo_track_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.svo_track_reset.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.svo_track_state.argtypes = [C.c_void_p, C.c_int, C.POINTER(TrackView)]
    L.svo_track_kp_capacity.argtypes = [C.c_void_p]
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def png_decode(data, out=None):
    """Input staging (host side, no GPU needed): PNG file bytes -> array in cv2.imread(IMREAD_UNCHANGED)'s layout
    ((h, w) u8 / u16 or (h, w, 3|4) u8 in BGR(A) order).  `out`: optional destination (e.g. a view of pinned memory)."""
    L = load()
    buf = np.frombuffer(data, np.uint8)
    w, h, ch, bd = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    rc = L.svo_png_info(_p(buf), buf.size, C.byref(w), C.byref(h), C.byref(ch), C.byref(bd))
    if rc < 0:
        raise SvoError(rc, "svo_png_info: not a PNG this decoder reads (8-bit gray/RGB/RGBA or 16-bit gray, non-interlaced)")
    shape = (h.value, w.value) if ch.value == 1 else (h.value, w.value, ch.value)
    dt = np.uint16 if bd.value == 16 else np.uint8
    if out is None:
        out = np.empty(shape, dt)
    assert out.shape == shape and out.dtype == dt and out.strides[-1] == out.itemsize
    rc = L.svo_png_decode(_p(buf), buf.size, _p(out), out.strides[0], out.strides[0] * h.value)
    if rc < 0:
        raise SvoError(rc, "svo_png_decode failed")
    return out


def _view(ptr, dtype, shape):
    n = int(np.prod(shape))
    if not ptr or n == 0:
        return np.zeros(shape, dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)


class Context:
    """One svo_ctx: device buffers, streams and pipeline lanes for one image size."""

    def __init__(self, width=1241, height=376, nfeatures=2000, nlevels=8, scale_factor=1.2, fast_threshold=20,
                 max_batch=1, lanes=1, max_rows=5000, device=0, stream=None, max_channels=1, distribution=0,
                 skip_match_score=False):
        self.lib = load()
        cfg = Config()
        self.lib.svo_default_config(C.byref(cfg))
        cfg.device, cfg.width, cfg.height, cfg.nfeatures = device, width, height, nfeatures
        cfg.nlevels, cfg.scale_factor, cfg.fast_threshold = nlevels, scale_factor, fast_threshold
        cfg.max_batch, cfg.lanes, cfg.max_rows, cfg.stream = max_batch, lanes, max_rows, stream
        cfg.max_channels = max_channels
        cfg.distribution = distribution      # DIST_RETAIN_BEST (cv::ORB parity) or DIST_OCTREE (opt-in, non-parity)
        cfg.skip_match_score = int(bool(skip_match_score))
        self.cfg = cfg
        self.h = C.c_void_p()
        rc = self.lib.svo_create(C.byref(cfg), C.byref(self.h))
        if rc != OK:
            msg = self.lib.svo_last_error(self.h).decode() if self.h else "svo_create failed"
            if self.h:
                self.lib.svo_destroy(self.h)
                self.h = None
            raise SvoError(rc, msg)
        self.width, self.height, self.nfeatures, self.max_rows, self.max_batch = width, height, nfeatures, max_rows, max_batch
        self._keep = {}

    def close(self):
        if getattr(self, "h", None):
            self.lib.svo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc < 0:
            raise SvoError(rc, self.lib.svo_last_error(self.h).decode())
        return rc

    def geometry(self):
        lw = np.zeros(8, np.int32); lh = np.zeros(8, np.int32); ls = np.zeros(8, np.float32); q = np.zeros(8, np.int32)
        n = self._chk(self.lib.svo_get_geometry(self.h, _p(lw), _p(lh), _p(ls), _p(q)))
        return lw[:n], lh[:n], ls[:n], q[:n]

    # ---- synchronous drop-ins -------------------------------------------------------
    def extract(self, gray, cam=0, cap=None):
        """frame::featuredetect -> (keypoints[KP_DTYPE], descriptors[n,32]).  (h, w) gray or (h, w, 3) BGR."""
        gray = np.asarray(gray, np.uint8)
        if gray.strides[-1] != 1 or (gray.ndim == 3 and gray.strides[1] != 3):
            gray = np.ascontiguousarray(gray)
        h, w = gray.shape[:2]
        cap = cap or (self.nfeatures * 2 + 1024)
        kp = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8)
        fn = self.lib.svo_extract_bgr if gray.ndim == 3 else self.lib.svo_extract
        n = self._chk(fn(self.h, cam, _p(gray), gray.strides[0], w, h, _p(kp), _p(desc), cap))
        n = min(n, cap)
        return kp[:n].copy(), desc[:n].copy()

    def stereo_sparse(self, bf, baseline, cap=None):
        cap = cap or (self.nfeatures * 2 + 1024)
        ur = np.zeros(cap, np.float32); dep = np.zeros(cap, np.float32)
        mr = np.zeros(cap, np.int32); sad = np.zeros(cap, np.int32)
        n = self._chk(self.lib.svo_stereo_sparse(self.h, bf, baseline, _p(ur), _p(dep), _p(mr), _p(sad), cap))
        n = min(n, cap)
        return ur[:n].copy(), dep[:n].copy(), mr[:n].copy(), sad[:n].copy()

    def match_bf(self, q, t):
        q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32); t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
        idx = np.zeros(len(q), np.int32); dist = np.zeros(len(q), np.int32); keep = np.zeros(len(q), np.uint8)
        self._chk(self.lib.svo_match_bf(self.h, _p(q), len(q), _p(t), len(t), _p(idx), _p(dist), _p(keep)))
        return idx, dist, keep

    def match_greedy(self, rows, cur, mode, claimed=None, row_live=None, row_base=0, claim_row=None,
                     win_uvr=None, cur_xy=None, veto=None, scores=True):
        rows = np.ascontiguousarray(rows, np.uint8).reshape(-1, 32); cur = np.ascontiguousarray(cur, np.uint8).reshape(-1, 32)
        M, N = len(rows), len(cur)
        claimed = np.zeros(N, np.uint8) if claimed is None else np.ascontiguousarray(claimed, np.uint8).copy()
        claim_row = np.full(N, -1, np.int32) if claim_row is None else np.ascontiguousarray(claim_row, np.int32).copy()
        if row_live is not None:
            row_live = np.ascontiguousarray(row_live, np.uint8)
        if win_uvr is not None:
            win_uvr = np.ascontiguousarray(win_uvr, np.float32)
        if cur_xy is not None:
            cur_xy = np.ascontiguousarray(cur_xy, np.float32)
        bi = np.full(M, -1, np.int32); b = np.full(M, 256, np.int32); s = np.full(M, 256, np.int32)
        rc = np.zeros(M, np.uint8); bad = np.zeros(M, np.uint8)
        v = None
        if veto is not None:
            boxes = np.ascontiguousarray(veto["boxes"], np.int32).reshape(-1, 4)
            F = np.ascontiguousarray(veto["F"], np.float64).reshape(9)
            rxy = np.ascontiguousarray(veto["row_xy"], np.float32); cxy = np.ascontiguousarray(veto["cur_xy"], np.float32)
            v = Veto(_p(boxes), len(boxes), _p(F), _p(rxy), _p(cxy))
            self._keep["veto"] = (boxes, F, rxy, cxy)
        self._chk(self.lib.svo_match_greedy(self.h, _p(rows), M, _p(cur), N, mode, _p(row_live), _p(claimed), _p(claim_row),
                                            row_base, _p(bi) if scores else None, _p(b) if scores else None,
                                            _p(s) if scores else None, _p(rc), _p(win_uvr), _p(cur_xy),
                                            C.byref(v) if v is not None else None, _p(bad)))
        return dict(best_idx=bi, best=b, second=s, row_claimed=rc, claimed=claimed, claim_row=claim_row, row_bad=bad)

    def project_map(self, xyz, octave, Tcw, K4, th=7.0):
        """Opt-in projection windows (u, v, r) of map points under the predicted pose."""
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        octave = None if octave is None else np.ascontiguousarray(octave, np.int32)
        T = np.ascontiguousarray(Tcw, np.float32).reshape(16)
        out = np.zeros((len(xyz), 3), np.float32)
        self._chk(self.lib.svo_project_map(self.h, _p(xyz), _p(octave), len(xyz), _p(T), float(K4[0]), float(K4[1]), float(K4[2]),
                                           float(K4[3]), float(th), _p(out)))
        return out

    def disp2depth(self, disp, bf):
        disp = np.ascontiguousarray(disp, np.float32)
        out = np.empty_like(disp)
        self._chk(self.lib.svo_disp2depth(self.h, _p(disp), _p(out), disp.size, bf))
        return out

    # ---- pose stage -------------------------------------------------------------------
    def _pose_problems(self, problems):
        """problems: list of dicts with pts3d (n,3), pts2d (n,2), K=(fx,fy,cx,cy) and optional Tcw (4,4)."""
        arr = (PoseProblem * len(problems))()
        keep = []
        for i, q in enumerate(problems):
            p3 = np.ascontiguousarray(q["pts3d"], np.float32).reshape(-1, 3)
            p2 = np.ascontiguousarray(q["pts2d"], np.float32).reshape(-1, 2)
            assert len(p3) == len(p2)
            keep += [p3, p2]
            arr[i].pts3d, arr[i].pts2d, arr[i].n = p3.ctypes.data, p2.ctypes.data, len(p3)
            arr[i].fx, arr[i].fy, arr[i].cx, arr[i].cy = [float(v) for v in q["K"]]
            T = np.ascontiguousarray(q.get("Tcw", np.eye(4)), np.float32).reshape(16)
            for k in range(16):
                arr[i].Tcw[k] = float(T[k])
        return arr, keep

    def pnp_ransac(self, problems, iterations=100, reproj_err=8.0, seed=1, refine_iters=10):
        """cv::solvePnPRansac's role (src/pnpmatch.cc:227) -> list of dicts (n_inliers, R, t, inliers, info)."""
        arr, keep = self._pose_problems(problems)
        res = (PnpResult * len(problems))()
        total = sum(a.n for a in arr)
        mask = np.zeros(max(total, 1), np.uint8)
        self._chk(self.lib.svo_pnp_ransac(self.h, arr, len(problems), iterations, reproj_err, seed, refine_iters, res, _p(mask)))
        out, off = [], 0
        for i in range(len(problems)):
            r = res[i]
            out.append(dict(n_inliers=r.n_inliers, R=np.array(r.R[:]).reshape(3, 3), t=np.array(r.t[:]),
                            inliers=mask[off:off + arr[i].n].copy(),
                            info=(r.best_iteration, r.best_solution, r.n_hypotheses)))
            off += arr[i].n
        return out

    def pose_optimize(self, problems, iterations=10):
        """Optimizer::PoseOptimization (src/Optimizer.cc:15-86) -> list of (Tcw[4,4] f32, iterations run, robust chi2)."""
        arr, keep = self._pose_problems(problems)
        T = np.zeros((len(problems), 4, 4), np.float32); st = np.zeros((len(problems), 2), np.float64)
        self._chk(self.lib.svo_pose_optimize(self.h, arr, len(problems), iterations, _p(T), _p(st)))
        return [(T[i].copy(), int(st[i, 0]), float(st[i, 1])) for i in range(len(problems))]

    # ---- batched pipeline ------------------------------------------------------------
    def batch_submit(self, lane, frames):
        """frames: list of dicts with left,right (u8 arrays or raw pointers + stride), bf, baseline and optional
        prev_desc, prev_live, map_desc, map_prev_row, map_win_uvr and the pass-1 veto inputs boxes (n,4 int32),
        F (3,3 f64), prev_xy (n_prev,2 f32).  Arrays are kept alive until the lane's next submit."""
        arr = (FrameIn * len(frames))()
        keep = []
        for i, f in enumerate(frames):
            fi = arr[i]
            for side in ("left", "right"):
                v = f[side]
                if isinstance(v, np.ndarray):
                    keep.append(v)
                    setattr(fi, side, v.ctypes.data)
                    fi.stride = v.strides[0]
                    fi.channels = 3 if v.ndim == 3 else 1
                else:
                    setattr(fi, side, v)
                    fi.stride = f["stride"]
                    fi.channels = f.get("channels", 1)
            fi.bf, fi.baseline = f["bf"], f["baseline"]
            for name, cnt in (("prev_desc", "n_prev"), ("map_desc", "n_map")):
                v = f.get(name)
                if v is None:
                    continue
                if isinstance(v, np.ndarray):
                    keep.append(v)
                    setattr(fi, name, v.ctypes.data); setattr(fi, cnt, len(v))
                else:
                    setattr(fi, name, v); setattr(fi, cnt, f[cnt])
            for name in ("prev_live", "map_prev_row", "map_win_uvr"):
                v = f.get(name)
                if isinstance(v, np.ndarray):
                    if name == "map_win_uvr":
                        v = np.ascontiguousarray(v, np.float32)
                    keep.append(v)
                    setattr(fi, name, v.ctypes.data)
                elif v is not None:
                    setattr(fi, name, int(v))          # raw (device) address
            if f.get("boxes") is not None:
                v = f["boxes"]
                if isinstance(v, np.ndarray) or isinstance(v, (list, tuple)):
                    v = np.ascontiguousarray(v, np.int32).reshape(-1, 4)
                    keep.append(v)
                    fi.boxes, fi.n_boxes = v.ctypes.data, len(v)
                else:
                    fi.boxes, fi.n_boxes = int(v), f["n_boxes"]
            if f.get("track_seq") is not None:          # device-resident tracker state: sequence index (0-based here)
                fi.track_seq, fi.frame_id = int(f["track_seq"]) + 1, int(f.get("frame_id", 0))
            if f.get("K") is not None:
                fi.fx, fi.fy, fi.cx, fi.cy = [float(v) for v in f["K"]]
                fi.proj_th = float(f.get("proj_th", 7.0))
            for name, dt in (("F", np.float64), ("prev_xy", np.float32), ("map_xyz", np.float32), ("map_octave", np.int32),
                             ("Tcw_pred", np.float32)):
                v = f.get(name)
                if isinstance(v, np.ndarray):
                    v = np.ascontiguousarray(v, dt)
                    keep.append(v)
                    setattr(fi, name, v.ctypes.data)
                elif v is not None:
                    setattr(fi, name, int(v))
        self._keep[("lane", lane)] = (arr, keep)
        self._chk(self.lib.svo_batch_submit(self.h, lane, arr, len(frames)))

    def batch_wait(self, lane):
        self._chk(self.lib.svo_batch_wait(self.h, lane))

    def batch_result(self, lane, i, copy=True):
        o = FrameOut()
        self._chk(self.lib.svo_batch_result(self.h, lane, i, C.byref(o)))
        arr, _ = self._keep[("lane", lane)]
        tracked = arr[i].track_seq != 0
        n_prev, n_map = (o.n_prev, o.n_map) if tracked else (arr[i].n_prev, arr[i].n_map)
        nl, nr = o.n_left, o.n_right
        r = dict(status=o.status, n_left=nl, n_right=nr, n_stereo=o.n_stereo,
                 kp_left=_view(o.kp_left, KP_DTYPE, (nl,)), kp_right=_view(o.kp_right, KP_DTYPE, (nr,)) if o.kp_right else None,
                 desc_left=_view(o.desc_left, np.uint8, (nl, 32)) if o.desc_left else None, desc_right=_view(o.desc_right, np.uint8, (nr, 32)) if o.desc_right else None,
                 u_right=_view(o.u_right, np.float32, (nl,)) if o.u_right else None, depth=_view(o.depth, np.float32, (nl,)),
                 claim_row=_view(o.claim_row, np.int32, (nl,)), n_prev=n_prev, n_map=n_map)
        if tracked:
            r.update(mp_create=_view(o.mp_create, np.int32, (nl,)), mp_xyz=_view(o.mp_xyz, np.float32, (nl, 3)))
        if n_prev and o.bf_idx:
            r.update(bf_idx=_view(o.bf_idx, np.int32, (nl,)), bf_dist=_view(o.bf_dist, np.int32, (nl,)),
                     bf_keep=_view(o.bf_keep, np.uint8, (nl,)),
                     p1_best_idx=_view(o.p1_best_idx, np.int32, (n_prev,)), p1_best=_view(o.p1_best, np.int32, (n_prev,)),
                     p1_second=_view(o.p1_second, np.int32, (n_prev,)),
                     p1_row_claimed=_view(o.p1_row_claimed, np.uint8, (n_prev,)),
                     p1_row_bad=_view(o.p1_row_bad, np.uint8, (n_prev,)))
        if n_map and o.p2_row_claimed:
            r.update(p2_row_claimed=_view(o.p2_row_claimed, np.uint8, (n_map,)))
        if copy:
            r = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in r.items()}
        return r

    # ---- device-resident tracker state (opt-in) ------------------------------------------
    def track_create(self, n_sequences, map_capacity=None, window=4):
        self._chk(self.lib.svo_track_create(self.h, n_sequences, map_capacity or self.max_rows, window))
        self.track_cap = map_capacity or self.max_rows

    def track_reset(self, seq, ballast=None):
        b = None if ballast is None else np.ascontiguousarray(ballast, np.uint8).reshape(-1, 32)
        self._chk(self.lib.svo_track_reset(self.h, seq, _p(b), 0 if b is None else len(b)))

    def track_state(self, seq, previous=False):
        """The sequence's current state as host arrays (test tap); previous=True: the state its last frame read."""
        K = self._chk(self.lib.svo_track_kp_capacity(self.h)); Cp = self.track_cap
        a = dict(last_desc=np.zeros((K, 32), np.uint8), prev_desc=np.zeros((K, 32), np.uint8), prev_live=np.zeros(K, np.uint8),
                 prev_map_row=np.zeros(K, np.int32), prev_create=np.zeros(K, np.int32), prev_xyz=np.zeros((K, 3), np.float32),
                 prev_xy=np.zeros((K, 2), np.float32), map_desc=np.zeros((Cp, 32), np.uint8), map_create=np.zeros(Cp, np.int32),
                 map_link=np.zeros(Cp, np.int32), map_xyz=np.zeros((Cp, 3), np.float32))
        v = TrackView()
        v.previous = int(bool(previous))
        for k, arr in a.items():
            setattr(v, k, arr.ctypes.data)
        self._chk(self.lib.svo_track_state(self.h, seq, C.byref(v)))
        out = {k: (arr[:v.n_map] if k.startswith("map_") else arr[:v.n_prev]) for k, arr in a.items()}
        out["n_prev"], out["n_map"] = v.n_prev, v.n_map
        return out

    def stage_ms(self, lane):
        ms = np.zeros(14, np.float32)
        self._chk(self.lib.svo_batch_stage_ms(self.h, lane, _p(ms), 14))
        return dict(zip(STAGES, ms.tolist()))

    def set_outputs(self, flags):
        """OUT_COMPACT | OUT_NO_RIGHT: what the batches submitted from now on copy back (0: everything at capacity)."""
        self._chk(self.lib.svo_set_outputs(self.h, int(flags)))

    def set_profiling(self, on):
        self._chk(self.lib.svo_set_profiling(self.h, int(bool(on))))

    def launch_count(self):
        return int(self.lib.svo_launch_count(self.h))

    def lane_stream(self, lane):
        return self.lib.svo_lane_stream(self.h, lane)

    def alloc_pinned(self, nbytes):
        p = self.lib.svo_alloc_pinned(self.h, nbytes)
        if not p:
            raise SvoError(E_NOMEM, "svo_alloc_pinned(%d)" % nbytes)
        return p

    def pinned_array(self, shape, dtype=np.uint8):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = self.alloc_pinned(n)
        return _view(p, dtype, shape)

    def alloc_device(self, nbytes):
        p = self.lib.svo_alloc_device(self.h, nbytes)
        if not p:
            raise SvoError(E_NOMEM, "svo_alloc_device(%d)" % nbytes)
        return p

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr)
        p = self.alloc_device(arr.nbytes)
        self._chk(self.lib.svo_copy_to_device(self.h, p, _p(arr), arr.nbytes))
        return p

    # ---- taps -------------------------------------------------------------------------
    def tap_image(self, cam, level, blurred=False):
        lw, lh, _, _ = self.geometry()
        out = np.zeros((int(lh[level]), int(lw[level])), np.uint8)
        self._chk(self.lib.svo_debug_tap(self.h, cam, TAP_BLUR if blurred else TAP_LEVEL, level, _p(out), out.nbytes))
        return out

    def tap_list(self, cam, what, level, cap=400000):
        out = np.zeros((cap, 3), np.int32)
        n = self._chk(self.lib.svo_debug_tap(self.h, cam, what, level, _p(out), out.nbytes))
        return out[:min(n, cap)].copy()

    def hamming_matrix(self, a, b):
        """Full distance matrix through the tensor-core tiles of the batch matchers (test tap)."""
        a = np.ascontiguousarray(a, np.uint8).reshape(-1, 32); b = np.ascontiguousarray(b, np.uint8).reshape(-1, 32)
        out = np.zeros((len(a), len(b)), np.int32)
        self._chk(self.lib.svo_debug_hamming_matrix(self.h, _p(a), len(a), _p(b), len(b), _p(out)))
        return out

    def tc_profile(self):
        st = np.zeros(1024, np.int64)
        self._chk(self.lib.svo_debug_tc_profile(self.h, _p(st), 1024))
        return st.reshape(4, 4, 64)

    def retain_best(self, resp, n_points, depth_limit=-1):
        resp = np.ascontiguousarray(resp, np.float32)
        idx = np.zeros(max(len(resp), 1), np.int32)
        k = self._chk(self.lib.svo_debug_retain_best(self.h, _p(resp), len(resp), n_points, depth_limit, _p(idx)))
        return idx[:k].copy()
