"""ctypes mirror of include/smb.h -- the only Python route into the CUDA library.

``SiftMatcher`` plays the role of ``colmap::MatchSiftFeaturesCPU`` as called by the reference op
(``/root/reference/integration/op_cpp/sequential_matching.cc:154``) plus the descriptor cache
that replaces its per-row ``read_matrix_from_element`` (``io.cc:181-194``).  Option names and
defaults are those of ``siftFeatureMatchingArgs`` (``colmap.proto:14-24``).

There is no CPU path: if ``libsmb.so`` is missing or no sm_100 device is present every compute
entry point raises ``SmbError``.
"""
from __future__ import annotations

import ctypes
import os
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsmb.so")
TEST_LIB_PATH = os.path.join(_HERE, "libsmb_test.so")   # + the CUDA-core cross-check engine; loaded by tests only
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "smb.h")

SMB_OK, SMB_EINVAL, SMB_ECUDA, SMB_ENOMEM, SMB_ENODEVICE, SMB_ECAPACITY = 0, -1, -2, -3, -4, -5
ENGINE_TCGEN05, ENGINE_DP4A = 0, 1
_ENGINES = {"tcgen05": ENGINE_TCGEN05, "dp4a": ENGINE_DP4A}


class SmbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"smb error {code}: {msg}")
        self.code = code


class smb_options(ctypes.Structure):
    _fields_ = [("max_ratio", ctypes.c_double), ("max_distance", ctypes.c_double),
                ("cross_check", ctypes.c_int32), ("max_num_matches", ctypes.c_int32),
                ("engine", ctypes.c_int32), ("profile", ctypes.c_int32)]


class smb_timing(ctypes.Structure):
    _fields_ = [("total_ms", ctypes.c_float), ("score_ms", ctypes.c_float), ("runner_up_ms", ctypes.c_float),
                ("decide_ms", ctypes.c_float), ("score_launches", ctypes.c_uint32), ("total_launches", ctypes.c_uint32),
                ("sub_batches", ctypes.c_uint32), ("plan_uploaded", ctypes.c_uint32),
                ("candidates", ctypes.c_uint64), ("ops", ctypes.c_uint64), ("cta_busy_max_over_mean", ctypes.c_float),
                ("pad_", ctypes.c_float)]


SMB_TVG_NO_WATERMARK, SMB_TVG_MULTIPLE_MODELS = 1, 2   # include/smb.h


class smb_tvg_options(ctypes.Structure):
    _fields_ = [("min_num_inliers", ctypes.c_int32), ("min_num_trials", ctypes.c_int32), ("max_num_trials", ctypes.c_int32),
                ("flags", ctypes.c_int32), ("max_error", ctypes.c_double), ("confidence", ctypes.c_double),
                ("min_inlier_ratio", ctypes.c_double), ("max_h_inlier_ratio", ctypes.c_double), ("seed", ctypes.c_uint64)]


class smb_tvg(ctypes.Structure):
    _fields_ = [("config", ctypes.c_int32), ("num_inliers_f", ctypes.c_int32), ("num_inliers_h", ctypes.c_int32),
                ("trials_f", ctypes.c_int32), ("trials_h", ctypes.c_int32), ("inlier_start", ctypes.c_uint32),
                ("inlier_count", ctypes.c_uint32), ("pad_", ctypes.c_uint32), ("F", ctypes.c_double * 9),
                ("H", ctypes.c_double * 9)]


class smb_match(ctypes.Structure):
    _fields_ = [("idx1", ctypes.c_uint32), ("idx2", ctypes.c_uint32)]


_libs = {}


def load_library(path: Optional[str] = None) -> ctypes.CDLL:
    """dlopen libsmb.so and declare every prototype of include/smb.h.  Raises if the library is absent:
    the product path must fail loudly rather than fall back."""
    p = path or os.environ.get("SMB_LIB") or LIB_PATH   # SMB_LIB: A/B a build variant (tools/bin/*.so) under the tests
    p = os.path.abspath(p)
    if p in _libs:
        return _libs[p]
    if not os.path.exists(p):
        raise SmbError(SMB_ENODEVICE, f"{p} not built -- run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = ctypes.CDLL(p)
    vp, szp = ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t)
    L.smb_default_options.argtypes = [ctypes.POINTER(smb_options)]
    L.smb_default_options.restype = None
    L.smb_abi_version.restype = ctypes.c_int
    L.smb_create.argtypes = [ctypes.c_int, ctypes.POINTER(smb_options), ctypes.POINTER(vp)]
    L.smb_destroy.argtypes = [vp]
    L.smb_destroy.restype = None
    L.smb_last_error.argtypes = [vp]
    L.smb_last_error.restype = ctypes.c_char_p
    L.smb_set_options.argtypes = [vp, ctypes.POINTER(smb_options)]
    L.smb_put_image.argtypes = [vp, ctypes.c_uint32, vp, ctypes.c_size_t, ctypes.c_size_t]
    L.smb_put_images.argtypes = [vp, vp, vp, vp, ctypes.c_size_t, ctypes.c_size_t]
    L.smb_put_images_async.argtypes = [vp, vp, vp, vp, ctypes.c_size_t, ctypes.c_size_t]
    L.smb_put_image_device.argtypes = [vp, ctypes.c_uint32, vp, ctypes.c_size_t, ctypes.c_size_t]
    L.smb_put_images_device.argtypes = [vp, vp, vp, vp, ctypes.c_size_t, ctypes.c_size_t]
    L.smb_put_images_device_async.argtypes = [vp, vp, vp, vp, ctypes.c_size_t, ctypes.c_size_t, vp]
    L.smb_has_image.argtypes = [vp, ctypes.c_uint32]
    L.smb_evict_image.argtypes = [vp, ctypes.c_uint32]
    L.smb_clear_images.argtypes = [vp]
    L.smb_image_device_ptr.argtypes = [vp, ctypes.c_uint32, ctypes.POINTER(vp), szp]
    L.smb_match_pairs.argtypes = [vp, vp, ctypes.c_size_t, ctypes.POINTER(vp)]
    L.smb_match_pairs_begin.argtypes = [vp, vp, ctypes.c_size_t, ctypes.POINTER(vp)]
    L.smb_result_wait.argtypes = [vp, vp]
    L.smb_result_num_pairs.argtypes = [vp]
    L.smb_result_num_pairs.restype = ctypes.c_size_t
    L.smb_result_matches.argtypes = [vp, ctypes.c_size_t, szp]
    L.smb_result_matches.restype = vp
    L.smb_result_total_matches.argtypes = [vp]
    L.smb_result_total_matches.restype = ctypes.c_size_t
    L.smb_result_release.argtypes = [vp, vp]
    L.smb_result_release.restype = None
    L.smb_match_descriptors.argtypes = [vp, vp, ctypes.c_size_t, vp, ctypes.c_size_t, vp, ctypes.c_size_t, szp]
    L.smb_get_timing.argtypes = [vp, ctypes.POINTER(smb_timing)]
    L.smb_get_filter.argtypes = [vp, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]
    L.smb_stream.argtypes = [vp]
    L.smb_stream.restype = vp
    L.smb_stream_wait_uploads.argtypes = [vp, vp]
    L.smb_wait_stream.argtypes = [vp, vp]
    L.smb_synchronize.argtypes = [vp]
    L.smb_default_tvg_options.argtypes = [ctypes.POINTER(smb_tvg_options)]
    L.smb_default_tvg_options.restype = None
    L.smb_put_keypoints.argtypes = [vp, ctypes.c_uint32, vp, ctypes.c_size_t, ctypes.c_size_t]
    L.smb_result_verify.argtypes = [vp, vp, ctypes.POINTER(smb_tvg_options)]
    L.smb_result_tvg.argtypes = [vp, ctypes.c_size_t, ctypes.POINTER(smb_tvg)]
    L.smb_result_inliers.argtypes = [vp, ctypes.c_size_t, szp]
    L.smb_result_inliers.restype = vp
    L.smb_alloc_pinned.argtypes = [ctypes.c_size_t, ctypes.POINTER(vp)]
    L.smb_free_pinned.argtypes = [vp]
    L.smb_free_pinned.restype = None
    if L.smb_abi_version() != 2:
        raise SmbError(SMB_EINVAL, f"{p}: ABI version {L.smb_abi_version()}, this mirror is written for 2 -- rebuild")
    _libs[p] = L
    return L


def _desc(d) -> np.ndarray:
    d = np.ascontiguousarray(d, dtype=np.uint8)
    if d.ndim != 2 or d.shape[1] != 128:
        raise ValueError("descriptors must be [n, 128] uint8 (FeatureDescriptors, io.cc:181-194)")
    return d


class SiftMatcher:
    """One matcher per GPU (one Scanner kernel instance owns one).

    Parameters mirror ``siftFeatureMatchingArgs``: ``max_ratio=0.8, max_distance=0.7,
    cross_check=True, max_num_matches=32768`` (accepted, not applied -- as on the reference CPU path).
    """

    def __init__(self, device: int = 0, max_ratio: float = 0.8, max_distance: float = 0.7, cross_check: bool = True,
                 max_num_matches: int = 32768, engine: str = "tcgen05", profile: bool = False):
        # the product library has one engine; the CUDA-core cross-check lives in the tests' build of the same sources
        self._L = load_library(TEST_LIB_PATH if engine != "tcgen05" and not os.environ.get("SMB_LIB") else None)
        self._h = ctypes.c_void_p()
        self._keepalive: list = []
        self._opts = smb_options(float(max_ratio), float(max_distance), int(bool(cross_check)), int(max_num_matches),
                                 _ENGINES[engine], int(bool(profile)))
        rc = self._L.smb_create(int(device), ctypes.byref(self._opts), ctypes.byref(self._h))
        if rc != SMB_OK:
            raise SmbError(rc, (self._L.smb_last_error(None) or b"").decode())
        self.device = int(device)

    # -- lifetime ------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._L.smb_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int) -> None:
        if rc != SMB_OK:
            raise SmbError(rc, (self._L.smb_last_error(self._h) or b"").decode())

    # -- options -------------------------------------------------------------------------------
    def set_options(self, **kw) -> None:
        for k, v in kw.items():
            if k == "engine":
                v = _ENGINES[v]
            if not hasattr(self._opts, k):
                raise TypeError(f"unknown option {k}")
            setattr(self._opts, k, type(getattr(self._opts, k))(v))
        self._check(self._L.smb_set_options(self._h, ctypes.byref(self._opts)))

    @property
    def filter(self) -> Tuple[int, int]:
        """(min_score, min_best): the integer pre-filter derived from the options."""
        a, b = ctypes.c_int32(), ctypes.c_int32()
        self._check(self._L.smb_get_filter(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    # -- descriptor cache ----------------------------------------------------------------------
    def put_image(self, image_id: int, descriptors) -> None:
        d = _desc(descriptors)
        self._check(self._L.smb_put_image(self._h, int(image_id), d.ctypes.data, d.shape[0], 128))

    def put_images(self, image_ids: Sequence[int], descriptors: Sequence[np.ndarray]) -> None:
        ds = [_desc(d) for d in descriptors]
        n = len(ds)
        ids = np.asarray(list(image_ids), dtype=np.uint32)
        ptrs = (ctypes.c_void_p * n)(*[d.ctypes.data for d in ds])
        ns = (ctypes.c_size_t * n)(*[d.shape[0] for d in ds])
        self._check(self._L.smb_put_images(self._h, ids.ctypes.data, ctypes.cast(ptrs, ctypes.c_void_p),
                                           ctypes.cast(ns, ctypes.c_void_p), n, 128))

    def put_images_async(self, image_ids: Sequence[int], descriptors: Sequence[np.ndarray]) -> None:
        """Queue the uploads and return at once.  ``descriptors`` must stay alive and unchanged (pinned memory for
        real overlap) until ``synchronize()`` or until a match call naming these images has returned."""
        ds = [_desc(d) for d in descriptors]
        n = len(ds)
        if len(self._keepalive) > 4096:      # a caller that never synchronises: drain, then the references can go
            self.synchronize()
        self._keepalive += ds                # the copies read these buffers until the uploads have landed
        ids = np.asarray(list(image_ids), dtype=np.uint32)
        ptrs = (ctypes.c_void_p * n)(*[d.ctypes.data for d in ds])
        ns = (ctypes.c_size_t * n)(*[d.shape[0] for d in ds])
        self._check(self._L.smb_put_images_async(self._h, ids.ctypes.data, ctypes.cast(ptrs, ctypes.c_void_p),
                                                 ctypes.cast(ns, ctypes.c_void_p), n, 128))

    def put_keypoints(self, image_id: int, keypoints) -> None:
        """Keypoint positions of a cached image, for ``MatchResult.verify``: float32 [n, 2] (x, y) or the reference's
        FeatureKeypoint rows float32 [n, 6] (x, y, a11, a12, a21, a22; io.cc:115-123) -- only x, y are read."""
        k = np.ascontiguousarray(keypoints, dtype=np.float32)
        if k.ndim != 2 or k.shape[1] < 2:
            raise ValueError("keypoints must be [n, >= 2] float32")
        self._check(self._L.smb_put_keypoints(self._h, int(image_id), k.ctypes.data, k.shape[0], k.strides[0]))

    def put_image_device(self, image_id: int, dev_ptr: int, n: int) -> None:
        self._check(self._L.smb_put_image_device(self._h, int(image_id), ctypes.c_void_p(dev_ptr), int(n), 128))

    def put_images_device(self, image_ids: Sequence[int], dev_ptrs: Sequence[int], ns: Sequence[int]) -> None:
        n = len(dev_ptrs)
        ids = np.asarray(list(image_ids), dtype=np.uint32)
        ptrs = (ctypes.c_void_p * n)(*[int(p) for p in dev_ptrs])
        cnt = (ctypes.c_size_t * n)(*[int(x) for x in ns])
        self._check(self._L.smb_put_images_device(self._h, ids.ctypes.data, ctypes.cast(ptrs, ctypes.c_void_p),
                                                  ctypes.cast(cnt, ctypes.c_void_p), n, 128))

    def put_images_device_async(self, image_ids: Sequence[int], dev_ptrs: Sequence[int], ns: Sequence[int],
                                producer_stream: int = 0) -> None:
        """Adopt device buffers that work already queued on `producer_stream` (a cudaStream_t handle, e.g.
        torch.cuda.current_stream().cuda_stream after an NCCL recv) is still filling; returns at once."""
        n = len(dev_ptrs)
        ids = np.asarray(list(image_ids), dtype=np.uint32)
        ptrs = (ctypes.c_void_p * n)(*[int(p) for p in dev_ptrs])
        cnt = (ctypes.c_size_t * n)(*[int(x) for x in ns])
        self._check(self._L.smb_put_images_device_async(self._h, ids.ctypes.data, ctypes.cast(ptrs, ctypes.c_void_p),
                                                        ctypes.cast(cnt, ctypes.c_void_p), n, 128,
                                                        ctypes.c_void_p(int(producer_stream))))

    def has_image(self, image_id: int) -> bool:
        return bool(self._L.smb_has_image(self._h, int(image_id)))

    def evict_image(self, image_id: int) -> None:
        self._check(self._L.smb_evict_image(self._h, int(image_id)))

    def clear_images(self) -> None:
        self._check(self._L.smb_clear_images(self._h))   # drains pending uploads
        self._keepalive.clear()

    def image_device_ptr(self, image_id: int) -> Tuple[int, int]:
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        self._check(self._L.smb_image_device_ptr(self._h, int(image_id), ctypes.byref(p), ctypes.byref(n)))
        return (p.value or 0), n.value

    # -- matching ------------------------------------------------------------------------------
    def match_pairs_begin(self, pairs) -> "MatchResult":
        """Queue the whole call on the GPU and return at once; ``.wait()`` blocks until the matches are in host
        memory.  One call in flight per matcher."""
        pr = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        res = ctypes.c_void_p()
        self._check(self._L.smb_match_pairs_begin(self._h, pr.ctypes.data, pr.shape[0], ctypes.byref(res)))
        return MatchResult(self, res, pr.shape[0])

    def match_pairs_result(self, pairs) -> "MatchResult":
        """match_pairs, but the (completed) result object is handed out instead of copies of every list: the
        benchmark keeps the last timed call's result and checks a sample of it afterwards."""
        r = self.match_pairs_begin(pairs)
        r.wait()
        return r

    def match_pairs(self, pairs) -> List[np.ndarray]:
        """Match (image_id1, image_id2) pairs of cached images.  Returns one uint32 [m, 2] array per
        pair: (idx1, idx2) in ascending idx1 -- the FeatureMatches of MatchSiftFeaturesCPU."""
        with self.match_pairs_result(pairs) as r:
            return [r.matches(i) for i in range(r.num_pairs)]

    def match_pairs_count(self, pairs) -> int:
        """Same work as match_pairs (results reach pinned host memory) but only the total is returned;
        used by the benchmark so Python list building stays out of the timed region."""
        pr = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        res = ctypes.c_void_p()
        self._check(self._L.smb_match_pairs(self._h, pr.ctypes.data, pr.shape[0], ctypes.byref(res)))
        total = self._L.smb_result_total_matches(res)
        self._L.smb_result_release(self._h, res)
        return int(total)

    def match(self, descriptors1, descriptors2) -> np.ndarray:
        """One-shot MatchSiftFeaturesCPU(options, descriptors1, descriptors2, &matches)."""
        d1, d2 = _desc(descriptors1), _desc(descriptors2)
        cap = max(d1.shape[0], 1)
        out = np.empty((cap, 2), dtype=np.uint32)
        cnt = ctypes.c_size_t()
        self._check(self._L.smb_match_descriptors(self._h, d1.ctypes.data, d1.shape[0], d2.ctypes.data, d2.shape[0],
                                                  out.ctypes.data, cap, ctypes.byref(cnt)))
        return out[:cnt.value].copy()

    # -- introspection -------------------------------------------------------------------------
    def timing(self) -> dict:
        t = smb_timing()
        self._check(self._L.smb_get_timing(self._h, ctypes.byref(t)))
        return {k: getattr(t, k) for k, _ in smb_timing._fields_}

    @property
    def stream(self) -> int:
        return self._L.smb_stream(self._h) or 0

    def stream_wait_uploads(self, stream: int) -> None:
        """Make `stream` (a cudaStream_t handle) wait on the device for every upload queued so far."""
        self._check(self._L.smb_stream_wait_uploads(self._h, ctypes.c_void_p(int(stream))))

    def wait_stream(self, stream: int) -> None:
        """Order the kernels of the match calls that follow behind the work queued so far on `stream`."""
        self._check(self._L.smb_wait_stream(self._h, ctypes.c_void_p(int(stream))))

    def synchronize(self) -> None:
        self._check(self._L.smb_synchronize(self._h))
        self._keepalive.clear()


class MatchResult:
    """Owns one smb_result (pinned host memory the device wrote directly).  Arrays handed out are copies, so they
    stay valid after ``release()`` hands the buffers back to the matcher's pool."""

    def __init__(self, matcher: SiftMatcher, handle: ctypes.c_void_p, npairs: int):
        self._m, self._r, self._n, self._waited = matcher, handle, npairs, False

    def wait(self) -> "MatchResult":
        if not self._waited:
            self._m._check(self._m._L.smb_result_wait(self._m._h, self._r))
            self._waited = True
        return self

    @property
    def num_pairs(self) -> int:
        return self._n

    @property
    def total(self) -> int:
        self.wait()
        return int(self._m._L.smb_result_total_matches(self._r))

    def matches(self, i: int) -> np.ndarray:
        self.wait()
        cnt = ctypes.c_size_t()
        ptr = self._m._L.smb_result_matches(self._r, int(i), ctypes.byref(cnt))
        if cnt.value == 0:
            return np.empty((0, 2), dtype=np.uint32)
        a = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint32)), shape=(cnt.value, 2))
        return a.copy()

    def verify(self, detect_watermark: bool = True, multiple_models: bool = False, **opts) -> None:
        """Two-view geometry verification of every pair on the GPU (TwoViewGeometry::Estimate with the reference's
        dummy cameras: uncalibrated F / H LORANSAC, then the watermark test).  Options: min_num_inliers, min_num_trials,
        max_num_trials, max_error, confidence, min_inlier_ratio, max_h_inlier_ratio, seed (defaults:
        colmap.proto:24-44); ``detect_watermark`` (COLMAP default true), ``multiple_models`` (colmap.proto:45:
        TwoViewGeometry::EstimateMultiple)."""
        self.wait()
        o = smb_tvg_options()
        self._m._L.smb_default_tvg_options(ctypes.byref(o))
        o.flags = (0 if detect_watermark else SMB_TVG_NO_WATERMARK) | (SMB_TVG_MULTIPLE_MODELS if multiple_models else 0)
        for k, v in opts.items():
            if not hasattr(o, k):
                raise TypeError(f"unknown option {k}")
            setattr(o, k, v)
        self._m._check(self._m._L.smb_result_verify(self._m._h, self._r, ctypes.byref(o)))

    def tvg(self, i: int) -> dict:
        t = smb_tvg()
        if self._m._L.smb_result_tvg(self._r, int(i), ctypes.byref(t)) != SMB_OK:
            raise SmbError(SMB_EINVAL, "result has not been verified")
        return {"config": t.config, "num_inliers_F": t.num_inliers_f, "num_inliers_H": t.num_inliers_h,
                "trials_F": t.trials_f, "trials_H": t.trials_h, "F": np.array(t.F, dtype=np.float64).reshape(3, 3),
                "H": np.array(t.H, dtype=np.float64).reshape(3, 3)}

    def inliers(self, i: int) -> np.ndarray:
        cnt = ctypes.c_size_t()
        ptr = self._m._L.smb_result_inliers(self._r, int(i), ctypes.byref(cnt))
        if cnt.value == 0:
            return np.empty((0, 2), dtype=np.uint32)
        return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint32)), shape=(cnt.value, 2)).copy()

    def release(self) -> None:
        if self._r is not None and self._m._h:
            self._m._L.smb_result_release(self._m._h, self._r)
        self._r = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.release()

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def sequential_pairs(image_ids: Sequence[int], overlap: int) -> np.ndarray:
    """Pairs the reference op produces over a table of ``image_ids`` with stencil ``range(0, overlap)``
    (feature_matching.py:43; sequential_matching.cc:139-146): row r is matched with the next
    overlap-1 DISTINCT ids; Scanner's REPEAT_EDGE halo at the table tail is absorbed by the dedup."""
    ids = list(image_ids)
    n = len(ids)
    out = []
    for r in range(n):
        seen = []
        for s in range(1, overlap):
            x = ids[min(r + s, n - 1)]
            if x == ids[r] or x in seen:
                continue
            seen.append(x)
            out.append((ids[r], x))
    return np.asarray(out, dtype=np.uint32).reshape(-1, 2)
