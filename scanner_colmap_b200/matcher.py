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
    _fields_ = [("total_ms", ctypes.c_float), ("score_ms", ctypes.c_float), ("decide_ms", ctypes.c_float),
                ("score_launches", ctypes.c_uint32), ("total_launches", ctypes.c_uint32),
                ("candidates", ctypes.c_uint64), ("ops", ctypes.c_uint64)]


class smb_match(ctypes.Structure):
    _fields_ = [("idx1", ctypes.c_uint32), ("idx2", ctypes.c_uint32)]


_lib = None


def load_library(path: Optional[str] = None) -> ctypes.CDLL:
    """dlopen libsmb.so and declare every prototype of include/smb.h.  Raises if the library is absent:
    the product path must fail loudly rather than fall back."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("SMB_LIB") or LIB_PATH   # SMB_LIB: A/B a build variant (tools/bin/*.so) under the tests
    if not os.path.exists(p):
        raise SmbError(SMB_ENODEVICE, f"{p} not built -- run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = ctypes.CDLL(p)
    vp, u32p, szp = ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_size_t)
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
    L.smb_synchronize.argtypes = [vp]
    del u32p
    if path is None:
        _lib = L
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
        self._L = load_library()
        self._h = ctypes.c_void_p()
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
        self._keepalive = getattr(self, "_keepalive", [])[-256:] + ds
        ids = np.asarray(list(image_ids), dtype=np.uint32)
        ptrs = (ctypes.c_void_p * n)(*[d.ctypes.data for d in ds])
        ns = (ctypes.c_size_t * n)(*[d.shape[0] for d in ds])
        self._check(self._L.smb_put_images_async(self._h, ids.ctypes.data, ctypes.cast(ptrs, ctypes.c_void_p),
                                                 ctypes.cast(ns, ctypes.c_void_p), n, 128))

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
        self._check(self._L.smb_clear_images(self._h))

    def image_device_ptr(self, image_id: int) -> Tuple[int, int]:
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        self._check(self._L.smb_image_device_ptr(self._h, int(image_id), ctypes.byref(p), ctypes.byref(n)))
        return (p.value or 0), n.value

    # -- matching ------------------------------------------------------------------------------
    def match_pairs(self, pairs, copy: bool = True) -> List[np.ndarray]:
        """Match (image_id1, image_id2) pairs of cached images.  Returns one uint32 [m, 2] array per
        pair: (idx1, idx2) in ascending idx1 -- the FeatureMatches of MatchSiftFeaturesCPU."""
        pr = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        res = ctypes.c_void_p()
        self._check(self._L.smb_match_pairs(self._h, pr.ctypes.data, pr.shape[0], ctypes.byref(res)))
        try:
            out = []
            cnt = ctypes.c_size_t()
            for i in range(pr.shape[0]):
                ptr = self._L.smb_result_matches(res, i, ctypes.byref(cnt))
                if cnt.value == 0:
                    out.append(np.empty((0, 2), dtype=np.uint32))
                    continue
                a = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint32)), shape=(cnt.value, 2))
                out.append(a.copy() if copy else a)
            return out
        finally:
            self._L.smb_result_release(self._h, res)

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

    def synchronize(self) -> None:
        self._check(self._L.smb_synchronize(self._h))


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
