"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Python face of the CPU oracle for the feature-matching hot path:

* ``match`` / ``match_many`` call the C restatement in ``sift_match_oracle.c``
  (COLMAP 3.5 ``MatchSiftFeaturesCPU`` as called by the reference at
  ``integration/op_cpp/sequential_matching.cc:154``).
* ``match_numpy`` is an independent numpy restatement of the same semantics
  (blocked int32 GEMM, top-2 by sort-free masking) used to cross-check the C
  code; both share the host libm ``acosf`` so float decisions are identical.
* ``sequential_pairs`` restates the reference's pair enumeration
  (``sequential_matching.cc:124-148`` with the stencil ``range(0, overlap)`` of
  ``integration/feature_matching.py:43``).

PARITY UNPINNED: the reference holds no golden vectors for this path and
COLMAP/Eigen/Scanner are not installable here; the oracle is
pinned by the hand-built known-answer tests in ``tests/test_oracle_kat.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Iterable, List, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

LUT_SIZE = 512 * 512 + 1  # dot products 0 .. 262144; larger values saturate


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "sift_match_oracle.c")
    if (force or not os.path.exists(_LIB_PATH)
            or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        L.oracle_dist_normed.restype = ctypes.c_float
        L.oracle_dist_normed.argtypes = [ctypes.c_int]
        L.oracle_acos_lut.restype = None
        L.oracle_acos_lut.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.oracle_distance_matrix.restype = None
        L.oracle_distance_matrix.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                             ctypes.c_size_t, ctypes.c_void_p]
        L.oracle_find_best_matches_one_way.restype = ctypes.c_size_t
        L.oracle_find_best_matches_one_way.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                                       ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                                       ctypes.c_void_p]
        L.oracle_find_best_matches.restype = ctypes.c_size_t
        L.oracle_find_best_matches.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                               ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
        L.oracle_match_sift_features_cpu.restype = ctypes.c_size_t
        L.oracle_match_sift_features_cpu.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                                     ctypes.c_size_t, ctypes.c_double, ctypes.c_double,
                                                     ctypes.c_int, ctypes.c_void_p]
        L.oracle_match_many.restype = ctypes.c_int
        L.oracle_match_many.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t),
                                        ctypes.c_void_p, ctypes.c_size_t, ctypes.c_double, ctypes.c_double,
                                        ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_int]
        L.oracle_num_procs.restype = ctypes.c_int
        del u8p
        _lib = L
    return _lib


def _as_desc(d) -> np.ndarray:
    d = np.ascontiguousarray(d, dtype=np.uint8)
    if d.ndim != 2 or d.shape[1] != 128:
        raise ValueError("descriptors must be [n, 128] uint8")
    return d


def acos_lut() -> np.ndarray:
    """float32[262145]: acosf(min(d / 512^2, 1)) from the host libm."""
    lut = np.empty(LUT_SIZE, dtype=np.float32)
    lib().oracle_acos_lut(lut.ctypes.data, LUT_SIZE)
    return lut


def distance_matrix(d1, d2) -> np.ndarray:
    d1, d2 = _as_desc(d1), _as_desc(d2)
    out = np.empty((d1.shape[0], d2.shape[0]), dtype=np.int32)
    lib().oracle_distance_matrix(d1.ctypes.data, d1.shape[0], d2.ctypes.data, d2.shape[0], out.ctypes.data)
    return out


def find_best_matches(dists, max_ratio=0.8, max_distance=0.7, cross_check=True) -> np.ndarray:
    """FindBestMatches on an explicit int32 matrix (lets KATs inject any value)."""
    dists = np.ascontiguousarray(dists, dtype=np.int32)
    n1, n2 = dists.shape
    out = np.empty((max(n1, 1), 2), dtype=np.uint32)
    c = lib().oracle_find_best_matches(dists.ctypes.data, n1, n2, np.float32(max_ratio),
                                       np.float32(max_distance), int(bool(cross_check)), out.ctypes.data)
    return out[:c].copy()


def one_way(dists, transposed=False, max_ratio=0.8, max_distance=0.7) -> np.ndarray:
    dists = np.ascontiguousarray(dists, dtype=np.int32)
    n1, n2 = dists.shape
    m = np.empty(max(n2 if transposed else n1, 1), dtype=np.int32)
    lib().oracle_find_best_matches_one_way(dists.ctypes.data, n1, n2, int(bool(transposed)),
                                           np.float32(max_ratio), np.float32(max_distance), m.ctypes.data)
    return m[: (n2 if transposed else n1)].copy()


def match(d1, d2, max_ratio=0.8, max_distance=0.7, cross_check=True) -> np.ndarray:
    """MatchSiftFeaturesCPU: uint32[m, 2] (idx1, idx2), ascending idx1."""
    d1, d2 = _as_desc(d1), _as_desc(d2)
    n1, n2 = d1.shape[0], d2.shape[0]
    out = np.empty((max(n1, 1), 2), dtype=np.uint32)
    c = lib().oracle_match_sift_features_cpu(d1.ctypes.data, n1, d2.ctypes.data, n2, float(max_ratio),
                                             float(max_distance), int(bool(cross_check)), out.ctypes.data)
    if c == ctypes.c_size_t(-1).value:
        raise MemoryError("oracle distance matrix allocation failed")
    return out[:c].copy()


def match_many(images: Sequence[np.ndarray], pairs, max_ratio=0.8, max_distance=0.7, cross_check=True,
               num_threads: int = 0) -> Tuple[List[np.ndarray], int]:
    """Match many (k1, k2) index pairs into ``images``; returns (list of uint32[m,2], threads used)."""
    imgs = [_as_desc(d) for d in images]
    pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
    npairs = pairs.shape[0]
    ptrs = (ctypes.c_void_p * len(imgs))(*[d.ctypes.data for d in imgs])
    ns = (ctypes.c_size_t * len(imgs))(*[d.shape[0] for d in imgs])
    caps = np.array([max(imgs[a].shape[0], 1) for a, _ in pairs], dtype=np.uint64)
    offs = np.zeros(npairs + 1, dtype=np.uint64)
    np.cumsum(caps, out=offs[1:])
    out = np.empty((int(offs[-1]), 2), dtype=np.uint32)
    counts = np.zeros(npairs, dtype=np.uint64)
    used = lib().oracle_match_many(ptrs, ns, pairs.ctypes.data, npairs, float(max_ratio), float(max_distance),
                                   int(bool(cross_check)), offs.ctypes.data, out.ctypes.data,
                                   counts.ctypes.data, int(num_threads))
    res = [out[int(offs[p]): int(offs[p]) + int(counts[p])].copy() for p in range(npairs)]
    return res, used


def num_procs() -> int:
    return lib().oracle_num_procs()


# --------------------------------------------------------------------------------------
# Independent numpy restatement (cross-checks the C code; small / medium sizes only)
# --------------------------------------------------------------------------------------

def _one_way_numpy(dists: np.ndarray, lut: np.ndarray, max_ratio: np.float32, max_distance: np.float32):
    """FindBestMatchesOneWay over the rows of ``dists`` (int32 [r, c])."""
    r, c = dists.shape
    res = np.full(r, -1, dtype=np.int64)
    if r == 0 or c == 0:
        return res
    best_i = np.argmax(dists, axis=1)  # first occurrence of the maximum == strict '>' ascending scan
    best = dists[np.arange(r), best_i].astype(np.int64)
    masked = dists.copy()
    masked[np.arange(r), best_i] = -1  # remove ONE instance: second-best of the multiset
    second = np.maximum(masked.max(axis=1), 0).astype(np.int64) if c > 1 else np.zeros(r, dtype=np.int64)
    has = best > 0
    bn = lut[np.minimum(best, LUT_SIZE - 1)]
    sn = lut[np.minimum(second, LUT_SIZE - 1)]
    ok = has & ~(bn > max_distance) & ~(bn >= (max_ratio * sn).astype(np.float32))
    res[ok] = best_i[ok]
    return res


def match_numpy(d1, d2, max_ratio=0.8, max_distance=0.7, cross_check=True) -> np.ndarray:
    d1, d2 = _as_desc(d1), _as_desc(d2)
    n1, n2 = d1.shape[0], d2.shape[0]
    if n1 == 0 or n2 == 0:
        return np.empty((0, 2), dtype=np.uint32)
    lut = acos_lut()
    mr, md = np.float32(max_ratio), np.float32(max_distance)
    # float64 GEMM is exact here: every partial sum < 2^53
    dists = (d1.astype(np.float64) @ d2.astype(np.float64).T).astype(np.int32)
    m12 = _one_way_numpy(dists, lut, mr, md)
    if cross_check:
        m21 = _one_way_numpy(np.ascontiguousarray(dists.T), lut, mr, md)
        i1 = np.nonzero(m12 >= 0)[0]
        keep = m21[m12[i1]] == i1
        i1 = i1[keep]
    else:
        i1 = np.nonzero(m12 >= 0)[0]
    return np.stack([i1, m12[i1]], axis=1).astype(np.uint32).reshape(-1, 2)


# --------------------------------------------------------------------------------------
# Pair enumeration of the reference op
# --------------------------------------------------------------------------------------

def stencil_rows(num_rows: int, row: int, overlap: int) -> List[int]:
    """Rows Scanner hands to execute() for ``row`` with stencil range(0, overlap):
    out-of-range rows repeat the table edge (REPEAT_EDGE) [ext]."""
    return [min(row + s, num_rows - 1) for s in range(overlap)]


def row_partners(image_ids: Sequence[int], row: int, overlap: int) -> List[int]:
    """sequential_matching.cc:139-146: partners of stencil[0], in stencil order, skipping the
    anchor id and ids already seen."""
    st = [image_ids[r] for r in stencil_rows(len(image_ids), row, overlap)]
    out: List[int] = []
    for x in st[1:]:
        if x == st[0] or x in out:
            continue
        out.append(x)
    return out


def sequential_pairs(image_ids: Sequence[int], overlap: int) -> List[Tuple[int, int]]:
    pairs: List[Tuple[int, int]] = []
    for r in range(len(image_ids)):
        for x in row_partners(image_ids, r, overlap):
            pairs.append((image_ids[r], x))
    return pairs
