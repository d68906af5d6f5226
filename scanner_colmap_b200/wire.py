"""Python mirror of the op's row formats (``/root/reference/integration/op_cpp/io.cc``), used by the fake
Scanner dispatch and the tests.  Little-endian, ``size_t`` = 8 bytes, exactly the bytes the C++ side
(``op/wire.h``) reads and writes:

* image id        8-byte ``size_t`` (prepare_image.cc:17); consumers read the low 4 bytes (io.cc:54-56)
* keypoints       ``[size_t n][n x 6 float32]``                                   (io.cc:151-162)
* descriptors     ``[size_t rows][size_t cols][rows*cols uint8]``                 (io.cc:198-212)
* pair_image_ids  ``[size_t n][n x uint32]``                                      (io.cc:151-176)
* two_view_geometries ``[size_t total][int32 n] n x {int32 config, 9+9+9+4+3+1 float64, size_t m, m x (u32,u32)}``
                                                                                  (io.cc:256-297)
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

TVG_FIXED = 4 + 35 * 8  # 284


def encode_image_id(image_id: int) -> bytes:
    return struct.pack("<Q", int(image_id))


def decode_image_id(buf: bytes) -> int:
    return struct.unpack_from("<I", buf, 0)[0]


def encode_keypoints(kp: np.ndarray) -> bytes:
    kp = np.ascontiguousarray(kp, dtype="<f4").reshape(-1, 6)
    return struct.pack("<Q", kp.shape[0]) + kp.tobytes()


def decode_keypoints(buf: bytes) -> np.ndarray:
    n = struct.unpack_from("<Q", buf, 0)[0]
    return np.frombuffer(buf, dtype="<f4", count=6 * n, offset=8).reshape(n, 6)


def encode_descriptors(desc: np.ndarray) -> bytes:
    desc = np.ascontiguousarray(desc, dtype=np.uint8)
    return struct.pack("<QQ", desc.shape[0], desc.shape[1]) + desc.tobytes()


def decode_descriptors(buf: bytes) -> np.ndarray:
    rows, cols = struct.unpack_from("<QQ", buf, 0)
    return np.frombuffer(buf, dtype=np.uint8, count=rows * cols, offset=16).reshape(rows, cols)


def encode_pair_ids(ids: Sequence[int]) -> bytes:
    a = np.asarray(list(ids), dtype="<u4")
    return struct.pack("<Q", a.size) + a.tobytes()


def decode_pair_ids(buf: bytes) -> List[int]:
    n = struct.unpack_from("<Q", buf, 0)[0]
    return np.frombuffer(buf, dtype="<u4", count=n, offset=8).tolist()


@dataclass
class TwoViewGeometry:
    config: int = 0
    E: np.ndarray = field(default_factory=lambda: np.zeros(9))
    F: np.ndarray = field(default_factory=lambda: np.zeros(9))
    H: np.ndarray = field(default_factory=lambda: np.zeros(9))
    qvec: np.ndarray = field(default_factory=lambda: np.zeros(4))
    tvec: np.ndarray = field(default_factory=lambda: np.zeros(3))
    tri_angle: float = 0.0
    inlier_matches: np.ndarray = field(default_factory=lambda: np.empty((0, 2), np.uint32))


def encode_two_view_geometries(tvgs: Sequence[TwoViewGeometry]) -> bytes:
    body = struct.pack("<i", len(tvgs))
    for t in tvgs:
        m = np.ascontiguousarray(t.inlier_matches, dtype="<u4").reshape(-1, 2)
        body += struct.pack("<i", int(t.config))
        for a, k in ((t.E, 9), (t.F, 9), (t.H, 9), (t.qvec, 4), (t.tvec, 3)):
            body += np.asarray(a, dtype="<f8").reshape(k).tobytes()
        body += struct.pack("<d", float(t.tri_angle)) + struct.pack("<Q", m.shape[0]) + m.tobytes()
    return struct.pack("<Q", 8 + len(body)) + body


def decode_two_view_geometries(buf: bytes) -> List[TwoViewGeometry]:
    total, n = struct.unpack_from("<Qi", buf, 0)
    if total != len(buf):
        raise ValueError("two_view_geometries length check failed (io.cc:249)")
    off = 12
    out = []
    for _ in range(n):
        config = struct.unpack_from("<i", buf, off)[0]
        off += 4
        d = np.frombuffer(buf, dtype="<f8", count=35, offset=off)
        off += 35 * 8
        m = struct.unpack_from("<Q", buf, off)[0]
        off += 8
        matches = np.frombuffer(buf, dtype="<u4", count=2 * m, offset=off).reshape(m, 2).copy()
        off += 8 * m
        out.append(TwoViewGeometry(config, d[0:9].copy(), d[9:18].copy(), d[18:27].copy(), d[27:31].copy(),
                                   d[31:34].copy(), float(d[34]), matches))
    if off != len(buf):
        raise ValueError("trailing bytes")
    return out


def encode_matching_args(*, max_ratio=None, max_distance=None, cross_check=None, max_num_matches=None,
                         min_num_inliers=None, overlap=None, max_error=None) -> bytes:
    """Serialise SequentialMatchingArgs (colmap.proto) by hand: only the given fields are emitted, the rest keep
    their proto2 defaults -- ``b""`` is what feature_matching.py effectively sends."""
    def varint(v: int) -> bytes:
        v &= (1 << 64) - 1
        out = b""
        while True:
            b = v & 0x7F
            v >>= 7
            out += bytes([b | (0x80 if v else 0)])
            if not v:
                return out
    sift = b""
    if max_ratio is not None:
        sift += bytes([3 << 3 | 1]) + struct.pack("<d", max_ratio)
    if max_distance is not None:
        sift += bytes([4 << 3 | 1]) + struct.pack("<d", max_distance)
    if cross_check is not None:
        sift += bytes([5 << 3 | 0]) + varint(int(bool(cross_check)))
    if max_num_matches is not None:
        sift += bytes([6 << 3 | 0]) + varint(max_num_matches)
    if max_error is not None:
        sift += bytes([7 << 3 | 5]) + struct.pack("<f", max_error)
    if min_num_inliers is not None:
        sift += bytes([12 << 3 | 0]) + varint(min_num_inliers)
    msg = b""
    if overlap is not None:
        msg += bytes([2 << 3 | 0]) + varint(overlap)
    if sift:
        msg += bytes([4 << 3 | 2]) + varint(len(sift)) + sift
    return msg
