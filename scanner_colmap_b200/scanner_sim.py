"""A stand-in for Scanner's dispatch of one stencilled, batched op over a table -- just enough to drive the
SequentialMatchingCPU op the way ``integration/feature_matching.py`` does (Scanner itself is not installable
here):

* the table has one row per image (columns image_id, keypoints, descriptors of the ``extraction`` table);
* ``stencil=range(0, overlap)``: row r sees rows r .. r+overlap-1, rows past the table end repeat the last row
  (REPEAT_EDGE), feature_matching.py:43;
* ``io_packet_size = work_packet_size = packet_size``: the kernel's execute() receives ``packet_size`` rows at a
  time (a new kernel instance per run, as Scanner makes one per pipeline instance), feature_matching.py:70-74;
* outputs: columns ``pair_image_ids`` and ``two_view_geometries``, one element per row.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Sequence, Tuple

import numpy as np

from . import wire

_HERE = os.path.dirname(os.path.abspath(__file__))
HARNESS_PATH = os.path.join(_HERE, "op", "build", "libsmb_op_harness.so")
OP_LIB_PATH = os.path.join(_HERE, "op", "build", "libsequential_matching.so")
_lib = None


def harness() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(HARNESS_PATH):
            raise RuntimeError(f"{HARNESS_PATH} not built (run __graft_entry__.build())")
        L = ctypes.CDLL(HARNESS_PATH)
        vp = ctypes.c_void_p
        L.smb_op_registered.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
        L.smb_op_new_kernel.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
        L.smb_op_new_kernel.restype = vp
        L.smb_op_delete_kernel.argtypes = [vp]
        L.smb_op_delete_kernel.restype = None
        L.smb_op_execute.argtypes = [vp, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, vp, vp, ctypes.c_size_t]
        L.smb_op_reset.argtypes = [vp]
        L.smb_op_reset.restype = None
        L.smb_op_new_stream.argtypes = [vp]
        L.smb_op_new_stream.restype = None
        L.smb_op_output_count.argtypes = [vp, ctypes.c_size_t]
        L.smb_op_output_count.restype = ctypes.c_size_t
        L.smb_op_output.argtypes = [vp, ctypes.c_size_t, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
        L.smb_op_output.restype = vp
        L.smb_wire_tvg_roundtrip.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
        L.smb_wire_tvg_roundtrip.restype = ctypes.c_size_t
        L.smb_wire_pair_ids.argtypes = [vp, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
        L.smb_wire_pair_ids.restype = ctypes.c_size_t
        L.smb_wire_descriptor_view.argtypes = [ctypes.c_char_p, ctypes.c_size_t] + [ctypes.POINTER(ctypes.c_size_t)] * 3
        L.smb_wire_image_id.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
        L.smb_wire_image_id.restype = ctypes.c_uint32
        L.smb_proto_parse.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_double),
                                      ctypes.POINTER(ctypes.c_double)] + [ctypes.POINTER(ctypes.c_int)] * 4 + \
                                     [ctypes.POINTER(ctypes.c_float)]
        _lib = L
    return _lib


def op_description(name: str = "SequentialMatchingCPU") -> str:
    buf = ctypes.create_string_buffer(512)
    if not harness().smb_op_registered(name.encode(), buf, 512):
        raise KeyError(name)
    return buf.value.decode()


class OpKernel:
    """One kernel instance of the op, as Scanner keeps one per pipeline instance: it outlives a single table (Scanner
    reuses instances across tasks and jobs and announces the change with reset() / new_stream())."""

    def __init__(self, args: bytes = b"", name: str = "SequentialMatchingCPU"):
        self._L = harness()
        self._k = self._L.smb_op_new_kernel(name.encode(), args, len(args))
        if not self._k:
            raise RuntimeError("kernel creation failed")

    def close(self):
        if self._k:
            self._L.smb_op_delete_kernel(self._k)
            self._k = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def reset(self):
        self._L.smb_op_reset(self._k)

    def new_stream(self):
        self._L.smb_op_new_stream(self._k)

    def run_table(self, image_ids: Sequence[int], keypoints: Sequence[np.ndarray], descriptors: Sequence[np.ndarray],
                  overlap: int = 10, packet_size: int = 25, decode: bool = True, encoded=None):
        """``feature_matching.py --overlap W --packet_size P`` over an in-memory extraction table.  Returns
        (pair_image_ids per row, two_view_geometries per row) -- decoded, or with ``decode=False`` the raw
        serialized rows (the op-level benchmark keeps Python decoding out of its timed region)."""
        L, k = self._L, self._k
        n = len(image_ids)
        cols = encoded or encode_table(image_ids, keypoints, descriptors)
        out_ids, out_tvg = [], []
        for start in range(0, n, packet_size):
            rows = list(range(start, min(start + packet_size, n)))
            batch = len(rows)
            count = 3 * batch * overlap
            bufs = (ctypes.c_char_p * count)()
            sizes = (ctypes.c_size_t * count)()
            x = 0
            for c in range(3):
                for r in rows:
                    for s in range(overlap):
                        e = cols[c][min(r + s, n - 1)]  # REPEAT_EDGE
                        bufs[x] = e
                        sizes[x] = len(e)
                        x += 1
            L.smb_op_execute(k, 3, batch, overlap, ctypes.cast(bufs, ctypes.c_void_p), ctypes.cast(sizes, ctypes.c_void_p), 2)
            for col, sink, dec in ((0, out_ids, wire.decode_pair_ids), (1, out_tvg, wire.decode_two_view_geometries)):
                assert L.smb_op_output_count(k, col) == batch
                for i in range(batch):
                    sz = ctypes.c_size_t()
                    p = L.smb_op_output(k, col, i, ctypes.byref(sz))
                    sink.append(dec(ctypes.string_at(p, sz.value)) if decode else sz.value)
        return out_ids, out_tvg


def encode_table(image_ids, keypoints, descriptors):
    """The three input columns of the ``extraction`` table as serialized elements (io.cc formats)."""
    return [[wire.encode_image_id(i) for i in image_ids],
            [wire.encode_keypoints(k) for k in keypoints],
            [wire.encode_descriptors(d) for d in descriptors]]


def run_feature_matching(image_ids: Sequence[int], keypoints: Sequence[np.ndarray], descriptors: Sequence[np.ndarray],
                         overlap: int = 10, packet_size: int = 25, args: bytes = b"") -> Tuple[List[List[int]], List[list]]:
    """A fresh kernel instance over one table (what one ``python3 feature_matching.py`` run amounts to)."""
    with OpKernel(args) as k:
        return k.run_table(image_ids, keypoints, descriptors, overlap=overlap, packet_size=packet_size)
