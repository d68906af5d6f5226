"""Byte-level tests of the op's row formats (io.cc) -- C++ (op/wire.h) against the Python mirror and against
committed golden bytes.  No GPU needed."""
import ctypes
import struct

import numpy as np
import pytest

from scanner_colmap_b200 import wire


@pytest.fixture(scope="module")
def H(built):
    from scanner_colmap_b200 import scanner_sim
    return scanner_sim.harness()


def test_op_registration_matches_reference(H):
    from scanner_colmap_b200 import scanner_sim
    # sequential_matching.cc:193-205
    assert scanner_sim.op_description() == ("SequentialMatchingCPU|stencil=1|in=image_ids,keypoints,descriptors,|"
                                            "out=pair_image_ids,two_view_geometries,|proto=featureMatchingArgs|"
                                            "device=CPU|batch=1|num_devices=1")


def test_image_id_is_low_half_of_size_t(H):
    b = wire.encode_image_id(0x1_0000_0007)            # prepare_image.cc:17 writes a size_t counter
    assert len(b) == 8 and H.smb_wire_image_id(b, 8) == 7 == wire.decode_image_id(b)


def test_descriptor_element_layout(H):
    d = (np.arange(3 * 128) % 251).astype(np.uint8).reshape(3, 128)
    b = wire.encode_descriptors(d)
    assert b[:16] == struct.pack("<QQ", 3, 128) and len(b) == 16 + 3 * 128
    r, c, o = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
    assert H.smb_wire_descriptor_view(b, len(b), ctypes.byref(r), ctypes.byref(c), ctypes.byref(o)) == 1
    assert (r.value, c.value, o.value) == (3, 128, 16)
    assert np.array_equal(wire.decode_descriptors(b), d)
    assert H.smb_wire_descriptor_view(b[:-1], len(b) - 1, ctypes.byref(r), ctypes.byref(c), ctypes.byref(o)) == 0  # truncated
    e = wire.encode_descriptors(np.empty((0, 128), np.uint8))
    assert H.smb_wire_descriptor_view(e, len(e), ctypes.byref(r), ctypes.byref(c), ctypes.byref(o)) == 1 and r.value == 0


def test_pair_ids_bytes(H):
    ids = np.array([5, 6, 4000000000], dtype=np.uint32)
    out = ctypes.create_string_buffer(64)
    n = H.smb_wire_pair_ids(ids.ctypes.data, 3, out, 64)
    assert out.raw[:n] == wire.encode_pair_ids(ids) == bytes.fromhex("0300000000000000" "05000000" "06000000" "00286bee")
    n0 = H.smb_wire_pair_ids(ids.ctypes.data, 0, out, 64)
    assert out.raw[:n0] == bytes(8)                     # last table row: n = 0


def test_two_view_geometry_golden_bytes_and_roundtrip(H):
    t0 = wire.TwoViewGeometry()                                    # default TVG: all zero, m = 0 (sequential_matching.cc:177)
    t1 = wire.TwoViewGeometry(config=2, E=np.arange(9), F=np.arange(9) * 0.5, H=-np.arange(9), qvec=[1, 0, 0, 0],
                              tvec=[0.1, 0.2, 0.3], tri_angle=0.25,
                              inlier_matches=np.array([[1, 2], [3, 40000]], np.uint32))
    b = wire.encode_two_view_geometries([t0, t1])
    assert len(b) == 12 + (284 + 8) + (284 + 8 + 16)               # 12 + sum(284 + 8 + 8 m_k)
    assert struct.unpack_from("<Qi", b, 0) == (len(b), 2)
    assert b[12:12 + 292] == bytes(292)                            # the default TVG serialises to zeros
    assert struct.unpack_from("<i", b, 12 + 292)[0] == 2
    assert struct.unpack_from("<9d", b, 12 + 292 + 4) == tuple(float(x) for x in range(9))   # E, Eigen storage order
    assert b[-24:] == struct.pack("<Q4I", 2, 1, 2, 3, 40000)
    out = ctypes.create_string_buffer(len(b))
    assert H.smb_wire_tvg_roundtrip(b, len(b), out, len(b)) == len(b) and out.raw == b      # C++ reader + writer
    assert H.smb_wire_tvg_roundtrip(b[:-1], len(b) - 1, out, len(b)) == 0                    # io.cc:249 length assert
    back = wire.decode_two_view_geometries(b)
    assert back[1].config == 2 and back[1].inlier_matches.tolist() == [[1, 2], [3, 40000]] and back[0].inlier_matches.shape == (0, 2)
    empty = wire.encode_two_view_geometries([])
    assert empty == struct.pack("<Qi", 12, 0)


def test_keypoints_layout():
    kp = np.arange(12, dtype=np.float32).reshape(2, 6)
    b = wire.encode_keypoints(kp)
    assert len(b) == 8 + 2 * 24 and np.array_equal(wire.decode_keypoints(b), kp)


def test_kernel_args_defaults_and_overrides(H):
    def parse(buf):
        mr, md = ctypes.c_double(), ctypes.c_double()
        cc, mn, mi, ov = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        me = ctypes.c_float()
        ok = H.smb_proto_parse(buf, len(buf), ctypes.byref(mr), ctypes.byref(md), ctypes.byref(cc), ctypes.byref(mn),
                               ctypes.byref(mi), ctypes.byref(ov), ctypes.byref(me))
        return ok, (mr.value, md.value, cc.value, mn.value, mi.value, ov.value, me.value)
    # feature_matching.py passes no args: proto2 defaults of colmap.proto:14-48,58
    assert parse(b"") == (1, (0.8, 0.7, 1, 32768, 15, 10, 4.0))
    a = wire.encode_matching_args(max_ratio=0.9, max_distance=1.1, cross_check=False, max_num_matches=100,
                                  min_num_inliers=3, overlap=20, max_error=2.5)
    assert parse(a) == (1, (0.9, 1.1, 0, 100, 3, 20, 2.5))
    assert parse(b"\x22\x7f")[0] == 0   # truncated embedded message is reported, defaults kept


def test_two_view_geometry_rows_roundtrip_property(H):
    """Random TwoViewGeometry lists (every configuration value incl. WATERMARK = 7 and MULTIPLE = 8, empty and long
    inlier lists) survive Python writer -> C++ reader -> C++ writer -> Python reader byte for byte (io.cc:224-297)."""
    from hypothesis import given, settings, strategies as st

    tvg = st.builds(
        lambda cfg, vals, m, seed: wire.TwoViewGeometry(
            config=cfg, E=vals[0:9], F=vals[9:18], H=vals[18:27], qvec=vals[27:31], tvec=vals[31:34], tri_angle=vals[34],
            inlier_matches=np.random.default_rng(seed).integers(0, 2 ** 32, size=(m, 2), dtype=np.uint64).astype(np.uint32)),
        st.integers(0, 8), st.lists(st.floats(-1e6, 1e6, allow_nan=False), min_size=35, max_size=35), st.integers(0, 300),
        st.integers(0, 2 ** 31))

    @settings(max_examples=40, deadline=None)
    @given(st.lists(tvg, min_size=0, max_size=5))
    def check(tvgs):
        b = wire.encode_two_view_geometries(tvgs)
        assert len(b) == 12 + sum(284 + 8 + 8 * len(t.inlier_matches) for t in tvgs)
        out = ctypes.create_string_buffer(max(len(b), 1))
        assert H.smb_wire_tvg_roundtrip(b, len(b), out, len(b)) == len(b) and out.raw[:len(b)] == b
        back = wire.decode_two_view_geometries(b)
        assert len(back) == len(tvgs)
        for x, y in zip(tvgs, back):
            assert x.config == y.config and np.array_equal(x.inlier_matches, y.inlier_matches)
            assert np.array_equal(np.asarray(x.F, float).ravel(), np.asarray(y.F, float).ravel())
        if len(b) > 12:
            assert H.smb_wire_tvg_roundtrip(b[:-3], len(b) - 3, out, len(b)) == 0     # truncated rows are refused

    check()
