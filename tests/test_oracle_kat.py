"""Known-answer tests that pin the CPU oracle (oracle/sift_match_oracle.c + the numpy mirror) to the
semantics of COLMAP 3.5 FindBestMatchesOneWay / FindBestMatches as called by the reference at
/root/reference/integration/op_cpp/sequential_matching.cc:154.  The reference ships no golden vectors for
this path (PARITY UNPINNED); these hand-built cases ARE the pin, each one derivable by hand
from the published algorithm."""
import json
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def oracle(built):
    from oracle import oracle as o
    return o


D = 512 * 512  # kDistNorm denominator


def test_lut_endpoints_and_threshold(oracle):
    lut = oracle.acos_lut()
    assert lut.shape == (D + 1,) and lut.dtype == np.float32
    assert lut[0] == np.float32(np.pi / 2) and lut[D] == 0.0
    assert np.all(np.diff(lut) <= 0)                      # monotone on this libm
    # default max_distance 0.7: the smallest passing score (200,499)
    assert int(np.argmax(lut <= np.float32(0.7))) == 200499


def test_tie_goes_to_lowest_index_and_is_then_rejected(oracle):
    # best == second-best: strict '>' keeps the first maximum, '>=' in the ratio test rejects it
    m = np.array([[250000, 250000, 10]], dtype=np.int32)
    assert oracle.one_way(m).tolist() == [-1]
    # with a ratio that can never reject (max_ratio large) the first index wins
    assert oracle.one_way(m, max_ratio=10.0).tolist() == [0]


def test_zero_dot_never_matches(oracle):
    m = np.zeros((2, 3), dtype=np.int32)
    assert oracle.one_way(m, max_ratio=10.0, max_distance=10.0).tolist() == [-1, -1]
    assert len(oracle.find_best_matches(m, 10.0, 10.0, False)) == 0


def test_saturation_above_512_squared(oracle):
    # scores above 512^2 all map to acos(1) = 0; best keeps the TRUE arg-max, second saturates too
    m = np.array([[D + 5, 3 * D, D + 7]], dtype=np.int32)
    assert oracle.one_way(m, max_ratio=0.8).tolist() == [-1]        # 0 >= 0.8 * 0 -> rejected
    m2 = np.array([[100, 3 * D, 50]], dtype=np.int32)
    assert oracle.one_way(m2, max_ratio=0.8).tolist() == [1]        # 0 < 0.8 * acos(100/D)


def test_max_distance_and_ratio_boundaries(oracle):
    lut = oracle.acos_lut()
    assert oracle.one_way(np.array([[200499, 0]], np.int32)).tolist() == [0]
    assert oracle.one_way(np.array([[200498, 0]], np.int32)).tolist() == [-1]   # fails max_distance
    # ratio test: best 250000; find the smallest second that rejects with float32 arithmetic
    bn = lut[250000]
    s = next(s for s in range(250000) if bn >= np.float32(0.8) * lut[s])
    assert oracle.one_way(np.array([[250000, s]], np.int32)).tolist() == [-1]
    assert oracle.one_way(np.array([[250000, s - 1]], np.int32)).tolist() == [0]


def test_cross_check_asymmetry(oracle):
    # row 0 -> col 1, row 1 -> col 1; col 1's best row is 1: only (1,1) survives the cross-check
    m = np.array([[10, 240000, 20], [30, 255000, 40]], dtype=np.int32)
    assert oracle.find_best_matches(m, 0.8, 0.7, False).tolist() == [[0, 1], [1, 1]]
    assert oracle.find_best_matches(m, 0.8, 0.7, True).tolist() == [[1, 1]]


def test_transposed_view_equals_explicit_transpose(oracle):
    rng = np.random.default_rng(0)
    m = rng.integers(0, 262144, size=(17, 23), dtype=np.int32)
    a = oracle.one_way(m, transposed=True, max_ratio=0.95, max_distance=1.3)
    b = oracle.one_way(np.ascontiguousarray(m.T), max_ratio=0.95, max_distance=1.3)
    assert a.tolist() == b.tolist()


def test_empty_and_single(oracle):
    e = np.empty((0, 128), np.uint8)
    one = np.full((1, 128), 46, np.uint8)      # |d|^2 = 128 * 46^2 = 270848 > 512^2 -> distance saturates to 0
    assert len(oracle.match(e, one)) == 0 and len(oracle.match(one, e)) == 0 and len(oracle.match(e, e)) == 0
    # one candidate: best distance 0 <= 0.7, second-best score 0 -> acos(0) = pi/2, 0 < 0.8 * pi/2 -> accepted
    assert oracle.match(one, one).tolist() == [[0, 0]]
    assert oracle.match(one, one, max_ratio=1.0, max_distance=0.0).tolist() == [[0, 0]]
    assert oracle.match(one, one, max_ratio=0.0).tolist() == []                     # 0 >= 0 * pi/2 -> rejected
    two = np.zeros((1, 128), np.uint8); two[0, :64] = 45   # dot with itself 129600
    assert oracle.match(two, two).tolist() == []            # acos(129600/262144) = 1.05 > 0.7
    assert oracle.match(two, two, max_distance=1.2).tolist() == [[0, 0]]


def test_numpy_mirror_agrees_with_c(oracle):
    from scanner_colmap_b200 import synth
    for n1, n2, cc in [(300, 257, True), (129, 400, False), (64, 64, True)]:
        a = synth.make_image(3, n1, track_step=8)
        b = synth.make_image(4, n2, track_step=8)
        assert np.array_equal(oracle.match(a, b, cross_check=cc), oracle.match_numpy(a, b, cross_check=cc))


def test_distance_matrix_is_exact_integer_dot(oracle):
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, (9, 128), dtype=np.uint8)
    b = rng.integers(0, 256, (7, 128), dtype=np.uint8)
    assert np.array_equal(oracle.distance_matrix(a, b), a.astype(np.int64) @ b.astype(np.int64).T)
    full = np.full((1, 128), 255, np.uint8)
    assert oracle.distance_matrix(full, full)[0, 0] == 128 * 255 * 255      # 8,323,200 < 2^24


def test_sequential_pairs_restates_op_loop(oracle):
    # feature_matching.py:43 stencil range(0, W); sequential_matching.cc:139-146 dedup of the REPEAT_EDGE halo
    ids = [7, 8, 9, 10, 11]
    assert oracle.sequential_pairs(ids, 3) == [(7, 8), (7, 9), (8, 9), (8, 10), (9, 10), (9, 11), (10, 11)]
    assert oracle.row_partners(ids, 4, 3) == []                 # last row: only repeated edge rows
    for n, w in [(20, 10), (100, 10), (1000, 20)]:
        assert len(oracle.sequential_pairs(list(range(n)), w)) == (w - 1) * n - w * (w - 1) // 2
    from scanner_colmap_b200 import sequential_pairs as product_pairs
    assert product_pairs(ids, 3).tolist() == [list(p) for p in oracle.sequential_pairs(ids, 3)]


def test_committed_golden_fixture(oracle):
    """tests/golden/matches_small.json was produced by tests/golden/make_golden.py (oracle C code on seeded
    synthetic descriptors); it freezes today's oracle so that later edits to it are caught."""
    from scanner_colmap_b200 import synth
    g = json.load(open(os.path.join(GOLDEN, "matches_small.json")))
    for case in g["cases"]:
        a = synth.make_image(case["id1"], case["n1"], track_step=case["track_step"])
        b = synth.make_image(case["id2"], case["n2"], track_step=case["track_step"])
        got = oracle.match(a, b, max_ratio=case["max_ratio"], max_distance=case["max_distance"],
                           cross_check=case["cross_check"])
        assert got.tolist() == case["matches"], case


def test_c_oracle_and_numpy_mirror_agree_property(oracle):
    """The two independent restatements (C, numpy) agree on random shapes, options and descriptor distributions that
    stress the decision boundaries: duplicated rows (ties), scaled-up rows (saturation above 512^2), sparse rows."""
    from hypothesis import given, settings, strategies as st
    from scanner_colmap_b200 import synth

    @settings(max_examples=40, deadline=None)
    @given(st.integers(0, 90), st.integers(0, 90), st.booleans(), st.sampled_from([0.0, 0.6, 0.8, 0.95, 1.0, 1.5]),
           st.sampled_from([0.0, 0.3, 0.7, 1.0, 1.5707964, 3.0]), st.integers(0, 2 ** 20), st.sampled_from(["sift", "dup", "hot", "sparse"]))
    def check(n1, n2, cc, ratio, dist, seed, kind):
        rng = np.random.default_rng(seed)
        a = synth.make_image(seed, n1, track_step=4) if n1 else np.empty((0, 128), np.uint8)
        b = synth.make_image(seed + 1, n2, track_step=4) if n2 else np.empty((0, 128), np.uint8)
        if kind == "dup" and n1 > 1 and n2 > 1:
            b[rng.integers(0, n2, n2 // 2)] = b[0]
            a[rng.integers(0, n1, n1 // 2)] = b[0]
        elif kind == "hot" and n1 and n2:
            a = np.minimum(a.astype(np.int32) * 3, 255).astype(np.uint8)
            b = np.minimum(b.astype(np.int32) * 3, 255).astype(np.uint8)
        elif kind == "sparse":
            a[:, 16:] = 0
            b[:, 16:] = 0
        got = oracle.match(a, b, max_ratio=ratio, max_distance=dist, cross_check=cc)
        assert np.array_equal(got, oracle.match_numpy(a, b, max_ratio=ratio, max_distance=dist, cross_check=cc))
        assert np.all(np.diff(got[:, 0].astype(np.int64)) > 0) if len(got) else True
        if cc and len(got):
            assert len(set(got[:, 1].tolist())) == len(got)                  # cross-check makes the matching one-to-one

    check()
