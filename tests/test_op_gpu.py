"""The Scanner op end to end on the GPU, driven the way integration/feature_matching.py drives it
(stencil range(0, overlap), packets of packet_size rows, REPEAT_EDGE at the table tail), against the oracle's
restatement of the reference op's pair loop (sequential_matching.cc:139-181)."""
import numpy as np
import pytest

from scanner_colmap_b200 import scanner_sim, synth, wire

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oracle(built):
    from oracle import oracle as o
    return o


@pytest.fixture(autouse=True)
def raw_matches(monkeypatch, request):
    """The matcher-parity tests compare the op's rows with the oracle's MATCHES: geometric verification off
    (SMB_OP_VERIFY=none); test_op_verifies_on_the_gpu switches it back on."""
    if "verifies" not in request.node.name:
        monkeypatch.setenv("SMB_OP_VERIFY", "none")


def _expected_rows(oracle, ids, descs, overlap, min_num_inliers=15, **opts):
    rows = []
    for r in range(len(ids)):
        partners = oracle.row_partners(ids, r, overlap)
        tv = []
        for pid in partners:
            m = oracle.match(descs[r], descs[ids.index(pid)], **opts)
            tv.append(m if len(m) >= min_num_inliers else np.empty((0, 2), np.uint32))   # sequential_matching.cc:173-178
        rows.append((partners, tv))
    return rows


@pytest.mark.parametrize("packet_size", [4, 25])
def test_feature_matching_config1_shape(oracle, packet_size):
    """BASELINE.json configs[0] (sequential overlap=10, packet_size=4 on 20 images), at 1024 descriptors per image
    so the CPU oracle finishes in seconds; test_gpu_parity covers 20 x 4096 through the C ABI."""
    n = 20
    ids = [1000 + 3 * k for k in range(n)]                     # ids need not be 0..n-1
    sizes = [1024 if k % 5 else 700 + k for k in range(n)]
    descs = [synth.make_image(i, s, track_step=48) for i, s in zip(ids, sizes)]
    kps = [np.zeros((s, 6), np.float32) for s in sizes]
    got_ids, got_tvg = scanner_sim.run_feature_matching(ids, kps, descs, overlap=10, packet_size=packet_size)
    want = _expected_rows(oracle, ids, descs, 10)
    assert len(got_ids) == len(got_tvg) == n
    total_pairs = 0
    for r, (partners, tv) in enumerate(want):
        assert got_ids[r] == partners, f"row {r}"
        assert len(got_tvg[r]) == len(partners)
        for g, w in zip(got_tvg[r], tv):
            assert np.array_equal(g.inlier_matches, w)
            assert g.config == 0 and not g.E.any() and not g.F.any() and not g.H.any()
        total_pairs += len(partners)
    assert total_pairs == 135 and got_ids[-1] == []            # last row: only repeated edge rows, n = 0
    assert sum(len(g.inlier_matches) for row in got_tvg for g in row) > 1000


def test_kernel_args_reach_the_matcher(oracle):
    ids = [1, 2, 3]
    descs = [synth.make_image(i, 400, track_step=16, noise=0.2) for i in ids]
    kps = [np.zeros((400, 6), np.float32)] * 3
    args = wire.encode_matching_args(max_ratio=0.95, max_distance=1.0, cross_check=False, min_num_inliers=1)
    got_ids, got_tvg = scanner_sim.run_feature_matching(ids, kps, descs, overlap=3, packet_size=2, args=args)
    want = _expected_rows(oracle, ids, descs, 3, min_num_inliers=1, max_ratio=0.95, max_distance=1.0, cross_check=False)
    for r, (partners, tv) in enumerate(want):
        assert got_ids[r] == partners
        for g, w in zip(got_tvg[r], tv):
            assert np.array_equal(g.inlier_matches, w)


def test_duplicate_ids_and_empty_images(oracle):
    """An id repeated inside the stencil is matched once; images without descriptors yield empty geometries."""
    ids = [5, 6, 6, 7, 8]
    sizes = [300, 200, 200, 0, 260]
    base = {i: synth.make_image(i, s, track_step=16) for i, s in zip(ids, sizes)}
    descs = [base[i] for i in ids]
    kps = [np.zeros((len(d), 6), np.float32) for d in descs]
    got_ids, got_tvg = scanner_sim.run_feature_matching(ids, kps, descs, overlap=4, packet_size=3)
    assert got_ids[0] == [6, 7] and got_ids[1] == [7, 8] and got_ids[2] == [7, 8] and got_ids[4] == []
    for r in range(len(ids)):
        for pid, g in zip(got_ids[r], got_tvg[r]):
            w = oracle.match(descs[r], base[pid])
            w = w if len(w) >= 15 else np.empty((0, 2), np.uint32)
            assert np.array_equal(g.inlier_matches, w)


def test_one_kernel_instance_two_tables_with_colliding_ids(oracle):
    """Image ids restart at 0 in every prepare_image kernel instance (prepare_image.cc:12), so they collide across
    tables and jobs, and Scanner reuses a kernel instance across them.  The op's descriptor cache must never serve
    table A's descriptors to table B: with new_stream() announced, and -- belt and braces -- without it (every cache
    hit is validated against the row's size and content signature)."""
    n, overlap = 8, 4
    ids = list(range(n))
    table_a = [synth.make_image(100 + i, 600 + 10 * i, track_step=24) for i in ids]
    table_b = [synth.make_image(200 + i, 600 + 10 * i, track_step=24) for i in ids]      # same ids, same sizes, other bytes
    kps = [np.zeros((len(d), 6), np.float32) for d in table_a]
    want_a = _expected_rows(oracle, ids, table_a, overlap)
    want_b = _expected_rows(oracle, ids, table_b, overlap)

    def check(got, want):
        got_ids, got_tvg = got
        for r, (partners, tv) in enumerate(want):
            assert got_ids[r] == partners
            for g, w in zip(got_tvg[r], tv):
                assert np.array_equal(g.inlier_matches, w), f"row {r}"

    with scanner_sim.OpKernel() as k:
        check(k.run_table(ids, kps, table_a, overlap=overlap, packet_size=3), want_a)
        check(k.run_table(ids, kps, table_b, overlap=overlap, packet_size=3), want_b)     # no reset announced
        k.new_stream()
        check(k.run_table(ids, kps, table_a, overlap=overlap, packet_size=8), want_a)
        k.reset()
        check(k.run_table(ids, kps, table_b, overlap=overlap, packet_size=2), want_b)


def test_kernel_instances_share_one_gpu_context(oracle):
    """Scanner makes one CPU kernel instance per pipeline instance; they must not each own a matcher (pool, log,
    accumulators, persistent kernels).  Two instances with different kernel args interleave on the shared handle."""
    ids = [11, 12, 13, 14]
    descs = [synth.make_image(i, 500, track_step=16, noise=0.2) for i in ids]
    kps = [np.zeros((500, 6), np.float32)] * 4
    args2 = wire.encode_matching_args(max_ratio=0.95, max_distance=1.0, cross_check=False, min_num_inliers=1)
    want1 = _expected_rows(oracle, ids, descs, 3)
    want2 = _expected_rows(oracle, ids, descs, 3, min_num_inliers=1, max_ratio=0.95, max_distance=1.0, cross_check=False)
    with scanner_sim.OpKernel() as k1, scanner_sim.OpKernel(args2) as k2:
        for _ in range(2):
            for k, want in ((k1, want1), (k2, want2)):
                got_ids, got_tvg = k.run_table(ids, kps, descs, overlap=3, packet_size=2)
                for r, (partners, tv) in enumerate(want):
                    assert got_ids[r] == partners
                    for g, w in zip(got_tvg[r], tv):
                        assert np.array_equal(g.inlier_matches, w)


def test_op_verifies_on_the_gpu(built):
    """Default op behaviour without COLMAP: every pair's matches go through the GPU two-view verification; rows carry
    config / F / H / inlier_matches, and pairs below min_num_inliers become default geometries (:173-178)."""
    from oracle import two_view_oracle as tv
    n = 2048
    p1, p2, m, truth = tv.synthetic_pair(n, n, 700, 200, 31)
    q1, q2, m2, truth2 = tv.synthetic_pair(n, n, 0, 60, 32)          # geometrically meaningless matches
    d = [synth.make_image(8100 + i, n, shared_frac=0.0) for i in range(3)]
    d[1][m[:, 1]] = d[0][m[:, 0]]                                     # image 0 <-> 1: a real scene
    free = np.setdiff1d(np.arange(n), m[:, 1])[:60]
    d[2][m2[:60, 1]] = d[1][free]                                     # image 1 <-> 2: 60 arbitrary matches
    kp = [np.zeros((n, 6), np.float32) for _ in range(3)]
    kp[0][:, :2], kp[1][:, :2] = p1, p2
    kp[2][:, :2] = q2
    got_ids, got_tvg = scanner_sim.run_feature_matching([40, 41, 42], kp, d, overlap=2, packet_size=3)
    assert got_ids == [[41], [42], []]
    g = got_tvg[0][0]
    assert g.config == tv.UNCALIBRATED and g.F.any() and not g.E.any()
    true_set = set(map(tuple, m[truth].tolist()))
    inl = set(map(tuple, g.inlier_matches.tolist()))
    assert len(inl & true_set) >= 0.95 * len(true_set) * 0.98 and len(inl - true_set) <= 0.05 * len(inl)
    x1, x2 = p1[g.inlier_matches[:, 0]].astype(np.float64), p2[g.inlier_matches[:, 1]].astype(np.float64)
    F = g.F.reshape(3, 3).T                                        # serialised in Eigen's column-major order (io.cc:283-293)
    assert np.median(tv.sampson_sq(F, x1, x2)) < 1.5               # ... and it explains its inliers
    g2 = got_tvg[1][0]
    assert len(g2.inlier_matches) == 0 or len(g2.inlier_matches) >= 15
