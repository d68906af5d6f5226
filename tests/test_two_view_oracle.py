"""Known-answer tests that pin oracle/two_view_oracle.py (the CPU restatement of COLMAP's uncalibrated two-view
geometry estimation, sequential_matching.cc:84-101 -> TwoViewGeometry::Estimate).  The reference ships no vectors for
this step either, so the pins are geometric identities with known answers."""
import math

import numpy as np
import pytest


@pytest.fixture(scope="module")
def tv():
    from oracle import two_view_oracle
    return two_view_oracle


def _scene(seed=0, n=60, planar=False):
    rng = np.random.default_rng(seed)
    K = np.array([[3000.0, 0, 2000], [0, 3000.0, 1500], [0, 0, 1]])
    a = 0.1
    R = np.array([[math.cos(a), 0, math.sin(a)], [0, 1, 0], [-math.sin(a), 0, math.cos(a)]])
    t = np.array([1.0, 0.2, 0.1])
    X = np.stack([rng.uniform(-3, 3, n), rng.uniform(-2, 2, n), np.full(n, 9.0) if planar else rng.uniform(6, 14, n)], axis=1)
    u1 = (K @ X.T).T
    u2 = (K @ (R @ X.T + t[:, None])).T
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    F = np.linalg.inv(K).T @ tx @ R @ np.linalg.inv(K)
    Hm = None
    if planar:                                  # plane z = 9: H = K (R + t n^T / d) K^-1 with n = (0, 0, 1), d = 9
        Hm = K @ (R + np.outer(t, [0, 0, 1.0]) / 9.0) @ np.linalg.inv(K)
    return u1[:, :2] / u1[:, 2:], u2[:, :2] / u2[:, 2:], F, Hm


def _same_up_to_scale(A, B, tol):
    A, B = A / np.linalg.norm(A), B / np.linalg.norm(B)
    return min(np.abs(A - B).max(), np.abs(A + B).max()) < tol


def test_residuals_vanish_on_the_true_models(tv):
    x1, x2, F, _ = _scene(1)
    assert tv.sampson_sq(F, x1, x2).max() < 1e-12
    p1, p2, _, H = _scene(2, planar=True)
    assert tv.transfer_sq(H, p1, p2).max() < 1e-12
    # a point moved d = 3 px off its epipolar line l2 = F x1: the Sampson error spreads the squared distance over both
    # images, d^2 * |l2|^2 / (|l2|^2 + |l1|^2) with l1 = F^T x2' -- about d^2 / 2 for this geometry
    l2 = F @ np.array([x1[0, 0], x1[0, 1], 1.0])
    moved = x2[:1] + 3.0 * l2[:2] / np.linalg.norm(l2[:2])
    l1 = F.T @ np.array([moved[0, 0], moved[0, 1], 1.0])
    a, b = l2[0] ** 2 + l2[1] ** 2, l1[0] ** 2 + l1[1] ** 2
    got = tv.sampson_sq(F, x1[:1], moved)[0]
    assert abs(got - 9.0 * a / (a + b)) < 1e-9 and 4.0 < got < 5.0


def test_minimal_and_least_squares_solvers_recover_the_models(tv):
    x1, x2, F, _ = _scene(3)
    models = tv.fundamental_seven_point(x1[:7], x2[:7])
    assert len(models) in (1, 3)
    assert min(tv.sampson_sq(M, x1, x2).max() for M in models) < 1e-6     # one root is the true geometry
    for M in models:
        assert abs(np.linalg.det(M / np.linalg.norm(M))) < 1e-9            # every root is singular
        assert tv.sampson_sq(M, x1[:7], x2[:7]).max() < 1e-8               # ... and fits its seven points
    (F8,) = tv.fundamental_eight_point(x1, x2)
    assert _same_up_to_scale(F8, F, 1e-6) and np.linalg.matrix_rank(F8 / np.linalg.norm(F8), tol=1e-9) == 2
    p1, p2, _, H = _scene(4, planar=True)
    (H4,) = tv.homography_dlt(p1[:4], p2[:4])
    (Hn,) = tv.homography_dlt(p1, p2)
    assert _same_up_to_scale(H4, H, 1e-6) and _same_up_to_scale(Hn, H, 1e-6)


def test_compute_num_trials_known_values(tv):
    # RANSAC::ComputeNumTrials: ceil(log(1 - confidence) / log(1 - ratio^k) * multiplier)
    assert tv.compute_num_trials(50, 100, 0.999, 7, 3.0) == math.ceil(math.log(0.001) / math.log(1 - 0.5 ** 7) * 3.0)
    assert tv.compute_num_trials(75, 100, 0.999, 7, 3.0) == 145               # log(0.001) / log(1 - 0.75^7) * 3 = 144.5
    assert tv.compute_num_trials(25000, 100000, 0.999, 4, 3.0) == 5295        # the cap min_inlier_ratio = 0.25 puts on H
    assert tv.compute_num_trials(100, 100, 0.999, 7, 3.0) == 1
    assert tv.compute_num_trials(0, 100, 0.999, 7, 3.0) > 10 ** 12


def test_estimate_uncalibrated_decisions(tv):
    p1, p2, m, truth = tv.synthetic_pair(3000, 3000, 400, 120, seed=8)
    g = tv.estimate_uncalibrated(p1, p2, m, seed=3)
    assert g.config == tv.UNCALIBRATED
    inl = set(map(tuple, g.inlier_matches.tolist()))
    true = set(map(tuple, m[truth].tolist()))
    assert len(inl & true) >= 0.98 * len(true) and len(inl - true) <= 0.03 * len(inl)
    assert np.all(np.diff(g.inlier_matches[:, 0].astype(np.int64)) > 0)       # order of the match list is kept
    q1, q2, m2, truth2 = tv.synthetic_pair(3000, 3000, 400, 120, seed=9, planar=True)
    h = tv.estimate_uncalibrated(q1, q2, m2, seed=3)
    assert h.config == tv.PLANAR_OR_PANORAMIC and h.num_inliers_H > 0.8 * h.num_inliers_F
    assert tv.estimate_uncalibrated(p1, p2, m[:14]).config == tv.DEGENERATE    # fewer than min_num_inliers matches
    assert tv.estimate_uncalibrated(p1, p2, m, tv.Options(min_num_inliers=1000)).config == tv.DEGENERATE
    # two seeds agree statistically (the criterion the GPU verifier is held to)
    g2 = tv.estimate_uncalibrated(p1, p2, m, seed=4)
    inl2 = set(map(tuple, g2.inlier_matches.tolist()))
    assert len(inl & inl2) / len(inl | inl2) >= 0.95


def test_watermark_detection_with_the_references_dummy_cameras(tv):
    # TwoViewGeometry::DetectWatermark: the inliers are one pure image translation.  With default-constructed cameras
    # (width = height = 0, sequential_matching.cc:89) the border-region condition holds for every inlier.
    p1, p2, m, truth = tv.synthetic_shift_pair(3000, 3000, 500, 150, seed=3)
    g = tv.estimate_uncalibrated(p1, p2, m, seed=1)
    assert g.config == tv.WATERMARK and 495 <= len(g.inlier_matches) <= 515   # the inlier matches are still reported
    assert tv.estimate_uncalibrated(p1, p2, m, tv.Options(detect_watermark=False), seed=1).config == tv.PLANAR_OR_PANORAMIC
    # a general scene is not a translation, whatever the seed
    q1, q2, m2, _ = tv.synthetic_pair(3000, 3000, 400, 120, seed=8)
    assert all(tv.estimate_uncalibrated(q1, q2, m2, seed=s).config == tv.UNCALIBRATED for s in range(3))
    # the test itself: 75 % of the correspondences on one translation passes, 60 % does not; real cameras change the
    # border condition (points in the image centre do not count)
    rng = np.random.default_rng(0)
    x1 = rng.uniform(200, 2800, size=(200, 2))
    x2 = x1 + np.array([25.0, -40.0])
    x2[150:] += rng.uniform(30, 300, size=(50, 2))
    assert tv.detect_watermark(x1, x2, tv.Options(), np.random.default_rng(1))
    x2[120:150] += rng.uniform(30, 300, size=(30, 2))
    assert not tv.detect_watermark(x1, x2, tv.Options(), np.random.default_rng(1))
    x2 = x1 + np.array([25.0, -40.0])
    assert not tv.detect_watermark(x1, x2, tv.Options(), np.random.default_rng(1), size1=(3000, 3000), size2=(3000, 3000))
    assert tv.translation_sq(np.array([25.0, -40.0]), x1, x2).max() < 1e-18
    (t,) = tv.translation_estimate(x1, x2)
    assert np.allclose(t, [25.0, -40.0])


def test_estimate_multiple(tv):
    # TwoViewGeometry::EstimateMultiple (multiple_models, sequential_matching.cc:94-96)
    p1, p2, m, group = tv.synthetic_two_motion_pair(3000, 3000, 400, 250, 120, seed=7)
    g = tv.estimate_multiple(p1, p2, m, seed=1)
    assert g.config == tv.MULTIPLE and not g.F.any() and not g.H.any()
    inl = set(map(tuple, g.inlier_matches.tolist()))
    for k in (0, 1):
        grp = set(map(tuple, m[group == k].tolist()))
        assert len(inl & grp) >= 0.97 * len(grp)
    assert len(inl & set(map(tuple, m[group == -1].tolist()))) <= 6
    single = tv.estimate_uncalibrated(p1, p2, m, seed=1)                      # one model explains one motion only
    assert single.config == tv.UNCALIBRATED and len(single.inlier_matches) < 0.7 * len(inl)
    # one motion: the single geometry is returned as it is
    q1, q2, m2, truth = tv.synthetic_pair(3000, 3000, 400, 120, seed=8)
    one = tv.estimate_multiple(q1, q2, m2, seed=3)
    assert one.config == tv.UNCALIBRATED and one.F.any() and 395 <= len(one.inlier_matches) <= 410
    # a watermark is skipped (multiple_ignore_watermark): nothing else left -> DEGENERATE, no inlier matches
    w1, w2, mw, _ = tv.synthetic_shift_pair(3000, 3000, 300, 60, seed=5)
    none = tv.estimate_multiple(w1, w2, mw, seed=1)
    assert none.config == tv.DEGENERATE and len(none.inlier_matches) == 0
    kept = tv.estimate_multiple(w1, w2, mw, tv.Options(multiple_ignore_watermark=False), seed=1)
    assert kept.config == tv.WATERMARK and len(kept.inlier_matches) >= 295
