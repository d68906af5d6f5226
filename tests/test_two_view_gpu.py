"""GPU two-view geometry verification against the CPU restatement of COLMAP's uncalibrated
F / H LORANSAC (oracle/two_view_oracle.py).  The contract is STATISTICAL, not bit parity -- COLMAP itself samples
from a history-dependent thread-local PRNG; the criterion (DESIGN.md "Two-view geometry"):

  * configuration: identical on scenes that are clearly general (UNCALIBRATED) or clearly planar (PLANAR_OR_PANORAMIC),
    DEGENERATE for fewer than min_num_inliers matches;
  * inlier set: intersection-over-union with the oracle's inlier matches >= 0.95 and inlier counts within 3 %;
  * model: the GPU's F explains the oracle's inliers (median squared Sampson error < 1 px^2, 99 % within max_error^2);
  * ground truth: >= 97 % of the planted true correspondences are inliers, <= 3 % of the inliers are planted outliers.
"""
import numpy as np
import pytest

from scanner_colmap_b200 import SiftMatcher, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tv(built):
    from oracle import two_view_oracle
    return two_view_oracle


def _iou(a, b):
    sa, sb = set(map(tuple, a.tolist())), set(map(tuple, b.tolist()))
    return len(sa & sb) / max(1, len(sa | sb))


def _scene(tv, seed, n_true, n_false, planar):
    """Descriptors that make the matcher return the planted match list: true and false correspondences get identical
    descriptors on both sides, everything else is clutter."""
    n1 = n2 = 2048
    p1, p2, m, truth = tv.synthetic_pair(n1, n2, n_true, n_false, seed, planar=planar)
    d1 = synth.make_image(7000 + 2 * seed, n1, shared_frac=0.0)
    d2 = synth.make_image(7001 + 2 * seed, n2, shared_frac=0.0)
    d2[m[:, 1]] = d1[m[:, 0]]
    return p1, p2, m, truth, d1, d2


@pytest.mark.parametrize("planar", [False, True])
def test_verification_matches_the_oracle_statistically(tv, planar):
    scenes = [_scene(tv, seed, 900, 300, planar) for seed in (11, 12, 13)]
    with SiftMatcher() as mt:
        pairs = []
        for k, (p1, p2, m, truth, d1, d2) in enumerate(scenes):
            mt.put_images([2 * k, 2 * k + 1], [d1, d2])
            mt.put_keypoints(2 * k, np.concatenate([p1, np.zeros((len(p1), 4), np.float32)], axis=1))   # FeatureKeypoint rows
            mt.put_keypoints(2 * k + 1, p2)                                                                # packed (x, y)
            pairs.append((2 * k, 2 * k + 1))
        with mt.match_pairs_result(np.array(pairs, dtype=np.uint32)) as r:
            r.verify(seed=5)
            for k, (p1, p2, m, truth, d1, d2) in enumerate(scenes):
                got_matches = r.matches(k)
                planted = set(map(tuple, m.tolist()))                       # (a few planted rows fail the distance test)
                assert set(map(tuple, got_matches.tolist())) <= planted and len(got_matches) >= 0.98 * len(m)
                keep = np.array([tuple(x) in set(map(tuple, got_matches.tolist())) for x in m.tolist()])
                m, truth = m[keep], truth[keep]
                assert np.array_equal(got_matches, m)
                want = tv.estimate_uncalibrated(p1, p2, m, seed=1)
                g, inl = r.tvg(k), r.inliers(k)
                assert g["config"] == want.config == (tv.PLANAR_OR_PANORAMIC if planar else tv.UNCALIBRATED)
                assert _iou(inl, want.inlier_matches) >= 0.95
                assert abs(len(inl) - len(want.inlier_matches)) <= 0.03 * len(want.inlier_matches)
                assert np.all(np.diff(inl[:, 0].astype(np.int64)) > 0)     # ascending idx1, like the match list
                x1, x2 = p1[want.inlier_matches[:, 0]].astype(np.float64), p2[want.inlier_matches[:, 1]].astype(np.float64)
                res = tv.sampson_sq(g["F"], x1, x2)
                assert np.median(res) < 1.0 and np.mean(res <= 16.0) >= 0.99
                true_set = set(map(tuple, m[truth].tolist()))
                got_set = set(map(tuple, inl.tolist()))
                assert len(got_set & true_set) >= 0.97 * len(true_set)
                assert len(got_set - true_set) <= 0.03 * len(got_set)
                if planar:
                    assert g["num_inliers_H"] > 0.8 * g["num_inliers_F"]
                    hres = tv.transfer_sq(g["H"], p1[m[truth][:, 0]].astype(np.float64), p2[m[truth][:, 1]].astype(np.float64))
                    assert np.mean(hres <= 16.0) >= 0.95


def test_too_few_matches_are_degenerate_and_options_are_honoured(tv):
    p1, p2, m, truth, d1, d2 = _scene(tv, 21, 10, 0, False)          # 10 matches < min_num_inliers = 15
    q1, q2, m2, truth2, e1, e2 = _scene(tv, 22, 40, 20, False)
    with SiftMatcher() as mt:
        mt.put_images([0, 1, 2, 3], [d1, d2, e1, e2])
        for i, p in enumerate((p1, p2, q1, q2)):
            mt.put_keypoints(i, p)
        with mt.match_pairs_result(np.array([[0, 1], [2, 3]], dtype=np.uint32)) as r:
            with pytest.raises(Exception):
                r.tvg(0)                                            # not verified yet
            r.verify()
            assert len(r.matches(0)) < 15 and r.tvg(0)["config"] == tv.DEGENERATE and len(r.inliers(0)) == 0
            g = r.tvg(1)
            assert g["config"] == tv.UNCALIBRATED and 36 <= len(r.inliers(1)) <= 46
            assert g["trials_F"] >= 30 and g["trials_F"] % 128 == 0
            r.verify(min_num_inliers=100)                           # neither model can reach 100 inliers
            assert r.tvg(1)["config"] == tv.DEGENERATE and len(r.inliers(1)) == 0
        with mt.match_pairs_result(np.array([[0, 1]], dtype=np.uint32)) as r2:
            mt.put_image(1, d2)                                     # descriptors replaced: the keypoints are gone
            with pytest.raises(Exception):
                r2.verify()
