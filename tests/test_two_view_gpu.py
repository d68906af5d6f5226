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


def _planted(tv, p1, p2, m, seed):
    """Descriptors for which the matcher returns the planted match list ``m`` (see _scene)."""
    d1 = synth.make_image(9000 + 2 * seed, len(p1), shared_frac=0.0)
    d2 = synth.make_image(9001 + 2 * seed, len(p2), shared_frac=0.0)
    d2[m[:, 1]] = d1[m[:, 0]]
    return d1, d2


def _surviving(got_matches, m, *labels):
    """The planted rows the matcher returned (a few fail the distance test), with their labels."""
    got = set(map(tuple, got_matches.tolist()))
    keep = np.array([tuple(x) in got for x in m.tolist()])
    assert got <= set(map(tuple, m.tolist())) and keep.sum() >= 0.98 * len(m)
    return (m[keep],) + tuple(l[keep] for l in labels)


def test_watermark_configuration(tv):
    # TwoViewGeometry::DetectWatermark with the reference's dummy cameras: inliers that are one image translation
    p1, p2, m, truth = tv.synthetic_shift_pair(2048, 2048, 600, 200, seed=31)
    q1, q2, m2, truth2 = tv.synthetic_pair(2048, 2048, 600, 200, seed=32)
    d1, d2 = _planted(tv, p1, p2, m, 31)
    e1, e2 = _planted(tv, q1, q2, m2, 32)
    with SiftMatcher() as mt:
        mt.put_images([0, 1, 2, 3], [d1, d2, e1, e2])
        for i, p in enumerate((p1, p2, q1, q2)):
            mt.put_keypoints(i, p)
        with mt.match_pairs_result(np.array([[0, 1], [2, 3]], dtype=np.uint32)) as r:
            r.verify(seed=3)
            m, truth = _surviving(r.matches(0), m, truth)
            want = tv.estimate_uncalibrated(p1, p2, m, seed=1)
            g, inl = r.tvg(0), r.inliers(0)
            assert g["config"] == want.config == tv.WATERMARK
            assert _iou(inl, want.inlier_matches) >= 0.95                    # the inlier matches are still those of F
            assert len(set(map(tuple, inl.tolist())) & set(map(tuple, m[truth].tolist()))) >= 0.97 * truth.sum()
            assert r.tvg(1)["config"] == tv.UNCALIBRATED                     # a general scene is not a translation
            r.verify(seed=3, detect_watermark=False)
            assert r.tvg(0)["config"] == tv.PLANAR_OR_PANORAMIC and r.tvg(1)["config"] == tv.UNCALIBRATED


def test_multiple_models(tv):
    # TwoViewGeometry::EstimateMultiple (siftMatchingArgs.multiple_models, sequential_matching.cc:94-96)
    p1, p2, m, group = tv.synthetic_two_motion_pair(2048, 2048, 400, 250, 120, seed=41)
    q1, q2, m2, truth2 = tv.synthetic_pair(2048, 2048, 500, 150, seed=42)
    w1, w2, mw, truthw = tv.synthetic_shift_pair(2048, 2048, 300, 80, seed=43)
    ds = [_planted(tv, a, b, mm, s) for a, b, mm, s in ((p1, p2, m, 41), (q1, q2, m2, 42), (w1, w2, mw, 43))]
    with SiftMatcher() as mt:
        mt.put_images(list(range(6)), [d for pair in ds for d in pair])
        for i, p in enumerate((p1, p2, q1, q2, w1, w2)):
            mt.put_keypoints(i, np.asarray(p, dtype=np.float32))
        with mt.match_pairs_result(np.array([[0, 1], [2, 3], [4, 5]], dtype=np.uint32)) as r:
            r.verify(seed=2, multiple_models=True)
            # two independently moving groups: MULTIPLE, the inlier matches of both models, F / H left at zero
            m, group = _surviving(r.matches(0), m, group)
            want = tv.estimate_multiple(p1, p2, m, seed=1)
            g, inl = r.tvg(0), r.inliers(0)
            assert g["config"] == want.config == tv.MULTIPLE and not g["F"].any() and not g["H"].any()
            got = set(map(tuple, inl.tolist()))
            assert len(got) == len(inl)                                      # no match is reported twice
            for k in (0, 1):
                grp = set(map(tuple, m[group == k].tolist()))
                assert len(got & grp) >= 0.97 * len(grp)
            assert len(got & set(map(tuple, m[group == -1].tolist()))) <= 8
            assert _iou(inl, want.inlier_matches) >= 0.95
            # one motion: the single geometry as it is (same contract as Estimate)
            m2, truth2 = _surviving(r.matches(1), m2, truth2)
            g1, inl1 = r.tvg(1), r.inliers(1)
            assert g1["config"] == tv.UNCALIBRATED and g1["F"].any()
            assert _iou(inl1, tv.estimate_multiple(q1, q2, m2, seed=1).inlier_matches) >= 0.95
            assert np.all(np.diff(inl1[:, 0].astype(np.int64)) > 0)
            # a watermark is skipped and nothing else is left: DEGENERATE, no inlier matches
            assert r.tvg(2)["config"] == tv.DEGENERATE and len(r.inliers(2)) == 0
            # the same result object verified again without the option: Estimate's answers
            r.verify(seed=2)
            assert r.tvg(0)["config"] == tv.UNCALIBRATED and len(r.inliers(0)) < 0.7 * len(got)
            assert r.tvg(2)["config"] == tv.WATERMARK and len(r.inliers(2)) >= 290
