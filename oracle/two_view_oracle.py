"""oracle/two_view_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy, double precision, SVD-based like COLMAP's Eigen code) of the geometric verification the
reference op runs on every matched pair:

    verifyTwoViewGeometry                      /root/reference/integration/op_cpp/sequential_matching.cc:84-101
      colmap::Camera camera1, camera2;         (dummy cameras: no prior focal length)                     :89
      two_view_geometry.Estimate(camera1, points1, camera2, points2, matches, options)                   :98
    options from siftFeatureMatchingArgs        sequential_matching.cc:63-75, defaults colmap.proto:24-44

With cameras that have no prior focal length COLMAP 3.5's TwoViewGeometry::Estimate takes EstimateUncalibrated [ext]:

    F  <- LORANSAC<FundamentalMatrixSevenPointEstimator, FundamentalMatrixEightPointEstimator>   (squared Sampson error)
    H  <- LORANSAC<HomographyMatrixEstimator, HomographyMatrixEstimator>                         (squared transfer error)
    DEGENERATE if neither succeeded or both have < min_num_inliers inliers
    config = PLANAR_OR_PANORAMIC if inliers(H) / inliers(F) > max_H_inlier_ratio (0.8) else UNCALIBRATED
    inlier_matches = the matches inside the F model's inlier mask
    config = WATERMARK if DetectWatermark(...) (Options::detect_watermark, default true): at least
             watermark_min_inlier_ratio (0.7) of the inliers lie in the border region AND fit one pure image translation
             (LORANSAC<TranslationTransformEstimator<2>> with min_inlier_ratio = 0.7).  With the reference's dummy
             cameras (width = height = 0, :89) the "inner" box is the single point (0, 0): every inlier counts as border,
             so only the translational test decides -- a pair whose inliers are one image shift is reported WATERMARK.
    multiple_models (colmap.proto:45, default false) selects EstimateMultiple instead: Estimate on the matches, remove
             the inliers, repeat until DEGENERATE; watermark models are skipped (multiple_ignore_watermark, default
             true); none -> DEGENERATE, one -> that geometry, several -> config MULTIPLE with all inlier matches.

[ext]: COLMAP is not vendored in the reference and not installable here; the algorithm below is
restated from the published COLMAP 3.5 sources (src/estimators/two_view_geometry.cc, src/optim/loransac.h,
src/estimators/fundamental_matrix.cc, homography_matrix.cc, src/optim/random_sampler.cc).  PARITY UNPINNED, and
inherently STATISTICAL: COLMAP draws its samples from a thread-local PRNG whose state depends on everything the
thread did before, so not even two runs of the reference agree bit for bit.  The GPU verifier is therefore checked
against this restatement with the criterion written down in DESIGN.md ("Two-view geometry"): same configuration,
inlier-set IoU, and agreement of the estimated F on the oracle's inliers -- not byte equality.

Only tests/ may import this module.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

# colmap::TwoViewGeometry::ConfigurationType
UNDEFINED, DEGENERATE, CALIBRATED, UNCALIBRATED, PLANAR, PANORAMIC, PLANAR_OR_PANORAMIC, WATERMARK, MULTIPLE = range(9)


@dataclass
class Options:
    """TwoViewGeometry::Options as the reference fills it (sequential_matching.cc:63-75; colmap.proto:24-44)."""
    min_num_inliers: int = 15
    max_error: float = 4.0
    confidence: float = 0.999
    min_num_trials: int = 30
    max_num_trials: int = 10000
    min_inlier_ratio: float = 0.25
    max_H_inlier_ratio: float = 0.8          # COLMAP default, not exposed by the reference's proto
    dyn_num_trials_multiplier: float = 3.0   # RANSACOptions default
    detect_watermark: bool = True            # TwoViewGeometry::Options defaults, not exposed by the reference's proto
    watermark_min_inlier_ratio: float = 0.7
    multiple_ignore_watermark: bool = True


@dataclass
class Report:
    success: bool = False
    num_trials: int = 0
    num_inliers: int = 0
    residual_sum: float = float("inf")
    model: Optional[np.ndarray] = None
    inlier_mask: np.ndarray = field(default_factory=lambda: np.zeros(0, bool))


@dataclass
class TwoViewGeometry:
    config: int = UNDEFINED
    F: np.ndarray = field(default_factory=lambda: np.zeros((3, 3)))
    H: np.ndarray = field(default_factory=lambda: np.zeros((3, 3)))
    inlier_matches: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.uint32))
    num_inliers_F: int = 0
    num_inliers_H: int = 0
    trials_F: int = 0
    trials_H: int = 0


# ------------------------------------------------------------------------------------------------- estimators
def _center_and_normalize(p: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """CenterAndNormalizeImagePoints: centroid to the origin, RMS distance to sqrt(2)."""
    c = p.mean(axis=0)
    rms = np.sqrt(((p - c) ** 2).sum(axis=1).mean())
    s = np.sqrt(2.0) / rms if rms > 0 else 1.0
    T = np.array([[s, 0, -s * c[0]], [0, s, -s * c[1]], [0, 0, 1.0]])
    return (p - c) * s, T


def _epipolar_rows(x1: np.ndarray, x2: np.ndarray) -> np.ndarray:
    """Rows of the linear system x2^T F x1 = 0 in the row-major entries of F."""
    o = np.ones(len(x1))
    return np.stack([x2[:, 0] * x1[:, 0], x2[:, 0] * x1[:, 1], x2[:, 0], x2[:, 1] * x1[:, 0], x2[:, 1] * x1[:, 1],
                     x2[:, 1], x1[:, 0], x1[:, 1], o], axis=1)


def fundamental_seven_point(x1: np.ndarray, x2: np.ndarray) -> List[np.ndarray]:
    """FundamentalMatrixSevenPointEstimator::Estimate: the two-dimensional null space of the 7 x 9 system, then the
    real roots of det(l * F1 + (1 - l) * F2) = 0: one or three models."""
    A = _epipolar_rows(x1, x2)
    _, _, vt = np.linalg.svd(A)
    F1, F2 = vt[7].reshape(3, 3), vt[8].reshape(3, 3)
    # the cubic through four evaluations (exact for a cubic)
    ls = np.array([0.0, 1.0, -1.0, 2.0])
    dets = np.array([np.linalg.det(l * F1 + (1 - l) * F2) for l in ls])
    coef = np.linalg.solve(np.vander(ls, 4), dets)
    if abs(coef[0]) < 1e-14 * max(1.0, np.abs(coef).max()):
        roots = np.roots(coef[1:]) if np.abs(coef[1:]).max() > 0 else np.zeros(0)
    else:
        roots = np.roots(coef)
    models = []
    for r in roots:
        if abs(r.imag) > 1e-10:
            continue
        F = r.real * F1 + (1 - r.real) * F2
        if abs(F[2, 2]) > 1e-12:
            F = F / F[2, 2]
        models.append(F)
    return models


def fundamental_eight_point(x1: np.ndarray, x2: np.ndarray) -> List[np.ndarray]:
    """FundamentalMatrixEightPointEstimator::Estimate: normalised eight-point algorithm, rank 2 enforced."""
    n1, T1 = _center_and_normalize(x1)
    n2, T2 = _center_and_normalize(x2)
    _, _, vt = np.linalg.svd(_epipolar_rows(n1, n2))
    Fh = vt[-1].reshape(3, 3)
    u, s, v = np.linalg.svd(Fh)
    s[2] = 0.0
    return [T2.T @ (u @ np.diag(s) @ v) @ T1]


def homography_dlt(x1: np.ndarray, x2: np.ndarray) -> List[np.ndarray]:
    """HomographyMatrixEstimator::Estimate: normalised DLT, x2 ~ H x1."""
    n1, T1 = _center_and_normalize(x1)
    n2, T2 = _center_and_normalize(x2)
    n = len(x1)
    A = np.zeros((2 * n, 9))
    A[0::2, 0:2] = -n1
    A[0::2, 2] = -1
    A[0::2, 6:8] = n1 * n2[:, :1]
    A[0::2, 8] = n2[:, 0]
    A[1::2, 3:5] = -n1
    A[1::2, 5] = -1
    A[1::2, 6:8] = n1 * n2[:, 1:2]
    A[1::2, 8] = n2[:, 1]
    _, _, vt = np.linalg.svd(A)
    Hh = vt[-1].reshape(3, 3)
    return [np.linalg.inv(T2) @ Hh @ T1]


def sampson_sq(F: np.ndarray, x1: np.ndarray, x2: np.ndarray) -> np.ndarray:
    """ComputeSquaredSampsonError."""
    h1 = np.concatenate([x1, np.ones((len(x1), 1))], axis=1)
    h2 = np.concatenate([x2, np.ones((len(x2), 1))], axis=1)
    Fx1 = h1 @ F.T
    Ftx2 = h2 @ F
    num = (h2 * Fx1).sum(axis=1) ** 2
    den = Fx1[:, 0] ** 2 + Fx1[:, 1] ** 2 + Ftx2[:, 0] ** 2 + Ftx2[:, 1] ** 2
    with np.errstate(divide="ignore", invalid="ignore"):
        r = num / den
    return np.where(den > 0, r, np.inf)


def transfer_sq(H: np.ndarray, x1: np.ndarray, x2: np.ndarray) -> np.ndarray:
    """HomographyMatrixEstimator::Residuals: squared forward transfer error."""
    h1 = np.concatenate([x1, np.ones((len(x1), 1))], axis=1)
    p = h1 @ H.T
    with np.errstate(divide="ignore", invalid="ignore"):
        q = p[:, :2] / p[:, 2:3]
    d = ((q - x2) ** 2).sum(axis=1)
    return np.where(np.isfinite(d), d, np.inf)


# ------------------------------------------------------------------------------------------------- LORANSAC
def compute_num_trials(num_inliers: int, num_samples: int, confidence: float, min_samples: int, multiplier: float) -> int:
    """RANSAC::ComputeNumTrials."""
    ratio = num_inliers / float(num_samples)
    nom = 1.0 - confidence
    if nom <= 0:
        return 2 ** 62
    denom = 1.0 - ratio ** min_samples
    if denom <= 0:
        return 1
    if denom >= 1.0:
        return 2 ** 62
    return int(np.ceil(np.log(nom) / np.log(denom) * multiplier))


def _better(n, s, bn, bs) -> bool:
    """InlierSupportMeasurer::Compare: more inliers, or as many with a smaller residual sum."""
    return n > bn or (n == bn and s < bs)


def loransac(x1: np.ndarray, x2: np.ndarray, minimal, k_min: int, local, k_local: int, residuals, opt: Options,
             rng: np.random.Generator) -> Report:
    """LORANSAC<..>::Estimate (src/optim/loransac.h) with COLMAP's RandomSampler (a fresh shuffle per trial)."""
    rep = Report()
    n = len(x1)
    rep.inlier_mask = np.zeros(n, bool)
    if n < k_min:
        return rep
    max_res = opt.max_error * opt.max_error
    cap = compute_num_trials(int(opt.min_inlier_ratio * 100000), 100000, opt.confidence, k_min, opt.dyn_num_trials_multiplier)
    max_trials = min(opt.max_num_trials, cap)
    dyn_max = max_trials
    best_n, best_s, best_model = 0, float("inf"), None
    trials = 0
    abort = False
    while trials < max_trials and not abort:
        idx = rng.permutation(n)[:k_min]
        for model in minimal(x1[idx], x2[idx]):
            r = residuals(model, x1, x2)
            inl = r <= max_res
            cn, cs = int(inl.sum()), float(r[inl].sum())
            if _better(cn, cs, best_n, best_s):
                best_n, best_s, best_model = cn, cs, model
                if cn > k_min and cn >= k_local:             # local optimisation on the new best model's inliers
                    for lm in local(x1[inl], x2[inl]):
                        lr = residuals(lm, x1, x2)
                        linl = lr <= max_res
                        ln, lsum = int(linl.sum()), float(lr[linl].sum())
                        if _better(ln, lsum, best_n, best_s):
                            best_n, best_s, best_model = ln, lsum, lm
                dyn_max = compute_num_trials(best_n, n, opt.confidence, k_min, opt.dyn_num_trials_multiplier)
            if trials >= dyn_max and trials >= opt.min_num_trials:
                abort = True
                break
        trials += 1
    rep.num_trials = trials
    rep.num_inliers, rep.residual_sum, rep.model = best_n, best_s, best_model
    if best_model is None or best_n < k_min:
        return rep
    rep.success = True
    rep.inlier_mask = residuals(best_model, x1, x2) <= max_res
    return rep


def estimate_uncalibrated(points1: np.ndarray, points2: np.ndarray, matches: np.ndarray, opt: Options = Options(),
                          seed: int = 0) -> TwoViewGeometry:
    """TwoViewGeometry::EstimateUncalibrated on (x, y) keypoint positions and FeatureMatches (uint32 [m, 2])."""
    tvg = TwoViewGeometry()
    matches = np.asarray(matches, dtype=np.uint32).reshape(-1, 2)
    if len(matches) < opt.min_num_inliers:
        tvg.config = DEGENERATE
        return tvg
    x1 = np.asarray(points1, dtype=np.float64)[matches[:, 0]]
    x2 = np.asarray(points2, dtype=np.float64)[matches[:, 1]]
    rng = np.random.default_rng(seed)
    fr = loransac(x1, x2, fundamental_seven_point, 7, fundamental_eight_point, 8, sampson_sq, opt, rng)
    hr = loransac(x1, x2, homography_dlt, 4, homography_dlt, 4, transfer_sq, opt, rng)
    tvg.num_inliers_F, tvg.num_inliers_H, tvg.trials_F, tvg.trials_H = fr.num_inliers, hr.num_inliers, fr.num_trials, hr.num_trials
    if fr.model is not None:
        tvg.F = fr.model
    if hr.model is not None:
        tvg.H = hr.model
    if (not fr.success and not hr.success) or (fr.num_inliers < opt.min_num_inliers and hr.num_inliers < opt.min_num_inliers):
        tvg.config = DEGENERATE
        return tvg
    ratio = hr.num_inliers / fr.num_inliers if fr.num_inliers else float("inf")
    tvg.config = PLANAR_OR_PANORAMIC if ratio > opt.max_H_inlier_ratio else UNCALIBRATED
    tvg.inlier_matches = matches[fr.inlier_mask]
    if opt.detect_watermark and detect_watermark(x1[fr.inlier_mask], x2[fr.inlier_mask], opt, rng):
        tvg.config = WATERMARK
    return tvg


def translation_estimate(x1: np.ndarray, x2: np.ndarray) -> List[np.ndarray]:
    """TranslationTransformEstimator<2>::Estimate (kMinNumSamples = 1): the mean displacement."""
    return [(x2 - x1).mean(axis=0)]


def translation_sq(t: np.ndarray, x1: np.ndarray, x2: np.ndarray) -> np.ndarray:
    d = x2 - (x1 + t)
    return (d * d).sum(axis=1)


def detect_watermark(x1_in: np.ndarray, x2_in: np.ndarray, opt: Options, rng: np.random.Generator,
                     size1=(0.0, 0.0), size2=(0.0, 0.0), border: float = 0.1) -> bool:
    """TwoViewGeometry::DetectWatermark on the inlier correspondences.  ``size*`` are the cameras' (width, height):
    the reference passes default-constructed cameras, i.e. (0, 0), which makes the inner box the point (0, 0)."""
    n = len(x1_in)
    if n == 0:
        return False

    def outside(p, size):
        d = border * float(np.hypot(size[0], size[1]))
        lo_x, lo_y, hi_x, hi_y = d, d, size[0] - d, size[1] - d
        inside = (p[:, 0] >= lo_x) & (p[:, 0] <= hi_x) & (p[:, 1] >= lo_y) & (p[:, 1] <= hi_y)
        return ~inside

    in_border = int((outside(x1_in, size1) & outside(x2_in, size2)).sum())
    if in_border / n < opt.watermark_min_inlier_ratio:
        return False
    from dataclasses import replace
    wopt = replace(opt, min_inlier_ratio=opt.watermark_min_inlier_ratio)
    rep = loransac(x1_in, x2_in, translation_estimate, 1, translation_estimate, 1, translation_sq, wopt, rng)
    return rep.num_inliers / n >= opt.watermark_min_inlier_ratio


def estimate_multiple(points1: np.ndarray, points2: np.ndarray, matches: np.ndarray, opt: Options = Options(),
                      seed: int = 0) -> TwoViewGeometry:
    """TwoViewGeometry::EstimateMultiple (multiple_models = true, sequential_matching.cc:94-96)."""
    remaining = np.asarray(matches, dtype=np.uint32).reshape(-1, 2)
    found: List[TwoViewGeometry] = []
    k = 0
    while True:
        g = estimate_uncalibrated(points1, points2, remaining, opt, seed + 7919 * k)
        k += 1
        if g.config == DEGENERATE:
            break
        if not (opt.multiple_ignore_watermark and g.config == WATERMARK):
            found.append(g)
        inl = {(int(a), int(b)) for a, b in g.inlier_matches}                    # ExtractOutlierMatches
        remaining = np.asarray([mm for mm in remaining if (int(mm[0]), int(mm[1])) not in inl], dtype=np.uint32).reshape(-1, 2)
    out = TwoViewGeometry()
    if not found:
        out.config = DEGENERATE
    elif len(found) == 1:
        out = found[0]
    else:
        out.config = MULTIPLE
        out.inlier_matches = np.concatenate([g.inlier_matches for g in found], axis=0)
    return out


# ------------------------------------------------------------------------------------------------- synthetic scenes
def synthetic_pair(n1: int, n2: int, n_true: int, n_false: int, seed: int, planar: bool = False, noise_px: float = 0.7,
                   size=(4000.0, 3000.0)):
    """Two views of a random 3-D scene (or of a plane): keypoint positions of both images and a match list with
    ``n_true`` correct correspondences (pixel noise ``noise_px``) and ``n_false`` random ones, shuffled like a real
    matcher's output (ascending idx1).  Returns (points1 [n1, 2], points2 [n2, 2], matches uint32 [m, 2], truth mask)."""
    rng = np.random.default_rng(seed)
    W, Hh = size
    f = 1.2 * W
    K = np.array([[f, 0, W / 2], [0, f, Hh / 2], [0, 0, 1.0]])
    a = 0.12 * rng.standard_normal(3)
    th = np.linalg.norm(a)
    kx = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]]) / max(th, 1e-12)
    R = np.eye(3) + np.sin(th) * kx + (1 - np.cos(th)) * kx @ kx
    t = np.array([1.0, 0.1, 0.05]) * (0.6 + 0.4 * rng.random())
    pts1, pts2 = [], []
    while len(pts1) < n_true:
        X = np.array([rng.uniform(-4, 4), rng.uniform(-3, 3), 8.0 if planar else rng.uniform(5, 14)])
        if planar:
            X[2] = 8.0 + 0.15 * X[0]
        u1 = K @ X
        u2 = K @ (R @ X + t)
        u1, u2 = u1[:2] / u1[2], u2[:2] / u2[2]
        if 0 <= u1[0] < W and 0 <= u1[1] < Hh and 0 <= u2[0] < W and 0 <= u2[1] < Hh:
            pts1.append(u1)
            pts2.append(u2)
    p1 = np.empty((n1, 2))
    p2 = np.empty((n2, 2))
    p1[:] = rng.uniform([0, 0], [W, Hh], size=(n1, 2))
    p2[:] = rng.uniform([0, 0], [W, Hh], size=(n2, 2))
    i1 = rng.choice(n1, size=n_true + n_false, replace=False)
    i2 = rng.choice(n2, size=n_true + n_false, replace=False)
    if n_true:
        p1[i1[:n_true]] = np.asarray(pts1) + noise_px * rng.standard_normal((n_true, 2))
        p2[i2[:n_true]] = np.asarray(pts2) + noise_px * rng.standard_normal((n_true, 2))
    m = np.stack([i1, i2], axis=1).astype(np.uint32)
    truth = np.zeros(len(m), bool)
    truth[:n_true] = True
    order = np.argsort(m[:, 0], kind="stable")
    return p1.astype(np.float32), p2.astype(np.float32), m[order], truth[order]


def synthetic_shift_pair(n1: int, n2: int, n_true: int, n_false: int, seed: int, shift=(37.0, -12.0), noise_px: float = 0.7,
                         size=(4000.0, 3000.0)):
    """Like synthetic_pair, but the correct correspondences are one pure image translation (what a watermark, or a
    distant scene under a small pan, looks like).  Same return values."""
    rng = np.random.default_rng(seed)
    W, Hh = size
    p1 = rng.uniform([0, 0], [W, Hh], size=(n1, 2))
    p2 = rng.uniform([0, 0], [W, Hh], size=(n2, 2))
    i1 = rng.choice(n1, size=n_true + n_false, replace=False)
    i2 = rng.choice(n2, size=n_true + n_false, replace=False)
    base = rng.uniform([100, 100], [W - 100, Hh - 100], size=(n_true, 2))
    p1[i1[:n_true]] = base + noise_px * rng.standard_normal((n_true, 2))
    p2[i2[:n_true]] = base + np.asarray(shift) + noise_px * rng.standard_normal((n_true, 2))
    m = np.stack([i1, i2], axis=1).astype(np.uint32)
    truth = np.zeros(len(m), bool)
    truth[:n_true] = True
    order = np.argsort(m[:, 0], kind="stable")
    return p1.astype(np.float32), p2.astype(np.float32), m[order], truth[order]


def synthetic_two_motion_pair(n1: int, n2: int, n_a: int, n_b: int, n_false: int, seed: int):
    """Two independently moving rigid groups in one image pair (n_a and n_b correct correspondences, each consistent
    with its own epipolar geometry) plus n_false random matches: the case EstimateMultiple exists for.  Returns
    (points1, points2, matches, group) with group 0 / 1 for the two motions and -1 for the random matches."""
    pa1, pa2, ma, ta = synthetic_pair(n1, n2, n_a, 0, seed)
    pb1, pb2, mb, tb = synthetic_pair(n1, n2, n_b, 0, seed + 1000)
    rng = np.random.default_rng(seed + 5)
    p1, p2 = pa1.copy(), pa2.copy()
    used1, used2 = set(ma[:, 0].tolist()), set(ma[:, 1].tolist())
    free1 = np.array([i for i in range(n1) if i not in used1])
    free2 = np.array([i for i in range(n2) if i not in used2])
    j1 = rng.choice(free1, size=n_b + n_false, replace=False)
    j2 = rng.choice(free2, size=n_b + n_false, replace=False)
    p1[j1[:n_b]] = pb1[mb[:, 0]]
    p2[j2[:n_b]] = pb2[mb[:, 1]]
    m = np.concatenate([ma, np.stack([j1, j2], axis=1).astype(np.uint32)], axis=0)
    group = np.concatenate([np.zeros(len(ma), int), np.ones(n_b, int), -np.ones(n_false, int)])
    order = np.argsort(m[:, 0], kind="stable")
    return p1, p2, m[order], group[order]
