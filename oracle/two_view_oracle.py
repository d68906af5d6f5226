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
    return tvg


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
