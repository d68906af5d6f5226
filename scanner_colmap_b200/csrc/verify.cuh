// Two-view geometry verification on the GPU: the step that follows the matcher in the reference op,
//   verifyTwoViewGeometry -> colmap::TwoViewGeometry::Estimate      /root/reference/integration/op_cpp/sequential_matching.cc:84-101,157-178
// With the reference's dummy cameras (no prior focal length, :89) COLMAP 3.5 takes EstimateUncalibrated [ext]:
// a fundamental matrix by LORANSAC (7-point minimal solver, normalised 8-point local optimisation, squared Sampson
// error), a homography by LORANSAC (normalised DLT, squared transfer error), then
//   DEGENERATE              fewer than min_num_inliers matches, or neither model reaches min_num_inliers
//   PLANAR_OR_PANORAMIC     inliers(H) / inliers(F) > max_H_inlier_ratio
//   UNCALIBRATED            otherwise
// and inlier_matches = the matches within max_error of F.
//
// This cannot be bit-exact against COLMAP (its samples come from a thread-local PRNG whose state depends on the
// thread's history); the contract is statistical and is written down in DESIGN.md ("Two-view geometry").
//
// One CTA of 128 threads per image pair.  A round evaluates 128 minimal samples at once (thread = sample: draw,
// solve, score every match); the best of the round goes through the local optimisation (moment matrix accumulated
// by the whole CTA, its smallest eigenvector by Jacobi rotations) and the trial budget is re-derived from the inlier
// ratio exactly like RANSAC::ComputeNumTrials -- so the number of samples drawn follows COLMAP's stopping rule in
// units of 128.  Solvers, local optimisation, every adopted model's score and the final inlier mask are computed in
// double like COLMAP's; only the bulk scoring of the 128 hypotheses of a round is single precision (residual_f).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace smb {
namespace tvg {

constexpr int kThreads = 128;
enum Config : int32_t { kUndefined = 0, kDegenerate = 1, kCalibrated = 2, kUncalibrated = 3, kPlanarOrPanoramic = 6, kWatermark = 7,
                        kMultiple = 8 };

struct Options {
  int32_t min_num_inliers;
  int32_t min_num_trials;
  int32_t max_num_trials;
  double max_error;
  double confidence;
  double min_inlier_ratio;
  double max_H_inlier_ratio;
  double watermark_min_inlier_ratio;  // TwoViewGeometry::Options default 0.7
  int32_t detect_watermark;           // TwoViewGeometry::Options default true
  int32_t pad_;
  unsigned long long seed;
};

struct Out {  // one per pair, written to pinned host memory
  int32_t config;
  int32_t num_inliers_F, num_inliers_H;
  int32_t trials_F, trials_H;
  uint32_t inlier_start, inlier_count;  // into the result's inlier buffer
  uint32_t pad_;
  double F[9], H[9];                    // row-major
};

__device__ __forceinline__ unsigned long long rng_next(unsigned long long& s) {  // splitmix64
  s += 0x9E3779B97F4A7C15ull;
  unsigned long long z = s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

template <int K>
__device__ __forceinline__ void sample_distinct(unsigned long long& s, uint32_t m, uint32_t (&idx)[K]) {
  for (int k = 0; k < K; ++k) {
    for (;;) {
      const uint32_t c = (uint32_t)(((rng_next(s) >> 32) * (unsigned long long)m) >> 32);
      bool dup = false;
      for (int j = 0; j < k; ++j) dup = dup || idx[j] == c;
      if (!dup) {
        idx[k] = c;
        break;
      }
    }
  }
}

// Null space of an R x 9 system by Gauss-Jordan elimination with full pivoting: basis[t], t < 9 - R.
template <int R>
__device__ bool nullspace(double (&A)[R][9], double (*basis)[9]) {
  int perm[9];
  for (int j = 0; j < 9; ++j) perm[j] = j;
  for (int r = 0; r < R; ++r) {
    int pi = r, pj = r;
    double best = 0.0;
    for (int i = r; i < R; ++i)
      for (int j = r; j < 9; ++j) {
        const double v = fabs(A[i][perm[j]]);
        if (v > best) { best = v; pi = i; pj = j; }
      }
    if (best < 1e-12) return false;
    if (pi != r)
      for (int j = 0; j < 9; ++j) { const double t = A[r][j]; A[r][j] = A[pi][j]; A[pi][j] = t; }
    { const int t = perm[r]; perm[r] = perm[pj]; perm[pj] = t; }
    const double inv = 1.0 / A[r][perm[r]];
    for (int j = 0; j < 9; ++j) A[r][j] *= inv;
    for (int i = 0; i < R; ++i) {
      if (i == r) continue;
      const double f = A[i][perm[r]];
      if (f != 0.0)
        for (int j = 0; j < 9; ++j) A[i][j] -= f * A[r][j];
    }
  }
  for (int t = 0; t < 9 - R; ++t) {
    const int fc = perm[R + t];
    for (int j = 0; j < 9; ++j) basis[t][j] = 0.0;
    basis[t][fc] = 1.0;
    for (int r = 0; r < R; ++r) basis[t][perm[r]] = -A[r][fc];
  }
  return true;
}

__device__ __forceinline__ double det3(const double* F) {
  return F[0] * (F[4] * F[8] - F[5] * F[7]) - F[1] * (F[3] * F[8] - F[5] * F[6]) + F[2] * (F[3] * F[7] - F[4] * F[6]);
}

// Real roots of a x^3 + b x^2 + c x + d (Cardano / trigonometric form, two Newton steps of polish).
__device__ int solve_cubic(double a, double b, double c, double d, double* roots) {
  const double scale = fmax(fmax(fabs(a), fabs(b)), fmax(fabs(c), fabs(d)));
  if (scale == 0.0) return 0;
  int n = 0;
  if (fabs(a) < 1e-14 * scale) {  // quadratic (or linear)
    if (fabs(b) < 1e-14 * scale) {
      if (fabs(c) < 1e-14 * scale) return 0;
      roots[0] = -d / c;
      return 1;
    }
    const double disc = c * c - 4.0 * b * d;
    if (disc < 0.0) return 0;
    const double sq = sqrt(disc), q = -0.5 * (c + (c >= 0 ? sq : -sq));
    roots[n++] = q / b;
    if (q != 0.0) roots[n++] = d / q;
    return n;
  }
  const double B = b / a, C = c / a, D = d / a;
  const double p = C - B * B / 3.0, q = 2.0 * B * B * B / 27.0 - B * C / 3.0 + D;
  const double disc = q * q / 4.0 + p * p * p / 27.0;
  if (disc > 0.0) {
    const double sq = sqrt(disc);
    roots[n++] = cbrt(-q / 2.0 + sq) + cbrt(-q / 2.0 - sq) - B / 3.0;
  } else {
    const double r = sqrt(fmax(-p / 3.0, 0.0));
    const double arg = r > 0.0 ? fmin(fmax(-q / (2.0 * r * r * r), -1.0), 1.0) : 0.0;
    const double phi = acos(arg);
    for (int k = 0; k < 3; ++k) roots[n++] = 2.0 * r * cos((phi + 2.0 * 3.14159265358979323846 * k) / 3.0) - B / 3.0;
  }
  for (int k = 0; k < n; ++k)
    for (int it = 0; it < 2; ++it) {
      const double x = roots[k], f = ((a * x + b) * x + c) * x + d, fp = (3.0 * a * x + 2.0 * b) * x + c;
      if (fp != 0.0) roots[k] = x - f / fp;
    }
  return n;
}

// squared Sampson error of (x1, y1) <-> (x2, y2) under x2^T F x1 = 0 (ComputeSquaredSampsonError)
__device__ __forceinline__ double sampson_sq(const double* F, double x1, double y1, double x2, double y2) {
  const double Fx0 = F[0] * x1 + F[1] * y1 + F[2], Fx1 = F[3] * x1 + F[4] * y1 + F[5], Fx2 = F[6] * x1 + F[7] * y1 + F[8];
  const double Ft0 = F[0] * x2 + F[3] * y2 + F[6], Ft1 = F[1] * x2 + F[4] * y2 + F[7];
  const double e = x2 * Fx0 + y2 * Fx1 + Fx2;
  const double den = Fx0 * Fx0 + Fx1 * Fx1 + Ft0 * Ft0 + Ft1 * Ft1;
  return den > 0.0 ? e * e / den : 1e300;
}
// squared forward transfer error under x2 ~ H x1 (HomographyMatrixEstimator::Residuals)
__device__ __forceinline__ double transfer_sq(const double* H, double x1, double y1, double x2, double y2) {
  const double w = H[6] * x1 + H[7] * y1 + H[8];
  if (w == 0.0) return 1e300;
  const double dx = (H[0] * x1 + H[1] * y1 + H[2]) / w - x2, dy = (H[3] * x1 + H[4] * y1 + H[5]) / w - y2;
  const double r = dx * dx + dy * dy;
  return r == r ? r : 1e300;
}

template <bool kIsF>
__device__ __forceinline__ double residual(const double* M, const float4 p) {
  return kIsF ? sampson_sq(M, p.x, p.y, p.z, p.w) : transfer_sq(M, p.x, p.y, p.z, p.w);
}

// Single-precision residuals for scoring the minimal-sample hypotheses (128 per round x every match: the bulk of the
// arithmetic).  Models are rescaled to unit largest entry first; a squared error near the 16 px^2 threshold carries a
// relative error of ~1e-4 -- a borderline match may flip for a HYPOTHESIS, never for an accepted model: every
// candidate that beats the best so far is re-scored in double before it is adopted, and so are the local models
// and the final inlier mask.
template <bool kIsF>
__device__ __forceinline__ float residual_f(const float* M, const float4 p) {
  if (kIsF) {
    const float Fx0 = M[0] * p.x + M[1] * p.y + M[2], Fx1 = M[3] * p.x + M[4] * p.y + M[5], Fx2 = M[6] * p.x + M[7] * p.y + M[8];
    const float Ft0 = M[0] * p.z + M[3] * p.w + M[6], Ft1 = M[1] * p.z + M[4] * p.w + M[7];
    const float e = p.z * Fx0 + p.w * Fx1 + Fx2;
    const float den = Fx0 * Fx0 + Fx1 * Fx1 + Ft0 * Ft0 + Ft1 * Ft1;
    return den > 0.f ? e * e / den : 3e38f;
  } else {
    const float w = M[6] * p.x + M[7] * p.y + M[8];
    const float iw = 1.f / w;
    const float dx = (M[0] * p.x + M[1] * p.y + M[2]) * iw - p.z, dy = (M[3] * p.x + M[4] * p.y + M[5]) * iw - p.w;
    const float r = dx * dx + dy * dy;
    return r == r ? r : 3e38f;  // NaN (w == 0) never counts
  }
}

// RANSAC::ComputeNumTrials
__device__ double num_trials_for(double inliers, double samples, double confidence, int k_min, double multiplier) {
  const double nom = 1.0 - confidence;
  if (nom <= 0.0) return 1e18;
  const double denom = 1.0 - pow(inliers / samples, (double)k_min);
  if (denom <= 0.0) return 1.0;
  if (denom >= 1.0) return 1e18;
  return ceil(log(nom) / log(denom) * multiplier);
}

// Similarity that moves the centroid of the selected points to the origin and their RMS distance to sqrt(2)
// (CenterAndNormalizeImagePoints): x' = s * (x - cx).
struct Norm { double cx, cy, s; };

// Smallest eigenvector of the symmetric 9 x 9 matrix in `a` (destroyed) by cyclic Jacobi rotations.
__device__ void smallest_eigenvector9(double (*a)[9], double (*v)[9], double* out) {
  for (int i = 0; i < 9; ++i)
    for (int j = 0; j < 9; ++j) v[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < 9; ++i) {
      diag += a[i][i] * a[i][i];
      for (int j = i + 1; j < 9; ++j) off += a[i][j] * a[i][j];
    }
    if (off <= 1e-30 * diag || off == 0.0) break;
    for (int p = 0; p < 8; ++p)
      for (int q = p + 1; q < 9; ++q) {
        if (a[p][q] == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 9; ++k) {
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 9; ++k) {
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 9; ++k) {
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int best = 0;
  for (int i = 1; i < 9; ++i)
    if (a[i][i] < a[best][best]) best = i;
  for (int k = 0; k < 9; ++k) out[k] = v[k][best];
}

// Closest rank-2 matrix: F - (F v3) v3^T with v3 the right singular vector of the smallest singular value,
// i.e. the smallest eigenvector of F^T F (3 x 3 Jacobi).
__device__ void enforce_rank2(double* F) {
  double a[3][3], v[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      a[i][j] = F[0 + i] * F[0 + j] + F[3 + i] * F[3 + j] + F[6 + i] * F[6 + j];
      v[i][j] = i == j ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 30; ++sweep) {
    const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    if (off <= 1e-32 * (a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2]) || off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) { const double x = a[k][p], y = a[k][q]; a[k][p] = c * x - s * y; a[k][q] = s * x + c * y; }
        for (int k = 0; k < 3; ++k) { const double x = a[p][k], y = a[q][k]; a[p][k] = c * x - s * y; a[q][k] = s * x + c * y; }
        for (int k = 0; k < 3; ++k) { const double x = v[k][p], y = v[k][q]; v[k][p] = c * x - s * y; v[k][q] = s * x + c * y; }
      }
  }
  int b = 0;
  for (int i = 1; i < 3; ++i)
    if (a[i][i] < a[b][b]) b = i;
  const double v3[3] = {v[0][b], v[1][b], v[2][b]};
  for (int r = 0; r < 3; ++r) {
    const double Fv = F[3 * r] * v3[0] + F[3 * r + 1] * v3[1] + F[3 * r + 2] * v3[2];
    for (int c = 0; c < 3; ++c) F[3 * r + c] -= Fv * v3[c];
  }
}

// M = T2^T * Mh * T1 (fundamental) or T2^-1 * Mh * T1 (homography), T = [[s,0,-s cx],[0,s,-s cy],[0,0,1]]
__device__ void denormalize(bool is_f, const double* Mh, const Norm& n1, const Norm& n2, double* M) {
  double T1[9] = {n1.s, 0, -n1.s * n1.cx, 0, n1.s, -n1.s * n1.cy, 0, 0, 1};
  double L[9];
  if (is_f) {
    const double t[9] = {n2.s, 0, 0, 0, n2.s, 0, -n2.s * n2.cx, -n2.s * n2.cy, 1};  // T2^T
    for (int k = 0; k < 9; ++k) L[k] = t[k];
  } else {
    const double is = 1.0 / n2.s;
    const double t[9] = {is, 0, n2.cx, 0, is, n2.cy, 0, 0, 1};  // T2^-1
    for (int k = 0; k < 9; ++k) L[k] = t[k];
  }
  double X[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) X[3 * r + c] = L[3 * r] * Mh[c] + L[3 * r + 1] * Mh[3 + c] + L[3 * r + 2] * Mh[6 + c];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) M[3 * r + c] = X[3 * r] * T1[c] + X[3 * r + 1] * T1[3 + c] + X[3 * r + 2] * T1[6 + c];
}

// The row(s) one correspondence contributes to the linear system, in normalised coordinates.
__device__ __forceinline__ void epipolar_row(double x1, double y1, double x2, double y2, double* r) {
  r[0] = x2 * x1; r[1] = x2 * y1; r[2] = x2; r[3] = y2 * x1; r[4] = y2 * y1; r[5] = y2; r[6] = x1; r[7] = y1; r[8] = 1.0;
}
__device__ __forceinline__ void homography_rows(double x1, double y1, double x2, double y2, double* ra, double* rb) {
  ra[0] = -x1; ra[1] = -y1; ra[2] = -1; ra[3] = 0; ra[4] = 0; ra[5] = 0; ra[6] = x1 * x2; ra[7] = y1 * x2; ra[8] = x2;
  rb[0] = 0; rb[1] = 0; rb[2] = 0; rb[3] = -x1; rb[4] = -y1; rb[5] = -1; rb[6] = x1 * y2; rb[7] = y1 * y2; rb[8] = y2;
}

struct Shared {
  double ata[45];           // upper triangle of the moment matrix (local optimisation)
  double sums[6];           // sum x1, y1, x2, y2, |x1 - c1|^2, |x2 - c2|^2 over the inliers
  double model[9];          // best model so far
  double cand[9];           // candidate (round winner / local model)
  double best_sum;
  double cand_sum;
  double a[9][9], v[9][9];  // Jacobi scratch
  int best_cnt, cand_cnt, cand_thread, count;
  int t_cnt[kThreads];
  double t_sum[kThreads];
  uint32_t warp_sums[kThreads / 32];
  uint32_t base;
};

// Score `M` over all matches with the whole CTA; result in sh.cand_cnt / sh.cand_sum (valid after the call).
template <bool kIsF>
__device__ void score_block(Shared& sh, const float4* __restrict__ pts, uint32_t m, const double* M, double thr) {
  if (threadIdx.x == 0) { sh.cand_cnt = 0; sh.cand_sum = 0.0; }
  __syncthreads();
  int cnt = 0;
  double sum = 0.0;
  for (uint32_t i = threadIdx.x; i < m; i += kThreads) {
    const double r = residual<kIsF>(M, pts[i]);
    if (r <= thr) { ++cnt; sum += r; }
  }
  atomicAdd(&sh.cand_cnt, cnt);
  atomicAdd(&sh.cand_sum, sum);
  __syncthreads();
}

// Local optimisation: least-squares model on the inliers of sh.model (normalised 8-point / DLT); leaves it in sh.cand.
template <bool kIsF>
__device__ void local_model(Shared& sh, const float4* __restrict__ pts, uint32_t m, double thr) {
  const int tid = threadIdx.x;
  if (tid < 45) sh.ata[tid] = 0.0;
  if (tid < 6) sh.sums[tid] = 0.0;
  if (tid == 0) sh.count = 0;
  __syncthreads();
  // pass 1: centroids of the inliers
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  int cnt = 0;
  for (uint32_t i = tid; i < m; i += kThreads) {
    const float4 p = pts[i];
    if (residual<kIsF>(sh.model, p) <= thr) { s0 += p.x; s1 += p.y; s2 += p.z; s3 += p.w; ++cnt; }
  }
  atomicAdd(&sh.sums[0], s0); atomicAdd(&sh.sums[1], s1); atomicAdd(&sh.sums[2], s2); atomicAdd(&sh.sums[3], s3);
  atomicAdd(&sh.count, cnt);
  __syncthreads();
  const double n = (double)sh.count;
  Norm n1{sh.sums[0] / n, sh.sums[1] / n, 1.0}, n2{sh.sums[2] / n, sh.sums[3] / n, 1.0};
  // pass 2: RMS distances
  double d1 = 0, d2 = 0;
  for (uint32_t i = tid; i < m; i += kThreads) {
    const float4 p = pts[i];
    if (residual<kIsF>(sh.model, p) <= thr) {
      d1 += (p.x - n1.cx) * (p.x - n1.cx) + (p.y - n1.cy) * (p.y - n1.cy);
      d2 += (p.z - n2.cx) * (p.z - n2.cx) + (p.w - n2.cy) * (p.w - n2.cy);
    }
  }
  atomicAdd(&sh.sums[4], d1); atomicAdd(&sh.sums[5], d2);
  __syncthreads();
  n1.s = sh.sums[4] > 0.0 ? sqrt(2.0) / sqrt(sh.sums[4] / n) : 1.0;
  n2.s = sh.sums[5] > 0.0 ? sqrt(2.0) / sqrt(sh.sums[5] / n) : 1.0;
  // pass 3: moment matrix A^T A of the normalised system
  double acc[45];
  for (int k = 0; k < 45; ++k) acc[k] = 0.0;
  for (uint32_t i = tid; i < m; i += kThreads) {
    const float4 p = pts[i];
    if (residual<kIsF>(sh.model, p) > thr) continue;
    const double x1 = (p.x - n1.cx) * n1.s, y1 = (p.y - n1.cy) * n1.s, x2 = (p.z - n2.cx) * n2.s, y2 = (p.w - n2.cy) * n2.s;
    double ra[9], rb[9];
    if (kIsF) epipolar_row(x1, y1, x2, y2, ra); else homography_rows(x1, y1, x2, y2, ra, rb);
    int k = 0;
    for (int a = 0; a < 9; ++a)
      for (int b = a; b < 9; ++b, ++k) acc[k] += ra[a] * ra[b] + (kIsF ? 0.0 : rb[a] * rb[b]);
  }
  for (int k = 0; k < 45; ++k)
    if (acc[k] != 0.0) atomicAdd(&sh.ata[k], acc[k]);
  __syncthreads();
  if (tid == 0) {
    int k = 0;
    for (int a = 0; a < 9; ++a)
      for (int b = a; b < 9; ++b, ++k) sh.a[a][b] = sh.a[b][a] = sh.ata[k];
    double Mh[9];
    smallest_eigenvector9(sh.a, sh.v, Mh);
    if (kIsF) enforce_rank2(Mh);
    denormalize(kIsF, Mh, n1, n2, sh.cand);
  }
  __syncthreads();
}

// Minimal solvers: models written to out[k][9], returns how many (F: 0, 1 or 3; H: 0 or 1).
__device__ int seven_point(const float4* __restrict__ pts, const uint32_t (&idx)[7], double (*out)[9]) {
  double A[7][9];
  for (int r = 0; r < 7; ++r) {
    const float4 p = pts[idx[r]];
    epipolar_row(p.x, p.y, p.z, p.w, A[r]);
  }
  double basis[2][9];
  if (!nullspace<7>(A, basis)) return 0;
  // det(l * F1 + (1 - l) * F2) is a cubic in l: recover it from four evaluations
  double G[9], dets[4];
  const double ls[4] = {0.0, 1.0, -1.0, 2.0};
  for (int e = 0; e < 4; ++e) {
    for (int k = 0; k < 9; ++k) G[k] = ls[e] * basis[0][k] + (1.0 - ls[e]) * basis[1][k];
    dets[e] = det3(G);
  }
  const double d = dets[0], b = 0.5 * (dets[1] + dets[2]) - d, s = 0.5 * (dets[1] - dets[2]);
  const double a = (dets[3] - 4.0 * b - d - 2.0 * s) / 6.0, c = s - a;
  double roots[3];
  const int nr = solve_cubic(a, b, c, d, roots);
  for (int k = 0; k < nr; ++k)
    for (int j = 0; j < 9; ++j) out[k][j] = roots[k] * basis[0][j] + (1.0 - roots[k]) * basis[1][j];
  return nr;
}

__device__ int four_point(const float4* __restrict__ pts, const uint32_t (&idx)[4], double (*out)[9]) {
  float4 p[4];
  Norm n1{0, 0, 1}, n2{0, 0, 1};
  for (int r = 0; r < 4; ++r) {
    p[r] = pts[idx[r]];
    n1.cx += 0.25 * p[r].x; n1.cy += 0.25 * p[r].y; n2.cx += 0.25 * p[r].z; n2.cy += 0.25 * p[r].w;
  }
  double d1 = 0, d2 = 0;
  for (int r = 0; r < 4; ++r) {
    d1 += (p[r].x - n1.cx) * (p[r].x - n1.cx) + (p[r].y - n1.cy) * (p[r].y - n1.cy);
    d2 += (p[r].z - n2.cx) * (p[r].z - n2.cx) + (p[r].w - n2.cy) * (p[r].w - n2.cy);
  }
  if (d1 <= 0.0 || d2 <= 0.0) return 0;
  n1.s = sqrt(2.0) / sqrt(0.25 * d1);
  n2.s = sqrt(2.0) / sqrt(0.25 * d2);
  double A[8][9];
  for (int r = 0; r < 4; ++r)
    homography_rows((p[r].x - n1.cx) * n1.s, (p[r].y - n1.cy) * n1.s, (p[r].z - n2.cx) * n2.s, (p[r].w - n2.cy) * n2.s, A[2 * r],
                    A[2 * r + 1]);
  double basis[1][9];
  if (!nullspace<8>(A, basis)) return 0;
  denormalize(false, basis[0], n1, n2, out[0]);
  return 1;
}

// LORANSAC for one model type over the matches of one pair.  Result: sh.model / sh.best_cnt / sh.best_sum; returns
// the number of samples drawn.
template <bool kIsF>
__device__ int loransac(Shared& sh, const float4* __restrict__ pts, uint32_t m, const Options& o, unsigned long long seed) {
  constexpr int kMin = kIsF ? 7 : 4, kLocal = kIsF ? 8 : 4;
  const int tid = threadIdx.x;
  const double thr = o.max_error * o.max_error;
  if (tid == 0) { sh.best_cnt = 0; sh.best_sum = 1e300; for (int k = 0; k < 9; ++k) sh.model[k] = 0.0; }
  __syncthreads();
  if (m < (uint32_t)kMin) return 0;
  const double cap = num_trials_for(o.min_inlier_ratio * 100000.0, 100000.0, o.confidence, kMin, 3.0);
  const int max_trials = (int)fmin((double)o.max_num_trials, cap);
  double dyn_max = (double)max_trials;
  int trials = 0;
  unsigned long long rs = seed ^ (0xD1B54A32D192ED03ull * (unsigned long long)(tid + 1));
  while (trials < max_trials) {
    // ---- 128 minimal samples: draw, solve, score every match
    uint32_t idx[kMin];
    sample_distinct<kMin>(rs, m, idx);
    double models[3][9];
    const int nm = kIsF ? seven_point(pts, reinterpret_cast<const uint32_t(&)[7]>(idx), models)
                        : four_point(pts, reinterpret_cast<const uint32_t(&)[4]>(idx), models);
    int my_cnt = -1, my_k = 0;
    double my_sum = 1e300;
    const float thr_f = (float)thr;
    for (int k = 0; k < nm; ++k) {
      double big = 0.0;
      for (int j = 0; j < 9; ++j) big = fmax(big, fabs(models[k][j]));
      if (!(big > 0.0)) continue;
      float fm[9];
      for (int j = 0; j < 9; ++j) fm[j] = (float)(models[k][j] / big);
      int cnt = 0;
      float sum = 0.f;
      for (uint32_t i = 0; i < m; ++i) {
        const float r = residual_f<kIsF>(fm, pts[i]);
        if (r <= thr_f) { ++cnt; sum += r; }
      }
      if (cnt > my_cnt || (cnt == my_cnt && (double)sum < my_sum)) { my_cnt = cnt; my_sum = (double)sum; my_k = k; }
    }
    sh.t_cnt[tid] = my_cnt;
    sh.t_sum[tid] = my_sum;
    __syncthreads();
    if (tid == 0) {
      int bt = 0;
      for (int t = 1; t < kThreads; ++t)
        if (sh.t_cnt[t] > sh.t_cnt[bt] || (sh.t_cnt[t] == sh.t_cnt[bt] && sh.t_sum[t] < sh.t_sum[bt])) bt = t;
      const bool better = sh.t_cnt[bt] > sh.best_cnt || (sh.t_cnt[bt] == sh.best_cnt && sh.t_sum[bt] < sh.best_sum);
      sh.cand_thread = better ? bt : -1;
    }
    __syncthreads();
    int winner = sh.cand_thread;
    if (winner >= 0) {  // uniform across the CTA: re-score the round's winner in double before adopting it
      if (tid == winner)
        for (int k = 0; k < 9; ++k) sh.cand[k] = models[my_k][k];
      __syncthreads();
      score_block<kIsF>(sh, pts, m, sh.cand, thr);
      const bool adopt = sh.cand_cnt > sh.best_cnt || (sh.cand_cnt == sh.best_cnt && sh.cand_sum < sh.best_sum);
      __syncthreads();
      if (adopt && tid == 0) {
        for (int k = 0; k < 9; ++k) sh.model[k] = sh.cand[k];
        sh.best_cnt = sh.cand_cnt;
        sh.best_sum = sh.cand_sum;
      }
      __syncthreads();
      if (!adopt) winner = -1;
    }
    if (winner >= 0) {
      if (sh.best_cnt > kMin && sh.best_cnt >= kLocal) {  // local optimisation on the new best model's inliers
        local_model<kIsF>(sh, pts, m, thr);
        score_block<kIsF>(sh, pts, m, sh.cand, thr);
        if (tid == 0 && (sh.cand_cnt > sh.best_cnt || (sh.cand_cnt == sh.best_cnt && sh.cand_sum < sh.best_sum))) {
          for (int k = 0; k < 9; ++k) sh.model[k] = sh.cand[k];
          sh.best_cnt = sh.cand_cnt;
          sh.best_sum = sh.cand_sum;
        }
        __syncthreads();
      }
      dyn_max = num_trials_for((double)sh.best_cnt, (double)m, o.confidence, kMin, 3.0);
    }
    trials += kThreads;
    if ((double)trials >= dyn_max && trials >= o.min_num_trials) break;
    __syncthreads();
  }
  __syncthreads();
  return trials;
}

// TwoViewGeometry::DetectWatermark on the n inlier correspondences q[0, n) of the accepted geometry, with the
// reference's default-constructed cameras (width = height = 0, sequential_matching.cc:89): the "inner" box is the point
// (0, 0), so every inlier not exactly there counts as lying in the border region, and what decides is whether at least
// watermark_min_inlier_ratio of the inliers fit ONE pure image translation within max_error.  COLMAP runs a LORANSAC
// over one-point samples for that (at most ComputeNumTrials(0.7, k = 1) = 18 trials); here every thread tries one
// random inlier's displacement (128 samples), the best goes through the local optimisation (mean displacement of its
// inliers), as in loransac() above.  Whole CTA; the verdict is uniform.
__device__ bool watermark_test(Shared& sh, const float4* __restrict__ q, uint32_t n, const Options& o, unsigned long long seed) {
  const int tid = threadIdx.x;
  const double thr = o.max_error * o.max_error;
  if (n == 0) return false;
  if (tid == 0) { sh.count = 0; sh.best_cnt = 0; sh.best_sum = 1e300; }
  __syncthreads();
  int nb = 0;
  for (uint32_t i = tid; i < n; i += kThreads) {
    const float4 p = q[i];
    if (!(p.x == 0.f && p.y == 0.f) && !(p.z == 0.f && p.w == 0.f)) ++nb;
  }
  atomicAdd(&sh.count, nb);
  __syncthreads();
  if ((double)sh.count / (double)n < o.watermark_min_inlier_ratio) return false;
  // ---- one-point samples
  unsigned long long rs = seed ^ (0xA0761D6478BD642Full * (unsigned long long)(tid + 1));
  const float4 s = q[(uint32_t)(rng_next(rs) % (unsigned long long)n)];
  const double tx = (double)s.z - (double)s.x, ty = (double)s.w - (double)s.y;
  int cnt = 0;
  double sum = 0.0;
  for (uint32_t i = 0; i < n; ++i) {
    const float4 p = q[i];
    const double dx = (double)p.z - (double)p.x - tx, dy = (double)p.w - (double)p.y - ty, r = dx * dx + dy * dy;
    if (r <= thr) { ++cnt; sum += r; }
  }
  sh.t_cnt[tid] = cnt;
  sh.t_sum[tid] = sum;
  __syncthreads();
  if (tid == 0) {
    int bt = 0;
    for (int t = 1; t < kThreads; ++t)
      if (sh.t_cnt[t] > sh.t_cnt[bt] || (sh.t_cnt[t] == sh.t_cnt[bt] && sh.t_sum[t] < sh.t_sum[bt])) bt = t;
    sh.cand_thread = bt;
    sh.best_cnt = sh.t_cnt[bt];
    sh.best_sum = sh.t_sum[bt];
  }
  __syncthreads();
  if (tid == sh.cand_thread) { sh.model[0] = tx; sh.model[1] = ty; }
  if (tid < 2) sh.sums[tid] = 0.0;
  __syncthreads();
  if (sh.best_cnt > 1) {  // local optimisation: the mean displacement of the best sample's inliers, adopted if it is better
    const double bx = sh.model[0], by = sh.model[1];
    double ax = 0.0, ay = 0.0;
    for (uint32_t i = tid; i < n; i += kThreads) {
      const float4 p = q[i];
      const double ux = (double)p.z - (double)p.x, uy = (double)p.w - (double)p.y;
      if ((ux - bx) * (ux - bx) + (uy - by) * (uy - by) <= thr) { ax += ux; ay += uy; }
    }
    atomicAdd(&sh.sums[0], ax);
    atomicAdd(&sh.sums[1], ay);
    if (tid == 0) { sh.cand_cnt = 0; sh.cand_sum = 0.0; }
    __syncthreads();
    const double mx = sh.sums[0] / (double)sh.best_cnt, my = sh.sums[1] / (double)sh.best_cnt;
    int c2 = 0;
    double s2 = 0.0;
    for (uint32_t i = tid; i < n; i += kThreads) {
      const float4 p = q[i];
      const double dx = (double)p.z - (double)p.x - mx, dy = (double)p.w - (double)p.y - my, r = dx * dx + dy * dy;
      if (r <= thr) { ++c2; s2 += r; }
    }
    atomicAdd(&sh.cand_cnt, c2);
    atomicAdd(&sh.cand_sum, s2);
    __syncthreads();
    if (tid == 0 && (sh.cand_cnt > sh.best_cnt || (sh.cand_cnt == sh.best_cnt && sh.cand_sum < sh.best_sum))) {
      sh.best_cnt = sh.cand_cnt;
      sh.best_sum = sh.cand_sum;
    }
    __syncthreads();
  }
  return (double)sh.best_cnt / (double)n >= o.watermark_min_inlier_ratio;
}

// grid = pairs; block = kThreads.  `matches` / `pair_out` are the matcher's result (pinned host memory, read through
// UVA), `xy` holds one float2 per descriptor-pool row, `pts` is device scratch of one float4 per match, `inliers` the
// pinned output buffer (same offsets as `matches`).
__global__ void __launch_bounds__(kThreads)
verify_kernel(const PairMeta* __restrict__ pairs, const uint2* __restrict__ matches, const PairOut* __restrict__ pair_out,
              const float2* __restrict__ xy, float4* __restrict__ pts_all, float4* __restrict__ ipts_all, Options o,
              Out* __restrict__ out, uint2* __restrict__ inliers) {
  // ipts_all: device scratch like pts_all; receives the inlier correspondences (for the watermark test)
  __shared__ Shared sh;
  const PairMeta pm = pairs[blockIdx.x];
  const PairOut po = pair_out[pm.out_slot];
  const uint32_t m = po.count;
  const uint2* mt = matches + po.start;
  float4* pts = pts_all + po.start;
  float4* ipts = ipts_all + po.start;
  const int tid = threadIdx.x;
  Out* res = out + pm.out_slot;
  for (uint32_t i = tid; i < m; i += kThreads) {
    const uint2 mm = mt[i];
    const float2 a = xy[pm.a_row0 + mm.x], b = xy[pm.b_row0 + mm.y];
    pts[i] = make_float4(a.x, a.y, b.x, b.y);
  }
  __syncthreads();
  if ((int)m < o.min_num_inliers) {
    if (tid == 0) {
      res->config = kDegenerate;
      res->num_inliers_F = res->num_inliers_H = res->trials_F = res->trials_H = 0;
      res->inlier_start = po.start;
      res->inlier_count = 0;
      for (int k = 0; k < 9; ++k) res->F[k] = res->H[k] = 0.0;
    }
    return;
  }
  const unsigned long long seed = o.seed * 0x9E3779B97F4A7C15ull + (unsigned long long)pm.out_slot * 0xC2B2AE3D27D4EB4Full;
  // homography first (its result is only counted), then the fundamental matrix, whose model stays in sh.model
  const int trials_h = loransac<false>(sh, pts, m, o, seed ^ 0x5851F42D4C957F2Dull);
  const int inl_h = sh.best_cnt;
  if (tid < 9) res->H[tid] = sh.model[tid];
  __syncthreads();
  const int trials_f = loransac<true>(sh, pts, m, o, seed);
  const int inl_f = sh.best_cnt;
  const bool ok_f = inl_f >= 7, ok_h = inl_h >= 4;
  int config;
  if ((!ok_f && !ok_h) || (inl_f < o.min_num_inliers && inl_h < o.min_num_inliers))
    config = kDegenerate;
  else
    config = (double)inl_h / (double)inl_f > o.max_H_inlier_ratio ? kPlanarOrPanoramic : kUncalibrated;
  // inlier matches of the F model, in match order
  const double thr = o.max_error * o.max_error;
  uint32_t running = 0;
  const uint32_t lane = tid & 31, wid = tid >> 5;
  for (uint32_t i0 = 0; i0 < m; i0 += kThreads) {
    const uint32_t i = i0 + tid;
    const bool in = config != kDegenerate && i < m && sampson_sq(sh.model, pts[i].x, pts[i].y, pts[i].z, pts[i].w) <= thr;
    const uint32_t ballot = __ballot_sync(0xffffffffu, in);
    __syncthreads();
    if (lane == 0) sh.warp_sums[wid] = __popc(ballot);
    __syncthreads();
    uint32_t before = 0, total = 0;
    for (int w = 0; w < kThreads / 32; ++w) {
      if (w < (int)wid) before += sh.warp_sums[w];
      total += sh.warp_sums[w];
    }
    if (in) {
      const uint32_t pos = running + before + __popc(ballot & ((1u << lane) - 1));
      inliers[po.start + pos] = mt[i];
      ipts[pos] = pts[i];
    }
    running += total;
  }
  if (tid < 9) res->F[tid] = sh.model[tid];  // before the watermark test reuses sh.model
  __syncthreads();
  if (config != kDegenerate && o.detect_watermark && watermark_test(sh, ipts, running, o, seed ^ 0x2545F4914F6CDD1Dull))
    config = kWatermark;
  if (tid == 0) {
    res->config = config;
    res->num_inliers_F = inl_f;
    res->num_inliers_H = inl_h;
    res->trials_F = trials_f;
    res->trials_H = trials_h;
    res->inlier_start = po.start;
    res->inlier_count = running;
  }
}

}  // namespace tvg
}  // namespace smb
