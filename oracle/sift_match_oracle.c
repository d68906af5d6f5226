/*
 * oracle/sift_match_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the arithmetic behind the reference's feature-matching
 * hot path.  The reference (garyjyzhang/scanner-colmap) calls
 *     colmap::MatchSiftFeaturesCPU(sift_options_, descriptors1, descriptors2, &featureMatches)
 * at integration/op_cpp/sequential_matching.cc:154.  COLMAP itself is an
 * un-vendored, un-pinned dependency (find_package(COLMAP REQUIRED),
 * integration/op_cpp/CMakeLists.txt:5; header layout = COLMAP 3.5), so this
 * file restates COLMAP 3.5's published src/feature/sift.cc algorithm:
 *
 *   ComputeSiftDistanceMatrix : dists(i,j) = sum_k int(d1[i,k]) * int(d2[j,k])
 *   FindBestMatchesOneWay     : per row, ascending scan with strict '>' for
 *                               best / second-best, acos distance mapping,
 *                               max_distance test, ratio test with '>='
 *   FindBestMatches           : optional cross-check, ascending idx1
 *   MatchSiftFeaturesCPU      : the two composed; max_num_matches is NOT
 *                               applied on the CPU path
 *
 * PARITY UNPINNED: the reference ships no test, fixture or golden vector for
 * this call and neither COLMAP nor Eigen can be
 * built here, so this oracle is pinned only by the hand-built known-answer
 * tests in tests/test_oracle_kat.py, which encode the semantics above.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path
 * (scanner_colmap_b200/csrc) never does.
 *
 * Float behaviour: the only float operations are (float)int * 2^-18 (exact),
 * fminf, acosf (host libm), one float multiply and two compares.  Build with
 * -ffp-contract=off and without -ffast-math.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_DIM 128

/* COLMAP 3.5 sift.cc FindBestMatchesOneWay: kDistNorm = 1.0f / (512.0f * 512.0f) */
static const float kDistNorm = 1.0f / (512.0f * 512.0f);

/* acosf(min(kDistNorm * d, 1.0f)) for an integer dot product d >= 0. */
float oracle_dist_normed(int d) {
  float x = kDistNorm * (float)d;
  if (x > 1.0f) x = 1.0f; /* std::min(x, 1.0f) */
  return acosf(x);
}

/* Table of oracle_dist_normed(d) for d = 0 .. 262144 (512*512); every larger
 * dot product saturates to entry 262144.  n must be 262145. */
void oracle_acos_lut(float *lut, int n) {
  for (int d = 0; d < n; ++d) lut[d] = oracle_dist_normed(d);
}

/* ComputeSiftDistanceMatrix: both operands cast to int, row-by-row dot
 * products into a dense row-major N1 x N2 int matrix. */
void oracle_distance_matrix(const uint8_t *d1, size_t n1, const uint8_t *d2,
                            size_t n2, int *dists) {
  int *a = (int *)malloc((n1 ? n1 : 1) * ORACLE_DIM * sizeof(int));
  int *b = (int *)malloc((n2 ? n2 : 1) * ORACLE_DIM * sizeof(int));
  for (size_t i = 0; i < n1 * ORACLE_DIM; ++i) a[i] = (int)d1[i];
  for (size_t i = 0; i < n2 * ORACLE_DIM; ++i) b[i] = (int)d2[i];
  for (size_t i1 = 0; i1 < n1; ++i1) {
    const int *ra = a + i1 * ORACLE_DIM;
    for (size_t i2 = 0; i2 < n2; ++i2) {
      const int *rb = b + i2 * ORACLE_DIM;
      int acc = 0;
      for (int k = 0; k < ORACLE_DIM; ++k) acc += ra[k] * rb[k];
      dists[i1 * n2 + i2] = acc;
    }
  }
  free(a);
  free(b);
}

/* FindBestMatchesOneWay over a (possibly transposed) view of dists:
 * element (r, c) lives at dists[r * rstride + c * cstride]. */
static size_t one_way(const int *dists, size_t rows, size_t cols,
                      size_t rstride, size_t cstride, float max_ratio,
                      float max_distance, int *matches) {
  size_t num = 0;
  for (size_t r = 0; r < rows; ++r) {
    matches[r] = -1;
    int best_i2 = -1;
    int best_dist = 0;
    int second_best_dist = 0;
    const int *row = dists + r * rstride;
    for (size_t c = 0; c < cols; ++c) {
      const int dist = row[c * cstride];
      if (dist > best_dist) {
        best_i2 = (int)c;
        second_best_dist = best_dist;
        best_dist = dist;
      } else if (dist > second_best_dist) {
        second_best_dist = dist;
      }
    }
    if (best_i2 == -1) continue; /* no positive dot product in the row */
    const float best_dist_normed = oracle_dist_normed(best_dist);
    if (best_dist_normed > max_distance) continue;
    const float second_best_dist_normed = oracle_dist_normed(second_best_dist);
    /* '>=' so that best == second-best is rejected */
    if (best_dist_normed >= max_ratio * second_best_dist_normed) continue;
    matches[r] = best_i2;
    ++num;
  }
  return num;
}

/* Exposed for tests: one-way matching on a dense matrix, transposed or not. */
size_t oracle_find_best_matches_one_way(const int *dists, size_t n1, size_t n2,
                                        int transposed, float max_ratio,
                                        float max_distance, int *matches) {
  if (!transposed)
    return one_way(dists, n1, n2, n2, 1, max_ratio, max_distance, matches);
  return one_way(dists, n2, n1, 1, n2, max_ratio, max_distance, matches);
}

/* FindBestMatches on a dense matrix. out = (idx1, idx2) uint32 pairs in
 * ascending idx1; returns the number of matches. */
size_t oracle_find_best_matches(const int *dists, size_t n1, size_t n2,
                                float max_ratio, float max_distance,
                                int cross_check, uint32_t *out) {
  int *m12 = (int *)malloc((n1 ? n1 : 1) * sizeof(int));
  size_t count = 0;
  one_way(dists, n1, n2, n2, 1, max_ratio, max_distance, m12);
  if (cross_check) {
    int *m21 = (int *)malloc((n2 ? n2 : 1) * sizeof(int));
    one_way(dists, n2, n1, 1, n2, max_ratio, max_distance, m21);
    for (size_t i1 = 0; i1 < n1; ++i1) {
      if (m12[i1] != -1 && m21[m12[i1]] != -1 && m21[m12[i1]] == (int)i1) {
        out[2 * count] = (uint32_t)i1;
        out[2 * count + 1] = (uint32_t)m12[i1];
        ++count;
      }
    }
    free(m21);
  } else {
    for (size_t i1 = 0; i1 < n1; ++i1) {
      if (m12[i1] != -1) {
        out[2 * count] = (uint32_t)i1;
        out[2 * count + 1] = (uint32_t)m12[i1];
        ++count;
      }
    }
  }
  free(m12);
  return count;
}

/* MatchSiftFeaturesCPU(options, d1, d2, &matches).  max_ratio / max_distance
 * are the double options narrowed to float at the FindBestMatches call, as in
 * COLMAP.  out must hold min(n1, n2) pairs when cross_check, else n1 pairs.
 * Returns the match count, or (size_t)-1 if the matrix cannot be allocated. */
size_t oracle_match_sift_features_cpu(const uint8_t *d1, size_t n1,
                                      const uint8_t *d2, size_t n2,
                                      double max_ratio, double max_distance,
                                      int cross_check, uint32_t *out) {
  if (n1 == 0 || n2 == 0) return 0;
  int *dists = (int *)malloc(n1 * n2 * sizeof(int));
  if (!dists) return (size_t)-1;
  oracle_distance_matrix(d1, n1, d2, n2, dists);
  size_t c = oracle_find_best_matches(dists, n1, n2, (float)max_ratio,
                                      (float)max_distance, cross_check, out);
  free(dists);
  return c;
}

/* Many independent pairs (the reference's per-row pair loop,
 * sequential_matching.cc:139-181, flattened), one OpenMP thread per pair.
 * desc[k] / n[k] describe image k; pairs = npairs x {k1, k2}; out_offsets
 * (npairs+1) are caller-provided capacities prefix (in matches); out_counts
 * receives per-pair counts.  Returns the thread count used. */
int oracle_match_many(const uint8_t *const *desc, const size_t *n,
                      const uint32_t *pairs, size_t npairs, double max_ratio,
                      double max_distance, int cross_check,
                      const uint64_t *out_offsets, uint32_t *out,
                      uint64_t *out_counts, int num_threads) {
  int used = 1;
#ifdef _OPENMP
  if (num_threads > 0) omp_set_num_threads(num_threads);
  used = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1)
#endif
  for (long p = 0; p < (long)npairs; ++p) {
    uint32_t a = pairs[2 * p], b = pairs[2 * p + 1];
    out_counts[p] = oracle_match_sift_features_cpu(
        desc[a], n[a], desc[b], n[b], max_ratio, max_distance, cross_check,
        out + 2 * out_offsets[p]);
  }
  return used;
}

int oracle_num_procs(void) {
#ifdef _OPENMP
  return omp_get_num_procs();
#else
  return 1;
#endif
}
