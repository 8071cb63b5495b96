// Host-side numerics shared by the LO-RANSAC drivers (H: ransac_host.cu, F: ransac_f_logic.hpp).  Plain C++ (no CUDA), so the same
// code is compiled into the library and into tests/native/ransac_f_cpu.cpp.
// Reference: degensac/rtools.c, utools.c, Htools.c, hash.c (line numbers at each function).
#pragma once
#include "parallel_host.hpp"
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

namespace mb2_ransac_common {
struct Score { unsigned I; double J; };

constexpr int ITER_SAM = 50, RAN_REP = 10, ILSQ_ITERS = 4, TC = 4, MWM = (9 / 4);  // rtools.h:7-35 (MWM is an int: 2)
constexpr int MAX_SAMPLES = 1000000;
constexpr double DEGENSAC_EPS = 2.2204e-16;
constexpr double CHECK_COEF = 9.0;
constexpr int MIN_GOOD_SYM_PTS = 5;

inline double truncQuad(double epsilon, double thr) {
  if (thr == 0) return 0;
  if (epsilon >= thr * 9 / 4) return 0;
  return 1 - (epsilon / (thr * 9 / 4));
}
inline int scoreLess(const Score a, const Score b) { return a.J < b.J; }  // __SCORE__ == SC_M
inline Score inlidxs(const double* err, int len, double th, int* inl) {
  Score s = {0, 0};
  for (int i = 0; i < len; ++i) {
    s.J += truncQuad(err[i], th);
    if (err[i] <= th) { inl[s.I] = i; ++(s.I); }
  }
  return s;
}
// same index list and count, without the MSAC sum (for the call sites that only read S.I and the list)
inline unsigned inlidxs_count(const double* err, int len, double th, int* inl) {
  unsigned I = 0;
  for (int i = 0; i < len; ++i) { inl[I] = i; I += err[i] <= th ? 1u : 0u; }
  return I;
}
inline int nsamples(int ninl, int ptNum, int samsiz, double conf) {
  double a = 1, b = 1;
  for (int i = 0; i < samsiz; i++) { a *= ninl - i; b *= ptNum - i; }
  a = a / b;
  if (a < DEGENSAC_EPS) return MAX_SAMPLES;
  a = 1 - a;
  if (a < DEGENSAC_EPS) return 1;
  b = std::log(1 - conf) / std::log(a);
  if (b > MAX_SAMPLES) return MAX_SAMPLES;
  return (int)std::ceil(b);
}
inline double det3(const double* A) {
  double r = (A[0] * A[4] * A[8] + A[2] * A[3] * A[7] + A[1] * A[5] * A[6]);
  r -= (A[2] * A[4] * A[6] + A[0] * A[5] * A[7] + A[1] * A[3] * A[8]);
  return r;
}

// glibc rand()/srand() stream with private state (TYPE_3, 128-byte state, same as the default generator)
struct LibcRand {
  struct random_data buf;
  char state[128];
  LibcRand() { std::memset(&buf, 0, sizeof buf); std::memset(state, 0, sizeof state); initstate_r(1, state, sizeof state, &buf); }
  LibcRand(const LibcRand& o) { copy_from(o); }
  LibcRand& operator=(const LibcRand& o) { if (this != &o) copy_from(o); return *this; }
  void seed(unsigned s) { srandom_r(s, &buf); }
  int next() { int32_t r; random_r(&buf, &r); return (int)r; }
 private:
  void copy_from(const LibcRand& o) {  // random_data holds pointers into `state`: rebase them
    std::memcpy(state, o.state, sizeof state);
    buf = o.buf;
    const ptrdiff_t shift = state - o.state;
    buf.fptr = (int32_t*)((char*)o.buf.fptr + shift); buf.rptr = (int32_t*)((char*)o.buf.rptr + shift);
    buf.state = (int32_t*)((char*)o.buf.state + shift); buf.end_ptr = (int32_t*)((char*)o.buf.end_ptr + shift);
  }
};

// rows 2q, 2q+1 of lin_hg's matrix (Htools.c:17-55) for correspondence q
inline void lin_rows(const double* u, int q, double* r0, double* r1) {
  const double* s = u + 6 * q;
  r0[0] = s[3]; r0[1] = 0; r0[2] = -s[0] * s[3]; r0[3] = s[4]; r0[4] = 0; r0[5] = -s[0] * s[4]; r0[6] = s[5]; r0[7] = 0; r0[8] = -s[0] * s[5];
  r1[0] = 0; r1[1] = s[3]; r1[2] = -s[1] * s[3]; r1[3] = 0; r1[4] = s[4]; r1[5] = -s[1] * s[4]; r1[6] = 0; r1[7] = s[5]; r1[8] = -s[1] * s[5];
}

// utools.c:97-167 (Gauss-Jordan with column pivoting bookkeeping; matrix row-wise)
inline int nullspace(double* matrix, double* nullsp, int n, int* buffer) {
  int* pnopivot = buffer; int nonpivot = 0;
  int* ppivot = buffer + n;
  int i = 0;
  const double tol = 1e-12;
  for (int j = 0; j < n; j++) {
    double pivot = std::fabs(matrix[n * i + j]); int max = i;
    for (int k = i + 1; k < n; k++) { double t = std::fabs(matrix[n * k + j]); if (pivot < t) { pivot = t; max = k; } }
    if (pivot < tol) {
      *(pnopivot++) = j; nonpivot++;
      for (int k = i; k < n; k++) matrix[n * k + j] = 0;
    } else {
      *(ppivot++) = j;
      for (int k = j; k < n; k++) { double t = matrix[i * n + k]; matrix[i * n + k] = matrix[max * n + k]; matrix[max * n + k] = t; }
      pivot = matrix[i * n + j];
      for (int k = j; k < n; k++) matrix[i * n + k] /= pivot;
      for (int k = 0; k < i; k++) { pivot = -matrix[k * n + j]; for (int l = j; l < n; l++) matrix[k * n + l] += pivot * matrix[i * n + l]; }
      for (int k = i + 1; k < n; k++) { pivot = matrix[k * n + j]; for (int l = j; l < n; l++) matrix[k * n + l] -= pivot * matrix[i * n + l]; }
      i++;
    }
  }
  for (int k = 0; k < nonpivot; k++) {
    int j = buffer[k];
    for (int l = 0; l < n - nonpivot; l++) nullsp[k * n + buffer[n + l]] = -matrix[l * n + j];
    for (int l = 0; l < nonpivot; l++) nullsp[k * n + buffer[l]] = (j == buffer[l]) ? 1 : 0;
  }
  return nonpivot;
}

// Htools.c:543-569
inline void cross3(double* out, const double* a, const double* b) {
  out[0] = a[1] * b[2] - a[2] * b[1]; out[1] = a[2] * b[0] - a[0] * b[2]; out[2] = a[0] * b[1] - a[1] * b[0];
}
inline int all_Hori_valid(const double* us, const int* idx) {
  double p[3], q[3];
  const double *a = us + 6 * idx[0], *b = us + 6 * idx[1], *c = us + 6 * idx[2], *d = us + 6 * idx[3];
  cross3(p, a, b); cross3(q, a + 3, b + 3);
  if ((p[0] * c[0] + p[1] * c[1] + p[2] * c[2]) * (q[0] * c[3] + q[1] * c[4] + q[2] * c[5]) < 0) return 0;
  if ((p[0] * d[0] + p[1] * d[1] + p[2] * d[2]) * (q[0] * d[3] + q[1] * d[4] + q[2] * d[5]) < 0) return 0;
  cross3(p, c, d); cross3(q, c + 3, d + 3);
  if ((p[0] * a[0] + p[1] * a[1] + p[2] * a[2]) * (q[0] * a[3] + q[1] * a[4] + q[2] * a[5]) < 0) return 0;
  if ((p[0] * b[0] + p[1] * b[1] + p[2] * b[2]) * (q[0] * b[3] + q[1] * b[4] + q[2] * b[5]) < 0) return 0;
  return 1;
}

// Smallest eigenvector of a symmetric 9x9 (stands in for LAPACK dsyev in lap_eig, lapwrap.c:67-97)
inline void smallest_eigvec9(const double* C, double* v) {
  const int n = 9;
  double A[81], V[81];
  std::memcpy(A, C, sizeof A);
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) V[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0, diag = 0;
    for (int i = 0; i < n; i++) { diag += A[i * n + i] * A[i * n + i]; for (int j = i + 1; j < n; j++) off += A[i * n + j] * A[i * n + j]; }
    if (off <= 1e-32 * diag || off == 0) break;
    for (int p = 0; p < n - 1; p++)
      for (int q = p + 1; q < n; q++) {
        const double apq = A[p * n + q];
        if (apq == 0) continue;
        const double theta = (A[q * n + q] - A[p * n + p]) / (2 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
        const double c = 1 / std::sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < n; k++) {
          const double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = c * akp - s * akq; A[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {
          const double apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = c * apk - s * aqk; A[q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {
          const double vkp = V[k * n + p], vkq = V[k * n + q];
          V[k * n + p] = c * vkp - s * vkq; V[k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  int best = 0;
  for (int i = 1; i < n; i++) if (A[i * n + i] < A[best * n + best]) best = i;
  for (int k = 0; k < n; k++) v[k] = V[k * n + best];
}

// utools.c:7-50
// Scratch of the large-set path of normu (per calling thread): the two distances of every listed correspondence, in list order.
inline std::vector<double>& normu_scratch() { static thread_local std::vector<double> g; return g; }
constexpr int NORMU_BIG = 4096;
inline void normu(const double* u, const int* inl, int len, double* A1, double* A2) {
  for (int j = 0; j < 3; j++) { A1[j] = 0; A2[j] = 0; }
  for (int j = 0; j < len; j++) { const double* p = u + 6 * inl[j]; A1[1] += p[0]; A1[2] += p[1]; A2[1] += p[3]; A2[2] += p[4]; }
  if (len > 0) for (int i = 1; i < 3; i++) { A1[i] /= len; A2[i] /= len; }
  if (len > NORMU_BIG) {
    // Large inlier sets (the LO steps of a view-sharded pair run on 200 k inliers): the square roots -- independent, correctly rounded -- are
    // taken in chunks on the host pool, the two sums stay single serial chains over the stored distances.  Same values, same additions.
    // (Measured and dropped: also gathering the coordinates into a dense array on the pool, so that the serial coordinate sums stream over
    // contiguous memory -- 3.4 ms per call at 200 k inliers against 2.2 ms: the serial thread then reads 10 MB written by other cores.)
    std::vector<double>& d = normu_scratch();
    if (d.size() < (size_t)2 * len) d.resize((size_t)2 * len);
    double* D = d.data();
    const int CHN = 8192, nch = (len + CHN - 1) / CHN;
    mb2par::parallel_chunks(nch, [&](int ck) {
      const int lo = ck * CHN, hi = std::min(len, lo + CHN);
      for (int j = lo; j < hi; j++) {
        const double* p = u + 6 * inl[j];
        const double a = p[0] - A1[1], b = p[1] - A1[2], c = p[3] - A2[1], e = p[4] - A2[2];
        D[2 * (size_t)j] = std::sqrt(a * a + b * b);
        D[2 * (size_t)j + 1] = std::sqrt(c * c + e * e);
      }
    });
    double s1 = A1[0], s2 = A2[0];
    for (int j = 0; j < len; j++) { s1 += D[2 * (size_t)j]; s2 += D[2 * (size_t)j + 1]; }
    A1[0] = s1; A2[0] = s2;
  } else {
    for (int j = 0; j < len; j++) {
      const double* p = u + 6 * inl[j];
      double a = p[0] - A1[1], b = p[1] - A1[2];
      A1[0] += std::sqrt(a * a + b * b);
      a = p[3] - A2[1]; b = p[4] - A2[2];
      A2[0] += std::sqrt(a * a + b * b);
    }
  }
  if (A1[0] != 0) A1[0] = len * std::sqrt(2) / A1[0];
  if (A2[0] != 0) A2[0] = len * std::sqrt(2) / A2[0];
  A1[1] *= -A1[0]; A1[2] *= -A1[0];
  A2[1] *= -A2[0]; A2[2] *= -A2[0];
}
// utools.c:70-89
inline void denormH(double* F, const double* A1, const double* A2) {
  double r = A2[0], x = A2[1], y = A2[2];
  F[6] += x * F[0] + y * F[3];
  F[7] += x * F[1] + y * F[4];
  F[8] += x * F[2] + y * F[5];
  F[0] *= r; F[1] *= r; F[2] *= r; F[3] *= r; F[4] *= r; F[5] *= r;
  r = 1 / A1[0]; x = -A1[1] * r; y = -A1[2] * r;
  for (int i = 0; i < 9; i += 3) { F[i] = r * F[i] + x * F[i + 2]; F[i + 1] = r * F[i + 1] + y * F[i + 2]; }
}
// Htools.c:98-130 (u2h): exact 4-point nullspace or normalised-DLT least squares
inline void u2h(const double* u, const int* inl, int len, double* H) {
  if (len < 4) return;
  if (len == 4) {
    // The reference fills a 9 x 8 column-major matrix (lin_hg, stride 2*len = 8) and then transposes it in place AS 9 x 9
    // (trnm(Z2, 9), Htools.c:108-109): the equations get mixed, and nine entries are read from uninitialised stack memory.  Reached
    // only when an inner sample has exactly 4 points (8 or 9 inliers).  The compiled reference behaves, in every run we compared
    // (516 seeded H and F runs over small and degenerate sets), as if those nine entries were 0: that arithmetic is reproduced here.
    double Z2[81], V[81];
    int nb[18];
    {
      double A[81], r0[9], r1[9];
      for (int i = 0; i < 81; i++) A[i] = 0.0;
      for (int i = 0; i < 4; i++) {   // lin_hg: element (row r, coefficient c) at c * 8 + r
        lin_rows(u, inl[i], r0, r1);
        for (int c = 0; c < 9; c++) { A[c * 8 + 2 * i] = r0[c]; A[c * 8 + 2 * i + 1] = r1[c]; }
      }
      for (int i = 0; i < 9; i++) for (int j = 0; j < 9; j++) Z2[i * 9 + j] = A[j * 9 + i];   // trnm(Z2, 9)
    }
    for (int i = 72; i < 81; ++i) Z2[i] = 0.0;
    std::memset(V, 0, sizeof V);
    nullspace(Z2, V, 9, nb);
    std::memcpy(H, V, 9 * sizeof(double));
    return;
  }
  double A1[3], A2[3], C[81];
  normu(u, inl, len, A1, A2);
  // lin_hgN (Htools.c:57-96) rows folded straight into the 9x9 covariance.  cov_mat (utools.c:170-185)
  // computes C[i][j] = sum_k Z[k][i] * Z[k][j] with k running over the rows in order; adding row 2i and
  // then row 2i+1 of every point to all 45 accumulators performs exactly those additions in exactly that
  // order, in one pass over the correspondences instead of 45 strided passes over a 2n x 9 matrix.
  // Large inlier sets (LO on tens of thousands of points) are summed in fixed chunks on the host pool (parallel_host.hpp);
  // the chunk boundaries do not depend on the thread count, so the result is machine-independent.  Short
  // lists (one chunk) keep the reference's single serial chain bit for bit.
  const int CH = 2048;
  const int nchunks = len <= 2 * CH ? 1 : (len + CH - 1) / CH;
  std::vector<double> part((size_t)nchunks * 45, 0.0);
  mb2par::parallel_chunks(nchunks, [&](int ck) {
    const int lo = nchunks == 1 ? 0 : ck * CH, hi = nchunks == 1 ? len : std::min(len, lo + CH);
    // Row 2i has entries only at columns S0 = {0, 2, 3, 5, 6, 8} (b0, -a0 b0, b1, -a0 b1, 1, -a0), row 2i + 1 only at S1 = {1, 2, 4, 5, 7, 8}
    // (b0, -a1 b0, b1, -a1 b1, 1, -a1); products with a structural zero are +-0 and leave a sum unchanged, so each of the 45 sums
    // receives, per point, its S0 x S0 product first and its S1 x S1 product second -- written out here as straight-line code on local
    // accumulators (15 sums fed by row 2i only, 15 by row 2i + 1 only, 6 -- columns {2, 5, 8} squared -- by both, in that order; the
    // other 9 stay 0).  Same additions in the same order as the generic double loop, about 3x fewer instructions.
    double e[15] = {0}, o[15] = {0}, m[6] = {0};
    for (int i = lo; i < hi; i++) {
      const double* s = u + 6 * inl[i];
      const double a0 = s[0] * A1[0] + A1[1], a1 = s[1] * A1[0] + A1[2];
      const double b0 = s[3] * A2[0] + A2[1], b1 = s[4] * A2[0] + A2[2];
      const double x0 = b0, x2 = -a0 * b0, x3 = b1, x5 = -a0 * b1, x8 = -a0 * 1.0;   // row 2i   (x6 = 1)
      const double y1 = b0, y2 = -a1 * b0, y4 = b1, y5 = -a1 * b1, y8 = -a1 * 1.0;   // row 2i+1 (y7 = 1)
      // row 2i only: (p, q) with p, q in S0, not both in {2, 5, 8}
      e[0] += x0 * x0;                                       // (0,0)
      e[1] += x2 * x0;                                       // (2,0)
      e[2] += x3 * x0; e[3] += x3 * x2; e[4] += x3 * x3;     // (3,0) (3,2) (3,3)
      e[5] += x5 * x0; e[6] += x5 * x3;                      // (5,0) (5,3)
      e[7] += 1.0 * x0; e[8] += 1.0 * x2; e[9] += 1.0 * x3; e[10] += 1.0 * x5; e[11] += 1.0 * 1.0;   // (6,0) (6,2) (6,3) (6,5) (6,6)
      e[12] += x8 * x0; e[13] += x8 * x3; e[14] += x8 * 1.0; // (8,0) (8,3) (8,6)
      // row 2i + 1 only
      o[0] += y1 * y1;                                       // (1,1)
      o[1] += y2 * y1;                                       // (2,1)
      o[2] += y4 * y1; o[3] += y4 * y2; o[4] += y4 * y4;     // (4,1) (4,2) (4,4)
      o[5] += y5 * y1; o[6] += y5 * y4;                      // (5,1) (5,4)
      o[7] += 1.0 * y1; o[8] += 1.0 * y2; o[9] += 1.0 * y4; o[10] += 1.0 * y5; o[11] += 1.0 * 1.0;   // (7,1) (7,2) (7,4) (7,5) (7,7)
      o[12] += y8 * y1; o[13] += y8 * y4; o[14] += y8 * 1.0; // (8,1) (8,4) (8,7)
      // both rows, row 2i first
      m[0] += x2 * x2; m[0] += y2 * y2;                      // (2,2)
      m[1] += x5 * x2; m[1] += y5 * y2;                      // (5,2)
      m[2] += x5 * x5; m[2] += y5 * y5;                      // (5,5)
      m[3] += x8 * x2; m[3] += y8 * y2;                      // (8,2)
      m[4] += x8 * x5; m[4] += y8 * y5;                      // (8,5)
      m[5] += x8 * x8; m[5] += y8 * y8;                      // (8,8)
    }
    double* acc = part.data() + (size_t)ck * 45;   // index t of (p, q), q <= p: p (p + 1) / 2 + q
    auto T = [](int p, int q) { return p * (p + 1) / 2 + q; };
    acc[T(0, 0)] = e[0]; acc[T(2, 0)] = e[1]; acc[T(3, 0)] = e[2]; acc[T(3, 2)] = e[3]; acc[T(3, 3)] = e[4]; acc[T(5, 0)] = e[5]; acc[T(5, 3)] = e[6];
    acc[T(6, 0)] = e[7]; acc[T(6, 2)] = e[8]; acc[T(6, 3)] = e[9]; acc[T(6, 5)] = e[10]; acc[T(6, 6)] = e[11]; acc[T(8, 0)] = e[12]; acc[T(8, 3)] = e[13]; acc[T(8, 6)] = e[14];
    acc[T(1, 1)] = o[0]; acc[T(2, 1)] = o[1]; acc[T(4, 1)] = o[2]; acc[T(4, 2)] = o[3]; acc[T(4, 4)] = o[4]; acc[T(5, 1)] = o[5]; acc[T(5, 4)] = o[6];
    acc[T(7, 1)] = o[7]; acc[T(7, 2)] = o[8]; acc[T(7, 4)] = o[9]; acc[T(7, 5)] = o[10]; acc[T(7, 7)] = o[11]; acc[T(8, 1)] = o[12]; acc[T(8, 4)] = o[13]; acc[T(8, 7)] = o[14];
    acc[T(2, 2)] = m[0]; acc[T(5, 2)] = m[1]; acc[T(5, 5)] = m[2]; acc[T(8, 2)] = m[3]; acc[T(8, 5)] = m[4]; acc[T(8, 8)] = m[5];
  });
  double acc[45];
  for (int t = 0; t < 45; t++) { acc[t] = part[t]; for (int ck = 1; ck < nchunks; ck++) acc[t] += part[(size_t)ck * 45 + t]; }
  {
    int t = 0;
    for (int p = 0; p < 9; p++)
      for (int q = 0; q <= p; q++, t++) { C[9 * p + q] = acc[t]; C[p + 9 * q] = acc[t]; }
  }
  smallest_eigvec9(C, H);
  denormH(H, A1, A2);
}

// hash.c
inline uint32_t SuperFastHash(const char* data, int len) {
  uint32_t hash = len, tmp;
  if (len <= 0 || data == 0) return 0;
  auto get16 = [](const char* d) { return (uint32_t)(((uint32_t)((const uint8_t*)d)[1]) << 8) + (uint32_t)((const uint8_t*)d)[0]; };
  int rem = len & 3;
  len >>= 2;
  for (; len > 0; len--) {
    hash += get16(data);
    tmp = (get16(data + 2) << 11) ^ hash;
    hash = (hash << 16) ^ tmp;
    data += 4;
    hash += hash >> 11;
  }
  switch (rem) {
    case 3: hash += get16(data); hash ^= hash << 16; hash ^= ((signed char)data[2]) << 18; hash += hash >> 11; break;
    case 2: hash += get16(data); hash ^= hash << 11; hash += hash >> 17; break;
    case 1: hash += (signed char)*data; hash ^= hash << 10; hash += hash >> 1;
  }
  hash ^= hash << 3; hash += hash >> 5; hash ^= hash << 4; hash += hash >> 17; hash ^= hash << 25; hash += hash >> 6;
  return hash;
}
struct HashTable {
  struct Field { uint32_t hash; int length, iterID; };
  std::vector<Field> f[64];  // newest last; the reference prepends, so scan backwards
  void insert(uint32_t h, int len, int id) { f[h % 64].push_back({h, len, id}); }
  int contains(uint32_t h, int len, int id) const {
    const auto& b = f[h % 64];
    for (size_t i = b.size(); i-- > 0;) if (b[i].hash == h && b[i].length == len && b[i].iterID == id) return id;
    for (size_t i = b.size(); i-- > 0;) if (b[i].hash == h && b[i].length == len) return b[i].iterID;
    return -1;
  }
};

}  // namespace mb2_ransac_common
