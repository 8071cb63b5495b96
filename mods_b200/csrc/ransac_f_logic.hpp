// LO-RANSAC for the fundamental matrix with the DEGENSAC plane-and-parallax branch: the sequential logic of
// exp_ransacFcustom, written against a `Scorer` that evaluates residuals of ALL correspondences for a model (the GPU in the
// library, `mb2_score_models`; the oracle's CPU scorers in tests/native/ransac_f_cpu.cpp -- test infrastructure only).
//
// Reference: degensac/exp_ranF.c:795-1192 (main loop), :610-736 (exp_iterFcustom), :739-790 (exp_inFranicustom);
// Ftools.c (lin_fm :14-38, slcm :40-93, rroots3 :211-255, lin_fmN :257-287, singulF :289-308, u2f :311-361, u2fw :363-422,
// epipole/getorisig/all_ori_valid :427-459, exFDs :148-171, exFDsSym :172-196); DegUtils.c (checksample :42-83, Hdetect :94-170,
// sortDs :173-193, dHDs :196-221, rFtH :263-445, innerFH :487-612, dual_sample :615-653, u2Fit :656-717, innerH :720-760);
// ranH.c (iterH :18-85, inHrani :88-137); utools.c (denormF :52-68, cov_mat :170-185); ccmath svduv / qrbdv / ldvmat
// (matutls/*.c) for the 3x3 SVD inside Hdetect, whose third right-singular column is taken UNSORTED.
//
// What is restructured, and why it is the same computation:
//  * Sample k is `srand(seed_k); 7 x random(); seed_{k+1} = rand()` with a pool permutation that never depends on scores, so the
//    7-point models (up to three real roots each) are generated in batches and FDS1 of every model over all correspondences is ONE
//    launch per batch (I = #inliers, J = MSAC sum); the reference's best-so-far / symmetric check / degeneracy / LO / adaptive
//    stop logic is replayed over those scores.  Anything that consumes rand() after the sample (LO, innerH, rFtH) starts from the
//    generator state the reference has at that point (re-seeded from seed_k lazily).
//  * Residual buffers errs[0..3] (+ errs[4] alias, + errorsBest) are modelled as "residuals of model f", materialised only when
//    the reference reads them; the rotations and the aliasing are kept literally.
//  * rFtH's 2-point epipole samples are scored in speculative batches (one launch per batch); at the first sample that improves
//    the consensus the generator / permutation state is rewound to that sample and the reference's inner steps run.
//  * LAPACK (`dsyev` in u2f/u2fw/u2h, `dgesvd` in singulF) is replaced by cyclic Jacobi / one-sided Jacobi, and the 9x8 `svduv`
//    of the 8-point case by a Householder null vector: same subspaces to ~1e-15, decisions agree unless a residual sits that
//    close to a threshold.  ccmath's 3x3 svduv and minv in Hdetect are restated step by step (their column order matters).
//  * Undefined corners of the reference that are refused or pinned here: u2f with fewer than 8 points reads an uninitialised
//    buffer in the reference -> the model is left unchanged; a run in which no model is ever accepted returns an all-zero mask.
#pragma once
#include "parallel_host.hpp"
#include "ransac_common.hpp"
#include "minv3.hpp"

namespace mb2_rf {
using namespace mb2_ransac_common;

constexpr double F_CHECK_COEF = 4.0, F_SYMM_COEF = 0.6;  // exp_ranF.c:19-20
constexpr int W_HDS = 0, W_FDS = 3, W_FDSSYM = 4, W_EXFDSSYM = 5;  // `which` of mb2_score_models

// ------------------------------------------------------------------------------------------------------------------
// Closed-form scorers for tiny point sets (the 7 points of checksample) and the LSQ weights; same expressions as ransac.cu.
inline double hds_point(const double* u, const double* H) {  // Htools.c:158-196 with lin_hg :17-55, pinvJ :132-156
  const double x1 = u[0], y1 = u[1];
  const double z1[9] = {u[3], 0, -x1 * u[3], u[4], 0, -x1 * u[4], u[5], 0, -x1 * u[5]};
  const double z2[9] = {0, u[3], -y1 * u[3], 0, u[4], -y1 * u[4], 0, u[5], -y1 * u[5]};
  double r1 = 0, r2 = 0;
  for (int j = 0; j < 9; j++) { r1 += H[j] * z1[j]; r2 += H[j] * z2[j]; }
  const double a = H[0] - H[2] * u[0], b = H[3] - H[5] * u[0], c = -H[8] - H[2] * u[3] - H[5] * u[4];
  const double d = H[1] - H[2] * u[1], e = H[4] - H[5] * u[1];
  double pJ[8];
  const double a2 = a * a, b2 = b * b, c2 = c * c, d2 = d * d, e2 = e * e;
  const double c2pd2 = c2 + d2, ab = a * b, de = d * e;
  const double Q = c * (c2pd2 + e2);
  pJ[0] = -b * de + a * (c2 + e2);
  pJ[1] = b * c2pd2 - a * de;
  pJ[2] = Q;
  pJ[3] = -c * (a * d + b * e);
  pJ[4] = d * (b2 + c2) - ab * e;
  pJ[5] = -ab * d + e * (a2 + c2);
  pJ[6] = pJ[3];
  pJ[7] = c * (a2 + b2 + c2);
  const double N = a * pJ[0] + b * pJ[1] + c * pJ[2];
  for (int q = 0; q < 8; q++) pJ[q] /= N;
  double s = 0;
  for (int j = 0; j < 4; j++) { const double v = pJ[j] * r1 + pJ[j + 4] * r2; s += v * v; }
  return s;
}
// LSQ weight of one correspondence (exFDs Ftools.c:148-171: 1/sqrt(denominator); exFDsSym :172-196: a b / (a + b))
inline double exfds_weight(const double* u, const double* F, bool sym) {
  const double u1 = u[0], u2 = u[1], u4 = u[3], u5 = u[4];
  const double rxc = F[0] * u4 + F[3] * u5 + F[6];
  const double ryc = F[1] * u4 + F[4] * u5 + F[7];
  const double rx = F[0] * u1 + F[1] * u2 + F[2];
  const double ry = F[3] * u1 + F[4] * u2 + F[5];
  if (!sym) return 1 / std::sqrt(rxc * rxc + ryc * ryc + rx * rx + ry * ry);
  const double a = rxc * rxc + ryc * ryc, b = rx * rx + ry * ry;
  return (a * b) / (a + b);
}

// ------------------------------------------------------------------------------------------------------------------
// 7-point solver pieces
inline void slcm(const double* A, double* B, double* p) {  // Ftools.c:40-93; B becomes A - B
  const double a11 = A[0], a12 = A[1], a13 = A[2], a21 = A[3], a22 = A[4], a23 = A[5], a31 = A[6], a32 = A[7], a33 = A[8];
  double b11 = B[0], b12 = B[1], b13 = B[2], b21 = B[3], b22 = B[4], b23 = B[5], b31 = B[6], b32 = B[7], b33 = B[8];
  p[0] = -(b13 * b22 * b31) + b12 * b23 * b31 + b13 * b21 * b32 - b11 * b23 * b32 - b12 * b21 * b33 + b11 * b22 * b33;
  p[1] = -(a33 * b12 * b21) + a32 * b13 * b21 + a33 * b11 * b22 - a31 * b13 * b22 - a32 * b11 * b23 + a31 * b12 * b23 +
         a23 * b12 * b31 - a22 * b13 * b31 - a13 * b22 * b31 + 3 * b13 * b22 * b31 + a12 * b23 * b31 - 3 * b12 * b23 * b31 -
         a23 * b11 * b32 + a21 * b13 * b32 + a13 * b21 * b32 - 3 * b13 * b21 * b32 - a11 * b23 * b32 + 3 * b11 * b23 * b32 +
         (a22 * b11 - a21 * b12 - a12 * b21 + 3 * b12 * b21 + a11 * b22 - 3 * b11 * b22) * b33;
  p[2] = -(a21 * a33 * b12) + a21 * a32 * b13 + a13 * a32 * b21 - a12 * a33 * b21 + 2 * a33 * b12 * b21 - 2 * a32 * b13 * b21 -
         a13 * a31 * b22 + a11 * a33 * b22 - 2 * a33 * b11 * b22 + 2 * a31 * b13 * b22 + a12 * a31 * b23 - a11 * a32 * b23 +
         2 * a32 * b11 * b23 - 2 * a31 * b12 * b23 + 2 * a13 * b22 * b31 - 3 * b13 * b22 * b31 - 2 * a12 * b23 * b31 +
         3 * b12 * b23 * b31 + a13 * a21 * b32 - 2 * a21 * b13 * b32 - 2 * a13 * b21 * b32 + 3 * b13 * b21 * b32 +
         2 * a11 * b23 * b32 - 3 * b11 * b23 * b32 +
         a23 * (-(a32 * b11) + a31 * b12 + a12 * b31 - 2 * b12 * b31 - a11 * b32 + 2 * b11 * b32) +
         (-(a12 * a21) + 2 * a21 * b12 + 2 * a12 * b21 - 3 * b12 * b21 - 2 * a11 * b22 + 3 * b11 * b22) * b33 +
         a22 * (a33 * b11 - a31 * b13 - a13 * b31 + 2 * b13 * b31 + a11 * b33 - 2 * b11 * b33);
  for (int i = 0; i < 9; i++) B[i] = A[i] - B[i];
  b11 = B[0]; b12 = B[1]; b13 = B[2]; b21 = B[3]; b22 = B[4]; b23 = B[5]; b31 = B[6]; b32 = B[7]; b33 = B[8];
  p[3] = -(b13 * b22 * b31) + b12 * b23 * b31 + b13 * b21 * b32 - b11 * b23 * b32 - b12 * b21 * b33 + b11 * b22 * b33;
}
inline int rroots3(const double* po, double* r) {  // Ftools.c:211-255
  const double lead = po[0], tail = po[3];
  const double b = po[1] / lead, c = po[2] / lead;
  const double b2 = b * b, bt = b / 3;
  const double p = (3 * c - b2) / 9;
  const double q = ((2 * b2 * b) / 27 - b * c / 3 + tail / lead) / 2;
  const double D = q * q + p * p * p;
  if (D > 0) {
    const double A = std::sqrt(D) - q;
    if (A > 0) { const double v = std::pow(A, 1.0 / 3); r[0] = v - p / v - bt; }
    else { const double v = std::pow(-A, 1.0 / 3); r[0] = p / v - v - bt; }
    return 1;
  }
  const double e = q > 0 ? 1 : -1;
  const double R = e * std::sqrt(-p), twoR = R * 2;
  double cosphi = q / (R * R * R);
  if (cosphi > 1) cosphi = 1; else if (cosphi < -1) cosphi = -1;
  const double phit = std::acos(cosphi) / 3, pit = 3.14159265358979 / 3;
  r[0] = -twoR * std::cos(phit) - bt;
  r[1] = twoR * std::cos(pit - phit) - bt;
  r[2] = twoR * std::cos(pit + phit) - bt;
  return 3;
}
inline void epipole(double* ec, const double* F) {  // Ftools.c:427-434
  const double xeps = 1.9984e-15;
  cross3(ec, F, F + 6);
  for (int i = 0; i < 3; i++) if ((ec[i] > xeps) || (ec[i] < -xeps)) return;
  cross3(ec, F + 3, F + 6);
}
inline int all_ori_valid(const double* F, const double* us, const int* idx, int N) {  // Ftools.c:436-459
  double ec[3];
  epipole(ec, F);
  auto sig_of = [&](const double* u) { const double s1 = F[0] * u[3] + F[3] * u[4] + F[6] * u[5], s2 = ec[1] * u[2] - ec[2] * u[1]; return s1 * s2; };
  const double sig1 = sig_of(us + 6 * idx[0]);
  for (int i = 1; i < N; i++) if (sig1 * sig_of(us + 6 * idx[i]) < 0) return 0;
  return 1;
}

// ------------------------------------------------------------------------------------------------------------------
// Linear algebra standing in for LAPACK / restating ccmath
// A = U diag(s) V^T for a 3x3 (row-major) by one-sided Jacobi; singular values sorted descending.
inline void svd3(const double* A, double* U, double* s, double* V) {
  double W[9];
  std::memcpy(W, A, sizeof W);
  for (int i = 0; i < 9; i++) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        double al = 0, be = 0, ga = 0;
        for (int k = 0; k < 3; k++) { al += W[k * 3 + p] * W[k * 3 + p]; be += W[k * 3 + q] * W[k * 3 + q]; ga += W[k * 3 + p] * W[k * 3 + q]; }
        if (ga == 0 || std::fabs(ga) <= 1e-17 * std::sqrt(al * be)) continue;
        off = std::max(off, std::fabs(ga) / std::sqrt(al * be));
        const double zeta = (be - al) / (2 * ga);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1 + zeta * zeta));
        const double c = 1 / std::sqrt(1 + t * t), sn = c * t;
        for (int k = 0; k < 3; k++) {
          const double wp = W[k * 3 + p], wq = W[k * 3 + q];
          W[k * 3 + p] = c * wp - sn * wq; W[k * 3 + q] = sn * wp + c * wq;
          const double vp = V[k * 3 + p], vq = V[k * 3 + q];
          V[k * 3 + p] = c * vp - sn * vq; V[k * 3 + q] = sn * vp + c * vq;
        }
      }
    if (off < 1e-16) break;
  }
  double nrm[3]; int ord[3] = {0, 1, 2};
  for (int j = 0; j < 3; j++) nrm[j] = std::sqrt(W[j] * W[j] + W[3 + j] * W[3 + j] + W[6 + j] * W[6 + j]);
  std::sort(ord, ord + 3, [&](int a, int b) { return nrm[a] > nrm[b]; });
  double Vs[9];
  for (int j = 0; j < 3; j++) {
    const int o = ord[j];
    s[j] = nrm[o];
    for (int k = 0; k < 3; k++) { U[k * 3 + j] = nrm[o] > 0 ? W[k * 3 + o] / nrm[o] : 0.0; Vs[k * 3 + j] = V[k * 3 + o]; }
  }
  std::memcpy(V, Vs, sizeof Vs);
}
// Ftools.c:289-308: closest rank-2 matrix (the transposes around lap_SVD cancel)
inline void singulF(double* F) {
  double U[9], s[3], V[9], out[9];
  svd3(F, U, s, V);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) out[i * 3 + j] = U[i * 3 + 0] * s[0] * V[j * 3 + 0] + U[i * 3 + 1] * s[1] * V[j * 3 + 1];
  std::memcpy(F, out, sizeof out);
}
// utools.c:52-68
inline void denormF(double* F, const double* A1, const double* A2) {
  double r = A2[0], x = A2[1], y = A2[2];
  F[6] += x * F[0] + y * F[3];
  F[7] += x * F[1] + y * F[4];
  F[8] += x * F[2] + y * F[5];
  F[0] *= r; F[1] *= r; F[2] *= r; F[3] *= r; F[4] *= r; F[5] *= r;
  r = A1[0]; x = A1[1]; y = A1[2];
  F[2] += x * F[0] + y * F[1];
  F[5] += x * F[3] + y * F[4];
  F[8] += x * F[6] + y * F[7];
  F[0] *= r; F[3] *= r; F[6] *= r; F[1] *= r; F[4] *= r; F[7] *= r;
}
// Left null vector of the 9 x n (n <= 8) matrix M (row-major, row stride n): the last column of the full U that
// svduv(D, Z, V, 9, U, 8) returns in u2f / u2fw (Ftools.c:333, :391), here by Householder QR.
inline void left_null9(double* M, int n, double* out) {
  double vs[8][9];
  for (int c = 0; c < n; c++) {
    double nrm = 0;
    for (int r = c; r < 9; r++) nrm += M[r * n + c] * M[r * n + c];
    nrm = std::sqrt(nrm);
    double* v = vs[c];
    for (int r = 0; r < 9; r++) v[r] = r < c ? 0.0 : M[r * n + c];
    if (nrm == 0) { for (int r = 0; r < 9; r++) v[r] = 0; continue; }
    v[c] += (M[c * n + c] >= 0 ? nrm : -nrm);
    double vn = 0;
    for (int r = c; r < 9; r++) vn += v[r] * v[r];
    vn = std::sqrt(vn);
    for (int r = c; r < 9; r++) v[r] /= vn;
    for (int k = c; k < n; k++) {
      double dot = 0;
      for (int r = c; r < 9; r++) dot += v[r] * M[r * n + k];
      for (int r = c; r < 9; r++) M[r * n + k] -= 2 * dot * v[r];
    }
  }
  double q[9] = {0, 0, 0, 0, 0, 0, 0, 0, 1};
  for (int c = n - 1; c >= 0; c--) {
    double dot = 0;
    for (int r = c; r < 9; r++) dot += vs[c][r] * q[r];
    for (int r = c; r < 9; r++) q[r] -= 2 * dot * vs[c][r];
  }
  std::memcpy(out, q, sizeof q);
}
// u2f (Ftools.c:311-361) and u2fw (:363-422, w != nullptr: weights indexed by correspondence).
inline void u2f_w(const double* u, const int* inl, const double* w, int len, double* F) {
  if (len < 8) return;  // the reference reads an uninitialised buffer here; pinned: the model stays as it is
  if (len == 8) {
    double Z[72];
    for (int i = 0; i < 8; i++) {
      const double* s = u + 6 * inl[i];
      for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) Z[(k * 3 + l) * 8 + i] = s[k + 3] * s[l];
    }
    // u2fw scales with `scalmul(Z + i, w[j], 9, 9)` (Ftools.c:388): stride 9 on a matrix whose row stride is len = 8, so the
    // weight of correspondence i lands on the entries i, i + 9, i + 18, ... (kept literally; entries past the matrix are dropped)
    if (w) for (int i = 0; i < 8; i++) for (int t = 0; t < 9; t++) if (i + 9 * t < 72) Z[i + 9 * t] *= w[inl[i]];
    left_null9(Z, 8, F);
    singulF(F);
    return;
  }
  double A1[3], A2[3];
  normu(u, inl, len, A1, A2);
  // lin_fmN rows (x their weight) folded into the 9x9 covariance in the order cov_mat adds them; long lists in fixed chunks
  const int CH = 2048;
  const int nchunks = len <= 2 * CH ? 1 : (len + CH - 1) / CH;
  std::vector<double> part((size_t)nchunks * 45, 0.0);
  mb2par::parallel_chunks(nchunks, [&](int ck) {
    double* acc = part.data() + (size_t)ck * 45;
    const int lo = nchunks == 1 ? 0 : ck * CH, hi = nchunks == 1 ? len : std::min(len, lo + CH);
    for (int i = lo; i < hi; i++) {
      const double* s = u + 6 * inl[i];
      double a[3], b[3], z[9];
      a[2] = 1; b[2] = 1;
      a[0] = s[0] * A1[0] + A1[1]; a[1] = s[1] * A1[0] + A1[2];
      b[0] = s[3] * A2[0] + A2[1]; b[1] = s[4] * A2[0] + A2[2];
      for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) z[k * 3 + l] = a[l] * b[k];
      if (w) { const double wi = w[inl[i]]; for (int r = 0; r < 9; r++) z[r] *= wi; }
      int t = 0;
      for (int p = 0; p < 9; p++) for (int q = 0; q <= p; q++, t++) acc[t] += z[p] * z[q];
    }
  });
  double C[81];
  {
    int t = 0;
    for (int p = 0; p < 9; p++)
      for (int q = 0; q <= p; q++, t++) {
        double v = part[t];
        for (int ck = 1; ck < nchunks; ck++) v += part[(size_t)ck * 45 + t];
        C[9 * p + q] = v; C[p + 9 * q] = v;
      }
  }
  smallest_eigvec9(C, F);
  singulF(F);
  denormF(F, A1, A2);
}

// ccmath svduv(d, a, u, 3, v, 3) restated for the right-singular matrix only (matutls/svduv.c, ldvmat.c, qrbdv.c): Householder
// bidiagonalisation, V from the single row reflector, implicit-shift QR sweeps rotating the columns of V.  Columns are NOT sorted.
inline void svd3_ccmath_V(const double* Ain, double* V) {
  const int n = 3;
  double a[9], d[3] = {0, 0, 0}, e[3] = {0, 0, 0}, w[3];
  std::memcpy(a, Ain, sizeof a);
  for (int i = 0; i < n; i++) {
    const int mm = n - i, nm = n - 1 - i;
    if (mm > 1) {
      double sv = 0, h = 0, s = 0;
      for (int j = 0; j < mm; j++) { w[j] = a[(i + j) * n + i]; s += w[j] * w[j]; }
      if (s > 0.) {
        h = std::sqrt(s); if (a[i * n + i] < 0.) h = -h;
        s += a[i * n + i] * h; s = 1. / s;
        w[0] += h;
        const double t = 1. / w[0];
        sv = 1. + std::fabs(a[i * n + i] / h);
        for (int k = 1; k < n - i; k++) {
          double r = 0.;
          for (int j = 0; j < mm; j++) r += w[j] * a[(i + j) * n + i + k];
          r *= s;
          for (int j = 0; j < mm; j++) a[(i + j) * n + i + k] -= r * w[j];
        }
        for (int j = 1; j < mm; j++) a[(i + j) * n + i] = t * w[j];
      }
      a[i * n + i] = sv; d[i] = -h;
    }
    if (mm == 1) d[i] = a[i * n + i];
    if (nm > 1) {
      double sv = 0, h = 0, s = 0;
      double* p1 = a + i * n + i + 1;
      for (int j = 0; j < nm; j++) s += p1[j] * p1[j];
      if (s > 0.) {
        h = std::sqrt(s); if (p1[0] < 0.) h = -h;
        sv = 1. + std::fabs(p1[0] / h);
        s += p1[0] * h; s = 1. / s;
        p1[0] += h;
        const double t = 1. / p1[0];
        for (int k = 1; k < n - i; k++) {
          double* pp = p1 + k * n;
          double r = 0.;
          for (int j = 0; j < nm; j++) r += p1[j] * pp[j];
          r *= s;
          for (int j = 0; j < nm; j++) pp[j] -= r * p1[j];
        }
        for (int j = 1; j < nm; j++) p1[j] *= t;
      }
      p1[0] = sv; e[i] = -h;
    }
    if (nm == 1) e[i] = a[i * n + i + 1];
  }
  // ldvmat for n = 3: one reflector acting on coordinates 1, 2
  for (int i = 0; i < 9; i++) V[i] = 0.;
  V[0] = 1.; V[8] = 1.;
  if (a[1] != 0.) {
    const double h = a[1], t2 = a[2];
    V[4] = 1. - h;
    V[7] = -h * t2;
    double s = V[8] * t2;
    s *= h;
    V[8] -= s * t2;
    V[5] = -s;
  } else { V[4] = 1.; V[5] = 0.; V[7] = 0.; }
  // qrbdv on (d, e), rotations applied to the columns of V only
  int m = n;
  double t = std::fabs(d[0]);
  for (int j = 1; j < m; j++) { const double s = std::fabs(d[j]) + std::fabs(e[j - 1]); if (s > t) t = s; }
  t *= 1.e-15;
  const int nmax = 100 * m;
  for (int j = 0; m > 1 && j < nmax; j++) {
    int k;
    for (k = m - 1; k > 0; --k) {
      if (std::fabs(e[k - 1]) < t) break;
      if (std::fabs(d[k - 1]) < t) {
        double s = 1., c = 0.;
        for (int i = k; i < m; i++) {
          const double aa = s * e[i - 1], bb = d[i];
          e[i - 1] *= c;
          const double uu = std::sqrt(aa * aa + bb * bb);
          d[i] = uu; s = -aa / uu; c = bb / uu;   // (the reference rotates the left vectors here; not needed)
        }
        break;
      }
    }
    double y = d[k], x = d[m - 1], uu = e[m - 2];
    double aa = (y + x) * (y - x) - uu * uu, s = y * e[k], bb = s + s;
    uu = std::sqrt(aa * aa + bb * bb);
    if (uu != 0.) {
      double c = std::sqrt((uu + aa) / (uu + uu));
      if (c != 0.) s /= (c * uu); else s = 1.;
      for (int i = k; i < m - 1; i++) {
        bb = e[i];
        if (i > k) {
          aa = s * e[i]; bb *= c;
          e[i - 1] = uu = std::sqrt(x * x + aa * aa);
          c = x / uu; s = aa / uu;
        }
        aa = c * y + s * bb; bb = c * bb - s * y;
        for (int r = 0; r < n; r++) {
          double* p = V + r * n + i;
          const double ww = c * p[0] + s * p[1];
          p[1] = c * p[1] - s * p[0]; p[0] = ww;
        }
        s *= d[i + 1]; d[i] = uu = std::sqrt(aa * aa + s * s);
        y = c * d[i + 1]; c = aa / uu; s /= uu;
        x = c * bb + s * y; y = c * y - s * bb;
      }
    }
    e[m - 2] = x; d[m - 1] = y;
    if (std::fabs(x) < t) --m;
    if (m == k + 1) --m;
  }
  for (int i = 0; i < n; i++) if (d[i] < 0.) for (int r = 0; r < n; r++) V[r * n + i] = -V[r * n + i];
}

inline void mat3_mul(double* c, const double* a, const double* b) {  // matutls/mmul.c
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0.; for (int k = 0; k < 3; k++) s += a[i * 3 + k] * b[k * 3 + j]; c[i * 3 + j] = s; }
}
inline void mat3_tr(double* a, const double* b) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) a[i * 3 + j] = b[j * 3 + i]; }
inline void skew_sym(const double* a, double* ax) { ax[0] = 0; ax[1] = -a[2]; ax[2] = a[1]; ax[3] = a[2]; ax[4] = 0; ax[5] = -a[0]; ax[6] = -a[1]; ax[7] = a[0]; ax[8] = 0; }

// DegUtils.c:94-170: homography compatible with F through three correspondences (Hartley & Zisserman, result 13.6)
inline void Hdetect(const double* F, const double* u7, const unsigned char* IDXS, double* H) {
  double V[9], ec[3], Ex[9], A[9], Ft[9], u3a[9], u3b[9], Au3b[9], u3aT[9], u3bT[9], p1T[9], p1[9], p2[9], b[3];
  mat3_tr(Ft, F);
  svd3_ccmath_V(F, V);
  ec[0] = V[2]; ec[1] = V[5]; ec[2] = V[8];
  skew_sym(ec, Ex);
  mat3_mul(A, Ex, Ft);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { u3a[i + j * 3] = u7[IDXS[i] * 6 + j]; u3b[i + j * 3] = u7[IDXS[i] * 6 + j + 3]; }
  mat3_mul(Au3b, A, u3b);
  mat3_tr(u3aT, u3a); mat3_tr(u3bT, Au3b);
  cross3(p1T, u3aT, u3bT); cross3(p1T + 3, u3aT + 3, u3bT + 3); cross3(p1T + 6, u3aT + 6, u3bT + 6);
  mat3_tr(p1, p1T);
  for (int i = 0; i < 9; i++) Ex[i] *= -1;
  mat3_mul(p2, Ex, u3a);
  for (int c = 0; c < 3; c++)
    b[c] = (p1[c] * p2[c] + p1[3 + c] * p2[3 + c] + p1[6 + c] * p2[6 + c]) / (p2[c] * p2[c] + p2[3 + c] * p2[3 + c] + p2[6 + c] * p2[6 + c]);
  mat3_tr(u3bT, u3b);
  const bool sing = !mb2_minv3_impl(u3bT);
  double x[3];
  for (int j = 0; j < 3; j++) { double z = 0.; for (int k = 0; k < 3; k++) z += u3bT[j * 3 + k] * b[k]; x[j] = z; }
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { const double z = 0. + ec[i] * x[j]; H[i + j * 3] = A[i * 3 + j] - z; }
  if (std::isnan(H[0]) || std::isinf(H[0]) || sing) { H[1] = H[2] = H[3] = H[5] = H[6] = H[7] = 0; H[0] = H[4] = H[8] = 1; }
}
// DegUtils.c:42-83
inline int checksample(const double* F, const double* u7, double th, double* H) {
  static const unsigned char IDXS[5][3] = {{0, 1, 2}, {3, 4, 5}, {0, 1, 6}, {3, 4, 6}, {2, 5, 6}};
  for (int i = 0; i < 5; ++i) {
    double Ds[7], sDs[7];
    unsigned char idx[7];
    int inl[7];
    Hdetect(F, u7, IDXS[i], H);
    for (int j = 0; j < 7; j++) Ds[j] = hds_point(u7 + 6 * j, H);
    std::memcpy(sDs, Ds, sizeof sDs);           // sortDs :173-193 (exchange sort, keeps its tie order)
    for (int a = 0; a < 7; ++a) idx[a] = (unsigned char)a;
    for (int a = 0; a < 7; ++a)
      for (int c = a + 1; c < 7; ++c)
        if (sDs[c] < sDs[a]) { std::swap(sDs[c], sDs[a]); std::swap(idx[c], idx[a]); }
    for (int j = 0; j < 5; ++j) inl[j] = idx[j];
    u2h(u7, inl, 5, H);
    int inlCount = 0;
    for (int j = 0; j < 7; ++j) if (hds_point(u7 + 6 * j, H) < th) ++inlCount;
    if (inlCount > 4) return 1;
  }
  return 0;
}

inline double pred(double x) { return std::nextafter(x, -INFINITY); }  // #{d < x} == #{d <= pred(x)}

// ------------------------------------------------------------------------------------------------------------------
template <class Scorer>
struct RansacF {
  Scorer& sc;
  const double* u; int len; double th; int errorType, doSymCheck, do_lo; unsigned inlLimit;
  int which_f, which_ex;
  int rc = 0;
  LibcRand rng;

  struct Buf { std::vector<double> d; double f[9]; bool valid = false, tagged = false; };
  struct BufSet {
    Buf b[4]; int errs[5] = {0, 1, 2, 3, 3};
  };
  BufSet main;
  HashTable ht;

  RansacF(Scorer& s, const double* u_, int len_, double th_, int errorType_, int doSymCheck_, int do_lo_, unsigned inlLimit_)
      : sc(s), u(u_), len(len_), th(th_), errorType(errorType_), doSymCheck(doSymCheck_), do_lo(do_lo_), inlLimit(inlLimit_) {
    which_f = errorType == 0 ? W_FDS : W_FDSSYM;
    which_ex = errorType == 0 ? W_FDS : W_EXFDSSYM;
  }

  // ---- scorer services (slot 0 = all correspondences)
  void resid(int which, const double* model, double* out, Score* S = nullptr, double th_ = 0, int slot = 0) {
    int r = sc.resid(slot, which, model, th_, out, S);
    if (r < 0 && rc == 0) rc = r;
  }
  void eval_into(BufSet& B, int id, int which, const double* f) {
    Buf& b = B.b[id];
    b.d.resize(len);
    resid(which, f, b.d.data());
    std::memcpy(b.f, f, sizeof b.f); b.valid = true; b.tagged = true;
  }
  void tag(int id, const double* f) { Buf& b = main.b[id]; std::memcpy(b.f, f, sizeof b.f); b.valid = false; b.tagged = true; }
  const double* data(int id) {
    Buf& b = main.b[id];
    if (!b.valid) {
      b.d.assign(len, 0.0);
      if (b.tagged) resid(which_f, b.f, b.d.data());
      b.valid = true;
    }
    return b.d.data();
  }
  int* randsubset(int* pool, int max_sz, int siz) {  // rtools.c:25-39
    for (int i = 0; i < siz; i++) {
      int s = rng.next() % (max_sz - i), j = max_sz - i - 1;
      int q = pool[s]; pool[s] = pool[j]; pool[j] = q;
    }
    return pool + max_sz - siz;
  }
  void u2f(const double* uu, const int* inl, int n, double* F) { u2f_w(uu, inl, nullptr, n, F); }

  // D3 "detached" subset rule of exp_iterFcustom (:631-649, :690-708)
  unsigned detached(unsigned have) const {
    unsigned dc = (unsigned)(int)(have * 1);   // D3_F_RATIO 1
    if (dc > inlLimit) dc = inlLimit;           // D3_F_MIN 0 never binds
    if (dc < 8) dc = 8;
    return dc;
  }

  // exp_ranF.c:610-736
  Score iterF(int* inliers, double ths, int iters, double* F, int iterID) {
    int d = main.errs[1];
    double f[9];
    Score S = {0, 0}, Ss, maxS;
    std::vector<double> w;
    const bool sym = errorType != 0;
    const double dth = (ths - th) / ILSQ_ITERS;
    maxS = inlidxs(data(main.errs[4]), len, th, inliers);
    if (maxS.I < 8) return S;
    S = inlidxs(data(main.errs[4]), len, th * MWM, inliers);
    std::memcpy(f, F, sizeof f);  // (u2f with < 8 points leaves the model as it is; the reference's f is uninitialised there)
    {
      const unsigned dc = detached(S.I);
      if (dc >= S.I) u2f(u, inliers, S.I, f);
      else { int* sub = randsubset(inliers, S.I, dc); u2f(u, sub, dc, f); }
    }
    for (int it = 0; it < iters; it++) {
      eval_into(main, d, which_ex, f);      // EXFDS1(u, f, d, w, len); the weights are formed on demand below
      const double* dd = main.b[d].d.data();
      S = inlidxs(dd, len, th, inliers);
      const uint32_t hash = SuperFastHash((const char*)inliers, S.I * sizeof(*inliers));
      const int ret = ht.contains(hash, S.I, iterID);
      if (ret != -1 && ret != iterID) { S.I = 0; S.J = 0; return S; }
      if (ret == -1) ht.insert(hash, S.I, iterID);
      if (scoreLess(maxS, S)) {
        maxS = S;
        main.errs[1] = main.errs[0]; main.errs[0] = d; d = main.errs[1];
        std::memcpy(F, f, sizeof f);
      }
      Ss = inlidxs(data(d), len, ths * MWM, inliers);
      if (Ss.I < 8) return maxS;
      {
        const unsigned dc = detached(Ss.I);
        int* use = inliers; unsigned cnt = Ss.I;
        if (dc < Ss.I) { use = randsubset(inliers, Ss.I, dc); cnt = dc; }
        w.assign(len, 0.0);
        double fw[9];
        std::memcpy(fw, f, sizeof fw);
        for (unsigned q = 0; q < cnt; q++) w[use[q]] = exfds_weight(u + 6 * use[q], fw, sym);
        u2f_w(u, use, w.data(), cnt, f);
      }
      ths -= dth;
    }
    eval_into(main, d, which_f, f);
    S = inlidxs(main.b[d].d.data(), len, th, inliers);
    if (scoreLess(maxS, S)) {
      maxS = S;
      main.errs[1] = main.errs[0]; main.errs[0] = d;
      std::memcpy(F, f, sizeof f);
    }
    return maxS;
  }

  // exp_ranF.c:739-790
  Score inFrani(int* inliers, int ninl, double* F, int* iterID) {
    Score S = {0, 0}, maxS = {0, 0};
    double f[9];
    if (ninl < 16) return maxS;
    std::vector<int> intbuff(len);
    unsigned ssiz = ninl / 2;
    if (ssiz > 14) ssiz = 14;
    std::swap(main.errs[2], main.errs[0]);
    std::memcpy(f, F, sizeof f);
    for (int i = 0; i < RAN_REP; i++) {
      int* sample = randsubset(inliers, ninl, ssiz);
      u2f(u, sample, ssiz, f);
      eval_into(main, main.errs[0], which_f, f);
      main.errs[4] = main.errs[0];
      S = iterF(intbuff.data(), TC * th, ILSQ_ITERS, f, ++*iterID);
      if (rc < 0) return maxS;
      if (scoreLess(maxS, S)) {
        maxS = S;
        std::swap(main.errs[2], main.errs[0]);
        std::memcpy(F, f, sizeof f);
      }
    }
    std::swap(main.errs[2], main.errs[0]);
    return maxS;
  }

  // __LSQ_BEFORE_LO__ block + exp_inFranicustom (exp_ranF.c:1014-1031 / 1141-1158); src = errs[4] or errorsBest
  Score local_optimisation(const double* src, int* inliers, double* f, int* iterID) {
    const int d = main.errs[0];
    Score S = inlidxs(src, len, TC * th * MWM, inliers);
    u2f(u, inliers, S.I, f);
    eval_into(main, d, which_f, f);
    S = inlidxs(main.b[d].d.data(), len, th, inliers);
    return inFrani(inliers, S.I, f, iterID);
  }

  // ---- DEGENSAC: homography side (ranH.c iterH / inHrani through DegUtils.c innerH)
  Score iterH_deg(BufSet& B, int* inliers, double thh, double ths, double* H, unsigned lim) {
    int d = B.errs[1];
    double h[9];
    Score S = {0, 0}, Ss, maxS;
    const double dth = (ths - thh) / ILSQ_ITERS;
    maxS = inlidxs(B.b[B.errs[4]].d.data(), len, thh, inliers);
    if (maxS.I < 4) return S;
    std::memcpy(h, H, sizeof h);
    if (maxS.I <= lim) u2h(u, inliers, maxS.I, h);
    else { int* sub = randsubset(inliers, maxS.I, lim); u2h(u, sub, lim, h); }
    for (int it = 0; it < ILSQ_ITERS; ++it) {
      eval_into(B, d, W_HDS, h);
      S = inlidxs(B.b[d].d.data(), len, thh, inliers);
      Ss = inlidxs(B.b[d].d.data(), len, ths, inliers);
      if (scoreLess(maxS, S)) {
        maxS = S;
        B.errs[1] = B.errs[0]; B.errs[0] = d; d = B.errs[1];
        std::memcpy(H, h, sizeof h);
      }
      if (Ss.I < 4) return maxS;
      if (Ss.I <= lim) u2h(u, inliers, Ss.I, h);
      else { int* sub = randsubset(inliers, Ss.I, lim); u2h(u, sub, lim, h); }
      ths -= dth;
    }
    eval_into(B, d, W_HDS, h);
    S = inlidxs(B.b[d].d.data(), len, thh, inliers);
    if (scoreLess(maxS, S)) {
      maxS = S;
      B.errs[1] = B.errs[0]; B.errs[0] = d;
      std::memcpy(H, h, sizeof h);
    }
    return maxS;
  }
  // DegUtils.c:720-760 (the `iters` argument lands in inHrani's inlLimit)
  unsigned innerH(double* H, double thh, unsigned lim, unsigned char* inl) {
    BufSet B;
    std::vector<int> inliers(len), intbuff(len);
    for (int i = 0; i < 4; i++) B.b[i].d.assign(len, 0.0);
    eval_into(B, B.errs[0], W_HDS, H);
    Score S = inlidxs(B.b[B.errs[0]].d.data(), len, thh, inliers.data());
    const int ninl = S.I;
    if (ninl >= 8) {  // inHrani, ranH.c:88-137
      Score maxS = {0, 0};
      double h[9];
      int ssiz = ninl / 2;
      if (ssiz > 12) ssiz = 12;
      std::swap(B.errs[2], B.errs[0]);
      std::memcpy(h, H, sizeof h);
      for (int i = 0; i < RAN_REP; ++i) {
        int* sample = randsubset(inliers.data(), ninl, ssiz);
        u2h(u, sample, ssiz, h);
        eval_into(B, B.errs[0], W_HDS, h);
        B.errs[4] = B.errs[0];
        S = iterH_deg(B, intbuff.data(), thh, TC * thh, h, lim);
        if (scoreLess(maxS, S)) {
          maxS = S;
          std::swap(B.errs[2], B.errs[0]);
          std::memcpy(H, h, sizeof h);
        }
      }
      std::swap(B.errs[2], B.errs[0]);
    }
    const double* d = B.b[B.errs[0]].d.data();
    unsigned I = 0;
    for (int j = 0; j < len; j++) { inl[j] = d[j] <= thh ? 1 : 0; I += inl[j]; }
    return I;
  }

  // DegUtils.c:656-717
  unsigned u2Fit(double* F, unsigned char* inl, double th0, double ths, unsigned iters, std::vector<double>& Ds, std::vector<int>& inlI) {
    const double dth = (ths - th0) / (iters - 1);
    unsigned no_i;
    for (unsigned iter = 0; iter < iters; ++iter) {
      resid(W_FDS, F, Ds.data());
      no_i = 0;
      for (int i = 0; i < len; ++i) { inl[i] = Ds[i] < ths ? 1 : 0; no_i += inl[i]; }
      if (no_i < 8) return no_i;
      no_i = 0;
      for (int i = 0; i < len; ++i) if (inl[i]) inlI[no_i++] = i;
      u2f(u, inlI.data(), no_i, F);
      ths -= dth;
    }
    resid(W_FDS, F, Ds.data());
    no_i = 0;
    for (int i = 0; i < len; ++i) { inl[i] = Ds[i] < th0 ? 1 : 0; no_i += inl[i]; }
    return no_i;
  }
  // DegUtils.c:487-612 with dual_sample :615-653
  void innerFH(const std::vector<double>& uH, unsigned lenH, const std::vector<double>& uO, unsigned lenO, double th0, unsigned repCount,
               unsigned sH, unsigned sO, double* F, unsigned char* inl) {
    std::vector<unsigned char> v(len);
    std::vector<double> usam(6 * (sH + sO)), Ds(len);
    std::vector<int> allInl(sH + sO), inlI(len);
    std::vector<unsigned> ptrA(lenH), ptrB(lenO);
    double aF[9];
    for (unsigned i = 0; i < sH + sO; ++i) allInl[i] = i;
    for (int i = 0; i < 9; ++i) { F[i] = 1; aF[i] = 1; }
    for (int i = 0; i < len; ++i) inl[i] = 0;
    unsigned max_i = 0, max_s = 0;
    for (unsigned rep = 0; rep < repCount; ++rep) {
      for (unsigned i = 0; i < lenH; ++i) ptrA[i] = i;
      for (unsigned i = 0; i < lenO; ++i) ptrB[i] = i;
      for (unsigned pos = 0; pos < sH; ++pos) { unsigned idx = rng.next() % lenH; std::swap(ptrA[pos], ptrA[idx]); }
      for (unsigned pos = 0; pos < sO; ++pos) { unsigned idx = rng.next() % lenO; std::swap(ptrB[pos], ptrB[idx]); }
      for (unsigned i = 0; i < sH; ++i) std::memcpy(&usam[6 * i], &uH[6 * ptrA[i]], 6 * sizeof(double));
      for (unsigned i = 0; i < sO; ++i) std::memcpy(&usam[6 * (i + sH)], &uO[6 * ptrB[i]], 6 * sizeof(double));
      u2f(usam.data(), allInl.data(), sH + sO, aF);
      resid(W_FDS, aF, Ds.data());
      unsigned no_i = 0;
      for (int i = 0; i < len; ++i) { v[i] = Ds[i] < th0 ? 1 : 0; no_i += v[i]; }
      if (max_i < no_i) { std::memcpy(inl, v.data(), len); std::memcpy(F, aF, sizeof aF); max_i = no_i; }
      if (no_i > max_s) {
        max_s = no_i;
        no_i = u2Fit(aF, v.data(), th0, th0 * 3, 4, Ds, inlI);
        if (max_i < no_i) { std::memcpy(inl, v.data(), len); std::memcpy(F, aF, sizeof aF); max_i = no_i; }
      }
      if (rc < 0) return;
    }
  }
  // DegUtils.c:263-445: plane-and-parallax -- epipoles from pairs of off-plane correspondences, F = [e]x H
  unsigned rFtH(const unsigned char* hinl, const double* H, double* F) {
    std::vector<double> Ds(len);
    std::vector<unsigned char> nhinl(len), inl(len);
    resid(W_HDS, H, Ds.data());
    unsigned nhinlCount = 0, hinlCount = 0;
    for (int i = 0; i < len; ++i) { nhinl[i] = Ds[i] > 100 * th ? 1 : 0; nhinlCount += nhinl[i]; if (hinl[i]) ++hinlCount; }
    std::vector<double> uN(6 * (size_t)nhinlCount), us(6 * (size_t)nhinlCount), uV, uH(6 * (size_t)hinlCount);
    nhinlCount = 0; hinlCount = 0;
    for (int i = 0; i < len; ++i) {
      if (nhinl[i]) {
        std::memcpy(&uN[6 * nhinlCount], u + 6 * i, 6 * sizeof(double));
        std::memcpy(&us[6 * nhinlCount], u + 6 * i, 3 * sizeof(double));
        us[6 * nhinlCount + 3] = H[0] * u[6 * i + 3] + H[3] * u[6 * i + 4] + H[6] * u[6 * i + 5];
        us[6 * nhinlCount + 4] = H[1] * u[6 * i + 3] + H[4] * u[6 * i + 4] + H[7] * u[6 * i + 5];
        us[6 * nhinlCount + 5] = H[2] * u[6 * i + 3] + H[5] * u[6 * i + 4] + H[8] * u[6 * i + 5];
        ++nhinlCount;
      }
      if (hinl[i]) { std::memcpy(&uH[6 * hinlCount], u + 6 * i, 6 * sizeof(double)); ++hinlCount; }
    }
    unsigned max_i = 3, m_i = 4, max_sam = 10000;
    const double conf = .999;
    if (nhinlCount < 4 || hinlCount < 6) return 0;
    std::vector<unsigned> ptr(nhinlCount);
    for (unsigned i = 0; i < nhinlCount; ++i) ptr[i] = i;
    if (sc.set_points(1, uN.data(), (int)nhinlCount) < 0) { rc = -1; return 0; }
    double Ht[9];
    mat3_tr(Ht, H);
    auto draw = [&](LibcRand& g, std::vector<unsigned>& pt, double* aFt) {  // one 2-point sample -> F = [e]x H (stored as FDs wants it)
      for (unsigned pos = 0; pos < 2; ++pos) { unsigned idx = pos + 1 + g.next() % (nhinlCount - pos - 1); std::swap(pt[pos], pt[idx]); }
      double c1[3], c2[3], ec[3], S[9], SH[9];
      cross3(c1, &us[6 * pt[0]], &us[6 * pt[0] + 3]);
      cross3(c2, &us[6 * pt[1]], &us[6 * pt[1] + 3]);
      cross3(ec, c1, c2);
      const double nrm = std::sqrt(ec[0] * ec[0] + ec[1] * ec[1] + ec[2] * ec[2]);
      ec[0] = ec[0] / nrm; ec[1] = ec[1] / nrm; ec[2] = ec[2] / nrm;
      skew_sym(ec, S);
      mat3_mul(SH, S, Ht);
      mat3_tr(aFt, SH);
    };
    const double th2 = pred(th * 2);
    std::vector<double> models, DsN(nhinlCount);
    std::vector<int> cnt;
    std::vector<unsigned char> v(nhinlCount);
    unsigned no_sam = 1;
    int batch = 64;
    while (no_sam < 2 * max_sam) {
      const unsigned want = std::min<unsigned>((unsigned)batch, 2 * max_sam - no_sam);
      LibcRand g = rng;
      std::vector<unsigned> pt = ptr;
      models.resize((size_t)want * 9); cnt.assign(want, 0);
      for (unsigned b = 0; b < want; b++) draw(g, pt, &models[(size_t)b * 9]);
      int r = sc.score(1, W_FDS, models.data(), (int)want, th2, cnt.data(), nullptr);
      if (r < 0) { rc = r; return 0; }
      unsigned hit = want;
      for (unsigned b = 0; b < want; b++) if ((unsigned)cnt[b] > m_i) { hit = b; break; }
      if (hit == want) { rng = g; ptr.swap(pt); no_sam += want; if (batch < 2048) batch *= 2; continue; }
      // rewind to the improving sample and run the reference's inner steps on it
      double aFt[9];
      for (unsigned b = 0; b <= hit; b++) draw(rng, ptr, aFt);
      no_sam += hit + 1;
      resid(W_FDS, aFt, DsN.data(), nullptr, 0, 1);
      unsigned no_i = 0;
      for (unsigned i = 0; i < nhinlCount; ++i) { v[i] = DsN[i] < th * 2 ? 1 : 0; no_i += v[i]; }
      if (no_i > m_i) {
        uV.resize(6 * (size_t)no_i);
        no_i = 0;
        for (unsigned i = 0; i < nhinlCount; ++i) if (v[i]) { std::memcpy(&uV[6 * no_i], &uN[6 * i], 6 * sizeof(double)); ++no_i; }
        m_i = no_i;
        double aF[9];
        innerFH(uH, hinlCount, uV, no_i, th, 15, 6, 4, aF, inl.data());
        if (rc < 0) return 0;
        unsigned ninl = 0;
        for (int i = 0; i < len; ++i) if (inl[i]) ++ninl;
        if (ninl > max_i) {
          max_i = ninl;
          std::memcpy(F, aF, sizeof aF);
          unsigned maxni = 0;
          for (int i = 0; i < len; ++i) if (inl[i] && nhinl[i]) ++maxni;
          const unsigned ns = (unsigned)nsamples((int)maxni, (int)nhinlCount, 2, conf);
          max_sam = max_sam > ns ? ns : max_sam;
        }
      }
    }
    return max_i;
  }
};

struct Result { int I; int samples; int lo; int Ih; double J; };

// exp_ranF.c:795-1192.  F: 9 doubles out; inl: len flags out.
template <class Scorer>
int ransac_f(Scorer& sc, const double* u, int len, double th, double conf, int max_sam, int errorType, int doSymCheck, int do_lo,
             unsigned inlLimit, long seed, double* F, unsigned char* inl, Result* res) {
  RansacF<Scorer> R(sc, u, len, th, errorType, doSymCheck, do_lo, inlLimit);
  for (int i = 0; i < 9; i++) F[i] = 0;
  for (int i = 0; i < len; i++) inl[i] = 0;
  if (res) { res->I = 0; res->samples = 0; res->lo = 0; res->Ih = 0; res->J = 0; }
  if (len < 8) return 0;   // (the reference would spin on an empty pool; LORANSACFiltering only calls with >= MIN_POINTS)
  if (sc.set_points(0, u, len) < 0) return -1;

  std::vector<int> pool(len), inliers(len);
  for (int i = 0; i < len; i++) pool[i] = i;
  int* samidx = pool.data() + len - 7;
  int samidxBest[7] = {0, 0, 0, 0, 0, 0, 0};
  double f[9] = {0}, H[9] = {0}, FBest[9] = {0}, u7[42];
  Score maxS = {8, 0}, maxSs = {8, 0}, S = {0, 0};
  int no_sam = 0, iter_cnt = 0, degen_cnt = 0, iterID = 0, Ihmax = 0, last_i = 0;
  unsigned non_degen_samples_count = 0;
  bool new_max = false, bad_model = false, any_accept = false;
  // errorsBest: a copy of the residuals of FBest (FDS1), materialised when read
  bool haveBest = false;
  std::vector<double> errorsBest;

  R.rng.seed((unsigned)seed);                       // srand(time(NULL)) in the reference (:827)
  unsigned cur_seed = (unsigned)R.rng.next();      // seed = rand()

  struct Samp { unsigned seed_before; int sam[7]; int nsol; double f[3][9]; unsigned char ori[3]; int first_model; };
  std::vector<Samp> batch;
  std::vector<double> models;
  std::vector<int> bI; std::vector<double> bJ;
  size_t bpos = 0;
  int batch_size = 64;
  LibcRand gen;   // generator used only for drawing samples
  auto refill = [&](int want) -> int {
    batch.clear(); models.clear(); bpos = 0;
    for (int k = 0; k < want; k++) {
      Samp sm; sm.seed_before = cur_seed; sm.nsol = -1; sm.first_model = (int)(models.size() / 9);
      double A[81], sol[81];
      int nb[18];
      gen.seed(cur_seed);
      for (int i = 0; i < 7; i++) {   // rsampleT(Z, 9, pool, 7, len, A): rows of lin_fm (Ftools.c:14-38)
        const int s = gen.next() % (len - i), j = len - i - 1;
        const int q = pool[s]; pool[s] = pool[j]; pool[j] = q;
        const double* p = u + 6 * q;
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) A[i * 9 + a * 3 + b] = p[a + 3] * p[b];
      }
      cur_seed = (unsigned)gen.next();
      std::memcpy(sm.sam, samidx, sizeof sm.sam);
      for (int i = 63; i < 81; ++i) A[i] = 0.0;
      std::memset(sol, 0, sizeof sol);
      const int nullsize = nullspace(A, sol, 9, nb);
      if (nullsize == 2) {
        double poly[4], roots[3];
        double* f1 = sol; double* f2 = sol + 9;
        slcm(f1, f2, poly);
        sm.nsol = rroots3(poly, roots);
        for (int i = 0; i < sm.nsol; i++) {
          for (int j = 0; j < 9; j++) sm.f[i][j] = f1[j] * roots[i] + f2[j] * (1 - roots[i]);
          sm.ori[i] = (unsigned char)all_ori_valid(sm.f[i], u, sm.sam, 7);
          if (sm.ori[i]) models.insert(models.end(), sm.f[i], sm.f[i] + 9);
        }
      }
      batch.push_back(sm);
    }
    const int K = (int)(models.size() / 9);
    bI.assign(std::max(K, 1), 0); bJ.assign(std::max(K, 1), 0.0);
    if (K > 0) { int r = sc.score(0, R.which_f, models.data(), K, th, bI.data(), bJ.data()); if (r < 0) return r; }
    return 0;
  };

  // The reference's generator state after drawing sample `sm` (srand(seed_k), 7 draws, 1 draw for the next seed); whatever runs
  // later in the same iteration (innerH, rFtH, LO) continues from there.
  bool synced = false; unsigned sync_seed = 0; bool any_sample = false;
  auto sync = [&]() { if (!synced && any_sample) { R.rng.seed(sync_seed); for (int i = 0; i < 8; i++) R.rng.next(); synced = true; } };

  // DEGENSAC branch shared by the main loop (:926-980) and the final block (:1083-1136)
  auto count_lt = [&](int which, const double* model, double thr) -> unsigned {
    Score s = {0, 0}; R.resid(which, model, nullptr, &s, pred(thr)); return s.I;
  };
  auto plane_and_parallax = [&](unsigned I, int slot_i) {   // from "if (I > Ihmax)" on; slot_i = the reference's errs[i]
    if ((int)I > Ihmax) Ihmax = (int)I;
    if (I > 6) {
      I = R.rFtH(inl, H, f);
      int d;
      if (I > maxS.I) {
        R.eval_into(R.main, R.main.errs[3], R.which_f, f);
        maxS.I = I; std::memcpy(F, f, sizeof f); new_max = true; any_accept = true;
        d = R.main.errs[3];
      } else {
        const int id = slot_i == 4 ? R.main.errs[4] : R.main.errs[slot_i];
        R.eval_into(R.main, id, R.which_f, f);
        d = id;
      }
      const double* dd = R.main.b[d].d.data();
      double jj = 0;
      for (int j = 0; j < len; j++) jj += truncQuad(dd[j], th);
      if (new_max) maxS.J = jj;
      ++degen_cnt;
    }
  };

  while (no_sam < max_sam) {
    if (bpos >= batch.size()) {
      const int want = std::min(batch_size, std::max(1, max_sam - no_sam));
      const int r = refill(want);
      if (r < 0) return r;
      if (batch_size < 2048) batch_size *= 2;
    }
    const Samp& sm = batch[bpos++];
    no_sam++;
    any_sample = true; synced = false; sync_seed = sm.seed_before;
    if (sm.nsol < 0) continue;
    new_max = false;
    bool do_iterate = false;
    int model_idx = sm.first_model;
    int i;
    for (i = 0; i < sm.nsol; i++) {
      std::memcpy(f, sm.f[i], sizeof f);
      if (!sm.ori[i]) continue;
      const int d = R.main.errs[i];
      R.tag(d, f);
      S.I = (unsigned)bI[model_idx]; S.J = bJ[model_idx]; model_idx++;
      if (scoreLess(maxS, S)) {
        if (doSymCheck) {
          Score sc2 = {0, 0};
          R.resid(W_FDSSYM, f, nullptr, &sc2, F_CHECK_COEF * th);
          const int SI_min = (int)std::floor(F_SYMM_COEF * S.I);
          bad_model = (int)sc2.I <= SI_min;
        }
        if (bad_model) continue;
        R.main.errs[i] = R.main.errs[3]; R.main.errs[3] = d;
        maxS = S; std::memcpy(F, f, sizeof f); new_max = true; any_accept = true;
      }
      if (scoreLess(maxSs, S)) {
        maxSs = S;
        for (int q = 0; q < 7; q++) std::memcpy(u7 + 6 * q, u + 6 * sm.sam[q], 6 * sizeof(double));
        if (checksample(f, u7, 3 * th, H)) {
          unsigned I = count_lt(W_HDS, H, th * 3);
          if (I < 8) break;
          sync();
          I = R.innerH(H, 16 * th, 10, inl);
          plane_and_parallax(I, i);
        } else {
          do_iterate = (do_lo > 0 && (no_sam > ITER_SAM));
          R.main.errs[4] = d;
          non_degen_samples_count++;
          std::memcpy(samidxBest, sm.sam, sizeof samidxBest);
          std::memcpy(FBest, f, sizeof FBest); haveBest = true;
        }
      }
      if (R.rc < 0) return R.rc;
    }
    last_i = i;
    if (do_lo > 0 && (no_sam == ITER_SAM) && non_degen_samples_count) do_iterate = true;
    if (do_iterate) {
      iter_cnt++;
      sync();
      S = R.local_optimisation(R.data(R.main.errs[4]), inliers.data(), f, &iterID);
      if (R.rc < 0) return R.rc;
      if (scoreLess(maxS, S)) {
        std::swap(R.main.errs[0], R.main.errs[3]);
        maxS = S; std::memcpy(F, f, sizeof f); new_max = true; any_accept = true;
      }
    }
    if (new_max) {
      const int new_sam = nsamples(maxS.I + 1, len, 7, conf);
      if (new_sam < max_sam) max_sam = new_sam;
    }
  }

  // "If there were no LOs, do at least one NOW!" (:1064-1172)
  if (do_lo && (!iter_cnt && !degen_cnt) && non_degen_samples_count) {
    sync();
    for (int q = 0; q < 7; q++) std::memcpy(u7 + 6 * q, u + 6 * samidxBest[q], 6 * sizeof(double));
    if (checksample(FBest, u7, 3 * th, H)) {
      unsigned I = count_lt(W_HDS, H, th * 3);
      if (I >= 8) I = R.innerH(H, 16 * th, 10, inl);
      plane_and_parallax(I, last_i > 3 ? 3 : last_i);
    } else {
      iter_cnt++;
      if (haveBest) { errorsBest.resize(len); R.resid(R.which_f, FBest, errorsBest.data()); } else errorsBest.assign(len, 0.0);
      S = R.local_optimisation(errorsBest.data(), inliers.data(), f, &iterID);
      if (R.rc < 0) return R.rc;
      if (scoreLess(maxS, S)) {
        std::swap(R.main.errs[0], R.main.errs[3]);
        maxS = S; std::memcpy(F, f, sizeof f); any_accept = true;
      }
    }
  }
  if (R.rc < 0) return R.rc;
  if (any_accept) {
    const double* d = R.data(R.main.errs[3]);
    if (R.rc < 0) return R.rc;
    for (int j = 0; j < len; j++) inl[j] = d[j] <= th ? 1 : 0;
  } else for (int j = 0; j < len; j++) inl[j] = 0;
  if (res) { res->I = (int)maxS.I; res->samples = no_sam; res->lo = iter_cnt; res->Ih = Ihmax; res->J = maxS.J; }
  return (int)maxS.I;
}

}  // namespace mb2_rf
