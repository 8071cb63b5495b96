// ccmath minv (matutls/minv.c) for n = 3, shared by the device scorers (ransac.cu) and the host-side DEGENSAC logic
// (ransac_f_logic.hpp: Hdetect).  Plain C++; MB2_HD3 adds the CUDA qualifiers under nvcc.
#pragma once
#include <cmath>
#if defined(__CUDACC__)
#define MB2_HD3 __host__ __device__
#else
#define MB2_HD3
#endif
// matutls/minv.c for n = 3, restated with indices (column-wise LU with row pivoting, in place).
MB2_HD3 inline bool mb2_minv3_impl(double* a) {
  const int n = 3;
  int le[3];
  double q0[3], tq = 0., zr = 1.e-15;
#define A_(r, c) a[(r) * n + (c)]
  for (int j = 0; j < n; ++j) {
    if (j > 0) {
      for (int i = 0; i < n; ++i) q0[i] = A_(i, j);
      for (int i = 1; i < n; ++i) {
        int lc = i < j ? i : j;
        double t = 0.;
        for (int k = 0; k < lc; ++k) t += A_(i, k) * q0[k];
        q0[i] -= t;
      }
      for (int i = 0; i < n; ++i) A_(i, j) = q0[i];
    }
    double s = fabs(A_(j, j));
    int lc = j;
    for (int k = j + 1; k < n; ++k) { double t = fabs(A_(k, j)); if (t > s) { s = t; lc = k; } }
    tq = tq > s ? tq : s;
    if (s < zr * tq) return false;
    le[j] = lc;
    if (lc != j) for (int k = 0; k < n; ++k) { double t = A_(j, k); A_(j, k) = A_(lc, k); A_(lc, k) = t; }
    double t = 1. / A_(j, j);
    for (int k = j + 1; k < n; ++k) A_(k, j) *= t;
    A_(j, j) = t;
  }
  for (int j = 1; j < n; ++j) for (int k = 0; k < j; ++k) A_(k, j) *= A_(j, j);
  for (int j = 1; j < n; ++j) {
    for (int i = 0; i < j; ++i) q0[i] = A_(i, j);
    for (int k = 0; k < j; ++k) { double t = 0.; for (int i = k; i < j; ++i) t -= A_(k, i) * q0[i]; q0[k] = t; }
    for (int i = 0; i < j; ++i) A_(i, j) = q0[i];
  }
  for (int j = n - 2; j >= 0; --j) {
    int m = n - j - 1;
    for (int i = 0; i < m; ++i) q0[i] = A_(j + 1 + i, j);
    for (int k = n - 1; k > j; --k) {
      double t = -A_(k, j);
      for (int i = j + 1, q = 0; i < k; ++i, ++q) t -= A_(k, i) * q0[q];
      q0[--m] = t;
    }
    m = n - j - 1;
    for (int i = 0; i < m; ++i) A_(j + 1 + i, j) = q0[i];
  }
  for (int k = 0; k < n - 1; ++k) {
    for (int i = 0; i < n; ++i) q0[i] = A_(i, k);
    for (int j = 0; j < n; ++j) {
      double t; int i;
      if (j > k) { t = 0.; i = j; } else { t = q0[j]; i = k + 1; }
      for (; i < n; ++i) t += A_(j, i) * q0[i];
      q0[j] = t;
    }
    for (int i = 0; i < n; ++i) A_(i, k) = q0[i];
  }
  for (int j = n - 2; j >= 0; --j)
    for (int k = 0; k < n; ++k) { double t = A_(k, j); A_(k, j) = A_(k, le[j]); A_(k, le[j]) = t; }
#undef A_
  return true;
}

