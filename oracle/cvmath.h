// TEST INFRASTRUCTURE (oracle) -- not part of the shipped product path.
//
// CPU restatement of the OpenCV 2.4.9 imgproc arithmetic the reference calls but
// does not vendor (README.md:11 pins "OpenCV 2.4.9"; it is neither in the tree nor
// a submodule).  Call sites in the reference:
//   detectors/helpers.cpp:717-731       cv::GaussianBlur(..., BORDER_REPLICATE)
//   detectors/affinedetectors/pyramid.cpp:520  cv::resize(.., Size(0,0), 0.5, 0.5, INTER_LINEAR)
//   synth-detection.cpp:67,545,868      cv::invert(3x3 double, DECOMP_LU)
//
// Restated from the published 2.4.9 algorithm:
//  * getGaussianKernel(n, sigma, CV_32F): t_i = exp(-0.5/sigma^2 * x_i^2) in double,
//    rounded to float, summed in double, each tap = (float)(tap * (1/sum)).
//  * sepFilter2D for CV_32F with ksize > 5: generic RowFilter (left-to-right
//    s = k0*S0; s += k_j*S_j), then SymmColumnFilter (s = k_c*S_c; s += k_j*(S_+j + S_-j)),
//    float accumulators, no FMA (the reference is built without -march, SSE2 only).
//    The row/column structure was checked against cv2 4.13 (AVX2+FMA build): with FMA
//    emulated and cv2's own kernel it is >99.6% bit identical, 2e-7 max rel
//    (tests/test_oracle_cv.py).
//  * resize(0.5, INTER_LINEAR) on exact factor 2 is re-routed by 2.4.9 to the fast
//    INTER_AREA path: out = (s00 + s01 + s10 + s11) * 0.25f, summed in that order in
//    float; cells that stick out of an odd-sized source average the valid samples.
//    Output size = cvRound(dim * 0.5) (round half to even).
//  * invert 3x3 (n <= 3 closed form): adjugate * (1/det), double.
//
// PARITY UNPINNED: no reference test or golden vector pins these.
#ifndef MB2_ORACLE_CVMATH_H
#define MB2_ORACLE_CVMATH_H

#include <cmath>
#include <cstring>
#include <vector>

namespace cvmath {

inline int gauss_ksize(float sigma) {      // helpers.cpp:720-721
  int size = (int)(2.0 * 3.0 * sigma + 1.0);
  if (size % 2 == 0) size++;
  return size;
}

inline std::vector<float> gauss_kernel(int n, double sigma) {
  std::vector<float> cf(n);
  double scale2X = -0.5 / (sigma * sigma);
  double sum = 0;
  for (int i = 0; i < n; i++) {
    double x = i - (n - 1) * 0.5;
    double t = std::exp(scale2X * x * x);
    cf[i] = (float)t;
    sum += cf[i];
  }
  sum = 1. / sum;
  for (int i = 0; i < n; i++) cf[i] = (float)(cf[i] * sum);
  return cf;
}

// Separable blur, BORDER_REPLICATE, possibly different kernels per axis.
// src and dst may alias.
inline void sep_filter(const float* src, float* dst, int rows, int cols,
                       const std::vector<float>& kx, const std::vector<float>& ky) {
  const int nx = (int)kx.size(), hx = nx / 2;
  const int ny = (int)ky.size(), hy = ny / 2;
  std::vector<float> tmp((size_t)rows * cols);
  std::vector<float> padded(cols + 2 * hx);
  for (int r = 0; r < rows; r++) {
    const float* s = src + (size_t)r * cols;
    for (int i = 0; i < hx; i++) padded[i] = s[0];
    std::memcpy(padded.data() + hx, s, sizeof(float) * cols);
    for (int i = 0; i < hx; i++) padded[hx + cols + i] = s[cols - 1];
    float* t = tmp.data() + (size_t)r * cols;
    for (int c = 0; c < cols; c++) {
      const float* p = padded.data() + c;
      float acc = kx[0] * p[0];
      for (int j = 1; j < nx; j++) acc = acc + kx[j] * p[j];
      t[c] = acc;
    }
  }
  for (int r = 0; r < rows; r++) {
    float* d = dst + (size_t)r * cols;
    const float* c0 = tmp.data() + (size_t)r * cols;
    for (int c = 0; c < cols; c++) d[c] = ky[hy] * c0[c];
    for (int j = 1; j <= hy; j++) {
      int rp = r + j; if (rp > rows - 1) rp = rows - 1;
      int rm = r - j; if (rm < 0) rm = 0;
      const float* a = tmp.data() + (size_t)rp * cols;
      const float* b = tmp.data() + (size_t)rm * cols;
      const float k = ky[hy + j];
      for (int c = 0; c < cols; c++) d[c] = d[c] + k * (a[c] + b[c]);
    }
  }
}

inline void gaussian_blur(const float* src, float* dst, int rows, int cols, float sigma) {
  int n = gauss_ksize(sigma);
  std::vector<float> k = gauss_kernel(n, sigma);
  sep_filter(src, dst, rows, cols, k, k);
}

inline int cv_round(double v) { return (int)std::nearbyint(v); }  // round half to even

inline void half_size(int rows, int cols, int* orows, int* ocols) {
  *orows = cv_round(rows * 0.5);
  *ocols = cv_round(cols * 0.5);
}

inline void resize_half(const float* src, int rows, int cols, float* dst) {
  int orows, ocols;
  half_size(rows, cols, &orows, &ocols);
  for (int r = 0; r < orows; r++) {
    for (int c = 0; c < ocols; c++) {
      int r0 = 2 * r, c0 = 2 * c;
      if (r0 + 1 < rows && c0 + 1 < cols) {
        const float* p = src + (size_t)r0 * cols + c0;
        float sum = 0.f;
        sum += p[0]; sum += p[1]; sum += p[cols]; sum += p[cols + 1];
        dst[(size_t)r * ocols + c] = sum * 0.25f;
      } else {
        float sum = 0.f; int count = 0;
        for (int sy = 0; sy < 2; sy++) {
          if (r0 + sy >= rows) break;
          for (int sx = 0; sx < 2; sx++) {
            if (c0 + sx >= cols) break;
            sum += src[(size_t)(r0 + sy) * cols + c0 + sx];
            count++;
          }
        }
        dst[(size_t)r * ocols + c] = (float)((float)sum / count);
      }
    }
  }
}

inline bool invert3x3(const double* S, double* D) {
  double d = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) +
             S[2] * (S[3] * S[7] - S[4] * S[6]);
  if (d == 0) { for (int i = 0; i < 9; i++) D[i] = 0; return false; }
  d = 1. / d;
  double t[9];
  t[0] = (S[4] * S[8] - S[5] * S[7]) * d;
  t[1] = (S[2] * S[7] - S[1] * S[8]) * d;
  t[2] = (S[1] * S[5] - S[2] * S[4]) * d;
  t[3] = (S[5] * S[6] - S[3] * S[8]) * d;
  t[4] = (S[0] * S[8] - S[2] * S[6]) * d;
  t[5] = (S[2] * S[3] - S[0] * S[5]) * d;
  t[6] = (S[3] * S[7] - S[4] * S[6]) * d;
  t[7] = (S[1] * S[6] - S[0] * S[7]) * d;
  t[8] = (S[0] * S[4] - S[1] * S[3]) * d;
  for (int i = 0; i < 9; i++) D[i] = t[i];
  return true;
}

}  // namespace cvmath
#endif
