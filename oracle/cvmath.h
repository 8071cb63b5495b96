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
//  * warpAffine(CV_32FC1, INTER_LINEAR, BORDER_CONSTANT) -- synth-detection.cpp:385,426: the forward matrix is inverted
//    in double (closed form), destination -> source coordinates are formed in FIXED POINT (AB_BITS = 10: per-column
//    terms cvRound(M0*x*1024), per-row terms cvRound((M1*y+M2)*1024) + 16), reduced to 1/32 pixel (INTER_BITS = 5), and
//    the four neighbours are blended with the float table w = {(1-fy)(1-fx), (1-fy)fx, fy(1-fx), fy*fx}, fx,fy = k/32,
//    as ((S00*w0 + S01*w1) + S10*w2) + S11*w3; samples outside the source read the border value.
//  * GaussianBlur(ksize (kx, ky), sigmaX, sigmaY, BORDER_REFLECT_101) -- synth-detection.cpp:411: same kernels as above;
//    symmetric row kernels of 3 or 5 taps use 2.4.9's SymmRowSmallFilter order  k0*S0 + k1*(S-1 + S+1) [+ k2*(S-2 + S+2)].
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

// BORDER_REFLECT_101 index (gfedcb|abcdefgh|gfedcba)
inline int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) { if (p < 0) p = -p; else p = 2 * len - 2 - p; }
  return p;
}
// Separable filter with BORDER_REFLECT_101 (cv::GaussianBlur's default border), src and dst may alias.
inline void sep_filter_reflect101(const float* src, float* dst, int rows, int cols, const std::vector<float>& kx, const std::vector<float>& ky) {
  const int nx = (int)kx.size(), hx = nx / 2;
  const int ny = (int)ky.size(), hy = ny / 2;
  std::vector<float> tmp((size_t)rows * cols);
  std::vector<float> padded(cols + 2 * hx);
  for (int r = 0; r < rows; r++) {
    const float* s = src + (size_t)r * cols;
    for (int i = 0; i < cols + 2 * hx; i++) padded[i] = s[reflect101(i - hx, cols)];
    float* t = tmp.data() + (size_t)r * cols;
    for (int c = 0; c < cols; c++) {
      const float* p = padded.data() + c;
      float acc;
      if (nx == 3) acc = p[1] * kx[1] + (p[0] + p[2]) * kx[2];
      else if (nx == 5) acc = p[2] * kx[2] + (p[1] + p[3]) * kx[3] + (p[0] + p[4]) * kx[4];
      else { acc = kx[0] * p[0]; for (int j = 1; j < nx; j++) acc = acc + kx[j] * p[j]; }
      t[c] = acc;
    }
  }
  for (int r = 0; r < rows; r++) {
    float* d = dst + (size_t)r * cols;
    const float* c0 = tmp.data() + (size_t)r * cols;
    for (int c = 0; c < cols; c++) d[c] = ky[hy] * c0[c];
    for (int j = 1; j <= hy; j++) {
      const float* a = tmp.data() + (size_t)reflect101(r + j, rows) * cols;
      const float* b = tmp.data() + (size_t)reflect101(r - j, rows) * cols;
      const float k = ky[hy + j];
      for (int c = 0; c < cols; c++) d[c] = d[c] + k * (a[c] + b[c]);
    }
  }
}

inline int cv_round_i(double v) { return (int)std::nearbyint(v); }   // saturate_cast<int>(double): round half to even

// cv::warpAffine(src, dst, M (forward, 2x3 double), dsize, INTER_LINEAR, BORDER_CONSTANT, borderValue) for one float channel
inline void warp_affine_linear(const float* src, int srows, int scols, const double* Mfwd, float* dst, int drows, int dcols, float cval) {
  double M[6];
  for (int i = 0; i < 6; i++) M[i] = Mfwd[i];
  {  // invertAffineTransform as warpAffine does it in place
    double D = M[0] * M[4] - M[1] * M[3];
    D = D != 0 ? 1. / D : 0;
    const double A11 = M[4] * D, A22 = M[0] * D;
    M[0] = A11; M[1] *= -D; M[3] *= -D; M[4] = A22;
    const double b1 = -M[0] * M[2] - M[1] * M[5];
    const double b2 = -M[3] * M[2] - M[4] * M[5];
    M[2] = b1; M[5] = b2;
  }
  const int AB_BITS = 10, AB_SCALE = 1 << AB_BITS, INTER_BITS = 5, INTER_TAB_SIZE = 1 << INTER_BITS;
  const int round_delta = AB_SCALE / INTER_TAB_SIZE / 2;
  std::vector<int> adelta(dcols), bdelta(dcols);
  for (int x = 0; x < dcols; x++) { adelta[x] = cv_round_i(M[0] * x * AB_SCALE); bdelta[x] = cv_round_i(M[3] * x * AB_SCALE); }
  float tab1[INTER_TAB_SIZE][2];
  for (int i = 0; i < INTER_TAB_SIZE; i++) { const float f = i * (1.f / INTER_TAB_SIZE); tab1[i][0] = 1.f - f; tab1[i][1] = f; }
  for (int y = 0; y < drows; y++) {
    const int X0 = cv_round_i((M[1] * y + M[2]) * AB_SCALE) + round_delta;
    const int Y0 = cv_round_i((M[4] * y + M[5]) * AB_SCALE) + round_delta;
    float* d = dst + (size_t)y * dcols;
    for (int x = 0; x < dcols; x++) {
      const int X = (X0 + adelta[x]) >> (AB_BITS - INTER_BITS), Y = (Y0 + bdelta[x]) >> (AB_BITS - INTER_BITS);
      int sx = X >> INTER_BITS, sy = Y >> INTER_BITS;
      sx = sx < -32768 ? -32768 : (sx > 32767 ? 32767 : sx);   // the map is stored as short
      sy = sy < -32768 ? -32768 : (sy > 32767 ? 32767 : sy);
      const int fx = X & (INTER_TAB_SIZE - 1), fy = Y & (INTER_TAB_SIZE - 1);
      const float w0 = tab1[fy][0] * tab1[fx][0], w1 = tab1[fy][0] * tab1[fx][1], w2 = tab1[fy][1] * tab1[fx][0], w3 = tab1[fy][1] * tab1[fx][1];
      if (sx >= 0 && sy >= 0 && sx < scols - 1 && sy < srows - 1) {
        const float* S = src + (size_t)sy * scols + sx;
        d[x] = S[0] * w0 + S[1] * w1 + S[scols] * w2 + S[scols + 1] * w3;
      } else if (sx >= scols || sx + 1 < 0 || sy >= srows || sy + 1 < 0) {
        d[x] = cval;
      } else {
        auto at = [&](int yy, int xx) { return (xx >= 0 && xx < scols && yy >= 0 && yy < srows) ? src[(size_t)yy * scols + xx] : cval; };
        const float v0 = at(sy, sx), v1 = at(sy, sx + 1), v2 = at(sy + 1, sx), v3 = at(sy + 1, sx + 1);
        d[x] = v0 * w0 + v1 * w1 + v2 * w2 + v3 * w3;
      }
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
