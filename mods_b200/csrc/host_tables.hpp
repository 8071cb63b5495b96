// Host-side constant tables, built exactly as the reference's constructors build them (libm on
// the host, so transcendental values are the reference's own).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace mb2host {

// detectors/helpers.cpp:30-72: atan(i/255) printed with 10 decimals; entries 32, 83 and 100 of the
// reference table end in "55" instead of "00" and are reproduced as they are.
inline void build_atan_lut(double* v) {
  for (int i = 0; i < 256; i++) {
    char buf[32];
    std::snprintf(buf, sizeof buf, "%.10f", std::atan(i / 255.0));
    v[i] = std::strtod(buf, nullptr);
  }
  v[32] = 0.1248376255; v[83] = 0.3146752558; v[100] = 0.3737268255;
}

// atan2LUTff (detectors/helpers.cpp:160-207) has 8 branches x 256 table values + the x == 0 corner: its result, and everything
// the hot kernels compute from that result alone, is a function of (branch, table index).  Built here with the very expressions the
// kernels used to evaluate per pixel (IEEE double / float on the host, no contraction):
//   sift_o[code]   = (float)(8.0 * ((double)ori + M_PI_DOUBLED) / M_PI_DOUBLED)          siftdesc.cpp:99-107 (orientation bin coordinate)
//   ori_bin[code]  = (int)(36.f * (ori / PIf + 1.f) * 0.5f)                              synth-detection.cpp:786 (36-bin histogram)
// code = branch * 256 + index; branch = (x > 0) * 4 + (y > 0) * 2 + (|x| > |y|); code 2048 = the x == 0 corner (ori = 0).
constexpr int ATAN_CODES = 2049;
inline float atan_branch_value(int branch, double t) {
  const double PI_2d = (double)1.57079632679489661923f, PId = (double)3.14159265358979323846f;
  switch (branch) {
    case 7: return (float)t;             // x > 0, y > 0, big
    case 6: return (float)(PI_2d - t);   // x > 0, y > 0
    case 5: return (float)(-t);          // x > 0, y <= 0, big
    case 4: return (float)(-PI_2d + t);  // x > 0, y <= 0
    case 3: return (float)(PId - t);     // x <= 0, y > 0, big
    case 2: return (float)(PI_2d + t);   // x <= 0, y > 0
    case 1: return (float)(-PId + t);    // x <= 0, y <= 0, big
    default: return (float)(-PI_2d - t); // x < 0, y <= 0
  }
}
inline void build_atan_derived(const double* lut, float* sift_o, unsigned char* ori_bin) {
  const double M_PI_DOUBLED = 6.28318530718;
  const float PIf = 3.14159265358979323846f;
  for (int code = 0; code < ATAN_CODES; code++) {
    const float ori = code == 2048 ? 0.f : atan_branch_value(code >> 8, lut[code & 255]);
    sift_o[code] = (float)(8.0 * ((double)ori + M_PI_DOUBLED) / M_PI_DOUBLED);
    volatile float a = ori / PIf; volatile float b = a + 1.0f; volatile float c = 36.f * b; volatile float d = c * 0.5f;
    ori_bin[code] = (unsigned char)(int)d;
  }
}

// computeGaussMask, detectors/helpers.cpp:411-440
inline void gauss_mask(float* mask, int size) {
  int halfSize = size >> 1;
  float scale = float(halfSize) / 3.0f;
  float scale2 = -2.0f * scale * scale;
  std::vector<float> tmp(halfSize + 1);
  for (int i = 0; i <= halfSize; i++) tmp[i] = std::exp((float(i * i) / scale2));
  int endSize = int(std::ceil(scale * 5.0f) - halfSize);
  for (int i = 1; i < endSize; i++) tmp[halfSize - i] += std::exp((float((i + halfSize) * (i + halfSize)) / scale2));
  for (int i = 0; i <= halfSize; i++)
    for (int j = 0; j <= halfSize; j++) {
      float v = tmp[i] * tmp[j];
      mask[(i + halfSize) * size + (-j + halfSize)] = v;
      mask[(-i + halfSize) * size + (j + halfSize)] = v;
      mask[(i + halfSize) * size + (j + halfSize)] = v;
      mask[(-i + halfSize) * size + (-j + halfSize)] = v;
    }
}

// computeCircularGaussMask, detectors/helpers.cpp:442-461
inline void circular_gauss_mask(float* mask, int size, float sigma = 0) {
  int halfSize = size >> 1;
  float r2 = float(halfSize * halfSize);
  float sigma2 = (sigma == 0) ? 0.9f * r2 : 2 * sigma * sigma;
  float* mp = mask;
  for (int i = 0; i < size; i++)
    for (int j = 0; j < size; j++) {
      float disq = float((i - halfSize) * (i - halfSize) + (j - halfSize) * (j - halfSize));
      *mp++ = (disq < r2) ? std::exp(-disq / sigma2) : 0;
    }
}

// SIFTDescriptor::precomputeBinsAndWeights, matching/siftdesc.cpp:22-71 (weights are float values)
inline void sift_bins(int patchSize, int spatialBins, int orientationBins, int* bin0, int* bin1, float* w0, float* w1) {
  int halfSize = patchSize >> 1;
  float step = float(spatialBins + 1) / (2 * halfSize);
  for (int i = 0; i < patchSize; i++) {
    float x = step * i;
    int xi = (int)(x);
    bin0[i] = xi - 1; bin1[i] = xi;
    w1[i] = x - xi; w0[i] = 1.0f - w1[i];
    if (bin0[i] < 0) { bin0[i] = 0; w0[i] = 0; }
    if (bin0[i] >= spatialBins) { bin0[i] = spatialBins - 1; w0[i] = 0; }
    if (bin1[i] < 0) { bin1[i] = 0; w1[i] = 0; }
    if (bin1[i] >= spatialBins) { bin1[i] = spatialBins - 1; w1[i] = 0; }
    bin0[i] *= orientationBins; bin1[i] *= orientationBins;
  }
}

// gaussianBlur's kernel size (detectors/helpers.cpp:720-721) and OpenCV 2.4.9 getGaussianKernel(CV_32F)
inline int gauss_ksize(float sigma) {
  int size = (int)(2.0 * 3.0 * sigma + 1.0);
  if (size % 2 == 0) size++;
  return size;
}
inline std::vector<float> gauss_kernel(int n, double sigma) {
  std::vector<float> cf(n);
  double scale2X = -0.5 / (sigma * sigma), sum = 0;
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

inline int cv_round(double v) { return (int)std::nearbyint(v); }

}  // namespace mb2host
