// TEST INFRASTRUCTURE (oracle) -- CPU restatement of the reference's hot path.  NOT shipped,
// NOT linked by the product; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
// may touch anything under oracle/.
//
// Every function cites the reference file:line it follows.  It is pinned, bit for bit, against
// the reference's own sources compiled in place (oracle/_ref/libmods_ref.so, see build_ref.sh)
// by tests/test_oracle_vs_ref.py, and against committed golden vectors generated from that
// build (tests/golden/).  The OpenCV 2.4.9 arithmetic the reference links but does not vendor
// (GaussianBlur / resize / invert) is restated in oracle/cvmath.h and is PARITY UNPINNED by the
// reference (no tests, no golden vectors) -- both the restatement and the
// reference build share that one definition.
//
// Floating-point contract: plain IEEE float/double, no FMA contraction (-ffp-contract=off),
// evaluation order exactly as written in the reference.
#ifndef MB2_MODS_ORACLE_HPP
#define MB2_MODS_ORACLE_HPP

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "cvmath.h"

namespace mo {

struct Image {
  int rows = 0, cols = 0;
  std::vector<float> px;
  Image() {}
  Image(int r, int c) : rows(r), cols(c), px((size_t)r * c, 0.f) {}
  float* row(int r) { return px.data() + (size_t)r * cols; }
  const float* row(int r) const { return px.data() + (size_t)r * cols; }
};

struct Key {  // detectors/structures.hpp:187-196 (AffineKeypoint)
  double x, y, a11, a12, a21, a22, s, response;
  int sub_type;
};

struct HessParams {  // [HessianAffine] of config_iter_mods_cviu.ini; structures.hpp:120-165, affine.h:28-63
  float threshold = 16.0f / 3.0f;
  int numberOfScales = 3;
  float initialSigma = 1.6f;
  float edgeEigenValueRatio = 10.0f;
  int border = 5;
  int maxIterations = 16;
  float convergenceThreshold = 0.05f;
  int smmWindowSize = 19;
  int doBaumberg = 1;
  int mode = 0;  // FIXED_TH
  int detectorType = 0;  // detector_type (structures.hpp): 0 DET_HESSIAN, 1 DET_DOG, 2 DET_HARRIS
  int reg_number = -1;
  float rel_threshold = -1;
  float rel_reg_number = -1;
  int patchSize = 41;
  float mrSize = 3.0f * 1.7320508f;
};

// ------------------------------------------------------------------------------------------
// detectors/helpers.cpp
// ------------------------------------------------------------------------------------------
extern const double* const ATAN_LUT_PTR;  // helpers.cpp:30-72, built in mods_oracle.cpp
#define ATAN_LUT ATAN_LUT_PTR

inline float atan2LUTff(float y, float x) {  // helpers.cpp:160-207
  const float PI_2f = 1.57079632679489661923f, PIf = 3.14159265358979323846f;
  if (x > 0.f) {
    if (y > 0.f) {
      if (x > y) return ATAN_LUT[(int)(255.f * y / x)];
      else return PI_2f - ATAN_LUT[(int)(255 * x / y)];
    } else {
      float absy = std::fabs(y);
      if (x > absy) return -ATAN_LUT[(int)(255.f * absy / x)];
      else return -PI_2f + ATAN_LUT[(int)(255.f * x / absy)];
    }
  } else if (y > 0.f) {
    float absx = std::fabs(x);
    if (absx > y) return PIf - ATAN_LUT[(int)(255.f * y / absx)];
    else return PI_2f + ATAN_LUT[(int)(255.f * absx / y)];
  } else {
    float absx = std::fabs(x), absy = std::fabs(y);
    if (absx > absy) return -PIf + ATAN_LUT[(int)(255.f * absy / absx)];
    else {
      if (x == 0.f) return 0.f;
      return -PI_2f - ATAN_LUT[(int)(255.f * absx / absy)];
    }
  }
}

inline void solveLinear3x3(float* A, float* b) {  // helpers.cpp:309-368
  int i = 0;
  float* pr = A;
  float vp = std::fabs(A[0]);
  float tmp = std::fabs(A[3]);
  if (tmp > vp) { pr = A + 3; i = 1; vp = tmp; }
  if (std::fabs(A[6]) > vp) { pr = A + 6; i = 2; }
  if (pr != A) {
    std::swap(pr[0], A[0]); std::swap(pr[1], A[1]); std::swap(pr[2], A[2]); std::swap(b[i], b[0]);
  }
  vp = A[3] / A[0];
  A[4] -= vp * A[1]; A[5] -= vp * A[2]; b[1] -= vp * b[0];
  vp = A[6] / A[0];
  A[7] -= vp * A[1]; A[8] -= vp * A[2]; b[2] -= vp * b[0];
  if (std::fabs(A[4]) < std::fabs(A[7])) {
    std::swap(A[7], A[4]); std::swap(A[8], A[5]); std::swap(b[2], b[1]);
  }
  vp = A[7] / A[4];
  A[8] -= vp * A[5]; b[2] -= vp * b[1];
  b[2] = (b[2]) / A[8];
  b[1] = (b[1] - A[5] * b[2]) / A[4];
  b[0] = (b[0] - A[2] * b[2] - A[1] * b[1]) / A[0];
}

inline void computeGaussMask(float* mask, int size) {  // helpers.cpp:411-440
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

inline void computeCircularGaussMask(float* mask, int size, float sigma = 0) {  // helpers.cpp:442-461
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

inline void invSqrt(float& a, float& b, float& c, float& l1, float& l2) {  // helpers.cpp:463-502
  double t, r;
  if (b != 0) {
    r = double(c - a) / (2 * b);
    if (r >= 0) t = 1.0 / (r + std::sqrt(1 + r * r));
    else t = -1.0 / (-r + std::sqrt(1 + r * r));
    r = 1.0 / std::sqrt(1 + t * t);
    t = t * r;
  } else { r = 1; t = 0; }
  double x, z, d;
  x = 1.0 / std::sqrt(r * r * a - 2 * r * t * b + t * t * c);
  z = 1.0 / std::sqrt(t * t * a + 2 * r * t * b + r * r * c);
  d = std::sqrt(x * z);
  x /= d; z /= d;
  if (x < z) { l1 = float(z); l2 = float(x); } else { l1 = float(x); l2 = float(z); }
  a = float(r * r * x + t * t * z);
  b = float(-r * t * x + t * r * z);
  c = float(t * t * x + r * r * z);
}

inline bool getEigenvalues(float a, float b, float c, float d, float& l1, float& l2) {  // helpers.cpp:504-515
  float trace = a + d;
  float delta1 = (trace * trace - 4 * (a * d - b * c));
  if (delta1 < 0) return false;
  float delta = std::sqrt(delta1);
  l1 = (trace + delta) / 2.0f;
  l2 = (trace - delta) / 2.0f;
  return true;
}

inline bool interpolateCheckBorders(int orig_img_w, int orig_img_h, float ofsx, float ofsy, float a11, float a12,
                                    float a21, float a22, int res_w, int res_h) {  // helpers.cpp:524-549
  const int width = orig_img_w - 2, height = orig_img_h - 2;
  const float halfWidth = std::ceil((float)res_w / 2.0);
  const float halfHeight = std::ceil((float)res_h / 2.0);
  float x[4] = {-halfWidth, -halfWidth, +halfWidth, +halfWidth};
  float y[4] = {-halfHeight, +halfHeight, -halfHeight, +halfHeight};
  for (int i = 0; i < 4; i++) {
    float imx = ofsx + x[i] * a11 + y[i] * a12;
    float imy = ofsy + x[i] * a21 + y[i] * a22;
    if (std::floor(imx) <= 0 || std::floor(imy) <= 0 || std::ceil(imx) >= width || std::ceil(imy) >= height) return true;
  }
  return false;
}

// res is (rh x rw), both odd in every caller.
inline bool interpolate(const float* im, int im_rows, int im_cols, float ofsx, float ofsy, float a11, float a12,
                        float a21, float a22, float* res, int rh, int rw) {  // helpers.cpp:551-626
  bool ret = false;
  const int width = im_cols - 1, height = im_rows - 1;
  const int halfWidth = rw >> 1, halfHeight = rh >> 1;
  float* out = res;
  float rx = ofsx - (float)halfHeight * a12;
  float ry = ofsy - (float)halfHeight * a22;
  bool touch = interpolateCheckBorders(im_cols, im_rows, ofsx, ofsy, a11, a12, a21, a22, rw, rh);
  if (!touch) {
    for (int j = -halfHeight; j <= halfHeight; ++j) {
      float WX = rx - (float)halfWidth * a11;
      float WY = ry - (float)halfWidth * a21;
      for (int i = -halfWidth; i <= halfWidth; ++i) {
        const int x = (int)(WX), y = (int)(WY);
        const float wx = WX - (float)x;
        const float* Row0 = im + (size_t)y * im_cols;
        const float* Row1 = Row0 + im_cols;
        const float I1 = wx * (Row0[x + 1] - Row0[x]) + Row0[x];
        *out++ = (WY - y) * (wx * (Row1[x + 1] - Row1[x]) + Row1[x] - I1) + I1;
        WX += a11; WY += a21;
      }
      rx += a12; ry += a22;
    }
  } else {
    for (int j = -halfHeight; j <= halfHeight; ++j) {
      float WX = rx - halfWidth * a11;
      float WY = ry - halfWidth * a21;
      for (int i = -halfWidth; i <= halfWidth; ++i) {
        const int x = (int)std::floor(WX), y = (int)std::floor(WY);
        if (WX >= 0 && WY >= 0 && x < width && y < height) {
          const float wx = WX - x;
          const float* Row0 = im + (size_t)y * im_cols;
          const float* Row1 = Row0 + im_cols;
          const float I1 = wx * (Row0[x + 1] - Row0[x]) + Row0[x];
          *out++ = (WY - y) * (wx * (Row1[x + 1] - Row1[x]) + Row1[x] - I1) + I1;
        } else { *out++ = 0; ret = true; }
        WX += a11; WY += a21;
      }
      rx += a12; ry += a22;
    }
  }
  return ret;
}

inline void photometricallyNormalize(float* image, const float* binaryMask, int width, int height) {  // helpers.cpp:666-715
  float sum = 0, gsum = 0;
  for (int j = 0; j < height; j++)
    for (int i = 0; i < width; i++)
      if (binaryMask[j * width + i] > 0) { sum += image[j * width + i]; gsum++; }
  sum = sum / gsum;
  float var = 0;
  for (int j = 0; j < height; j++)
    for (int i = 0; i < width; i++)
      if (binaryMask[j * width + i] > 0) var += (sum - image[j * width + i]) * (sum - image[j * width + i]);
  var = std::sqrt(var / gsum);
  if (var < 0.0001) return;
  float fac = 50.0f / var;
  for (int k = 0; k < width * height; k++) {
    image[k] = 128 + fac * (image[k] - sum);
    if (image[k] > 255) image[k] = 255;
    if (image[k] < 0) image[k] = 0;
  }
}

inline void computeGradient(const float* img, int width, int height, float* gx, float* gy) {  // helpers.cpp:779-797
  for (int r = 0; r < height; ++r)
    for (int c = 0; c < width; ++c) {
      float xgrad, ygrad;
      if (c == 0) xgrad = img[r * width + c + 1] - img[r * width + c];
      else if (c == width - 1) xgrad = img[r * width + c] - img[r * width + c - 1];
      else xgrad = img[r * width + c + 1] - img[r * width + c - 1];
      if (r == 0) ygrad = img[(r + 1) * width + c] - img[r * width + c];
      else if (r == height - 1) ygrad = img[r * width + c] - img[(r - 1) * width + c];
      else ygrad = img[(r + 1) * width + c] - img[(r - 1) * width + c];
      gx[r * width + c] = xgrad; gy[r * width + c] = ygrad;
    }
}

inline void rectifyTransformation(double& a11, double& a12, double& a21, double& a22) {  // synth-detection.cpp:46-55
  double a = a11, b = a12, c = a21, d = a22;
  double det = std::sqrt(std::fabs(a * d - b * c));
  double b2a2 = std::sqrt(b * b + a * a);
  a11 = b2a2 / det; a12 = 0; a21 = (d * b + c * a) / (b2a2 * det); a22 = det / b2a2;
}

// ------------------------------------------------------------------------------------------
// detectors/affinedetectors/pyramid.cpp  +  affine.cpp  +  scale-space-detector.*
// ------------------------------------------------------------------------------------------
inline Image gaussianBlur(const Image& in, float sigma) {  // helpers.cpp:717-724
  Image out(in.rows, in.cols);
  cvmath::gaussian_blur(in.px.data(), out.px.data(), in.rows, in.cols, sigma);
  return out;
}

inline Image hessianResponse(const Image& in, float norm) {  // pyramid.cpp:223-281
  Image out(in.rows, in.cols);  // reference leaves the 1-px frame uninitialised; we zero it
  const float norm2 = norm * norm;
  for (int r = 1; r < in.rows - 1; ++r) {
    const float* p0 = in.row(r - 1); const float* p1 = in.row(r); const float* p2 = in.row(r + 1);
    float* o = out.row(r);
    for (int c = 1; c < in.cols - 1; ++c) {
      const float v11 = p0[c - 1], v12 = p0[c], v13 = p0[c + 1];
      const float v21 = p1[c - 1], v22 = p1[c], v23 = p1[c + 1];
      const float v31 = p2[c - 1], v32 = p2[c], v33 = p2[c + 1];
      float Lxx = (v21 - 2 * v22 + v23);
      float Lyy = (v12 - 2 * v22 + v32);
      float Lxy = (v13 - v11 + v31 - v33) / 4.0f;
      o[c] = (Lxx * Lyy - Lxy * Lxy) * norm2;
    }
  }
  return out;
}

inline Image dogResponse(const Image& in, float norm) {  // pyramid.cpp:176-181: the level minus its own blur with sigma = norm (= curSigma^2)
  Image nb = gaussianBlur(in, norm), out(in.rows, in.cols);
  for (size_t i = 0; i < out.px.size(); i++) out.px[i] = in.px[i] - nb.px[i];
  return out;
}

// pyramid.cpp:283-305: Harris measure of the level's gradient, every Mat operation rounded to float on its own as cv::Mat arithmetic does
// (Lx.mul(Lx), the three blurs with sigma = sqrt(0.6 norm), scalar * Mat with the scalar widened to double, +, -)
inline Image harrisResponse(const Image& in, float norm) {
  const int rows = in.rows, cols = in.cols;
  const float sigmasq = 0.6 * norm;
  const float sigma = std::sqrt(sigmasq);
  Image Lx(rows, cols), Ly(rows, cols), xx(rows, cols), yy(rows, cols), xy(rows, cols), out(rows, cols);
  computeGradient(in.px.data(), cols, rows, Lx.px.data(), Ly.px.data());
  for (size_t i = 0; i < out.px.size(); i++) { xx.px[i] = Lx.px[i] * Lx.px[i]; yy.px[i] = Ly.px[i] * Ly.px[i]; xy.px[i] = Lx.px[i] * Ly.px[i]; }
  Image bxx = gaussianBlur(xx, sigma), byy = gaussianBlur(yy, sigma), bxy = gaussianBlur(xy, sigma);
  for (size_t i = 0; i < out.px.size(); i++) {
    const float dx2 = (float)(bxx.px[i] * (double)sigmasq), dy2 = (float)(byy.px[i] * (double)sigmasq), dxdy = (float)(bxy.px[i] * (double)sigmasq);
    const float sum = dx2 + dy2;
    const float a = dx2 * dy2, b = dxdy * dxdy, c = sum * sum;
    const float ab = a - b;
    out.px[i] = ab - (float)(c * 0.04);
  }
  return out;
}

struct Candidate { int r, c; };  // 3x3x3 extremum before localisation

struct HessianAffineDetector {
  HessParams par;
  std::vector<Key> keys;
  std::vector<float> smmMask;  // computeGaussMask(19x19), affine.h:104
  // statistics / intermediate dumps for kernel-level parity tests
  int extrema_points = 0, localized_points = 0;
  struct LevelDump { int octave, level, rows, cols; float pixelDistance, curSigma; };
  bool keep_levels = false;
  std::vector<Image> dump_blur, dump_resp;
  std::vector<LevelDump> dump_info;
  // (x, y, s, pixelDistance, type, response, octave, level) of every localized point, in detection order
  std::vector<std::vector<float>> localized;

  double edgeScoreThreshold;
  float finalThreshold, positiveThreshold, negativeThreshold;

  explicit HessianAffineDetector(const HessParams& p) : par(p), smmMask((size_t)p.smmWindowSize * p.smmWindowSize) {
    // pyramid.h:46-69.  edgeEigenValueRatio is a double in PyramidParams.
    double er = par.edgeEigenValueRatio;
    edgeScoreThreshold = (er + 1.0f) * (er + 1.0f) / er;
    finalThreshold = par.threshold;
    positiveThreshold = (float)(0.8 * finalThreshold);
    negativeThreshold = -positiveThreshold;
    if (par.detectorType == 0) finalThreshold = par.threshold * par.threshold;  // DET_HESSIAN only (pyramid.h:56-57)
    if (par.mode != 0) finalThreshold = positiveThreshold = negativeThreshold = 0.0f;
    computeGaussMask(smmMask.data(), par.smmWindowSize);
  }

  // affine.cpp:26-169
  bool findAffineShape(const Image& blur, float x, float y, float s, float pixelDistance, int type, float response) {
    float eigen_ratio_act = 0.0f, eigen_ratio_bef = 0.0f;
    float u11 = 1.0f, u12 = 0.0f, u21 = 0.0f, u22 = 1.0f, l1 = 1.0f, l2 = 1.0f;
    float lx = x / pixelDistance, ly = y / pixelDistance;
    float ratio = s / (par.initialSigma * pixelDistance);
    const int W = par.smmWindowSize, maskPixels = W * W;
    std::vector<float> img(maskPixels), fx(maskPixels), fy(maskPixels);
    if (!par.doBaumberg) { pushKey(x, y, s, u11, u12, u21, u22, type, response); return true; }
    for (int l = 0; l < par.maxIterations; l++) {
      float a = 0, b = 0, c = 0;
      interpolate(blur.px.data(), blur.rows, blur.cols, lx, ly, u11 * ratio, u12 * ratio, u21 * ratio, u22 * ratio,
                  img.data(), W, W);
      computeGradient(img.data(), W, W, fx.data(), fy.data());
      for (int i = 0; i < maskPixels; ++i) {
        const float v = smmMask[i], gxx = fx[i], gyy = fy[i];
        const float gxy = gxx * gyy;
        a += gxx * gxx * v;
        b += gxy * v;
        c += gyy * gyy * v;
      }
      a /= maskPixels; b /= maskPixels; c /= maskPixels;
      invSqrt(a, b, c, l1, l2);
      if ((a != a) || (b != b) || (c != c)) break;
      eigen_ratio_bef = eigen_ratio_act;
      eigen_ratio_act = 1.0 - l2 / l1;
      float u11t = u11, u12t = u12;
      u11 = a * u11t + b * u21;
      u12 = a * u12t + b * u22;
      u21 = b * u11t + c * u21;
      u22 = b * u12t + c * u22;
      if (!getEigenvalues(u11, u12, u21, u22, l1, l2)) break;
      if ((l1 / l2 > 6) || (l2 / l1 > 6)) break;
      if (eigen_ratio_act < par.convergenceThreshold && eigen_ratio_bef < par.convergenceThreshold) {
        pushKey(x, y, s, u11, u12, u21, u22, type, response);
        return true;
      }
    }
    return false;
  }
  void pushKey(float x, float y, float s, float a11, float a12, float a21, float a22, int type, float response) {
    Key k;  // scale-space-detector.hpp:70-88
    k.x = x; k.y = y; k.s = s; k.a11 = a11; k.a12 = a12; k.a21 = a21; k.a22 = a22; k.response = response; k.sub_type = type;
    keys.push_back(k);
  }

  // pyramid.cpp:308-430
  void localizeKeypoint(int r, int c, float curScale, float pixelDistance, const Image& low, const Image& cur,
                        const Image& high, const Image& blur, const Image& prevBlur, std::vector<unsigned char>& octaveMap,
                        int octave, int level) {
    const int cols = cur.cols, rows = cur.rows;
    float b[3] = {};
    float val = 0;
    int nr = r, nc = c;
    for (int iter = 0; iter < 5; iter++) {
      r = nr; c = nc;
      const float* cur0 = cur.row(r - 1); const float* cur1 = cur.row(r); const float* cur2 = cur.row(r + 1);
      const float* low0 = low.row(r - 1); const float* low1 = low.row(r); const float* low2 = low.row(r + 1);
      const float* high0 = high.row(r - 1); const float* high1 = high.row(r); const float* high2 = high.row(r + 1);
      float dxx = cur1[c - 1] - 2.0f * cur1[c] + cur1[c + 1];
      float dyy = cur0[c] - 2.0f * cur1[c] + cur2[c];
      float dss = low1[c] - 2.0f * cur1[c] + high1[c];
      float dxy = 0.25f * (cur2[c + 1] - cur2[c - 1] - cur0[c + 1] + cur0[c - 1]);
      if (0 == iter) {
        float edgeScore = (dxx + dyy) * (dxx + dyy) / (dxx * dyy - dxy * dxy);
        if (edgeScore >= edgeScoreThreshold || edgeScore < 0) return;
      }
      float dxs = 0.25f * (high1[c + 1] - high1[c - 1] - low1[c + 1] + low1[c - 1]);
      float dys = 0.25f * (high2[c] - high0[c] - low2[c] + low0[c]);
      float A[9] = {dxx, dxy, dxs, dxy, dyy, dys, dxs, dys, dss};
      float dx = 0.5f * (cur1[c + 1] - cur1[c - 1]);
      float dy = 0.5f * (cur2[c] - cur0[c]);
      float ds = 0.5f * (high1[c] - low1[c]);
      b[0] = -dx; b[1] = -dy; b[2] = -ds;
      solveLinear3x3(A, b);
      if (std::isnan(b[0]) || std::isnan(b[1]) || std::isnan(b[2])) return;
      val = cur1[c] + 0.5f * (dx * b[0] + dy * b[1] + ds * b[2]);
      if (b[0] > 0.6) { if (c < cols - 3) nc++; else return; }
      if (b[1] > 0.6) { if (r < rows - 3) nr++; else return; }
      if (b[0] < -0.6) { if (c > 3) nc--; else return; }
      if (b[1] < -0.6) { if (r > 3) nr--; else return; }
      if (nr == r && nc == c) break;
    }
    if (std::fabs(b[0]) > 1.5 || std::fabs(b[1]) > 1.5 || std::fabs(b[2]) > 1.5 || std::fabs(val) < finalThreshold ||
        octaveMap[(size_t)r * cols + c] > 0)
      return;
    octaveMap[(size_t)r * cols + c] = 1;
    float scale = curScale * std::pow(2.0f, b[2] / par.numberOfScales);
    int type;  // getPointType, pyramid.cpp:66-130
    if (par.detectorType == 1) type = val < 0 ? 11 : 10;   // DOG_BRIGHT : DOG_DARK (pyramid.h:36-37)
    else if (par.detectorType == 2) type = val < 0 ? 31 : 30;   // HARRIS_BRIGHT : HARRIS_DARK (pyramid.h:38-39)
    else if (val < 0) type = 2;
    else {
      const float* p = blur.row(r) + c;
      float Lxx = (p[-1] - 2 * p[0] + p[1]);
      type = (Lxx < 0) ? 0 : 1;
    }
    localized_points++;
    float kx = pixelDistance * (c + b[0]), ky = pixelDistance * (r + b[1]), ks = pixelDistance * scale;
    localized.push_back({kx, ky, ks, pixelDistance, (float)type, val, (float)octave, (float)level});
    findAffineShape(prevBlur, kx, ky, ks, pixelDistance, type, val);  // scale-space-detector.hpp:48-56
  }

  static bool isMax(float val, const Image& pix, int row, int col) {  // pyramid.cpp:42-52
    for (int r = row - 1; r <= row + 1; r++) { const float* p = pix.row(r); for (int c = col - 1; c <= col + 1; c++) if (p[c] > val) return false; }
    return true;
  }
  static bool isMin(float val, const Image& pix, int row, int col) {  // pyramid.cpp:54-64
    for (int r = row - 1; r <= row + 1; r++) { const float* p = pix.row(r); for (int c = col - 1; c <= col + 1; c++) if (p[c] < val) return false; }
    return true;
  }

  // pyramid.cpp:455-538 (detectOctaveKeypoints) with findLevelKeypoints (:432-452) inlined
  void detectOctave(const Image& firstLevel, float pixelDistance, Image& nextOctaveFirstLevel, int octave) {
    std::vector<unsigned char> octaveMap((size_t)firstLevel.rows * firstLevel.cols, 0);
    float sigmaStep = std::pow(2.0f, 1.0f / (float)par.numberOfScales);
    float curSigma = par.initialSigma;
    int numLevels = 1;
    Image blur = firstLevel, prevBlur, low, cur, high;
    auto response = [&](const Image& im, float norm) {   // pyramid.cpp:132-175
      return par.detectorType == 1 ? dogResponse(im, norm) : par.detectorType == 2 ? harrisResponse(im, norm) : hessianResponse(im, norm); };
    cur = response(blur, curSigma * curSigma);
    if (keep_levels) { dump_blur.push_back(blur); dump_resp.push_back(cur); dump_info.push_back({octave, 0, blur.rows, blur.cols, pixelDistance, curSigma}); }
    for (int i = 1; i < par.numberOfScales + 2; i++) {
      float sigma = curSigma * std::sqrt(sigmaStep * sigmaStep - 1.0f);
      Image nextBlur = gaussianBlur(blur, sigma);
      sigma = curSigma * sigmaStep;
      high = response(nextBlur, sigma * sigma);
      if (keep_levels) { dump_blur.push_back(nextBlur); dump_resp.push_back(high); dump_info.push_back({octave, i, blur.rows, blur.cols, pixelDistance, sigma}); }
      numLevels++;
      if (numLevels == 3) {
        const int rows = cur.rows, cols = cur.cols;
        for (int r = par.border; r < (rows - par.border); r++) {
          const float* curPtr = cur.row(r);
          for (int c = par.border; c < (cols - par.border); c++) {
            const float val = curPtr[c];
            if ((val > positiveThreshold && (isMax(val, cur, r, c) && isMax(val, low, r, c) && isMax(val, high, r, c))) ||
                (val < negativeThreshold && (isMin(val, cur, r, c) && isMin(val, low, r, c) && isMin(val, high, r, c)))) {
              extrema_points++;
              localizeKeypoint(r, c, curSigma, pixelDistance, low, cur, high, blur, prevBlur, octaveMap, octave, i - 1);
            }
          }
        }
        numLevels--;
      }
      if (i == par.numberOfScales) {
        int orows, ocols;
        cvmath::half_size(nextBlur.rows, nextBlur.cols, &orows, &ocols);
        nextOctaveFirstLevel = Image(orows, ocols);
        cvmath::resize_half(nextBlur.px.data(), nextBlur.rows, nextBlur.cols, nextOctaveFirstLevel.px.data());
      }
      prevBlur = blur; blur = nextBlur; low = cur; cur = high;
      curSigma *= sigmaStep;
    }
  }

  // pyramid.cpp:540-573 (upscaleInputImage == 0 only)
  void detectPyramidKeypoints(const Image& image) {
    float curSigma = 0.5f, pixelDistance = 1.0f;
    Image firstLevel = image;
    if (par.initialSigma > curSigma) {
      float sigma = std::sqrt(par.initialSigma * par.initialSigma - curSigma * curSigma);
      cvmath::gaussian_blur(firstLevel.px.data(), firstLevel.px.data(), firstLevel.rows, firstLevel.cols, sigma);
    }
    int minSize = 2 * par.border + 2, octave = 0;
    while (firstLevel.rows > minSize && firstLevel.cols > minSize) {
      Image next;
      detectOctave(firstLevel, pixelDistance, next, octave++);
      pixelDistance *= 2.0;
      firstLevel = next;
    }
  }

  // scale-space-detector.hpp:127-198 (prepareKeysForExport)
  void prepareKeysForExport() {
    if (keys.empty() || par.mode == 0) return;
    // sortKeys (:115-118): the reference's own std::sort (unstable) with responseCompareInvOrder -- the same library routine on the
    // same input order is the only way to reproduce the order of equal responses
    std::sort(keys.begin(), keys.end(), [](const Key& a, const Key& b) { return std::fabs(a.response) > std::fabs(b.response); });
    int regNumber = (int)keys.size();
    auto lower = [&](double thr) {
      int lo = 0;  // first index whose |response| is NOT > thr  (std::lower_bound with responseCompareInvOrder)
      while (lo < regNumber && std::fabs(keys[lo].response) > std::fabs(thr)) lo++;
      return lo;
    };
    switch (par.mode) {
      case 1: keys.resize(lower(std::fabs(keys[0].response) * par.rel_threshold)); break;
      case 2: {
        int n = par.reg_number;
        if (par.doBaumberg) n = (int)std::floor(3.0 * (double)n);
        if (n < regNumber && n >= 0) keys.resize(n);
        break;
      }
      case 3: keys.resize((int)std::floor(par.rel_reg_number * (double)keys.size())); break;
      case 4: {
        int fix = lower(par.threshold);
        if (fix < par.reg_number) keys.resize(std::min(par.reg_number, regNumber));
        else keys.resize(std::min(fix, regNumber));
        break;
      }
    }
    if (par.mode == 2 && (int)keys.size() > par.reg_number) keys.resize(par.reg_number);
  }
};

// scale-space-detector.cpp:43-85 (tilt/zoom rescale of reg_number included)
inline std::vector<Key> detectAffineKeypoints(const Image& img, HessParams p, double tilt, double zoom,
                                              HessianAffineDetector** keep = nullptr) {
  if ((tilt > 2.0) || (zoom < 0.5)) p.reg_number = (int)std::floor(zoom * (double)p.reg_number / tilt);
  HessianAffineDetector* d = new HessianAffineDetector(p);
  if (keep) d->keep_levels = true;
  d->detectPyramidKeypoints(img);
  d->prepareKeysForExport();
  std::vector<Key> out = d->keys;
  if (keep) *keep = d; else delete d;
  return out;
}

// synth-detection.hpp:93-126 (DetectAffineRegions: scale by sqrt|det|, rectify)
inline void toRegions(std::vector<Key>& keys) {
  for (Key& k : keys) {
    k.s = k.s * std::sqrt(std::fabs(k.a11 * k.a22 - k.a12 * k.a21));
    rectifyTransformation(k.a11, k.a12, k.a21, k.a22);
  }
}

// ------------------------------------------------------------------------------------------
// synth-detection.cpp:236-430  GenerateSynthImageCorr (gray input; the (B+G+R)/3 conversion is the caller's)
// ------------------------------------------------------------------------------------------
struct SynthView { Image pixels; double H[9]; double tilt, zoom, rotation; bool identity; };
inline void generateSynthImage(const Image& in, double tilt, double phi, double zoom, double InitSigma, int doBlur, SynthView& out) {
  bool vertical_tilt = false;
  if (tilt < 0) { tilt = -tilt; vertical_tilt = true; }
  const int zoomed = std::fabs(zoom - 1.0f) >= 0.05 ? 1 : 0;
  const int w = in.cols, h = in.rows;
  const int wS1 = (int)(w * zoom), hS1 = (int)(h * zoom);
  for (int i = 0; i < 9; i++) out.H[i] = (i % 4 == 0) ? 1.0 : 0.0;
  out.identity = false;
  if ((std::fabs(tilt - 1.) <= 0.1) && (std::fabs(phi) <= 0.2) && (std::fabs(zoom - 1.) <= 0.1)) {   // :278 (abs is std::abs(double): `using namespace std`, :30)
    out.rotation = 0; out.tilt = 1; out.zoom = 1; out.pixels = in; out.identity = true;
    return;
  }
  double d, d2, w_new, h_new, kV = 1., kH = 1.;
  if (zoomed) { kV = (double)w / (double)wS1; kH = (double)h / (double)hS1; }
  const double c = std::cos(phi), s = std::sin(phi);
  const bool q1 = (phi >= 0) && (phi < M_PI / 2);
  const double sx = vertical_tilt ? kH : tilt * kH, sy = vertical_tilt ? tilt * kV : kV;   // :301-342
  if (q1) {
    w_new = std::floor((0.5 + c * w + s * h) / sx); h_new = std::floor((0.5 + s * w + c * h) / sy);
    out.H[0] = c / sx; out.H[1] = s / sx; out.H[2] = 0;
    out.H[3] = -s / sy; out.H[4] = c / sy; out.H[5] = std::floor(0.5 + s * w / sy);
  } else {
    w_new = std::floor((0.5 - c * w + s * h) / sx); h_new = std::floor((0.5 + s * w - c * h) / sy);
    d = -std::floor(c * w / sx); d2 = std::floor(0.5 + (s * w - c * h) / sy);
    out.H[0] = c / sx; out.H[1] = s / sx; out.H[2] = d;
    out.H[3] = -s / sy; out.H[4] = c / sy; out.H[5] = d2;
  }
  out.H[6] = 0; out.H[7] = 0; out.H[8] = 1;
  out.rotation = phi * 180 / M_PI; out.tilt = tilt; out.zoom = zoom;
  const double sigma_aa_2 = zoomed ? InitSigma / (4.0 * zoom) : InitSigma / 2.0;
  const double sigma_aa = InitSigma * tilt / (2.0 * zoom);
  const double sigma_x = vertical_tilt ? sigma_aa_2 : sigma_aa, sigma_y = vertical_tilt ? sigma_aa : sigma_aa_2;
  int w_new_rot, h_new_rot; double warpRot[6];
  if (q1) {
    w_new_rot = (int)std::floor(0.5 + c * w + s * h); h_new_rot = (int)std::floor(0.5 + s * w + c * h);
    warpRot[0] = c; warpRot[1] = s; warpRot[2] = 0; warpRot[3] = -s; warpRot[4] = c; warpRot[5] = std::floor(0.5 + s * w);
  } else {
    w_new_rot = (int)std::floor(0.5 - c * w + s * h); h_new_rot = (int)std::floor(0.5 + s * w - c * h);
    d = -std::floor(c * w); d2 = std::floor(0.5 + (s * w - c * h));
    warpRot[0] = c; warpRot[1] = s; warpRot[2] = d; warpRot[3] = -s; warpRot[4] = c; warpRot[5] = d2;
  }
  Image rot(h_new_rot, w_new_rot);
  cvmath::warp_affine_linear(in.px.data(), h, w, warpRot, rot.px.data(), h_new_rot, w_new_rot, 128.f);   // :385
  if (doBlur) {                                                                                               // :398-412
    int kx = (int)std::floor(2.0 * 3.0 * sigma_x + 1.0); if (kx % 2 == 0) kx++; if (kx < 3) kx = 3;
    int ky = (int)std::floor(2.0 * 3.0 * sigma_y + 1.0); if (ky % 2 == 0) ky++; if (ky < 3) ky = 3;
    cvmath::sep_filter_reflect101(rot.px.data(), rot.px.data(), rot.rows, rot.cols, cvmath::gauss_kernel(kx, sigma_x), cvmath::gauss_kernel(ky, sigma_y));
  }
  const double wtz[6] = {1.0 / sx, 0, 0, 0, 1.0 / sy, 0};                                                    // :415-427
  out.pixels = Image((int)h_new, (int)w_new);
  cvmath::warp_affine_linear(rot.px.data(), rot.rows, rot.cols, wtz, out.pixels.px.data(), (int)h_new, (int)w_new, 128.f);
}

// ------------------------------------------------------------------------------------------
// synth-detection.cpp: orientation, reprojection;  synth-detection.hpp: DescribeRegions
// ------------------------------------------------------------------------------------------
const double k_sigma_sd = 2 * 3.0 * 1.7320508075688772;  // synth-detection.cpp:28  (2*3*sqrt(3))

// synth-detection.cpp:746-839.  img is patch x patch.
inline void estimateDominantAngles(const float* img, int pS, const float* orimask, std::vector<float>& angles1, double max_th,
                                   int maxAngles, std::vector<float>& gmag, std::vector<float>& gori, bool doHalfSIFT = false) {
  angles1.clear();
  if (maxAngles == 0) return;
  const int bins = 36;
  float hist[bins + 1];
  std::vector<float> peak_values;
  for (int i = 0; i < bins; i++) hist[i] = 0.0f;
  float hist36 = 0.f;  // bin 36 (ori == +pi) lands in an uninitialised slot that is never read
  // computeGradientMagnitudeAndOrientation, helpers.cpp:840-863 (border of gmag/gori stays 0)
  for (int r = 1; r < pS - 1; ++r)
    for (int c = 1; c < pS - 1; ++c) {
      float xgrad = img[r * pS + c + 1] - img[r * pS + c - 1];
      float ygrad = img[(r + 1) * pS + c] - img[(r - 1) * pS + c];
      gmag[r * pS + c] = std::sqrt(xgrad * xgrad + ygrad * ygrad);
      gori[r * pS + c] = atan2LUTff(ygrad, xgrad);
    }
  const int maskPixels = pS * (pS - 2);
  const float* maskptr = orimask + pS; const float* pmag = gmag.data() + pS; const float* pori = gori.data() + pS;
  for (int i = 0; i < maskPixels; ++i) {
    if (maskptr[i] > 0 && pmag[i] > 1.0) {
      int bin = (int)(bins * (pori[i] / float(M_PI) + 1.0f) / 2.0f);
      if (bin == bins) hist36 += pmag[i] * maskptr[i]; else hist[bin] += pmag[i] * maskptr[i];
    }
  }
  (void)hist36;
  for (int it = 0; it < 6; it++) {  // smoothCircularBuffer<36>, synth-detection.cpp:721-732
    float first = hist[0], prev = hist[bins - 1];
    for (int i = 0; i < bins - 1; i++) { float cur = hist[i]; hist[i] = prev + cur + hist[i + 1]; prev = cur; }
    hist[bins - 1] = prev + hist[bins - 1] + first;
  }
  float thresh = 0.0;
  for (int i = 0; i < bins; i++) if (hist[i] > thresh) thresh = hist[i];
  thresh *= max_th;
  if (doHalfSIFT) {  // synth-detection.cpp:801-808: orientations modulo pi
    const int halfbins = bins / 2;
    for (int i = 0; i < halfbins; i++) { hist[i] += hist[i + halfbins]; hist[i + halfbins] = 0; }
  }
  auto addPeak = [&](int a, int b, int c) {  // synth-detection.cpp:734-744
    if (hist[b] >= thresh && hist[b] > hist[a] && hist[b] > hist[c]) {
      float pp = (hist[a] - hist[c]) / (hist[a] - 2.0f * hist[b] + hist[c]) / 2.0f;
      angles1.push_back(2.0f * float(M_PI) * (b + 0.5f + pp) / bins - float(M_PI));
      peak_values.push_back(hist[b]);
    }
  };
  addPeak(bins - 1, 0, 1);
  for (int i = 1; i < bins - 1; i++) addPeak(i - 1, i, i + 1);
  addPeak(bins - 2, bins - 1, 0);
  if (maxAngles == -1) maxAngles = 100000000;
  maxAngles = std::min(maxAngles, (int)peak_values.size());
  if (maxAngles > 0) {
    std::vector<float> ang_tmp;
    for (int ang = 0; ang < maxAngles; ang++) { if (peak_values[ang] >= thresh) ang_tmp.push_back(angles1[ang]); else break; }
    angles1 = ang_tmp;
  } else angles1.clear();
}

// synth-detection.cpp:841-919 (addUpRight = false)
inline std::vector<Key> detectOrientation(const std::vector<Key>& in, const Image& img, double mrSize, int patchSize,
                                          int maxAngNum, double th, bool doHalfSIFT = false) {
  std::vector<Key> out;
  double mrScale = (double)mrSize;
  int patchImageSize = 2 * int(mrScale) + 1;
  double imageToPatchScale = double(patchImageSize) / (double)patchSize;
  std::vector<float> patch((size_t)patchSize * patchSize), orimask((size_t)patchSize * patchSize);
  std::vector<float> gmag((size_t)patchSize * patchSize, 0.f), gori((size_t)patchSize * patchSize, 0.f);
  computeCircularGaussMask(orimask.data(), patchSize, patchSize / 3.0f);
  std::vector<float> angles1;
  for (size_t i = 0; i < in.size(); i++) {
    const Key& k = in[i];
    float curr_sc = imageToPatchScale * k.s;
    if (interpolateCheckBorders(img.cols, img.rows, (float)k.x, (float)k.y, (float)k.a11, (float)k.a12, (float)k.a21,
                                (float)k.a22, (int)(k_sigma_sd * k.s), (int)(k_sigma_sd * k.s)))
      continue;
    if (maxAngNum > 0) {
      interpolate(img.px.data(), img.rows, img.cols, (float)k.x, (float)k.y, (float)k.a11 * curr_sc, (float)k.a12 * curr_sc,
                  (float)k.a21 * curr_sc, (float)k.a22 * curr_sc, patch.data(), patchSize, patchSize);
      estimateDominantAngles(patch.data(), patchSize, orimask.data(), angles1, th, maxAngNum, gmag, gori, doHalfSIFT);
      for (size_t j = 0; j < angles1.size(); j++) {
        double ci = std::cos(-angles1[j]), si = std::sin(-angles1[j]);
        Key t = k;
        t.a11 = k.a11 * ci - k.a12 * si;
        t.a12 = k.a11 * si + k.a12 * ci;
        t.a21 = k.a21 * ci - k.a22 * si;
        t.a22 = k.a21 * si + k.a22 * ci;
        out.push_back(t);
      }
    }
  }
  return out;
}

inline bool HIsEye(const double* H) {  // synth-detection.cpp:56-61
  return (std::fabs(H[0] - 1.0) + std::fabs(H[1]) + std::fabs(H[2]) + std::fabs(H[3]) + std::fabs(H[4] - 1.0) + std::fabs(H[5]) +
              std::fabs(H[6]) + std::fabs(H[7]) + std::fabs(H[8] - 1.0) < 0.01);
}

// synth-detection.cpp:541-616 (which=0, halfsize k_sigma*s) and :63-102 (which=1, halfsize mrSize*s).
// det/reproj are filtered in lockstep.
inline void reprojectRegions(std::vector<Key>& det, std::vector<Key>& reproj, const double* H, int orig_w, int orig_h,
                             int which, double mrSize) {
  double Hinv[9];
  cvmath::invert3x3(H, Hinv);
  reproj = det;
  if (!HIsEye(H)) {
    for (size_t i = 0; i < det.size(); i++) {  // ReprojectByH, synth-detection.cpp:490-498
      const Key& in = det[i]; Key& o = reproj[i];
      o.x = (Hinv[0] * in.x + Hinv[1] * in.y + Hinv[2]);
      o.y = (Hinv[3] * in.x + Hinv[4] * in.y + Hinv[5]);
      o.a11 = (Hinv[0] * in.a11 + Hinv[1] * in.a21);
      o.a12 = (Hinv[0] * in.a12 + Hinv[1] * in.a22);
      o.a21 = (Hinv[3] * in.a11 + Hinv[4] * in.a21);
      o.a22 = (Hinv[3] * in.a12 + Hinv[4] * in.a22);
    }
  }
  std::vector<Key> d2, r2;
  const double half = which == 0 ? k_sigma_sd : mrSize;
  for (size_t i = 0; i < det.size(); i++) {
    const Key& p = reproj[i];
    if ((p.x < orig_w) && (p.y < orig_h) && (p.x > 0) && (p.y > 0)) {
      if (!interpolateCheckBorders(orig_w, orig_h, (float)p.x, (float)p.y, (float)p.a11, (float)p.a12, (float)p.a21, (float)p.a22,
                                   (int)(half * p.s), (int)(half * p.s))) {
        d2.push_back(det[i]); r2.push_back(reproj[i]);
      }
    }
  }
  det = d2; reproj = r2;
}

// matching/siftdesc.{h,cpp}: 4x4x8 SIFT / RootSIFT on a 41x41 patch
struct SIFTDescriptor {
  int patchSize = 41, spatialBins = 4, orientationBins = 8;
  double maxBinValue = 0.2f;  // siftdesc.h:56 (a float literal stored in a double)
  bool useRootSIFT = false;
  bool doHalfSIFT = false;   // with useRootSIFT: HalfRootSIFT (64 entries; rows stay 128 wide, the upper 64 are 0)
  bool doNorm = true;        // siftdesc.h:47; false: the raw votes come back (DSPSIFT, imagerepresentation.cpp:1549-1551)
  std::vector<float> mask, grad, ori;
  std::vector<int> bin0, bin1;
  std::vector<double> w0, w1, vec;
  SIFTDescriptor(int ps, bool root) : patchSize(ps), useRootSIFT(root), mask((size_t)ps * ps), grad((size_t)ps * ps), ori((size_t)ps * ps),
                                      bin0(ps), bin1(ps), w0(ps), w1(ps), vec(128) {
    computeCircularGaussMask(mask.data(), ps);
    int halfSize = ps >> 1;  // precomputeBinsAndWeights, siftdesc.cpp:22-71
    float step = float(spatialBins + 1) / (2 * halfSize);
    for (int i = 0; i < ps; i++) {
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
  void samplePatch() {  // siftdesc.cpp:73-131
    const double M_PI_DOUBLED = 6.28318530718;
    for (int r = 0; r < patchSize; ++r) {
      const int br0 = spatialBins * bin0[r]; const float wr0 = w0[r];
      const int br1 = spatialBins * bin1[r]; const float wr1 = w1[r];
      for (int c = 0; c < patchSize; ++c) {
        float val = 0.0f * 1.0 + (1.0 - 0.0f) * mask[r * patchSize + c] * grad[r * patchSize + c];
        const int bc0 = bin0[c]; const float wc0 = w0[c] * val;
        const int bc1 = bin1[c]; const float wc1 = w1[c] * val;
        const float o = float(orientationBins) * (ori[r * patchSize + c] + M_PI_DOUBLED) / M_PI_DOUBLED;
        int bo0 = (int)o;
        const float wo1 = o - bo0;
        bo0 %= orientationBins;
        int bo1 = (bo0 + 1) % orientationBins;
        const float wo0 = 1.0f - wo1;
        val = wr0 * wc0; if (val > 0) { vec[br0 + bc0 + bo0] += val * wo0; vec[br0 + bc0 + bo1] += val * wo1; }
        val = wr0 * wc1; if (val > 0) { vec[br0 + bc1 + bo0] += val * wo0; vec[br0 + bc1 + bo1] += val * wo1; }
        val = wr1 * wc0; if (val > 0) { vec[br1 + bc0 + bo0] += val * wo0; vec[br1 + bc0 + bo1] += val * wo1; }
        val = wr1 * wc1; if (val > 0) { vec[br1 + bc1 + bo0] += val * wo0; vec[br1 + bc1 + bo1] += val * wo1; }
      }
    }
  }
  static double normalize(std::vector<double>& v) {  // siftdesc.cpp:133-159
    double len = 0.0;
    for (size_t i = 0; i < v.size(); i += 4) {
      const double sq0 = v[i] * v[i], sq1 = v[i + 1] * v[i + 1], sq2 = v[i + 2] * v[i + 2], sq3 = v[i + 3] * v[i + 3];
      len += sq0 + sq1 + sq2 + sq3;
    }
    len = std::sqrt(len);
    const double fac = 1.0 / len;
    for (size_t i = 0; i < v.size(); i++) v[i] *= fac;
    return len;
  }
  void finish(std::vector<double>& vec) {  // SIFTnorm / RootSIFTnorm (double overloads), siftdesc.cpp:199-222, 247-262
    normalize(vec);
    bool changed = false;
    for (size_t i = 0; i < vec.size(); i++) if (vec[i] > maxBinValue) { vec[i] = maxBinValue; changed = true; }
    if (changed) normalize(vec);
    if (useRootSIFT) {
      double sum = 0.;
      for (size_t i = 0; i < vec.size(); i++) sum += std::fabs(vec[i]);
      for (size_t i = 0; i < vec.size(); i++) vec[i] = std::sqrt(vec[i] / sum);
      for (size_t i = 0; i < vec.size(); i++) { int b = std::max(0, std::min((int)(512.0 * vec[i] + 0.5), 255)); vec[i] = double(b); }
    } else {
      for (size_t i = 0; i < vec.size(); i++) { int b = std::max(0, std::min((int)(512.0f * vec[i] + 0.5), 255)); vec[i] = double(b); }
    }
  }
  void operator()(const float* patch, float* desc) {  // siftdesc.cpp:290-381, 401-442
    const int width = patchSize, height = patchSize;
    for (int r = 0; r < height; ++r)
      for (int c = 0; c < width; ++c) {
        float xgrad, ygrad;
        if (c == 0) xgrad = patch[r * width + c + 1] - patch[r * width + c];
        else if (c == width - 1) xgrad = patch[r * width + c] - patch[r * width + c - 1];
        else xgrad = patch[r * width + c + 1] - patch[r * width + c - 1];
        if (r == 0) ygrad = patch[(r + 1) * width + c] - patch[r * width + c];
        else if (r == height - 1) ygrad = patch[r * width + c] - patch[(r - 1) * width + c];
        else ygrad = patch[(r + 1) * width + c] - patch[(r - 1) * width + c];
        grad[r * width + c] = std::sqrt(xgrad * xgrad + ygrad * ygrad);
        ori[r * width + c] = atan2LUTff(ygrad, xgrad);
      }
    for (size_t i = 0; i < vec.size(); i++) vec[i] = 0;
    samplePatch();
    if (doHalfSIFT) {  // siftdesc.cpp:408-431: the un-normalised votes of opposite orientation bins are summed, then RootSIFTnorm on 64 entries
      const int spBins = spatialBins * spatialBins, oriHalf = orientationBins / 2;
      std::vector<double> half_vec((size_t)spBins * oriHalf);
      int b = 0;
      for (int i = 0; i < spBins; i++)
        for (int j = 0; j < oriHalf; j++) half_vec[b++] = vec[i * orientationBins + j] + vec[i * orientationBins + j + oriHalf];
      finish(half_vec);
      for (size_t i = 0; i < vec.size(); i++) desc[i] = i < half_vec.size() ? (float)half_vec[i] : 0.0f;
      return;
    }
    if (doNorm) finish(vec);
    for (size_t i = 0; i < vec.size(); i++) desc[i] = (float)vec[i];
  }
  // SIFTnorm(std::vector<float>&) with the float normalize (siftdesc.cpp:160-184, 263-278): what DSPSIFT applies to the summed votes
  void finishFloat(float* v) const {
    auto normalizef = [](float* x) {
      float len = 0.0f;
      for (int i = 0; i < 128; i += 4) {
        const float sq0 = x[i] * x[i], sq1 = x[i + 1] * x[i + 1], sq2 = x[i + 2] * x[i + 2], sq3 = x[i + 3] * x[i + 3];
        len += sq0 + sq1 + sq2 + sq3;
      }
      len = (float)std::sqrt((double)len);
      const double fac = 1.0 / len;
      for (int i = 0; i < 128; i++) x[i] *= (float)fac;
    };
    normalizef(v);
    bool changed = false;
    for (int i = 0; i < 128; i++) if (v[i] > maxBinValue) { v[i] = (float)maxBinValue; changed = true; }
    if (changed) normalizef(v);
    for (int i = 0; i < 128; i++) { int b = std::max(0, std::min((int)(512.0f * v[i] + 0.5), 255)); v[i] = float(b); }
  }
};

// synth-detection.hpp:169-255 (DescribeRegions); patch_out (n x ps x ps) optional for kernel-level parity
inline void describeRegions(const std::vector<Key>& kps, const Image& img, SIFTDescriptor& D, double mrSize, int patchSize,
                            bool fast_extraction, bool photoNorm, float* desc_out, float* patch_out = nullptr) {
  std::vector<float> patch((size_t)patchSize * patchSize), mask((size_t)patchSize * patchSize), workspace;
  computeCircularGaussMask(mask.data(), patchSize);
  for (size_t i = 0; i < kps.size(); i++) {
    const Key& k = kps[i];
    if (!fast_extraction) {
      float mrScale = std::ceil(k.s * mrSize);
      int patchImageSize = 2 * int(mrScale) + 1;
      float imageToPatchScale = float(patchImageSize) / float(patchSize);
      if (imageToPatchScale > 0.4) {
        patchImageSize += 2;
        workspace.resize((size_t)patchImageSize * patchImageSize);
        interpolate(img.px.data(), img.rows, img.cols, (float)k.x, (float)k.y, (float)k.a11, (float)k.a12, (float)k.a21,
                    (float)k.a22, workspace.data(), patchImageSize, patchImageSize);
        cvmath::gaussian_blur(workspace.data(), workspace.data(), patchImageSize, patchImageSize, 1.5f * imageToPatchScale);
        interpolate(workspace.data(), patchImageSize, patchImageSize, (float)(patchImageSize >> 1), (float)(patchImageSize >> 1),
                    imageToPatchScale, 0, 0, imageToPatchScale, patch.data(), patchSize, patchSize);
      } else {
        interpolate(img.px.data(), img.rows, img.cols, (float)k.x, (float)k.y, (float)k.a11 * imageToPatchScale,
                    (float)k.a12 * imageToPatchScale, (float)k.a21 * imageToPatchScale, (float)k.a22 * imageToPatchScale,
                    patch.data(), patchSize, patchSize);
      }
    } else {
      double mrScale = (double)mrSize * k.s;
      int patchImageSize = 2 * int(mrScale) + 1;
      double imageToPatchScale = double(patchImageSize) / (double)patchSize;
      float curr_sc = imageToPatchScale;
      interpolate(img.px.data(), img.rows, img.cols, (float)k.x, (float)k.y, (float)k.a11 * curr_sc, (float)k.a12 * curr_sc,
                  (float)k.a21 * curr_sc, (float)k.a22 * curr_sc, patch.data(), patchSize, patchSize);
    }
    if (photoNorm) photometricallyNormalize(patch.data(), mask.data(), patchSize, patchSize);
    if (patch_out) std::memcpy(patch_out + i * patch.size(), patch.data(), sizeof(float) * patch.size());
    D(patch.data(), desc_out + i * 128);
  }
}

// imagerepresentation.cpp:1547-1598 ("DSPSIFT"): the un-normalised SIFT votes described at numScales + 1 measurement-region sizes
// mrSize * (start + i (end - start) / numScales), summed in float, then SIFTnorm on the float vector
inline void describeRegionsDSP(const std::vector<Key>& kps, const Image& img, double mrSize, int patchSize, bool fast_extraction, bool photoNorm,
                               int numScales, double startCoef, double endCoef, float* desc_out) {
  SIFTDescriptor D(patchSize, false);
  D.doNorm = false;
  std::vector<float> tmp(kps.size() * 128);
  for (int dsp_idx = 0; dsp_idx < numScales + 1; dsp_idx++) {
    const double curr_mrSize = mrSize * (startCoef + dsp_idx * (endCoef - startCoef) / numScales);
    describeRegions(kps, img, D, curr_mrSize, patchSize, fast_extraction, photoNorm, tmp.data());
    for (size_t i = 0; i < tmp.size(); i++) desc_out[i] = dsp_idx == 0 ? tmp[i] : desc_out[i] + tmp[i];
  }
  D.doNorm = true;
  for (size_t k = 0; k < kps.size(); k++) D.finishFloat(desc_out + k * 128);
}

// ------------------------------------------------------------------------------------------
// matching/matching.cpp:357-461  MatchFlannFGINN with vector_matcher = linear (exact kNN).
// FLANN's order among equal distances is unpinned; we define ties = lower train index first.
// desc are integer-valued (0..255), so squared distances are exact in float.
// out rows: (query, idx0, idxJ, idx1, d0, dJ, d1)
// ------------------------------------------------------------------------------------------
struct Tentative { int q, i0, iJ, i1; float d0, dJ, d1; };

inline std::vector<Tentative> matchFGINN(const float* q, int nq, const float* t, int nt, const double* txy /* nt x 2 */,
                                         double matchRatio, double contradDist, int nn = 50) {
  std::vector<Tentative> out;
  if (nq == 0 || nt == 0) return out;
  const double sqminratio = matchRatio * matchRatio, contrDistSq = contradDist * contradDist;
  const int k = std::min(nn, nt);
  // queries are independent: all host threads (OpenMP), results put back in query order -- the same list as the serial loop.
  // Descriptor entries are integers 0..255 held in floats (siftdesc.cpp:218-274), so the float sum of squared differences the
  // reference forms (cvflann::L2<float>) is an exact integer < 2^24 in any summation order: it is evaluated here in 32-bit integer
  // arithmetic on u8 copies (vectorisable) and converted back to float -- bit-identical, several times faster than a serial float sum.
  bool integral = true;
  for (size_t i = 0; i < (size_t)nq * 128 && integral; i++) integral = q[i] >= 0 && q[i] <= 255 && q[i] == (float)(int)q[i];
  for (size_t i = 0; i < (size_t)nt * 128 && integral; i++) integral = t[i] >= 0 && t[i] <= 255 && t[i] == (float)(int)t[i];
  std::vector<unsigned char> q8, t8;
  if (integral) {
    q8.resize((size_t)nq * 128); t8.resize((size_t)nt * 128);
    for (size_t i = 0; i < q8.size(); i++) q8[i] = (unsigned char)q[i];
    for (size_t i = 0; i < t8.size(); i++) t8[i] = (unsigned char)t[i];
  }
  std::vector<Tentative> per_q(nq);
  std::vector<char> has(nq, 0);
#pragma omp parallel
  {
    std::vector<std::pair<float, int>> d(nt);
#pragma omp for schedule(dynamic, 16)
    for (int i = 0; i < nq; i++) {
      if (integral) {
        const unsigned char* a = q8.data() + (size_t)i * 128;
        for (int j = 0; j < nt; j++) {
          const unsigned char* b = t8.data() + (size_t)j * 128;
          int s = 0;
          for (int e = 0; e < 128; e++) { const int df = (int)a[e] - (int)b[e]; s += df * df; }
          d[j] = std::make_pair((float)s, j);
        }
      } else {
        const float* a = q + (size_t)i * 128;
        for (int j = 0; j < nt; j++) {
          const float* b = t + (size_t)j * 128;
          float s = 0;
          for (int e = 0; e < 128; e++) { float df = a[e] - b[e]; s += df * df; }
          d[j] = std::make_pair(s, j);
        }
      }
      std::partial_sort(d.begin(), d.begin() + k, d.end());
      for (int j = 1; j < k; j++) {
        double ratio = d[0].first / d[j].first;  // float / float, as in the reference
        double dx = txy[2 * d[0].second] - txy[2 * d[j].second], dy = txy[2 * d[0].second + 1] - txy[2 * d[j].second + 1];
        double dist1 = dx * dx + dy * dy;
        if (sqminratio >= 1.0) {
          if ((j == nn - 1) || (dist1 > contrDistSq)) { per_q[i] = {i, d[0].second, d[j].second, d[1].second, d[0].first, d[j].first, d[1].first}; has[i] = 1; break; }
        } else {
          if (ratio <= sqminratio) { per_q[i] = {i, d[0].second, d[j].second, d[1].second, d[0].first, d[j].first, d[1].first}; has[i] = 1; break; }
          if (dist1 > contrDistSq) break;
        }
      }
    }
  }
  for (int i = 0; i < nq; i++) if (has[i]) out.push_back(per_q[i]);
  return out;
}

// ------------------------------------------------------------------------------------------
// matching/matching.cpp:607-666  MatchFLANNDistance with binary_matcher = linear (exact 2-NN), binary_dist = Hamming.
// Descriptor entries are floored to bytes (:631-632, :638-639); distance = number of differing bits (cvflann::Hamming, int);
// a query is kept when its first distance is <= (int)(float)matchDistanceThreshold (:609, :653); ratio = d1 / d2 in double (:659,
// inf / nan for d2 = 0 as in the reference).  Ties between equal distances: lower train index first (FLANN's order is unpinned).
// Fewer than 2 trains is undefined in the reference (knnSearch with knn = 2): returns nothing here.
// ------------------------------------------------------------------------------------------
struct HammingTentative { int q, i0, i1; int d0, d1; double ratio; };

inline std::vector<HammingTentative> matchHamming(const float* q, int nq, const float* t, int nt, int dim, double matchDistanceThreshold) {
  std::vector<HammingTentative> out;
  if (nq == 0 || nt < 2) return out;
  const int max_distance = (int)float(matchDistanceThreshold);
  std::vector<unsigned char> q8((size_t)nq * dim), t8((size_t)nt * dim);
  for (size_t i = 0; i < q8.size(); i++) q8[i] = (unsigned char)std::floor(q[i]);
  for (size_t i = 0; i < t8.size(); i++) t8[i] = (unsigned char)std::floor(t[i]);
  std::vector<HammingTentative> per_q(nq);
  std::vector<char> has(nq, 0);
#pragma omp parallel for schedule(dynamic, 16)
  for (int i = 0; i < nq; i++) {
    const unsigned char* a = q8.data() + (size_t)i * dim;
    int b0 = -1, b1 = -1, d0 = 1 << 30, d1 = 1 << 30;
    for (int j = 0; j < nt; j++) {
      const unsigned char* b = t8.data() + (size_t)j * dim;
      int s = 0;
      for (int e = 0; e < dim; e++) s += __builtin_popcount((unsigned)(a[e] ^ b[e]));
      if (s < d0) { d1 = d0; b1 = b0; d0 = s; b0 = j; }
      else if (s < d1) { d1 = s; b1 = j; }
    }
    if (d0 <= max_distance) { per_q[i] = {i, b0, b1, d0, d1, (double)d0 / (double)d1}; has[i] = 1; }
  }
  for (int i = 0; i < nq; i++) if (has[i]) out.push_back(per_q[i]);
  return out;
}

// ------------------------------------------------------------------------------------------
// degensac: scorers (closed form, f64)
// ------------------------------------------------------------------------------------------
// Htools.c:17-55 lin_hg + :132-196 pinvJ/HDs, fused: lin is never materialised.
// Z rows (Htools.c:28-50): r1 = h0*x2 + h3*y2 + h6 + h2*(-x1*x2) + h5*(-x1*y2) + h8*(-x1), accumulated for j=0..8 in
// the order of the 9 columns of Z.
inline void HDs(const double* u, const double* H, double* p, int len) {
  for (int i = 0; i < len; i++, u += 6) {
    const double x1 = u[0], y1 = u[1];
    const double z1[9] = {u[3], 0, -x1 * u[3], u[4], 0, -x1 * u[4], u[5], 0, -x1 * u[5]};
    const double z2[9] = {0, u[3], -y1 * u[3], 0, u[4], -y1 * u[4], 0, u[5], -y1 * u[5]};
    double r1 = 0, r2 = 0;
    for (int j = 0; j < 9; j++) { r1 += H[j] * z1[j]; r2 += H[j] * z2[j]; }
    double a = H[0] - H[2] * u[0];
    double b = H[3] - H[5] * u[0];
    double c = -H[8] - H[2] * u[3] - H[5] * u[4];
    double d = H[1] - H[2] * u[1];
    double e = H[4] - H[5] * u[1];
    double pJ[8];
    {  // pinvJ, Htools.c:132-156
      double a2 = a * a, b2 = b * b, c2 = c * c, d2 = d * d, e2 = e * e;
      double c2pd2 = c2 + d2, ab = a * b, de = d * e;
      double Q = c * (c2pd2 + e2);
      pJ[0] = -b * de + a * (c2 + e2);
      pJ[1] = b * c2pd2 - a * de;
      pJ[2] = Q;
      pJ[3] = -c * (a * d + b * e);
      pJ[4] = d * (b2 + c2) - ab * e;
      pJ[5] = -ab * d + e * (a2 + c2);
      pJ[6] = pJ[3];
      pJ[7] = c * (a2 + b2 + c2);
      double N = a * pJ[0] + b * pJ[1] + c * pJ[2];
      for (int q = 0; q < 8; q++) pJ[q] /= N;
    }
    double s = 0;
    for (int j = 0; j < 4; j++) { double v = pJ[j] * r1 + pJ[j + 4] * r2; s += v * v; }
    p[i] = s;
  }
}

// matutls/minv.c is a general Gauss-Jordan inverse; for parity we call the same algorithm
// restated for n = 3 (partial pivoting, in place).  Returns false when singular.
bool minv3(double* a);

inline void HDsSymImpl(const double* u, const double* H, double* p, int len, bool useMax) {  // Htools.c:199-282
  double Hinv[9] = {H[0], H[3], H[6], H[1], H[4], H[7], H[2], H[5], H[8]}, H1[9];
  for (int i = 0; i < 9; i++) H1[i] = Hinv[i];
  minv3(H1);
  for (int i = 0; i < len; i++, u += 6) {
    double a = H1[6] * u[0] + H1[7] * u[1] + H1[8];
    double b = Hinv[6] * u[3] + Hinv[7] * u[4] + Hinv[8];
    double xa = (H1[0] * u[0] + H1[1] * u[1] + H1[2]) / a;
    double ya = (H1[3] * u[0] + H1[4] * u[1] + H1[5]) / a;
    double xdiff = u[3] - xa, ydiff = u[4] - ya;
    double d1 = xdiff * xdiff + ydiff * ydiff;
    xa = (Hinv[0] * u[3] + Hinv[1] * u[4] + Hinv[2]) / b;
    ya = (Hinv[3] * u[3] + Hinv[4] * u[4] + Hinv[5]) / b;
    xdiff = u[0] - xa; ydiff = u[1] - ya;
    double d2 = xdiff * xdiff + ydiff * ydiff;
    p[i] = useMax ? (d1 < d2 ? d2 : d1) : d1 + d2;
  }
}

// sym: 0 FDs (Ftools.c:82-100), 1 FDsSym (:102-123), 2 the residual exFDsSym returns (:172-196: r^2 / (ab / (a + b)))
inline void FDsImpl(const double* u, const double* F, double* p, int len, int sym) {
  for (int i = 0; i < len; i++, u += 6) {
    const double u1 = u[0], u2 = u[1], u4 = u[3], u5 = u[4];
    double rxc = F[0] * u4 + F[3] * u5 + F[6];
    double ryc = F[1] * u4 + F[4] * u5 + F[7];
    double rwc = F[2] * u4 + F[5] * u5 + F[8];
    double r = (u1 * rxc + u2 * ryc + rwc);
    double rx = F[0] * u1 + F[1] * u2 + F[2];
    double ry = F[3] * u1 + F[4] * u2 + F[5];
    if (!sym) p[i] = r * r / (rxc * rxc + ryc * ryc + rx * rx + ry * ry);
    else {
      double a = rxc * rxc + ryc * ryc, b = rx * rx + ry * ry;
      if (sym == 1) p[i] = r * r * (a + b) / (a * b);
      else { const double w = (a * b) / (a + b); p[i] = r * r / w; }
    }
  }
}

}  // namespace mo
#endif
