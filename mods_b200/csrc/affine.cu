// Baumberg affine-shape iteration, one warp per localized scale-space point.
//
// Reference: AffineShape::findAffineShape (detectors/affinedetectors/affine.cpp:26-169) with
// interpolate (helpers.cpp:551-626), computeGradient (:779-797), invSqrt (:463-502),
// getEigenvalues (:504-515); samples the level BELOW the detection level (pyramid.cpp:429).
//
// Bit parity: the second-moment sums a, b, c are float sums over the 19x19 window in raster
// order; float addition is not associative, so three lanes walk the 361 products serially while
// the sampling / gradient / product phases use the whole warp.
#include "common.cuh"
#undef MB2_NS
#define MB2_NS mb2_affine_detail
#include "pyramid.cuh"

namespace MB2_NS {

constexpr int WARPS_PER_BLOCK = 4;
constexpr int MAXW = 19;  // smmWindowSize

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
k_affine_shape(const OctaveLevels* __restrict__ octaves, KeypointRec* __restrict__ kps, int n, AffineParams ap,
               const float* __restrict__ smm_mask) {
  __shared__ float s_img[WARPS_PER_BLOCK][MAXW * MAXW];
  __shared__ float s_pa[WARPS_PER_BLOCK][MAXW * MAXW];
  __shared__ float s_pb[WARPS_PER_BLOCK][MAXW * MAXW];
  __shared__ float s_pc[WARPS_PER_BLOCK][MAXW * MAXW];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kidx = blockIdx.x * WARPS_PER_BLOCK + warp;
  if (kidx >= n) return;
  KeypointRec kp = kps[kidx];
  if (!ap.doBaumberg) { if (lane == 0) kps[kidx].ok = 1; return; }
  const ImgView blur = octaves[kp.octave].blur[kp.level - 1];  // prevBlur
  const int W = ap.smmWindowSize, NP = W * W;
  float* img = s_img[warp]; float* pa = s_pa[warp]; float* pb = s_pb[warp]; float* pc = s_pc[warp];

  float eigen_ratio_act = 0.0f, eigen_ratio_bef = 0.0f;
  float u11 = 1.0f, u12 = 0.0f, u21 = 0.0f, u22 = 1.0f, l1 = 1.0f, l2 = 1.0f;
  const float lx = fdiv(kp.x, kp.pixelDistance), ly = fdiv(kp.y, kp.pixelDistance);
  const float ratio = fdiv(kp.s, fmul(ap.initialSigma, kp.pixelDistance));
  int ok = 0;
  for (int l = 0; l < ap.maxIterations; l++) {
    const float A11 = fmul(u11, ratio), A12 = fmul(u12, ratio), A21 = fmul(u21, ratio), A22 = fmul(u22, ratio);
    const bool touch = interpolateCheckBorders_dev(blur.cols, blur.rows, lx, ly, A11, A12, A21, A22, W, W);
    if (lane < W)
      interpolate_row(blur.p, blur.rows, blur.cols, blur.pitch, lx, ly, A11, A12, A21, A22, W, W, touch, lane,
                      [&](int i, float v) { img[lane * W + i] = v; });
    __syncwarp();
    for (int p = lane; p < NP; p += 32) {
      const int r = p / W, c = p - r * W;
      float gx, gy;  // computeGradient, helpers.cpp:779-797
      if (c == 0) gx = fsub(img[p + 1], img[p]);
      else if (c == W - 1) gx = fsub(img[p], img[p - 1]);
      else gx = fsub(img[p + 1], img[p - 1]);
      if (r == 0) gy = fsub(img[p + W], img[p]);
      else if (r == W - 1) gy = fsub(img[p], img[p - W]);
      else gy = fsub(img[p + W], img[p - W]);
      const float v = smm_mask[p];
      const float gxy = fmul(gx, gy);
      pa[p] = fmul(fmul(gx, gx), v);
      pb[p] = fmul(gxy, v);
      pc[p] = fmul(fmul(gy, gy), v);
    }
    __syncwarp();
    float acc = 0.f;
    if (lane < 3) {
      const float* arr = lane == 0 ? pa : (lane == 1 ? pb : pc);
      for (int p = 0; p < NP; ++p) acc = fadd(acc, arr[p]);
    }
    float a = __shfl_sync(0xffffffffu, acc, 0), b = __shfl_sync(0xffffffffu, acc, 1), c = __shfl_sync(0xffffffffu, acc, 2);
    __syncwarp();
    a = fdiv(a, (float)NP); b = fdiv(b, (float)NP); c = fdiv(c, (float)NP);
    {  // invSqrt, helpers.cpp:463-502 (double; -fmad=false keeps the written order)
      double t, r;
      if (b != 0) {
        r = (double)fsub(c, a) / (double)fmul(2.f, b);
        if (r >= 0) t = 1.0 / (r + sqrt(1 + r * r));
        else t = -1.0 / (-r + sqrt(1 + r * r));
        r = 1.0 / sqrt(1 + t * t);
        t = t * r;
      } else { r = 1; t = 0; }
      double x = 1.0 / sqrt(r * r * a - 2 * r * t * b + t * t * c);
      double z = 1.0 / sqrt(t * t * a + 2 * r * t * b + r * r * c);
      double d = sqrt(x * z);
      x /= d; z /= d;
      if (x < z) { l1 = (float)z; l2 = (float)x; } else { l1 = (float)x; l2 = (float)z; }
      a = (float)(r * r * x + t * t * z);
      b = (float)(-r * t * x + t * r * z);
      c = (float)(t * t * x + r * r * z);
    }
    if ((a != a) || (b != b) || (c != c)) break;
    eigen_ratio_bef = eigen_ratio_act;
    eigen_ratio_act = (float)(1.0 - (double)fdiv(l2, l1));
    const float u11t = u11, u12t = u12;
    u11 = fadd(fmul(a, u11t), fmul(b, u21));
    u12 = fadd(fmul(a, u12t), fmul(b, u22));
    u21 = fadd(fmul(b, u11t), fmul(c, u21));
    u22 = fadd(fmul(b, u12t), fmul(c, u22));
    {  // getEigenvalues, helpers.cpp:504-515
      const float trace = fadd(u11, u22);
      const float delta1 = fsub(fmul(trace, trace), fmul(4.f, fsub(fmul(u11, u22), fmul(u12, u21))));
      if (delta1 < 0) break;
      const float delta = sqrtf(delta1);
      l1 = fdiv(fadd(trace, delta), 2.0f);
      l2 = fdiv(fsub(trace, delta), 2.0f);
    }
    if ((fdiv(l1, l2) > 6) || (fdiv(l2, l1) > 6)) break;
    if (eigen_ratio_act < ap.convergenceThreshold && eigen_ratio_bef < ap.convergenceThreshold) { ok = 1; break; }
  }
  if (lane == 0) {
    KeypointRec* o = kps + kidx;
    o->ok = ok;
    if (ok) { o->a11 = u11; o->a12 = u12; o->a21 = u21; o->a22 = u22; }
  }
}

// AffineDetector::onNormalizedPatchAvailable (scale-space-detector.hpp:70-88) stores the float
// results in the double AffineKeypoint; DetectAffineRegions (synth-detection.hpp:110-124) then
// scales s by sqrt|det A| and rectifies A (synth-detection.cpp:46-55).
__global__ void k_export(const KeypointRec* __restrict__ kps, int n, KeyOut* __restrict__ out, int as_regions) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const KeypointRec k = kps[i];
  KeyOut o;
  double a11 = k.a11, a12 = k.a12, a21 = k.a21, a22 = k.a22, s = k.s;
  if (as_regions) {
    s = s * sqrt(fabs(a11 * a22 - a12 * a21));
    const double a = a11, b = a12, c = a21, d = a22;
    const double det = sqrt(fabs(a * d - b * c));
    const double b2a2 = sqrt(b * b + a * a);
    a11 = b2a2 / det; a12 = 0; a21 = (d * b + c * a) / (b2a2 * det); a22 = det / b2a2;
  }
  o.v[0] = k.x; o.v[1] = k.y; o.v[2] = a11; o.v[3] = a12; o.v[4] = a21; o.v[5] = a22; o.v[6] = s;
  o.v[7] = k.response; o.v[8] = (double)k.type;
  o.order = k.order; o.keep = k.ok; o.pad = 0;
  out[i] = o;
}

}  // namespace
using namespace MB2_NS;

void mb2_launch_affine_shape(mb2_ctx* ctx, const OctaveLevels* d_octaves, int n_octaves, KeypointRec* kps, int n,
                             const AffineParams& ap, const float* d_smm_mask) {
  (void)n_octaves;
  if (!n) return;
  MB2_LAUNCH(ctx, k_affine_shape, (n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, WARPS_PER_BLOCK * 32, 0, d_octaves, kps, n, ap,
             d_smm_mask);
}

void mb2_launch_export(mb2_ctx* ctx, const KeypointRec* kps, int n, KeyOut* out, int as_regions) {
  if (!n) return;
  MB2_LAUNCH(ctx, k_export, (n + 127) / 128, 128, 0, kps, n, out, as_regions);
}
