// Dominant orientation (one warp per region) and reprojection / boundary filtering.
//
// Reference: DetectOrientation (synth-detection.cpp:841-919), EstimateDominantAnglesFunctor
// (:746-839), smoothCircularBuffer / addPeakAngle (:721-744), computeGradientMagnitudeAndOrientation
// (detectors/helpers.cpp:840-863), ReprojectRegions / ReprojectByH (synth-detection.cpp:541-616,
// 490-498).
//
// Bit parity: the 36-bin histogram is a float sum per bin in raster order of the 39x41 inner
// patch; each lane owns one (or two) bins and scans the pixels serially.
#include "common.cuh"
#undef MB2_NS
#define MB2_NS mb2_orient_detail
#include "pyramid.cuh"

namespace MB2_NS {

constexpr int PS = 41, NPIX = PS * PS, OW = 4;  // warps per block
constexpr double K_SIGMA = 2 * 3.0 * 1.7320508075688772;  // synth-detection.cpp:28

__global__ void __launch_bounds__(OW * 32)
k_orientation(ImgView img, const KeyOut* __restrict__ in, int n, OrientParams op, const float* __restrict__ orimask,
              KeyOut* __restrict__ out, int* __restrict__ out_count) {
  __shared__ float s_patch[OW][NPIX];
  __shared__ float s_hist[OW][40];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kidx = blockIdx.x * OW + warp;
  if (kidx >= n) return;
  const KeyOut k = in[kidx];
  float* patch = s_patch[warp]; float* hist = s_hist[warp];
  const int maxA = op.maxAngles;
  KeyOut* dst = out + (size_t)kidx * maxA;
  for (int j = lane; j < maxA; j += 32) dst[j].keep = 0;
  if (lane == 0) out_count[kidx] = 0;

  const float x = (float)k.v[0], y = (float)k.v[1];
  const float a11 = (float)k.v[2], a12 = (float)k.v[3], a21 = (float)k.v[4], a22 = (float)k.v[5];
  const double s = k.v[6];
  const int res = (int)(K_SIGMA * s);
  if (interpolateCheckBorders_dev(img.cols, img.rows, x, y, a11, a12, a21, a22, res, res)) return;
  if (maxA <= 0) return;

  const double mrScale = op.mrSize;
  const int patchImageSize = 2 * (int)mrScale + 1;
  const double imageToPatchScale = (double)patchImageSize / (double)op.patchSize;
  const float curr_sc = (float)(imageToPatchScale * s);
  const float A11 = fmul(a11, curr_sc), A12 = fmul(a12, curr_sc), A21 = fmul(a21, curr_sc), A22 = fmul(a22, curr_sc);
  const bool touch = interpolateCheckBorders_dev(img.cols, img.rows, x, y, A11, A12, A21, A22, PS, PS);
  for (int row = lane; row < PS; row += 32)
    interpolate_row(img.p, img.rows, img.cols, img.pitch, x, y, A11, A12, A21, A22, PS, PS, touch, row,
                    [&](int i, float v) { patch[row * PS + i] = v; });
  __syncwarp();
  // Gradient magnitude / orientation on the inner 39x39 (helpers.cpp:840-863), then the 36-bin
  // histogram: bin b is a float sum over its pixels in raster order.  Lane L owns bins L and L+32;
  // each slab of 32 consecutive pixels is evaluated one pixel per lane and then replayed in pixel
  // order through shuffles, so every bin sees its contributions in the reference's order.
  const float PIf = 3.14159265358979323846f;
  {
    float h0 = 0.f, h1 = 0.f;
    const int b0 = lane, b1 = lane + 32;
    for (int p0 = PS; p0 < NPIX - PS; p0 += 32) {
      const int p = p0 + lane;
      int b = 255; float w = 0.f;
      if (p < NPIX - PS) {
        const int r = p / PS, c = p - r * PS;
        if (c >= 1 && c < PS - 1) {
          const float xg = fsub(patch[p + 1], patch[p - 1]);
          const float yg = fsub(patch[p + PS], patch[p - PS]);
          const float mag = sqrtf(fadd(fmul(xg, xg), fmul(yg, yg)));
          const float m = orimask[p];
          if (m > 0 && (double)mag > 1.0) {
            const float ori = atan2LUTff_dev(yg, xg);
            b = (int)fdiv(fmul(36.f, fadd(fdiv(ori, PIf), 1.0f)), 2.0f);  // 36 (ori == +pi): a slot the reference never reads
            w = fmul(mag, m);
          }
        }
      }
      if (__ballot_sync(0xffffffffu, b != 255) == 0) continue;
#pragma unroll
      for (int j = 0; j < 32; j++) {
        const int bj = __shfl_sync(0xffffffffu, b, j);
        const float wj = __shfl_sync(0xffffffffu, w, j);
        if (bj == b0) h0 = fadd(h0, wj);
        else if (bj == b1) h1 = fadd(h1, wj);
      }
    }
    hist[b0] = h0;
    if (b1 < 36) hist[b1] = h1;
  }
  __syncwarp();
  if (lane == 0) {
    float h[36];
    for (int i = 0; i < 36; i++) h[i] = hist[i];
    for (int it = 0; it < 6; it++) {  // smoothCircularBuffer<36>
      float first = h[0], prev = h[35];
      for (int i = 0; i < 35; i++) { float cur = h[i]; h[i] = fadd(fadd(prev, cur), h[i + 1]); prev = cur; }
      h[35] = fadd(fadd(prev, h[35]), first);
    }
    float thresh = 0.0f;
    for (int i = 0; i < 36; i++) if (h[i] > thresh) thresh = h[i];
    thresh = (float)((double)thresh * op.threshold);
    int cnt = 0;
    for (int q = 0; q < 36 && cnt < maxA; q++) {
      // order of addPeakAngle calls: (35,0,1), (i-1,i,i+1) for i = 1..34, (34,35,0)
      const int a = (q == 0) ? 35 : q - 1, b = q, c = (q == 35) ? 0 : q + 1;
      if (h[b] >= thresh && h[b] > h[a] && h[b] > h[c]) {
        const float pp = fdiv(fdiv(fsub(h[a], h[c]), fadd(fsub(h[a], fmul(2.0f, h[b])), h[c])), 2.0f);
        const float ang = fsub(fdiv(fmul(fmul(2.0f, PIf), fadd(fadd((float)b, 0.5f), pp)), 36.f), PIf);
        // The rotation needs cos(-ang), sin(-ang), which the reference takes from the host libm in
        // FLOAT (std::cos(float), synth-detection.cpp:902-903); libm's cosf/sinf are not correctly
        // rounded and differ between libm builds, so the host adapter evaluates them
        // and k_apply_rotation finishes the frame.
        KeyOut o = k;
        o.keep = 1;
        o.order = k.order;  // peak rank is implied by the slot index
        o.pad = __float_as_int(ang);
        dst[cnt++] = o;
      }
    }
    out_count[kidx] = cnt;
  }
}

// A <- A * R(-theta) with (ci, si) = (cos(-theta), sin(-theta)) from the host (synth-detection.cpp:902-910)
__global__ void k_extract_angles(const KeyOut* __restrict__ keys, int n, float* __restrict__ ang) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ang[i] = __int_as_float(keys[i].pad);
}
__global__ void k_apply_rotation(KeyOut* __restrict__ keys, const double* __restrict__ cs, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double ci = cs[2 * i], si = cs[2 * i + 1];
  KeyOut k = keys[i];
  const double a11 = k.v[2], a12 = k.v[3], a21 = k.v[4], a22 = k.v[5];
  k.v[2] = a11 * ci - a12 * si;
  k.v[3] = a11 * si + a12 * ci;
  k.v[4] = a21 * ci - a22 * si;
  k.v[5] = a21 * si + a22 * ci;
  k.pad = 0;
  keys[i] = k;
}

// ReprojectRegions (synth-detection.cpp:541-616): reproj_kp = Hinv (affine part) applied to centre
// and shape (identity view: copy), then drop regions whose centre or k_sigma*s box leaves the
// ORIGINAL image.  det[i].keep is cleared in place for dropped regions; reproj gets the new frame.
__global__ void k_reproject(KeyOut* __restrict__ det, int n, const double* __restrict__ Hinv, int h_is_eye, int orig_w, int orig_h,
                            KeyOut* __restrict__ reproj) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  KeyOut d = det[i];
  KeyOut o = d;
  if (!h_is_eye) {
    o.v[0] = (Hinv[0] * d.v[0] + Hinv[1] * d.v[1] + Hinv[2]);
    o.v[1] = (Hinv[3] * d.v[0] + Hinv[4] * d.v[1] + Hinv[5]);
    o.v[2] = (Hinv[0] * d.v[2] + Hinv[1] * d.v[4]);
    o.v[3] = (Hinv[0] * d.v[3] + Hinv[1] * d.v[5]);
    o.v[4] = (Hinv[3] * d.v[2] + Hinv[4] * d.v[4]);
    o.v[5] = (Hinv[3] * d.v[3] + Hinv[4] * d.v[5]);
  }
  int keep = d.keep;
  if (keep) {
    keep = 0;
    if ((o.v[0] < orig_w) && (o.v[1] < orig_h) && (o.v[0] > 0) && (o.v[1] > 0)) {
      const int res = (int)(K_SIGMA * o.v[6]);
      if (!interpolateCheckBorders_dev(orig_w, orig_h, (float)o.v[0], (float)o.v[1], (float)o.v[2], (float)o.v[3], (float)o.v[4],
                                       (float)o.v[5], res, res))
        keep = 1;
    }
  }
  o.keep = keep;
  det[i].keep = keep;
  reproj[i] = o;
}

}  // namespace
using namespace MB2_NS;

void mb2_launch_orientation(mb2_ctx* ctx, const ImgView& img, const KeyOut* in, int n, const OrientParams& op,
                            const float* d_orimask, KeyOut* out, int* out_count_per_kp) {
  if (!n) return;
  MB2_LAUNCH(ctx, k_orientation, (n + OW - 1) / OW, OW * 32, 0, img, in, n, op, d_orimask, out, out_count_per_kp);
}

void mb2_launch_extract_angles(mb2_ctx* ctx, const KeyOut* keys, int n, float* d_ang) {
  if (!n) return;
  MB2_LAUNCH(ctx, k_extract_angles, (n + 127) / 128, 128, 0, keys, n, d_ang);
}

void mb2_launch_apply_rotation(mb2_ctx* ctx, KeyOut* keys, const double* d_cs, int n) {
  if (!n) return;
  MB2_LAUNCH(ctx, k_apply_rotation, (n + 127) / 128, 128, 0, keys, d_cs, n);
}

void mb2_launch_reproject(mb2_ctx* ctx, const KeyOut* det, int n, const double* Hinv9, int h_is_eye, int orig_w, int orig_h,
                          KeyOut* reproj) {
  if (!n) return;
  MB2_LAUNCH(ctx, k_reproject, (n + 127) / 128, 128, 0, (KeyOut*)det, n, Hinv9, h_is_eye, orig_w, orig_h, reproj);
}
