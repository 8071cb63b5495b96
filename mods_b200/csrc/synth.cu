// View synthesis on the GPU: GenerateSynthImageCorr (synth-detection.cpp:236-430) =
//   cv::warpAffine(rotation, INTER_LINEAR, border 128) -> cv::GaussianBlur(anisotropic anti-aliasing, BORDER_REFLECT_101)
//   -> cv::warpAffine(tilt / zoom squeeze).
// The OpenCV arithmetic (fixed-point destination -> source coordinates at 1/32 px, float bilinear table, separable float
// filter) is the one restated in oracle/cvmath.h; the kernels spell the same operation order (library built with
// -fmad=false), so the synthesised view is bit-identical to the oracle's.  Pure streaming work: HBM bound.
#include "common.cuh"
#undef MB2_NS
#define MB2_NS mb2_synth_detail
#include "pyramid.cuh"

#include <cmath>
#include <vector>

namespace MB2_NS {

struct AffineInv { double m[6]; };   // inverted 2x3 map (destination -> source), as cv::warpAffine forms it

// one thread per destination pixel
__global__ void k_warp_affine(const float* __restrict__ src, int srows, int scols, int spitch, AffineInv A, float* __restrict__ dst, int drows,
                              int dcols, int dpitch, float cval) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= dcols || y >= drows) return;
  const int AB_BITS = 10, INTER_BITS = 5, TAB = 1 << INTER_BITS;
  const double AB_SCALE = 1024.0;
  const int adelta = __double2int_rn(__dmul_rn(__dmul_rn(A.m[0], (double)x), AB_SCALE));
  const int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(A.m[3], (double)x), AB_SCALE));
  const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(A.m[1], (double)y), A.m[2]), AB_SCALE)) + 16;
  const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(A.m[4], (double)y), A.m[5]), AB_SCALE)) + 16;
  const int X = (X0 + adelta) >> (AB_BITS - INTER_BITS), Y = (Y0 + bdelta) >> (AB_BITS - INTER_BITS);
  int sx = X >> INTER_BITS, sy = Y >> INTER_BITS;
  sx = max(-32768, min(32767, sx)); sy = max(-32768, min(32767, sy));
  const int fx = X & (TAB - 1), fy = Y & (TAB - 1);
  const float tx1 = fmul((float)fx, 1.f / 32), tx0 = fsub(1.f, tx1), ty1 = fmul((float)fy, 1.f / 32), ty0 = fsub(1.f, ty1);
  const float w0 = fmul(ty0, tx0), w1 = fmul(ty0, tx1), w2 = fmul(ty1, tx0), w3 = fmul(ty1, tx1);
  float v;
  if (sx >= 0 && sy >= 0 && sx < scols - 1 && sy < srows - 1) {
    const float* S = src + (size_t)sy * spitch + sx;
    v = fadd(fadd(fadd(fmul(S[0], w0), fmul(S[1], w1)), fmul(S[spitch], w2)), fmul(S[spitch + 1], w3));
  } else if (sx >= scols || sx + 1 < 0 || sy >= srows || sy + 1 < 0) {
    v = cval;
  } else {
    const bool x0 = sx >= 0 && sx < scols, x1 = sx + 1 >= 0 && sx + 1 < scols, y0 = sy >= 0 && sy < srows, y1 = sy + 1 >= 0 && sy + 1 < srows;
    const float v0 = (x0 && y0) ? src[(size_t)sy * spitch + sx] : cval, v1 = (x1 && y0) ? src[(size_t)sy * spitch + sx + 1] : cval;
    const float v2 = (x0 && y1) ? src[(size_t)(sy + 1) * spitch + sx] : cval, v3 = (x1 && y1) ? src[(size_t)(sy + 1) * spitch + sx + 1] : cval;
    v = fadd(fadd(fadd(fmul(v0, w0), fmul(v1, w1)), fmul(v2, w2)), fmul(v3, w3));
  }
  dst[(size_t)y * dpitch + x] = v;
}

__device__ __forceinline__ int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) { if (p < 0) p = -p; else p = 2 * len - 2 - p; }
  return p;
}
// row pass: generic left-to-right sum, or 2.4.9's symmetric small-kernel order for 3 / 5 taps
__global__ void k_blur101_rows(const float* __restrict__ src, int rows, int cols, int pitch, const float* __restrict__ k, int n, float* __restrict__ dst) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= cols || y >= rows) return;
  const float* s = src + (size_t)y * pitch;
  const int h = n / 2;
  float acc;
  if (n == 3) acc = fadd(fmul(s[x], k[1]), fmul(fadd(s[reflect101(x - 1, cols)], s[reflect101(x + 1, cols)]), k[2]));
  else if (n == 5)
    acc = fadd(fadd(fmul(s[x], k[2]), fmul(fadd(s[reflect101(x - 1, cols)], s[reflect101(x + 1, cols)]), k[3])),
               fmul(fadd(s[reflect101(x - 2, cols)], s[reflect101(x + 2, cols)]), k[4]));
  else {
    acc = fmul(k[0], s[reflect101(x - h, cols)]);
    for (int j = 1; j < n; j++) acc = fadd(acc, fmul(k[j], s[reflect101(x - h + j, cols)]));
  }
  dst[(size_t)y * pitch + x] = acc;
}
// column pass: centre tap, then pairs (SymmColumnFilter order)
__global__ void k_blur101_cols(const float* __restrict__ src, int rows, int cols, int pitch, const float* __restrict__ k, int n, float* __restrict__ dst) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= cols || y >= rows) return;
  const int h = n / 2;
  float d = fmul(k[h], src[(size_t)y * pitch + x]);
  for (int j = 1; j <= h; j++)
    d = fadd(d, fmul(k[h + j], fadd(src[(size_t)reflect101(y + j, rows) * pitch + x], src[(size_t)reflect101(y - j, rows) * pitch + x])));
  dst[(size_t)y * pitch + x] = d;
}

inline void invert_affine(const double* Mf, AffineInv* out) {   // cv::warpAffine's in-place inversion
  double M[6];
  for (int i = 0; i < 6; i++) M[i] = Mf[i];
  double D = M[0] * M[4] - M[1] * M[3];
  D = D != 0 ? 1. / D : 0;
  const double A11 = M[4] * D, A22 = M[0] * D;
  M[0] = A11; M[1] *= -D; M[3] *= -D; M[4] = A22;
  const double b1 = -M[0] * M[2] - M[1] * M[5];
  const double b2 = -M[3] * M[2] - M[4] * M[5];
  M[2] = b1; M[5] = b2;
  for (int i = 0; i < 6; i++) out->m[i] = M[i];
}
inline std::vector<float> gauss_kernel(int n, double sigma) {   // cv::getGaussianKernel(n, sigma, CV_32F)
  std::vector<float> cf(n);
  const double scale2X = -0.5 / (sigma * sigma);
  double sum = 0;
  for (int i = 0; i < n; i++) { const double x = i - (n - 1) * 0.5; cf[i] = (float)std::exp(scale2X * x * x); sum += cf[i]; }
  sum = 1. / sum;
  for (int i = 0; i < n; i++) cf[i] = (float)(cf[i] * sum);
  return cf;
}
inline int pitch_of(int cols) { return (cols + 31) & ~31; }

}  // namespace MB2_NS

// GenerateSynthImageCorr on the device image `in`.  The view ends in ctx->synth_b (or is `in` itself for the identity
// view); H = SynthImage::H (original -> view).  Geometry, sigmas and the cos / sin are host arithmetic, as in the reference.
int mb2_synth_core(mb2_ctx* ctx, const ImgView& in, const mb2_view_params& vp, ImgView* out, double* H) {
  using namespace MB2_NS;
  double tilt = vp.tilt; const double phi = vp.phi, zoom = vp.zoom, InitSigma = vp.InitSigma;
  bool vertical_tilt = false;
  if (tilt < 0) { tilt = -tilt; vertical_tilt = true; }
  const int zoomed = std::fabs(zoom - 1.0f) >= 0.05 ? 1 : 0;
  const int w = in.cols, h = in.rows;
  const int wS1 = (int)(w * zoom), hS1 = (int)(h * zoom);
  for (int i = 0; i < 9; i++) H[i] = (i % 4 == 0) ? 1.0 : 0.0;
  if ((std::fabs(tilt - 1.) <= 0.1) && (std::fabs(phi) <= 0.2) && (std::fabs(zoom - 1.) <= 0.1)) { *out = in; return 1; }   // :278-289
  if (wS1 <= 0 || hS1 <= 0) { ctx->set_error("synth view: zoom too small"); return MB2_ERR_ARG; }
  double d, d2, w_new, h_new, kV = 1., kH = 1.;
  if (zoomed) { kV = (double)w / (double)wS1; kH = (double)h / (double)hS1; }
  const double c = std::cos(phi), s = std::sin(phi);
  const bool q1 = (phi >= 0) && (phi < M_PI / 2);
  const double sx = vertical_tilt ? kH : tilt * kH, sy = vertical_tilt ? tilt * kV : kV;
  double warpRot[6]; int w_rot, h_rot;
  if (q1) {
    w_new = std::floor((0.5 + c * w + s * h) / sx); h_new = std::floor((0.5 + s * w + c * h) / sy);
    H[0] = c / sx; H[1] = s / sx; H[2] = 0; H[3] = -s / sy; H[4] = c / sy; H[5] = std::floor(0.5 + s * w / sy);
    w_rot = (int)std::floor(0.5 + c * w + s * h); h_rot = (int)std::floor(0.5 + s * w + c * h);
    warpRot[0] = c; warpRot[1] = s; warpRot[2] = 0; warpRot[3] = -s; warpRot[4] = c; warpRot[5] = std::floor(0.5 + s * w);
  } else {
    w_new = std::floor((0.5 - c * w + s * h) / sx); h_new = std::floor((0.5 + s * w - c * h) / sy);
    d = -std::floor(c * w / sx); d2 = std::floor(0.5 + (s * w - c * h) / sy);
    H[0] = c / sx; H[1] = s / sx; H[2] = d; H[3] = -s / sy; H[4] = c / sy; H[5] = d2;
    w_rot = (int)std::floor(0.5 - c * w + s * h); h_rot = (int)std::floor(0.5 + s * w - c * h);
    d = -std::floor(c * w); d2 = std::floor(0.5 + (s * w - c * h));
    warpRot[0] = c; warpRot[1] = s; warpRot[2] = d; warpRot[3] = -s; warpRot[4] = c; warpRot[5] = d2;
  }
  H[6] = 0; H[7] = 0; H[8] = 1;
  const int wn = (int)w_new, hn = (int)h_new;
  if (w_rot <= 0 || h_rot <= 0 || wn <= 0 || hn <= 0) { ctx->set_error("synth view: empty view"); return MB2_ERR_ARG; }
  const double sigma_aa_2 = zoomed ? InitSigma / (4.0 * zoom) : InitSigma / 2.0;
  const double sigma_aa = InitSigma * tilt / (2.0 * zoom);
  const double sigma_x = vertical_tilt ? sigma_aa_2 : sigma_aa, sigma_y = vertical_tilt ? sigma_aa : sigma_aa_2;

  const int prot = pitch_of(w_rot), pout = pitch_of(wn);
  MB2_CUDA_CHECK(ctx, ctx->synth_a.reserve((size_t)prot * h_rot * 4));
  MB2_CUDA_CHECK(ctx, ctx->synth_c.reserve((size_t)prot * h_rot * 4));
  MB2_CUDA_CHECK(ctx, ctx->synth_b.reserve((size_t)pout * hn * 4));
  float* rot = ctx->synth_a.as<float>(); float* tmp = ctx->synth_c.as<float>(); float* outp = ctx->synth_b.as<float>();
  AffineInv A;
  invert_affine(warpRot, &A);
  const dim3 blk(32, 8);
  MB2_LAUNCH(ctx, k_warp_affine, dim3((w_rot + 31) / 32, (h_rot + 7) / 8), blk, 0, in.p, in.rows, in.cols, in.pitch, A, rot, h_rot, w_rot, prot, 128.f);
  if (vp.doBlur) {
    int kx = (int)std::floor(2.0 * 3.0 * sigma_x + 1.0); if (kx % 2 == 0) kx++; if (kx < 3) kx = 3;
    int ky = (int)std::floor(2.0 * 3.0 * sigma_y + 1.0); if (ky % 2 == 0) ky++; if (ky < 3) ky = 3;
    std::vector<float> taps = gauss_kernel(kx, sigma_x), tky = gauss_kernel(ky, sigma_y);
    taps.insert(taps.end(), tky.begin(), tky.end());
    MB2_CUDA_CHECK(ctx, ctx->synth_k.reserve(taps.size() * 4));
    MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->synth_k.p, taps.data(), taps.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));   // `taps` is a stack-owned staging buffer
    MB2_LAUNCH(ctx, k_blur101_rows, dim3((w_rot + 31) / 32, (h_rot + 7) / 8), blk, 0, rot, h_rot, w_rot, prot, ctx->synth_k.as<float>(), kx, tmp);
    MB2_LAUNCH(ctx, k_blur101_cols, dim3((w_rot + 31) / 32, (h_rot + 7) / 8), blk, 0, tmp, h_rot, w_rot, prot, ctx->synth_k.as<float>() + kx, ky, rot);
  }
  const double wtz[6] = {1.0 / sx, 0, 0, 0, 1.0 / sy, 0};
  invert_affine(wtz, &A);
  MB2_LAUNCH(ctx, k_warp_affine, dim3((wn + 31) / 32, (hn + 7) / 8), blk, 0, rot, h_rot, w_rot, prot, A, outp, hn, wn, pout, 128.f);
  out->p = outp; out->rows = hn; out->cols = wn; out->pitch = pout;
  return 0;
}
