// TEST INFRASTRUCTURE (oracle) -- see mods_oracle.hpp.  Out-of-line pieces + C API for ctypes.
#include "mods_oracle.hpp"
#include "mser_oracle.hpp"
#include <cstdio>

namespace mo {

// helpers.cpp:30-72.  The reference table is atan(i/255) printed with 10 decimals, except three
// entries whose trailing zeros were overwritten with "55" (32, 83, 100); they are part of the
// reference's behaviour and are reproduced here.
static struct AtanLutInit {
  double v[256];
  AtanLutInit() {
    for (int i = 0; i < 256; i++) {
      char buf[32];
      std::snprintf(buf, sizeof buf, "%.10f", std::atan(i / 255.0));
      v[i] = std::strtod(buf, nullptr);
    }
    v[32] = 0.1248376255; v[83] = 0.3146752558; v[100] = 0.3737268255;
  }
} g_lut_init;
const double* const ATAN_LUT_PTR = g_lut_init.v;
}  // namespace mo

namespace mo {
// matutls/minv.c (ccmath, in-place inverse via column-wise LU with row pivoting), n = 3,
// restated with explicit indices; same operation order, so HDsSym is bit-identical.
bool minv3(double* a) {
  const int n = 3;
  int le[3];
  double q0[3], tq = 0., zr = 1.e-15;
  auto A = [&](int r, int c) -> double& { return a[r * n + c]; };
  for (int j = 0; j < n; ++j) {
    if (j > 0) {
      for (int i = 0; i < n; ++i) q0[i] = A(i, j);
      for (int i = 1; i < n; ++i) {
        int lc = i < j ? i : j;
        double t = 0.;
        for (int k = 0; k < lc; ++k) t += A(i, k) * q0[k];
        q0[i] -= t;
      }
      for (int i = 0; i < n; ++i) A(i, j) = q0[i];
    }
    double s = std::fabs(A(j, j));
    int lc = j;
    for (int k = j + 1; k < n; ++k) { double t = std::fabs(A(k, j)); if (t > s) { s = t; lc = k; } }
    tq = tq > s ? tq : s;
    if (s < zr * tq) return false;
    le[j] = lc;
    if (lc != j) for (int k = 0; k < n; ++k) std::swap(A(j, k), A(lc, k));
    double t = 1. / A(j, j);
    for (int k = j + 1; k < n; ++k) A(k, j) *= t;
    A(j, j) = t;
  }
  for (int j = 1; j < n; ++j) for (int k = 0; k < j; ++k) A(k, j) *= A(j, j);
  for (int j = 1; j < n; ++j) {
    for (int i = 0; i < j; ++i) q0[i] = A(i, j);
    for (int k = 0; k < j; ++k) { double t = 0.; for (int i = k; i < j; ++i) t -= A(k, i) * q0[i]; q0[k] = t; }
    for (int i = 0; i < j; ++i) A(i, j) = q0[i];
  }
  for (int j = n - 2; j >= 0; --j) {
    int m = n - j - 1;
    for (int i = 0; i < m; ++i) q0[i] = A(j + 1 + i, j);
    for (int k = n - 1; k > j; --k) {
      double t = -A(k, j);
      for (int i = j + 1, q = 0; i < k; ++i, ++q) t -= A(k, i) * q0[q];
      q0[--m] = t;
    }
    m = n - j - 1;
    for (int i = 0; i < m; ++i) A(j + 1 + i, j) = q0[i];
  }
  for (int k = 0; k < n - 1; ++k) {
    for (int i = 0; i < n; ++i) q0[i] = A(i, k);
    for (int j = 0; j < n; ++j) {
      double t; int i;
      if (j > k) { t = 0.; i = j; } else { t = q0[j]; i = k + 1; }
      for (; i < n; ++i) t += A(j, i) * q0[i];
      q0[j] = t;
    }
    for (int i = 0; i < n; ++i) A(i, k) = q0[i];
  }
  for (int j = n - 2; j >= 0; --j)
    for (int k = 0; k < n; ++k) std::swap(A(k, j), A(k, le[j]));
  return true;
}
}  // namespace mo

// ---------------------------------------------------------------------------------------------
// C API (ctypes).  Mirrors oracle/ref_api.cpp one to one (orc_* vs ref_*), so tests loop over both.
// Keypoint record (KP = 9 doubles): x y a11 a12 a21 a22 s response sub_type.
// ---------------------------------------------------------------------------------------------
using namespace mo;
namespace {
const int KP = 9;
struct HessParamsC {
  float threshold; int numberOfScales; float initialSigma; float edgeEigenValueRatio; int border;
  int maxIterations; float convergenceThreshold; int smmWindowSize; int doBaumberg;
  int mode; int reg_number; float rel_threshold; float rel_reg_number; int patchSize; float mrSize;
  int detectorType;   // 0 DET_HESSIAN, 1 DET_DOG
};
HessParams to_par(const HessParamsC& c) {
  HessParams p;
  p.threshold = c.threshold; p.numberOfScales = c.numberOfScales; p.initialSigma = c.initialSigma;
  p.edgeEigenValueRatio = c.edgeEigenValueRatio; p.border = c.border; p.maxIterations = c.maxIterations;
  p.convergenceThreshold = c.convergenceThreshold; p.smmWindowSize = c.smmWindowSize; p.doBaumberg = c.doBaumberg;
  p.mode = c.mode; p.reg_number = c.reg_number; p.rel_threshold = c.rel_threshold; p.rel_reg_number = c.rel_reg_number;
  p.patchSize = c.patchSize; p.mrSize = c.mrSize; p.detectorType = c.detectorType;
  return p;
}
void kp_out(const Key& k, double* o) {
  o[0] = k.x; o[1] = k.y; o[2] = k.a11; o[3] = k.a12; o[4] = k.a21; o[5] = k.a22; o[6] = k.s; o[7] = k.response; o[8] = k.sub_type;
}
std::vector<Key> keys_in(const double* kps, int n) {
  std::vector<Key> v(n);
  for (int i = 0; i < n; i++) {
    const double* o = kps + (size_t)i * KP;
    v[i] = Key{o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7], (int)o[8]};
  }
  return v;
}
Image image_in(const float* img, int w, int h) {
  Image im(h, w);
  std::memcpy(im.px.data(), img, sizeof(float) * (size_t)w * h);
  return im;
}
}  // namespace

extern "C" {

void orc_gaussian_blur(const float* src, float* dst, int w, int h, float sigma) { cvmath::gaussian_blur(src, dst, h, w, sigma); }
void orc_resize_half(const float* src, int w, int h, float* dst) { cvmath::resize_half(src, h, w, dst); }
void orc_hessian_response(const float* src, float* dst, int w, int h, float norm) {
  Image o = hessianResponse(image_in(src, w, h), norm);
  std::memcpy(dst, o.px.data(), sizeof(float) * (size_t)w * h);
}
float orc_atan2LUTff(float y, float x) { return atan2LUTff(y, x); }
int orc_interpolate(const float* img, int w, int h, float ox, float oy, float a11, float a12, float a21, float a22, float* out,
                    int ow, int oh) {
  return interpolate(img, h, w, ox, oy, a11, a12, a21, a22, out, oh, ow) ? 1 : 0;
}

int orc_hessaff_detect(const float* img, int w, int h, const HessParamsC* hp, int raw, double* out, int max_out) {
  std::vector<Key> keys = detectAffineKeypoints(image_in(img, w, h), to_par(*hp), 1.0, 1.0);
  if (!raw) toRegions(keys);
  for (size_t i = 0; i < keys.size() && (int)i < max_out; i++) kp_out(keys[i], out + i * KP);
  return (int)keys.size();
}

// Kernel-level dumps for the GPU parity tests: every pyramid level (blur + response) and the
// localized points (before Baumberg) in detection order.
struct OrcPyramid { HessianAffineDetector* d; };
void* orc_pyramid_build(const float* img, int w, int h, const HessParamsC* hp) {
  HessianAffineDetector* d = nullptr;
  detectAffineKeypoints(image_in(img, w, h), to_par(*hp), 1.0, 1.0, &d);
  return d;
}
int orc_pyramid_num_levels(void* p) { return (int)((HessianAffineDetector*)p)->dump_info.size(); }
void orc_pyramid_level_info(void* p, int i, int* octave, int* level, int* rows, int* cols, float* sigma) {
  auto& li = ((HessianAffineDetector*)p)->dump_info[i];
  *octave = li.octave; *level = li.level; *rows = li.rows; *cols = li.cols; *sigma = li.curSigma;
}
void orc_pyramid_level_data(void* p, int i, float* blur, float* resp) {
  auto* d = (HessianAffineDetector*)p;
  std::memcpy(blur, d->dump_blur[i].px.data(), sizeof(float) * d->dump_blur[i].px.size());
  std::memcpy(resp, d->dump_resp[i].px.data(), sizeof(float) * d->dump_resp[i].px.size());
}
int orc_pyramid_counts(void* p, int* extrema, int* localized) {
  auto* d = (HessianAffineDetector*)p;
  *extrema = d->extrema_points; *localized = d->localized_points;
  return (int)d->keys.size();
}
void orc_pyramid_localized(void* p, float* out /* localized x 8 */) {
  auto* d = (HessianAffineDetector*)p;
  for (size_t i = 0; i < d->localized.size(); i++) std::memcpy(out + i * 8, d->localized[i].data(), 8 * sizeof(float));
}
void orc_pyramid_free(void* p) { delete (HessianAffineDetector*)p; }

int orc_detect_orientation(const float* img, int w, int h, const double* kps, int n, double mrSize, int patchSize, int maxAngles,
                           double th, double* out, int max_out) {
  std::vector<Key> res = detectOrientation(keys_in(kps, n), image_in(img, w, h), mrSize, patchSize, maxAngles, th);
  for (size_t i = 0; i < res.size() && (int)i < max_out; i++) kp_out(res[i], out + i * KP);
  return (int)res.size();
}
int orc_detect_orientation_half(const float* img, int w, int h, const double* kps, int n, double mrSize, int patchSize, int maxAngles,
                                double th, int doHalfSIFT, double* out, int max_out) {
  std::vector<Key> res = detectOrientation(keys_in(kps, n), image_in(img, w, h), mrSize, patchSize, maxAngles, th, doHalfSIFT != 0);
  for (size_t i = 0; i < res.size() && (int)i < max_out; i++) kp_out(res[i], out + i * KP);
  return (int)res.size();
}
int orc_reproject(const double* kps, int n, const double* H, int w, int h, int which, double mrSize, double* out_det,
                  double* out_reproj) {
  std::vector<Key> det = keys_in(kps, n), rep;
  reprojectRegions(det, rep, H, w, h, which, mrSize);
  for (size_t i = 0; i < det.size(); i++) { kp_out(det[i], out_det + i * KP); kp_out(rep[i], out_reproj + i * KP); }
  return (int)det.size();
}
void orc_describe(const float* img, int w, int h, const double* kps, int n, double mrSize, int patchSize, int fast, int photoNorm,
                  int rootsift, float* desc, float* patches /* may be null */) {
  SIFTDescriptor D(patchSize, (rootsift & 1) != 0); D.doHalfSIFT = (rootsift & 2) != 0;   // flags: bit0 RootSIFT, bit1 Half descriptor
  describeRegions(keys_in(kps, n), image_in(img, w, h), D, mrSize, patchSize, fast != 0, photoNorm != 0, desc, patches);
}
void orc_describe_dsp(const float* img, int w, int h, const double* kps, int n, double mrSize, int patchSize, int fast, int photoNorm, int numScales,
                      double startCoef, double endCoef, float* desc) {
  describeRegionsDSP(keys_in(kps, n), image_in(img, w, h), mrSize, patchSize, fast != 0, photoNorm != 0, numScales, startCoef, endCoef, desc);
}
void orc_sift_patch(const float* patch41, int rootsift, float* desc128) {
  SIFTDescriptor D(41, (rootsift & 1) != 0); D.doHalfSIFT = (rootsift & 2) != 0;
  D(patch41, desc128);
}
// MSER: DetectMSERs (extrema.cpp:284) -- raw keys, or regions after DetectAffineRegions (synth-detection.hpp:93)
int orc_mser_detect(const float* img, int w, int h, double max_area, int min_size, double min_margin, int mode, int reg_number,
                    int raw, double* out, int max_out) {
  mser::Params mp; mp.max_area = max_area; mp.min_size = min_size; mp.min_margin = min_margin; mp.mode = mode; mp.reg_number = reg_number;
  std::vector<mser::MKey> mk = mser::detectMSERs(img, w, h, mp, 1.0, 1.0);
  std::vector<Key> keys;
  for (const mser::MKey& k : mk) keys.push_back(Key{k.x, k.y, k.a11, k.a12, k.a21, k.a22, k.s, k.response, k.sub_type});
  if (!raw) toRegions(keys);
  for (size_t i = 0; i < keys.size() && (int)i < max_out; i++) kp_out(keys[i], out + i * KP);
  return (int)keys.size();
}
// raw region list of getRLEExtrema (libExtrema.cpp:462): rows of 13 doubles
//   polarity minI maxI threshold margin area border nruns cx cy sxx sxy syy
int orc_mser_regions(const float* img, int w, int h, double max_area, int min_size, double min_margin, double* out, int max_out) {
  mser::Params mp; mp.max_area = max_area; mp.min_size = min_size; mp.min_margin = min_margin;
  std::vector<mser::OutRegion> regs;
  mser::detectMSERs(img, w, h, mp, 1.0, 1.0, &regs);
  for (size_t i = 0; i < regs.size() && (int)i < max_out; i++) {
    const mser::OutRegion& r = regs[i]; double* o = out + i * 13;
    o[0] = r.polarity; o[1] = r.minI; o[2] = r.maxI; o[3] = r.threshold; o[4] = r.margin; o[5] = r.area; o[6] = r.border;
    o[7] = (double)r.rle.size(); o[8] = r.cx; o[9] = r.cy; o[10] = r.sxx; o[11] = r.sxy; o[12] = r.syy;
  }
  return (int)regs.size();
}
int orc_view_pipeline(const float* img, int w, int h, int detector, const HessParamsC* hp, double mser_max_area, int mser_min_size, double mser_min_margin,
                      double ori_mrSize, int ori_patch, int maxAngles, double ori_th, double desc_mrSize, int desc_patch,
                      int photoNorm, int rootsift, double* det_out, double* reproj_out, float* desc_out, int max_out) {
  Image im = image_in(img, w, h);
  std::vector<Key> kp1;
  if (detector == 0) kp1 = detectAffineKeypoints(im, to_par(*hp), 1.0, 1.0);
  else {
    mser::Params mp; mp.max_area = mser_max_area; mp.min_size = mser_min_size; mp.min_margin = mser_min_margin;
    for (const mser::MKey& k : mser::detectMSERs(img, w, h, mp, 1.0, 1.0))
      kp1.push_back(Key{k.x, k.y, k.a11, k.a12, k.a21, k.a22, k.s, k.response, k.sub_type});
  }
  toRegions(kp1);
  std::vector<Key> det = detectOrientation(kp1, im, ori_mrSize, ori_patch, maxAngles, ori_th, (rootsift & 4) != 0), rep;
  const double H[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  reprojectRegions(det, rep, H, w, h, 0, 0);
  int n = (int)det.size();
  if (n > max_out) return -n;
  SIFTDescriptor D(desc_patch, (rootsift & 1) != 0); D.doHalfSIFT = (rootsift & 2) != 0;   // flags: bit0 RootSIFT, bit1 Half descriptor, bit2 orientation modulo pi
  describeRegions(det, im, D, desc_mrSize, desc_patch, false, photoNorm != 0, desc_out);
  for (int i = 0; i < n; i++) { kp_out(det[i], det_out + (size_t)i * KP); kp_out(rep[i], reproj_out + (size_t)i * KP); }
  return n;
}

int orc_synth_view(const float* img, int w, int h, double tilt, double phi, double zoom, double InitSigma, int doBlur,
                   float* out, int capacity, int* ow, int* oh, double* H) {
  SynthView v;
  generateSynthImage(image_in(img, w, h), tilt, phi, zoom, InitSigma, doBlur, v);
  *ow = v.pixels.cols; *oh = v.pixels.rows;
  for (int i = 0; i < 9; i++) H[i] = v.H[i];
  if ((size_t)*ow * *oh <= (size_t)capacity) std::memcpy(out, v.pixels.px.data(), sizeof(float) * (size_t)*ow * *oh);
  return v.identity ? 1 : 0;
}
int orc_view_pipeline_synth(const float* img, int w, int h, int detector, const HessParamsC* hp, double mser_max_area, int mser_min_size,
                            double mser_min_margin, double ori_mrSize, int ori_patch, int maxAngles, double ori_th, double desc_mrSize,
                            int desc_patch, int photoNorm, int rootsift, double tilt, double phi, double zoom, double InitSigma, int doBlur,
                            double* det_out, double* reproj_out, float* desc_out, int max_out) {
  SynthView v;
  generateSynthImage(image_in(img, w, h), tilt, phi, zoom, InitSigma, doBlur, v);
  const Image& im = v.pixels;
  std::vector<Key> kp1;
  if (detector == 0) kp1 = detectAffineKeypoints(im, to_par(*hp), v.tilt, v.zoom);
  else {
    mser::Params mp; mp.max_area = mser_max_area; mp.min_size = mser_min_size; mp.min_margin = mser_min_margin;
    for (const mser::MKey& k : mser::detectMSERs(im.px.data(), im.cols, im.rows, mp, v.tilt, v.zoom))
      kp1.push_back(Key{k.x, k.y, k.a11, k.a12, k.a21, k.a22, k.s, k.response, k.sub_type});
  }
  toRegions(kp1);
  std::vector<Key> det = detectOrientation(kp1, im, ori_mrSize, ori_patch, maxAngles, ori_th, (rootsift & 4) != 0), rep;
  reprojectRegions(det, rep, v.H, w, h, 0, 0);
  int n = (int)det.size();
  if (n > max_out) return -n;
  SIFTDescriptor D(desc_patch, (rootsift & 1) != 0); D.doHalfSIFT = (rootsift & 2) != 0;   // flags: bit0 RootSIFT, bit1 Half descriptor, bit2 orientation modulo pi
  describeRegions(det, im, D, desc_mrSize, desc_patch, false, photoNorm != 0, desc_out);
  for (int i = 0; i < n; i++) { kp_out(det[i], det_out + (size_t)i * KP); kp_out(rep[i], reproj_out + (size_t)i * KP); }
  return n;
}

// which: 0 HDs, 1 HDsSym, 2 HDsSymMax, 3 FDs, 4 FDsSym, 5 exFDsSym's residual
void orc_score(int which, const double* u, const double* M, double* d, int len) {
  switch (which) {
    case 0: HDs(u, M, d, len); break;
    case 1: HDsSymImpl(u, M, d, len, false); break;
    case 2: HDsSymImpl(u, M, d, len, true); break;
    case 3: FDsImpl(u, M, d, len, 0); break;
    case 4: FDsImpl(u, M, d, len, 1); break;
    default: FDsImpl(u, M, d, len, 2); break;
  }
}
// out rows of 7 doubles: q idx0 idxJ idx1 d0 dJ d1
int orc_match_fginn(const float* q, int nq, const float* t, int nt, const double* txy, double ratio, double contradDist, int nn,
                    double* out, int max_out) {
  std::vector<Tentative> m = matchFGINN(q, nq, t, nt, txy, ratio, contradDist, nn);
  for (size_t i = 0; i < m.size() && (int)i < max_out; i++) {
    double* o = out + i * 7;
    o[0] = m[i].q; o[1] = m[i].i0; o[2] = m[i].iJ; o[3] = m[i].i1; o[4] = m[i].d0; o[5] = m[i].dJ; o[6] = m[i].d1;
  }
  return (int)m.size();
}

// out rows of 6 doubles: q idx0 idx1 d0 d1 ratio
int orc_match_hamming(const float* q, int nq, const float* t, int nt, int dim, double matchDistanceThreshold, double* out, int max_out) {
  std::vector<HammingTentative> m = matchHamming(q, nq, t, nt, dim, matchDistanceThreshold);
  for (size_t i = 0; i < m.size() && (int)i < max_out; i++) {
    double* o = out + i * 6;
    o[0] = m[i].q; o[1] = m[i].i0; o[2] = m[i].i1; o[3] = m[i].d0; o[4] = m[i].d1; o[5] = m[i].ratio;
  }
  return (int)m.size();
}

}  // extern "C"
