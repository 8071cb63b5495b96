// TEST INFRASTRUCTURE (oracle) -- not part of the shipped product path.
//
// extern "C" doorway into the UNMODIFIED reference sources that oracle/build_ref.sh
// compiles in place from /root/reference into oracle/_ref/libmods_ref.so.  Nothing here
// restates reference arithmetic except the per-view glue of
// ImageRepresentation::SynthDetectDescribeKeypoints (imagerepresentation.cpp:717-720,
// 1254-1341), which cannot be compiled here (it pulls every detector family + OpenCV
// nonfree): ref_view_pipeline() calls the same reference functions in the same order.
//
// Keypoint record (KP = 9 doubles): x y a11 a12 a21 a22 s response sub_type.
#include <cstring>
#include <vector>
#include <opencv2/core/core.hpp>
#include "detectors/structures.hpp"
#include "detectors/helpers.h"
#include "detectors/affinedetectors/scale-space-detector.hpp"
#include "detectors/mser/extrema/extrema.h"
#include "detectors/mser/extrema/libExtrema.h"
#include "matching/siftdesc.h"
#include "synth-detection.hpp"

extern "C" {
// Prototypes restated from degensac/{Htools,Ftools,rtools,utools,exp_ranH}.h -- those headers
// #define one-letter macros (u1.., _f1..) that cannot be included into C++ next to other code.
typedef struct { unsigned I; double J; } Score;                                   // rtools.h:17-23
typedef void (*HDsPtr)(const double*, const double*, const double*, double*, int);  // Htools.h:1
typedef void (*HDsiPtr)(const double*, const double*, const double*, double*, int, int*, int);
typedef void (*HDsidxPtr)(const double*, const double*, const double*, double*, int, int*, int);
void lin_hg(const double* u, double* dst, const int* inl, int len);
void u2h(const double* u, const int* inl, int len, double* H, double* buffer);
void HDs(const double*, const double*, const double*, double*, int);
void HDsSym(const double*, const double*, const double*, double*, int);
void HDsSymMax(const double*, const double*, const double*, double*, int);
void HDsi(const double*, const double*, const double*, double*, int, int*, int);
void HDsiSym(const double*, const double*, const double*, double*, int, int*, int);
void HDsiSymMax(const double*, const double*, const double*, double*, int, int*, int);
void HDsidx(const double*, const double*, const double*, double*, int, int*, int);
void HDsSymidx(const double*, const double*, const double*, double*, int, int*, int);
void HDsSymidxMax(const double*, const double*, const double*, double*, int, int*, int);
void FDs(const double* u, const double* F, double* p, int len);
typedef void (*FDsPtr)(const double*, const double*, double*, int);                  // Fcustomdef.h:3
typedef void (*exFDsPtr)(const double*, const double*, double*, double*, int);      // Fcustomdef.h:4
void exFDs(const double* u, const double* F, double* p, double* w, int len);
void exFDsSym(const double* u, const double* F, double* p, double* w, int len);
int exp_ransacFcustom(double* u, int len, double th, double conf, int max_sam, double* F, unsigned char* inl, int* data_out, int do_lo,
                      unsigned inlLimit, double** resids, double* H_best, int* Ih, exFDsPtr EXFDS1, FDsPtr FDS1, int doSymCheck);   // exp_ranF.h:70-72
void FDsSym(const double* u, const double* F, double* p, int len);
int nsamples(int ninl, int ptNum, int samsiz, double conf);
Score exp_ransacHcustom(double* u, int len, double th, double conf, int max_sam, double* H, unsigned char* inl,
                        int iter_type, int* data_out, int oriented_constraint, unsigned inlLimit, double** resids,
                        HDsPtr HDS1, HDsiPtr HDSi1, HDsidxPtr HDSidx1, int doSymCheck);   // exp_ranH.h:32
int svduv(double* d, double* a, double* u, int m, double* v, int n);                                    // matutls/svduv.c
void u2f(const double* u, const int* inl, int len, double* F, double* buffer);                          // Ftools.c:311
void u2fw(const double* u, const int* inl, const double* w, int len, double* F, double* buffer);        // Ftools.c:363
int checksample(double* F, double* u7, double th, double* H);                                           // DegUtils.c:42
long mb2_ref_seed = 1;
}

namespace {
const int KP = 9;

struct HessParamsC {       // mirrors the [HessianAffine] section of config_iter_mods_cviu.ini
  float threshold; int numberOfScales; float initialSigma; float edgeEigenValueRatio; int border;
  int maxIterations; float convergenceThreshold; int smmWindowSize; int doBaumberg;
  int mode; int reg_number; float rel_threshold; float rel_reg_number; int patchSize; float mrSize;
  int detectorType;   // 0 DET_HESSIAN, 1 DET_DOG, 2 DET_HARRIS
};

ScaleSpaceDetectorParams to_ref(const HessParamsC& p) {
  ScaleSpaceDetectorParams sp;
  sp.PyramidPars.threshold = p.threshold;
  sp.PyramidPars.numberOfScales = p.numberOfScales;
  sp.PyramidPars.initialSigma = p.initialSigma;
  sp.PyramidPars.edgeEigenValueRatio = p.edgeEigenValueRatio;
  sp.PyramidPars.border = p.border;
  sp.PyramidPars.DetectorMode = (detection_mode_t)p.mode;
  sp.PyramidPars.reg_number = p.reg_number;
  sp.PyramidPars.rel_threshold = p.rel_threshold;
  sp.PyramidPars.rel_reg_number = p.rel_reg_number;
  sp.PyramidPars.DetectorType = p.detectorType == 1 ? DET_DOG : p.detectorType == 2 ? DET_HARRIS : DET_HESSIAN;
  sp.AffineShapePars.maxIterations = p.maxIterations;
  sp.AffineShapePars.convergenceThreshold = p.convergenceThreshold;
  sp.AffineShapePars.smmWindowSize = p.smmWindowSize;
  sp.AffineShapePars.doBaumberg = p.doBaumberg;
  sp.AffineShapePars.patchSize = p.patchSize;
  sp.AffineShapePars.initialSigma = p.initialSigma;
  sp.AffineShapePars.mrSize = p.mrSize;
  return sp;
}

void kp_out(const AffineKeypoint& k, double* o) {
  o[0] = k.x; o[1] = k.y; o[2] = k.a11; o[3] = k.a12; o[4] = k.a21; o[5] = k.a22;
  o[6] = k.s; o[7] = k.response; o[8] = k.sub_type;
}
void kp_in(const double* o, AffineKeypoint& k) {
  k.x = o[0]; k.y = o[1]; k.a11 = o[2]; k.a12 = o[3]; k.a21 = o[4]; k.a22 = o[5];
  k.s = o[6]; k.response = o[7]; k.sub_type = (int)o[8]; k.octave_number = 0; k.pyramid_scale = 0;
}
void identity_view(SynthImage& v, const float* img, int w, int h) {
  v.id = 0; v.tilt = 1.0; v.rotation = 0.0; v.zoom = 1.0;
  for (int i = 0; i < 9; i++) v.H[i] = (i % 4 == 0) ? 1.0 : 0.0;
  v.pixels = cv::Mat(h, w, CV_32FC1);
  std::memcpy(v.pixels.data, img, sizeof(float) * (size_t)w * h);
}
AffineRegionList regions_in(const double* kps, int n) {
  AffineRegionList l(n);
  for (int i = 0; i < n; i++) {
    l[i].img_id = 0; l[i].img_reproj_id = 0; l[i].id = i; l[i].parent_id = 0; l[i].type = DET_HESSIAN;
    kp_in(kps + (size_t)i * KP, l[i].det_kp);
    l[i].reproj_kp = l[i].det_kp;
  }
  return l;
}

struct ExposedDetector : public ScaleSpaceDetector {
  ExposedDetector(const PyramidParams& p) : ScaleSpaceDetector(p) {}
  cv::Mat hessian(const cv::Mat& in, float norm) { return HessianResponse(in, norm); }
};
}  // namespace

extern "C" {

// ---- image helpers -------------------------------------------------------------------------
void ref_gaussian_blur(const float* src, float* dst, int w, int h, float sigma) {
  cv::Mat in(h, w, CV_32FC1, (void*)src);
  cv::Mat out = gaussianBlur(in, sigma);                         // helpers.cpp:717
  std::memcpy(dst, out.data, sizeof(float) * (size_t)w * h);
}
void ref_hessian_response(const float* src, float* dst, int w, int h, float norm) {
  PyramidParams p;
  ExposedDetector d(p);
  cv::Mat in(h, w, CV_32FC1, (void*)src);
  cv::Mat out = d.hessian(in, norm);                             // pyramid.cpp:223
  std::memcpy(dst, out.data, sizeof(float) * (size_t)w * h);     // NB: 1-px frame is uninitialised
}
float ref_atan2LUTff(float y, float x) { return atan2LUTff(y, x); }   // helpers.cpp:160
int ref_interpolate(const float* img, int w, int h, float ox, float oy, float a11, float a12, float a21,
                    float a22, float* out, int ow, int oh) {
  cv::Mat in(h, w, CV_32FC1, (void*)img);
  cv::Mat res(oh, ow, CV_32FC1, (void*)out);
  return interpolate(in, ox, oy, a11, a12, a21, a22, res) ? 1 : 0;   // helpers.cpp:551
}

// ---- Hessian-Affine ------------------------------------------------------------------------
// raw=1: DetectAffineKeypoints output (scale-space-detector.cpp:43); raw=0: after
// DetectAffineRegions (synth-detection.hpp:93-126: s*=sqrt|det A|, rectifyTransformation).
int ref_hessaff_detect(const float* img, int w, int h, const HessParamsC* hp, int raw, double* out, int max_out) {
  ScaleSpaceDetectorParams sp = to_ref(*hp);
  SynthImage view; identity_view(view, img, w, h);
  int n = 0;
  if (raw) {
    std::vector<AffineKeypoint> keys;
    ScalePyramid pyr;
    DetectAffineKeypoints(view.pixels, keys, sp, pyr, 1.0, 1.0);
    n = (int)keys.size();
    for (int i = 0; i < n && i < max_out; i++) kp_out(keys[i], out + (size_t)i * KP);
  } else {
    AffineRegionList regs;
    DetectAffineRegions(view, regs, sp, DET_HESSIAN, DetectAffineKeypoints);
    n = (int)regs.size();
    for (int i = 0; i < n && i < max_out; i++) kp_out(regs[i].det_kp, out + (size_t)i * KP);
  }
  return n;
}

// ---- MSER ----------------------------------------------------------------------------------
int ref_mser_detect(const float* img, int w, int h, double max_area, int min_size, double min_margin,
                    int mode, int reg_number, int raw, double* out, int max_out) {
  extrema::ExtremaParams ep;
  ep.max_area = max_area; ep.min_size = min_size; ep.min_margin = min_margin;
  ep.DetectorMode = (detection_mode_t)mode; ep.reg_number = reg_number;
  SynthImage view; identity_view(view, img, w, h);
  int n = 0;
  if (raw) {
    std::vector<AffineKeypoint> keys;
    ScalePyramid pyr;
    DetectMSERs(view.pixels, keys, ep, pyr, 1.0, 1.0);           // extrema.cpp:284
    n = (int)keys.size();
    for (int i = 0; i < n && i < max_out; i++) kp_out(keys[i], out + (size_t)i * KP);
  } else {
    AffineRegionList regs;
    DetectAffineRegions(view, regs, ep, DET_MSER, DetectMSERs);
    n = (int)regs.size();
    for (int i = 0; i < n && i < max_out; i++) kp_out(regs[i].det_kp, out + (size_t)i * KP);
  }
  return n;
}

// raw region list of getRLEExtrema (libExtrema.cpp:462) as DetectMSERs calls it (extrema.cpp:405): rows of 13 doubles
//   polarity minI maxI threshold margin area border nruns cx cy sxx sxy syy
int ref_mser_regions(const float* img, int w, int h, double max_area, int min_size, double min_margin, double* out, int max_out) {
  extrema::ExtremaParams ep;
  ep.max_area = max_area; ep.min_size = min_size; ep.min_margin = min_margin;
  extrema::ExtremaImage im;
  im.height = h; im.width = w; im.channels = 1;
  std::vector<unsigned char> px((size_t)w * h);
  for (size_t i = 0; i < px.size(); i++) px[i] = (unsigned char)img[i];    // extrema.cpp:401-403
  im.data = px.data();
  extrema::RLEExtrema res = extrema::getRLEExtrema(ep, im);
  int n = 0;
  for (int pol = 0; pol < 2; pol++) {
    const std::vector<extrema::RLERegion>& v = pol ? res.MSERmin : res.MSERplus;
    for (size_t i = 0; i < v.size(); i++, n++) {
      if (n >= max_out) continue;
      const extrema::RLERegion& r = v[i]; double* o = out + (size_t)n * 13;
      o[0] = pol; o[1] = r.minI; o[2] = r.maxI; o[3] = r.threshold; o[4] = r.margin; o[5] = r.area; o[6] = r.border;
      o[7] = (double)r.rle.size(); o[8] = r.cx; o[9] = r.cy; o[10] = r.sxx; o[11] = r.sxy; o[12] = r.syy;
    }
  }
  return n;
}

// ---- orientation / reprojection / description -----------------------------------------------
int ref_detect_orientation(const float* img, int w, int h, const double* kps, int n, double mrSize,
                           int patchSize, int maxAngles, double th, double* out, int max_out) {
  SynthImage view; identity_view(view, img, w, h);
  AffineRegionList in = regions_in(kps, n), res;
  DetectOrientation(in, res, view, mrSize, patchSize, 0, maxAngles, th, false);   // synth-detection.cpp:841
  int m = (int)res.size();
  for (int i = 0; i < m && i < max_out; i++) kp_out(res[i].det_kp, out + (size_t)i * KP);
  return m;
}
int ref_detect_orientation_half(const float* img, int w, int h, const double* kps, int n, double mrSize,
                                int patchSize, int maxAngles, double th, int doHalfSIFT, double* out, int max_out) {
  SynthImage view; identity_view(view, img, w, h);
  AffineRegionList in = regions_in(kps, n), res;
  DetectOrientation(in, res, view, mrSize, patchSize, doHalfSIFT, maxAngles, th, false);   // synth-detection.cpp:841
  int m = (int)res.size();
  for (int i = 0; i < m && i < max_out; i++) kp_out(res[i].det_kp, out + (size_t)i * KP);
  return m;
}
// which=0: ReprojectRegions (synth-detection.cpp:541); which=1: ...AndRemoveTouchBoundary (:63)
int ref_reproject(const double* kps, int n, const double* H, int w, int h, int which, double mrSize,
                  double* out_det, double* out_reproj) {
  AffineRegionList l = regions_in(kps, n);
  double Hc[9]; std::memcpy(Hc, H, sizeof(Hc));
  if (which == 0) ReprojectRegions(l, Hc, w, h);
  else ReprojectRegionsAndRemoveTouchBoundary(l, Hc, w, h, mrSize);
  for (size_t i = 0; i < l.size(); i++) {
    kp_out(l[i].det_kp, out_det + i * KP);
    kp_out(l[i].reproj_kp, out_reproj + i * KP);
  }
  return (int)l.size();
}
void ref_describe(const float* img, int w, int h, const double* kps, int n, double mrSize, int patchSize,
                  int fast, int photoNorm, int rootsift, float* desc /* n x 128 */) {
  SynthImage view; identity_view(view, img, w, h);
  AffineRegionList l = regions_in(kps, n);
  SIFTDescriptorParams sp;
  sp.useRootSIFT = rootsift & 1; sp.doHalfSIFT = (rootsift & 2) ? 1 : 0; sp.PEParam.patchSize = patchSize; sp.PEParam.mrSize = mrSize;
  SIFTDescriptor D(sp);
  DescribeRegions(l, view, D, mrSize, patchSize, fast != 0, photoNorm != 0);      // synth-detection.hpp:169
  for (int i = 0; i < n; i++)
    for (int j = 0; j < 128; j++) desc[(size_t)i * 128 + j] = j < (int)l[i].desc.vec.size() ? l[i].desc.vec[j] : 0.f;
}
// The "DSPSIFT" branch of SynthDetectDescribeKeypoints (imagerepresentation.cpp:1547-1598), which cannot be compiled here (it pulls
// every descriptor family): the same reference calls in the same order -- DescribeRegions with an un-normalised plain-SIFT functor at
// numScales + 1 measurement-region sizes, float sums, SIFTnorm(std::vector<float>&) of a functor with doNorm on.
void ref_describe_dsp(const float* img, int w, int h, const double* kps, int n, double mrSize, int patchSize, int fast, int photoNorm, int numScales,
                      double startCoef, double endCoef, float* desc /* n x 128 */) {
  SynthImage view; identity_view(view, img, w, h);
  AffineRegionList temp_kp1_desc = regions_in(kps, n), dsp_desc;
  SIFTDescriptorParams dspsiftparams;
  dspsiftparams.PEParam.patchSize = patchSize; dspsiftparams.PEParam.mrSize = mrSize;
  dspsiftparams.useRootSIFT = false;
  dspsiftparams.doNorm = false;
  SIFTDescriptor DSPSIFTdesc(dspsiftparams);
  const int num_domains = numScales;
  for (int dsp_idx = 0; dsp_idx < num_domains + 1; dsp_idx++) {
    dsp_desc = temp_kp1_desc;
    const double curr_mrSize = mrSize * (startCoef + dsp_idx * (endCoef - startCoef) / num_domains);
    DescribeRegions(dsp_desc, view, DSPSIFTdesc, curr_mrSize, patchSize, fast != 0, photoNorm != 0);
    for (size_t kp_idx = 0; kp_idx < dsp_desc.size(); kp_idx++) {
      const int desc_dim = (int)dsp_desc[kp_idx].desc.vec.size();
      temp_kp1_desc[kp_idx].desc.vec.resize(desc_dim);
      for (int e = 0; e < desc_dim; e++) {
        if (dsp_idx == 0) temp_kp1_desc[kp_idx].desc.vec[e] = dsp_desc[kp_idx].desc.vec[e];
        else temp_kp1_desc[kp_idx].desc.vec[e] += dsp_desc[kp_idx].desc.vec[e];
      }
    }
  }
  dspsiftparams.doNorm = true;
  SIFTDescriptor DSPSIFTdesc1(dspsiftparams);
  for (size_t kp_idx = 0; kp_idx < temp_kp1_desc.size(); kp_idx++) DSPSIFTdesc1.SIFTnorm(temp_kp1_desc[kp_idx].desc.vec);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < 128; j++) desc[(size_t)i * 128 + j] = j < (int)temp_kp1_desc[i].desc.vec.size() ? temp_kp1_desc[i].desc.vec[j] : 0.f;
}
void ref_sift_patch(const float* patch41, int rootsift, float* desc128) {
  SIFTDescriptorParams sp; sp.useRootSIFT = rootsift & 1; sp.doHalfSIFT = (rootsift & 2) ? 1 : 0;
  SIFTDescriptor D(sp);
  cv::Mat p(41, 41, CV_32FC1);
  std::memcpy(p.data, patch41, sizeof(float) * 41 * 41);
  std::vector<float> v;
  D(p, v);                                                                         // siftdesc.cpp:401
  for (int j = 0; j < 128; j++) desc128[j] = j < (int)v.size() ? v[j] : 0.f;
}

// One (detector, view) pass of SynthDetectDescribeKeypoints for the identity view with a
// SIFT-like descriptor (imagerepresentation.cpp:717-720 | 1035-1038, 1254-1341).
// detector: 0 HessianAffine, 3 MSER.  Returns number of described regions.
int ref_view_pipeline(const float* img, int w, int h, int detector, const HessParamsC* hp,
                      double mser_max_area, int mser_min_size, double mser_min_margin,
                      double ori_mrSize, int ori_patch, int maxAngles, double ori_th,
                      double desc_mrSize, int desc_patch, int photoNorm, int rootsift,
                      double* det_out, double* reproj_out, float* desc_out, int max_out) {
  SynthImage view; identity_view(view, img, w, h);
  AffineRegionList kp1;
  if (detector == 0) {
    ScaleSpaceDetectorParams sp = to_ref(*hp);
    DetectAffineRegions(view, kp1, sp, DET_HESSIAN, DetectAffineKeypoints);
  } else {
    extrema::ExtremaParams ep;
    ep.max_area = mser_max_area; ep.min_size = mser_min_size; ep.min_margin = mser_min_margin;
    DetectAffineRegions(view, kp1, ep, DET_MSER, DetectMSERs);
  }
  AffineRegionList oriented;
  DetectOrientation(kp1, oriented, view, ori_mrSize, ori_patch, (rootsift & 4) ? 1 : 0, maxAngles, ori_th, false);
  AffineRegionList desc_list = oriented;
  ReprojectRegions(desc_list, view.H, w, h);
  SIFTDescriptorParams sp;
  sp.useRootSIFT = rootsift & 1; sp.doHalfSIFT = (rootsift & 2) ? 1 : 0; sp.PEParam.patchSize = desc_patch; sp.PEParam.mrSize = desc_mrSize;
  SIFTDescriptor D(sp);
  DescribeRegions(desc_list, view, D, desc_mrSize, desc_patch, false, photoNorm != 0);
  int n = (int)desc_list.size();
  for (int i = 0; i < n && i < max_out; i++) {
    kp_out(desc_list[i].det_kp, det_out + (size_t)i * KP);
    kp_out(desc_list[i].reproj_kp, reproj_out + (size_t)i * KP);
    for (int j = 0; j < 128; j++) desc_out[(size_t)i * 128 + j] = j < (int)desc_list[i].desc.vec.size() ? desc_list[i].desc.vec[j] : 0.f;
  }
  return n;
}

// ---- view synthesis (synth-detection.cpp:236-430) ----------------------------------------------------
// out: pixels of the synthesised view (capacity floats), its size and H (original -> view).  Returns 1 for the identity view.
int ref_synth_view(const float* img, int w, int h, double tilt, double phi, double zoom, double InitSigma, int doBlur,
                   float* out, int capacity, int* ow, int* oh, double* H) {
  cv::Mat in(h, w, CV_32FC1);
  std::memcpy(in.data, img, sizeof(float) * (size_t)w * h);
  SynthImage view;
  GenerateSynthImageCorr(in, view, "img", tilt, phi, zoom, InitSigma, doBlur, 1, false);
  *ow = view.pixels.cols; *oh = view.pixels.rows;
  for (int i = 0; i < 9; i++) H[i] = view.H[i];
  if ((size_t)*ow * *oh <= (size_t)capacity)
    for (int r = 0; r < *oh; r++) std::memcpy(out + (size_t)r * *ow, view.pixels.ptr<float>(r), sizeof(float) * *ow);
  return view.id == 0 ? 1 : 0;
}
// One (detector, view) pass of SynthDetectDescribeKeypoints for a synthesised view (imagerepresentation.cpp:621-1341)
int ref_view_pipeline_synth(const float* img, int w, int h, int detector, const HessParamsC* hp,
                            double mser_max_area, int mser_min_size, double mser_min_margin,
                            double ori_mrSize, int ori_patch, int maxAngles, double ori_th,
                            double desc_mrSize, int desc_patch, int photoNorm, int rootsift,
                            double tilt, double phi, double zoom, double InitSigma, int doBlur,
                            double* det_out, double* reproj_out, float* desc_out, int max_out) {
  cv::Mat in(h, w, CV_32FC1);
  std::memcpy(in.data, img, sizeof(float) * (size_t)w * h);
  SynthImage view;
  GenerateSynthImageCorr(in, view, "img", tilt, phi, zoom, InitSigma, doBlur, 1, false);
  AffineRegionList kp1;
  if (detector == 0) {
    ScaleSpaceDetectorParams sp = to_ref(*hp);
    DetectAffineRegions(view, kp1, sp, DET_HESSIAN, DetectAffineKeypoints);
  } else {
    extrema::ExtremaParams ep;
    ep.max_area = mser_max_area; ep.min_size = mser_min_size; ep.min_margin = mser_min_margin;
    DetectAffineRegions(view, kp1, ep, DET_MSER, DetectMSERs);
  }
  AffineRegionList oriented;
  DetectOrientation(kp1, oriented, view, ori_mrSize, ori_patch, (rootsift & 4) ? 1 : 0, maxAngles, ori_th, false);
  AffineRegionList desc_list = oriented;
  ReprojectRegions(desc_list, view.H, w, h);
  SIFTDescriptorParams sp;
  sp.useRootSIFT = rootsift & 1; sp.doHalfSIFT = (rootsift & 2) ? 1 : 0; sp.PEParam.patchSize = desc_patch; sp.PEParam.mrSize = desc_mrSize;
  SIFTDescriptor D(sp);
  DescribeRegions(desc_list, view, D, desc_mrSize, desc_patch, false, photoNorm != 0);
  int n = (int)desc_list.size();
  for (int i = 0; i < n && i < max_out; i++) {
    kp_out(desc_list[i].det_kp, det_out + (size_t)i * KP);
    kp_out(desc_list[i].reproj_kp, reproj_out + (size_t)i * KP);
    for (int j = 0; j < 128; j++) desc_out[(size_t)i * 128 + j] = j < (int)desc_list[i].desc.vec.size() ? desc_list[i].desc.vec[j] : 0.f;
  }
  return n;
}

// ---- DEGENSAC ------------------------------------------------------------------------------
void ref_set_seed(long s) { mb2_ref_seed = s; }
void ref_lin_hg(const double* u, double* Z, int len) {
  std::vector<int> pool(len);
  for (int i = 0; i < len; i++) pool[i] = i;
  lin_hg(u, Z, pool.data(), len);                                  // Htools.c:17
}
// which: 0 HDs (Htools.c:158), 1 HDsSym (:199), 2 HDsSymMax (:241), 3 FDs (Ftools.c:82), 4 FDsSym (:102)
void ref_score(int which, const double* u, const double* M, double* d, int len) {
  std::vector<double> Z;
  if (which < 3) { Z.resize((size_t)len * 18); ref_lin_hg(u, Z.data(), len); }
  switch (which) {
    case 0: HDs(Z.data(), u, M, d, len); break;
    case 1: HDsSym(Z.data(), u, M, d, len); break;
    case 2: HDsSymMax(Z.data(), u, M, d, len); break;
    case 3: FDs(u, M, d, len); break;
    default: FDsSym(u, M, d, len); break;
  }
}
void ref_u2h(const double* u, const int* inl, int len, double* H) { u2h(u, inl, len, H, 0); }   // Htools.c:98
int ref_nsamples(int ninl, int ptNum, int samsiz, double conf) { return nsamples(ninl, ptNum, samsiz, conf); }
// exp_ransacHcustom (exp_ranH.c:796) exactly as LORANSACFiltering calls it (matching.cpp:891);
// srand(time(NULL)) sees mb2_ref_seed through the build recipe's time() hook.
// errorType: 0 Sampson, 1 SymmMax, 2 SymmSum.  out: I, samples, LO count, rejected.
double ref_exp_ransacH(const double* u, int len, double th, double conf, int max_sam, int errorType,
                       int doSymCheck, long seed, double* H, unsigned char* inl, int* out4) {
  mb2_ref_seed = seed;
  std::vector<double> uc(u, u + (size_t)len * 6);
  std::vector<int> data_out((size_t)len * 18 + 8, 0);
  double* resids = 0;
  HDsPtr a = errorType == 0 ? &HDs : errorType == 1 ? &HDsSymMax : &HDsSym;
  HDsiPtr b = errorType == 0 ? &HDsi : errorType == 1 ? &HDsiSymMax : &HDsiSym;
  HDsidxPtr c = errorType == 0 ? &HDsidx : errorType == 1 ? &HDsSymidxMax : &HDsSymidx;
  Score s = exp_ransacHcustom(uc.data(), len, th, conf, max_sam, H, inl, 4, data_out.data(), 1, 0, &resids,
                              a, b, c, doSymCheck);
  free(resids);
  out4[0] = (int)s.I; out4[1] = data_out[0]; out4[2] = data_out[1]; out4[3] = data_out[2];
  return s.J;
}

// exp_ransacFcustom (exp_ranF.c:795) exactly as LORANSACFiltering calls it in F mode (matching.cpp:883): 7-point LO-RANSAC with the
// DEGENSAC plane-and-parallax branch; srand(time(NULL)) sees mb2_ref_seed.  errorType: 0 Sampson, otherwise symmetric epipolar.
// out4: I, samples, LO count, Ih (inliers of the best homography met on the way).
// inlLimit: LORANSACFiltering passes 0 (every LSQ of the LO then runs on 8 random inliers); ref_exp_ransacF keeps the "no limit"
// form (inlLimit = len) the first golden vectors were made with.
int ref_exp_ransacF2(const double* u, int len, double th, double conf, int max_sam, int errorType, int doSymCheck, long seed, unsigned inlLimit,
                     double* F, unsigned char* inl, int* out4);
int ref_exp_ransacF3(const double* u, int len, double th, double conf, int max_sam, int errorType, int doSymCheck, long seed, unsigned inlLimit,
                     int do_lo, double* F, unsigned char* inl, int* out4);   // do_lo = pars.localOptimization (matching.cpp:808)
int ref_exp_ransacF(const double* u, int len, double th, double conf, int max_sam, int errorType, int doSymCheck, long seed, double* F,
                    unsigned char* inl, int* out4) {
  return ref_exp_ransacF2(u, len, th, conf, max_sam, errorType, doSymCheck, seed, (unsigned)len, F, inl, out4);
}
int ref_exp_ransacF3(const double* u, int len, double th, double conf, int max_sam, int errorType, int doSymCheck, long seed, unsigned inlLimit,
                     int do_lo, double* F, unsigned char* inl, int* out4) {
  mb2_ref_seed = seed;
  std::vector<double> uc(u, u + (size_t)len * 6);
  std::vector<int> data_out((size_t)len * 18 + 8, 0);
  double* resids = 0; double Hbest[9] = {0}; int Ih = 0;
  FDsPtr a = errorType == 0 ? &FDs : &FDsSym;
  exFDsPtr b = errorType == 0 ? &exFDs : &exFDsSym;
  const int I = exp_ransacFcustom(uc.data(), len, th, conf, max_sam, F, inl, data_out.data(), do_lo, inlLimit, &resids, Hbest, &Ih, b, a, doSymCheck);
  free(resids);
  out4[0] = I; out4[1] = data_out[0]; out4[2] = data_out[1]; out4[3] = Ih;
  return I;
}
int ref_exp_ransacF2(const double* u, int len, double th, double conf, int max_sam, int errorType, int doSymCheck, long seed, unsigned inlLimit,
                     double* F, unsigned char* inl, int* out4) {
  return ref_exp_ransacF3(u, len, th, conf, max_sam, errorType, doSymCheck, seed, inlLimit, 1, F, inl, out4);
}
// pieces of the F path on their own (unit comparisons of the restated numerics)
void ref_svduv3_V(const double* A, double* V) { double a[9], d[3], U[9]; std::memcpy(a, A, sizeof a); svduv(d, a, U, 3, V, 3); }
void ref_u2f(const double* u, const int* inl, const double* w, int len, double* F) {
  std::vector<double> buf((size_t)9 * (len < 9 ? 9 : len));
  if (w) u2fw(u, inl, w, len, F, buf.data()); else u2f(u, inl, len, F, buf.data());
}
int ref_checksample(const double* F, const double* u7, double th, double* H) {
  double f[9], uu[42]; std::memcpy(f, F, sizeof f); std::memcpy(uu, u7, sizeof uu);
  return checksample(f, uu, th, H);
}

}  // extern "C"
