// Shared records of the detection kernels (pyramid.cu, affine.cu, orient.cu, describe.cu).
#pragma once
#include "common.cuh"

struct BlurTaps { int n; float k[33]; };

struct Candidate { int r, c, level, pad; };

struct Localized {
  unsigned long long key;  // level*rows*cols + r0*cols + c0 : the reference's processing order inside an octave
  int valid, r, c, level;
  float b0, b1, b2, val;
};

// A localized scale-space point on its way through Baumberg (affine.cu).
struct KeypointRec {
  unsigned long long order;  // (octave << 56) | Localized::key  == detection order of the reference
  float x, y, s, pixelDistance, response;
  float a11, a12, a21, a22;
  float b2;                  // scale offset of the quadratic fit; s is filled in by the host leg (powf)
  int type, octave, level, ok;
};

#define MB2_MAX_LEVELS 8
struct OctaveLevels {
  ImgView blur[MB2_MAX_LEVELS];
  ImgView resp[MB2_MAX_LEVELS];
  int nlevels;
};

struct LocalizeParams {
  double edgeScoreThreshold;
  float finalThreshold;
  float pixelDistance;
  int numberOfScales;
  float levelSigma[MB2_MAX_LEVELS];
  int detectorType;   // 0 DET_HESSIAN, 1 DET_DOG, 2 DET_HARRIS (getPointType, pyramid.cpp:66-130)
};

// DET_DOG response (pyramid.cpp:176-181): resp = level - GaussianBlur(level, ksize(sigma), sigma, BORDER_REPLICATE); d_taps: n taps on
// the device, d_tmp: one plane of the level's size (row pass)
void mb2_launch_dog(mb2_ctx* ctx, const ImgView& level, float* resp, int resp_pitch, const float* d_taps, int n, float* d_tmp);
// DET_HARRIS response (pyramid.cpp:283-305); taps: Gaussian of sigma = sqrt(0.6 norm), sigmasq = (float)(0.6 norm); d_tmp6: six planes of the level's size
int mb2_launch_harris(mb2_ctx* ctx, const ImgView& level, float* resp, int resp_pitch, const BlurTaps& taps, float sigmasq, float* d_tmp6);
int mb2_launch_blur(mb2_ctx* ctx, const ImgView& src, float* dst_blur, float* dst_resp, int dst_pitch, const BlurTaps& taps,
                    float norm2, int want_resp);
// in-level part of the 3x3x3 extremum test, run by the kernel that produces a detection level (k_blur_hess_tma)
struct PrefilterArgs { int enable, border; float posThr, negThr; int level; Candidate* pre; int* pre_count; int pre_cap; };
int mb2_launch_blur_tma(mb2_ctx* ctx, const ImgView& src, float* dst_blur, float* dst_resp, int dst_pitch, const BlurTaps& taps, float norm2,
                        const PrefilterArgs& pf);
void mb2_launch_nms_finish(mb2_ctx* ctx, const OctaveLevels& oct, const Candidate* pre, const int* pre_count, int pre_cap, Candidate* out, int* count,
                           int capacity);
void mb2_launch_hessian(mb2_ctx* ctx, const ImgView& src, float* dst, int dst_pitch, float norm2);
void mb2_launch_resize_half(mb2_ctx* ctx, const ImgView& src, float* dst, int orows, int ocols, int dst_pitch);
void mb2_launch_nms(mb2_ctx* ctx, const ImgView& low, const ImgView& cur, const ImgView& high, int border, float posThr,
                    float negThr, int level, Candidate* out, int* count, int capacity);
void mb2_launch_fill_u64(mb2_ctx* ctx, unsigned long long* p, size_t n, unsigned long long v);
void mb2_launch_localize(mb2_ctx* ctx, const OctaveLevels& oct, Candidate* cand, int n, const LocalizeParams& lp,
                         unsigned long long* octmap, Localized* out);
void mb2_launch_emit(mb2_ctx* ctx, const OctaveLevels& oct, const Localized* loc, int n, const unsigned long long* octmap,
                     const LocalizeParams& lp, int octave, KeypointRec* out, int* count, int capacity);

struct ScaleReq { float b2; int level; int octave; };
void mb2_launch_scale_requests(mb2_ctx* ctx, const KeypointRec* kps, int n, ScaleReq* out);
void mb2_launch_set_scales(mb2_ctx* ctx, KeypointRec* kps, int n, const float* d_s);

// affine.cu
struct AffineParams {
  int maxIterations, smmWindowSize, doBaumberg;
  float convergenceThreshold, initialSigma;
};
void mb2_launch_affine_shape(mb2_ctx* ctx, const OctaveLevels* d_octaves, int n_octaves, KeypointRec* kps, int n,
                             const AffineParams& ap, const float* d_smm_mask);

// final keypoint record (doubles) as the C ABI returns it
struct KeyOut { double v[MB2_KP]; unsigned long long order; int keep; int pad; };
void mb2_launch_export(mb2_ctx* ctx, const KeypointRec* kps, int n, KeyOut* out, int as_regions);

// orient.cu
struct OrientParams { double mrSize; int patchSize; int maxAngles; double threshold; int half; };
void mb2_launch_orientation(mb2_ctx* ctx, const ImgView& img, const KeyOut* in, int n, const OrientParams& op,
                            const float* d_orimask, KeyOut* out /* n * maxAngles; .pad = float bits of the angle */, int* out_count_per_kp);
void mb2_launch_extract_angles(mb2_ctx* ctx, const KeyOut* keys, int n, float* d_ang);
void mb2_launch_apply_rotation(mb2_ctx* ctx, KeyOut* keys, const double* d_cs, int n);

// describe.cu
struct DescribeParams { double mrSize; int patchSize; int photoNorm; int rootSIFT; int fast; int half; int raw; };   // raw: leave the un-normalised votes (DSPSIFT)
void mb2_launch_dsp_accumulate(mb2_ctx* ctx, const double* d_vecT, int n, float* d_acc, int first);
void mb2_launch_dsp_norm(mb2_ctx* ctx, float* d_acc, int n, uint8_t* d_desc);
struct DescTables {      // precomputed on the host exactly as the reference's constructors do
  float mask[41 * 41];   // computeCircularGaussMask(41), siftdesc.h:87 / synth-detection.hpp:181
  int bin0[41], bin1[41];
  float w0[41], w1[41];  // floats stored in doubles by the reference (siftdesc.cpp:44-45)
};
int mb2_launch_describe(mb2_ctx* ctx, const ImgView& img, const KeyOut* kps, int n, const DescribeParams& dp,
                        const DescTables* d_tables, uint8_t* d_desc, float* d_patches /* may be null */);
// reprojection + boundary filter (synth-detection.cpp:541-616)
void mb2_launch_reproject(mb2_ctx* ctx, const KeyOut* det, int n, const double* Hinv9, int h_is_eye, int orig_w, int orig_h,
                          KeyOut* reproj /* keep flag set */);

// mser.cu: MSER+ / MSER- keys of the device image -> ctx->kp_b (reference order)
int mb2_mser_core(mb2_ctx* ctx, const ImgView& img, const mb2_mser_params& par, double tilt, double zoom, int as_regions, int* n_out,
                  double* d_table_out, int capacity);
int mb2_mser_core_pair(mb2_ctx* ctx, const ImgView& img1, const ImgView& img2, const mb2_mser_params& par, int as_regions, int* n1, int* n2);
void mb2_mser_release(mb2_ctx* ctx);

// synth.cu: GenerateSynthImageCorr on the device; returns 1 for the identity view (out = in), 0 otherwise, < 0 on error
int mb2_synth_core(mb2_ctx* ctx, const ImgView& in, const mb2_view_params& vp, ImgView* out, double* H);
