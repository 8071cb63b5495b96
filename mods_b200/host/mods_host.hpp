// Host-side mirror of the reference's plugin surface for the hot path, on top of the C ABI
// (include/mods_b200.h).  Same type / method names, argument meaning and error behaviour as the
// reference, so that mods.cpp's iteration loop (mods.cpp:229-415) reads the same:
//
//   ImageRepresentation::SynthDetectDescribeKeypoints   imagerepresentation.h:44-47, .cpp:603-2047
//   CorrespondenceBank::MatchImgReps                    correspondencebank.h:29-31, .cpp:237-351
//   MatchFlannFGINN                                     matching/matching.hpp:268-269, .cpp:357-461
//   MatchFLANNDistance                                  matching/matching.hpp:273, .cpp:607-666
//   DuplicateFiltering                                  matching/matching.hpp:300, .cpp:2983-3047
//   LORANSACFiltering (+ NaiveHCheck, H_LAF_check)      matching/matching.hpp:284-286, .cpp:806-980, 1171-1200, 251-309
//
// cv::Mat is replaced by a plain float image; everything else keeps the reference's layout
// (detectors/structures.hpp:187-245).  Only the HessianAffine and MSER detectors and the SIFT / RootSIFT
// descriptors are wired (the other branches are out of scope); unknown detector or
// descriptor names are skipped silently, exactly as the reference skips detectors that have no
// views configured.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "../../include/mods_b200.h"

namespace mods {

enum detector_type { DET_HESSIAN = 0, DET_DOG = 1, DET_HARRIS = 2, DET_MSER = 3, DET_UNKNOWN = 1000 };
enum descriptor_type { DESC_SIFT = 0, DESC_ROOT_SIFT = 1, DESC_HALF_SIFT = 2, DESC_HALF_ROOT_SIFT = 3, DESC_DSP_SIFT = 19, DESC_UNKNOWN = 1000 };   // detectors/structures.hpp:76-97
enum RANSAC_error_t { SAMPSON, SYMM_MAX, SYMM_SUM };
const int MODE_RANDOM = 0, MODE_FGINN = 1, MODE_DISTANCE = 2, MODE_BIGGER_REGION = 3;  // configuration.hpp:31-34

struct GrayImage {  // stands in for the CV_32F cv::Mat OriginalImg (gray = (B+G+R)/3, synth-detection.cpp:256-263)
  const float* data = nullptr;  // host or device pointer, row-major, cols floats per row
  int rows = 0, cols = 0;
};

struct AffineKeypoint {  // detectors/structures.hpp:187-196
  double x, y;
  double a11, a12, a21, a22;
  double s;
  double response;
  int octave_number;
  double pyramid_scale;
  int sub_type;
};
struct ViewSynthParameters {  // detectors/structures.hpp:198-211
  double zoom = 1.0, tilt = 1.0, phi = 0.0, InitSigma = 0.5;
  int doBlur = 1, DSPlevels = 0;
  double minSigma = 1.0, maxSigma = 1.0;
  std::vector<std::string> descriptors;
  std::map<std::string, double> FGINNThreshold;
  std::map<std::string, double> DistanceThreshold;
};
typedef std::map<std::string, std::vector<ViewSynthParameters> > IterationViewsynthesisParam;
struct Descriptor { descriptor_type type = DESC_UNKNOWN; std::vector<float> vec; };
struct AffineRegion {  // detectors/structures.hpp:219-230
  int img_id = 0, img_reproj_id = 0, id = 0, parent_id = 0;
  detector_type type = DET_UNKNOWN;
  AffineKeypoint det_kp, reproj_kp;
  Descriptor desc;
};
typedef std::vector<AffineRegion> AffineRegionVector;
typedef std::vector<AffineRegion> AffineRegionList;
typedef std::map<std::string, AffineRegionVector> AffineRegionVectorMap;

struct WhatToMatch {  // detectors/structures.hpp:247-253
  std::vector<std::string> group_detectors, group_descriptors, separate_detectors, separate_descriptors;
};
struct TimeLog {  // detectors/structures.hpp:51-74
  double SynthTime = 0, DetectTime = 0, OrientTime = 0, DescTime = 0, MatchingTime = 0, RANSACTime = 0, MiscTime = 0, TotalTime = 0;
};

struct DetectorsParameters { mb2_hessaff_params HessParam; mb2_mser_params MSERParam; DetectorsParameters(); };
struct DescriptorsParameters {
  mb2_sift_params SIFTParam, RootSIFTParam, HalfRootSIFTParam, HalfSIFTParam;
  int DSPScales = 3; double DSPStartCoef = 0.5, DSPEndCoef = 1.5;   // SIFTParam.DSPParam (DomainSizePolingParams, siftdesc.h:19-30)
  DescriptorsParameters();
};
struct DominantOrientationParams {  // descriptors_parameters.hpp:23-36 + [DominantOrientation]
  int maxAngles = 1; float threshold = 0.8f; bool addUpRight = false; bool halfSIFTMode = false;
  double mrSize = 1.0; int patchSize = 41;
};

struct TentativeCorrespExt {  // matching/matching.hpp:35-52.  Regions are carried WITHOUT their descriptor payload.
  AffineRegion first, second, secondbadby2ndcl, secondbad;
  double d1 = 0, d2 = 0, d2by2ndcl = 0, d2byDB = 0, ratio = 0;
  int isTrue = 0;
};
struct TentativeCorrespListExt {
  std::vector<TentativeCorrespExt> TCList;
  double H[9];
  TentativeCorrespListExt() { for (int i = 0; i < 9; i++) H[i] = -1; }
};
struct MatchPars {  // matching/matching.hpp:99-144 (the fields the float path reads)
  std::map<std::string, double> FGINNThreshold, DistanceThreshold;
  double currMatchRatio = -1.0, matchDistanceThreshold = 0, contradDist = 10.0;
  int minMatches = 15, maxSteps = 4;
};
struct RANSACPars {  // matching/matching.hpp:146-171
  int useF = 0;
  double err_threshold = 2.0, confidence = 0.99;
  int max_samples = 100000, localOptimization = 1;
  double LAFCoef = 3.0, HLAFCoef = 10.0;
  RANSAC_error_t errorType = SYMM_SUM;
  int doSymmCheck = 0, justMarkOutliers = 0;
  long seed = 0;  // 0 => time(NULL) like the reference (exp_ranH.c:823); anything else is used as the srand seed
};

// SetVSPars (synth-detection.cpp:103-234): expands ScaleSet x TiltSet x rotations (n = floor(180 t / Phi), delta = pi / n) into
// view parameters, drops the views already generated in earlier steps (prev_par, tolerance 0.01) and appends the new ones to it.
int SetVSPars(const std::vector<double>& scale_set, const std::vector<double>& tilt_set, const double phi_base,
              const std::vector<double>& FGINNThreshold, const std::vector<double>& DistanceThreshold,
              const std::vector<std::string> descriptors, std::vector<ViewSynthParameters>& par, std::vector<ViewSynthParameters>& prev_par,
              const double InitSigma = 0.5, const int doBlur = 1, const int dsplevels = 0, const double minSigma = 1.0, const double maxSigma = 1.0);

class CorrespondenceBank;

class ImageRepresentation {
 public:
  ImageRepresentation(mb2_ctx* ctx, GrayImage img, std::string name, int device_slot);
  descriptor_type GetDescriptorType(std::string desc_name) const;
  detector_type GetDetectorType(std::string det_name) const;
  TimeLog GetTimeSpent() const { return TimeSpent; }
  int GetRegionsNumber(std::string det_name = "All") const;
  int GetDescriptorsNumber(std::string desc_name = "All", std::string det_name = "All") const;
  AffineRegionVector GetAffineRegionVector(std::string desc_name, std::string det_name) const;
  void SynthDetectDescribeKeypoints(IterationViewsynthesisParam& synth_par, DetectorsParameters& det_par,
                                    DescriptorsParameters& desc_par, DominantOrientationParams& dom_ori_par);
  // Appends the n regions of the most recent view pass of context `from` (still resident there) as view `synth` of
  // (det, desc): what SynthDetectDescribeKeypoints does after every view, for passes the caller ran itself (the pair
  // driver batches the MSER detection of both images, mb2_mser_detect_pair).
  // creates the (det, desc) entries up front, so that passes of different detectors may then fill them from different threads
  void Prepare(const std::string& det, const std::string& desc) { Blocks[det][desc]; slot_state[det]; }
  // Feature cache, text format of the reference (imagerepresentation.cpp:2139-2215; record = saveAR, :89-99): files written here are
  // read by the reference's `read_pre_extracted` flow and by build/read_features.m, and vice versa.  Loaded regions live on the host
  // only (MatchImgReps then uploads them; the device-resident fast path needs regions detected in this process).
  void SaveRegions(std::string fname, int mode = 0) const;
  void LoadRegions(std::string fname);
  void AppendViewFrom(mb2_ctx* from, const std::string& det, const std::string& desc, int n, int synth);
  GrayImage OriginalImg;
  friend class CorrespondenceBank;

  // Regions live as the SoA blocks the C ABI returns; the AoS AffineRegionVector the reference keeps in
  // RegionVectorMap[det][desc] (imagerepresentation.h:66) is materialised on demand (GetAffineRegionVector).
  struct RegionBlock {
    int n = 0;
    detector_type det = DET_UNKNOWN; descriptor_type desc = DESC_UNKNOWN;
    std::vector<double> det_kp, reproj_kp;  // n x MB2_KP
    std::vector<uint8_t> desc_u8;           // n x 128
    std::vector<int> img_id;                // view index per region
    AffineRegion region(int i, bool with_desc) const;
  };
 public:
  const RegionBlock* block(const std::string& det, const std::string& desc) const {
    auto d = Blocks.find(det);
    if (d == Blocks.end()) return nullptr;
    auto e = d->second.find(desc);
    return e == d->second.end() ? nullptr : &e->second;
  }
 protected:
  mb2_ctx* ctx;
  TimeLog TimeSpent;
  std::map<std::string, std::map<std::string, RegionBlock> > Blocks;  // [det][desc]
  std::string Name;
  // device-resident copies of RegionVectorMap[det][desc] for the matcher: HessianAffine in `slot`, MSER in `slot + 2`
  int slot;
  // The class API keeps the reference's silent behaviour (a failing view yields empty lists, imagerepresentation.h:44-47 has no return
  // code); the first negative code of a view call is remembered here so that the C entry points (mb2_mods_pair / mb2_mods_pairs) can fail.
  int last_rc = 0;
 public:
  int LastError() const { return last_rc; }
  void ClearError() { last_rc = 0; }
  int SlotCount(const std::string& det) const { auto it = slot_state.find(det); return it == slot_state.end() ? 0 : it->second.count; }
 protected:
  struct SlotState { std::string desc; int count = 0; };   // which descriptor the slot currently holds ("" = none)
  std::map<std::string, SlotState> slot_state;              // per detector
  int slot_of(const std::string& det) const { return slot < 0 ? -1 : (det == "MSER" ? slot + 2 : slot); }
};

class CorrespondenceBank {
 public:
  explicit CorrespondenceBank(mb2_ctx* ctx) : ctx(ctx) {}
  int GetCorrespondencesNumber(std::string desc_name = "All", std::string det_name = "All") const;
  TentativeCorrespListExt GetCorresponcesVector(std::string desc_name = "All", std::string det_name = "All") const;
  // same content as GetCorresponcesVector("All","All") but hands the lists over instead of deep-copying them
  TentativeCorrespListExt TakeCorrespondences();
  int MatchImgReps(ImageRepresentation& imgrep1, ImageRepresentation& imgrep2, IterationViewsynthesisParam& synth_par,
                   const WhatToMatch WhatToMatchNow, const MatchPars& par, const DescriptorsParameters& desc_pars);
  void ClearCorrespondences(std::string det_name, std::string desc_name);

 protected:
  mb2_ctx* ctx;
  std::map<std::string, std::map<std::string, TentativeCorrespListExt> > CorrespondencesMapMap;  // [desc][det]
};

int MatchFlannFGINN(mb2_ctx* ctx, const AffineRegionList& list1, const AffineRegionList& list2, TentativeCorrespListExt& corresp,
                    const MatchPars& par, const int nn = 50);
int MatchFLANNDistance(mb2_ctx* ctx, const AffineRegionList& list1, const AffineRegionList& list2, TentativeCorrespListExt& corresp,
                       const MatchPars& par, const int nn = 50);   // matching.hpp:273, .cpp:607-666 (binary descriptors, Hamming 2-NN)
void DuplicateFiltering(TentativeCorrespListExt& in_corresp, const double r = 3.0, const int mode = MODE_RANDOM);
int LORANSACFiltering(mb2_ctx* ctx, TentativeCorrespListExt& in_corresp, TentativeCorrespListExt& out_corresp, double* H,
                      const RANSACPars pars);
// the same two steps on plain arrays (what the class API calls underneath)
std::vector<int> duplicate_filter_core(const double* xy, const double* key, int n, double r, bool sorted);
int loransac_core(mb2_ctx* ctx, const double* frames14, int n, const RANSACPars& pars, std::vector<unsigned char>& inl,
                  std::vector<int>& verified, double* H);

}  // namespace mods

// ---- C door for bench.py / tests: one MODS iteration on one image pair (mods.cpp:229-415, step 0 of
// an iters file that has a single HessianAffine view tier with the identity view) --------------------
extern "C" {
#define MB2_MAX_PAIR_VIEWS 64
typedef struct {
  mb2_hessaff_params det;
  mb2_orientation_params ori;
  mb2_sift_params desc;
  double matchRatio, contradDist, duplicateDist;
  double err_threshold, confidence, HLAFCoef;
  int max_samples, errorType, doSymmCheck;
  long seed;
  /* second detector of the step ([MSER0] / [MSER1] tiers of iters_mods_cviu.ini), matched separately like the reference's
   * "separate detectors" loop (correspondencebank.cpp:291-347) and verified together with the HessianAffine tentatives */
  int use_mser;
  mb2_mser_params mser;
  double mserMatchRatio;
  /* view tiers of the step (rows of SetVSPars; iters_mods_cviu.ini [HessianAffine4] has 11, [MSER2] 3): n == 0 means the identity view
   * only.  The views of one detector are appended in this order (imagerepresentation.cpp:2044-2045). */
  int n_hess_views, n_mser_views;
  mb2_view_params hess_views[MB2_MAX_PAIR_VIEWS], mser_views[MB2_MAX_PAIR_VIEWS];
  /* verification model (RANSACPars, matching.hpp:146-171): useF = 1 runs the epipolar branch of LORANSACFiltering
   * (exp_ransacFcustom + F_LAF_check) instead of the homography one */
  int useF, localOptimization;
  double LAFCoef;
  /* 1: every tier describes `RootSIFT,HalfRootSIFT` (the WxBS tiers): orientations modulo pi for both (the Half* name is sticky,
   * imagerepresentation.cpp:693-714), each (detector, descriptor) group matched separately, all tentatives verified together.
   * regions1/2 and mser_regions1/2 count the RootSIFT regions; tentatives / mser_tentatives sum over both descriptors. */
  int halfRootSIFT, reserved;
} mb2_pair_config;
typedef struct {
  int regions1, regions2, tentatives, unique_tentatives, ransac_inliers, verified;
  double H[9];
  double ms_detect_describe, ms_match, ms_duplicate, ms_ransac, ms_total;
  int mser_regions1, mser_regions2, mser_tentatives;   /* the MSER share of regions1 / regions2 / tentatives */
} mb2_pair_result;
void mb2_pair_config_default(mb2_pair_config* c);
/* images: gray f32 [H|D].  verified_out (optional): capacity rows of 4 doubles (x1 y1 x2 y2).  Returns verified count or < 0. */
int mb2_mods_pair(mb2_ctx* ctx, const float* img1, int w1, int h1, const float* img2, int w2, int h2, const mb2_pair_config* cfg,
                  mb2_pair_result* res, double* verified_out, int capacity);
/* n_pairs independent pairs, software-pipelined (verification of pair k overlaps detection of pair k+1);
 * res[k] (and verified_out[k] / capacity[k] when given) are exactly what mb2_mods_pair returns for pair k. */
int mb2_mods_pairs(mb2_ctx* ctx, int n_pairs, const float* const* img1, const int* w1, const int* h1, const float* const* img2,
                   const int* w2, const int* h2, const mb2_pair_config* cfg, mb2_pair_result* res, double* const* verified_out,
                   const int* capacity);
/* Other callers of the same path.  mb2_mods_multi = mods_multi.cpp:232-330 (1-to-N: the query image is described once);
 * mb2_extract_features = extract_features.cpp (describe + SaveRegions in the reference's text format); mb2_mods_pair_cached = the
 * read_pre_extracted flow of mods.cpp:224-240 (LoadRegions for both images, then match + verify). */
int mb2_mods_multi(mb2_ctx* ctx, const float* img1, int w1, int h1, int n, const float* const* imgs2, const int* w2, const int* h2,
                   const mb2_pair_config* cfg, mb2_pair_result* res, double* const* verified_out, const int* capacity);
int mb2_extract_features(mb2_ctx* ctx, const float* img, int w, int h, const mb2_pair_config* cfg, const char* fname);
int mb2_mods_pair_cached(mb2_ctx* ctx, const char* cache1, const char* cache2, const mb2_pair_config* cfg, mb2_pair_result* res,
                         double* verified_out, int capacity);
/* mods.cpp:298-415 on plain arrays: DuplicateFiltering(MODE_FGINN) + LORANSACFiltering of n tentatives.  frames: n rows of 14 doubles
 * (reproj_kp x y a11 a12 a21 a22 s of the first region, then of the second), key[n] = sqrt(d1/d2) ratios.  Used by the view-sharded
 * driver (mods_b200/sharding.py), where rank 0 verifies the tentatives gathered from all ranks.  Returns the verified count. */
int mb2_host_verify(mb2_ctx* ctx, const double* frames14, const double* key, int n, const mb2_pair_config* cfg, mb2_pair_result* res,
                    double* verified_out, int capacity);
/* ---- view-sharded pair over the ranks of one node (BASELINE config C4; mods_sharded.cpp) ----------------------------
 * mb2_dist_unique_id (rank 0) -> hand the 128 bytes to every rank -> mb2_dist_comm_create on every rank (its ncclComm_t).  NCCL is
 * bound at run time (libnccl.so.2).  mb2_views_sharded_pair: see mods_sharded.cpp; img1 / img2 are the same on every rank, the result
 * is complete on rank 0; digest (4 x u64, optional) is identical on every rank and for every world size. */
int mb2_dist_unique_id(unsigned char* id128);
int mb2_dist_comm_create(mb2_ctx* ctx, int rank, int world, const unsigned char* id128, void** comm);
void mb2_dist_comm_destroy(void* comm);
int mb2_views_sharded_pair(mb2_ctx* ctx, void* comm, int rank, int world, const float* img1, int w1, int h1, const float* img2, int w2, int h2,
                           const mb2_pair_config* cfg, mb2_pair_result* res, double* verified_out, int capacity, unsigned long long* digest, double* stats);
/* Dataset form: a list of independent pairs, views sharded over the ranks as above; pair k is VERIFIED by rank k % world on a helper thread
 * while all ranks go on with the views of pair k + 1 (after the tentative exchange every rank holds all records and rows of the pair).
 * res[k] is complete on every rank (one all-gather of the result records at the end); verified_out[k] on rank k % world only.
 * digest: 4 x n_pairs, stats: 8 x n_pairs, both optional. */
int mb2_views_sharded_pairs(mb2_ctx* ctx, void* comm, int rank, int world, int n_pairs, const float* const* img1, const int* w1, const int* h1,
                            const float* const* img2, const int* w2, const int* h2, const mb2_pair_config* cfg, mb2_pair_result* res,
                            double* const* verified_out, const int* capacity, unsigned long long* digest, double* stats);
/* the unit plan on its own (checked without a GPU): cost of a view, longest-processing-time-first owners, offsets in the gathered buffer */
double mb2_shard_view_cost(int w, int h, double tilt, double zoom);
void mb2_shard_assign(const double* costs, int n_units, int world, int* owner);
int mb2_shard_layout(const int* owner, const int* counts, int n_units, int world, int* src_off);
/* test doors: feature cache of one (detector, descriptor) set through ImageRepresentation::SaveRegions / LoadRegions */
int mb2_host_save_regions(const char* fname, const char* det, const char* desc, int n, const double* det_kp, const double* reproj_kp, const uint8_t* desc_u8);
int mb2_host_load_regions(const char* fname, const char* det, const char* desc, int capacity, double* det_kp, double* reproj_kp, uint8_t* desc_u8);
/* test door: SetVSPars for one step; out rows = (zoom, tilt, phi); prev (n_prev rows, same layout) = views of earlier steps */
int mb2_host_set_vs_pars(const double* scales, int n_scales, const double* tilts, int n_tilts, double phi_base, const double* prev, int n_prev,
                         double* out, int capacity);
/* kernels launched by ctx and by the helper contexts mb2_mods_pair(s) keep (second image, verification) */
long long mb2_mods_launch_count(mb2_ctx* ctx);
/* helper context `which` (0..5) of ctx: same device, its own stream; created on first use, released by mb2_mods_release */
mb2_ctx* mb2_mods_sibling(mb2_ctx* ctx, int which);
/* destroys the helper context of ctx (call before mb2_ctx_destroy(ctx)) */
void mb2_mods_release(mb2_ctx* ctx);
}
