/* mods_b200 -- C ABI of the B200-native MODS feature pipeline (libmods_b200.so).
 *
 * Plain pointers and sizes only.  Every entry point names the reference interface it stands in
 * for (paths relative to the ducha-aiki/mods tree).  Conventions kept from the reference
 * (SURVEY.md 8b): integer status/count returns, nothing is thrown across the ABI, an empty
 * output on failure.  Negative return = error (mb2_last_error() has the text).
 *
 * All `const float* / const uint8_t* / double*` data arguments marked [H|D] may be host or
 * device pointers (decided with cudaPointerGetAttributes); host buffers are staged through
 * pinned memory on the context's stream.
 *
 * Keypoint record: MB2_KP = 9 doubles  { x, y, a11, a12, a21, a22, s, response, sub_type }
 * (AffineKeypoint, detectors/structures.hpp:187-196; octave_number / pyramid_scale are never
 * initialised by the reference for these detectors and are not carried).
 */
#ifndef MODS_B200_H
#define MODS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB2_KP 9
#define MB2_DESC_DIM 128

#define MB2_OK 0
#define MB2_ERR_CUDA (-1)
#define MB2_ERR_ARG (-2)
#define MB2_ERR_UNSUPPORTED (-3)
#define MB2_ERR_CAPACITY (-4)

typedef struct mb2_ctx mb2_ctx;

/* [HessianAffine] section of config_iter_mods_cviu.ini == PyramidParams + AffineShapeParams
 * (detectors/structures.hpp:120-165, detectors/affinedetectors/affine.h:28-63). */
typedef struct {
  float threshold;            /* 16/3 */
  int numberOfScales;         /* 3 */
  float initialSigma;         /* 1.6 */
  float edgeEigenValueRatio;  /* 10 */
  int border;                 /* 5 */
  int maxIterations;          /* 16 */
  float convergenceThreshold; /* 0.05 */
  int smmWindowSize;          /* 19 */
  int doBaumberg;             /* 1 */
  int mode;                   /* detection_mode_t: 0 FIXED_TH .. 4 NOT_LESS_THAN_REGIONS */
  int reg_number;
  float rel_threshold;
  float rel_reg_number;
  int patchSize;              /* 41 */
  float mrSize;               /* 3*sqrt(3) */
  int detectorType;           /* detector_type (structures.hpp): 0 DET_HESSIAN, 1 DET_DOG, 2 DET_HARRIS (pyramid.cpp:283-305) -- for DET_DOG the response of every level becomes
                               * level - GaussianBlur(level, sigma = curSigma^2) (pyramid.cpp:176-181, Response() :132-175), the
                               * threshold is used un-squared (pyramid.h:56-57) and the point type is the sign (DOG_DARK 10 /
                               * DOG_BRIGHT 11, pyramid.cpp:92-99); everything else is the same scale-space detector */
} mb2_hessaff_params;

/* [MSER] section of config_iter_mods_cviu.ini == extrema::ExtremaParams (detectors/mser/extrema/extremaParams.h:56-93). */
typedef struct {
  double max_area;      /* 0.05 */
  int min_size;         /* 30 (>= 2) */
  double min_margin;    /* 8 */
  int relative;         /* 0 (relative margins: unsupported) */
  int mode;             /* detection_mode_t: 0 FIXED_TH .. 4 NOT_LESS_THAN_REGIONS */
  int reg_number;
  float rel_threshold;
  float rel_reg_number;
} mb2_mser_params;

/* One synthesised view == ViewSynthParameters (detectors/structures.hpp:198-211) as passed to GenerateSynthImageCorr
 * (synth-detection.cpp:236-245): tilt (< 0: vertical tilt), rotation phi in radians, zoom, anti-aliasing InitSigma. */
typedef struct {
  double tilt, phi, zoom, InitSigma;   /* 1, 0, 1, 0.5 */
  int doBlur;                          /* 1 */
} mb2_view_params;

/* [DominantOrientation] (descriptors_parameters.hpp) as passed to DetectOrientation
 * (synth-detection.cpp:841-849). */
typedef struct {
  double mrSize;   /* 1.0 in config_iter_mods_cviu.ini */
  int patchSize;   /* 41 */
  int maxAngles;   /* 1 */
  double threshold;/* 0.8 */
  int doHalfSIFT;  /* 0; 1 = orientations modulo pi for the Half* descriptors: the upper half of the smoothed histogram is folded onto
                    * the lower one before the peak search (synth-detection.cpp:801-808) */
  int reserved;
} mb2_orientation_params;

/* [SIFTDescriptor] as passed to DescribeRegions<SIFTDescriptor> (synth-detection.hpp:169-172,
 * matching/siftdesc.h:32-79). */
typedef struct {
  double mrSize;   /* 5.1962 */
  int patchSize;   /* 41 */
  int photoNorm;   /* 1 */
  int rootSIFT;    /* 1 = RootSIFT, 0 = SIFT */
  int fastPatchExtraction; /* 0 */
  int doHalfSIFT;  /* 0; 1 with rootSIFT = HalfRootSIFT (siftdesc.cpp:401-442): opposite orientation bins summed -> 64-D, RootSIFT
                    * normalisation on the 64 entries.  Rows of desc_u8 stay 128 wide: entries 0..63 hold the descriptor, 64..127 are 0
                    * (L2 distances between such rows equal the 64-D distances).  doHalfSIFT without rootSIFT is refused: the reference's
                    * SIFTnorm reads 128 entries of the 64-entry vector (siftdesc.cpp:251, 267). */
  int dspScales;   /* 0 = off.  > 0 = DSPSIFT (imagerepresentation.cpp:1547-1598; DomainSizePolingParams, siftdesc.h:19-30: numScales 3,
                    * startCoef 0.5, endCoef 1.5): the un-normalised plain-SIFT votes described at dspScales + 1 measurement-region sizes
                    * mrSize * (dspStartCoef + i (dspEndCoef - dspStartCoef) / dspScales), summed in float, then SIFTnorm on the float
                    * vector.  rootSIFT / doHalfSIFT are ignored (the reference forces plain SIFT here). */
  double dspStartCoef, dspEndCoef;
} mb2_sift_params;

/* ---- context --------------------------------------------------------------------------- */
int mb2_ctx_create(int device, mb2_ctx** out);
/* Same, with the context's streams at the device's highest priority when high_priority != 0: for the context whose work sits on the
 * critical path when several contexts share one GPU (its pending blocks are scheduled ahead of the others'). */
int mb2_ctx_create_prio(int device, int high_priority, mb2_ctx** out);
void mb2_ctx_destroy(mb2_ctx* ctx);
const char* mb2_last_error(const mb2_ctx* ctx);
/* Blocks until everything queued on the context's stream has finished. */
int mb2_ctx_sync(mb2_ctx* ctx);
/* cudaStream_t of the context (for CUDA-event timing by the caller). */
void* mb2_ctx_stream(mb2_ctx* ctx);
int mb2_ctx_device(const mb2_ctx* ctx);
int mb2_ctx_profiling(const mb2_ctx* ctx);   /* 1 between mb2_ctx_profile_begin and _end */
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
long long mb2_ctx_launch_count(const mb2_ctx* ctx);

/* Per-kernel timing for bench.py's roofline: between begin and end every kernel launch of this context
 * is bracketed by CUDA events on the context's stream.  end() writes "name\tlaunches\ttotal_ms\n" lines
 * (plus "__extract_gather_bytes__", the algorithmic gather bytes of the patch extraction) into buf and
 * returns the number of distinct kernels. */
int mb2_ctx_profile_begin(mb2_ctx* ctx);
int mb2_ctx_profile_end(mb2_ctx* ctx, char* buf, int buflen);

/* ---- detection ------------------------------------------------------------------------- */
/* Replaces the detector hook `int DetectAffineKeypoints(cv::Mat&, vector<AffineKeypoint>&,
 * ScaleSpaceDetectorParams, ScalePyramid&, tilt, zoom)` (scale-space-detector.hpp:231,
 * scale-space-detector.cpp:43-85) *plus* the post-step of DetectAffineRegions<>
 * (synth-detection.hpp:93-126) when `as_regions` != 0 (s *= sqrt|det A|, rectifyTransformation).
 * pixels: w*h f32 gray image [H|D].  out_kp: capacity*MB2_KP doubles [H].
 * Returns the number of keypoints (detection order of the reference); > capacity => truncated
 * output and MB2_ERR_CAPACITY. */
int mb2_hessaff_detect(mb2_ctx* ctx, const float* pixels, int w, int h, const mb2_hessaff_params* par,
                       double tilt, double zoom, int as_regions, double* out_kp, int capacity);

/* Replaces the detector hook `int DetectMSERs(cv::Mat&, vector<AffineKeypoint>&, extrema::ExtremaParams, ScalePyramid&,
 * tilt, zoom)` (detectors/mser/extrema/extrema.h:11, extrema.cpp:284-473; doOnNormal branch) and everything below it
 * (getRLEExtrema, libExtrema.cpp:462) *plus* the post-step of DetectAffineRegions<> when `as_regions` != 0.
 * MSER+ regions first, then MSER-, each in the reference's list order; response = margin, sub_type 21 / 20.
 * Returns the number of keypoints; MB2_ERR_UNSUPPORTED when the reference itself leaves the defined behaviour of its
 * packed label fields (a min_reg label absorbing >= 32768 pixels, getExtrema.cpp:322). */
int mb2_mser_detect(mb2_ctx* ctx, const float* pixels, int w, int h, const mb2_mser_params* par, double tilt, double zoom,
                    int as_regions, double* out_kp, int capacity);
/* Diagnostics / parity: the raw region list of getRLEExtrema (libExtrema.cpp:462-482), one row of 13 doubles per
 * (region, threshold): polarity minI maxI threshold margin area border nruns cx cy sxx sxy syy.  FIXED_TH semantics. */
int mb2_mser_regions(mb2_ctx* ctx, const float* pixels, int w, int h, const mb2_mser_params* par, double* out_rows, int capacity);

/* Replaces `int DetectOrientation(AffineRegionList&, AffineRegionList&, SynthImage&, mrSize,
 * patchSize, doHalfSIFT=0, maxAngNum, th, addUpRight=false)` (synth-detection.cpp:841-919). */
int mb2_detect_orientation(mb2_ctx* ctx, const float* pixels, int w, int h, const double* in_kp, int n,
                           const mb2_orientation_params* par, double* out_kp, int capacity);

/* Replaces `template<FuncType> void DescribeRegions(AffineRegionList&, SynthImage&, SIFTDescriptor,
 * mrSize, patchSize, fast_extraction, photoNorm)` (synth-detection.hpp:169-255) with the
 * SIFTDescriptor functor (matching/siftdesc.cpp:401-442).  desc_u8: n*128 bytes (the reference
 * stores the same integers 0..255 as float); patches (optional, may be NULL): n*ps*ps f32. */
int mb2_describe_sift(mb2_ctx* ctx, const float* pixels, int w, int h, const double* kp, int n,
                      const mb2_sift_params* par, uint8_t* desc_u8, float* patches);

/* One (detector = HessianAffine, view) pass of ImageRepresentation::SynthDetectDescribeKeypoints
 * (imagerepresentation.cpp:717-720, 1254-1341): detect -> DetectOrientation -> ReprojectRegions
 * (synth-detection.cpp:541-616) -> DescribeRegions.  `H` is SynthImage::H (original -> view,
 * affine, row-major 3x3); pixels is the view's image.  Outputs [H]: det_kp / reproj_kp
 * (capacity*MB2_KP doubles each), desc_u8 (capacity*128).  The described regions also stay
 * resident on the device as region set `slot` (0..MB2_MAX_SLOTS-1) for mb2_match_slots().
 * Returns the region count. */
#define MB2_MAX_SLOTS 8
int mb2_detect_describe_view(mb2_ctx* ctx, const float* pixels, int w, int h, const double* H,
                             int orig_w, int orig_h, const mb2_hessaff_params* det,
                             const mb2_orientation_params* ori, const mb2_sift_params* desc,
                             int slot, int append, double* det_kp, double* reproj_kp, uint8_t* desc_u8,
                             int capacity);

/* Same pass with detector = MSER (imagerepresentation.cpp:1035-1038). */
int mb2_detect_describe_view_mser(mb2_ctx* ctx, const float* pixels, int w, int h, const double* H,
                                  int orig_w, int orig_h, const mb2_mser_params* det,
                                  const mb2_orientation_params* ori, const mb2_sift_params* desc,
                                  int slot, int append, double* det_kp, double* reproj_kp, uint8_t* desc_u8,
                                  int capacity);

/* Batched form of the MSER detector hook for the two images of a pair (mods.cpp:255-271 runs the two images side by side):
 * the component tree is built level by level and is latency bound, so two same-size images in ONE pass cost hardly more
 * than one.  FIXED_TH only.  Detection only: the keys (as regions) stay on the device; *n1 / *n2 = regions per image.
 * mb2_describe_view_of_pair then runs DetectOrientation -> ReprojectRegions -> DescribeRegions for image `which` (0 / 1)
 * exactly as mb2_detect_describe_view_mser does after its detection (same outputs, same slot semantics); `src` is the
 * context that ran mb2_mser_detect_pair (NULL = ctx itself), so the two images can be finished on two contexts at once. */
int mb2_mser_detect_pair(mb2_ctx* ctx, const float* pixels1, const float* pixels2, int w, int h, const mb2_mser_params* par,
                         int* n1, int* n2);
int mb2_describe_view_of_pair(mb2_ctx* ctx, mb2_ctx* src, int which, const double* H, int orig_w, int orig_h,
                              const mb2_orientation_params* ori, const mb2_sift_params* desc, int slot, int append,
                              double* det_kp, double* reproj_kp, uint8_t* desc_u8, int capacity);

/* Replaces `void GenerateSynthImageCorr(const cv::Mat& in, SynthImage& out, name, tilt, phi, zoom, InitSigma, doBlur, img_id,
 * convert2gray)` (synth-detection.cpp:236-430) for a gray f32 image [H|D]: rotation warp (border 128) -> anisotropic
 * anti-aliasing blur -> tilt / zoom warp, with OpenCV 2.4.9's fixed-point bilinear warpAffine.  out (optional, [H]):
 * capacity floats, receives the view row-major when it fits; *ow, *oh its size; H9 = SynthImage::H (original -> view).
 * Returns 1 for the identity view (pixels untouched), 0 otherwise. */
int mb2_synth_view(mb2_ctx* ctx, const float* pixels, int w, int h, const mb2_view_params* view, float* out, int capacity,
                   int* ow, int* oh, double* H9);
/* One (detector, view) pass of SynthDetectDescribeKeypoints for an arbitrary view (imagerepresentation.cpp:621-1341):
 * GenerateSynthImageCorr -> detect on the view (detector 0 = HessianAffine with `hess`, 3 = MSER with `mser`) ->
 * DetectOrientation -> ReprojectRegions to the original frame (regions leaving the original image are dropped) ->
 * DescribeRegions on the view.  Outputs and slot semantics as mb2_detect_describe_view. */
int mb2_detect_describe_synth_view(mb2_ctx* ctx, const float* pixels, int w, int h, const mb2_view_params* view, int detector,
                                   const mb2_hessaff_params* hess, const mb2_mser_params* mser,
                                   const mb2_orientation_params* ori, const mb2_sift_params* desc, int slot, int append,
                                   double* det_kp, double* reproj_kp, uint8_t* desc_u8, int capacity);

/* Scheduling aid for callers that run several contexts on one GPU: the MSER component-tree kernel is bound by memory latency
 * and slows down badly next to bandwidth-hungry kernels.  mb2_ctx_tree_epoch(src) = number of tree kernels `src` has launched so
 * far (safe to poll from another host thread); mb2_ctx_wait_tree(ctx, src) makes everything submitted to ctx from now on wait for
 * the most recent tree kernel of src (stream-ordered, no host block). */
long long mb2_ctx_tree_epoch(const mb2_ctx* ctx);
int mb2_ctx_wait_tree(mb2_ctx* ctx, mb2_ctx* src);

/* Copies the regions of the most recent mb2_detect_describe_view (which may be called with NULL
 * outputs to learn the count first) to the host.  Returns that count. */
int mb2_view_fetch(mb2_ctx* ctx, double* det_kp, double* reproj_kp, uint8_t* desc_u8, int capacity);

/* ---- device-resident region records (view-sharded multi-GPU path) --------------------------------------------------
 * One record per described region, MB2_REGION_RECORD_BYTES = 184: the 128-byte descriptor followed by 7 doubles of reproj_kp
 * (x y a11 a12 a21 a22 s) -- everything MatchFlannFGINN, DuplicateFiltering and the LAF checks read of an AffineRegion
 * (imagerepresentation.h:66 holds them per (detector, descriptor); the reference appends the views of an image in view-index order,
 * imagerepresentation.cpp:2044-2045).  Records never leave the device between detection and matching: ranks exchange them with one
 * ncclAllGather (mods_b200/host: mb2_views_sharded_pair).
 * mb2_view_pack: records of the most recent mb2_detect_describe*_view into d_dst [D] (capacity records); returns the region count.
 * mb2_slot_from_records: region set `slot` (descriptors + centres, as the view calls leave it) from n records [D].
 * mb2_match_slots_range: mb2_match_slots for the queries [q_lo, q_hi) only (row-sharded N1 x N2 matching); rows carry the query's
 * index in the whole set. */
#define MB2_REGION_RECORD_BYTES 184
/* device scratch + stream-ordered copies on the context's stream for callers above the C ABI that keep data on the device (the
 * view-sharded driver): kind 0 host->device, 1 device->host, 2 device->device.  mb2_ctx_sync() waits for them. */
/* measured FP64 FMA throughput of the device (TFLOP/s, a DFMA micro-kernel): the denominator of the scorer's roofline in bench.py */
int mb2_debug_fp64_peak(mb2_ctx* ctx, double* tflops);
int mb2_ctx_make_current(mb2_ctx* ctx);   /* cudaSetDevice(device of ctx) on the calling thread */
int mb2_is_device_pointer(const void* p);   /* 1: device (or managed) memory, 0: host */
void* mb2_dev_alloc(mb2_ctx* ctx, size_t bytes);
void mb2_dev_free(mb2_ctx* ctx, void* p);
int mb2_dev_copy(mb2_ctx* ctx, void* dst, const void* src, size_t bytes, int kind);
int mb2_view_pack(mb2_ctx* ctx, void* d_dst, int capacity);
int mb2_slot_from_records(mb2_ctx* ctx, int slot, const void* d_records, int n);
int mb2_match_slots_range(mb2_ctx* ctx, int q_slot, int t_slot, int q_lo, int q_hi, double matchRatio, double contradDist, int nn,
                          double* out, int capacity);
/* frames14 [H]: n rows of 14 doubles = the 7 reproj_kp doubles of record q_idx[i] of d_q followed by those of record t_idx[i] of d_t
 * (q_idx / t_idx [H]) -- what DuplicateFiltering / LORANSACFiltering read of a TentativeCorrespExt, without moving whole record sets. */
int mb2_records_gather_frames(mb2_ctx* ctx, const void* d_q, const void* d_t, const int* q_idx, const int* t_idx, int n, double* frames14);
/* order-sensitive 64-bit checksum of n records [D], computed on the device: sum over i of (i + 1) * FNV-1a(record i) modulo 2^64 */
int mb2_records_checksum(mb2_ctx* ctx, const void* d_records, int n, unsigned long long* out);

/* Hands a device-resident region set over to another context on the same GPU (no copy).  Lets two host
 * threads run mb2_detect_describe_view for the two images of a pair on two contexts (= two streams), the
 * way mods.cpp:255-271 runs them as two OpenMP tasks, and match them afterwards on one. */
int mb2_slot_move(mb2_ctx* dst, int dst_slot, mb2_ctx* src, int src_slot);

/* ---- matching -------------------------------------------------------------------------- */
/* Replaces `int MatchFlannFGINN(const AffineRegionList& q, const AffineRegionList& t,
 * TentativeCorrespListExt&, const MatchPars&, int nn = 50)` (matching/matching.cpp:357-461) for
 * vector_matcher = linear (exact kNN): squared-L2 first NN, first-geometrically-inconsistent
 * ratio test.  q_desc/t_desc: n*128 u8 [H|D]; t_xy: nt*2 doubles = reproj_kp.(x,y) of the trains
 * [H|D].  out [H]: rows of 7 doubles { query, idx0, idxJ, idx1, d0, dJ, d1 } in query order,
 * i.e. TentativeCorrespExt{first, second, secondbad, secondbadby2ndcl, d1, d2, d2by2ndcl}.
 * Ties between equal distances: lower train index first (FLANN's order is unspecified).
 * matchRatio >= 1 selects the "all points" branch (matching.cpp:397-428: NN paired with its first geometrically inconsistent neighbour,
 * or with neighbour nn - 1), evaluated from an exact sorted k-NN table per query (2 <= nn <= 64).
 * Returns the number of tentatives. */
int mb2_match_fginn(mb2_ctx* ctx, const uint8_t* q_desc, int nq, const uint8_t* t_desc, int nt,
                    const double* t_xy, double matchRatio, double contradDist, int nn, double* out,
                    int capacity);
/* Replaces `int MatchFLANNDistance(const AffineRegionList& q, const AffineRegionList& t, TentativeCorrespListExt&, const MatchPars&,
 * int nn = 50)` (matching/matching.cpp:607-666; called from correspondencebank.cpp:284,343 for binary descriptors when
 * matchDistanceThreshold > 0) for binary_matcher = linear, binary_dist = Hamming: exact 2-NN on the number of differing bits, a query
 * is kept when its first distance is <= (int)(float)matchDistanceThreshold.  q_desc / t_desc: n * desc_bytes bytes [H|D]
 * (the reference floors AffineRegion::desc.vec to bytes itself), desc_bytes 1..64 (ORB 32, BRISK / FREAK 64); nt >= 2.
 * out [H]: rows of 7 doubles in the layout of mb2_match_fginn { query, idx0, idx1, idx1, d0, d1, d1 }: TentativeCorrespExt{first,
 * second, d1, d2}, ratio = d0 / d1 in double is the caller's (matching.cpp:659).  Ties: lower train index first.
 * Returns the number of tentatives. */
int mb2_match_hamming(mb2_ctx* ctx, const uint8_t* q_desc, int nq, const uint8_t* t_desc, int nt, int desc_bytes,
                      double matchDistanceThreshold, double* out, int capacity);
/* Same, on two device-resident region sets left by mb2_detect_describe_view. */
int mb2_match_slots(mb2_ctx* ctx, int q_slot, int t_slot, double matchRatio, double contradDist, int nn,
                    double* out, int capacity);

/* ---- verification ---------------------------------------------------------------------- */
/* Batched form of the DEGENSAC scorer hooks `typedef void (*HDsPtr)(const double* lin, const
 * double* u, const double* H, double* p, int len)` (degensac/Htools.h:1; HDs Htools.c:158-196,
 * HDsSym :199-240, HDsSymMax :241-282) and `FDsPtr` (degensac/Fcustomdef.h:3; FDs Ftools.c:82-100,
 * FDsSym :102-123).  u: len*6 doubles (x1 y1 1 x2 y2 1) [H|D]; models: K*9 doubles (h stored as in
 * DEGENSAC, column-wise, 2nd image -> 1st) [H|D].  which: 0 HDs, 1 HDsSym, 2 HDsSymMax, 3 FDs,
 * 4 FDsSym, 5 the residual of exFDsSym (Ftools.c:172-196; same quantity as FDsSym, formed as r^2 / (ab/(a+b)), which is what the
 * F-matrix LO thresholds).  Outputs [H], each may be NULL: resid K*len doubles; I[K] = #{d <= th};
 * J[K] = sum truncQuad(d, th) (rtools.c:228-236). */
int mb2_score_models(mb2_ctx* ctx, int which, const double* u, int len, const double* models, int K,
                     double th, double* resid, int* I, double* J);

/* Replaces `Score exp_ransacHcustom(double* u, int len, double th, double conf, int max_sam,
 * double* H, unsigned char* inl, int iter_type = 4, int* data_out, int oriented_constraint = 1,
 * unsigned inlLimit = 0, double** resids, HDsPtr, HDsiPtr, HDsidxPtr, int doSymCheck)`
 * (degensac/exp_ranH.h:32-36, exp_ranH.c:796-1236) as LORANSACFiltering calls it
 * (matching/matching.cpp:891): LO-RANSAC homography with MSAC scoring.  Hypotheses are generated
 * on the host exactly as the reference does (same libc rand() stream from `seed`, where the
 * reference uses time(NULL)), scored in batches on the GPU, and replayed through the reference's
 * sequential best-so-far / LO / adaptive-stop logic.  errorType: 0 Sampson, 1 SymmMax, 2 SymmSum.
 * data_out[0..2] = samples, LO count, rejected-by-orientation.  Returns #inliers (Score.I). */
int mb2_ransac_h(mb2_ctx* ctx, const double* u, int len, double th, double conf, int max_sam,
                 int errorType, int doSymCheck, long seed, double* H, unsigned char* inl, int* data_out,
                 double* J);

/* Replaces `int exp_ransacFcustom(double* u, int len, double th, double conf, int max_sam, double* F,
 * unsigned char* inl, int* data_out, int do_lo, unsigned inlLimit, double** resids, double* H_best, int* Ih,
 * exFDsPtr EXFDS1, FDsPtr FDS1, int doSymCheck)` (degensac/exp_ranF.h:70-72, exp_ranF.c:795-1192) as
 * LORANSACFiltering calls it in F mode (matching/matching.cpp:883, which passes inlLimit = 0 and
 * do_lo = pars.localOptimization): 7-point LO-RANSAC with MSAC scoring, oriented-epipolar and symmetric-distance
 * checks, the DEGENSAC test of every so-far-the-best sample (checksample -> innerH -> plane-and-parallax rFtH)
 * and the inner-RANSAC + iterated weighted LSQ local optimisation.  The 7-point models (up to 3 per sample) are
 * generated on the host from the reference's rand() stream (`seed` stands in for time(NULL)), scored in batches
 * on the GPU and replayed through the reference's sequential logic; all later residual vectors are GPU launches
 * too.  errorType: 0 Sampson (FDs / exFDs), otherwise symmetric epipolar (FDsSym / exFDsSym).
 * u [H]: len*6 doubles; F [H]: 9 doubles; inl [H]: len flags.  data_out[0..3] = samples, LO count, Ih (inliers of
 * the best homography met by the degeneracy test), scorer launches.  Returns #inliers (maxS.I), < 0 on error. */
int mb2_ransac_f(mb2_ctx* ctx, const double* u, int len, double th, double conf, int max_sam,
                 int errorType, int doSymCheck, int do_lo, unsigned inlLimit, long seed, double* F,
                 unsigned char* inl, int* data_out, double* J);

/* ---- diagnostics ------------------------------------------------------------------------ */
/* Copies one plane of the most recent scale-space pyramid (ScalePyramid / Octave::blurs,
 * detectors/structures.hpp:167-185, which DetectAffineKeypoints exports through its
 * `ScalePyramid&` argument) to the host: want_resp = 0 the blurred level, 1 its Hessian response.
 * out may be NULL to query rows/cols.  Returns the number of octaves. */
int mb2_debug_pyramid_level(mb2_ctx* ctx, int octave, int level, int want_resp, float* out, int* rows, int* cols);

#ifdef __cplusplus
}
#endif
#endif
