// TEST INFRASTRUCTURE (oracle) -- not part of the shipped product path.
//
// extern "C" doorway into the reference's UNMODIFIED matching/matching.cpp, compiled in place by oracle/build_ref.sh into
// oracle/_ref/libmods_ref.so: MatchFlannFGINN (:357-461), DuplicateFiltering (:2983-3047), LORANSACFiltering (:806-980) with its
// post-checks NaiveHCheck (:1171-1200), H_LAF_check (:251-309), F_LAF_check (:193-250).  OpenCV's FLANN is not in the reference tree:
// cv::flann::Index is answered by oracle/shim/opencv2/flann/flann.hpp (exact linear k-NN, the `vector_matcher=linear` setting).
// Nothing here restates reference arithmetic; the functions only move plain arrays in and out of the reference's own structs.
//
// Region record (KP = 9 doubles): x y a11 a12 a21 a22 s response sub_type -- the reproj_kp of an AffineRegion.
// Tentative rows (7 doubles): query, idx0 (second), idxJ (secondbad), idx1 (secondbadby2ndcl), d0, dJ, d1.
#include <cstring>
#include <vector>
#include <opencv2/core/core.hpp>
#include "detectors/structures.hpp"
#include "matching/matching.hpp"

extern "C" long mb2_ref_seed;   // srand(time(NULL)) of exp_ransac*custom sees this through the build recipe's time() hook

// ORSAFiltering (matching.cpp:982, ver_type 3) is outside the hot path; orsa.cpp is not compiled.  The symbol only has to resolve.
struct Match;
float orsa(int, int, std::vector<Match>&, std::vector<float>&, int, int, int, int, int, double*) { cv::shim_unsupported("orsa"); }

namespace {
const int KP = 9;
void kp_in(const double* o, AffineKeypoint& k) {
  k.x = o[0]; k.y = o[1]; k.a11 = o[2]; k.a12 = o[3]; k.a21 = o[4]; k.a22 = o[5];
  k.s = o[6]; k.response = o[7]; k.sub_type = (int)o[8]; k.octave_number = 0; k.pyramid_scale = 0;
}
AffineRegionList regions_in(const double* kps, const float* desc, int n, int dim) {
  AffineRegionList l(n);
  for (int i = 0; i < n; i++) {
    l[i].img_id = 0; l[i].img_reproj_id = 0; l[i].id = i; l[i].parent_id = 0; l[i].type = DET_HESSIAN;
    kp_in(kps + (size_t)i * KP, l[i].reproj_kp);
    l[i].det_kp = l[i].reproj_kp;
    l[i].desc.type = DESC_UNKNOWN;
    if (desc) l[i].desc.vec.assign(desc + (size_t)i * dim, desc + (size_t)(i + 1) * dim);
  }
  return l;
}
MatchPars match_pars(double ratio, double contradDist) {
  MatchPars p;
  p.currMatchRatio = ratio; p.contradDist = contradDist;
  p.vector_matcher = cvflann::FLANN_INDEX_LINEAR; p.vector_dist = cvflann::FLANN_DIST_L2;
  return p;
}
RANSACPars ransac_pars(int useF, double err_threshold, double confidence, int max_samples, int localOptimization, double LAFCoef, double HLAFCoef,
                       int errorType, int doSymmCheck) {
  RANSACPars p;
  p.useF = useF; p.err_threshold = err_threshold; p.confidence = confidence; p.max_samples = max_samples; p.localOptimization = localOptimization;
  p.LAFCoef = LAFCoef; p.HLAFCoef = HLAFCoef; p.errorType = (RANSAC_error_t)errorType; p.doSymmCheck = doSymmCheck; p.justMarkOutliers = 0;
  return p;
}
// frames14 row: reproj_kp of `first` (x y a11 a12 a21 a22 s) then of `second`; ids carry the row number
TentativeCorrespListExt tentatives_in(const double* frames14, const double* ratio, int n) {
  TentativeCorrespListExt L;
  L.TCList.resize(n);
  for (int i = 0; i < n; i++) {
    TentativeCorrespExt& t = L.TCList[i];
    const double* f = frames14 + (size_t)i * 14;
    double a[KP] = {f[0], f[1], f[2], f[3], f[4], f[5], f[6], 0, 0}, b[KP] = {f[7], f[8], f[9], f[10], f[11], f[12], f[13], 0, 0};
    kp_in(a, t.first.reproj_kp); kp_in(b, t.second.reproj_kp);
    t.first.det_kp = t.first.reproj_kp; t.second.det_kp = t.second.reproj_kp;
    t.first.id = i; t.second.id = i; t.first.img_id = t.second.img_id = 0; t.first.img_reproj_id = t.second.img_reproj_id = 0;
    t.first.parent_id = t.second.parent_id = 0; t.first.type = t.second.type = DET_HESSIAN;
    t.secondbad = t.second; t.secondbadby2ndcl = t.second;
    t.d1 = 0; t.d2 = 0; t.d2by2ndcl = 0; t.d2byDB = 0; t.ratio = ratio ? ratio[i] : 0; t.isTrue = 0;
  }
  return L;
}
}  // namespace

extern "C" {

// MatchFlannFGINN (matching.cpp:357) on plain arrays.  t_kps: nt x 9 reprojected regions (only x, y are read: distanceSq).
int ref_match_fginn(const float* q, int nq, const float* t, int nt, const double* t_kps, double ratio, double contradDist, int nn,
                    double* out, int max_out) {
  std::vector<double> qk((size_t)nq * KP, 0.0);
  AffineRegionList l1 = regions_in(qk.data(), q, nq, 128), l2 = regions_in(t_kps, t, nt, 128);
  TentativeCorrespListExt tents;
  MatchFlannFGINN(l1, l2, tents, match_pars(ratio, contradDist), nn);
  const int n = (int)tents.TCList.size();
  for (int i = 0; i < n && i < max_out; i++) {
    const TentativeCorrespExt& c = tents.TCList[i];
    double* o = out + (size_t)i * 7;
    o[0] = c.first.id; o[1] = c.second.id; o[2] = c.secondbad.id; o[3] = c.secondbadby2ndcl.id; o[4] = c.d1; o[5] = c.d2; o[6] = c.d2by2ndcl;
  }
  return n;
}

// MatchFLANNDistance (matching.cpp:607-666) on plain arrays: 2-NN of binary descriptors in Hamming distance, a query is kept when its
// first distance is <= (int)(float)matchDistanceThreshold.  Descriptors arrive as floats (AffineRegion::desc.vec) and are floored to
// bytes by the reference itself.  out rows (5 doubles): query, second.id, d1, d2, ratio.
int ref_match_hamming(const float* q, int nq, const float* t, int nt, int dim, double matchDistanceThreshold, double* out, int max_out) {
  std::vector<double> qk((size_t)nq * KP, 0.0), tk((size_t)nt * KP, 0.0);
  AffineRegionList l1 = regions_in(qk.data(), q, nq, dim), l2 = regions_in(tk.data(), t, nt, dim);
  MatchPars p;
  p.matchDistanceThreshold = matchDistanceThreshold;
  p.binary_matcher = cvflann::FLANN_INDEX_LINEAR; p.binary_dist = cvflann::FLANN_DIST_HAMMING;
  TentativeCorrespListExt tents;
  MatchFLANNDistance(l1, l2, tents, p, 2);
  const int n = (int)tents.TCList.size();
  for (int i = 0; i < n && i < max_out; i++) {
    const TentativeCorrespExt& c = tents.TCList[i];
    double* o = out + (size_t)i * 5;
    o[0] = c.first.id; o[1] = c.second.id; o[2] = c.d1; o[3] = c.d2; o[4] = c.ratio;
  }
  return n;
}

// DuplicateFiltering (matching.cpp:2983): kept_idx receives the surviving rows in the reference's output order.
int ref_duplicate_filter(const double* frames14, const double* ratio, int n, double r, int mode, int* kept_idx) {
  TentativeCorrespListExt L = tentatives_in(frames14, ratio, n);
  DuplicateFiltering(L, r, mode);
  for (size_t i = 0; i < L.TCList.size(); i++) kept_idx[i] = L.TCList[i].first.id;
  return (int)L.TCList.size();
}

// LORANSACFiltering (matching.cpp:806) with a fixed seed.  verified_idx: rows that survive RANSAC + NaiveHCheck + LAF check, in
// output order; inl: RANSAC inlier flag per input row (TentativeCorrespExt::isTrue); H: 9 doubles as LORANSACFiltering returns them.
int ref_loransac_filtering(const double* frames14, int n, int useF, double err_threshold, double confidence, int max_samples, int localOptimization,
                           double LAFCoef, double HLAFCoef, int errorType, int doSymmCheck, long seed, int* verified_idx, unsigned char* inl, double* H) {
  TentativeCorrespListExt in = tentatives_in(frames14, nullptr, n), out;
  mb2_ref_seed = seed;
  for (int i = 0; i < 9; i++) H[i] = 0;
  const int k = LORANSACFiltering(in, out, H, ransac_pars(useF, err_threshold, confidence, max_samples, localOptimization, LAFCoef, HLAFCoef, errorType, doSymmCheck));
  if (useF) for (int i = 0; i < 9; i++) H[i] = out.H[i];
  for (int i = 0; i < n; i++) inl[i] = n >= 8 ? (unsigned char)(in.TCList[i].isTrue != 0) : 0;
  for (size_t i = 0; i < out.TCList.size(); i++) verified_idx[i] = out.TCList[i].first.id;
  return k < (int)out.TCList.size() ? (int)out.TCList.size() : k;
}

// The back half of one mods.cpp iteration (mods.cpp:290-351) on G (detector, descriptor) groups: MatchFlannFGINN per group
// (correspondencebank.cpp:340), the groups appended in order (GetCorresponcesVector("All","All")), DuplicateFiltering(MODE_FGINN),
// LORANSACFiltering.  Regions carry their descriptors, as in the reference (every TentativeCorrespExt holds four AffineRegions).
// tent_out: rows of 8 doubles = group + the 7 tentative columns; counts: [tentatives, unique, ransac inliers, verified].
int ref_pair_back(int G, const int* nq, const double* const* q_kps, const float* const* q_desc, const int* nt, const double* const* t_kps,
                  const float* const* t_desc, const double* ratio, double contradDist, int nn, double duplicateDist, int useF, double err_threshold,
                  double confidence, int max_samples, int localOptimization, double LAFCoef, double HLAFCoef, int errorType, int doSymmCheck, long seed,
                  double* tent_out, int tent_cap, int* kept_idx, int* verified_idx, double* H, int* counts) {
  TentativeCorrespListExt all;
  int n_all = 0;
  for (int g = 0; g < G; g++) {
    AffineRegionList l1 = regions_in(q_kps[g], q_desc[g], nq[g], 128), l2 = regions_in(t_kps[g], t_desc[g], nt[g], 128);
    TentativeCorrespListExt tents;
    MatchFlannFGINN(l1, l2, tents, match_pars(ratio[g], contradDist), nn);
    for (size_t i = 0; i < tents.TCList.size(); i++) {
      TentativeCorrespExt& c = tents.TCList[i];
      if (n_all < tent_cap) {
        double* o = tent_out + (size_t)n_all * 8;
        o[0] = g; o[1] = c.first.id; o[2] = c.second.id; o[3] = c.secondbad.id; o[4] = c.secondbadby2ndcl.id; o[5] = c.d1; o[6] = c.d2; o[7] = c.d2by2ndcl;
      }
      c.first.parent_id = n_all;   // row in the appended list ("id" keeps the region index)
      n_all++;
    }
    AddMatchingsToList(all, tents);
  }
  counts[0] = n_all; counts[1] = counts[2] = counts[3] = 0;
  DuplicateFiltering(all, duplicateDist, MODE_FGINN);
  counts[1] = (int)all.TCList.size();
  for (size_t i = 0; i < all.TCList.size(); i++) kept_idx[i] = all.TCList[i].first.parent_id;
  TentativeCorrespListExt verified;
  mb2_ref_seed = seed;
  for (int i = 0; i < 9; i++) H[i] = 0;
  LORANSACFiltering(all, verified, H, ransac_pars(useF, err_threshold, confidence, max_samples, localOptimization, LAFCoef, HLAFCoef, errorType, doSymmCheck));
  if (useF) for (int i = 0; i < 9; i++) H[i] = verified.H[i];
  if (all.TCList.size() >= 8) for (size_t i = 0; i < all.TCList.size(); i++) counts[2] += all.TCList[i].isTrue != 0;
  counts[3] = (int)verified.TCList.size();
  for (size_t i = 0; i < verified.TCList.size(); i++) verified_idx[i] = verified.TCList[i].first.parent_id;
  return counts[3];
}

}  // extern "C"
