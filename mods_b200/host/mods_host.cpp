// Host-side mirror of the reference's plugin surface (see mods_host.hpp).  Everything numeric runs
// in libmods_b200.so through the C ABI; what stays here is what the reference also does in host C++:
// region bookkeeping, the O(T) (grid-accelerated, same result) duplicate filter and the small
// post-RANSAC checks.
#include "mods_host.hpp"
#include "../csrc/parallel_host.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <condition_variable>
#include <deque>
#include <fstream>
#include <iostream>
#include <memory>
#include <mutex>
#include <thread>
#include <unordered_map>

namespace mods {

namespace {
double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
const char* kDetectorNames[] = {"HessianAffine", "DoG", "HarrisAffine", "MSER"};  // detectors/structures.hpp:39-44 (prefix)
AffineKeypoint kp_from(const double* v) {
  AffineKeypoint k;
  k.x = v[0]; k.y = v[1]; k.a11 = v[2]; k.a12 = v[3]; k.a21 = v[4]; k.a22 = v[5]; k.s = v[6]; k.response = v[7];
  k.sub_type = (int)v[8]; k.octave_number = 0; k.pyramid_scale = 0;
  return k;
}
}  // namespace

DetectorsParameters::DetectorsParameters() {
  // [HessianAffine] of config_iter_mods_cviu.ini
  HessParam = mb2_hessaff_params{5.3333f, 3, 1.6f, 10.0f, 5, 16, 0.05f, 19, 1, 0, 2000, -1.f, -1.f, 41, 3.0f * std::sqrt(3.0f), 0};
  // [MSER] of config_iter_mods_cviu.ini
  MSERParam = mb2_mser_params{0.05, 30, 8.0, 0, 0, -1, -1.f, -1.f};
}
DescriptorsParameters::DescriptorsParameters() {
  SIFTParam = mb2_sift_params{5.1962, 41, 1, 0, 0, 0, 0, 0.5, 1.5};
  RootSIFTParam = mb2_sift_params{5.1962, 41, 1, 1, 0, 0, 0, 0.5, 1.5};
  HalfRootSIFTParam = RootSIFTParam; HalfRootSIFTParam.doHalfSIFT = 1;   // io_mods.cpp:752-753
  HalfSIFTParam = HalfRootSIFTParam;                                     // io_mods.cpp:756 (HalfSIFT is HalfRootSIFT in the reference)
}

// ---- SetVSPars (synth-detection.cpp:103-234) --------------------------------------------------------
int SetVSPars(const std::vector<double>& scale_set, const std::vector<double>& tilt_set, const double phi_base,
              const std::vector<double>& FGINNThreshold, const std::vector<double>& DistanceThreshold,
              const std::vector<std::string> descriptors, std::vector<ViewSynthParameters>& par, std::vector<ViewSynthParameters>& prev_par,
              const double InitSigma, const int doBlur, const int dsplevels, const double minSigma, const double maxSigma) {
  const double eps1 = 0.01;   // synth-detection.cpp:29
  par.clear();
  std::vector<ViewSynthParameters> pars_tmp;
  auto make = [&](double phi, double tilt, double zoom, int blur) {
    ViewSynthParameters t;
    t.phi = phi; t.tilt = tilt; t.zoom = zoom; t.InitSigma = InitSigma; t.doBlur = blur; t.DSPlevels = dsplevels;
    t.minSigma = minSigma; t.maxSigma = maxSigma; t.descriptors = descriptors;
    for (size_t d = 0; d < descriptors.size(); d++) {
      t.DistanceThreshold[descriptors[d]] = d < DistanceThreshold.size() ? DistanceThreshold[d] : 0;
      t.FGINNThreshold[descriptors[d]] = d < FGINNThreshold.size() ? FGINNThreshold[d] : 0;
    }
    return t;
  };
  if (scale_set.empty() || tilt_set.empty()) pars_tmp.push_back(make(0, 0, 0, 0));   // :120-136
  for (size_t sc = 0; sc < scale_set.size(); sc++)
    for (size_t t = 0; t < tilt_set.size(); t++) {
      if (std::fabs(tilt_set[t] - 1) > eps1) {
        int n_rot1 = (int)std::floor(180.0 * tilt_set[t] / phi_base);
        double delta_phi = M_PI / n_rot1;
        if (n_rot1 < 0) {   // no rotation mode if negative: one vertical tilt and one horizontal one (:144-169)
          n_rot1 = 1; delta_phi = 0;
          pars_tmp.push_back(make(0, -tilt_set[t], scale_set[sc], doBlur));
        }
        for (int r = 0; r < n_rot1; r++) pars_tmp.push_back(make(delta_phi * r, tilt_set[t], scale_set[sc], doBlur));
      } else
        pars_tmp.push_back(make(0, tilt_set[t], scale_set[sc], doBlur));
    }
  for (const ViewSynthParameters& p : pars_tmp) {
    bool unique = true;
    for (const ViewSynthParameters& q : prev_par)
      if ((std::fabs(p.zoom - q.zoom) <= eps1) && (std::fabs(p.tilt - q.tilt) <= eps1) && (std::fabs(p.phi - q.phi) <= eps1)) { unique = false; break; }
    if (unique) par.push_back(p);
  }
  for (const ViewSynthParameters& p : par) prev_par.push_back(p);
  return (int)par.size();
}

// ---- ImageRepresentation ------------------------------------------------------------------------
ImageRepresentation::ImageRepresentation(mb2_ctx* c, GrayImage img, std::string name, int device_slot)
    : OriginalImg(img), ctx(c), Name(name), slot(device_slot) {}

descriptor_type ImageRepresentation::GetDescriptorType(std::string n) const {
  if (n == "SIFT") return DESC_SIFT;
  if (n == "RootSIFT") return DESC_ROOT_SIFT;
  if (n == "HalfSIFT") return DESC_HALF_SIFT;
  if (n == "HalfRootSIFT") return DESC_HALF_ROOT_SIFT;
  if (n == "DSPSIFT") return DESC_DSP_SIFT;
  return DESC_UNKNOWN;
}
detector_type ImageRepresentation::GetDetectorType(std::string n) const {
  for (int i = 0; i < 4; i++) if (n == kDetectorNames[i]) return (detector_type)i;
  return DET_UNKNOWN;
}
AffineRegion ImageRepresentation::RegionBlock::region(int i, bool with_desc) const {
  AffineRegion r;
  r.img_id = img_id[i]; r.img_reproj_id = 0; r.id = i; r.parent_id = 0; r.type = det;  // ids carry no information
  r.det_kp = kp_from(&det_kp[(size_t)i * MB2_KP]);
  r.reproj_kp = kp_from(&reproj_kp[(size_t)i * MB2_KP]);
  r.desc.type = desc;
  if (with_desc) r.desc.vec.assign(desc_u8.begin() + (size_t)i * 128, desc_u8.begin() + (size_t)(i + 1) * 128);
  return r;
}
int ImageRepresentation::GetRegionsNumber(std::string det_name) const {  // imagerepresentation.cpp:264-283 ("None" lists)
  int n = 0;
  for (auto& d : Blocks) {
    if (det_name != "All" && d.first != det_name) continue;
    for (auto& e : d.second) { n += e.second.n; break; }
  }
  return n;
}
int ImageRepresentation::GetDescriptorsNumber(std::string desc_name, std::string det_name) const {  // :284-330
  int n = 0;
  for (auto& d : Blocks) {
    if (det_name != "All" && d.first != det_name) continue;
    for (auto& e : d.second) if (desc_name == "All" || e.first == desc_name) n += e.second.n;
  }
  return n;
}
AffineRegionVector ImageRepresentation::GetAffineRegionVector(std::string desc_name, std::string det_name) const {  // :420-437
  AffineRegionVector out;
  auto d = Blocks.find(det_name);
  if (d == Blocks.end()) return out;
  auto e = d->second.find(desc_name);
  if (e == d->second.end()) return out;
  out.reserve(e->second.n);
  for (int i = 0; i < e->second.n; i++) out.push_back(e->second.region(i, true));
  return out;
}

void ImageRepresentation::SynthDetectDescribeKeypoints(IterationViewsynthesisParam& synth_par, DetectorsParameters& det_par,
                                                       DescriptorsParameters& desc_par, DominantOrientationParams& dom_ori_par) {
  // imagerepresentation.cpp:603-2047: HessianAffine branch (:717-720) and MSER branch (:1035-1038) with SIFT-like
  // descriptors (:1254-1341); every view goes through GenerateSynthImageCorr on the device (mb2_detect_describe_synth_view).
  for (int det = 0; det < 4; det++) {
    const std::string curr_det = kDetectorNames[det];
    auto it = synth_par.find(curr_det);
    if (it == synth_par.end() || (curr_det != "HessianAffine" && curr_det != "MSER")) continue;
    const bool is_mser = curr_det == "MSER";
    SlotState& ss = slot_state[curr_det];
    for (size_t synth = 0; synth < it->second.size(); synth++) {   // views are appended in view-index order (:2044-2045)
      const ViewSynthParameters& v = it->second[synth];
      const bool identity = (std::fabs(v.tilt - 1.) <= 0.1) && (std::fabs(v.phi) <= 0.2) && (std::fabs(v.zoom - 1.) <= 0.1);  // synth-detection.cpp:278
      const double H[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      const mb2_view_params vp{v.tilt, v.phi, v.zoom, v.InitSigma, v.doBlur};
      // HalfSIFT_like_desc is sticky over the view's descriptor list (imagerepresentation.cpp:693-714): one Half* name makes
      // EVERY SIFT-like descriptor of the view use the regions oriented modulo pi (:1257-1262, :1288-1291)
      bool half_like = false;
      for (const std::string& d : v.descriptors) half_like = half_like || d.find("Half") != std::string::npos;
      for (const std::string& curr_desc : v.descriptors) {
        const descriptor_type dt = GetDescriptorType(curr_desc);
        if (dt == DESC_UNKNOWN) continue;
        mb2_sift_params sp = dt == DESC_ROOT_SIFT ? desc_par.RootSIFTParam : dt == DESC_HALF_ROOT_SIFT ? desc_par.HalfRootSIFTParam
                             : dt == DESC_HALF_SIFT ? desc_par.HalfSIFTParam : desc_par.SIFTParam;
        if (dt == DESC_DSP_SIFT) { sp.rootSIFT = 0; sp.doHalfSIFT = 0; sp.dspScales = desc_par.DSPScales; sp.dspStartCoef = desc_par.DSPStartCoef; sp.dspEndCoef = desc_par.DSPEndCoef; }   // imagerepresentation.cpp:1547-1598
        mb2_orientation_params op{dom_ori_par.mrSize, dom_ori_par.patchSize, dom_ori_par.maxAngles, (double)dom_ori_par.threshold, half_like ? 1 : 0, 0};
        const double t0 = now_ms();
        const bool use_slot = slot >= 0 && (ss.desc.empty() || ss.desc == curr_desc);
        const int dev_slot = use_slot ? slot_of(curr_det) : MB2_MAX_SLOTS - 1;
        int n;
        if (!identity)
          n = mb2_detect_describe_synth_view(ctx, OriginalImg.data, OriginalImg.cols, OriginalImg.rows, &vp, is_mser ? 3 : 0, &det_par.HessParam,
                                             &det_par.MSERParam, &op, &sp, dev_slot, use_slot && ss.count > 0, nullptr, nullptr, nullptr, 0);
        else
          n = is_mser ? mb2_detect_describe_view_mser(ctx, OriginalImg.data, OriginalImg.cols, OriginalImg.rows, H, OriginalImg.cols, OriginalImg.rows,
                                                      &det_par.MSERParam, &op, &sp, dev_slot, use_slot && ss.count > 0, nullptr, nullptr, nullptr, 0)
                      : mb2_detect_describe_view(ctx, OriginalImg.data, OriginalImg.cols, OriginalImg.rows, H, OriginalImg.cols, OriginalImg.rows,
                                                 &det_par.HessParam, &op, &sp, dev_slot, use_slot && ss.count > 0, nullptr, nullptr, nullptr, 0);
        if (n < 0) { if (last_rc == 0) last_rc = n; continue; }  // silent for the class API like the reference (empty lists); LastError() keeps the code
        // host copy first: the device slot (already appended by the call) and the host block must never go out of step
        std::vector<double> dk((size_t)n * MB2_KP), rk((size_t)n * MB2_KP); std::vector<unsigned char> du((size_t)n * 128);
        if (n > 0) {
          const int frc = mb2_view_fetch(ctx, dk.data(), rk.data(), du.data(), n);
          if (frc < 0) { if (last_rc == 0) last_rc = frc; if (use_slot) { ss.desc = curr_desc; ss.count = -1; } continue; }   // count -1: slot unusable, pair_front refuses to match it
        }
        if (use_slot && ss.count >= 0) { ss.desc = curr_desc; ss.count += n; }
        RegionBlock& B = Blocks[curr_det][curr_desc];
        B.det = is_mser ? DET_MSER : DET_HESSIAN; B.desc = dt;
        const size_t base = (size_t)B.n;
        B.det_kp.resize((base + n) * MB2_KP); B.reproj_kp.resize((base + n) * MB2_KP); B.desc_u8.resize((base + n) * 128);
        B.img_id.resize(base + n, (int)synth);
        if (n > 0) {
          std::memcpy(&B.det_kp[base * MB2_KP], dk.data(), dk.size() * sizeof(double)); std::memcpy(&B.reproj_kp[base * MB2_KP], rk.data(), rk.size() * sizeof(double));
          std::memcpy(&B.desc_u8[base * 128], du.data(), du.size());
        }
        B.n = (int)base + n;
        TimeSpent.DetectTime += (now_ms() - t0) / 1000.0;  // detect + orient + describe are one fused call here
      }
    }
  }
}

// ---- feature cache (imagerepresentation.cpp:34-37 saveKP, :89-99 saveAR, :124-147 loadKP / loadAR, :2139-2215) ----------------
void ImageRepresentation::SaveRegions(std::string fname, int mode) const {
  if (mode != 0) return;   // the reference's binary branch is empty (:2141-2143)
  std::ofstream kpfile(fname);
  if (!kpfile.is_open()) { std::cerr << "Cannot open file " << fname << " to save keypoints" << std::endl; return; }
  auto saveKP = [&](const double* v) {   // x y a11 a12 a21 a22 pyramid_scale octave_number s sub_type  (response is not stored)
    kpfile << v[0] << " " << v[1] << " " << v[2] << " " << v[3] << " " << v[4] << " " << v[5] << " ";
    kpfile << 0.0 << " " << 0 << " " << v[6] << " " << (int)v[8] << " ";
  };
  kpfile << Blocks.size() << std::endl;
  for (auto& d : Blocks) {
    kpfile << d.first << " " << d.second.size() << std::endl;
    for (auto& e : d.second) {
      const RegionBlock& B = e.second;
      kpfile << e.first << " " << B.n << std::endl;
      if (B.n > 0) kpfile << 128 << std::endl;
      for (int i = 0; i < B.n; i++) {
        kpfile << i << " " << B.img_id[i] << " " << 0 << " " << 0 << " ";   // id img_id img_reproj_id parent_id
        saveKP(&B.det_kp[(size_t)i * MB2_KP]);
        saveKP(&B.reproj_kp[(size_t)i * MB2_KP]);
        kpfile << " " << 128 << " ";
        for (int j = 0; j < 128; j++) kpfile << (float)B.desc_u8[(size_t)i * 128 + j] << " ";
        kpfile << std::endl;
      }
    }
  }
}
void ImageRepresentation::LoadRegions(std::string fname) {
  std::ifstream kpfile(fname);
  if (!kpfile.is_open()) { std::cerr << "Cannot open file " << fname << " to load keypoints" << std::endl; return; }
  int numberOfDetectors = 0;
  kpfile >> numberOfDetectors;
  for (int det = 0; det < numberOfDetectors && kpfile; det++) {
    std::string det_name; int num_of_descs = 0;
    kpfile >> det_name >> num_of_descs;
    for (int desc = 0; desc < num_of_descs && kpfile; desc++) {
      std::string desc_name; int num_of_kp = 0, desc_size = 0;
      kpfile >> desc_name >> num_of_kp;
      if (num_of_kp > 0) kpfile >> desc_size;
      const descriptor_type dt = GetDescriptorType(desc_name);
      RegionBlock tmp;
      RegionBlock& B = (dt != DESC_UNKNOWN) ? Blocks[det_name][desc_name] : tmp;   // "None" and foreign descriptors are parsed and dropped
      B.det = GetDetectorType(det_name); B.desc = dt;
      for (int kp = 0; kp < num_of_kp && kpfile; kp++) {
        int id, img_id, img_reproj_id, parent_id, size1 = 0;
        kpfile >> id >> img_id >> img_reproj_id >> parent_id;
        double k[2][MB2_KP];
        for (int w = 0; w < 2; w++) {   // loadKP: x y a11 a12 a21 a22 pyramid_scale octave_number s sub_type
          double ps; int oct, st;
          kpfile >> k[w][0] >> k[w][1] >> k[w][2] >> k[w][3] >> k[w][4] >> k[w][5] >> ps >> oct >> k[w][6] >> st;
          k[w][7] = 0; k[w][8] = st;
        }
        kpfile >> size1;
        std::vector<float> vec(size1 > 0 ? size1 : 0);
        for (int j = 0; j < size1; j++) kpfile >> vec[j];
        if (!kpfile) break;
        B.det_kp.insert(B.det_kp.end(), k[0], k[0] + MB2_KP);
        B.reproj_kp.insert(B.reproj_kp.end(), k[1], k[1] + MB2_KP);
        for (int j = 0; j < 128; j++) B.desc_u8.push_back(j < size1 ? (uint8_t)std::max(0.f, std::min(255.f, vec[j])) : 0);
        B.img_id.push_back(img_id);
        B.n++;
      }
    }
  }
  slot_state.clear();   // nothing of this is resident on the device
}

void ImageRepresentation::AppendViewFrom(mb2_ctx* from, const std::string& det, const std::string& desc, int n, int synth) {
  const descriptor_type dt = GetDescriptorType(desc);
  if (dt == DESC_UNKNOWN || n < 0) return;
  SlotState& ss = slot_state[det];
  ss.desc = desc; ss.count += n;
  RegionBlock& B = Blocks[det][desc];
  B.det = GetDetectorType(det); B.desc = dt;
  const size_t base = (size_t)B.n;
  B.det_kp.resize((base + n) * MB2_KP); B.reproj_kp.resize((base + n) * MB2_KP); B.desc_u8.resize((base + n) * 128);
  B.img_id.resize(base + n, synth);
  if (n > 0 && mb2_view_fetch(from, &B.det_kp[base * MB2_KP], &B.reproj_kp[base * MB2_KP], &B.desc_u8[base * 128], n) < 0) n = 0;
  B.n = (int)base + n;
  B.det_kp.resize((size_t)B.n * MB2_KP); B.reproj_kp.resize((size_t)B.n * MB2_KP); B.desc_u8.resize((size_t)B.n * 128); B.img_id.resize(B.n);
}

// ---- matching ---------------------------------------------------------------------------------
namespace {
void fill_tentatives(const double* rows, int n, const AffineRegionList& list1, const AffineRegionList& list2, TentativeCorrespListExt& out) {
  out.TCList.reserve(out.TCList.size() + n);
  auto light = [](const AffineRegion& r) { AffineRegion c; c.img_id = r.img_id; c.img_reproj_id = r.img_reproj_id; c.id = r.id;
                                           c.parent_id = r.parent_id; c.type = r.type; c.det_kp = r.det_kp; c.reproj_kp = r.reproj_kp;
                                           c.desc.type = r.desc.type; return c; };
  for (int i = 0; i < n; i++) {
    const double* r = rows + (size_t)i * 7;
    TentativeCorrespExt t;
    t.first = light(list1[(int)r[0]]);
    t.second = light(list2[(int)r[1]]);
    t.secondbad = light(list2[(int)r[2]]);
    t.secondbadby2ndcl = light(list2[(int)r[3]]);
    t.d1 = r[4]; t.d2 = r[5]; t.d2by2ndcl = r[6];
    t.ratio = std::sqrt((double)((float)r[4] / (float)r[5]));  // sqrt(ratio), ratio = float d0 / float dJ (matching.cpp:435,448)
    out.TCList.push_back(std::move(t));
  }
}
template <class Get1, class Get2>
void fill_tentatives_from(const double* rows, int n, Get1 get1, Get2 get2, TentativeCorrespListExt& out) {
  out.TCList.resize(n);
  for (int i = 0; i < n; i++) {
    const double* r = rows + (size_t)i * 7;
    TentativeCorrespExt& t = out.TCList[i];
    t.first = get1((int)r[0]);
    t.second = get2((int)r[1]);
    t.secondbad = get2((int)r[2]);
    t.secondbadby2ndcl = get2((int)r[3]);
    t.d1 = r[4]; t.d2 = r[5]; t.d2by2ndcl = r[6];
    t.ratio = std::sqrt((double)((float)r[4] / (float)r[5]));
  }
}
}  // namespace

int MatchFlannFGINN(mb2_ctx* ctx, const AffineRegionList& list1, const AffineRegionList& list2, TentativeCorrespListExt& corresp,
                    const MatchPars& par, const int nn) {
  if (list1.empty() || list2.empty()) return 0;
  std::vector<uint8_t> q(list1.size() * 128), t(list2.size() * 128);
  std::vector<double> txy(list2.size() * 2);
  for (size_t i = 0; i < list1.size(); i++) for (int j = 0; j < 128; j++) q[i * 128 + j] = (uint8_t)list1[i].desc.vec[j];
  for (size_t i = 0; i < list2.size(); i++) {
    for (int j = 0; j < 128; j++) t[i * 128 + j] = (uint8_t)list2[i].desc.vec[j];
    txy[2 * i] = list2[i].reproj_kp.x; txy[2 * i + 1] = list2[i].reproj_kp.y;
  }
  std::vector<double> rows(list1.size() * 7);
  int n = mb2_match_fginn(ctx, q.data(), (int)list1.size(), t.data(), (int)list2.size(), txy.data(), par.currMatchRatio, par.contradDist, nn,
                          rows.data(), (int)list1.size());
  if (n < 0) return 0;
  fill_tentatives(rows.data(), n, list1, list2, corresp);
  return n;
}

// matching/matching.cpp:607-666.  Binary descriptors: entries floored to bytes (:631-632), exact Hamming 2-NN on the device,
// kept when d1 <= (int)(float)matchDistanceThreshold; ratio = d1 / d2 in double (:659), no square root here.
int MatchFLANNDistance(mb2_ctx* ctx, const AffineRegionList& list1, const AffineRegionList& list2, TentativeCorrespListExt& corresp,
                       const MatchPars& par, const int /*nn*/) {
  if (list1.empty() || list2.empty()) return 0;
  const size_t dim = list1[0].desc.vec.size();
  std::vector<uint8_t> q(list1.size() * dim), t(list2.size() * dim);
  for (size_t i = 0; i < list1.size(); i++) for (size_t j = 0; j < dim; j++) q[i * dim + j] = (uint8_t)std::floor(list1[i].desc.vec[j]);
  for (size_t i = 0; i < list2.size(); i++) for (size_t j = 0; j < dim; j++) t[i * dim + j] = (uint8_t)std::floor(list2[i].desc.vec[j]);
  std::vector<double> rows(list1.size() * 7);
  const int n = mb2_match_hamming(ctx, q.data(), (int)list1.size(), t.data(), (int)list2.size(), (int)dim, par.matchDistanceThreshold, rows.data(),
                                  (int)list1.size());
  corresp.TCList.clear();
  if (n < 0) return 0;
  corresp.TCList.resize(n);
  for (int i = 0; i < n; i++) {
    const double* r = rows.data() + (size_t)i * 7;
    TentativeCorrespExt& c = corresp.TCList[i];
    c.first = list1[(int)r[0]]; c.second = list2[(int)r[1]];
    c.d1 = r[4]; c.d2 = r[5];
    c.ratio = (double)c.d1 / (double)c.d2;
  }
  return n;
}

int CorrespondenceBank::GetCorrespondencesNumber(std::string desc_name, std::string det_name) const {
  int n = 0;
  for (auto& d : CorrespondencesMapMap) {
    if (desc_name != "All" && d.first != desc_name) continue;
    for (auto& e : d.second) if (det_name == "All" || e.first == det_name) n += (int)e.second.TCList.size();
  }
  return n;
}
TentativeCorrespListExt CorrespondenceBank::GetCorresponcesVector(std::string desc_name, std::string det_name) const {
  TentativeCorrespListExt out;
  for (auto& d : CorrespondencesMapMap) {
    if (desc_name != "All" && d.first != desc_name) continue;
    for (auto& e : d.second)
      if (det_name == "All" || e.first == det_name) out.TCList.insert(out.TCList.end(), e.second.TCList.begin(), e.second.TCList.end());
  }
  return out;
}
TentativeCorrespListExt CorrespondenceBank::TakeCorrespondences() {
  TentativeCorrespListExt out;
  for (auto& d : CorrespondencesMapMap)
    for (auto& e : d.second) {
      if (out.TCList.empty()) out.TCList.swap(e.second.TCList);
      else { out.TCList.insert(out.TCList.end(), std::make_move_iterator(e.second.TCList.begin()), std::make_move_iterator(e.second.TCList.end())); e.second.TCList.clear(); }
    }
  return out;
}
void CorrespondenceBank::ClearCorrespondences(std::string det_name, std::string desc_name) {
  auto d = CorrespondencesMapMap.find(desc_name);
  if (d == CorrespondencesMapMap.end()) return;
  auto e = d->second.find(det_name);
  if (e != d->second.end()) e->second.TCList.clear();
}

int CorrespondenceBank::MatchImgReps(ImageRepresentation& imgrep1, ImageRepresentation& imgrep2, IterationViewsynthesisParam& synth_par,
                                     const WhatToMatch WhatToMatchNow, const MatchPars& par, const DescriptorsParameters&) {
  // correspondencebank.cpp:237-351, "individual detectors" loop (:291-347); the grouped loop (:248-289)
  // matches concatenated lists the same way.
  for (const std::string& curr_det : WhatToMatchNow.separate_detectors) {
    auto vs = synth_par.find(curr_det);
    if (vs == synth_par.end() || vs->second.empty()) continue;
    const ViewSynthParameters& current_VS_params = vs->second[0];
    for (const std::string& curr_desc : WhatToMatchNow.separate_descriptors) {
      ClearCorrespondences(curr_det, curr_desc);
      MatchPars cur = par;
      auto th = current_VS_params.FGINNThreshold.find(curr_desc);
      cur.currMatchRatio = th != current_VS_params.FGINNThreshold.end() ? th->second : 0;
      if (!(cur.currMatchRatio > 0)) continue;
      TentativeCorrespListExt tents;
      auto d1 = imgrep1.Blocks.find(curr_det), d2 = imgrep2.Blocks.find(curr_det);
      if (d1 == imgrep1.Blocks.end() || d2 == imgrep2.Blocks.end()) continue;
      auto b1 = d1->second.find(curr_desc), b2 = d2->second.find(curr_desc);
      if (b1 == d1->second.end() || b2 == d2->second.end()) continue;
      const ImageRepresentation::RegionBlock& Q = b1->second;
      const ImageRepresentation::RegionBlock& T = b2->second;
      if (Q.n == 0 || T.n == 0) continue;
      std::vector<double> rows((size_t)Q.n * 7);
      auto s1 = imgrep1.slot_state.find(curr_det), s2 = imgrep2.slot_state.find(curr_det);
      const bool resident = imgrep1.slot >= 0 && imgrep2.slot >= 0 && s1 != imgrep1.slot_state.end() && s2 != imgrep2.slot_state.end() &&
                            s1->second.desc == curr_desc && s2->second.desc == curr_desc && s1->second.count == Q.n && s2->second.count == T.n;
      int n;
      if (resident) {  // descriptors are still on the device in exactly this order: no re-upload
        n = mb2_match_slots(ctx, imgrep1.slot_of(curr_det), imgrep2.slot_of(curr_det), cur.currMatchRatio, cur.contradDist, 50, rows.data(), Q.n);
      } else {
        std::vector<double> txy((size_t)T.n * 2);
        for (int i = 0; i < T.n; i++) { txy[2 * i] = T.reproj_kp[(size_t)i * MB2_KP]; txy[2 * i + 1] = T.reproj_kp[(size_t)i * MB2_KP + 1]; }
        n = mb2_match_fginn(ctx, Q.desc_u8.data(), Q.n, T.desc_u8.data(), T.n, txy.data(), cur.currMatchRatio, cur.contradDist, 50, rows.data(), Q.n);
      }
      if (n > 0)
        fill_tentatives_from(rows.data(), n, [&](int i) { return Q.region(i, false); }, [&](int i) { return T.region(i, false); }, tents);
      CorrespondencesMapMap[curr_desc][curr_det] = std::move(tents);
    }
  }
  return 0;
}

// ---- DuplicateFiltering (matching.cpp:2983-3047) -------------------------------------------------
// The reference's O(T^2) double loop keeps i and drops every later j whose endpoints are both within r
// of i's.  The kept set is decided greedily in list order, so a uniform grid over the first image's
// coordinates (cell = r) gives the identical result in O(T) (spatial join + greedy pass over the close pairs, below).
// Core on plain arrays: xy = n x (x1 y1 x2 y2), key = sort key (ignored when !sorted).  Returns the kept
// original indices in processing order.
std::vector<int> duplicate_filter_core(const double* xy_in, const double* key, int T, double r, bool sorted) {
  std::vector<int> order(T);
  if (sorted) {
    // std::sort(TCList, CompareCorrespondenceByRatio), matching.cpp:2999: unstable, so the order of EQUAL keys is whatever the library's
    // introsort leaves.  Its moves depend only on the comparison results, not on the element type: the same routine with the same
    // comparator on (key, row) pairs yields the reference's permutation, ties included (checked against the compiled reference).
    std::vector<std::pair<double, int> > ki(T);
    for (int i = 0; i < T; i++) ki[i] = std::make_pair(key[i], i);
    std::sort(ki.begin(), ki.end(), [](const std::pair<double, int>& a, const std::pair<double, int>& b) { return std::fabs(a.first) < std::fabs(b.first); });
    for (int i = 0; i < T; i++) order[i] = ki[i].second;
  } else
    for (int i = 0; i < T; i++) order[i] = i;
  std::vector<double> xy((size_t)T * 4);
  for (int j = 0; j < T; j++) std::memcpy(&xy[4 * (size_t)j], xy_in + 4 * (size_t)order[j], 4 * sizeof(double));
  const double r_sq = r * r;
  // Two phases instead of one hash probe per tentative (random memory accesses: 260 ms for the 419k tentatives of the C4 workload).
  // (1) Spatial join, cell by cell: the tentatives are bucketed by the grid cell (cell >= r) of their first-image point; for every
  //     tentative the EARLIER ones (in processing order) that lie within r in both images are listed -- contiguous buckets, independent
  //     per tentative, so the loop runs on all host threads.
  // (2) The reference's greedy rule in processing order over those short lists: j is dropped iff one of its listed predecessors is kept.
  // The close-pair relation and the order are the reference's, so the kept set is identical.
  // Cell size: any cell >= r keeps every close pair inside a 3 x 3 neighbourhood.  Cells of exactly r make a dense grid of millions
  // of cells for a 4096 x 3072 image (building and scanning it cost 6 of the 6.8 ms this step took for 30k tentatives), so the cell
  // grows until there are about as many cells as tentatives; the distance tests below are unchanged, so is the kept set.
  double cell = r;
  {
    double lox = 0, hix = 0, loy = 0, hiy = 0;
    for (int j = 0; j < T; j++) {
      const double x = xy[4 * (size_t)j], y = xy[4 * (size_t)j + 1];
      if (j == 0) { lox = hix = x; loy = hiy = y; }
      lox = std::min(lox, x); hix = std::max(hix, x); loy = std::min(loy, y); hiy = std::max(hiy, y);
    }
    const double area = (hix - lox + r) * (hiy - loy + r);
    if (T > 0 && area > 0 && std::isfinite(area)) cell = std::max(r, std::sqrt(area / (double)T));
  }
  std::vector<long long> cx(T), cy(T);
  long long minx = 0, miny = 0, maxx = 0, maxy = 0;
  for (int j = 0; j < T; j++) {
    cx[j] = (long long)std::floor(xy[4 * (size_t)j] / cell); cy[j] = (long long)std::floor(xy[4 * (size_t)j + 1] / cell);
    if (j == 0) { minx = maxx = cx[j]; miny = maxy = cy[j]; }
    minx = std::min(minx, cx[j]); maxx = std::max(maxx, cx[j]); miny = std::min(miny, cy[j]); maxy = std::max(maxy, cy[j]);
  }
  const long long gw = T ? maxx - minx + 1 : 1, gh = T ? maxy - miny + 1 : 1;
  std::vector<int> kept;
  kept.reserve(T);
  if (T > 0 && gw * gh <= 64LL * 1000 * 1000) {
    std::vector<int> cell_start((size_t)(gw * gh) + 1, 0), by_cell(T);
    for (int j = 0; j < T; j++) cell_start[(size_t)((cy[j] - miny) * gw + (cx[j] - minx)) + 1]++;
    for (size_t c = 0; c < (size_t)(gw * gh); c++) cell_start[c + 1] += cell_start[c];
    { std::vector<int> cur(cell_start.begin(), cell_start.end() - 1); for (int j = 0; j < T; j++) by_cell[cur[(size_t)((cy[j] - miny) * gw + (cx[j] - minx))]++] = j; }   // ascending j inside a cell
    // coordinates gathered in bucket order: the join then reads contiguous memory (in processing order every candidate was a cache
    // miss: 8 of the 12 ms this function took for 30k tentatives on one core)
    std::vector<double> xyc((size_t)T * 4);
    for (int k = 0; k < T; k++) std::memcpy(&xyc[4 * (size_t)k], &xy[4 * (size_t)by_cell[k]], 4 * sizeof(double));
    // one sweep over the cells, bands of grid rows dealt out to the host pool (parallel_host.hpp); the predecessors of a tentative go to
    // its band's own list, remembered per tentative as (band, offset, count)
    const int BAND = 4;
    const int n_bands = T > 20000 ? (int)((gh + BAND - 1) / BAND) : 1;
    const long long band_rows = n_bands > 1 ? BAND : gh;
    std::vector<std::vector<int> > preds(n_bands);
    std::vector<int> pred_off(T, 0), npred(T, 0), pred_band(T, 0);
    mb2par::parallel_chunks(n_bands, [&](int band) {
      std::vector<int>& mine = preds[band];
      for (long long gy = band * band_rows, gy_end = std::min<long long>(gh, gy + band_rows); gy < gy_end; gy++)
      for (long long gx = 0; gx < gw; gx++) {
        const size_t c0 = (size_t)(gy * gw + gx);
        for (int kj = cell_start[c0]; kj < cell_start[c0 + 1]; kj++) {
          const int j = by_cell[kj];
          const double x1 = xyc[4 * (size_t)kj], y1 = xyc[4 * (size_t)kj + 1], x2 = xyc[4 * (size_t)kj + 2], y2 = xyc[4 * (size_t)kj + 3];
          const size_t first = mine.size();
          for (long long dy = -1; dy <= 1; dy++) {
            const long long yy = gy + dy;
            if (yy < 0 || yy >= gh) continue;
            for (long long dx = -1; dx <= 1; dx++) {
              const long long xx = gx + dx;
              if (xx < 0 || xx >= gw) continue;
              const size_t c = (size_t)(yy * gw + xx);
              for (int k = cell_start[c]; k < cell_start[c + 1]; k++) {
                const int i = by_cell[k];
                if (i >= j) break;                   // ascending inside the cell: only earlier tentatives can drop j
                double ddx = xyc[4 * (size_t)k] - x1, ddy = xyc[4 * (size_t)k + 1] - y1;
                if (ddx * ddx + ddy * ddy > r_sq) continue;
                ddx = xyc[4 * (size_t)k + 2] - x2; ddy = xyc[4 * (size_t)k + 3] - y2;
                if (ddx * ddx + ddy * ddy <= r_sq) mine.push_back(i);
              }
            }
          }
          pred_off[j] = (int)first; npred[j] = (int)(mine.size() - first); pred_band[j] = band;
        }
      }
    });
    std::vector<char> is_kept(T, 0);
    for (int j = 0; j < T; j++) {
      bool dup = false;
      const int* pl = npred[j] ? preds[pred_band[j]].data() + pred_off[j] : nullptr;
      for (int k = 0; k < npred[j] && !dup; k++) dup = is_kept[pl[k]] != 0;
      if (!dup) { is_kept[j] = 1; kept.push_back(order[j]); }
    }
    return kept;
  }
  // (a degenerate coordinate range would make the dense grid too large: the reference's double loop, with early exits)
  std::vector<char> flag(T, 1);
  for (int i = 0; i < T; i++) {
    if (!flag[i]) continue;
    kept.push_back(order[i]);
    for (int j = i + 1; j < T; j++) {
      if (!flag[j]) continue;
      double ddx = xy[4 * (size_t)i] - xy[4 * (size_t)j], ddy = xy[4 * (size_t)i + 1] - xy[4 * (size_t)j + 1];
      if (ddx * ddx + ddy * ddy > r_sq) continue;
      ddx = xy[4 * (size_t)i + 2] - xy[4 * (size_t)j + 2]; ddy = xy[4 * (size_t)i + 3] - xy[4 * (size_t)j + 3];
      if (ddx * ddx + ddy * ddy <= r_sq) flag[j] = 0;
    }
  }
  return kept;
}

void DuplicateFiltering(TentativeCorrespListExt& in_corresp, const double r, const int mode) {
  if (r <= 0) return;
  std::vector<TentativeCorrespExt>& L0 = in_corresp.TCList;
  const int T = (int)L0.size();
  std::vector<double> key(T), xy((size_t)T * 4);
  for (int i = 0; i < T; i++) {
    const TentativeCorrespExt& t = L0[i];
    switch (mode) {
      case MODE_FGINN: key[i] = std::fabs(t.ratio); break;
      case MODE_DISTANCE: key[i] = std::fabs(t.d1); break;
      case MODE_BIGGER_REGION: key[i] = std::fabs(t.first.reproj_kp.s); break;
      default: key[i] = 0.0;
    }
    xy[4 * (size_t)i] = t.first.reproj_kp.x; xy[4 * (size_t)i + 1] = t.first.reproj_kp.y;
    xy[4 * (size_t)i + 2] = t.second.reproj_kp.x; xy[4 * (size_t)i + 3] = t.second.reproj_kp.y;
  }
  const bool sorted = mode == MODE_FGINN || mode == MODE_DISTANCE || mode == MODE_BIGGER_REGION;
  std::vector<int> kept = duplicate_filter_core(xy.data(), key.data(), T, r, sorted);
  std::vector<TentativeCorrespExt> L;
  L.reserve(kept.size());
  for (int i : kept) L.push_back(std::move(L0[i]));
  L0.swap(L);
}

// ---- LORANSACFiltering (matching.cpp:806-980), homography and epipolar modes -------------------------------------
namespace {
bool invert3(const double* S, double* D) {  // cv::invert, 3x3 closed form
  double d = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6]);
  if (d == 0) { for (int i = 0; i < 9; i++) D[i] = 0; return false; }
  d = 1. / d;
  D[0] = (S[4] * S[8] - S[5] * S[7]) * d; D[1] = (S[2] * S[7] - S[1] * S[8]) * d; D[2] = (S[1] * S[5] - S[2] * S[4]) * d;
  D[3] = (S[5] * S[6] - S[3] * S[8]) * d; D[4] = (S[0] * S[8] - S[2] * S[6]) * d; D[5] = (S[2] * S[3] - S[0] * S[5]) * d;
  D[6] = (S[3] * S[7] - S[4] * S[6]) * d; D[7] = (S[1] * S[6] - S[0] * S[7]) * d; D[8] = (S[0] * S[4] - S[1] * S[3]) * d;
  return true;
}
}  // namespace

// Core on plain arrays.  frames: n x 14 doubles = (x y a11 a12 a21 a22 s) of the first and of the second region
// (reproj_kp).  inl[n] = DEGENSAC's inlier flags; verified = indices that survive NaiveHCheck / H_LAF_check.
int loransac_core(mb2_ctx* ctx, const double* frames, int tent_size, const RANSACPars& pars, std::vector<unsigned char>& inl,
                  std::vector<int>& verified, double* H) {
  const int MIN_POINTS = 8;
  inl.assign(std::max(tent_size, 0), 0);
  verified.clear();
  int max_samples = pars.max_samples;
  if (tent_size <= 20) max_samples = 1000;
  if (tent_size < MIN_POINTS) return 0;
  std::vector<double> u((size_t)tent_size * 6);
  for (int i = 0; i < tent_size; i++) {
    const double* f = frames + (size_t)i * 14;
    double* p = &u[(size_t)i * 6];
    p[0] = f[0]; p[1] = f[1]; p[2] = 1.; p[3] = f[7]; p[4] = f[8]; p[5] = 1.;
  }
  double Hloran[9];
  int data_out[4];
  double J = 0;
  const long seed = pars.seed ? pars.seed : (long)time(NULL);
  if (pars.useF) {  // epipolar mode (matching.cpp:875-887, 959-971): exp_ransacFcustom with inlLimit 0, then F_LAF_check
    const int I = mb2_ransac_f(ctx, u.data(), tent_size, pars.err_threshold * pars.err_threshold, pars.confidence, pars.max_samples,
                               pars.errorType == SAMPSON ? 0 : 1, pars.doSymmCheck, pars.localOptimization, 0u, seed, Hloran, inl.data(), data_out, &J);
    if (I < 0) return 0;
    for (int i = 0; i < tent_size; i++) if (inl[i] || pars.justMarkOutliers) verified.push_back(i);
    for (int i = 0; i < 9; i++) H[i] = Hloran[i];
    // F_LAF_check (matching.cpp:193-250): the three points of each local affine frame against F, sum of the three distances
    const double affineFerror = pars.LAFCoef * pars.err_threshold;
    if (affineFerror > 0 && !verified.empty()) {
      const double k_sigma = 3.0;  // matching.cpp:172
      const size_t n = verified.size();
      std::vector<double> u3(n * 18), err(n * 3);
      for (size_t l = 0; l < n; l++) {
        const double* f = frames + (size_t)verified[l] * 14;
        double* q = &u3[l * 18];
        q[0] = f[0]; q[1] = f[1]; q[2] = 1.0;
        q[3] = f[7]; q[4] = f[8]; q[5] = 1.0;
        q[6] = q[0] + k_sigma * f[3] * f[6]; q[7] = q[1] + k_sigma * f[5] * f[6]; q[8] = 1.0;
        q[9] = q[3] + k_sigma * f[10] * f[13]; q[10] = q[4] + k_sigma * f[12] * f[13]; q[11] = 1.0;
        q[12] = q[0] + k_sigma * f[2] * f[6]; q[13] = q[1] + k_sigma * f[4] * f[6]; q[14] = 1.0;
        q[15] = q[3] + k_sigma * f[9] * f[13]; q[16] = q[4] + k_sigma * f[11] * f[13]; q[17] = 1.0;
      }
      if (mb2_score_models(ctx, pars.errorType == SAMPSON ? 3 : 4, u3.data(), (int)(n * 3), Hloran, 1, 0.0, err.data(), nullptr, nullptr) < 0) { verified.clear(); return 0; }
      std::vector<int> good;
      good.reserve(n);
      for (size_t l = 0; l < n; l++) {
        const double sumErr = std::sqrt(err[3 * l]) + std::sqrt(err[3 * l + 1]) + std::sqrt(err[3 * l + 2]);
        if (!(sumErr > affineFerror)) good.push_back(verified[l]);
      }
      verified.swap(good);
    }
    if ((int)verified.size() < MIN_POINTS) verified.clear();
    return (int)verified.size();
  }
  static const bool vtrace = getenv("MB2_VERIFY_TRACE") != nullptr;   // diagnostics: wall time of the legs of the verification stage
  const double tv0 = now_ms();
  int I = mb2_ransac_h(ctx, u.data(), tent_size, pars.err_threshold * pars.err_threshold, pars.confidence, max_samples,
                       (int)pars.errorType, pars.doSymmCheck, seed, Hloran, inl.data(), data_out, &J);
  const double tv1 = now_ms();
  if (I < 0) return 0;
  for (int i = 0; i < tent_size; i++) if (inl[i] || pars.justMarkOutliers) verified.push_back(i);
  // H = inv(Hloran^T) (matching.cpp:920-938)
  const double Ht[9] = {Hloran[0], Hloran[3], Hloran[6], Hloran[1], Hloran[4], Hloran[7], Hloran[2], Hloran[5], Hloran[8]};
  double Hinv[9];
  invert3(Ht, Hinv);
  bool nonzero = false;
  for (int i = 0; i < 9; i++) nonzero = nonzero || (Hinv[i] != 0.0);
  if (!nonzero) { verified.clear(); return 0; }
  for (int i = 0; i < 9; i++) H[i] = Hinv[i];
  {  // NaiveHCheck (matching.cpp:1171-1200), DO_TRANSFER_H_CHECK
    const double err_sq = 10.0 * 10.0;
    double Hi[9];
    invert3(H, Hi);
    int corr_numb = 0;
    for (int i : verified) {
      const double* f = frames + (size_t)i * 14;
      const double x1 = f[0], y1 = f[1], x2 = f[7], y2 = f[8];
      double xa = (H[0] * x1 + H[1] * y1 + H[2]) / (H[6] * x1 + H[7] * y1 + H[8]);
      double ya = (H[3] * x1 + H[4] * y1 + H[5]) / (H[6] * x1 + H[7] * y1 + H[8]);
      const double d1 = (x2 - xa) * (x2 - xa) + (y2 - ya) * (y2 - ya);
      xa = (Hi[0] * x2 + Hi[1] * y2 + Hi[2]) / (Hi[6] * x2 + Hi[7] * y2 + Hi[8]);
      ya = (Hi[3] * x2 + Hi[4] * y2 + Hi[5]) / (Hi[6] * x2 + Hi[7] * y2 + Hi[8]);
      const double d2 = (x1 - xa) * (x1 - xa) + (y1 - ya) * (y1 - ya);
      if ((d1 <= err_sq) && (d2 <= (err_sq))) corr_numb++;
    }
    if (corr_numb < MIN_POINTS) verified.clear();
  }
  const double tv2 = now_ms();
  // H_LAF_check (matching.cpp:251-309): three points per local affine frame scored with HDsSymMax in one batch
  const double affineFerror = 3.0 * pars.HLAFCoef * pars.err_threshold;
  if (affineFerror > 0 && !verified.empty()) {
    const double k_sigma = 3.0;  // matching.cpp:172
    const size_t n = verified.size();
    std::vector<double> u3(n * 18), err(n * 3);
    for (size_t l = 0; l < n; l++) {
      const double* f = frames + (size_t)verified[l] * 14;   // x y a11 a12 a21 a22 s | x y a11 a12 a21 a22 s
      double* q = &u3[l * 18];
      q[0] = f[0]; q[1] = f[1]; q[2] = 1.0;
      q[3] = f[7]; q[4] = f[8]; q[5] = 1.0;
      q[6] = q[0] + k_sigma * f[3] * f[6]; q[7] = q[1] + k_sigma * f[5] * f[6]; q[8] = 1.0;
      q[9] = q[3] + k_sigma * f[10] * f[13]; q[10] = q[4] + k_sigma * f[12] * f[13]; q[11] = 1.0;
      q[12] = q[0] + k_sigma * f[2] * f[6]; q[13] = q[1] + k_sigma * f[4] * f[6]; q[14] = 1.0;
      q[15] = q[3] + k_sigma * f[9] * f[13]; q[16] = q[4] + k_sigma * f[11] * f[13]; q[17] = 1.0;
    }
    if (mb2_score_models(ctx, 2 /*HDsSymMax*/, u3.data(), (int)(n * 3), Hloran, 1, 0.0, err.data(), nullptr, nullptr) < 0) { verified.clear(); return 0; }
    std::vector<int> good;
    good.reserve(n);
    for (size_t l = 0; l < n; l++) {
      const double sumErr = std::sqrt(err[3 * l] + err[3 * l + 1] + err[3 * l + 2]);
      if (!(sumErr > affineFerror)) good.push_back(verified[l]);
    }
    verified.swap(good);
  }
  if (vtrace) fprintf(stderr, "[verify] %d tentatives: mb2_ransac_h %.2f ms, inv(H) + NaiveHCheck %.2f ms, H_LAF_check %.2f ms\n", tent_size, tv1 - tv0, tv2 - tv1, now_ms() - tv2);
  if ((int)verified.size() < MIN_POINTS) verified.clear();
  return (int)verified.size();
}

namespace {
void frame14(const AffineKeypoint& a, const AffineKeypoint& b, double* f) {
  f[0] = a.x; f[1] = a.y; f[2] = a.a11; f[3] = a.a12; f[4] = a.a21; f[5] = a.a22; f[6] = a.s;
  f[7] = b.x; f[8] = b.y; f[9] = b.a11; f[10] = b.a12; f[11] = b.a21; f[12] = b.a22; f[13] = b.s;
}
}  // namespace

int LORANSACFiltering(mb2_ctx* ctx, TentativeCorrespListExt& in_corresp, TentativeCorrespListExt& ransac_corresp, double* H,
                      const RANSACPars pars) {
  const int n = (int)in_corresp.TCList.size();
  ransac_corresp.TCList.clear();
  std::vector<double> frames((size_t)n * 14);
  for (int i = 0; i < n; i++) frame14(in_corresp.TCList[i].first.reproj_kp, in_corresp.TCList[i].second.reproj_kp, &frames[(size_t)i * 14]);
  std::vector<unsigned char> inl;
  std::vector<int> verified;
  double Hout[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  const int k = loransac_core(ctx, frames.data(), n, pars, inl, verified, Hout);
  for (int i = 0; i < n && i < (int)inl.size(); i++) in_corresp.TCList[i].isTrue = inl[i];
  for (int i : verified) ransac_corresp.TCList.push_back(in_corresp.TCList[i]);
  // homography mode writes both; epipolar mode leaves the caller's H alone and stores F in the list (matching.cpp:932-936, 969-970)
  for (int i = 0; i < 9; i++) { ransac_corresp.H[i] = Hout[i]; if (!pars.useF) H[i] = Hout[i]; }
  return k;
}

}  // namespace mods

// ---- test door: DuplicateFiltering on plain arrays (CPU-only logic, used by tests/test_host_logic.py) ------
extern "C" int mb2_host_duplicate_filter(const double* xy /* n x 4: x1 y1 x2 y2 */, const double* ratio, int n, double r, int mode,
                                         int* kept_out /* original indices, in output order */) {
  mods::TentativeCorrespListExt L;
  L.TCList.resize(n);
  for (int i = 0; i < n; i++) {
    mods::TentativeCorrespExt& t = L.TCList[i];
    t.first.reproj_kp.x = xy[4 * i]; t.first.reproj_kp.y = xy[4 * i + 1]; t.second.reproj_kp.x = xy[4 * i + 2]; t.second.reproj_kp.y = xy[4 * i + 3];
    t.ratio = ratio[i]; t.first.id = i;
  }
  mods::DuplicateFiltering(L, r, mode);
  for (size_t i = 0; i < L.TCList.size(); i++) kept_out[i] = L.TCList[i].first.id;
  return (int)L.TCList.size();
}

extern "C" int mb2_host_save_regions(const char* fname, const char* det, const char* desc, int n, const double* det_kp, const double* reproj_kp,
                                     const uint8_t* desc_u8) {
  struct Rep : mods::ImageRepresentation {
    Rep() : mods::ImageRepresentation(nullptr, mods::GrayImage(), "cache", -1) {}
    RegionBlock& blk(const std::string& d, const std::string& e) { return Blocks[d][e]; }
  } rep;
  mods::ImageRepresentation::RegionBlock& B = rep.blk(det, desc);
  B.n = n; B.det = rep.GetDetectorType(det); B.desc = rep.GetDescriptorType(desc);
  B.det_kp.assign(det_kp, det_kp + (size_t)n * MB2_KP); B.reproj_kp.assign(reproj_kp, reproj_kp + (size_t)n * MB2_KP);
  B.desc_u8.assign(desc_u8, desc_u8 + (size_t)n * 128); B.img_id.assign(n, 0);
  rep.SaveRegions(fname);
  return n;
}
extern "C" int mb2_host_load_regions(const char* fname, const char* det, const char* desc, int capacity, double* det_kp, double* reproj_kp,
                                     uint8_t* desc_u8) {
  mods::ImageRepresentation rep(nullptr, mods::GrayImage(), "cache", -1);
  rep.LoadRegions(fname);
  const mods::ImageRepresentation::RegionBlock* B = rep.block(det, desc);
  if (!B) return 0;
  const int m = std::min(B->n, capacity);
  std::memcpy(det_kp, B->det_kp.data(), (size_t)m * MB2_KP * 8); std::memcpy(reproj_kp, B->reproj_kp.data(), (size_t)m * MB2_KP * 8);
  std::memcpy(desc_u8, B->desc_u8.data(), (size_t)m * 128);
  return B->n;
}

extern "C" int mb2_host_set_vs_pars(const double* scales, int n_scales, const double* tilts, int n_tilts, double phi_base, const double* prev,
                                    int n_prev, double* out, int capacity) {
  std::vector<mods::ViewSynthParameters> par, prev_par(n_prev);
  for (int i = 0; i < n_prev; i++) { prev_par[i].zoom = prev[3 * i]; prev_par[i].tilt = prev[3 * i + 1]; prev_par[i].phi = prev[3 * i + 2]; }
  const int n = mods::SetVSPars(std::vector<double>(scales, scales + n_scales), std::vector<double>(tilts, tilts + n_tilts), phi_base, {}, {}, {}, par,
                                prev_par);
  for (int i = 0; i < n && i < capacity; i++) { out[3 * i] = par[i].zoom; out[3 * i + 1] = par[i].tilt; out[3 * i + 2] = par[i].phi; }
  return n;
}

// ---- one MODS iteration on one pair ---------------------------------------------------------------
namespace {
// Helper contexts (own stream + scratch) kept per primary context, created on first use:
//   [0] the second image of a pair (detected / described concurrently with the first),
//   [1] verification (duplicate filter + LO-RANSAC) of pair k while pair k+1 is detected (mb2_mods_pairs),
//   [2] the MSER pass of both images of a pair (one batched detection, mb2_mser_detect_pair), next to the two HessianAffine passes,
//   [3] orientation + description of the second image's MSER regions while [2] does the first image's.
std::mutex g_sib_mutex;
//   [4], [5] further MSER contexts: with MB2_MSER_AHEAD=n the dataset call (mb2_mods_pairs) runs the MSER detection of the next n pairs
//   while the current pair is in its HessianAffine / description stage; a detection owns its context until it is described.
//   [6] the primary context of the SECOND LANE of the dataset call (mb2_mods_pairs runs two pairs' front stages side by side); it has
//   helper contexts [0], [2], [3], [5] of its own.
struct Helpers { mb2_ctx* c[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; bool tried[7] = {false, false, false, false, false, false, false}; };
std::unordered_map<mb2_ctx*, Helpers> g_siblings;
mb2_ctx* sibling_ctx(mb2_ctx* ctx, int which = 0) {
  std::lock_guard<std::mutex> lk(g_sib_mutex);
  Helpers& h = g_siblings[ctx];
  if (!h.tried[which]) {
    h.tried[which] = true;
    // [2] and [3] carry the MSER chain, the critical path of a pair: their kernels go ahead of the HessianAffine ones
    static const bool prio = getenv("MB2_NO_PRIO") == nullptr;
    if (mb2_ctx_create_prio(mb2_ctx_device(ctx), (which >= 2 && prio) ? 1 : 0, &h.c[which]) != MB2_OK) h.c[which] = nullptr;
  }
  return h.c[which];
}
}  // namespace

// device staging of mb2_mods_pairs' host images: two buffer pairs per primary context, grow-only, released by mb2_mods_release
namespace {
struct Staged { void* d[2] = {nullptr, nullptr}; size_t cap[2] = {0, 0}; const float* use[2] = {nullptr, nullptr}; int rc = MB2_OK; mb2_ctx* owner = nullptr; };
std::unordered_map<mb2_ctx*, Staged*> g_staged;
Staged* staged_of(mb2_ctx* ctx) {
  std::lock_guard<std::mutex> lk(g_sib_mutex);
  Staged*& s = g_staged[ctx];
  if (!s) s = new Staged[2];
  return s;
}
}  // namespace

// helper context `which` of a primary context (same device, its own stream); created on first use, released by mb2_mods_release
extern "C" mb2_ctx* mb2_mods_sibling(mb2_ctx* ctx, int which) { return (ctx && which >= 0 && which < 6) ? sibling_ctx(ctx, which) : nullptr; }

extern "C" long long mb2_mods_launch_count(mb2_ctx* ctx) {
  std::lock_guard<std::mutex> lk(g_sib_mutex);
  long long n = mb2_ctx_launch_count(ctx);
  auto it = g_siblings.find(ctx);
  if (it != g_siblings.end()) {
    for (mb2_ctx* c : it->second.c) if (c) n += mb2_ctx_launch_count(c);
    if (mb2_ctx* lane = it->second.c[6]) {   // the second lane's own helpers
      auto jt = g_siblings.find(lane);
      if (jt != g_siblings.end()) for (mb2_ctx* c : jt->second.c) if (c) n += mb2_ctx_launch_count(c);
    }
  }
  return n;
}
// mods_sharded.cpp: the view-sharded driver's device scratch of this context (weak: the CPU test harness builds this file alone)
extern "C" void mb2_sharded_release(mb2_ctx* ctx) __attribute__((weak));
extern "C" void mb2_mods_release(mb2_ctx* ctx) {
  if (mb2_sharded_release) mb2_sharded_release(ctx);
  mb2_ctx* lane = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_sib_mutex);
    auto it = g_siblings.find(ctx);
    if (it != g_siblings.end()) lane = it->second.c[6];
  }
  if (lane) mb2_mods_release(lane);   // the second lane's helpers and staging buffers first (the lane context itself goes with ctx's helpers below)
  std::lock_guard<std::mutex> lk(g_sib_mutex);
  {
    auto st = g_staged.find(ctx);
    if (st != g_staged.end()) {
      for (int b = 0; b < 2; b++) for (int im = 0; im < 2; im++) if (st->second[b].d[im]) mb2_dev_free(st->second[b].owner, st->second[b].d[im]);
      delete[] st->second;
      g_staged.erase(st);
    }
  }
  auto it = g_siblings.find(ctx);
  if (it == g_siblings.end()) return;
  for (mb2_ctx* c : it->second.c) if (c) mb2_ctx_destroy(c);
  g_siblings.erase(it);
}

extern "C" void mb2_pair_config_default(mb2_pair_config* c) {
  mods::DetectorsParameters dp; mods::DescriptorsParameters sp;
  c->det = dp.HessParam;
  c->ori = mb2_orientation_params{1.0, 41, 1, 0.8, 0, 0};
  c->desc = sp.RootSIFTParam;
  c->matchRatio = 0.8; c->contradDist = 30.0; c->duplicateDist = 2.0;              // iters_mods_cviu.ini:62, config_iter_mods_cviu.ini:151,157
  c->err_threshold = 3.0; c->confidence = 0.99; c->HLAFCoef = 12.0;                // config_iter_mods_cviu.ini:163-172
  c->max_samples = 100000; c->errorType = 0; c->doSymmCheck = 1;
  c->seed = 1;
  c->use_mser = 0; c->mser = dp.MSERParam; c->mserMatchRatio = 0.8;                 // iters_mods_cviu.ini:36 ([MSER0] FGINNThreshold)
  c->n_hess_views = c->n_mser_views = 0;
  std::memset(c->hess_views, 0, sizeof c->hess_views); std::memset(c->mser_views, 0, sizeof c->mser_views);
  c->halfRootSIFT = 0;
  c->useF = 0; c->localOptimization = 1; c->LAFCoef = 2.0;                           // config_iter_mods_cviu.ini:168-169
}

namespace {
using namespace mods;

struct PairSetup {  // what getCLIparam would read from config_iter_mods_cviu.ini / an iters file with one HessianAffine tier
  DetectorsParameters det_par; DescriptorsParameters desc_par; DominantOrientationParams dom;
  std::string desc_name; IterationViewsynthesisParam iters; RANSACPars rp;
  // descriptors of every tier: {desc_name}, or {"RootSIFT", "HalfRootSIFT"} for the WxBS tiers (iters_mods_cviu_wxbs.ini:35,48,61)
  std::vector<std::string> descs;
  bool mser_identity_only = true;   // the batched two-image MSER pass covers the identity view only
  explicit PairSetup(const mb2_pair_config* cfg) {
    det_par.HessParam = cfg->det;
    desc_par.RootSIFTParam = cfg->desc; desc_par.SIFTParam = cfg->desc;
    desc_par.HalfRootSIFTParam = cfg->desc; desc_par.HalfRootSIFTParam.rootSIFT = 1; desc_par.HalfRootSIFTParam.doHalfSIFT = 1;
    desc_par.HalfSIFTParam = desc_par.HalfRootSIFTParam;
    dom.maxAngles = cfg->ori.maxAngles; dom.threshold = (float)cfg->ori.threshold; dom.mrSize = cfg->ori.mrSize; dom.patchSize = cfg->ori.patchSize;
    desc_name = cfg->desc.rootSIFT ? "RootSIFT" : "SIFT";
    descs.push_back(desc_name);
    if (cfg->halfRootSIFT && cfg->desc.rootSIFT) descs.push_back("HalfRootSIFT");
    auto tier = [&](const char* det, double ratio, int n, const mb2_view_params* views) {
      auto add = [&](ViewSynthParameters v) { for (const std::string& d : descs) { v.descriptors.push_back(d); v.FGINNThreshold[d] = ratio; } iters[det].push_back(v); };
      if (n <= 0) { add(ViewSynthParameters()); return; }
      for (int i = 0; i < n && i < MB2_MAX_PAIR_VIEWS; i++) {
        ViewSynthParameters v; v.tilt = views[i].tilt; v.phi = views[i].phi; v.zoom = views[i].zoom; v.InitSigma = views[i].InitSigma; v.doBlur = views[i].doBlur;
        add(v);
      }
    };
    tier("HessianAffine", cfg->matchRatio, cfg->n_hess_views, cfg->hess_views);
    mser_identity_only = true;
    if (cfg->use_mser) {
      det_par.MSERParam = cfg->mser;
      tier("MSER", cfg->mserMatchRatio, cfg->n_mser_views, cfg->mser_views);
      const std::vector<ViewSynthParameters>& mv = iters["MSER"];
      mser_identity_only = mv.size() == 1 && std::fabs(mv[0].tilt - 1.) <= 0.1 && std::fabs(mv[0].phi) <= 0.2 && std::fabs(mv[0].zoom - 1.) <= 0.1 &&
                           cfg->mser.mode == 0 &&   // the batched pass is FIXED_TH only (the other modes sort on the host per image)
                           descs.size() == 1;       // ... and describes one descriptor with plain orientations
    }
    rp.err_threshold = cfg->err_threshold; rp.confidence = cfg->confidence; rp.max_samples = cfg->max_samples;
    rp.HLAFCoef = cfg->HLAFCoef; rp.errorType = (RANSAC_error_t)cfg->errorType; rp.doSymmCheck = cfg->doSymmCheck; rp.seed = cfg->seed;
    rp.useF = cfg->useF; rp.localOptimization = cfg->localOptimization; rp.LAFCoef = cfg->LAFCoef;
  }
};

// Everything the verification stage needs from the detection / matching stage, on the host.
struct PairFront {
  std::shared_ptr<ImageRepresentation> rep1, rep2;   // shared: the 1-to-N caller (mb2_mods_multi) keeps one query image for all pairs
  // tentatives per (descriptor, detector), in the order GetCorresponcesVector("All", "All") concatenates them
  // (CorrespondencesMapMap[desc][det], std::map order: "HalfRootSIFT" < "RootSIFT", "HessianAffine" < "MSER")
  struct Group { const char* det; std::string desc; std::vector<double> rows; int nt = 0; };
  Group groups[4];
  int n_groups = 0, nt = 0, rc = MB2_OK;
  int slot1 = 0, slot2 = 1;   // device slots of the HessianAffine sets of the two images (MSER: + 2)
  double t_start = 0;
};

// mods.cpp:229-330 for one pair: SynthDetectDescribeKeypoints of both images on two host threads
// (mods.cpp:255-271 runs them as two OpenMP tasks), each with its own context / stream, then MatchImgReps.
// A MSER detection of both images of a pair started ahead of its pair_front (mb2_mods_pairs, MB2_MSER_AHEAD): mb2_mser_detect_pair
// running on context `c` in its own thread.
struct MserAhead {
  mb2_ctx* c = nullptr; std::thread th; int rc = MB2_OK, n1 = 0, n2 = 0; double ms = 0;
  ~MserAhead() { if (th.joinable()) th.join(); }
};

// MatchImgReps of the pair (correspondencebank.cpp:237-351): every (descriptor, detector) group separately; fills out.groups / res.
void match_reps(mb2_ctx* ctx, const mb2_pair_config* cfg, PairSetup& ps, mb2_pair_result* res, PairFront& out) {
  double t0;
  // a view call that failed (CUDA error, capacity, tap-table overflow) fails the pair: no success with partial regions
  if (out.rep1->LastError() < 0) { out.rc = out.rep1->LastError(); return; }
  if (out.rep2->LastError() < 0) { out.rc = out.rep2->LastError(); return; }
  res->regions1 = out.rep1->GetDescriptorsNumber(ps.desc_name); res->regions2 = out.rep2->GetDescriptorsNumber(ps.desc_name);
  res->mser_regions1 = out.rep1->GetDescriptorsNumber(ps.desc_name, "MSER"); res->mser_regions2 = out.rep2->GetDescriptorsNumber(ps.desc_name, "MSER");
  t0 = now_ms();
  const char* dets[2] = {"HessianAffine", "MSER"};
  std::vector<std::string> desc_order = ps.descs;
  std::sort(desc_order.begin(), desc_order.end());
  for (const std::string& desc : desc_order)
    for (int g = 0; g < (cfg->use_mser ? 2 : 1); g++) {   // MatchImgReps, separate detectors x separate descriptors (correspondencebank.cpp:291-347)
      PairFront::Group& G = out.groups[out.n_groups];
      G.det = dets[g]; G.desc = desc;
      const ImageRepresentation::RegionBlock* Q = out.rep1->block(dets[g], desc);
      const ImageRepresentation::RegionBlock* T = out.rep2->block(dets[g], desc);
      out.n_groups++;
      if (!(Q && T && Q->n > 0 && T->n > 0)) continue;
      G.rows.resize((size_t)Q->n * 7);
      const double ratio = g == 0 ? cfg->matchRatio : cfg->mserMatchRatio;
      // the first descriptor of the tier is the one left resident on the device; the slots must then hold exactly the host blocks'
      // regions (the same check MatchImgReps makes: count == Q.n).  Regions read from a feature cache live on the host only.
      const int sq = out.rep1->SlotCount(dets[g]), st = out.rep2->SlotCount(dets[g]);
      if (desc == ps.desc_name && (sq < 0 || st < 0 || (sq > 0 && sq != Q->n) || (st > 0 && st != T->n))) { out.rc = MB2_ERR_CUDA; return; }
      if (desc == ps.desc_name && sq == Q->n && st == T->n)
        G.nt = mb2_match_slots(ctx, g == 0 ? out.slot1 : out.slot1 + 2, g == 0 ? out.slot2 : out.slot2 + 2, ratio, cfg->contradDist, 50, G.rows.data(), Q->n);
      else {
        std::vector<double> txy((size_t)T->n * 2);
        for (int i = 0; i < T->n; i++) { txy[2 * i] = T->reproj_kp[(size_t)i * MB2_KP]; txy[2 * i + 1] = T->reproj_kp[(size_t)i * MB2_KP + 1]; }
        G.nt = mb2_match_fginn(ctx, Q->desc_u8.data(), Q->n, T->desc_u8.data(), T->n, txy.data(), ratio, cfg->contradDist, 50, G.rows.data(), Q->n);
      }
      if (G.nt < 0) { out.rc = G.nt; G.nt = 0; return; }
      out.nt += G.nt;
      if (g == 1) res->mser_tentatives += G.nt;
    }
  res->ms_match = now_ms() - t0;
  res->tentatives = out.nt;
}


void pair_front(mb2_ctx* ctx, const float* img1, int w1, int h1, const float* img2, int w2, int h2, const mb2_pair_config* cfg,
                PairSetup& ps, mb2_pair_result* res, PairFront& out, MserAhead* pre = nullptr) {
  out.t_start = now_ms();
  mb2_ctx* ctx2 = mb2_ctx_profiling(ctx) ? nullptr : sibling_ctx(ctx, 0);   // per-kernel profiling keeps everything on one stream
  out.rep1.reset(new ImageRepresentation(ctx, GrayImage{img1, h1, w1}, "img1", 0));
  out.rep2.reset(new ImageRepresentation(ctx2 ? ctx2 : ctx, GrayImage{img2, h2, w2}, "img2", 1));
  double t0 = now_ms();
  // MSER of both images in ONE batched pass on a third context when the images have the same size (the component-tree
  // kernel is latency bound: two images cost hardly more than one); otherwise per image, after HessianAffine.
  mb2_ctx* ctx3 = (ctx2 && cfg->use_mser && ps.mser_identity_only && w1 == w2 && h1 == h2) ? (pre ? pre->c : sibling_ctx(ctx, 2)) : nullptr;
  IterationViewsynthesisParam iters_here = ps.iters;
  if (ctx3) {
    iters_here.erase("MSER");
    for (auto* r : {out.rep1.get(), out.rep2.get()}) { r->Prepare("HessianAffine", ps.desc_name); r->Prepare("MSER", ps.desc_name); }
  }
  int mser_rc = MB2_OK;
  auto mser_post = [&](mb2_ctx* run, int which, int* rc_out) {   // orientation + description of one image's MSER regions
    const double Hid[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    mb2_sift_params sp = cfg->desc.rootSIFT ? ps.desc_par.RootSIFTParam : ps.desc_par.SIFTParam;
    mb2_orientation_params op{ps.dom.mrSize, ps.dom.patchSize, ps.dom.maxAngles, (double)ps.dom.threshold, 0, 0};
    const int k = mb2_describe_view_of_pair(run, ctx3, which, Hid, w1, h1, &op, &sp, 2 + which, 0, nullptr, nullptr, nullptr, 0);
    if (k < 0) { *rc_out = k; return; }
    (which ? out.rep2 : out.rep1)->AppendViewFrom(run, "MSER", ps.desc_name, k, 0);
  };
  static const bool timing = getenv("MB2_PAIR_TIMING") != nullptr;   // diagnostics: where the front stage spends its time
  double tm_detect = 0, tm_post = 0, th1 = 0, th2 = 0;
  auto mser_pair = [&] {
    int n1 = 0, n2 = 0;
    const double ta = now_ms();
    if (pre) {   // started while the previous pair was in its HessianAffine / description stage
      if (pre->th.joinable()) pre->th.join();
      if ((mser_rc = pre->rc) < 0) return;
      n1 = pre->n1; n2 = pre->n2;
      tm_detect = pre->ms;
    } else {
      if ((mser_rc = mb2_mser_detect_pair(ctx3, img1, img2, w1, h1, &ps.det_par.MSERParam, &n1, &n2)) < 0) return;
      tm_detect = now_ms() - ta;
    }
    struct PostTimer { double& t; double t0; ~PostTimer() { t = now_ms() - t0; } } post_timer{tm_post, now_ms()};
    mser_rc = MB2_OK;
    // the two images are finished side by side: image 1 here, image 2 on a fourth context
    mb2_ctx* ctx4 = sibling_ctx(ctx, 3);
    int rc2 = MB2_OK;
    if (ctx4) {
      std::thread t2([&] { mser_post(ctx4, 1, &rc2); });
      mser_post(ctx3, 0, &mser_rc);
      t2.join();
      if (rc2 < 0) mser_rc = rc2;
      if (mser_rc >= 0 && mb2_slot_move(ctx3, 3, ctx4, 3) < 0) mser_rc = MB2_ERR_CUDA;
    } else { mser_post(ctx3, 0, &mser_rc); if (mser_rc >= 0) mser_post(ctx3, 1, &mser_rc); }
  };
  if (ctx2) {
    // The tree kernel of the MSER pass goes first and alone: the HessianAffine streams are ordered behind it (they then overlap
    // with the rest of the MSER pass), because next to their bandwidth-hungry kernels the latency-bound tree kernel takes 2-3x longer.
    static const bool tree_first_env = getenv("MB2_NO_TREE_FIRST") == nullptr;
    const bool tree_first = tree_first_env && !pre;   // a detection started ahead belongs to another time window: nothing to order behind
    const long long epoch0 = ctx3 ? mb2_ctx_tree_epoch(ctx3) : 0;
    auto wait_tree = [&](mb2_ctx* c) {
      if (!ctx3 || !tree_first) return;
      const double t_give_up = now_ms() + 200.0;   // the MSER thread reaches the launch within a few ms; never block for good
      while (mb2_ctx_tree_epoch(ctx3) == epoch0 && mser_rc >= 0 && now_ms() < t_give_up) std::this_thread::sleep_for(std::chrono::microseconds(50));
      if (mb2_ctx_tree_epoch(ctx3) != epoch0) mb2_ctx_wait_tree(c, ctx3);
    };
    std::thread tm;
    if (ctx3) tm = std::thread(mser_pair);
    std::thread th([&] { wait_tree(ctx2); const double ta = now_ms(); out.rep2->SynthDetectDescribeKeypoints(iters_here, ps.det_par, ps.desc_par, ps.dom); th2 = now_ms() - ta; });
    { wait_tree(ctx); const double ta = now_ms(); out.rep1->SynthDetectDescribeKeypoints(iters_here, ps.det_par, ps.desc_par, ps.dom); th1 = now_ms() - ta; }
    th.join();
    if (ctx3) tm.join();
    if (timing) fprintf(stderr, "[pair front] hessaff img1 %.1f ms, img2 %.1f ms | mser detect (both) %.1f ms, orient+describe %.1f ms\n", th1, th2, tm_detect, tm_post);
    if (mser_rc < 0) { out.rc = mser_rc; return; }
    if (mb2_slot_move(ctx, 1, ctx2, 1) < 0) { out.rc = MB2_ERR_CUDA; return; }
    if (ctx3) { if (mb2_slot_move(ctx, 2, ctx3, 2) < 0 || mb2_slot_move(ctx, 3, ctx3, 3) < 0) { out.rc = MB2_ERR_CUDA; return; } }
    else if (cfg->use_mser && mb2_slot_move(ctx, 3, ctx2, 3) < 0) { out.rc = MB2_ERR_CUDA; return; }
  } else {
    out.rep1->SynthDetectDescribeKeypoints(ps.iters, ps.det_par, ps.desc_par, ps.dom);
    out.rep2->SynthDetectDescribeKeypoints(ps.iters, ps.det_par, ps.desc_par, ps.dom);
  }
  res->ms_detect_describe = now_ms() - t0;
  match_reps(ctx, cfg, ps, res, out);
}

// mods.cpp:330-415: DuplicateFiltering(MODE_FGINN) + LORANSACFiltering on the region blocks and index lists
// directly (the AoS TentativeCorrespListExt of the reference is only materialised by the class API).
int pair_back(mb2_ctx* vctx, const mb2_pair_config* cfg, PairSetup& ps, PairFront& in, mb2_pair_result* res, double* verified_out, int capacity) {
  int n = 0;
  const int nt = in.nt;
  if (nt > 0) {
    double t0 = now_ms();
    // tentatives["All"] = GetCorresponcesVector() (mods.cpp:298): the per-detector lists one after the other
    std::vector<double> xy((size_t)nt * 4), key(nt);
    std::vector<const double*> fa(nt), fb(nt);   // reproj_kp records of the two regions of every tentative
    int o = 0;
    for (int g = 0; g < in.n_groups; g++) {
      const PairFront::Group& G = in.groups[g];
      if (G.nt <= 0) continue;
      const ImageRepresentation::RegionBlock* Q = in.rep1->block(G.det, G.desc);
      const ImageRepresentation::RegionBlock* T = in.rep2->block(G.det, G.desc);
      for (int i = 0; i < G.nt; i++, o++) {
        const double* r = &G.rows[(size_t)i * 7];
        const double* a = &Q->reproj_kp[(size_t)r[0] * MB2_KP];
        const double* b = &T->reproj_kp[(size_t)r[1] * MB2_KP];
        fa[o] = a; fb[o] = b;
        xy[4 * (size_t)o] = a[0]; xy[4 * (size_t)o + 1] = a[1]; xy[4 * (size_t)o + 2] = b[0]; xy[4 * (size_t)o + 3] = b[1];
        key[o] = std::fabs(std::sqrt((double)((float)r[4] / (float)r[5])));
      }
    }
    std::vector<int> kept = duplicate_filter_core(xy.data(), key.data(), nt, cfg->duplicateDist, true);
    res->ms_duplicate = now_ms() - t0;
    res->unique_tentatives = (int)kept.size();
    t0 = now_ms();
    std::vector<double> frames(kept.size() * 14);
    for (size_t i = 0; i < kept.size(); i++) {
      const double* a = fa[kept[i]];
      const double* b = fb[kept[i]];
      double* f = &frames[i * 14];
      for (int j = 0; j < 7; j++) { f[j] = a[j]; f[7 + j] = b[j]; }   // KP layout starts with x y a11 a12 a21 a22 s
    }
    std::vector<unsigned char> inl;
    std::vector<int> verified;
    n = loransac_core(vctx, frames.data(), (int)kept.size(), ps.rp, inl, verified, res->H);
    res->ms_ransac = now_ms() - t0;
    for (unsigned char b : inl) res->ransac_inliers += b;
    res->verified = n;
    if (verified_out)
      for (int i = 0; i < n && i < capacity; i++) {
        const double* f = &frames[(size_t)verified[i] * 14];
        verified_out[4 * i] = f[0]; verified_out[4 * i + 1] = f[1]; verified_out[4 * i + 2] = f[7]; verified_out[4 * i + 3] = f[8];
      }
  }
  res->ms_total = res->ms_detect_describe + res->ms_match + res->ms_duplicate + res->ms_ransac;
  return n;
}
}  // namespace

extern "C" int mb2_host_verify(mb2_ctx* ctx, const double* frames14, const double* key, int n, const mb2_pair_config* cfg, mb2_pair_result* res,
                               double* verified_out, int capacity) {
  if (!ctx || !cfg || !res || n < 0 || (n > 0 && (!frames14 || !key))) return MB2_ERR_ARG;
  PairSetup ps(cfg);
  std::vector<double> xy((size_t)n * 4);
  for (int i = 0; i < n; i++) {
    const double* f = frames14 + (size_t)i * 14;
    xy[4 * (size_t)i] = f[0]; xy[4 * (size_t)i + 1] = f[1]; xy[4 * (size_t)i + 2] = f[7]; xy[4 * (size_t)i + 3] = f[8];
  }
  double t0 = now_ms();
  std::vector<int> kept = duplicate_filter_core(xy.data(), key, n, cfg->duplicateDist, true);
  res->ms_duplicate = now_ms() - t0;
  res->tentatives = n; res->unique_tentatives = (int)kept.size();
  t0 = now_ms();
  std::vector<double> frames(kept.size() * 14);
  for (size_t i = 0; i < kept.size(); i++) std::memcpy(&frames[i * 14], frames14 + (size_t)kept[i] * 14, 14 * sizeof(double));
  std::vector<unsigned char> inl;
  std::vector<int> verified;
  const int k = loransac_core(ctx, frames.data(), (int)kept.size(), ps.rp, inl, verified, res->H);
  res->ms_ransac = now_ms() - t0;
  res->ransac_inliers = 0;
  for (unsigned char b : inl) res->ransac_inliers += b;
  res->verified = k;
  if (verified_out)
    for (int i = 0; i < k && i < capacity; i++) {
      const double* f = &frames[(size_t)verified[i] * 14];
      verified_out[4 * i] = f[0]; verified_out[4 * i + 1] = f[1]; verified_out[4 * i + 2] = f[7]; verified_out[4 * i + 3] = f[8];
    }
  return k;
}

// ---- other callers of the same path: 1-to-N matching and the feature cache ---------------------------------------
// mods_multi.cpp:232-330: the query image is detected / described ONCE, then every other image is detected, matched against it,
// duplicate-filtered and verified.  res[i] / verified_out[i] as mb2_mods_pair returns them for the pair (img1, imgs2[i]).
extern "C" int mb2_mods_multi(mb2_ctx* ctx, const float* img1, int w1, int h1, int n, const float* const* imgs2, const int* w2, const int* h2,
                              const mb2_pair_config* cfg, mb2_pair_result* res, double* const* verified_out, const int* capacity) {
  if (!ctx || !img1 || !cfg || !res || n < 0 || (n > 0 && (!imgs2 || !w2 || !h2))) return MB2_ERR_ARG;
  PairSetup ps(cfg);
  std::shared_ptr<ImageRepresentation> rep1(new ImageRepresentation(ctx, GrayImage{img1, h1, w1}, "img1", 0));
  const double t0 = now_ms();
  rep1->SynthDetectDescribeKeypoints(ps.iters, ps.det_par, ps.desc_par, ps.dom);   // mods_multi.cpp:244-247
  const double t_img1 = now_ms() - t0;
  for (int i = 0; i < n; i++) {
    std::memset(&res[i], 0, sizeof res[i]);
    PairFront f;
    f.t_start = now_ms();
    f.rep1 = rep1;
    f.rep2.reset(new ImageRepresentation(ctx, GrayImage{imgs2[i], h2[i], w2[i]}, "img2", 1));
    f.rep2->SynthDetectDescribeKeypoints(ps.iters, ps.det_par, ps.desc_par, ps.dom);   // :252-256
    res[i].ms_detect_describe = now_ms() - f.t_start + (i == 0 ? t_img1 : 0.0);
    match_reps(ctx, cfg, ps, &res[i], f);                                              // :275-277
    if (f.rc < 0) return f.rc;
    const int k = pair_back(ctx, cfg, ps, f, &res[i], verified_out ? verified_out[i] : nullptr, capacity ? capacity[i] : 0);   // :291-330
    if (k < 0) return k;
    res[i].ms_total = now_ms() - f.t_start;
  }
  return n;
}

// extract_features.cpp: SynthDetectDescribeKeypoints + SaveRegions -- the feature cache the reference's `read_pre_extracted` flow
// (mods.cpp:224-240) and build/read_features.m read.  Returns the number of described regions written.
extern "C" int mb2_extract_features(mb2_ctx* ctx, const float* img, int w, int h, const mb2_pair_config* cfg, const char* fname) {
  if (!ctx || !img || !cfg || !fname) return MB2_ERR_ARG;
  PairSetup ps(cfg);
  ImageRepresentation rep(ctx, GrayImage{img, h, w}, "img", -1);
  rep.SynthDetectDescribeKeypoints(ps.iters, ps.det_par, ps.desc_par, ps.dom);
  if (rep.LastError() < 0) return rep.LastError();
  rep.SaveRegions(fname, 0);
  return rep.GetDescriptorsNumber();
}

// mods.cpp:224-240 + 290-415 on two feature caches: LoadRegions for both images, then MatchImgReps -> DuplicateFiltering ->
// LORANSACFiltering.  The descriptors are uploaded for the matching (loaded regions are not resident on the device).
extern "C" int mb2_mods_pair_cached(mb2_ctx* ctx, const char* cache1, const char* cache2, const mb2_pair_config* cfg, mb2_pair_result* res,
                                    double* verified_out, int capacity) {
  if (!ctx || !cache1 || !cache2 || !cfg || !res) return MB2_ERR_ARG;
  std::memset(res, 0, sizeof *res);
  PairSetup ps(cfg);
  PairFront f;
  f.t_start = now_ms();
  f.rep1.reset(new ImageRepresentation(ctx, GrayImage{nullptr, 0, 0}, "img1", -1));
  f.rep2.reset(new ImageRepresentation(ctx, GrayImage{nullptr, 0, 0}, "img2", -1));
  f.rep1->LoadRegions(cache1); f.rep2->LoadRegions(cache2);
  match_reps(ctx, cfg, ps, res, f);
  if (f.rc < 0) return f.rc;
  const int k = pair_back(ctx, cfg, ps, f, res, verified_out, capacity);
  res->ms_total = now_ms() - f.t_start;
  return k;
}

extern "C" int mb2_mods_pair(mb2_ctx* ctx, const float* img1, int w1, int h1, const float* img2, int w2, int h2,
                             const mb2_pair_config* cfg, mb2_pair_result* res, double* verified_out, int capacity) {
  if (!ctx || !img1 || !img2 || !cfg || !res) return MB2_ERR_ARG;
  std::memset(res, 0, sizeof *res);
  PairSetup ps(cfg);
  PairFront f;
  pair_front(ctx, img1, w1, h1, img2, w2, h2, cfg, ps, res, f);
  if (f.rc < 0) return f.rc;
  const int n = pair_back(ctx, cfg, ps, f, res, verified_out, capacity);
  res->ms_total = now_ms() - f.t_start;
  return n;
}

// A list of independent pairs (a dataset run: mods.cpp is started once per pair by the EVD / WxBS scripts),
// software-pipelined over two host threads: while pair k is being verified (duplicate filter + LO-RANSAC on
// the helper context [1]), pair k+1 is already in detection / description / matching on the primary one.
// Host images of the next pair are uploaded on the copy engine meanwhile; for images up to 4 Mpx two pairs' front stages run side by side
// (two "lanes", each on its own set of contexts: one such pair does not fill the GPU).
// Results are identical to calling mb2_mods_pair on every pair in turn.
extern "C" int mb2_mods_pairs(mb2_ctx* ctx, int n_pairs, const float* const* img1, const int* w1, const int* h1, const float* const* img2,
                              const int* w2, const int* h2, const mb2_pair_config* cfg, mb2_pair_result* res, double* const* verified_out,
                              const int* capacity) {
  if (!ctx || n_pairs < 0 || !cfg || !res || (n_pairs > 0 && (!img1 || !img2 || !w1 || !h1 || !w2 || !h2))) return MB2_ERR_ARG;
  mb2_ctx* vctx = mb2_ctx_profiling(ctx) ? nullptr : sibling_ctx(ctx, 1);
  PairSetup ps(cfg);
  if (!vctx) {  // no helper context: plain loop
    for (int k = 0; k < n_pairs; k++) {
      int r = mb2_mods_pair(ctx, img1[k], w1[k], h1[k], img2[k], w2[k], h2[k], cfg, &res[k], verified_out ? verified_out[k] : nullptr,
                            capacity ? capacity[k] : 0);
      if (r < 0) return r;
    }
    return n_pairs;
  }
  std::mutex m;
  std::condition_variable cv;
  std::deque<std::pair<int, std::unique_ptr<PairFront> > > q;
  bool done = false;
  int rc = MB2_OK;
  std::thread back([&] {
    PairSetup ps_back(cfg);
    for (;;) {
      std::pair<int, std::unique_ptr<PairFront> > item;
      {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return done || !q.empty(); });
        if (q.empty()) return;
        item = std::move(q.front()); q.pop_front();
      }
      cv.notify_all();
      const int k = item.first;
      int r = pair_back(vctx, cfg, ps_back, *item.second, &res[k], verified_out ? verified_out[k] : nullptr, capacity ? capacity[k] : 0);
      if (r < 0) { std::lock_guard<std::mutex> lk(m); rc = r; }
    }
  });
  // MB2_MSER_AHEAD=n (0..2): the MSER detection (component tree: latency bound, ~10 % SM use) of the next n pairs runs while the
  // current pair is in its HessianAffine / description stage, each on its own context.  Same results, different schedule.
  int ahead = 0;
  if (const char* e = getenv("MB2_MSER_AHEAD")) ahead = std::max(0, std::min(2, atoi(e)));
  bool same_size = true;
  for (int k = 0; k < n_pairs; k++) same_size = same_size && w1[k] == w2[k] && h1[k] == h2[k];
  if (!(cfg->use_mser && ps.mser_identity_only && same_size && sibling_ctx(ctx, 0))) ahead = 0;
  mb2_ctx* pool[3] = {nullptr, nullptr, nullptr};
  if (ahead > 0) {
    static const int which[3] = {2, 4, 5};
    for (int i = 0; i <= ahead; i++) if (!(pool[i] = sibling_ctx(ctx, which[i]))) ahead = 0;
  }
  std::deque<std::unique_ptr<MserAhead> > inflight;
  int next_launch = 0;
  auto launch_ahead = [&](int k) {
    std::unique_ptr<MserAhead> a(new MserAhead);
    a->c = pool[k % (ahead + 1)];
    MserAhead* p = a.get();
    p->th = std::thread([p, k, &ps, img1, img2, w1, h1] {
      const double t0 = now_ms();
      p->rc = mb2_mser_detect_pair(p->c, img1[k], img2[k], w1[k], h1[k], &ps.det_par.MSERParam, &p->n1, &p->n2);
      p->ms = now_ms() - t0;
    });
    inflight.push_back(std::move(a));
  };
  // LANES (MB2_LANES=1|2; default: 2 for images up to 4 Mpx, else 1): with 2, the front stages of pairs k and k + 1 run side by side, each
  // lane on its own set of contexts (lane 1: helper context [6] as its primary, with helpers of its own).  Results do not depend on the
  // schedule (every pair is computed by the same calls on private contexts); verification of finished pairs stays on the one helper
  // thread below.  Measured on one B200 (tools/e2e_probe.py): 1920 x 1080 pairs (BASELINE config 5's size) 7.0 -> 5.9 ms per pair
  // (142 -> 168 pairs/s) -- one pair's kernels do not fill the GPU and its host legs leave gaps; 4096 x 3072 pairs 31.6 -> 33.7 ms -- the
  // three chains of ONE pair already keep the GPU throughput bound there, a second pair's kernels only slow the first pair's down.
  int n_lanes = 1;
  {
    long long max_px = 0;
    for (int k = 0; k < n_pairs; k++) max_px = std::max(max_px, std::max((long long)w1[k] * h1[k], (long long)w2[k] * h2[k]));
    if (n_pairs > 0 && max_px <= 4000000LL) n_lanes = 2;
  }
  if (const char* e = getenv("MB2_LANES")) n_lanes = std::max(1, std::min(2, atoi(e)));
  if (ahead > 0 || n_pairs < 2) n_lanes = 1;
  mb2_ctx* lane_ctx[2] = {ctx, n_lanes > 1 ? sibling_ctx(ctx, 6) : nullptr};
  if (n_lanes > 1 && (!lane_ctx[1] || !sibling_ctx(lane_ctx[1], 0))) n_lanes = 1;
  std::atomic<bool> failed(false);
  // Host images: the two images of a lane's next pair travel to the device (copy engine, own stream) while its current pair is computed,
  // into one of two buffer pairs per lane; the front stage then starts from device-resident images.  (Staged inside the view calls, every
  // context uploaded its own copy -- HessianAffine and MSER contexts: 200 MB per pair -- ahead of its first kernel.)
  auto run_lane = [&](const int lane) {
    mb2_ctx* lctx = lane_ctx[lane];
    PairSetup ps_lane(cfg);
    mb2_ctx* cctx = ahead == 0 ? sibling_ctx(lctx, 5) : nullptr;
    Staged* staged = staged_of(lctx);   // two buffer pairs, kept per context between calls (cudaMalloc / cudaFree of 200 MB per call otherwise)
    auto upload = [&](int k) {
      Staged& S = staged[(k / n_lanes) & 1];
      S.rc = MB2_OK;
      const double t_up = now_ms();
      for (int im = 0; im < 2; im++) {
        const float* src = im ? img2[k] : img1[k];
        S.use[im] = src;
        if (!cctx || mb2_is_device_pointer(src)) continue;
        const size_t bytes = (size_t)(im ? w2[k] : w1[k]) * (im ? h2[k] : h1[k]) * 4;
        if (bytes > S.cap[im]) {
          if (S.d[im]) mb2_dev_free(cctx, S.d[im]);
          S.d[im] = mb2_dev_alloc(cctx, bytes); S.cap[im] = S.d[im] ? bytes : 0; S.owner = cctx;
          if (!S.d[im]) { S.rc = MB2_ERR_CUDA; return; }
        }
        if (mb2_dev_copy(cctx, S.d[im], src, bytes, 0) != MB2_OK) { S.rc = MB2_ERR_CUDA; return; }
        S.use[im] = (const float*)S.d[im];
      }
      if (cctx && mb2_ctx_sync(cctx) != MB2_OK) S.rc = MB2_ERR_CUDA;
      static const bool timing = getenv("MB2_PAIR_TIMING") != nullptr;
      if (timing) fprintf(stderr, "[pair upload] pair %d: %.2f ms\n", k, now_ms() - t_up);
    };
    auto fail = [&](int code) { std::lock_guard<std::mutex> lk(m); if (rc >= 0) rc = code; failed = true; };
    std::thread up;
    if (lane < n_pairs) upload(lane);
    for (int k = lane; k < n_pairs && !failed; k += n_lanes) {
      std::memset(&res[k], 0, sizeof res[k]);
      std::unique_ptr<PairFront> f(new PairFront);
      std::unique_ptr<MserAhead> pre;
      if (ahead > 0) {   // (single lane only)
        while (next_launch < n_pairs && next_launch <= k + ahead) launch_ahead(next_launch++);
        pre = std::move(inflight.front()); inflight.pop_front();
      }
      if (up.joinable()) up.join();                       // pair k is on the device
      Staged& S = staged[(k / n_lanes) & 1];
      if (S.rc < 0) { fail(S.rc); break; }
      if (k + n_lanes < n_pairs) up = std::thread(upload, k + n_lanes);   // the lane's next pair follows while pair k is computed (the other buffer pair)
      pair_front(lctx, S.use[0], w1[k], h1[k], S.use[1], w2[k], h2[k], cfg, ps_lane, &res[k], *f, pre.get());
      pre.reset();
      if (f->rc < 0) { fail(f->rc); break; }
      std::unique_lock<std::mutex> lk(m);
      cv.wait(lk, [&] { return q.size() < 2; });   // at most two pairs waiting for verification
      q.emplace_back(k, std::move(f));
      lk.unlock();
      cv.notify_all();
    }
    if (up.joinable()) up.join();
  };
  std::thread lane1;
  if (n_lanes > 1) lane1 = std::thread(run_lane, 1);
  run_lane(0);
  if (lane1.joinable()) lane1.join();
  inflight.clear();   // joins detections that were started for pairs an error kept us from reaching
  { std::lock_guard<std::mutex> lk(m); done = true; }
  cv.notify_all();
  back.join();
  return rc < 0 ? rc : n_pairs;
}
