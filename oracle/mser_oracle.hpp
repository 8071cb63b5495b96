// TEST INFRASTRUCTURE (oracle) -- CPU restatement of the reference's MSER detector.
// NOT shipped, NOT linked by the product.  Sequential, like the reference: pixels in increasing intensity
// (raster order within a level), union-find with the reference's survivor rules.  Pinned bit for bit against
// the reference's own MSER sources compiled in place (oracle/_ref, tests/test_oracle_vs_ref.py).
//
//   DetectMSERs (6-arg)            detectors/mser/extrema/extrema.cpp:284-473
//   getRLEExtrema                  detectors/mser/extrema/libExtrema.cpp:462-482
//   CalcHistogram / BinSortPixels  detectors/mser/extrema/sortPixels.cpp:62-128, invert :131-153
//   GetExtrema & friends           detectors/mser/extrema/getExtrema.cpp:103-437
//   FastSetOptThresholds4Stable... detectors/mser/extrema/optThresh.cpp:15-166
//   RegionBoundaries               detectors/mser/extrema/boundary.cpp:107-207  (result restated: row runs of the
//                                  4-connected component {I <= thresh} that contains the region's first pixel)
//   RLE2Ellipse                    detectors/mser/extrema/libExtrema.cpp:117-159
//   Matrix2::schur_sym / sqrt      detectors/mser/utls/matrix.cpp:185-217, 163-169
#ifndef MB2_MSER_ORACLE_HPP
#define MB2_MSER_ORACLE_HPP

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace mo {
namespace mser {

struct Params {  // extremaParams.h:56-93 + config_iter_mods_cviu.ini [MSER]
  double max_area = 0.05;
  int min_size = 30;
  double min_margin = 8;
  int mode = 0;  // FIXED_TH
  int reg_number = -1;
  float rel_threshold = -1, rel_reg_number = -1;
};

struct Thresh { int thresh, pos, margin; };
struct Run { int line, col1, col2; };
struct OutRegion {  // libExtrema.h:42-75 (Region + RLERegion), one per (region, threshold)
  int polarity;     // 0 = MSER+, 1 = MSER- (inverted image)
  int minI, maxI, threshold, margin, area, border;
  int birth_level, seed;  // bookkeeping for debugging (seed = padded offset of the region's first pixel)
  double cx, cy, sxx, sxy, syy;
  std::vector<Run> rle;
};

struct RegionRec {  // extremaTypes.h:47-59 (t_region)
  int minimum_int, maximum_int, pixel_total, border_total;
  uint32_t seed;
  bool alive;
  std::vector<Thresh> th;
  int pixels[256], borders[256];
};

// label slot kinds (getExtrema.cpp:19-36: pointer / min_reg / region encoded in the two low bits)
enum { UNLABELLED = 0, POINTER = 1, MINREG = 2, REGION = 3 };

struct Extrema {
  int W, H, cols;  // cols = W + 2: padded label / intensity images (preprocess.cpp:22: BAry(-1,h,-1,w))
  std::vector<uint8_t> img;
  std::vector<uint8_t> kind;
  std::vector<uint64_t> val;  // POINTER: slot index; MINREG: packed counters (size<<2 | border<<17); REGION: region index
  std::vector<RegionRec*> regions;
  int min_size, min_size_int, max_size;
  double min_margin;
  // GetLabelled state
  int labelled[4], label_num, border_num;

  ~Extrema() { for (auto* r : regions) delete r; }

  int findRoot(int slot) {  // FindEquivLabel, getExtrema.cpp:187-214 (with the same flattening)
    int root = (int)val[slot];
    if (kind[root] != POINTER) return root;
    do root = (int)val[root]; while (kind[root] == POINTER);
    int p = slot;
    while (kind[p] == POINTER) { int nx = (int)val[p]; val[p] = (uint64_t)root; p = nx; }
    return root;
  }
  void getLabelled(int ofs) {  // getExtrema.cpp:216-263
    const int nb[4] = {ofs - cols, ofs - 1, ofs + 1, ofs + cols};
    label_num = 0; border_num = 0;
    for (int k = 0; k < 4; k++) {
      int l = nb[k];
      if (kind[l] == UNLABELLED) continue;
      if (kind[l] == POINTER) l = findRoot(l);
      bool dup = false;
      for (int j = 0; j < label_num; j++) dup |= (labelled[j] == l);
      // the reference compares l2 with l1, l3 with l1,l2, ... where an unlabelled li keeps its own (distinct) address
      if (!dup) labelled[label_num++] = l;
      border_num++;
    }
    border_num *= 2;
  }
  RegionRec* upgradeRegion(int slot, int intensity) {  // getExtrema.cpp:103-141
    RegionRec* r = new RegionRec;
    uint64_t mr = val[slot];
    r->pixel_total = (int)((mr & 0x1fffcull) >> 2);
    r->border_total = (int)(mr >> 17);
    r->seed = (uint32_t)slot;
    r->minimum_int = r->maximum_int = intensity;
    std::memset(r->pixels, 0, sizeof(r->pixels)); std::memset(r->borders, 0, sizeof(r->borders));
    r->pixels[intensity] = r->pixel_total; r->borders[intensity] = r->border_total;
    r->alive = true;
    regions.push_back(r);
    kind[slot] = REGION; val[slot] = (uint64_t)(regions.size() - 1);
    return r;
  }
  void insMarkPixel(int root, int ofs, int intensity) {  // getExtrema.cpp:144-171
    kind[ofs] = POINTER; val[ofs] = (uint64_t)root;
    if (kind[root] == MINREG) {
      val[root] += 0x00080004ull - ((uint64_t)border_num << 17);
      if ((int)(val[root] & 0x1fffcull) >= min_size_int) upgradeRegion(root, intensity);
    } else {
      RegionRec* r = regions[val[root]];
      r->maximum_int = intensity;
      r->pixel_total++; r->border_total += 4 - border_num;
      r->pixels[intensity]++; r->borders[intensity] += 4 - border_num;
    }
  }
  void mergeRegions(int ofs, int intensity) {  // getExtrema.cpp:267-361
    unsigned maxSize = 0; int maxLabel = labelled[0], num_large = 0;
    for (int i = 0; i < label_num; i++)
      if (kind[labelled[i]] == REGION) {
        RegionRec* r = regions[val[labelled[i]]];
        unsigned size = (unsigned)(r->pixel_total - r->pixels[intensity]);
        num_large++;
        if (size > maxSize) { maxSize = size; maxLabel = labelled[i]; }
      }
    if (!num_large) {
      for (int i = 1; i < label_num; i++) {
        val[maxLabel] += val[labelled[i]];  // both are packed min_reg counters (flag bits masked in the reference)
        kind[labelled[i]] = POINTER; val[labelled[i]] = (uint64_t)maxLabel;
      }
    } else {
      bool max_has_minstats = kind[maxLabel] == MINREG;
      RegionRec* maxRegion = max_has_minstats ? nullptr : regions[val[maxLabel]];
      for (int i = 0; i < label_num; i++) {
        int label = labelled[i];
        if (label == maxLabel) continue;
        bool merging_min_reg = kind[label] == MINREG;
        uint64_t mr = val[label];
        RegionRec* region = merging_min_reg ? nullptr : regions[mr];
        kind[label] = POINTER; val[label] = (uint64_t)maxLabel;
        int pixel_total, border_total;
        if (merging_min_reg) { pixel_total = (int)((mr & 0x1fffcull) >> 2); border_total = (int)(mr >> 17); }
        else { pixel_total = region->pixel_total; border_total = region->border_total; }
        if (max_has_minstats) {  // :322: the sum is formed in (wrapping) 32-bit int, then widened to the 64-bit label
          int packed = (int)(((uint32_t)pixel_total << 2) + ((uint32_t)border_total << 17));
          val[maxLabel] += (uint64_t)(int64_t)packed;
        }
        else {
          maxRegion->pixel_total += pixel_total; maxRegion->border_total += border_total;
          maxRegion->pixels[intensity] += pixel_total; maxRegion->borders[intensity] += border_total;
        }
        if (!merging_min_reg) {
          if ((intensity - region->minimum_int + 1) <= min_margin) region->alive = false;
          else {
            region->maximum_int = intensity;
            setOptThresholds(region);
            if (region->th.empty()) region->alive = false;
          }
        }
      }
    }
    insMarkPixel(maxLabel, ofs, intensity);
  }

  void setOptThresholds(RegionRec* r) {  // optThresh.cpp:69-166
    if (r->pixel_total < min_size) return;
    int* cA = r->pixels; int* cB = r->borders;
    for (int i = r->minimum_int + 1; i <= r->maximum_int; i++) { cA[i] += cA[i - 1]; cB[i] += cB[i - 1]; }
    int up, localMaxMargin = -1, localMaxPos = -1;
    int i = r->minimum_int;
    auto emit = [&]() {
      Thresh t; t.thresh = localMaxPos + localMaxMargin / 2;
      if (cA[t.thresh] <= max_size && cA[t.thresh] > min_size) { t.pos = localMaxPos; t.margin = localMaxMargin; r->th.push_back(t); }
    };
    do {
      int area_i = cA[i], radius_i = cB[i];
      up = (int)(i + min_margin);
      if (up > r->maximum_int) break;
      while ((cA[up] - area_i < radius_i) && (up < r->maximum_int)) up++;
      int margin = up - i;
      double quality = (double)margin;
      if (quality > min_margin && margin >= localMaxMargin) { localMaxMargin = margin; localMaxPos = i; }
      else {
        if (localMaxPos >= 0) { emit(); localMaxPos = -1; }
        localMaxMargin = margin;
      }
      i++;
    } while (up < r->maximum_int);
    if (localMaxPos >= 0) emit();
    // SuppresOverlappingTresholds4StableRegions, optThresh.cpp:15-65
    std::vector<Thresh>& T = r->th;
    for (int k = 0; k < (int)T.size(); k++) {
      while (k >= 0 && k + 1 < (int)T.size()) {
        Thresh& a = T[k]; Thresh& b = T[k + 1];
        if ((a.pos + a.margin < b.thresh) && (a.thresh < b.pos)) break;
        if (b.margin <= a.margin) T.erase(T.begin() + k + 1);
        else { T.erase(T.begin() + k); k--; break; }
      }
    }
    for (int k = 0; k < (int)T.size(); k++) {
      while (k + 1 < (int)T.size()) {
        Thresh& a = T[k]; Thresh& b = T[k + 1];
        if (a.pos + a.margin < b.pos) break;
        if (cA[b.thresh] - cA[a.thresh] <= 0.1 * cA[a.thresh]) {
          a.margin = b.pos - a.pos + b.margin; a.thresh = a.pos + a.margin / 2;
          T.erase(T.begin() + k + 1);
        } else break;
      }
    }
  }

  // GetExtrema, getExtrema.cpp:390-437, on the padded u8 image `img`
  void run(double max_area) {
    size_t n = (size_t)(H + 2) * cols;
    kind.assign(n, UNLABELLED); val.assign(n, 0);
    min_size_int = std::min(10000, min_size) * 4;
    max_size = (int)((double)W * (double)H * max_area);  // PrepareThresholds :377: (cols-2)*(rows-2) int product * max_area
    // counting sort of padded offsets, raster order inside a level (sortPixels.cpp:62-128)
    std::vector<uint32_t> start(257, 0);
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) start[img[(size_t)(y + 1) * cols + x + 1] + 1]++;
    for (int i = 0; i < 256; i++) start[i + 1] += start[i];
    std::vector<uint32_t> order((size_t)W * H), cur(start.begin(), start.end() - 1);
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) { uint32_t o = (uint32_t)((y + 1) * cols + x + 1); order[cur[img[o]]++] = o; }
    for (int level = 0; level < 256; level++)
      for (uint32_t k = start[level]; k < start[level + 1]; k++) {
        int ofs = (int)order[k];
        getLabelled(ofs);
        if (label_num == 0) { kind[ofs] = MINREG; val[ofs] = 0x00080004ull; }  // ConsRegion :174-178
        else if (label_num == 1) insMarkPixel(labelled[0], ofs, level);
        else mergeRegions(ofs, level);
      }
    int root = cols + 1;
    if (kind[root] == POINTER) root = findRoot(root);
    if (kind[root] == REGION) setOptThresholds(regions[val[root]]);
  }

  // RegionBoundaries + OutputRLEAndEll: row runs (raster order) of the component at each selected threshold
  void output(int polarity, std::vector<OutRegion>& out) {
    std::vector<uint32_t> stamp((size_t)(H + 2) * cols, 0), stack;
    uint32_t cur = 0;
    for (RegionRec* r : regions) {
      if (!r->alive || r->th.empty()) continue;
      for (const Thresh& t : r->th) {
        OutRegion o; o.polarity = polarity; o.minI = r->minimum_int; o.maxI = r->maximum_int; o.threshold = t.thresh; o.margin = t.margin;
        o.area = r->pixels[t.thresh]; o.border = r->borders[t.thresh]; o.birth_level = r->minimum_int; o.seed = (int)r->seed;
        cur++;
        int y0 = H + 2, y1 = -1, x0 = cols, x1 = -1;
        stack.clear(); stack.push_back(r->seed); stamp[r->seed] = cur;
        while (!stack.empty()) {
          uint32_t p = stack.back(); stack.pop_back();
          int y = p / cols, x = p % cols;
          y0 = std::min(y0, y); y1 = std::max(y1, y); x0 = std::min(x0, x); x1 = std::max(x1, x);
          const uint32_t nb[4] = {p + cols, p - cols, p + 1, p - 1};
          for (uint32_t q : nb) {
            int qy = q / cols, qx = q % cols;
            if (qy < 1 || qy > H || qx < 1 || qx > W) continue;
            if (stamp[q] == cur || img[q] > t.thresh) continue;
            stamp[q] = cur; stack.push_back(q);
          }
        }
        for (int y = y0; y <= y1; y++) {
          int x = x0;
          while (x <= x1) {
            if (stamp[(size_t)y * cols + x] != cur) { x++; continue; }
            int xs = x;
            while (x <= x1 && stamp[(size_t)y * cols + x] == cur) x++;
            o.rle.push_back(Run{y - 1, xs - 1, x - 2});
          }
        }
        rle2Ellipse(o);
        out.push_back(std::move(o));
      }
    }
  }
  static void rle2Ellipse(OutRegion& o) {  // libExtrema.cpp:117-159
    double area = 0, sumX = 0, sumY = 0;
    for (const Run& r : o.rle) {
      double line = r.line, m = r.col1, n = 1 + r.col2;
      sumX += (n * n - m * m) / 2;
      sumY += (n - m) * (2 * line + 1) / 2;
      area += n - m;
    }
    o.cx = sumX / area; o.cy = sumY / area;
    double sumX2 = 0, sumY2 = 0, sumXY = 0;
    for (const Run& r : o.rle) {
      double line = r.line - o.cy, m = r.col1 - o.cx, n = 1 + r.col2 - o.cx;
      double l2 = line * line, m2 = m * m, n2 = n * n;
      sumX2 += (n2 * n - m2 * m) / 3;
      sumY2 += (n - m) * (3 * l2 + 3 * line + 1) / 3;
      sumXY += -.25 * (m2 - n2) * (2 * line + 1);
    }
    o.sxx = sumX2 / area; o.syy = sumY2 / area; o.sxy = sumXY / area;
  }
};

// getRLEExtrema(params, image): MSER+ on the image, then MSER- on 255 - image (libExtrema.cpp:462-482)
inline std::vector<OutRegion> rleExtrema(const uint8_t* gray, int w, int h, const Params& par, double min_margin_eff) {
  std::vector<OutRegion> out;
  for (int pol = 0; pol < 2; pol++) {
    Extrema e; e.W = w; e.H = h; e.cols = w + 2;
    e.img.assign((size_t)(h + 2) * e.cols, 0);
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
      uint8_t v = gray[(size_t)y * w + x];
      e.img[(size_t)(y + 1) * e.cols + x + 1] = pol ? (uint8_t)(255 - v) : v;
    }
    e.min_size = par.min_size; e.min_margin = min_margin_eff;
    e.run(par.max_area);
    e.output(pol, out);
  }
  return out;
}

struct MKey { double x, y, a11, a12, a21, a22, s, response; int sub_type; };

// sqrtm of the covariance: C.schur_sym(U,T); A = U * T.sqrt() * U.transpose()  (extrema.cpp:417-427, matrix.cpp)
inline void ellipseToA(double sxx, double sxy, double syy, double* A) {
  double r, t;
  if (sxy != 0) {
    r = (syy - sxx) / (2 * sxy);
    if (r >= 0) t = 1.0 / (r + std::sqrt(1 + r * r)); else t = -1.0 / (-r + std::sqrt(1 + r * r));
    r = 1.0 / std::sqrt(1 + t * t);
    t = t * r;
  } else { r = 1; t = 0; }
  const double Q[4] = {r, t, -t, r}, Qt[4] = {r, -t, t, r}, C[4] = {sxx, sxy, sxy, syy};
  auto mul = [](const double* a, const double* b, double* c) {
    c[0] = a[0] * b[0] + a[1] * b[2]; c[1] = a[0] * b[1] + a[1] * b[3];
    c[2] = a[2] * b[0] + a[3] * b[2]; c[3] = a[2] * b[1] + a[3] * b[3];
  };
  double M1[4], T[4], M2[4];
  mul(Qt, C, M1); mul(M1, Q, T);
  T[1] = 0; T[2] = 0;
  const double S[4] = {std::sqrt(T[0]), std::sqrt(T[1]), std::sqrt(T[2]), std::sqrt(T[3])};
  mul(Q, S, M2); mul(M2, Qt, A);
}

inline bool marginCompareInvOrder(const MKey& a, const MKey& b) { return std::fabs(a.response) > std::fabs(b.response); }

// prepareKeysForExport, extrema.cpp:31-90 (std::sort is the reference's own, unstable, sort)
inline void prepareKeysForExport(std::vector<MKey>& keys, const Params& par, int reg_number) {
  if (keys.empty() || par.mode == 0) return;
  std::sort(keys.begin(), keys.end(), marginCompareInvOrder);
  double maxResponse = std::fabs(keys[0].response);
  int regNumber = (int)keys.size();
  switch (par.mode) {
    case 1: {  // RELATIVE_TH
      MKey tmp = keys[0]; tmp.response = maxResponse * par.rel_threshold;
      keys.resize(std::lower_bound(keys.begin(), keys.end(), tmp, marginCompareInvOrder) - keys.begin());
      break; }
    case 2: if (reg_number < regNumber && reg_number >= 0) keys.resize(reg_number); break;  // FIXED_REG_NUMBER
    case 3: keys.resize((int)std::floor(par.rel_reg_number * (double)keys.size())); break;     // RELATIVE_REG_NUMBER
    case 4: {  // NOT_LESS_THAN_REGIONS
      MKey tmp = keys[0]; tmp.response = par.min_margin;
      int fix = (int)(std::lower_bound(keys.begin(), keys.end(), tmp, marginCompareInvOrder) - keys.begin());
      if (fix < reg_number) keys.resize(std::min(reg_number, regNumber)); else keys.resize(std::min(fix, regNumber));
      break; }
    default: break;
  }
}

// DetectMSERs (6-arg, doOnNormal branch), extrema.cpp:284-473.  `regions_out` (optional) receives the raw region list.
inline std::vector<MKey> detectMSERs(const float* img, int w, int h, Params par, double tilt, double zoom,
                                     std::vector<OutRegion>* regions_out = nullptr) {
  int reg_number = par.reg_number;
  if ((tilt > 2.0) || (zoom < 0.5)) reg_number = (int)std::floor(zoom * 2.0 * reg_number / tilt);
  double finalThreshold = par.mode != 0 ? 1.0 : par.min_margin;
  std::vector<uint8_t> gray((size_t)w * h);
  for (size_t i = 0; i < gray.size(); i++) gray[i] = (unsigned char)img[i];
  std::vector<OutRegion> regs = rleExtrema(gray.data(), w, h, par, finalThreshold);
  std::vector<MKey> keys; keys.reserve(regs.size());
  for (const OutRegion& r : regs) {
    MKey k; double A[4];
    ellipseToA(r.sxx, r.sxy, r.syy, A);
    k.x = r.cx; k.y = r.cy; k.a11 = A[0]; k.a12 = A[1]; k.a21 = A[2]; k.a22 = A[3]; k.s = 1.0;
    k.response = r.margin; k.sub_type = r.polarity ? 20 : 21;
    keys.push_back(k);
  }
  Params pe = par; pe.min_margin = finalThreshold;
  prepareKeysForExport(keys, pe, reg_number);
  if (regions_out) *regions_out = std::move(regs);
  return keys;
}

}  // namespace mser
}  // namespace mo
#endif
