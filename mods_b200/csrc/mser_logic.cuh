// MSER on a canonical component tree -- the per-node / per-region decision logic.
//
// The reference (detectors/mser/extrema/getExtrema.cpp:103-437) grows components pixel by pixel in
// (intensity, raster) order with a union-find.  Everything it outputs is a function of the component tree of the
// level sets {I <= t}, except three order-dependent details, which are replayed exactly (sequentially, but only
// inside the one tree node concerned):
//   * which of several equally large tracked regions survives a merge      (MergeRegions, :267-289, first-in-
//     neighbour-order among equals),
//   * when, inside a level, a growing component is first promoted to a tracked region -- this fixes the position
//     of the region in the output list                                      (InsMarkPixel/UpgradeRegion, :103-171),
//   * regions born and re-born inside one level (a min_reg label absorbing a freshly promoted region, :311-324).
//
// Tree representation (built by mser.cu on the GPU, by tests/native/mser_tree_cpu.cpp on the host for the logic
// tests): one node per (component, level) that owns at least one pixel of that level; the node is named by its
// representative pixel `rep`, some pixel of that level in the component (which one does not matter);
//   parent[x] = rep of x's node if x is not a rep, else rep of the parent node (root: itself);
//   area[rep] = pixels in the component; nedge[rep] = 4-adjacent pixel pairs inside the component, so that the
//   reference's border count (+4 - 2*labelled neighbours per pixel, :150-168) is 4*area - 2*nedge.
//
// This header is plain C++ (no CUDA intrinsics): MB2_HD functions are compiled for the device by mser.cu and for the
// host by the logic tests.  It is not a CPU fallback of the product: nothing in libmods_b200.so calls it on the host.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define MB2_HD __host__ __device__ __forceinline__
#else
#define MB2_HD inline
#endif

namespace mser_logic {

static const uint32_t NONE = 0xffffffffu;

// Several images of the same size may be stacked into one pixel index space (mser.cu stacks MSER+ on top of MSER-, the
// inverted image): pixel p lies in image p / (W*H), and rows of different images are never neighbours.  Every image has
// its own root (the only pixels with parent[x] == x).
struct Tree {
  int W, H;                // size of ONE image
  const uint8_t* lev;      // intensity (255 - I in the MSER- image)
  const uint32_t* parent;
  const uint32_t* area;
  const uint32_t* nedge;
  int track_size;          // min(10000, min_size): promotion threshold (PrepareThresholds, getExtrema.cpp:372-381)
};

MB2_HD bool is_root(const Tree& t, uint32_t x) { return t.parent[x] == x; }
MB2_HD bool is_rep(const Tree& t, uint32_t x) { const uint32_t p = t.parent[x]; return p == x || t.lev[p] > t.lev[x]; }
MB2_HD uint32_t node_of(const Tree& t, uint32_t x) { return is_rep(t, x) ? x : t.parent[x]; }
MB2_HD int border_of(const Tree& t, uint32_t rep) { return (int)(4u * t.area[rep] - 2u * t.nedge[rep]); }
MB2_HD bool tracked(const Tree& t, uint32_t rep) { return (int)t.area[rep] >= t.track_size; }

// child of node v whose subtree contains pixel q (q must lie strictly below v)
MB2_HD uint32_t child_containing(const Tree& t, uint32_t v, uint32_t q) {
  uint32_t u = node_of(t, q);
  while (t.parent[u] != v) u = t.parent[u];
  return u;
}

// ---- replay of one node's level (getExtrema.cpp: GetLabelled :216-263, ProcessPixel :363-375) -----------------
// Elements of the small union-find: own pixels (index = pixel) and children (index = N + child rep, state initialised
// lazily); all arrays hold 2N entries.  A pixel is an own pixel of exactly one node and a child of at most one, so nodes
// replayed concurrently never share an element.
struct EmuScratch {
  uint32_t N;
  uint32_t* uf;       // NONE = untouched
  uint32_t* size;     // pixels in the set (root only)
  uint32_t* presize;  // tracked sets: size before this level (0 for a region promoted inside the level)
  uint32_t* ident;    // tracked sets: child rep that carries the region, NONE for a region born in this level
  uint32_t* birth;    // tracked sets born here: raster index of the pixel whose insertion promoted them
  uint8_t* kind;      // 0 min_reg, 1 tracked
};

MB2_HD uint32_t emu_find(EmuScratch& s, uint32_t e) {
  uint32_t r = e;
  while (s.uf[r] != r) r = s.uf[r];
  while (s.uf[e] != r) { uint32_t n = s.uf[e]; s.uf[e] = r; e = n; }
  return r;
}

struct EmuResult {
  uint32_t survivor;  // child rep whose region continues through this node, or NONE
  uint32_t birth;     // if survivor == NONE and the node is tracked: promotion time (raster index)
  int overflow;       // 1 if a min_reg label absorbed >= 32768 pixels: the reference's 15-bit size field wraps there
};

MB2_HD void emu_add_pixel(const Tree& t, EmuScratch& s, uint32_t root, uint32_t p, EmuResult& res) {  // InsMarkPixel :144-171
  s.uf[p] = root;
  s.size[root]++;
  if (s.kind[root] == 0) {
    if (s.size[root] >= 32768u) res.overflow = 1;
    if ((int)s.size[root] >= t.track_size) { s.kind[root] = 1; s.presize[root] = 0; s.ident[root] = NONE; s.birth[root] = p; }
  }
}

// own: the node's own pixels in raster order (the rep is one of them), as the low 32 bits of sorted (node<<32 | pixel) keys
MB2_HD EmuResult emulate_node(const Tree& t, EmuScratch& s, uint32_t v, const unsigned long long* own, int n_own) {
  EmuResult res; res.survivor = NONE; res.birth = NONE; res.overflow = 0;
  const int L = t.lev[v], W = t.W;
  for (int k = 0; k < n_own; k++) {
    const uint32_t p = (uint32_t)own[k];
    const int x = (int)(p % (uint32_t)W), y = (int)((p / (uint32_t)W) % (uint32_t)t.H);   // row inside its own image
    uint32_t lab[4]; int n = 0;
    for (int d = 0; d < 4; d++) {  // up, left, right, down
      uint32_t q;
      if (d == 0) { if (y == 0) continue; q = p - W; }
      else if (d == 1) { if (x == 0) continue; q = p - 1; }
      else if (d == 2) { if (x == W - 1) continue; q = p + 1; }
      else { if (y == t.H - 1) continue; q = p + W; }
      const int lq = t.lev[q];
      uint32_t e;
      if (lq < L) {
        const uint32_t c = child_containing(t, v, q);
        e = s.N + c;
        if (s.uf[e] == NONE) {  // first contact with this child
          s.uf[e] = e; s.size[e] = t.area[c];
          if (tracked(t, c)) { s.kind[e] = 1; s.presize[e] = t.area[c]; s.ident[e] = c; s.birth[e] = NONE; }
          else s.kind[e] = 0;
        }
      } else if (lq == L && q < p) e = q;
      else continue;
      const uint32_t r = emu_find(s, e);
      bool dup = false;
      for (int j = 0; j < n; j++) dup |= (lab[j] == r);
      if (!dup) lab[n++] = r;
    }
    if (n == 0) { s.uf[p] = p; s.size[p] = 1; s.kind[p] = 0; continue; }       // ConsRegion :174-178
    if (n == 1) { emu_add_pixel(t, s, lab[0], p, res); continue; }
    // MergeRegions :267-361
    uint32_t maxSize = 0, maxLabel = lab[0]; int num_large = 0;
    for (int i = 0; i < n; i++)
      if (s.kind[lab[i]] == 1) { num_large++; if (s.presize[lab[i]] > maxSize) { maxSize = s.presize[lab[i]]; maxLabel = lab[i]; } }
    for (int i = 0; i < n; i++) {
      const uint32_t l = lab[i];
      if (l == maxLabel) continue;
      s.size[maxLabel] += s.size[l];   // min_reg label or region totals: the pixels end up in maxLabel either way
      s.uf[l] = maxLabel;
    }
    if (s.kind[maxLabel] == 0 && s.size[maxLabel] >= 32768u) res.overflow = 1;
    (void)num_large;
    emu_add_pixel(t, s, maxLabel, p, res);
  }
  const uint32_t r = emu_find(s, v);
  if (s.kind[r] == 1) { res.survivor = s.ident[r]; res.birth = s.birth[r]; }
  return res;
}

// ---- a tracked region = chain of nodes v0 -> ... (survivor links) ------------------------------------------------
// surv[v] = child rep whose region continues through v (NONE: none).  Returns the region's maximum_int and whether
// it is the region alive at the root.
MB2_HD int region_extent(const Tree& t, const uint32_t* surv, uint32_t v0, bool* at_root, uint32_t* last) {
  uint32_t u = v0;
  for (;;) {
    if (is_root(t, u)) { *at_root = true; *last = u; return t.lev[u]; }
    const uint32_t p = t.parent[u];
    if (surv[p] != u) { *at_root = false; *last = u; return t.lev[p]; }
    u = p;
  }
}

struct Thresh { int thresh, pos, margin; };

// FastSetOptThresholds4StableRegion + SuppresOverlappingTresholds4StableRegions (optThresh.cpp:15-166).
// cA/cB: 256 cumulative areas/borders, element i at cA[i*stride].  T: room for max_t thresholds.
// Returns the number of thresholds, or -1 when more than max_t would have been needed.
MB2_HD int select_thresholds(const int* cA, const int* cB, int stride, int minimum_int, int maximum_int, double min_margin,
                             int min_size, int max_size, Thresh* T, int max_t) {
  int nt = 0; bool over = false;
  int up, localMaxMargin = -1, localMaxPos = -1;
  int i = minimum_int;
#define MB2_MSER_EMIT()                                                                             \
  do {                                                                                              \
    const int th_ = localMaxPos + localMaxMargin / 2;                                               \
    const int a_ = cA[th_ * stride];                                                                \
    if (a_ <= max_size && a_ > min_size) {                                                          \
      if (nt < max_t) { T[nt].thresh = th_; T[nt].pos = localMaxPos; T[nt].margin = localMaxMargin; nt++; } \
      else over = true;                                                                             \
    }                                                                                               \
  } while (0)
  do {
    const int area_i = cA[i * stride], radius_i = cB[i * stride];
    up = (int)(i + min_margin);
    if (up > maximum_int) break;
    while ((cA[up * stride] - area_i < radius_i) && (up < maximum_int)) up++;
    const int margin = up - i;
    const double quality = (double)margin;
    if (quality > min_margin && margin >= localMaxMargin) { localMaxMargin = margin; localMaxPos = i; }
    else {
      if (localMaxPos >= 0) { MB2_MSER_EMIT(); localMaxPos = -1; }
      localMaxMargin = margin;
    }
    i++;
  } while (up < maximum_int);
  if (localMaxPos >= 0) MB2_MSER_EMIT();
#undef MB2_MSER_EMIT
  if (over) return -1;
  // overlap suppression: keep the larger margin (ties: the earlier one)
  for (int k = 0; k < nt; k++) {
    while (k >= 0 && k + 1 < nt) {
      const Thresh a = T[k], b = T[k + 1];
      if ((a.pos + a.margin < b.thresh) && (a.thresh < b.pos)) break;
      if (b.margin <= a.margin) { for (int j = k + 1; j + 1 < nt; j++) T[j] = T[j + 1]; nt--; }
      else { for (int j = k; j + 1 < nt; j++) T[j] = T[j + 1]; nt--; k--; break; }
    }
  }
  // merge neighbours whose areas differ by at most 10 %
  for (int k = 0; k < nt; k++) {
    while (k + 1 < nt) {
      const Thresh b = T[k + 1];
      if (T[k].pos + T[k].margin < b.pos) break;
      if ((double)(cA[b.thresh * stride] - cA[T[k].thresh * stride]) <= 0.1 * (double)cA[T[k].thresh * stride]) {
        T[k].margin = b.pos - T[k].pos + b.margin; T[k].thresh = T[k].pos + T[k].margin / 2;
        for (int j = k + 1; j + 1 < nt; j++) T[j] = T[j + 1];
        nt--;
      } else break;
    }
  }
  return nt;
}

// Fills the cumulative histograms of the region born at v0 (t_region::pixels / borders after the prefix sums of
// optThresh.cpp:84-88) for levels minimum_int .. maximum_int-1; the entry at maximum_int is never read by the
// reference's selection and is set to the value below it.
MB2_HD void region_histograms(const Tree& t, const uint32_t* surv, uint32_t v0, int maximum_int, bool at_root, int* cA, int* cB, int stride) {
  uint32_t u = v0;
  int lv = t.lev[u];
  for (;;) {
    const int a = (int)t.area[u], b = border_of(t, u);
    int next_lv; uint32_t nu = NONE;
    if (is_root(t, u)) next_lv = 256;
    else { const uint32_t p = t.parent[u]; if (surv[p] == u) { nu = p; next_lv = t.lev[p]; } else next_lv = 256; }
    const int hi = (next_lv < maximum_int + 1) ? next_lv : maximum_int + 1;
    for (int i = lv; i < hi; i++) { cA[i * stride] = a; cB[i * stride] = b; }
    if (nu == NONE) break;
    u = nu; lv = next_lv;
  }
  (void)at_root;
}

// node of the region's chain that is the component at threshold `thresh`
MB2_HD uint32_t region_node_at(const Tree& t, const uint32_t* surv, uint32_t v0, int thresh) {
  uint32_t u = v0;
  for (;;) {
    if (is_root(t, u)) return u;
    const uint32_t p = t.parent[u];
    if (surv[p] != u || (int)t.lev[p] > thresh) return u;
    u = p;
  }
}

// ---- ellipse from row runs (RLE2Ellipse, libExtrema.cpp:117-159) and its square root (extrema.cpp:417-427) --------
struct Moments { double cx, cy, sxx, sxy, syy; };

#ifdef __CUDA_ARCH__
#define MB2_DMUL(a, b) __dmul_rn(a, b)
#define MB2_DADD(a, b) __dadd_rn(a, b)
#define MB2_DSUB(a, b) __dsub_rn(a, b)
#define MB2_DDIV(a, b) __ddiv_rn(a, b)
#define MB2_DSQRT(a) __dsqrt_rn(a)
#else
#define MB2_DMUL(a, b) ((a) * (b))
#define MB2_DADD(a, b) ((a) + (b))
#define MB2_DSUB(a, b) ((a) - (b))
#define MB2_DDIV(a, b) ((a) / (b))
#define MB2_DSQRT(a) sqrt(a)
#endif

// starts/ends: sorted keys (slot<<32 | line<<16 | col) of the first / last pixel of every run of one region
MB2_HD Moments moments_from_runs(const unsigned long long* starts, const unsigned long long* ends, int n) {
  double area = 0, sumX = 0, sumY = 0;
  for (int j = 0; j < n; j++) {
    const double line = (double)(int)((starts[j] >> 16) & 0xffff), m = (double)(int)(starts[j] & 0xffff);
    const double nn = (double)(1 + (int)(ends[j] & 0xffff));
    sumX = MB2_DADD(sumX, MB2_DDIV(MB2_DSUB(MB2_DMUL(nn, nn), MB2_DMUL(m, m)), 2.0));
    sumY = MB2_DADD(sumY, MB2_DDIV(MB2_DMUL(MB2_DSUB(nn, m), MB2_DADD(MB2_DMUL(2.0, line), 1.0)), 2.0));
    area = MB2_DADD(area, MB2_DSUB(nn, m));
  }
  Moments o;
  o.cx = MB2_DDIV(sumX, area); o.cy = MB2_DDIV(sumY, area);
  double sumX2 = 0, sumY2 = 0, sumXY = 0;
  for (int j = 0; j < n; j++) {
    const double line = MB2_DSUB((double)(int)((starts[j] >> 16) & 0xffff), o.cy);
    const double m = MB2_DSUB((double)(int)(starts[j] & 0xffff), o.cx);
    const double nn = MB2_DSUB((double)(1 + (int)(ends[j] & 0xffff)), o.cx);
    const double l2 = MB2_DMUL(line, line), m2 = MB2_DMUL(m, m), n2 = MB2_DMUL(nn, nn);
    sumX2 = MB2_DADD(sumX2, MB2_DDIV(MB2_DSUB(MB2_DMUL(n2, nn), MB2_DMUL(m2, m)), 3.0));
    sumY2 = MB2_DADD(sumY2, MB2_DDIV(MB2_DMUL(MB2_DSUB(nn, m), MB2_DADD(MB2_DADD(MB2_DMUL(3.0, l2), MB2_DMUL(3.0, line)), 1.0)), 3.0));
    sumXY = MB2_DADD(sumXY, MB2_DMUL(MB2_DMUL(-.25, MB2_DSUB(m2, n2)), MB2_DADD(MB2_DMUL(2.0, line), 1.0)));
  }
  o.sxx = MB2_DDIV(sumX2, area); o.syy = MB2_DDIV(sumY2, area); o.sxy = MB2_DDIV(sumXY, area);
  return o;
}

// C.schur_sym(U, T); A = U * T.sqrt() * U.transpose()   (matrix.cpp:185-217, :163-169, :65-73)
MB2_HD void ellipse_to_A(double sxx, double sxy, double syy, double* A) {
  double r, t;
  if (sxy != 0) {
    r = MB2_DDIV(MB2_DSUB(syy, sxx), MB2_DMUL(2.0, sxy));
    if (r >= 0) t = MB2_DDIV(1.0, MB2_DADD(r, MB2_DSQRT(MB2_DADD(1.0, MB2_DMUL(r, r)))));
    else t = MB2_DDIV(-1.0, MB2_DADD(-r, MB2_DSQRT(MB2_DADD(1.0, MB2_DMUL(r, r)))));
    r = MB2_DDIV(1.0, MB2_DSQRT(MB2_DADD(1.0, MB2_DMUL(t, t))));
    t = MB2_DMUL(t, r);
  } else { r = 1; t = 0; }
  const double Q[4] = {r, t, -t, r}, Qt[4] = {r, -t, t, r}, C[4] = {sxx, sxy, sxy, syy};
  double M1[4], T[4], M2[4];
#define MB2_MUL2(a, b, c)                                                   \
  do {                                                                      \
    c[0] = MB2_DADD(MB2_DMUL(a[0], b[0]), MB2_DMUL(a[1], b[2]));            \
    c[1] = MB2_DADD(MB2_DMUL(a[0], b[1]), MB2_DMUL(a[1], b[3]));            \
    c[2] = MB2_DADD(MB2_DMUL(a[2], b[0]), MB2_DMUL(a[3], b[2]));            \
    c[3] = MB2_DADD(MB2_DMUL(a[2], b[1]), MB2_DMUL(a[3], b[3]));            \
  } while (0)
  MB2_MUL2(Qt, C, M1); MB2_MUL2(M1, Q, T);
  T[1] = 0; T[2] = 0;
  const double S[4] = {MB2_DSQRT(T[0]), MB2_DSQRT(T[1]), MB2_DSQRT(T[2]), MB2_DSQRT(T[3])};
  MB2_MUL2(Q, S, M2); MB2_MUL2(M2, Qt, A);
#undef MB2_MUL2
}

}  // namespace mser_logic
