// TEST INFRASTRUCTURE -- host harness for mods_b200/csrc/mser_logic.cuh.
// Builds the canonical component tree with the same level-by-level union-find the GPU uses (sequentially), then runs the
// shared decision logic (survivors, replay of births / ties, stability thresholds, run moments) in plain loops, so that
// the logic can be checked against the oracle / reference on the CPU.  Not part of the product, not a fallback: the
// library never calls this.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include <atomic>
#include <random>
#include <thread>

#include "../../mods_b200/csrc/mser_logic.cuh"
#include "../../mods_b200/csrc/mser_tree_build.cuh"

using namespace mser_logic;

// how the tree is built: 0 = level-by-level union-find (the original sequential builder), 1 = the lock-free merge of
// mser_tree_build.cuh run sequentially (tiles first, then tile borders, edges in a seeded random order), 2 = the same on several host
// threads at once (real compare-and-swap races), every edge in random order
static int g_mode = 0, g_tile = 64, g_threads = 4; static unsigned g_seed = 1;
extern "C" void mser_tree_set_mode(int mode, int tile, int threads, unsigned seed) { g_mode = mode; g_tile = tile; g_threads = threads; g_seed = seed; }

namespace {
struct HostMem {   // plain words, or real atomics when several threads run
  uint32_t* w;
  uint32_t load(uint32_t i) { return __atomic_load_n(w + i, __ATOMIC_RELAXED); }
  void store(uint32_t i, uint32_t v) { __atomic_store_n(w + i, v, __ATOMIC_RELAXED); }
  bool cas(uint32_t i, uint32_t expect, uint32_t desired) { return __atomic_compare_exchange_n(w + i, &expect, desired, false, __ATOMIC_ACQ_REL, __ATOMIC_RELAXED); }
};
typedef mser_tree::Key<24> K24;

struct Build {
  int W, H, N;
  std::vector<uint8_t> lev;
  std::vector<uint32_t> zpar, parent, area, nedge;
  uint32_t root;
  uint32_t find(uint32_t x) {
    while (zpar[x] != x) { zpar[x] = zpar[zpar[x]]; x = zpar[x]; }
    return x;
  }
  bool less(uint32_t a, uint32_t b) const { return lev[a] < lev[b] || (lev[a] == lev[b] && a > b); }  // same hooking order as mser.cu
  // the lock-free merge: par words, canonical parents, then own counts summed up the tree level by level (as mser.cu does)
  void run_merge() {
    N = W * H;
    std::vector<uint32_t> par(N);
    for (int i = 0; i < N; i++) par[i] = K24::make(lev[i], (uint32_t)i);
    HostMem m{par.data()};
    struct E { uint32_t a, b; };
    std::vector<E> inner, border;
    for (int y = 0; y < H; y++)
      for (int x = 0; x < W; x++) {
        const uint32_t p = (uint32_t)(y * W + x);
        if (x + 1 < W) ((x + 1) % g_tile == 0 ? border : inner).push_back(E{p, p + 1});
        if (y + 1 < H) ((y + 1) % g_tile == 0 ? border : inner).push_back(E{p, p + (uint32_t)W});
      }
    std::mt19937 rng(g_seed);
    std::shuffle(inner.begin(), inner.end(), rng); std::shuffle(border.begin(), border.end(), rng);
    auto key = [&](uint32_t p) { return K24::make(lev[p], p); };
    auto work = [&](const std::vector<E>& e, int t, int nt) { HostMem mm{par.data()}; for (size_t i = t; i < e.size(); i += nt) mser_tree::connect<K24>(mm, key(e[i].a), key(e[i].b)); };
    if (g_mode == 1) { work(inner, 0, 1); work(border, 0, 1); }
    else {
      for (const std::vector<E>* e : {&inner, &border}) {
        std::vector<std::thread> th;
        for (int t = 0; t < g_threads; t++) th.emplace_back(work, std::cref(*e), t, g_threads);
        for (auto& t : th) t.join();
      }
    }
    parent.resize(N); area.assign(N, 0); nedge.assign(N, 0);
    for (int i = 0; i < N; i++) parent[i] = K24::idx(mser_tree::canonical_parent<K24>(m, key(i)));
    std::vector<std::vector<uint32_t>> by_level(256);
    for (int p = 0; p < N; p++) {
      const bool rep = parent[p] == (uint32_t)p || lev[parent[p]] > lev[p];
      const uint32_t r = rep ? (uint32_t)p : parent[p];
      if (rep) by_level[lev[p]].push_back(p);
      const int x = p % W, y = p / W;
      uint32_t e = 0;
      const int dx[4] = {0, -1, 1, 0}, dy[4] = {-1, 0, 0, 1};
      for (int d = 0; d < 4; d++) {
        const int qx = x + dx[d], qy = y + dy[d];
        if (qx < 0 || qy < 0 || qx >= W || qy >= H) continue;
        const uint32_t q = qy * W + qx;
        if (lev[q] < lev[p] || (lev[q] == lev[p] && q < (uint32_t)p)) e++;
      }
      area[r] += 1; nedge[r] += e;
    }
    for (int L = 0; L < 256; L++)
      for (uint32_t v : by_level[L]) if (parent[v] != v) { area[parent[v]] += area[v]; nedge[parent[v]] += nedge[v]; }
    root = 0; while (parent[root] != root) root = parent[root];
  }
  void run() {
    if (g_mode != 0) { run_merge(); return; }
    N = W * H;
    zpar.resize(N); parent.resize(N); area.assign(N, 1); nedge.assign(N, 0);
    for (int i = 0; i < N; i++) zpar[i] = parent[i] = i;
    std::vector<uint32_t> start(257, 0), order(N);
    for (int i = 0; i < N; i++) start[lev[i] + 1]++;
    for (int i = 0; i < 256; i++) start[i + 1] += start[i];
    { std::vector<uint32_t> cur(start.begin(), start.end() - 1); for (int i = 0; i < N; i++) order[cur[lev[i]]++] = i; }
    std::vector<uint32_t> hooked;
    for (int L = 0; L < 256; L++) {
      hooked.clear();
      for (uint32_t k = start[L]; k < start[L + 1]; k++) {
        const uint32_t p = order[k];
        const int x = p % W, y = p / W;
        uint32_t e = 0;
        const int dx[4] = {0, -1, 1, 0}, dy[4] = {-1, 0, 0, 1};
        for (int d = 0; d < 4; d++) {
          const int qx = x + dx[d], qy = y + dy[d];
          if (qx < 0 || qy < 0 || qx >= W || qy >= H) continue;
          const uint32_t q = qy * W + qx;
          if (!(lev[q] < L || (lev[q] == L && q < p))) continue;
          e++;
          uint32_t ra = find(p), rb = find(q);
          if (ra == rb) continue;
          if (!less(ra, rb)) std::swap(ra, rb);
          zpar[ra] = rb; hooked.push_back(ra);
        }
        nedge[p] = e;
      }
      for (uint32_t x : hooked) {
        const uint32_t r = find(x);
        parent[x] = r; area[r] += area[x]; nedge[r] += nedge[x];
      }
    }
    root = find(0);
  }
};
}  // namespace

extern "C" int mser_tree_regions(const float* img, int w, int h, double max_area, int min_size, double min_margin, double* out, int max_out,
                                 int* stats /* [0] births [1] tie nodes [2] overflow flags [3] threshold overflow */) {
  int n_out = 0;
  stats[0] = stats[1] = stats[2] = stats[3] = 0;
  const int N = w * h;
  for (int pol = 0; pol < 2; pol++) {
    Build b; b.W = w; b.H = h; b.lev.resize(N);
    for (int i = 0; i < N; i++) { uint8_t v = (unsigned char)img[i]; b.lev[i] = pol ? (uint8_t)(255 - v) : v; }
    b.run();
    Tree t; t.W = w; t.H = h; t.lev = b.lev.data(); t.parent = b.parent.data(); t.area = b.area.data(); t.nedge = b.nedge.data();
    t.track_size = std::min(10000, min_size);
    const int max_size = (int)((double)w * (double)h * max_area);
    // survivors: largest tracked child; ties and births are replayed
    std::vector<uint32_t> surv(N, NONE), bestArea(N, 0), nbest(N, 0);
    for (int x = 0; x < N; x++) {
      if (is_root(t, x) || !is_rep(t, x) || !tracked(t, x)) continue;
      const uint32_t v = t.parent[x];
      if (t.area[x] > bestArea[v]) { bestArea[v] = t.area[x]; surv[v] = x; nbest[v] = 1; }
      else if (t.area[x] == bestArea[v]) nbest[v]++;
    }
    std::vector<uint32_t> emu_nodes;
    for (int v = 0; v < N; v++) {
      if (!is_rep(t, v) || !tracked(t, v)) continue;
      if (surv[v] == NONE) { emu_nodes.push_back(v); stats[0]++; }
      else if (nbest[v] > 1) { emu_nodes.push_back(v); stats[1]++; }
    }
    // own pixels of the replayed nodes, raster order
    std::vector<std::vector<unsigned long long>> own(emu_nodes.size());
    {
      std::vector<int> slot(N, -1);
      for (size_t i = 0; i < emu_nodes.size(); i++) slot[emu_nodes[i]] = (int)i;
      for (int x = 0; x < N; x++) { const int s = slot[node_of(t, x)]; if (s >= 0) own[s].push_back(((unsigned long long)emu_nodes[s] << 32) | (unsigned)x); }
    }
    std::vector<uint32_t> uf(2 * N, NONE), esz(2 * N), epre(2 * N), eid(2 * N), ebirth(2 * N), birth(N, NONE);
    std::vector<uint8_t> ekind(2 * N);
    EmuScratch s{(uint32_t)N, uf.data(), esz.data(), epre.data(), eid.data(), ebirth.data(), ekind.data()};
    for (size_t i = 0; i < emu_nodes.size(); i++) {
      const uint32_t v = emu_nodes[i];
      EmuResult r = emulate_node(t, s, v, own[i].data(), (int)own[i].size());
      stats[2] += r.overflow;
      surv[v] = r.survivor; birth[v] = r.birth;
    }
    // regions
    struct Sel { unsigned long long key; uint32_t node; int minI, maxI, thresh, margin, area, border; };
    std::vector<Sel> sel;
    for (int v0 = 0; v0 < N; v0++) {
      if (!is_rep(t, v0) || !tracked(t, v0) || surv[v0] != NONE) continue;
      bool at_root; uint32_t last;
      const int maxI = region_extent(t, surv.data(), v0, &at_root, &last), minI = t.lev[v0];
      if (!at_root && (maxI - minI + 1) <= min_margin) continue;
      int cA[256], cB[256];
      region_histograms(t, surv.data(), v0, maxI, at_root, cA, cB, 1);
      Thresh T[128];
      int nt = select_thresholds(cA, cB, 1, minI, maxI, min_margin, min_size, max_size, T, 128);
      if (nt < 0) { stats[3]++; continue; }
      for (int k = 0; k < nt; k++) {
        Sel e; e.key = ((unsigned long long)minI << 40) | ((unsigned long long)birth[v0] << 8) | (unsigned)k;
        e.node = region_node_at(t, surv.data(), v0, T[k].thresh);
        e.minI = minI; e.maxI = maxI; e.thresh = T[k].thresh; e.margin = T[k].margin; e.area = cA[T[k].thresh]; e.border = cB[T[k].thresh];
        sel.push_back(e);
      }
    }
    std::sort(sel.begin(), sel.end(), [](const Sel& a, const Sel& b) { return a.key < b.key; });
    // runs + moments (plain: mark the subtree of each selected node)
    for (const Sel& e : sel) {
      std::vector<unsigned long long> starts, ends;
      std::vector<uint8_t> in(N, 0);
      for (int x = 0; x < N; x++) {
        uint32_t u = node_of(t, x);
        while (u != e.node && !is_root(t, u) && t.lev[u] <= t.lev[e.node]) u = t.parent[u];
        in[x] = (u == e.node);
      }
      for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
          if (!in[y * w + x]) continue;
          if (x == 0 || !in[y * w + x - 1]) starts.push_back(((unsigned long long)y << 16) | (unsigned)x);
          if (x == w - 1 || !in[y * w + x + 1]) ends.push_back(((unsigned long long)y << 16) | (unsigned)x);
        }
      Moments m = moments_from_runs(starts.data(), ends.data(), (int)starts.size());
      if (n_out < max_out) {
        double* o = out + (size_t)n_out * 13;
        o[0] = pol; o[1] = e.minI; o[2] = e.maxI; o[3] = e.thresh; o[4] = e.margin; o[5] = e.area; o[6] = e.border; o[7] = (double)starts.size();
        o[8] = m.cx; o[9] = m.cy; o[10] = m.sxx; o[11] = m.sxy; o[12] = m.syy;
      }
      n_out++;
    }
  }
  return n_out;
}

// node-level comparison of two builders on one image: for every pixel the (level, area, edge count) of its node and of its parent node
extern "C" int mser_tree_compare(const float* img, int w, int h, int pol) {
  const int N = w * h;
  auto build = [&](int mode) {
    const int keep = g_mode; g_mode = mode;
    Build b; b.W = w; b.H = h; b.lev.resize(N);
    for (int i = 0; i < N; i++) { uint8_t v = (unsigned char)img[i]; b.lev[i] = pol ? (uint8_t)(255 - v) : v; }
    b.run(); g_mode = keep;
    return b;
  };
  Build a = build(0), b = build(g_mode);
  int bad = 0;
  auto sig = [&](Build& t, int p, uint32_t* o) {
    const bool rep = t.parent[p] == (uint32_t)p || t.lev[t.parent[p]] > t.lev[p];
    const uint32_t r = rep ? (uint32_t)p : t.parent[p];
    if (t.lev[r] != t.lev[p]) { o[0] = 0xdeadbeef; return; }
    const uint32_t pr = t.parent[r];
    o[0] = t.area[r]; o[1] = t.nedge[r]; o[2] = pr == r ? 256u : t.lev[pr]; o[3] = t.area[pr]; o[4] = t.nedge[pr];
    o[5] = (pr == r || t.parent[pr] == pr || t.lev[t.parent[pr]] > t.lev[pr]) ? 1u : 0u;   // the parent pointer of a representative names a representative
  };
  for (int p = 0; p < N; p++) {
    uint32_t sa[6] = {0}, sb[6] = {0};
    sig(a, p, sa); sig(b, p, sb);
    if (std::memcmp(sa, sb, sizeof sa) != 0 || sb[5] != 1u) bad++;
  }
  return bad;
}

extern "C" void mser_tree_ellipse_to_A(double sxx, double sxy, double syy, double* A) { ellipse_to_A(sxx, sxy, syy, A); }
