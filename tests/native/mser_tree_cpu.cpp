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

#include "../../mods_b200/csrc/mser_logic.cuh"

using namespace mser_logic;

namespace {
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
  void run() {
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

extern "C" void mser_tree_ellipse_to_A(double sxx, double sxy, double syy, double* A) { ellipse_to_A(sxx, sxy, syy, A); }
