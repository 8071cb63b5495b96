// MSER detector on the GPU: replaces DetectMSERs (detectors/mser/extrema/extrema.cpp:284-473)
// and everything below it (libExtrema.cpp, sortPixels.cpp, getExtrema.cpp, optThresh.cpp, boundary.cpp).
//
// The reference is sequential in (intensity, raster) order.  Here:
//   1. k_mser_prep / radix sort      u8 image of the polarity, pixels ordered by (level, raster)      [HBM streaming]
//   2. k_mser_union / k_mser_final   component tree of the level sets, level by level: lock-free union-find with
//                                    "larger (level, index) wins" hooking, so the representative of a component is
//                                    its canonical pixel; every element that stops being a root is recorded once
//                                    (`hooked`, grouped by level), gets its canonical parent and adds its area / inner
//                                    edge count to the new representative                              [L2 latency]
//   3. k_mser_best .. k_mser_emulate survivor of every merge = largest tracked child; nodes where the reference's
//                                    answer depends on its pixel order (births, equal sizes) are replayed by one
//                                    thread each (mser_logic.cuh: emulate_node)
//   4. k_mser_regions_*              one thread per tracked region: lifetime, cumulative area/border histograms,
//                                    stability thresholds (FastSetOptThresholds4StableRegion)
//   5. k_mser_down / k_mser_runs     selected components -> row runs (start / end events), sorted per region
//   6. k_mser_moments / k_mser_keys  RLE2Ellipse in the reference's summation order, sqrtm, AffineKeypoint record
// Integer / index work is bit-exact by construction; the f64 moment sums follow the reference's order run by run.
#include "common.cuh"
#undef MB2_NS
#define MB2_NS mb2_mser_detail
#include "pyramid.cuh"
#include "mser_logic.cuh"

#include <cub/cub.cuh>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace MB2_NS {
using namespace mser_logic;
typedef unsigned long long u64;

struct MserCounters {
  uint32_t hook_cnt, emu_nodes, own_keys, long_regions, n_sel, n_slots, n_starts, n_ends;
  uint32_t overflow, thr_overflow, root, cap_overflow;
  uint32_t hist[256];
  uint32_t lvl_off[257];
  uint32_t hook_off[258];
};

struct SelRec {  // one (region, threshold): a row of getRLEExtrema's output
  u64 key;       // birth level << 40 | promotion time << 8 | threshold rank: the reference's list order
  uint32_t node, slot;
  int minI, maxI, thresh, margin, area, border;
};

struct MserBufs {
  DevBuf lev, zpar, parent, area, nedge, order, order_in, keys8, hooked, best, surv, birth, flag, emu_nodes, own_a, own_b;
  DevBuf uf, esz, epre, eid, ebirth, ekind, longr, sel, sel_sorted, selkey_a, selkey_b, selidx_a, selidx_b, slot_of_node, node_of_slot;
  DevBuf sa, up_sel, ev_a, ev_b, ev_c, ev_d, mom, cub_tmp, counters, table;
  HostBuf h_counters;
  void release() {
    DevBuf* all[] = {&lev, &zpar, &parent, &area, &nedge, &order, &order_in, &keys8, &hooked, &best, &surv, &birth, &flag, &emu_nodes, &own_a,
                     &own_b, &uf, &esz, &epre, &eid, &ebirth, &ekind, &longr, &sel, &sel_sorted, &selkey_a, &selkey_b, &selidx_a, &selidx_b,
                     &slot_of_node, &node_of_slot, &sa, &up_sel, &ev_a, &ev_b, &ev_c, &ev_d, &mom, &cub_tmp, &counters, &table};
    for (DevBuf* b : all) b->release();
    h_counters.release();
  }
};

// ---- union-find on zpar (L2-coherent accesses: other SMs hook concurrently) ---------------------------------------
__device__ __forceinline__ uint32_t uf_find(uint32_t* zpar, uint32_t x) {
  for (;;) {
    const uint32_t p = __ldcg(zpar + x);
    if (p == x) return x;
    const uint32_t g = __ldcg(zpar + p);
    if (g != p) __stcg(zpar + x, g);  // path halving; values only ever move towards the root
    x = g;
  }
}
__device__ __forceinline__ bool key_less(const uint8_t* lev, uint32_t a, uint32_t b) {
  const int la = lev[a], lb = lev[b];
  return la < lb || (la == lb && a < b);
}
// joins the sets of a and b; returns the element that stopped being a root (NONE if already joined)
__device__ __forceinline__ uint32_t uf_unite(uint32_t* zpar, const uint8_t* lev, uint32_t a, uint32_t b) {
  for (;;) {
    uint32_t ra = uf_find(zpar, a), rb = uf_find(zpar, b);
    if (ra == rb) return NONE;
    if (key_less(lev, rb, ra)) { const uint32_t t = ra; ra = rb; rb = t; }
    if (atomicCAS(zpar + ra, ra, rb) == ra) return ra;
  }
}

// ---- 1. preparation ------------------------------------------------------------------------------------------------
// float -> u8 as extrema.cpp:401-403 does ((unsigned char) of the float: truncation), inverted for MSER-
__global__ void k_mser_prep(const float* __restrict__ img, int pitch, int W, int H, int pol, uint8_t* __restrict__ lev,
                            uint32_t* __restrict__ zpar, uint32_t* __restrict__ parent, uint32_t* __restrict__ area,
                            uint32_t* __restrict__ order_in, MserCounters* __restrict__ C) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t N = (uint32_t)W * H;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    const int y = i / W, x = i - y * W;
    const int v = ((int)img[(size_t)y * pitch + x]) & 0xff;
    const uint8_t l = (uint8_t)(pol ? 255 - v : v);
    lev[i] = l; zpar[i] = i; parent[i] = i; area[i] = 1; order_in[i] = i;
    atomicAdd(&h[l], 1u);
  }
  __syncthreads();
  if (h[threadIdx.x]) atomicAdd(&C->hist[threadIdx.x], h[threadIdx.x]);
}
__global__ void k_mser_offsets(MserCounters* C) {
  if (threadIdx.x || blockIdx.x) return;
  uint32_t s = 0;
  for (int i = 0; i < 256; i++) { C->lvl_off[i] = s; s += C->hist[i]; }
  C->lvl_off[256] = s;
  C->hook_off[0] = 0;
}

// ---- 2. component tree, one level per launch pair ---------------------------------------------------------------------
// A pixel joins every 4-neighbour that the reference has already labelled when it reaches the pixel: lower level, or the
// same level and earlier in raster order (getExtrema.cpp:216-263).  nedge[p] counts them (border_num / 2).
__global__ void k_mser_union(int L, int W, int H, const uint8_t* __restrict__ lev, const uint32_t* __restrict__ order, uint32_t* zpar,
                             uint32_t* __restrict__ nedge, uint32_t* __restrict__ hooked, MserCounters* C) {
  const uint32_t beg = C->lvl_off[L], end = C->lvl_off[L + 1];
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t base = beg + warp * 32; base < end; base += nwarps * 32) {
    const uint32_t k = base + lane;
    uint32_t hk[4]; int cnt = 0;
    if (k < end) {
      const uint32_t p = order[k];
      const int y = p / W, x = p - y * W;
      uint32_t e = 0;
#pragma unroll
      for (int d = 0; d < 4; d++) {
        uint32_t q;
        if (d == 0) { if (y == 0) continue; q = p - W; }
        else if (d == 1) { if (x == 0) continue; q = p - 1; }
        else if (d == 2) { if (x == W - 1) continue; q = p + 1; }
        else { if (y == H - 1) continue; q = p + W; }
        const int lq = lev[q];
        if (lq < L || (lq == L && q < p)) {
          e++;
          const uint32_t h = uf_unite(zpar, lev, p, q);
          if (h != NONE) hk[cnt++] = h;
        }
      }
      nedge[p] = e;
    }
    __syncwarp();
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
    const int tot = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t b0 = 0;
    if (lane == 31 && tot) b0 = atomicAdd(&C->hook_cnt, (uint32_t)tot);
    b0 = __shfl_sync(0xffffffffu, b0, 31);
    for (int j = 0; j < cnt; j++) hooked[b0 + incl - cnt + j] = hk[j];
  }
}
// Every element hooked during level L now learns its canonical parent (the representative of the level-L node) and hands
// its totals over; lanes that share a representative combine first.
__global__ void k_mser_final(int L, uint32_t* zpar, uint32_t* __restrict__ parent, uint32_t* area, uint32_t* nedge,
                             const uint32_t* __restrict__ hooked, MserCounters* C) {
  const uint32_t beg = C->hook_off[L], end = C->hook_cnt;
  if (blockIdx.x == 0 && threadIdx.x == 0) C->hook_off[L + 1] = end;
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t base = beg + warp * 32; base < end; base += nwarps * 32) {
    const uint32_t i = base + lane;
    const bool act = i < end;
    const unsigned mask = __ballot_sync(0xffffffffu, act);
    if (act) {
      const uint32_t x = hooked[i];
      const uint32_t r = uf_find(zpar, x);
      parent[x] = r;
      const uint32_t a = area[x], e = nedge[x];
      const unsigned grp = __match_any_sync(mask, r);
      const uint32_t sa = __reduce_add_sync(grp, a), se = __reduce_add_sync(grp, e);
      if (lane == __ffs(grp) - 1) { atomicAdd(area + r, sa); atomicAdd(nedge + r, se); }
    }
  }
}
__global__ void k_mser_root(uint32_t* zpar, MserCounters* C) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { C->root = uf_find(zpar, 0); C->hook_off[257] = C->hook_cnt; }
}

// ---- 3. survivors -------------------------------------------------------------------------------------------------------
struct TreeDev {  // Tree with the root read from the counters block
  int W, H; const uint8_t* lev; const uint32_t* parent; const uint32_t* area; const uint32_t* nedge; const MserCounters* C; int track_size;
  __device__ __forceinline__ Tree get() const {
    Tree t; t.W = W; t.H = H; t.lev = lev; t.parent = parent; t.area = area; t.nedge = nedge; t.root = C->root; t.track_size = track_size;
    return t;
  }
};
// best[v] = (area << 32 | ~rep) of the largest tracked child of v
__global__ void k_mser_best(TreeDev td, uint32_t N, u64* __restrict__ best) {
  const Tree t = td.get();
  for (uint32_t x = blockIdx.x * blockDim.x + threadIdx.x; x < N; x += gridDim.x * blockDim.x) {
    if (x == t.root || !is_rep(t, x) || !tracked(t, x)) continue;
    atomicMax(best + t.parent[x], ((u64)t.area[x] << 32) | (u64)(~x));
  }
}
// flag: 1 = region born in this node, 2 = equally large tracked children (survivor decided by replay), 0 otherwise
__global__ void k_mser_ties(TreeDev td, uint32_t N, const u64* __restrict__ best, uint8_t* __restrict__ flag) {
  const Tree t = td.get();
  for (uint32_t x = blockIdx.x * blockDim.x + threadIdx.x; x < N; x += gridDim.x * blockDim.x) {
    if (x == t.root || !is_rep(t, x) || !tracked(t, x)) continue;
    const u64 b = best[t.parent[x]];
    if ((uint32_t)(b >> 32) == t.area[x] && (uint32_t)(~b) != x) flag[t.parent[x]] = 2;
  }
}
__global__ void k_mser_classify(TreeDev td, uint32_t N, const u64* __restrict__ best, uint8_t* __restrict__ flag, uint32_t* __restrict__ surv,
                                uint32_t* __restrict__ emu_nodes, uint32_t cap, MserCounters* C) {
  const Tree t = td.get();
  for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < N; v += gridDim.x * blockDim.x) {
    if (!is_rep(t, v) || !tracked(t, v)) continue;
    const u64 b = best[v];
    if (b == 0) flag[v] = 1; else surv[v] = (uint32_t)(~b);
    if (flag[v]) { const uint32_t i = atomicAdd(&C->emu_nodes, 1u); if (i < cap) emu_nodes[i] = v; else C->cap_overflow = 1; }
  }
}
// own pixels of the nodes to replay, as keys node << 32 | pixel (sorted afterwards: raster order inside a node)
__global__ void k_mser_ownkeys(TreeDev td, uint32_t N, const uint8_t* __restrict__ flag, u64* __restrict__ keys, uint32_t cap, MserCounters* C) {
  const Tree t = td.get();
  const int lane = threadIdx.x & 31;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < N; base += stride) {
    const uint32_t x = base + lane;
    uint32_t v = 0; bool want = false;
    if (x < N) { v = node_of(t, x); want = flag[v] != 0; }
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (!m) continue;
    uint32_t b0 = 0;
    if (lane == 0) b0 = atomicAdd(&C->own_keys, (uint32_t)__popc(m));
    b0 = __shfl_sync(0xffffffffu, b0, 0);
    if (want) {
      const uint32_t i = b0 + __popc(m & ((1u << lane) - 1));
      if (i < cap) keys[i] = ((u64)v << 32) | x; else C->cap_overflow = 1;
    }
  }
}
__device__ __forceinline__ uint32_t lower_bound_u64(const u64* a, uint32_t n, u64 key) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
  return lo;
}
__global__ void k_mser_emulate(TreeDev td, uint32_t N, const uint32_t* __restrict__ emu_nodes, uint32_t n_nodes, const u64* __restrict__ keys,
                               uint32_t n_keys, EmuScratch s, uint32_t* __restrict__ surv, uint32_t* __restrict__ birth, MserCounters* C) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  const Tree t = td.get();
  const uint32_t v = emu_nodes[i];
  const uint32_t lo = lower_bound_u64(keys, n_keys, (u64)v << 32), hi = lower_bound_u64(keys, n_keys, ((u64)v + 1) << 32);
  const EmuResult r = emulate_node(t, s, v, keys + lo, (int)(hi - lo));
  if (r.overflow) atomicAdd(&C->overflow, 1u);
  surv[v] = r.survivor;
  birth[v] = r.birth;
}

// ---- 4. regions --------------------------------------------------------------------------------------------------------
struct LongRegion { uint32_t v0; int maxI; int at_root; };
__global__ void k_mser_regions_a(TreeDev td, const uint32_t* __restrict__ emu_nodes, uint32_t n_nodes, const uint32_t* __restrict__ surv,
                                 double min_margin, LongRegion* __restrict__ out, uint32_t cap, MserCounters* C) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  const Tree t = td.get();
  const uint32_t v0 = emu_nodes[i];
  if (surv[v0] != NONE) return;  // a tie node: the region passing through it was born elsewhere
  bool at_root; uint32_t last;
  const int maxI = region_extent(t, surv, v0, &at_root, &last), minI = t.lev[v0];
  if (!at_root && (maxI - minI + 1) <= min_margin) return;  // getExtrema.cpp:337-338
  const uint32_t k = atomicAdd(&C->long_regions, 1u);
  if (k < cap) { out[k].v0 = v0; out[k].maxI = maxI; out[k].at_root = at_root ? 1 : 0; } else C->cap_overflow = 1;
}
#define MSER_RB_THREADS 32
#define MSER_MAX_T 128
__global__ void __launch_bounds__(MSER_RB_THREADS) k_mser_regions_b(TreeDev td, const LongRegion* __restrict__ regs, uint32_t n, const uint32_t* __restrict__ surv,
                                 const uint32_t* __restrict__ birth, double min_margin, int min_size, int max_size, uint32_t* slot_of_node,
                                 uint32_t* __restrict__ node_of_slot, SelRec* __restrict__ sel, uint32_t cap, MserCounters* C) {
  extern __shared__ int sm_hist[];  // cA[256][32], cB[256][32]
  const uint32_t i = blockIdx.x * MSER_RB_THREADS + threadIdx.x;
  if (i >= n) return;
  const Tree t = td.get();
  int* cA = sm_hist + threadIdx.x;
  int* cB = sm_hist + 256 * MSER_RB_THREADS + threadIdx.x;
  const LongRegion r = regs[i];
  const int minI = t.lev[r.v0];
  region_histograms(t, surv, r.v0, r.maxI, r.at_root != 0, cA, cB, MSER_RB_THREADS);
  if ((int)t.area[r.v0] < min_size) return;  // optThresh.cpp:71 (only when min_size > 10000)
  Thresh T[MSER_MAX_T];
  const int nt = select_thresholds(cA, cB, MSER_RB_THREADS, minI, r.maxI, min_margin, min_size, max_size, T, MSER_MAX_T);
  if (nt < 0) { atomicAdd(&C->thr_overflow, 1u); return; }
  for (int k = 0; k < nt; k++) {
    const uint32_t node = region_node_at(t, surv, r.v0, T[k].thresh);
    const uint32_t o = atomicAdd(&C->n_sel, 1u);
    if (o >= cap) { C->cap_overflow = 1; continue; }
    // one run list per distinct component: the smallest record index of a node names its slot
    node_of_slot[o] = node;
    atomicMin(slot_of_node + node, o);
    const uint32_t slot = o;
    SelRec e;
    e.key = ((u64)minI << 40) | ((u64)birth[r.v0] << 8) | (u64)k;
    e.node = node; e.slot = slot; e.minI = minI; e.maxI = r.maxI; e.thresh = T[k].thresh; e.margin = T[k].margin;
    e.area = cA[T[k].thresh * MSER_RB_THREADS]; e.border = cB[T[k].thresh * MSER_RB_THREADS];
    sel[o] = e;
  }
}
__global__ void k_mser_selkeys(const SelRec* __restrict__ sel, uint32_t n, u64* __restrict__ keys, uint32_t* __restrict__ idx) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { keys[i] = sel[i].key; idx[i] = i; }
}

// ---- 5. selected components -> row runs ------------------------------------------------------------------------------------
// sa[x] = slot of the nearest selected node on the path from x to the root (x included); levels from the top down.
__global__ void k_mser_down_init(uint32_t* __restrict__ sa, const uint32_t* __restrict__ slot_of_node, const MserCounters* C) {
  if (threadIdx.x == 0 && blockIdx.x == 0) sa[C->root] = slot_of_node[C->root];
}
__global__ void k_mser_down(int L, const uint32_t* __restrict__ parent, const uint32_t* __restrict__ hooked, const uint32_t* __restrict__ slot_of_node,
                            uint32_t* __restrict__ sa, const MserCounters* C) {
  const uint32_t beg = C->hook_off[L], end = C->hook_off[L + 1];
  for (uint32_t i = beg + blockIdx.x * blockDim.x + threadIdx.x; i < end; i += gridDim.x * blockDim.x) {
    const uint32_t x = hooked[i];
    const uint32_t s = slot_of_node[x];
    sa[x] = (s != NONE) ? s : sa[parent[x]];
  }
}
__global__ void k_mser_upsel(const uint32_t* __restrict__ parent, const uint32_t* __restrict__ node_of_slot, uint32_t n_slots,
                             const uint32_t* __restrict__ sa, uint32_t* __restrict__ up_sel, const MserCounters* C) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  const uint32_t node = node_of_slot[s];
  up_sel[s] = (node == C->root) ? NONE : sa[parent[node]];
}
__device__ __forceinline__ bool in_slot(const uint32_t* sa, const uint32_t* up_sel, uint32_t q, uint32_t s) {
  uint32_t t = sa[q];
  while (t != NONE && t != s) t = up_sel[t];
  return t == s;
}
__device__ __forceinline__ void agg_append(u64* list, uint32_t* counter, uint32_t cap, u64 key, uint32_t* overflow) {
  const unsigned m = __activemask();
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  uint32_t b0 = 0;
  if (lane == leader) b0 = atomicAdd(counter, (uint32_t)__popc(m));
  b0 = __shfl_sync(m, b0, leader);
  const uint32_t i = b0 + __popc(m & ((1u << lane) - 1));
  if (i < cap) list[i] = key; else *overflow = 1;
}
// first / last pixel of every row run of every selected component: key = slot << 32 | line << 16 | column
__global__ void k_mser_runs(int W, int H, const uint32_t* __restrict__ sa, const uint32_t* __restrict__ up_sel, u64* __restrict__ starts,
                            u64* __restrict__ ends, uint32_t cap, MserCounters* C) {
  const uint32_t N = (uint32_t)W * H;
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < N; p += gridDim.x * blockDim.x) {
    uint32_t s = sa[p];
    if (s == NONE) continue;
    const int y = p / W, x = p - y * W;
    while (s != NONE) {
      const bool l = x > 0 && in_slot(sa, up_sel, p - 1, s);
      const bool r = x < W - 1 && in_slot(sa, up_sel, p + 1, s);
      const u64 key = ((u64)s << 32) | ((u64)y << 16) | (u64)x;
      if (!l) agg_append(starts, &C->n_starts, cap, key, &C->cap_overflow);
      if (!r) agg_append(ends, &C->n_ends, cap, key, &C->cap_overflow);
      s = up_sel[s];
    }
  }
}

// ---- 6. moments, keys ---------------------------------------------------------------------------------------------------------
struct SlotMoments { Moments m; int nruns; int pad; };
__global__ void k_mser_moments(const u64* __restrict__ starts, const u64* __restrict__ ends, uint32_t n_runs, uint32_t n_slots,
                               SlotMoments* __restrict__ out) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  const uint32_t lo = lower_bound_u64(starts, n_runs, (u64)s << 32), hi = lower_bound_u64(starts, n_runs, ((u64)s + 1) << 32);
  SlotMoments o;
  o.m = moments_from_runs(starts + lo, ends + lo, (int)(hi - lo));
  o.nruns = (int)(hi - lo); o.pad = 0;
  out[s] = o;
}
// AffineKeypoint as DetectMSERs fills it (extrema.cpp:409-433) [+ DetectAffineRegions' post-step, synth-detection.hpp:110-124]
__global__ void k_mser_keys(const SelRec* __restrict__ sel, const uint32_t* __restrict__ order, uint32_t n, const SlotMoments* __restrict__ mom,
                            const uint32_t* __restrict__ slot_of_node, int pol, int as_regions, KeyOut* __restrict__ out, double* __restrict__ table) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const SelRec e = sel[order[i]];
  const SlotMoments sm = mom[slot_of_node[e.node]];
  double A[4];
  ellipse_to_A(sm.m.sxx, sm.m.sxy, sm.m.syy, A);
  double a11 = A[0], a12 = A[1], a21 = A[2], a22 = A[3], s = 1.0;
  if (as_regions) {
    s = s * sqrt(fabs(a11 * a22 - a12 * a21));
    const double a = a11, b = a12, c = a21, d = a22;
    const double det = sqrt(fabs(a * d - b * c));
    const double b2a2 = sqrt(b * b + a * a);
    a11 = b2a2 / det; a12 = 0; a21 = (d * b + c * a) / (b2a2 * det); a22 = det / b2a2;
  }
  KeyOut o;
  o.v[0] = sm.m.cx; o.v[1] = sm.m.cy; o.v[2] = a11; o.v[3] = a12; o.v[4] = a21; o.v[5] = a22; o.v[6] = s;
  o.v[7] = (double)e.margin; o.v[8] = pol ? 20.0 : 21.0;
  o.order = i; o.keep = 1; o.pad = 0;
  out[i] = o;
  if (table) {
    double* r = table + (size_t)i * 13;
    r[0] = pol; r[1] = e.minI; r[2] = e.maxI; r[3] = e.thresh; r[4] = e.margin; r[5] = e.area; r[6] = e.border; r[7] = sm.nruns;
    r[8] = sm.m.cx; r[9] = sm.m.cy; r[10] = sm.m.sxx; r[11] = sm.m.sxy; r[12] = sm.m.syy;
  }
}

inline int grid_for(uint32_t n, int threads, int max_blocks) {
  const long long b = ((long long)n + threads - 1) / threads;
  return (int)std::max(1LL, std::min<long long>(b, max_blocks));
}

MserBufs* mser_bufs(mb2_ctx* ctx) {
  if (!ctx->mser_state) ctx->mser_state = new MserBufs();
  return (MserBufs*)ctx->mser_state;
}

#define MSER_SYNC_COUNTERS()                                                                                                  \
  do {                                                                                                                        \
    MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(hc, dC, sizeof(MserCounters), cudaMemcpyDeviceToHost, st));                           \
    MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(st));                                                                           \
    if (hc->cap_overflow) { ctx->set_error("mser: internal list capacity exceeded"); return MB2_ERR_CAPACITY; }               \
  } while (0)

// One polarity.  Appends the keys (reference order) to d_out[*n_out ...]; table (optional) receives 13 doubles per region.
int mser_polarity(mb2_ctx* ctx, const ImgView& img, int pol, const mb2_mser_params& par, double min_margin, int as_regions, KeyOut* d_out,
                  int out_cap, int* n_out, double* d_table) {
  MserBufs& B = *mser_bufs(ctx);
  cudaStream_t st = ctx->stream;
  const int W = img.cols, H = img.rows;
  const uint32_t N = (uint32_t)W * H;
  const int G = ctx->num_sms * 8;
  const int track_size = std::min(10000, par.min_size);
  const int max_size = (int)((double)W * (double)H * par.max_area);
  const uint32_t list_cap = N;  // every list below holds at most one entry per pixel, except runs (checked on the device)

  MB2_CUDA_CHECK(ctx, B.counters.reserve(sizeof(MserCounters)));
  MB2_CUDA_CHECK(ctx, B.h_counters.reserve(sizeof(MserCounters)));
  MserCounters* dC = B.counters.as<MserCounters>();
  MserCounters* hc = B.h_counters.as<MserCounters>();
  MB2_CUDA_CHECK(ctx, B.lev.reserve(N)); MB2_CUDA_CHECK(ctx, B.keys8.reserve(N));
  DevBuf* u32bufs[] = {&B.zpar, &B.parent, &B.area, &B.nedge, &B.order, &B.order_in, &B.hooked, &B.surv, &B.birth, &B.slot_of_node, &B.sa};
  for (DevBuf* b : u32bufs) MB2_CUDA_CHECK(ctx, b->reserve((size_t)N * 4));
  MB2_CUDA_CHECK(ctx, B.best.reserve((size_t)N * 8));
  MB2_CUDA_CHECK(ctx, B.flag.reserve(N));
  uint8_t* lev = B.lev.as<uint8_t>();
  uint32_t *zpar = B.zpar.as<uint32_t>(), *parent = B.parent.as<uint32_t>(), *area = B.area.as<uint32_t>(), *nedge = B.nedge.as<uint32_t>();
  uint32_t *order = B.order.as<uint32_t>(), *order_in = B.order_in.as<uint32_t>(), *hooked = B.hooked.as<uint32_t>();
  uint32_t *surv = B.surv.as<uint32_t>(), *birth = B.birth.as<uint32_t>(), *slot_of_node = B.slot_of_node.as<uint32_t>(), *sa = B.sa.as<uint32_t>();

  // 1. u8 image, histogram, (level, raster) order
  MB2_CUDA_CHECK(ctx, cudaMemsetAsync(dC, 0, sizeof(MserCounters), st));
  MB2_LAUNCH(ctx, k_mser_prep, grid_for(N, 256, G), 256, 0, img.p, img.pitch, W, H, pol, lev, zpar, parent, area, order_in, dC);
  MB2_LAUNCH(ctx, k_mser_offsets, 1, 32, 0, dC);
  {
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, lev, B.keys8.as<uint8_t>(), order_in, order, (int)N, 0, 8, st);
    MB2_CUDA_CHECK(ctx, B.cub_tmp.reserve(tmp));
    cub::DeviceRadixSort::SortPairs(B.cub_tmp.p, tmp, lev, B.keys8.as<uint8_t>(), order_in, order, (int)N, 0, 8, st);
    ctx->launches += 3;
  }
  // 2. component tree
  for (int L = 0; L < 256; L++) {
    MB2_LAUNCH(ctx, k_mser_union, G, 256, 0, L, W, H, lev, order, zpar, nedge, hooked, dC);
    MB2_LAUNCH(ctx, k_mser_final, G, 256, 0, L, zpar, parent, area, nedge, hooked, dC);
  }
  MB2_LAUNCH(ctx, k_mser_root, 1, 32, 0, zpar, dC);
  // 3. survivors
  TreeDev td{W, H, lev, parent, area, nedge, dC, track_size};
  MB2_CUDA_CHECK(ctx, cudaMemsetAsync(B.best.p, 0, (size_t)N * 8, st));
  MB2_CUDA_CHECK(ctx, cudaMemsetAsync(B.flag.p, 0, N, st));
  MB2_CUDA_CHECK(ctx, cudaMemsetAsync(surv, 0xff, (size_t)N * 4, st));
  MB2_CUDA_CHECK(ctx, cudaMemsetAsync(birth, 0xff, (size_t)N * 4, st));
  MB2_CUDA_CHECK(ctx, B.emu_nodes.reserve((size_t)N * 4));
  MB2_CUDA_CHECK(ctx, B.own_a.reserve((size_t)N * 8));
  MB2_CUDA_CHECK(ctx, B.own_b.reserve((size_t)N * 8));
  MB2_LAUNCH(ctx, k_mser_best, grid_for(N, 256, G), 256, 0, td, N, B.best.as<u64>());
  MB2_LAUNCH(ctx, k_mser_ties, grid_for(N, 256, G), 256, 0, td, N, B.best.as<u64>(), B.flag.as<uint8_t>());
  MB2_LAUNCH(ctx, k_mser_classify, grid_for(N, 256, G), 256, 0, td, N, B.best.as<u64>(), B.flag.as<uint8_t>(), surv, B.emu_nodes.as<uint32_t>(), list_cap, dC);
  MB2_LAUNCH(ctx, k_mser_ownkeys, grid_for(N, 256, G), 256, 0, td, N, B.flag.as<uint8_t>(), B.own_a.as<u64>(), list_cap, dC);
  MSER_SYNC_COUNTERS();
  const uint32_t n_emu = hc->emu_nodes, n_own = hc->own_keys;
  const u64* own_sorted = B.own_a.as<u64>();
  if (n_own > 0) {
    int bits = 33; while (bits < 64 && (1ull << (bits - 32)) < (u64)N) bits++;
    size_t tmp = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp, B.own_a.as<u64>(), B.own_b.as<u64>(), (int)n_own, 0, bits, st);
    MB2_CUDA_CHECK(ctx, B.cub_tmp.reserve(tmp));
    cub::DeviceRadixSort::SortKeys(B.cub_tmp.p, tmp, B.own_a.as<u64>(), B.own_b.as<u64>(), (int)n_own, 0, bits, st);
    ctx->launches += 4;
    own_sorted = B.own_b.as<u64>();
  }
  if (n_emu > 0) {
    DevBuf* e32[] = {&B.uf, &B.esz, &B.epre, &B.eid, &B.ebirth};
    for (DevBuf* b : e32) MB2_CUDA_CHECK(ctx, b->reserve((size_t)N * 8));
    MB2_CUDA_CHECK(ctx, B.ekind.reserve((size_t)N * 2));
    MB2_CUDA_CHECK(ctx, cudaMemsetAsync(B.uf.p, 0xff, (size_t)N * 8, st));
    EmuScratch es{N, B.uf.as<uint32_t>(), B.esz.as<uint32_t>(), B.epre.as<uint32_t>(), B.eid.as<uint32_t>(), B.ebirth.as<uint32_t>(), B.ekind.as<uint8_t>()};
    MB2_LAUNCH(ctx, k_mser_emulate, (n_emu + 63) / 64, 64, 0, td, N, B.emu_nodes.as<uint32_t>(), n_emu, own_sorted, n_own, es, surv, birth, dC);
  }
  // 4. regions and their stability thresholds
  int n_sel = 0, n_slots = 0;
  MB2_CUDA_CHECK(ctx, cudaMemsetAsync(slot_of_node, 0xff, (size_t)N * 4, st));
  if (n_emu > 0) {
    MB2_CUDA_CHECK(ctx, B.longr.reserve((size_t)n_emu * sizeof(LongRegion)));
    MB2_LAUNCH(ctx, k_mser_regions_a, (n_emu + 127) / 128, 128, 0, td, B.emu_nodes.as<uint32_t>(), n_emu, surv, min_margin, B.longr.as<LongRegion>(), n_emu, dC);
    MSER_SYNC_COUNTERS();
    if (hc->overflow) { ctx->set_error("mser: a min_reg label absorbed >= 32768 pixels (the reference's 15-bit size field wraps here)"); return MB2_ERR_UNSUPPORTED; }
    const uint32_t n_long = hc->long_regions;
    if (n_long > 0) {
      const uint32_t sel_cap = n_long * 8 + 1024;
      MB2_CUDA_CHECK(ctx, B.sel.reserve((size_t)sel_cap * sizeof(SelRec)));
      MB2_CUDA_CHECK(ctx, B.node_of_slot.reserve((size_t)sel_cap * 4));
      const size_t smem = (size_t)2 * 256 * MSER_RB_THREADS * sizeof(int);
      MB2_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_mser_regions_b, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      MB2_LAUNCH(ctx, k_mser_regions_b, (n_long + MSER_RB_THREADS - 1) / MSER_RB_THREADS, MSER_RB_THREADS, smem, td, B.longr.as<LongRegion>(), n_long, surv,
                 birth, min_margin, par.min_size, max_size, slot_of_node, B.node_of_slot.as<uint32_t>(), B.sel.as<SelRec>(), sel_cap, dC);
      MSER_SYNC_COUNTERS();
      if (hc->thr_overflow) { ctx->set_error("mser: more than 128 stability thresholds in one region"); return MB2_ERR_CAPACITY; }
      n_sel = (int)hc->n_sel; n_slots = n_sel;
    }
  }
  if (n_sel == 0) return MB2_OK;
  if (*n_out + n_sel > out_cap) { ctx->set_error("mser: output capacity too small"); *n_out += n_sel; return MB2_ERR_CAPACITY; }
  // reference list order
  MB2_CUDA_CHECK(ctx, B.selkey_a.reserve((size_t)n_sel * 8)); MB2_CUDA_CHECK(ctx, B.selkey_b.reserve((size_t)n_sel * 8));
  MB2_CUDA_CHECK(ctx, B.selidx_a.reserve((size_t)n_sel * 4)); MB2_CUDA_CHECK(ctx, B.selidx_b.reserve((size_t)n_sel * 4));
  MB2_LAUNCH(ctx, k_mser_selkeys, (n_sel + 255) / 256, 256, 0, B.sel.as<SelRec>(), (uint32_t)n_sel, B.selkey_a.as<u64>(), B.selidx_a.as<uint32_t>());
  {
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, B.selkey_a.as<u64>(), B.selkey_b.as<u64>(), B.selidx_a.as<uint32_t>(), B.selidx_b.as<uint32_t>(), n_sel, 0, 48, st);
    MB2_CUDA_CHECK(ctx, B.cub_tmp.reserve(tmp));
    cub::DeviceRadixSort::SortPairs(B.cub_tmp.p, tmp, B.selkey_a.as<u64>(), B.selkey_b.as<u64>(), B.selidx_a.as<uint32_t>(), B.selidx_b.as<uint32_t>(), n_sel, 0, 48, st);
    ctx->launches += 4;
  }
  // 5. runs of the selected components
  MB2_CUDA_CHECK(ctx, cudaMemsetAsync(sa, 0xff, (size_t)N * 4, st));
  MB2_CUDA_CHECK(ctx, B.up_sel.reserve((size_t)n_slots * 4));
  MB2_LAUNCH(ctx, k_mser_down_init, 1, 32, 0, sa, slot_of_node, dC);
  for (int L = 255; L >= 0; L--) MB2_LAUNCH(ctx, k_mser_down, G, 256, 0, L, parent, hooked, slot_of_node, sa, dC);
  MB2_LAUNCH(ctx, k_mser_upsel, (n_slots + 255) / 256, 256, 0, parent, B.node_of_slot.as<uint32_t>(), (uint32_t)n_slots, sa, B.up_sel.as<uint32_t>(), dC);
  const uint32_t run_cap = 2 * N + 1024;
  MB2_CUDA_CHECK(ctx, B.ev_a.reserve((size_t)run_cap * 8)); MB2_CUDA_CHECK(ctx, B.ev_b.reserve((size_t)run_cap * 8));
  MB2_LAUNCH(ctx, k_mser_runs, grid_for(N, 256, G), 256, 0, W, H, sa, B.up_sel.as<uint32_t>(), B.ev_a.as<u64>(), B.ev_b.as<u64>(), run_cap, dC);
  MSER_SYNC_COUNTERS();
  const uint32_t n_runs = hc->n_starts;
  if (hc->n_ends != n_runs) { ctx->set_error("mser: run start/end mismatch"); return MB2_ERR_CUDA; }
  MB2_CUDA_CHECK(ctx, B.ev_c.reserve((size_t)n_runs * 8 + 8)); MB2_CUDA_CHECK(ctx, B.ev_d.reserve((size_t)n_runs * 8 + 8));
  {
    int bits = 33; while (bits < 64 && (1ull << (bits - 32)) < (u64)n_slots) bits++;
    size_t tmp = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp, B.ev_a.as<u64>(), B.ev_c.as<u64>(), (int)n_runs, 0, bits, st);
    MB2_CUDA_CHECK(ctx, B.cub_tmp.reserve(tmp));
    cub::DeviceRadixSort::SortKeys(B.cub_tmp.p, tmp, B.ev_a.as<u64>(), B.ev_c.as<u64>(), (int)n_runs, 0, bits, st);
    cub::DeviceRadixSort::SortKeys(B.cub_tmp.p, tmp, B.ev_b.as<u64>(), B.ev_d.as<u64>(), (int)n_runs, 0, bits, st);
    ctx->launches += 8;
  }
  // 6. moments and keys
  MB2_CUDA_CHECK(ctx, B.mom.reserve((size_t)n_slots * sizeof(SlotMoments)));
  MB2_LAUNCH(ctx, k_mser_moments, (n_slots + 63) / 64, 64, 0, B.ev_c.as<u64>(), B.ev_d.as<u64>(), n_runs, (uint32_t)n_slots, B.mom.as<SlotMoments>());
  MB2_LAUNCH(ctx, k_mser_keys, (n_sel + 127) / 128, 128, 0, B.sel.as<SelRec>(), B.selidx_b.as<uint32_t>(), (uint32_t)n_sel, B.mom.as<SlotMoments>(), slot_of_node, pol,
             as_regions, d_out + *n_out, d_table ? d_table + (size_t)*n_out * 13 : nullptr);
  *n_out += n_sel;
  return MB2_OK;
}

}  // namespace MB2_NS

void mb2_mser_release(mb2_ctx* ctx) {
  if (ctx->mser_state) { ((MB2_NS::MserBufs*)ctx->mser_state)->release(); delete (MB2_NS::MserBufs*)ctx->mser_state; ctx->mser_state = nullptr; }
}

// MSER+ then MSER- (getRLEExtrema, libExtrema.cpp:462-482) on the device image; FIXED_TH only on the device, the other
// DetectorModes re-order the keys with the reference's own std::sort on the host (extrema.cpp:31-90).
// Result: ctx->kp_b holds *n_out KeyOut records in the reference's order.
int mb2_mser_core(mb2_ctx* ctx, const ImgView& img, const mb2_mser_params& par, double tilt, double zoom, int as_regions, int* n_out,
                  double* d_table_out /* optional device table, capacity rows */, int capacity) {
  using namespace MB2_NS;
  *n_out = 0;
  if (par.min_size < 2) { ctx->set_error("mser: min_size < 2 is not supported"); return MB2_ERR_UNSUPPORTED; }
  if (par.relative) { ctx->set_error("mser: relative margins are not supported"); return MB2_ERR_UNSUPPORTED; }
  if (img.cols > 65535 || img.rows > 65535 || (long long)img.cols * img.rows >= (1LL << 31)) { ctx->set_error("mser: image too large"); return MB2_ERR_ARG; }
  const double min_margin = par.mode != 0 ? 1.0 : par.min_margin;  // extrema.cpp:294-299
  const int cap = capacity > 0 ? capacity : (int)std::min<long long>((long long)img.cols * img.rows / 16 + 4096, 4000000);
  MB2_CUDA_CHECK(ctx, ctx->kp_b.reserve((size_t)cap * sizeof(KeyOut)));
  int rc;
  const bool sort_on_host = par.mode != 0;
  for (int pol = 0; pol < 2; pol++)
    if ((rc = mser_polarity(ctx, img, pol, par, min_margin, sort_on_host ? 0 : as_regions, ctx->kp_b.as<KeyOut>(), cap, n_out, d_table_out))) return rc;
  if (sort_on_host && *n_out > 0) {
    // prepareKeysForExport (extrema.cpp:31-90): the reference's own (unstable) std::sort decides the order of equal margins,
    // so this step runs through the same library routine on the host.
    int reg_number = par.reg_number;
    if ((tilt > 2.0) || (zoom < 0.5)) reg_number = (int)std::floor(zoom * 2.0 * reg_number / tilt);
    const int n = *n_out;
    std::vector<KeyOut> keys(n);
    MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(keys.data(), ctx->kp_b.p, (size_t)n * sizeof(KeyOut), cudaMemcpyDeviceToHost, ctx->stream));
    MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    auto cmp = [](const KeyOut& a, const KeyOut& b) { return std::fabs(a.v[7]) > std::fabs(b.v[7]); };
    std::sort(keys.begin(), keys.end(), cmp);
    const double maxResponse = std::fabs(keys[0].v[7]);
    KeyOut probe = keys[0];
    switch (par.mode) {
      case 1: probe.v[7] = maxResponse * par.rel_threshold; keys.resize(std::lower_bound(keys.begin(), keys.end(), probe, cmp) - keys.begin()); break;
      case 2: if (reg_number < n && reg_number >= 0) keys.resize(reg_number); break;
      case 3: keys.resize((int)std::floor(par.rel_reg_number * (double)n)); break;
      case 4: {
        probe.v[7] = min_margin;
        const int fix = (int)(std::lower_bound(keys.begin(), keys.end(), probe, cmp) - keys.begin());
        keys.resize(fix < reg_number ? std::min(reg_number, n) : std::min(fix, n));
        break; }
      default: break;
    }
    if (as_regions)
      for (KeyOut& k : keys) {  // DetectAffineRegions post-step (synth-detection.hpp:110-124, synth-detection.cpp:46-55)
        double a = k.v[2], b = k.v[3], c = k.v[4], d = k.v[5];
        k.v[6] = k.v[6] * std::sqrt(std::fabs(a * d - b * c));
        const double det = std::sqrt(std::fabs(a * d - b * c)), b2a2 = std::sqrt(b * b + a * a);
        k.v[2] = b2a2 / det; k.v[3] = 0; k.v[4] = (d * b + c * a) / (b2a2 * det); k.v[5] = det / b2a2;
      }
    for (size_t i = 0; i < keys.size(); i++) keys[i].order = i;
    *n_out = (int)keys.size();
    if (*n_out > 0) {
      MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->kp_b.p, keys.data(), keys.size() * sizeof(KeyOut), cudaMemcpyHostToDevice, ctx->stream));
      MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    }
  }
  return MB2_OK;
}
