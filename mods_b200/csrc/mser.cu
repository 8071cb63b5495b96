// MSER detector on the GPU: replaces DetectMSERs (detectors/mser/extrema/extrema.cpp:284-473)
// and everything below it (libExtrema.cpp, sortPixels.cpp, getExtrema.cpp, optThresh.cpp, boundary.cpp).
//
// The reference is sequential in (intensity, raster) order.  Here both polarities are handled as ONE problem: the
// inverted image (MSER-) is stacked under the image (MSER+) in one pixel index space of 2*W*H pixels whose halves are never
// neighbours, so every pass below serves both at once:
//   1. k_mtree_tiles                 u8 image of both polarities; component tree of every 64 x 64 tile built in shared memory by
//                                    lock-free merging of all inner pixel pairs at once (mser_tree_build.cuh)         [shared memory]
//   2. k_mtree_local / borders /     canonical form + area / inner-edge sums of every subtree inside its tile; the same merge over the
//      fix                           pixel pairs across tile borders (3 % of the edges) on global words; final parents, and the local
//                                    sums added along the gaps the border merge opened (nothing is sorted, no pass per level)
//                                                                                                   [L2 atomics / HBM streaming]
//   3. k_mser_best .. k_mser_emulate survivor of every merge = largest tracked child; nodes where the reference's
//                                    answer depends on its pixel order (births, equal sizes) are replayed by one
//                                    thread each (mser_logic.cuh: emulate_node)
//   4. k_mser_regions_*              one thread per tracked region: lifetime, cumulative area/border histograms,
//                                    stability thresholds (FastSetOptThresholds4StableRegion)
//   5. k_mser_sa / k_mser_runs       selected components -> row runs (start / end events), sorted per region
//   6. k_mser_moments / k_mser_keys  RLE2Ellipse in the reference's summation order, sqrtm, AffineKeypoint record
// Integer / index work is bit-exact by construction; the f64 moment sums follow the reference's order run by run.
#include "common.cuh"
#undef MB2_NS
#define MB2_NS mb2_mser_detail
#include "pyramid.cuh"
#include "mser_logic.cuh"
#include "mser_tree_build.cuh"

#include <cooperative_groups.h>
#include <cub/cub.cuh>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace MB2_NS {
using namespace mser_logic;
namespace cg = cooperative_groups;
typedef unsigned long long u64;

struct MserCounters {
  uint32_t hook_cnt, n_walk, emu_nodes, own_keys, long_regions, n_sel, n_slots, n_starts, n_ends;
  uint32_t overflow, thr_overflow, cap_overflow, n_stack, root[8], n_img[4];   // n_stack = 2 x images; n_img = regions per image
};

struct SelRec {  // one (region, threshold): a row of getRLEExtrema's output
  u64 key;       // birth level << 40 | promotion time << 8 | threshold rank: the reference's list order
  uint32_t node, slot;
  int minI, maxI, thresh, margin, area, border;
};

struct MserBufs {
  DevBuf lev, gpar, parent, area, nedge, lsum, hi, best, surv, birth, flag, emu_nodes, own_a, own_b;
  DevBuf uf, esz, epre, eid, ebirth, ekind, longr, sel, sel_sorted, selkey_a, selkey_b, selidx_a, selidx_b, slot_of_node, node_of_slot;
  DevBuf sa, up_sel, ev_a, ev_b, ev_c, ev_d, mom, cub_tmp, counters, table;
  HostBuf h_counters;
  void release() {
    DevBuf* all[] = {&lev, &gpar, &parent, &area, &nedge, &lsum, &hi, &best, &surv, &birth, &flag, &emu_nodes, &own_a,
                     &own_b, &uf, &esz, &epre, &eid, &ebirth, &ekind, &longr, &sel, &sel_sorted, &selkey_a, &selkey_b, &selidx_a, &selidx_b,
                     &slot_of_node, &node_of_slot, &sa, &up_sel, &ev_a, &ev_b, &ev_c, &ev_d, &mom, &cub_tmp, &counters, &table};
    for (DevBuf* b : all) b->release();
    h_counters.release();
  }
};

// ---- 1 + 2. component trees by lock-free merging (mser_tree_build.cuh) -----------------------------------------------------------
// k_mtree_tiles    one CTA per 64 x 64 tile: float -> u8 as extrema.cpp:401-403 does ((unsigned char) of the float: truncation) and its
//                  inversion (InvertImageAndHistogram, sortPixels.cpp:131-153); the trees of BOTH polarities of the tile are built in
//                  shared memory by connect() over the tile's inner edges, all threads at once; written out as global words
// k_mtree_local    per (tile, polarity): canonical words, area / inner-edge sums of every subtree inside the tile
// k_mtree_borders  connect() over the edges that cross tile borders (3 % of all edges), one thread each, on the global words
// k_mtree_fix      canonical parent of every pixel in the final tree; local sums added along the gaps the border merge opened
// Words: 32 bit (level << 24 | pixel index inside its image) up to 2^24 pixels per image, else 64 bit (level << 32 | index).
typedef mser_tree::KeyT<uint32_t, 24> GKey32;
typedef mser_tree::KeyT<unsigned long long, 32> GKey64;
typedef mser_tree::KeyT<uint32_t, 12> TKey;      // inside a tile: 4096 elements
#define MT_TILE 64
#define MT_THREADS 256
#define MT_TILES_THREADS 512   // k_mtree_tiles: 3 CTAs of 512 per SM beat 6 of 256 (late stages have few pixel pairs per tile)

struct SmemWords {
  uint32_t* w;
  __device__ __forceinline__ uint32_t load(uint32_t i) { return *(volatile uint32_t*)(w + i); }
  __device__ __forceinline__ void store(uint32_t i, uint32_t v) { *(volatile uint32_t*)(w + i) = v; }
  __device__ __forceinline__ bool cas(uint32_t i, uint32_t expect, uint32_t desired) { return atomicCAS(w + i, expect, desired) == expect; }
};
struct SmemWordsOwned {   // words only the calling thread touches (the sequential 4 x 4 phase of k_mtree_tiles)
  uint32_t* w;
  __device__ __forceinline__ uint32_t load(uint32_t i) { return w[i]; }
  __device__ __forceinline__ void store(uint32_t i, uint32_t v) { w[i] = v; }
  __device__ __forceinline__ bool cas(uint32_t i, uint32_t, uint32_t desired) { w[i] = desired; return true; }
};
template <class WordT>
struct GmemWords {   // L2-coherent accesses: a stale L1 line would make a failed compare-and-swap repeat forever
  WordT* w;
  __device__ __forceinline__ WordT load(uint32_t i) { return __ldcg(w + i); }
  __device__ __forceinline__ void store(uint32_t i, WordT v) { __stcg(w + i, v); }
  __device__ __forceinline__ bool cas(uint32_t i, WordT expect, WordT desired) { return atomicCAS(w + i, expect, desired) == expect; }
};

template <class WordT>
struct GmemWordsRO {   // kernels that only READ the finished words (k_mtree_fix, k_mtree_walk): through L1 -- the pixels of a warp mostly name the same few nodes
  const WordT* w;
  __device__ __forceinline__ WordT load(uint32_t i) { return __ldg(w + i); }
};

struct MserImages { const float* p[4]; int pitch[4]; int n; };   // same-size images processed together (a pair: n = 2)

template <class GK>
__global__ void __launch_bounds__(MT_TILES_THREADS) k_mtree_tiles(MserImages im, int W, int H, uint8_t* __restrict__ lev, typename GK::word* __restrict__ gpar,
                                                            MserCounters* __restrict__ C) {
  __shared__ uint32_t par[2][MT_TILE * MT_TILE];
  __shared__ uint8_t sv[MT_TILE * MT_TILE];
  const int k = blockIdx.z, x0 = blockIdx.x * MT_TILE, y0 = blockIdx.y * MT_TILE;
  const int tw = min(MT_TILE, W - x0), th = min(MT_TILE, H - y0);
  const uint32_t Nimg = (uint32_t)W * H;
  for (int i = threadIdx.x; i < MT_TILE * MT_TILE; i += MT_TILES_THREADS) {
    const int lx = i & (MT_TILE - 1), ly = i >> 6;
    int v = 0;
    if (lx < tw && ly < th) v = ((int)im.p[k][(size_t)(y0 + ly) * im.pitch[k] + x0 + lx]) & 0xff;
    par[0][i] = TKey::make(v, i); par[1][i] = TKey::make(255 - v, i); sv[i] = (uint8_t)v;
  }
  __syncthreads();
  // Merge order: small blocks first, up to the two halves of the tile (12 stages; stage s joins blocks across b = 1 << (s >> 1) wide gaps,
  // even s left-right, odd s top-bottom).  A join of two trees walks both root paths once ("zipping"), so joins should happen while the
  // trees are small; the other pixel pairs of a block border find their paths already merged.
  //   stages 0-3: ONE thread builds the tree of a whole 4 x 4 block (24 pixel pairs in turn; nobody else touches the block's words, so
  //               plain loads / stores instead of compare-and-swap, no barriers, every lane busy with the same amount of work);
  //   stages 4-11: one thread per pixel pair across the block border, all pairs of a border zipping concurrently (connect() is lock free).
  // Measured at 4096 x 3072, both polarities (tools/micro/mtree_tiles.cu, every variant checked node by node against this one):
  // all inner edges at once in arbitrary order 6.1 ms; 12 cooperative stages, 256 threads 1.87 ms; 4 x 4 blocks per thread + 8 stages,
  // 512 threads 1.33 ms (this kernel); 8 x 8 per thread 1.49; block pairs joined by one thread each up to stage 8 / 12: 2.0 / 3.8 ms; one
  // middle pair of every border first, then the rest: 1.6 - 2.2 ms; pixels in level order with union-find shortcuts and a barrier per
  // level 11 ms (too few pixels per level and tile to keep the warps busy between barriers).
  for (int t = threadIdx.x; t < 2 * (MT_TILE / 4) * (MT_TILE / 4); t += MT_TILES_THREADS) {
    const int pol = t & 1, u = t >> 1, bx = (u & 15) * 4, by = (u >> 4) * 4;
    SmemWordsOwned m{par[pol]};
    for (int s = 0; s < 4; s++) {
      const bool horiz = !(s & 1);
      const int b = 1 << (s >> 1);
      for (int line = 0; line < 4; line++)
        for (int c = b - 1; c < 3; c += 2 * b) {
          const int lx = bx + (horiz ? c : line), ly = by + (horiz ? line : c);
          if ((horiz ? lx + 1 : lx) >= tw || (horiz ? ly : ly + 1) >= th) continue;
          const int i = ly * MT_TILE + lx, j = horiz ? i + 1 : i + MT_TILE;
          const int vi = pol ? 255 - sv[i] : sv[i], vj = pol ? 255 - sv[j] : sv[j];
          mser_tree::connect<TKey>(m, TKey::make(vi, i), TKey::make(vj, j));
        }
    }
  }
  __syncthreads();
  for (int s = 4; s < 12; s++) {
    const bool horiz = !(s & 1);
    const int b = 1 << (s >> 1), per_line = (MT_TILE / 2) / b, ntask = 2 * MT_TILE * per_line;
    for (int t = threadIdx.x; t < ntask; t += MT_TILES_THREADS) {
      const int pol = t & 1, u = t >> 1, line = u / per_line, c = (2 * (u - line * per_line) + 1) * b - 1;
      const int lx = horiz ? c : line, ly = horiz ? line : c;
      if ((horiz ? lx + 1 : lx) >= tw || (horiz ? ly : ly + 1) >= th) continue;
      const int i = ly * MT_TILE + lx, j = horiz ? i + 1 : i + MT_TILE;
      const int vi = pol ? 255 - sv[i] : sv[i], vj = pol ? 255 - sv[j] : sv[j];   // the word of i may already point elsewhere: its own level is the image value
      SmemWords m{par[pol]};
      mser_tree::connect<TKey>(m, TKey::make(vi, i), TKey::make(vj, j));
    }
    __syncthreads();
  }
  // write-out: tile-local index -> index inside the image; sub-image 2k = MSER+, 2k + 1 = MSER-
  for (int i = threadIdx.x; i < MT_TILE * MT_TILE; i += MT_TILES_THREADS) {
    const int lx = i & (MT_TILE - 1), ly = i >> 6;
    if (lx >= tw || ly >= th) continue;
    const uint32_t pi = (uint32_t)(y0 + ly) * W + x0 + lx;
#pragma unroll
    for (int pol = 0; pol < 2; pol++) {
      const uint32_t w = par[pol][i];
      const uint32_t j = TKey::idx(w);
      const uint32_t pj = (uint32_t)(y0 + (j >> 6)) * W + x0 + (j & (MT_TILE - 1));
      const size_t a = (size_t)(2 * k + pol) * Nimg + pi;
      gpar[a] = GK::make(TKey::lev(w), pj);
    }
    const int v = sv[i];
    lev[(size_t)(2 * k) * Nimg + pi] = (uint8_t)v; lev[(size_t)(2 * k + 1) * Nimg + pi] = (uint8_t)(255 - v);
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) C->n_stack = 2 * im.n;
}
// Per (tile, polarity), after the tile's tree is complete and before the borders are merged: the tile's words are rewritten in canonical
// form (a pixel points to the representative of its node, a representative to the representative of its parent node), and every
// representative gets the area / inner-edge count of its subtree INSIDE the tile (lsum: area | edges << 16) together with the level
// of its parent in the tile (hi; 256 for the tile's root).  k_mtree_fix turns these local sums into the sums of the final tree.
// nedge of a pixel = 4-neighbours the reference has already labelled when it reaches the pixel: lower level, or the same level and
// earlier in raster order (getExtrema.cpp:216-263) -- neighbours across the tile border included; summed over a component it counts the
// pixel pairs inside, so that the reference's border count (+4 - 2 x labelled neighbours per pixel) is 4 area - 2 edges.
template <class GK>
__global__ void __launch_bounds__(MT_THREADS) k_mtree_local(int W, int H, const uint8_t* __restrict__ lev, typename GK::word* __restrict__ gpar,
                                                            uint32_t* __restrict__ lsum, uint16_t* __restrict__ hi, uint32_t* __restrict__ area,
                                                            uint32_t* __restrict__ nedge) {
  __shared__ uint32_t par[MT_TILE * MT_TILE];
  __shared__ uint32_t acc[MT_TILE * MT_TILE];
  __shared__ uint32_t pend[MT_TILE * MT_TILE / 2];        // children (representatives) that have not yet delivered their subtree sums: 16-bit counters, two per word
  __shared__ uint8_t sl[(MT_TILE + 2) * (MT_TILE + 2)];   // levels with a one-pixel halo
  const int sub = blockIdx.z, x0 = blockIdx.x * MT_TILE, y0 = blockIdx.y * MT_TILE;
  const int tw = min(MT_TILE, W - x0), th = min(MT_TILE, H - y0);
  const uint32_t Nimg = (uint32_t)W * H;
  const uint8_t* l = lev + (size_t)sub * Nimg;
  typename GK::word* gp = gpar + (size_t)sub * Nimg;
  const int SW = MT_TILE + 2;
  for (int i = threadIdx.x; i < SW * SW; i += MT_THREADS) {
    const int gx = x0 + (i % SW) - 1, gy = y0 + (i / SW) - 1;
    sl[i] = (gx >= 0 && gy >= 0 && gx < W && gy < H) ? l[(size_t)gy * W + gx] : 0;
  }
  for (int i = threadIdx.x; i < MT_TILE * MT_TILE; i += MT_THREADS) {
    const int lx = i & (MT_TILE - 1), ly = i >> 6;
    acc[i] = 0; if (!(i & 1)) pend[i >> 1] = 0;
    if (lx < tw && ly < th) {
      const typename GK::word w = gp[(size_t)(y0 + ly) * W + x0 + lx];
      const uint32_t pj = GK::idx(w);
      const int jy = (int)(pj / (uint32_t)W) - y0, jx = (int)(pj % (uint32_t)W) - x0;
      par[i] = TKey::make(GK::lev(w), (uint32_t)(jy * MT_TILE + jx));
    } else par[i] = TKey::make(0, i);
  }
  __syncthreads();
  // canonical form, in place: a word replaced by its canonical value is still a valid pointer for every other walk
  SmemWords m{par};
  for (int i = threadIdx.x; i < MT_TILE * MT_TILE; i += MT_THREADS) {
    const int lx = i & (MT_TILE - 1), ly = i >> 6;
    if (lx >= tw || ly >= th) continue;
    const int L = sl[(ly + 1) * SW + lx + 1];
    par[i] = mser_tree::canonical_parent<TKey>(m, TKey::make(L, i));
  }
  __syncthreads();
  // own counts of every node; every representative announces itself to its parent
  for (int i = threadIdx.x; i < MT_TILE * MT_TILE; i += MT_THREADS) {
    const int lx = i & (MT_TILE - 1), ly = i >> 6;
    if (lx >= tw || ly >= th) continue;
    const int c = (ly + 1) * SW + lx + 1, L = sl[c];
    const uint32_t w = par[i];
    const bool self = w == TKey::make(L, i), rep = self || TKey::lev(w) > L;
    uint32_t e = 0;
    if (y0 + ly > 0) e += sl[c - SW] <= L;              // q < p: lower or equal level
    if (x0 + lx > 0) e += sl[c - 1] <= L;
    if (x0 + lx < W - 1) e += sl[c + 1] < L;            // q > p: strictly lower level
    if (y0 + ly < H - 1) e += sl[c + SW] < L;
    atomicAdd(&acc[rep ? i : TKey::idx(w)], 1u | (e << 16));
    if (rep && !self) atomicAdd(&pend[TKey::idx(w) >> 1], 1u << (16 * (TKey::idx(w) & 1)));
  }
  __syncthreads();
  // Subtree sums inside the tile WITHOUT a pass per level: the leaves start, every representative hands its finished sums to its parent,
  // and whoever delivers the LAST outstanding child of a node carries on with that node (one thread climbs the trunk at the end).  The
  // leaves are fixed before anybody climbs (a node whose counter drops to zero later must not be mistaken for one).
  uint32_t leaf_mask = 0;
  for (int i = threadIdx.x, q = 0; i < MT_TILE * MT_TILE; i += MT_THREADS, q++) {
    const int lx = i & (MT_TILE - 1), ly = i >> 6;
    if (lx >= tw || ly >= th) continue;
    const int L = sl[(ly + 1) * SW + lx + 1];
    const uint32_t w = par[i];
    if ((w == TKey::make(L, i) || TKey::lev(w) > L) && ((pend[i >> 1] >> (16 * (i & 1))) & 0xffffu) == 0) leaf_mask |= 1u << q;
  }
  __syncthreads();
  for (int i = threadIdx.x, q = 0; i < MT_TILE * MT_TILE; i += MT_THREADS, q++) {
    if (!((leaf_mask >> q) & 1u)) continue;
    uint32_t u = (uint32_t)i;
    for (;;) {
      const uint32_t p = TKey::idx(par[u]);
      if (p == u) break;                                        // the tile's root
      atomicAdd(&acc[p], *(volatile uint32_t*)&acc[u]);          // u's sums are complete: all its children delivered before the counter reached zero
      __threadfence_block();
      if (((atomicSub(&pend[p >> 1], 1u << (16 * (p & 1))) >> (16 * (p & 1))) & 0xffffu) != 1u) break;                 // somebody else delivers p's last child and goes on from there
      __threadfence_block();                                    // acquire side: p's other children added their sums before they decremented the counter
      u = p;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < MT_TILE * MT_TILE; i += MT_THREADS) {
    const int lx = i & (MT_TILE - 1), ly = i >> 6;
    if (lx >= tw || ly >= th) continue;
    const int L = sl[(ly + 1) * SW + lx + 1];
    const uint32_t w = par[i], j = TKey::idx(w);
    const size_t g = (size_t)(y0 + ly) * W + x0 + lx;
    gp[g] = GK::make(TKey::lev(w), (uint32_t)(y0 + (j >> 6)) * W + x0 + (j & (MT_TILE - 1)));
    const bool self = w == TKey::make(L, i), rep = self || TKey::lev(w) > L;
    const uint32_t a = rep ? acc[i] : 0u;
    lsum[(size_t)sub * Nimg + g] = a;
    hi[(size_t)sub * Nimg + g] = (uint16_t)(self ? 256 : TKey::lev(w));
    area[(size_t)sub * Nimg + g] = a & 0xffffu; nedge[(size_t)sub * Nimg + g] = a >> 16;   // a representative's own contribution to its own final node
  }
}

// edges across tile borders: first the horizontal borders (pixel pairs (x, y - 1) / (x, y) with y a multiple of the tile size: threads
// run along x), then the vertical ones
template <class GK>
__global__ void __launch_bounds__(256) k_mtree_borders(int W, int H, int n_sub, const uint8_t* __restrict__ lev, typename GK::word* gpar, int phase) {
  const uint32_t Nimg = (uint32_t)W * H;
  const uint32_t nh = (uint32_t)((H - 1) / MT_TILE) * W, nv = (uint32_t)((W - 1) / MT_TILE) * H, per = nh + nv;
  const unsigned long long total = (unsigned long long)per * n_sub;
  for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t sub = (uint32_t)(t / per), e = (uint32_t)(t - (unsigned long long)sub * per);
    uint32_t p, q;
    // phase 0: one pair in the middle of every tile border (these join the tile trees: the long walks up two root paths, without the
    // other 63 pairs of the border swapping the same words); phase 1: the rest
    if (e < nh) { const uint32_t b = e / W, x = e - b * W; if (phase < 2 && ((x & (MT_TILE - 1)) == MT_TILE / 2) != (phase == 0)) continue; q = (b + 1) * MT_TILE * W + x; p = q - W; }
    else { const uint32_t f = e - nh, b = f / H, y = f - b * H; if (phase < 2 && ((y & (MT_TILE - 1)) == MT_TILE / 2) != (phase == 0)) continue; q = y * W + (b + 1) * MT_TILE; p = q - 1; }
    const uint8_t* l = lev + (size_t)sub * Nimg;
    GmemWords<typename GK::word> m{gpar + (size_t)sub * Nimg};
    mser_tree::connect<GK>(m, GK::make(l[p], p), GK::make(l[q], q));
  }
}

// After the borders: canonical parent of every pixel in the final tree, and the final sums.  A representative u of a tile tree carries
// the sums of its subtree inside the tile; they belong to every node of the FINAL tree between u's node and the node of u's parent in
// the tile (exclusive: that one receives them through the parent's own local sums) -- one node when nothing was inserted from other
// tiles, the whole trunk for a tile's root.  Every representative walks that gap and adds its sums; no level-by-level pass is needed.
template <class GK>
__global__ void __launch_bounds__(256) k_mtree_fix(int W, int H, uint32_t N, const uint8_t* __restrict__ lev, const typename GK::word* __restrict__ gpar,
                                                   const uint32_t* __restrict__ lsum, const uint16_t* __restrict__ hi,
                                                   uint32_t* __restrict__ parent, uint32_t* __restrict__ work, MserCounters* C) {
  const uint32_t Nimg = (uint32_t)W * H;
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;   // one pixel per thread: the chains are short, the loads independent
  bool walk = false;
  if (x < N) {
    const uint32_t sub = x / Nimg, pi = x - sub * Nimg, off = sub * Nimg;
    const int L = lev[x];
    GmemWordsRO<typename GK::word> m{gpar + (size_t)off};
    const typename GK::word xk = GK::make(L, pi);
    const typename GK::word ck = mser_tree::canonical_parent<GK>(m, xk);
    parent[x] = off + GK::idx(ck);
    // a representative of a tile tree whose sums belong to more than its own node: it was merged into another node of its level, or
    // nodes of other tiles were inserted below its parent in the tile
    if (lsum[x] != 0) walk = ck == xk ? false : (GK::lev(ck) == L || GK::lev(ck) < (int)hi[x]);
  }
  const unsigned mk = __ballot_sync(0xffffffffu, walk);
  if (mk) {
    const int lane = threadIdx.x & 31;
    uint32_t b0 = 0;
    if (lane == 0) b0 = atomicAdd(&C->n_walk, (uint32_t)__popc(mk));
    b0 = __shfl_sync(0xffffffffu, b0, 0);
    if (walk) work[b0 + __popc(mk & ((1u << lane) - 1))] = x;
  }
}
template <class GK>
__global__ void __launch_bounds__(128) k_mtree_walk(int W, int H, const uint8_t* __restrict__ lev, const typename GK::word* __restrict__ gpar,
                                                    const uint32_t* __restrict__ lsum, const uint16_t* __restrict__ hi, const uint32_t* __restrict__ work,
                                                    const MserCounters* __restrict__ C, uint32_t* area, uint32_t* nedge) {
  const uint32_t Nimg = (uint32_t)W * H, n = C->n_walk;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t x = work[i];
    const uint32_t sub = x / Nimg, pi = x - sub * Nimg, off = sub * Nimg;
    const int L = lev[x];
    GmemWordsRO<typename GK::word> m{gpar + (size_t)off};
    const typename GK::word xk = GK::make(L, pi);
    const typename GK::word ck = mser_tree::canonical_parent<GK>(m, xk);
    const uint32_t ls = lsum[x], sa = ls & 0xffffu, se = ls >> 16;
    const int H_ = hi[x];
    typename GK::word vk = (ck == xk || GK::lev(ck) > L) ? xk : ck;   // the final node of x
    for (;;) {
      if (vk != xk) {                                                 // (x's own node already holds x's sums: k_mtree_local)
        atomicAdd(area + off + GK::idx(vk), sa);
        if (se) atomicAdd(nedge + off + GK::idx(vk), se);
      }
      const typename GK::word z = m.load(GK::idx(vk));
      if (z == vk) break;                                             // the root
      typename GK::word r = z;                                        // representative of the parent node
      for (;;) { const typename GK::word v = m.load(GK::idx(r)); if (v == r || GK::lev(v) != GK::lev(r)) break; r = v; }
      if (GK::lev(r) >= H_) break;
      vk = r;
    }
  }
}

// ---- 3. survivors -------------------------------------------------------------------------------------------------------
struct TreeDev {
  int W, H; const uint8_t* lev; const uint32_t* parent; const uint32_t* area; const uint32_t* nedge; int track_size;
  __device__ __forceinline__ Tree get() const {
    Tree t; t.W = W; t.H = H; t.lev = lev; t.parent = parent; t.area = area; t.nedge = nedge; t.track_size = track_size;
    return t;
  }
};
// best[v] = (area << 32 | ~rep) of the largest tracked child of v
__global__ void k_mser_best(TreeDev td, uint32_t N, u64* __restrict__ best) {
  const Tree t = td.get();
  for (uint32_t x = blockIdx.x * blockDim.x + threadIdx.x; x < N; x += gridDim.x * blockDim.x) {
    if (is_root(t, x) || !is_rep(t, x) || !tracked(t, x)) continue;
    atomicMax(best + t.parent[x], ((u64)t.area[x] << 32) | (u64)(~x));
  }
}
// flag: 1 = region born in this node, 2 = equally large tracked children (survivor decided by replay), 0 otherwise
__global__ void k_mser_ties(TreeDev td, uint32_t N, const u64* __restrict__ best, uint8_t* __restrict__ flag) {
  const Tree t = td.get();
  for (uint32_t x = blockIdx.x * blockDim.x + threadIdx.x; x < N; x += gridDim.x * blockDim.x) {
    if (is_root(t, x) || !is_rep(t, x) || !tracked(t, x)) continue;
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

// Warp-per-node form of the replay.  emulate_node (mser_logic.cuh) spends most of its time in loads that do NOT depend on the replay
// state: the level of the four neighbours of every own pixel and, for a neighbour below the node's level, the walk up to the child
// of v that contains it (child_containing) plus that child's area.  The 32 lanes do this for 32 consecutive own pixels at once; the
// order-dependent part -- the small union-find over {children, own pixels} -- is then replayed by lane 0 pixel by pixel, in raster
// order, from the shuffled results.  Same decisions as emulate_node, statement by statement.
__global__ void __launch_bounds__(256)
k_mser_emulate_warp(TreeDev td, uint32_t N, const uint32_t* __restrict__ emu_nodes, uint32_t n_nodes, const u64* __restrict__ keys, uint32_t n_keys,
                    EmuScratch s, uint32_t* __restrict__ surv, uint32_t* __restrict__ birth, MserCounters* C) {
  const int lane = threadIdx.x & 31;
  const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= n_nodes) return;
  const Tree t = td.get();
  const uint32_t v = emu_nodes[wid];
  const uint32_t lo = lower_bound_u64(keys, n_keys, (u64)v << 32), hi = lower_bound_u64(keys, n_keys, ((u64)v + 1) << 32);
  const u64* own = keys + lo;
  const int n_own = (int)(hi - lo);
  const int L = t.lev[v], W = t.W;
  EmuResult res; res.survivor = NONE; res.birth = NONE; res.overflow = 0;
  for (int k0 = 0; k0 < n_own; k0 += 32) {
    // ---- order-independent part, one own pixel per lane
    const int k = k0 + lane;
    uint32_t p = 0, id[4] = {0, 0, 0, 0}, ar[4] = {0, 0, 0, 0};
    unsigned codes = 0;   // 2 bits per direction (up, left, right, down): 0 none, 1 child of v (id = its rep, ar = its area), 2 same-level pixel
    if (k < n_own) {
      p = (uint32_t)own[k];
      const int x = (int)(p % (uint32_t)W), y = (int)((p / (uint32_t)W) % (uint32_t)t.H);
#pragma unroll
      for (int d = 0; d < 4; d++) {
        uint32_t q;
        if (d == 0) { if (y == 0) continue; q = p - W; }
        else if (d == 1) { if (x == 0) continue; q = p - 1; }
        else if (d == 2) { if (x == W - 1) continue; q = p + 1; }
        else { if (y == t.H - 1) continue; q = p + W; }
        const int lq = t.lev[q];
        if (lq < L) { const uint32_t c = child_containing(t, v, q); id[d] = c; ar[d] = t.area[c]; codes |= 1u << (2 * d); }
        else if (lq == L && q < p) { id[d] = q; codes |= 2u << (2 * d); }
      }
    }
    // ---- order-dependent part: lane 0 replays the chunk in raster order
    const int cnt = min(32, n_own - k0);
    for (int i = 0; i < cnt; i++) {
      const uint32_t pi = __shfl_sync(0xffffffffu, p, i);
      const unsigned ci = __shfl_sync(0xffffffffu, codes, i);
      uint32_t idi[4], ari[4];
#pragma unroll
      for (int d = 0; d < 4; d++) { idi[d] = __shfl_sync(0xffffffffu, id[d], i); ari[d] = __shfl_sync(0xffffffffu, ar[d], i); }
      if (lane != 0) continue;
      uint32_t lab[4]; int n = 0;
#pragma unroll
      for (int d = 0; d < 4; d++) {
        const unsigned code = (ci >> (2 * d)) & 3u;
        if (code == 0) continue;
        uint32_t e;
        if (code == 1) {
          const uint32_t c = idi[d];
          e = s.N + c;
          if (s.uf[e] == NONE) {  // first contact with this child
            s.uf[e] = e; s.size[e] = ari[d];
            if ((int)ari[d] >= t.track_size) { s.kind[e] = 1; s.presize[e] = ari[d]; s.ident[e] = c; s.birth[e] = NONE; }
            else s.kind[e] = 0;
          }
        } else e = idi[d];
        const uint32_t r = emu_find(s, e);
        bool dup = false;
        for (int j = 0; j < n; j++) dup |= (lab[j] == r);
        if (!dup) lab[n++] = r;
      }
      if (n == 0) { s.uf[pi] = pi; s.size[pi] = 1; s.kind[pi] = 0; continue; }       // ConsRegion :174-178
      if (n == 1) { emu_add_pixel(t, s, lab[0], pi, res); continue; }
      // MergeRegions :267-361
      uint32_t maxSize = 0, maxLabel = lab[0];
      for (int j = 0; j < n; j++)
        if (s.kind[lab[j]] == 1 && s.presize[lab[j]] > maxSize) { maxSize = s.presize[lab[j]]; maxLabel = lab[j]; }
      for (int j = 0; j < n; j++) {
        const uint32_t l = lab[j];
        if (l == maxLabel) continue;
        s.size[maxLabel] += s.size[l];
        s.uf[l] = maxLabel;
      }
      if (s.kind[maxLabel] == 0 && s.size[maxLabel] >= 32768u) res.overflow = 1;
      emu_add_pixel(t, s, maxLabel, pi, res);
    }
    __syncwarp();
  }
  if (lane == 0) {
    const uint32_t r = emu_find(s, v);
    if (s.kind[r] == 1) { res.survivor = s.ident[r]; res.birth = s.birth[r]; }
    if (res.overflow) atomicAdd(&C->overflow, 1u);
    surv[v] = res.survivor;
    birth[v] = res.birth;
  }
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
                                 const uint32_t* __restrict__ birth, uint32_t Nimg, double min_margin, int min_size, int max_size, uint32_t* slot_of_node,
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
    e.key = ((u64)(r.v0 / Nimg) << 48) | ((u64)minI << 40) | ((u64)birth[r.v0] << 8) | (u64)k;
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
// sa[x] = slot of the nearest selected node on the path from x's node to the root (the node included).  Selected components are
// at most max_size pixels large and areas only grow towards the root, so every pixel simply walks up until it meets a selected
// node or a component that is too large to have selected ancestors: background pixels stop at once, pixels inside a blob after a few
// levels.  (A level-synchronous top-down pass costs a grid barrier per level instead.)
__global__ void k_mser_sa(TreeDev td, uint32_t N, const uint32_t* __restrict__ slot_of_node, uint32_t max_size, uint32_t* __restrict__ sa) {
  const Tree t = td.get();
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < N; p += gridDim.x * blockDim.x) {
    uint32_t u = node_of(t, p), s = NONE;
    for (;;) {
      if (t.area[u] > max_size) break;
      s = slot_of_node[u];
      if (s != NONE) break;
      const uint32_t up = t.parent[u];
      if (up == u) break;
      u = up;
    }
    sa[p] = s;
  }
}
__global__ void k_mser_upsel(const uint32_t* __restrict__ parent, const uint32_t* __restrict__ node_of_slot, uint32_t n_slots,
                             const uint32_t* __restrict__ sa, uint32_t* __restrict__ up_sel) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  const uint32_t node = node_of_slot[s];
  up_sel[s] = (parent[node] == node) ? NONE : sa[parent[node]];
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
  const uint32_t N2 = C->n_stack * (uint32_t)W * H;
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < N2; p += gridDim.x * blockDim.x) {
    uint32_t s = sa[p];
    if (s == NONE) continue;
    const int yy = p / W, x = p - yy * W, y = yy % H;
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
                            const uint32_t* __restrict__ slot_of_node, uint32_t Nimg, int as_regions, KeyOut* __restrict__ out, double* __restrict__ table,
                            uint32_t* __restrict__ n_img) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const SelRec e = sel[order[i]];
  const int pol = (int)((e.node / Nimg) & 1u);
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
  atomicAdd(n_img + e.node / (2 * Nimg), 1u);   // regions per image (keys are image-major)
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

// K same-size images, both polarities each, at once (stacked, see the header comment).  Writes the keys image by image
// (MSER+ first, reference order inside each) to d_out, n_per_img[k] of them for image k; table (optional) receives 13
// doubles per region.
int mser_stack(mb2_ctx* ctx, const ImgView* imgs, int K, const mb2_mser_params& par, double min_margin, int as_regions, KeyOut* d_out,
               int out_cap, int* n_per_img, double* d_table) {
  MserBufs& B = *mser_bufs(ctx);
  cudaStream_t st = ctx->stream;
  const int W = imgs[0].cols, H = imgs[0].rows;
  const uint32_t Nimg = (uint32_t)W * H, N = 2 * Nimg * (uint32_t)K;
  MserImages mi; mi.n = K;
  for (int k = 0; k < K; k++) { mi.p[k] = imgs[k].p; mi.pitch[k] = imgs[k].pitch; n_per_img[k] = 0; }
  const int G = ctx->num_sms * 8;
  const int track_size = std::min(10000, par.min_size);
  const int max_size = (int)((double)W * (double)H * par.max_area);
  const uint32_t list_cap = N;  // every list below holds at most one entry per pixel, except runs (checked on the device)

  MB2_CUDA_CHECK(ctx, B.counters.reserve(sizeof(MserCounters)));
  MB2_CUDA_CHECK(ctx, B.h_counters.reserve(sizeof(MserCounters)));
  MserCounters* dC = B.counters.as<MserCounters>();
  MserCounters* hc = B.h_counters.as<MserCounters>();
  const bool wide = Nimg > (1u << 24);          // 64-bit tree words above 2^24 pixels per image
  MB2_CUDA_CHECK(ctx, B.lev.reserve(N));
  MB2_CUDA_CHECK(ctx, B.gpar.reserve((size_t)N * (wide ? 8 : 4)));
  DevBuf* u32bufs[] = {&B.parent, &B.area, &B.nedge, &B.lsum, &B.surv, &B.birth, &B.slot_of_node, &B.sa};
  for (DevBuf* b : u32bufs) MB2_CUDA_CHECK(ctx, b->reserve((size_t)N * 4));
  MB2_CUDA_CHECK(ctx, B.hi.reserve((size_t)N * 2));
  MB2_CUDA_CHECK(ctx, B.best.reserve((size_t)N * 8));
  MB2_CUDA_CHECK(ctx, B.flag.reserve(N));
  uint8_t* lev = B.lev.as<uint8_t>();
  uint32_t *parent = B.parent.as<uint32_t>(), *area = B.area.as<uint32_t>(), *nedge = B.nedge.as<uint32_t>(), *lsum = B.lsum.as<uint32_t>();
  uint16_t* hi = B.hi.as<uint16_t>();
  uint32_t *surv = B.surv.as<uint32_t>(), *birth = B.birth.as<uint32_t>(), *slot_of_node = B.slot_of_node.as<uint32_t>(), *sa = B.sa.as<uint32_t>();

  // 1 + 2. component trees: tiles in shared memory, local sums, tile borders in global memory, final parents and sums
  MB2_CUDA_CHECK(ctx, cudaMemsetAsync(dC, 0, sizeof(MserCounters), st));
  {
    const dim3 tg((W + MT_TILE - 1) / MT_TILE, (H + MT_TILE - 1) / MT_TILE, K), lg(tg.x, tg.y, 2 * K);
    const bool borders_one = getenv("MB2_BORDERS_TWO") == nullptr;   // all border pairs in ONE launch (0.73 ms per image; one middle pair of every border first, then the rest: 0.76 -- MB2_BORDERS_TWO=1)
    const unsigned long long n_border = ((unsigned long long)((H - 1) / MT_TILE) * W + (unsigned long long)((W - 1) / MT_TILE) * H) * 2 * K;
    const int bg = (int)std::max<unsigned long long>(1, std::min<unsigned long long>((n_border + 255) / 256, (unsigned long long)ctx->num_sms * 64));
    if (wide) {
      unsigned long long* gp = B.gpar.as<unsigned long long>();
      MB2_LAUNCH(ctx, k_mtree_tiles<GKey64>, tg, MT_TILES_THREADS, 0, mi, W, H, lev, gp, dC);
      MB2_LAUNCH(ctx, k_mtree_local<GKey64>, lg, MT_THREADS, 0, W, H, lev, gp, lsum, hi, area, nedge);
      if (n_border && borders_one) MB2_LAUNCH(ctx, k_mtree_borders<GKey64>, bg, 256, 0, W, H, 2 * K, lev, gp, 2);
      else if (n_border) { MB2_LAUNCH(ctx, k_mtree_borders<GKey64>, bg, 256, 0, W, H, 2 * K, lev, gp, 0); MB2_LAUNCH(ctx, k_mtree_borders<GKey64>, bg, 256, 0, W, H, 2 * K, lev, gp, 1); }
      MB2_LAUNCH(ctx, k_mtree_fix<GKey64>, (N + 255) / 256, 256, 0, W, H, N, lev, gp, lsum, hi, parent, sa, dC);
      MB2_LAUNCH(ctx, k_mtree_walk<GKey64>, ctx->num_sms * 16, 128, 0, W, H, lev, gp, lsum, hi, sa, dC, area, nedge);
    } else {
      uint32_t* gp = B.gpar.as<uint32_t>();
      MB2_LAUNCH(ctx, k_mtree_tiles<GKey32>, tg, MT_TILES_THREADS, 0, mi, W, H, lev, gp, dC);
      MB2_LAUNCH(ctx, k_mtree_local<GKey32>, lg, MT_THREADS, 0, W, H, lev, gp, lsum, hi, area, nedge);
      if (n_border && borders_one) MB2_LAUNCH(ctx, k_mtree_borders<GKey32>, bg, 256, 0, W, H, 2 * K, lev, gp, 2);
      else if (n_border) { MB2_LAUNCH(ctx, k_mtree_borders<GKey32>, bg, 256, 0, W, H, 2 * K, lev, gp, 0); MB2_LAUNCH(ctx, k_mtree_borders<GKey32>, bg, 256, 0, W, H, 2 * K, lev, gp, 1); }
      MB2_LAUNCH(ctx, k_mtree_fix<GKey32>, (N + 255) / 256, 256, 0, W, H, N, lev, gp, lsum, hi, parent, sa, dC);
      MB2_LAUNCH(ctx, k_mtree_walk<GKey32>, ctx->num_sms * 16, 128, 0, W, H, lev, gp, lsum, hi, sa, dC, area, nedge);
    }
    if (ctx->ev_tree) { cudaEventRecord(ctx->ev_tree, st); __sync_synchronize(); ctx->tree_epoch = ctx->tree_epoch + 1; }
  }
  // 3. survivors
  TreeDev td{W, H, lev, parent, area, nedge, track_size};
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
    static const bool emu_warp = getenv("MB2_MSER_EMU_THREAD") == nullptr;   // A/B: the thread-per-node form
    if (emu_warp) MB2_LAUNCH(ctx, k_mser_emulate_warp, (n_emu + 7) / 8, 256, 0, td, N, B.emu_nodes.as<uint32_t>(), n_emu, own_sorted, n_own, es, surv, birth, dC);
    else MB2_LAUNCH(ctx, k_mser_emulate, (n_emu + 63) / 64, 64, 0, td, N, B.emu_nodes.as<uint32_t>(), n_emu, own_sorted, n_own, es, surv, birth, dC);
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
                 birth, Nimg, min_margin, par.min_size, max_size, slot_of_node, B.node_of_slot.as<uint32_t>(), B.sel.as<SelRec>(), sel_cap, dC);
      MSER_SYNC_COUNTERS();
      if (hc->thr_overflow) { ctx->set_error("mser: more than 128 stability thresholds in one region"); return MB2_ERR_CAPACITY; }
      n_sel = (int)hc->n_sel; n_slots = n_sel;
    }
  }
  if (n_sel == 0) return MB2_OK;
  if (n_sel > out_cap) { ctx->set_error("mser: output capacity too small"); return MB2_ERR_CAPACITY; }
  // reference list order (polarity, birth level, promotion time, threshold rank)
  MB2_CUDA_CHECK(ctx, B.selkey_a.reserve((size_t)n_sel * 8)); MB2_CUDA_CHECK(ctx, B.selkey_b.reserve((size_t)n_sel * 8));
  MB2_CUDA_CHECK(ctx, B.selidx_a.reserve((size_t)n_sel * 4)); MB2_CUDA_CHECK(ctx, B.selidx_b.reserve((size_t)n_sel * 4));
  MB2_LAUNCH(ctx, k_mser_selkeys, (n_sel + 255) / 256, 256, 0, B.sel.as<SelRec>(), (uint32_t)n_sel, B.selkey_a.as<u64>(), B.selidx_a.as<uint32_t>());
  {
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, B.selkey_a.as<u64>(), B.selkey_b.as<u64>(), B.selidx_a.as<uint32_t>(), B.selidx_b.as<uint32_t>(), n_sel, 0, 52, st);
    MB2_CUDA_CHECK(ctx, B.cub_tmp.reserve(tmp));
    cub::DeviceRadixSort::SortPairs(B.cub_tmp.p, tmp, B.selkey_a.as<u64>(), B.selkey_b.as<u64>(), B.selidx_a.as<uint32_t>(), B.selidx_b.as<uint32_t>(), n_sel, 0, 52, st);
    ctx->launches += 4;
  }
  // 5. runs of the selected components
  MB2_CUDA_CHECK(ctx, B.up_sel.reserve((size_t)n_slots * 4));
  MB2_LAUNCH(ctx, k_mser_sa, grid_for(N, 256, G), 256, 0, td, N, slot_of_node, (uint32_t)max_size, sa);
  MB2_LAUNCH(ctx, k_mser_upsel, (n_slots + 255) / 256, 256, 0, parent, B.node_of_slot.as<uint32_t>(), (uint32_t)n_slots, sa, B.up_sel.as<uint32_t>());
  const uint32_t run_cap = N + 1024;
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
  MB2_LAUNCH(ctx, k_mser_keys, (n_sel + 127) / 128, 128, 0, B.sel.as<SelRec>(), B.selidx_b.as<uint32_t>(), (uint32_t)n_sel, B.mom.as<SlotMoments>(), slot_of_node,
             Nimg, as_regions, d_out, d_table, &dC->n_img[0]);
  MSER_SYNC_COUNTERS();
  for (int k = 0; k < K; k++) n_per_img[k] = (int)hc->n_img[k];
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
  if (img.cols > 65535 || img.rows > 65535 || (long long)img.cols * img.rows >= (1LL << 30)) { ctx->set_error("mser: image too large"); return MB2_ERR_ARG; }
  const double min_margin = par.mode != 0 ? 1.0 : par.min_margin;  // extrema.cpp:294-299
  const int cap = capacity > 0 ? capacity : (int)std::min<long long>((long long)img.cols * img.rows / 16 + 4096, 4000000);
  MB2_CUDA_CHECK(ctx, ctx->kp_b.reserve((size_t)cap * sizeof(KeyOut)));
  int rc;
  const bool sort_on_host = par.mode != 0;
  if ((rc = mser_stack(ctx, &img, 1, par, min_margin, sort_on_host ? 0 : as_regions, ctx->kp_b.as<KeyOut>(), cap, n_out, d_table_out))) return rc;
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

// Two same-size images in one pass (FIXED_TH): the level-synchronous tree kernel is latency bound, so a pair costs
// hardly more than one image.  Keys of image 0 then image 1 in ctx->kp_b.
int mb2_mser_core_pair(mb2_ctx* ctx, const ImgView& img1, const ImgView& img2, const mb2_mser_params& par, int as_regions, int* n1, int* n2) {
  using namespace MB2_NS;
  *n1 = *n2 = 0;
  if (par.min_size < 2 || par.relative || par.mode != 0) { ctx->set_error("mser pair: FIXED_TH, absolute margins, min_size >= 2 only"); return MB2_ERR_UNSUPPORTED; }
  if (img1.cols != img2.cols || img1.rows != img2.rows) { ctx->set_error("mser pair: images must have the same size"); return MB2_ERR_ARG; }
  if (img1.cols > 65535 || img1.rows > 65535 || (long long)img1.cols * img1.rows >= (1LL << 29)) { ctx->set_error("mser: image too large"); return MB2_ERR_ARG; }
  const int cap = (int)std::min<long long>((long long)img1.cols * img1.rows / 8 + 8192, 8000000);
  MB2_CUDA_CHECK(ctx, ctx->kp_b.reserve((size_t)cap * sizeof(KeyOut)));
  const ImgView imgs[2] = {img1, img2};
  int n[2] = {0, 0};
  const int rc = mser_stack(ctx, imgs, 2, par, par.min_margin, as_regions, ctx->kp_b.as<KeyOut>(), cap, n, nullptr);
  *n1 = n[0]; *n2 = n[1];
  return rc;
}
