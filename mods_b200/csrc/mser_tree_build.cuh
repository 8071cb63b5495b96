// Component tree of the level sets {I <= t} by LOCK-FREE MERGING (replaces the level-synchronous k_mser_tree).
//
// Every pixel x carries one word par[x] = key of another pixel, key(p) = (level(p) << IDX_BITS) | index(p) -- a strict total order that
// refines the level order.  Invariants, true at every instant:
//   * par[x] == key(x) (x is the root of its current tree) or par[x] > key(x): pointers only go up in the total order (no cycles);
//   * par[x] is a pixel of x's own node (same level, same component: x is then NOT a level root, and never becomes one again), or a pixel
//     of a strict ancestor node of x's node in the FINAL tree (x is a level root: the representative of its node so far).
// connect(a, b) is called once for every 4-adjacent pixel pair, in ANY order and concurrently: it carries the obligation "a and b lie in
// one component at level max(level a, level b)" up the two root paths like the merge of two sorted lists -- Wilkinson's sequential
// branch merge made atomic: one compare-and-swap per link, a failed swap re-reads and continues.  A successful swap of par[x] from z to y
// leaves the obligation (y, z), so whatever was reachable from x stays reachable.  When all calls have returned, following same-level
// pointers from x ends at the representative of x's node, and the representative's pointer ends in its parent node: the canonical
// tree mser_logic.cuh expects (which pixel of its level names a node is irrelevant to every output).
// Path halving inside a level keeps flat zones shallow; nothing is sorted and there is no barrier between levels.
//
// The same code runs on a 64 x 64 tile in shared memory (all edges inside the tile), on the tile borders in global memory, and --
// compiled for the host by tests/native/mser_tree_cpu.cpp -- sequentially and on several host threads in the logic tests.
// `Mem` supplies: word load(uint32_t idx); bool cas(uint32_t idx, word expect, word desired); void store(uint32_t idx, word w).
#pragma once
#include <stdint.h>

#ifndef MB2_HD
#ifdef __CUDACC__
#define MB2_HD __host__ __device__ __forceinline__
#else
#define MB2_HD inline
#endif
#endif

namespace mser_tree {

// 32-bit words hold images up to 2^24 pixels (level in the top byte); larger images use 64-bit words (KeyT<unsigned long long, 32>)
template <class WordT, int IDX_BITS>
struct KeyT {
  typedef WordT word;
  static MB2_HD WordT make(int lev, uint32_t idx) { return ((WordT)(uint32_t)lev << IDX_BITS) | (WordT)idx; }
  static MB2_HD int lev(WordT k) { return (int)(k >> IDX_BITS); }
  static MB2_HD uint32_t idx(WordT k) { return (uint32_t)(k & (((WordT)1 << IDX_BITS) - (WordT)1)); }
};
template <int IDX_BITS> using Key = KeyT<uint32_t, IDX_BITS>;

// representative (level root) of x's node at this moment; xk = key of x
template <class K, class Mem>
MB2_HD typename K::word levroot(Mem& m, typename K::word xk) {
  for (;;) {
    const typename K::word w = m.load(K::idx(xk));
    if (w == xk || K::lev(w) != K::lev(xk)) return xk;
    const typename K::word ww = m.load(K::idx(w));
    if (ww == w || K::lev(ww) != K::lev(xk)) return w;   // w is a level root
    m.store(K::idx(xk), ww);                              // path halving: xk is not a level root, so nobody swaps its word
    xk = ww;
  }
}

template <class K, class Mem>
MB2_HD void connect(Mem& m, typename K::word ak, typename K::word bk) {
  typename K::word x = levroot<K>(m, ak), y = levroot<K>(m, bk);
  for (;;) {
    if (x == y) return;
    if (x > y) { const typename K::word t = x; x = y; y = t; }
    const typename K::word z = m.load(K::idx(x));
    const bool self = z == x;
    if (!self && K::lev(z) == K::lev(x)) { x = levroot<K>(m, x); continue; }    // x stopped being a level root
    if (K::lev(x) == K::lev(y)) {                // one node: x goes under y, x's parent becomes y's obligation
      if (m.cas(K::idx(x), z, y)) { if (self) return; x = y; y = levroot<K>(m, z); }
      else x = levroot<K>(m, x);
      continue;
    }
    if (self) { if (m.cas(K::idx(x), z, y)) return; x = levroot<K>(m, x); continue; }
    if (K::lev(z) <= K::lev(y)) { x = levroot<K>(m, z); continue; }              // the parent is still below (or at) y's level: move up
    if (m.cas(K::idx(x), z, y)) { x = y; y = levroot<K>(m, z); }                 // y between x and z
    else x = levroot<K>(m, x);
  }
}

// after all connects: canonical parent of x as mser_logic.cuh defines it -- the representative of x's node if x is not one, else the
// representative of the parent node (the root: itself).  Returns a KEY.
template <class K, class Mem>
MB2_HD typename K::word canonical_parent(Mem& m, typename K::word xk) {
  typename K::word r = xk;
  for (;;) {
    const typename K::word w = m.load(K::idx(r));
    if (w == r) return r == xk ? xk : r;
    if (K::lev(w) != K::lev(r)) {
      if (r != xk) return r;
      r = w;                                     // xk is the representative: continue inside the parent node
      for (;;) {
        const typename K::word v = m.load(K::idx(r));
        if (v == r || K::lev(v) != K::lev(r)) return r;
        r = v;
      }
    }
    r = w;
  }
}

}  // namespace mser_tree
