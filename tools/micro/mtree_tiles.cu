// Variants of the per-tile component-tree kernel (k_mtree_tiles, mods_b200/csrc/mser.cu) timed side by side on a real image, each checked
// against the shipped formulation node by node (canonical form: smallest pixel of a node, smallest pixel of its parent node).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I mods_b200/csrc -o tools/micro/mtree_tiles.bin tools/micro/mtree_tiles.cu
//   tools/micro/mtree_tiles.bin image.u8 W H        (u8 image as extrema.cpp:401-403 makes it: (unsigned char) of the float)
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "mser_tree_build.cuh"

typedef mser_tree::KeyT<uint32_t, 12> TKey;
#define MT_TILE 64

// SWZ: word i lives at i ^ ((i >> 8) & 3): rows of different 4-row bands start in different banks, so the 32 lanes of a warp that own 32
// different 4 x 4 blocks (8 across x 4 down) hit 32 different banks when they step through their blocks in lock step
template <bool SWZ> __device__ __forceinline__ uint32_t phys(uint32_t i) { return SWZ ? i ^ ((i >> 8) & 3u) : i; }
template <bool SWZ>
struct SmemWords {
  uint32_t* w;
  __device__ __forceinline__ uint32_t load(uint32_t i) { return *(volatile uint32_t*)(w + phys<SWZ>(i)); }
  __device__ __forceinline__ void store(uint32_t i, uint32_t v) { *(volatile uint32_t*)(w + phys<SWZ>(i)) = v; }
  __device__ __forceinline__ bool cas(uint32_t i, uint32_t expect, uint32_t desired) { return atomicCAS(w + phys<SWZ>(i), expect, desired) == expect; }
};
// one thread owns the words it touches (sequential phases): plain loads / stores, no atomics
template <bool SWZ>
struct SmemWordsSeq {
  uint32_t* w;
  __device__ __forceinline__ uint32_t load(uint32_t i) { return w[phys<SWZ>(i)]; }
  __device__ __forceinline__ void store(uint32_t i, uint32_t v) { w[phys<SWZ>(i)] = v; }
  __device__ __forceinline__ bool cas(uint32_t i, uint32_t, uint32_t desired) { w[phys<SWZ>(i)] = desired; return true; }
};

// Stages (s = 0 .. 11: b = 1 << (s >> 1); even s joins b x b blocks into 2b x b, odd s joins 2b x b blocks into 2b x 2b):
//   s < 2 SEQ        ONE thread builds a whole (1 << SEQ)^2 block sequentially,
//   s < PAIR_UPTO    one thread per PAIR of blocks joins them: all border edges of the pair in turn, plain loads / stores,
//   else             one thread per border edge, compare-and-swap (the shipped kernel does this for all 12 stages).
template <int THREADS, int SEQ, int PAIR_UPTO, bool SWZ, bool PROF, int TWO_FROM = 12, bool KMAJOR = false>
__global__ void __launch_bounds__(THREADS) k_tiles(const uint8_t* __restrict__ img, int W, int H, uint32_t* __restrict__ out, unsigned long long* prof) {
  __shared__ uint32_t par[2][MT_TILE * MT_TILE];
  __shared__ uint8_t sv[MT_TILE * MT_TILE];
  const int x0 = blockIdx.x * MT_TILE, y0 = blockIdx.y * MT_TILE;
  const int tw = min(MT_TILE, W - x0), th = min(MT_TILE, H - y0);
  const uint32_t Nimg = (uint32_t)W * H;
  for (int i = threadIdx.x; i < MT_TILE * MT_TILE; i += THREADS) {
    const int lx = i & (MT_TILE - 1), ly = i >> 6;
    int v = 0;
    if (lx < tw && ly < th) v = img[(size_t)(y0 + ly) * W + x0 + lx];
    par[0][phys<SWZ>(i)] = TKey::make(v, i); par[1][phys<SWZ>(i)] = TKey::make(255 - v, i); sv[i] = (uint8_t)v;
  }
  __syncthreads();
  long long t_prev = 0;
  if (PROF && threadIdx.x == 0) t_prev = clock64();
  if (SEQ > 0) {
    const int BS = 1 << SEQ, nb = MT_TILE / BS, ntask = 2 * nb * nb;
    for (int t = threadIdx.x; t < ntask; t += THREADS) {
      int pol, bx, by;
      if (SWZ && SEQ == 2) {   // 8 blocks across x 4 down per warp; polarity in the top bit of the task index
        pol = t >> 8; const int u = t & 255;
        bx = ((u & 7) | (((u >> 5) & 1) << 3)) * BS; by = (((u >> 3) & 3) | ((u >> 6) << 2)) * BS;
      } else { pol = t & 1; const int u = t >> 1; bx = (u % nb) * BS; by = (u / nb) * BS; }
      SmemWordsSeq<SWZ> m{par[pol]};
      for (int s = 0; s < 2 * SEQ; s++) {
        const bool horiz = !(s & 1);
        const int b = 1 << (s >> 1);
        for (int line = 0; line < BS; line++)
          for (int c = b - 1; c < BS - 1; c += 2 * b) {
            const int lx = bx + (horiz ? c : line), ly = by + (horiz ? line : c);
            if ((horiz ? lx + 1 : lx) >= tw || (horiz ? ly : ly + 1) >= th) continue;
            const int i = ly * MT_TILE + lx, j = horiz ? i + 1 : i + MT_TILE;
            const int vi = pol ? 255 - sv[i] : sv[i], vj = pol ? 255 - sv[j] : sv[j];
            mser_tree::connect<TKey>(m, TKey::make(vi, i), TKey::make(vj, j));
          }
      }
    }
    __syncthreads();
    if (PROF && threadIdx.x == 0) { const long long t = clock64(); atomicAdd(&prof[12], (unsigned long long)(t - t_prev)); t_prev = t; }
  }
  for (int s = 2 * SEQ; s < 12; s++) {
    const bool horiz = !(s & 1);
    const int b = 1 << (s >> 1), per_line = (MT_TILE / 2) / b;
    if (s < PAIR_UPTO) {
      const int n_edges = horiz ? b : 2 * b, groups = MT_TILE / n_edges, ntask = 2 * per_line * groups;   // pairs of blocks x polarities
      for (int t = threadIdx.x; t < ntask; t += THREADS) {
        const int pol = t & 1, u = t >> 1, k = u % per_line, grp = u / per_line, c = (2 * k + 1) * b - 1;
        SmemWordsSeq<SWZ> m{par[pol]};
        for (int e = 0; e < n_edges; e++) {
          const int line = grp * n_edges + ((e + n_edges / 2) % n_edges);   // the middle of the border first
          const int lx = horiz ? c : line, ly = horiz ? line : c;
          if ((horiz ? lx + 1 : lx) >= tw || (horiz ? ly : ly + 1) >= th) continue;
          const int i = ly * MT_TILE + lx, j = horiz ? i + 1 : i + MT_TILE;
          const int vi = pol ? 255 - sv[i] : sv[i], vj = pol ? 255 - sv[j] : sv[j];
          mser_tree::connect<TKey>(m, TKey::make(vi, i), TKey::make(vj, j));
        }
      }
    } else {
      // TWO_FROM: from that stage on, ONE edge in the middle of every block border goes first (it zips the two trees while nobody else
      // swaps the same words), the other edges follow after a barrier and mostly find their paths merged
      const int ntask = 2 * MT_TILE * per_line, n_edges = horiz ? b : 2 * b;
      for (int ph = (s >= TWO_FROM ? 0 : 1); ph < 2; ph++) {
        for (int t = threadIdx.x; t < ntask; t += THREADS) {
          const int pol = t & 1, u = t >> 1;
          int line, k;
          if (KMAJOR) { k = u / MT_TILE; line = u - k * MT_TILE; } else { line = u / per_line; k = u - line * per_line; }
          const int c = (2 * k + 1) * b - 1;
          if (s >= TWO_FROM && ((line & (n_edges - 1)) == n_edges / 2) != (ph == 0)) continue;
          const int lx = horiz ? c : line, ly = horiz ? line : c;
          if ((horiz ? lx + 1 : lx) >= tw || (horiz ? ly : ly + 1) >= th) continue;
          const int i = ly * MT_TILE + lx, j = horiz ? i + 1 : i + MT_TILE;
          const int vi = pol ? 255 - sv[i] : sv[i], vj = pol ? 255 - sv[j] : sv[j];
          SmemWords<SWZ> m{par[pol]};
          mser_tree::connect<TKey>(m, TKey::make(vi, i), TKey::make(vj, j));
        }
        if (ph == 0) __syncthreads();
      }
    }
    __syncthreads();
    if (PROF && threadIdx.x == 0) { const long long t = clock64(); atomicAdd(&prof[s], (unsigned long long)(t - t_prev)); t_prev = t; }
  }
  for (int i = threadIdx.x; i < MT_TILE * MT_TILE; i += THREADS) {
    const int lx = i & (MT_TILE - 1), ly = i >> 6;
    if (lx >= tw || ly >= th) continue;
    const uint32_t pi = (uint32_t)(y0 + ly) * W + x0 + lx;
    out[pi] = par[0][phys<SWZ>(i)]; out[Nimg + pi] = par[1][phys<SWZ>(i)];
  }
}

// ---- 8 x 8 blocks built by ONE thread each with the sequential union-find algorithm (pixels in key order, Berger et al.) ------------
// instead of pair-by-pair merging: a counting sort of the block's 64 pixels (two 4-bit passes, byte counters packed in registers), then
// every pixel in turn adopts the current tops of its already processed neighbours' components.  The words come out with the same
// invariants connect() keeps (pointers go up in key order; same-level chains end at the node's representative, whose word names a pixel
// of the parent node), so the cooperative stages 6-11 continue from them.
#define BT 128   // tasks per tile: 64 blocks x 2 polarities
__device__ __forceinline__ uint32_t sc_addr(int t, int i) { return (uint32_t)((((i >> 2) * BT + t) << 2) | (i & 3)); }   // byte i of task t, word-interleaved over the tasks
template <int THREADS, bool PROF>
__global__ void __launch_bounds__(THREADS) k_tiles_seq8(const uint8_t* __restrict__ img, int W, int H, uint32_t* __restrict__ out, unsigned long long* prof) {
  extern __shared__ uint32_t dyn[];
  uint32_t (*par)[MT_TILE * MT_TILE] = reinterpret_cast<uint32_t (*)[MT_TILE * MT_TILE]>(dyn);
  uint8_t* sv = reinterpret_cast<uint8_t*>(dyn + 2 * MT_TILE * MT_TILE);
  uint8_t* ordb = sv + MT_TILE * MT_TILE;          // [64 x BT] sorted pixel lists
  uint8_t* zpb = ordb + 64 * BT;                    // [64 x BT] union-find forest (first: scratch of the sort)
  const int x0 = blockIdx.x * MT_TILE, y0 = blockIdx.y * MT_TILE;
  const int tw = min(MT_TILE, W - x0), th = min(MT_TILE, H - y0);
  const uint32_t Nimg = (uint32_t)W * H;
  for (int i = threadIdx.x; i < MT_TILE * MT_TILE; i += THREADS) {
    const int lx = i & (MT_TILE - 1), ly = i >> 6;
    int v = 0;
    if (lx < tw && ly < th) v = img[(size_t)(y0 + ly) * W + x0 + lx];
    par[0][i] = TKey::make(v, i); par[1][i] = TKey::make(255 - v, i); sv[i] = (uint8_t)v;
  }
  __syncthreads();
  long long t_prev = 0;
  if (PROF && threadIdx.x == 0) t_prev = clock64();
  for (int t = threadIdx.x; t < BT; t += THREADS) {
    const int pol = t & 1, u = t >> 1, bx = (u & 7) * 8, by = (u >> 3) * 8, base = by * MT_TILE + bx;
    const uint32_t flip = pol ? 255u : 0u;
    auto level = [&](int j) -> uint32_t { return (uint32_t)sv[base + (j >> 3) * MT_TILE + (j & 7)] ^ flip; };   // 255 - v == v ^ 255
    // counting sort, stable, by (level, j): pass 1 on the low nibble j -> zpb, pass 2 on the high nibble zpb -> ordb
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
      uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll 4
      for (int k = 0; k < 64; k++) {
        const int j = pass ? zpb[sc_addr(t, k)] : k;
        const uint32_t d = (level(j) >> (4 * pass)) & 15u, inc = 1u << ((d & 3u) * 8u), m = d >> 2;
        c0 += m == 0 ? inc : 0u; c1 += m == 1 ? inc : 0u; c2 += m == 2 ? inc : 0u; c3 += m == 3 ? inc : 0u;
      }
      const uint32_t p0 = c0 * 0x01010101u, p1 = c1 * 0x01010101u, p2 = c2 * 0x01010101u, p3 = c3 * 0x01010101u;
      const uint32_t t0 = p0 >> 24, t1 = t0 + (p1 >> 24), t2 = t1 + (p2 >> 24);
      c0 = p0 - c0; c1 = p1 - c1 + t0 * 0x01010101u; c2 = p2 - c2 + t1 * 0x01010101u; c3 = p3 - c3 + t2 * 0x01010101u;   // exclusive offsets per digit
#pragma unroll 4
      for (int k = 0; k < 64; k++) {
        const int j = pass ? zpb[sc_addr(t, k)] : k;
        const uint32_t d = (level(j) >> (4 * pass)) & 15u, sh = (d & 3u) * 8u, inc = 1u << sh, m = d >> 2;
        const uint32_t cm = m == 0 ? c0 : m == 1 ? c1 : m == 2 ? c2 : c3;
        const uint32_t o = (cm >> sh) & 0xffu;
        (pass ? ordb : zpb)[sc_addr(t, (int)o)] = (uint8_t)j;
        c0 += m == 0 ? inc : 0u; c1 += m == 1 ? inc : 0u; c2 += m == 2 ? inc : 0u; c3 += m == 3 ? inc : 0u;
      }
    }
    // union-find in key order
    uint32_t* pw = par[pol];
#pragma unroll 1
    for (int k = 0; k < 64; k++) {
      const int p = ordb[sc_addr(t, k)];
      const int px = p & 7, py = p >> 3, ip = base + py * MT_TILE + px;
      const uint32_t lp = level(p), kp = TKey::make((int)lp, (uint32_t)ip);
      zpb[sc_addr(t, p)] = (uint8_t)p;
      if (bx + px >= tw || by + py >= th) continue;          // outside the image: stays alone
#pragma unroll
      for (int nb = 0; nb < 4; nb++) {
        const int qx = px + (nb == 0 ? -1 : nb == 1 ? 1 : 0), qy = py + (nb == 2 ? -1 : nb == 3 ? 1 : 0);
        if (qx < 0 || qx > 7 || qy < 0 || qy > 7 || bx + qx >= tw || by + qy >= th) continue;
        const int q = qy * 8 + qx;
        const uint32_t lq = level(q);
        if (!(lq < lp || (lq == lp && q < p))) continue;     // not processed yet
        int r = q;
        for (;;) {                                            // find with path halving
          const int z = zpb[sc_addr(t, r)];
          if (z == r) break;
          const int zz = zpb[sc_addr(t, z)];
          zpb[sc_addr(t, r)] = (uint8_t)zz;
          r = zz;
        }
        if (r != p) { pw[base + (r >> 3) * MT_TILE + (r & 7)] = kp; zpb[sc_addr(t, r)] = (uint8_t)p; }
      }
    }
  }
  __syncthreads();
  if (PROF && threadIdx.x == 0) { const long long t = clock64(); atomicAdd(&prof[12], (unsigned long long)(t - t_prev)); t_prev = t; }
  for (int s = 6; s < 12; s++) {
    const bool horiz = !(s & 1);
    const int b = 1 << (s >> 1), per_line = (MT_TILE / 2) / b, ntask = 2 * MT_TILE * per_line;
    for (int t = threadIdx.x; t < ntask; t += THREADS) {
      const int pol = t & 1, u = t >> 1, line = u / per_line, c = (2 * (u - line * per_line) + 1) * b - 1;
      const int lx = horiz ? c : line, ly = horiz ? line : c;
      if ((horiz ? lx + 1 : lx) >= tw || (horiz ? ly : ly + 1) >= th) continue;
      const int i = ly * MT_TILE + lx, j = horiz ? i + 1 : i + MT_TILE;
      const int vi = pol ? 255 - sv[i] : sv[i], vj = pol ? 255 - sv[j] : sv[j];
      SmemWords<false> m{par[pol]};
      mser_tree::connect<TKey>(m, TKey::make(vi, i), TKey::make(vj, j));
    }
    __syncthreads();
    if (PROF && threadIdx.x == 0) { const long long t = clock64(); atomicAdd(&prof[s], (unsigned long long)(t - t_prev)); t_prev = t; }
  }
  for (int i = threadIdx.x; i < MT_TILE * MT_TILE; i += THREADS) {
    const int lx = i & (MT_TILE - 1), ly = i >> 6;
    if (lx >= tw || ly >= th) continue;
    const uint32_t pi = (uint32_t)(y0 + ly) * W + x0 + lx;
    out[pi] = par[0][i]; out[Nimg + pi] = par[1][i];
  }
}
#define SEQ8_SMEM (2 * MT_TILE * MT_TILE * 4 + MT_TILE * MT_TILE + 2 * 64 * BT)
template <int T, bool P> static void launch8(dim3 g, const uint8_t* img, int W, int H, uint32_t* out, unsigned long long* prof) {
  cudaFuncSetAttribute(k_tiles_seq8<T, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, SEQ8_SMEM);
  k_tiles_seq8<T, P><<<g, T, SEQ8_SMEM>>>(img, W, H, out, prof);
}

// ---- host: canonical form of one tile's tree ---------------------------------------------------------------------------------------
static void canon(const std::vector<uint32_t>& out, const std::vector<uint8_t>& img, int W, int H, int tx, int ty, int pol, std::vector<int>& A, std::vector<int>& B) {
  const int x0 = tx * MT_TILE, y0 = ty * MT_TILE, tw = std::min(MT_TILE, W - x0), th = std::min(MT_TILE, H - y0);
  const size_t N = (size_t)W * H;
  std::vector<uint32_t> w(4096, 0); std::vector<int> L(4096, -1), rep(4096, -1), mn(4096, 1 << 30);
  for (int ly = 0; ly < th; ly++) for (int lx = 0; lx < tw; lx++) {
    const int i = ly * 64 + lx; const size_t pi = (size_t)(y0 + ly) * W + x0 + lx;
    w[i] = out[pol * N + pi]; L[i] = pol ? 255 - img[pi] : img[pi];
  }
  auto find_rep = [&](int i) { int r = i; for (int g = 0; g < 5000; g++) { const uint32_t v = w[r]; if ((int)TKey::idx(v) == r || TKey::lev(v) != L[r]) return r; r = TKey::idx(v); } return -2; };
  for (int ly = 0; ly < th; ly++) for (int lx = 0; lx < tw; lx++) { const int i = ly * 64 + lx; rep[i] = find_rep(i); if (rep[i] >= 0) mn[rep[i]] = std::min(mn[rep[i]], i); }
  A.assign(4096, -1); B.assign(4096, -1);
  for (int ly = 0; ly < th; ly++) for (int lx = 0; lx < tw; lx++) {
    const int i = ly * 64 + lx, r = rep[i];
    if (r < 0) { A[i] = -2; continue; }
    A[i] = mn[r];
    const uint32_t v = w[r];
    if ((int)TKey::idx(v) == r) B[i] = -1;
    else { const int pr = find_rep(TKey::idx(v)); B[i] = pr < 0 ? -2 : mn[pr] | (L[pr] << 16); }
  }
}

typedef void (*Launch)(dim3, const uint8_t*, int, int, uint32_t*, unsigned long long*);
template <int T, int S, int PU, bool Z, bool P, int TF = 12, bool KM = false> static void launch(dim3 g, const uint8_t* img, int W, int H, uint32_t* out, unsigned long long* prof) { k_tiles<T, S, PU, Z, P, TF, KM><<<g, T>>>(img, W, H, out, prof); }

int main(int argc, char** argv) {
  if (argc < 4) { printf("usage: %s image.u8 W H\n", argv[0]); return 2; }
  const int W = atoi(argv[2]), H = atoi(argv[3]);
  const size_t N = (size_t)W * H;
  std::vector<uint8_t> img(N);
  FILE* f = fopen(argv[1], "rb");
  if (!f || fread(img.data(), 1, N, f) != N) { printf("cannot read %s\n", argv[1]); return 2; }
  fclose(f);
  uint8_t* d_img; uint32_t* d_out; unsigned long long* d_prof;
  cudaMalloc(&d_img, N); cudaMalloc(&d_out, 2 * N * 4); cudaMalloc(&d_prof, 16 * 8);
  cudaMemcpy(d_img, img.data(), N, cudaMemcpyHostToDevice);
  dim3 g((W + 63) / 64, (H + 63) / 64);
  // every formulation DESIGN.md section 7 quotes a time for (ms per 4096 x 3072 image, both polarities, one B200)
  struct V { const char* name; Launch fn; } vs[] = {
      {"round-2a kernel: 256 threads, 12 cooperative stages", launch<256, 0, 0, false, false>},       // 1.87
      {"12 cooperative stages, 128 threads", launch<128, 0, 0, false, false>},                       // 2.70
      {"12 cooperative stages, 512 threads", launch<512, 0, 0, false, false>},                       // 1.78
      {"seq 2x2 per thread + 10 stages, 256", launch<256, 1, 0, false, false>},                      // 1.72
      {"seq 4x4 per thread + 8 stages, 256", launch<256, 2, 0, false, false>},                       // 1.48
      {"seq 4x4 per thread + 8 stages, 512 (SHIPPED)", launch<512, 2, 0, false, false>},             // 1.33
      {"seq 4x4 per thread + 8 stages, 1024", launch<1024, 2, 0, false, false>},                     // 1.85
      {"seq 8x8 per thread + 6 stages, 256", launch<256, 3, 0, false, false>},                       // 1.49
      {"seq 4x4 bank-swizzled, 512", launch<512, 2, 0, true, false>},                                // 1.42
      {"seq 4x4 + block pairs by one thread up to stage 6, 512", launch<512, 2, 6, false, false>},   // 1.36
      {"seq 4x4 + block pairs by one thread up to stage 8, 512", launch<512, 2, 8, false, false>},   // 2.00
      {"seq 4x4 + block pairs by one thread up to stage 12, 512", launch<512, 2, 12, false, false>}, // 3.79
      {"block pairs by one thread from stage 0 up to 8, 512", launch<512, 0, 8, false, false>},      // 2.26
      {"seq 4x4, 512, middle pair of a border first from stage 10", launch<512, 2, 0, false, false, 10>},   // 1.57
      {"seq 4x4, 512, middle pair of a border first from stage 4", launch<512, 2, 0, false, false, 4>},     // 2.23
      {"seq 4x4, 512, k-major task order", launch<512, 2, 0, false, false, 12, true>},               // 1.37
      {"seq 8x8 by sorted union-find + 6 stages, 512", launch8<512, false>},                         // 1.56
      {"seq 8x8 by sorted union-find + 6 stages, 256", launch8<256, false>},                         // 1.76
  };
  std::vector<uint32_t> ref, out(2 * N);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (const V& v : vs) {
    v.fn(g, d_img, W, H, d_out, d_prof);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) v.fn(g, d_img, W, H, d_out, d_prof);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const cudaError_t err = cudaGetLastError();
    cudaMemcpy(out.data(), d_out, 2 * N * 4, cudaMemcpyDeviceToHost);
    int bad = 0, checked = 0;
    if (ref.empty()) ref = out;
    else {
      std::vector<int> A0, B0, A1, B1;
      for (int t = 0; t < (int)(g.x * g.y); t += 7) {   // every 7th tile, both polarities
        for (int pol = 0; pol < 2; pol++) {
          canon(ref, img, W, H, t % g.x, t / g.x, pol, A0, B0); canon(out, img, W, H, t % g.x, t / g.x, pol, A1, B1);
          checked++; if (A0 != A1 || B0 != B1) bad++;
        }
      }
    }
    printf("%-48s %.3f ms per image   %s   trees checked %d, different %d\n", v.name, ms / 5, err ? cudaGetErrorString(err) : "ok", checked, bad);
  }
  // per-stage share of the shipped kernel and of the 4x4 variant (cycles of thread 0 between barriers, summed over the CTAs)
  for (int which = 0; which < 2; which++) {
    cudaMemset(d_prof, 0, 16 * 8);
    if (which == 0) launch<512, 2, 0, false, true>(g, d_img, W, H, d_out, d_prof); else launch8<512, true>(g, d_img, W, H, d_out, d_prof);
    unsigned long long p[16]; cudaMemcpy(p, d_prof, sizeof p, cudaMemcpyDeviceToHost);
    double tot = 0; for (int s = 0; s < 13; s++) tot += (double)p[s];
    printf("%s stage cycles per CTA:", which ? "seq 8x8 sorted union-find (512)" : "seq 4x4 (512)");
    printf("  [seq] %.0f (%.0f%%)", (double)p[12] / (g.x * g.y), 100.0 * p[12] / tot);
    for (int s = 0; s < 12; s++) printf("  [%d] %.0f (%.0f%%)", s, (double)p[s] / (g.x * g.y), 100.0 * p[s] / tot);
    printf("\n");
  }
  return 0;
}
