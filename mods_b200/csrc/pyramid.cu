// Scale-space pyramid for the Hessian-Affine detector: separable Gaussian blur fused with the
// Hessian response, 2x2 decimation, 3x3x3 non-maximum suppression, sub-pixel localisation and the
// per-octave "first come" de-duplication.
//
// Reference: detectors/affinedetectors/pyramid.cpp (detectPyramidKeypoints :540-573,
// detectOctaveKeypoints :455-538, HessianResponse :223-281, findLevelKeypoints :432-452,
// localizeKeypoint :308-430), detectors/helpers.cpp (gaussianBlur :717-731, solveLinear3x3
// :309-368).  The OpenCV 2.4.9 GaussianBlur / resize arithmetic is the definition restated in
// oracle/cvmath.h (row pass left-to-right, column pass symmetric pairs, no FMA).
//
// All planes are f32 with a row pitch that is a multiple of 32 floats (128 B).
#include "common.cuh"
#undef MB2_NS
#define MB2_NS mb2_pyramid_detail
#include "pyramid.cuh"
#include <cuda.h>

// ---------------------------------------------------------------------------------------------
// Blur (+ optional Hessian response of the blurred image) -- one pass over HBM per level:
// reads the source level once (halo re-reads are L2 hits), writes blur and response.
// Tile = 64 x 32 output pixels per CTA of 256 threads; halo = taps/2 (+1 for the 3x3 Hessian).
// ---------------------------------------------------------------------------------------------
namespace MB2_NS {

constexpr int TW = 64, TH = 32, BLUR_THREADS = 256;
constexpr int RS = 6;    // row-pass outputs per work item  (TW + 2 = 66 = 11 * 6)
constexpr int CS = 12;   // column-pass outputs per work item (TH + 2 = 34 -> 12 + 12 + 10)

// The kernel is instruction-bound unless the taps are applied from registers: every work item of the
// row pass loads a window of RS + NT - 1 neighbours once and produces RS outputs from it; the column
// pass does the same down a column.  Arithmetic order per output is unchanged (row: left-to-right
// accumulation; column: centre tap, then symmetric pairs), so results are bit-identical to the oracle.
template <int NT>  // number of taps (odd)
__global__ void __launch_bounds__(BLUR_THREADS)
k_blur_hess(ImgView src, float* __restrict__ dst_blur, float* __restrict__ dst_resp, int dst_pitch, BlurTaps taps,
            float norm2, int want_resp) {
  constexpr int H = NT / 2;
  constexpr int IN_W = TW + 2 * H + 2, IN_H = TH + 2 * H + 2;   // source tile incl. halo (IN_W is even)
  constexpr int IN_WP = IN_W + 2;                               // padded pitch (even, so float2 loads stay aligned)
  constexpr int RP_W = TW + 2, RP_WP = RP_W + 2;                // row-pass result: IN_H x RP_W (pitch 68: conflict-free columns)
  constexpr int CP_W = TW + 2, CP_H = TH + 2, CP_WP = RP_WP;    // col-pass result, aliases the source tile
  extern __shared__ float smem[];
  float* s_in = smem;                        // IN_H * IN_WP
  float* s_rp = s_in + IN_H * IN_WP;         // IN_H * RP_WP
  float* s_cp = s_in;                        // CP_H * CP_WP (the source tile is dead after the row pass)
  static_assert(CP_H * CP_WP <= IN_H * IN_WP, "col-pass tile must fit in the source tile");

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  float k[NT];
#pragma unroll
  for (int j = 0; j < NT; j++) k[j] = taps.k[j];
  // 1. load with clamped (replicated) coordinates, one warp per tile row
  for (int ly = warp; ly < IN_H; ly += BLUR_THREADS / 32) {
    const int gy = min(max(y0 + ly - H - 1, 0), src.rows - 1);
    const float* row = src.p + (size_t)gy * src.pitch;
    for (int lx = lane; lx < IN_W; lx += 32) {
      const int gx = min(max(x0 + lx - H - 1, 0), src.cols - 1);
      s_in[ly * IN_WP + lx] = row[gx];
    }
  }
  __syncthreads();
  // 2. row pass: item = (row ly, strip of RS outputs); output lx is centred on source column lx + H
  for (int item = tid; item < IN_H * (RP_W / RS); item += BLUR_THREADS) {
    const int ly = item / (RP_W / RS), st = item - ly * (RP_W / RS);
    const float* p = s_in + ly * IN_WP + st * RS;
    float w[RS + NT - 1];
#pragma unroll
    for (int j = 0; j < RS + NT - 1; j += 2) {
      const float2 v = *reinterpret_cast<const float2*>(p + j);
      w[j] = v.x;
      if (j + 1 < RS + NT - 1) w[j + 1] = v.y;
    }
    float* o = s_rp + ly * RP_WP + st * RS;
#pragma unroll
    for (int q = 0; q < RS; q++) {
      float acc = fmul(k[0], w[q]);
#pragma unroll
      for (int j = 1; j < NT; j++) acc = fadd(acc, fmul(k[j], w[q + j]));
      o[q] = acc;
    }
  }
  __syncthreads();
  // 3. column pass (symmetric pairs): item = (column lx, strip of CS rows); output ly is centred on row ly + H
  constexpr int NSTRIP = (CP_H + CS - 1) / CS;
  for (int item = tid; item < CP_W * NSTRIP; item += BLUR_THREADS) {
    const int st = item / CP_W, lx = item - st * CP_W;
    const int ly0 = st * CS;
    const float* p = s_rp + ly0 * RP_WP + lx;
    float w[CS + 2 * H];
#pragma unroll
    for (int j = 0; j < CS + 2 * H; j++) w[j] = (ly0 + j < IN_H) ? p[j * RP_WP] : 0.f;
#pragma unroll
    for (int q = 0; q < CS; q++) {
      if (ly0 + q < CP_H) {
        float acc = fmul(k[H], w[q + H]);
#pragma unroll
        for (int j = 1; j <= H; j++) acc = fadd(acc, fmul(k[H + j], fadd(w[q + H + j], w[q + H - j])));
        s_cp[(ly0 + q) * CP_WP + lx] = acc;
      }
    }
  }
  __syncthreads();
  // 4. write blur + Hessian response (pyramid.cpp:223-281; 1-px frame of the response is left 0)
  for (int ly = warp; ly < TH; ly += BLUR_THREADS / 32) {
    const int gy = y0 + ly;
    if (gy >= src.rows) break;
#pragma unroll
    for (int half = 0; half < TW / 32; half++) {
      const int lx = lane + 32 * half, gx = x0 + lx;
      if (gx >= src.cols) continue;
      const float* c = s_cp + (ly + 1) * CP_WP + (lx + 1);
      dst_blur[(size_t)gy * dst_pitch + gx] = c[0];
      if (want_resp) {
        float r = 0.f;
        if (gy >= 1 && gy < src.rows - 1 && gx >= 1 && gx < src.cols - 1) {
          const float v11 = c[-CP_WP - 1], v12 = c[-CP_WP], v13 = c[-CP_WP + 1];
          const float v21 = c[-1], v22 = c[0], v23 = c[1];
          const float v31 = c[CP_WP - 1], v32 = c[CP_WP], v33 = c[CP_WP + 1];
          const float Lxx = fadd(fsub(v21, fmul(2.f, v22)), v23);
          const float Lyy = fadd(fsub(v12, fmul(2.f, v22)), v32);
          const float Lxy = fmul(fsub(fadd(fsub(v13, v11), v31), v33), 0.25f) /* == x / 4.0f exactly: a power-of-two scale */;
          r = fmul(fsub(fmul(Lxx, Lyy), fmul(Lxy, Lxy)), norm2);
        }
        dst_resp[(size_t)gy * dst_pitch + gx] = r;
      }
    }
  }
}

__device__ __forceinline__ bool is_max9(const ImgView& im, float val, int r, int c);
__device__ __forceinline__ bool is_min9(const ImgView& im, float val, int r, int c);

// ---------------------------------------------------------------------------------------------
// TMA-staged form of the same pass (the one the library runs for DET_HESSIAN): the source tile with its halo arrives in shared memory
// through ONE cp.async.bulk.tensor.2d per CTA (a 2-D tensor map over the source plane; out-of-image elements come back as zeros and
// are then overwritten with the replicated border -- BORDER_REPLICATE -- from the tile itself, on border tiles only), so the load costs
// no address arithmetic in the SM.  The tile carries one more pixel of halo than k_blur_hess: the response is formed on (TW + 2) x
// (TH + 2) pixels, which lets the kernel that PRODUCES a detection level also run the in-level part of the 3x3x3 extremum test
// (pyramid.cpp:42-64, 432-452) from shared memory.  Only the pixels that pass it (|val| beyond the gate and a non-strict 3x3 extremum of
// their own level, ~1 % of the pixels) are listed; k_nms_finish checks the levels below and above for those.  The response planes are
// no longer re-read in full by a separate pass.
// Arithmetic per output is the one of k_blur_hess (row pass left to right, column pass centre tap then symmetric pairs, no FMA).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pyr_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NT>
struct TmaBlurGeom {
  static constexpr int H = NT / 2, HALO = H + 2;
  // The inner start coordinate of a tensor-map box must be a multiple of 16 bytes (measured: tools/micro/tma2d.cu), so the tile starts
  // HX >= HALO columns left of the output tile, HX a multiple of 4; the row pass starts XS - E columns into the tile (an even offset: its
  // windows are read as float2) and its output column lx stands for image column x0 - 2 - E + lx.
  static constexpr int HX = (HALO + 3) / 4 * 4, XS = HX - HALO, E = XS & 1;
  static constexpr int RP_W = 72, RP_WP = 74;                       // row-pass result: 72 columns (12 strips of 6; 68 + E are needed), pitch 74
  static constexpr int IN_H = TH + 2 * HALO;
  static constexpr int BOXW = ((XS - E + RP_W + 2 * H) + 3) / 4 * 4; // TMA box width: whole 16-byte units
  static constexpr int CP_W = TW + 4, CP_H = TH + 4, CP_WP = RP_WP; // column-pass result (aliases the source tile)
  static constexpr int RS_W = TW + 2, RS_H = TH + 2, RS_WP = TW + 4;// response incl. a one-pixel halo (aliases the row-pass result)
  static constexpr size_t SMEM = 128 /*align*/ + sizeof(float) * (size_t)(IN_H * BOXW + IN_H * RP_WP) + 16 /*mbarrier*/;
  static_assert(CP_H * CP_WP <= IN_H * BOXW, "col-pass tile must fit in the source tile");
  static_assert(RS_H * RS_WP <= IN_H * RP_WP, "response tile must fit in the row-pass tile");
  static_assert(BOXW <= 256 && IN_H <= 256, "TMA box limits");
};

template <int NT>
__global__ void __launch_bounds__(BLUR_THREADS)
k_blur_hess_tma(const __grid_constant__ CUtensorMap tmap, int rows, int cols, float* __restrict__ dst_blur, float* __restrict__ dst_resp, int dst_pitch,
                BlurTaps taps, float norm2, int prefilter, int border, float posThr, float negThr, int level, Candidate* __restrict__ pre,
                int* __restrict__ pre_count, int pre_cap) {
  typedef TmaBlurGeom<NT> G;
  constexpr int H = G::H, HALO = G::HALO, IN_H = G::IN_H, BOXW = G::BOXW, RP_W = G::RP_W, RP_WP = G::RP_WP;
  constexpr int CP_W = G::CP_W, CP_H = G::CP_H, CP_WP = G::CP_WP, RS_W = G::RS_W, RS_H = G::RS_H, RS_WP = G::RS_WP;
  extern __shared__ unsigned char smem_raw_[];
  float* s_in = reinterpret_cast<float*>(((uintptr_t)smem_raw_ + 127) & ~(uintptr_t)127);   // IN_H x BOXW, written by the TMA unit
  float* s_rp = s_in + IN_H * BOXW;                                                          // IN_H x RP_WP
  float* s_cp = s_in;                                                                        // CP_H x CP_WP
  float* s_rs = s_rp;                                                                        // RS_H x RS_WP
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_rp + IN_H * RP_WP);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const int gx0 = x0 - G::HX, gy0 = y0 - HALO;           // image coordinates of the tile's first element (gx0 is a multiple of 4)
  float k[NT];
#pragma unroll
  for (int j = 0; j < NT; j++) k[j] = taps.k[j];
  // 1. one bulk tensor copy brings the whole tile
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pyr_smem_u32(bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pyr_smem_u32(bar)), "r"((uint32_t)(IN_H * BOXW * sizeof(float))) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(pyr_smem_u32(s_in)),
                 "l"(&tmap), "r"(pyr_smem_u32(bar)), "r"(gx0), "r"(gy0)
                 : "memory");
  }
  __syncthreads();
  {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(pyr_smem_u32(bar)), "r"(0u) : "memory");
  }
  // BORDER_REPLICATE on tiles that reach over the image: every out-of-image element takes the value of the nearest image pixel, which
  // lies inside the tile (never itself overwritten)
  if (gx0 < 0 || gy0 < 0 || gx0 + BOXW > cols || gy0 + IN_H > rows) {
    for (int i = tid; i < IN_H * BOXW; i += BLUR_THREADS) {
      const int ly = i / BOXW, lx = i - ly * BOXW;
      const int gy = gy0 + ly, gx = gx0 + lx;
      if (gy < 0 || gy >= rows || gx < 0 || gx >= cols) {
        const int cy = min(max(gy, 0), rows - 1) - gy0, cx = min(max(gx, 0), cols - 1) - gx0;
        s_in[i] = s_in[cy * BOXW + cx];
      }
    }
    __syncthreads();
  }
  // 2. row pass: item = (row ly, strip of RS outputs); output column lx is centred on source column lx + H
  for (int item = tid; item < IN_H * (RP_W / RS); item += BLUR_THREADS) {
    const int ly = item / (RP_W / RS), st = item - ly * (RP_W / RS);
    const float* p = s_in + ly * BOXW + (G::XS - G::E) + st * RS;
    float w[RS + NT - 1];
#pragma unroll
    for (int j = 0; j < RS + NT - 1; j += 2) {
      const float2 v = *reinterpret_cast<const float2*>(p + j);
      w[j] = v.x;
      if (j + 1 < RS + NT - 1) w[j + 1] = v.y;
    }
    float* o = s_rp + ly * RP_WP + st * RS;
#pragma unroll
    for (int q = 0; q < RS; q++) {
      float acc = fmul(k[0], w[q]);
#pragma unroll
      for (int j = 1; j < NT; j++) acc = fadd(acc, fmul(k[j], w[q + j]));
      o[q] = acc;
    }
  }
  __syncthreads();
  // 3. column pass (symmetric pairs): item = (column lx, strip of CS rows); output row ly is centred on row ly + H
  constexpr int NSTRIP = CP_H / CS;
  static_assert(CP_H % CS == 0, "column strips");
  for (int item = tid; item < CP_W * NSTRIP; item += BLUR_THREADS) {
    const int st = item / CP_W, lx = item - st * CP_W;
    const int ly0 = st * CS;
    const float* p = s_rp + ly0 * RP_WP + lx + G::E;   // s_cp column lx <-> image column x0 - 2 + lx
    float w[CS + 2 * H];
#pragma unroll
    for (int j = 0; j < CS + 2 * H; j++) w[j] = p[j * RP_WP];
#pragma unroll
    for (int q = 0; q < CS; q++) {
      float acc = fmul(k[H], w[q + H]);
#pragma unroll
      for (int j = 1; j <= H; j++) acc = fadd(acc, fmul(k[H + j], fadd(w[q + H + j], w[q + H - j])));
      s_cp[(ly0 + q) * CP_WP + lx] = acc;
    }
  }
  __syncthreads();
  // 4. Hessian response (pyramid.cpp:223-281; 0 on the one-pixel frame of the image) on the tile and a one-pixel ring around it
  for (int i = tid; i < RS_H * RS_W; i += BLUR_THREADS) {
    const int ly = i / RS_W, lx = i - ly * RS_W;         // response (ly, lx) <-> image (y0 - 1 + ly, x0 - 1 + lx) <-> s_cp (ly + 1, lx + 1)
    const int gy = y0 - 1 + ly, gx = x0 - 1 + lx;
    float r = 0.f;
    if (gy >= 1 && gy < rows - 1 && gx >= 1 && gx < cols - 1) {
      const float* c = s_cp + (ly + 1) * CP_WP + (lx + 1);
      const float v11 = c[-CP_WP - 1], v12 = c[-CP_WP], v13 = c[-CP_WP + 1];
      const float v21 = c[-1], v22 = c[0], v23 = c[1];
      const float v31 = c[CP_WP - 1], v32 = c[CP_WP], v33 = c[CP_WP + 1];
      const float Lxx = fadd(fsub(v21, fmul(2.f, v22)), v23);
      const float Lyy = fadd(fsub(v12, fmul(2.f, v22)), v32);
      const float Lxy = fmul(fsub(fadd(fsub(v13, v11), v31), v33), 0.25f) /* == x / 4.0f exactly: a power-of-two scale */;
      r = fmul(fsub(fmul(Lxx, Lyy), fmul(Lxy, Lxy)), norm2);
    }
    s_rs[ly * RS_WP + lx] = r;
  }
  // blur goes out while the response tile settles (s_cp is read-only from here on)
  for (int ly = warp; ly < TH; ly += BLUR_THREADS / 32) {
    const int gy = y0 + ly;
    if (gy >= rows) break;
#pragma unroll
    for (int half = 0; half < TW / 32; half++) {
      const int lx = lane + 32 * half, gx = x0 + lx;
      if (gx < cols) dst_blur[(size_t)gy * dst_pitch + gx] = s_cp[(ly + 2) * CP_WP + lx + 2];
    }
  }
  __syncthreads();
  // 5. response out + in-level extremum test
  for (int ly = warp; ly < TH; ly += BLUR_THREADS / 32) {
    const int gy = y0 + ly;
    if (gy >= rows) break;
#pragma unroll
    for (int half = 0; half < TW / 32; half++) {
      const int lx = lane + 32 * half, gx = x0 + lx;
      bool hit = false;
      if (gx < cols) {
        const float* c = s_rs + (ly + 1) * RS_WP + lx + 1;
        const float val = c[0];
        dst_resp[(size_t)gy * dst_pitch + gx] = val;
        if (prefilter && gy >= border && gy < rows - border && gx >= border && gx < cols - border) {
          if (val > posThr) {
            hit = true;
#pragma unroll
            for (int dr = -1; dr <= 1; dr++)
#pragma unroll
              for (int dc = -1; dc <= 1; dc++) hit = hit && !(c[dr * RS_WP + dc] > val);
          } else if (val < negThr) {
            hit = true;
#pragma unroll
            for (int dr = -1; dr <= 1; dr++)
#pragma unroll
              for (int dc = -1; dc <= 1; dc++) hit = hit && !(c[dr * RS_WP + dc] < val);
          }
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (m) {
        int base = 0;
        if (lane == 0) base = atomicAdd(pre_count, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0) + __popc(m & ((1u << lane) - 1));
        if (hit && base < pre_cap) pre[base] = Candidate{gy, gx, level, 0};
      }
    }
  }
}

// The levels below and above, for the pixels that passed the in-level test (pyramid.cpp:42-64: non-strict, ties pass)
__global__ void __launch_bounds__(128) k_nms_finish(OctaveLevels oct, const Candidate* __restrict__ pre, const int* __restrict__ pre_count, int pre_cap,
                                                    Candidate* __restrict__ out, int* __restrict__ count, int capacity) {
  const int n = min(*pre_count, pre_cap), lane = threadIdx.x & 31;
  for (int base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; base < n; base += gridDim.x * blockDim.x) {
    const int i = base + lane;
    bool hit = false;
    Candidate cd{0, 0, 0, 0};
    if (i < n) {
      cd = pre[i];
      const ImgView low = oct.resp[cd.level - 1], cur = oct.resp[cd.level], high = oct.resp[cd.level + 1];
      const float val = cur.at(cd.r, cd.c);
      hit = val > 0 ? (is_max9(low, val, cd.r, cd.c) && is_max9(high, val, cd.r, cd.c)) : (is_min9(low, val, cd.r, cd.c) && is_min9(high, val, cd.r, cd.c));
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (m) {
      int b0 = 0;
      if (lane == 0) b0 = atomicAdd(count, __popc(m));
      b0 = __shfl_sync(0xffffffffu, b0, 0) + __popc(m & ((1u << lane) - 1));
      if (hit && b0 < capacity) out[b0] = cd;
    }
  }
}

// Hessian response of an existing level (first level of every octave after the 0th).
__global__ void k_hessian(ImgView src, float* __restrict__ dst, int dst_pitch, float norm2) {
  int gx = blockIdx.x * blockDim.x + threadIdx.x, gy = blockIdx.y * blockDim.y + threadIdx.y;
  if (gx >= src.cols || gy >= src.rows) return;
  float r = 0.f;
  if (gy >= 1 && gy < src.rows - 1 && gx >= 1 && gx < src.cols - 1) {
    const float v11 = src.at(gy - 1, gx - 1), v12 = src.at(gy - 1, gx), v13 = src.at(gy - 1, gx + 1);
    const float v21 = src.at(gy, gx - 1), v22 = src.at(gy, gx), v23 = src.at(gy, gx + 1);
    const float v31 = src.at(gy + 1, gx - 1), v32 = src.at(gy + 1, gx), v33 = src.at(gy + 1, gx + 1);
    const float Lxx = fadd(fsub(v21, fmul(2.f, v22)), v23);
    const float Lyy = fadd(fsub(v12, fmul(2.f, v22)), v32);
    const float Lxy = fmul(fsub(fadd(fsub(v13, v11), v31), v33), 0.25f) /* == x / 4.0f exactly: a power-of-two scale */;
    r = fmul(fsub(fmul(Lxx, Lyy), fmul(Lxy, Lxy)), norm2);
  }
  dst[(size_t)gy * dst_pitch + gx] = r;
}

// cv::resize(x0.5, INTER_LINEAR) == fast INTER_AREA 2x2 mean (oracle/cvmath.h resize_half).
__global__ void k_resize_half(ImgView src, float* __restrict__ dst, int orows, int ocols, int dst_pitch) {
  int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y * blockDim.y + threadIdx.y;
  if (c >= ocols || r >= orows) return;
  int r0 = 2 * r, c0 = 2 * c;
  float v;
  if (r0 + 1 < src.rows && c0 + 1 < src.cols) {
    float sum = 0.f;
    sum = fadd(sum, src.at(r0, c0)); sum = fadd(sum, src.at(r0, c0 + 1));
    sum = fadd(sum, src.at(r0 + 1, c0)); sum = fadd(sum, src.at(r0 + 1, c0 + 1));
    v = fmul(sum, 0.25f);
  } else {
    float sum = 0.f; int count = 0;
    for (int sy = 0; sy < 2; sy++) {
      if (r0 + sy >= src.rows) break;
      for (int sx = 0; sx < 2; sx++) {
        if (c0 + sx >= src.cols) break;
        sum = fadd(sum, src.at(r0 + sy, c0 + sx)); count++;
      }
    }
    v = fdiv(sum, (float)count);
  }
  dst[(size_t)r * dst_pitch + c] = v;
}

// ---------------------------------------------------------------------------------------------
// 3x3x3 extrema (pyramid.cpp:42-64, 432-452).  Non-strict: ties pass.  Candidates are appended
// unordered; `key` = level * rows*cols + r*cols + c restores the reference's processing order.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool is_max9(const ImgView& im, float val, int r, int c) {
#pragma unroll
  for (int dr = -1; dr <= 1; dr++)
#pragma unroll
    for (int dc = -1; dc <= 1; dc++)
      if (im.at(r + dr, c + dc) > val) return false;
  return true;
}
__device__ __forceinline__ bool is_min9(const ImgView& im, float val, int r, int c) {
#pragma unroll
  for (int dr = -1; dr <= 1; dr++)
#pragma unroll
    for (int dc = -1; dc <= 1; dc++)
      if (im.at(r + dr, c + dc) < val) return false;
  return true;
}

// Four pixels per thread from one aligned 16-byte load of the centre plane (the scan of `cur` is the
// whole HBM cost: only the ~0.5% of pixels beyond the threshold ever touch their 26 neighbours).
__global__ void __launch_bounds__(256)
k_nms(ImgView low, ImgView cur, ImgView high, int border, float posThr, float negThr, int level,
      Candidate* __restrict__ out, int* __restrict__ count, int capacity) {
  const int c0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x), r = blockIdx.y * blockDim.y + threadIdx.y + border;
  const int lane = (threadIdx.y * blockDim.x + threadIdx.x) & 31;
  unsigned hits = 0;
  if (r < cur.rows - border && c0 < cur.cols) {
    const float4 v4 = *reinterpret_cast<const float4*>(cur.p + (size_t)r * cur.pitch + c0);  // pitch and c0 are multiples of 4
    const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int c = c0 + q;
      if (c < border || c >= cur.cols - border) continue;
      const float val = v[q];
      bool hit = false;
      if (val > posThr) hit = is_max9(cur, val, r, c) && is_max9(low, val, r, c) && is_max9(high, val, r, c);
      else if (val < negThr) hit = is_min9(cur, val, r, c) && is_min9(low, val, r, c) && is_min9(high, val, r, c);
      hits |= (unsigned)hit << q;
    }
  }
  // warp-aggregated append
  const int mine = __popc(hits);
  int prefix = mine;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) { int t = __shfl_up_sync(0xffffffffu, prefix, off); if (lane >= off) prefix += t; }
  const int total = __shfl_sync(0xffffffffu, prefix, 31);
  if (total) {
    int base = 0;
    if (lane == 31) base = atomicAdd(count, total);
    base = __shfl_sync(0xffffffffu, base, 31) + prefix - mine;
#pragma unroll
    for (int q = 0; q < 4; q++)
      if (hits & (1u << q)) { if (base < capacity) out[base] = Candidate{r, c0 + q, level, 0}; base++; }
  }
}

// helpers.cpp:309-368
__device__ __forceinline__ void swapf(float& a, float& b) { float t = a; a = b; b = t; }
__device__ void solveLinear3x3_dev(float* A, float* b) {
  int i = 0, pr = 0;
  float vp = fabsf(A[0]);
  float tmp = fabsf(A[3]);
  if (tmp > vp) { pr = 3; i = 1; vp = tmp; }
  if (fabsf(A[6]) > vp) { pr = 6; i = 2; }
  if (pr != 0) { swapf(A[pr], A[0]); swapf(A[pr + 1], A[1]); swapf(A[pr + 2], A[2]); swapf(b[i], b[0]); }
  vp = fdiv(A[3], A[0]);
  A[4] = fsub(A[4], fmul(vp, A[1])); A[5] = fsub(A[5], fmul(vp, A[2])); b[1] = fsub(b[1], fmul(vp, b[0]));
  vp = fdiv(A[6], A[0]);
  A[7] = fsub(A[7], fmul(vp, A[1])); A[8] = fsub(A[8], fmul(vp, A[2])); b[2] = fsub(b[2], fmul(vp, b[0]));
  if (fabsf(A[4]) < fabsf(A[7])) { swapf(A[7], A[4]); swapf(A[8], A[5]); swapf(b[2], b[1]); }
  vp = fdiv(A[7], A[4]);
  A[8] = fsub(A[8], fmul(vp, A[5])); b[2] = fsub(b[2], fmul(vp, b[1]));
  b[2] = fdiv(b[2], A[8]);
  b[1] = fdiv(fsub(b[1], fmul(A[5], b[2])), A[4]);
  b[0] = fdiv(fsub(fsub(b[0], fmul(A[2], b[2])), fmul(A[1], b[1])), A[0]);
}

// localizeKeypoint (pyramid.cpp:308-430) up to, not including, the octaveMap test.  Survivors
// claim their final cell with atomicMin(key): the earliest candidate in the reference's processing
// order (level, then raster) wins, exactly what the sequential octaveMap does.
__global__ void k_localize(OctaveLevels oct, Candidate* __restrict__ cand, int n, LocalizeParams lp,
                           unsigned long long* __restrict__ octmap, Localized* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Candidate cd = cand[i];
  const ImgView low = oct.resp[cd.level - 1], cur = oct.resp[cd.level], high = oct.resp[cd.level + 1];
  const int cols = cur.cols, rows = cur.rows;
  int r = cd.r, c = cd.c, nr = r, nc = c;
  float b[3] = {0.f, 0.f, 0.f};
  float val = 0.f;
  Localized L;
  L.valid = 0;
  L.key = (unsigned long long)cd.level * (unsigned long long)(rows * cols) + (unsigned long long)(cd.r * cols + cd.c);
  bool ok = true;
  for (int iter = 0; iter < 5 && ok; iter++) {
    r = nr; c = nc;
    const float c00 = cur.at(r - 1, c - 1), c01 = cur.at(r - 1, c), c02 = cur.at(r - 1, c + 1);
    const float c10 = cur.at(r, c - 1), c11 = cur.at(r, c), c12 = cur.at(r, c + 1);
    const float c20 = cur.at(r + 1, c - 1), c21 = cur.at(r + 1, c), c22 = cur.at(r + 1, c + 1);
    const float l01 = low.at(r - 1, c), l10 = low.at(r, c - 1), l11 = low.at(r, c), l12 = low.at(r, c + 1), l21 = low.at(r + 1, c);
    const float h01 = high.at(r - 1, c), h10 = high.at(r, c - 1), h11 = high.at(r, c), h12 = high.at(r, c + 1), h21 = high.at(r + 1, c);
    float dxx = fadd(fsub(c10, fmul(2.0f, c11)), c12);
    float dyy = fadd(fsub(c01, fmul(2.0f, c11)), c21);
    float dss = fadd(fsub(l11, fmul(2.0f, c11)), h11);
    float dxy = fmul(0.25f, fadd(fsub(fsub(c22, c20), c02), c00));
    if (iter == 0) {
      float s = fadd(dxx, dyy);
      float edgeScore = fdiv(fmul(s, s), fsub(fmul(dxx, dyy), fmul(dxy, dxy)));
      if ((double)edgeScore >= lp.edgeScoreThreshold || edgeScore < 0) { ok = false; break; }
    }
    float dxs = fmul(0.25f, fadd(fsub(fsub(h12, h10), l12), l10));
    float dys = fmul(0.25f, fadd(fsub(fsub(h21, h01), l21), l01));
    float A[9] = {dxx, dxy, dxs, dxy, dyy, dys, dxs, dys, dss};
    float dx = fmul(0.5f, fsub(c12, c10));
    float dy = fmul(0.5f, fsub(c21, c01));
    float ds = fmul(0.5f, fsub(h11, l11));
    b[0] = -dx; b[1] = -dy; b[2] = -ds;
    solveLinear3x3_dev(A, b);
    if (isnan(b[0]) || isnan(b[1]) || isnan(b[2])) { ok = false; break; }
    val = fadd(c11, fmul(0.5f, fadd(fadd(fmul(dx, b[0]), fmul(dy, b[1])), fmul(ds, b[2]))));
    // MAX_SUBPIXEL_SHIFT is the double literal 0.6; POINT_SAFETY_BORDER = 3
    if ((double)b[0] > 0.6) { if (c < cols - 3) nc++; else { ok = false; break; } }
    if ((double)b[1] > 0.6) { if (r < rows - 3) nr++; else { ok = false; break; } }
    if ((double)b[0] < -0.6) { if (c > 3) nc--; else { ok = false; break; } }
    if ((double)b[1] < -0.6) { if (r > 3) nr--; else { ok = false; break; } }
    if (nr == r && nc == c) break;
  }
  if (ok) {
    if (fabs((double)b[0]) > 1.5 || fabs((double)b[1]) > 1.5 || fabs((double)b[2]) > 1.5 || fabsf(val) < lp.finalThreshold) ok = false;
  }
  if (ok) {
    L.valid = 1; L.r = r; L.c = c; L.level = cd.level;
    L.b0 = b[0]; L.b1 = b[1]; L.b2 = b[2]; L.val = val;
    atomicMin(&octmap[(size_t)r * cols + c], L.key);
  }
  out[i] = L;
}

// Winners of the octaveMap race become keypoints (pyramid.cpp:420-429): scale, blob type, and the
// arguments of onKeypointDetected (x, y, s in pixels of the original view).
__global__ void k_emit_keypoints(OctaveLevels oct, const Localized* __restrict__ loc, int n,
                                 const unsigned long long* __restrict__ octmap, LocalizeParams lp, int octave,
                                 KeypointRec* __restrict__ out, int* __restrict__ count, int capacity) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Localized L = loc[i];
  if (!L.valid) return;
  const int cols = oct.resp[0].cols;
  if (octmap[(size_t)L.r * cols + L.c] != L.key) return;
  // scale = curScale * pow(2.0f, b[2] / numberOfScales) (pyramid.cpp:421) goes through the host libm:
  // glibc powf is not correctly rounded (and has FMA / non-FMA builds), so the host adapter
  // evaluates it -- see detect_core; here we only carry b2.
  int type;
  if (lp.detectorType == 1) type = L.val < 0 ? 11 : 10;   // DOG_BRIGHT : DOG_DARK
  else if (lp.detectorType == 2) type = L.val < 0 ? 31 : 30;   // HARRIS_BRIGHT : HARRIS_DARK (pyramid.h:38-39)
  else if (L.val < 0) type = 2;
  else {
    const ImgView blur = oct.blur[L.level];
    const float Lxx = fadd(fsub(blur.at(L.r, L.c - 1), fmul(2.f, blur.at(L.r, L.c))), blur.at(L.r, L.c + 1));
    type = (Lxx < 0) ? 0 : 1;
  }
  int slot = atomicAdd(count, 1);
  if (slot >= capacity) return;
  KeypointRec k;
  const float pd = lp.pixelDistance;
  k.x = fmul(pd, fadd((float)L.c, L.b0));
  k.y = fmul(pd, fadd((float)L.r, L.b1));
  k.s = 0.f;
  k.b2 = L.b2;
  k.pixelDistance = pd;
  k.response = L.val;
  k.type = type;
  k.octave = octave;
  k.level = L.level;
  k.order = ((unsigned long long)octave << 56) | L.key;
  k.a11 = 1.f; k.a12 = 0.f; k.a21 = 0.f; k.a22 = 1.f;
  k.ok = 0;
  out[slot] = k;
}

__global__ void k_scale_requests(const KeypointRec* __restrict__ kps, int n, ScaleReq* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = ScaleReq{kps[i].b2, kps[i].level, kps[i].octave};
}
__global__ void k_set_scales(KeypointRec* __restrict__ kps, int n, const float* __restrict__ s) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) kps[i].s = s[i];
}

__global__ void k_fill_u64(unsigned long long* p, size_t n, unsigned long long v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

template <int NT>
void launch_blur_nt(mb2_ctx* ctx, const ImgView& src, float* dst_blur, float* dst_resp, int dst_pitch, const BlurTaps& taps,
                    float norm2, int want_resp) {
  constexpr int H = NT / 2;
  constexpr int IN_W = TW + 2 * H + 2, IN_H = TH + 2 * H + 2;
  size_t smem = sizeof(float) * (IN_H * (IN_W + 2) + IN_H * (TW + 4));
  static unsigned long long attr_devs = 0;
  if (mb2_first_use_on_device(&attr_devs, ctx->device)) cudaFuncSetAttribute(k_blur_hess<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((src.cols + TW - 1) / TW, (src.rows + TH - 1) / TH);
  MB2_LAUNCH(ctx, k_blur_hess<NT>, grid, BLUR_THREADS, smem, src, dst_blur, dst_resp, dst_pitch, taps, norm2, want_resp);
}


typedef CUresult (*PFN_pyrEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// 2-D tensor map over one f32 plane (cols x rows, row pitch in floats: a multiple of 4), box = the kernel's tile with halo
int make_plane_tmap(mb2_ctx* ctx, CUtensorMap* m, const ImgView& src, int boxw, int boxh) {
  if (!ctx->tmap_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || !fn) { ctx->set_error("cuTensorMapEncodeTiled entry point not available"); return MB2_ERR_CUDA; }
    ctx->tmap_encode = fn;
  }
  cuuint64_t dims[2] = {(cuuint64_t)src.cols, (cuuint64_t)src.rows};
  cuuint64_t strides[1] = {(cuuint64_t)src.pitch * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)boxw, (cuuint32_t)boxh};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ((PFN_pyrEncodeTiled)ctx->tmap_encode)(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)src.p, dims, strides, box, estr,
                                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { ctx->set_error("cuTensorMapEncodeTiled (pyramid plane) failed: " + std::to_string((int)r)); return MB2_ERR_CUDA; }
  return MB2_OK;
}

template <int NT>
int launch_blur_tma_nt(mb2_ctx* ctx, const ImgView& src, float* dst_blur, float* dst_resp, int dst_pitch, const BlurTaps& taps, float norm2,
                       const PrefilterArgs& pf) {
  typedef TmaBlurGeom<NT> G;
  CUtensorMap tm;
  int rc = make_plane_tmap(ctx, &tm, src, G::BOXW, G::IN_H);
  if (rc) return rc;
  static unsigned long long attr_devs = 0;
  if (mb2_first_use_on_device(&attr_devs, ctx->device)) cudaFuncSetAttribute(k_blur_hess_tma<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
  dim3 grid((src.cols + TW - 1) / TW, (src.rows + TH - 1) / TH);
  MB2_LAUNCH(ctx, k_blur_hess_tma<NT>, grid, BLUR_THREADS, G::SMEM, tm, src.rows, src.cols, dst_blur, dst_resp, dst_pitch, taps, norm2, pf.enable, pf.border,
             pf.posThr, pf.negThr, pf.level, pf.pre, pf.pre_count, pf.pre_cap);
  return MB2_OK;
}

}  // namespace
using namespace MB2_NS;

// ---------------------------------------------------------------------------------------------
// DET_DOG response: wide separable Gaussian in OpenCV's order (generic row filter left to right, symmetric column filter centre
// then pairs, BORDER_REPLICATE; oracle/cvmath.h sep_filter), the subtraction fused into the column pass.  Tap counts reach ~100
// (sigma = curSigma^2), so the taps live in global memory and every thread walks its own window through L1/L2.
__global__ void k_wide_rows(ImgView src, const float* __restrict__ k, int n, float* __restrict__ dst) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= src.cols || y >= src.rows) return;
  const float* s = src.p + (size_t)y * src.pitch;
  const int h = n / 2, last = src.cols - 1;
  float acc = fmul(k[0], s[max(x - h, 0)]);
  for (int j = 1; j < n; j++) acc = fadd(acc, fmul(k[j], s[min(max(x - h + j, 0), last)]));
  dst[(size_t)y * src.pitch + x] = acc;
}
__global__ void k_dog_cols(ImgView level, const float* __restrict__ tmp, const float* __restrict__ k, int n, float* __restrict__ resp, int resp_pitch) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= level.cols || y >= level.rows) return;
  const int h = n / 2, last = level.rows - 1;
  float d = fmul(k[h], tmp[(size_t)y * level.pitch + x]);
  for (int j = 1; j <= h; j++)
    d = fadd(d, fmul(k[h + j], fadd(tmp[(size_t)min(y + j, last) * level.pitch + x], tmp[(size_t)max(y - j, 0) * level.pitch + x])));
  resp[(size_t)y * resp_pitch + x] = fsub(level.p[(size_t)y * level.pitch + x], d);
}

// DET_HARRIS response (pyramid.cpp:283-305): products of the level's gradient (computeGradient, helpers.cpp:779-797: central differences,
// one-sided on the image frame), each blurred with sigma = sqrt(0.6 norm) by the separable blur above, then the Harris measure -- every
// cv::Mat operation of the reference rounded to float on its own (scalar * Mat with the scalar widened to double).
__global__ void k_harris_products(ImgView im, float* __restrict__ xx, float* __restrict__ yy, float* __restrict__ xy) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y * blockDim.y + threadIdx.y;
  if (c >= im.cols || r >= im.rows) return;
  float gx, gy;
  if (c == 0) gx = fsub(im.at(r, c + 1), im.at(r, c));
  else if (c == im.cols - 1) gx = fsub(im.at(r, c), im.at(r, c - 1));
  else gx = fsub(im.at(r, c + 1), im.at(r, c - 1));
  if (r == 0) gy = fsub(im.at(r + 1, c), im.at(r, c));
  else if (r == im.rows - 1) gy = fsub(im.at(r, c), im.at(r - 1, c));
  else gy = fsub(im.at(r + 1, c), im.at(r - 1, c));
  const size_t o = (size_t)r * im.pitch + c;
  xx[o] = fmul(gx, gx); yy[o] = fmul(gy, gy); xy[o] = fmul(gx, gy);
}
__global__ void k_harris_combine(int rows, int cols, int pitch, const float* __restrict__ bxx, const float* __restrict__ byy, const float* __restrict__ bxy,
                                 float sigmasq, float* __restrict__ resp, int resp_pitch) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y * blockDim.y + threadIdx.y;
  if (c >= cols || r >= rows) return;
  const size_t o = (size_t)r * pitch + c;
  const double s = (double)sigmasq;
  const float dx2 = (float)d_mul((double)bxx[o], s), dy2 = (float)d_mul((double)byy[o], s), dxdy = (float)d_mul((double)bxy[o], s);
  const float sum = fadd(dx2, dy2);
  const float ab = fsub(fmul(dx2, dy2), fmul(dxdy, dxdy));
  resp[(size_t)r * resp_pitch + c] = fsub(ab, (float)d_mul((double)fmul(sum, sum), 0.04));
}

// host launchers
// ---------------------------------------------------------------------------------------------
int mb2_launch_harris(mb2_ctx* ctx, const ImgView& level, float* resp, int resp_pitch, const BlurTaps& taps, float sigmasq, float* d_tmp6) {
  const size_t plane = (size_t)level.pitch * level.rows;
  float *xx = d_tmp6, *yy = d_tmp6 + plane, *xy = d_tmp6 + 2 * plane, *bxx = d_tmp6 + 3 * plane, *byy = d_tmp6 + 4 * plane, *bxy = d_tmp6 + 5 * plane;
  dim3 block(32, 8), grid((level.cols + 31) / 32, (level.rows + 7) / 8);
  MB2_LAUNCH(ctx, k_harris_products, grid, block, 0, level, xx, yy, xy);
  int rc;
  const float* src[3] = {xx, yy, xy}; float* dst[3] = {bxx, byy, bxy};
  for (int i = 0; i < 3; i++)
    if ((rc = mb2_launch_blur(ctx, ImgView{src[i], level.rows, level.cols, level.pitch}, dst[i], nullptr, level.pitch, taps, 0.f, 0))) return rc;
  MB2_LAUNCH(ctx, k_harris_combine, grid, block, 0, level.rows, level.cols, level.pitch, (const float*)bxx, (const float*)byy, (const float*)bxy, sigmasq, resp,
             resp_pitch);
  return MB2_OK;
}
void mb2_launch_dog(mb2_ctx* ctx, const ImgView& level, float* resp, int resp_pitch, const float* d_taps, int n, float* d_tmp) {
  dim3 block(32, 8), grid((level.cols + 31) / 32, (level.rows + 7) / 8);
  MB2_LAUNCH(ctx, k_wide_rows, grid, block, 0, level, d_taps, n, d_tmp);
  MB2_LAUNCH(ctx, k_dog_cols, grid, block, 0, level, (const float*)d_tmp, d_taps, n, resp, resp_pitch);
}

int mb2_launch_blur(mb2_ctx* ctx, const ImgView& src, float* dst_blur, float* dst_resp, int dst_pitch, const BlurTaps& taps,
                    float norm2, int want_resp) {
  switch (taps.n) {
    case 3: launch_blur_nt<3>(ctx, src, dst_blur, dst_resp, dst_pitch, taps, norm2, want_resp); break;
    case 5: launch_blur_nt<5>(ctx, src, dst_blur, dst_resp, dst_pitch, taps, norm2, want_resp); break;
    case 7: launch_blur_nt<7>(ctx, src, dst_blur, dst_resp, dst_pitch, taps, norm2, want_resp); break;
    case 9: launch_blur_nt<9>(ctx, src, dst_blur, dst_resp, dst_pitch, taps, norm2, want_resp); break;
    case 11: launch_blur_nt<11>(ctx, src, dst_blur, dst_resp, dst_pitch, taps, norm2, want_resp); break;
    case 13: launch_blur_nt<13>(ctx, src, dst_blur, dst_resp, dst_pitch, taps, norm2, want_resp); break;
    case 15: launch_blur_nt<15>(ctx, src, dst_blur, dst_resp, dst_pitch, taps, norm2, want_resp); break;
    case 17: launch_blur_nt<17>(ctx, src, dst_blur, dst_resp, dst_pitch, taps, norm2, want_resp); break;
    case 19: launch_blur_nt<19>(ctx, src, dst_blur, dst_resp, dst_pitch, taps, norm2, want_resp); break;
    case 21: launch_blur_nt<21>(ctx, src, dst_blur, dst_resp, dst_pitch, taps, norm2, want_resp); break;
    default:
      ctx->set_error("pyramid blur: unsupported tap count " + std::to_string(taps.n));
      return MB2_ERR_UNSUPPORTED;
  }
  return MB2_OK;
}

// TMA-staged blur + Hessian (+ in-level extremum pre-filter).  The plane must satisfy the tensor-map rules: 16-byte aligned base, pitch a
// multiple of 4 floats (every plane of the library has a pitch that is a multiple of 32 floats).
int mb2_launch_blur_tma(mb2_ctx* ctx, const ImgView& src, float* dst_blur, float* dst_resp, int dst_pitch, const BlurTaps& taps, float norm2,
                        const PrefilterArgs& pf) {
  if (((uintptr_t)src.p & 15) != 0 || (src.pitch & 3) != 0) { ctx->set_error("pyramid blur: plane not aligned for a tensor map"); return MB2_ERR_ARG; }
  switch (taps.n) {
    case 3: return launch_blur_tma_nt<3>(ctx, src, dst_blur, dst_resp, dst_pitch, taps, norm2, pf);
    case 5: return launch_blur_tma_nt<5>(ctx, src, dst_blur, dst_resp, dst_pitch, taps, norm2, pf);
    case 7: return launch_blur_tma_nt<7>(ctx, src, dst_blur, dst_resp, dst_pitch, taps, norm2, pf);
    case 9: return launch_blur_tma_nt<9>(ctx, src, dst_blur, dst_resp, dst_pitch, taps, norm2, pf);
    case 11: return launch_blur_tma_nt<11>(ctx, src, dst_blur, dst_resp, dst_pitch, taps, norm2, pf);
    case 13: return launch_blur_tma_nt<13>(ctx, src, dst_blur, dst_resp, dst_pitch, taps, norm2, pf);
    case 15: return launch_blur_tma_nt<15>(ctx, src, dst_blur, dst_resp, dst_pitch, taps, norm2, pf);
    case 17: return launch_blur_tma_nt<17>(ctx, src, dst_blur, dst_resp, dst_pitch, taps, norm2, pf);
    case 19: return launch_blur_tma_nt<19>(ctx, src, dst_blur, dst_resp, dst_pitch, taps, norm2, pf);
    case 21: return launch_blur_tma_nt<21>(ctx, src, dst_blur, dst_resp, dst_pitch, taps, norm2, pf);
    default:
      ctx->set_error("pyramid blur: unsupported tap count " + std::to_string(taps.n));
      return MB2_ERR_UNSUPPORTED;
  }
}
void mb2_launch_nms_finish(mb2_ctx* ctx, const OctaveLevels& oct, const Candidate* pre, const int* pre_count, int pre_cap, Candidate* out, int* count,
                           int capacity) {
  MB2_LAUNCH(ctx, k_nms_finish, ctx->num_sms * 4, 128, 0, oct, pre, pre_count, pre_cap, out, count, capacity);
}

void mb2_launch_hessian(mb2_ctx* ctx, const ImgView& src, float* dst, int dst_pitch, float norm2) {
  dim3 block(32, 8), grid((src.cols + 31) / 32, (src.rows + 7) / 8);
  MB2_LAUNCH(ctx, k_hessian, grid, block, 0, src, dst, dst_pitch, norm2);
}

void mb2_launch_resize_half(mb2_ctx* ctx, const ImgView& src, float* dst, int orows, int ocols, int dst_pitch) {
  dim3 block(32, 8), grid((ocols + 31) / 32, (orows + 7) / 8);
  MB2_LAUNCH(ctx, k_resize_half, grid, block, 0, src, dst, orows, ocols, dst_pitch);
}

void mb2_launch_nms(mb2_ctx* ctx, const ImgView& low, const ImgView& cur, const ImgView& high, int border, float posThr,
                    float negThr, int level, Candidate* out, int* count, int capacity) {
  int w = cur.cols - 2 * border, h = cur.rows - 2 * border;
  if (w <= 0 || h <= 0) return;
  dim3 block(32, 8), grid((cur.cols + 127) / 128, (h + 7) / 8);
  MB2_LAUNCH(ctx, k_nms, grid, block, 0, low, cur, high, border, posThr, negThr, level, out, count, capacity);
}

void mb2_launch_scale_requests(mb2_ctx* ctx, const KeypointRec* kps, int n, ScaleReq* out) {
  if (!n) return;
  MB2_LAUNCH(ctx, k_scale_requests, (n + 255) / 256, 256, 0, kps, n, out);
}
void mb2_launch_set_scales(mb2_ctx* ctx, KeypointRec* kps, int n, const float* d_s) {
  if (!n) return;
  MB2_LAUNCH(ctx, k_set_scales, (n + 255) / 256, 256, 0, kps, n, d_s);
}

void mb2_launch_fill_u64(mb2_ctx* ctx, unsigned long long* p, size_t n, unsigned long long v) {
  if (!n) return;
  MB2_LAUNCH(ctx, k_fill_u64, (unsigned)((n + 255) / 256), 256, 0, p, n, v);
}

void mb2_launch_localize(mb2_ctx* ctx, const OctaveLevels& oct, Candidate* cand, int n, const LocalizeParams& lp,
                         unsigned long long* octmap, Localized* out) {
  if (!n) return;
  MB2_LAUNCH(ctx, k_localize, (n + 127) / 128, 128, 0, oct, cand, n, lp, octmap, out);
}

void mb2_launch_emit(mb2_ctx* ctx, const OctaveLevels& oct, const Localized* loc, int n, const unsigned long long* octmap,
                     const LocalizeParams& lp, int octave, KeypointRec* out, int* count, int capacity) {
  if (!n) return;
  MB2_LAUNCH(ctx, k_emit_keypoints, (n + 127) / 128, 128, 0, oct, loc, n, octmap, lp, octave, out, count, capacity);
}
