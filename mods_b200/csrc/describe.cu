// Affine-normalised 41x41 patch extraction + SIFT / RootSIFT, one CTA per region.
//
// Reference: DescribeRegions<> (synth-detection.hpp:169-255): sample a (P+2)^2 patch with the
// det-1 affine frame, blur it with sigma = 1.5*P/41 (helpers.cpp:726-731 -> OpenCV GaussianBlur,
// restated in oracle/cvmath.h), resample to 41x41, photometricallyNormalize (helpers.cpp:666-715);
// SIFTDescriptor (matching/siftdesc.cpp:22-71 bins, :73-131 samplePatch, :133-159 normalize,
// :199-278 SIFTnorm/RootSIFTnorm, :290-381 gradients).
//
// Design notes
//  * The final 41x41 bilinear resample reads the blurred patch at <= 82 distinct rows x 82 distinct
//    columns (the resampling matrix is diagonal), so the row pass is evaluated only at those
//    columns and the column pass only at those 82x82 points -- same values, O(P*82*taps) instead of
//    O(P^2*taps) work for large regions.
//  * Gaussian taps depend only on the integer half-size m = ceil(s*mrSize) and are built on the
//    host with libm (exactly the reference's arithmetic), one table per context.
//  * Bit parity: sample positions use the reference's running float sums; photometric sums are
//    serial float sums in raster order; every SIFT bin is accumulated in double by one thread
//    walking the patch in raster order (the reference touches a bin at most once per pixel).
#include "common.cuh"
#undef MB2_NS
#define MB2_NS mb2_describe_detail
#include "pyramid.cuh"
#include "describe.cuh"

namespace MB2_NS {

constexpr int PS = 41, NPIX = PS * PS, DT = 128;  // threads per CTA
constexpr int NEED = 2 * PS;                      // needed rows / columns of the blurred patch

struct KpGeom {  // per-region geometry shared by the planning and the describe kernels
  int m;          // int(ceil(s * mrSize))
  int P2;         // patchImageSize + 2 (0 => direct path)
  float scale;    // imageToPatchScale
};

__device__ __forceinline__ KpGeom kp_geom(const KeyOut& k, const DescribeParams& dp) {
  KpGeom g;
  if (!dp.fast) {
    const float mrScale = (float)ceil(k.v[6] * dp.mrSize);
    g.m = (int)mrScale;
    const int patchImageSize = 2 * g.m + 1;
    g.scale = fdiv((float)patchImageSize, (float)dp.patchSize);
    g.P2 = ((double)g.scale > 0.4) ? patchImageSize + 2 : 0;
  } else {
    const double mrScale = dp.mrSize * k.v[6];
    g.m = (int)mrScale;
    const int patchImageSize = 2 * g.m + 1;
    g.scale = (float)((double)patchImageSize / (double)dp.patchSize);
    g.P2 = 0;
  }
  return g;
}

// Size classes of the blur path.  Small regions keep the whole (P+2)^2 patch in shared memory and blur it
// densely; medium ones keep it in shared memory and blur only the <= 82 needed columns / rows; the rest
// (and the direct path) go through the global-scratch kernel.  Bounds are on m = ceil(s * mrSize).
constexpr int N_CLASSES = 6;
// class order = launch order = sort order: big regions first
enum { CLS_GLOBAL = 0, CLS_NEED_L = 1, CLS_NEED_S = 2, CLS_DENSE_L = 3, CLS_DENSE_M = 4, CLS_DENSE_S = 5 };
__host__ __device__ __forceinline__ int size_class(int m, int P2) {
  if (P2 == 0 || m > MB2_CLS_M_NEED_L) return CLS_GLOBAL;
  if (m > MB2_CLS_M_NEED_S) return CLS_NEED_L;
  if (m > MB2_CLS_M_DENSE_L) return CLS_NEED_S;
  if (m > MB2_CLS_M_DENSE_M) return CLS_DENSE_L;
  if (m > MB2_CLS_M_DENSE_S) return CLS_DENSE_M;
  return CLS_DENSE_S;
}

// scratch floats needed by region i: (P2*P2) sampled patch + (P2*NEED) row-pass columns (global-scratch class only);
// sort key = class | descending m | index, so that one radix sort yields the per-class lists, big regions first.
__global__ void k_plan(const KeyOut* __restrict__ kps, int n, DescribeParams dp, int max_m, unsigned long long* __restrict__ need,
                       int* __restrict__ too_big, unsigned long long* __restrict__ sum_p2sq, unsigned long long* __restrict__ keys,
                       int* __restrict__ cls_cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  KpGeom g = kp_geom(kps[i], dp);
  if (g.P2 > 0 && g.m > max_m) { atomicMax(too_big, g.m); }
  const int cls = size_class(g.m, g.P2);
  need[i] = (g.P2 > 0 && cls == CLS_GLOBAL) ? (unsigned long long)g.P2 * (unsigned long long)(g.P2 + NEED) : 0ull;
  if (g.P2 > 0) atomicAdd(sum_p2sq, (unsigned long long)g.P2 * (unsigned long long)g.P2);
  const int mm = g.m < 0 ? 0 : (g.m > 65535 ? 65535 : g.m);
  keys[i] = ((unsigned long long)cls << 48) | ((unsigned long long)(65535 - mm) << 32) | (unsigned long long)(unsigned)i;
  atomicAdd(&cls_cnt[cls], 1);
}

constexpr int MAXT_SMEM = 384;   // Gaussian taps kept in shared memory (larger kernels read them from global)

__global__ void __launch_bounds__(512)
k_extract(ImgView img, const KeyOut* __restrict__ kps, const unsigned long long* __restrict__ list, int n, DescribeParams dp,
          const TapTable taps, const unsigned long long* __restrict__ scratch_off, float* __restrict__ scratch,
          float* __restrict__ patch_out) {
  __shared__ float s_patch[NPIX];
  __shared__ float s_small[NEED * NEED]; // blurred patch at the needed rows x columns
  __shared__ int s_cols[NEED], s_rows[NEED];
  __shared__ float s_taps[MAXT_SMEM];

  // list = this class's slice of the sorted keys (big regions first, so they do not form a tail)
  const int tid = threadIdx.x, NT = blockDim.x;
  if ((int)blockIdx.x >= n) return;
  const int kidx = (int)(list[blockIdx.x] & 0xffffffffull);
  const KeyOut k = kps[kidx];
  const KpGeom g = kp_geom(k, dp);
  const float x = (float)k.v[0], y = (float)k.v[1];
  const float a11 = (float)k.v[2], a12 = (float)k.v[3], a21 = (float)k.v[4], a22 = (float)k.v[5];

  if (g.P2 == 0) {
    // heavy oversampling (or fast extraction): affine-normalise straight from the image, 3 threads per row
    const float A11 = fmul(a11, g.scale), A12 = fmul(a12, g.scale), A21 = fmul(a21, g.scale), A22 = fmul(a22, g.scale);
    const bool touch = interpolateCheckBorders_dev(img.cols, img.rows, x, y, A11, A12, A21, A22, PS, PS);
    if (tid < 3 * PS) {
      const int row = tid / 3, sg = tid - row * 3;
      interpolate_seg(img.p, img.rows, img.cols, img.pitch, x, y, A11, A12, A21, A22, PS, PS, touch, row, sg * 14, 14,
                      [&](int i, float v) { s_patch[row * PS + i] = v; });
    }
    __syncthreads();
  } else {
    const int P2 = g.P2;
    float* bufA = scratch + scratch_off[kidx];          // P2 x P2 sampled patch
    float* bufB = bufA + (size_t)P2 * P2;               // P2 x NEED row pass at the needed columns
    const int nt = taps.n[g.m], h = nt >> 1;
    const float* kw = taps.w + taps.off[g.m];
    if (nt <= MAXT_SMEM) { for (int j = tid; j < nt; j += NT) s_taps[j] = kw[j]; kw = s_taps; }
    // A. sample with the det-1 frame: rows are split into segments so that all threads sample
    {
      const bool touch = interpolateCheckBorders_dev(img.cols, img.rows, x, y, a11, a12, a21, a22, P2, P2);
      const int segs = P2 >= NT ? 1 : NT / P2;                 // segments per row
      const int slen = (P2 + segs - 1) / segs;
      for (int item = tid; item < P2 * segs; item += NT) {
        const int row = item / segs, sg = item - row * segs;
        float* orow = bufA + (size_t)row * P2;
        interpolate_seg(img.p, img.rows, img.cols, img.pitch, x, y, a11, a12, a21, a22, P2, P2, touch, row, sg * slen, slen,
                        [&](int i, float v) { orow[i] = v; });
      }
    }
    // positions of the second interpolate(): ofs = P2>>1, A = diag(scale); the reference's running sums
    const float ofs = (float)(P2 >> 1), sc = g.scale;
    const bool touch2 = interpolateCheckBorders_dev(P2, P2, ofs, ofs, sc, 0.f, 0.f, sc, PS, PS);
    if (tid == 0) {
      // columns: WX starts at rx - 20*a11 with rx = ofs - 20*a12 = ofs - 0, then += a11
      float rx = fsub(ofs, fmul((float)(PS >> 1), 0.f));
      float WX = fsub(rx, fmul((float)(PS >> 1), sc));
      for (int i = 0; i < PS; i++) {
        int xi = touch2 ? (int)floorf(WX) : (int)WX;
        s_cols[2 * i] = xi; s_cols[2 * i + 1] = xi + 1;
        WX = fadd(WX, sc);
      }
    } else if (tid == 32) {
      // rows: ry starts at ofs - 20*a22, += a22 per row; WY = ry - 20*a21 = ry - 0
      float ry = fsub(ofs, fmul((float)(PS >> 1), sc));
      for (int j = 0; j < PS; j++) {
        float WY = fsub(ry, fmul((float)(PS >> 1), 0.f));
        int yi = touch2 ? (int)floorf(WY) : (int)WY;
        s_rows[2 * j] = yi; s_rows[2 * j + 1] = yi + 1;
        ry = fadd(ry, sc);
      }
    }
    __syncthreads();
    // B. row pass (left-to-right accumulation, replicate border) at the needed columns, all rows.  The needed
    // columns come in adjacent pairs (x_i, x_i + 1): both are produced from one sliding window, one load per tap.
    for (int idx = tid; idx < P2 * PS; idx += NT) {
      const int r = idx / PS, i = idx - r * PS;
      const int c = s_cols[2 * i];
      float acc0 = 0.f, acc1 = 0.f;
      if (c >= 0 && c + 1 < P2) {
        const float* rowp = bufA + (size_t)r * P2;
        int cc = c - h;                                  // source column of tap 0 for output c
        float a = rowp[cc < 0 ? 0 : cc];
        float b = rowp[cc + 1 < 0 ? 0 : (cc + 1 > P2 - 1 ? P2 - 1 : cc + 1)];
        acc0 = fmul(kw[0], a); acc1 = fmul(kw[0], b);
        for (int j = 1; j < nt; j++) {
          a = b;
          const int cn = cc + j + 1;
          b = rowp[cn < 0 ? 0 : (cn > P2 - 1 ? P2 - 1 : cn)];
          const float kj = kw[j];
          acc0 = fadd(acc0, fmul(kj, a)); acc1 = fadd(acc1, fmul(kj, b));
        }
      } else {
        for (int q = 0; q < 2; q++) {                    // generic (never taken for in-range sample positions)
          const int cq = c + q;
          float acc = 0.f;
          if (cq >= 0 && cq < P2) {
            const float* rowp = bufA + (size_t)r * P2;
            int cc = cq - h; cc = cc < 0 ? 0 : cc;
            acc = fmul(kw[0], rowp[cc]);
            for (int j = 1; j < nt; j++) {
              cc = cq - h + j; cc = cc < 0 ? 0 : (cc > P2 - 1 ? P2 - 1 : cc);
              acc = fadd(acc, fmul(kw[j], rowp[cc]));
            }
          }
          if (q == 0) acc0 = acc; else acc1 = acc;
        }
      }
      bufB[(size_t)r * NEED + 2 * i] = acc0;
      bufB[(size_t)r * NEED + 2 * i + 1] = acc1;
    }
    __syncthreads();
    // C. column pass (symmetric pairs) at the needed rows x needed columns; the needed rows also come in
    // adjacent pairs (y_j, y_j + 1) that share all but two of their inputs.
    for (int idx = tid; idx < PS * NEED; idx += NT) {
      const int jr = idx / NEED, ci = idx - jr * NEED;
      const int r = s_rows[2 * jr], c = s_cols[ci];
      float acc0 = 0.f, acc1 = 0.f;
      if (c >= 0 && c < P2) {
        const float* col = bufB + ci;
        auto at = [&](int rr) { rr = rr < 0 ? 0 : (rr > P2 - 1 ? P2 - 1 : rr); return col[(size_t)rr * NEED]; };
        if (r >= 0 && r + 1 < P2) {
          float up0 = at(r + 1), dn1 = at(r);            // out1 is centred on r + 1: its centre tap is up0, out0's centre is dn1
          acc0 = fmul(kw[h], dn1); acc1 = fmul(kw[h], up0);
          float dn0_prev = dn1;                          // B[r - (j-1)] for out0 == B[r + 1 - j] for out1
          for (int j = 1; j <= h; j++) {
            const float up1 = at(r + 1 + j);             // out1: B[r+1+j];  out0 at the next j: B[r+j+1]
            const float dn0 = at(r - j);                 // out0: B[r-j]
            const float kj = kw[h + j];
            acc0 = fadd(acc0, fmul(kj, fadd(up0, dn0)));       // B[r+j] + B[r-j]
            acc1 = fadd(acc1, fmul(kj, fadd(up1, dn0_prev)));  // B[r+1+j] + B[r+1-j]
            up0 = up1; dn0_prev = dn0;
          }
        } else {
          for (int q = 0; q < 2; q++) {
            const int rq = r + q;
            float acc = 0.f;
            if (rq >= 0 && rq < P2) {
              acc = fmul(kw[h], at(rq));
              for (int j = 1; j <= h; j++) acc = fadd(acc, fmul(kw[h + j], fadd(at(rq + j), at(rq - j))));
            }
            if (q == 0) acc0 = acc; else acc1 = acc;
          }
        }
      }
      s_small[(2 * jr) * NEED + ci] = acc0;
      s_small[(2 * jr + 1) * NEED + ci] = acc1;
    }
    __syncthreads();
    // D. bilinear resample to 41x41 (interpolate(), helpers.cpp:551-626, with the same running sums), 3 threads per row
    if (tid < 3 * PS) {
      const int j = tid / 3, sg = tid - j * 3;
      float ry = fsub(ofs, fmul((float)(PS >> 1), sc));
      for (int q = 0; q < j; q++) ry = fadd(ry, sc);
      float rx = fsub(ofs, fmul((float)(PS >> 1), 0.f));
      for (int q = 0; q < j; q++) rx = fadd(rx, 0.f);
      float WX = fsub(rx, fmul((float)(PS >> 1), sc));
      float WY = fsub(ry, fmul((float)(PS >> 1), 0.f));
      const int ib = sg * 14, ie = min(PS, ib + 14);
      for (int i = 0; i < ib; i++) { WX = fadd(WX, sc); WY = fadd(WY, 0.f); }
      const int width = P2 - 1, height = P2 - 1;
      for (int i = ib; i < ie; i++) {
        const int xi = s_cols[2 * i], yi = s_rows[2 * j];
        float v = 0.f;
        const bool inside = touch2 ? (WX >= 0 && WY >= 0 && xi < width && yi < height) : true;
        if (inside) {
          const float wx = fsub(WX, (float)xi);
          const float r0x = s_small[(2 * j) * NEED + 2 * i], r0x1 = s_small[(2 * j) * NEED + 2 * i + 1];
          const float r1x = s_small[(2 * j + 1) * NEED + 2 * i], r1x1 = s_small[(2 * j + 1) * NEED + 2 * i + 1];
          const float I1 = fadd(fmul(wx, fsub(r0x1, r0x)), r0x);
          v = fadd(fmul(fsub(WY, (float)yi), fsub(fadd(fmul(wx, fsub(r1x1, r1x)), r1x), I1)), I1);
        }
        s_patch[j * PS + i] = v;
        WX = fadd(WX, sc); WY = fadd(WY, 0.f);
      }
    }
    __syncthreads();
  }

  for (int p = tid; p < NPIX; p += NT) patch_out[(size_t)kidx * NPIX + p] = s_patch[p];
}

// ---- shared-memory variants --------------------------------------------------------------------------
// Same arithmetic as k_extract (every output is the same chain of __fmul_rn / __fadd_rn), different staging:
// the sampled patch A lives in shared memory with h replicated columns on both sides (row stride SA = P2 + 2h,
// odd: conflict-free when consecutive lanes take consecutive rows), the row-pass result B with h replicated
// rows above and below, so neither pass clamps an index.
extern __shared__ float s_dyn[];

struct ResamplePos {   // second interpolate() (ofs = P2 >> 1, A = diag(scale)): per output column / row
  int xi[PS], yi[PS];
  float wx[PS], wy[PS];
};

__device__ __forceinline__ void resample_positions(ResamplePos& rp, int P2, float sc, bool touch2, int tid) {
  const float ofs = (float)(P2 >> 1);
  if (tid == 0) {
    // columns: WX starts at rx - 20*a11 with rx = ofs - 20*a12 = ofs - 0, then += a11
    float rx = fsub(ofs, fmul((float)(PS >> 1), 0.f));
    float WX = fsub(rx, fmul((float)(PS >> 1), sc));
    for (int i = 0; i < PS; i++) { rp.xi[i] = touch2 ? (int)floorf(WX) : (int)WX; rp.wx[i] = WX; WX = fadd(WX, sc); }
  } else if (tid == 32) {
    // rows: ry starts at ofs - 20*a22, += a22 per row; WY = ry - 20*a21 = ry - 0
    float ry = fsub(ofs, fmul((float)(PS >> 1), sc));
    for (int j = 0; j < PS; j++) {
      const float WY = fsub(ry, fmul((float)(PS >> 1), 0.f));
      rp.yi[j] = touch2 ? (int)floorf(WY) : (int)WY; rp.wy[j] = WY;
      ry = fadd(ry, sc);
    }
  }
}

// phase A: sample the P2 x P2 patch with the det-1 frame into A (padded rows), then replicate the borders
__device__ __forceinline__ void sample_to_smem(const ImgView& img, float x, float y, float a11, float a12, float a21, float a22,
                                               int P2, int h, int SA, float* A) {
  const int tid = threadIdx.x, NT = blockDim.x;
  const bool touch = interpolateCheckBorders_dev(img.cols, img.rows, x, y, a11, a12, a21, a22, P2, P2);
  const int segs = P2 >= NT ? 1 : NT / P2;                 // segments per row
  const int slen = (P2 + segs - 1) / segs;
  for (int item = tid; item < P2 * segs; item += NT) {
    const int row = item / segs, sg = item - row * segs;
    float* orow = A + row * SA + h;
    interpolate_seg(img.p, img.rows, img.cols, img.pitch, x, y, a11, a12, a21, a22, P2, P2, touch, row, sg * slen, slen,
                    [&](int i, float v) { orow[i] = v; });
  }
  __syncthreads();
  const int h2 = 2 * h;
  for (int t = tid; t < P2 * h2; t += NT) {
    const int r = t / h2, q = t - r * h2;
    float* row = A + r * SA;
    if (q < h) row[q] = row[h]; else row[P2 + q] = row[h + P2 - 1];
  }
}

// Dense variant (P2 <= ~83): every column / row of the blurred patch is produced, four adjacent outputs per
// thread from one sliding window.  Shared memory: A [P2][SA] (+4), B [P2 + 2h][P2], taps.
__global__ void __launch_bounds__(256)
k_extract_dense(ImgView img, const KeyOut* __restrict__ kps, const unsigned long long* __restrict__ list, int n, DescribeParams dp,
                const TapTable taps, float* __restrict__ patch_out) {
  __shared__ ResamplePos rp;
  const int tid = threadIdx.x, NT = blockDim.x;
  if ((int)blockIdx.x >= n) return;
  const int kidx = (int)(list[blockIdx.x] & 0xffffffffull);
  const KeyOut k = kps[kidx];
  const KpGeom g = kp_geom(k, dp);
  const int P2 = g.P2, nt = taps.n[g.m], h = nt >> 1, SA = P2 + 2 * h;
  float* A = s_dyn;
  float* B = A + P2 * SA + 4;
  float* kw = B + (P2 + 2 * h) * P2;
  float* S = A;   // blurred patch, P2 x P2, overwrites A after the row pass
  for (int j = tid; j < nt; j += NT) kw[j] = taps.w[taps.off[g.m] + j];
  const float sc = g.scale;
  const bool touch2 = interpolateCheckBorders_dev(P2, P2, (float)(P2 >> 1), (float)(P2 >> 1), sc, 0.f, 0.f, sc, PS, PS);
  resample_positions(rp, P2, sc, touch2, tid);
  sample_to_smem(img, (float)k.v[0], (float)k.v[1], (float)k.v[2], (float)k.v[3], (float)k.v[4], (float)k.v[5], P2, h, SA, A);
  __syncthreads();
  // row pass (left-to-right accumulation): lanes -> consecutive rows, 4 adjacent columns per item
  const int G = (P2 + 3) >> 2;
  for (int item = tid; item < P2 * G; item += NT) {
    const int g4 = item / P2, r = item - g4 * P2;
    const int c0 = min(4 * g4, P2 - 4);                    // last group overlaps the previous one (same values)
    const float* ap = A + r * SA + c0;                     // padded index of source column c0 - h
    float v0 = ap[0], v1 = ap[1], v2 = ap[2], v3 = ap[3];
    float kk = kw[0];
    float a0 = fmul(kk, v0), a1 = fmul(kk, v1), a2 = fmul(kk, v2), a3 = fmul(kk, v3);
#pragma unroll 4
    for (int j = 1; j < nt; j++) {
      v0 = v1; v1 = v2; v2 = v3; v3 = ap[j + 3];
      kk = kw[j];
      a0 = fadd(a0, fmul(kk, v0)); a1 = fadd(a1, fmul(kk, v1)); a2 = fadd(a2, fmul(kk, v2)); a3 = fadd(a3, fmul(kk, v3));
    }
    float* bp = B + (r + h) * P2 + c0;
    bp[0] = a0; bp[1] = a1; bp[2] = a2; bp[3] = a3;
  }
  __syncthreads();
  for (int t = tid; t < 2 * h * P2; t += NT) {            // replicate rows above / below
    const int q = t / P2, c = t - q * P2;
    if (q < h) B[q * P2 + c] = B[h * P2 + c]; else B[(P2 + q) * P2 + c] = B[(h + P2 - 1) * P2 + c];
  }
  __syncthreads();
  // column pass (symmetric pairs): lanes -> consecutive columns, 4 adjacent rows per item
  for (int item = tid; item < G * P2; item += NT) {
    const int gr = item / P2, c = item - gr * P2;
    const int r0 = min(4 * gr, P2 - 4);
    const float* bp = B + (r0 + h) * P2 + c;
    const float c0 = bp[0], c1 = bp[P2], c2 = bp[2 * P2], c3 = bp[3 * P2];
    float kk = kw[h];
    float a0 = fmul(kk, c0), a1 = fmul(kk, c1), a2 = fmul(kk, c2), a3 = fmul(kk, c3);
    float u0 = c1, u1 = c2, u2 = c3, u3, d0, d1 = c0, d2 = c1, d3 = c2;
#pragma unroll 4
    for (int j = 1; j <= h; j++) {
      u3 = bp[(3 + j) * P2]; d0 = bp[-j * P2];
      kk = kw[h + j];
      a0 = fadd(a0, fmul(kk, fadd(u0, d0))); a1 = fadd(a1, fmul(kk, fadd(u1, d1)));
      a2 = fadd(a2, fmul(kk, fadd(u2, d2))); a3 = fadd(a3, fmul(kk, fadd(u3, d3)));
      u0 = u1; u1 = u2; u2 = u3; d3 = d2; d2 = d1; d1 = d0;
    }
    float* sp = S + r0 * P2 + c;
    sp[0] = a0; sp[P2] = a1; sp[2 * P2] = a2; sp[3 * P2] = a3;
  }
  __syncthreads();
  // bilinear resample to 41x41 (interpolate(), helpers.cpp:551-626)
  const int width = P2 - 1, height = P2 - 1;
  for (int p = tid; p < NPIX; p += NT) {
    const int j = p / PS, i = p - j * PS;
    const int xi = rp.xi[i], yi = rp.yi[j];
    const float WX = rp.wx[i], WY = rp.wy[j];
    float v = 0.f;
    const bool inside = touch2 ? (WX >= 0 && WY >= 0 && xi < width && yi < height) : true;
    if (inside) {
      const float* s0 = S + yi * P2 + xi;
      const float wx = fsub(WX, (float)xi);
      const float r0x = s0[0], r0x1 = s0[1], r1x = s0[P2], r1x1 = s0[P2 + 1];
      const float I1 = fadd(fmul(wx, fsub(r0x1, r0x)), r0x);
      v = fadd(fmul(fsub(WY, (float)yi), fsub(fadd(fmul(wx, fsub(r1x1, r1x)), r1x), I1)), I1);
    }
    patch_out[(size_t)kidx * NPIX + p] = v;
  }
}

// Needed-columns variant (P2 up to ~150): A in shared memory, row pass only at the <= 82 columns the resample
// reads (pairs (x_i, x_i + 1) from one sliding window, two rows per item), column pass at the 82 x 82 points.
// Shared memory: A [P2][SA] (+4), B [P2 + 2h][SBN], taps;  the 82 x 82 result overwrites A.
constexpr int SBN = NEED + 1;
__global__ void __launch_bounds__(512)
k_extract_needed(ImgView img, const KeyOut* __restrict__ kps, const unsigned long long* __restrict__ list, int n, DescribeParams dp,
                 const TapTable taps, float* __restrict__ patch_out) {
  __shared__ ResamplePos rp;
  const int tid = threadIdx.x, NT = blockDim.x;
  if ((int)blockIdx.x >= n) return;
  const int kidx = (int)(list[blockIdx.x] & 0xffffffffull);
  const KeyOut k = kps[kidx];
  const KpGeom g = kp_geom(k, dp);
  const int P2 = g.P2, nt = taps.n[g.m], h = nt >> 1, SA = P2 + 2 * h;
  float* A = s_dyn;
  float* B = A + P2 * SA + 4;
  float* kw = B + (P2 + 2 * h) * SBN;
  float* S = A;   // 82 x 82 blurred values at the needed rows x columns (stride NEED)
  for (int j = tid; j < nt; j += NT) kw[j] = taps.w[taps.off[g.m] + j];
  const float sc = g.scale;
  const bool touch2 = interpolateCheckBorders_dev(P2, P2, (float)(P2 >> 1), (float)(P2 >> 1), sc, 0.f, 0.f, sc, PS, PS);
  resample_positions(rp, P2, sc, touch2, tid);
  sample_to_smem(img, (float)k.v[0], (float)k.v[1], (float)k.v[2], (float)k.v[3], (float)k.v[4], (float)k.v[5], P2, h, SA, A);
  __syncthreads();
  // row pass: item = (column pair i, rows rr and rr + R2); lanes -> consecutive rows
  const int R2 = (P2 + 1) >> 1;
  for (int item = tid; item < PS * R2; item += NT) {
    const int i = item / R2, rr = item - i * R2;
    const int rA = rr, rB = min(rr + R2, P2 - 1);
    const int c = rp.xi[i];
    float A0 = 0.f, A1 = 0.f, B0 = 0.f, B1 = 0.f;
    if (c >= 0 && c + 1 < P2) {
      const float* pa = A + rA * SA + c;                   // padded index of source column c - h
      const float* pb = A + rB * SA + c;
      float a = pa[0], b = pa[1], e = pb[0], f = pb[1];
      float kk = kw[0];
      A0 = fmul(kk, a); A1 = fmul(kk, b); B0 = fmul(kk, e); B1 = fmul(kk, f);
#pragma unroll 4
      for (int j = 1; j < nt; j++) {
        a = b; b = pa[j + 1]; e = f; f = pb[j + 1];
        kk = kw[j];
        A0 = fadd(A0, fmul(kk, a)); A1 = fadd(A1, fmul(kk, b)); B0 = fadd(B0, fmul(kk, e)); B1 = fadd(B1, fmul(kk, f));
      }
    } else {
      for (int q = 0; q < 2; q++) {                        // generic (only for sample positions outside the patch)
        const int cq = c + q;
        if (cq < 0 || cq >= P2) continue;
        const float* pa = A + rA * SA + cq;
        const float* pb = A + rB * SA + cq;
        float s0 = fmul(kw[0], pa[0]), s1 = fmul(kw[0], pb[0]);
        for (int j = 1; j < nt; j++) { s0 = fadd(s0, fmul(kw[j], pa[j])); s1 = fadd(s1, fmul(kw[j], pb[j])); }
        if (q == 0) { A0 = s0; B0 = s1; } else { A1 = s0; B1 = s1; }
      }
    }
    float* ba = B + (rA + h) * SBN + 2 * i;
    float* bb = B + (rB + h) * SBN + 2 * i;
    ba[0] = A0; ba[1] = A1; bb[0] = B0; bb[1] = B1;
  }
  __syncthreads();
  for (int t = tid; t < 2 * h * NEED; t += NT) {          // replicate rows above / below
    const int q = t / NEED, c = t - q * NEED;
    if (q < h) B[q * SBN + c] = B[h * SBN + c]; else B[(P2 + q) * SBN + c] = B[(h + P2 - 1) * SBN + c];
  }
  __syncthreads();
  // column pass (symmetric pairs) at the needed rows (y_j, y_j + 1) x needed columns
  for (int idx = tid; idx < PS * NEED; idx += NT) {
    const int jr = idx / NEED, ci = idx - jr * NEED;
    const int r = rp.yi[jr], c = rp.xi[ci >> 1] + (ci & 1);
    float acc0 = 0.f, acc1 = 0.f;
    if (c >= 0 && c < P2) {
      const float* col = B + h * SBN + ci;                 // col[rr * SBN] = row-pass value of patch row rr, rr in [-h, P2 - 1 + h]
      if (r >= 0 && r + 1 < P2) {
        const float* bp = col + r * SBN;
        float up0 = bp[SBN], dn1 = bp[0];                  // out1 is centred on r + 1, out0 on r
        acc0 = fmul(kw[h], dn1); acc1 = fmul(kw[h], up0);
        float dn0_prev = dn1;
#pragma unroll 4
        for (int j = 1; j <= h; j++) {
          const float up1 = bp[(1 + j) * SBN];
          const float dn0 = bp[-j * SBN];
          const float kj = kw[h + j];
          acc0 = fadd(acc0, fmul(kj, fadd(up0, dn0)));
          acc1 = fadd(acc1, fmul(kj, fadd(up1, dn0_prev)));
          up0 = up1; dn0_prev = dn0;
        }
      } else {
        for (int q = 0; q < 2; q++) {
          const int rq = r + q;
          float acc = 0.f;
          if (rq >= 0 && rq < P2) {
            const float* bp = col + rq * SBN;
            acc = fmul(kw[h], bp[0]);
            for (int j = 1; j <= h; j++) acc = fadd(acc, fmul(kw[h + j], fadd(bp[j * SBN], bp[-j * SBN])));
          }
          if (q == 0) acc0 = acc; else acc1 = acc;
        }
      }
    }
    S[(2 * jr) * NEED + ci] = acc0;
    S[(2 * jr + 1) * NEED + ci] = acc1;
  }
  __syncthreads();
  const int width = P2 - 1, height = P2 - 1;
  for (int p = tid; p < NPIX; p += NT) {
    const int j = p / PS, i = p - j * PS;
    const int xi = rp.xi[i], yi = rp.yi[j];
    const float WX = rp.wx[i], WY = rp.wy[j];
    float v = 0.f;
    const bool inside = touch2 ? (WX >= 0 && WY >= 0 && xi < width && yi < height) : true;
    if (inside) {
      const float wx = fsub(WX, (float)xi);
      const float r0x = S[(2 * j) * NEED + 2 * i], r0x1 = S[(2 * j) * NEED + 2 * i + 1];
      const float r1x = S[(2 * j + 1) * NEED + 2 * i], r1x1 = S[(2 * j + 1) * NEED + 2 * i + 1];
      const float I1 = fadd(fmul(wx, fsub(r0x1, r0x)), r0x);
      v = fadd(fmul(fsub(WY, (float)yi), fsub(fadd(fmul(wx, fsub(r1x1, r1x)), r1x), I1)), I1);
    }
    patch_out[(size_t)kidx * NPIX + p] = v;
  }
}

// photometricallyNormalize statistics (helpers.cpp:666-694): two serial float sums over the masked
// pixels in raster order.  Float addition is not associative, so the sums stay serial, but 32 regions
// are summed side by side by one warp: lane = region, the patches are staged through shared memory in
// coalesced 32-pixel slabs (padded rows: conflict-free columns).
// The slabs arrive through a ring of cp.async groups three slabs ahead of the sums (one warp per CTA cannot hide a DRAM round trip per
// slab otherwise: 320 us per 31k regions before, bound by exposed latency at 10 % occupancy).
constexpr int PN_T = 32, PN_SLAB = 32, PN_RING = 4, PN_NSLAB = (NPIX + PN_SLAB - 1) / PN_SLAB;
__global__ void __launch_bounds__(PN_T)
k_photonorm_stats(const float* __restrict__ patches, int n, const DescTables* __restrict__ tab, float2* __restrict__ stats) {
  __shared__ float s_tile[PN_RING][PN_T][PN_SLAB + 1];
  __shared__ unsigned s_mbits[PN_NSLAB];   // bit k of word j: pixel 32 j + k is inside the mask
  const int lane = threadIdx.x;
  const int r0 = blockIdx.x * PN_T;
  int inside = 0;
  for (int j = 0; j < PN_NSLAB; j++) {
    const int p = j * PN_SLAB + lane;
    const unsigned m = __ballot_sync(0xffffffffu, p < NPIX && tab->mask[p] > 0);
    if (lane == 0) s_mbits[j] = m;
    inside += __popc(m);
  }
  const float gsum = (float)inside;   // the reference adds 1.f per masked pixel: exact integers
  auto issue = [&](int s) {           // slab s of the two passes (pass 1 re-reads the patches) into ring slot s % PN_RING
    if (s < 2 * PN_NSLAB) {
      const int p = min((s % PN_NSLAB) * PN_SLAB + lane, NPIX - 1);   // clamped: the surplus columns of the last slab are never summed
      float* dst = &s_tile[s % PN_RING][0][lane];
#pragma unroll 8
      for (int rr = 0; rr < PN_T; rr++) {
        const int region = min(r0 + rr, n - 1);
        const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + rr * (PN_SLAB + 1));
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(patches + (size_t)region * NPIX + p) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int s = 0; s < PN_RING - 1; s++) issue(s);
  float sum = 0.f, var = 0.f;
  for (int s = 0; s < 2 * PN_NSLAB; s++) {
    issue(s + PN_RING - 1);   // into the slot consumed in the previous iteration
    asm volatile("cp.async.wait_group %0;" ::"n"(PN_RING - 1) : "memory");
    __syncwarp();
    const int j = s % PN_NSLAB;
    const unsigned mb = s_mbits[j];
    const float* mine = s_tile[s % PN_RING][lane];
    if (s < PN_NSLAB) {
#pragma unroll
      for (int k = 0; k < PN_SLAB; k++) if ((mb >> k) & 1u) sum = fadd(sum, mine[k]);
      if (s == PN_NSLAB - 1) sum = fdiv(sum, gsum);
    } else {
#pragma unroll
      for (int k = 0; k < PN_SLAB; k++) if ((mb >> k) & 1u) { const float d = fsub(sum, mine[k]); var = fadd(var, fmul(d, d)); }
    }
    __syncwarp();
  }
  var = sqrtf(fdiv(var, gsum));
  if (r0 + lane < n) stats[r0 + lane] = make_float2(sum, var);
}

// Normalisation apply (helpers.cpp:695-714) + per-pixel gradient record of the SIFT descriptor
// (siftdesc.cpp:290-345, 99-107): rec.x = mask * |grad|, rec.y = o = 8 * (ori + 2pi) / 2pi (its integer part
// mod 8 is the lower orientation bin, its fraction the weight of the upper one).  One CTA per region.
__global__ void __launch_bounds__(DT)
k_sift_grad(float* __restrict__ patches, int n, DescribeParams dp, const DescTables* __restrict__ tab,
            const float2* __restrict__ stats, float2* __restrict__ rec) {
  __shared__ float s_patch[NPIX];
  const int kidx = blockIdx.x, tid = threadIdx.x;
  if (kidx >= n) return;
  float* gp = patches + (size_t)kidx * NPIX;
  bool normalise = false;
  float sum = 0.f, fac = 0.f;
  if (dp.photoNorm) {
    const float2 st = stats[kidx];
    sum = st.x;
    if (!((double)st.y < 0.0001)) { normalise = true; fac = fdiv(50.0f, st.y); }   // helpers.cpp:695-697
  }
  for (int p = tid; p < NPIX; p += DT) {
    float v = gp[p];
    if (normalise) {
      v = fadd(128.f, fmul(fac, fsub(v, sum)));
      if (v > 255) v = 255;
      if (v < 0) v = 0;
      gp[p] = v;   // the normalised patch is what DescribeRegions hands to the descriptor
    }
    s_patch[p] = v;
  }
  __syncthreads();
  float2* out = rec + (size_t)kidx * NPIX;
  for (int p = tid; p < NPIX; p += DT) {
    const int r = p / PS, c = p - r * PS;
    float xg, yg;
    if (c == 0) xg = fsub(s_patch[p + 1], s_patch[p]);
    else if (c == PS - 1) xg = fsub(s_patch[p], s_patch[p - 1]);
    else xg = fsub(s_patch[p + 1], s_patch[p - 1]);
    if (r == 0) yg = fsub(s_patch[p + PS], s_patch[p]);
    else if (r == PS - 1) yg = fsub(s_patch[p], s_patch[p - PS]);
    else yg = fsub(s_patch[p + PS], s_patch[p - PS]);
    const float grad = sqrtf(fadd(fmul(xg, xg), fmul(yg, yg)));
    // o = (float)(8.0 * ((double)atan2LUTff(yg, xg) + 2 pi) / (2 pi)): a function of the LUT branch and index alone, tabulated on the
    // host with exactly that expression (host_tables.hpp) -- no double arithmetic per pixel
    const float o = __ldg(&g_atan_sift_o[atan2LUT_code(yg, xg)]);
    out[p] = make_float2(fmul(tab->mask[p], grad), o);
  }
}

// 4x4x8 votes (siftdesc.cpp:73-131), one warp per region.  Lane = (rb, cb, h): spatial bin (rb, cb) and
// one of its two 8-column segments (h = 0: columns 8cb..8cb+7 reach the bin through (bin1, w1); h = 1:
// columns 8cb+8..8cb+15 through (bin0, w0)); every lane walks its 16 x 8 pixels in raster order and adds
// into its own 8 orientation accumulators (double, in shared memory), then the two segments are summed.
// (Sums are in double: the association differs from the reference's single raster chain only at 1e-16
// relative, far below the 1/512 quantisation of the descriptor.)
constexpr int VW = 4;  // warps (regions) per CTA
__global__ void __launch_bounds__(VW * 32)
k_sift_votes(const float2* __restrict__ rec, int n, const DescTables* __restrict__ tab, double* __restrict__ vecT /* [128][n] */) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kidx = blockIdx.x * VW + warp;
  if (kidx >= n) return;
  const int rb = lane >> 3, cb = (lane >> 1) & 3, h = lane & 1;
  // the lane's 8 orientation accumulators in shared memory, bin-major: s_acc[b][lane] puts the 32 lanes of a warp on distinct
  // banks whatever their bins (a 64-bit access of a full warp is then the minimal 2 wavefronts; the lane-major layout used before
  // cost 4+ and made the kernel shared-memory bound).  The two votes of a pixel go to different bins (bo1 = bo0 + 1 mod 8), so both
  // are loaded before either is stored: one dependent shared-memory round trip per pixel instead of two.
  __shared__ double s_acc[VW][8][32];
  double* acc = &s_acc[warp][0][lane];
#pragma unroll
  for (int b = 0; b < 8; b++) acc[b * 32] = 0.0;
  const float2* R = rec + (size_t)kidx * NPIX;
  const int c_lo = 8 * cb + 8 * h;
  float wcol[8];
#pragma unroll
  for (int t = 0; t < 8; t++) {
    const int c = min(c_lo + t, PS - 1);
    const float w = h == 0 ? (tab->bin1[c] == cb * 8 ? tab->w1[c] : 0.f) : (tab->bin0[c] == cb * 8 ? tab->w0[c] : 0.f);
    wcol[t] = c_lo + t < PS ? w : 0.f;
  }
  for (int tr = 0; tr < 16; tr++) {
    const int r = 8 * rb + tr;
    if (r >= PS) break;
    float wr = 0.f;
    if (tr < 8) { if (tab->bin1[r] == rb * 8) wr = tab->w1[r]; }
    else { if (tab->bin0[r] == rb * 8) wr = tab->w0[r]; }
    if (!(wr > 0)) continue;
    const float2* row = R + r * PS + c_lo;
#pragma unroll
    for (int t = 0; t < 8; t++) {
      if (c_lo + t >= PS) break;
      const float2 px = __ldg(row + t);
      const float wc = fmul(wcol[t], px.x);
      const float val = fmul(wr, wc);
      if (val > 0) {
        const int io = (int)px.y;
        const float wo1 = fsub(px.y, (float)io);
        const int bo0 = io & 7, bo1 = (bo0 + 1) & 7;   // io >= 0: io % 8 == io & 7
        const double a0 = acc[bo0 * 32], a1 = acc[bo1 * 32];
        acc[bo0 * 32] = a0 + (double)fmul(val, fsub(1.0f, wo1));
        acc[bo1 * 32] = a1 + (double)fmul(val, wo1);
      }
    }
  }
  __syncwarp();
  // lanes (rb, cb, 0) and (rb, cb, 1) are neighbours: h = 0 segment first (lower columns), as in a raster walk
  for (int b = lane & 1 ? 4 : 0, e = b + 4; b < e; b++) {
    const double v = s_acc[warp][b][lane & ~1] + s_acc[warp][b][lane | 1];
    vecT[(size_t)((rb * 4 + cb) * 8 + b) * n + kidx] = v;
  }
}

// SIFTnorm / RootSIFTnorm (siftdesc.cpp:133-159, 199-222, 247-262) + quantisation, one thread per
// region (serial double sums in index order, as the reference); vecT is bin-major so the 128 reads
// of neighbouring threads coalesce.
__global__ void __launch_bounds__(128)
k_sift_finish(double* __restrict__ vecT, int n, DescribeParams dp, uint8_t* __restrict__ desc_out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const double maxBinValue = (double)0.2f;
  int D = 128;
  if (dp.half) {
    // HalfSIFT (siftdesc.cpp:408-421): the raw votes (doNorm is off for this pass) of opposite orientations are summed,
    // half_vec[i*4 + j] = vec[i*8 + j] + vec[i*8 + j + 4]; the normalisation below then runs on 64 entries
    D = 64;
    for (int i = 0; i < 16; i++)
      for (int j = 0; j < 4; j++) {
        const double h = vecT[(size_t)(i * 8 + j) * n + r] + vecT[(size_t)(i * 8 + j + 4) * n + r];
        vecT[(size_t)(i * 4 + j) * n + r] = h;   // i*4+j <= i*8+j: entries still to be read are never overwritten
      }
  }
  for (int pass = 0; pass < 2; pass++) {
    double len = 0.0;
    for (int i = 0; i < D; i += 4) {
      const double v0 = vecT[(size_t)i * n + r], v1 = vecT[(size_t)(i + 1) * n + r], v2 = vecT[(size_t)(i + 2) * n + r],
                   v3 = vecT[(size_t)(i + 3) * n + r];
      const double sq0 = v0 * v0, sq1 = v1 * v1, sq2 = v2 * v2, sq3 = v3 * v3;
      len += sq0 + sq1 + sq2 + sq3;
    }
    len = sqrt(len);
    const double fac = 1.0 / len;
    bool changed = false;
    for (int i = 0; i < D; i++) {
      double v = vecT[(size_t)i * n + r] * fac;
      if (pass == 0 && v > maxBinValue) { v = maxBinValue; changed = true; }
      vecT[(size_t)i * n + r] = v;
    }
    if (pass == 1 || !changed) break;
  }
  double sum = 0.;
  if (dp.rootSIFT) for (int i = 0; i < D; i++) sum += fabs(vecT[(size_t)i * n + r]);
  for (int i = 0; i < D; i++) {
    double v = vecT[(size_t)i * n + r];
    if (dp.rootSIFT) v = sqrt(v / sum);
    // (int)(512.0 * v + 0.5) for RootSIFT, (int)(512.0f * v + 0.5) for SIFT: identical in double
    int b = (int)(512.0 * v + 0.5);
    b = b < 0 ? 0 : (b > 255 ? 255 : b);
    desc_out[(size_t)r * 128 + i] = (uint8_t)b;
  }
  for (int i = D; i < 128; i++) desc_out[(size_t)r * 128 + i] = 0;
}

// DSPSIFT (imagerepresentation.cpp:1547-1598): the raw votes of one measurement-region size, narrowed to float as SIFTDescriptor::operator()
// hands them out (desc[i] = (float)vec[i], siftdesc.cpp:436-440), added in float to the running sum of the domains
__global__ void k_dsp_accumulate(const double* __restrict__ vecT, int n, float* __restrict__ acc, int first) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 128) return;
  const int r = i >> 7, b = i & 127;
  const float v = (float)vecT[(size_t)b * n + r];
  acc[i] = first ? v : fadd(acc[i], v);
}
// SIFTnorm(std::vector<float>&) with the float normalize (siftdesc.cpp:160-184, 263-278), one thread per region
__global__ void __launch_bounds__(128) k_dsp_norm(float* __restrict__ acc, int n, uint8_t* __restrict__ desc_out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  float* v = acc + (size_t)r * 128;
  const double maxBinValue = (double)0.2f;
  for (int pass = 0; pass < 2; pass++) {
    float len = 0.0f;
    for (int i = 0; i < 128; i += 4) {
      const float sq0 = fmul(v[i], v[i]), sq1 = fmul(v[i + 1], v[i + 1]), sq2 = fmul(v[i + 2], v[i + 2]), sq3 = fmul(v[i + 3], v[i + 3]);
      len = fadd(len, fadd(fadd(fadd(sq0, sq1), sq2), sq3));
    }
    len = (float)__dsqrt_rn((double)len);
    const float fac = (float)__ddiv_rn(1.0, (double)len);
    bool changed = false;
    for (int i = 0; i < 128; i++) {
      float x = fmul(v[i], fac);
      if (pass == 0 && (double)x > maxBinValue) { x = (float)maxBinValue; changed = true; }
      v[i] = x;
    }
    if (pass == 1 || !changed) break;
  }
  for (int i = 0; i < 128; i++) {
    int b = (int)d_add((double)fmul(512.0f, v[i]), 0.5);
    b = b < 0 ? 0 : (b > 255 ? 255 : b);
    desc_out[(size_t)r * 128 + i] = (uint8_t)b;
  }
}

}  // namespace
using namespace MB2_NS;

void mb2_launch_dsp_accumulate(mb2_ctx* ctx, const double* d_vecT, int n, float* d_acc, int first) {
  if (n) MB2_LAUNCH(ctx, k_dsp_accumulate, (n * 128 + 255) / 256, 256, 0, d_vecT, n, d_acc, first);
}
void mb2_launch_dsp_norm(mb2_ctx* ctx, float* d_acc, int n, uint8_t* d_desc) {
  if (n) MB2_LAUNCH(ctx, k_dsp_norm, (n + 127) / 128, 128, 0, d_acc, n, d_desc);
}

int mb2_describe_plan(mb2_ctx* ctx, const KeyOut* kps, int n, const DescribeParams& dp, int max_m, unsigned long long* d_need,
                      int* d_too_big, unsigned long long* d_sum_p2sq, unsigned long long* d_keys, int* d_cls_cnt) {
  if (!n) return MB2_OK;
  MB2_LAUNCH(ctx, k_plan, (n + 127) / 128, 128, 0, kps, n, dp, max_m, d_need, d_too_big, d_sum_p2sq, d_keys, d_cls_cnt);
  return MB2_OK;
}

namespace {
// dynamic shared memory of a class = the largest region it may hold
size_t class_smem_bytes(int cls, const int* tap_n, int max_m) {
  int lo, hi; bool dense;
  switch (cls) {
    case CLS_DENSE_S: lo = 0; hi = MB2_CLS_M_DENSE_S; dense = true; break;
    case CLS_DENSE_M: lo = MB2_CLS_M_DENSE_S + 1; hi = MB2_CLS_M_DENSE_M; dense = true; break;
    case CLS_DENSE_L: lo = MB2_CLS_M_DENSE_M + 1; hi = MB2_CLS_M_DENSE_L; dense = true; break;
    case CLS_NEED_S: lo = MB2_CLS_M_DENSE_L + 1; hi = MB2_CLS_M_NEED_S; dense = false; break;
    case CLS_NEED_L: lo = MB2_CLS_M_NEED_S + 1; hi = MB2_CLS_M_NEED_L; dense = false; break;
    default: return 0;
  }
  size_t best = 0;
  for (int m = lo; m <= hi && m <= max_m; m++) {
    const int nt = tap_n[m];
    if (!nt) continue;
    const size_t P2 = 2 * (size_t)m + 3, h = nt >> 1, SA = P2 + 2 * h;
    const size_t fl = P2 * SA + 4 + (P2 + 2 * h) * (dense ? P2 : (size_t)SBN) + nt + 4;
    best = std::max(best, fl * 4);
  }
  return best;
}
}  // namespace

int mb2_launch_describe_kernel(mb2_ctx* ctx, const ImgView& img, const KeyOut* kps, int n, const DescribeParams& dp,
                               const DescTables* d_tables, const TapTable& taps, const int* tap_n_host, const ExtractPlan& plan,
                               const unsigned long long* d_off, float* d_scratch,
                               uint8_t* d_desc, float* d_patches, float2* d_stats, double* d_vecT, float2* d_rec) {
  if (!n) return MB2_OK;
  static unsigned long long attr_devs = 0;
  if (mb2_first_use_on_device(&attr_devs, ctx->device)) {
    cudaFuncSetAttribute(k_extract_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_extract_needed, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  }
  // The few very large regions of the global-scratch class take as long as thousands of small ones: they run on
  // the side stream, next to the shared-memory classes (per-kernel profiling keeps one stream).
  const bool fork = !ctx->profiling && plan.count[CLS_GLOBAL] > 0 && plan.count[CLS_GLOBAL] < n;
  int first = 0;
  for (int cls = 0; cls < N_CLASSES; cls++) {
    const int cnt = plan.count[cls];
    if (cnt > 0) {
      const unsigned long long* list = plan.sorted + first;
      const size_t smem = class_smem_bytes(cls, tap_n_host, taps.max_m);
      switch (cls) {
        case CLS_GLOBAL:
          if (fork) {
            cudaEventRecord(ctx->ev_fork, ctx->stream);
            cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0);
            MB2_LAUNCH_ON(ctx, ctx->side, k_extract, cnt, 512, 0, img, kps, list, cnt, dp, taps, d_off, d_scratch, d_patches);
            cudaEventRecord(ctx->ev_join, ctx->side);
          } else {
            MB2_LAUNCH(ctx, k_extract, cnt, 512, 0, img, kps, list, cnt, dp, taps, d_off, d_scratch, d_patches);
          }
          break;
        case CLS_NEED_L: MB2_LAUNCH(ctx, k_extract_needed, cnt, 512, smem, img, kps, list, cnt, dp, taps, d_patches); break;
        case CLS_NEED_S: MB2_LAUNCH(ctx, k_extract_needed, cnt, 256, smem, img, kps, list, cnt, dp, taps, d_patches); break;
        case CLS_DENSE_L: MB2_LAUNCH(ctx, k_extract_dense, cnt, 256, smem, img, kps, list, cnt, dp, taps, d_patches); break;
        default: MB2_LAUNCH(ctx, k_extract_dense, cnt, 128, smem, img, kps, list, cnt, dp, taps, d_patches); break;
      }
    }
    first += cnt;
  }
  if (fork) cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0);
  if (dp.photoNorm) MB2_LAUNCH(ctx, k_photonorm_stats, (n + PN_T - 1) / PN_T, PN_T, 0, d_patches, n, d_tables, d_stats);
  MB2_LAUNCH(ctx, k_sift_grad, n, DT, 0, d_patches, n, dp, d_tables, d_stats, d_rec);
  MB2_LAUNCH(ctx, k_sift_votes, (n + VW - 1) / VW, VW * 32, 0, d_rec, n, d_tables, d_vecT);
  if (!dp.raw) MB2_LAUNCH(ctx, k_sift_finish, (n + 127) / 128, 128, 0, d_vecT, n, dp, d_desc);   // raw: the caller takes the votes (DSPSIFT)
  return MB2_OK;
}
