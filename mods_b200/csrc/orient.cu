// Dominant orientation (one warp per region) and reprojection / boundary filtering.
//
// Reference: DetectOrientation (synth-detection.cpp:841-919), EstimateDominantAnglesFunctor
// (:746-839), smoothCircularBuffer / addPeakAngle (:721-744), computeGradientMagnitudeAndOrientation
// (detectors/helpers.cpp:840-863), ReprojectRegions / ReprojectByH (synth-detection.cpp:541-616,
// 490-498).
//
// Bit parity: the 36-bin histogram is a float sum per bin in raster order of the 39x39 inner
// patch; the votes are bucketed per bin (stable) and each lane adds the list of its own bin(s) in order.
#include "common.cuh"
#undef MB2_NS
#define MB2_NS mb2_orient_detail
#include "pyramid.cuh"

namespace MB2_NS {

constexpr int PS = 41, NPIX = PS * PS, OW = 4;  // warps per block
constexpr double K_SIGMA = 2 * 3.0 * 1.7320508075688772;  // synth-detection.cpp:28

// Per-warp working set (dynamic shared memory).  The histogram is built in chunks of CH_ROWS patch rows so that
// the vote buffers stay small (more resident warps): the per-bin chains simply continue from chunk to chunk.
constexpr int CH_ROWS = 8, CHP = ((CH_ROWS * PS + 31) / 32) * 32;
struct OriWarp {
  float patch[NPIX];
  float w[CHP], lists[CHP];
  unsigned char b[CHP];
  float hist[40];
  int cnt[40], cur[40];
};

extern __shared__ unsigned char s_ori_raw[];

__global__ void __launch_bounds__(OW * 32)
k_orientation(ImgView img, const KeyOut* __restrict__ in, int n, OrientParams op, const float* __restrict__ orimask,
              KeyOut* __restrict__ out, int* __restrict__ out_count) {
  OriWarp* s_w = reinterpret_cast<OriWarp*>(s_ori_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kidx = blockIdx.x * OW + warp;
  if (kidx >= n) return;
  const KeyOut k = in[kidx];
  OriWarp& W = s_w[warp];
  float* patch = W.patch; float* hist = W.hist;
  const int maxA = op.maxAngles;
  KeyOut* dst = out + (size_t)kidx * maxA;
  for (int j = lane; j < maxA; j += 32) dst[j].keep = 0;
  if (lane == 0) out_count[kidx] = 0;

  const float x = (float)k.v[0], y = (float)k.v[1];
  const float a11 = (float)k.v[2], a12 = (float)k.v[3], a21 = (float)k.v[4], a22 = (float)k.v[5];
  const double s = k.v[6];
  const int res = (int)(K_SIGMA * s);
  if (interpolateCheckBorders_dev(img.cols, img.rows, x, y, a11, a12, a21, a22, res, res)) return;
  if (maxA <= 0) return;

  const double mrScale = op.mrSize;
  const int patchImageSize = 2 * (int)mrScale + 1;
  const double imageToPatchScale = (double)patchImageSize / (double)op.patchSize;
  const float curr_sc = (float)(imageToPatchScale * s);
  const float A11 = fmul(a11, curr_sc), A12 = fmul(a12, curr_sc), A21 = fmul(a21, curr_sc), A22 = fmul(a22, curr_sc);
  const bool touch = interpolateCheckBorders_dev(img.cols, img.rows, x, y, A11, A12, A21, A22, PS, PS);
  // 41 rows x 3 segments of 14 samples: every lane busy in all four rounds
  for (int item = lane; item < PS * 3; item += 32) {
    const int row = item / 3, sg = item - row * 3;
    interpolate_seg(img.p, img.rows, img.cols, img.pitch, x, y, A11, A12, A21, A22, PS, PS, touch, row, sg * 14, 14,
                    [&](int i, float v) { patch[row * PS + i] = v; });
  }
  __syncwarp();
  // Gradient magnitude / orientation on the inner 39x39 (helpers.cpp:840-863), then the 36-bin histogram: bin b is
  // a float sum over ITS pixels in raster order.  Per chunk of rows: pass 1 evaluates the pixels (two per lane and
  // iteration, independent chains) and counts the votes per bin; pass 2 scatters the weights into per-bin lists, stable
  // in raster order (rank inside a slab of 32 pixels from __match_any_sync); pass 3: the owner lane of a bin adds its
  // list front to back -- the reference's chain of float additions for that bin, without replaying every pixel on
  // every lane.
  const float PIf = 3.14159265358979323846f;
  const unsigned lt_mask = (1u << lane) - 1u;
  auto vote = [&](int p, int& b, float& w) {
    b = 255; w = 0.f;
    const int r = p / PS, c = p - r * PS;
    if (c >= 1 && c < PS - 1) {
      const float xg = fsub(patch[p + 1], patch[p - 1]);
      const float yg = fsub(patch[p + PS], patch[p - PS]);
      const float mag = sqrtf(fadd(fmul(xg, xg), fmul(yg, yg)));
      const float m = orimask[p];
      if (m > 0 && (double)mag > 1.0) {
        // bin = (int)(36 (atan2LUTff(yg, xg) / pi + 1) / 2): a function of the LUT branch and index alone, tabulated on the host with
        // that float expression (host_tables.hpp); 36 (ori == +pi) is a slot the reference never reads
        b = g_atan_ori_bin[atan2LUT_code(yg, xg)];
        w = fmul(mag, m);
      }
    }
  };
  float h0 = 0.f, h1 = 0.f;   // bins lane and 32 + lane
  for (int row0 = 1; row0 < PS - 1; row0 += CH_ROWS) {
    const int pbeg = row0 * PS, npx = (min(row0 + CH_ROWS, PS - 1) - row0) * PS;
    for (int i = lane; i < 40; i += 32) W.cnt[i] = 0;
    __syncwarp();
    for (int q0 = 0; q0 < npx; q0 += 64) {
      const int qa = q0 + lane, qb = qa + 32;
      int ba = 255, bb = 255; float wa = 0.f, wb = 0.f;
      if (qa < npx) vote(pbeg + qa, ba, wa);
      if (qb < npx) vote(pbeg + qb, bb, wb);
      if (qa < npx) { W.b[qa] = (unsigned char)ba; W.w[qa] = wa; }
      if (qb < npx) { W.b[qb] = (unsigned char)bb; W.w[qb] = wb; }
      const unsigned ga = __match_any_sync(0xffffffffu, ba), gb = __match_any_sync(0xffffffffu, bb);
      if (ba != 255 && (ga & lt_mask) == 0) W.cnt[ba] += __popc(ga);   // lowest lane of the group
      __syncwarp();
      if (bb != 255 && (gb & lt_mask) == 0) W.cnt[bb] += __popc(gb);
      __syncwarp();
    }
    {  // exclusive scan of the 37 counts -> list starts
      const int c0 = W.cnt[lane], c1 = lane < 8 ? W.cnt[32 + lane] : 0;
      int i0 = c0, i1 = c1;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t0 = __shfl_up_sync(0xffffffffu, i0, d), t1 = __shfl_up_sync(0xffffffffu, i1, d);
        if (lane >= d) { i0 += t0; i1 += t1; }
      }
      const int tot0 = __shfl_sync(0xffffffffu, i0, 31);
      W.cur[lane] = i0 - c0;
      if (lane < 8) W.cur[32 + lane] = tot0 + i1 - c1;
    }
    __syncwarp();
    for (int q0 = 0; q0 < npx; q0 += 32) {
      const int q = q0 + lane;
      int b = 255; float w = 0.f;
      if (q < npx) { b = W.b[q]; w = W.w[q]; }
      const unsigned grp = __match_any_sync(0xffffffffu, b);
      if (b != 255) W.lists[W.cur[b] + __popc(grp & lt_mask)] = w;
      __syncwarp();
      if (b != 255 && (grp & lt_mask) == 0) W.cur[b] += __popc(grp);
      __syncwarp();
    }
    {  // after the scatter cur[b] is the END of list b
      const int e0 = W.cur[lane];
      for (int q = e0 - W.cnt[lane]; q < e0; q++) h0 = fadd(h0, W.lists[q]);
      if (lane < 4) { const int e1 = W.cur[32 + lane]; for (int q = e1 - W.cnt[32 + lane]; q < e1; q++) h1 = fadd(h1, W.lists[q]); }
    }
    __syncwarp();
  }
  hist[lane] = h0;
  if (lane < 4) hist[32 + lane] = h1;
  __syncwarp();
  // smoothCircularBuffer<36>, six times: new[i] = (old[i-1] + old[i]) + old[i+1] -- every element from the OLD values,
  // exactly as the reference's in-place loop with its `prev` / `first` temporaries (synth-detection.cpp:721-733)
  const int i1 = 32 + (lane & 3);
  for (int it = 0; it < 6; it++) {
    const float n0 = fadd(fadd(hist[lane == 0 ? 35 : lane - 1], hist[lane]), hist[lane + 1]);   // lane 31 reads hist[32]
    const float n1 = fadd(fadd(hist[i1 - 1], hist[i1]), hist[i1 == 35 ? 0 : i1 + 1]);
    __syncwarp();
    hist[lane] = n0;
    if (lane < 4) hist[i1] = n1;
    __syncwarp();
  }
  float thresh = fmaxf(hist[lane], lane < 4 ? hist[i1] : 0.0f);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) thresh = fmaxf(thresh, __shfl_xor_sync(0xffffffffu, thresh, d));
  thresh = fmaxf(thresh, 0.0f);
  thresh = (float)((double)thresh * op.threshold);
  if (op.half) {  // doHalfSIFT: hist[i] += hist[i + 18], upper half cleared (synth-detection.cpp:801-808), after the threshold is fixed
    const float lo = lane < 18 ? fadd(hist[lane], hist[lane + 18]) : 0.0f;
    __syncwarp();
    hist[lane] = lo;              // lanes 18..31 clear bins 18..31
    if (lane < 4) hist[32 + lane] = 0.0f;
    __syncwarp();
  }
  // peaks in the order of the reference's addPeakAngle calls: (35,0,1), (i-1,i,i+1) for i = 1..34, (34,35,0); first maxAngles
  auto is_peak = [&](int q) {
    const int a = (q == 0) ? 35 : q - 1, c = (q == 35) ? 0 : q + 1;
    return hist[q] >= thresh && hist[q] > hist[a] && hist[q] > hist[c];
  };
  const bool pk0 = is_peak(lane), pk1 = lane < 4 && is_peak(i1);
  const unsigned m0 = __ballot_sync(0xffffffffu, pk0), m1 = __ballot_sync(0xffffffffu, pk1);
  auto emit = [&](int q, int rank) {
    const int a = (q == 0) ? 35 : q - 1, c = (q == 35) ? 0 : q + 1;
    const float ha = hist[a], hb = hist[q], hc = hist[c];
    const float pp = fmul(fdiv(fsub(ha, hc), fadd(fsub(ha, fmul(2.0f, hb)), hc)), 0.5f);   // / 2.0f, exactly
    const float ang = fsub(fdiv(fmul(fmul(2.0f, PIf), fadd(fadd((float)q, 0.5f), pp)), 36.f), PIf);
    // The rotation needs cos(-ang), sin(-ang), which the reference takes from the host libm in
    // FLOAT (std::cos(float), synth-detection.cpp:902-903); libm's cosf/sinf are not correctly
    // rounded and differ between libm builds, so the host adapter evaluates them
    // and k_apply_rotation finishes the frame.
    KeyOut o = k;
    o.keep = 1;
    o.order = k.order;  // peak rank is implied by the slot index
    o.pad = __float_as_int(ang);
    dst[rank] = o;
  };
  const int r0 = __popc(m0 & lt_mask), r1 = __popc(m0) + __popc(m1 & lt_mask);
  if (pk0 && r0 < maxA) emit(lane, r0);
  if (pk1 && r1 < maxA) emit(i1, r1);
  if (lane == 0) out_count[kidx] = min(maxA, __popc(m0) + __popc(m1));
}

// A <- A * R(-theta) with (ci, si) = (cos(-theta), sin(-theta)) from the host (synth-detection.cpp:902-910)
__global__ void k_extract_angles(const KeyOut* __restrict__ keys, int n, float* __restrict__ ang) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ang[i] = __int_as_float(keys[i].pad);
}
__global__ void k_apply_rotation(KeyOut* __restrict__ keys, const double* __restrict__ cs, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double ci = cs[2 * i], si = cs[2 * i + 1];
  KeyOut k = keys[i];
  const double a11 = k.v[2], a12 = k.v[3], a21 = k.v[4], a22 = k.v[5];
  k.v[2] = a11 * ci - a12 * si;
  k.v[3] = a11 * si + a12 * ci;
  k.v[4] = a21 * ci - a22 * si;
  k.v[5] = a21 * si + a22 * ci;
  k.pad = 0;
  keys[i] = k;
}

// ReprojectRegions (synth-detection.cpp:541-616): reproj_kp = Hinv (affine part) applied to centre
// and shape (identity view: copy), then drop regions whose centre or k_sigma*s box leaves the
// ORIGINAL image.  det[i].keep is cleared in place for dropped regions; reproj gets the new frame.
__global__ void k_reproject(KeyOut* __restrict__ det, int n, const double* __restrict__ Hinv, int h_is_eye, int orig_w, int orig_h,
                            KeyOut* __restrict__ reproj) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  KeyOut d = det[i];
  KeyOut o = d;
  if (!h_is_eye) {
    o.v[0] = (Hinv[0] * d.v[0] + Hinv[1] * d.v[1] + Hinv[2]);
    o.v[1] = (Hinv[3] * d.v[0] + Hinv[4] * d.v[1] + Hinv[5]);
    o.v[2] = (Hinv[0] * d.v[2] + Hinv[1] * d.v[4]);
    o.v[3] = (Hinv[0] * d.v[3] + Hinv[1] * d.v[5]);
    o.v[4] = (Hinv[3] * d.v[2] + Hinv[4] * d.v[4]);
    o.v[5] = (Hinv[3] * d.v[3] + Hinv[4] * d.v[5]);
  }
  int keep = d.keep;
  if (keep) {
    keep = 0;
    if ((o.v[0] < orig_w) && (o.v[1] < orig_h) && (o.v[0] > 0) && (o.v[1] > 0)) {
      const int res = (int)(K_SIGMA * o.v[6]);
      if (!interpolateCheckBorders_dev(orig_w, orig_h, (float)o.v[0], (float)o.v[1], (float)o.v[2], (float)o.v[3], (float)o.v[4],
                                       (float)o.v[5], res, res))
        keep = 1;
    }
  }
  o.keep = keep;
  det[i].keep = keep;
  reproj[i] = o;
}

}  // namespace
using namespace MB2_NS;

void mb2_launch_orientation(mb2_ctx* ctx, const ImgView& img, const KeyOut* in, int n, const OrientParams& op,
                            const float* d_orimask, KeyOut* out, int* out_count_per_kp) {
  if (!n) return;
  const size_t smem = OW * sizeof(OriWarp);
  static unsigned long long attr_devs = 0;
  if (mb2_first_use_on_device(&attr_devs, ctx->device)) cudaFuncSetAttribute(k_orientation, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  MB2_LAUNCH(ctx, k_orientation, (n + OW - 1) / OW, OW * 32, smem, img, in, n, op, d_orimask, out, out_count_per_kp);
}

void mb2_launch_extract_angles(mb2_ctx* ctx, const KeyOut* keys, int n, float* d_ang) {
  if (!n) return;
  MB2_LAUNCH(ctx, k_extract_angles, (n + 127) / 128, 128, 0, keys, n, d_ang);
}

void mb2_launch_apply_rotation(mb2_ctx* ctx, KeyOut* keys, const double* d_cs, int n) {
  if (!n) return;
  MB2_LAUNCH(ctx, k_apply_rotation, (n + 127) / 128, 128, 0, keys, d_cs, n);
}

void mb2_launch_reproject(mb2_ctx* ctx, const KeyOut* det, int n, const double* Hinv9, int h_is_eye, int orig_w, int orig_h,
                          KeyOut* reproj) {
  if (!n) return;
  MB2_LAUNCH(ctx, k_reproject, (n + 127) / 128, 128, 0, (KeyOut*)det, n, Hinv9, h_is_eye, orig_w, orig_h, reproj);
}
