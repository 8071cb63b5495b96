// mods_b200: shared device/host helpers.
//
// Floating-point contract: the reference is compiled without -march, so it
// never executes an FMA; every parity-critical kernel is built with -fmad=false and writes the
// reference's evaluation order out explicitly.  Where an FMA is provably exact (integer-valued
// descriptor arithmetic) it is requested explicitly with __fmaf_rn.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/mods_b200.h"

#define MB2_CUDA_CHECK(ctx, expr)                                                              \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      (ctx)->set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                    \
      return MB2_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

struct DevBuf {  // grow-only device scratch
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  template <class T> T* as() { return (T*)p; }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct HostBuf {  // grow-only pinned staging
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  template <class T> T* as() { return (T*)p; }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// A device-resident set of described regions (what RegionVectorMap[det][desc] holds on the host
// side of the reference, imagerepresentation.h:66) -- enough of it for MatchFlannFGINN.
struct RegionSlot {
  DevBuf desc;   // n x 128 u8
  DevBuf xy;     // n x 2 f64  (reproj_kp.x, reproj_kp.y)
  int n = 0;
};

struct mb2_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t side = nullptr;            // forked from `stream` for launches that would otherwise leave a long tail
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::string err;
  long long launches = 0;
  int num_sms = 148;
  // scratch, reused across calls
  DevBuf img, pyr, resp, cand, misc, kp_a, kp_b, kp_c, desc_u8, patch_scratch, nn_a, nn_b, nn_c, nn_d, rs_a, rs_b, rs_c, rs_u, rs_v, dog_taps, dog_tmp;
  DevBuf octmap;
  HostBuf h_a, h_b, h_c;
  RegionSlot slots[MB2_MAX_SLOTS];
  void* tmap_encode = nullptr;  // cuTensorMapEncodeTiled, fetched through the runtime
  int last_view_n = 0;          // regions of the most recent mb2_detect_describe_view (still in kp_b / rs_c / desc_u8)
  // optional per-kernel timing (mb2_ctx_profile_begin/end): CUDA events around every launch, on `stream`
  struct ProfRec { const char* name; cudaEvent_t a, b; };
  bool profiling = false;
  std::vector<ProfRec> prof;
  void* mser_state = nullptr;   // MserBufs (mser.cu), released by mb2_mser_release
  DevBuf img2, pair_keys;       // mb2_mser_detect_pair: second image, keys of both images
  DevBuf synth_a, synth_b, synth_c, synth_k;   // view synthesis: rotated image, view, blur scratch, taps
  int pair_n[2] = {0, 0}, pair_w = 0, pair_h = 0;
  // the component-tree kernel is latency bound and suffers badly from bandwidth-hungry neighbours: other contexts can order their
  // work behind it (mb2_ctx_wait_tree).  tree_epoch counts the tree kernels launched so far (read by other host threads).
  cudaEvent_t ev_tree = nullptr;
  volatile long long tree_epoch = 0;
  unsigned long long prof_extract_bytes = 0;  // algorithmic gather bytes of the patch-extraction launches
  cudaError_t launch_error = cudaSuccess;     // first failed kernel launch (MB2_LAUNCH); reported by mb2_ctx_sync and the C entry points
  void set_error(const std::string& s) { err = s; }
  // every device buffer of the context in one list, so that mb2_ctx_destroy cannot forget a member
  void release_buffers() {
    DevBuf* all[] = {&img, &pyr, &resp, &cand, &misc, &kp_a, &kp_b, &kp_c, &desc_u8, &patch_scratch, &nn_a, &nn_b, &nn_c, &nn_d, &rs_a, &rs_b, &rs_c, &rs_u, &rs_v,
                     &dog_taps, &dog_tmp, &octmap, &img2, &pair_keys, &synth_a, &synth_b, &synth_c, &synth_k};
    for (DevBuf* b : all) b->release();
    HostBuf* hall[] = {&h_a, &h_b, &h_c};
    for (HostBuf* b : hall) b->release();
    for (RegionSlot& s : slots) { s.desc.release(); s.xy.release(); s.n = 0; }
  }
};
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE setting: one bit per device in a per-call-site mask
inline bool mb2_first_use_on_device(unsigned long long* mask, int device) {
  const unsigned long long bit = 1ull << (device & 63);
  if (*mask & bit) return false;
  *mask |= bit;
  return true;
}

// ---- exact (non-contracted) float helpers ---------------------------------------------------
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double d_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double d_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double d_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double d_div(double a, double b) { return __ddiv_rn(a, b); }

// ---- device image view ----------------------------------------------------------------------
struct ImgView {
  const float* p;
  int rows, cols, pitch;  // pitch in floats
  __device__ __forceinline__ float at(int r, int c) const { return p[(size_t)r * pitch + c]; }
};

// atan2LUTff table (detectors/helpers.cpp:30-72), uploaded once per process by capi.cu.
// The library is built as ONE translation unit (mods_b200.cu includes every kernel file), so this
// is the single definition.
__constant__ double c_atan_lut[256];
// what the hot kernels derive from atan2LUTff's result alone, per (branch, table index): host_tables.hpp build_atan_derived
__device__ float g_atan_sift_o[2049];
__device__ unsigned char g_atan_ori_bin[2049];

// detectors/helpers.cpp:160-207.  The eight branches of the reference differ only in which of |x|, |y| is the
// numerator and in the final affine map of the table value, so the table is read ONCE (index from the same
// float multiply / divide as the taken branch) and the quadrant formula is selected afterwards: identical
// results without eight divergent paths.  lut: 256 doubles (shared-memory copy of c_atan_lut in the hot kernels).
__device__ __forceinline__ float atan2LUTff_dev(float y, float x, const double* __restrict__ lut) {
  const double PI_2d = (double)1.57079632679489661923f, PId = (double)3.14159265358979323846f;
  const float ax = fabsf(x), ay = fabsf(y);
  const bool big = ax > ay;                       // x > y, x > absy, absx > y, absx > absy in the four quadrants
  const float num = big ? ay : ax, den = big ? ax : ay;
  const double t = lut[(int)(fdiv(fmul(255.f, num), den))];
  if (x > 0.f) {
    if (y > 0.f) return big ? (float)t : (float)d_sub(PI_2d, t);
    return big ? (float)(-t) : (float)d_add(-PI_2d, t);
  }
  if (y > 0.f) return big ? (float)d_sub(PId, t) : (float)d_add(PI_2d, t);
  if (big) return (float)d_add(-PId, t);
  if (x == 0.f) return 0.f;
  return (float)d_sub(-PI_2d, t);
}

// (branch, table index) of atan2LUTff for (y, x): same index arithmetic and branch conditions as atan2LUTff_dev above
__device__ __forceinline__ int atan2LUT_code(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const bool big = ax > ay;
  const float num = big ? ay : ax, den = big ? ax : ay;
  const int idx = (int)(fdiv(fmul(255.f, num), den));
  if (!(x > 0.f) && !(y > 0.f) && !big && x == 0.f) return 2048;
  return (((x > 0.f) ? 4 : 0) | ((y > 0.f) ? 2 : 0) | (big ? 1 : 0)) * 256 + idx;
}

// detectors/helpers.cpp:524-549
__device__ __forceinline__ bool interpolateCheckBorders_dev(int orig_w, int orig_h, float ofsx, float ofsy, float a11, float a12,
                                                            float a21, float a22, int res_w, int res_h) {
  const int width = orig_w - 2, height = orig_h - 2;
  const float halfWidth = (float)ceil((double)(float)res_w / 2.0);
  const float halfHeight = (float)ceil((double)(float)res_h / 2.0);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float xs = (i < 2) ? -halfWidth : halfWidth;
    const float ys = (i & 1) ? halfHeight : -halfHeight;
    float imx = fadd(fadd(ofsx, fmul(xs, a11)), fmul(ys, a12));
    float imy = fadd(fadd(ofsy, fmul(xs, a21)), fmul(ys, a22));
    if (floorf(imx) <= 0 || floorf(imy) <= 0 || ceilf(imx) >= width || ceilf(imy) >= height) return true;
  }
  return false;
}

// A run of `cnt` samples starting at column i0 of output row `row` of interpolate()
// (detectors/helpers.cpp:551-626).  The reference advances the sample position with running float
// sums: row `row` is reached by `row` additions of (a12, a22) starting at (ofs - halfHeight*a12,
// ofs - halfHeight*a22), column i by i additions of (a11, a21) from the row start; a thread that starts
// in the middle of a row first replays those additions (2 FADDs per skipped sample), so any split of the
// patch over threads yields the reference's positions bit for bit.  im may be global or shared memory.
template <class Out>
__device__ __forceinline__ void interpolate_seg(const float* __restrict__ im, int im_rows, int im_cols, int im_pitch,
                                                float ofsx, float ofsy, float a11, float a12, float a21, float a22,
                                                int rw, int rh, bool touch, int row, int i0, int cnt, Out out /* out(i, value) */) {
  const int halfWidth = rw >> 1, halfHeight = rh >> 1;
  float rx = fsub(ofsx, fmul((float)halfHeight, a12));
  float ry = fsub(ofsy, fmul((float)halfHeight, a22));
  for (int j = 0; j < row; ++j) { rx = fadd(rx, a12); ry = fadd(ry, a22); }
  float WX = fsub(rx, fmul((float)halfWidth, a11));
  float WY = fsub(ry, fmul((float)halfWidth, a21));
  for (int i = 0; i < i0; ++i) { WX = fadd(WX, a11); WY = fadd(WY, a21); }
  const int width = im_cols - 1, height = im_rows - 1;
  const int i1 = min(rw, i0 + cnt);
  for (int i = i0; i < i1; ++i) {
    float v;
    if (!touch) {
      const int x = (int)WX, y = (int)WY;
      const float wx = fsub(WX, (float)x);
      const float* R0 = im + y * im_pitch + x;
      const float* R1 = R0 + im_pitch;
      const float I1 = fadd(fmul(wx, fsub(R0[1], R0[0])), R0[0]);
      v = fadd(fmul(fsub(WY, (float)y), fsub(fadd(fmul(wx, fsub(R1[1], R1[0])), R1[0]), I1)), I1);
    } else {
      const int x = (int)floorf(WX), y = (int)floorf(WY);
      if (WX >= 0 && WY >= 0 && x < width && y < height) {
        const float wx = fsub(WX, (float)x);
        const float* R0 = im + y * im_pitch + x;
        const float* R1 = R0 + im_pitch;
        const float I1 = fadd(fmul(wx, fsub(R0[1], R0[0])), R0[0]);
        v = fadd(fmul(fsub(WY, (float)y), fsub(fadd(fmul(wx, fsub(R1[1], R1[0])), R1[0]), I1)), I1);
      } else v = 0.f;
    }
    out(i, v);
    WX = fadd(WX, a11); WY = fadd(WY, a21);
  }
}

// One whole output row.
template <class Out>
__device__ __forceinline__ void interpolate_row(const float* __restrict__ im, int im_rows, int im_cols, int im_pitch,
                                                float ofsx, float ofsy, float a11, float a12, float a21, float a22,
                                                int rw, int rh, bool touch, int row, Out out /* out(i, value) */) {
  interpolate_seg(im, im_rows, im_cols, im_pitch, ofsx, ofsy, a11, a12, a21, a22, rw, rh, touch, row, 0, rw, out);
}

// Launch bookkeeping
#define MB2_LAUNCH_ON(ctx, strm, kernel, grid, block, smem, ...)           \
  do {                                                                     \
    mb2_ctx::ProfRec _pr{#kernel, nullptr, nullptr};                       \
    if ((ctx)->profiling) {                                                \
      cudaEventCreate(&_pr.a); cudaEventCreate(&_pr.b);                    \
      cudaEventRecord(_pr.a, (strm));                                      \
    }                                                                      \
    kernel<<<(grid), (block), (smem), (strm)>>>(__VA_ARGS__);              \
    { const cudaError_t _le = cudaPeekAtLastError();                       \
      if (_le != cudaSuccess && (ctx)->launch_error == cudaSuccess) {      \
        (ctx)->launch_error = _le; (ctx)->set_error(std::string("launch of " #kernel ": ") + cudaGetErrorString(_le)); } } \
    if ((ctx)->profiling) { cudaEventRecord(_pr.b, (strm)); (ctx)->prof.push_back(_pr); } \
    (ctx)->launches++;                                                     \
  } while (0)
#define MB2_LAUNCH(ctx, kernel, grid, block, smem, ...) MB2_LAUNCH_ON(ctx, (ctx)->stream, kernel, grid, block, smem, __VA_ARGS__)

// host-side helpers implemented in capi.cu
int mb2_stage_in(mb2_ctx* ctx, const void* src, size_t bytes, DevBuf& dst, const void** dev_ptr);
bool mb2_is_device_ptr(const void* p);
