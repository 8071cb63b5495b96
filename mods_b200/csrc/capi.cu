// C ABI of libmods_b200.so (include/mods_b200.h): host orchestration of the kernels.
// No oracle / CPU fallback anywhere on this path: every entry point either runs the CUDA kernels
// or fails with an error code.
#include "common.cuh"
#include "parallel_host.hpp"
#undef MB2_NS
#define MB2_NS mb2_capi_detail
#include "pyramid.cuh"
#include "describe.cuh"
#include "nn.cuh"
#include "ransac.cuh"
#include "host_tables.hpp"

#include <cub/cub.cuh>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>

namespace MB2_NS {

struct CtxTables {
  DevBuf smm_mask, orimask, desc_tables, tap_n, tap_off, tap_w, octaves;
  int smm_size = 0, tap_max_m = -1;
  std::vector<int> tap_n_host;
  double tap_mrsize_key = -1;
  int tap_patch = 0;
};
struct CtxPriv {
  CtxTables t;
  std::vector<OctaveLevels> last_octaves;  // geometry of the most recent pyramid (diagnostics)
  float* last_patches = nullptr;           // n x 41 x 41 normalised patches of the most recent describe
};
std::mutex g_lut_mutex;
bool g_lut_uploaded[64] = {false};

CtxPriv* priv(mb2_ctx* ctx);

int upload_lut(mb2_ctx* ctx) {
  std::lock_guard<std::mutex> lk(g_lut_mutex);
  if (ctx->device < 64 && g_lut_uploaded[ctx->device]) return MB2_OK;
  double lut[256];
  mb2host::build_atan_lut(lut);
  MB2_CUDA_CHECK(ctx, cudaMemcpyToSymbol(c_atan_lut, lut, sizeof lut));
  static float sift_o[mb2host::ATAN_CODES];
  static unsigned char ori_bin[mb2host::ATAN_CODES];
  mb2host::build_atan_derived(lut, sift_o, ori_bin);
  MB2_CUDA_CHECK(ctx, cudaMemcpyToSymbol(g_atan_sift_o, sift_o, sizeof sift_o));
  MB2_CUDA_CHECK(ctx, cudaMemcpyToSymbol(g_atan_ori_bin, ori_bin, sizeof ori_bin));
  if (ctx->device < 64) g_lut_uploaded[ctx->device] = true;
  return MB2_OK;
}

inline int pitch_of(int cols) { return (cols + 31) & ~31; }

// ---- cub helpers (bookkeeping only: ordering + stable compaction of short lists) ------------------
__global__ void k_iota_keys(const KeyOut* __restrict__ in, int n, unsigned long long* keys, int* idx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { keys[i] = in[i].order; idx[i] = i; }
}
__global__ void k_gather(const KeyOut* __restrict__ in, const int* __restrict__ idx, int n, KeyOut* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[idx[i]];
}
__global__ void k_flags(const KeyOut* __restrict__ in, int n, int* __restrict__ flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = in[i].keep ? 1 : 0;
}
__global__ void k_scatter(const KeyOut* __restrict__ in, const KeyOut* __restrict__ in2, const int* __restrict__ flags,
                          const int* __restrict__ pos, int n, KeyOut* __restrict__ out, KeyOut* __restrict__ out2, int* __restrict__ total) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (flags[i]) { out[pos[i]] = in[i]; if (in2) out2[pos[i]] = in2[i]; }
  if (i == n - 1) *total = pos[i] + flags[i];
}
__global__ void k_scan_offsets_total(const unsigned long long* need, const unsigned long long* off, int n, unsigned long long* total) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *total = n ? off[n - 1] + need[n - 1] : 0ull;
}
__global__ void k_kp_from_doubles(const double* __restrict__ kp, int n, KeyOut* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  KeyOut o;
  for (int j = 0; j < MB2_KP; j++) o.v[j] = kp[(size_t)i * MB2_KP + j];
  o.order = (unsigned long long)i; o.keep = 1; o.pad = 0;
  out[i] = o;
}
__global__ void k_kp_to_doubles(const KeyOut* __restrict__ in, int n, double* __restrict__ kp) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int j = 0; j < MB2_KP; j++) kp[(size_t)i * MB2_KP + j] = in[i].v[j];
}
__global__ void k_xy_from_keys(const KeyOut* __restrict__ in, int n, double* __restrict__ xy) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { xy[2 * i] = in[i].v[0]; xy[2 * i + 1] = in[i].v[1]; }
}
__global__ void k_match_flags(const int* __restrict__ accept, int n, int* flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = accept[i];
}
__global__ void k_match_scatter(const MatchRow* __restrict__ rows, const int* __restrict__ flags, const int* __restrict__ pos, int n,
                                double* __restrict__ out, int capacity, int* __restrict__ total) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (flags[i] && pos[i] < capacity) {
    const MatchRow r = rows[i];
    double* o = out + (size_t)pos[i] * 7;
    o[0] = r.q; o[1] = r.idx0; o[2] = r.idxJ; o[3] = r.idx1; o[4] = r.d0; o[5] = r.dJ; o[6] = r.d1;
  }
  if (i == n - 1) *total = pos[i] + flags[i];
}

#define LAUNCH1D(ctx, kernel, n, ...) \
  do { if ((n) > 0) MB2_LAUNCH(ctx, kernel, ((n) + 255) / 256, 256, 0, __VA_ARGS__); } while (0)

// sort `n` records by KeyOut::order (ascending) into out
int sort_by_order(mb2_ctx* ctx, const KeyOut* in, int n, KeyOut* out, DevBuf& tmp) {
  if (n <= 0) return MB2_OK;
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int*)nullptr,
                                  (int*)nullptr, n, 0, 64, ctx->stream);
  size_t need = (size_t)n * (8 + 8 + 4 + 4) + cub_bytes + 64;
  MB2_CUDA_CHECK(ctx, tmp.reserve(need));
  unsigned long long* k0 = tmp.as<unsigned long long>();
  unsigned long long* k1 = k0 + n;
  int* i0 = (int*)(k1 + n);
  int* i1 = i0 + n;
  void* cub_tmp = (void*)(((uintptr_t)(i1 + n) + 15) & ~(uintptr_t)15);
  LAUNCH1D(ctx, k_iota_keys, n, in, n, k0, i0);
  MB2_CUDA_CHECK(ctx, cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, k0, k1, i0, i1, n, 0, 64, ctx->stream));
  ctx->launches += 3;
  LAUNCH1D(ctx, k_gather, n, in, i1, n, out);
  return MB2_OK;
}

// stable compaction of records with keep != 0 (optionally a second array in lockstep).  *n_out on host.
int compact(mb2_ctx* ctx, const KeyOut* in, const KeyOut* in2, int n, KeyOut* out, KeyOut* out2, DevBuf& tmp, int* n_out) {
  *n_out = 0;
  if (n <= 0) return MB2_OK;
  size_t cub_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (int*)nullptr, (int*)nullptr, n, ctx->stream);
  size_t need = (size_t)n * 8 + 16 + cub_bytes + 64;
  MB2_CUDA_CHECK(ctx, tmp.reserve(need));
  int* flags = tmp.as<int>();
  int* pos = flags + n;
  int* total = pos + n;
  void* cub_tmp = (void*)(((uintptr_t)(total + 1) + 15) & ~(uintptr_t)15);
  LAUNCH1D(ctx, k_flags, n, in, n, flags);
  MB2_CUDA_CHECK(ctx, cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, flags, pos, n, ctx->stream));
  ctx->launches += 1;
  LAUNCH1D(ctx, k_scatter, n, in, in2, flags, pos, n, out, out2, total);
  MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(n_out, total, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return MB2_OK;
}

// ---- per-context tables -----------------------------------------------------------------------
std::vector<std::pair<mb2_ctx*, CtxPriv*>> g_privs;
std::mutex g_priv_mutex;
CtxPriv* priv(mb2_ctx* ctx) {
  std::lock_guard<std::mutex> lk(g_priv_mutex);
  for (auto& p : g_privs) if (p.first == ctx) return p.second;
  g_privs.emplace_back(ctx, new CtxPriv());
  return g_privs.back().second;
}
void drop_priv(mb2_ctx* ctx) {
  std::lock_guard<std::mutex> lk(g_priv_mutex);
  for (size_t i = 0; i < g_privs.size(); i++)
    if (g_privs[i].first == ctx) {
      CtxTables& t = g_privs[i].second->t;
      t.smm_mask.release(); t.orimask.release(); t.desc_tables.release(); t.tap_n.release(); t.tap_off.release(); t.tap_w.release();
      t.octaves.release();
      delete g_privs[i].second;
      g_privs.erase(g_privs.begin() + i);
      return;
    }
}

int ensure_smm_mask(mb2_ctx* ctx, int size) {
  CtxTables& t = priv(ctx)->t;
  if (t.smm_size == size) return MB2_OK;
  std::vector<float> m((size_t)size * size);
  mb2host::gauss_mask(m.data(), size);
  MB2_CUDA_CHECK(ctx, t.smm_mask.reserve(m.size() * 4));
  MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(t.smm_mask.p, m.data(), m.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  t.smm_size = size;
  return MB2_OK;
}
int ensure_orimask(mb2_ctx* ctx) {
  CtxTables& t = priv(ctx)->t;
  if (t.orimask.p) return MB2_OK;
  std::vector<float> m(41 * 41);
  mb2host::circular_gauss_mask(m.data(), 41, 41 / 3.0f);  // synth-detection.cpp:762
  MB2_CUDA_CHECK(ctx, t.orimask.reserve(m.size() * 4));
  MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(t.orimask.p, m.data(), m.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return MB2_OK;
}
int ensure_desc_tables(mb2_ctx* ctx) {
  CtxTables& t = priv(ctx)->t;
  if (t.desc_tables.p) return MB2_OK;
  DescTables* h = new DescTables();
  mb2host::circular_gauss_mask(h->mask, 41);
  mb2host::sift_bins(41, 4, 8, h->bin0, h->bin1, h->w0, h->w1);
  MB2_CUDA_CHECK(ctx, t.desc_tables.reserve(sizeof(DescTables)));
  cudaError_t e = cudaMemcpyAsync(t.desc_tables.p, h, sizeof(DescTables), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  delete h;
  MB2_CUDA_CHECK(ctx, e);
  return MB2_OK;
}
// taps for every m in [0, max_m]
int ensure_taps(mb2_ctx* ctx, int max_m, int patchSize, TapTable* out) {
  CtxTables& t = priv(ctx)->t;
  if (t.tap_max_m < max_m || t.tap_patch != patchSize) {
    int want = std::max(max_m, 256);
    std::vector<int> n(want + 1, 0), off(want + 1, 0);
    std::vector<float> w;
    for (int m = 0; m <= want; m++) {
      const int patchImageSize = 2 * m + 1;
      const float scale = float(patchImageSize) / float(patchSize);
      off[m] = (int)w.size();
      if (scale > 0.4) {
        const float sigma = 1.5f * scale;  // synth-detection.hpp:208
        const int ks = mb2host::gauss_ksize(sigma);
        std::vector<float> k = mb2host::gauss_kernel(ks, sigma);
        n[m] = ks;
        w.insert(w.end(), k.begin(), k.end());
      }
    }
    MB2_CUDA_CHECK(ctx, t.tap_n.reserve(n.size() * 4));
    MB2_CUDA_CHECK(ctx, t.tap_off.reserve(off.size() * 4));
    MB2_CUDA_CHECK(ctx, t.tap_w.reserve(w.size() * 4 + 4));
    MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(t.tap_n.p, n.data(), n.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(t.tap_off.p, off.data(), off.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(t.tap_w.p, w.data(), w.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    t.tap_max_m = want; t.tap_patch = patchSize; t.tap_n_host = n;
  }
  out->n = t.tap_n.as<int>(); out->off = t.tap_off.as<int>(); out->w = t.tap_w.as<float>(); out->max_m = t.tap_max_m;
  return MB2_OK;
}

// ---- staging -------------------------------------------------------------------------------
// alias_ok: the caller only READS the image during this call -- a device-resident image whose rows are already at the plane pitch is
// used where it lies (no 50 MB copy per synthesised view)
int stage_image(mb2_ctx* ctx, const float* pixels, int w, int h, ImgView* view, bool alias_ok = false) {
  const int pitch = pitch_of(w);
  if (alias_ok && pitch == w && ((uintptr_t)pixels & 127) == 0 && mb2_is_device_ptr(pixels)) {
    view->p = pixels; view->rows = h; view->cols = w; view->pitch = pitch;
    return MB2_OK;
  }
  MB2_CUDA_CHECK(ctx, ctx->img.reserve((size_t)pitch * h * 4));
  cudaMemcpyKind kind = mb2_is_device_ptr(pixels) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  MB2_CUDA_CHECK(ctx, cudaMemcpy2DAsync(ctx->img.p, (size_t)pitch * 4, pixels, (size_t)w * 4, (size_t)w * 4, h, kind, ctx->stream));
  view->p = ctx->img.as<float>(); view->rows = h; view->cols = w; view->pitch = pitch;
  return MB2_OK;
}

struct PyramidLayout {
  int n_octaves = 0;
  std::vector<int> rows, cols, pitch;
  std::vector<size_t> off;  // float offset of level 0 of each octave; levels are consecutive planes
  size_t total = 0;
};

// ---- detection core: pixels on device -> ordered KeyOut list on device --------------------------
// Result: ctx->kp_b holds n KeyOut (ordered like the reference's key vector, keep == 1 for all).
int detect_core(mb2_ctx* ctx, const ImgView& img, const mb2_hessaff_params& par, double tilt, double zoom, int as_regions, int* n_out) {
  *n_out = 0;
  if (par.mode < 0 || par.mode > 4) { ctx->set_error("hessaff: unknown DetectorMode"); return MB2_ERR_ARG; }
  if (par.detectorType < 0 || par.detectorType > 2) { ctx->set_error("scale-space detector: DET_HESSIAN (0), DET_DOG (1) and DET_HARRIS (2) are built"); return MB2_ERR_UNSUPPORTED; }
  if (par.numberOfScales + 2 > MB2_MAX_LEVELS || par.smmWindowSize != 19 || par.border < 2) {
    ctx->set_error("hessaff: unsupported numberOfScales / smmWindowSize / border"); return MB2_ERR_ARG;
  }
  int rc;
  if ((rc = upload_lut(ctx))) return rc;
  if ((rc = ensure_smm_mask(ctx, par.smmWindowSize))) return rc;
  const int S = par.numberOfScales, NL = S + 2;
  // octave geometry (pyramid.cpp:564-572, cv::resize output size = cvRound(dim * 0.5))
  PyramidLayout L;
  {
    int r = img.rows, c = img.cols;
    const int minSize = 2 * par.border + 2;
    while (r > minSize && c > minSize) {
      L.rows.push_back(r); L.cols.push_back(c); L.pitch.push_back(pitch_of(c));
      L.off.push_back(L.total);
      L.total += (size_t)L.pitch.back() * r * NL;
      r = mb2host::cv_round(r * 0.5); c = mb2host::cv_round(c * 0.5);
    }
    L.n_octaves = (int)L.rows.size();
  }
  if (L.n_octaves == 0) return MB2_OK;
  MB2_CUDA_CHECK(ctx, ctx->pyr.reserve(L.total * 4));
  MB2_CUDA_CHECK(ctx, ctx->resp.reserve(L.total * 4));
  float* pyr = ctx->pyr.as<float>(); float* resp = ctx->resp.as<float>();

  // sigmas exactly as pyramid.cpp:459-487 computes them in float
  const float sigmaStep = std::pow(2.0f, 1.0f / (float)S);
  std::vector<float> levelSigma(NL), incSigma(NL);
  {
    float curSigma = par.initialSigma;
    levelSigma[0] = curSigma;
    for (int i = 1; i < NL; i++) {
      incSigma[i] = curSigma * std::sqrt(sigmaStep * sigmaStep - 1.0f);
      levelSigma[i] = curSigma * sigmaStep;  // "sigma = curSigma*sigmaStep" used for the response norm
      curSigma *= sigmaStep;
    }
  }
  auto make_taps = [](float sigma) {
    BlurTaps t; t.n = mb2host::gauss_ksize(sigma);
    std::vector<float> k = mb2host::gauss_kernel(t.n, sigma);
    for (int i = 0; i < t.n && i < 33; i++) t.k[i] = k[i];
    return t;
  };
  std::vector<OctaveLevels> octs(L.n_octaves);
  for (int o = 0; o < L.n_octaves; o++) {
    octs[o].nlevels = NL;
    for (int l = 0; l < NL; l++) {
      size_t off = L.off[o] + (size_t)l * L.pitch[o] * L.rows[o];
      octs[o].blur[l] = ImgView{pyr + off, L.rows[o], L.cols[o], L.pitch[o]};
      octs[o].resp[l] = ImgView{resp + off, L.rows[o], L.cols[o], L.pitch[o]};
    }
  }
  CtxTables& T = priv(ctx)->t;
  priv(ctx)->last_octaves = octs;
  MB2_CUDA_CHECK(ctx, T.octaves.reserve(sizeof(OctaveLevels) * L.n_octaves));
  MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(T.octaves.p, octs.data(), sizeof(OctaveLevels) * L.n_octaves, cudaMemcpyHostToDevice, ctx->stream));

  // thresholds (pyramid.h:46-69)
  // every mode but FIXED_TH opens all response gates: every 3x3x3 extremum is localised and goes through Baumberg (pyramid.h:59-60)
  const bool harris = par.detectorType == 2;
  const bool dog = par.detectorType != 0;   // "the response comes from its own kernels after the blur" (DET_DOG and DET_HARRIS)
  const float finalThreshold = par.mode != 0 ? 0.0f : (dog ? par.threshold : par.threshold * par.threshold);   // squared for DET_HESSIAN only
  const float positiveThreshold = par.mode != 0 ? 0.0f : (float)(0.8 * par.threshold), negativeThreshold = -positiveThreshold;
  const double er = par.edgeEigenValueRatio;
  const double edgeScoreThreshold = (er + 1.0f) * (er + 1.0f) / er;

  // capacity of the per-octave candidate list and the global keypoint list
  const size_t px0 = (size_t)L.rows[0] * L.cols[0];
  const int cand_cap = (int)std::min<size_t>(px0 / 8 + 4096, (size_t)1 << 26);
  const int kp_cap = (int)std::min<size_t>(px0 / 16 + 4096, (size_t)1 << 25);
  MB2_CUDA_CHECK(ctx, ctx->cand.reserve((size_t)cand_cap * (2 * sizeof(Candidate) + sizeof(Localized)) + 64));
  Candidate* d_cand = ctx->cand.as<Candidate>();
  Localized* d_loc = (Localized*)(d_cand + cand_cap);
  Candidate* d_pre = (Candidate*)(d_loc + cand_cap);   // pixels that passed the in-level extremum test (k_blur_hess_tma)
  MB2_CUDA_CHECK(ctx, ctx->kp_a.reserve((size_t)kp_cap * sizeof(KeypointRec) + 64));
  KeypointRec* d_kp = ctx->kp_a.as<KeypointRec>();
  MB2_CUDA_CHECK(ctx, ctx->octmap.reserve(px0 * 8 + 64));
  unsigned long long* d_octmap = ctx->octmap.as<unsigned long long>();
  MB2_CUDA_CHECK(ctx, ctx->misc.reserve(4096));
  int* d_counts = ctx->misc.as<int>();  // [0] = candidates (per octave), [1] = keypoints (global)
  MB2_CUDA_CHECK(ctx, cudaMemsetAsync(d_counts, 0, 64, ctx->stream));

  // DET_DOG: the response of level l is level - GaussianBlur(level, sigma = levelSigma[l]^2) in every octave: taps of the NL wide
  // kernels on the device (ctx->dog_taps), one scratch plane for the row pass
  std::vector<int> dog_n(NL, 0), dog_off(NL, 0);
  if (harris) MB2_CUDA_CHECK(ctx, ctx->dog_tmp.reserve((size_t)L.pitch[0] * L.rows[0] * 4 * 6));   // gradient products and their blurs
  if (dog && !harris) {
    std::vector<float> all;
    for (int l = 0; l < NL; l++) {
      const float sg = levelSigma[l] * levelSigma[l];
      dog_n[l] = mb2host::gauss_ksize(sg); dog_off[l] = (int)all.size();
      const std::vector<float> k = mb2host::gauss_kernel(dog_n[l], sg);
      all.insert(all.end(), k.begin(), k.end());
    }
    MB2_CUDA_CHECK(ctx, ctx->dog_taps.reserve(all.size() * 4));
    MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dog_taps.p, all.data(), all.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));   // `all` is a host temporary
    MB2_CUDA_CHECK(ctx, ctx->dog_tmp.reserve((size_t)L.pitch[0] * L.rows[0] * 4));
  }
  int resp_rc = MB2_OK;   // first failure of a separate response pass
  auto dog_response = [&](const OctaveLevels& oc, int l) {
    if (harris) {   // HarrisResponse(level, norm = levelSigma^2): sigmasq = 0.6 norm, sigma = sqrt(sigmasq) (pyramid.cpp:287-288)
      const float norm = levelSigma[l] * levelSigma[l];
      const float sigmasq = 0.6 * norm;
      const int hrc = mb2_launch_harris(ctx, oc.blur[l], (float*)oc.resp[l].p, oc.resp[l].pitch, make_taps(std::sqrt(sigmasq)), sigmasq, ctx->dog_tmp.as<float>());
      if (hrc && !resp_rc) resp_rc = hrc;
      return;
    }
    mb2_launch_dog(ctx, oc.blur[l], (float*)oc.resp[l].p, oc.resp[l].pitch, ctx->dog_taps.as<float>() + dog_off[l], dog_n[l], ctx->dog_tmp.as<float>());
  };

  // ---- scale space -----------------------------------------------------------------------------
  float pixelDistance = 1.0f;
  for (int o = 0; o < L.n_octaves; o++) {
    const OctaveLevels& oc = octs[o];
    if (o == 0) {
      // first level: initial blur sqrt(initialSigma^2 - 0.5^2) (pyramid.cpp:555-562) fused with its response
      const float curSigma0 = 0.5f;
      float n0 = levelSigma[0] * levelSigma[0];
      if (par.initialSigma > curSigma0) {
        const float sigma = std::sqrt(par.initialSigma * par.initialSigma - curSigma0 * curSigma0);
        if (dog) rc = mb2_launch_blur(ctx, img, (float*)oc.blur[0].p, (float*)oc.resp[0].p, oc.blur[0].pitch, make_taps(sigma), n0 * n0, 0);
        else rc = mb2_launch_blur_tma(ctx, img, (float*)oc.blur[0].p, (float*)oc.resp[0].p, oc.blur[0].pitch, make_taps(sigma), n0 * n0, PrefilterArgs{0, 0, 0, 0, 0, nullptr, nullptr, 0});
        if (rc) return rc;
        if (dog) dog_response(oc, 0);
      } else {
        MB2_CUDA_CHECK(ctx, cudaMemcpy2DAsync((void*)oc.blur[0].p, (size_t)oc.blur[0].pitch * 4, img.p, (size_t)img.pitch * 4,
                                              (size_t)img.cols * 4, img.rows, cudaMemcpyDeviceToDevice, ctx->stream));
        if (dog) dog_response(oc, 0); else mb2_launch_hessian(ctx, oc.blur[0], (float*)oc.resp[0].p, oc.resp[0].pitch, n0 * n0);
      }
    } else {
      // next octave = cv::resize(level S, 0.5) of the previous one (pyramid.cpp:517-520)
      mb2_launch_resize_half(ctx, octs[o - 1].blur[S], (float*)oc.blur[0].p, oc.blur[0].rows, oc.blur[0].cols, oc.blur[0].pitch);
      float n0 = levelSigma[0] * levelSigma[0];
      if (dog) dog_response(oc, 0); else mb2_launch_hessian(ctx, oc.blur[0], (float*)oc.resp[0].p, oc.resp[0].pitch, n0 * n0);
    }
    MB2_CUDA_CHECK(ctx, cudaMemsetAsync(d_counts, 0, sizeof(int), ctx->stream));
    MB2_CUDA_CHECK(ctx, cudaMemsetAsync(d_counts + 2, 0, sizeof(int), ctx->stream));
    for (int i = 1; i < NL; i++) {
      float nrm = levelSigma[i] * levelSigma[i];  // Response(nextBlur, sigma*sigma); HessianResponse squares it again
      if (dog) rc = mb2_launch_blur(ctx, oc.blur[i - 1], (float*)oc.blur[i].p, (float*)oc.resp[i].p, oc.blur[i].pitch, make_taps(incSigma[i]), nrm * nrm, 0);
      else   // detection levels 1..S: the kernel that produces the level also lists its in-level extrema beyond the gate
        rc = mb2_launch_blur_tma(ctx, oc.blur[i - 1], (float*)oc.blur[i].p, (float*)oc.resp[i].p, oc.blur[i].pitch, make_taps(incSigma[i]), nrm * nrm,
                                 PrefilterArgs{i <= S ? 1 : 0, par.border, positiveThreshold, negativeThreshold, i, d_pre, d_counts + 2, cand_cap});
      if (rc) return rc;
      if (dog) dog_response(oc, i);
    }
    // ---- extrema -> localisation -> de-duplication, levels 1..S in the reference's order --------
    if (dog) {
      for (int lv = 1; lv <= S; lv++)
        mb2_launch_nms(ctx, oc.resp[lv - 1], oc.resp[lv], oc.resp[lv + 1], par.border, positiveThreshold, negativeThreshold, lv, d_cand,
                       d_counts, cand_cap);
    } else if (oc.blur[0].cols - 2 * par.border > 0 && oc.blur[0].rows - 2 * par.border > 0)
      mb2_launch_nms_finish(ctx, oc, d_pre, d_counts + 2, cand_cap, d_cand, d_counts, cand_cap);
    int n_cand = 0;
    MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(&n_cand, d_counts, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_cand > cand_cap) { ctx->set_error("hessaff: candidate list overflow"); return MB2_ERR_CAPACITY; }
    if (n_cand > 0) {
      LocalizeParams lp;
      lp.edgeScoreThreshold = edgeScoreThreshold; lp.finalThreshold = finalThreshold; lp.pixelDistance = pixelDistance;
      lp.numberOfScales = S; lp.detectorType = par.detectorType;
      // findLevelKeypoints(curSigma): curSigma at level lv is initialSigma * sigmaStep^lv accumulated in float
      {
        float cs = par.initialSigma;
        for (int l = 0; l < NL; l++) { lp.levelSigma[l] = cs; cs *= sigmaStep; }
      }
      mb2_launch_fill_u64(ctx, d_octmap, (size_t)L.rows[o] * L.cols[o], ~0ull);
      mb2_launch_localize(ctx, oc, d_cand, n_cand, lp, d_octmap, d_loc);
      mb2_launch_emit(ctx, oc, d_loc, n_cand, d_octmap, lp, o, d_kp, d_counts + 1, kp_cap);
    }
    pixelDistance *= 2.0;
  }
  if (resp_rc) return resp_rc;
  int n_kp = 0;
  MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(&n_kp, d_counts + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  if (n_kp > kp_cap) { ctx->set_error("hessaff: keypoint list overflow"); return MB2_ERR_CAPACITY; }
  if (n_kp == 0) return MB2_OK;

  // ---- host leg: keypoint scale through the host libm (pyramid.cpp:421) ----------------------------
  {
    MB2_CUDA_CHECK(ctx, ctx->rs_a.reserve((size_t)n_kp * (sizeof(ScaleReq) + 4) + 64));
    ScaleReq* d_req = ctx->rs_a.as<ScaleReq>();
    float* d_s = (float*)(d_req + n_kp);
    MB2_CUDA_CHECK(ctx, ctx->h_a.reserve((size_t)n_kp * sizeof(ScaleReq)));
    MB2_CUDA_CHECK(ctx, ctx->h_b.reserve((size_t)n_kp * 4));
    mb2_launch_scale_requests(ctx, d_kp, n_kp, d_req);
    MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->h_a.p, d_req, (size_t)n_kp * sizeof(ScaleReq), cudaMemcpyDeviceToHost, ctx->stream));
    MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    const ScaleReq* req = ctx->h_a.as<ScaleReq>();
    float* hs = ctx->h_b.as<float>();
    float lvlSigma[MB2_MAX_LEVELS];
    { float cs = par.initialSigma; for (int l = 0; l < NL; l++) { lvlSigma[l] = cs; cs *= sigmaStep; } }
    const int nscales = par.numberOfScales;
    mb2par::parallel_chunks((n_kp + 4095) / 4096, [&](int ck) {
      for (int i = ck * 4096, e = std::min(n_kp, i + 4096); i < e; i++) {
        const float scale = lvlSigma[req[i].level] * std::pow(2.0f, req[i].b2 / nscales);
        float pd = 1.0f;
        for (int o = 0; o < req[i].octave; o++) pd *= 2.0f;
        hs[i] = pd * scale;
      }
    });
    MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(d_s, hs, (size_t)n_kp * 4, cudaMemcpyHostToDevice, ctx->stream));
    mb2_launch_set_scales(ctx, d_kp, n_kp, d_s);
  }

  // ---- Baumberg, export, ordering ---------------------------------------------------------------
  AffineParams ap;
  ap.maxIterations = par.maxIterations; ap.smmWindowSize = par.smmWindowSize; ap.doBaumberg = par.doBaumberg;
  ap.convergenceThreshold = par.convergenceThreshold; ap.initialSigma = par.initialSigma;
  mb2_launch_affine_shape(ctx, T.octaves.as<OctaveLevels>(), L.n_octaves, d_kp, n_kp, ap, T.smm_mask.as<float>());
  MB2_CUDA_CHECK(ctx, ctx->kp_b.reserve((size_t)n_kp * sizeof(KeyOut)));
  MB2_CUDA_CHECK(ctx, ctx->kp_c.reserve((size_t)n_kp * sizeof(KeyOut)));
  mb2_launch_export(ctx, d_kp, n_kp, ctx->kp_c.as<KeyOut>(), as_regions);
  if ((rc = sort_by_order(ctx, ctx->kp_c.as<KeyOut>(), n_kp, ctx->kp_b.as<KeyOut>(), ctx->misc))) return rc;
  int n_keep = 0;
  if ((rc = compact(ctx, ctx->kp_b.as<KeyOut>(), nullptr, n_kp, ctx->kp_c.as<KeyOut>(), nullptr, ctx->misc, &n_keep))) return rc;
  std::swap(ctx->kp_b, ctx->kp_c);
  *n_out = n_keep;
  if (par.mode != 0 && n_keep > 0) {
    // prepareKeysForExport (scale-space-detector.hpp:127-198): sort by |response| with the reference's own (unstable) std::sort and
    // truncate.  Host leg: only the same library routine on the same input order reproduces the order of equal responses.
    int reg_number = par.reg_number;
    if ((tilt > 2.0) || (zoom < 0.5)) reg_number = (int)std::floor(zoom * (double)reg_number / tilt);   // scale-space-detector.cpp:50-51
    std::vector<KeyOut> keys(n_keep);
    MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(keys.data(), ctx->kp_b.p, (size_t)n_keep * sizeof(KeyOut), cudaMemcpyDeviceToHost, ctx->stream));
    MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    auto cmp = [](const KeyOut& a, const KeyOut& b) { return std::fabs(a.v[7]) > std::fabs(b.v[7]); };
    std::sort(keys.begin(), keys.end(), cmp);
    const int regNumber = n_keep;
    KeyOut probe = keys[0];
    switch (par.mode) {
      case 1: probe.v[7] = std::fabs(keys[0].v[7]) * par.rel_threshold; keys.resize(std::lower_bound(keys.begin(), keys.end(), probe, cmp) - keys.begin()); break;
      case 2: {
        int n = reg_number;
        if (par.doBaumberg) n = (int)std::floor(3.0 * (double)n);
        if (n < regNumber && n >= 0) keys.resize(n);
        break; }
      case 3: keys.resize((int)std::floor(par.rel_reg_number * (double)keys.size())); break;
      case 4: {
        probe.v[7] = par.threshold;
        const int fix = (int)(std::lower_bound(keys.begin(), keys.end(), probe, cmp) - keys.begin());
        keys.resize(fix < reg_number ? std::min(reg_number, regNumber) : std::min(fix, regNumber));
        break; }
      default: break;
    }
    if (par.mode == 2 && (int)keys.size() > reg_number) keys.resize(std::max(reg_number, 0));
    for (size_t i = 0; i < keys.size(); i++) keys[i].order = i;
    *n_out = (int)keys.size();
    if (*n_out > 0) {
      MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->kp_b.p, keys.data(), keys.size() * sizeof(KeyOut), cudaMemcpyHostToDevice, ctx->stream));
      MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    }
  }
  return MB2_OK;
}

int download_keys(mb2_ctx* ctx, const KeyOut* d_keys, int n, double* out, int capacity, DevBuf& tmp) {
  const int m = std::min(n, capacity);
  if (m <= 0) return MB2_OK;
  MB2_CUDA_CHECK(ctx, tmp.reserve((size_t)m * MB2_KP * 8));
  LAUNCH1D(ctx, k_kp_to_doubles, m, d_keys, m, tmp.as<double>());
  MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(out, tmp.p, (size_t)m * MB2_KP * 8, cudaMemcpyDeviceToHost, ctx->stream));
  return MB2_OK;
}

// orientation on ctx->kp_b (n records) -> ctx->kp_c (n_out records), ordered
int orient_core(mb2_ctx* ctx, const ImgView& img, int n, const mb2_orientation_params& op, int* n_out) {
  *n_out = 0;
  if (n <= 0) return MB2_OK;
  if (op.patchSize != 41) { ctx->set_error("orientation: patchSize must be 41"); return MB2_ERR_ARG; }
  int rc;
  if ((rc = upload_lut(ctx))) return rc;
  if ((rc = ensure_orimask(ctx))) return rc;
  const int maxA = std::max(op.maxAngles, 0);
  if (maxA == 0) return MB2_OK;
  if (maxA > 36) { ctx->set_error("orientation: maxAngles > 36"); return MB2_ERR_ARG; }
  const size_t slots = (size_t)n * maxA;
  MB2_CUDA_CHECK(ctx, ctx->kp_a.reserve(slots * sizeof(KeyOut) + (size_t)n * 4 + 64));

  KeyOut* d_slots = ctx->kp_a.as<KeyOut>();
  int* d_cnt = (int*)(d_slots + slots);
  OrientParams p{op.mrSize, op.patchSize, maxA, op.threshold, op.doHalfSIFT != 0 ? 1 : 0};
  mb2_launch_orientation(ctx, img, ctx->kp_b.as<KeyOut>(), n, p, priv(ctx)->t.orimask.as<float>(), d_slots, d_cnt);
  MB2_CUDA_CHECK(ctx, ctx->kp_c.reserve(slots * sizeof(KeyOut)));
  if ((rc = compact(ctx, d_slots, nullptr, (int)slots, ctx->kp_c.as<KeyOut>(), nullptr, ctx->misc, n_out))) return rc;
  const int m = *n_out;
  if (m > 0) {
    // host leg: ci = cos(-angle), si = sin(-angle) in float through the host libm, exactly as
    // DetectOrientation does (synth-detection.cpp:902-903); the device finishes A <- A R.
    MB2_CUDA_CHECK(ctx, ctx->rs_a.reserve((size_t)m * (4 + 16) + 64));
    float* d_ang = ctx->rs_a.as<float>();
    double* d_cs = (double*)(((uintptr_t)(d_ang + m) + 15) & ~(uintptr_t)15);
    MB2_CUDA_CHECK(ctx, ctx->h_a.reserve((size_t)m * 4));
    MB2_CUDA_CHECK(ctx, ctx->h_b.reserve((size_t)m * 16));
    mb2_launch_extract_angles(ctx, ctx->kp_c.as<KeyOut>(), m, d_ang);
    MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->h_a.p, d_ang, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->stream));
    MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    const float* ang = ctx->h_a.as<float>();
    double* cs = ctx->h_b.as<double>();
    mb2par::parallel_chunks((m + 4095) / 4096, [&](int ck) {
      for (int i = ck * 4096, e = std::min(m, i + 4096); i < e; i++) { cs[2 * i] = std::cos(-ang[i]); cs[2 * i + 1] = std::sin(-ang[i]); }
    });
    MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(d_cs, cs, (size_t)m * 16, cudaMemcpyHostToDevice, ctx->stream));
    mb2_launch_apply_rotation(ctx, ctx->kp_c.as<KeyOut>(), d_cs, m);
  }
  return MB2_OK;
}

// describe the n records at d_keys; descriptors -> ctx->desc_u8 (device)
int describe_core_one(mb2_ctx* ctx, const ImgView& img, const KeyOut* d_keys, int n, const mb2_sift_params& sp, float* d_dsp_acc, int dsp_first);
// DescribeRegions for one descriptor; DSPSIFT (sp.dspScales > 0) = the raw plain-SIFT votes at dspScales + 1 measurement-region sizes,
// summed in float, then the float SIFTnorm (imagerepresentation.cpp:1547-1598)
int describe_core(mb2_ctx* ctx, const ImgView& img, const KeyOut* d_keys, int n, const mb2_sift_params& sp) {
  if (n <= 0) return MB2_OK;
  if (sp.dspScales <= 0) return describe_core_one(ctx, img, d_keys, n, sp, nullptr, 0);
  MB2_CUDA_CHECK(ctx, ctx->rs_u.reserve((size_t)n * 128 * 4));
  float* acc = ctx->rs_u.as<float>();
  for (int i = 0; i < sp.dspScales + 1; i++) {
    mb2_sift_params one = sp;
    one.rootSIFT = 0; one.doHalfSIFT = 0;
    one.mrSize = sp.mrSize * (sp.dspStartCoef + i * (sp.dspEndCoef - sp.dspStartCoef) / sp.dspScales);
    const int rc = describe_core_one(ctx, img, d_keys, n, one, acc, i == 0);
    if (rc) return rc;
  }
  mb2_launch_dsp_norm(ctx, acc, n, ctx->desc_u8.as<uint8_t>());
  return MB2_OK;
}
int describe_core_one(mb2_ctx* ctx, const ImgView& img, const KeyOut* d_keys, int n, const mb2_sift_params& sp, float* d_dsp_acc, int dsp_first) {
  if (n <= 0) return MB2_OK;
  if (sp.patchSize != 41) { ctx->set_error("describe: patchSize must be 41"); return MB2_ERR_ARG; }
  int rc;
  if ((rc = upload_lut(ctx))) return rc;
  if ((rc = ensure_desc_tables(ctx))) return rc;
  if (sp.doHalfSIFT && !sp.rootSIFT) {
    ctx->set_error("describe: HalfSIFT without RootSIFT is undefined in the reference (SIFTnorm reads past the 64-entry vector)");
    return MB2_ERR_UNSUPPORTED;
  }
  DescribeParams dp{sp.mrSize, sp.patchSize, sp.photoNorm, sp.rootSIFT, sp.fastPatchExtraction, sp.doHalfSIFT != 0 ? 1 : 0, d_dsp_acc ? 1 : 0};
  // largest possible m: the region must fit in the image for the earlier boundary tests, bound by the image diagonal
  const int max_m = (int)std::ceil(std::sqrt((double)img.rows * img.rows + (double)img.cols * img.cols)) + 8;
  TapTable taps;
  if ((rc = ensure_taps(ctx, max_m, sp.patchSize, &taps))) return rc;
  // plan: need[i] scratch floats per region (exclusive scan -> offsets) and the size-class lists (one radix sort)
  size_t scan_bytes = 0, sort_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr, n, ctx->stream);
  cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr, n, 0, 64, ctx->stream);
  const size_t cub_bytes = std::max(scan_bytes, sort_bytes);
  MB2_CUDA_CHECK(ctx, ctx->rs_a.reserve((size_t)n * 32 + 128 + cub_bytes + 64));
  unsigned long long* d_need = ctx->rs_a.as<unsigned long long>();
  unsigned long long* d_off = d_need + n;
  unsigned long long* d_skeys = d_off + n;        // unsorted, then sorted class keys
  unsigned long long* d_sorted = d_skeys + n;
  unsigned long long* d_total = d_sorted + n;     // [0] scratch floats, [1] sum of P2^2
  int* d_toobig = (int*)(d_total + 2);            // [0] too big, [1] pad, [2..7] class counts
  int* d_cls = d_toobig + 2;
  void* cub_tmp = (void*)(((uintptr_t)(d_cls + MB2_N_CLASSES) + 15) & ~(uintptr_t)15);
  MB2_CUDA_CHECK(ctx, cudaMemsetAsync(d_total, 0, 16 + 8 + 4 * MB2_N_CLASSES, ctx->stream));
  mb2_describe_plan(ctx, d_keys, n, dp, taps.max_m, d_need, d_toobig, d_total + 1, d_skeys, d_cls);
  MB2_CUDA_CHECK(ctx, cub::DeviceScan::ExclusiveSum(cub_tmp, scan_bytes, d_need, d_off, n, ctx->stream));
  MB2_LAUNCH(ctx, k_scan_offsets_total, 1, 32, 0, d_need, d_off, n, d_total);
  MB2_CUDA_CHECK(ctx, cub::DeviceRadixSort::SortKeys(cub_tmp, sort_bytes, d_skeys, d_sorted, n, 0, 64, ctx->stream));
  ctx->launches += 4;
  unsigned long long total = 0, tot2[2] = {0, 0}; int hostc[2 + MB2_N_CLASSES] = {0};
  MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(tot2, d_total, 16, cudaMemcpyDeviceToHost, ctx->stream));
  MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(hostc, d_toobig, sizeof hostc, cudaMemcpyDeviceToHost, ctx->stream));
  MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  const int toobig = hostc[0];
  ExtractPlan plan;
  plan.sorted = d_sorted;
  for (int c = 0; c < MB2_N_CLASSES; c++) plan.count[c] = hostc[2 + c];
  total = tot2[0];
  // SURVEY 8d gather model: 4 bilinear taps x 4 B per sampled pixel ((P+2)^2 + 41^2 per region) + 128 B out
  if (ctx->profiling) ctx->prof_extract_bytes += tot2[1] * 16ull + (unsigned long long)n * (1681ull * 16ull + 128ull);
  if (toobig) { ctx->set_error("describe: region larger than the tap table (m=" + std::to_string(toobig) + ")"); return MB2_ERR_CAPACITY; }
  if (total > ((size_t)24 << 30) / 4) { ctx->set_error("describe: patch scratch would exceed 24 GiB"); return MB2_ERR_CAPACITY; }
  // scratch = per-region sampling buffers, then n normalised 41x41 patches, photometric stats, bin-major votes
  const size_t f_patches = ((size_t)total + 31) & ~(size_t)31;
  const size_t f_stats = (f_patches + (size_t)n * 41 * 41 + 1) & ~(size_t)1;
  const size_t f_vec = (f_stats + (size_t)n * 2 + 1) & ~(size_t)1;
  const size_t f_rec = f_vec + (size_t)n * 128 * 2;
  MB2_CUDA_CHECK(ctx, ctx->patch_scratch.reserve((f_rec + (size_t)n * 41 * 41 * 2) * 4 + 64));
  MB2_CUDA_CHECK(ctx, ctx->desc_u8.reserve((size_t)n * 128));
  float* base = ctx->patch_scratch.as<float>();
  priv(ctx)->last_patches = base + f_patches;
  rc = mb2_launch_describe_kernel(ctx, img, d_keys, n, dp, priv(ctx)->t.desc_tables.as<DescTables>(), taps, priv(ctx)->t.tap_n_host.data(), plan, d_off, base,
                                  ctx->desc_u8.as<uint8_t>(), base + f_patches, (float2*)(base + f_stats), (double*)(base + f_vec),
                                  (float2*)(base + f_rec));
  if (rc == MB2_OK && d_dsp_acc) mb2_launch_dsp_accumulate(ctx, (const double*)(base + f_vec), n, d_dsp_acc, dsp_first);
  return rc;
}

// ordered compaction of the accepted rows of a matcher; out rows on the host
int compact_rows(mb2_ctx* ctx, const MatchRow* rows, const int* accept, int nq, double* out, int capacity) {
  size_t cub_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (int*)nullptr, (int*)nullptr, nq, ctx->stream);
  const int cap = std::min(capacity, nq);
  MB2_CUDA_CHECK(ctx, ctx->nn_d.reserve((size_t)nq * 4 + 16 + cub_bytes + 64 + (size_t)std::max(cap, 1) * 7 * 8));
  int* pos = ctx->nn_d.as<int>();
  int* total = pos + nq;
  void* cub_tmp = (void*)(((uintptr_t)(total + 1) + 15) & ~(uintptr_t)15);
  double* d_out = (double*)(((uintptr_t)((uint8_t*)cub_tmp + cub_bytes) + 15) & ~(uintptr_t)15);
  MB2_CUDA_CHECK(ctx, cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, accept, pos, nq, ctx->stream));
  ctx->launches += 1;
  LAUNCH1D(ctx, k_match_scatter, nq, rows, accept, pos, nq, d_out, cap, total);
  int n_match = 0;
  MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(&n_match, total, 4, cudaMemcpyDeviceToHost, ctx->stream));
  MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  const int m = std::min(n_match, cap);
  if (m > 0) {
    MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(out, d_out, (size_t)m * 7 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  }
  if (n_match > capacity) { ctx->set_error("match: output capacity too small"); return MB2_ERR_CAPACITY; }
  return n_match;
}

// FGINN on device-resident data.  out rows on host.
int match_core(mb2_ctx* ctx, const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, const double* d_txy, double matchRatio,
               double contradDist, int nn, double* out, int capacity) {
  if (nq <= 0 || nt <= 0) return 0;
  if (!(matchRatio > 0)) { ctx->set_error("match: matchRatio must be > 0"); return MB2_ERR_ARG; }
  const double sqminratio = matchRatio * matchRatio, contr2 = contradDist * contradDist;
  const bool all_points = sqminratio >= 1.0;   // matching.cpp:397-428
  if (all_points && (nn < 2 || nn > 64)) { ctx->set_error("match: matchRatio >= 1 needs 2 <= nn <= 64"); return MB2_ERR_ARG; }
  const int nq_pad = (nq + 255) & ~255, nt_pad = (nt + 255) & ~255;
  MB2_CUDA_CHECK(ctx, ctx->nn_a.reserve((size_t)nq_pad * 128 * 2 + (size_t)nq_pad * 4));
  MB2_CUDA_CHECK(ctx, ctx->nn_b.reserve((size_t)nt_pad * 128 * 2 + (size_t)nt_pad * 4));
  void* q_bf16 = ctx->nn_a.p; float* qn = (float*)((uint8_t*)q_bf16 + (size_t)nq_pad * 256);
  void* t_bf16 = ctx->nn_b.p; float* tn = (float*)((uint8_t*)t_bf16 + (size_t)nt_pad * 256);
  mb2_nn_prepare(ctx, d_q, nq, nq_pad, q_bf16, qn, 0.f);
  mb2_nn_prepare(ctx, d_t, nt, nt_pad, t_bf16, tn, 3.0e38f);
  // per-query state
  const size_t per_q = 8 * 3 + 4 * 6;
  MB2_CUDA_CHECK(ctx, ctx->nn_c.reserve((size_t)nq * (per_q + sizeof(MatchRow) + 16) + 256));
  uint8_t* base = ctx->nn_c.as<uint8_t>();
  NNState st;
  st.best0 = (unsigned long long*)base; st.best1 = st.best0 + nq; st.bestP = st.best1 + nq;
  st.cnt = (int*)(st.bestP + nq); st.incons = st.cnt + nq; st.idx0 = st.incons + nq;
  st.d0 = (float*)(st.idx0 + nq); st.thr = st.d0 + nq; st.thr_rel = st.thr + nq;
  MatchRow* rows = (MatchRow*)(((uintptr_t)(st.thr_rel + nq) + 15) & ~(uintptr_t)15);
  int* accept = (int*)(rows + nq);
  mb2_nn_init_state(ctx, st, nq);
  const char* impl = std::getenv("MB2_NN_IMPL");
  const bool simt = impl && std::strcmp(impl, "simt") == 0;
  int rc;
  if (all_points) mb2_nn_topk(ctx, d_q, nq, d_t, nt, qn, tn, d_txy, contr2, nn, rows, accept);
  else {
    for (int pass = 1; pass <= 2; pass++) {
      if (simt) mb2_nn_pass_simt(ctx, pass, d_q, nq, d_t, nt, qn, tn, st, d_txy, contr2);
      else if ((rc = mb2_nn_pass_tc(ctx, pass, q_bf16, nq, nq_pad, t_bf16, nt_pad, qn, tn, st, d_txy, contr2, std::min(nn, nt) - 1))) return rc;
      // the tcgen05 passes record a minimum with its 32-column block: recover the train index of the one winning block per query
      mb2_nn_resolve(ctx, pass == 1 ? st.best0 : st.bestP, nq, d_q, d_t, nt, qn, tn);
      if (pass == 1) mb2_nn_threshold(ctx, st, nq, qn, sqminratio);
    }
    mb2_nn_finalize(ctx, st, nq, nt, nn, rows, accept);
  }
  return compact_rows(ctx, rows, accept, nq, out, capacity);
}

// MatchFLANNDistance (matching.cpp:607-666) on device-resident byte descriptors
int match_hamming_core(mb2_ctx* ctx, const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, int desc_bytes, double matchDistanceThreshold,
                       double* out, int capacity) {
  if (nq <= 0 || nt <= 0) return 0;
  if (desc_bytes < 1 || desc_bytes > 64) { ctx->set_error("hamming match: descriptors of 1..64 bytes"); return MB2_ERR_UNSUPPORTED; }
  if (nt < 2) { ctx->set_error("hamming match: fewer than 2 trains (knnSearch with knn = 2 is undefined in the reference)"); return MB2_ERR_ARG; }
  const int W = desc_bytes <= 16 ? 4 : desc_bytes <= 32 ? 8 : 16;
  const int max_distance = (int)float(matchDistanceThreshold);   // matching.cpp:609
  const int n_chunks = mb2_nn_hamming_chunks(ctx, nq, nt);
  MB2_CUDA_CHECK(ctx, ctx->nn_a.reserve((size_t)nq * W * 4));
  MB2_CUDA_CHECK(ctx, ctx->nn_b.reserve((size_t)nt * W * 4));
  MB2_CUDA_CHECK(ctx, ctx->nn_c.reserve((size_t)nq * n_chunks * 16 + (size_t)nq * (sizeof(MatchRow) + 4) + 64));
  uint32_t *qw = ctx->nn_a.as<uint32_t>(), *tw = ctx->nn_b.as<uint32_t>();
  unsigned long long* part = ctx->nn_c.as<unsigned long long>();
  MatchRow* rows = (MatchRow*)(part + (size_t)nq * n_chunks * 2);
  int* accept = (int*)(rows + nq);
  mb2_nn_hamming_words(ctx, d_q, nq, desc_bytes, W, qw);
  mb2_nn_hamming_words(ctx, d_t, nt, desc_bytes, W, tw);
  mb2_nn_hamming(ctx, qw, nq, tw, nt, W, n_chunks, max_distance, part, rows, accept);
  return compact_rows(ctx, rows, accept, nq, out, capacity);
}

}  // namespace
using namespace MB2_NS;

bool mb2_is_device_ptr(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int mb2_stage_in(mb2_ctx* ctx, const void* src, size_t bytes, DevBuf& dst, const void** dev_ptr) {
  if (mb2_is_device_ptr(src)) { *dev_ptr = src; return MB2_OK; }
  MB2_CUDA_CHECK(ctx, dst.reserve(bytes ? bytes : 1));
  if (bytes) MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(dst.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  *dev_ptr = dst.p;
  return MB2_OK;
}

// =================================================================================================
extern "C" {

int mb2_ctx_create(int device, mb2_ctx** out) { return mb2_ctx_create_prio(device, 0, out); }
int mb2_ctx_create_prio(int device, int high_priority, mb2_ctx** out) {
  if (!out) return MB2_ERR_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) { cudaGetLastError(); return MB2_ERR_CUDA; }
  if (cudaSetDevice(device) != cudaSuccess) return MB2_ERR_CUDA;
  // MB2_SCHED=yield|block: how host threads wait in cudaStreamSynchronize.  A pair keeps ~5 host threads waiting on the GPU; with one
  // rank per GPU and few host cores per rank the default spin-wait oversubscribes the cores.
  if (const char* sched = getenv("MB2_SCHED")) {
    const unsigned f = !strcmp(sched, "yield") ? cudaDeviceScheduleYield : !strcmp(sched, "block") ? cudaDeviceScheduleBlockingSync : cudaDeviceScheduleAuto;
    if (cudaSetDeviceFlags(f) != cudaSuccess) cudaGetLastError();   // best effort
  }
  mb2_ctx* c = new mb2_ctx();
  c->device = device;
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);   // numerically lower = higher priority
  const int prio = high_priority ? prio_hi : prio_lo;
  if (cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio) != cudaSuccess) { delete c; return MB2_ERR_CUDA; }
  if (cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, prio) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_tree, cudaEventDisableTiming) != cudaSuccess) { mb2_ctx_destroy(c); return MB2_ERR_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->num_sms = prop.multiProcessorCount;
  *out = c;
  return MB2_OK;
}

void mb2_ctx_destroy(mb2_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ctx->release_buffers();
  mb2_mser_release(ctx);
  drop_priv(ctx);
  if (ctx->side) { cudaStreamSynchronize(ctx->side); cudaStreamDestroy(ctx->side); }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->ev_tree) cudaEventDestroy(ctx->ev_tree);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* mb2_last_error(const mb2_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int mb2_ctx_sync(mb2_ctx* ctx) {
  MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->launch_error != cudaSuccess) { ctx->launch_error = cudaSuccess; return MB2_ERR_CUDA; }   // a kernel launch failed since the last sync (message in mb2_last_error)
  return MB2_OK;
}
void* mb2_ctx_stream(mb2_ctx* ctx) { return (void*)ctx->stream; }
int mb2_ctx_device(const mb2_ctx* ctx) { return ctx ? ctx->device : -1; }
int mb2_ctx_profiling(const mb2_ctx* ctx) { return ctx && ctx->profiling ? 1 : 0; }

int mb2_slot_move(mb2_ctx* dst, int dst_slot, mb2_ctx* src, int src_slot) {
  if (!dst || !src || dst_slot < 0 || dst_slot >= MB2_MAX_SLOTS || src_slot < 0 || src_slot >= MB2_MAX_SLOTS || dst->device != src->device)
    return MB2_ERR_ARG;
  if (dst == src && dst_slot == src_slot) return MB2_OK;
  cudaSetDevice(dst->device);
  MB2_CUDA_CHECK(src, cudaStreamSynchronize(src->stream));
  MB2_CUDA_CHECK(dst, cudaStreamSynchronize(dst->stream));
  RegionSlot& d = dst->slots[dst_slot];
  RegionSlot& s = src->slots[src_slot];
  std::swap(d.desc, s.desc); std::swap(d.xy, s.xy); std::swap(d.n, s.n);   // buffers change owner, nothing is copied
  s.n = 0;
  return MB2_OK;
}
long long mb2_ctx_launch_count(const mb2_ctx* ctx) { return ctx->launches; }

int mb2_ctx_profile_begin(mb2_ctx* ctx) {
  if (!ctx) return MB2_ERR_ARG;
  cudaSetDevice(ctx->device);
  MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->prof.clear(); ctx->prof_extract_bytes = 0; ctx->profiling = true;
  return MB2_OK;
}

int mb2_ctx_profile_end(mb2_ctx* ctx, char* buf, int buflen) {
  if (!ctx || !buf || buflen <= 0) return MB2_ERR_ARG;
  cudaSetDevice(ctx->device);
  ctx->profiling = false;
  MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  std::vector<std::pair<std::string, std::pair<int, double> > > agg;
  for (auto& r : ctx->prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    std::string name = r.name;
    size_t lt = name.find('<');
    bool found = false;
    for (auto& a : agg) if (a.first == name) { a.second.first++; a.second.second += ms; found = true; break; }
    if (!found) agg.push_back({name, {1, (double)ms}});
    (void)lt;
  }
  ctx->prof.clear();
  std::string out;
  for (auto& a : agg) out += a.first + "\t" + std::to_string(a.second.first) + "\t" + std::to_string(a.second.second) + "\n";
  out += "__extract_gather_bytes__\t1\t" + std::to_string((double)ctx->prof_extract_bytes) + "\n";
  if ((int)out.size() + 1 > buflen) { ctx->set_error("profile_end: buffer too small"); return MB2_ERR_CAPACITY; }
  std::memcpy(buf, out.c_str(), out.size() + 1);
  return (int)agg.size();
}

int mb2_hessaff_detect(mb2_ctx* ctx, const float* pixels, int w, int h, const mb2_hessaff_params* par, double tilt, double zoom,
                       int as_regions, double* out_kp, int capacity) {
  if (!ctx || !pixels || !par || w <= 0 || h <= 0) return MB2_ERR_ARG;
  cudaSetDevice(ctx->device);
  ImgView img;
  int rc, n = 0;
  if ((rc = stage_image(ctx, pixels, w, h, &img))) return rc;
  if ((rc = detect_core(ctx, img, *par, tilt, zoom, as_regions, &n))) return rc;
  if ((rc = download_keys(ctx, ctx->kp_b.as<KeyOut>(), n, out_kp, capacity, ctx->rs_b))) return rc;
  MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  if (n > capacity) { ctx->set_error("hessaff: output capacity too small"); return MB2_ERR_CAPACITY; }
  return n;
}

int mb2_detect_orientation(mb2_ctx* ctx, const float* pixels, int w, int h, const double* in_kp, int n,
                           const mb2_orientation_params* par, double* out_kp, int capacity) {
  if (!ctx || !pixels || !par || w <= 0 || h <= 0 || n < 0) return MB2_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (n == 0) return 0;
  ImgView img;
  int rc, m = 0;
  if ((rc = stage_image(ctx, pixels, w, h, &img))) return rc;
  const void* d_in;
  if ((rc = mb2_stage_in(ctx, in_kp, (size_t)n * MB2_KP * 8, ctx->rs_b, &d_in))) return rc;
  MB2_CUDA_CHECK(ctx, ctx->kp_b.reserve((size_t)n * sizeof(KeyOut)));
  LAUNCH1D(ctx, k_kp_from_doubles, n, (const double*)d_in, n, ctx->kp_b.as<KeyOut>());
  if ((rc = orient_core(ctx, img, n, *par, &m))) return rc;
  if ((rc = download_keys(ctx, ctx->kp_c.as<KeyOut>(), m, out_kp, capacity, ctx->rs_b))) return rc;
  MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  if (m > capacity) { ctx->set_error("orientation: output capacity too small"); return MB2_ERR_CAPACITY; }
  return m;
}

int mb2_describe_sift(mb2_ctx* ctx, const float* pixels, int w, int h, const double* kp, int n, const mb2_sift_params* par,
                      uint8_t* desc_u8, float* patches) {
  if (!ctx || !pixels || !par || !desc_u8 || w <= 0 || h <= 0 || n < 0) return MB2_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (n == 0) return 0;
  ImgView img;
  int rc;
  if ((rc = stage_image(ctx, pixels, w, h, &img))) return rc;
  const void* d_in;
  if ((rc = mb2_stage_in(ctx, kp, (size_t)n * MB2_KP * 8, ctx->rs_b, &d_in))) return rc;
  MB2_CUDA_CHECK(ctx, ctx->kp_b.reserve((size_t)n * sizeof(KeyOut)));
  LAUNCH1D(ctx, k_kp_from_doubles, n, (const double*)d_in, n, ctx->kp_b.as<KeyOut>());
  if ((rc = describe_core(ctx, img, ctx->kp_b.as<KeyOut>(), n, *par))) return rc;
  MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(desc_u8, ctx->desc_u8.p, (size_t)n * 128, cudaMemcpyDeviceToHost, ctx->stream));
  if (patches)
    MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(patches, priv(ctx)->last_patches, (size_t)n * 41 * 41 * 4, cudaMemcpyDeviceToHost, ctx->stream));
  MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return n;
}

static int view_post(mb2_ctx* ctx, const ImgView& img, int n, const double* H, int orig_w, int orig_h, const mb2_orientation_params* ori,
                     const mb2_sift_params* desc, int slot, int append, double* det_kp, double* reproj_kp, uint8_t* desc_u8, int capacity);

// One (detector, view) pass; det_h != NULL: HessianAffine, det_m != NULL: MSER.
static int view_core(mb2_ctx* ctx, const float* pixels, int w, int h, const double* H, int orig_w, int orig_h,
                     const mb2_hessaff_params* det_h, const mb2_mser_params* det_m, const mb2_orientation_params* ori, const mb2_sift_params* desc,
                     int slot, int append, double* det_kp, double* reproj_kp, uint8_t* desc_u8, int capacity) {
  if (!ctx || !pixels || !H || (!det_h && !det_m) || !ori || !desc || slot < 0 || slot >= MB2_MAX_SLOTS) return MB2_ERR_ARG;
  cudaSetDevice(ctx->device);
  ImgView img;
  int rc, n = 0;
  ctx->last_view_n = 0;
  if ((rc = stage_image(ctx, pixels, w, h, &img))) return rc;
  if (det_h) { if ((rc = detect_core(ctx, img, *det_h, 1.0, 1.0, 1, &n))) return rc; }
  else if ((rc = mb2_mser_core(ctx, img, *det_m, 1.0, 1.0, 1, &n, nullptr, 0))) return rc;
  return view_post(ctx, img, n, H, orig_w, orig_h, ori, desc, slot, append, det_kp, reproj_kp, desc_u8, capacity);
}

// DetectOrientation -> ReprojectRegions -> DescribeRegions on the n detected regions in ctx->kp_b (imagerepresentation.cpp:1254-1341)
static int view_post(mb2_ctx* ctx, const ImgView& img, int n, const double* H, int orig_w, int orig_h, const mb2_orientation_params* ori,
                     const mb2_sift_params* desc, int slot, int append, double* det_kp, double* reproj_kp, uint8_t* desc_u8, int capacity) {
  int rc, m = 0, k = 0;
  ctx->last_view_n = 0;
  if ((rc = orient_core(ctx, img, n, *ori, &m))) return rc;
  RegionSlot& rs = ctx->slots[slot];
  if (!append) rs.n = 0;
  if (m > 0) {
    // ReprojectRegions (synth-detection.cpp:541-616): Hinv via cv::invert (closed form for 3x3)
    double Hinv[9];
    {
      const double* S = H;
      double d = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6]);
      if (d == 0) { ctx->set_error("view: singular H"); return MB2_ERR_ARG; }
      d = 1. / d;
      Hinv[0] = (S[4] * S[8] - S[5] * S[7]) * d; Hinv[1] = (S[2] * S[7] - S[1] * S[8]) * d; Hinv[2] = (S[1] * S[5] - S[2] * S[4]) * d;
      Hinv[3] = (S[5] * S[6] - S[3] * S[8]) * d; Hinv[4] = (S[0] * S[8] - S[2] * S[6]) * d; Hinv[5] = (S[2] * S[3] - S[0] * S[5]) * d;
      Hinv[6] = (S[3] * S[7] - S[4] * S[6]) * d; Hinv[7] = (S[1] * S[6] - S[0] * S[7]) * d; Hinv[8] = (S[0] * S[4] - S[1] * S[3]) * d;
    }
    const int is_eye = (std::fabs(H[0] - 1.0) + std::fabs(H[1]) + std::fabs(H[2]) + std::fabs(H[3]) + std::fabs(H[4] - 1.0) +
                            std::fabs(H[5]) + std::fabs(H[6]) + std::fabs(H[7]) + std::fabs(H[8] - 1.0) < 0.01) ? 1 : 0;
    MB2_CUDA_CHECK(ctx, ctx->rs_b.reserve(9 * 8));
    MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->rs_b.p, Hinv, sizeof Hinv, cudaMemcpyHostToDevice, ctx->stream));
    // oriented regions are in kp_c; reproj -> kp_a; compact both -> kp_b (det), kp_c... use rs_c for reproj
    MB2_CUDA_CHECK(ctx, ctx->kp_a.reserve((size_t)m * sizeof(KeyOut)));
    mb2_launch_reproject(ctx, ctx->kp_c.as<KeyOut>(), m, ctx->rs_b.as<double>(), is_eye, orig_w, orig_h, ctx->kp_a.as<KeyOut>());
    MB2_CUDA_CHECK(ctx, ctx->kp_b.reserve((size_t)m * sizeof(KeyOut)));
    MB2_CUDA_CHECK(ctx, ctx->rs_c.reserve((size_t)m * sizeof(KeyOut)));
    if ((rc = compact(ctx, ctx->kp_c.as<KeyOut>(), ctx->kp_a.as<KeyOut>(), m, ctx->kp_b.as<KeyOut>(), ctx->rs_c.as<KeyOut>(), ctx->misc, &k)))
      return rc;
  }
  if (k > 0) {
    if ((rc = describe_core(ctx, img, ctx->kp_b.as<KeyOut>(), k, *desc))) return rc;
    // keep on device for matching
    const int base = rs.n;
    {
      DevBuf nd, nx;
      if (base > 0) {  // grow-preserving append
        MB2_CUDA_CHECK(ctx, nd.reserve((size_t)(base + k) * 128));
        MB2_CUDA_CHECK(ctx, nx.reserve((size_t)(base + k) * 16));
        MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(nd.p, rs.desc.p, (size_t)base * 128, cudaMemcpyDeviceToDevice, ctx->stream));
        MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(nx.p, rs.xy.p, (size_t)base * 16, cudaMemcpyDeviceToDevice, ctx->stream));
        MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        rs.desc.release(); rs.xy.release();
        rs.desc = nd; rs.xy = nx;
      } else {
        MB2_CUDA_CHECK(ctx, rs.desc.reserve((size_t)k * 128));
        MB2_CUDA_CHECK(ctx, rs.xy.reserve((size_t)k * 16));
      }
    }
    MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(rs.desc.as<uint8_t>() + (size_t)base * 128, ctx->desc_u8.p, (size_t)k * 128, cudaMemcpyDeviceToDevice,
                                        ctx->stream));
    LAUNCH1D(ctx, k_xy_from_keys, k, ctx->rs_c.as<KeyOut>(), k, rs.xy.as<double>() + (size_t)base * 2);
    rs.n = base + k;
    ctx->last_view_n = k;
    // results to the host
    const int mcap = std::min(k, capacity);
    if (det_kp && (rc = download_keys(ctx, ctx->kp_b.as<KeyOut>(), k, det_kp, capacity, ctx->rs_a))) return rc;
    MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    if (reproj_kp && (rc = download_keys(ctx, ctx->rs_c.as<KeyOut>(), k, reproj_kp, capacity, ctx->rs_a))) return rc;
    if (desc_u8 && mcap > 0)
      MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(desc_u8, ctx->desc_u8.p, (size_t)mcap * 128, cudaMemcpyDeviceToHost, ctx->stream));
    MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    if (k > capacity && (det_kp || reproj_kp || desc_u8)) { ctx->set_error("view: output capacity too small"); return MB2_ERR_CAPACITY; }
  }
  return k;
}

int mb2_detect_describe_view(mb2_ctx* ctx, const float* pixels, int w, int h, const double* H, int orig_w, int orig_h,
                             const mb2_hessaff_params* det, const mb2_orientation_params* ori, const mb2_sift_params* desc, int slot,
                             int append, double* det_kp, double* reproj_kp, uint8_t* desc_u8, int capacity) {
  if (!det) return MB2_ERR_ARG;
  return view_core(ctx, pixels, w, h, H, orig_w, orig_h, det, nullptr, ori, desc, slot, append, det_kp, reproj_kp, desc_u8, capacity);
}
int mb2_detect_describe_view_mser(mb2_ctx* ctx, const float* pixels, int w, int h, const double* H, int orig_w, int orig_h,
                                  const mb2_mser_params* det, const mb2_orientation_params* ori, const mb2_sift_params* desc, int slot,
                                  int append, double* det_kp, double* reproj_kp, uint8_t* desc_u8, int capacity) {
  if (!det) return MB2_ERR_ARG;
  return view_core(ctx, pixels, w, h, H, orig_w, orig_h, nullptr, det, ori, desc, slot, append, det_kp, reproj_kp, desc_u8, capacity);
}

long long mb2_ctx_tree_epoch(const mb2_ctx* ctx) { return ctx ? ctx->tree_epoch : 0; }
int mb2_ctx_wait_tree(mb2_ctx* ctx, mb2_ctx* src) {
  if (!ctx || !src || ctx->device != src->device || !src->ev_tree) return MB2_ERR_ARG;
  cudaSetDevice(ctx->device);
  MB2_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, src->ev_tree, 0));
  return MB2_OK;
}

int mb2_synth_view(mb2_ctx* ctx, const float* pixels, int w, int h, const mb2_view_params* view, float* out, int capacity, int* ow, int* oh,
                   double* H9) {
  if (!ctx || !pixels || !view || !ow || !oh || !H9 || w <= 0 || h <= 0) return MB2_ERR_ARG;
  cudaSetDevice(ctx->device);
  ImgView img, v;
  int rc;
  if ((rc = stage_image(ctx, pixels, w, h, &img))) return rc;
  const int ident = mb2_synth_core(ctx, img, *view, &v, H9);
  if (ident < 0) return ident;
  *ow = v.cols; *oh = v.rows;
  if (out && (long long)v.cols * v.rows <= capacity)
    MB2_CUDA_CHECK(ctx, cudaMemcpy2DAsync(out, (size_t)v.cols * 4, v.p, (size_t)v.pitch * 4, (size_t)v.cols * 4, v.rows, cudaMemcpyDeviceToHost, ctx->stream));
  MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return ident;
}
int mb2_detect_describe_synth_view(mb2_ctx* ctx, const float* pixels, int w, int h, const mb2_view_params* view, int detector,
                                   const mb2_hessaff_params* hess, const mb2_mser_params* mser, const mb2_orientation_params* ori,
                                   const mb2_sift_params* desc, int slot, int append, double* det_kp, double* reproj_kp, uint8_t* desc_u8,
                                   int capacity) {
  if (!ctx || !pixels || !view || !ori || !desc || slot < 0 || slot >= MB2_MAX_SLOTS || (detector == 0 && !hess) || (detector == 3 && !mser) ||
      (detector != 0 && detector != 3))
    return MB2_ERR_ARG;
  cudaSetDevice(ctx->device);
  ImgView img, v;
  int rc, n = 0;
  double H[9];
  ctx->last_view_n = 0;
  if ((rc = stage_image(ctx, pixels, w, h, &img, true))) return rc;
  if ((rc = mb2_synth_core(ctx, img, *view, &v, H)) < 0) return rc;
  const double tilt = std::fabs(view->tilt), zoom = view->zoom;   // SynthImage::tilt / zoom as DetectAffineRegions passes them on
  if (detector == 0) { if ((rc = detect_core(ctx, v, *hess, tilt, zoom, 1, &n))) return rc; }
  else if ((rc = mb2_mser_core(ctx, v, *mser, tilt, zoom, 1, &n, nullptr, 0))) return rc;
  return view_post(ctx, v, n, H, w, h, ori, desc, slot, append, det_kp, reproj_kp, desc_u8, capacity);
}

int mb2_mser_detect_pair(mb2_ctx* ctx, const float* pixels1, const float* pixels2, int w, int h, const mb2_mser_params* par, int* n1, int* n2) {
  if (!ctx || !pixels1 || !pixels2 || !par || !n1 || !n2 || w <= 0 || h <= 0) return MB2_ERR_ARG;
  cudaSetDevice(ctx->device);
  ImgView a, b;
  int rc;
  ctx->pair_n[0] = ctx->pair_n[1] = 0;
  if ((rc = stage_image(ctx, pixels1, w, h, &a))) return rc;
  {  // second image next to the first
    const int pitch = pitch_of(w);
    MB2_CUDA_CHECK(ctx, ctx->img2.reserve((size_t)pitch * h * 4));
    cudaMemcpyKind kind = mb2_is_device_ptr(pixels2) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    MB2_CUDA_CHECK(ctx, cudaMemcpy2DAsync(ctx->img2.p, (size_t)pitch * 4, pixels2, (size_t)w * 4, (size_t)w * 4, h, kind, ctx->stream));
    b.p = ctx->img2.as<float>(); b.rows = h; b.cols = w; b.pitch = pitch;
  }
  if ((rc = mb2_mser_core_pair(ctx, a, b, *par, 1, n1, n2))) return rc;
  const int n = *n1 + *n2;
  if (n > 0) {
    MB2_CUDA_CHECK(ctx, ctx->pair_keys.reserve((size_t)n * sizeof(KeyOut)));
    MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->pair_keys.p, ctx->kp_b.p, (size_t)n * sizeof(KeyOut), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  ctx->pair_n[0] = *n1; ctx->pair_n[1] = *n2; ctx->pair_w = w; ctx->pair_h = h;
  MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));   // the keys may be picked up by another context (mb2_describe_view_of_pair)
  return n;
}
int mb2_describe_view_of_pair(mb2_ctx* ctx, mb2_ctx* src, int which, const double* H, int orig_w, int orig_h, const mb2_orientation_params* ori,
                              const mb2_sift_params* desc, int slot, int append, double* det_kp, double* reproj_kp, uint8_t* desc_u8,
                              int capacity) {
  if (!src) src = ctx;
  if (!ctx || which < 0 || which > 1 || !H || !ori || !desc || slot < 0 || slot >= MB2_MAX_SLOTS || src->pair_w <= 0 || src->device != ctx->device)
    return MB2_ERR_ARG;
  cudaSetDevice(ctx->device);
  ImgView img;   // the staged images and the keys stay where mb2_mser_detect_pair left them (same device: readable from any context)
  img.p = which ? src->img2.as<float>() : src->img.as<float>(); img.rows = src->pair_h; img.cols = src->pair_w; img.pitch = pitch_of(src->pair_w);
  const int n = src->pair_n[which];
  if (n > 0) {
    MB2_CUDA_CHECK(ctx, ctx->kp_b.reserve((size_t)n * sizeof(KeyOut)));
    MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->kp_b.p, src->pair_keys.as<KeyOut>() + (which ? src->pair_n[0] : 0), (size_t)n * sizeof(KeyOut),
                                        cudaMemcpyDeviceToDevice, ctx->stream));
  }
  return view_post(ctx, img, n, H, orig_w, orig_h, ori, desc, slot, append, det_kp, reproj_kp, desc_u8, capacity);
}

int mb2_mser_detect(mb2_ctx* ctx, const float* pixels, int w, int h, const mb2_mser_params* par, double tilt, double zoom, int as_regions,
                    double* out_kp, int capacity) {
  if (!ctx || !pixels || !par || w <= 0 || h <= 0) return MB2_ERR_ARG;
  cudaSetDevice(ctx->device);
  ImgView img;
  int rc, n = 0;
  if ((rc = stage_image(ctx, pixels, w, h, &img))) return rc;
  if ((rc = mb2_mser_core(ctx, img, *par, tilt, zoom, as_regions, &n, nullptr, 0))) return rc;
  if ((rc = download_keys(ctx, ctx->kp_b.as<KeyOut>(), n, out_kp, capacity, ctx->rs_b))) return rc;
  MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  if (n > capacity) { ctx->set_error("mser: output capacity too small"); return MB2_ERR_CAPACITY; }
  return n;
}
int mb2_mser_regions(mb2_ctx* ctx, const float* pixels, int w, int h, const mb2_mser_params* par, double* out_rows, int capacity) {
  if (!ctx || !pixels || !par || !out_rows || w <= 0 || h <= 0 || capacity <= 0) return MB2_ERR_ARG;
  cudaSetDevice(ctx->device);
  ImgView img;
  int rc, n = 0;
  if ((rc = stage_image(ctx, pixels, w, h, &img))) return rc;
  MB2_CUDA_CHECK(ctx, ctx->rs_u.reserve((size_t)capacity * 13 * 8));
  mb2_mser_params p = *par; p.mode = 0;
  if ((rc = mb2_mser_core(ctx, img, p, 1.0, 1.0, 0, &n, ctx->rs_u.as<double>(), capacity))) return rc;
  MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(out_rows, ctx->rs_u.p, (size_t)std::min(n, capacity) * 13 * 8, cudaMemcpyDeviceToHost, ctx->stream));
  MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return n;
}

int mb2_view_fetch(mb2_ctx* ctx, double* det_kp, double* reproj_kp, uint8_t* desc_u8, int capacity) {
  if (!ctx) return MB2_ERR_ARG;
  cudaSetDevice(ctx->device);
  const int k = ctx->last_view_n, m = std::min(k, capacity);
  if (m <= 0) return k;
  int rc;
  if (det_kp && (rc = download_keys(ctx, ctx->kp_b.as<KeyOut>(), m, det_kp, m, ctx->rs_a))) return rc;
  if (reproj_kp && (rc = download_keys(ctx, ctx->rs_c.as<KeyOut>(), m, reproj_kp, m, ctx->rs_b))) return rc;
  if (desc_u8) MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(desc_u8, ctx->desc_u8.p, (size_t)m * 128, cudaMemcpyDeviceToHost, ctx->stream));
  MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return k;
}

int mb2_match_fginn(mb2_ctx* ctx, const uint8_t* q_desc, int nq, const uint8_t* t_desc, int nt, const double* t_xy, double matchRatio,
                    double contradDist, int nn, double* out, int capacity) {
  if (!ctx || nq < 0 || nt < 0 || !out) return MB2_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (nq == 0 || nt == 0) return 0;
  const void *dq, *dt, *dxy;
  int rc;
  if ((rc = mb2_stage_in(ctx, q_desc, (size_t)nq * 128, ctx->rs_a, &dq))) return rc;
  if ((rc = mb2_stage_in(ctx, t_desc, (size_t)nt * 128, ctx->rs_b, &dt))) return rc;
  if ((rc = mb2_stage_in(ctx, t_xy, (size_t)nt * 16, ctx->rs_c, &dxy))) return rc;
  return match_core(ctx, (const uint8_t*)dq, nq, (const uint8_t*)dt, nt, (const double*)dxy, matchRatio, contradDist, nn, out, capacity);
}

int mb2_match_hamming(mb2_ctx* ctx, const uint8_t* q_desc, int nq, const uint8_t* t_desc, int nt, int desc_bytes, double matchDistanceThreshold,
                      double* out, int capacity) {
  if (!ctx || nq < 0 || nt < 0 || !out || desc_bytes <= 0) return MB2_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (nq == 0 || nt == 0) return 0;
  const void *dq, *dt;
  int rc;
  if ((rc = mb2_stage_in(ctx, q_desc, (size_t)nq * desc_bytes, ctx->rs_a, &dq))) return rc;
  if ((rc = mb2_stage_in(ctx, t_desc, (size_t)nt * desc_bytes, ctx->rs_b, &dt))) return rc;
  return match_hamming_core(ctx, (const uint8_t*)dq, nq, (const uint8_t*)dt, nt, desc_bytes, matchDistanceThreshold, out, capacity);
}

int mb2_match_slots(mb2_ctx* ctx, int q_slot, int t_slot, double matchRatio, double contradDist, int nn, double* out, int capacity) {
  if (!ctx || q_slot < 0 || q_slot >= MB2_MAX_SLOTS || t_slot < 0 || t_slot >= MB2_MAX_SLOTS || !out) return MB2_ERR_ARG;
  cudaSetDevice(ctx->device);
  RegionSlot &q = ctx->slots[q_slot], &t = ctx->slots[t_slot];
  return match_core(ctx, q.desc.as<uint8_t>(), q.n, t.desc.as<uint8_t>(), t.n, t.xy.as<double>(), matchRatio, contradDist, nn, out, capacity);
}

int mb2_match_slots_range(mb2_ctx* ctx, int q_slot, int t_slot, int q_lo, int q_hi, double matchRatio, double contradDist, int nn, double* out,
                          int capacity) {
  if (!ctx || q_slot < 0 || q_slot >= MB2_MAX_SLOTS || t_slot < 0 || t_slot >= MB2_MAX_SLOTS || !out) return MB2_ERR_ARG;
  cudaSetDevice(ctx->device);
  RegionSlot &q = ctx->slots[q_slot], &t = ctx->slots[t_slot];
  if (q_lo < 0 || q_hi > q.n || q_lo > q_hi) return MB2_ERR_ARG;
  if (q_hi == q_lo || t.n == 0) return 0;
  const int n = match_core(ctx, q.desc.as<uint8_t>() + (size_t)q_lo * 128, q_hi - q_lo, t.desc.as<uint8_t>(), t.n, t.xy.as<double>(), matchRatio, contradDist,
                           nn, out, capacity);
  for (int i = 0; i < n && i < capacity; i++) out[(size_t)i * 7] += q_lo;
  return n;
}

namespace MB2_NS {
__global__ void k_pack_records(const KeyOut* __restrict__ reproj, const uint8_t* __restrict__ desc, int n, uint8_t* __restrict__ dst) {
  // one warp per record: 128 descriptor bytes as 32 words, then 7 doubles
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= n) return;
  uint8_t* r = dst + (size_t)i * MB2_REGION_RECORD_BYTES;
  reinterpret_cast<uint32_t*>(r)[lane] = reinterpret_cast<const uint32_t*>(desc + (size_t)i * 128)[lane];
  if (lane < 7) reinterpret_cast<double*>(r + 128)[lane] = reproj[i].v[lane];
}
__global__ void k_unpack_records(const uint8_t* __restrict__ src, int n, uint8_t* __restrict__ desc, double* __restrict__ xy) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= n) return;
  const uint8_t* r = src + (size_t)i * MB2_REGION_RECORD_BYTES;
  reinterpret_cast<uint32_t*>(desc + (size_t)i * 128)[lane] = reinterpret_cast<const uint32_t*>(r)[lane];
  if (lane < 2) xy[(size_t)i * 2 + lane] = reinterpret_cast<const double*>(r + 128)[lane];
}
__global__ void k_gather_frames(const uint8_t* __restrict__ q, const uint8_t* __restrict__ t, const int* __restrict__ qi, const int* __restrict__ ti, int n,
                                double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 14) return;
  const int r = i / 14, c = i - r * 14;
  const uint8_t* rec = c < 7 ? q + (size_t)qi[r] * MB2_REGION_RECORD_BYTES : t + (size_t)ti[r] * MB2_REGION_RECORD_BYTES;
  out[i] = reinterpret_cast<const double*>(rec + 128)[c < 7 ? c : c - 7];
}
}  // namespace MB2_NS

namespace MB2_NS {
__global__ void k_records_checksum(const uint8_t* __restrict__ rec, int n, unsigned long long* __restrict__ out) {
  unsigned long long acc = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(rec + (size_t)i * MB2_REGION_RECORD_BYTES);
    unsigned long long h = 1469598103934665603ull;
    for (int k = 0; k < MB2_REGION_RECORD_BYTES / 4; k++) { h ^= w[k]; h *= 1099511628211ull; }
    acc += h * (unsigned long long)(i + 1);
  }
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}
}  // namespace MB2_NS
int mb2_records_checksum(mb2_ctx* ctx, const void* d_records, int n, unsigned long long* out) {
  if (!ctx || !out || n < 0 || (n > 0 && !d_records)) return MB2_ERR_ARG;
  cudaSetDevice(ctx->device);
  *out = 0;
  if (n == 0) return MB2_OK;
  MB2_CUDA_CHECK(ctx, ctx->misc.reserve(4096));
  unsigned long long* d = ctx->misc.as<unsigned long long>() + 64;
  MB2_CUDA_CHECK(ctx, cudaMemsetAsync(d, 0, 8, ctx->stream));
  MB2_LAUNCH(ctx, MB2_NS::k_records_checksum, ctx->num_sms * 4, 256, 0, (const uint8_t*)d_records, n, d);
  MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(out, d, 8, cudaMemcpyDeviceToHost, ctx->stream));
  MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return MB2_OK;
}

int mb2_records_gather_frames(mb2_ctx* ctx, const void* d_q, const void* d_t, const int* q_idx, const int* t_idx, int n, double* frames14) {
  if (!ctx || n < 0 || (n > 0 && (!d_q || !d_t || !q_idx || !t_idx || !frames14))) return MB2_ERR_ARG;
  if (n == 0) return 0;
  cudaSetDevice(ctx->device);
  MB2_CUDA_CHECK(ctx, ctx->rs_a.reserve((size_t)n * 8));
  MB2_CUDA_CHECK(ctx, ctx->rs_b.reserve((size_t)n * 14 * 8));
  int* d_qi = ctx->rs_a.as<int>(); int* d_ti = d_qi + n;
  MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(d_qi, q_idx, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(d_ti, t_idx, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  MB2_LAUNCH(ctx, MB2_NS::k_gather_frames, (n * 14 + 255) / 256, 256, 0, (const uint8_t*)d_q, (const uint8_t*)d_t, d_qi, d_ti, n, ctx->rs_b.as<double>());
  MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(frames14, ctx->rs_b.p, (size_t)n * 14 * 8, cudaMemcpyDeviceToHost, ctx->stream));
  MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return n;
}

namespace MB2_NS {
// FP64 peak probe for the scorer's roofline (MEASURED_PEAKS.json has no FP64 figure): 8 independent DFMA chains per thread
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, double a, double b, int iters) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++)
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = __fma_rn(x[i], a, b);
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace MB2_NS
int mb2_debug_fp64_peak(mb2_ctx* ctx, double* tflops) {
  if (!ctx || !tflops) return MB2_ERR_ARG;
  cudaSetDevice(ctx->device);
  const int grid = ctx->num_sms * 8, iters = 4096;
  MB2_CUDA_CHECK(ctx, ctx->rs_a.reserve((size_t)grid * 256 * 8));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  MB2_NS::k_dfma_peak<<<grid, 256, 0, ctx->stream>>>(ctx->rs_a.as<double>(), 1.0000001, 1e-9, iters);
  cudaEventRecord(e0, ctx->stream);
  for (int r = 0; r < 5; r++) MB2_NS::k_dfma_peak<<<grid, 256, 0, ctx->stream>>>(ctx->rs_a.as<double>(), 1.0000001, 1e-9, iters);
  cudaEventRecord(e1, ctx->stream);
  MB2_CUDA_CHECK(ctx, cudaEventSynchronize(e1));
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *tflops = 5.0 * 2.0 * grid * 256.0 * 8.0 * iters / (ms * 1e-3) / 1e12;
  return MB2_OK;
}

int mb2_ctx_make_current(mb2_ctx* ctx) { if (!ctx) return MB2_ERR_ARG; MB2_CUDA_CHECK(ctx, cudaSetDevice(ctx->device)); return MB2_OK; }
void* mb2_dev_alloc(mb2_ctx* ctx, size_t bytes) {
  if (!ctx) return nullptr;
  cudaSetDevice(ctx->device);
  void* p = nullptr;
  if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) { ctx->set_error("mb2_dev_alloc: out of device memory"); return nullptr; }
  return p;
}
void mb2_dev_free(mb2_ctx* ctx, void* p) { if (ctx && p) { cudaSetDevice(ctx->device); cudaFree(p); } }
int mb2_is_device_pointer(const void* p) { return p && mb2_is_device_ptr(p) ? 1 : 0; }
int mb2_dev_copy(mb2_ctx* ctx, void* dst, const void* src, size_t bytes, int kind) {
  if (!ctx || kind < 0 || kind > 2 || (bytes && (!dst || !src))) return MB2_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (bytes) MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(dst, src, bytes, kind == 0 ? cudaMemcpyHostToDevice : kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, ctx->stream));
  return MB2_OK;
}

int mb2_view_pack(mb2_ctx* ctx, void* d_dst, int capacity) {
  if (!ctx || (!d_dst && capacity > 0)) return MB2_ERR_ARG;
  cudaSetDevice(ctx->device);
  const int k = ctx->last_view_n;
  if (k > capacity) { ctx->set_error("view_pack: capacity too small"); return MB2_ERR_CAPACITY; }
  if (k > 0) MB2_LAUNCH(ctx, MB2_NS::k_pack_records, (k * 32 + 255) / 256, 256, 0, ctx->rs_c.as<KeyOut>(), ctx->desc_u8.as<uint8_t>(), k, (uint8_t*)d_dst);
  return k;
}

int mb2_slot_from_records(mb2_ctx* ctx, int slot, const void* d_records, int n) {
  if (!ctx || slot < 0 || slot >= MB2_MAX_SLOTS || n < 0 || (n > 0 && !d_records)) return MB2_ERR_ARG;
  cudaSetDevice(ctx->device);
  RegionSlot& s = ctx->slots[slot];
  MB2_CUDA_CHECK(ctx, s.desc.reserve((size_t)std::max(n, 1) * 128));
  MB2_CUDA_CHECK(ctx, s.xy.reserve((size_t)std::max(n, 1) * 16));
  if (n > 0) MB2_LAUNCH(ctx, MB2_NS::k_unpack_records, (n * 32 + 255) / 256, 256, 0, (const uint8_t*)d_records, n, s.desc.as<uint8_t>(), s.xy.as<double>());
  s.n = n;
  return n;
}

int mb2_score_models(mb2_ctx* ctx, int which, const double* u, int len, const double* models, int K, double th, double* resid, int* I,
                     double* J) {
  if (!ctx || !u || !models || len < 0 || K < 0 || which < 0 || which > 5) return MB2_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (len == 0 || K == 0) return 0;
  const void *du, *dm;
  int rc;
  if ((rc = mb2_stage_in(ctx, u, (size_t)len * 48, ctx->rs_a, &du))) return rc;
  // models and results go through pinned staging: these calls sit inside RANSAC's sequential loop, where a
  // pageable cudaMemcpyAsync would cost more than the kernel
  const size_t rbytes = resid ? (size_t)K * len * 8 : 0;
  const size_t out_bytes = (size_t)K * 16 + rbytes + 64;
  MB2_CUDA_CHECK(ctx, ctx->h_c.reserve((size_t)K * 72 + out_bytes));
  uint8_t* hp = ctx->h_c.as<uint8_t>();
  double* h_models = (double*)hp;
  double* h_J = (double*)(hp + (((size_t)K * 72 + 15) & ~(size_t)15));
  int* h_I = (int*)(h_J + K);
  double* h_res = (double*)(((uintptr_t)(h_I + K) + 15) & ~(uintptr_t)15);
  if (mb2_is_device_ptr(models)) dm = models;
  else {
    std::memcpy(h_models, models, (size_t)K * 72);
    MB2_CUDA_CHECK(ctx, ctx->rs_b.reserve((size_t)K * 72));
    MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->rs_b.p, h_models, (size_t)K * 72, cudaMemcpyHostToDevice, ctx->stream));
    dm = ctx->rs_b.p;
  }
  const size_t pbytes = (J || I) ? mb2_score_partial_bytes(len, K) : 0;
  MB2_CUDA_CHECK(ctx, ctx->rs_c.reserve(rbytes + (size_t)K * 16 + pbytes + 128));
  double* d_J = ctx->rs_c.as<double>();
  int* d_I = (int*)(d_J + K);
  double* d_res = (double*)(((uintptr_t)(d_I + K) + 15) & ~(uintptr_t)15);
  void* d_part = (void*)(((uintptr_t)(d_res + (rbytes / 8)) + 15) & ~(uintptr_t)15);
  mb2_launch_score(ctx, which, (const double*)du, len, (const double*)dm, K, th, resid ? d_res : nullptr, (J || I) ? d_I : nullptr,
                   (J || I) ? d_J : nullptr, d_part);
  if (J || I) MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(h_J, d_J, (size_t)K * 12, cudaMemcpyDeviceToHost, ctx->stream));
  if (resid) MB2_CUDA_CHECK(ctx, cudaMemcpyAsync(h_res, d_res, rbytes, cudaMemcpyDeviceToHost, ctx->stream));
  MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  if (J) std::memcpy(J, h_J, (size_t)K * 8);
  if (I) std::memcpy(I, h_I, (size_t)K * 4);
  if (resid) std::memcpy(resid, h_res, rbytes);
  return K;
}

int mb2_debug_pyramid_level(mb2_ctx* ctx, int octave, int level, int want_resp, float* out, int* rows, int* cols) {
  if (!ctx) return MB2_ERR_ARG;
  cudaSetDevice(ctx->device);
  std::vector<OctaveLevels>& o = priv(ctx)->last_octaves;
  if (octave < 0 || octave >= (int)o.size() || level < 0 || level >= o[octave].nlevels) return MB2_ERR_ARG;
  const ImgView v = want_resp ? o[octave].resp[level] : o[octave].blur[level];
  if (rows) *rows = v.rows;
  if (cols) *cols = v.cols;
  if (out) {
    MB2_CUDA_CHECK(ctx, cudaMemcpy2DAsync(out, (size_t)v.cols * 4, v.p, (size_t)v.pitch * 4, (size_t)v.cols * 4, v.rows,
                                          cudaMemcpyDeviceToHost, ctx->stream));
    MB2_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return (int)o.size();
}

}  // extern "C"
