// LO-RANSAC fundamental matrix (DEGENSAC, exp_ransacFcustom) with GPU-batched hypothesis scoring: the C ABI entry over the
// sequential logic in ransac_f_logic.hpp (reference map and restructuring notes there).  Every residual vector over all
// correspondences -- FDs / FDsSym / exFDsSym of 7-point and LSQ models, HDs of the DEGENSAC homographies, FDs of the
// plane-and-parallax epipole samples -- is a launch of k_score (ransac.cu); the correspondences (and the off-plane subset rFtH
// samples from) stay resident on the device for the whole run.  No CPU scorer exists in the library.
#include "common.cuh"
#include "ransac_f_logic.hpp"

namespace mb2_ransac_f_host_detail {

struct GpuScorer {
  mb2_ctx* ctx;
  const double* dev[2] = {nullptr, nullptr};
  int n[2] = {0, 0};
  long launches = 0;
  int set_points(int slot, const double* u, int len) {
    DevBuf& b = slot == 0 ? ctx->rs_u : ctx->rs_v;
    if (b.reserve((size_t)std::max(len, 1) * 48) != cudaSuccess) { ctx->set_error("ransac_f: cudaMalloc failed"); return MB2_ERR_CUDA; }
    // synchronous copy: `u` of slot 1 is a temporary of the caller
    if (len > 0 && cudaMemcpy(b.p, u, (size_t)len * 48, cudaMemcpyHostToDevice) != cudaSuccess) return MB2_ERR_CUDA;
    dev[slot] = b.as<double>(); n[slot] = len;
    return MB2_OK;
  }
  int resid(int slot, int which, const double* model, double th, double* out, mb2_ransac_common::Score* S) {
    int I = 0; double J = 0;
    launches++;
    const int r = mb2_score_models(ctx, which, dev[slot], n[slot], model, 1, th, out, S ? &I : nullptr, S ? &J : nullptr);
    if (r < 0) return r;
    if (S) { S->I = (unsigned)I; S->J = J; }
    return MB2_OK;
  }
  int score(int slot, int which, const double* models, int K, double th, int* I, double* J) {
    launches++;
    const int r = mb2_score_models(ctx, which, dev[slot], n[slot], models, K, th, nullptr, I, J);
    return r < 0 ? r : MB2_OK;
  }
};

}  // namespace mb2_ransac_f_host_detail

extern "C" int mb2_ransac_f(mb2_ctx* ctx, const double* u, int len, double th, double conf, int max_sam, int errorType, int doSymCheck,
                            int do_lo, unsigned inlLimit, long seed, double* F, unsigned char* inl, int* data_out, double* Jout) {
  if (!ctx || !u || len < 0 || !F || !inl) return MB2_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (mb2_is_device_ptr(u)) { ctx->set_error("ransac_f: u must be a host pointer (the LO logic runs on the host)"); return MB2_ERR_ARG; }
  mb2_ransac_f_host_detail::GpuScorer sc;
  sc.ctx = ctx;
  mb2_rf::Result res;
  const int I = mb2_rf::ransac_f(sc, u, len, th, conf, max_sam, errorType, doSymCheck, do_lo, inlLimit, seed, F, inl, &res);
  if (I < 0) return I;
  if (data_out) { data_out[0] = res.samples; data_out[1] = res.lo; data_out[2] = res.Ih; data_out[3] = (int)sc.launches; }
  if (Jout) *Jout = res.J;
  return I;
}
