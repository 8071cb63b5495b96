// TEST INFRASTRUCTURE: the verification stage of the host mirror (mods_b200/host/mods_host.cpp, included unchanged: duplicate filter,
// LORANSACFiltering glue with NaiveHCheck / H_LAF_check / F_LAF_check) built with its three GPU services replaced, so that the host
// arithmetic around them can be checked on the CPU: mb2_score_models calls the ORACLE's residual functions, mb2_ransac_h / mb2_ransac_f
// return a model and an inlier mask planted by the test.  Every other mb2_* symbol resolves to the real library and is never called here.
#include <cstring>
#include <vector>
#include "../../include/mods_b200.h"

typedef void (*score_fn)(int which, const double* u, const double* M, double* d, int len);
static score_fn g_score = nullptr;
static double g_model[9];
static std::vector<unsigned char> g_mask;
static int g_last_inlLimit = -1, g_last_errorType = -1, g_last_max_sam = -1;

extern "C" {
int t_score_models(mb2_ctx*, int which, const double* u, int len, const double* models, int K, double, double* resid, int*, int*) {
  for (int k = 0; k < K; k++) g_score(which, u, models + (size_t)k * 9, resid + (size_t)k * len, len);
  return K;
}
static int planted(int len, double* M, unsigned char* inl) {
  int I = 0;
  for (int i = 0; i < 9; i++) M[i] = g_model[i];
  for (int i = 0; i < len; i++) { inl[i] = i < (int)g_mask.size() ? g_mask[i] : 0; I += inl[i]; }
  return I;
}
int t_ransac_h(mb2_ctx*, const double*, int len, double, double, int max_sam, int errorType, int, long, double* H, unsigned char* inl, int* data_out, double* J) {
  g_last_errorType = errorType; g_last_max_sam = max_sam; g_last_inlLimit = -1;
  if (data_out) data_out[0] = data_out[1] = data_out[2] = 0;
  if (J) *J = 0;
  return planted(len, H, inl);
}
int t_ransac_f(mb2_ctx*, const double*, int len, double, double, int max_sam, int errorType, int, int, unsigned inlLimit, long, double* F, unsigned char* inl, int* data_out, double* J) {
  g_last_errorType = errorType; g_last_max_sam = max_sam; g_last_inlLimit = (int)inlLimit;
  if (data_out) data_out[0] = data_out[1] = data_out[2] = data_out[3] = 0;
  if (J) *J = 0;
  return planted(len, F, inl);
}
}
#define mb2_score_models(ctx, which, u, len, models, K, th, resid, I, J) t_score_models(ctx, which, u, len, models, K, th, resid, nullptr, nullptr)
#define mb2_ransac_h t_ransac_h
#define mb2_ransac_f t_ransac_f
#include "../../mods_b200/host/mods_host.cpp"
#undef mb2_score_models
#undef mb2_ransac_h
#undef mb2_ransac_f

extern "C" void t_plant(score_fn fn, const double* model9, const unsigned char* mask, int n) {
  g_score = fn; std::memcpy(g_model, model9, sizeof g_model); g_mask.assign(mask, mask + n);
}
extern "C" void t_last_call(int* out3) { out3[0] = g_last_errorType; out3[1] = g_last_max_sam; out3[2] = g_last_inlLimit; }
