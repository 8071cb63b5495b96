// TEST INFRASTRUCTURE: the F-matrix LO-RANSAC driver logic (mods_b200/csrc/ransac_f_logic.hpp -- the very code the library
// compiles) instantiated with a CPU scorer that calls the ORACLE's residual functions (a function pointer handed in by the test), so
// that the sequential logic can be compared with the compiled reference (oracle/_ref, exp_ransacFcustom) without a GPU.  Nothing in
// the library links or calls this file.
#include "../../mods_b200/csrc/ransac_f_logic.hpp"

typedef void (*score_fn)(int which, const double* u, const double* M, double* d, int len);

struct CpuScorer {
  score_fn fn;
  const double* pts[2] = {nullptr, nullptr}; int n[2] = {0, 0};
  std::vector<double> own1, tmp;
  long calls = 0, models = 0;
  int set_points(int slot, const double* u, int len) {
    if (slot == 1) { own1.assign(u, u + (size_t)len * 6); pts[1] = own1.data(); } else pts[slot] = u;
    n[slot] = len; return 0;
  }
  int resid(int slot, int which, const double* model, double th, double* out, mb2_ransac_common::Score* S) {
    calls++; models++;
    if (!out) { tmp.resize(n[slot]); out = tmp.data(); }
    fn(which, pts[slot], model, out, n[slot]);
    if (S) { S->I = 0; S->J = 0; for (int i = 0; i < n[slot]; i++) { S->J += mb2_ransac_common::truncQuad(out[i], th); if (out[i] <= th) S->I++; } }
    return 0;
  }
  int score(int slot, int which, const double* models_, int K, double th, int* I, double* J) {
    calls++; models += K;
    tmp.resize(n[slot]);
    for (int k = 0; k < K; k++) {
      fn(which, pts[slot], models_ + (size_t)k * 9, tmp.data(), n[slot]);
      int c = 0; double j = 0;
      for (int i = 0; i < n[slot]; i++) { j += mb2_ransac_common::truncQuad(tmp[i], th); if (tmp[i] <= th) c++; }
      if (I) I[k] = c;
      if (J) J[k] = j;
    }
    return 0;
  }
};

extern "C" {
// out5: I, samples, LO count, Ih, scorer calls
int t_ransac_f(score_fn fn, const double* u, int len, double th, double conf, int max_sam, int errorType, int doSymCheck, int do_lo,
               unsigned inlLimit, long seed, double* F, unsigned char* inl, int* out5) {
  CpuScorer sc; sc.fn = fn;
  mb2_rf::Result r;
  const int I = mb2_rf::ransac_f(sc, u, len, th, conf, max_sam, errorType, doSymCheck, do_lo, inlLimit, seed, F, inl, &r);
  out5[0] = r.I; out5[1] = r.samples; out5[2] = r.lo; out5[3] = r.Ih; out5[4] = (int)sc.calls;
  return I;
}
void t_svd3_ccmath_V(const double* A, double* V) { mb2_rf::svd3_ccmath_V(A, V); }
void t_u2f(const double* u, const int* inl, const double* w, int len, double* F) { mb2_rf::u2f_w(u, inl, w, len, F); }
int t_checksample(const double* F, const double* u7, double th, double* H) { return mb2_rf::checksample(F, u7, th, H); }
}
