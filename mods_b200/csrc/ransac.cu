// Batched-hypothesis scorer for DEGENSAC / LO-RANSAC: K models x n correspondences in one launch.
//
// Reference scorers (all closed form, f64): HDs (degensac/Htools.c:158-196) with lin_hg (:17-55) and
// pinvJ (:132-156) fused so the n x 18 `lin` matrix is never built; HDsSym (:199-240), HDsSymMax
// (:241-282) with ccmath minv (matutls/minv.c) for the 3x3 inverse; FDs / FDsSym
// (degensac/Ftools.c:82-123); inlier count d <= th and MSAC gain truncQuad (rtools.c:228-236).
// Built with -fmad=false: residuals are bit-identical to the CPU scorers; the MSAC sum J is a
// fixed-order tree sum (deterministic, differs from the serial CPU sum only in the last bits).
#include "common.cuh"
#undef MB2_NS
#define MB2_NS mb2_ransac_detail
#include "ransac.cuh"

#include "minv3.hpp"
__host__ __device__ bool mb2_minv3(double* a) { return mb2_minv3_impl(a); }

namespace MB2_NS {

__device__ __forceinline__ double score_HDs(const double* u, const double* H) {
  const double x1 = u[0], y1 = u[1];
  const double z1[9] = {u[3], 0, -x1 * u[3], u[4], 0, -x1 * u[4], u[5], 0, -x1 * u[5]};
  const double z2[9] = {0, u[3], -y1 * u[3], 0, u[4], -y1 * u[4], 0, u[5], -y1 * u[5]};
  double r1 = 0, r2 = 0;
#pragma unroll
  for (int j = 0; j < 9; j++) { r1 += H[j] * z1[j]; r2 += H[j] * z2[j]; }
  const double a = H[0] - H[2] * u[0];
  const double b = H[3] - H[5] * u[0];
  const double c = -H[8] - H[2] * u[3] - H[5] * u[4];
  const double d = H[1] - H[2] * u[1];
  const double e = H[4] - H[5] * u[1];
  double pJ[8];
  const double a2 = a * a, b2 = b * b, c2 = c * c, d2 = d * d, e2 = e * e;
  const double c2pd2 = c2 + d2, ab = a * b, de = d * e;
  const double Q = c * (c2pd2 + e2);
  pJ[0] = -b * de + a * (c2 + e2);
  pJ[1] = b * c2pd2 - a * de;
  pJ[2] = Q;
  pJ[3] = -c * (a * d + b * e);
  pJ[4] = d * (b2 + c2) - ab * e;
  pJ[5] = -ab * d + e * (a2 + c2);
  pJ[6] = pJ[3];
  pJ[7] = c * (a2 + b2 + c2);
  const double N = a * pJ[0] + b * pJ[1] + c * pJ[2];
#pragma unroll
  for (int q = 0; q < 8; q++) pJ[q] /= N;
  double s = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) { const double v = pJ[j] * r1 + pJ[j + 4] * r2; s += v * v; }
  return s;
}

__device__ __forceinline__ double score_HDsSym(const double* u, const double* Hinv, const double* H1, bool useMax) {
  const double a = H1[6] * u[0] + H1[7] * u[1] + H1[8];
  const double b = Hinv[6] * u[3] + Hinv[7] * u[4] + Hinv[8];
  double xa = (H1[0] * u[0] + H1[1] * u[1] + H1[2]) / a;
  double ya = (H1[3] * u[0] + H1[4] * u[1] + H1[5]) / a;
  double xdiff = u[3] - xa, ydiff = u[4] - ya;
  const double d1 = xdiff * xdiff + ydiff * ydiff;
  xa = (Hinv[0] * u[3] + Hinv[1] * u[4] + Hinv[2]) / b;
  ya = (Hinv[3] * u[3] + Hinv[4] * u[4] + Hinv[5]) / b;
  xdiff = u[0] - xa; ydiff = u[1] - ya;
  const double d2 = xdiff * xdiff + ydiff * ydiff;
  return useMax ? (d1 < d2 ? d2 : d1) : d1 + d2;
}

// sym: 0 FDs, 1 FDsSym, 2 the residual exFDsSym returns (Ftools.c:172-196: r^2 / (a b / (a + b)), used inside the F-matrix LO)
__device__ __forceinline__ double score_FDs(const double* u, const double* F, int sym) {
  const double u1 = u[0], u2 = u[1], u4 = u[3], u5 = u[4];
  const double rxc = F[0] * u4 + F[3] * u5 + F[6];
  const double ryc = F[1] * u4 + F[4] * u5 + F[7];
  const double rwc = F[2] * u4 + F[5] * u5 + F[8];
  const double r = (u1 * rxc + u2 * ryc + rwc);
  const double rx = F[0] * u1 + F[1] * u2 + F[2];
  const double ry = F[3] * u1 + F[4] * u2 + F[5];
  if (!sym) return r * r / (rxc * rxc + ryc * ryc + rx * rx + ry * ry);
  const double a = rxc * rxc + ryc * ryc, b = rx * rx + ry * ry;
  if (sym == 1) return r * r * (a + b) / (a * b);
  const double w = (a * b) / (a + b);
  return r * r / w;
}

// rtools.c:228-236
__device__ __forceinline__ double truncQuad_dev(double epsilon, double thr) {
  if (thr == 0) return 0;
  if (epsilon >= thr * 9 / 4) return 0;
  return 1 - (epsilon / (thr * 9 / 4));
}

constexpr int ST = 256;

struct Partial { double J; int I; int pad; };

__device__ __forceinline__ void load_model(int which, const double* __restrict__ models, int k, double* sM, double* sHinv, double* sH1) {
  const int tid = threadIdx.x;
  if (tid < 9) sM[tid] = models[(size_t)k * 9 + tid];
  __syncthreads();
  if (tid == 0 && (which == 1 || which == 2)) {
    const double t[9] = {sM[0], sM[3], sM[6], sM[1], sM[4], sM[7], sM[2], sM[5], sM[8]};
    double inv[9];
    for (int i = 0; i < 9; i++) { sHinv[i] = t[i]; inv[i] = t[i]; }
    mb2_minv3(inv);
    for (int i = 0; i < 9; i++) sH1[i] = inv[i];
  }
  __syncthreads();
}

// grid = (K models, S segments of `chunk` correspondences).  A CTA walks its segment with coalesced loads of
// the 6-double records, writes the residuals (optional) and one (I, J) partial; k_score_sum adds the S
// partials of a model in segment order, so the MSAC sum is a fixed-order tree sum whatever the grid is.
__global__ void __launch_bounds__(ST)
k_score(int which, const double* __restrict__ u, int len, int chunk, const double* __restrict__ models, double th,
        double* __restrict__ resid, Partial* __restrict__ part) {
  __shared__ double sM[9], sHinv[9], sH1[9];
  __shared__ double sJ[ST];
  __shared__ int sI[ST];
  const int k = blockIdx.x, seg = blockIdx.y, tid = threadIdx.x;
  load_model(which, models, k, sM, sHinv, sH1);
  const int lo = seg * chunk, hi = min(len, lo + chunk);
  double J = 0; int I = 0;
  for (int i = lo + tid; i < hi; i += ST) {
    double p[6];
#pragma unroll
    for (int j = 0; j < 6; j++) p[j] = u[(size_t)i * 6 + j];
    double d;
    switch (which) {
      case 0: d = score_HDs(p, sM); break;
      case 1: d = score_HDsSym(p, sHinv, sH1, false); break;
      case 2: d = score_HDsSym(p, sHinv, sH1, true); break;
      case 3: d = score_FDs(p, sM, 0); break;
      case 4: d = score_FDs(p, sM, 1); break;
      default: d = score_FDs(p, sM, 2); break;
    }
    if (resid) resid[(size_t)k * len + i] = d;
    if (d <= th) I++;
    J += truncQuad_dev(d, th);
  }
  if (!part) return;
  sJ[tid] = J; sI[tid] = I;
  __syncthreads();
  for (int s = ST / 2; s > 0; s >>= 1) {
    if (tid < s) { sJ[tid] += sJ[tid + s]; sI[tid] += sI[tid + s]; }
    __syncthreads();
  }
  if (tid == 0) { Partial P; P.J = sJ[0]; P.I = sI[0]; P.pad = 0; part[(size_t)k * gridDim.y + seg] = P; }
}

__global__ void k_score_sum(const Partial* __restrict__ part, int K, int S, int* __restrict__ I_out, double* __restrict__ J_out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  double J = 0; int I = 0;
  for (int s = 0; s < S; s++) { J += part[(size_t)k * S + s].J; I += part[(size_t)k * S + s].I; }
  if (I_out) I_out[k] = I;
  if (J_out) J_out[k] = J;
}

}  // namespace
using namespace MB2_NS;

// Segment plan: enough CTAs to fill the 148 SMs a few times over when there are few models.
void mb2_score_plan(int len, int K, int* S, int* chunk) {
  int s = (4 * 148 + K - 1) / K;
  const int max_s = (len + ST - 1) / ST;
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  int c = (len + s - 1) / s;
  c = (c + ST - 1) / ST * ST;
  *chunk = c; *S = (len + c - 1) / c;
}
size_t mb2_score_partial_bytes(int len, int K) { int S, c; mb2_score_plan(len, K, &S, &c); return (size_t)K * S * sizeof(Partial); }

void mb2_launch_score(mb2_ctx* ctx, int which, const double* d_u, int len, const double* d_models, int K, double th, double* d_resid,
                      int* d_I, double* d_J, void* d_partials) {
  if (K <= 0 || len <= 0) return;
  int S, chunk;
  mb2_score_plan(len, K, &S, &chunk);
  const bool want = d_I || d_J;
  MB2_LAUNCH(ctx, k_score, dim3(K, S), ST, 0, which, d_u, len, chunk, d_models, th, d_resid, want ? (Partial*)d_partials : (Partial*)nullptr);
  if (want) MB2_LAUNCH(ctx, k_score_sum, (K + 127) / 128, 128, 0, (const Partial*)d_partials, K, S, d_I, d_J);
}
