// TEST INFRASTRUCTURE: the H-matrix LO-RANSAC driver (mods_b200/csrc/ransac_host.cu -- the very file the library compiles, included
// unchanged) built for the host with the device services replaced: device buffers are host memory, and mb2_score_models is a CPU
// scorer that calls the ORACLE's residual functions (a function pointer handed in by the test).  This lets the sequential logic of
// exp_ransacHcustom be compared with the compiled reference (oracle/_ref) on many seeded scenes without a GPU.  Nothing in the library
// links or calls this file.  Build: nvcc -std=c++17 -O2 -shared -Xcompiler -fPIC,-fopenmp,-ffp-contract=off (host code only).
#include <cuda_runtime.h>
#include <cstdlib>
#include <cstring>

static cudaError_t t_malloc(void** p, size_t n) { *p = std::malloc(n); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static cudaError_t t_free(void* p) { std::free(p); return cudaSuccess; }
static cudaError_t t_copy(void* d, const void* s, size_t n) { std::memcpy(d, s, n); return cudaSuccess; }
#define cudaMalloc(p, n) t_malloc((void**)(p), (n))
#define cudaFree(p) t_free(p)
#define cudaFreeHost(p) t_free(p)
#define cudaMallocHost(p, n) t_malloc((void**)(p), (n))
#define cudaHostAlloc(p, n, f) t_malloc((void**)(p), (n))
#define cudaMemcpyAsync(d, s, n, kind, stream) t_copy((d), (s), (n))
#define cudaSetDevice(d) cudaSuccess
#define mb2_score_models t_score_models
#define mb2_ransac_h t_ransac_h_impl

#include "../../mods_b200/csrc/common.cuh"

typedef void (*score_fn)(int which, const double* u, const double* M, double* d, int len);
static score_fn g_score = nullptr;
static long g_calls = 0;

extern "C" int t_score_models(mb2_ctx*, int which, const double* u, int len, const double* models, int K, double th, double* resid, int* I, double* J) {
  g_calls++;
  std::vector<double> tmp((size_t)len);
  for (int k = 0; k < K; k++) {
    double* out = resid ? resid + (size_t)k * len : tmp.data();
    g_score(which, u, models + (size_t)k * 9, out, len);
    int c = 0; double j = 0;
    for (int i = 0; i < len; i++) {
      const double e = out[i];
      if (e <= th) c++;
      j += (th == 0 || e >= th * 9 / 4) ? 0.0 : 1 - (e / (th * 9 / 4));   // truncQuad, rtools.c:228-236
    }
    if (I) I[k] = c;
    if (J) J[k] = j;
  }
  return K;
}
bool mb2_is_device_ptr(const void*) { return false; }

#include "../../mods_b200/csrc/ransac_host.cu"

extern "C" int t_ransac_h(score_fn fn, const double* u, int len, double th, double conf, int max_sam, int errorType, int doSymCheck, long seed,
                          double* H, unsigned char* inl, int* data_out, double* J) {
  g_score = fn; g_calls = 0;
  mb2_ctx ctx;
  ctx.device = 0;
  const int r = t_ransac_h_impl(&ctx, u, len, th, conf, max_sam, errorType, doSymCheck, seed, H, inl, data_out, J);
  data_out[3] = (int)g_calls;
  return r;
}

// the least-squares set-up alone (ransac_common.hpp: u2h), for the comparison with the compiled reference's u2h at sizes up to the
// chunked / pooled large-set path
extern "C" void t_u2h(const double* u, const int* inl, int len, double* H) { mb2_ransac_common::u2h(u, inl, len, H); }
