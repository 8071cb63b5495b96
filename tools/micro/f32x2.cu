// Throughput of scalar FMUL+FADD vs packed FMUL2 / FADD2 on sm_100a (no FMA: parity code may not contract).
// nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -o /tmp/f32x2 tools/micro/f32x2.cu && /tmp/f32x2
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float2 up(u64 v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
constexpr int CH = 8, IT = 2048;
// per iteration and chain: acc = acc + k * v for two lanes (4 scalar FP instr)
__global__ void k_scalar(float* out, float k, float v) {
  float a[CH][2];
  for (int c = 0; c < CH; c++) { a[c][0] = threadIdx.x + c; a[c][1] = c; }
  for (int i = 0; i < IT; i++)
#pragma unroll
    for (int c = 0; c < CH; c++) { a[c][0] = __fadd_rn(a[c][0], __fmul_rn(k, a[c][1])); a[c][1] = __fadd_rn(a[c][1], __fmul_rn(v, a[c][0])); }
  float s = 0; for (int c = 0; c < CH; c++) s += a[c][0] + a[c][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// packed multiply, scalar adds (3 instr)
__global__ void k_mul2(float* out, float k, float v) {
  float a[CH][2];
  for (int c = 0; c < CH; c++) { a[c][0] = threadIdx.x + c; a[c][1] = c; }
  const u64 kv = pk(k, v);
  for (int i = 0; i < IT; i++)
#pragma unroll
    for (int c = 0; c < CH; c++) {
      u64 m; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(m) : "l"(kv), "l"(pk(a[c][1], a[c][0])));
      float2 p = up(m);
      a[c][0] = __fadd_rn(a[c][0], p.x); a[c][1] = __fadd_rn(a[c][1], p.y);
    }
  float s = 0; for (int c = 0; c < CH; c++) s += a[c][0] + a[c][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// scalar multiplies, packed add (3 instr)
__global__ void k_add2(float* out, float k, float v) {
  u64 a[CH];
  for (int c = 0; c < CH; c++) a[c] = pk(threadIdx.x + c, c);
  for (int i = 0; i < IT; i++)
#pragma unroll
    for (int c = 0; c < CH; c++) {
      float2 x = up(a[c]);
      float p0 = __fmul_rn(k, x.y), p1 = __fmul_rn(v, x.x);
      asm("add.rn.f32x2 %0, %1, %2;" : "=l"(a[c]) : "l"(a[c]), "l"(pk(p0, p1)));
    }
  float s = 0; for (int c = 0; c < CH; c++) { float2 x = up(a[c]); s += x.x + x.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// packed add only (1 instr per two lanes)
__global__ void k_addonly2(float* out, float k, float v) {
  u64 a[CH];
  for (int c = 0; c < CH; c++) a[c] = pk(threadIdx.x + c, c);
  const u64 kv = pk(k, v);
  for (int i = 0; i < IT; i++)
#pragma unroll
    for (int c = 0; c < CH; c++) asm("add.rn.f32x2 %0, %1, %2;" : "=l"(a[c]) : "l"(a[c]), "l"(kv));
  float s = 0; for (int c = 0; c < CH; c++) { float2 x = up(a[c]); s += x.x + x.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_addonly1(float* out, float k, float v) {
  float a[CH][2];
  for (int c = 0; c < CH; c++) { a[c][0] = threadIdx.x + c; a[c][1] = c; }
  for (int i = 0; i < IT; i++)
#pragma unroll
    for (int c = 0; c < CH; c++) { a[c][0] = __fadd_rn(a[c][0], k); a[c][1] = __fadd_rn(a[c][1], v); }
  float s = 0; for (int c = 0; c < CH; c++) s += a[c][0] + a[c][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> void run(const char* name, F f, float* d, double lane_ops) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f<<<148 * 8, 256>>>(d, 1.0001f, 0.9999f);
  cudaEventRecord(a);
  for (int r = 0; r < 5; r++) f<<<148 * 8, 256>>>(d, 1.0001f, 0.9999f);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
  printf("%-12s %.3f ms  %.2f T lane-ops/s\n", name, ms, lane_ops / ms / 1e9);
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  const double macs = 148.0 * 8 * 256 * CH * IT * 2;   // lane mul+add pairs
  run("scalar", k_scalar, d, macs * 2);
  run("mul2+add", k_mul2, d, macs * 2);
  run("mul+add2", k_add2, d, macs * 2);
  run("add only", k_addonly1, d, macs);
  run("add2 only", k_addonly2, d, macs);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
