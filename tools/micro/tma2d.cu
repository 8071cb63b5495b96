// TMA 2-D float tile load probe: tools/micro/tma2d.bin boxw boxh cx cy cols rows  -> OK / mismatch / CUDA error
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tm, int boxw, int boxh, int cx, int cy, float* out) {
  extern __shared__ unsigned char raw[];
  float* s = reinterpret_cast<float*>(((uintptr_t)raw + 127) & ~(uintptr_t)127);
  uint64_t* bar = reinterpret_cast<uint64_t*>(s + boxw * boxh);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"((uint32_t)(boxw * boxh * 4)) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(s)), "l"(&tm), "r"(s32(bar)), "r"(cx), "r"(cy) : "memory");
  }
  __syncthreads();
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(s32(bar)), "r"(0u) : "memory");
  for (int i = threadIdx.x; i < boxw * boxh; i += blockDim.x) out[i] = s[i];
}
int main(int argc, char** argv) {
  int boxw = atoi(argv[1]), boxh = atoi(argv[2]), cx = atoi(argv[3]), cy = atoi(argv[4]), cols = atoi(argv[5]), rows = atoi(argv[6]);
  int pitch = (cols + 31) / 32 * 32;
  std::vector<float> h((size_t)pitch * rows);
  for (int y = 0; y < rows; y++) for (int x = 0; x < pitch; x++) h[(size_t)y * pitch + x] = y * 1000 + x + 1;
  float *d, *o; cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, boxw * boxh * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  CUtensorMap tm; cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, str[1] = {(cuuint64_t)pitch * 4}; cuuint32_t box[2] = {(cuuint32_t)boxw, (cuuint32_t)boxh}, es[2] = {1, 1};
  CUresult r = ((PFN)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r) { printf("encode failed %d\n", (int)r); return 1; }
  size_t sm = 128 + boxw * boxh * 4 + 16;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  k<<<1, 128, sm>>>(tm, boxw, boxh, cx, cy, o);
  cudaError_t e = cudaDeviceSynchronize();
  if (e) { printf("box %dx%d at (%d,%d) in %dx%d: CUDA error %s\n", boxw, boxh, cx, cy, cols, rows, cudaGetErrorString(e)); return 2; }
  std::vector<float> g(boxw * boxh); cudaMemcpy(g.data(), o, g.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int y = 0; y < boxh; y++) for (int x = 0; x < boxw; x++) {
    int gx = cx + x, gy = cy + y; float exp = (gx >= 0 && gy >= 0 && gx < cols && gy < rows) ? gy * 1000 + gx + 1 : 0.f;
    if (g[y * boxw + x] != exp) bad++;
  }
  printf("box %dx%d at (%d,%d) in %dx%d: %s (%d mismatches)\n", boxw, boxh, cx, cy, cols, rows, bad ? "MISMATCH" : "OK", bad);
  return 0;
}
