// Exact all-pairs squared-L2 nearest neighbours over 128-D u8 descriptors with the FGINN
// (first geometrically inconsistent) ratio test -- MatchFlannFGINN (matching/matching.cpp:357-461)
// for vector_matcher = linear.
//
// Arithmetic: descriptors are integers 0..255 (siftdesc.cpp:218,242,258,274), exact in bf16;
// products <= 65025 and 128-term sums <= 8.4e6 < 2^24 are exact in fp32, so
// d(q,t) = |q|^2 + |t|^2 - 2 q.t from a bf16 x bf16 -> fp32 tensor-core contraction is bit-exact
// and independent of accumulation order.
//
// Instead of materialising the reference's 50-NN table, the decision is evaluated exactly in two
// streaming passes over the N1 x N2 distance matrix (distances never leave the SM):
//   pass 1: (d0, idx0) = min over trains, ties -> lower index.
//   between: thr(q) = the smallest integer distance dJ for which the reference's float test
//            (double)(d0 / dJ) <= ratio^2 holds (monotone in dJ).
//   pass 2: "failers" are trains with d < thr (they precede every "passer" in the sorted kNN list):
//            count them, flag any failer farther than contradDist from idx0 in the image, keep the
//            closest failer != idx0 (2nd NN) and the closest passer (the neighbour the loop accepts).
//   accept  <=> no inconsistent failer, #failers <= min(nn, N2) - 1, a passer exists.
//
// Epilogue economy (round 2): the streaming loops only keep running MINIMA (one FFMA + half an FMNMX3 per distance); a minimum is
// recorded with the 32-column block it came from, and the train index inside that block is recovered afterwards for the one winning
// block per query (k_nn_resolve: 32 dp4a distances per query).  Pass 2 looks at single distances only in blocks that hold a
// "failer" (d < thr), and a query that can no longer be accepted (an inconsistent failer, or more failers than the k-NN list is
// long) stops looking altogether.
//
// The contraction runs on tcgen05 (UMMA 128x128x16, bf16 -> fp32 in TMEM), operands staged by TMA
// with 128B swizzle; a SIMT dp4a kernel computes the same per-(query, chunk) partials and is used
// as the on-GPU cross-check (MB2_NN_IMPL=simt) -- both feed the same merge + finalize kernels.
#include "common.cuh"
#undef MB2_NS
#define MB2_NS mb2_nn_detail
#include "nn.cuh"
#include <cuda.h>
#include <cuda_bf16.h>

namespace MB2_NS {

constexpr unsigned long long KEY_INIT = 0xFFFFFFFFFFFFFFFFull;

__device__ __forceinline__ unsigned long long make_key(float d, int idx) {
  return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)idx;  // d >= 0: bit pattern is monotone
}

// ---------------------------------------------------------------------------------------------
// preparation: u8 -> bf16 rows (padded to a multiple of 256 rows) + squared norms
// ---------------------------------------------------------------------------------------------
__global__ void k_prepare(const uint8_t* __restrict__ desc, int n, int n_pad, __nv_bfloat16* __restrict__ out,
                          float* __restrict__ norms, float pad_norm) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n_pad) return;
  unsigned w = 0;
  if (row < n) w = ((const unsigned*)(desc + (size_t)row * 128))[lane];
  const int b0 = w & 255, b1 = (w >> 8) & 255, b2 = (w >> 16) & 255, b3 = w >> 24;
  __nv_bfloat162 lo = __floats2bfloat162_rn((float)b0, (float)b1), hi = __floats2bfloat162_rn((float)b2, (float)b3);
  __nv_bfloat162* o = (__nv_bfloat162*)(out + (size_t)row * 128 + lane * 4);
  o[0] = lo; o[1] = hi;
  int s = b0 * b0 + b1 * b1 + b2 * b2 + b3 * b3;
  for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) norms[row] = row < n ? (float)s : pad_norm;
}

__global__ void k_init_state(NNState st, int nq) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  st.best0[i] = KEY_INIT; st.best1[i] = KEY_INIT; st.bestP[i] = KEY_INIT; st.cnt[i] = 0; st.incons[i] = 0;
}

// thr(q): smallest integer dJ >= 1 with (double)(float(d0) / float(dJ)) <= sqminratio  (matching.cpp:435-437)
__global__ void k_threshold(NNState st, int nq, const float* __restrict__ qn, double sqminratio) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const unsigned long long k = st.best0[i];
  const float d0 = __uint_as_float((unsigned)(k >> 32));
  st.d0[i] = d0; st.idx0[i] = (int)(unsigned)k;
  double guess = floor((double)d0 / sqminratio);
  if (guess < 1) guess = 1;
  // No squared distance between two u8 descriptors reaches TMAX = 128 * 255^2 + 1: a threshold at or above it classifies every train as a
  // "failer", exactly like TMAX itself.  Clamping keeps every value below 2^24, where t +- 1.f is exact (above it `t += 1.f` would not
  // move and the walk below would never end for small ratios with a large d0).
  const float TMAX = 8323201.f;
  float t;
  if (guess >= (double)TMAX) t = TMAX;
  else {
    t = (float)guess;
    // walk to the exact boundary of the float test (a couple of steps at most)
    while (t > 1.f && (double)__fdiv_rn(d0, t - 1.f) <= sqminratio) t -= 1.f;
    while (t < TMAX && !((double)__fdiv_rn(d0, t) <= sqminratio)) t += 1.f;
  }
  st.thr[i] = t;
  st.thr_rel[i] = t - qn[i];  // compared against |t|^2 - 2 q.t (exact: all integers < 2^24)
}

// ---------------------------------------------------------------------------------------------
// per-thread accumulators of one query row over a range of trains
// ---------------------------------------------------------------------------------------------
struct Pass1Acc {
  float best; int idx;
  __device__ __forceinline__ void init() { best = 3.0e38f; idx = -1; }
  __device__ __forceinline__ void add(float v, int col) { if (v < best) { best = v; idx = col; } }
};

struct Pass2Acc {
  float bestP; int idxP; float best1; int idx1; int cnt; int incons;
  __device__ __forceinline__ void init() { bestP = 3.0e38f; idxP = -1; best1 = 3.0e38f; idx1 = -1; cnt = 0; incons = 0; }
  __device__ __forceinline__ void add(float v, int col, float thr_rel, int idx0, const double* __restrict__ txy, double contr2) {
    if (v < thr_rel) {
      cnt++;
      if (col != idx0) {
        if (v < best1) { best1 = v; idx1 = col; }
        const double dx = txy[2 * idx0] - txy[2 * col], dy = txy[2 * idx0 + 1] - txy[2 * col + 1];
        if (dx * dx + dy * dy > contr2) incons = 1;   // distanceSq, matching.cpp:174-179
      }
    } else if (v < bestP) { bestP = v; idxP = col; }
  }
};

__device__ __forceinline__ float fmin3(float a, float b, float c) {
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));   // FMNMX3
  return r;
}

// Rare path of pass 2, kept out of line so the streaming loop stays small: one 32-column block that holds at least one failer.
// vals = the 32 distances (relative: |t|^2 - 2 q.t).  The failers are replayed in column order; the block's closest passer is returned.
__device__ __noinline__ float pass2_block(Pass2Acc& a, const uint32_t* vals, float thr_rel, int col0, int idx0, const double* __restrict__ txy,
                                          double contr2) {
  unsigned failmask = 0;
  float mp = 3.0e38f;
#pragma unroll
  for (int j = 0; j < 32; j++) {
    const float v = __uint_as_float(vals[j]);
    const bool f = v < thr_rel;
    failmask |= (unsigned)f << j;
    mp = fminf(mp, f ? 3.0e38f : v);
  }
  while (failmask) {
    const int j = __ffs(failmask) - 1;
    failmask &= failmask - 1;
    const float v = __uint_as_float(vals[j]);
    const int col = col0 + j;
    a.cnt++;
    if (col != idx0) {
      if (v < a.best1) { a.best1 = v; a.idx1 = col; }
      const double dx = txy[2 * idx0] - txy[2 * col], dy = txy[2 * idx0 + 1] - txy[2 * col + 1];
      if (dx * dx + dy * dy > contr2) a.incons = 1;   // distanceSq, matching.cpp:174-179
    }
  }
  return mp;
}

// keys of best0 / bestP carry (distance, train index) from the SIMT kernel and (distance, first column of the 32-column block) from
// the tcgen05 kernel; both order equal distances by the lower index, and k_nn_resolve turns either into the exact index.
__device__ __forceinline__ void merge_pass1(NNState& st, int q, float qn, const Pass1Acc& a) {
  if (a.idx >= 0) atomicMin(&st.best0[q], make_key(a.best + qn, a.idx));
}
__device__ __forceinline__ void merge_pass2(NNState& st, int q, float qn, const Pass2Acc& a) {
  if (a.cnt) atomicAdd(&st.cnt[q], a.cnt);
  if (a.incons) atomicOr(&st.incons[q], 1);
  if (a.idx1 >= 0) atomicMin(&st.best1[q], make_key(a.best1 + qn, a.idx1));
  if (a.idxP >= 0) atomicMin(&st.bestP[q], make_key(a.bestP + qn, a.idxP));
}

// The low word of a key names a 32-column block (index & ~31): find the lowest train index inside it whose distance equals the key's.
// One warp per query, lane = column of the block; the distance is formed exactly as in the streaming kernels.
__global__ void __launch_bounds__(256)
k_nn_resolve(unsigned long long* __restrict__ keys, int nq, const uint8_t* __restrict__ q, const uint8_t* __restrict__ t, int nt,
             const float* __restrict__ qn, const float* __restrict__ tn) {
  const int qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (qi >= nq) return;
  const unsigned long long key = keys[qi];
  if (key == KEY_INIT) return;
  const float d = __uint_as_float((unsigned)(key >> 32));
  const int col = (int)((unsigned)key & ~31u) + lane;
  bool hit = false;
  if (col < nt) {
    const uint4* qp = (const uint4*)(q + (size_t)qi * 128);
    const uint4* tp = (const uint4*)(t + (size_t)col * 128);
    unsigned dot = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const uint4 a = __ldg(qp + i), b = __ldg(tp + i);
      dot = __dp4a(a.x, b.x, dot); dot = __dp4a(a.y, b.y, dot); dot = __dp4a(a.z, b.z, dot); dot = __dp4a(a.w, b.w, dot);
    }
    hit = __fadd_rn(__fmaf_rn(-2.f, (float)dot, tn[col]), qn[qi]) == d;
  }
  const unsigned m = __ballot_sync(0xffffffffu, hit);
  if (lane == 0 && m) keys[qi] = ((unsigned long long)(unsigned)(key >> 32) << 32) | (unsigned)(((unsigned)key & ~31u) + __ffs(m) - 1);
}

// ---------------------------------------------------------------------------------------------
// SIMT cross-check kernel: one thread per query, __dp4a on the raw u8 descriptors
// ---------------------------------------------------------------------------------------------
template <int PASS>
__global__ void __launch_bounds__(128)
k_nn_simt(const uint8_t* __restrict__ q, int nq, const uint8_t* __restrict__ t, int nt, const float* __restrict__ qn,
          const float* __restrict__ tn, NNState st, const double* __restrict__ txy, double contr2, int chunk) {
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  const int t0 = blockIdx.y * chunk, t1 = min(nt, t0 + chunk);
  if (qi >= nq) return;
  unsigned qa[32];
#pragma unroll
  for (int i = 0; i < 32; i++) qa[i] = ((const unsigned*)(q + (size_t)qi * 128))[i];
  Pass1Acc a1; Pass2Acc a2; a1.init(); a2.init();
  float thr_rel = 0.f; int idx0 = -1;
  if (PASS == 2) { thr_rel = st.thr_rel[qi]; idx0 = st.idx0[qi]; }
  for (int j = t0; j < t1; j++) {
    const unsigned* tp = (const unsigned*)(t + (size_t)j * 128);
    unsigned dot = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) dot = __dp4a(qa[i], tp[i], dot);
    const float v = __fmaf_rn(-2.f, (float)dot, tn[j]);
    if (PASS == 1) a1.add(v, j); else a2.add(v, j, thr_rel, idx0, txy, contr2);
  }
  if (PASS == 1) merge_pass1(st, qi, qn[qi], a1); else merge_pass2(st, qi, qn[qi], a2);
}

// ---------------------------------------------------------------------------------------------
// tcgen05 kernel
// ---------------------------------------------------------------------------------------------
constexpr int BM = 256;            // queries per work item (two UMMA M=128 halves)
constexpr int BN = 128;            // trains per tile (UMMA N)
constexpr int BK = 64;             // bf16 per swizzle-128B row
constexpr int KBLOCKS = 2;         // 128 / BK
constexpr int STAGES = 3;          // B-operand smem ring
constexpr int TSTAGES = 2;         // TMEM accumulator ring (2 x 256 columns)
constexpr int TILE_BYTES = 128 * BK * 2;          // one [128 x 64] bf16 box = 16 KB
constexpr int A_BYTES = 2 * KBLOCKS * TILE_BYTES; // 64 KB
constexpr int B_STAGE_BYTES = KBLOCKS * TILE_BYTES;  // 32 KB
// warp 0 TMA, warp 1 MMA, warps 2.. epilogue: 8 warps = one per 32 rows x 128 columns of a tile (pass 1: the double-buffered TMEM loads
// need the registers), 16 warps = one per 32 rows x 64 columns (pass 2: the rare failer path stalls a warp on dependent loads, more
// warps hide it).  Measured at 30k x 30k / 60k x 60k: pass 1 0.215 / 0.807 ms with 8 warps (0.234 / 0.912 with 16), pass 2 0.273 /
// 0.918 ms with 16 (0.299 / 1.055 with 8).
constexpr int SMEM_BYTES = A_BYTES + STAGES * B_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void tma_load_2d(const void* tmap, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// K-major, 128B-swizzled operand tile [rows][64 bf16]: start address, SBO = 1024 B (8 rows), version 1 (sm_100),
// layout type 2 (SWIZZLE_128B).  (cute::UMMA::SmemDescriptor, cute/arch/mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // bits [0,14)  start address >> 4
  d |= (uint64_t)0 << 16;                            // bits [16,30) leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                  // bits [32,46) stride byte offset
  d |= (uint64_t)1 << 46;                            // bits [46,48) descriptor version
  d |= (uint64_t)2 << 61;                            // bits [61,64) SWIZZLE_128B
  return d;
}
// kind::f16, BF16 x BF16 -> F32, K-major A and B, M = 128, N = 128 (cute::UMMA::InstrDescriptor)
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

template <int PASS, int EW>
__global__ void __launch_bounds__(32 * (2 + EW), 1)
k_nn_tc(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_t, int nq, int nt_pad,
        const float* __restrict__ qn, const float* __restrict__ tn, NNState st, const double* __restrict__ txy, double contr2,
        int tiles_per_chunk, int n_qblocks, int n_chunks, int kmax) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                          // [half][kblock] 16 KB tiles
  uint8_t* sB = smem + A_BYTES;                // [stage][kblock]
  uint64_t* bars = (uint64_t*)(smem + A_BYTES + STAGES * B_STAGE_BYTES);
  uint64_t* full = bars;                       // [STAGES]   TMA -> MMA
  uint64_t* empty = bars + STAGES;             // [STAGES]   MMA -> TMA
  uint64_t* tfull = bars + 2 * STAGES;         // [TSTAGES]  MMA -> epilogue
  uint64_t* tempty = tfull + TSTAGES;          // [TSTAGES]  epilogue -> MMA
  uint64_t* a_full = tempty + TSTAGES;         // A tile landed
  uint64_t* a_empty = a_full + 1;              // A tile no longer read by the tensor core
  uint32_t* tmem_slot = (uint32_t*)(a_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = n_qblocks * n_chunks;
  const int total_tiles = nt_pad / BN;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < TSTAGES; i++) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], EW); }
    mbar_init(a_full, 1); mbar_init(a_empty, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: all 512 columns (2 stages x 2 halves x 128 fp32 columns)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, a_phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int ch = item / n_qblocks, qb = item - ch * n_qblocks;   // chunk-major: the query blocks of one train chunk run side by side
        const int tile0 = ch * tiles_per_chunk, tile1 = min(total_tiles, tile0 + tiles_per_chunk);
        mbar_wait(a_empty, a_phase ^ 1);
        mbar_expect_tx(a_full, A_BYTES);
        for (int half = 0; half < 2; half++)
          for (int kb = 0; kb < KBLOCKS; kb++)
            tma_load_2d(&tmap_q, a_full, sA + (half * KBLOCKS + kb) * TILE_BYTES, kb * BK, qb * BM + half * 128);
        a_phase ^= 1;
        for (int tile = tile0; tile < tile1; tile++) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], B_STAGE_BYTES);
          for (int kb = 0; kb < KBLOCKS; kb++)
            tma_load_2d(&tmap_t, &full[stage], sB + (stage * KBLOCKS + kb) * TILE_BYTES, kb * BK, tile * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, tstage = 0, tphase = 0, a_phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int ch = item / n_qblocks;
        const int tile0 = ch * tiles_per_chunk, tile1 = min(total_tiles, tile0 + tiles_per_chunk);
        mbar_wait(a_full, a_phase);
        a_phase ^= 1;
        for (int tile = tile0; tile < tile1; tile++) {
          mbar_wait(&tempty[tstage], tphase ^ 1);
          mbar_wait(&full[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          for (int half = 0; half < 2; half++) {
            const uint32_t d_tmem = tmem_base + tstage * 256 + half * 128;
            for (int kb = 0; kb < KBLOCKS; kb++) {
              const uint64_t da0 = umma_desc(smem_u32(sA + (half * KBLOCKS + kb) * TILE_BYTES));
              const uint64_t db0 = umma_desc(smem_u32(sB + (stage * KBLOCKS + kb) * TILE_BYTES));
#pragma unroll
              for (int k = 0; k < BK / 16; k++)  // +32 B per UMMA_K step inside the swizzle atom
                umma_bf16(d_tmem, da0 + (uint64_t)(k * 2), db0 + (uint64_t)(k * 2), (kb | k) != 0);
            }
          }
          umma_commit(&empty[stage]);   // B slot reusable once these MMAs have read it
          umma_commit(&tfull[tstage]);  // accumulators ready
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          if (++tstage == TSTAGES) { tstage = 0; tphase ^= 1; }
        }
        umma_commit(a_empty);           // A tile reusable
      }
    }
  } else {
    // ===== epilogue: warps 2..; TMEM lane group = warp % 4 (a warp only reaches its own 32 TMEM lanes), query half and -- with 16
    // warps -- column half of the tile from (warp - 2) / 4 =====
    // A thread owns one query row.  Per 32-column block: 32 FFMA (|t|^2 - 2 q.t) and a tree of FMNMX3 give the block minimum; that
    // is all pass 1 needs (the winning block is recorded, the index inside it is recovered by k_nn_resolve), and all pass 2 needs
    // for a block without failers (block minimum >= thr: every column is a passer).  The TMEM load of the next block is in flight
    // while the current one is reduced.
    constexpr int CSPLIT = EW / 8, NBLK = BN / 32 / CSPLIT;   // column split of a tile among warps, 32-column blocks per warp
    const int lg = warp & 3, half = ((warp - 2) >> 2) / CSPLIT, cpart = ((warp - 2) >> 2) % CSPLIT;
    uint32_t tstage = 0, tphase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int ch = item / n_qblocks, qb = item - ch * n_qblocks;
      const int tile0 = ch * tiles_per_chunk, tile1 = min(total_tiles, tile0 + tiles_per_chunk);
      const int q = qb * BM + half * 128 + lg * 32 + lane;
      const bool qvalid = q < nq;
      Pass1Acc a1; Pass2Acc a2; a1.init(); a2.init();
      float thr_rel = -3.0e38f; int idx0 = -1;
      if (PASS == 2 && qvalid) {
        // a query that earlier work items already proved unacceptable (inconsistent failer, too many failers) looks at nothing
        if (__ldcg(&st.incons[q]) == 0 && __ldcg(&st.cnt[q]) <= kmax) { thr_rel = st.thr_rel[q]; idx0 = st.idx0[q]; }
      }
      for (int tile = tile0; tile < tile1; tile++) {
        mbar_wait(&tfull[tstage], tphase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + tstage * 256 + half * 128 + cpart * NBLK * 32;
        uint32_t ra[32], rb[32];
        tmem_ld32(taddr, ra);
#pragma unroll
        for (int c = 0; c < NBLK; c++) {
          uint32_t* cur = (c & 1) ? rb : ra;
          uint32_t* nxt = (c & 1) ? ra : rb;
          const int col0 = tile * BN + (cpart * NBLK + c) * 32;
          float tnv[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 t4 = __ldg((const float4*)(tn + col0 + j));
            tnv[j] = t4.x; tnv[j + 1] = t4.y; tnv[j + 2] = t4.z; tnv[j + 3] = t4.w;
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (c + 1 < NBLK) tmem_ld32(taddr + (c + 1) * 32, nxt);
          float m[4] = {3.0e38f, 3.0e38f, 3.0e38f, 3.0e38f};
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const float v0 = __fmaf_rn(-2.f, __uint_as_float(cur[j + 2 * k]), tnv[j + 2 * k]);
              const float v1 = __fmaf_rn(-2.f, __uint_as_float(cur[j + 2 * k + 1]), tnv[j + 2 * k + 1]);
              if (PASS == 2) { cur[j + 2 * k] = __float_as_uint(v0); cur[j + 2 * k + 1] = __float_as_uint(v1); }
              m[k] = fmin3(m[k], v0, v1);
            }
          }
          const float mn = fminf(fmin3(m[0], m[1], m[2]), m[3]);
          if (PASS == 1) {
            if (mn < a1.best) { a1.best = mn; a1.idx = col0; }
          } else {
            float mp = mn;
            if (mn < thr_rel) {   // the block holds a failer
              uint32_t tmp[32];   // the out-of-line call takes the block through local memory; only this rare path pays for it
#pragma unroll
              for (int j = 0; j < 32; j++) tmp[j] = cur[j];
              mp = pass2_block(a2, tmp, thr_rel, col0, idx0, txy, contr2);
              if (a2.incons | (a2.cnt > kmax)) thr_rel = -3.0e38f;   // cannot be accepted any more
            }
            if (mp < a2.bestP) { a2.bestP = mp; a2.idxP = col0; }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[tstage]);
        if (++tstage == TSTAGES) { tstage = 0; tphase ^= 1; }
      }
      if (qvalid) {
        if (PASS == 1) merge_pass1(st, q, qn[q], a1); else merge_pass2(st, q, qn[q], a2);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// finalize: one row per accepted query, in query order (stable compaction done by the caller)
// ---------------------------------------------------------------------------------------------
__global__ void k_finalize(NNState st, int nq, int nt, int nn, MatchRow* __restrict__ rows, int* __restrict__ accept) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const int k = nn < nt ? nn : nt;
  const unsigned long long kp = st.bestP[i], k1 = st.best1[i];
  const int cnt = st.cnt[i];  // failers incl. the NN itself == rank of the first passer
  const bool ok = (st.incons[i] == 0) && (kp != KEY_INIT) && (cnt >= 1) && (cnt <= k - 1);
  accept[i] = ok ? 1 : 0;
  MatchRow r;
  r.q = i; r.idx0 = st.idx0[i]; r.d0 = st.d0[i];
  r.idxJ = (int)(unsigned)kp; r.dJ = __uint_as_float((unsigned)(kp >> 32));
  const unsigned long long second = cnt > 1 ? k1 : kp;   // indices[1], dists[1]
  r.idx1 = (int)(unsigned)second; r.d1 = __uint_as_float((unsigned)(second >> 32));
  r.pad = 0;
  rows[i] = r;
}

// ---------------------------------------------------------------------------------------------
// matchRatio >= 1: the "all points" branch of MatchFlannFGINN (matching.cpp:397-428), which needs the sorted k-NN table itself: the
// tentative of a query is its NN paired with the first neighbour (in distance order, ties by train index) that lies farther than
// contradDist from the NN in the image, or with neighbour nn - 1 when there is none.  One thread per query keeps the exact nn nearest
// (distance, index) pairs in an insertion-sorted list while it walks all trains (every lane of a warp reads the same train row: one
// broadcast load); distances are exact integers in f32.  Rarely used (no shipped iters file sets a ratio >= 1), not tuned.
// ---------------------------------------------------------------------------------------------
#define MB2_NN_MAXK 64
__global__ void __launch_bounds__(128)
k_nn_topk(const uint8_t* __restrict__ q, int nq, const uint8_t* __restrict__ t, int nt, const float* __restrict__ qn, const float* __restrict__ tn,
          const double* __restrict__ txy, double contr2, int nn, MatchRow* __restrict__ rows, int* __restrict__ accept) {
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  unsigned qa[32];
#pragma unroll
  for (int i = 0; i < 32; i++) qa[i] = ((const unsigned*)(q + (size_t)qi * 128))[i];
  const int k = nn < nt ? nn : nt;
  float dk[MB2_NN_MAXK]; int ik[MB2_NN_MAXK];
  int cnt = 0;
  const float qq = qn[qi];
  for (int j = 0; j < nt; j++) {
    const unsigned* tp = (const unsigned*)(t + (size_t)j * 128);
    unsigned dot = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) dot = __dp4a(qa[i], tp[i], dot);
    const float v = __fadd_rn(qq, __fmaf_rn(-2.f, (float)dot, tn[j]));   // |q|^2 + |t|^2 - 2 q.t: integers below 2^24, exact
    if (cnt < k || v < dk[cnt - 1]) {      // trains come in index order: an equal distance stays behind the earlier one
      int p = cnt < k ? cnt : k - 1;
      while (p > 0 && dk[p - 1] > v) { dk[p] = dk[p - 1]; ik[p] = ik[p - 1]; p--; }
      dk[p] = v; ik[p] = j;
      if (cnt < k) cnt++;
    }
  }
  MatchRow r; r.q = qi; r.idx0 = ik[0]; r.d0 = dk[0]; r.idxJ = -1; r.dJ = 0; r.idx1 = k > 1 ? ik[1] : -1; r.d1 = k > 1 ? dk[1] : 0; r.pad = 0;
  int ok = 0;
  for (int j = 1; j < k; j++) {
    const double dx = txy[2 * ik[0]] - txy[2 * ik[j]], dy = txy[2 * ik[0] + 1] - txy[2 * ik[j] + 1];
    if (j == nn - 1 || dx * dx + dy * dy > contr2) { r.idxJ = ik[j]; r.dJ = dk[j]; ok = 1; break; }
  }
  rows[qi] = r; accept[qi] = ok;
}

// ---------------------------------------------------------------------------------------------
// Hamming 2-NN of binary descriptors (MatchFLANNDistance, matching/matching.cpp:607-666, binary_matcher = linear): one thread per query
// with the query's words in registers, the trains of the CTA's chunk staged through shared memory (every lane reads the same train
// word: broadcast), XOR + POPC.  The two smallest (distance << 32 | train index) keys per (query, chunk) go to `part`; keys are
// unique, so ties resolve to the lower train index.  k_hamming_finish merges the chunks and applies the distance threshold (:653).
// ---------------------------------------------------------------------------------------------
#define HM_TILE 128
template <int W>
__global__ void __launch_bounds__(128)
k_hamming_2nn(const uint32_t* __restrict__ q, int nq, const uint32_t* __restrict__ t, int nt, int chunk, unsigned long long* __restrict__ part) {
  __shared__ uint32_t st[HM_TILE * W];
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t qa[W];
#pragma unroll
  for (int i = 0; i < W; i++) qa[i] = qi < nq ? q[(size_t)qi * W + i] : 0u;
  const int j_lo = blockIdx.y * chunk, j_hi = min(nt, j_lo + chunk);
  unsigned long long b0 = ~0ull, b1 = ~0ull;
  for (int j0 = j_lo; j0 < j_hi; j0 += HM_TILE) {
    const int m = min(HM_TILE, j_hi - j0);
    __syncthreads();
    for (int i = threadIdx.x; i < m * W; i += blockDim.x) st[i] = t[(size_t)j0 * W + i];
    __syncthreads();
    for (int j = 0; j < m; j++) {
      unsigned d = 0;
#pragma unroll
      for (int i = 0; i < W; i++) d += __popc(qa[i] ^ st[j * W + i]);
      const unsigned long long key = ((unsigned long long)d << 32) | (unsigned)(j0 + j);
      if (key < b1) { if (key < b0) { b1 = b0; b0 = key; } else b1 = key; }
    }
  }
  if (qi < nq) { part[((size_t)qi * gridDim.y + blockIdx.y) * 2] = b0; part[((size_t)qi * gridDim.y + blockIdx.y) * 2 + 1] = b1; }
}
__global__ void k_hamming_finish(const unsigned long long* __restrict__ part, int nq, int n_chunks, int max_distance, MatchRow* __restrict__ rows,
                                 int* __restrict__ accept) {
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  unsigned long long b0 = ~0ull, b1 = ~0ull;
  for (int c = 0; c < 2 * n_chunks; c++) {
    const unsigned long long key = part[(size_t)qi * n_chunks * 2 + c];
    if (key < b1) { if (key < b0) { b1 = b0; b0 = key; } else b1 = key; }
  }
  MatchRow r; r.q = qi; r.idx0 = (int)(unsigned)b0; r.d0 = (float)(unsigned)(b0 >> 32);
  r.idxJ = r.idx1 = (int)(unsigned)b1; r.dJ = r.d1 = (float)(unsigned)(b1 >> 32); r.pad = 0;
  rows[qi] = r; accept[qi] = (int)(unsigned)(b0 >> 32) <= max_distance;
}
// rows of `bytes` bytes -> rows of W zero-padded 32-bit words (equal padding adds nothing to a bit count)
__global__ void k_hamming_words(const uint8_t* __restrict__ src, int n, int bytes, int W, uint32_t* __restrict__ dst) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * W) return;
  const size_t r = i / W; const int w = (int)(i - r * W);
  uint32_t v = 0;
  for (int b = 0; b < 4; b++) { const int e = w * 4 + b; if (e < bytes) v |= (uint32_t)src[r * bytes + e] << (8 * b); }
  dst[i] = v;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap(mb2_ctx* ctx, CUtensorMap* m, const void* base, int rows) {
  if (!ctx->tmap_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || !fn) { ctx->set_error("cuTensorMapEncodeTiled entry point not available"); return MB2_ERR_CUDA; }
    ctx->tmap_encode = fn;
  }
  cuuint64_t dims[2] = {128, (cuuint64_t)rows};
  cuuint64_t strides[1] = {128 * sizeof(__nv_bfloat16)};
  cuuint32_t box[2] = {(cuuint32_t)BK, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ((PFN_encodeTiled)ctx->tmap_encode)(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr,
                                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { ctx->set_error("cuTensorMapEncodeTiled failed: " + std::to_string((int)r)); return MB2_ERR_CUDA; }
  return MB2_OK;
}

}  // namespace
using namespace MB2_NS;

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
void mb2_nn_prepare(mb2_ctx* ctx, const uint8_t* d_desc, int n, int n_pad, void* d_bf16, float* d_norms, float pad_norm) {
  MB2_LAUNCH(ctx, k_prepare, (n_pad + 3) / 4, 128, 0, d_desc, n, n_pad, (__nv_bfloat16*)d_bf16, d_norms, pad_norm);
}
void mb2_nn_init_state(mb2_ctx* ctx, const NNState& st, int nq) {
  MB2_LAUNCH(ctx, k_init_state, (nq + 255) / 256, 256, 0, st, nq);
}
void mb2_nn_threshold(mb2_ctx* ctx, const NNState& st, int nq, const float* qn, double sqminratio) {
  MB2_LAUNCH(ctx, k_threshold, (nq + 255) / 256, 256, 0, st, nq, qn, sqminratio);
}
void mb2_nn_topk(mb2_ctx* ctx, const uint8_t* q, int nq, const uint8_t* t, int nt, const float* qn, const float* tn, const double* txy, double contr2,
                 int nn, MatchRow* rows, int* accept) {
  MB2_LAUNCH(ctx, k_nn_topk, (nq + 127) / 128, 128, 0, q, nq, t, nt, qn, tn, txy, contr2, nn, rows, accept);
}
void mb2_nn_hamming_words(mb2_ctx* ctx, const uint8_t* src, int n, int bytes, int W, uint32_t* dst) {
  const size_t total = (size_t)n * W;
  if (total) MB2_LAUNCH(ctx, k_hamming_words, (unsigned)((total + 255) / 256), 256, 0, src, n, bytes, W, dst);
}
int mb2_nn_hamming_chunks(mb2_ctx* ctx, int nq, int nt) {
  // enough CTAs to fill the SMs a few times over; a chunk is a whole number of shared-memory tiles
  const int qblocks = (nq + 127) / 128;
  int n_chunks = (4 * ctx->num_sms + qblocks - 1) / qblocks;
  const int max_chunks = (nt + HM_TILE - 1) / HM_TILE;
  if (n_chunks > max_chunks) n_chunks = max_chunks;
  if (n_chunks < 1) n_chunks = 1;
  if (n_chunks > 4096) n_chunks = 4096;
  return n_chunks;
}
void mb2_nn_hamming(mb2_ctx* ctx, const uint32_t* q, int nq, const uint32_t* t, int nt, int W, int n_chunks, int max_distance,
                    unsigned long long* part, MatchRow* rows, int* accept) {
  int chunk = (nt + n_chunks - 1) / n_chunks;
  chunk = (chunk + HM_TILE - 1) / HM_TILE * HM_TILE;
  dim3 grid((nq + 127) / 128, n_chunks);
  if (W == 4) MB2_LAUNCH(ctx, k_hamming_2nn<4>, grid, 128, 0, q, nq, t, nt, chunk, part);
  else if (W == 8) MB2_LAUNCH(ctx, k_hamming_2nn<8>, grid, 128, 0, q, nq, t, nt, chunk, part);
  else MB2_LAUNCH(ctx, k_hamming_2nn<16>, grid, 128, 0, q, nq, t, nt, chunk, part);
  MB2_LAUNCH(ctx, k_hamming_finish, (nq + 255) / 256, 256, 0, part, nq, n_chunks, max_distance, rows, accept);
}
void mb2_nn_finalize(mb2_ctx* ctx, const NNState& st, int nq, int nt, int nn, MatchRow* rows, int* accept) {
  MB2_LAUNCH(ctx, k_finalize, (nq + 255) / 256, 256, 0, st, nq, nt, nn, rows, accept);
}
void mb2_nn_pass_simt(mb2_ctx* ctx, int pass, const uint8_t* q, int nq, const uint8_t* t, int nt, const float* qn, const float* tn,
                      const NNState& st, const double* txy, double contr2) {
  const int chunk = 4096;
  dim3 grid((nq + 127) / 128, (nt + chunk - 1) / chunk);
  if (pass == 1) MB2_LAUNCH(ctx, k_nn_simt<1>, grid, 128, 0, q, nq, t, nt, qn, tn, st, txy, contr2, chunk);
  else MB2_LAUNCH(ctx, k_nn_simt<2>, grid, 128, 0, q, nq, t, nt, qn, tn, st, txy, contr2, chunk);
}
void mb2_nn_resolve(mb2_ctx* ctx, unsigned long long* keys, int nq, const uint8_t* q, const uint8_t* t, int nt, const float* qn, const float* tn) {
  if (nq > 0) MB2_LAUNCH(ctx, k_nn_resolve, (nq + 7) / 8, 256, 0, keys, nq, q, t, nt, qn, tn);
}
int mb2_nn_pass_tc(mb2_ctx* ctx, int pass, const void* q_bf16, int nq, int nq_pad, const void* t_bf16, int nt_pad, const float* qn,
                   const float* tn, const NNState& st, const double* txy, double contr2, int kmax) {
  CUtensorMap mq, mt;
  int rc = make_tmap(ctx, &mq, q_bf16, nq_pad);
  if (rc) return rc;
  rc = make_tmap(ctx, &mt, t_bf16, nt_pad);
  if (rc) return rc;
  const int n_qblocks = nq_pad / BM, total_tiles = nt_pad / BN;
  // aim for ~8 work items per SM so the persistent CTAs stay balanced
  int n_chunks = (8 * ctx->num_sms + n_qblocks - 1) / n_qblocks;
  if (n_chunks < 1) n_chunks = 1;
  if (n_chunks > total_tiles) n_chunks = total_tiles;
  const int tiles_per_chunk = (total_tiles + n_chunks - 1) / n_chunks;
  n_chunks = (total_tiles + tiles_per_chunk - 1) / tiles_per_chunk;
  const int n_items = n_qblocks * n_chunks;
  const int grid = n_items < ctx->num_sms ? n_items : ctx->num_sms;
  static unsigned long long attr_devs = 0;
  if (mb2_first_use_on_device(&attr_devs, ctx->device)) {
    cudaFuncSetAttribute(k_nn_tc<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaFuncSetAttribute(k_nn_tc<2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaFuncSetAttribute(k_nn_tc<1, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaFuncSetAttribute(k_nn_tc<2, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  }
  const char* ew_env = std::getenv("MB2_NN_EPI_WARPS");   // A/B switch for measurements
  const int ew = ew_env ? std::atoi(ew_env) : (pass == 1 ? 8 : 16);
#define MB2_NN_GO(P, E) MB2_LAUNCH(ctx, (k_nn_tc<P, E>), grid, 32 * (2 + E), SMEM_BYTES, mq, mt, nq, nt_pad, qn, tn, st, txy, contr2, tiles_per_chunk, n_qblocks, n_chunks, kmax)
  if (ew == 8) { if (pass == 1) MB2_NN_GO(1, 8); else MB2_NN_GO(2, 8); }
  else { if (pass == 1) MB2_NN_GO(1, 16); else MB2_NN_GO(2, 16); }
#undef MB2_NN_GO
  return MB2_OK;
}
