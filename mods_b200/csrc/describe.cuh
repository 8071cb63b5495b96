// Records shared between describe.cu and capi.cu.
#pragma once
#include "common.cuh"
#include "pyramid.cuh"

// Gaussian taps of the per-region pre-blur, indexed by m = int(ceil(s*mrSize)):
// sigma = 1.5f * float(2m+1)/41, kernel as oracle/cvmath.h / OpenCV 2.4.9 getGaussianKernel.
struct TapTable {
  const int* n;      // taps per m (0 when the region takes the direct path)
  const int* off;    // offset into w
  const float* w;
  int max_m;
};

// Size classes of the patch extraction (describe.cu): upper bounds on m = ceil(s * mrSize) of the three dense
// shared-memory classes and the two needed-columns shared-memory classes; larger regions use global scratch.
constexpr int MB2_CLS_M_DENSE_S = 23, MB2_CLS_M_DENSE_M = 31, MB2_CLS_M_DENSE_L = 40, MB2_CLS_M_NEED_S = 54, MB2_CLS_M_NEED_L = 73;
constexpr int MB2_N_CLASSES = 6;
struct ExtractPlan {
  const unsigned long long* sorted = nullptr;   // keys sorted by (class, descending m, index); low 32 bits = region index
  int count[MB2_N_CLASSES] = {0, 0, 0, 0, 0, 0};
};

int mb2_describe_plan(mb2_ctx* ctx, const KeyOut* kps, int n, const DescribeParams& dp, int max_m, unsigned long long* d_need,
                      int* d_too_big, unsigned long long* d_sum_p2sq, unsigned long long* d_keys, int* d_cls_cnt);
int mb2_launch_describe_kernel(mb2_ctx* ctx, const ImgView& img, const KeyOut* kps, int n, const DescribeParams& dp,
                               const DescTables* d_tables, const TapTable& taps, const int* tap_n_host, const ExtractPlan& plan,
                               const unsigned long long* d_off, float* d_scratch,
                               uint8_t* d_desc, float* d_patches /* n x 41 x 41, required */, float2* d_stats /* n */,
                               double* d_vecT /* 128 x n */, float2* d_rec /* n x 1681 gradient records */);
