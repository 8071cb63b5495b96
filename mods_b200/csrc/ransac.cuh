#pragma once
#include "common.cuh"
__host__ __device__ bool mb2_minv3(double* a);
void mb2_launch_score(mb2_ctx* ctx, int which, const double* d_u, int len, const double* d_models, int K, double th, double* d_resid,
                      int* d_I, double* d_J, void* d_partials);
size_t mb2_score_partial_bytes(int len, int K);
