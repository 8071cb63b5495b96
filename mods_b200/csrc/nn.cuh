// Records shared between nn.cu and capi.cu.
#pragma once
#include "common.cuh"

struct NNState {               // one entry per query
  unsigned long long* best0;   // pass 1: (d, idx) of the nearest train
  unsigned long long* best1;   // pass 2: nearest failer other than idx0
  unsigned long long* bestP;   // pass 2: nearest passer
  int* cnt;                    // pass 2: number of failers (incl. idx0)
  int* incons;                 // pass 2: some failer is geometrically inconsistent with idx0
  float* d0; int* idx0; float* thr; float* thr_rel;
};

struct MatchRow { int q, idx0, idxJ, idx1; float d0, dJ, d1; int pad; };

void mb2_nn_prepare(mb2_ctx* ctx, const uint8_t* d_desc, int n, int n_pad, void* d_bf16, float* d_norms, float pad_norm);
void mb2_nn_init_state(mb2_ctx* ctx, const NNState& st, int nq);
void mb2_nn_threshold(mb2_ctx* ctx, const NNState& st, int nq, const float* qn, double sqminratio);
void mb2_nn_finalize(mb2_ctx* ctx, const NNState& st, int nq, int nt, int nn, MatchRow* rows, int* accept);
void mb2_nn_topk(mb2_ctx* ctx, const uint8_t* q, int nq, const uint8_t* t, int nt, const float* qn, const float* tn, const double* txy, double contr2,
                 int nn, MatchRow* rows, int* accept);   // matchRatio >= 1 (matching.cpp:397-428): exact sorted k-NN table per query, nn <= 64
void mb2_nn_pass_simt(mb2_ctx* ctx, int pass, const uint8_t* q, int nq, const uint8_t* t, int nt, const float* qn, const float* tn,
                      const NNState& st, const double* txy, double contr2);
int mb2_nn_pass_tc(mb2_ctx* ctx, int pass, const void* q_bf16, int nq, int nq_pad, const void* t_bf16, int nt_pad, const float* qn,
                   const float* tn, const NNState& st, const double* txy, double contr2);
