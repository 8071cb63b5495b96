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
                   const float* tn, const NNState& st, const double* txy, double contr2, int kmax);   // kmax = min(nn, nt) - 1 failers at most
// low word of every key: 32-column block (or any index inside it) -> exact lowest train index with the key's distance
void mb2_nn_resolve(mb2_ctx* ctx, unsigned long long* keys, int nq, const uint8_t* q, const uint8_t* t, int nt, const float* qn, const float* tn);
// Hamming 2-NN of binary descriptors (MatchFLANNDistance, matching.cpp:607-666): descriptors as rows of W (4 / 8 / 16) 32-bit words
void mb2_nn_hamming_words(mb2_ctx* ctx, const uint8_t* src, int n, int bytes, int W, uint32_t* dst);
int mb2_nn_hamming_chunks(mb2_ctx* ctx, int nq, int nt);
void mb2_nn_hamming(mb2_ctx* ctx, const uint32_t* q, int nq, const uint32_t* t, int nt, int W, int n_chunks, int max_distance,
                    unsigned long long* part, MatchRow* rows, int* accept);
