// LO-RANSAC homography (DEGENSAC, exp_ransacHcustom) with GPU-batched hypothesis scoring.
//
// Reference: degensac/exp_ranH.c:796-1236 (main loop, iter_type 4 = LSQ-before-LO + inner RANSAC),
// exp_inHranicustom (:741-793), exp_iterHcustom (:617-737); rtools.c (sample :12-23, randsubset
// :25-39, multirsampleT :136-156, inlidxs :160-171, nsamples :202-225, truncQuad :228-236,
// scoreLess :238-250); Htools.c (lin_hg :17-55, lin_hgN :57-96, u2h :98-130, all_Hori_valid
// :543-569); utools.c (normu :7-50, denormH :70-89, nullspace :97-167, cov_mat :170-185, det3
// :196-202); hash.c (SuperFastHash, htContains / htInsert).
//
// What is restructured, and why it is the same computation:
//  * The reference draws sample k with `srand(seed_k); 4 x random(); seed_{k+1} = rand()` and a pool
//    permutation that never depends on scores, so the hypothesis sequence is a pure function of the
//    initial seed.  We generate hypotheses in batches on the host with exactly those calls
//    (glibc random_r with a private state: same stream as srand/rand, but re-entrant), score a whole
//    batch in ONE launch (I = #inliers, J = MSAC sum per hypothesis), and then replay the reference's
//    sequential best-so-far / symmetric-check / LO / adaptive-stop logic over the scores.
//  * The reference keeps residual vectors in four rotating buffers (errs[0..3], errs[4] aliases one
//    of them).  A buffer is modelled as "the residuals of hypothesis h": it is only materialised
//    (one GPU launch) when the reference would actually read it (symmetric check, LO, final mask),
//    including the reference's aliasing quirk where errs[4] may point at a buffer that later samples
//    overwrite.
//  * The 9x9 symmetric eigenproblem of the normalised DLT (u2h) uses a cyclic Jacobi solver instead
//    of LAPACK dsyev (the reference links an unpinned system LAPACK): H agrees to ~1e-13 relative up
//    to sign, decisions agree unless a residual sits within that distance of a threshold.
#include "common.cuh"
#undef MB2_NS
#define MB2_NS mb2_ransac_host_detail
#include "ransac.cuh"
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#include "ransac_common.hpp"

namespace MB2_NS {
using namespace mb2_ransac_common;

// ---------------------------------------------------------------------------------------------
struct Trace {  // MB2_RANSAC_TRACE=1: where the wall time of one mb2_ransac_h call goes (stderr)
  bool on = std::getenv("MB2_RANSAC_TRACE") != nullptr;
  double t[6] = {0, 0, 0, 0, 0, 0}; int n[6] = {0, 0, 0, 0, 0, 0};
  static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
};
struct TraceScope {
  Trace& T; int k; double t0;
  TraceScope(Trace& T, int k) : T(T), k(k), t0(T.on ? Trace::now() : 0) {}
  ~TraceScope() { if (T.on) { T.t[k] += Trace::now() - t0; T.n[k]++; } }
};

struct RansacH {
  mb2_ctx* ctx;
  Trace tr;
  const double* u; int len; double th; int which, doSymCheck;
  const double* d_u = nullptr;
  // residual buffers: errs[0..3] are buffer ids, errs[4] aliases one of them
  struct Buf { std::vector<double> d; double h[9]; bool valid = false; bool tagged = false; };
  Buf bufs[4];
  int errs[5] = {0, 1, 2, 3, 3};
  HashTable ht;
  LibcRand rng;
  int rc = MB2_OK;

  // ---- GPU services
  int gpu_resid(int which_, const double* h, double* out, Score* S, double th_) {
    TraceScope ts(tr, 0);
    int I = 0; double J = 0;
    int r = mb2_score_models(ctx, which_, d_u, len, h, 1, th_, out, &I, &J);
    if (r < 0) { rc = r; return r; }
    if (S) { S->I = (unsigned)I; S->J = J; }
    return MB2_OK;
  }
  // eager "HDS1(Z, u, h, d, len)" into buffer id b
  void eval_into(int b, const double* h) {
    Buf& B = bufs[b];
    B.d.resize(len);
    gpu_resid(which, h, B.d.data(), nullptr, th);
    std::memcpy(B.h, h, sizeof B.h); B.valid = true; B.tagged = true;
  }
  // lazy form used by the main sampling loop
  void tag(int b, const double* h) { Buf& B = bufs[b]; std::memcpy(B.h, h, sizeof B.h); B.valid = false; B.tagged = true; }
  const double* data(int b) {
    Buf& B = bufs[b];
    if (!B.valid) {
      B.d.assign(len, 0.0);
      if (B.tagged) gpu_resid(which, B.h, B.d.data(), nullptr, th);
      B.valid = true;
    }
    return B.d.data();
  }
  bool sym_check_bad(const double* h) {  // exp_ranH.c:917-927
    Score S = {0, 0};
    gpu_resid(1 /*HDsSym*/, h, nullptr, &S, CHECK_COEF * th);
    return S.I <= (unsigned)MIN_GOOD_SYM_PTS;
  }

  // rtools.c:25-39
  int* randsubset(int* pool, int max_sz, int siz) {
    for (int i = 0; i < siz; i++) {
      int s = rng.next() % (max_sz - i), j = max_sz - i - 1;
      int q = pool[s]; pool[s] = pool[j]; pool[j] = q;
    }
    return pool + max_sz - siz;
  }

  // exp_ranH.c:617-737
  Score iterH(int* inliers, double ths, int steps, double* H, int iterID) {
    int dbuf = errs[1];
    double h[9] = {0};
    Score maxS = {0, 0}, S = {0, 0}, Ss;
    const double dth = (ths - th) / (steps);
    maxS = inlidxs(data(errs[4]), len, th, inliers);
    if (maxS.I < 4) return S;
    S.I = inlidxs_count(data(errs[4]), len, th * MWM, inliers);
    { TraceScope ts(tr, 1); u2h(u, inliers, S.I, h); }  // __D3__ with D3_H_RATIO 1 and an unlimited inlLimit: all inliers
    for (int it = 0; it < steps; it++) {
      eval_into(dbuf, h);
      Ss = inlidxs(bufs[dbuf].d.data(), len, th, inliers);
      const uint32_t hash = SuperFastHash((const char*)inliers, Ss.I * sizeof(*inliers));
      const int ret = ht.contains(hash, Ss.I, iterID);
      if (ret != -1 && ret != iterID) { S.I = 0; S.J = 0; return S; }
      if (ret == -1) ht.insert(hash, Ss.I, iterID);
      S.I = inlidxs_count(bufs[dbuf].d.data(), len, ths * MWM, inliers);
      if (scoreLess(maxS, Ss)) {
        maxS = Ss;
        errs[1] = errs[0]; errs[0] = dbuf; dbuf = errs[1];
        std::memcpy(H, h, 9 * sizeof(double));
      }
      if (S.I < 4) return maxS;
      { TraceScope ts(tr, 1); u2h(u, inliers, S.I, h); }
      ths -= dth;
    }
    eval_into(dbuf, h);
    S = inlidxs(bufs[dbuf].d.data(), len, th, inliers);
    if (scoreLess(maxS, S)) {
      maxS = S;
      errs[1] = errs[0]; errs[0] = dbuf;
      std::memcpy(H, h, 9 * sizeof(double));
    }
    return maxS;
  }

  // exp_ranH.c:741-793
  Score inHrani(int* inliers, int ninl, double* H, int rep, int* iterID) {
    Score S, maxS = {0, 0};
    double h[9] = {0};
    std::vector<int> intbuff(len);
    if (ninl < 8) return maxS;
    int ssiz = ninl / 2;
    if (ssiz > 12) ssiz = 12;
    std::swap(errs[2], errs[0]);
    for (int i = 0; i < rep; i++) {
      int* sample = randsubset(inliers, ninl, ssiz);
      u2h(u, sample, ssiz, h);
      eval_into(errs[0], h);
      errs[4] = errs[0];
      S = iterH(intbuff.data(), TC * th, ILSQ_ITERS, h, ++*iterID);
      if (scoreLess(maxS, S)) {
        maxS = S;
        std::swap(errs[2], errs[0]);
        std::memcpy(H, h, 9 * sizeof(double));
      }
    }
    std::swap(errs[2], errs[0]);
    return maxS;
  }

  // case 4 of the iteration switch (exp_ranH.c:1012-1035 / 1151-1174)
  Score local_optimisation(int* inliers, double* h, int* iterID) {
    const int d = errs[0];
    unsigned I = inlidxs_count(data(errs[4]), len, TC * th * MWM, inliers);
    { TraceScope ts(tr, 1); u2h(u, inliers, I, h); }
    eval_into(d, h);
    I = inlidxs_count(bufs[d].d.data(), len, th, inliers);
    return inHrani(inliers, I, h, RAN_REP, iterID);
  }
};

}  // namespace
using namespace MB2_NS;

extern "C" int mb2_ransac_h(mb2_ctx* ctx, const double* u, int len, double th, double conf, int max_sam, int errorType, int doSymCheck,
                            long seed, double* H, unsigned char* inl, int* data_out, double* Jout) {
  if (!ctx || !u || len < 0 || !H || !inl) return MB2_ERR_ARG;
  cudaSetDevice(ctx->device);
  for (int i = 0; i < 9; i++) H[i] = 0;
  for (int i = 0; i < len; i++) inl[i] = 0;
  if (data_out) data_out[0] = data_out[1] = data_out[2] = 0;
  if (Jout) *Jout = 0;
  if (len < 4) return 0;
  if (mb2_is_device_ptr(u)) { ctx->set_error("ransac_h: u must be a host pointer (the LO logic runs on the host)"); return MB2_ERR_ARG; }

  RansacH R;
  R.ctx = ctx; R.u = u; R.len = len; R.th = th; R.doSymCheck = doSymCheck;
  R.which = errorType == 0 ? 0 : (errorType == 1 ? 2 : 1);  // Sampson -> HDs, SymmMax -> HDsSymMax, SymmSum -> HDsSym
  // correspondences stay resident on the device for the whole run
  DevBuf& d_u_buf = ctx->rs_u;
  if (d_u_buf.reserve((size_t)len * 48) != cudaSuccess) { ctx->set_error("ransac_h: cudaMalloc failed"); return MB2_ERR_CUDA; }
  if (cudaMemcpyAsync(d_u_buf.p, u, (size_t)len * 48, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return MB2_ERR_CUDA;
  R.d_u = d_u_buf.as<double>();

  std::vector<int> pool(len), inliers(len);
  for (int i = 0; i < len; i++) pool[i] = i;
  int* samidx = pool.data() + len - 4;
  int bestsamidx[4] = {0, 0, 0, 0};
  double sol[81], *h = sol, M[81];
  int nullspace_buff[18];
  Score maxS = {0, 0}, maxSs = {0, 0}, S = {0, 0};
  int no_sam = 0, iter_cnt = 0, no_rej = 0, iterID = 0;
  bool new_max = false, bad_model = false;
  std::memset(sol, 0, sizeof sol);

  R.rng.seed((unsigned)seed);           // srand(time(NULL)) in the reference (exp_ranH.c:823)
  unsigned cur_seed = (unsigned)R.rng.next();  // seed = rand()

  // ---- batched hypothesis generation -------------------------------------------------------------
  struct Hyp { double h[9]; unsigned seed_before; int sam[4]; int status; /* 0 ok, 1 orientation reject, 2 skipped */ };
  std::vector<Hyp> batch;
  std::vector<double> models;
  std::vector<int> bI; std::vector<double> bJ;
  size_t bpos = 0;
  int batch_size = 64;
  auto refill = [&](int want) -> int {
    TraceScope ts(R.tr, 2);
    batch.clear(); models.clear(); bpos = 0;
    for (int k = 0; k < want; k++) {
      Hyp hy; hy.status = 0; hy.seed_before = cur_seed;
      R.rng.seed(cur_seed);
      // multirsampleT(Z, 9, 2, pool, 4, len, M): 4 x sample() (rtools.c:12-23), rows of Z into M
      for (int i = 0; i < 4; i++) {
        int s = R.rng.next() % (len - i), j = len - i - 1;
        int q = pool[s]; pool[s] = pool[j]; pool[j] = q;
        lin_rows(u, q, M + (2 * i) * 9, M + (2 * i + 1) * 9);
      }
      cur_seed = (unsigned)R.rng.next();
      std::memcpy(hy.sam, samidx, sizeof hy.sam);
      if (!all_Hori_valid(u, samidx)) hy.status = 1;
      else {
        for (int i = 72; i < 81; ++i) M[i] = 0.0;
        double hs[81];
        std::memset(hs, 0, sizeof hs);
        int nullsize = nullspace(M, hs, 9, nullspace_buff);
        if (nullsize != 1) hy.status = 2;
        else {
          double v = det3(hs), tol = hs[8];
          if (tol == 0) { for (int i = 0; i < 9; ++i) tol += hs[i] * hs[i]; tol = std::sqrt(tol); tol *= 0.001; }
          tol = tol * tol * tol;
          if (std::fabs(v / tol) < 10e-2) hy.status = 2;
          else std::memcpy(hy.h, hs, sizeof hy.h);
        }
      }
      if (hy.status == 0) models.insert(models.end(), hy.h, hy.h + 9);
      batch.push_back(hy);
    }
    const int K = (int)(models.size() / 9);
    bI.assign(std::max(K, 1), 0); bJ.assign(std::max(K, 1), 0.0);
    if (K > 0) {
      TraceScope ts2(R.tr, 3);
      int r = mb2_score_models(ctx, R.which, R.d_u, len, models.data(), K, th, nullptr, bI.data(), bJ.data());
      if (r < 0) return r;
    }
    return MB2_OK;
  };

  auto tol_of = [](const double* hh) {
    double tol = hh[8];
    if (tol == 0) { for (int i = 0; i < 9; ++i) tol += hh[i] * hh[i]; tol = std::sqrt(tol); tol *= 0.001; }
    return tol * tol * tol;
  };

  size_t model_idx = 0;
  unsigned last_seed_before = cur_seed;
  bool any_sample = false;
  while (no_sam < max_sam) {
    if (bpos >= batch.size()) {
      int want = std::min(batch_size, std::max(1, max_sam - no_sam));
      int r = refill(want);
      if (r < 0) return r;
      model_idx = 0;
      if (batch_size < 4096) batch_size *= 2;
    }
    const Hyp& hy = batch[bpos++];
    no_sam++;
    last_seed_before = hy.seed_before; any_sample = true;
    if (hy.status == 1) { no_rej++; continue; }
    if (hy.status == 2) continue;
    std::memcpy(h, hy.h, 9 * sizeof(double));
    const int dbuf = R.errs[0];
    R.tag(dbuf, h);                                   // d = errs[0]; HDS1(Z, u, h, d, len)
    S.I = (unsigned)bI[model_idx]; S.J = bJ[model_idx]; model_idx++;
    // The batch scorer's J is a tree sum whose partition depends on the batch; the reference (and every LO model here) sums the MSAC
    // terms serially.  Scores are ordered by J alone (scoreLess, __SCORE__ == SC_M), so a hypothesis that could win or tie -- J within
    // 1e-9 of a best-so-far or above it -- gets the reference's serial sum from its residuals before it is compared: the bests always
    // carry serial sums, and a sample drawn twice (identical residuals) ties exactly instead of flipping on the last bit.
    if (S.J >= std::min(maxS.J, maxSs.J) * (1.0 - 1e-9)) {
      const double* dd = R.data(dbuf);
      if (R.rc < 0) return R.rc;
      double Js = 0;
      for (int i = 0; i < len; ++i) Js += truncQuad(dd[i], th);
      S.J = Js;
    }
    bool do_iterate;
    if (scoreLess(maxS, S)) {
      if (doSymCheck) bad_model = R.sym_check_bad(h);
      if (bad_model) continue;
      R.errs[0] = R.errs[3]; R.errs[3] = dbuf;
      maxS = S; new_max = true;
      std::memcpy(H, h, 9 * sizeof(double));
    }
    if (scoreLess(maxSs, S)) {
      do_iterate = no_sam > ITER_SAM;
      maxSs = S;
      R.errs[4] = dbuf;
      std::memcpy(bestsamidx, hy.sam, sizeof bestsamidx);
    } else do_iterate = false;
    if ((no_sam >= ITER_SAM) && (iter_cnt == 0) && (maxSs.I > 4)) do_iterate = true;
    if (do_iterate) {
      iter_cnt++;
      // the reference's generator state here: srand(seed_k); 4 draws; 1 draw for the next seed
      R.rng.seed(hy.seed_before);
      for (int i = 0; i < 5; i++) R.rng.next();
      { TraceScope ts(R.tr, 4); S = R.local_optimisation(inliers.data(), h, &iterID); }
      if (R.rc < 0) return R.rc;
      if (scoreLess(maxS, S) && (std::fabs(det3(h) / tol_of(h)) > 10e-2)) {
        if (doSymCheck) bad_model = R.sym_check_bad(h);
        if (!bad_model) {
          const int d0 = R.errs[0];
          R.errs[0] = R.errs[3]; R.errs[3] = d0;
          maxS = S; new_max = true;
          std::memcpy(H, h, 9 * sizeof(double));
        }
      }
    }
    if (new_max) {
      const int new_sam = nsamples(maxS.I + 1, len, 4, conf);
      if (new_sam < max_sam) max_sam = new_sam;
      new_max = false;
    }
  }
  if (iter_cnt == 0) {  // "If there were no LOs, do at least one NOW!" (exp_ranH.c:1124-1204)
    iter_cnt++;
    if (any_sample) { R.rng.seed(last_seed_before); for (int i = 0; i < 5; i++) R.rng.next(); }
    S = R.local_optimisation(inliers.data(), h, &iterID);
    if (R.rc < 0) return R.rc;
    if (scoreLess(maxS, S) && (std::fabs(det3(h) / tol_of(h)) > 10e-2)) {
      if (doSymCheck) bad_model = R.sym_check_bad(h);
      if (!bad_model) {
        const int d0 = R.errs[0];
        R.errs[0] = R.errs[3]; R.errs[3] = d0;
        maxS = S;
        std::memcpy(H, h, 9 * sizeof(double));
      }
    }
  }
  const double* d = R.data(R.errs[3]);
  if (R.rc < 0) return R.rc;
  for (int j = 0; j < len; j++) inl[j] = (maxS.J > 0 || maxS.I > 0) ? (d[j] <= th ? 1 : 0) : 0;
  if (data_out) { data_out[0] = no_sam; data_out[1] = iter_cnt; data_out[2] = no_rej; }
  if (R.tr.on)
    std::fprintf(stderr, "[mb2_ransac_h] len %d samples %d LOs %d I %u | resid %d x %.3f ms | u2h %d x %.3f ms | refill %d: %.2f ms (score %.2f) | LO total %.2f ms\n",
                 len, no_sam, iter_cnt, maxS.I, R.tr.n[0], R.tr.n[0] ? R.tr.t[0] / R.tr.n[0] : 0., R.tr.n[1], R.tr.n[1] ? R.tr.t[1] / R.tr.n[1] : 0.,
                 R.tr.n[2], R.tr.t[2], R.tr.t[3], R.tr.t[4]);
  if (Jout) *Jout = maxS.J;
  return (int)maxS.I;
}
