// View-sharded pair driver (BASELINE config C4): the synthesised views of both images -- the reference's own unit of
// parallelism (`#pragma omp parallel for` over views, imagerepresentation.cpp:621) -- are dealt out over the ranks of ONE node, one
// process per GPU.  Every rank detects / describes its views (mb2_detect_describe_synth_view) and leaves the regions on its device as
// 184-byte records; ONE ncclAllGather of the padded record buffers (plus an all-gather of the per-unit counts) gives every rank all
// regions; they are put in the reference's order -- detector, then view index, then detection order (imagerepresentation.cpp:2044-2045)
// -- by device-to-device copies; the N1 x N2 matching of every detector is split by query rows (mb2_match_slots_range); the tentative
// rows are all-gathered; rank 0 runs DuplicateFiltering + LORANSACFiltering (mb2_host_verify).  Nothing on this path goes through host
// numpy; the host sees counts, the tentative rows and (rank 0) the 7-double frames of the regions for the verification.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: in a process that imported torch this is torch's own copy), so libmods_host.so has
// no link-time dependency on it and single-GPU users never load it.
#include "mods_host.hpp"

#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {
// ---- the five NCCL entry points, restated from nccl.h (2.x ABI) ------------------------------------------------------------------
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void* ncclComm_h;
typedef int (*fn_GetUniqueId)(ncclUniqueId_t*);
typedef int (*fn_CommInitRank)(ncclComm_h*, int, ncclUniqueId_t, int);
typedef int (*fn_CommDestroy)(ncclComm_h);
typedef int (*fn_AllGather)(const void*, void*, size_t, int /*ncclDataType_t*/, ncclComm_h, void* /*cudaStream_t*/);
typedef const char* (*fn_GetErrorString)(int);
struct Nccl {
  void* lib = nullptr;
  fn_GetUniqueId GetUniqueId = nullptr; fn_CommInitRank CommInitRank = nullptr; fn_CommDestroy CommDestroy = nullptr;
  fn_AllGather AllGather = nullptr; fn_GetErrorString GetErrorString = nullptr;
  bool load() {
    if (lib) return true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) { lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) return false;
    GetUniqueId = (fn_GetUniqueId)dlsym(lib, "ncclGetUniqueId"); CommInitRank = (fn_CommInitRank)dlsym(lib, "ncclCommInitRank");
    CommDestroy = (fn_CommDestroy)dlsym(lib, "ncclCommDestroy"); AllGather = (fn_AllGather)dlsym(lib, "ncclAllGather");
    GetErrorString = (fn_GetErrorString)dlsym(lib, "ncclGetErrorString");
    return GetUniqueId && CommInitRank && CommDestroy && AllGather;
  }
};
Nccl g_nccl;
const int NCCL_INT8 = 0;   // ncclInt8 / ncclChar (nccl.h: ncclDataType_t)

struct DBuf {   // device scratch of one call (through the C ABI: the host library does not link the CUDA runtime)
  mb2_ctx* ctx = nullptr; void* p = nullptr; size_t cap = 0;
  bool reserve(size_t n) {
    if (n <= cap) return true;
    if (p) mb2_dev_free(ctx, p);
    p = mb2_dev_alloc(ctx, n + 256); cap = p ? n + 256 : 0;
    return p != nullptr;
  }
  ~DBuf() { if (p) mb2_dev_free(ctx, p); }
};
double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// Device scratch of mb2_views_sharded_pair, kept per primary context between calls (grow-only): a call moves ~1 GB through these
// buffers at C4, and cudaMalloc / cudaFree of that much memory cost up to several hundred ms per call when they were per-call objects.
struct ShardScratch { DBuf send, recv, d_counts, ordered[4], d_rows, d_img[2], wbuf[4]; };
std::mutex g_scratch_mutex;
std::unordered_map<mb2_ctx*, ShardScratch*> g_scratch;
ShardScratch* scratch_of(mb2_ctx* ctx) {
  std::lock_guard<std::mutex> lk(g_scratch_mutex);
  ShardScratch*& s = g_scratch[ctx];
  if (!s) s = new ShardScratch();
  return s;
}
}  // namespace

// called by mb2_mods_release (before the helper contexts go away: worker buffers belong to them)
extern "C" void mb2_sharded_release(mb2_ctx* ctx) {
  std::lock_guard<std::mutex> lk(g_scratch_mutex);
  auto it = g_scratch.find(ctx);
  if (it == g_scratch.end()) return;
  delete it->second;
  g_scratch.erase(it);
}

extern "C" {

// ---- unit plan (plain C, checked without a GPU by tests/test_sharding_gloo.py) --------------------------------------------------------
// Pixel count of a synthesised view: tilt t shrinks one side by 1 / t, zoom z both sides by z (synth-detection.cpp:301-342).
double mb2_shard_view_cost(int w, int h, double tilt, double zoom) { return ((double)w * zoom) * ((double)h * zoom) / std::max(std::fabs(tilt), 1e-9); }

// Longest-processing-time-first over `world` ranks, deterministic (ties: lower unit, lower rank): owner[u] = rank of unit u.
void mb2_shard_assign(const double* costs, int n_units, int world, int* owner) {
  std::vector<int> order(n_units);
  for (int i = 0; i < n_units; i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return costs[a] > costs[b]; });
  std::vector<double> load(std::max(world, 1), 0.0);
  for (int u : order) {
    int r = 0;
    for (int k = 1; k < world; k++) if (load[k] < load[r]) r = k;
    owner[u] = r; load[r] += costs[u];
  }
}

// Where the records of every unit sit in the all-gathered buffer (world blocks of `stride` records; a rank packs its units in ascending
// unit order): src_off[u] in records.  Returns the stride = the largest per-rank total (at least 1).
int mb2_shard_layout(const int* owner, const int* counts, int n_units, int world, int* src_off) {
  std::vector<long long> tot(std::max(world, 1), 0);
  for (int u = 0; u < n_units; u++) tot[owner[u]] += counts[u];
  long long stride = 1;
  for (long long t : tot) stride = std::max(stride, t);
  std::vector<long long> run(std::max(world, 1), 0);
  for (int u = 0; u < n_units; u++) { src_off[u] = (int)(owner[u] * stride + run[owner[u]]); run[owner[u]] += counts[u]; }
  return (int)stride;
}

// ---- communicator ----------------------------------------------------------------------------------------------------------------
int mb2_dist_unique_id(unsigned char* id128) {
  if (!id128 || !g_nccl.load()) return MB2_ERR_UNSUPPORTED;
  ncclUniqueId_t id;
  if (g_nccl.GetUniqueId(&id) != 0) return MB2_ERR_CUDA;
  std::memcpy(id128, id.internal, 128);
  return MB2_OK;
}
int mb2_dist_comm_create(mb2_ctx* ctx, int rank, int world, const unsigned char* id128, void** comm) {
  if (!ctx || !comm || !id128 || rank < 0 || rank >= world || !g_nccl.load()) return MB2_ERR_ARG;
  if (mb2_ctx_make_current(ctx) != MB2_OK) return MB2_ERR_CUDA;   // the communicator lives on the context's device
  ncclUniqueId_t id; std::memcpy(id.internal, id128, 128);
  ncclComm_h c = nullptr;
  if (g_nccl.CommInitRank(&c, world, id, rank) != 0) return MB2_ERR_CUDA;
  *comm = c;
  return MB2_OK;
}
void mb2_dist_comm_destroy(void* comm) { if (comm && g_nccl.lib) g_nccl.CommDestroy((ncclComm_h)comm); }

// ---- one pair, views sharded over the ranks ------------------------------------------------------------------------------------------
// img1 / img2: gray f32 [H|D], the SAME on every rank.  comm: mb2_dist_comm_create (may be NULL when world == 1).  The result is
// complete on rank 0 (other ranks fill regions / tentatives only).  digest (optional, 4 x u64): order-sensitive checksums of the
// gathered region records and of the tentative rows -- identical on every rank and for every world size, which is what the tests
// and bench.py assert.  stats (optional, 8 doubles): ms views, ms gather, ms match, ms tentative gather, ms verify, all-gather payload
// bytes per rank, regions of both images, units of this rank.
}  // extern "C"
namespace {
// What the verification stage needs of a pair, on the host: the 14-double frames of the tentatives and their sort keys.
struct ShardFront { int T = 0; std::vector<double> frames, key; double t_start = 0, t_gather = 0, t_match = 0; };

// Everything up to and including the tentative exchange; `verifier` = this rank will verify the pair (its frames are fetched).
int sharded_front(mb2_ctx* ctx, void* comm, int rank, int world, const float* img1, int w1, int h1, const float* img2, int w2, int h2,
                  const mb2_pair_config* cfg, mb2_pair_result* res, unsigned long long* digest, double* stats, bool verifier, ShardFront& F) {
  if (!ctx || !img1 || !img2 || !cfg || !res || world < 1 || rank < 0 || rank >= world || (world > 1 && (!comm || !g_nccl.load()))) return MB2_ERR_ARG;
  std::memset(res, 0, sizeof *res);
  void* st = mb2_ctx_stream(ctx);
  const double t_start = now_ms();
  // ---- units: (image, detector, view), detector-major inside an image as RegionVectorMap orders them
  struct Unit { int image, det; mb2_view_params vp; };
  std::vector<Unit> units;
  const mb2_view_params ident{1.0, 0.0, 1.0, 0.5, 1};
  for (int im = 0; im < 2; im++)
    for (int det = 0; det < (cfg->use_mser ? 2 : 1); det++) {
      const int n = det == 0 ? cfg->n_hess_views : cfg->n_mser_views;
      const mb2_view_params* v = det == 0 ? cfg->hess_views : cfg->mser_views;
      if (n <= 0) units.push_back(Unit{im, det, ident});
      for (int i = 0; i < n && i < MB2_MAX_PAIR_VIEWS; i++) units.push_back(Unit{im, det, v[i]});
    }
  const int U = (int)units.size();
  std::vector<double> costs(U);
  for (int u = 0; u < U; u++) costs[u] = mb2_shard_view_cost(units[u].image ? w2 : w1, units[u].image ? h2 : h1, units[u].vp.tilt, units[u].vp.zoom);
  std::vector<int> owner(U), counts(U, 0);
  mb2_shard_assign(costs.data(), U, world, owner.data());

  // ---- my views.  A view pipeline alone leaves the GPU idle between its kernels (host legs, count read-backs: 35 % of the 693 ms the
  // 176 views of C4 took one after the other on one GPU), so the units of this rank are run by several host threads, each on its
  // own context (stream) of the device, longest view first; every worker packs its regions into its own device buffer, and the send
  // buffer is assembled from them in ascending unit order -- the layout mb2_shard_layout describes -- by device-to-device copies.
  const int REC = MB2_REGION_RECORD_BYTES;
  ShardScratch& S = *scratch_of(ctx);
  DBuf &send = S.send, &recv = S.recv, &d_counts = S.d_counts, &d_rows = S.d_rows;
  DBuf* ordered = S.ordered; DBuf* d_img = S.d_img;
  for (DBuf* b : {&send, &recv, &d_counts, &ordered[0], &ordered[1], &ordered[2], &ordered[3], &d_rows, &d_img[0], &d_img[1]}) b->ctx = ctx;
  // host images go to the device once (every view call would stage its own copy otherwise)
  const float* dimg[2] = {img1, img2};
  for (int im = 0; im < 2; im++) {
    const float* src = im ? img2 : img1;
    if (mb2_is_device_pointer(src)) continue;
    const size_t bytes = (size_t)(im ? w2 : w1) * (im ? h2 : h1) * 4;
    if (!d_img[im].reserve(bytes)) return MB2_ERR_CUDA;
    mb2_dev_copy(ctx, d_img[im].p, src, bytes, 0);
    dimg[im] = (const float*)d_img[im].p;
  }
  if (mb2_ctx_sync(ctx) != MB2_OK) return MB2_ERR_CUDA;
  std::vector<int> mine;
  for (int u = 0; u < U; u++) if (owner[u] == rank) mine.push_back(u);
  const int my_units = (int)mine.size();
  std::stable_sort(mine.begin(), mine.end(), [&](int a, int b) { return costs[a] > costs[b]; });
  int n_workers = world >= 4 ? 2 : 4;   // host threads are the limit when 8 ranks share one node's cores
  if (const char* e = getenv("MB2_VIEW_WORKERS")) n_workers = std::max(1, std::min(4, atoi(e)));
  if (mb2_ctx_profiling(ctx)) n_workers = 1;   // per-kernel profiling keeps everything on one stream
  n_workers = std::max(1, std::min(n_workers, my_units));
  struct Worker { mb2_ctx* c = nullptr; DBuf* bufp = nullptr; size_t cap = 0, used = 0; int rc = MB2_OK; };
  std::vector<Worker> workers(n_workers);
  for (int wk = 0; wk < n_workers; wk++) {
    workers[wk].c = wk == 0 ? ctx : mb2_mods_sibling(ctx, wk - 1);
    if (!workers[wk].c) { n_workers = wk; break; }
    workers[wk].bufp = &S.wbuf[wk];
    workers[wk].bufp->ctx = workers[wk].c;
  }
  workers.resize(n_workers);
  std::vector<int> unit_worker(U, -1);
  std::vector<size_t> unit_off(U, 0);
  auto run = [&](int wk) {
    Worker& W = workers[wk];
    DBuf& buf = *W.bufp;
    if (!buf.reserve(((size_t)std::max(w1 * h1, w2 * h2) / 16 + 65536) * REC)) { W.rc = MB2_ERR_CUDA; return; }
    W.cap = (buf.cap - 256) / REC;
    // static deal of the cost-sorted units (worker wk takes every n_workers-th): the same worker meets the same views in every call,
    // so its context's buffers stop growing after the first call
    for (int i = wk; i < my_units; i += n_workers) {
      const int u = mine[i];
      const Unit& un = units[u];
      const int n = mb2_detect_describe_synth_view(W.c, dimg[un.image], un.image ? w2 : w1, un.image ? h2 : h1, &un.vp, un.det == 0 ? 0 : 3, &cfg->det,
                                                   &cfg->mser, &cfg->ori, &cfg->desc, MB2_MAX_SLOTS - 1, 0, nullptr, nullptr, nullptr, 0);
      if (n < 0) { W.rc = n; break; }
      if (W.used + (size_t)n > W.cap) {   // grow, keeping what is already packed
        const size_t ncap = std::max(W.cap * 2, W.used + (size_t)n);
        void* np_ = mb2_dev_alloc(W.c, ncap * REC + 256);
        if (!np_) { W.rc = MB2_ERR_CUDA; break; }
        mb2_dev_copy(W.c, np_, buf.p, W.used * REC, 2);
        mb2_ctx_sync(W.c);
        mb2_dev_free(W.c, buf.p); buf.p = np_; buf.cap = ncap * REC + 256; W.cap = ncap;
      }
      const int k = mb2_view_pack(W.c, (unsigned char*)buf.p + W.used * REC, (int)(W.cap - W.used));
      if (k < 0) { W.rc = k; break; }
      counts[u] = k; unit_worker[u] = wk; unit_off[u] = W.used; W.used += (size_t)k;
    }
    if (mb2_ctx_sync(W.c) != MB2_OK && W.rc >= 0) W.rc = MB2_ERR_CUDA;
  };
  {
    std::vector<std::thread> th;
    for (int wk = 1; wk < n_workers; wk++) th.emplace_back(run, wk);
    run(0);
    for (auto& t : th) t.join();
  }
  size_t used = 0;
  for (Worker& W : workers) { if (W.rc < 0) return W.rc; used += W.used; }
  if (!send.reserve(std::max<size_t>(used, 1) * REC)) return MB2_ERR_CUDA;
  size_t send_cap = (send.cap - 256) / REC;   // the buffer is kept between calls: it may be larger than this call needs
  {
    size_t o = 0;
    for (int u = 0; u < U; u++) {
      if (owner[u] != rank || counts[u] <= 0) continue;
      mb2_dev_copy(ctx, (unsigned char*)send.p + o * REC, (const unsigned char*)workers[unit_worker[u]].bufp->p + unit_off[u] * REC, (size_t)counts[u] * REC, 2);
      o += (size_t)counts[u];
    }
  }
  const double t_views = now_ms();

  // ---- counts of every unit on every rank (each rank contributes its own units, zeros elsewhere), then ONE all-gather of the records
  std::vector<int> all_counts(counts);
  if (world > 1) {
    if (!d_counts.reserve((size_t)U * 4 * (world + 1))) return MB2_ERR_CUDA;
    int* d_mine = (int*)d_counts.p; int* d_all = d_mine + U;
    mb2_dev_copy(ctx, d_mine, counts.data(), (size_t)U * 4, 0);
    if (g_nccl.AllGather(d_mine, d_all, (size_t)U * 4, NCCL_INT8, (ncclComm_h)comm, st) != 0) return MB2_ERR_CUDA;
    std::vector<int> g((size_t)U * world);
    mb2_dev_copy(ctx, g.data(), d_all, g.size() * 4, 1);
    if (mb2_ctx_sync(ctx) != MB2_OK) return MB2_ERR_CUDA;
    for (int u = 0; u < U; u++) all_counts[u] = g[(size_t)owner[u] * U + u];
  }
  std::vector<int> src_off(U);
  const int stride = mb2_shard_layout(owner.data(), all_counts.data(), U, world, src_off.data());
  const unsigned char* gathered = (const unsigned char*)send.p;
  if (world > 1) {
    if ((size_t)stride > send_cap) {   // the padded block must be readable: grow the send buffer to the stride
      void* np_ = mb2_dev_alloc(ctx, (size_t)stride * REC + 256);
      if (!np_) return MB2_ERR_CUDA;
      mb2_dev_copy(ctx, np_, send.p, used * REC, 2);
      mb2_ctx_sync(ctx);
      mb2_dev_free(ctx, send.p); send.p = np_; send.cap = (size_t)stride * REC + 256; send_cap = stride;
    }
    if (!recv.reserve((size_t)stride * REC * world)) return MB2_ERR_CUDA;
    if (g_nccl.AllGather(send.p, recv.p, (size_t)stride * REC, NCCL_INT8, (ncclComm_h)comm, st) != 0) return MB2_ERR_CUDA;
    gathered = (const unsigned char*)recv.p;
  }
  // ---- the reference's order: per (image, detector) the units in view-index order
  int set_n[4] = {0, 0, 0, 0};
  for (int u = 0; u < U; u++) set_n[units[u].image * 2 + units[u].det] += all_counts[u];
  for (int s = 0; s < 4; s++) if (!ordered[s].reserve((size_t)std::max(set_n[s], 1) * REC)) return MB2_ERR_CUDA;
  {
    int fill[4] = {0, 0, 0, 0};
    for (int u = 0; u < U; u++) {
      const int s = units[u].image * 2 + units[u].det;
      if (all_counts[u] > 0)
        mb2_dev_copy(ctx, (unsigned char*)ordered[s].p + (size_t)fill[s] * REC, gathered + (size_t)src_off[u] * REC, (size_t)all_counts[u] * REC, 2);
      fill[s] += all_counts[u];
    }
  }
  // slots: image 0 -> 0 (HessianAffine), 2 (MSER); image 1 -> 1, 3 -- the numbering mb2_mods_pair uses
  for (int s = 0; s < 4; s++) {
    const int rc = mb2_slot_from_records(ctx, (s >> 1) + 2 * (s & 1), ordered[s].p, set_n[s]);
    if (rc < 0) return rc;
  }
  if (mb2_ctx_sync(ctx) != MB2_OK) return MB2_ERR_CUDA;
  const double t_gather = now_ms();
  res->regions1 = set_n[0] + set_n[1]; res->regions2 = set_n[2] + set_n[3]; res->mser_regions1 = set_n[1]; res->mser_regions2 = set_n[3];

  // ---- matching, split by query rows; rows of all ranks gathered per detector
  std::vector<double> rows_all;   // 8 doubles: detector + the 7 tentative columns, detector-major, query order
  double t_match = 0, t_tgather = 0;
  for (int det = 0; det < (cfg->use_mser ? 2 : 1); det++) {
    const int nq = set_n[det], nt = set_n[2 + det];
    const int lo = (int)((long long)nq * rank / world), hi = (int)((long long)nq * (rank + 1) / world);
    std::vector<double> rows((size_t)std::max(hi - lo, 1) * 7);
    int n = 0;
    const double tm0 = now_ms();
    if (hi > lo && nt > 0) {
      n = mb2_match_slots_range(ctx, det == 0 ? 0 : 2, det == 0 ? 1 : 3, lo, hi, det == 0 ? cfg->matchRatio : cfg->mserMatchRatio, cfg->contradDist, 50, rows.data(), hi - lo);
      if (n < 0) return n;
    }
    const double tm1 = now_ms();
    t_match += tm1 - tm0;
    std::vector<int> ncnt(world, n);
    std::vector<double> got;
    if (world > 1) {
      if (!d_counts.reserve((size_t)4 * (world + 1))) return MB2_ERR_CUDA;
      int* d_mine = (int*)d_counts.p; int* d_all = d_mine + 1;
      mb2_dev_copy(ctx, d_mine, &n, 4, 0);
      if (g_nccl.AllGather(d_mine, d_all, 4, NCCL_INT8, (ncclComm_h)comm, st) != 0) return MB2_ERR_CUDA;
      mb2_dev_copy(ctx, ncnt.data(), d_all, (size_t)4 * world, 1);
      if (mb2_ctx_sync(ctx) != MB2_OK) return MB2_ERR_CUDA;
      int mx = 1;
      for (int c : ncnt) mx = std::max(mx, c);
      if (!d_rows.reserve((size_t)mx * 56 * (world + 1))) return MB2_ERR_CUDA;
      unsigned char* d_mr = (unsigned char*)d_rows.p; unsigned char* d_ar = d_mr + (size_t)mx * 56;
      mb2_dev_copy(ctx, d_mr, rows.data(), (size_t)n * 56, 0);
      if (g_nccl.AllGather(d_mr, d_ar, (size_t)mx * 56, NCCL_INT8, (ncclComm_h)comm, st) != 0) return MB2_ERR_CUDA;
      got.resize((size_t)mx * 7 * world);
      mb2_dev_copy(ctx, got.data(), d_ar, got.size() * 8, 1);
      if (mb2_ctx_sync(ctx) != MB2_OK) return MB2_ERR_CUDA;
      for (int r = 0; r < world; r++) {
        const size_t base = rows_all.size();
        rows_all.resize(base + (size_t)ncnt[r] * 8);
        for (int i = 0; i < ncnt[r]; i++) { double* o = &rows_all[base + (size_t)i * 8]; o[0] = det; std::memcpy(o + 1, &got[((size_t)r * mx + i) * 7], 56); }
      }
    } else {
      const size_t base = rows_all.size();
      rows_all.resize(base + (size_t)n * 8);
      for (int i = 0; i < n; i++) { double* o = &rows_all[base + (size_t)i * 8]; o[0] = det; std::memcpy(o + 1, &rows[(size_t)i * 7], 56); }
    }
    int tot = 0;
    for (int c : ncnt) tot += c;
    if (world == 1) tot = n;
    res->tentatives += tot;
    if (det == 1) res->mser_tentatives = tot;
    t_tgather += now_ms() - tm1;
  }
  // ---- checksums (order-sensitive): records of the four sets on the device, tentative rows on the host
  if (digest) {
    digest[0] = digest[1] = digest[2] = digest[3] = 0;
    for (int s = 0; s < 4; s++) {
      unsigned long long a = 0;
      const int rc = mb2_records_checksum(ctx, ordered[s].p, set_n[s], &a);
      if (rc < 0) return rc;
      digest[s >> 1] ^= (a + (unsigned long long)set_n[s]) * (unsigned long long)(2 * (s & 1) + 1);
    }
    unsigned long long a = 1469598103934665603ull;
    const unsigned long long* rb = (const unsigned long long*)rows_all.data();
    for (size_t i = 0; i < rows_all.size(); i++) { a ^= rb[i]; a *= 1099511628211ull; }
    digest[2] = a; digest[3] = (unsigned long long)res->tentatives;
  }
  // ---- frames of the tentatives for the rank that verifies (mods.cpp:298-415), straight from the device records (per detector: image-0 set
  // x image-1 set); keys from the distances
  const double t_v0 = now_ms();
  F.T = 0;
  if (verifier && res->tentatives > 0) {
    const int T = res->tentatives;
    F.frames.resize((size_t)T * 14); F.key.resize(T);
    std::vector<int> qi(T), ti(T);
    int done = 0;
    for (int det = 0; det < (cfg->use_mser ? 2 : 1); det++) {
      int n = 0;
      for (int i = done; i < T && (int)rows_all[(size_t)i * 8] == det; i++, n++) {
        const double* r = &rows_all[(size_t)i * 8];
        qi[i] = (int)r[1]; ti[i] = (int)r[2];
        F.key[i] = std::fabs(std::sqrt((double)((float)r[5] / (float)r[6])));   // TentativeCorrespExt::ratio = sqrt(d1 / d2), matching.cpp:449
      }
      const int rc = mb2_records_gather_frames(ctx, ordered[det].p, ordered[2 + det].p, qi.data() + done, ti.data() + done, n, F.frames.data() + (size_t)done * 14);
      if (rc < 0) return rc;
      done += n;
    }
    F.T = T;
  }
  F.t_start = t_start; F.t_gather = t_gather; F.t_match = t_match;
  res->ms_detect_describe = t_gather - t_start; res->ms_match = t_match; res->ms_total = now_ms() - t_start;
  if (stats) {
    stats[0] = t_views - t_start; stats[1] = t_gather - t_views; stats[2] = t_match; stats[3] = t_tgather; stats[4] = now_ms() - t_v0;
    stats[5] = world > 1 ? (double)stride * REC : 0.0; stats[6] = res->regions1 + res->regions2; stats[7] = my_units;
  }
  return MB2_OK;
}

// DuplicateFiltering + LORANSACFiltering of a pair's tentatives on context vctx (mb2_host_verify); fills the verification fields of res.
int sharded_back(mb2_ctx* vctx, const mb2_pair_config* cfg, ShardFront& F, mb2_pair_result* res, double* verified_out, int capacity, double* stats) {
  if (F.T <= 0) return 0;
  const double t0 = now_ms();
  mb2_pair_result vr; std::memset(&vr, 0, sizeof vr);
  const int k = mb2_host_verify(vctx, F.frames.data(), F.key.data(), F.T, cfg, &vr, verified_out, capacity);
  if (k < 0) return k;
  res->unique_tentatives = vr.unique_tentatives; res->ransac_inliers = vr.ransac_inliers; res->verified = vr.verified;
  std::memcpy(res->H, vr.H, sizeof vr.H); res->ms_duplicate = vr.ms_duplicate; res->ms_ransac = vr.ms_ransac;
  if (stats) stats[4] += now_ms() - t0;
  return k;
}
}  // namespace

extern "C" {
int mb2_views_sharded_pair(mb2_ctx* ctx, void* comm, int rank, int world, const float* img1, int w1, int h1, const float* img2, int w2, int h2,
                           const mb2_pair_config* cfg, mb2_pair_result* res, double* verified_out, int capacity, unsigned long long* digest, double* stats) {
  ShardFront F;
  const int rc = sharded_front(ctx, comm, rank, world, img1, w1, h1, img2, w2, h2, cfg, res, digest, stats, rank == 0, F);
  if (rc < 0) return rc;
  const int k = rank == 0 ? sharded_back(ctx, cfg, F, res, verified_out, capacity, stats) : 0;
  if (k < 0) return k;
  res->ms_total = now_ms() - F.t_start;
  return k;
}

// A list of independent pairs (a dataset run), views sharded over the ranks as above, with the part that does NOT shard inside one pair --
// duplicate filter + LO-RANSAC + LAF checks, 250 ms of a C4 pair -- spread over the ranks pair by pair and taken off the critical path:
// after the tentative exchange every rank holds all records and all tentative rows of pair k, so pair k is verified by rank k % world,
// on a helper thread and context of that rank, while all ranks go on with the views of pair k + 1.  At the end the small result records
// are all-gathered: res[k] is complete on EVERY rank; verified_out[k] (optional) is filled on rank k % world only.
// digest: 4 x n_pairs, stats: 8 x n_pairs (both optional).  Results are identical to calling mb2_views_sharded_pair pair by pair.
int mb2_views_sharded_pairs(mb2_ctx* ctx, void* comm, int rank, int world, int n_pairs, const float* const* img1, const int* w1, const int* h1,
                            const float* const* img2, const int* w2, const int* h2, const mb2_pair_config* cfg, mb2_pair_result* res,
                            double* const* verified_out, const int* capacity, unsigned long long* digest, double* stats) {
  if (!ctx || n_pairs < 0 || !cfg || !res || world < 1 || rank < 0 || rank >= world || (n_pairs > 0 && (!img1 || !img2 || !w1 || !h1 || !w2 || !h2))) return MB2_ERR_ARG;
  if (world > 1 && (!comm || !g_nccl.load())) return MB2_ERR_ARG;
  mb2_ctx* vctx = mb2_mods_sibling(ctx, 4);   // [0..2] run views next to ctx; [4] verifies
  if (!vctx) return MB2_ERR_CUDA;
  std::mutex m;
  std::condition_variable cv;
  std::deque<std::pair<int, std::unique_ptr<ShardFront> > > q;
  bool done = false;
  int rc = MB2_OK;
  std::thread back([&] {
    for (;;) {
      std::pair<int, std::unique_ptr<ShardFront> > item;
      {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return done || !q.empty(); });
        if (q.empty()) return;
        item = std::move(q.front()); q.pop_front();
      }
      cv.notify_all();
      const int k = item.first;
      const int r = sharded_back(vctx, cfg, *item.second, &res[k], verified_out ? verified_out[k] : nullptr, capacity ? capacity[k] : 0, stats ? stats + 8 * k : nullptr);
      if (r < 0) { std::lock_guard<std::mutex> lk(m); if (rc >= 0) rc = r; }
    }
  });
  for (int k = 0; k < n_pairs; k++) {
    std::unique_ptr<ShardFront> F(new ShardFront);
    const bool mine = (k % world) == rank;
    const int r = sharded_front(ctx, comm, rank, world, img1[k], w1[k], h1[k], img2[k], w2[k], h2[k], cfg, &res[k], digest ? digest + 4 * k : nullptr,
                                stats ? stats + 8 * k : nullptr, mine, *F);
    if (r < 0) { std::lock_guard<std::mutex> lk(m); if (rc >= 0) rc = r; break; }   // a failed collective fails on every rank alike
    if (!mine) continue;
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [&] { return q.size() < 2; });   // at most two of this rank's pairs waiting for verification
    q.emplace_back(k, std::move(F));
    lk.unlock();
    cv.notify_all();
  }
  { std::lock_guard<std::mutex> lk(m); done = true; }
  cv.notify_all();
  back.join();
  // every rank learns every pair's result: one all-gather of the result records, pair k taken from rank k % world
  if (world > 1 && n_pairs > 0) {
    int bad = rc < 0 ? 1 : 0;
    const size_t bytes = (size_t)n_pairs * sizeof(mb2_pair_result) + 8;
    ShardScratch& S = *scratch_of(ctx);
    S.d_rows.ctx = ctx;
    if (!S.d_rows.reserve(bytes * (world + 1))) return MB2_ERR_CUDA;
    unsigned char* d_mine = (unsigned char*)S.d_rows.p; unsigned char* d_all = d_mine + bytes;
    std::vector<unsigned char> h(bytes * world, 0);
    std::memcpy(h.data(), res, (size_t)n_pairs * sizeof(mb2_pair_result));
    std::memcpy(h.data() + (size_t)n_pairs * sizeof(mb2_pair_result), &bad, sizeof bad);
    mb2_dev_copy(ctx, d_mine, h.data(), bytes, 0);
    if (g_nccl.AllGather(d_mine, d_all, bytes, NCCL_INT8, (ncclComm_h)comm, mb2_ctx_stream(ctx)) != 0) return MB2_ERR_CUDA;
    mb2_dev_copy(ctx, h.data(), d_all, bytes * world, 1);
    if (mb2_ctx_sync(ctx) != MB2_OK) return MB2_ERR_CUDA;
    for (int r = 0; r < world; r++) {
      int b = 0;
      std::memcpy(&b, h.data() + (size_t)r * bytes + (size_t)n_pairs * sizeof(mb2_pair_result), sizeof b);
      if (b && rc >= 0) rc = MB2_ERR_CUDA;   // some rank failed: the call fails everywhere
    }
    for (int k = 0; k < n_pairs; k++) {
      std::memcpy(&res[k], h.data() + (size_t)(k % world) * bytes + (size_t)k * sizeof(mb2_pair_result), sizeof(mb2_pair_result));
    }
  }
  return rc < 0 ? rc : n_pairs;
}

}  // extern "C"
