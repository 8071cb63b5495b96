// Host-side parallel loop for the short CPU legs of the path (least-squares set-ups, libm legs, the duplicate filter's join).
//
// Why not OpenMP: libgomp's worker threads SPIN for a while after every parallel region (OMP_WAIT_POLICY defaults to active).  These
// legs are 0.1 - 2 ms bursts issued dozens of times per image pair from several host threads that otherwise wait on the GPU; the
// spinning pool threads took the cores from exactly those threads (measured with two ranks on one node: 35.8 ms per pair with the
// OpenMP regions, 31.9 ms with OMP_WAIT_POLICY=passive).  The wait policy can only be set through the environment before libgomp
// loads, which a library cannot rely on -- so the path brings its own small pool whose workers BLOCK on a condition variable.
//
// parallel_chunks(n, fn): fn(c) for c in [0, n), dealt out dynamically; the caller takes part; returns when all are done.  Results
// must not depend on which thread ran a chunk (callers sum per-chunk partials in chunk order afterwards).
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace mb2par {

class Pool {
 public:
  explicit Pool(int n) {
    for (int i = 0; i < n; i++)
      th_.emplace_back([this] {
        for (;;) {
          std::function<void()> f;
          {
            std::unique_lock<std::mutex> lk(m_);
            cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
            if (q_.empty()) return;
            f = std::move(q_.front()); q_.pop_front();
          }
          f();
        }
      });
  }
  ~Pool() {
    { std::lock_guard<std::mutex> lk(m_); stop_ = true; }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  int size() const { return (int)th_.size(); }
  void submit(int copies, const std::function<void()>& f) {
    { std::lock_guard<std::mutex> lk(m_); for (int i = 0; i < copies; i++) q_.push_back(f); }
    if (copies == 1) cv_.notify_one(); else cv_.notify_all();
  }

 private:
  std::mutex m_;
  std::condition_variable cv_;
  std::deque<std::function<void()> > q_;
  std::vector<std::thread> th_;
  bool stop_ = false;
};

// MB2_HOST_THREADS: helper threads of the pool (default: a quarter of the hardware threads, at most 6 -- the legs are short, and under
// torchrun every rank has its own pool); 0 = everything on the calling thread
inline Pool& pool() {
  static Pool p([] {
    if (const char* e = std::getenv("MB2_HOST_THREADS")) return std::max(0, std::min(64, std::atoi(e)));
    const int hw = (int)std::thread::hardware_concurrency();
    return std::max(1, std::min(6, hw / 4));
  }());
  return p;
}

template <class F>
void parallel_chunks(int n, F&& fn) {
  if (n <= 0) return;
  Pool& P = pool();
  const int helpers = std::min(P.size(), n - 1);
  if (helpers <= 0) { for (int c = 0; c < n; c++) fn(c); return; }
  struct Job { std::atomic<int> next{0}, done{0}; int n = 0; std::mutex m; std::condition_variable cv; };
  std::shared_ptr<Job> job = std::make_shared<Job>();
  job->n = n;
  auto* fp = &fn;   // only dereferenced for a chunk index below n, i.e. while this call is still waiting
  std::function<void()> work = [job, fp] {
    for (;;) {
      const int c = job->next.fetch_add(1);
      if (c >= job->n) return;
      (*fp)(c);
      if (job->done.fetch_add(1) + 1 == job->n) { std::lock_guard<std::mutex> lk(job->m); job->cv.notify_all(); }
    }
  };
  P.submit(helpers, work);
  work();
  std::unique_lock<std::mutex> lk(job->m);
  job->cv.wait(lk, [&] { return job->done.load() == job->n; });
}

}  // namespace mb2par
