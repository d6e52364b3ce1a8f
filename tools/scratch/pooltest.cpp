#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <queue>
#include <thread>
#include <vector>
#include <memory>
#include <sys/mman.h>
class Pool {
 public:
  explicit Pool(int n) { for (int i = 0; i < n; ++i) th_.emplace_back([this] { run(); }); }
  ~Pool() { { std::lock_guard<std::mutex> l(mu_); quit_ = true; } cv_.notify_all(); for (auto& t : th_) t.join(); }
  void submit(std::function<void()> f) { { std::lock_guard<std::mutex> l(mu_); q_.push(std::move(f)); } cv_.notify_one(); }
 private:
  void run() { for (;;) { std::function<void()> f; { std::unique_lock<std::mutex> l(mu_); cv_.wait(l, [this] { return quit_ || !q_.empty(); }); if (q_.empty()) return; f = std::move(q_.front()); q_.pop(); } f(); } }
  std::vector<std::thread> th_; std::queue<std::function<void()>> q_; std::mutex mu_; std::condition_variable cv_; bool quit_ = false;
};
struct TaskGroup { std::mutex mu; std::condition_variable cv; int left = 0; void add(int n){std::lock_guard<std::mutex> l(mu); left+=n;} void done(){std::lock_guard<std::mutex> l(mu); if(--left==0) cv.notify_all();} void wait(){std::unique_lock<std::mutex> l(mu); cv.wait(l,[this]{return left==0;});} };
int main() {
  const size_t total = 419430400, CH = 4 << 20; const int NSTG = 8;
  std::vector<char*> views;
  for (int i = 0; i < 100; ++i) { char* p = (char*)malloc(total / 100); memset(p, i, total / 100); views.push_back(p); }
  std::vector<char*> stg; for (int i = 0; i < NSTG; ++i) { char* p = (char*)malloc(CH); memset(p, 1, CH); mlock(p, CH); stg.push_back(p); }
  Pool pool(32);
  for (int lag = 1; lag <= 6; ++lag) {
    for (int rep = 0; rep < 2; ++rep) {
      auto t0 = std::chrono::steady_clock::now();
      std::vector<std::unique_ptr<TaskGroup>> grp(100);
      for (int k = 0; k < 100; ++k) {
        grp[k].reset(new TaskGroup()); grp[k]->add(1);
        char* d = stg[k % NSTG]; const char* s = views[k]; TaskGroup* g = grp[k].get();
        pool.submit([=] { memcpy(d, s, CH); g->done(); });
        if (k >= lag) grp[k - lag]->wait();
      }
      for (int k = 100 - lag; k < 100; ++k) grp[k]->wait();
      double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (rep) printf("lag %d: %.1f GB/s\n", lag, total / dt / 1e9);
    }
  }
  for (int T : {1, 2, 4, 8, 16}) {   // static partition: raw multi-thread memcpy bandwidth, no task hand-off
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) th.emplace_back([&, t] { for (int k = t; k < 100; k += T) memcpy(stg[(k % NSTG)], views[k], CH); });
    for (auto& x : th) x.join();
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("static %d threads: %.1f GB/s\n", T, total / dt / 1e9);
  }
  // single thread
  auto t0 = std::chrono::steady_clock::now();
  for (int k = 0; k < 100; ++k) memcpy(stg[k % NSTG], views[k], CH);
  double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  printf("single thread: %.1f GB/s\n", total / dt / 1e9);
}
