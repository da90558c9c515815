// lanes.h -- TEST INFRASTRUCTURE ONLY.  A group of N host threads that run a kernel core in lock
// step: lane(), size(), sync() and the reductions have the semantics of WarpGroup / BlockGroup in
// trgt_b200/csrc/coop.h, so the collective logic of the cores (leader election, broadcasts,
// barrier placement) is exercised on the CPU, under a real memory model, before it ever runs on
// a GPU.  Every collective is two barriers around a shared slot array.
#pragma once
#include <pthread.h>

#include <climits>
#include <functional>
#include <thread>
#include <vector>

namespace trgt_test {

struct LaneShared {
  pthread_barrier_t bar;
  std::vector<int> slot;
  int n;
  explicit LaneShared(int n_) : slot(n_), n(n_) { pthread_barrier_init(&bar, nullptr, (unsigned)n_); }
  ~LaneShared() { pthread_barrier_destroy(&bar); }
};

struct LaneGroup {
  int lane_;
  LaneShared *sh;
  int lane() const { return lane_; }
  int size() const { return sh->n; }
  void sync() const { pthread_barrier_wait(&sh->bar); }
  template <class F>
  int reduce(int v, F f) const {
    sh->slot[lane_] = v;
    sync();
    int r = sh->slot[0];
    for (int i = 1; i < sh->n; i++) r = f(r, sh->slot[i]);
    sync();
    return r;
  }
  int min_i(int v) const { return reduce(v, [](int a, int b) { return a < b ? a : b; }); }
  int max_i(int v) const { return reduce(v, [](int a, int b) { return a > b ? a : b; }); }
  int any(int p) const { return reduce(p ? 1 : 0, [](int a, int b) { return a | b; }); }
  int bcast(int v, int src) const {
    sh->slot[lane_] = v;
    sync();
    const int r = sh->slot[src];
    sync();
    return r;
  }
  int bcast0(int v) const { return bcast(v, 0); }
  int excl_scan_i(int v, int *total) const {
    sh->slot[lane_] = v;
    sync();
    int before = 0, all = 0;
    for (int i = 0; i < sh->n; i++) {
      if (i < lane_) before += sh->slot[i];
      all += sh->slot[i];
    }
    sync();
    *total = all;
    return before;
  }
};

// run fn(group) on n lanes; returns when all are done
inline void run_lanes(int n, const std::function<void(const LaneGroup &)> &fn) {
  LaneShared sh(n);
  std::vector<std::thread> th;
  for (int i = 0; i < n; i++) th.emplace_back([&, i]() { LaneGroup g{i, &sh}; fn(g); });
  for (auto &t : th) t.join();
}

}  // namespace trgt_test
