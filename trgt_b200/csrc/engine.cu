// engine.cu -- C ABI of libtrgt_b200.so (include/trgt_engine.h): device memory, streams, launches.
//
// There is no CPU path in this library: every entry point that computes launches the sm_100a
// kernels of kernels.cuh, and trgt_engine_create fails without a CUDA device.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <cub/device/device_scan.cuh>
#include <thrust/iterator/transform_iterator.h>
#include <algorithm>
#include <chrono>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/trgt_engine.h"
#include "hmm_host.h"
#include "kernels.cuh"

using namespace trgt;

// ------------------------------------------------------------------ plumbing --------------

namespace {

thread_local std::string g_create_error;

struct KernelStat {
  const char *name;
  uint64_t launches;
  double total_ms;
};

struct PendingEvent {
  int stat;
  cudaEvent_t a, b;
};

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
};

// grow-only pinned host buffer: D2H copies land here at full PCIe rate and callers read it in place
struct PinBuf {
  void *p = nullptr;
  size_t cap = 0;
  template <class T> T *as() { return (T *)p; }
};

struct CastU64 {
  __host__ __device__ unsigned long long operator()(const uint32_t &v) const { return (unsigned long long)v; }
};

}  // namespace

struct trgt_engine {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // H2D of read chunks, overlapped with phase A kernels
  cudaEvent_t sleep_ev = nullptr;      // blocking-sync event: waits on the stream sleep instead of spinning (TRGT_SYNC=sleep)
  int sm_count = 0;
  int smem_optin = 0;
  std::string error;
  std::mutex mu;
  bool profiling = false;
  uint64_t launches = 0;
  std::vector<KernelStat> stats;
  std::vector<PendingEvent> pending;
  std::vector<cudaEvent_t> event_pool;
  // HMM constants
  HmmConsts hmm_consts;
  HmmJumpTable jump;
  DevBuf d_mm_off, d_mm_lp;
  size_t jump_uploaded_len = 0;
  // scan scratch
  DevBuf d_scan_tmp;
  // pinned arena for host-built lists on their way to the device (see h2d_staged)
  PinBuf h_stage;
  size_t stage_used = 0;
  // pinned staging for small read-backs
  Counters *h_ctr = nullptr;
  unsigned long long *h_u64 = nullptr;
  // one-shot batches (results stay valid until the next call)
  trgt_flank_batch_t *one_flank = nullptr;
  trgt_align_batch_t *one_align = nullptr;
  trgt_hmm_batch_t *one_hmm = nullptr;
  DevBuf d_ed[6];
  DevBuf d_cl[5];  // cluster glue: scratch slots, group ids, central reads, group counts, read index
  size_t workspace_budget = (size_t)24 << 30;  // cap on back-pointer / trace workspace per wave
  int band_budget = 16;  // cost cap of the banded flank fallback (0: always use the full-width path)
  bool hmm_lane = true;  // single-motif loci take the register / packed-word HMM kernels (k_hmm_lane_*)
};

namespace {

int fail(trgt_engine *e, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (e) e->error = buf;
  return code;
}

#define CU(e, call)                                                                          \
  do {                                                                                       \
    cudaError_t err_ = (call);                                                               \
    if (err_ != cudaSuccess)                                                                 \
      return fail((e), TRGT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err_), \
                  __FILE__, __LINE__);                                                       \
  } while (0)

#define TRY(expr)          \
  do {                     \
    int rc_ = (expr);      \
    if (rc_ != 0) return rc_; \
  } while (0)

// Wait for everything queued on the engine's stream.  A host that drives one engine per thread, several threads per
// GPU, has as many threads inside such waits as it has engines: spinning (the driver's default) takes a core each
// away from the host's own work, sleeping on a blocking-sync event gives it back at the price of a wake-up.
cudaError_t engine_wait(trgt_engine *e) {
  if (e->sleep_ev == nullptr) return cudaStreamSynchronize(e->stream);
  cudaError_t err = cudaEventRecord(e->sleep_ev, e->stream);
  if (err != cudaSuccess) return err;
  return cudaEventSynchronize(e->sleep_ev);
}

// Buffers of resident batches that have been freed wait here for the next batch (process-wide, per device): a host
// that streams chunk after chunk through trgt_*_upload / _free would otherwise pay a cudaMalloc / cudaFree (each a
// device-wide synchronisation) and a cudaMallocHost per array and chunk.  Emptied when the last engine of the
// device is destroyed.
struct PoolEntry { void *p; size_t cap; int device; };
struct BufPool {
  std::mutex mu;
  std::vector<PoolEntry> dev, pin;
  size_t dev_bytes = 0, pin_bytes = 0;
  int engines[64] = {0};
  unsigned long long hits = 0, misses = 0, puts = 0, rejected = 0;  // TRGT_TRACE prints them when a device's last engine goes
};
BufPool g_pool;
const size_t POOL_DEV_MAX = 32ull << 30, POOL_PIN_MAX = 4ull << 30;

// best fit: the smallest pooled buffer that holds `bytes` without being wastefully large
void *pool_take(std::vector<PoolEntry> &v, size_t &total, int device, size_t bytes, size_t *cap_out) {
  std::lock_guard<std::mutex> lk(g_pool.mu);
  int best = -1;
  for (size_t i = 0; i < v.size(); i++)
    if (v[i].device == device && v[i].cap >= bytes && v[i].cap <= 4 * bytes + (4u << 20) &&
        (best < 0 || v[i].cap < v[(size_t)best].cap))
      best = (int)i;
  if (best < 0) { g_pool.misses++; return nullptr; }
  g_pool.hits++;
  void *p = v[(size_t)best].p;
  *cap_out = v[(size_t)best].cap;
  total -= v[(size_t)best].cap;
  v[(size_t)best] = v.back();
  v.pop_back();
  return p;
}

bool pool_put(std::vector<PoolEntry> &v, size_t &total, size_t max_total, int device, void *p, size_t cap) {
  std::lock_guard<std::mutex> lk(g_pool.mu);
  if (device < 0 || device >= 64 || g_pool.engines[device] <= 0 || total + cap > max_total) { g_pool.rejected++; return false; }
  g_pool.puts++;
  v.push_back(PoolEntry{p, cap, device});
  total += cap;
  return true;
}

void pool_release(int device) {
  std::vector<PoolEntry> d, h;
  {
    std::lock_guard<std::mutex> lk(g_pool.mu);
    for (size_t i = 0; i < g_pool.dev.size();)
      if (g_pool.dev[i].device == device) {
        d.push_back(g_pool.dev[i]); g_pool.dev_bytes -= g_pool.dev[i].cap;
        g_pool.dev[i] = g_pool.dev.back(); g_pool.dev.pop_back();
      } else i++;
    for (size_t i = 0; i < g_pool.pin.size();)
      if (g_pool.pin[i].device == device) {
        h.push_back(g_pool.pin[i]); g_pool.pin_bytes -= g_pool.pin[i].cap;
        g_pool.pin[i] = g_pool.pin.back(); g_pool.pin.pop_back();
      } else i++;
  }
  for (auto &x : d) cudaFree(x.p);
  for (auto &x : h) cudaFreeHost(x.p);
}

int dev_reserve(trgt_engine *e, DevBuf &b, size_t bytes, bool keep = false) {
  if (bytes <= b.cap && b.p) return 0;
  size_t ncap = bytes + bytes / 8 + 256;  // (a pooled buffer only has to hold `bytes`: the slack is for fresh ones)
  void *np = pool_take(g_pool.dev, g_pool.dev_bytes, e->device, bytes, &ncap);
  if (!np) CU(e, cudaMalloc(&np, ncap));
  if (keep && b.p && b.cap) CU(e, cudaMemcpyAsync(np, b.p, b.cap, cudaMemcpyDeviceToDevice, e->stream));
  if (b.p) {
    CU(e, engine_wait(e));
    if (!pool_put(g_pool.dev, g_pool.dev_bytes, POOL_DEV_MAX, e->device, b.p, b.cap)) CU(e, cudaFree(b.p));
  }
  b.p = np;
  b.cap = ncap;
  return 0;
}

int pin_reserve(trgt_engine *e, PinBuf &b, size_t bytes) {
  if (bytes <= b.cap && b.p) return 0;
  if (b.p && !pool_put(g_pool.pin, g_pool.pin_bytes, POOL_PIN_MAX, e->device, b.p, b.cap)) cudaFreeHost(b.p);
  b.p = nullptr;
  b.cap = 0;
  size_t ncap = bytes + bytes / 8 + 256;
  b.p = pool_take(g_pool.pin, g_pool.pin_bytes, e->device, bytes, &ncap);
  if (!b.p) CU(e, cudaMallocHost(&b.p, ncap));
  b.cap = ncap;
  return 0;
}

// device >= 0: the buffer goes to the pool of that device (the caller has made sure nothing in flight uses it)
void pin_free(PinBuf &b, int device = -1) {
  if (b.p && !pool_put(g_pool.pin, g_pool.pin_bytes, POOL_PIN_MAX, device, b.p, b.cap)) cudaFreeHost(b.p);
  b.p = nullptr;
  b.cap = 0;
}

void dev_free(DevBuf &b, int device = -1) {
  if (b.p && !pool_put(g_pool.dev, g_pool.dev_bytes, POOL_DEV_MAX, device, b.p, b.cap)) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}

int h2d(trgt_engine *e, DevBuf &b, const void *src, size_t bytes, size_t pad = 16) {
  TRY(dev_reserve(e, b, bytes + pad));
  if (bytes) CU(e, cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, e->stream));
  return 0;
}

// Lists the engine builds on the host (std::vector: pageable memory) go up through a pinned arena: an upload from
// pageable memory makes the runtime synchronise the stream first and stage the bytes itself, one blocking step per
// list behind whatever the stream still has queued on a shared link; from the arena they are ordinary asynchronous
// copies.  stage_begin reserves the arena for one call's lists (the call ends with a wait on the stream, so the arena is
// free again for the next call).
int stage_begin(trgt_engine *e, size_t total_bytes) {
  TRY(pin_reserve(e, e->h_stage, total_bytes + 256));
  e->stage_used = 0;
  return 0;
}
int h2d_staged(trgt_engine *e, DevBuf &b, const void *src, size_t bytes, size_t pad = 16) {
  TRY(dev_reserve(e, b, bytes + pad));
  if (!bytes) return 0;
  const size_t at = (e->stage_used + 15) & ~(size_t)15;
  if (at + bytes > e->h_stage.cap) return h2d(e, b, src, bytes, pad);  // (not reserved for: the plain path)
  memcpy((uint8_t *)e->h_stage.p + at, src, bytes);
  e->stage_used = at + bytes;
  CU(e, cudaMemcpyAsync(b.p, (const uint8_t *)e->h_stage.p + at, bytes, cudaMemcpyHostToDevice, e->stream));
  return 0;
}

int stat_index(trgt_engine *e, const char *name) {
  for (size_t i = 0; i < e->stats.size(); i++)
    if (e->stats[i].name == name || strcmp(e->stats[i].name, name) == 0) return (int)i;
  e->stats.push_back(KernelStat{name, 0, 0.0});
  return (int)e->stats.size() - 1;
}

cudaEvent_t get_event(trgt_engine *e) {
  if (!e->event_pool.empty()) {
    cudaEvent_t ev = e->event_pool.back();
    e->event_pool.pop_back();
    return ev;
  }
  cudaEvent_t ev;
  cudaEventCreate(&ev);
  return ev;
}

struct LaunchScope {
  trgt_engine *e;
  int si;
  cudaEvent_t a = nullptr, b = nullptr;
  LaunchScope(trgt_engine *e_, const char *name) : e(e_) {
    si = stat_index(e, name);
    e->stats[si].launches++;
    e->launches++;
    if (e->profiling) {
      a = get_event(e);
      b = get_event(e);
      cudaEventRecord(a, e->stream);
    }
  }
  ~LaunchScope() {
    if (e->profiling) {
      cudaEventRecord(b, e->stream);
      e->pending.push_back(PendingEvent{si, a, b});
    }
  }
};

void resolve_pending(trgt_engine *e) {
  if (e->pending.empty()) return;
  engine_wait(e);
  for (auto &p : e->pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) e->stats[p.stat].total_ms += ms;
    e->event_pool.push_back(p.a);
    e->event_pool.push_back(p.b);
  }
  e->pending.clear();
}

int check_launch(trgt_engine *e, const char *name) {
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(e, TRGT_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(err));
  return 0;
}

template <class K>
int persistent_grid(trgt_engine *e, K kernel, int block, size_t smem, int *grid_out) {
  if (smem > (size_t)e->smem_optin) return fail(e, TRGT_ERR_INTERNAL, "kernel wants %zu bytes of shared memory", smem);
  int per_sm = 0;
  CU(e, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem));
  if (per_sm < 1) return fail(e, TRGT_ERR_INTERNAL, "kernel does not fit on an SM (smem %zu)", smem);
  *grid_out = per_sm * e->sm_count;
  return 0;
}

// A persistent kernel whose every slot (CTA, or warp when slots_per_block > 1) owns `stride_ints` ints of global
// scratch: shrink the grid until the scratch fits the engine's workspace budget (fewer slots, same result).
int cap_grid_to_budget(trgt_engine *e, int *grid, int slots_per_block, size_t stride_ints, const char *what) {
  if (stride_ints == 0) return 0;
  const size_t budget_ints = e->workspace_budget / sizeof(int);
  const size_t per_block = stride_ints * (size_t)slots_per_block;
  if (stride_ints > budget_ints)
    return fail(e, TRGT_ERR_INTERNAL, "%s: one scratch slot of %zu ints exceeds the workspace budget", what, stride_ints);
  size_t fit = budget_ints / per_block;
  if (fit == 0) fit = 1;  // one CTA of several slots: over the budget by less than slots_per_block, but it has to run
  if ((size_t)*grid > fit) *grid = (int)fit;
  return 0;
}

int check_seqs(trgt_engine *e, const trgt_seqs_t *s, const char *what) {
  if (!s || (s->n && (!s->offsets))) return fail(e, TRGT_ERR_ARG, "%s: null sequence set", what);
  for (uint64_t i = 0; i < s->n; i++)
    if (s->offsets[i + 1] < s->offsets[i]) return fail(e, TRGT_ERR_ARG, "%s: offsets not monotone at %llu", what, (unsigned long long)i);
  if (s->n && s->offsets[s->n] > s->offsets[0] && !s->data) return fail(e, TRGT_ERR_ARG, "%s: null data", what);
  return 0;
}

uint64_t max_len(const trgt_seqs_t *s) {
  uint64_t m = 0;
  for (uint64_t i = 0; i < s->n; i++) {
    const uint64_t l = s->offsets[i + 1] - s->offsets[i];
    if (l > m) m = l;
  }
  return m;
}

int upload_seqs(trgt_engine *e, const trgt_seqs_t *s, DevBuf &data, DevBuf &off) {
  static const uint64_t zero_off[1] = {0};
  const uint64_t total = s->n ? s->offsets[s->n] : 0;
  TRY(h2d(e, data, s->data, (size_t)total));
  TRY(h2d(e, off, s->n ? s->offsets : zero_off, (size_t)(s->n + 1) * sizeof(uint64_t)));
  return 0;
}

int exclusive_scan_u32(trgt_engine *e, const uint32_t *d_in, unsigned long long *d_out, size_t n) {
  auto it = thrust::make_transform_iterator(d_in, CastU64());
  size_t tmp = 0;
  CU(e, cub::DeviceScan::ExclusiveSum(nullptr, tmp, it, d_out, (int)n, e->stream));
  TRY(dev_reserve(e, e->d_scan_tmp, tmp));
  LaunchScope ls(e, "cub_exclusive_scan");
  CU(e, cub::DeviceScan::ExclusiveSum(e->d_scan_tmp.p, tmp, it, d_out, (int)n, e->stream));
  return 0;
}

}  // namespace

// ------------------------------------------------------------------ lifetime --------------

extern "C" {

const char *trgt_last_create_error(void) { return g_create_error.c_str(); }

int32_t trgt_engine_create(int32_t device, trgt_engine_t **out) {
  if (!out) return TRGT_ERR_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t err = cudaGetDeviceCount(&count);
  if (err != cudaSuccess || count == 0) {
    g_create_error = std::string("no CUDA device: ") + (err != cudaSuccess ? cudaGetErrorString(err) : "device count is 0");
    return TRGT_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= count) {
    g_create_error = "device index out of range";
    return TRGT_ERR_ARG;
  }
  cudaDeviceProp prop;
  if ((err = cudaSetDevice(device)) != cudaSuccess || (err = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
    g_create_error = std::string("cudaSetDevice failed: ") + cudaGetErrorString(err);
    return TRGT_ERR_CUDA;
  }
  if (prop.major < 10) {
    g_create_error = "libtrgt_b200 is built for sm_100a only; device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor);
    return TRGT_ERR_NO_DEVICE;
  }
  trgt_engine *e = new trgt_engine();
  e->device = device;
  e->sm_count = prop.multiProcessorCount;
  e->smem_optin = (int)prop.sharedMemPerBlockOptin;
  if ((err = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (err = cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (err = cudaMallocHost((void **)&e->h_ctr, sizeof(Counters))) != cudaSuccess ||
      (err = cudaMallocHost((void **)&e->h_u64, 8 * sizeof(unsigned long long))) != cudaSuccess) {
    g_create_error = std::string("engine setup failed: ") + cudaGetErrorString(err);
    delete e;
    return TRGT_ERR_CUDA;
  }
  {
    const char *lane = getenv("TRGT_HMM_LANE");  // "0": single-motif loci take the generic HMM kernels too (measurements)
    if (lane && lane[0] == '0') e->hmm_lane = false;
  }
  {
    const char *mode = getenv("TRGT_SYNC");
    if (mode && strcmp(mode, "sleep") == 0 &&
        (err = cudaEventCreateWithFlags(&e->sleep_ev, cudaEventBlockingSync | cudaEventDisableTiming)) != cudaSuccess) {
      g_create_error = std::string("engine setup failed: ") + cudaGetErrorString(err);
      trgt_engine_destroy(e);
      return TRGT_ERR_CUDA;
    }
  }
  // Dynamic shared memory limits are per-function process state: raise them once to the device
  // maximum (never per launch, engines on other host threads may be launching concurrently).
  {
    const int mx = e->smem_optin;
    if ((err = cudaFuncSetAttribute(k_wfa_score<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx)) != cudaSuccess ||
        (err = cudaFuncSetAttribute(k_wfa_score<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx)) != cudaSuccess ||
        (err = cudaFuncSetAttribute(k_wfa_trace, cudaFuncAttributeMaxDynamicSharedMemorySize, mx)) != cudaSuccess ||
        (err = cudaFuncSetAttribute(k_flank_band1, cudaFuncAttributeMaxDynamicSharedMemorySize, mx)) != cudaSuccess ||
        (err = cudaFuncSetAttribute(k_e2e_lane, cudaFuncAttributeMaxDynamicSharedMemorySize, mx)) != cudaSuccess ||
        (err = cudaFuncSetAttribute(k_hmm_viterbi, cudaFuncAttributeMaxDynamicSharedMemorySize, mx)) != cudaSuccess ||
        (err = cudaFuncSetAttribute(k_hmm_viterbi_thread, cudaFuncAttributeMaxDynamicSharedMemorySize, mx)) != cudaSuccess) {
      g_create_error = std::string("cudaFuncSetAttribute failed: ") + cudaGetErrorString(err);
      trgt_engine_destroy(e);
      return TRGT_ERR_CUDA;
    }
  }
  e->hmm_consts = hmm_make_consts();
  {
    std::lock_guard<std::mutex> lk(g_pool.mu);
    if (e->device >= 0 && e->device < 64) g_pool.engines[e->device]++;
  }
  *out = e;
  return TRGT_OK;
}

void trgt_engine_destroy(trgt_engine_t *e) {
  if (!e) return;
  cudaSetDevice(e->device);
  engine_wait(e);
  if (e->one_flank) trgt_flank_free(e, e->one_flank);
  if (e->one_align) trgt_align_free(e, e->one_align);
  if (e->one_hmm) trgt_hmm_free(e, e->one_hmm);
  resolve_pending(e);
  for (auto ev : e->event_pool) cudaEventDestroy(ev);
  dev_free(e->d_mm_off);
  dev_free(e->d_mm_lp);
  dev_free(e->d_scan_tmp);
  for (auto &b : e->d_ed) dev_free(b);
  for (auto &b : e->d_cl) dev_free(b);
  pin_free(e->h_stage);
  if (e->h_ctr) cudaFreeHost(e->h_ctr);
  if (e->h_u64) cudaFreeHost(e->h_u64);
  if (e->sleep_ev) cudaEventDestroy(e->sleep_ev);
  cudaStreamDestroy(e->stream);
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  bool last = false;
  {
    std::lock_guard<std::mutex> lk(g_pool.mu);
    if (e->device >= 0 && e->device < 64) last = --g_pool.engines[e->device] == 0;
  }
  if (last) {
    if (getenv("TRGT_TRACE"))
      fprintf(stderr, "[trgt] buffer pool: %llu hits, %llu misses, %llu puts, %llu not pooled\n", g_pool.hits, g_pool.misses,
              g_pool.puts, g_pool.rejected);
    pool_release(e->device);
  }
  delete e;
}

const char *trgt_engine_last_error(const trgt_engine_t *e) { return e ? e->error.c_str() : ""; }
void *trgt_engine_stream(trgt_engine_t *e) { return e ? (void *)e->stream : nullptr; }
int32_t trgt_engine_sm_count(const trgt_engine_t *e) { return e ? e->sm_count : 0; }

int32_t trgt_engine_sync(trgt_engine_t *e) {
  if (!e) return TRGT_ERR_ARG;
  CU(e, engine_wait(e));
  return 0;
}

void *trgt_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
  return p;
}
void trgt_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

void trgt_engine_set_profiling(trgt_engine_t *e, int32_t on) {
  if (!e) return;
  resolve_pending(e);
  e->profiling = on != 0;
}
void trgt_engine_reset_stats(trgt_engine_t *e) {
  if (!e) return;
  resolve_pending(e);
  e->stats.clear();
  e->launches = 0;
}
int32_t trgt_engine_kernel_count(trgt_engine_t *e) {
  if (!e) return 0;
  resolve_pending(e);
  return (int32_t)e->stats.size();
}
int32_t trgt_engine_kernel_stat(trgt_engine_t *e, int32_t i, const char **name, uint64_t *launches, double *total_ms) {
  if (!e || i < 0 || (size_t)i >= e->stats.size()) return TRGT_ERR_ARG;
  resolve_pending(e);
  if (name) *name = e->stats[i].name;
  if (launches) *launches = e->stats[i].launches;
  if (total_ms) *total_ms = e->stats[i].total_ms;
  return 0;
}
uint64_t trgt_engine_launches(const trgt_engine_t *e) { return e ? e->launches : 0; }

}  // extern "C"

// ------------------------------------------------------------------ phase A ----------------

struct trgt_flank_batch {
  uint32_t n_loci = 0, n_reads = 0;
  int Pmax = 0, Pmin = 0, Tmax = 0;
  trgt_scoring_t scoring{2, 5, 1};
  double frac = 0.7;
  DevBuf reads, read_off, lp, lp_off, rp, rp_off, locus_read_off, read_locus;
  DevBuf hits, spans, work, work2, list1, ends, ctr, gring, gws;
  DevBuf kidx;                         // 8-mer indexes of both pieces of every locus (k_flank_exact_t -> k_flank_band)
  bool kidx_valid = false;
  bool ran = false;                    // spans / hits hold results
  DevBuf tr_len, tr_off, tr_data;      // trgt_flank_trs: repeat sequences of the spanning reads
  PinBuf h_tr_off, h_tr_data;
  DevBuf seq4, seq4_starts, seq4_len;  // BAM 4-bit input (trgt_flank_*_seq4): decoded into `reads` on the device
  uint32_t last_n_work = 0, last_n_tier2 = 0, last_n_wide = 0;
};

namespace {

WfaSrc flank_src(const trgt_flank_batch *b) {
  WfaSrc s;
  memset(&s, 0, sizeof s);
  s.mode = WFA_MODE_FLANK;
  s.x = b->scoring.mismatch;
  s.oe = b->scoring.gap_open + b->scoring.gap_extend;
  s.e = b->scoring.gap_extend;
  s.reads = (const uint8_t *)b->reads.p;
  s.read_off = (const uint64_t *)b->read_off.p;
  s.read_locus = (const uint32_t *)b->read_locus.p;
  s.lp = (const uint8_t *)b->lp.p;
  s.lp_off = (const uint64_t *)b->lp_off.p;
  s.rp = (const uint8_t *)b->rp.p;
  s.rp_off = (const uint64_t *)b->rp_off.p;
  return s;
}

size_t ring_ints_bound(int x, int oe, int e, int Pmax, int Tmax) {
  WfaProb pr;
  memset(&pr, 0, sizeof pr);
  pr.x = x; pr.oe = oe; pr.e = e; pr.P = Pmax; pr.T = Tmax;
  wfa_unband(pr);
  return wfa_ring_ints(pr);
}

int check_scoring(trgt_engine *e, int x, int o, int ext) {
  if (x < 1 || o < 0 || ext < 1 || x > 4096 || o > 4096 || ext > 4096)
    return fail(e, TRGT_ERR_ARG, "scoring (%d,%d,%d) unsupported: need mismatch>=1, gap_open>=0, gap_extend>=1", x, o, ext);
  return 0;
}

// WFA pass 2 launch shared by phases A and B
int launch_trace(trgt_engine *e, const WfaSrc &src, const uint32_t *work, const unsigned int *n_work_ptr,
                 uint32_t n_work_host, const WfaEnd *ends, unsigned long long max_trace_ints, DevBuf &gws,
                 double frac, trgt_flank_hit_t *hits, uint32_t *pool, unsigned long long pool_cap,
                 unsigned long long *cig_off, uint32_t *cig_n, int32_t *status, Counters *ctr) {
  if (n_work_host == 0) return 0;
  const int block = 128, wpb = block / 32;
  // 12 KB per warp: cones up to cost ~25 stay on chip.  A batch whose largest cone is far beyond that (alleles of
  // kilobases, costs in the hundreds) works out of the global slots anyway: it gets a token of shared memory and
  // four times the resident warps instead.
#ifndef TRGT_TRACE_LONG_FACTOR
#define TRGT_TRACE_LONG_FACTOR 16
#endif
  const size_t smem_cap_ints = max_trace_ints > (unsigned long long)TRGT_TRACE_LONG_FACTOR * 3072ull ? 512 : 3072;
  const int smem_ws_ints = (int)(max_trace_ints < smem_cap_ints ? (max_trace_ints ? max_trace_ints : 1) : smem_cap_ints);
  const size_t smem = (size_t)wpb * smem_ws_ints * sizeof(int);
  int grid = 0;
  TRY(persistent_grid(e, k_wfa_trace, block, smem, &grid));
  const uint32_t need_blocks = (n_work_host + wpb - 1) / wpb;
  if ((uint32_t)grid > need_blocks) grid = (int)need_blocks;
  size_t stride = 0;
  int *gws_p = nullptr;
  if (max_trace_ints > (unsigned long long)smem_ws_ints) {
    stride = (size_t)max_trace_ints;
    // fewer resident warps when one slot is huge
    size_t slots = (size_t)grid * wpb;
    const size_t budget_ints = e->workspace_budget / sizeof(int);
    if (stride > budget_ints) return fail(e, TRGT_ERR_INTERNAL, "trace workspace of %zu ints exceeds the budget", stride);
    if (slots * stride > budget_ints) {
      slots = budget_ints / stride;
      grid = (int)((slots + wpb - 1) / wpb);
      if (grid < 1) grid = 1;
      if ((size_t)grid * wpb > slots && grid > 1) grid--;
    }
    TRY(dev_reserve(e, gws, (size_t)grid * wpb * stride * sizeof(int)));
    gws_p = (int *)gws.p;
  }
  LaunchScope ls(e, "k_wfa_trace");
  k_wfa_trace<<<grid, block, smem, e->stream>>>(src, work, n_work_ptr, ends, gws_p, stride, smem_ws_ints, frac, hits,
                                                pool, pool_cap, cig_off, cig_n, status, ctr);
  return check_launch(e, "k_wfa_trace");
}

}  // namespace

extern "C" {

void trgt_flank_free(trgt_engine_t *e, trgt_flank_batch_t *b) {
  if (!b) return;
  if (e) {
    cudaSetDevice(e->device);
    engine_wait(e);
    if (e->one_flank == b) e->one_flank = nullptr;
  }
  DevBuf *all[] = {&b->reads, &b->read_off, &b->lp, &b->lp_off, &b->rp, &b->rp_off, &b->locus_read_off,
                   &b->read_locus, &b->hits, &b->spans, &b->work, &b->work2, &b->list1, &b->ends, &b->ctr, &b->gring, &b->gws,
                   &b->seq4, &b->seq4_starts, &b->seq4_len, &b->tr_len, &b->tr_off, &b->tr_data, &b->kidx};
  const int device = e ? e->device : -1;  // with an engine (its stream has drained): the buffers wait for the next batch
  for (auto *d : all) dev_free(*d, device);
  pin_free(b->h_tr_off, device);
  pin_free(b->h_tr_data, device);
  delete b;
}

static int flank_upload_into(trgt_engine_t *e, trgt_flank_batch *b, const trgt_seqs_t *left_pieces,
                             const trgt_seqs_t *right_pieces, const trgt_seqs_t *reads,
                             const uint32_t *locus_read_offsets, uint32_t n_loci, trgt_scoring_t scoring,
                             double min_flank_id_frac, bool defer_read_bytes = false,
                             bool skip_read_offsets = false) {
  TRY(check_seqs(e, left_pieces, "left_pieces"));
  TRY(check_seqs(e, right_pieces, "right_pieces"));
  TRY(check_seqs(e, reads, "reads"));
  TRY(check_scoring(e, scoring.mismatch, scoring.gap_open, scoring.gap_extend));
  if (left_pieces->n != n_loci || right_pieces->n != n_loci) return fail(e, TRGT_ERR_ARG, "need one left and one right piece per locus");
  if (n_loci && !locus_read_offsets) return fail(e, TRGT_ERR_ARG, "locus_read_offsets is null");
  if (reads->n > 0x7fffffffull) return fail(e, TRGT_ERR_ARG, "too many reads in one batch");
  for (uint32_t l = 0; l < n_loci; l++)
    if (locus_read_offsets[l + 1] < locus_read_offsets[l]) return fail(e, TRGT_ERR_ARG, "locus_read_offsets not monotone");
  if (n_loci && (locus_read_offsets[0] != 0 || locus_read_offsets[n_loci] != reads->n))
    return fail(e, TRGT_ERR_ARG, "locus_read_offsets must cover all reads");
  const uint64_t pm = max_len(left_pieces) > max_len(right_pieces) ? max_len(left_pieces) : max_len(right_pieces);
  const uint64_t tm = max_len(reads);
  if (pm + tm > 0x3fffffffull) return fail(e, TRGT_ERR_ARG, "sequence too long");
  b->n_loci = n_loci;
  b->n_reads = (uint32_t)reads->n;
  b->ran = false;
  b->Pmax = (int)pm;
  {
    uint64_t mn = pm;
    for (uint32_t l = 0; l < n_loci; l++) {
      const uint64_t a = left_pieces->offsets[l + 1] - left_pieces->offsets[l];
      const uint64_t c = right_pieces->offsets[l + 1] - right_pieces->offsets[l];
      if (a < mn) mn = a;
      if (c < mn) mn = c;
    }
    b->Pmin = (int)mn;
  }
  b->Tmax = (int)tm;
  b->scoring = scoring;
  b->frac = min_flank_id_frac;
  CU(e, cudaSetDevice(e->device));
  if (defer_read_bytes) {  // the caller streams the read bytes in chunks (flank_oneshot_locked)
    static const uint64_t zero_off[1] = {0};
    TRY(dev_reserve(e, b->reads, (size_t)(reads->n ? reads->offsets[reads->n] : 0) + 16));
    if (!skip_read_offsets)  // (seq4 input: the device derives the offsets from the read lengths)
      TRY(h2d(e, b->read_off, reads->n ? reads->offsets : zero_off, (size_t)(reads->n + 1) * sizeof(uint64_t)));
  } else {
    TRY(upload_seqs(e, reads, b->reads, b->read_off));
  }
  TRY(upload_seqs(e, left_pieces, b->lp, b->lp_off));
  TRY(upload_seqs(e, right_pieces, b->rp, b->rp_off));
  static const uint32_t zero32[1] = {0};
  TRY(h2d(e, b->locus_read_off, n_loci ? locus_read_offsets : zero32, ((size_t)n_loci + 1) * sizeof(uint32_t)));
  TRY(dev_reserve(e, b->read_locus, ((size_t)b->n_reads + 1) * sizeof(uint32_t)));
  TRY(dev_reserve(e, b->hits, ((size_t)b->n_reads * 2 + 1) * sizeof(trgt_flank_hit_t)));
  TRY(dev_reserve(e, b->spans, ((size_t)b->n_reads + 1) * sizeof(trgt_span_t)));
  TRY(dev_reserve(e, b->work, ((size_t)b->n_reads * 2 + 1) * sizeof(uint32_t)));
  TRY(dev_reserve(e, b->work2, ((size_t)b->n_reads * 2 + 1) * sizeof(uint32_t)));
  TRY(dev_reserve(e, b->list1, ((size_t)b->n_reads * 2 + 1) * sizeof(uint2)));
  TRY(dev_reserve(e, b->ends, ((size_t)b->n_reads * 2 + 1) * sizeof(WfaEnd)));
  TRY(dev_reserve(e, b->ctr, sizeof(Counters)));
  if (n_loci) {
    LaunchScope ls(e, "k_expand_offsets");
    k_expand_offsets<<<(n_loci + 255) / 256, 256, 0, e->stream>>>((const uint32_t *)b->locus_read_off.p, n_loci,
                                                                   (uint32_t *)b->read_locus.p);
    TRY(check_launch(e, "k_expand_offsets"));
  }
  return 0;
}

int32_t trgt_flank_upload(trgt_engine_t *e, const trgt_seqs_t *left_pieces, const trgt_seqs_t *right_pieces,
                          const trgt_seqs_t *reads, const uint32_t *locus_read_offsets, uint32_t n_loci,
                          trgt_scoring_t scoring, double min_flank_id_frac, trgt_flank_batch_t **out) {
  if (!e || !out) return TRGT_ERR_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  *out = nullptr;
  trgt_flank_batch *b = new trgt_flank_batch();
  int rc = flank_upload_into(e, b, left_pieces, right_pieces, reads, locus_read_offsets, n_loci, scoring,
                             min_flank_id_frac);
  // the copies were queued from the caller's buffers: with pinned sources they are asynchronous to the host, and
  // the header promises that no pointer is retained after return
  if (rc == 0 && engine_wait(e) != cudaSuccess) rc = fail(e, TRGT_ERR_CUDA, "flank upload failed");
  if (rc != 0) {
    trgt_flank_free(nullptr, b);
    return rc;
  }
  *out = b;
  return 0;
}

static int flank_launch_locate(trgt_engine_t *e, trgt_flank_batch *b, const WfaSrc &src, uint32_t l0, uint32_t l1) {
  if (l1 <= l0) return 0;
  const int block = 32;  // one warp per CTA, one locus at a time per warp
  b->kidx_valid = false;
  if (b->Pmin >= 16 && b->Pmax <= FXT_PMAX) {  // usual piece lengths: one lane per (read, flank) pair
    const int tb = 32 * FXT_WARPS;
    if (e->band_budget > 0) {
      TRY(dev_reserve(e, b->kidx, (size_t)b->n_loci * 2 * TRGT_KIDX_SLOTS * sizeof(uint16_t) + 16));
      b->kidx_valid = true;
    }
    int grid = 0;
    TRY(persistent_grid(e, k_flank_exact_t, tb, 0, &grid));
    const uint32_t need = (l1 - l0 + FXT_WARPS - 1) / FXT_WARPS;
    if ((uint32_t)grid > need) grid = (int)need;
    LaunchScope ls(e, "k_flank_exact_t");
    k_flank_exact_t<<<grid, tb, 0, e->stream>>>(src, (const uint32_t *)b->locus_read_off.p, l0, l1, e->band_budget,
                                                (trgt_flank_hit_t *)b->hits.p, (uint32_t *)b->work.p,
                                                (Counters *)b->ctr.p, b->kidx_valid ? (uint16_t *)b->kidx.p : nullptr);
    TRY(check_launch(e, "k_flank_exact_t"));
  } else {
    int grid = 0;
    TRY(persistent_grid(e, k_flank_exact, block, 0, &grid));
    if ((uint32_t)grid > l1 - l0) grid = (int)(l1 - l0);
    LaunchScope ls(e, "k_flank_exact");
    k_flank_exact<<<grid, block, 0, e->stream>>>(src, (const uint32_t *)b->locus_read_off.p, l0, l1, e->band_budget,
                                                 (trgt_flank_hit_t *)b->hits.p, (uint32_t *)b->work.p,
                                                 (Counters *)b->ctr.p);
    TRY(check_launch(e, "k_flank_exact"));
  }
  if (e->band_budget > 0 && b->kidx_valid) {  // first cost tier of the fallback, seed pass (the band pass runs in flank_finish)
    int grid = 0;
    TRY(persistent_grid(e, k_flank_seed, 128, 0, &grid));
    const uint32_t need = (l1 - l0 + FB_LOCI - 1) / FB_LOCI;
    if ((uint32_t)grid > need) grid = (int)need;
    LaunchScope ls(e, "k_flank_seed");
    k_flank_seed<<<grid, 128, 0, e->stream>>>(src, (const uint32_t *)b->locus_read_off.p, l0, l1, e->band_budget,
                                              (trgt_flank_hit_t *)b->hits.p, (uint2 *)b->list1.p, (uint32_t *)b->work2.p,
                                              (Counters *)b->ctr.p, (const uint16_t *)b->kidx.p);
    TRY(check_launch(e, "k_flank_seed"));
  } else if (e->band_budget > 0) {  // unusual piece lengths: the one-pass first tier
    int grid = 0;
    TRY(persistent_grid(e, k_flank_band, 128, 0, &grid));
    const uint32_t need = (l1 - l0 + FB_LOCI - 1) / FB_LOCI;
    if ((uint32_t)grid > need) grid = (int)need;
    LaunchScope ls(e, "k_flank_band");
    k_flank_band<<<grid, 128, 0, e->stream>>>(src, (const uint32_t *)b->locus_read_off.p, l0, l1, e->band_budget,
                                                b->frac, (trgt_flank_hit_t *)b->hits.p, (uint32_t *)b->work2.p,
                                                (Counters *)b->ctr.p, nullptr, /*prefetch=*/1);
    TRY(check_launch(e, "k_flank_band"));
  }
  return 0;
}

// pairs the on-chip path deferred: full-width score pass + cone trace; then the combine rule
static int flank_finish(trgt_engine_t *e, trgt_flank_batch *b, const WfaSrc &src) {
  Counters *ctr = (Counters *)b->ctr.p;
  if (e->band_budget > 0 && b->kidx_valid) {  // band pass of the first cost tier over the pairs the seed pass listed
    const int cap1 = e->band_budget < (src.x > src.oe ? src.x : src.oe) ? e->band_budget : (src.x > src.oe ? src.x : src.oe);
    int rows = 1;
    if (cap1 <= FT1_SMAX) {
      const unsigned live = ft1_live_scores(src.x, src.oe, src.e, cap1, nullptr);
      rows = __builtin_popcount(live);
    }
    const size_t smem = fb1_smem_bytes(rows);
    int grid = 0;
    TRY(persistent_grid(e, k_flank_band1, FB1_THREADS, smem, &grid));
    LaunchScope ls(e, "k_flank_band1");
    k_flank_band1<<<grid, FB1_THREADS, smem, e->stream>>>(src, (const uint2 *)b->list1.p, &ctr->n_list1, e->band_budget,
                                                          b->frac, src.reads + b->reads.cap, rows,
                                                          (trgt_flank_hit_t *)b->hits.p, (uint32_t *)b->work2.p, ctr);
    TRY(check_launch(e, "k_flank_band1"));
  }
  if (e->band_budget > 0) {  // second cost tier for what the first handed on, one warp per pair
    int grid = 0;
    TRY(persistent_grid(e, k_flank_band2, 32, 0, &grid));
    LaunchScope ls(e, "k_flank_band2");
    k_flank_band2<<<grid, 32, 0, e->stream>>>(src, (const uint32_t *)b->work2.p, &ctr->n_tier2, e->band_budget, b->frac,
                                              (trgt_flank_hit_t *)b->hits.p, (uint32_t *)b->work.p, ctr,
                                              b->kidx_valid ? (const uint16_t *)b->kidx.p : nullptr);
    TRY(check_launch(e, "k_flank_band2"));
  }
  if (e->band_budget > 0) {
    int grid = 0;
    TRY(persistent_grid(e, k_flank_band_wide, 32, 0, &grid));
    LaunchScope ls(e, "k_flank_band_wide");
    k_flank_band_wide<<<grid, 32, 0, e->stream>>>(src, (uint32_t *)b->work.p, &ctr->n_work, b->frac,
                                                  (trgt_flank_hit_t *)b->hits.p, ctr,
                                                  b->kidx_valid ? (const uint16_t *)b->kidx.p : nullptr);
    TRY(check_launch(e, "k_flank_band_wide"));
  }
  {
    const int block = 128;
    const size_t bound = ring_ints_bound(src.x, src.oe, src.e, b->Pmax, b->Tmax);
    const size_t cap_ints = 24 * 1024;  // 96 KB: two CTAs per SM
    const int smem_ring_ints = (int)(bound < cap_ints ? bound : cap_ints);
    const size_t smem = (size_t)(40 + smem_ring_ints) * sizeof(int);
    int grid = 0;
    TRY(persistent_grid(e, k_wfa_score<true>, block, smem, &grid));
    const uint32_t max_items = b->n_reads * 2;
    if ((uint32_t)grid > max_items) grid = (int)max_items;
    int *gring = nullptr;
    size_t stride = 0;
    if (bound > (size_t)smem_ring_ints) {
      stride = bound;
      TRY(cap_grid_to_budget(e, &grid, 1, stride, "flank ring"));
      TRY(dev_reserve(e, b->gring, (size_t)grid * stride * sizeof(int)));
      gring = (int *)b->gring.p;
    }
    LaunchScope ls(e, "k_wfa_score_block");
    k_wfa_score<true><<<grid, block, smem, e->stream>>>(src, (const uint32_t *)b->work.p, &ctr->n_work, 0,
                                                         (WfaEnd *)b->ends.p, gring, stride, smem_ring_ints, nullptr,
                                                         nullptr, ctr);
    TRY(check_launch(e, "k_wfa_score_block"));
  }
  CU(e, cudaMemcpyAsync(e->h_ctr, ctr, sizeof(Counters), cudaMemcpyDeviceToHost, e->stream));
  CU(e, engine_wait(e));
  b->last_n_work = e->h_ctr->n_work - e->h_ctr->n_banded;  // pairs that needed the full-width kernels
  b->last_n_tier2 = e->h_ctr->n_tier2;
  b->last_n_wide = e->h_ctr->n_work;
  TRY(launch_trace(e, src, (const uint32_t *)b->work.p, &ctr->n_work, e->h_ctr->n_work, (const WfaEnd *)b->ends.p,
                   e->h_ctr->max_trace_ints, b->gws, b->frac, (trgt_flank_hit_t *)b->hits.p, nullptr, 0, nullptr,
                   nullptr, nullptr, ctr));
  {
    LaunchScope ls(e, "k_flank_combine");
    const uint32_t grid = (b->n_reads + 255) / 256;
    k_flank_combine<<<grid, 256, 0, e->stream>>>((const trgt_flank_hit_t *)b->hits.p, b->n_reads, (trgt_span_t *)b->spans.p);
    TRY(check_launch(e, "k_flank_combine"));
  }
  b->ran = true;
  return 0;
}

static int flank_run_locked(trgt_engine_t *e, trgt_flank_batch *b) {
  CU(e, cudaSetDevice(e->device));
  if (b->n_reads == 0) return 0;
  const WfaSrc src = flank_src(b);
  CU(e, cudaMemsetAsync(b->ctr.p, 0, sizeof(Counters), e->stream));
  TRY(flank_launch_locate(e, b, src, 0, b->n_loci));
  return flank_finish(e, b, src);
}

// One-shot phase A from host buffers: the read bytes go up in chunks on the copy stream while the
// locate kernel works on the chunks that have landed.
static int flank_oneshot_locked(trgt_engine_t *e, trgt_flank_batch *b, const trgt_seqs_t *reads,
                                const uint32_t *locus_read_offsets, uint32_t n_loci) {
  CU(e, cudaSetDevice(e->device));
  if (b->n_reads == 0) return 0;
  const WfaSrc src = flank_src(b);
  CU(e, cudaMemsetAsync(b->ctr.p, 0, sizeof(Counters), e->stream));
  // the copy stream may start once everything queued so far (allocations, small uploads) is done
  cudaEvent_t ready = get_event(e);
  CU(e, cudaEventRecord(ready, e->stream));
  CU(e, cudaStreamWaitEvent(e->copy_stream, ready, 0));
  const uint64_t total = reads->offsets[reads->n];
  const uint64_t chunk_bytes = 192ull << 20;
  std::vector<cudaEvent_t> evs;
  uint32_t l0 = 0;
  while (l0 < n_loci) {
    uint32_t l1 = l0 + 1;
    const uint64_t b0 = reads->offsets[locus_read_offsets[l0]];
    while (l1 < n_loci && reads->offsets[locus_read_offsets[l1 + 1]] - b0 <= chunk_bytes) l1++;
    const uint64_t b1 = reads->offsets[locus_read_offsets[l1]];
    if (b1 > b0)
      CU(e, cudaMemcpyAsync((uint8_t *)b->reads.p + b0, reads->data + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, e->copy_stream));
    cudaEvent_t ev = get_event(e);
    evs.push_back(ev);
    CU(e, cudaEventRecord(ev, e->copy_stream));
    CU(e, cudaStreamWaitEvent(e->stream, ev, 0));
    TRY(flank_launch_locate(e, b, src, l0, l1));
    l0 = l1;
  }
  (void)total;
  const int rc = flank_finish(e, b, src);  // synchronises e->stream, hence every chunk event
  e->event_pool.push_back(ready);
  for (auto ev : evs) e->event_pool.push_back(ev);
  return rc;
}

// ---- reads handed over as BAM 4-bit bases (trgt_seq4_t) ----

#define SEQ4_PAD 16  // bytes of padding on both sides of the packed buffer (seq4_decode16 reads whole words)

static int check_seq4(trgt_engine_t *e, const trgt_seq4_t *r, uint64_t *total_out, uint64_t *max_out) {
  if (!r || (r->n && (!r->starts || !r->lengths))) return fail(e, TRGT_ERR_ARG, "reads: null seq4 set");
  if (r->n > 0x7fffffffull) return fail(e, TRGT_ERR_ARG, "too many reads in one batch");
  uint64_t total = 0, mx = 0, prev = 0;
  for (uint64_t i = 0; i < r->n; i++) {
    const uint64_t s0 = r->starts[i], len = r->lengths[i];
    if (s0 < prev) return fail(e, TRGT_ERR_ARG, "reads: seq4 starts must be non-decreasing (read %llu)", (unsigned long long)i);
    if (s0 + len > 2 * r->data_bytes) return fail(e, TRGT_ERR_ARG, "reads: seq4 read %llu runs past data_bytes", (unsigned long long)i);
    prev = s0;
    total += len;
    if (len > mx) mx = len;
  }
  if (total && !r->data) return fail(e, TRGT_ERR_ARG, "reads: null data");
  *total_out = total;
  *max_out = mx;
  return 0;
}

// device side of a seq4 read set: starts / lengths up, ASCII CSR offsets by a scan of the lengths;
// the packed bytes themselves are copied by the caller (at once or in chunks)
static int seq4_prepare(trgt_engine_t *e, const trgt_seq4_t *r, uint64_t total, DevBuf &d_seq4, DevBuf &d_starts,
                        DevBuf &d_len, DevBuf &d_ascii, DevBuf &d_off) {
  static const uint64_t zero64[1] = {0};
  static const uint32_t zero32[1] = {0};
  TRY(dev_reserve(e, d_seq4, (size_t)r->data_bytes + 2 * SEQ4_PAD));
  TRY(h2d(e, d_starts, r->n ? r->starts : zero64, (size_t)(r->n ? r->n : 1) * sizeof(uint64_t)));
  TRY(h2d(e, d_len, r->n ? r->lengths : zero32, (size_t)(r->n ? r->n : 1) * sizeof(uint32_t)));
  CU(e, cudaMemsetAsync((uint8_t *)d_len.p + (size_t)r->n * sizeof(uint32_t), 0, sizeof(uint32_t), e->stream));
  TRY(dev_reserve(e, d_ascii, (size_t)total + 16));
  TRY(dev_reserve(e, d_off, (size_t)(r->n + 1) * sizeof(uint64_t) + 16));
  TRY(exclusive_scan_u32(e, (const uint32_t *)d_len.p, (unsigned long long *)d_off.p, (size_t)r->n + 1));
  return 0;
}

static int launch_unpack(trgt_engine_t *e, const DevBuf &d_seq4, const DevBuf &d_starts, const DevBuf &d_len,
                         const DevBuf &d_off, DevBuf &d_ascii, uint32_t r0, uint32_t r1) {
  if (r1 <= r0) return 0;
  const int block = 256, wpb = block / 32;
  int grid = 0;
  TRY(persistent_grid(e, k_unpack_seq4, block, 0, &grid));
  const uint32_t need = (r1 - r0 + wpb - 1) / wpb;
  if ((uint32_t)grid > need) grid = (int)need;
  LaunchScope ls(e, "k_unpack_seq4");
  k_unpack_seq4<<<grid, block, 0, e->stream>>>((const uint8_t *)d_seq4.p + SEQ4_PAD, (const uint64_t *)d_starts.p,
                                               (const uint32_t *)d_len.p, (const unsigned long long *)d_off.p, r0, r1,
                                               (uint8_t *)d_ascii.p);
  return check_launch(e, "k_unpack_seq4");
}

// upload of everything but the packed read bytes
static int flank_upload_seq4_into(trgt_engine_t *e, trgt_flank_batch *b, const trgt_seqs_t *left_pieces,
                                  const trgt_seqs_t *right_pieces, const trgt_seq4_t *reads,
                                  const uint32_t *locus_read_offsets, uint32_t n_loci, trgt_scoring_t scoring,
                                  double min_flank_id_frac) {
  uint64_t total = 0, mx = 0;
  TRY(check_seq4(e, reads, &total, &mx));
  // flank_upload_into validates and sizes everything from a CSR view of the reads: build the ASCII
  // offsets once on the host (never uploaded: the device derives its own by a scan of the lengths)
  std::vector<uint64_t> offs((size_t)reads->n + 1);
  offs[0] = 0;
  for (uint64_t i = 0; i < reads->n; i++) offs[i + 1] = offs[i] + reads->lengths[i];
  static const uint8_t dummy = 0;
  trgt_seqs_t view;
  view.data = &dummy; view.offsets = offs.data(); view.n = reads->n;
  TRY(flank_upload_into(e, b, left_pieces, right_pieces, &view, locus_read_offsets, n_loci, scoring, min_flank_id_frac,
                        /*defer_read_bytes=*/true, /*skip_read_offsets=*/true));
  return seq4_prepare(e, reads, total, b->seq4, b->seq4_starts, b->seq4_len, b->reads, b->read_off);
}

int32_t trgt_flank_upload_seq4(trgt_engine_t *e, const trgt_seqs_t *left_pieces, const trgt_seqs_t *right_pieces,
                               const trgt_seq4_t *reads, const uint32_t *locus_read_offsets, uint32_t n_loci,
                               trgt_scoring_t scoring, double min_flank_id_frac, trgt_flank_batch_t **out) {
  if (!e || !out) return TRGT_ERR_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  *out = nullptr;
  trgt_flank_batch *b = new trgt_flank_batch();
  int rc = flank_upload_seq4_into(e, b, left_pieces, right_pieces, reads, locus_read_offsets, n_loci, scoring,
                                  min_flank_id_frac);
  if (rc == 0 && reads->data_bytes) {
    cudaError_t err = cudaMemcpyAsync((uint8_t *)b->seq4.p + SEQ4_PAD, reads->data, (size_t)reads->data_bytes,
                                      cudaMemcpyHostToDevice, e->stream);
    if (err != cudaSuccess) rc = fail(e, TRGT_ERR_CUDA, "seq4 upload failed: %s", cudaGetErrorString(err));
  }
  if (rc == 0) rc = launch_unpack(e, b->seq4, b->seq4_starts, b->seq4_len, b->read_off, b->reads, 0, b->n_reads);
  if (rc == 0 && engine_wait(e) != cudaSuccess) rc = fail(e, TRGT_ERR_CUDA, "seq4 decode failed");
  if (rc != 0) {
    trgt_flank_free(nullptr, b);
    return rc;
  }
  *out = b;
  return 0;
}

// One-shot phase A from BAM 4-bit reads: packed bytes go up in chunks of loci on the copy stream;
// each chunk is decoded and searched as soon as it has landed.
static int flank_oneshot_seq4_locked(trgt_engine_t *e, trgt_flank_batch *b, const trgt_seq4_t *reads,
                                     const uint32_t *locus_read_offsets, uint32_t n_loci) {
  CU(e, cudaSetDevice(e->device));
  if (b->n_reads == 0) return 0;
  const WfaSrc src = flank_src(b);
  CU(e, cudaMemsetAsync(b->ctr.p, 0, sizeof(Counters), e->stream));
  cudaEvent_t ready = get_event(e);
  CU(e, cudaEventRecord(ready, e->stream));
  CU(e, cudaStreamWaitEvent(e->copy_stream, ready, 0));
  const uint64_t chunk_bytes = 96ull << 20;
  std::vector<cudaEvent_t> evs;
  // byte range of the packed buffer that reads [ra, rb) touch (starts are non-decreasing)
  auto first_byte = [&](uint32_t r) { return r < b->n_reads ? reads->starts[r] >> 1 : reads->data_bytes; };
  uint64_t sent = 0;  // bytes [0, sent) are queued
  uint32_t l0 = 0;
  while (l0 < n_loci) {
    uint32_t l1 = l0 + 1;
    const uint64_t b0 = first_byte(locus_read_offsets[l0]);
    while (l1 < n_loci && first_byte(locus_read_offsets[l1 + 1]) - b0 <= chunk_bytes) l1++;
    const uint32_t ra = locus_read_offsets[l0], rb = locus_read_offsets[l1];
    // everything up to the end of the chunk's last read (a later read never starts before an earlier one,
    // but an earlier read may end after a later one starts: take the maximum end)
    uint64_t end = sent;
    for (uint32_t r = ra; r < rb; r++) {
      const uint64_t e_r = (reads->starts[r] + reads->lengths[r] + 1) >> 1;
      if (e_r > end) end = e_r;
    }
    if (l1 == n_loci) end = reads->data_bytes > end ? reads->data_bytes : end;
    if (end > reads->data_bytes) end = reads->data_bytes;
    if (end > sent) {
      CU(e, cudaMemcpyAsync((uint8_t *)b->seq4.p + SEQ4_PAD + sent, reads->data + sent, (size_t)(end - sent),
                            cudaMemcpyHostToDevice, e->copy_stream));
      sent = end;
    }
    cudaEvent_t ev = get_event(e);
    evs.push_back(ev);
    CU(e, cudaEventRecord(ev, e->copy_stream));
    CU(e, cudaStreamWaitEvent(e->stream, ev, 0));
    TRY(launch_unpack(e, b->seq4, b->seq4_starts, b->seq4_len, b->read_off, b->reads, ra, rb));
    TRY(flank_launch_locate(e, b, src, l0, l1));
    l0 = l1;
  }
  const int rc = flank_finish(e, b, src);
  e->event_pool.push_back(ready);
  for (auto ev : evs) e->event_pool.push_back(ev);
  return rc;
}

int32_t trgt_flank_run(trgt_engine_t *e, trgt_flank_batch_t *b) {
  if (!e || !b) return TRGT_ERR_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  return flank_run_locked(e, b);
}

static int flank_download_locked(trgt_engine_t *e, trgt_flank_batch *b, trgt_span_t *spans_out,
                                 trgt_flank_hit_t *hits_out) {
  if (b->n_reads && !b->ran) return fail(e, TRGT_ERR_ARG, "trgt_flank_download before trgt_flank_run");
  CU(e, cudaSetDevice(e->device));
  if (b->n_reads) {
    if (spans_out)
      CU(e, cudaMemcpyAsync(spans_out, b->spans.p, (size_t)b->n_reads * sizeof(trgt_span_t), cudaMemcpyDeviceToHost, e->stream));
    if (hits_out)
      CU(e, cudaMemcpyAsync(hits_out, b->hits.p, (size_t)b->n_reads * 2 * sizeof(trgt_flank_hit_t), cudaMemcpyDeviceToHost, e->stream));
  }
  CU(e, engine_wait(e));
  return 0;
}

int32_t trgt_flank_download(trgt_engine_t *e, trgt_flank_batch_t *b, trgt_span_t *spans_out, trgt_flank_hit_t *hits_out) {
  if (!e || !b) return TRGT_ERR_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  return flank_download_locked(e, b, spans_out, hits_out);
}

int32_t trgt_flank_spans(trgt_engine_t *e, const trgt_seqs_t *left_pieces, const trgt_seqs_t *right_pieces,
                         const trgt_seqs_t *reads, const uint32_t *locus_read_offsets, uint32_t n_loci,
                         trgt_scoring_t scoring, double min_flank_id_frac, trgt_span_t *spans_out,
                         trgt_flank_hit_t *hits_out) {
  if (!e) return TRGT_ERR_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  if (!e->one_flank) e->one_flank = new trgt_flank_batch();
  TRY(flank_upload_into(e, e->one_flank, left_pieces, right_pieces, reads, locus_read_offsets, n_loci, scoring,
                        min_flank_id_frac, /*defer_read_bytes=*/true));
  TRY(flank_oneshot_locked(e, e->one_flank, reads, locus_read_offsets, n_loci));
  return flank_download_locked(e, e->one_flank, spans_out, hits_out);
}

int32_t trgt_flank_spans_seq4(trgt_engine_t *e, const trgt_seqs_t *left_pieces, const trgt_seqs_t *right_pieces,
                              const trgt_seq4_t *reads, const uint32_t *locus_read_offsets, uint32_t n_loci,
                              trgt_scoring_t scoring, double min_flank_id_frac, trgt_span_t *spans_out,
                              trgt_flank_hit_t *hits_out) {
  if (!e) return TRGT_ERR_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  if (!e->one_flank) e->one_flank = new trgt_flank_batch();
  static const bool trace = getenv("TRGT_TRACE") != nullptr;  // host-side stage times of this call, to stderr
  const auto t0 = std::chrono::steady_clock::now();
  TRY(flank_upload_seq4_into(e, e->one_flank, left_pieces, right_pieces, reads, locus_read_offsets, n_loci, scoring,
                             min_flank_id_frac));
  const auto t1 = std::chrono::steady_clock::now();
  TRY(flank_oneshot_seq4_locked(e, e->one_flank, reads, locus_read_offsets, n_loci));
  const auto t2 = std::chrono::steady_clock::now();
  const int rc = flank_download_locked(e, e->one_flank, spans_out, hits_out);
  if (trace) {
    const auto t3 = std::chrono::steady_clock::now();
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
      return std::chrono::duration<double, std::milli>(b - a).count();
    };
    fprintf(stderr, "[trgt] flank_spans_seq4: checks + small uploads %.3f ms, reads + kernels %.3f ms, download %.3f ms\n",
            ms(t0, t1), ms(t1, t2), ms(t2, t3));
  }
  return rc;
}

int32_t trgt_seq4_decode(trgt_engine_t *e, const trgt_seq4_t *reads, uint8_t *ascii_out, uint64_t *ascii_offsets_out) {
  if (!e || !ascii_offsets_out) return TRGT_ERR_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  CU(e, cudaSetDevice(e->device));
  uint64_t total = 0, mx = 0;
  TRY(check_seq4(e, reads, &total, &mx));
  if (total && !ascii_out) return fail(e, TRGT_ERR_ARG, "ascii_out is null");
  DevBuf *d = e->d_ed;  // scratch shared with trgt_edit_dist (calls on one engine serialise)
  TRY(seq4_prepare(e, reads, total, d[0], d[1], d[2], d[3], d[4]));
  if (reads->data_bytes)
    CU(e, cudaMemcpyAsync((uint8_t *)d[0].p + SEQ4_PAD, reads->data, (size_t)reads->data_bytes, cudaMemcpyHostToDevice, e->stream));
  TRY(launch_unpack(e, d[0], d[1], d[2], d[4], d[3], 0, (uint32_t)reads->n));
  if (total) CU(e, cudaMemcpyAsync(ascii_out, d[3].p, (size_t)total, cudaMemcpyDeviceToHost, e->stream));
  CU(e, cudaMemcpyAsync(ascii_offsets_out, d[4].p, (size_t)(reads->n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
  CU(e, engine_wait(e));
  return 0;
}

int32_t trgt_bamlet_clip(trgt_engine_t *e, trgt_flank_batch_t *b, const uint32_t *cigar_ops,
                         const uint64_t *cigar_offsets, const int64_t *ref_starts, uint32_t flank_len,
                         trgt_bamlet_clip_t *clips_out) {
  if (!e) return TRGT_ERR_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  if (!b) b = e->one_flank;  // the batch of the last one-shot trgt_flank_spans* call
  if (!b || (b->n_reads && !b->ran)) return fail(e, TRGT_ERR_ARG, "trgt_bamlet_clip: the flank batch has not been run");
  CU(e, cudaSetDevice(e->device));
  const uint64_t n_reads = b->n_reads;
  if (n_reads == 0) return 0;
  if (!cigar_offsets || !ref_starts || !clips_out) return fail(e, TRGT_ERR_ARG, "trgt_bamlet_clip: null argument");
  for (uint64_t i = 0; i < n_reads; i++)
    if (cigar_offsets[i + 1] < cigar_offsets[i]) return fail(e, TRGT_ERR_ARG, "cigar_offsets not monotone");
  const uint64_t n_ops = cigar_offsets[n_reads];
  if (n_ops && !cigar_ops) return fail(e, TRGT_ERR_ARG, "cigar_ops is null");
  DevBuf *d = e->d_ed;
  TRY(h2d(e, d[0], cigar_ops, (size_t)n_ops * sizeof(uint32_t)));
  TRY(h2d(e, d[1], cigar_offsets, (size_t)(n_reads + 1) * sizeof(uint64_t)));
  TRY(h2d(e, d[2], ref_starts, (size_t)n_reads * sizeof(int64_t)));
  TRY(dev_reserve(e, d[5], (size_t)n_reads * sizeof(trgt_bamlet_clip_t)));
  {
    int grid = 0;
    TRY(persistent_grid(e, k_bamlet_clip, 256, 0, &grid));
    const uint32_t need = (uint32_t)((n_reads + 7) / 8);
    if ((uint32_t)grid > need) grid = (int)need;
    LaunchScope ls(e, "k_bamlet_clip");
    k_bamlet_clip<<<grid, 256, 0, e->stream>>>((const uint8_t *)b->reads.p, (const uint64_t *)b->read_off.p,
                                               (const trgt_span_t *)b->spans.p, (const uint32_t *)d[0].p,
                                               (const uint64_t *)d[1].p, (const long long *)d[2].p, flank_len,
                                               (uint32_t)n_reads, (trgt_bamlet_clip_t *)d[5].p);
    TRY(check_launch(e, "k_bamlet_clip"));
  }
  CU(e, cudaMemcpyAsync(clips_out, d[5].p, (size_t)n_reads * sizeof(trgt_bamlet_clip_t), cudaMemcpyDeviceToHost, e->stream));
  CU(e, engine_wait(e));
  return 0;
}

int32_t trgt_clip_reads(trgt_engine_t *e, const uint32_t *cigar_ops, const uint64_t *cigar_offsets,
                        const int64_t *ref_starts, uint64_t n_reads, const int64_t *regions,
                        const uint32_t *locus_read_offsets, uint32_t n_loci, trgt_clip_t *clips_out) {
  if (!e) return TRGT_ERR_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  CU(e, cudaSetDevice(e->device));
  if (n_reads == 0) return 0;
  if (!cigar_offsets || !ref_starts || !regions || !locus_read_offsets || !clips_out || n_loci == 0)
    return fail(e, TRGT_ERR_ARG, "trgt_clip_reads: null argument");
  if (n_reads > 0x7fffffffull) return fail(e, TRGT_ERR_ARG, "too many reads in one batch");
  for (uint64_t i = 0; i < n_reads; i++)
    if (cigar_offsets[i + 1] < cigar_offsets[i]) return fail(e, TRGT_ERR_ARG, "cigar_offsets not monotone");
  for (uint32_t l = 0; l < n_loci; l++)
    if (locus_read_offsets[l + 1] < locus_read_offsets[l]) return fail(e, TRGT_ERR_ARG, "locus_read_offsets not monotone");
  if (locus_read_offsets[0] != 0 || locus_read_offsets[n_loci] != n_reads)
    return fail(e, TRGT_ERR_ARG, "locus_read_offsets must cover all reads");
  const uint64_t n_ops = cigar_offsets[n_reads];
  if (n_ops && !cigar_ops) return fail(e, TRGT_ERR_ARG, "cigar_ops is null");
  DevBuf *d = e->d_ed;
  TRY(h2d(e, d[0], cigar_ops, (size_t)n_ops * sizeof(uint32_t)));
  TRY(h2d(e, d[1], cigar_offsets, (size_t)(n_reads + 1) * sizeof(uint64_t)));
  TRY(h2d(e, d[2], ref_starts, (size_t)n_reads * sizeof(int64_t)));
  TRY(h2d(e, d[3], regions, (size_t)n_loci * 2 * sizeof(int64_t)));
  TRY(h2d(e, d[4], locus_read_offsets, ((size_t)n_loci + 1) * sizeof(uint32_t)));
  // read -> locus map and the clips share one buffer: [n_reads] uint32 (padded to 16 B) then trgt_clip_t[n_reads]
  const size_t map_bytes = (((size_t)n_reads + 1) * sizeof(uint32_t) + 15) & ~(size_t)15;
  TRY(dev_reserve(e, d[5], map_bytes + (size_t)n_reads * sizeof(trgt_clip_t)));
  uint32_t *d_map = (uint32_t *)d[5].p;
  trgt_clip_t *d_clips = (trgt_clip_t *)((uint8_t *)d[5].p + map_bytes);
  {
    LaunchScope ls(e, "k_expand_offsets");
    k_expand_offsets<<<(n_loci + 255) / 256, 256, 0, e->stream>>>((const uint32_t *)d[4].p, n_loci, d_map);
    TRY(check_launch(e, "k_expand_offsets"));
  }
  {
    int grid = 0;
    TRY(persistent_grid(e, k_clip_cigar, 256, 0, &grid));
    const uint32_t need = (uint32_t)((n_reads + 255) / 256);
    if ((uint32_t)grid > need) grid = (int)need;
    LaunchScope ls(e, "k_clip_cigar");
    k_clip_cigar<<<grid, 256, 0, e->stream>>>((const uint32_t *)d[0].p, (const uint64_t *)d[1].p, (const long long *)d[2].p,
                                              d_map, (const long long *)d[3].p, (uint32_t)n_reads, d_clips);
    TRY(check_launch(e, "k_clip_cigar"));
  }
  CU(e, cudaMemcpyAsync(clips_out, d_clips, (size_t)n_reads * sizeof(trgt_clip_t), cudaMemcpyDeviceToHost, e->stream));
  CU(e, engine_wait(e));
  return 0;
}

int32_t trgt_flank_trs(trgt_engine_t *e, trgt_flank_batch_t *b, trgt_seqs_out_t *out) {
  if (!e || !out) return TRGT_ERR_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  if (!b) b = e->one_flank;
  if (!b || (b->n_reads && !b->ran)) return fail(e, TRGT_ERR_ARG, "trgt_flank_trs: the flank batch has not been run");
  CU(e, cudaSetDevice(e->device));
  memset(out, 0, sizeof *out);
  const uint32_t n = b->n_reads;
  TRY(pin_reserve(e, b->h_tr_off, ((size_t)n + 1) * sizeof(uint64_t)));
  uint64_t *h_off = b->h_tr_off.as<uint64_t>();
  h_off[0] = 0;
  out->n = n;
  out->offsets = h_off;
  if (n == 0) return 0;
  TRY(dev_reserve(e, b->tr_len, ((size_t)n + 1) * sizeof(uint32_t)));
  TRY(dev_reserve(e, b->tr_off, ((size_t)n + 1) * sizeof(uint64_t)));
  {
    LaunchScope ls(e, "k_tr_len");
    k_tr_len<<<(n + 256) / 256, 256, 0, e->stream>>>((const trgt_span_t *)b->spans.p, n, (uint32_t *)b->tr_len.p);
    TRY(check_launch(e, "k_tr_len"));
  }
  TRY(exclusive_scan_u32(e, (const uint32_t *)b->tr_len.p, (unsigned long long *)b->tr_off.p, (size_t)n + 1));
  CU(e, cudaMemcpyAsync(h_off, b->tr_off.p, ((size_t)n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
  CU(e, engine_wait(e));
  const uint64_t total = h_off[n];
  TRY(dev_reserve(e, b->tr_data, (size_t)total + 16));
  TRY(pin_reserve(e, b->h_tr_data, (size_t)total + 16));
  if (total) {
    int grid = 0;
    TRY(persistent_grid(e, k_tr_gather, 256, 0, &grid));
    const uint32_t need = (n + 7) / 8;
    if ((uint32_t)grid > need) grid = (int)need;
    {
      LaunchScope ls(e, "k_tr_gather");
      k_tr_gather<<<grid, 256, 0, e->stream>>>((const uint8_t *)b->reads.p, (const uint64_t *)b->read_off.p,
                                               (const trgt_span_t *)b->spans.p, (const unsigned long long *)b->tr_off.p, n,
                                               (uint8_t *)b->tr_data.p);
      TRY(check_launch(e, "k_tr_gather"));
    }
    CU(e, cudaMemcpyAsync(b->h_tr_data.p, b->tr_data.p, (size_t)total, cudaMemcpyDeviceToHost, e->stream));
    CU(e, engine_wait(e));
  }
  out->data = b->h_tr_data.as<uint8_t>();
  return 0;
}

/* device pointers of a resident flank batch (for device-side consumers such as the pipeline) */
int32_t trgt_flank_device_views(trgt_flank_batch_t *b, const void **d_reads, const void **d_read_off,
                                const void **d_spans, const void **d_hits, uint32_t *n_reads, uint32_t *n_wfa) {
  if (!b) return TRGT_ERR_ARG;
  if (d_reads) *d_reads = b->reads.p;
  if (d_read_off) *d_read_off = b->read_off.p;
  if (d_spans) *d_spans = b->spans.p;
  if (d_hits) *d_hits = b->hits.p;
  if (n_reads) *n_reads = b->n_reads;
  if (n_wfa) *n_wfa = b->last_n_work;
  return 0;
}

/* how the last run settled the pairs that missed the exact search: out[0] = handed to the second cost
 * tier, out[1] = handed to the wide-band kernel, out[2] = needed the full-width kernels */
int32_t trgt_flank_fallback_counts(trgt_flank_batch_t *b, uint32_t out[3]) {
  if (!b || !out) return TRGT_ERR_ARG;
  out[0] = b->last_n_tier2; out[1] = b->last_n_wide; out[2] = b->last_n_work;
  return 0;
}

}  // extern "C"

// ------------------------------------------------------------------ phase B: align ---------

struct trgt_align_batch {
  uint32_t n_groups = 0, n_seqs = 0;
  int Pmax = 0, Tmax = 0;
  DevBuf bb, bb_off, seqs, seq_off, group_off, seq_group;
  DevBuf ends, trace_work, resid, diff, cig_n, cig_off, pool, ctr, gring, gws;
  DevBuf out_off, out_words, scores, status;
  DevBuf cons_counts, cons_recs, cons_len, cons_status, cons_off, cons_data;
  PinBuf h_cons_off, h_cons_data, h_cons_status;
  // host copies handed out by download
  PinBuf h_off, h_words, h_scores, h_status;
  unsigned long long total_words = 0;
  bool ran = false;
};

namespace {

WfaSrc align_src(const trgt_align_batch *b) {
  WfaSrc s;
  memset(&s, 0, sizeof s);
  s.mode = WFA_MODE_E2E;
  s.x = 2; s.oe = 5 + 1; s.e = 1;  // create_thread_local_ga_aligner: affine(2,5,1), commands/genotype.rs:82-86
  s.bb = (const uint8_t *)b->bb.p;
  s.bb_off = (const uint64_t *)b->bb_off.p;
  s.seqs = (const uint8_t *)b->seqs.p;
  s.seq_off = (const uint64_t *)b->seq_off.p;
  s.seq_group = (const uint32_t *)b->seq_group.p;
  return s;
}

}  // namespace

extern "C" {

void trgt_align_free(trgt_engine_t *e, trgt_align_batch_t *b) {
  if (!b) return;
  if (e) {
    cudaSetDevice(e->device);
    engine_wait(e);
    if (e->one_align == b) e->one_align = nullptr;
  }
  DevBuf *all[] = {&b->bb, &b->bb_off, &b->seqs, &b->seq_off, &b->group_off, &b->seq_group, &b->ends, &b->trace_work,
                   &b->cig_n, &b->cig_off, &b->pool, &b->ctr, &b->gring, &b->gws, &b->out_off, &b->out_words,
                   &b->scores, &b->status, &b->cons_counts, &b->cons_recs, &b->cons_len, &b->cons_status,
                   &b->cons_off, &b->cons_data, &b->resid, &b->diff};
  for (auto *d : all) dev_free(*d);
  pin_free(b->h_off); pin_free(b->h_words); pin_free(b->h_scores); pin_free(b->h_status);
  pin_free(b->h_cons_off); pin_free(b->h_cons_data); pin_free(b->h_cons_status);
  delete b;
}

static int align_prepare(trgt_engine_t *e, trgt_align_batch *b, const uint32_t *group_seq_offsets, uint32_t n_groups,
                         uint32_t n_seqs, int pm, int tm);

static int align_upload_into(trgt_engine_t *e, trgt_align_batch *b, const trgt_seqs_t *backbones,
                             const trgt_seqs_t *seqs, const uint32_t *group_seq_offsets, uint32_t n_groups) {
  TRY(check_seqs(e, backbones, "backbones"));
  TRY(check_seqs(e, seqs, "seqs"));
  if (backbones->n != n_groups) return fail(e, TRGT_ERR_ARG, "need one backbone per group");
  if (n_groups && !group_seq_offsets) return fail(e, TRGT_ERR_ARG, "group_seq_offsets is null");
  if (seqs->n > 0x7fffffffull) return fail(e, TRGT_ERR_ARG, "too many sequences in one batch");
  for (uint32_t g = 0; g < n_groups; g++)
    if (group_seq_offsets[g + 1] < group_seq_offsets[g]) return fail(e, TRGT_ERR_ARG, "group_seq_offsets not monotone");
  if (n_groups && (group_seq_offsets[0] != 0 || group_seq_offsets[n_groups] != seqs->n))
    return fail(e, TRGT_ERR_ARG, "group_seq_offsets must cover all sequences");
  if (!n_groups && seqs->n) return fail(e, TRGT_ERR_ARG, "sequences without groups");
  const uint64_t pm = max_len(backbones), tm = max_len(seqs);
  if (pm + tm > 0x0fffffffull) return fail(e, TRGT_ERR_ARG, "sequence too long");
  CU(e, cudaSetDevice(e->device));
  TRY(upload_seqs(e, backbones, b->bb, b->bb_off));
  TRY(upload_seqs(e, seqs, b->seqs, b->seq_off));
  return align_prepare(e, b, group_seq_offsets, n_groups, (uint32_t)seqs->n, (int)pm, (int)tm);
}

// the rest of an align batch once backbones and member sequences are on the device (uploaded, or gathered there)
static int align_prepare(trgt_engine_t *e, trgt_align_batch *b, const uint32_t *group_seq_offsets, uint32_t n_groups,
                         uint32_t n_seqs, int pm, int tm) {
  b->n_groups = n_groups;
  b->n_seqs = n_seqs;
  b->Pmax = pm;
  b->Tmax = tm;
  b->ran = false;
  static const uint32_t zero32[1] = {0};
  TRY(h2d(e, b->group_off, n_groups ? group_seq_offsets : zero32, ((size_t)n_groups + 1) * sizeof(uint32_t)));
  const size_t n = b->n_seqs;
  TRY(dev_reserve(e, b->seq_group, (n + 1) * sizeof(uint32_t)));
  TRY(dev_reserve(e, b->ends, (n + 1) * sizeof(WfaEnd)));
  TRY(dev_reserve(e, b->trace_work, (n + 1) * sizeof(uint32_t)));
  TRY(dev_reserve(e, b->cig_n, (n + 1) * sizeof(uint32_t)));
  TRY(dev_reserve(e, b->cig_off, (n + 1) * sizeof(unsigned long long)));
  TRY(dev_reserve(e, b->out_off, (n + 1) * sizeof(unsigned long long)));
  TRY(dev_reserve(e, b->scores, (n + 1) * sizeof(int32_t)));
  TRY(dev_reserve(e, b->status, (n + 1) * sizeof(int32_t)));
  TRY(dev_reserve(e, b->ctr, sizeof(Counters)));
  if (n_groups) {
    LaunchScope ls(e, "k_expand_offsets");
    k_expand_offsets<<<(n_groups + 255) / 256, 256, 0, e->stream>>>((const uint32_t *)b->group_off.p, n_groups,
                                                                     (uint32_t *)b->seq_group.p);
    TRY(check_launch(e, "k_expand_offsets"));
  }
  return 0;
}

int32_t trgt_align_upload(trgt_engine_t *e, const trgt_seqs_t *backbones, const trgt_seqs_t *seqs,
                          const uint32_t *group_seq_offsets, uint32_t n_groups, trgt_align_batch_t **out) {
  if (!e || !out) return TRGT_ERR_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  *out = nullptr;
  trgt_align_batch *b = new trgt_align_batch();
  int rc = align_upload_into(e, b, backbones, seqs, group_seq_offsets, n_groups);
  // as trgt_flank_upload: the caller's buffers are free again when this returns
  if (rc == 0 && engine_wait(e) != cudaSuccess) rc = fail(e, TRGT_ERR_CUDA, "align upload failed");
  if (rc != 0) {
    trgt_align_free(nullptr, b);
    return rc;
  }
  *out = b;
  return 0;
}

static int align_run_locked(trgt_engine_t *e, trgt_align_batch *b) {
  CU(e, cudaSetDevice(e->device));
  b->ran = true;
  b->total_words = 0;
  if (b->n_seqs == 0) return 0;
  const WfaSrc src = align_src(b);
  const uint32_t n = b->n_seqs;
  Counters *ctr = (Counters *)b->ctr.p;
  CU(e, cudaMemsetAsync(ctr, 0, sizeof(Counters), e->stream));
  CU(e, cudaMemsetAsync(b->cig_n.p, 0, ((size_t)n + 1) * sizeof(uint32_t), e->stream));
  CU(e, cudaMemsetAsync(b->status.p, 0, ((size_t)n + 1) * sizeof(int32_t), e->stream));
  const size_t bound = ring_ints_bound(src.x, src.oe, src.e, b->Pmax, b->Tmax);
  unsigned long long pool_cap1 = 0;
  if (false) {
    // (a CTA per pair: the wide-and-shallow shape of the flank problem.  End-to-end wavefronts of similar alleles stay
    // within a few hundred diagonals, which the banded on-chip ring of the warp kernel covers: measured on config 5,
    // 2 x 422 ms with a warp per pair and that ring against 2 x 501 ms with a CTA per pair on the global ring.)
    const int block = 128;
    const size_t cap_ints = 24 * 1024;
    const int smem_ring_ints = (int)(bound < cap_ints ? bound : cap_ints);
    const size_t smem = (size_t)(40 + smem_ring_ints) * sizeof(int);
    int grid = 0;
    TRY(persistent_grid(e, k_wfa_score<true>, block, smem, &grid));
    if ((uint32_t)grid > n) grid = (int)n;
    int *gring = nullptr;
    size_t stride = 0;
    if (bound > (size_t)smem_ring_ints) {
      stride = bound;
      TRY(cap_grid_to_budget(e, &grid, 1, stride, "align ring"));
      TRY(dev_reserve(e, b->gring, (size_t)grid * stride * sizeof(int)));
      gring = (int *)b->gring.p;
    }
    LaunchScope ls(e, "k_wfa_score_block");
    k_wfa_score<true><<<grid, block, smem, e->stream>>>(src, nullptr, nullptr, n, (WfaEnd *)b->ends.p, gring, stride,
                                                         smem_ring_ints, (uint32_t *)b->trace_work.p,
                                                         (uint32_t *)b->cig_n.p, ctr);
    TRY(check_launch(e, "k_wfa_score_block"));
  } else {
    const int block = 128, wpb = 4;
#ifndef TRGT_E2E_RING_INTS
#define TRGT_E2E_RING_INTS 1536  // 6 KB per warp: the rings of long pairs live in L2 anyway, and resident warps are what hides their latency
                                 // (config 5, k_wfa_score_warp: 301 ms at 24 KB, 180 at 12, 157 at 8, 144 at 6)
#endif
    const size_t cap_ints = TRGT_E2E_RING_INTS;
    size_t want = bound < cap_ints ? bound : cap_ints;
    if (want < (size_t)(E2E_NARROW_INTS + E2E_NARROW_WORDS)) want = E2E_NARROW_INTS + E2E_NARROW_WORDS;
    const int smem_ring_ints = (int)want;
    const size_t smem = (size_t)wpb * smem_ring_ints * sizeof(int);
    int grid = 0;
    TRY(persistent_grid(e, k_wfa_score<false>, block, smem, &grid));
    // the kernel walks the list of members the lane-per-pair kernels handed on, one member per warp and round
    // (a handful of 20 kb pairs must not queue 32 deep behind four warps)
    const uint32_t need = (n + wpb - 1) / wpb;
    if ((uint32_t)grid > need) grid = (int)need;
    int *gring = nullptr;
    size_t stride = 0;
    if (bound > (size_t)smem_ring_ints) {
      stride = bound;
      TRY(cap_grid_to_budget(e, &grid, wpb, stride, "align ring"));
      TRY(dev_reserve(e, b->gring, (size_t)grid * wpb * stride * sizeof(int)));
      gring = (int *)b->gring.p;
    }
    // CIGAR pool of the one-pass path (pairs it cannot place fall through to the two-pass path)
    pool_cap1 = 2ull * n + 65536ull;
    TRY(dev_reserve(e, b->pool, (size_t)(pool_cap1 + 1) * sizeof(uint32_t)));
    // one lane per pair first: identical members and cheap alignments end there
    TRY(dev_reserve(e, b->resid, ((size_t)n + 1) * sizeof(uint32_t)));
    TRY(dev_reserve(e, b->diff, ((size_t)n + 1) * sizeof(uint32_t)));
    {
      int igrid = 0;
      TRY(persistent_grid(e, k_e2e_identity, 256, 0, &igrid));
      const uint32_t ineed = (n + 255) / 256;
      if ((uint32_t)igrid > ineed) igrid = (int)ineed;
      LaunchScope ls(e, "k_e2e_identity");
      k_e2e_identity<<<igrid, 256, 0, e->stream>>>(src, n, (WfaEnd *)b->ends.p, (uint32_t *)b->cig_n.p,
                                                   (uint32_t *)b->diff.p, ctr);
      TRY(check_launch(e, "k_e2e_identity"));
    }
    {
      // the members that differ, one lane each, history rows (the scores the scoring allows up to the lane cap) on chip
      const int rows = __builtin_popcount(ft1_live_scores(src.x, src.oe, src.e, E2T_COST, nullptr));
      const size_t lsmem = (size_t)E2L_THREADS * rows * 3 * E2L_WMAX * sizeof(int16_t);
      int tgrid = 0;
      TRY(persistent_grid(e, k_e2e_lane, E2L_THREADS, lsmem, &tgrid));
      const uint32_t tneed = (n + E2L_THREADS - 1) / E2L_THREADS;
      if ((uint32_t)tgrid > tneed) tgrid = (int)tneed;
      LaunchScope ls(e, "k_e2e_lane");
      k_e2e_lane<<<tgrid, E2L_THREADS, lsmem, e->stream>>>(src, (const uint32_t *)b->diff.p, &ctr->n_diff, (WfaEnd *)b->ends.p,
                                                   (uint32_t *)b->cig_n.p, (unsigned long long *)b->cig_off.p,
                                                   (uint32_t *)b->pool.p, pool_cap1, (uint32_t *)b->resid.p, ctr, rows);
      TRY(check_launch(e, "k_e2e_lane"));
    }
    LaunchScope ls(e, "k_wfa_score_warp");
    k_wfa_score<false><<<grid, block, smem, e->stream>>>(src, (const uint32_t *)b->resid.p, &ctr->n_resid, 0, (WfaEnd *)b->ends.p, gring, stride,
                                                          smem_ring_ints, (uint32_t *)b->trace_work.p,
                                                          (uint32_t *)b->cig_n.p, ctr, (uint32_t *)b->pool.p, pool_cap1,
                                                          (unsigned long long *)b->cig_off.p);
    TRY(check_launch(e, "k_wfa_score_warp"));
  }
  CU(e, cudaMemcpyAsync(e->h_ctr, ctr, sizeof(Counters), cudaMemcpyDeviceToHost, e->stream));
  CU(e, engine_wait(e));
  // pool = words the one-pass path already placed + room for everything the trace pass may emit
  const unsigned long long words_bound = e->h_ctr->pool_used + e->h_ctr->words_bound;
  TRY(dev_reserve(e, b->pool, (size_t)(words_bound + 1) * sizeof(uint32_t), /*keep=*/true));
  TRY(launch_trace(e, src, (const uint32_t *)b->trace_work.p, &ctr->n_trace, e->h_ctr->n_trace, (const WfaEnd *)b->ends.p,
                   e->h_ctr->max_trace_ints, b->gws, 0.0, nullptr, (uint32_t *)b->pool.p, words_bound,
                   (unsigned long long *)b->cig_off.p, (uint32_t *)b->cig_n.p, (int32_t *)b->status.p, ctr));
  TRY(exclusive_scan_u32(e, (const uint32_t *)b->cig_n.p, (unsigned long long *)b->out_off.p, (size_t)n + 1));
  TRY(dev_reserve(e, b->out_words, (size_t)(words_bound + n + 1) * sizeof(uint32_t)));
  {
    LaunchScope ls(e, "k_cigar_gather");
    k_cigar_gather<<<(n + 255) / 256, 256, 0, e->stream>>>(src, n, (const WfaEnd *)b->ends.p, (const uint32_t *)b->pool.p,
                                                           (const unsigned long long *)b->cig_off.p,
                                                           (const uint32_t *)b->cig_n.p,
                                                           (const unsigned long long *)b->out_off.p,
                                                           (uint32_t *)b->out_words.p, (int32_t *)b->scores.p,
                                                           (int32_t *)b->status.p);
    TRY(check_launch(e, "k_cigar_gather"));
  }
  return 0;
}

int32_t trgt_align_run(trgt_engine_t *e, trgt_align_batch_t *b) {
  if (!e || !b) return TRGT_ERR_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  return align_run_locked(e, b);
}

static int align_download_locked(trgt_engine_t *e, trgt_align_batch *b, trgt_cigars_t *out) {
  if (!out) return fail(e, TRGT_ERR_ARG, "out is null");
  if (!b->ran) return fail(e, TRGT_ERR_ARG, "trgt_align_download before trgt_align_run");
  CU(e, cudaSetDevice(e->device));
  const size_t n = b->n_seqs;
  TRY(pin_reserve(e, b->h_off, (n + 1) * sizeof(uint64_t)));
  TRY(pin_reserve(e, b->h_scores, (n + 1) * sizeof(int32_t)));
  TRY(pin_reserve(e, b->h_status, (n + 1) * sizeof(int32_t)));
  b->h_off.as<uint64_t>()[0] = 0;
  b->total_words = 0;
  if (n) {
    CU(e, cudaMemcpyAsync(b->h_off.p, b->out_off.p, (n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaMemcpyAsync(b->h_scores.p, b->scores.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaMemcpyAsync(b->h_status.p, b->status.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    CU(e, engine_wait(e));
    b->total_words = b->h_off.as<uint64_t>()[n];
  }
  TRY(pin_reserve(e, b->h_words, (size_t)(b->total_words + 1) * sizeof(uint32_t)));
  if (b->total_words) {
    CU(e, cudaMemcpyAsync(b->h_words.p, b->out_words.p, (size_t)b->total_words * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    CU(e, engine_wait(e));
  }
  out->n = n;
  out->offsets = b->h_off.as<uint64_t>();
  out->words = b->h_words.as<uint32_t>();
  out->scores = b->h_scores.as<int32_t>();
  out->status = b->h_status.as<int32_t>();
  return 0;
}

int32_t trgt_align_download(trgt_engine_t *e, trgt_align_batch_t *b, trgt_cigars_t *out) {
  if (!e || !b) return TRGT_ERR_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  return align_download_locked(e, b, out);
}

int32_t trgt_align_e2e(trgt_engine_t *e, const trgt_seqs_t *backbones, const trgt_seqs_t *seqs,
                       const uint32_t *group_seq_offsets, uint32_t n_groups, trgt_cigars_t *out) {
  if (!e) return TRGT_ERR_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  if (!e->one_align) e->one_align = new trgt_align_batch();
  static const bool trace = getenv("TRGT_TRACE") != nullptr;  // host-side stage times of this call, to stderr
  const auto t0 = std::chrono::steady_clock::now();
  TRY(align_upload_into(e, e->one_align, backbones, seqs, group_seq_offsets, n_groups));
  const auto t1 = std::chrono::steady_clock::now();
  TRY(align_run_locked(e, e->one_align));
  const auto t2 = std::chrono::steady_clock::now();
  const int rc = align_download_locked(e, e->one_align, out);
  if (trace) {
    const auto t3 = std::chrono::steady_clock::now();
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
      return std::chrono::duration<double, std::milli>(b - a).count();
    };
    fprintf(stderr, "[trgt] align_e2e: checks + uploads queued %.3f ms, kernels (two read-backs) %.3f ms, download %.3f ms\n",
            ms(t0, t1), ms(t1, t2), ms(t2, t3));
  }
  return rc;
}

// ------------------------------------------------------------------ consensus (next row) ------

static int consensus_run_locked(trgt_engine_t *e, trgt_align_batch *b, trgt_seqs_out_t *out);

int32_t trgt_consensus(trgt_engine_t *e, const trgt_seqs_t *backbones, const trgt_seqs_t *seqs,
                       const uint32_t *group_seq_offsets, uint32_t n_groups, trgt_seqs_out_t *out) {
  if (!e || !out) return TRGT_ERR_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  if (!e->one_align) e->one_align = new trgt_align_batch();
  trgt_align_batch *b = e->one_align;
  TRY(align_upload_into(e, b, backbones, seqs, group_seq_offsets, n_groups));
  return consensus_run_locked(e, b, out);
}

// utils::align + repair_consensus on an align batch whose sequences are in place
static int consensus_run_locked(trgt_engine_t *e, trgt_align_batch *b, trgt_seqs_out_t *out) {
  const uint32_t n_groups = b->n_groups;
  TRY(align_run_locked(e, b));  // CIGARs now sit in out_off / out_words on the device
  const size_t ng = n_groups;
  TRY(pin_reserve(e, b->h_cons_off, (ng + 1) * sizeof(uint64_t)));
  TRY(pin_reserve(e, b->h_cons_status, (ng + 1) * sizeof(int32_t)));
  b->h_cons_off.as<uint64_t>()[0] = 0;
  out->n = ng;
  out->offsets = b->h_cons_off.as<uint64_t>();
  out->status = b->h_cons_status.as<int32_t>();
  out->data = nullptr;
  if (ng == 0) return 0;
  const WfaSrc src = align_src(b);
  TRY(dev_reserve(e, b->cons_len, (ng + 2) * sizeof(uint32_t)));
  TRY(dev_reserve(e, b->cons_status, (ng + 1) * sizeof(int32_t)));
  TRY(dev_reserve(e, b->cons_off, (ng + 2) * sizeof(unsigned long long)));
  // total CIGAR words bound the insertion records (one slot per word, plus one per group)
  CU(e, cudaMemcpyAsync(&e->h_u64[2], (unsigned long long *)b->out_off.p + b->n_seqs, sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
  CU(e, engine_wait(e));
  const unsigned long long total_words = b->n_seqs ? e->h_u64[2] : 0;
  TRY(dev_reserve(e, b->cons_recs, (size_t)(total_words + ng + 1) * sizeof(ConsRec)));
  // long backbones: a CTA per group (the members' CIGAR walks and the columns over four warps); else a warp per group
  const bool per_cta = b->Pmax > 2048;
  const int block = 128, wpb = per_cta ? 1 : 4;
  int grid = 0;
  if (per_cta) TRY(persistent_grid(e, k_consensus_vote<false, true>, block, 0, &grid));
  else TRY(persistent_grid(e, k_consensus_vote<false, false>, block, 0, &grid));
  const uint32_t need = (uint32_t)((ng + wpb - 1) / wpb);
  if ((uint32_t)grid > need) grid = (int)need;
  const size_t stride = (size_t)6 * (size_t)(b->Pmax > 0 ? b->Pmax : 1);
  // fewer resident warps when one column-count slot is large
  {
    const size_t budget_ints = e->workspace_budget / sizeof(int);
    size_t slots = (size_t)grid * wpb;
    if (stride > budget_ints) return fail(e, TRGT_ERR_INTERNAL, "consensus workspace exceeds the budget");
    if (slots * stride > budget_ints) {
      slots = budget_ints / stride;
      grid = (int)(slots / wpb);
      if (grid < 1) grid = 1;
    }
  }
  TRY(dev_reserve(e, b->cons_counts, (size_t)grid * wpb * stride * sizeof(int)));
  {
    LaunchScope ls(e, "k_consensus_vote_count");
    if (per_cta)
      k_consensus_vote<false, true><<<grid, block, 0, e->stream>>>(
          src, (const uint32_t *)b->group_off.p, n_groups, (const uint32_t *)b->out_words.p,
          (const unsigned long long *)b->out_off.p, (const int32_t *)b->status.p, (int *)b->cons_counts.p, stride,
          (ConsRec *)b->cons_recs.p, (uint32_t *)b->cons_len.p, (int32_t *)b->cons_status.p, nullptr, nullptr);
    else
      k_consensus_vote<false, false><<<grid, block, 0, e->stream>>>(
          src, (const uint32_t *)b->group_off.p, n_groups, (const uint32_t *)b->out_words.p,
          (const unsigned long long *)b->out_off.p, (const int32_t *)b->status.p, (int *)b->cons_counts.p, stride,
          (ConsRec *)b->cons_recs.p, (uint32_t *)b->cons_len.p, (int32_t *)b->cons_status.p, nullptr, nullptr);
    TRY(check_launch(e, "k_consensus_vote_count"));
  }
  CU(e, cudaMemsetAsync((uint32_t *)b->cons_len.p + ng, 0, sizeof(uint32_t), e->stream));
  TRY(exclusive_scan_u32(e, (const uint32_t *)b->cons_len.p, (unsigned long long *)b->cons_off.p, ng + 1));
  CU(e, cudaMemcpyAsync(b->h_cons_off.p, b->cons_off.p, (ng + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
  CU(e, cudaMemcpyAsync(b->h_cons_status.p, b->cons_status.p, ng * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
  CU(e, engine_wait(e));
  const unsigned long long total = b->h_cons_off.as<uint64_t>()[ng];
  TRY(dev_reserve(e, b->cons_data, (size_t)total + 16));
  TRY(pin_reserve(e, b->h_cons_data, (size_t)total + 16));
  if (total) {
    LaunchScope ls(e, "k_consensus_vote_write");
    if (per_cta)
      k_consensus_vote<true, true><<<grid, block, 0, e->stream>>>(
          src, (const uint32_t *)b->group_off.p, n_groups, (const uint32_t *)b->out_words.p,
          (const unsigned long long *)b->out_off.p, (const int32_t *)b->status.p, (int *)b->cons_counts.p, stride,
          (ConsRec *)b->cons_recs.p, (uint32_t *)b->cons_len.p, (int32_t *)b->cons_status.p,
          (const unsigned long long *)b->cons_off.p, (uint8_t *)b->cons_data.p);
    else
      k_consensus_vote<true, false><<<grid, block, 0, e->stream>>>(
          src, (const uint32_t *)b->group_off.p, n_groups, (const uint32_t *)b->out_words.p,
          (const unsigned long long *)b->out_off.p, (const int32_t *)b->status.p, (int *)b->cons_counts.p, stride,
          (ConsRec *)b->cons_recs.p, (uint32_t *)b->cons_len.p, (int32_t *)b->cons_status.p,
          (const unsigned long long *)b->cons_off.p, (uint8_t *)b->cons_data.p);
    TRY(check_launch(e, "k_consensus_vote_write"));
    CU(e, cudaMemcpyAsync(b->h_cons_data.p, b->cons_data.p, (size_t)total, cudaMemcpyDeviceToHost, e->stream));
    CU(e, engine_wait(e));
  }
  out->data = b->h_cons_data.as<uint8_t>();
  return 0;
}

// ------------------------------------------------------------------ phase B: edit distance ---

// get_dist_matrix of every locus into d_ed[4] (condensed, locus after locus at pair_off).  Sequences: d_seqs /
// d_seq_off as a CSR set, or -- with d_index and d_spans -- the repeat sequences of a flank batch read in place.
static int edit_dist_device(trgt_engine_t *e, const uint8_t *d_seqs, const uint64_t *d_seq_off, const uint32_t *d_index,
                            const trgt_span_t *d_spans, const uint32_t *locus_seq_offsets, uint32_t n_loci, uint64_t n_seqs,
                            unsigned long long *total_out, uint32_t *n_max_out) {
  if (locus_seq_offsets[0] != 0 || locus_seq_offsets[n_loci] != n_seqs) return fail(e, TRGT_ERR_ARG, "locus_seq_offsets must cover all sequences");
  std::vector<unsigned long long> pair_off((size_t)n_loci + 1, 0);
  uint32_t n_max = 0;
  for (uint32_t l = 0; l < n_loci; l++) {
    if (locus_seq_offsets[l + 1] < locus_seq_offsets[l]) return fail(e, TRGT_ERR_ARG, "locus_seq_offsets not monotone");
    const unsigned long long n = locus_seq_offsets[l + 1] - locus_seq_offsets[l];
    pair_off[l + 1] = pair_off[l] + (n < 2 ? 0 : n * (n - 1) / 2);
    if (n > n_max) n_max = (uint32_t)n;
  }
  const unsigned long long total = pair_off[n_loci];
  *total_out = total;
  if (n_max_out) *n_max_out = n_max;
  TRY(h2d(e, e->d_ed[2], locus_seq_offsets, ((size_t)n_loci + 1) * sizeof(uint32_t)));
  TRY(h2d(e, e->d_ed[3], pair_off.data(), pair_off.size() * sizeof(unsigned long long)));
  TRY(dev_reserve(e, e->d_ed[4], (size_t)(total + 1) * sizeof(double)));
  CU(e, engine_wait(e));  // pair_off is a local
  if (total == 0) return 0;
  int grid = 0;
  TRY(persistent_grid(e, k_edit_dist, 128, 0, &grid));
  if ((uint32_t)grid > n_loci) grid = (int)n_loci;
  LaunchScope ls(e, "k_edit_dist");
  k_edit_dist<<<grid, 128, 0, e->stream>>>(d_seqs, d_seq_off, (const uint32_t *)e->d_ed[2].p,
                                           (const unsigned long long *)e->d_ed[3].p, n_loci, (double *)e->d_ed[4].p,
                                           d_index, d_spans);
  return check_launch(e, "k_edit_dist");
}

int32_t trgt_edit_dist(trgt_engine_t *e, const trgt_seqs_t *seqs, const uint32_t *locus_seq_offsets, uint32_t n_loci,
                       double *dists_out) {
  if (!e) return TRGT_ERR_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  TRY(check_seqs(e, seqs, "seqs"));
  if (n_loci && !locus_seq_offsets) return fail(e, TRGT_ERR_ARG, "locus_seq_offsets is null");
  if (n_loci == 0) return 0;
  CU(e, cudaSetDevice(e->device));
  TRY(upload_seqs(e, seqs, e->d_ed[0], e->d_ed[1]));
  unsigned long long total = 0;
  TRY(edit_dist_device(e, (const uint8_t *)e->d_ed[0].p, (const uint64_t *)e->d_ed[1].p, nullptr, nullptr,
                       locus_seq_offsets, n_loci, seqs->n, &total, nullptr));
  if (total == 0) return 0;
  if (!dists_out) return fail(e, TRGT_ERR_ARG, "dists_out is null");
  CU(e, cudaMemcpyAsync(dists_out, e->d_ed[4].p, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  CU(e, engine_wait(e));
  return 0;
}

// ------------------------------------------------------------------ cluster-genotyper glue (next row, rank 3) ---

// cluster() + group1 / group2 + central_read on the matrices edit_dist_device left in d_ed[4]
static int cluster_device(trgt_engine_t *e, uint32_t n_loci, uint64_t n_seqs, uint32_t n_max, int32_t *group_out,
                          uint32_t *central_out, uint32_t *n_groups_out) {
  const int block = 128, wpb = 4;
  int grid = 0;
  TRY(persistent_grid(e, k_cluster_ward, block, 0, &grid));
  const uint32_t need = (n_loci + wpb - 1) / wpb;
  if ((uint32_t)grid > need) grid = (int)need;
  const size_t stride = (cl_ws_bytes(n_max) + 15) & ~(size_t)15;
  TRY(dev_reserve(e, e->d_cl[0], (size_t)grid * wpb * stride));
  TRY(dev_reserve(e, e->d_cl[1], (size_t)(n_seqs + 1) * sizeof(int32_t)));
  TRY(dev_reserve(e, e->d_cl[2], ((size_t)n_loci * 2 + 2) * sizeof(uint32_t)));
  TRY(dev_reserve(e, e->d_cl[3], ((size_t)n_loci + 1) * sizeof(uint32_t)));
  {
    LaunchScope ls(e, "k_cluster_ward");
    k_cluster_ward<<<grid, block, 0, e->stream>>>((const uint32_t *)e->d_ed[2].p, (const unsigned long long *)e->d_ed[3].p,
                                                  n_loci, (double *)e->d_ed[4].p, (unsigned char *)e->d_cl[0].p, stride,
                                                  (int32_t *)e->d_cl[1].p, (uint32_t *)e->d_cl[2].p, (uint32_t *)e->d_cl[3].p);
    TRY(check_launch(e, "k_cluster_ward"));
  }
  if (n_seqs) CU(e, cudaMemcpyAsync(group_out, e->d_cl[1].p, (size_t)n_seqs * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
  CU(e, cudaMemcpyAsync(central_out, e->d_cl[2].p, (size_t)n_loci * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
  if (n_groups_out) CU(e, cudaMemcpyAsync(n_groups_out, e->d_cl[3].p, (size_t)n_loci * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
  CU(e, engine_wait(e));
  return 0;
}

int32_t trgt_cluster(trgt_engine_t *e, const trgt_seqs_t *seqs, const uint32_t *locus_seq_offsets, uint32_t n_loci,
                     int32_t *group_out, uint32_t *central_out, uint32_t *n_groups_out) {
  if (!e) return TRGT_ERR_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  TRY(check_seqs(e, seqs, "seqs"));
  if (n_loci && !locus_seq_offsets) return fail(e, TRGT_ERR_ARG, "locus_seq_offsets is null");
  if (n_loci == 0) return 0;
  if (!central_out || (seqs->n && !group_out)) return fail(e, TRGT_ERR_ARG, "trgt_cluster: null output");
  CU(e, cudaSetDevice(e->device));
  TRY(upload_seqs(e, seqs, e->d_ed[0], e->d_ed[1]));
  unsigned long long total = 0;
  uint32_t n_max = 0;
  TRY(edit_dist_device(e, (const uint8_t *)e->d_ed[0].p, (const uint64_t *)e->d_ed[1].p, nullptr, nullptr,
                       locus_seq_offsets, n_loci, seqs->n, &total, &n_max));
  return cluster_device(e, n_loci, seqs->n, n_max, group_out, central_out, n_groups_out);
}

// read indices of a flank batch handed in by the caller: all in range
static int check_read_index(trgt_engine_t *e, const trgt_flank_batch *b, const uint32_t *reads, uint64_t n, const char *what) {
  if (n && !reads) return fail(e, TRGT_ERR_ARG, "%s: null read index", what);
  for (uint64_t i = 0; i < n; i++)
    if (reads[i] >= b->n_reads) return fail(e, TRGT_ERR_ARG, "%s: read %u out of range", what, reads[i]);
  return 0;
}

int32_t trgt_cluster_trs(trgt_engine_t *e, trgt_flank_batch_t *b, const uint32_t *reads, const uint32_t *locus_offsets,
                         uint32_t n_loci, int32_t *group_out, uint32_t *central_out, uint32_t *n_groups_out) {
  if (!e) return TRGT_ERR_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  if (!b) b = e->one_flank;
  if (!b || (b->n_reads && !b->ran)) return fail(e, TRGT_ERR_ARG, "trgt_cluster_trs: the flank batch has not been run");
  if (n_loci == 0) return 0;
  if (!locus_offsets || !central_out) return fail(e, TRGT_ERR_ARG, "trgt_cluster_trs: null argument");
  const uint64_t n_sel = locus_offsets[n_loci];
  if (n_sel && !group_out) return fail(e, TRGT_ERR_ARG, "trgt_cluster_trs: null output");
  TRY(check_read_index(e, b, reads, n_sel, "trgt_cluster_trs"));
  CU(e, cudaSetDevice(e->device));
  static const uint32_t zero32[1] = {0};
  TRY(h2d(e, e->d_cl[4], n_sel ? reads : zero32, (size_t)(n_sel ? n_sel : 1) * sizeof(uint32_t)));
  unsigned long long total = 0;
  uint32_t n_max = 0;
  TRY(edit_dist_device(e, (const uint8_t *)b->reads.p, (const uint64_t *)b->read_off.p, (const uint32_t *)e->d_cl[4].p,
                       (const trgt_span_t *)b->spans.p, locus_offsets, n_loci, n_sel, &total, &n_max));
  return cluster_device(e, n_loci, n_sel, n_max, group_out, central_out, n_groups_out);
}

// Backbones and members named by read index of a flank batch -> the engine's one-shot align batch, sequences gathered
// on the device from the batch's reads and spans (index, lengths, scan, bytes) and everything else prepared.
static int align_gather_trs(trgt_engine_t *e, trgt_flank_batch_t *fb, const uint32_t *backbone_reads,
                            const uint32_t *member_reads, const uint32_t *group_offsets, uint32_t n_groups,
                            const char *what, trgt_align_batch **b_out) {
  if (!fb) fb = e->one_flank;
  if (!fb || (fb->n_reads && !fb->ran)) return fail(e, TRGT_ERR_ARG, "%s: the flank batch has not been run", what);
  if (n_groups && (!group_offsets || !backbone_reads)) return fail(e, TRGT_ERR_ARG, "%s: null argument", what);
  const uint32_t n_seqs = n_groups ? group_offsets[n_groups] : 0;
  if (n_groups && group_offsets[0] != 0) return fail(e, TRGT_ERR_ARG, "group_offsets must start at 0");
  for (uint32_t g = 0; g < n_groups; g++)
    if (group_offsets[g + 1] < group_offsets[g]) return fail(e, TRGT_ERR_ARG, "group_offsets not monotone");
  TRY(check_read_index(e, fb, backbone_reads, n_groups, what));
  TRY(check_read_index(e, fb, member_reads, n_seqs, what));
  CU(e, cudaSetDevice(e->device));
  if (!e->one_align) e->one_align = new trgt_align_batch();
  trgt_align_batch *b = e->one_align;
  // Lengths, offsets (a scan) and bytes of both sets are produced on the device from the batch's spans; only the two
  // index lists go up, and the two totals and the two longest lengths (they size buffers and wavefront rings) come back.
  struct Part { const uint32_t *idx; uint32_t n; DevBuf *data, *off, *d_idx, *d_len; } parts[2] = {
      {backbone_reads, n_groups, &b->bb, &b->bb_off, &e->d_cl[4], &e->d_ed[1]},
      {member_reads, n_seqs, &b->seqs, &b->seq_off, &e->d_ed[0], &e->d_ed[2]}};
  unsigned long long *h = e->h_u64;  // pinned: [0..1] totals, [2..3] longest lengths (as 32-bit words)
  TRY(dev_reserve(e, e->d_ed[3], 4 * sizeof(unsigned int)));
  CU(e, cudaMemsetAsync(e->d_ed[3].p, 0, 4 * sizeof(unsigned int), e->stream));
  static const uint32_t zero32[1] = {0};
  for (int k = 0; k < 2; k++) {
    Part &p = parts[k];
    TRY(h2d(e, *p.d_idx, p.n ? p.idx : zero32, (size_t)(p.n ? p.n : 1) * sizeof(uint32_t)));
    TRY(dev_reserve(e, *p.d_len, ((size_t)p.n + 1) * sizeof(uint32_t)));
    TRY(dev_reserve(e, *p.off, ((size_t)p.n + 1) * sizeof(unsigned long long) + 16));
    {
      LaunchScope ls(e, "k_trs_len");
      k_trs_len<<<(p.n + 256) / 256, 256, 0, e->stream>>>((const trgt_span_t *)fb->spans.p, (const uint32_t *)p.d_idx->p, p.n,
                                                         (uint32_t *)p.d_len->p, (unsigned int *)e->d_ed[3].p + k);
      TRY(check_launch(e, "k_trs_len"));
    }
    TRY(exclusive_scan_u32(e, (const uint32_t *)p.d_len->p, (unsigned long long *)p.off->p, (size_t)p.n + 1));
    CU(e, cudaMemcpyAsync(&h[k], (const unsigned long long *)p.off->p + p.n, sizeof(unsigned long long),
                          cudaMemcpyDeviceToHost, e->stream));
  }
  CU(e, cudaMemcpyAsync(&h[2], e->d_ed[3].p, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, e->stream));
  CU(e, engine_wait(e));
  const unsigned int *h_max = (const unsigned int *)&h[2];
  const int pm = (int)h_max[0], tm = (int)h_max[1];
  for (int k = 0; k < 2; k++) {
    Part &p = parts[k];
    const unsigned long long total = h[k];
    TRY(dev_reserve(e, *p.data, (size_t)total + 16));
    if (total) {
      int grid = 0;
      TRY(persistent_grid(e, k_trs_gather, 256, 0, &grid));
      const uint32_t need = (p.n + 7) / 8;
      if ((uint32_t)grid > need) grid = (int)need;
      LaunchScope ls(e, "k_trs_gather");
      k_trs_gather<<<grid, 256, 0, e->stream>>>((const uint8_t *)fb->reads.p, (const uint64_t *)fb->read_off.p,
                                                (const trgt_span_t *)fb->spans.p, (const uint32_t *)p.d_idx->p,
                                                (const unsigned long long *)p.off->p, p.n, (uint8_t *)p.data->p);
      TRY(check_launch(e, "k_trs_gather"));
    }
  }
  if ((uint64_t)pm + (uint64_t)tm > 0x0fffffffull) return fail(e, TRGT_ERR_ARG, "sequence too long");
  TRY(align_prepare(e, b, group_offsets, n_groups, n_seqs, pm, tm));
  *b_out = b;
  return 0;
}

int32_t trgt_consensus_trs(trgt_engine_t *e, trgt_flank_batch_t *fb, const uint32_t *backbone_reads,
                           const uint32_t *member_reads, const uint32_t *group_offsets, uint32_t n_groups,
                           trgt_seqs_out_t *out) {
  if (!e || !out) return TRGT_ERR_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  trgt_align_batch *b = nullptr;
  TRY(align_gather_trs(e, fb, backbone_reads, member_reads, group_offsets, n_groups, "trgt_consensus_trs", &b));
  return consensus_run_locked(e, b, out);
}

int32_t trgt_align_trs(trgt_engine_t *e, trgt_flank_batch_t *fb, const uint32_t *backbone_reads,
                       const uint32_t *member_reads, const uint32_t *group_offsets, uint32_t n_groups,
                       trgt_cigars_t *out) {
  if (!e || !out) return TRGT_ERR_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  trgt_align_batch *b = nullptr;
  TRY(align_gather_trs(e, fb, backbone_reads, member_reads, group_offsets, n_groups, "trgt_align_trs", &b));
  TRY(align_run_locked(e, b));
  return align_download_locked(e, b, out);
}

}  // extern "C"

// ------------------------------------------------------------------ phase C: HMM -----------

struct trgt_hmm_batch {
  uint32_t n_loci = 0, n_alleles = 0;
  int want_paths = 0;
  int S_max = 0, nb_max = 0, mbytes_max = 0;
  size_t warp_bytes = 0;
  DevBuf motifs, motif_off, locus_motif_off, alleles, allele_off, allele_locus, bp_off, mc_off;
  DevBuf bp, mc, purity, n_spans, span_off, spans, path_len, path_off, paths, status;
  std::vector<unsigned long long> h_bp_off, h_mc_off, h_scr_off;
  DevBuf span_scratch;
  // the alleles the generic kernels take (those the lane kernels do not): h_bp_off / h_scr_off index this list
  std::vector<uint32_t> h_list;
  DevBuf list, scr_off;
  // single-motif loci: slots sorted by (motif length, allele length), 32 slots = one group = one warp
  std::vector<uint32_t> h_slots;
  std::vector<unsigned long long> h_group_off;   // [n_groups+1] words
  std::vector<uint8_t> h_group_n;                // motif length of each group
  std::vector<std::pair<uint32_t, uint32_t>> lane_waves;  // group ranges whose packed words fit the budget
  DevBuf slots, group_off, group_n, lane_bp, lane_scratch;
  std::vector<uint32_t> h_locus_allele_off;  // [n_loci+1] when the alleles are grouped by locus, else empty
  DevBuf locus_allele_off, vcf_len, vcf_off, vcf_data;
  PinBuf r_vcf_off, r_vcf_data;
  std::vector<std::pair<uint32_t, uint32_t>> waves;
  // host results
  PinBuf r_mc, r_span_off, r_spans, r_purity, r_status, r_path_off, r_paths;
  unsigned long long total_spans = 0, total_path = 0;
  bool ran = false;
};

extern "C" {

void trgt_hmm_free(trgt_engine_t *e, trgt_hmm_batch_t *b) {
  if (!b) return;
  if (e) {
    cudaSetDevice(e->device);
    engine_wait(e);
    if (e->one_hmm == b) e->one_hmm = nullptr;
  }
  DevBuf *all[] = {&b->motifs, &b->motif_off, &b->locus_motif_off, &b->alleles, &b->allele_off, &b->allele_locus,
                   &b->bp_off, &b->mc_off, &b->bp, &b->mc, &b->purity, &b->n_spans, &b->span_off, &b->spans,
                   &b->path_len, &b->path_off, &b->paths, &b->status, &b->locus_allele_off, &b->vcf_len, &b->vcf_off,
                   &b->vcf_data, &b->span_scratch, &b->list, &b->scr_off, &b->slots, &b->group_off, &b->group_n, &b->lane_bp,
                   &b->lane_scratch};
  for (auto *d : all) dev_free(*d);
  PinBuf *pins[] = {&b->r_mc, &b->r_span_off, &b->r_spans, &b->r_purity, &b->r_status, &b->r_path_off, &b->r_paths,
                    &b->r_vcf_off, &b->r_vcf_data};
  for (auto *q : pins) pin_free(*q);
  delete b;
}

static int hmm_upload_into(trgt_engine_t *e, trgt_hmm_batch *b, const trgt_seqs_t *motifs,
                           const uint32_t *locus_motif_offsets, uint32_t n_loci, const trgt_seqs_t *alleles,
                           const uint32_t *allele_locus, int32_t want_paths) {
  TRY(check_seqs(e, motifs, "motifs"));
  TRY(check_seqs(e, alleles, "alleles"));
  if (n_loci && !locus_motif_offsets) return fail(e, TRGT_ERR_ARG, "locus_motif_offsets is null");
  if (alleles->n && !allele_locus) return fail(e, TRGT_ERR_ARG, "allele_locus is null");
  if (alleles->n > 0x7fffffffull) return fail(e, TRGT_ERR_ARG, "too many alleles in one batch");
  if (n_loci && (locus_motif_offsets[0] != 0 || locus_motif_offsets[n_loci] != motifs->n))
    return fail(e, TRGT_ERR_ARG, "locus_motif_offsets must cover all motifs");
  // per-locus model size
  std::vector<uint32_t> locus_S(n_loci, 0), locus_mb(n_loci, 0);
  int S_max = 7, nb_max = 1, mb_max = 0;
  uint64_t len_max = 0;
  for (uint32_t l = 0; l < n_loci; l++) {
    if (locus_motif_offsets[l + 1] < locus_motif_offsets[l]) return fail(e, TRGT_ERR_ARG, "locus_motif_offsets not monotone");
    uint64_t S = 7, mb = 0;
    const uint32_t nm = locus_motif_offsets[l + 1] - locus_motif_offsets[l];
    for (uint32_t m = locus_motif_offsets[l]; m < locus_motif_offsets[l + 1]; m++) {
      const uint64_t n = motifs->offsets[m + 1] - motifs->offsets[m];
      S += 3 * n + 1;
      mb += n;
      if (n > len_max) len_max = n;
    }
    if (S > 65535 || nm + 1 > 255) return fail(e, TRGT_ERR_ARG, "locus %u: model too large (%llu states, %u motifs)", l, (unsigned long long)S, nm);
    locus_S[l] = (uint32_t)S;
    locus_mb[l] = (uint32_t)mb;
    if ((int)S > S_max) S_max = (int)S;
    if ((int)nm + 1 > nb_max) nb_max = (int)nm + 1;
    if ((int)mb > mb_max) mb_max = (int)mb;
  }
  const size_t n = alleles->n;
  b->h_mc_off.assign(n + 1, 0);
  // Two classes of alleles.  Single-motif loci with a motif of 1..HMM_LANE_NMAX bases (every locus of a genome-wide
  // catalog) take the lane kernels: counting sort by (motif length, allele length) into slots, 32 slots = one warp.
  // Everything else (several motifs, long motifs, empty alleles) is listed for the generic kernels.
  const uint32_t LB = 1024;  // alleles of LB-1 bases and more share the last bin of their motif length (sorted below)
  std::vector<uint32_t> bin_of(n, 0xFFFFFFFFu);
  std::vector<uint32_t> bin_cnt((size_t)(HMM_LANE_NMAX + 1) * LB + 1, 0);
  b->h_list.clear();
  for (size_t a = 0; a < n; a++) {
    const uint32_t l = allele_locus[a];
    if (l >= n_loci) return fail(e, TRGT_ERR_ARG, "allele %zu: locus %u out of range", a, l);
    const uint64_t L = alleles->offsets[a + 1] - alleles->offsets[a];
    if (L > 0x3ffffff0ull) return fail(e, TRGT_ERR_ARG, "allele too long");
    const uint32_t nm = locus_motif_offsets[l + 1] - locus_motif_offsets[l];
    b->h_mc_off[a + 1] = b->h_mc_off[a] + nm;
    const uint64_t n1 = nm == 1 ? motifs->offsets[locus_motif_offsets[l] + 1] - motifs->offsets[locus_motif_offsets[l]] : 0;
    if (e->hmm_lane && n1 >= 1 && n1 <= HMM_LANE_NMAX && L > 0 && L < (1ull << 26)) {  // (32-bit event counters)
      bin_of[a] = (uint32_t)n1 * LB + (uint32_t)(L < LB - 1 ? L : LB - 1);
      bin_cnt[bin_of[a] + 1]++;
    } else {
      b->h_list.push_back((uint32_t)a);
    }
  }
  const size_t n_gen = b->h_list.size(), n_lane = n - n_gen;
  b->h_slots.clear();
  b->h_group_off.assign(1, 0);
  b->h_group_n.clear();
  if (n_lane) {
    for (size_t k = 1; k < bin_cnt.size(); k++) bin_cnt[k] += bin_cnt[k - 1];  // bin_cnt[k] = first position of bin k
    std::vector<uint32_t> sorted(n_lane);
    {
      std::vector<uint32_t> pos(bin_cnt.begin(), bin_cnt.end() - 1);
      for (size_t a = 0; a < n; a++)
        if (bin_of[a] != 0xFFFFFFFFu) sorted[pos[bin_of[a]]++] = (uint32_t)a;
    }
    auto len_of = [&](uint32_t a) { return alleles->offsets[a + 1] - alleles->offsets[a]; };
    struct Group { uint64_t lmax; uint32_t first, count, n1; };
    std::vector<Group> groups;
    for (uint32_t n1 = 1; n1 <= HMM_LANE_NMAX; n1++) {
      const size_t s0 = bin_cnt[(size_t)n1 * LB], s1 = bin_cnt[(size_t)(n1 + 1) * LB];
      const size_t t0 = bin_cnt[(size_t)n1 * LB + LB - 1];  // the long ones of this motif length, by length
      std::sort(sorted.begin() + t0, sorted.begin() + s1, [&](uint32_t x, uint32_t y) {
        const uint64_t lx = len_of(x), ly = len_of(y);
        return lx != ly ? lx < ly : x < y;
      });
      for (size_t k = s0; k < s1; k += 32) {  // one group: its columns are as many as its longest allele has bases
        const uint32_t cnt = (uint32_t)(s1 - k < 32 ? s1 - k : 32);
        uint64_t lmax = 0;
        for (uint32_t j = 0; j < cnt; j++) lmax = len_of(sorted[k + j]) > lmax ? len_of(sorted[k + j]) : lmax;
        groups.push_back(Group{lmax, (uint32_t)k, cnt, n1});
      }
    }
    // longest first: a lane walks its columns one after the other, so the long alleles should start first
    std::stable_sort(groups.begin(), groups.end(), [](const Group &x, const Group &y) { return x.lmax > y.lmax; });
    b->h_slots.reserve(groups.size() * 32);
    for (const Group &g : groups) {
      for (uint32_t j = 0; j < 32; j++) b->h_slots.push_back(j < g.count ? sorted[g.first + j] : 0xFFFFFFFFu);
      b->h_group_off.push_back(b->h_group_off.back() + 32ull * g.lmax);
      b->h_group_n.push_back((uint8_t)g.n1);
    }
  }
  // back-pointer bytes and span scratch slots of the generic alleles, indexed by their position in the list
  b->h_bp_off.assign(n_gen + 1, 0);
  b->h_scr_off.assign(n_gen + 1, 0);
  for (size_t i = 0; i < n_gen; i++) {
    const uint32_t a = b->h_list[i];
    const uint64_t L = alleles->offsets[a + 1] - alleles->offsets[a];
    b->h_bp_off[i + 1] = b->h_bp_off[i] + (L ? (L + 2) * (unsigned long long)locus_S[allele_locus[a]] : 0);
    b->h_scr_off[i + 1] = b->h_scr_off[i] + L;
  }
  {  // allele range of every locus, if the alleles come grouped by locus (trgt_vcf_fields needs that)
    bool sorted = true;
    for (size_t a = 1; a < n && sorted; a++) sorted = allele_locus[a] >= allele_locus[a - 1];
    b->h_locus_allele_off.clear();
    if (sorted) {
      b->h_locus_allele_off.assign((size_t)n_loci + 1, 0);
      for (size_t a = 0; a < n; a++) b->h_locus_allele_off[allele_locus[a] + 1]++;
      for (uint32_t l = 0; l < n_loci; l++) b->h_locus_allele_off[l + 1] += b->h_locus_allele_off[l];
    }
  }
  b->n_loci = n_loci;
  b->n_alleles = (uint32_t)n;
  b->want_paths = want_paths;
  b->S_max = S_max;
  b->nb_max = nb_max;
  b->mbytes_max = mb_max;
  b->warp_bytes = hmm_onchip_bytes(S_max, nb_max, mb_max);
  b->ran = false;
  if (b->warp_bytes * 4 > (size_t)e->smem_optin) return fail(e, TRGT_ERR_ARG, "HMM with %d states does not fit in shared memory", S_max);
  // waves: consecutive listed alleles / groups whose back-pointers fit the workspace budget (state paths are
  // written by a second walk after all offsets are known, which needs every back-pointer: one wave then)
  b->waves.clear();
  b->lane_waves.clear();
  {
    size_t a0 = 0;
    while (a0 < n_gen) {
      size_t a1 = a0 + 1;
      while (a1 < n_gen && (want_paths || b->h_bp_off[a1 + 1] - b->h_bp_off[a0] <= e->workspace_budget)) a1++;
      b->waves.push_back({(uint32_t)a0, (uint32_t)a1});
      a0 = a1;
    }
    const size_t n_groups = b->h_group_off.size() - 1;
    size_t g0 = 0;
    while (g0 < n_groups) {
      size_t g1 = g0 + 1;
      while (g1 < n_groups && (want_paths || (b->h_group_off[g1 + 1] - b->h_group_off[g0]) * sizeof(uint32_t) <= e->workspace_budget)) g1++;
      b->lane_waves.push_back({(uint32_t)g0, (uint32_t)g1});
      g0 = g1;
    }
  }
  CU(e, cudaSetDevice(e->device));
  // jump-in ln table
  e->jump.ensure((int)(len_max > HMM_LANE_NMAX ? len_max : HMM_LANE_NMAX));
  if (e->jump_uploaded_len != e->jump.lp.size() || !e->d_mm_lp.p) {
    TRY(h2d(e, e->d_mm_off, e->jump.off.data(), e->jump.off.size() * sizeof(uint32_t)));
    TRY(h2d(e, e->d_mm_lp, e->jump.lp.data(), e->jump.lp.size() * sizeof(double)));
    CU(e, engine_wait(e));
    e->jump_uploaded_len = e->jump.lp.size();
  }
  TRY(upload_seqs(e, motifs, b->motifs, b->motif_off));
  TRY(upload_seqs(e, alleles, b->alleles, b->allele_off));
  static const uint32_t zero32[1] = {0};
  TRY(h2d(e, b->locus_motif_off, n_loci ? locus_motif_offsets : zero32, ((size_t)n_loci + 1) * sizeof(uint32_t)));
  TRY(h2d(e, b->allele_locus, allele_locus, n * sizeof(uint32_t)));
  TRY(stage_begin(e, (n_gen + 1) * 2 * sizeof(unsigned long long) + n_gen * sizeof(uint32_t) +
                         b->h_slots.size() * sizeof(uint32_t) + b->h_group_off.size() * sizeof(unsigned long long) +
                         b->h_group_n.size() + (n + 1) * sizeof(unsigned long long) + 8 * 16));
  if (n_gen) {
    TRY(h2d_staged(e, b->bp_off, b->h_bp_off.data(), (n_gen + 1) * sizeof(unsigned long long)));
    TRY(h2d_staged(e, b->scr_off, b->h_scr_off.data(), (n_gen + 1) * sizeof(unsigned long long)));
    if (n_gen != n) TRY(h2d_staged(e, b->list, b->h_list.data(), n_gen * sizeof(uint32_t)));  // (all of them: identity)
  }
  if (n_lane) {
    TRY(h2d_staged(e, b->slots, b->h_slots.data(), b->h_slots.size() * sizeof(uint32_t)));
    TRY(h2d_staged(e, b->group_off, b->h_group_off.data(), b->h_group_off.size() * sizeof(unsigned long long)));
    TRY(h2d_staged(e, b->group_n, b->h_group_n.data(), b->h_group_n.size()));
  }
  TRY(h2d_staged(e, b->mc_off, b->h_mc_off.data(), (n + 1) * sizeof(unsigned long long)));
  CU(e, engine_wait(e));  // h_* vectors may be reallocated by the next upload
  TRY(dev_reserve(e, b->mc, (size_t)(b->h_mc_off[n] + 1) * sizeof(uint32_t)));
  TRY(dev_reserve(e, b->purity, (n + 1) * sizeof(double)));
  TRY(dev_reserve(e, b->n_spans, (n + 2) * sizeof(uint32_t)));
  TRY(dev_reserve(e, b->span_off, (n + 2) * sizeof(unsigned long long)));
  TRY(dev_reserve(e, b->status, (n + 1) * sizeof(int32_t)));
  if (want_paths) {
    TRY(dev_reserve(e, b->path_len, (n + 2) * sizeof(unsigned long long)));
    TRY(dev_reserve(e, b->path_off, (n + 2) * sizeof(unsigned long long)));
  }
  size_t bp_need = 0, lane_need = 0;
  for (auto &w : b->waves) {
    const size_t bytes = (size_t)(b->h_bp_off[w.second] - b->h_bp_off[w.first]);
    if (bytes > bp_need) bp_need = bytes;
  }
  for (auto &w : b->lane_waves) {
    const size_t bytes = (size_t)(b->h_group_off[w.second] - b->h_group_off[w.first]) * sizeof(uint32_t);
    if (bytes > lane_need) lane_need = bytes;
  }
  TRY(dev_reserve(e, b->bp, bp_need + 16));
  TRY(dev_reserve(e, b->lane_bp, lane_need + 16));
  // span scratch of the whole batch: one slot per base (generic alleles) / per packed word (lane groups)
  if (!want_paths) TRY(dev_reserve(e, b->span_scratch, (size_t)(b->h_scr_off[n_gen] + 1) * sizeof(trgt_motif_span_t)));
  TRY(dev_reserve(e, b->lane_scratch, (size_t)(b->h_group_off.back() + 1) * sizeof(trgt_motif_span_t)));
  return 0;
}

int32_t trgt_hmm_upload(trgt_engine_t *e, const trgt_seqs_t *motifs, const uint32_t *locus_motif_offsets,
                        uint32_t n_loci, const trgt_seqs_t *alleles, const uint32_t *allele_locus,
                        int32_t want_paths, trgt_hmm_batch_t **out) {
  if (!e || !out) return TRGT_ERR_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  *out = nullptr;
  trgt_hmm_batch *b = new trgt_hmm_batch();
  const int rc = hmm_upload_into(e, b, motifs, locus_motif_offsets, n_loci, alleles, allele_locus, want_paths);
  if (rc != 0) {
    trgt_hmm_free(nullptr, b);
    return rc;
  }
  *out = b;
  return 0;
}

// inclusive-to-exclusive helper for u64 path lengths: offsets[i] = sum of len[0..i)
__global__ void k_scan_u64_serial(const unsigned long long *len, unsigned long long *off, uint32_t a0, uint32_t a1,
                                  unsigned long long base) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long acc = base;
    for (uint32_t a = a0; a < a1; a++) {
      off[a] = acc;
      acc += len[a];
    }
    off[a1] = acc;
  }
}

__global__ void k_add_base_u64(unsigned long long *v, uint32_t a0, uint32_t a1, unsigned long long base) {
  const uint32_t gsz = gridDim.x * blockDim.x;
  for (uint32_t a = a0 + blockIdx.x * blockDim.x + threadIdx.x; a <= a1; a += gsz) v[a] += base;
}

static int hmm_run_locked(trgt_engine_t *e, trgt_hmm_batch *b) {
  CU(e, cudaSetDevice(e->device));
  b->ran = true;
  b->total_spans = 0;
  b->total_path = 0;
  if (b->n_alleles == 0) return 0;
  const uint32_t n = b->n_alleles, n_gen = (uint32_t)b->h_list.size();
  const uint32_t n_groups = (uint32_t)(b->h_group_off.size() - 1);
  HmmBatch hb;
  hb.motifs = (const uint8_t *)b->motifs.p;
  hb.motif_off = (const uint64_t *)b->motif_off.p;
  hb.locus_motif_off = (const uint32_t *)b->locus_motif_off.p;
  hb.alleles = (const uint8_t *)b->alleles.p;
  hb.allele_off = (const uint64_t *)b->allele_off.p;
  hb.allele_locus = (const uint32_t *)b->allele_locus.p;
  hb.list = n_gen != n ? (const uint32_t *)b->list.p : nullptr;
  hb.bp_off = (const unsigned long long *)b->bp_off.p;
  hb.scr_off = (const unsigned long long *)b->scr_off.p;
  hb.mc_off = (const unsigned long long *)b->mc_off.p;
  hb.mm_off = (const uint32_t *)e->d_mm_off.p;
  hb.mm_lp = (const double *)e->d_mm_lp.p;
  hb.c = e->hmm_consts;
  hb.S_max = b->S_max;
  hb.nb_max = b->nb_max;
  hb.mbytes_max = b->mbytes_max;
  hb.warp_bytes = b->warp_bytes;
  HmmLaneBatch lb;
  lb.motifs = hb.motifs; lb.motif_off = hb.motif_off; lb.locus_motif_off = hb.locus_motif_off;
  lb.alleles = hb.alleles; lb.allele_off = hb.allele_off; lb.allele_locus = hb.allele_locus;
  lb.mc_off = hb.mc_off;
  lb.slots = (const uint32_t *)b->slots.p;
  lb.group_off = (const unsigned long long *)b->group_off.p;
  lb.group_n = (const uint8_t *)b->group_n.p;
  lb.c = e->hmm_consts;
  for (int n1 = 0; n1 <= HMM_LANE_NMAX; n1++)
    for (int i = 0; i < HMM_LANE_NMAX; i++)
      lb.jump[n1][i] = (n1 >= 1 && i < n1) ? e->jump.lp[e->jump.off[n1] + i] : 0.0;
  unsigned long long *path_len = b->want_paths ? (unsigned long long *)b->path_len.p : nullptr;
  // ---- generic alleles: Viterbi (a lane or a warp per allele) and the counting walk, wave by wave ----
  const int block = 128, wpb = 4;
  const size_t smem = b->warp_bytes * wpb;
  int grid_v = 0;
  if (n_gen) TRY(persistent_grid(e, k_hmm_viterbi, block, smem, &grid_v));
  for (auto &w : b->waves) {
    const uint32_t a0 = w.first, a1 = w.second, cnt = a1 - a0;
    const uint32_t need = (cnt + wpb - 1) / wpb;
    const uint32_t tgrid = (cnt + 127) / 128;  // one thread per allele for the walks
    const unsigned long long bp_base = b->h_bp_off[a0];
    {  // small models: one allele per thread
      const int s_cap = b->S_max < HMM_THREAD_S ? b->S_max : HMM_THREAD_S;  // two score columns per thread
      const size_t tsmem = (size_t)2 * s_cap * 128 * sizeof(double);
      int grid = 0;
      TRY(persistent_grid(e, k_hmm_viterbi_thread, 128, tsmem, &grid));
      if ((uint32_t)grid > tgrid) grid = (int)tgrid;
      LaunchScope ls(e, "k_hmm_viterbi_thread");
      k_hmm_viterbi_thread<<<grid, 128, tsmem, e->stream>>>(hb, a0, a1, bp_base, (uint8_t *)b->bp.p,
                                                             (int32_t *)b->status.p, s_cap);
      TRY(check_launch(e, "k_hmm_viterbi_thread"));
    }
    if (b->S_max > HMM_THREAD_S) {  // larger models: one allele per warp
      LaunchScope ls(e, "k_hmm_viterbi");
      const int grid = (uint32_t)grid_v > need ? (int)need : grid_v;
      k_hmm_viterbi<<<grid, block, smem, e->stream>>>(hb, a0, a1, bp_base, (uint8_t *)b->bp.p, (int32_t *)b->status.p, 1);
      TRY(check_launch(e, "k_hmm_viterbi"));
    }
    // without state paths the walk leaves the spans in scratch slots and a copy kernel places them;
    // with state paths a second walk (k_hmm_emit) writes spans and paths at their offsets
    {
      LaunchScope ls(e, "k_hmm_walk");
      k_hmm_walk<<<tgrid, 128, 0, e->stream>>>(hb, a0, a1, bp_base, (const uint8_t *)b->bp.p, (uint32_t *)b->mc.p,
                                               (double *)b->purity.p, (uint32_t *)b->n_spans.p, path_len,
                                               (int32_t *)b->status.p,
                                               b->want_paths ? nullptr : (trgt_motif_span_t *)b->span_scratch.p);
      TRY(check_launch(e, "k_hmm_walk"));
    }
  }
  // ---- single-motif loci: lane kernels per motif length, wave by wave ----
  for (auto &w : b->lane_waves) {
    const unsigned long long word_base = b->h_group_off[w.first];
    const uint32_t lgrid = (w.second - w.first + 3) / 4;
    {
      LaunchScope ls(e, "k_hmm_lane_viterbi");
      k_hmm_lane_viterbi<<<lgrid, 128, 0, e->stream>>>(lb, w.first, w.second, word_base, (uint32_t *)b->lane_bp.p,
                                                       (int32_t *)b->status.p);
      TRY(check_launch(e, "k_hmm_lane_viterbi"));
    }
    {
      LaunchScope ls(e, "k_hmm_lane_walk");
      k_hmm_lane_walk<<<lgrid, 128, 0, e->stream>>>(lb, w.first, w.second, word_base, (const uint32_t *)b->lane_bp.p,
                                                    (uint32_t *)b->mc.p, (double *)b->purity.p, (uint32_t *)b->n_spans.p,
                                                    path_len, (int32_t *)b->status.p,
                                                    (trgt_motif_span_t *)b->lane_scratch.p);
      TRY(check_launch(e, "k_hmm_lane_walk"));
    }
  }
  // ---- span (and path) offsets of the whole batch ----
  CU(e, cudaMemsetAsync((uint32_t *)b->n_spans.p + n, 0, sizeof(uint32_t), e->stream));
  TRY(exclusive_scan_u32(e, (const uint32_t *)b->n_spans.p, (unsigned long long *)b->span_off.p, (size_t)n + 1));
  CU(e, cudaMemcpyAsync(&e->h_u64[0], (unsigned long long *)b->span_off.p + n, sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
  if (b->want_paths) {
    LaunchScope ls(e, "k_scan_u64_serial");
    k_scan_u64_serial<<<1, 32, 0, e->stream>>>((const unsigned long long *)b->path_len.p, (unsigned long long *)b->path_off.p, 0, n, 0);
    TRY(check_launch(e, "k_scan_u64_serial"));
    CU(e, cudaMemcpyAsync(&e->h_u64[1], (unsigned long long *)b->path_off.p + n, sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
  }
  CU(e, engine_wait(e));
  const unsigned long long span_end = e->h_u64[0];
  const unsigned long long path_end = b->want_paths ? e->h_u64[1] : 0;
  TRY(dev_reserve(e, b->spans, (size_t)(span_end + 1) * sizeof(trgt_motif_span_t)));
  if (b->want_paths) TRY(dev_reserve(e, b->paths, (size_t)(path_end + 1) * sizeof(uint32_t)));
  // ---- spans to their CSR offsets; state paths by a second walk ----
  if (n_gen && !b->want_paths && span_end) {
    LaunchScope ls(e, "k_hmm_spans");
    k_hmm_spans<<<(n_gen + 127) / 128, 128, 0, e->stream>>>(hb, 0, n_gen, (const trgt_motif_span_t *)b->span_scratch.p,
                                                            (const uint32_t *)b->n_spans.p,
                                                            (const unsigned long long *)b->span_off.p,
                                                            (trgt_motif_span_t *)b->spans.p);
    TRY(check_launch(e, "k_hmm_spans"));
  }
  if (n_gen && b->want_paths && (span_end || path_end)) {  // (one wave: every back-pointer is still in place)
    LaunchScope ls(e, "k_hmm_emit");
    k_hmm_emit<<<(n_gen + 127) / 128, 128, 0, e->stream>>>(hb, 0, n_gen, 0, (const uint8_t *)b->bp.p,
                                                           (const uint32_t *)b->n_spans.p,
                                                           (const unsigned long long *)b->span_off.p,
                                                           (trgt_motif_span_t *)b->spans.p,
                                                           (const unsigned long long *)b->path_off.p,
                                                           (uint32_t *)b->paths.p, (const int32_t *)b->status.p);
    TRY(check_launch(e, "k_hmm_emit"));
  }
  if (n_groups && span_end) {
    LaunchScope ls(e, "k_hmm_lane_spans");
    k_hmm_lane_spans<<<(n_groups + 3) / 4, 128, 0, e->stream>>>(lb, 0, n_groups, (const trgt_motif_span_t *)b->lane_scratch.p,
                                                                (const uint32_t *)b->n_spans.p,
                                                                (const unsigned long long *)b->span_off.p,
                                                                (trgt_motif_span_t *)b->spans.p);
    TRY(check_launch(e, "k_hmm_lane_spans"));
  }
  if (n_groups && b->want_paths && path_end) {
    LaunchScope ls(e, "k_hmm_lane_emit");
    k_hmm_lane_emit<<<(n_groups + 3) / 4, 128, 0, e->stream>>>(lb, 0, n_groups, 0ull, (const uint32_t *)b->lane_bp.p,
                                                               (const unsigned long long *)b->path_off.p,
                                                               (uint32_t *)b->paths.p, (const int32_t *)b->status.p);
    TRY(check_launch(e, "k_hmm_lane_emit"));
  }
  b->total_spans = span_end;
  b->total_path = path_end;
  return 0;
}

int32_t trgt_hmm_run(trgt_engine_t *e, trgt_hmm_batch_t *b) {
  if (!e || !b) return TRGT_ERR_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  return hmm_run_locked(e, b);
}

static int hmm_download_locked(trgt_engine_t *e, trgt_hmm_batch *b, trgt_annotations_t *out) {
  if (!out) return fail(e, TRGT_ERR_ARG, "out is null");
  if (!b->ran) return fail(e, TRGT_ERR_ARG, "trgt_hmm_download before trgt_hmm_run");
  CU(e, cudaSetDevice(e->device));
  const size_t n = b->n_alleles;
  static const uint64_t zero_off[1] = {0};
  const size_t n_mc = b->h_mc_off.empty() ? 0 : (size_t)b->h_mc_off[n];
  TRY(pin_reserve(e, b->r_mc, (n_mc + 1) * sizeof(uint32_t)));
  TRY(pin_reserve(e, b->r_span_off, (n + 1) * sizeof(uint64_t)));
  TRY(pin_reserve(e, b->r_spans, (size_t)(b->total_spans + 1) * sizeof(trgt_motif_span_t)));
  TRY(pin_reserve(e, b->r_purity, (n + 1) * sizeof(double)));
  TRY(pin_reserve(e, b->r_status, (n + 1) * sizeof(int32_t)));
  TRY(pin_reserve(e, b->r_path_off, (n + 1) * sizeof(uint64_t)));
  TRY(pin_reserve(e, b->r_paths, (size_t)(b->total_path + 1) * sizeof(uint32_t)));
  b->r_span_off.as<uint64_t>()[0] = 0;
  b->r_path_off.as<uint64_t>()[0] = 0;
  if (n) {
    if (n_mc) CU(e, cudaMemcpyAsync(b->r_mc.p, b->mc.p, n_mc * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaMemcpyAsync(b->r_span_off.p, b->span_off.p, (n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
    if (b->total_spans)
      CU(e, cudaMemcpyAsync(b->r_spans.p, b->spans.p, (size_t)b->total_spans * sizeof(trgt_motif_span_t), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaMemcpyAsync(b->r_purity.p, b->purity.p, n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaMemcpyAsync(b->r_status.p, b->status.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    if (b->want_paths) {
      CU(e, cudaMemcpyAsync(b->r_path_off.p, b->path_off.p, (n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
      if (b->total_path)
        CU(e, cudaMemcpyAsync(b->r_paths.p, b->paths.p, (size_t)b->total_path * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    }
    CU(e, engine_wait(e));
  }
  out->n = n;
  out->motif_count_offsets = b->h_mc_off.empty() ? zero_off : (const uint64_t *)b->h_mc_off.data();
  out->motif_counts = b->r_mc.as<uint32_t>();
  out->span_offsets = b->r_span_off.as<uint64_t>();
  out->spans = b->r_spans.as<trgt_motif_span_t>();
  out->purity = b->r_purity.as<double>();
  out->status = b->r_status.as<int32_t>();
  out->path_offsets = b->want_paths ? b->r_path_off.as<uint64_t>() : nullptr;
  out->paths = b->want_paths ? b->r_paths.as<uint32_t>() : nullptr;
  return 0;
}

int32_t trgt_vcf_fields(trgt_engine_t *e, trgt_hmm_batch_t *b, trgt_seqs_out_t *out) {
  if (!e || !out) return TRGT_ERR_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  if (!b) b = e->one_hmm;
  if (!b || !b->ran) return fail(e, TRGT_ERR_ARG, "trgt_vcf_fields: no HMM batch has been run");
  if (b->h_locus_allele_off.empty() && b->n_alleles)
    return fail(e, TRGT_ERR_ARG, "trgt_vcf_fields: the alleles of a locus must be consecutive (allele_locus non-decreasing)");
  CU(e, cudaSetDevice(e->device));
  memset(out, 0, sizeof *out);
  const size_t nf = 4 * (size_t)b->n_loci;
  TRY(pin_reserve(e, b->r_vcf_off, (nf + 1) * sizeof(uint64_t)));
  uint64_t *h_off = b->r_vcf_off.as<uint64_t>();
  h_off[0] = 0;
  out->n = nf;
  out->offsets = h_off;
  if (nf == 0) return 0;
  if (nf + 1 > 0x7fffffffull) return fail(e, TRGT_ERR_ARG, "too many loci in one batch");
  if (b->h_locus_allele_off.empty()) b->h_locus_allele_off.assign((size_t)b->n_loci + 1, 0);
  TRY(h2d(e, b->locus_allele_off, b->h_locus_allele_off.data(), ((size_t)b->n_loci + 1) * sizeof(uint32_t)));
  TRY(dev_reserve(e, b->vcf_len, (nf + 1) * sizeof(uint32_t)));
  TRY(dev_reserve(e, b->vcf_off, (nf + 1) * sizeof(uint64_t)));
  CU(e, cudaMemsetAsync((uint32_t *)b->vcf_len.p + nf, 0, sizeof(uint32_t), e->stream));
  int grid = 0;
  TRY(persistent_grid(e, k_vcf_fields<false>, 128, 0, &grid));
  const uint32_t need = (b->n_loci + 127) / 128;
  if ((uint32_t)grid > need) grid = (int)need;
#define VCF_ARGS(outp)                                                                                                \
  b->n_loci, (const uint32_t *)b->locus_allele_off.p, (const uint64_t *)b->allele_off.p,                              \
      (const unsigned long long *)b->mc_off.p, (const uint32_t *)b->mc.p, (const unsigned long long *)b->span_off.p,  \
      (const trgt_motif_span_t *)b->spans.p, (const double *)b->purity.p, (const int32_t *)b->status.p,               \
      (uint32_t *)b->vcf_len.p, (const unsigned long long *)b->vcf_off.p, (outp)
  {
    LaunchScope ls(e, "k_vcf_fields_count");
    k_vcf_fields<false><<<grid, 128, 0, e->stream>>>(VCF_ARGS((uint8_t *)nullptr));
    TRY(check_launch(e, "k_vcf_fields_count"));
  }
  TRY(exclusive_scan_u32(e, (const uint32_t *)b->vcf_len.p, (unsigned long long *)b->vcf_off.p, nf + 1));
  CU(e, cudaMemcpyAsync(h_off, b->vcf_off.p, (nf + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
  CU(e, engine_wait(e));
  const uint64_t total = h_off[nf];
  TRY(dev_reserve(e, b->vcf_data, (size_t)total + 16));
  TRY(pin_reserve(e, b->r_vcf_data, (size_t)total + 16));
  {
    LaunchScope ls(e, "k_vcf_fields_write");
    k_vcf_fields<true><<<grid, 128, 0, e->stream>>>(VCF_ARGS((uint8_t *)b->vcf_data.p));
    TRY(check_launch(e, "k_vcf_fields_write"));
  }
#undef VCF_ARGS
  if (total) CU(e, cudaMemcpyAsync(b->r_vcf_data.p, b->vcf_data.p, (size_t)total, cudaMemcpyDeviceToHost, e->stream));
  CU(e, engine_wait(e));
  out->data = b->r_vcf_data.as<uint8_t>();
  return 0;
}

int32_t trgt_hmm_download(trgt_engine_t *e, trgt_hmm_batch_t *b, trgt_annotations_t *out) {
  if (!e || !b) return TRGT_ERR_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  return hmm_download_locked(e, b, out);
}

int32_t trgt_hmm_label(trgt_engine_t *e, const trgt_seqs_t *motifs, const uint32_t *locus_motif_offsets,
                       uint32_t n_loci, const trgt_seqs_t *alleles, const uint32_t *allele_locus, int32_t want_paths,
                       trgt_annotations_t *out) {
  if (!e) return TRGT_ERR_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  if (!e->one_hmm) e->one_hmm = new trgt_hmm_batch();
  TRY(hmm_upload_into(e, e->one_hmm, motifs, locus_motif_offsets, n_loci, alleles, allele_locus, want_paths));
  TRY(hmm_run_locked(e, e->one_hmm));
  return hmm_download_locked(e, e->one_hmm, out);
}

void trgt_engine_set_workspace_budget(trgt_engine_t *e, size_t bytes) {
  if (e && bytes >= ((size_t)1 << 20)) e->workspace_budget = bytes;
}

void trgt_engine_set_hmm_lane_path(trgt_engine_t *e, int32_t on) {
  if (e) e->hmm_lane = on != 0;
}

void trgt_engine_set_flank_band_budget(trgt_engine_t *e, int32_t max_cost) {
  if (e) e->band_budget = max_cost < 0 ? 0 : (max_cost > 64 ? 64 : max_cost);
}

}  // extern "C"
