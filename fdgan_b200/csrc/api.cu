// Error reporting, launch accounting and version of the fdgan_b200 C ABI.
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace fdg {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// One process may drive several GPUs from several threads (nn.DataParallel, demo.py:89): function attributes, __constant__
// uploads and the SM count are per device, so the kernels' one-time launch state is indexed by the current device.
int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  return dev < 64 ? dev : 63;
}

int device_sm_count() {
  static std::atomic<int> sms[64];
  const int dev = current_device();
  int n = sms[dev].load(std::memory_order_relaxed);
  if (n <= 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    sms[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("FDG_PDL"); return e ? atoi(e) != 0 : true; }();
  return on;
}

bool epi_warp_stores() {
  static const bool on = [] { const char* e = getenv("FDG_EPI_WARP"); return e ? atoi(e) != 0 : true; }();
  return on;
}

static cudaEvent_t* g_event_ring[64] = {nullptr};      // per device: the ring fdg_event_record hands handles out of

static int g_dbg = 0;
int dbg_flags() { return g_dbg; }
void set_dbg_flags(int v) { g_dbg = v; }

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
    return FDG_ECUDA;
  }
  count_launch(1);
  return FDG_OK;
}

// ---------------------------------------------------------------- per-launch event profiling
// A measurement aid for bench.py (one process = one GPU).  The event pool belongs to the device that was current when
// profiling was enabled; launches on any other device (nn.DataParallel worker threads) are simply not recorded.
struct ProfRec { int family; double flops, bytes; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof_recs;
static std::vector<cudaEvent_t> g_prof_pool;
static std::atomic<int> g_prof_on{0};
static std::atomic<int> g_prof_dev{0};

static cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

ProfScope::ProfScope(int family, double flops, double bytes, cudaStream_t stream) : idx(-1), st(stream) {
  if (!g_prof_on.load(std::memory_order_relaxed) || current_device() != g_prof_dev.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r{family, flops, bytes, prof_event(), prof_event()};
  cudaEventRecord(r.a, st);
  g_prof_recs.push_back(r);
  idx = (int)g_prof_recs.size() - 1;
}

ProfScope::~ProfScope() {
  if (idx < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (idx < (int)g_prof_recs.size()) cudaEventRecord(g_prof_recs[idx].b, st);
}

}  // namespace fdg

extern "C" {

int fdg_profile_enable(int on) {
  if (on) fdg::g_prof_dev.store(fdg::current_device());
  fdg::g_prof_on.store(on ? 1 : 0);
  return FDG_OK;
}

// Sums the recorded launches per family (ms, algorithmic flops, algorithmic bytes, launch count) and clears them.
// Arrays have FDG_PROF_FAMILIES entries.  Synchronises on the recorded events.
int fdg_profile_collect(double* ms, double* flops, double* bytes, int64_t* launches) {
  FDG_REQUIRE(ms && flops && bytes && launches, "fdg_profile_collect: null output array");
  std::lock_guard<std::mutex> lk(fdg::g_prof_mu);
  for (int f = 0; f < fdg::PF_COUNT; ++f) { ms[f] = 0; flops[f] = 0; bytes[f] = 0; launches[f] = 0; }
  for (auto& r : fdg::g_prof_recs) {
    float t = 0.f;
    if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
      ms[r.family] += t; flops[r.family] += r.flops; bytes[r.family] += r.bytes; launches[r.family] += 1;
    }
    fdg::g_prof_pool.push_back(r.a);
    fdg::g_prof_pool.push_back(r.b);
  }
  fdg::g_prof_recs.clear();
  cudaGetLastError();
  return FDG_OK;
}

// Light-weight stream fork / join for the executors (one ctypes call each instead of a torch Event object + context manager):
// a per-device ring of timing-disabled events.  fdg_event_record enqueues a record on `stream` and returns a handle;
// fdg_stream_wait makes `stream` wait for that point.  Legal under stream capture (they become graph dependencies).  A handle
// stays valid for FDG_EVENT_RING further records on its device.
int fdg_event_record(fdg_stream_t stream) {
  constexpr int RING = 1024;
  static cudaEvent_t ring[64][RING];
  static std::atomic<unsigned> next[64];
  static std::mutex mu;
  const int dev = fdg::current_device();
  const unsigned slot = next[dev].fetch_add(1u) % RING;
  cudaEvent_t& ev = ring[dev][slot];
  if (!ev) {
    std::lock_guard<std::mutex> lk(mu);
    if (!ev && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { fdg::set_error("fdg_event_record: cannot create an event"); return FDG_ECUDA; }
  }
  if (cudaEventRecord(ev, (cudaStream_t)stream) != cudaSuccess) { fdg::set_error("fdg_event_record: cudaEventRecord failed: %s", cudaGetErrorString(cudaGetLastError())); return FDG_ECUDA; }
  fdg::g_event_ring[dev] = &ring[dev][0];
  return (int)slot;
}

int fdg_stream_wait(fdg_stream_t stream, int handle) {
  const int dev = fdg::current_device();
  FDG_REQUIRE(handle >= 0 && handle < 1024 && fdg::g_event_ring[dev] && fdg::g_event_ring[dev][handle], "fdg_stream_wait: bad event handle %d", handle);
  if (cudaStreamWaitEvent((cudaStream_t)stream, fdg::g_event_ring[dev][handle], 0) != cudaSuccess) {
    fdg::set_error("fdg_stream_wait: cudaStreamWaitEvent failed: %s", cudaGetErrorString(cudaGetLastError()));
    return FDG_ECUDA;
  }
  return FDG_OK;
}

const char* fdg_last_error(void) { return fdg::g_err; }

int fdg_version(void) { return 100; }

int64_t fdg_launch_count(void) { return fdg::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
