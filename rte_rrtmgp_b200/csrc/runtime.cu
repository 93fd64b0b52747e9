// runtime.cu - backend plumbing of the CUDA library (include/rrtmgp_b200_ext.h, first section)
//
// Threading contract (SURVEY 8b "Threading / re-entrancy"): the reference kernels are serial and stateless, so a host
// may call them concurrently from many threads on disjoint column blocks (its own OpenMP-over-blocks idiom,
// examples/rfmip-clear-sky/rrtmgp_rfmip_lw.F90:177-178).  Here every HOST THREAD has its own launch stream
// (thread_local; cudaStreamPerThread until the thread calls rrtmgpb_set_stream), scratch is allocated and freed
// stream-ordered on that stream, the profiler's records are guarded by a mutex, table caches by theirs, and the
// process-wide switches (solver variant, checks, ...) are atomics that are meant to be set once at start-up.
#include <algorithm>
#include <atomic>
#include <mutex>
#include <map>
#include <string>
#include <vector>
#include <cstring>
#include "common.cuh"
#include "rrtmgp_b200_ext.h"

namespace rrtmgpb {

thread_local const char* tl_op_name = nullptr;
thread_local ExpressSolverMode tl_express;
thread_local bool tl_trust_device_ptrs = false;
// cudaStreamPerThread is a blocking stream: it orders itself against the legacy default stream, so hosts that mix this
// library with legacy-stream work (torch's default stream, plain cudaMemcpy) keep the ordering they had
static thread_local cudaStream_t tl_stream = cudaStreamPerThread;
static std::atomic<long long> g_launches{0};

cudaStream_t stream() { return tl_stream; }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// keep freed scratch in the pool instead of returning it to the driver after every sync - once per DEVICE
static std::atomic<unsigned long long> g_pool_ready{0};  // bit d: device d's default pool is configured
static std::mutex g_pool_mutex;
static void init_pool() {
  int dev = 0;
  RB_CUDA_CHECK(cudaGetDevice(&dev));
  const unsigned long long bit = 1ull << (dev & 63);
  if (g_pool_ready.load(std::memory_order_acquire) & bit) return;
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  if (g_pool_ready.load(std::memory_order_relaxed) & bit) return;
  cudaMemPool_t pool;
  RB_CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, dev));
  unsigned long long thresh = ~0ull;
  RB_CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
  g_pool_ready.fetch_or(bit, std::memory_order_release);
}

void* dev_alloc(size_t bytes) {
  init_pool();
  void* p = nullptr;
  RB_CUDA_CHECK(cudaMallocAsync(&p, bytes ? bytes : 16, stream()));
  return p;
}
void dev_free(void* p) {
  if (p) RB_CUDA_CHECK(cudaFreeAsync(p, stream()));
}

// ---- per-kernel event timing (records of all threads in one list, guarded by a mutex) -----------
struct TimerRec { const char* name; cudaEvent_t a, b; };
static std::atomic<bool> g_profile{false};
static std::mutex g_prof_mutex;
static std::vector<TimerRec> g_recs;
static std::vector<cudaEvent_t> g_free_events;
static cudaEvent_t get_event() {  // g_prof_mutex held
  if (!g_free_events.empty()) { cudaEvent_t e = g_free_events.back(); g_free_events.pop_back(); return e; }
  cudaEvent_t e; RB_CUDA_CHECK(cudaEventCreate(&e)); return e;
}
KernelTimer::KernelTimer(const char* name) : slot(-1), b_(nullptr) {
  if (!g_profile.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  TimerRec r{name, get_event(), get_event()};
  RB_CUDA_CHECK(cudaEventRecord(r.a, stream()));
  g_recs.push_back(r);
  slot = (int)g_recs.size() - 1;
  b_ = r.b;
}
KernelTimer::~KernelTimer() {
  if (slot >= 0) RB_CUDA_CHECK(cudaEventRecord(static_cast<cudaEvent_t>(b_), stream()));
}

bool is_device_ptr(const void* p) {
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

}  // namespace rrtmgpb

using namespace rrtmgpb;

extern "C" {
const char* rrtmgpb_backend_name(void) { return "cuda-sm_100a"; }
int rrtmgpb_float_bytes(void) { return (int)sizeof(Float); }
void* rrtmgpb_mem_alloc(size_t bytes) { return dev_alloc(bytes); }
void rrtmgpb_mem_free(void* p) {
  table_cache_release(p);  // no-op unless p is a k-distribution table with transposed copies
  table_cache_release_abi(p);
  dev_free(p);
}
// (cudaMemcpyDefault: the direction follows from the pointers - a host program that hands HOST arrays to the frontend
// mirror, as a stock Fortran host does with the extern kernels, reaches these with either kind of pointer)
void rrtmgpb_mem_to_backend(void* d, const void* s, size_t n) {
  RB_CUDA_CHECK(cudaMemcpyAsync(d, s, n, cudaMemcpyDefault, stream()));
}
void rrtmgpb_mem_to_host(void* d, const void* s, size_t n) {
  RB_CUDA_CHECK(cudaMemcpyAsync(d, s, n, cudaMemcpyDefault, stream()));
  RB_CUDA_CHECK(cudaStreamSynchronize(stream()));
}
void rrtmgpb_mem_copy(void* d, const void* s, size_t n) {
  RB_CUDA_CHECK(cudaMemcpyAsync(d, s, n, cudaMemcpyDefault, stream()));
  if (!is_device_ptr(d) || !is_device_ptr(s)) RB_CUDA_CHECK(cudaStreamSynchronize(stream()));  // pageable host memory involved
}
/* the CALLING THREAD's launch stream (every host thread has its own; default cudaStreamPerThread) */
void rrtmgpb_set_stream(void* s) { tl_stream = static_cast<cudaStream_t>(s); }
void* rrtmgpb_get_stream(void) { return tl_stream; }
void rrtmgpb_set_device(int d) { RB_CUDA_CHECK(cudaSetDevice(d)); }
void rrtmgpb_sync(void) { RB_CUDA_CHECK(cudaStreamSynchronize(stream())); }
void rrtmgpb_profile_enable(int on) { g_profile.store(on != 0); }
// Writes "name count total_ms\n" per kernel (sorted by time) into buf; clears the records.
int rrtmgpb_profile_report(char* buf, size_t buflen) {
  RB_CUDA_CHECK(cudaStreamSynchronize(stream()));
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  std::map<std::string, std::pair<int, double>> agg;
  for (auto& r : g_recs) {
    float ms = 0;
    RB_CUDA_CHECK(cudaEventSynchronize(r.b));  // records of other host threads' streams
    RB_CUDA_CHECK(cudaEventElapsedTime(&ms, r.a, r.b));
    auto& e = agg[r.name];
    e.first += 1; e.second += ms;
    g_free_events.push_back(r.a); g_free_events.push_back(r.b);
  }
  g_recs.clear();
  std::vector<std::pair<double, std::string>> order;
  for (auto& kv : agg) order.push_back({-kv.second.second, kv.first});
  std::sort(order.begin(), order.end());
  std::string out;
  for (auto& o : order) {
    char line[256];
    std::snprintf(line, sizeof line, "%s %d %.6f\n", o.second.c_str(), agg[o.second].first, agg[o.second].second);
    out += line;
  }
  if (buf && buflen) { std::strncpy(buf, out.c_str(), buflen - 1); buf[buflen - 1] = 0; }
  return (int)order.size();
}
long long rrtmgpb_launch_count(int reset) {
  long long v = g_launches.load();
  if (reset) g_launches.store(0);
  return v;
}
}
