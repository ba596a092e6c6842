// api.cu -- ABI version, error plumbing, launch counter and the optional per-kernel event timer.
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include "common.cuh"
namespace dr4sr { int fused_fwd_set_trace(int* host_mapped); int attn_bwd_set_trace(int* host_mapped); }

namespace dr4sr {
static thread_local char g_err[256] = "";
void set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
}

std::atomic<int> g_gemm_backend{0};
std::atomic<int> g_attn_backend{2};
std::atomic<int> g_fused_backend{2};
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

struct ProfRec { const char* name; cudaEvent_t beg, end; };
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;

ProfScope::ProfScope(const char* n, cudaStream_t s) : name(n), st(s), slot(-1) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  ProfRec r{n, nullptr, nullptr};
  if (cudaEventCreate(&r.beg) != cudaSuccess || cudaEventCreate(&r.end) != cudaSuccess) return;
  cudaEventRecord(r.beg, s);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  slot = (int)g_prof.size();
  g_prof.push_back(r);
}
ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_prof[slot].end, st);
}
// Auxiliary stream for short independent kernels (fork / join with events; legal inside a CUDA-graph capture of the main
// stream).  One per device, created on first use, never destroyed.  aux_fork: work queued on the returned stream starts
// once everything enqueued on `st` so far has finished; aux_join: `st` continues after that work.  Falls back to `st` itself.
struct AuxStream { cudaStream_t s = nullptr; cudaEvent_t fork = nullptr, join = nullptr; bool ok = false; };
static AuxStream& aux_stream() {
  static AuxStream per_dev[64];
  int dev = 0;
  cudaGetDevice(&dev);
  AuxStream& x = per_dev[dev & 63];
  if (!x.ok) {
    bool good = cudaStreamCreateWithFlags(&x.s, cudaStreamNonBlocking) == cudaSuccess;
    good = good && cudaEventCreateWithFlags(&x.fork, cudaEventDisableTiming) == cudaSuccess;
    good = good && cudaEventCreateWithFlags(&x.join, cudaEventDisableTiming) == cudaSuccess;
    x.ok = good;
  }
  return x;
}
cudaStream_t aux_fork(cudaStream_t st) {
  AuxStream& x = aux_stream();
  if (!x.ok || cudaEventRecord(x.fork, st) != cudaSuccess || cudaStreamWaitEvent(x.s, x.fork, 0) != cudaSuccess) return st;
  return x.s;
}
int aux_join(cudaStream_t aux, cudaStream_t st) {
  if (aux == st) return DR4SR_OK;
  AuxStream& x = aux_stream();
  if (cudaEventRecord(x.join, aux) != cudaSuccess || cudaStreamWaitEvent(st, x.join, 0) != cudaSuccess) {
    set_cuda_error(cudaGetLastError(), "aux join");
    return DR4SR_ECUDA;
  }
  return DR4SR_OK;
}
// Background stream: one longer kernel that may run under everything the main stream does until bg_join (the scatter-add of
// the target / negative gradient rows runs under the whole encoder backward).  Same fork / join protocol, its own stream, so
// the short kernels of the auxiliary stream never queue behind it.
struct BgStream { cudaStream_t s = nullptr; cudaEvent_t fork = nullptr, join = nullptr; bool ok = false, pending = false; };
static BgStream& bg_stream() {
  static BgStream per_dev[64];
  int dev = 0;
  cudaGetDevice(&dev);
  BgStream& x = per_dev[dev & 63];
  if (!x.ok) {
    bool good = cudaStreamCreateWithFlags(&x.s, cudaStreamNonBlocking) == cudaSuccess;
    good = good && cudaEventCreateWithFlags(&x.fork, cudaEventDisableTiming) == cudaSuccess;
    good = good && cudaEventCreateWithFlags(&x.join, cudaEventDisableTiming) == cudaSuccess;
    x.ok = good;
  }
  return x;
}
cudaStream_t bg_fork(cudaStream_t st) {
  BgStream& x = bg_stream();
  if (!x.ok || cudaEventRecord(x.fork, st) != cudaSuccess || cudaStreamWaitEvent(x.s, x.fork, 0) != cudaSuccess) return st;
  return x.s;
}
int bg_mark(cudaStream_t bg, cudaStream_t st) {
  if (bg == st) return DR4SR_OK;
  BgStream& x = bg_stream();
  if (cudaEventRecord(x.join, bg) != cudaSuccess) {
    set_cuda_error(cudaGetLastError(), "background stream record");
    return DR4SR_ECUDA;
  }
  x.pending = true;
  return DR4SR_OK;
}
int bg_join(cudaStream_t st) {
  BgStream& x = bg_stream();
  if (!x.ok || !x.pending) return DR4SR_OK;
  x.pending = false;
  if (cudaStreamWaitEvent(st, x.join, 0) != cudaSuccess) {
    set_cuda_error(cudaGetLastError(), "background stream join");
    return DR4SR_ECUDA;
  }
  return DR4SR_OK;
}
}  // namespace dr4sr

using namespace dr4sr;

extern "C" int dr4sr_abi_version(void) { return DR4SR_ABI_VERSION; }
extern "C" int dr4sr_enable_peer_access(int peer_device) {
  int cur = 0, can = 0;
  if (cudaGetDevice(&cur) != cudaSuccess) { set_cuda_error(cudaGetLastError(), "peer access"); return DR4SR_ECUDA; }
  if (peer_device == cur) return DR4SR_OK;
  if (cudaDeviceCanAccessPeer(&can, cur, peer_device) != cudaSuccess || !can) { cudaGetLastError(); return DR4SR_EINVAL; }
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_cuda_error(e, "cudaDeviceEnablePeerAccess"); return DR4SR_ECUDA; }
  cudaGetLastError();
  return DR4SR_OK;
}
extern "C" const char* dr4sr_last_cuda_error(void) { return g_err; }
extern "C" int dr4sr_set_gemm_backend(int backend) {
  if (backend != 0 && backend != 1) return DR4SR_EINVAL;
  g_gemm_backend.store(backend);
  return DR4SR_OK;
}
extern "C" int dr4sr_set_attn_backend(int backend) {
  if (backend < 0 || backend > 2) return DR4SR_EINVAL;
  g_attn_backend.store(backend);
  return DR4SR_OK;
}
extern "C" int dr4sr_set_fused_backend(int backend) {
  if (backend < 0 || backend > 2) return DR4SR_EINVAL;
  g_fused_backend.store(backend);
  return DR4SR_OK;
}
extern "C" int dr4sr_debug_trace(int* host_mapped) {
  const int rc = fused_fwd_set_trace(host_mapped);
  return rc != DR4SR_OK ? rc : attn_bwd_set_trace(host_mapped);
}
extern "C" long long dr4sr_launch_count(void) { return g_launches.load(); }

extern "C" int dr4sr_prof_enable(int on) {
  g_prof_on.store(on ? 1 : 0);
  return DR4SR_OK;
}

// Waits for the recorded events, writes "name,launches,total_ms\n" lines (sorted by total time) into buf and
// clears the records.  Returns the number of bytes written (0 if nothing was recorded).
extern "C" size_t dr4sr_prof_collect(char* buf, size_t cap) {
  std::vector<ProfRec> recs;
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    recs.swap(g_prof);
  }
  std::map<std::string, std::pair<long long, double>> agg;
  for (auto& r : recs) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.end) == cudaSuccess && cudaEventElapsedTime(&ms, r.beg, r.end) == cudaSuccess) {
      auto& a = agg[r.name];
      a.first += 1; a.second += ms;
    }
    cudaEventDestroy(r.beg); cudaEventDestroy(r.end);
  }
  std::vector<std::pair<std::string, std::pair<long long, double>>> rows(agg.begin(), agg.end());
  std::sort(rows.begin(), rows.end(), [](auto& a, auto& b) { return a.second.second > b.second.second; });
  size_t off = 0;
  for (auto& r : rows) {
    char line[160];
    int n = snprintf(line, sizeof(line), "%s,%lld,%.6f\n", r.first.c_str(), r.second.first, r.second.second);
    if (n < 0 || off + (size_t)n >= cap) break;
    memcpy(buf + off, line, (size_t)n);
    off += (size_t)n;
  }
  if (cap) buf[off < cap ? off : cap - 1] = 0;
  return off;
}
