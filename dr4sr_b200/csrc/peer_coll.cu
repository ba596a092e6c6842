// peer_coll.cu -- the two per-step synchronisation points of the peer-sharded table as kernels over peer memory (NVLink /
// NVSwitch loads and stores), instead of NCCL collectives of one float and of 0.8 MB:
//
//   dr4sr_peer_barrier   : stream-ordered barrier over the ranks (+ optional sum of one int32 per rank: the number of valid
//                          targets).  Rank r stores the barrier's epoch into slot r of every peer's flag array (release,
//                          system scope) and waits until every slot of its own array has reached the epoch (acquire).
//   dr4sr_peer_allreduce : the same barrier, then out[i] = sum over ranks of stage_r[i] in RANK ORDER, every rank reading all
//                          staging buffers through its peer mappings -- the result is bit-identical on every rank.
//
// No reference counterpart (the reference is single-device).  Why not NCCL: measured on 8 x B200 every small collective costs
// the step 40-100 us (launch + protocol latency at 8 ranks, and c10d's host path when the loop is host-bound), three of them per
// step; a flag exchange over NVLink is a few microseconds and one kernel launch.  Epochs increase monotonically (host counter,
// identical call sequence on every rank), so a fast rank's next signal never confuses a slow one (the wait is `>= epoch`).
// Every spin is bounded by a wall-clock timeout (a missing rank traps after 60 s instead of hanging the GPU).
#include "internal.cuh"

namespace dr4sr {
namespace {

struct PeerComm {
  int32_t* flags[DR4SR_MAX_SHARDS];
  int32_t* slots[DR4SR_MAX_SHARDS];
  const float* stage[DR4SR_MAX_SHARDS];
  int world, rank;
};

__device__ __forceinline__ void st_release_sys(int32_t* p, int32_t v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(int32_t* p, int32_t v) {
  asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int32_t ld_acquire_sys(const int32_t* p) {
  int32_t v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

constexpr unsigned long long kBarrierTimeoutNs = 60ull * 1000ull * 1000ull * 1000ull;   // 60 s: a rank is missing

// one CTA of 32 threads; thread p talks to rank p
__global__ void __launch_bounds__(32) peer_barrier_kernel(const PeerComm c, int32_t epoch, int32_t* count_inout) {
  __shared__ int32_t s_cnt[DR4SR_MAX_SHARDS];
  const int p = threadIdx.x;
  const int32_t mine = count_inout ? *count_inout : 0;
  if (p < c.world) {
    if (count_inout) st_relaxed_sys(c.slots[p] + c.rank, mine);          // my count into rank p's slot array
    __threadfence_system();                                              // everything this GPU wrote before the barrier ...
    st_release_sys(c.flags[p] + c.rank, epoch);                          // ... is visible before the signal
    const int32_t* my_flag = c.flags[c.rank] + p;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(my_flag) < epoch) {
      if (global_ns() - t0 > kBarrierTimeoutNs) __trap();
      __nanosleep(64);
    }
    s_cnt[p] = count_inout ? ld_acquire_sys(c.slots[c.rank] + p) : 0;    // rank p's count (written before its signal)
  }
  __syncthreads();
  if (count_inout && p == 0) {
    int32_t s = 0;
    for (int r = 0; r < c.world; ++r) s += s_cnt[r];
    *count_inout = s;
  }
}

__global__ void __launch_bounds__(256) peer_sum_kernel(const PeerComm c, int64_t n, float* __restrict__ out) {
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 acc = reinterpret_cast<const float4*>(c.stage[0])[i];
    for (int r = 1; r < c.world; ++r) {                                   // fixed rank order: the same bits on every rank
      const float4 v = reinterpret_cast<const float4*>(c.stage[r])[i];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(out)[i] = acc;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    float acc = c.stage[0][i];
    for (int r = 1; r < c.world; ++r) acc += c.stage[r][i];
    out[i] = acc;
  }
}

bool comm_from(const dr4sr_peer_comm* m, PeerComm* c) {
  if (!m || m->world < 1 || m->world > DR4SR_MAX_SHARDS || m->rank < 0 || m->rank >= m->world) return false;
  for (int r = 0; r < m->world; ++r) {
    if (!m->flags[r] || !m->slots[r]) return false;
    c->flags[r] = m->flags[r]; c->slots[r] = m->slots[r]; c->stage[r] = m->stage[r];
  }
  c->world = m->world; c->rank = m->rank;
  return true;
}

}  // namespace
}  // namespace dr4sr

using namespace dr4sr;

extern "C" int dr4sr_peer_barrier(const dr4sr_peer_comm* comm, int32_t epoch, int32_t* count_inout, dr4sr_stream_t stream) {
  PeerComm c{};
  if (!comm_from(comm, &c) || epoch <= 0) return DR4SR_EINVAL;
  cudaStream_t st = as_stream(stream);
  ProfScope prof("peer_barrier", st);
  peer_barrier_kernel<<<1, 32, 0, st>>>(c, epoch, count_inout);
  DR4SR_LAUNCH_CHECK("peer_barrier_kernel");
  return DR4SR_OK;
}

extern "C" int dr4sr_peer_allreduce(const dr4sr_peer_comm* comm, int32_t epoch, int64_t n, float* out, dr4sr_stream_t stream) {
  PeerComm c{};
  if (!comm_from(comm, &c) || epoch <= 0 || n <= 0 || !out) return DR4SR_EINVAL;
  for (int r = 0; r < c.world; ++r)
    if (!c.stage[r] || (reinterpret_cast<uintptr_t>(c.stage[r]) & 15)) return DR4SR_EINVAL;
  if (reinterpret_cast<uintptr_t>(out) & 15) return DR4SR_EINVAL;
  cudaStream_t st = as_stream(stream);
  {
    ProfScope prof("peer_barrier", st);
    peer_barrier_kernel<<<1, 32, 0, st>>>(c, epoch, nullptr);
    DR4SR_LAUNCH_CHECK("peer_barrier_kernel");
  }
  ProfScope prof("peer_allreduce", st);
  const int blocks = ceil_div(n >> 2, 256) < 2 * kNumSMs ? (ceil_div(n >> 2, 256) > 0 ? ceil_div(n >> 2, 256) : 1) : 2 * kNumSMs;
  peer_sum_kernel<<<blocks, 256, 0, st>>>(c, n, out);
  DR4SR_LAUNCH_CHECK("peer_sum_kernel");
  return DR4SR_OK;
}
