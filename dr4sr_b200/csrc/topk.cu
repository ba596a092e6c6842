// topk.cu -- full-catalog scoring + top-k, replaces BaseModel.topk (reference model/basemodel.py:354-365).
//
//   scores = q @ E[:N].T ; -inf at ids outside the eval domain and at the user's history ; top-k.
// Default path (D in {64, 128}, k <= 128): logits_tc.cu -- tcgen05 scoring with the selection fused into the epilogue,
// exact fp32 re-scoring of the survivors; no B x N buffer.  This file keeps the exact-fp32 reference path of the library
// (dr4sr_set_gemm_backend(1), other shapes): FFMA GEMM with the domain mask folded in as a column bias (0 / -inf),
// a scatter of -inf at the history ids, then one CTA per row doing an 8-bit MSD radix select for the
// k-th key followed by an index-ordered gather and a bitonic sort of the k survivors.
// Ties are broken by the lower item id (deterministic; torch.topk leaves tie order unspecified).
#include <atomic>
#include "gemm_simt.cuh"
#include "internal.cuh"
namespace dr4sr { extern std::atomic<int> g_gemm_backend; }

namespace dr4sr {
namespace {

// column bias of the scoring GEMM: 0 for live items, -inf for dead ones and for the row padding
__global__ void __launch_bounds__(256) dead_bias_kernel(const uint8_t* __restrict__ dead, int64_t N, int64_t n_al,
                                                        float* __restrict__ bias) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_al; i += (int64_t)gridDim.x * blockDim.x)
    bias[i] = (i >= N || (dead && dead[i])) ? -INFINITY : 0.0f;
}

__global__ void __launch_bounds__(256) mask_hist_kernel(const int64_t* __restrict__ hist, int B, int H, int64_t N, int64_t ld,
                                                        float* __restrict__ scores) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  const int64_t id = hist[i];
  if (id >= 0 && id < N) scores[(size_t)(i / H) * ld + id] = -INFINITY;
}

// monotone map float -> uint32 (larger float => larger key); -inf -> smallest finite-class key
__device__ __forceinline__ uint32_t order_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

constexpr int kSelThreads = 512;

__global__ void __launch_bounds__(kSelThreads) topk_select_kernel(const float* __restrict__ scores, int64_t N, int64_t ld, int k,
                                                                  int kpad, float* __restrict__ out_scores,
                                                                  int64_t* __restrict__ out_ids) {
  extern __shared__ uint32_t sm_u[];
  uint32_t* hist = sm_u;                       // [256]
  uint32_t* ck = hist + 256;                   // [kpad] candidate keys
  int32_t* ci = reinterpret_cast<int32_t*>(ck + kpad);   // [kpad] candidate ids
  __shared__ uint32_t s_prefix, s_mask;
  __shared__ int s_need, s_base, s_tiebase, s_warp[kSelThreads / 32][2];
  const float* row = scores + (size_t)blockIdx.x * ld;
  const int tid = threadIdx.x;

  // ---- radix select: find key K such that count(key > K) < k <= count(key >= K) ----
  if (tid == 0) { s_prefix = 0u; s_mask = 0u; s_need = k; }
  __syncthreads();
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = tid; i < 256; i += kSelThreads) hist[i] = 0u;
    __syncthreads();
    const uint32_t prefix = s_prefix, mask = s_mask;
    for (int64_t start = 0; start < N; start += kSelThreads) {   // uniform trip count: match_any needs full warps
      const int64_t i = start + tid;
      bool live = false;
      uint32_t digit = 0x100u | (uint32_t)(tid & 31);              // unique per lane => never aggregated
      if (i < N) {
        const uint32_t key = order_key(row[i]);
        if ((key & mask) == prefix) { live = true; digit = (key >> shift) & 255u; }
      }
      const uint32_t peers = __match_any_sync(0xffffffffu, digit);  // warp-aggregated histogram update
      if (live && (tid & 31) == __ffs(peers) - 1) atomicAdd(&hist[digit], (uint32_t)__popc(peers));
    }
    __syncthreads();
    if (tid == 0) {
      int need = s_need;
      int d = 255;
      for (; d > 0; --d) {
        const int c = (int)hist[d];
        if (c >= need) break;
        need -= c;
      }
      s_need = need;   // how many of the selected digit's bucket still belong to the top-k
      s_prefix = prefix | ((uint32_t)d << shift);
      s_mask = mask | (255u << shift);
    }
    __syncthreads();
  }
  const uint32_t kth = s_prefix;
  const int n_tie = s_need;              // elements equal to kth that are kept (lowest ids first)
  const int n_gt = k - n_tie;

  // ---- ordered gather: keys > kth to slots [0, n_gt), first n_tie keys == kth to [n_gt, k) ----
  if (tid == 0) { s_base = 0; s_tiebase = 0; }
  for (int i = tid; i < kpad; i += kSelThreads) { ck[i] = 0u; ci[i] = 0x7fffffff; }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  for (int64_t start = 0; start < N; start += kSelThreads) {
    const int64_t i = start + tid;
    uint32_t key = 0u;
    bool gt = false, eq = false;
    if (i < N) { key = order_key(row[i]); gt = key > kth; eq = key == kth; }
    const uint32_t bg = __ballot_sync(0xffffffffu, gt), be = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) { s_warp[warp][0] = __popc(bg); s_warp[warp][1] = __popc(be); }
    __syncthreads();
    int og = s_base, oe = s_tiebase;
    for (int wv = 0; wv < warp; ++wv) { og += s_warp[wv][0]; oe += s_warp[wv][1]; }
    og += __popc(bg & ((1u << lane) - 1u));
    oe += __popc(be & ((1u << lane) - 1u));
    if (gt) { ck[og] = key; ci[og] = (int32_t)i; }
    if (eq && oe < n_tie) { ck[n_gt + oe] = key; ci[n_gt + oe] = (int32_t)i; }
    __syncthreads();
    if (tid == 0) {
      int tg = 0, te = 0;
      for (int wv = 0; wv < kSelThreads / 32; ++wv) { tg += s_warp[wv][0]; te += s_warp[wv][1]; }
      s_base += tg; s_tiebase += te;
    }
    __syncthreads();
  }

  // ---- bitonic sort of kpad (key desc, id asc); padding entries (key 0, id max) sink to the end ----
  for (int size = 2; size <= kpad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < kpad; i += kSelThreads) {
        const int j = i ^ stride;
        if (j > i) {
          const bool up = (i & size) == 0;   // 'up' blocks hold the better elements first
          const uint32_t ki = ck[i], kj = ck[j];
          const int32_t ii = ci[i], ij = ci[j];
          const bool i_better = ki > kj || (ki == kj && ii < ij);
          if (up ? !i_better : i_better) { ck[i] = kj; ck[j] = ki; ci[i] = ij; ci[j] = ii; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < k; i += kSelThreads) {
    const int32_t id = ci[i];
    out_ids[(size_t)blockIdx.x * k + i] = id;
    out_scores[(size_t)blockIdx.x * k + i] = row[id];
  }
}

}  // namespace
}  // namespace dr4sr

using namespace dr4sr;

static bool topk_use_tc(int32_t D, int32_t k) { return g_gemm_backend.load(std::memory_order_relaxed) == 0 && logits_tc_supported(D, k); }

extern "C" size_t dr4sr_topk_workspace_bytes(int32_t B, int64_t N, int32_t k) {
  // tcgen05 path: bf16 hi/lo images of the table + per-row candidate lists -- O(N D + B), independent of B x N.  The
  // signature carries neither D nor H: sized for D = 128 and H = 64 (the reference's max_seq_len is 50).
  if (topk_use_tc(128, k)) return logits_tc_workspace_bytes(B, N, 128, 64);
  const size_t n_al = ((size_t)N + 127) & ~(size_t)127;
  return sizeof(float) * ((size_t)B * n_al + n_al) + 256;      // exact-fp32 FFMA path: materialised logits
}

extern "C" int dr4sr_topk(const float* q, const float* table, const uint8_t* item_dead, const int64_t* user_hist, int32_t B,
                          int32_t D, int64_t N, int32_t H, int32_t k, float* out_scores, int64_t* out_ids, void* ws,
                          size_t ws_bytes, dr4sr_stream_t stream) {
  if (!q || !table || !out_scores || !out_ids || !ws || B <= 0 || D % 4 || N <= 0 || N > 0x7fffff00) return DR4SR_EINVAL;
  if (k <= 0 || k > 1024 || k > N) return DR4SR_EINVAL;
  cudaStream_t st = as_stream(stream);
  if (topk_use_tc(D, k) && H <= 64)
    return launch_logits_topk_tc(q, table, item_dead, user_hist, B, D, N, H, k, out_scores, out_ids, ws, ws_bytes, st);
  {
    const size_t n_al0 = ((size_t)N + 127) & ~(size_t)127;
    if (ws_bytes < sizeof(float) * ((size_t)B * n_al0 + n_al0) + 256) return DR4SR_EWORKSPACE;
  }
  const size_t n_al = ((size_t)N + 127) & ~(size_t)127;   // padded row stride: 16-byte stores, whole tiles
  float* bias = reinterpret_cast<float*>(ws);
  float* scores = bias + n_al;
  {
  ProfScope prof("topk_dead_bias", st);
  dead_bias_kernel<<<ceil_div(n_al, 256) < 4 * kNumSMs ? ceil_div(n_al, 256) : 4 * kNumSMs, 256, 0, st>>>(item_dead, N, (int64_t)n_al, bias);
  DR4SR_LAUNCH_CHECK("dead_bias_kernel");
  }
  {
    GemmArgs g = gemm_args(q, D, table, D, scores, (int)n_al, B, (int)n_al, D, nullptr);
    g.bias = bias; g.b_rows = (int)N; g.tag = "topk_logits_gemm";
    DR4SR_TRY((launch_gemm<128, 128, true, true, false>(g, st)));
  }
  if (user_hist && H > 0) {
    ProfScope prof("topk_mask_hist", st);
    mask_hist_kernel<<<ceil_div((int64_t)B * H, 256), 256, 0, st>>>(user_hist, B, H, N, (int64_t)n_al, scores);
    DR4SR_LAUNCH_CHECK("mask_hist_kernel");
  }
  int kpad = 32;
  while (kpad < k) kpad <<= 1;
  const size_t smem = sizeof(uint32_t) * (256 + 2 * (size_t)kpad);
  ProfScope prof("topk_select", st);
  topk_select_kernel<<<B, kSelThreads, smem, st>>>(scores, N, (int64_t)n_al, k, kpad, out_scores, out_ids);
  DR4SR_LAUNCH_CHECK("topk_select_kernel");
  return DR4SR_OK;
}

// ---- ranking metrics of an eval batch, accumulated on the device (reference evaluation/__init__.py:9-36,107-134 through
// model/basemodel.py:337-352: one relevant item per user) ---------------------------------------------------------------
// hit position r of `target` in the user's top-k ids -> ndcg@c += 1 / log2(r + 2), recall@c += 1 for every cutoff c > r.
// One thread per user; block partials in double, one atomicAdd(double) per block and metric: the epoch's sums stay on the
// device (no host read per batch) and are read once at the end.
namespace dr4sr {
struct Cutoffs { int32_t c[8]; int32_t n; };
__global__ void __launch_bounds__(256) rank_metrics_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ target, int B, int k,
                                                           const Cutoffs cut, double* __restrict__ sums) {
  __shared__ double s_part[16][8];                            // [2 * n_cut][warps]
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  int r = -1;
  if (u < B) {
    const int64_t t = target[u];
    const int64_t* row = ids + (size_t)u * k;
    for (int j = 0; j < k; ++j)
      if (row[j] == t) { r = j; break; }
  }
  const float gain = r >= 0 ? 1.0f / log2f((float)r + 2.0f) : 0.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = 0; i < cut.n; ++i) {
    const bool in = r >= 0 && r < cut.c[i];
    double nd = in ? (double)gain : 0.0, rc = in ? 1.0 : 0.0;
    for (int o = 16; o > 0; o >>= 1) {
      nd += __shfl_xor_sync(0xffffffffu, nd, o);
      rc += __shfl_xor_sync(0xffffffffu, rc, o);
    }
    if (lane == 0) { s_part[2 * i][warp] = nd; s_part[2 * i + 1][warp] = rc; }
  }
  __syncthreads();
  if (threadIdx.x < 2 * cut.n) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_part[threadIdx.x][w];
    atomicAdd(sums + threadIdx.x, t);
  }
}
}  // namespace dr4sr

extern "C" int dr4sr_rank_metrics(const int64_t* topk_ids, const int64_t* target, int32_t B, int32_t k, const int32_t* cutoffs,
                                  int32_t n_cutoffs, double* sums, dr4sr_stream_t stream) {
  if (!topk_ids || !target || !cutoffs || !sums || B <= 0 || k <= 0 || n_cutoffs <= 0 || n_cutoffs > 8) return DR4SR_EINVAL;
  Cutoffs cut{};
  cut.n = n_cutoffs;
  for (int i = 0; i < n_cutoffs; ++i) {
    if (cutoffs[i] <= 0 || cutoffs[i] > k) return DR4SR_EINVAL;
    cut.c[i] = cutoffs[i];
  }
  cudaStream_t st = as_stream(stream);
  ProfScope prof("rank_metrics", st);
  rank_metrics_kernel<<<ceil_div(B, 256), 256, 0, st>>>(topk_ids, target, B, k, cut, sums);
  DR4SR_LAUNCH_CHECK("rank_metrics_kernel");
  return DR4SR_OK;
}
