// table_grad_sorted.cu -- deterministic embedding-gradient reduction: sort the (item id, gradient-row source) entries of the
// step, then reduce every run of equal ids with warps in a FIXED order (no float atomics).  Same result contract as
// dr4sr_table_grad (loss_table.cu): dE[in_id] += dx0 ; dE[item_id] += ds+ q ; dE[neg] += ds- q, padding id 0 skipped
// (reference: nn.Embedding's backward under model/basemodel.py:204-214; torch's own CUDA path sorts as well).
//
//   entries   e = type * Tcap + packed_row, type 0 = input row (dx0), 1 = target (ds+ q), 2 = negative (ds- q), Tcap = B L
//   keys      item id of the entry, or the sentinel N for entries that carry no gradient (pad rows, id 0, slots without target)
//   sort      stable LSB radix sort of (key, e) over ceil(log2(N+1)) bits -- cub::DeviceRadixSort (library code: a plain sort);
//             stability keeps the entries of one id in increasing e, which fixes the summation order
//   reduce    one warp per chunk of kChunk sorted entries, lanes over the D columns (float4 each): runs that lie inside the
//             chunk are added to the table-gradient row by their only writer; a run that crosses a chunk boundary leaves a
//             "head" (began earlier) / "tail" (continues) partial row per chunk
//   merge     one warp per tail: tail + heads of the following chunks in chunk order -> the row.  Hot ids (Zipf catalogs)
//             therefore cost ceil(run / kChunk) row adds in the merge instead of `run` serialized atomics on one L2 line.
//
// Every row of dE has exactly one writer per launch sequence and every sum has a fixed shape, so two runs on the same inputs
// give bit-identical gradients (tests/test_cuda_table_grad.py).  HBM-bound: reads 3 U of source rows once (U = T D 4 bytes).
#include <cub/device/device_radix_sort.cuh>

#include "internal.cuh"

namespace dr4sr {
namespace {

constexpr int kChunk = 16;                  // sorted entries per warp

__global__ void __launch_bounds__(256) tg_keys_kernel(const int64_t* __restrict__ in_ids, const int64_t* __restrict__ item_id,
                                                      const int64_t* __restrict__ neg_item, const int32_t* __restrict__ tok_off,
                                                      const int32_t* __restrict__ row_seq, const int32_t* __restrict__ counts, int L,
                                                      int Tcap, int with_in, int sentinel, int32_t* __restrict__ keys,
                                                      int32_t* __restrict__ vals) {
  const int T = counts[0];
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 3 * Tcap; e += gridDim.x * blockDim.x) {
    const int type = e / Tcap, row = e - type * Tcap;
    int key = sentinel;
    if (row < T) {
      const int b = row_seq[row];
      const size_t slot = (size_t)b * L + (row - tok_off[b]);
      if (type == 0) {
        const int64_t id = with_in ? in_ids[slot] : 0;
        if (id != 0) key = (int)id;
      } else if (item_id) {
        const int64_t pid = item_id[slot];
        if (pid != 0) key = type == 1 ? (int)pid : (int)neg_item[slot];
      }
    }
    keys[e] = key;
    vals[e] = e;
  }
}

struct ChunkMeta { int head_key, head_ends, tail_key, pad; };

__device__ __forceinline__ void row_add(float* __restrict__ dst, const float4* acc, int D, int lane) {
  for (int j = 0, c = lane * 4; c < D; ++j, c += 128) {
    float4 o = *reinterpret_cast<float4*>(dst + c);
    o.x += acc[j].x; o.y += acc[j].y; o.z += acc[j].z; o.w += acc[j].w;
    *reinterpret_cast<float4*>(dst + c) = o;
  }
}
__device__ __forceinline__ void row_store(float* __restrict__ dst, const float4* acc, int D, int lane) {
  for (int j = 0, c = lane * 4; c < D; ++j, c += 128) *reinterpret_cast<float4*>(dst + c) = acc[j];
}

__global__ void __launch_bounds__(256) tg_segment_kernel(const int32_t* __restrict__ keys, const int32_t* __restrict__ vals,
                                                         const float* __restrict__ dx0, const float* __restrict__ q,
                                                         const float* __restrict__ dscore, int n, int Tcap, int D, int sentinel,
                                                         float* __restrict__ grad, float* __restrict__ partial,
                                                         ChunkMeta* __restrict__ meta, int n_chunks) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= n_chunks) return;
  const int start = w * kChunk, end = min(n, start + kChunk);
  ChunkMeta m{-1, 0, -1, 0};
  const int key_prev = start > 0 ? keys[start - 1] : -1;
  float4 acc[2];
  acc[0] = acc[1] = make_float4(0.f, 0.f, 0.f, 0.f);
  int seg_first = start;                                     // first entry (inside this chunk) of the run being accumulated
  for (int i = start; i < end; ++i) {
    const int k = keys[i];
    if (k == sentinel) break;                                // sorted: nothing but gradient-free entries from here on
    const int e = vals[i];
    const int type = e / Tcap, row = e - type * Tcap;
    const float* src = (type == 0 ? dx0 : q) + (size_t)row * D;
    const float s = type == 0 ? 1.f : dscore[2 * (size_t)row + (type - 1)];
    for (int j = 0, c = lane * 4; c < D; ++j, c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(src + c);
      acc[j].x = fmaf(s, v.x, acc[j].x); acc[j].y = fmaf(s, v.y, acc[j].y);
      acc[j].z = fmaf(s, v.z, acc[j].z); acc[j].w = fmaf(s, v.w, acc[j].w);
    }
    const int k_next = i + 1 < n ? keys[i + 1] : -2;
    const bool run_ends = k_next != k;                        // in the sorted order, not just in this chunk
    if (run_ends || i + 1 == end) {
      const bool starts_here = seg_first > start || key_prev != k;
      if (starts_here && run_ends) {
        row_add(grad + (size_t)k * D, acc, D, lane);          // the run lies inside the chunk: this warp is the row's only writer
      } else if (!starts_here) {                              // began in an earlier chunk: head partial (may be the whole chunk)
        row_store(partial + ((size_t)2 * w) * D, acc, D, lane);
        m.head_key = k;
        m.head_ends = run_ends ? 1 : 0;
      } else {                                                // begins here, continues in the next chunk: tail partial
        row_store(partial + ((size_t)2 * w + 1) * D, acc, D, lane);
        m.tail_key = k;
      }
      acc[0] = acc[1] = make_float4(0.f, 0.f, 0.f, 0.f);
      seg_first = i + 1;
    }
  }
  if (lane == 0) meta[w] = m;
}

__global__ void __launch_bounds__(256) tg_merge_kernel(const ChunkMeta* __restrict__ meta, const float* __restrict__ partial, int D,
                                                       float* __restrict__ grad, int n_chunks) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= n_chunks) return;
  const int key = meta[w].tail_key;
  if (key < 0) return;
  int last = w + 1;                                           // the run's heads: chunks w+1 .. last (flags first, rows after)
  while (last < n_chunks - 1 && !meta[last].head_ends) ++last;
  float4 acc[2];
  for (int j = 0, c = lane * 4; c < D; ++j, c += 128) acc[j] = *reinterpret_cast<const float4*>(partial + ((size_t)2 * w + 1) * D + c);
#pragma unroll 4
  for (int ch = w + 1; ch <= last; ++ch) {
    if (ch >= n_chunks || meta[ch].head_key != key) break;    // (defensive: a tail is always followed by its head)
    for (int j = 0, c = lane * 4; c < D; ++j, c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(partial + ((size_t)2 * ch) * D + c);
      acc[j].x += v.x; acc[j].y += v.y; acc[j].z += v.z; acc[j].w += v.w;
    }
  }
  row_add(grad + (size_t)key * D, acc, D, lane);
}

inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
struct SortedPlan { int n, n_chunks, bits; size_t off_keys, off_vals, off_keys2, off_vals2, off_meta, off_partial, off_cub, cub_bytes, off_pos, bytes; };
SortedPlan make_plan(int B, int L, int D, int64_t N) {
  SortedPlan p{};
  p.n = 3 * B * L;
  p.n_chunks = (p.n + kChunk - 1) / kChunk;
  p.bits = 1;
  while (((int64_t)1 << p.bits) <= N) ++p.bits;              // keys are 0 .. N (N = the sentinel)
  size_t o = 0;
  p.off_keys = o; o = al256(o + sizeof(int32_t) * p.n);
  p.off_vals = o; o = al256(o + sizeof(int32_t) * p.n);
  p.off_keys2 = o; o = al256(o + sizeof(int32_t) * p.n);
  p.off_vals2 = o; o = al256(o + sizeof(int32_t) * p.n);
  p.off_meta = o; o = al256(o + sizeof(ChunkMeta) * p.n_chunks);
  p.off_partial = o; o = al256(o + sizeof(float) * (size_t)2 * p.n_chunks * D);
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, p.n, 0, p.bits, (cudaStream_t)0);
  p.cub_bytes = cub_bytes;
  p.off_cub = o; o = al256(o + cub_bytes);
  p.off_pos = o; o = al256(o + dr4sr_table_grad_workspace_bytes(L, D));
  p.bytes = o;
  return p;
}

}  // namespace
}  // namespace dr4sr

using namespace dr4sr;

extern "C" size_t dr4sr_table_grad_sorted_workspace_bytes(int32_t B, int32_t L, int32_t D, int64_t N) {
  if (B <= 0 || L <= 0 || D <= 0 || N <= 0) return 0;
  return make_plan(B, L, D, N).bytes;
}

extern "C" int dr4sr_table_grad_sorted(const float* dx0_packed, const float* q_packed, const float* dscore,
                                       const int64_t* in_item_id, const int64_t* item_id, const int64_t* neg_item,
                                       const int32_t* tok_off, const int32_t* row_seq, const int32_t* counts, int32_t B, int32_t L,
                                       int32_t D, int64_t N, float* table_grad, float* pos_grad, void* ws, size_t ws_bytes,
                                       dr4sr_stream_t stream) {
  if (!in_item_id || !tok_off || !row_seq || !counts || !table_grad || !ws || D % 4 || D > 256 || N <= 0 || N >= 0x7fffffff)
    return DR4SR_EINVAL;
  if (item_id && (!q_packed || !dscore || !neg_item)) return DR4SR_EINVAL;
  if ((int64_t)3 * B * L >= 0x7fffffff) return DR4SR_EINVAL;
  const SortedPlan p = make_plan(B, L, D, N);
  if (ws_bytes < p.bytes) return DR4SR_EWORKSPACE;
  cudaStream_t st = as_stream(stream);
  char* base = reinterpret_cast<char*>(ws);
  int32_t* keys = reinterpret_cast<int32_t*>(base + p.off_keys);
  int32_t* vals = reinterpret_cast<int32_t*>(base + p.off_vals);
  int32_t* keys2 = reinterpret_cast<int32_t*>(base + p.off_keys2);
  int32_t* vals2 = reinterpret_cast<int32_t*>(base + p.off_vals2);
  ChunkMeta* meta = reinterpret_cast<ChunkMeta*>(base + p.off_meta);
  float* partial = reinterpret_cast<float*>(base + p.off_partial);
  // the positional gradient: the deterministic column reduction of the one-call path, on the auxiliary stream
  cudaStream_t sa = st;
  if (pos_grad && dx0_packed) {
    sa = aux_fork(st);
    DR4SR_TRY(launch_pos_grad(dx0_packed, tok_off, B, L, D, pos_grad, base + p.off_pos, sa));
  }
  const int Tcap = B * L;
  {
    ProfScope prof("table_grad_keys", st);
    const int blocks = ceil_div(p.n, 256) < 4 * kNumSMs ? ceil_div(p.n, 256) : 4 * kNumSMs;
    tg_keys_kernel<<<blocks, 256, 0, st>>>(in_item_id, item_id, neg_item, tok_off, row_seq, counts, L, Tcap, dx0_packed ? 1 : 0, (int)N,
                                          keys, vals);
    DR4SR_LAUNCH_CHECK("tg_keys_kernel");
  }
  {
    ProfScope prof("table_grad_sort", st);
    size_t cub_bytes = p.cub_bytes;
    if (cub::DeviceRadixSort::SortPairs(base + p.off_cub, cub_bytes, keys, keys2, vals, vals2, p.n, 0, p.bits, st) != cudaSuccess) {
      set_cuda_error(cudaGetLastError(), "cub radix sort");
      return DR4SR_ECUDA;
    }
  }
  {
    ProfScope prof("table_grad_segments", st);
    tg_segment_kernel<<<ceil_div(p.n_chunks, 8), 256, 0, st>>>(keys2, vals2, dx0_packed, q_packed, dscore, p.n, Tcap, D, (int)N, table_grad,
                                                             partial, meta, p.n_chunks);
    DR4SR_LAUNCH_CHECK("tg_segment_kernel");
    tg_merge_kernel<<<ceil_div(p.n_chunks, 8), 256, 0, st>>>(meta, partial, D, table_grad, p.n_chunks);
    DR4SR_LAUNCH_CHECK("tg_merge_kernel");
  }
  if (sa != st) DR4SR_TRY(aux_join(sa, st));
  return DR4SR_OK;
}
