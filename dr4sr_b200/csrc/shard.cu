// shard.cu -- device side of the row-sharded item table (multi-GPU, no reference counterpart:
// the reference is single-device).
//
// Rank r owns rows [lo_r, hi_r) of E with their Adam state.  Per step a rank needs three row sets for
// its local sequences (input ids, positive targets, negatives).  They are fetched with ONE all-to-all
// of row ids and ONE all-to-all of rows (8x less traffic than all-gathering activations), staged in a
// small local table `loc` whose row 0 is the zero pad row, and every kernel of the single-GPU path then
// runs unchanged on (loc, remapped ids).  The backward sends the rows of the local gradient table back
// the same way and the owners scatter-add them into their shard.
//
//   dr4sr_shard_plan        requests of the live slots -> bucketed by owner, ids remapped to loc rows
//   dr4sr_gather_rows       owner side: rows of the shard for the received ids
//   dr4sr_scatter_add_rows  owner side: received gradient rows += into the shard's gradient
#include "internal.cuh"

namespace dr4sr {
namespace {

__device__ __forceinline__ int owner_of_row(int64_t id, int64_t base, int64_t extra) {
  const int64_t cut = extra * (base + 1);
  return id < cut ? (int)(id / (base + 1)) : (int)(extra + (id - cut) / (base > 0 ? base : 1));
}

// request j = 3 * row + k of packed live row `row`: k = 0 input id, 1 positive target, 2 negative
// (targets only where item_id != 0); id 0 (pad) is never requested.
__device__ __forceinline__ int64_t request_id(int j, const int64_t* in_ids, const int64_t* item_id, const int64_t* neg_item,
                                              const int32_t* tok_off, const int32_t* row_seq, int L) {
  const int row = j / 3, k = j % 3;
  const int b = row_seq[row];
  const size_t slot = (size_t)b * L + (row - tok_off[b]);
  if (k == 0) return in_ids[slot];
  const int64_t pid = item_id ? item_id[slot] : 0;
  if (pid == 0) return 0;
  return k == 1 ? pid : neg_item[slot];
}

__global__ void __launch_bounds__(256) plan_count_kernel(const int64_t* __restrict__ in_ids, const int64_t* __restrict__ item_id,
                                                         const int64_t* __restrict__ neg_item, const int32_t* __restrict__ tok_off,
                                                         const int32_t* __restrict__ row_seq, const int32_t* __restrict__ counts, int L,
                                                         int64_t base, int64_t extra, int world, int32_t* __restrict__ send_counts) {
  __shared__ int hist[64];
  if (threadIdx.x < 64) hist[threadIdx.x] = 0;
  __syncthreads();
  const int R = 3 * counts[0];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < R; j += gridDim.x * blockDim.x) {
    const int64_t id = request_id(j, in_ids, item_id, neg_item, tok_off, row_seq, L);
    if (id != 0) atomicAdd(&hist[owner_of_row(id, base, extra)], 1);
  }
  __syncthreads();
  if (threadIdx.x < world && hist[threadIdx.x]) atomicAdd(&send_counts[threadIdx.x], hist[threadIdx.x]);
}

__global__ void plan_offsets_kernel(const int32_t* __restrict__ send_counts, int world, int32_t* __restrict__ cursor,
                                    int32_t* __restrict__ total) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int s = 0;
    for (int r = 0; r < world; ++r) { cursor[r] = s; s += send_counts[r]; }
    total[0] = s;
  }
}

__global__ void __launch_bounds__(256) plan_fill_kernel(const int64_t* __restrict__ in_ids, const int64_t* __restrict__ item_id,
                                                        const int64_t* __restrict__ neg_item, const int32_t* __restrict__ tok_off,
                                                        const int32_t* __restrict__ row_seq, const int32_t* __restrict__ counts, int L,
                                                        int64_t base, int64_t extra, int32_t* __restrict__ cursor,
                                                        int64_t* __restrict__ send_ids, int64_t* __restrict__ in_loc,
                                                        int64_t* __restrict__ item_loc, int64_t* __restrict__ neg_loc) {
  const int R = 3 * counts[0];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < R; j += gridDim.x * blockDim.x) {
    const int row = j / 3, k = j % 3;
    const int b = row_seq[row];
    const size_t slot = (size_t)b * L + (row - tok_off[b]);
    const int64_t id = request_id(j, in_ids, item_id, neg_item, tok_off, row_seq, L);
    int64_t loc = 0;
    if (id != 0) {
      const int pos = atomicAdd(&cursor[owner_of_row(id, base, extra)], 1);
      send_ids[pos] = id;
      loc = 1 + pos;                                   // row of the staged local table (row 0 = pad)
    }
    if (k == 0) in_loc[slot] = loc;
    else if (k == 1) { if (item_loc) item_loc[slot] = loc; }
    else { if (neg_loc) neg_loc[slot] = loc; }
  }
}

__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, const int64_t* __restrict__ ids, int64_t lo,
                                                          int64_t m, int D, float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < m; r += (int64_t)gridDim.x * 8) {
    const float* s = src + (size_t)(ids[r] - lo) * D;
    for (int c = lane * 4; c < D; c += 128) *reinterpret_cast<float4*>(out + (size_t)r * D + c) = *reinterpret_cast<const float4*>(s + c);
  }
}

__global__ void __launch_bounds__(256) scatter_add_rows_kernel(float* __restrict__ dst, const int64_t* __restrict__ ids, int64_t lo,
                                                               int64_t m, int D, const float* __restrict__ rows) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < m; r += (int64_t)gridDim.x * 8) {
    float* d = dst + (size_t)(ids[r] - lo) * D;
    for (int c = lane * 4; c < D; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(rows + (size_t)r * D + c);
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + c), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
  }
}

}  // namespace
}  // namespace dr4sr

using namespace dr4sr;

extern "C" int dr4sr_shard_plan(const int64_t* in_item_id, const int64_t* item_id, const int64_t* neg_item, const int32_t* tok_off,
                                const int32_t* row_seq, const int32_t* counts, int32_t B, int32_t L, int64_t num_rows, int32_t world,
                                int32_t* send_counts, int32_t* scratch, int64_t* send_ids, int64_t* in_loc, int64_t* item_loc,
                                int64_t* neg_loc, dr4sr_stream_t stream) {
  if (!in_item_id || !tok_off || !row_seq || !counts || !send_counts || !scratch || !send_ids || !in_loc) return DR4SR_EINVAL;
  if (world < 1 || world > 64 || num_rows < world) return DR4SR_EINVAL;
  if (item_id && (!neg_item || !item_loc || !neg_loc)) return DR4SR_EINVAL;
  cudaStream_t st = as_stream(stream);
  const int64_t base = num_rows / world, extra = num_rows % world;
  const int R_cap = 3 * B * L;
  const int blocks = ceil_div(R_cap, 256) < 4 * kNumSMs ? ceil_div(R_cap, 256) : 4 * kNumSMs;
  ProfScope prof("shard_plan", st);
  cudaMemsetAsync(send_counts, 0, sizeof(int32_t) * world, st);
  cudaMemsetAsync(in_loc, 0, sizeof(int64_t) * (size_t)B * L, st);
  if (item_loc) cudaMemsetAsync(item_loc, 0, sizeof(int64_t) * (size_t)B * L, st);
  if (neg_loc) cudaMemsetAsync(neg_loc, 0, sizeof(int64_t) * (size_t)B * L, st);
  plan_count_kernel<<<blocks, 256, 0, st>>>(in_item_id, item_id, neg_item, tok_off, row_seq, counts, L, base, extra, world, send_counts);
  DR4SR_LAUNCH_CHECK("plan_count_kernel");
  plan_offsets_kernel<<<1, 32, 0, st>>>(send_counts, world, scratch, scratch + world);
  DR4SR_LAUNCH_CHECK("plan_offsets_kernel");
  plan_fill_kernel<<<blocks, 256, 0, st>>>(in_item_id, item_id, neg_item, tok_off, row_seq, counts, L, base, extra, scratch, send_ids,
                                           in_loc, item_loc, neg_loc);
  DR4SR_LAUNCH_CHECK("plan_fill_kernel");
  return DR4SR_OK;
}

extern "C" int dr4sr_gather_rows(const float* src, const int64_t* ids, int64_t lo, int64_t m, int32_t D, float* out,
                                 dr4sr_stream_t stream) {
  if (!src || !out || D % 4 || m < 0) return DR4SR_EINVAL;
  if (m == 0) return DR4SR_OK;
  if (!ids) return DR4SR_EINVAL;
  const int blocks = ceil_div(m, 8) < 8 * kNumSMs ? ceil_div(m, 8) : 8 * kNumSMs;
  ProfScope prof("shard_gather_rows", as_stream(stream));
  gather_rows_kernel<<<blocks, 256, 0, as_stream(stream)>>>(src, ids, lo, m, D, out);
  DR4SR_LAUNCH_CHECK("gather_rows_kernel");
  return DR4SR_OK;
}

extern "C" int dr4sr_scatter_add_rows(float* dst, const int64_t* ids, int64_t lo, int64_t m, int32_t D, const float* rows,
                                      dr4sr_stream_t stream) {
  if (!dst || !rows || D % 4 || m < 0) return DR4SR_EINVAL;
  if (m == 0) return DR4SR_OK;
  if (!ids) return DR4SR_EINVAL;
  const int blocks = ceil_div(m, 8) < 8 * kNumSMs ? ceil_div(m, 8) : 8 * kNumSMs;
  ProfScope prof("shard_scatter_add_rows", as_stream(stream));
  scatter_add_rows_kernel<<<blocks, 256, 0, as_stream(stream)>>>(dst, ids, lo, m, D, rows);
  DR4SR_LAUNCH_CHECK("scatter_add_rows_kernel");
  return DR4SR_OK;
}
