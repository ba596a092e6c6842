// loss_table.cu -- sampled scoring + BCE (forward and backward in one pass), embedding-gradient
// scatter-add, dense Adam.  All HBM-bound: one warp per row, 16-byte lanes, no intermediate [T,D]
// gathers are materialised (the reference writes E[item_id], E[neg_item] and their products to HBM).
#include "internal.cuh"

namespace dr4sr {
namespace {

// One warp per (b, t) slot of the dense [B, L] target grid.
//   t <  len : q = packed row; s+ = <q,E+>, s- = <q,E->  (skipped when item_id == 0)
//   t >= len : the reference's 'origin' pooling has zeroed q, so a non-pad target there contributes
//              (-logsig(0) + softplus(0)) / n and no gradient.
//
// KIND 0: sampled BCE (model/loss_func.py:9-35, the loss training_step uses);  KIND 1: BPR as model/loss_func.py:40-49
// intends it, l = -logsig(s+ - s-) / n with one negative (unreachable through the reference's training_step, which
// passes `reduce` to a two-argument forward -- the "fixed BPR" extension of SURVEY.md Appendix C.6).
template <int KIND>
__global__ void __launch_bounds__(256) score_bce_kernel(const float* __restrict__ q, const __grid_constant__ ShardView table,
                                                        const int64_t* __restrict__ item_id, const int64_t* __restrict__ neg_item,
                                                        const int32_t* __restrict__ tok_off, const int32_t* __restrict__ counts,
                                                        int B, int L, int D, const float* __restrict__ loss_weight,
                                                        const float* __restrict__ upstream, float* __restrict__ loss_pos,
                                                        float* __restrict__ dscore, float* __restrict__ dq) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_valid = counts[1];
  const float inv_n = 1.0f / (float)n_valid;
  const float up = upstream ? *upstream : 1.0f;
  const int total = B * L;
  for (int slot = blockIdx.x * 8 + warp; slot < total; slot += gridDim.x * 8) {
    const int b = slot / L, t = slot % L;
    const int off = tok_off[b], len = tok_off[b + 1] - off;
    const int64_t pid = item_id[slot];
    const bool in_seq = t < len;
    const int row = off + t;
    if (pid == 0) {
      if (lane == 0) loss_pos[slot] = 0.f;
      if (in_seq) {
        if (lane < 2) dscore[2 * (size_t)row + lane] = 0.f;
        if (dq) for (int c = lane * 4; c < D; c += 128) *reinterpret_cast<float4*>(dq + (size_t)row * D + c) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      continue;
    }
    if (!in_seq) {
      if (lane == 0) loss_pos[slot] = (KIND == 0 ? 2.f : 1.f) * 0.69314718055994531f * inv_n;
      continue;
    }
    const int64_t nid = neg_item[slot];
    const float* qr = q + (size_t)row * D;
    const float* ep = table.row(pid, D);          // local HBM, or the owner's HBM through peer memory
    const float* en = table.row(nid, D);
    float sp = 0.f, sn = 0.f;
    float4 qv[2], pv[2], nv[2];   // D <= 256
    int k = 0;
    for (int c = lane * 4; c < D; c += 128, ++k) {
      qv[k] = *reinterpret_cast<const float4*>(qr + c);
      pv[k] = *reinterpret_cast<const float4*>(ep + c);
      nv[k] = *reinterpret_cast<const float4*>(en + c);
      sp += (qv[k].x * pv[k].x + qv[k].y * pv[k].y) + (qv[k].z * pv[k].z + qv[k].w * pv[k].w);
      sn += (qv[k].x * nv[k].x + qv[k].y * nv[k].y) + (qv[k].z * nv[k].z + qv[k].w * nv[k].w);
    }
    sp = warp_sum(sp);
    sn = warp_sum(sn);
    const float w = (loss_weight ? loss_weight[slot] : 1.0f);
    float dsp, dsn;
    if (KIND == 0) {
      if (lane == 0) loss_pos[slot] = (-log_sigmoid_f(sp) + softplus_f(sn)) * inv_n;
      dsp = -sigmoid_f(-sp) * inv_n * w * up;
      dsn = sigmoid_f(sn) * inv_n * w * up;
    } else {
      const float x = sp - sn;
      if (lane == 0) loss_pos[slot] = -log_sigmoid_f(x) * inv_n;
      dsp = -sigmoid_f(-x) * inv_n * w * up;
      dsn = -dsp;
    }
    if (lane == 0) { dscore[2 * (size_t)row] = dsp; dscore[2 * (size_t)row + 1] = dsn; }
    if (dq) {
      k = 0;
      for (int c = lane * 4; c < D; c += 128, ++k) {
        float4 o;
        o.x = dsp * pv[k].x + dsn * nv[k].x; o.y = dsp * pv[k].y + dsn * nv[k].y;
        o.z = dsp * pv[k].z + dsn * nv[k].z; o.w = dsp * pv[k].w + dsn * nv[k].w;
        *reinterpret_cast<float4*>(dq + (size_t)row * D + c) = o;
      }
    }
  }
}

// fixed-shape tree: every thread strides the array, then a block reduction; single CTA => deterministic
__global__ void __launch_bounds__(1024) sum_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;           // four loads in flight per thread; fixed combination order
  const int64_t step = blockDim.x;
  int64_t i = threadIdx.x;
  for (; i + 3 * step < n; i += 4 * step) { s0 += x[i]; s1 += x[i + step]; s2 += x[i + 2 * step]; s3 += x[i + 3 * step]; }
  for (; i < n; i += step) s0 += x[i];
  float s = (s0 + s1) + (s2 + s3);
  s = warp_sum(s);
  __shared__ float part[32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = part[threadIdx.x];
    v = warp_sum(v);
    if (threadIdx.x == 0) out[0] = v;
  }
}

__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// dE[in_id] += dx0 ; dE[item_id] += ds+ q ; dE[neg] += ds- q.  One warp per packed row, vector reds into L2.
__global__ void __launch_bounds__(256) table_grad_kernel(const float* __restrict__ dx0, const float* __restrict__ q,
                                                         const float* __restrict__ dscore, const int64_t* __restrict__ in_ids,
                                                         const int64_t* __restrict__ item_id, const int64_t* __restrict__ neg_item,
                                                         const int32_t* __restrict__ tok_off, const int32_t* __restrict__ row_seq,
                                                         const int32_t* __restrict__ counts, int L, int D, const __grid_constant__ ShardView tg) {
  const int T = counts[0];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = blockIdx.x * 8 + warp; row < T; row += gridDim.x * 8) {
    const int b = row_seq[row];
    const int t = row - tok_off[b];
    const size_t slot = (size_t)b * L + t;
    const int64_t iid = in_ids[slot];
    if (dx0 && iid != 0) {   // padding_idx = 0 receives no gradient (nn.Embedding)
      for (int c = lane * 4; c < D; c += 128)
        red_add_v4(tg.grad_row(iid, D) + c, *reinterpret_cast<const float4*>(dx0 + (size_t)row * D + c));
    }
    const int64_t pid = item_id ? item_id[slot] : 0;
    if (pid != 0) {
      const float dsp = dscore[2 * (size_t)row], dsn = dscore[2 * (size_t)row + 1];
      const int64_t nid = neg_item[slot];
      float* gp_row = tg.grad_row(pid, D);           // red.global.add into the owner's accumulator (NVLink atomics for remote rows)
      float* gn_row = tg.grad_row(nid, D);
      for (int c = lane * 4; c < D; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(q + (size_t)row * D + c);
        red_add_v4(gp_row + c, make_float4(dsp * v.x, dsp * v.y, dsp * v.z, dsp * v.w));
        red_add_v4(gn_row + c, make_float4(dsn * v.x, dsn * v.y, dsn * v.z, dsn * v.w));
      }
    }
  }
}

// dP[t] = sum_b dx0[(b,t)].  Each CTA streams the packed rows of a contiguous slice of sequences (the rows
// of a sequence are adjacent, so this is a linear read of dx0) and accumulates them per position in
// shared memory: warp w owns the positions t % 8 == w and lane c the columns 4c.., so every accumulator
// has a single writer and rows are added in a fixed order.  The [L, D] partial of each CTA goes to the
// workspace and the fixed-order reducer sums the kPosChunks partials -> deterministic.
constexpr int kPosChunks = 2 * kNumSMs;
__global__ void __launch_bounds__(256) pos_grad_kernel(const float* __restrict__ dx0, const int32_t* __restrict__ tok_off, int B, int L,
                                                       int D, float* __restrict__ partial) {
  extern __shared__ __align__(16) float acc[];           // [L][D]
  for (int e = threadIdx.x; e < L * D; e += blockDim.x) acc[e] = 0.f;
  __syncthreads();
  const int per = (B + gridDim.x - 1) / gridDim.x;
  const int b0 = min(B, (int)blockIdx.x * per), b1 = min(B, b0 + per);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = b0; b < b1; ++b) {
    const int off = tok_off[b], len = min(tok_off[b + 1] - off, L);
    for (int t = warp; t < len; t += 8) {
      for (int c = lane * 4; c < D; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(dx0 + (size_t)(off + t) * D + c);
        float4* a = reinterpret_cast<float4*>(acc + t * D + c);
        float4 o = *a;
        o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
        *a = o;
      }
    }
  }
  __syncthreads();
  float* out = partial + (size_t)blockIdx.x * L * D;
  for (int e = threadIdx.x; e < L * D; e += blockDim.x) out[e] = acc[e];
}

// torch.optim.Adam (_single_tensor_adam op order): g += wd p; m = lerp(m, g, 1-b1);
// v = b2 v + (1-b2) g g; p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n4, int64_t n, float step_size,
                                                   float inv_bc2_sqrt, float b1, float b2, float eps, float wd, int zero_grad) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i], gg = reinterpret_cast<float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float* P = &pp.x; float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gk = G[k];
      if (wd != 0.f) gk = fmaf(wd, P[k], gk);
      M[k] = M[k] + (gk - M[k]) * (1.0f - b1);
      V[k] = V[k] * b2 + (1.0f - b2) * gk * gk;
      const float denom = sqrtf(V[k]) * inv_bc2_sqrt + eps;
      P[k] = P[k] - step_size * (M[k] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // scalar tail (n not a multiple of 4)
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float gk = g[i];
    if (wd != 0.f) gk = fmaf(wd, p[i], gk);
    const float mk = m[i] + (gk - m[i]) * (1.0f - b1);
    const float vk = v[i] * b2 + (1.0f - b2) * gk * gk;
    p[i] = p[i] - step_size * (mk / (sqrtf(vk) * inv_bc2_sqrt + eps));
    m[i] = mk; v[i] = vk;
    if (zero_grad) g[i] = 0.f;
  }
}

// dq, dscore *= *upstream over the live rows; leaves immediately when the factor is 1 (the usual loss.backward())
__global__ void __launch_bounds__(256) scale_grads_kernel(const float* __restrict__ upstream, const int32_t* __restrict__ counts,
                                                          int D, float* __restrict__ dq, float* __restrict__ dscore) {
  const float up = *upstream;
  if (up == 1.0f) return;
  const int T = counts[0];
  const int64_t nq = (int64_t)T * D / 4, ns = (int64_t)T * 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<float4*>(dq)[i];
    v.x *= up; v.y *= up; v.z *= up; v.w *= up;
    reinterpret_cast<float4*>(dq)[i] = v;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ns; i += (int64_t)gridDim.x * blockDim.x) dscore[i] *= up;
}

}  // namespace
}  // namespace dr4sr

using namespace dr4sr;

extern "C" int dr4sr_scale_grads(const float* upstream, const int32_t* counts, int32_t D, float* dq_packed, float* dscore,
                                 dr4sr_stream_t stream) {
  if (!upstream || !counts || !dq_packed || !dscore || D % 4) return DR4SR_EINVAL;
  ProfScope prof("scale_grads", as_stream(stream));
  scale_grads_kernel<<<2 * kNumSMs, 256, 0, as_stream(stream)>>>(upstream, counts, D, dq_packed, dscore);
  DR4SR_LAUNCH_CHECK("scale_grads_kernel");
  return DR4SR_OK;
}

static int score_loss_impl(int32_t kind, const float* q_packed, const ShardView& tv, const int64_t* item_id,
                           const int64_t* neg_item, const int32_t* tok_off, const int32_t* counts,
                           int32_t B, int32_t L, int32_t D, const float* loss_weight, const float* upstream, float* loss_pos,
                           float* dscore, float* dq_packed, dr4sr_stream_t stream) {
  if (!q_packed || !tv.table[0] || !item_id || !neg_item || !tok_off || !counts || !loss_pos || !dscore) return DR4SR_EINVAL;
  if (D % 4 || D > 256 || (kind != DR4SR_LOSS_BCE && kind != DR4SR_LOSS_BPR)) return DR4SR_EINVAL;
  const int total = B * L;
  const int blocks = ceil_div(total, 8) < 8 * kNumSMs ? ceil_div(total, 8) : 8 * kNumSMs;
  ProfScope prof(kind == DR4SR_LOSS_BCE ? "score_bce" : "score_bpr", as_stream(stream));
  if (kind == DR4SR_LOSS_BCE)
    score_bce_kernel<0><<<blocks, 256, 0, as_stream(stream)>>>(q_packed, tv, item_id, neg_item, tok_off, counts, B, L, D,
                                                                loss_weight, upstream, loss_pos, dscore, dq_packed);
  else
    score_bce_kernel<1><<<blocks, 256, 0, as_stream(stream)>>>(q_packed, tv, item_id, neg_item, tok_off, counts, B, L, D,
                                                                loss_weight, upstream, loss_pos, dscore, dq_packed);
  DR4SR_LAUNCH_CHECK("score_bce_kernel");
  return DR4SR_OK;
}

extern "C" int dr4sr_score_loss(int32_t kind, const float* q_packed, const float* table, const int64_t* item_id,
                                const int64_t* neg_item, const int32_t* tok_off, const int32_t* row_seq, const int32_t* counts,
                                int32_t B, int32_t L, int32_t D, const float* loss_weight, const float* upstream, float* loss_pos,
                                float* dscore, float* dq_packed, dr4sr_stream_t stream) {
  (void)row_seq;
  return score_loss_impl(kind, q_packed, shard_view_local(table, nullptr, 0), item_id, neg_item, tok_off, counts, B, L, D, loss_weight,
                         upstream, loss_pos, dscore, dq_packed, stream);
}

extern "C" int dr4sr_score_loss_sharded(int32_t kind, const float* q_packed, const dr4sr_shard_map* map, const int64_t* item_id,
                                        const int64_t* neg_item, const int32_t* tok_off, const int32_t* row_seq, const int32_t* counts,
                                        int32_t B, int32_t L, int32_t D, const float* loss_weight, const float* upstream,
                                        float* loss_pos, float* dscore, float* dq_packed, dr4sr_stream_t stream) {
  (void)row_seq;
  ShardView tv;
  if (!shard_view_from(map, &tv)) return DR4SR_EINVAL;
  return score_loss_impl(kind, q_packed, tv, item_id, neg_item, tok_off, counts, B, L, D, loss_weight, upstream, loss_pos, dscore,
                         dq_packed, stream);
}

extern "C" int dr4sr_score_bce(const float* q_packed, const float* table, const int64_t* item_id, const int64_t* neg_item,
                               const int32_t* tok_off, const int32_t* row_seq, const int32_t* counts, int32_t B, int32_t L,
                               int32_t D, const float* loss_weight, const float* upstream, float* loss_pos, float* dscore,
                               float* dq_packed, dr4sr_stream_t stream) {
  return dr4sr_score_loss(DR4SR_LOSS_BCE, q_packed, table, item_id, neg_item, tok_off, row_seq, counts, B, L, D, loss_weight,
                          upstream, loss_pos, dscore, dq_packed, stream);
}

extern "C" int dr4sr_sum(const float* x, int64_t n, float* out, dr4sr_stream_t stream) {
  if (!x || !out || n < 0) return DR4SR_EINVAL;
  ProfScope prof("loss_sum", as_stream(stream));
  sum_kernel<<<1, 1024, 0, as_stream(stream)>>>(x, n, out);
  DR4SR_LAUNCH_CHECK("sum_kernel");
  return DR4SR_OK;
}

extern "C" size_t dr4sr_table_grad_workspace_bytes(int32_t L, int32_t D) { return sizeof(float) * (size_t)kPosChunks * L * D; }

namespace dr4sr {
// dP[t] = sum_b dx0[(b, t)]: per-CTA partials in `ws` (dr4sr_table_grad_workspace_bytes), summed in a fixed order
int launch_pos_grad(const float* dx0_packed, const int32_t* tok_off, int B, int L, int D, float* pos_grad, void* ws, cudaStream_t sa) {
  float* partial = reinterpret_cast<float*>(ws);
  const size_t smem = sizeof(float) * (size_t)L * D;
  {
    ProfScope prof("pos_grad", sa);
    if (cudaFuncSetAttribute(pos_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_cuda_error(cudaGetLastError(), "pos_grad smem attribute");
      return DR4SR_ECUDA;
    }
    pos_grad_kernel<<<kPosChunks, 256, smem, sa>>>(dx0_packed, tok_off, B, L, D, partial);
    DR4SR_LAUNCH_CHECK("pos_grad_kernel");
  }
  ReduceTable tab{};
  tab.seg[0] = ReduceSeg{partial, pos_grad, kPosChunks, (int64_t)L * D, L * D};
  tab.count = 1;
  return launch_reduce_segments(tab, sa);
}
}  // namespace dr4sr

static int table_grad_impl(const float* dx0_packed, const float* q_packed, const float* dscore,
                           const int64_t* in_item_id, const int64_t* item_id, const int64_t* neg_item,
                           const int32_t* tok_off, const int32_t* row_seq, const int32_t* counts, int32_t B, int32_t L,
                           int32_t D, const ShardView& gv, float* pos_grad, void* ws, size_t ws_bytes,
                           dr4sr_stream_t stream) {
  float* table_grad = gv.grad[0];
  if (!in_item_id || !tok_off || !row_seq || !counts || !table_grad || D % 4) return DR4SR_EINVAL;
  if (item_id && (!q_packed || !dscore || !neg_item)) return DR4SR_EINVAL;
  cudaStream_t st = as_stream(stream);
  const int T_cap = B * L;
  const int blocks = ceil_div(T_cap, 8) < 8 * kNumSMs ? ceil_div(T_cap, 8) : 8 * kNumSMs;
  // the positional gradient (a deterministic column reduction of dx0) shares nothing with the scatter-add but its input:
  // it runs on the auxiliary stream, beside the scatter
  const bool with_pos = pos_grad && dx0_packed;
  cudaStream_t sa = st;
  if (with_pos) {
    if (!ws || ws_bytes < dr4sr_table_grad_workspace_bytes(L, D)) return DR4SR_EWORKSPACE;
    sa = aux_fork(st);
    DR4SR_TRY(launch_pos_grad(dx0_packed, tok_off, B, L, D, pos_grad, ws, sa));
  }
  {
  ProfScope prof("table_grad_scatter", st);
  table_grad_kernel<<<blocks, 256, 0, st>>>(dx0_packed, q_packed, dscore, in_item_id, item_id, neg_item, tok_off, row_seq,
                                            counts, L, D, gv);
  DR4SR_LAUNCH_CHECK("table_grad_kernel");
  }
  if (with_pos) DR4SR_TRY(aux_join(sa, st));
  return DR4SR_OK;
}

extern "C" int dr4sr_table_grad(const float* dx0_packed, const float* q_packed, const float* dscore,
                                const int64_t* in_item_id, const int64_t* item_id, const int64_t* neg_item,
                                const int32_t* tok_off, const int32_t* row_seq, const int32_t* counts, int32_t B, int32_t L,
                                int32_t D, int64_t N, float* table_grad, float* pos_grad, void* ws, size_t ws_bytes,
                                dr4sr_stream_t stream) {
  return table_grad_impl(dx0_packed, q_packed, dscore, in_item_id, item_id, neg_item, tok_off, row_seq, counts, B, L, D,
                         shard_view_local(nullptr, table_grad, N), pos_grad, ws, ws_bytes, stream);
}

extern "C" int dr4sr_table_grad_sharded(const float* dx0_packed, const float* q_packed, const float* dscore,
                                        const int64_t* in_item_id, const int64_t* item_id, const int64_t* neg_item,
                                        const int32_t* tok_off, const int32_t* row_seq, const int32_t* counts, int32_t B, int32_t L,
                                        int32_t D, const dr4sr_shard_map* map, float* pos_grad, void* ws, size_t ws_bytes,
                                        dr4sr_stream_t stream) {
  ShardView gv;
  if (!shard_view_from(map, &gv)) return DR4SR_EINVAL;
  return table_grad_impl(dx0_packed, q_packed, dscore, in_item_id, item_id, neg_item, tok_off, row_seq, counts, B, L, D, gv, pos_grad,
                         ws, ws_bytes, stream);
}

// The target / negative rows of the scatter-add depend on the loss kernel only (ds, q), not on the encoder backward: queued on
// the background stream they run under it; dr4sr_table_grad_targets_join makes `stream` wait for them.
static int table_grad_targets_impl(const float* q_packed, const float* dscore, const int64_t* item_id, const int64_t* neg_item,
                                   const int32_t* tok_off, const int32_t* row_seq, const int32_t* counts, int32_t B, int32_t L,
                                   int32_t D, const ShardView& gv, dr4sr_stream_t stream) {
  if (!q_packed || !dscore || !item_id || !neg_item || !tok_off || !row_seq || !counts || !gv.grad[0] || D % 4) return DR4SR_EINVAL;
  cudaStream_t st = as_stream(stream);
  cudaStream_t sb = bg_fork(st);
  const int T_cap = B * L;
  const int blocks = ceil_div(T_cap, 8) < 8 * kNumSMs ? ceil_div(T_cap, 8) : 8 * kNumSMs;
  {
    ProfScope prof("table_grad_targets", sb);
    table_grad_kernel<<<blocks, 256, 0, sb>>>(nullptr, q_packed, dscore, item_id, item_id, neg_item, tok_off, row_seq, counts, L, D, gv);
    DR4SR_LAUNCH_CHECK("table_grad_kernel (targets)");
  }
  return bg_mark(sb, st);
}
extern "C" int dr4sr_table_grad_targets_async(const float* q_packed, const float* dscore, const int64_t* item_id,
                                              const int64_t* neg_item, const int32_t* tok_off, const int32_t* row_seq,
                                              const int32_t* counts, int32_t B, int32_t L, int32_t D, int64_t N, float* table_grad,
                                              dr4sr_stream_t stream) {
  return table_grad_targets_impl(q_packed, dscore, item_id, neg_item, tok_off, row_seq, counts, B, L, D,
                                 shard_view_local(nullptr, table_grad, N), stream);
}
extern "C" int dr4sr_table_grad_targets_async_sharded(const float* q_packed, const float* dscore, const int64_t* item_id,
                                                      const int64_t* neg_item, const int32_t* tok_off, const int32_t* row_seq,
                                                      const int32_t* counts, int32_t B, int32_t L, int32_t D,
                                                      const dr4sr_shard_map* map, dr4sr_stream_t stream) {
  ShardView gv;
  if (!shard_view_from(map, &gv)) return DR4SR_EINVAL;
  return table_grad_targets_impl(q_packed, dscore, item_id, neg_item, tok_off, row_seq, counts, B, L, D, gv, stream);
}
extern "C" int dr4sr_table_grad_targets_join(dr4sr_stream_t stream) { return bg_join(as_stream(stream)); }

extern "C" int dr4sr_adam(float* p, float* g, float* m, float* v, int64_t n, int64_t step, float lr, float beta1, float beta2,
                          float eps, float weight_decay, int32_t zero_grad, dr4sr_stream_t stream) {
  if (!p || !g || !m || !v || n < 0 || step < 1) return DR4SR_EINVAL;
  if (n == 0) return DR4SR_OK;
  if (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) return DR4SR_EINVAL;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  const int64_t n4 = n / 4;
  int64_t want = (n4 + 255) / 256;
  const int blocks = (int)(want < 1 ? 1 : (want > 16 * kNumSMs ? 16 * kNumSMs : want));
  ProfScope prof(n >= (1 << 22) ? "adam_table" : "adam_small", as_stream(stream));
  adam_kernel<<<blocks, 256, 0, as_stream(stream)>>>(p, g, m, v, n4, n, step_size, inv_bc2_sqrt, beta1, beta2, eps,
                                                     weight_decay, zero_grad);
  DR4SR_LAUNCH_CHECK("adam_kernel");
  return DR4SR_OK;
}
