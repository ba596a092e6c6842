// gru.cu -- GRU4Rec encoder forward / backward over packed rows.
//
// Replaces GRU4Rec.query_encoder (reference model/gru4rec.py:12-22, module/layers.py:117-136):
//   x0 = drop(E[ids]);  n_layer x bias-free GRU (h0 = 0, gate rows [r; z; n]);  y = W_out h + b_out
// Spec: SURVEY.md Appendix C.3.  Pads are trailing, so only slots t < seqlen exist (packed layout) and a
// sequence simply stops stepping at its length.
//
// Structure per layer
//   * input projection for all live tokens at once: gi = x W_ih^T  (tcgen05 / FFMA GEMM, dense.cuh)
//   * the recurrence as ONE cluster-resident kernel: a cluster of 8 CTAs owns a tile of 32 sequences
//     (sequences are counting-sorted by length so a tile steps in lock-step); CTA r keeps the W_hh rows
//     of its H/8 hidden units (r-, z- and n-gate rows: 3H/8 x H fp32, ~100 KB for H = 256) in shared
//     memory for the whole sequence, so the recurrent weights are read from HBM/L2 once per layer, not
//     once per step.  Each step: gh = h_{t-1} W_slice^T (register-tiled FFMA), gates, then the new h of
//     the CTA's units is pushed to the 7 peers through distributed shared memory (cluster barrier).
//   * backward-through-time mirrors it: dgh_slice (elementwise from the stored gates), partial
//     dh_{t-1} = dgh_slice W_slice over all H columns, reduce-scattered across the cluster via DSMEM.
//   * all weight gradients are token-batched GEMMs after the loop (dW_ih = dgi^T x, dW_hh = dgh^T h_prev).
#include <cooperative_groups.h>
#include "dense.cuh"

namespace cg = cooperative_groups;

namespace dr4sr {
namespace {

constexpr int kTile = 32;        // sequences per cluster (more, smaller tiles: long tiles start first, short ones back-fill the SMs)
constexpr int kCluster = 8;      // CTAs per cluster
constexpr int kGruThreads = 256;

struct GruOffsets { size_t w_ih[4], w_hh[4], w_out, b_out, total; };
GruOffsets gru_offsets(int D, int H, int n_layer) {
  GruOffsets o{};
  size_t p = 0;
  for (int l = 0; l < n_layer; ++l) {
    o.w_ih[l] = p; p += (size_t)3 * H * (l == 0 ? D : H);
    o.w_hh[l] = p; p += (size_t)3 * H * H;
  }
  o.w_out = p; p += (size_t)D * H;
  o.b_out = p; p += D;
  o.total = p;
  return o;
}

struct GruWs {
  float* x0;
  struct Layer { float *gi, *gates, *h, *hprev, *dgi, *dgh; Img ih_f, ih_b, hh_f; } layer[4];
  float *dh, *g0, *dy_mask;        // dh [T,H]: upstream gradient into a layer's outputs; g0 [T,D]
  float *part_w, *part_cs;         // weight-gradient partials [kSplit][max], colsum partials
  Img out_f, out_b;
  int32_t* order;                  // [B] sequences sorted by decreasing length
  size_t bytes;
};

GruWs carve(const dr4sr_gru_cfg& c, void* base) {
  GruWs w{};
  float* p = reinterpret_cast<float*>(base);
  size_t off = 0;
  const size_t T = (size_t)c.B * c.L, D = c.D, H = c.H;
  auto take = [&](size_t n) { float* r = p ? p + off : nullptr; off += ws_align(n); return r; };
  auto take_img = [&](size_t elems) { Img im; im.hi = reinterpret_cast<uint16_t*>(take((elems + 1) / 2)); im.lo = reinterpret_cast<uint16_t*>(take((elems + 1) / 2)); return im; };
  w.x0 = take(T * D);
  for (int l = 0; l < c.n_layer; ++l) {
    auto& y = w.layer[l];
    const size_t in = l == 0 ? D : H;
    y.gi = take(T * 3 * H); y.gates = take(T * 4 * H); y.h = take(T * H); y.hprev = take(T * H);
    y.dgi = take(T * 3 * H); y.dgh = take(T * 3 * H);
    y.ih_f = take_img(3 * H * in); y.ih_b = take_img(3 * H * in);
    y.hh_f = take_img(3 * H * H);                    // W_hh images: the tcgen05 recurrence copies its 96-row slices out of them
  }
  w.dh = take(T * H); w.g0 = take(T * D); w.dy_mask = nullptr;
  const size_t wmax = 3 * H * (H > D ? H : D);
  w.part_w = take((size_t)kSplit * (2 * wmax > D * H ? 2 * wmax : D * H));
  w.part_cs = take((size_t)kColsumBlocks * (D > 4 ? D : 4));
  w.out_f = take_img(D * H); w.out_b = take_img(D * H);
  w.order = reinterpret_cast<int32_t*>(take((size_t)c.B));
  w.bytes = off * sizeof(float);
  return w;
}

int check_cfg(const dr4sr_gru_cfg* c) {
  if (!c) return DR4SR_EINVAL;
  if (c->B <= 0 || c->B > 32768 || c->L <= 0 || c->L > 255 || c->n_layer < 1 || c->n_layer > 4) return DR4SR_EINVAL;
  if (c->D != 64 && c->D != 128) return DR4SR_EINVAL;
  if (c->H != 64 && c->H != 128 && c->H != 256) return DR4SR_EINVAL;
  if (c->dropout_p < 0.f || c->dropout_p >= 1.f) return DR4SR_EINVAL;
  return DR4SR_OK;
}

// order[] = sequence ids sorted by decreasing length (counting sort over L+1 buckets, stable); one CTA
__global__ void __launch_bounds__(1024) gru_order_kernel(const int32_t* __restrict__ tok_off, int B, int L, int32_t* __restrict__ order) {
  __shared__ int cnt[257], start[257];
  extern __shared__ unsigned char lens[];        // [B] lengths (L <= 255)
  for (int i = threadIdx.x; i <= L; i += blockDim.x) cnt[i] = 0;
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const int len = min(tok_off[b + 1] - tok_off[b], L);
    lens[b] = (unsigned char)len;
    atomicAdd(&cnt[len], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int len = L; len >= 0; --len) { start[len] = s; s += cnt[len]; }
  }
  __syncthreads();
  if (threadIdx.x <= L) {                        // stable: one thread per length walks the sequences in order
    const int len = threadIdx.x;
    int pos = start[len];
    for (int b = 0; b < B; ++b)
      if (lens[b] == len) order[pos++] = b;
  }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// thread tiling of one CTA's (sequence, unit) work for hidden size H
template <int H>
struct Tiling {
  static constexpr int UPC = H / kCluster;                  // hidden units per CTA
  static constexpr int UB = UPC >= 16 ? 16 : UPC;           // unit lanes
  static constexpr int UT = UPC / UB;                       // units per thread (stride UB)
  static constexpr int BB = kGruThreads / UB;               // batch lanes
  static constexpr int BT = kTile / BB;                     // sequences per thread (stride BB)
  static constexpr int LD = H + 4;                          // smem row stride: conflict-free 16-byte rows
  static constexpr int ROWS = 3 * UPC;                      // W_hh rows held by the CTA: local row = gate * UPC + unit
};

template <int H>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kGruThreads, 1)
gru_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ w_hh, const int32_t* __restrict__ tok_off,
               const int32_t* __restrict__ order, int B, float* __restrict__ h_out, float* __restrict__ hprev_out,
               float* __restrict__ gates) {
  using TL = Tiling<H>;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int tile = blockIdx.x / kCluster;
  extern __shared__ __align__(16) float sm[];
  float* Ws = sm;                              // [ROWS][LD]
  float* hb = Ws + TL::ROWS * TL::LD;          // [kTile][LD]  full h_{t-1} of the tile
  __shared__ int s_off[kTile], s_len[kTile];

  const int tid = threadIdx.x;
  for (int i = tid; i < kTile; i += kGruThreads) {
    const int bi = tile * kTile + i;
    int off = 0, len = 0;
    if (bi < B) { const int b = order[bi]; off = tok_off[b]; len = tok_off[b + 1] - off; }
    s_off[i] = off; s_len[i] = len;
  }
  // resident slice of W_hh: rows {g * H + rank * UPC + u}
  for (int e = tid; e < TL::ROWS * (H / 4); e += kGruThreads) {
    const int lr = e / (H / 4), c = (e % (H / 4)) * 4;
    const int g = lr / TL::UPC, u = lr % TL::UPC;
    *reinterpret_cast<float4*>(Ws + lr * TL::LD + c) =
        *reinterpret_cast<const float4*>(w_hh + (size_t)(g * H + rank * TL::UPC + u) * H + c);
  }
  for (int e = tid; e < kTile * TL::LD; e += kGruThreads) hb[e] = 0.f;   // h0 = 0
  __syncthreads();
  int maxlen = 0;
  for (int i = 0; i < kTile; ++i) maxlen = max(maxlen, s_len[i]);        // sorted: the first entry, but stay general
  cluster.sync();

  const int ub = tid % TL::UB, bb = tid / TL::UB;
  float* hs = hb + kTile * TL::LD;              // [kTile][UPC] staging of this CTA's new hidden units
  for (int t = 0; t < maxlen; ++t) {
    // input-projection terms of this step: issued now, consumed after the recurrent product (latency hidden)
    float gi_r[TL::BT][TL::UT], gi_z[TL::BT][TL::UT], gi_n[TL::BT][TL::UT];
    bool live[TL::BT];
#pragma unroll
    for (int i = 0; i < TL::BT; ++i) {
      const int bl = bb + i * TL::BB;
      live[i] = t < s_len[bl];
#pragma unroll
      for (int u = 0; u < TL::UT; ++u) {
        gi_r[i][u] = gi_z[i][u] = gi_n[i][u] = 0.f;
        if (live[i]) {
          const float* gir = gi + (size_t)(s_off[bl] + t) * 3 * H + rank * TL::UPC + ub + u * TL::UB;
          gi_r[i][u] = gir[0]; gi_z[i][u] = gir[H]; gi_n[i][u] = gir[2 * H];
        }
      }
    }
    // gh[b][row] = <h_{t-1}[b], W[row]> for the thread's BT sequences x (3 gates x UT units)
    float acc[TL::BT][3][TL::UT];
#pragma unroll
    for (int i = 0; i < TL::BT; ++i)
#pragma unroll
      for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int u = 0; u < TL::UT; ++u) acc[i][g][u] = 0.f;
#pragma unroll 2
    for (int k = 0; k < H; k += 4) {
      float4 hv[TL::BT], wv[3][TL::UT];
#pragma unroll
      for (int i = 0; i < TL::BT; ++i) hv[i] = *reinterpret_cast<const float4*>(hb + (bb + i * TL::BB) * TL::LD + k);
#pragma unroll
      for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int u = 0; u < TL::UT; ++u) wv[g][u] = *reinterpret_cast<const float4*>(Ws + (g * TL::UPC + ub + u * TL::UB) * TL::LD + k);
#pragma unroll
      for (int i = 0; i < TL::BT; ++i)
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
          for (int u = 0; u < TL::UT; ++u) {
            acc[i][g][u] = fmaf(hv[i].x, wv[g][u].x, acc[i][g][u]);
            acc[i][g][u] = fmaf(hv[i].y, wv[g][u].y, acc[i][g][u]);
            acc[i][g][u] = fmaf(hv[i].z, wv[g][u].z, acc[i][g][u]);
            acc[i][g][u] = fmaf(hv[i].w, wv[g][u].w, acc[i][g][u]);
          }
    }
    // gates; the new h of this CTA's units goes to the staging tile (finished sequences keep their state)
#pragma unroll
    for (int i = 0; i < TL::BT; ++i) {
      const int bl = bb + i * TL::BB;
      const size_t row = (size_t)(s_off[bl] + t);
#pragma unroll
      for (int u = 0; u < TL::UT; ++u) {
        const int ul = ub + u * TL::UB, unit = rank * TL::UPC + ul;
        const float hp = hb[bl * TL::LD + unit];
        float hv = hp;
        if (live[i]) {
          const float r = sigmoidf_(gi_r[i][u] + acc[i][0][u]);
          const float z = sigmoidf_(gi_z[i][u] + acc[i][1][u]);
          const float hn = acc[i][2][u];
          const float n = tanhf(gi_n[i][u] + r * hn);
          hv = (1.0f - z) * n + z * hp;
          h_out[row * H + unit] = hv;
          hprev_out[row * H + unit] = hp;
          float* gr = gates + row * 4 * H;
          gr[unit] = r; gr[H + unit] = z; gr[2 * H + unit] = n; gr[3 * H + unit] = hn;
        }
        hs[bl * TL::UPC + ul] = hv;
      }
    }
    cluster.sync();                               // every CTA is done reading h_{t-1}; staging tiles are complete
    // push the staged [kTile x UPC] block into every CTA's h buffer: 16-byte distributed-shared-memory stores
    constexpr int F4 = TL::UPC / 4;
    for (int e = tid; e < kCluster * kTile * F4; e += kGruThreads) {
      const int r = e / (kTile * F4), rem = e % (kTile * F4), bl = rem / F4, c = (rem % F4) * 4;
      *reinterpret_cast<float4*>(cluster.map_shared_rank(hb, r) + bl * TL::LD + rank * TL::UPC + c) =
          *reinterpret_cast<const float4*>(hs + bl * TL::UPC + c);
    }
    cluster.sync();                               // h_t visible in every CTA
  }
}

// backward through time of one layer.  dh_up [T,H]: gradient arriving at every h_t from above (read-only).
template <int H>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kGruThreads, 1)
gru_bwd_kernel(const float* __restrict__ dh_up, const float* __restrict__ w_hh, const float* __restrict__ gates,
               const float* __restrict__ hprev, const int32_t* __restrict__ tok_off, const int32_t* __restrict__ order, int B,
               float* __restrict__ dgi, float* __restrict__ dgh) {
  using TL = Tiling<H>;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int tile = blockIdx.x / kCluster;
  extern __shared__ __align__(16) float sm[];
  constexpr int LDG = TL::ROWS + 4;
  float* Ws = sm;                              // [ROWS][LD]
  float* part = Ws + TL::ROWS * TL::LD;        // [kTile][LD]   partial dh_{t-1} over all H columns (this CTA's gate rows)
  float* dg = part + kTile * TL::LD;           // [kTile][LDG]  dgh of this CTA's rows at the current step
  float* dhc = dg + kTile * LDG;               // [kTile][UPC]  carried dh (recurrent part) for this CTA's units
  __shared__ int s_off[kTile], s_len[kTile];

  const int tid = threadIdx.x;
  for (int i = tid; i < kTile; i += kGruThreads) {
    const int bi = tile * kTile + i;
    int off = 0, len = 0;
    if (bi < B) { const int b = order[bi]; off = tok_off[b]; len = tok_off[b + 1] - off; }
    s_off[i] = off; s_len[i] = len;
  }
  for (int e = tid; e < TL::ROWS * (H / 4); e += kGruThreads) {
    const int lr = e / (H / 4), c = (e % (H / 4)) * 4;
    const int g = lr / TL::UPC, u = lr % TL::UPC;
    *reinterpret_cast<float4*>(Ws + lr * TL::LD + c) =
        *reinterpret_cast<const float4*>(w_hh + (size_t)(g * H + rank * TL::UPC + u) * H + c);
  }
  for (int e = tid; e < kTile * TL::UPC; e += kGruThreads) dhc[e] = 0.f;
  __syncthreads();
  int maxlen = 0;
  for (int i = 0; i < kTile; ++i) maxlen = max(maxlen, s_len[i]);
  cluster.sync();

  // column tiling of the partial product: thread owns BTc sequences x 4*CT columns
  constexpr int CL = 16, CT = H / (4 * CL);          // 16 column lanes, CT float4 per thread (4 for H = 256)
  constexpr int BL = kGruThreads / CL, BTc = kTile / BL;   // 16 batch lanes x 4 sequences
  const int cl = tid % CL, bl0 = tid / CL;

  for (int t = maxlen - 1; t >= 0; --t) {
    // (1) elementwise gate gradients for this CTA's units
    for (int e = tid; e < kTile * TL::UPC; e += kGruThreads) {
      const int bl = e / TL::UPC, ul = e % TL::UPC;
      const int unit = rank * TL::UPC + ul;
      float dar = 0.f, daz = 0.f, dan = 0.f, dhn = 0.f, carry = 0.f;
      if (t < s_len[bl]) {
        const size_t row = (size_t)(s_off[bl] + t);
        const float* gr = gates + row * 4 * H;
        const float r = gr[unit], z = gr[H + unit], n = gr[2 * H + unit], hn = gr[3 * H + unit];
        const float hp = hprev[row * H + unit];
        const float dh = dh_up[row * H + unit] + dhc[bl * TL::UPC + ul];
        const float dn = dh * (1.0f - z);
        const float dz = dh * (hp - n);
        dan = dn * (1.0f - n * n);
        daz = dz * z * (1.0f - z);
        dar = dan * hn * r * (1.0f - r);
        dhn = dan * r;
        carry = dh * z;
        float* o1 = dgi + row * 3 * H;
        o1[unit] = dar; o1[H + unit] = daz; o1[2 * H + unit] = dan;
        float* o2 = dgh + row * 3 * H;
        o2[unit] = dar; o2[H + unit] = daz; o2[2 * H + unit] = dhn;
      }
      dg[bl * LDG + ul] = dar; dg[bl * LDG + TL::UPC + ul] = daz; dg[bl * LDG + 2 * TL::UPC + ul] = dhn;
      dhc[bl * TL::UPC + ul] = carry;                // direct path dh * z; the W_hh path is added after the exchange
    }
    __syncthreads();
    // (2) part[b][j] = sum_rows dg[b][row] * W[row][j]
    {
      float4 acc[BTc][CT];
#pragma unroll
      for (int i = 0; i < BTc; ++i)
#pragma unroll
        for (int c = 0; c < CT; ++c) acc[i][c] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < TL::ROWS; ++r) {
        float4 wv[CT];
        float gv[BTc];
#pragma unroll
        for (int c = 0; c < CT; ++c) wv[c] = *reinterpret_cast<const float4*>(Ws + r * TL::LD + (cl + c * CL) * 4);
#pragma unroll
        for (int i = 0; i < BTc; ++i) gv[i] = dg[(bl0 + i * BL) * LDG + r];
#pragma unroll
        for (int i = 0; i < BTc; ++i)
#pragma unroll
          for (int c = 0; c < CT; ++c) {
            acc[i][c].x = fmaf(gv[i], wv[c].x, acc[i][c].x); acc[i][c].y = fmaf(gv[i], wv[c].y, acc[i][c].y);
            acc[i][c].z = fmaf(gv[i], wv[c].z, acc[i][c].z); acc[i][c].w = fmaf(gv[i], wv[c].w, acc[i][c].w);
          }
      }
#pragma unroll
      for (int i = 0; i < BTc; ++i)
#pragma unroll
        for (int c = 0; c < CT; ++c) *reinterpret_cast<float4*>(part + (bl0 + i * BL) * TL::LD + (cl + c * CL) * 4) = acc[i][c];
    }
    cluster.sync();                                   // all partials written
    // (3) reduce-scatter: this CTA's units gather their columns from every CTA's partial (fixed rank order),
    //     16-byte distributed-shared-memory loads
    {
      constexpr int F4 = TL::UPC / 4;
      for (int e = tid; e < kTile * F4; e += kGruThreads) {
        const int bl = e / F4, c = (e % F4) * 4;
        float4 sacc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < kCluster; ++r) {
          const float4 v = *reinterpret_cast<const float4*>(cluster.map_shared_rank(part, r) + bl * TL::LD + rank * TL::UPC + c);
          sacc.x += v.x; sacc.y += v.y; sacc.z += v.z; sacc.w += v.w;
        }
        float4* d = reinterpret_cast<float4*>(dhc + bl * TL::UPC + c);
        float4 o = *d;
        o.x += sacc.x; o.y += sacc.y; o.z += sacc.z; o.w += sacc.w;
        *d = o;
      }
    }
    cluster.sync();                                   // partials consumed before the next step overwrites them
  }
}

template <int H>
size_t fwd_smem() { using TL = Tiling<H>; return sizeof(float) * (size_t)(TL::ROWS * TL::LD + kTile * TL::LD + kTile * TL::UPC); }
template <int H>
size_t bwd_smem() {
  using TL = Tiling<H>;
  return sizeof(float) * (size_t)(TL::ROWS * TL::LD + kTile * TL::LD + kTile * (TL::ROWS + 4) + kTile * TL::UPC);
}

template <int H>
int launch_gru_fwd_t(const float* gi, const float* w_hh, const int32_t* tok_off, const int32_t* order, int B, float* h, float* hprev,
                     float* gates, cudaStream_t st) {
  const size_t smem = fwd_smem<H>();
  ProfScope prof("gru_recurrence_fwd", st);
  if (cudaFuncSetAttribute(gru_fwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_cuda_error(cudaGetLastError(), "gru_fwd smem attribute");
    return DR4SR_ECUDA;
  }
  const int tiles = ceil_div(B, kTile);
  gru_fwd_kernel<H><<<tiles * kCluster, kGruThreads, smem, st>>>(gi, w_hh, tok_off, order, B, h, hprev, gates);
  DR4SR_LAUNCH_CHECK("gru_fwd_kernel");
  return DR4SR_OK;
}
template <int H>
int launch_gru_bwd_t(const float* dh_up, const float* w_hh, const float* gates, const float* hprev, const int32_t* tok_off,
                     const int32_t* order, int B, float* dgi, float* dgh, cudaStream_t st) {
  const size_t smem = bwd_smem<H>();
  ProfScope prof("gru_recurrence_bwd", st);
  if (cudaFuncSetAttribute(gru_bwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_cuda_error(cudaGetLastError(), "gru_bwd smem attribute");
    return DR4SR_ECUDA;
  }
  const int tiles = ceil_div(B, kTile);
  gru_bwd_kernel<H><<<tiles * kCluster, kGruThreads, smem, st>>>(dh_up, w_hh, gates, hprev, tok_off, order, B, dgi, dgh);
  DR4SR_LAUNCH_CHECK("gru_bwd_kernel");
  return DR4SR_OK;
}
int launch_gru_fwd(int H, const float* gi, const float* w_hh, const int32_t* tok_off, const int32_t* order, int B, float* h,
                   float* hprev, float* gates, cudaStream_t st) {
  if (H == 256) return launch_gru_fwd_t<256>(gi, w_hh, tok_off, order, B, h, hprev, gates, st);
  if (H == 128) return launch_gru_fwd_t<128>(gi, w_hh, tok_off, order, B, h, hprev, gates, st);
  return launch_gru_fwd_t<64>(gi, w_hh, tok_off, order, B, h, hprev, gates, st);
}
int launch_gru_bwd(int H, const float* dh_up, const float* w_hh, const float* gates, const float* hprev, const int32_t* tok_off,
                   const int32_t* order, int B, float* dgi, float* dgh, cudaStream_t st) {
  if (H == 256) return launch_gru_bwd_t<256>(dh_up, w_hh, gates, hprev, tok_off, order, B, dgi, dgh, st);
  if (H == 128) return launch_gru_bwd_t<128>(dh_up, w_hh, gates, hprev, tok_off, order, B, dgi, dgh, st);
  return launch_gru_bwd_t<64>(dh_up, w_hh, gates, hprev, tok_off, order, B, dgi, dgh, st);
}

int build_images(const dr4sr_gru_cfg& c, const float* params, const GruWs& w, const GruOffsets& lo, cudaStream_t st) {
  if (!tc_enabled()) return DR4SR_OK;
  const int D = c.D, H = c.H;
  tc::ImageTable tab{};
  auto add = [&](const float* src, int ld, int N, int K, int tr, const Img& im) {
    if (tc::tc_supported(N, K, false)) tab.job[tab.count++] = tc::ImageJob{src, ld, N, K, tr, im.hi, im.lo};
  };
  for (int l = 0; l < c.n_layer; ++l) {
    const int in = l == 0 ? D : H;
    add(params + lo.w_ih[l], in, 3 * H, in, 0, w.layer[l].ih_f);      // gi = x W_ih^T
    add(params + lo.w_ih[l], in, in, 3 * H, 1, w.layer[l].ih_b);      // dx = dgi W_ih
    if (gru_tc_supported(H)) add(params + lo.w_hh[l], H, 3 * H, H, 0, w.layer[l].hh_f);   // gh = h W_hh^T (recurrence)
  }
  add(params + lo.w_out, H, D, H, 0, w.out_f);                        // y = h W_out^T
  add(params + lo.w_out, H, H, D, 1, w.out_b);                        // dh = dy W_out
  return tc::launch_weight_images(tab, st);
}

}  // namespace
}  // namespace dr4sr

using namespace dr4sr;

extern "C" size_t dr4sr_gru_param_count(const dr4sr_gru_cfg* c) {
  if (check_cfg(c) != DR4SR_OK) return 0;
  return gru_offsets(c->D, c->H, c->n_layer).total;
}
extern "C" size_t dr4sr_gru_workspace_bytes(const dr4sr_gru_cfg* c) {
  if (check_cfg(c) != DR4SR_OK) return 0;
  return carve(*c, nullptr).bytes;
}

extern "C" int dr4sr_gru_fwd(const dr4sr_gru_cfg* c, const float* table, const float* params, const int64_t* in_item_id,
                             const int32_t* tok_off, const int32_t* row_seq, const int32_t* counts, void* ws, size_t ws_bytes,
                             int32_t train, float* q_packed, float* q_last, float* q_dense, dr4sr_stream_t stream) {
  DR4SR_TRY(check_cfg(c));
  if (!table || !params || !in_item_id || !tok_off || !row_seq || !counts || !ws || !q_packed) return DR4SR_EINVAL;
  GruWs w = carve(*c, ws);
  if (ws_bytes < w.bytes) return DR4SR_EWORKSPACE;
  cudaStream_t st = as_stream(stream);
  const int T = c->B * c->L, D = c->D, H = c->H;
  const bool tr = train != 0;
  const GruOffsets lo = gru_offsets(D, H, c->n_layer);

  DR4SR_TRY(dr4sr_embed_fwd(table, nullptr, in_item_id, tok_off, row_seq, counts, c->B, c->L, D, tr ? c->dropout_p : 0.f, c->seed,
                            c->step, w.x0, stream));
  {
    ProfScope prof("gru_order", st);
    gru_order_kernel<<<1, 1024, (size_t)c->B, st>>>(tok_off, c->B, c->L, w.order);
    DR4SR_LAUNCH_CHECK("gru_order_kernel");
  }
  DR4SR_TRY(build_images(*c, params, w, lo, st));
  const float* x = w.x0;
  int in = D;
  for (int l = 0; l < c->n_layer; ++l) {
    auto& y = w.layer[l];
    {
      GemmArgs g = gemm_args(x, in, params + lo.w_ih[l], in, y.gi, 3 * H, T, 3 * H, in, counts);
      g.tag = "gru_gemm_gi";
      DR4SR_TRY(gemm_nt(g, y.ih_f, st));
    }
    if (tc_enabled() && gru_tc_supported(H))
      DR4SR_TRY(launch_gru_fwd_tc(y.gi, y.hh_f.hi, y.hh_f.lo, tok_off, w.order, c->B, y.h, y.hprev, y.gates, st));
    else
      DR4SR_TRY(launch_gru_fwd(H, y.gi, params + lo.w_hh[l], tok_off, w.order, c->B, y.h, y.hprev, y.gates, st));
    x = y.h;
    in = H;
  }
  {
    GemmArgs g = gemm_args(x, H, params + lo.w_out, H, q_packed, D, T, D, H, counts);
    g.bias = params + lo.b_out; g.tag = "gru_gemm_out";
    DR4SR_TRY(gemm_nt(g, w.out_f, st));
  }
  return dr4sr_unpack_rows(q_packed, tok_off, c->B, c->L, D, q_last, q_dense, stream);
}

extern "C" int dr4sr_gru_bwd(const dr4sr_gru_cfg* c, const float* table, const float* params, const int64_t* in_item_id,
                             const int32_t* tok_off, const int32_t* row_seq, const int32_t* counts, void* ws, size_t ws_bytes,
                             float* dq_packed, float* grads, float* dx0_packed, dr4sr_stream_t stream) {
  (void)table; (void)in_item_id; (void)row_seq;
  DR4SR_TRY(check_cfg(c));
  if (!params || !tok_off || !counts || !ws || !dq_packed || !grads || !dx0_packed) return DR4SR_EINVAL;
  GruWs w = carve(*c, ws);
  if (ws_bytes < w.bytes) return DR4SR_EWORKSPACE;
  cudaStream_t st = as_stream(stream);
  const int T = c->B * c->L, D = c->D, H = c->H;
  const float p = c->dropout_p;
  const bool tr = p > 0.f;
  const GruOffsets lo = gru_offsets(D, H, c->n_layer);
  const int top = c->n_layer - 1;

  // output projection: dW_out = dy^T h, db_out = colsum(dy), dh_top = dy W_out
  DR4SR_TRY(launch_colsum(dq_packed, D, T, counts, w.part_cs, st));
  {
    GemmArgs g = gemm_args(dq_packed, D, params + lo.w_out, H, w.dh, H, T, H, D, counts);
    g.tag = "gru_gemm_bwd_dh";
    DR4SR_TRY(gemm_nn(g, w.out_b, st));
  }
  const bool tc_w = tc_enabled();
  auto wgrad = [&](const float* A, int M_out, const float* Bm, int N_out, float* partial, const char* tag) -> int {
    if (tc_w && tc::wgrad_supported(M_out, N_out)) {
      tc::WgradTable tab{};
      tab.job[0] = tc::WgradJob{A, M_out, PRO_NONE, no_dropout(), Bm, N_out, PRO_NONE, no_dropout(), M_out, N_out, partial, 0};
      tab.count = 1; tab.T_cap = T; tab.tok_dev = counts; tab.n_split = kSplit;
      return tc::launch_wgrad_tc(tab, st);
    }
    GemmArgs g = gemm_args(A, M_out, Bm, N_out, nullptr, N_out, M_out, N_out, T, counts);
    g.tag = tag;
    return gemm_tn(g, partial, st);
  };
  auto reduce1 = [&](const float* src, float* dst, int ns, int64_t stride, int n) -> int {
    ReduceTable tab{};
    tab.seg[0] = ReduceSeg{src, dst, ns, stride, n};
    tab.count = 1;
    return launch_reduce_segments(tab, st);
  };
  DR4SR_TRY(wgrad(dq_packed, D, w.layer[top].h, H, w.part_w, "gru_wgrad_out"));
  {
    ReduceTable tab{};
    tab.seg[0] = ReduceSeg{w.part_w, grads + lo.w_out, kSplit, (int64_t)D * H, D * H};
    tab.seg[1] = ReduceSeg{w.part_cs, grads + lo.b_out, kColsumBlocks, D, D};
    tab.count = 2;
    DR4SR_TRY(launch_reduce_segments(tab, st));
  }
  for (int l = top; l >= 0; --l) {
    auto& y = w.layer[l];
    const int in = l == 0 ? D : H;
    const float* xin = l == 0 ? w.x0 : w.layer[l - 1].h;
    DR4SR_TRY(launch_gru_bwd(H, w.dh, params + lo.w_hh[l], y.gates, y.hprev, tok_off, w.order, c->B, y.dgi, y.dgh, st));
    // dW_hh = dgh^T h_prev ; dW_ih = dgi^T x
    DR4SR_TRY(wgrad(y.dgh, 3 * H, y.hprev, H, w.part_w, "gru_wgrad_hh"));
    DR4SR_TRY(reduce1(w.part_w, grads + lo.w_hh[l], kSplit, (int64_t)3 * H * H, 3 * H * H));
    DR4SR_TRY(wgrad(y.dgi, 3 * H, xin, in, w.part_w, "gru_wgrad_ih"));
    DR4SR_TRY(reduce1(w.part_w, grads + lo.w_ih[l], kSplit, (int64_t)3 * H * in, 3 * H * in));
    {  // dx = dgi W_ih : gradient into the layer below (layer 0: into the gathered rows, through the embedding dropout)
      float* dst = l == 0 ? dx0_packed : w.dh;
      GemmArgs g = gemm_args(y.dgi, 3 * H, params + lo.w_ih[l], in, dst, in, T, in, 3 * H, counts);
      if (l == 0) g.dropE = make_dropout(p, c->seed, c->step, SITE_EMBED, tr);
      g.tag = "gru_gemm_bwd_dx";
      DR4SR_TRY(gemm_nn(g, y.ih_b, st));
    }
  }
  return DR4SR_OK;
}
