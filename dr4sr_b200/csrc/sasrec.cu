// sasrec.cu -- SASRec encoder forward / backward over packed rows (orchestration of the kernels).
//
// Replaces SASRecQueryEncoder.forward (reference model/sasrec.py:39-75) = embedding + learned
// positions + dropout -> 2 x post-norm TransformerEncoderLayer (model/sasrec.py:21-34) -> pooling,
// and its autograd backward.  Spec: SURVEY.md Appendix C.1, C.2, C.5.
#include <stdlib.h>
#include "dense.cuh"

namespace dr4sr {
namespace {


struct LayerOffsets {   // offsets (floats) inside one layer's slice of the flat parameter buffer
  size_t in_w, in_b, out_w, out_b, w1, b1, w2, b2, g1, be1, g2, be2, total;
};
LayerOffsets layer_offsets(int D, int F) {
  LayerOffsets o;
  size_t p = 0;
  o.in_w = p; p += (size_t)3 * D * D;
  o.in_b = p; p += (size_t)3 * D;
  o.out_w = p; p += (size_t)D * D;
  o.out_b = p; p += D;
  o.w1 = p; p += (size_t)F * D;
  o.b1 = p; p += F;
  o.w2 = p; p += (size_t)D * F;
  o.b2 = p; p += D;
  o.g1 = p; p += D;
  o.be1 = p; p += D;
  o.g2 = p; p += D;
  o.be2 = p; p += D;
  o.total = p;
  return o;
}

struct Workspace {
  // saved activations
  float* x0;
  // fused forward: hm = mask * gelu(pre), gp = mask * gelu'(pre) instead of pre; per-operator forward: pre
  struct Layer { float *qkv, *attn, *z1, *st1, *x1, *pre, *hm, *gp, *z2, *st2, *x2; } layer[8];
  // backward scratch.  g0 / g2 belong to the data-gradient chain; everything in `Bwd` is per layer because the
  // weight-gradient work of layer l runs on a side stream while the main stream already works on layer l-1
  float *g0, *g2;
  struct Bwd {
    float *g1, *g3, *dqkv, *dpre;
    float *dx1, *gout;                // fused backward: dx1 rows (LN1' input, re-read by the side stream) and this layer's dx
    float *part_w;                    // [kSplit][3DD + DD + FD + DF] weight-gradient partials
    float *part_ln2, *part_ln1;       // [kLnBwdBlocks][3D]
    float *part_cs_in, *part_cs_b1;   // [kColsumBlocks][3D], [kColsumBlocks][F]
  } bwd[8];
  // bf16 hi/lo weight images (UMMA SW128 K-major) of every layer: forward operands W[n,k] and the
  // transposed backward-data operands W^T, rebuilt at the start of every forward
  struct LayerImg { Img in_f, out_f, w1_f, w2_f, in_b, out_b, w1_b, w2_b; } img[8];
  int32_t* tile_first;   // attention tiles (tcgen05 path): first sequence of every <= 128-row tile
  int32_t* fused_tiles;  // fused encoder kernels: {n_tiles, first sequence of every greedy <= 128-row tile..., B}
  size_t bytes;
};

inline size_t align_up(size_t x) { return (x + 63) & ~(size_t)63; }   // 256-byte granules (in floats: 64)

Workspace carve(const dr4sr_sasrec_cfg& c, void* base) {
  Workspace w{};
  float* p = reinterpret_cast<float*>(base);
  size_t off = 0;
  const size_t T = (size_t)c.B * c.L, D = c.D, F = c.F;
  auto take = [&](size_t n) { float* r = p ? p + off : nullptr; off += align_up(n); return r; };
  w.x0 = take(T * D);
  for (int l = 0; l < c.n_layer; ++l) {
    auto& y = w.layer[l];
    y.qkv = take(T * 3 * D); y.attn = take(T * D); y.z1 = take(T * D); y.st1 = take(T * 2); y.x1 = take(T * D);
    y.pre = take(T * F); y.hm = take(T * F); y.gp = y.pre;   // gp and pre never coexist (fused vs per-operator forward)
    y.z2 = take(T * D); y.st2 = take(T * 2); y.x2 = take(T * D);
  }
  w.g0 = take(T * D); w.g2 = take(T * D);
  for (int l = 0; l < c.n_layer; ++l) {
    auto& b = w.bwd[l];
    b.g1 = take(T * D); b.g3 = take(T * D); b.dqkv = take(T * 3 * D); b.dpre = take(T * F);
    b.dx1 = take(T * D); b.gout = take(T * D);
    b.part_w = take((size_t)kSplit * (3 * D * D + D * D + 2 * F * D));
    b.part_ln2 = take((size_t)kLnBwdBlocks * 3 * D);
    b.part_ln1 = take((size_t)kLnBwdBlocks * 3 * D);
    b.part_cs_in = take((size_t)kColsumBlocks * 3 * D);
    b.part_cs_b1 = take((size_t)kColsumBlocks * F);
  }
  auto take_img = [&](size_t elems) {   // hi + lo images, bf16; 1 KB aligned (align_up keeps 256 B, images are multiples of 16 KB)
    Img im;
    im.hi = reinterpret_cast<uint16_t*>(take((elems + 1) / 2));
    im.lo = reinterpret_cast<uint16_t*>(take((elems + 1) / 2));
    return im;
  };
  for (int l = 0; l < c.n_layer; ++l) {
    auto& m = w.img[l];
    m.in_f = take_img(3 * D * D); m.out_f = take_img(D * D); m.w1_f = take_img(F * D); m.w2_f = take_img(D * F);
    m.in_b = take_img(3 * D * D); m.out_b = take_img(D * D); m.w1_b = take_img(F * D); m.w2_b = take_img(D * F);
  }
  w.tile_first = reinterpret_cast<int32_t*>(take((size_t)attn_tc_num_tiles(c.B, c.L) + 2));
  w.fused_tiles = reinterpret_cast<int32_t*>(take((size_t)fused_tiles_cap(c.B, c.L) + 4));
  w.bytes = off * sizeof(float);
  return w;
}

int check_cfg(const dr4sr_sasrec_cfg* c) {
  if (!c) return DR4SR_EINVAL;
  if (c->B <= 0 || c->L <= 0 || c->L > 64 || c->n_layer < 1 || c->n_layer > 8) return DR4SR_EINVAL;
  if (c->D != 64 && c->D != 128) return DR4SR_EINVAL;            // LN-epilogue tiles are instantiated for these
  if (c->F % 64 || c->F <= 0 || c->F > 1024) return DR4SR_EINVAL;
  if (c->n_head <= 0 || c->D % c->n_head || (c->D / c->n_head) % 4) return DR4SR_EINVAL;
  if (c->dropout_p < 0.f || c->dropout_p >= 1.f) return DR4SR_EINVAL;
  return DR4SR_OK;
}

// bf16 hi/lo images of all layers' weights (one launch per <= 2 layers)
int build_weight_images(const dr4sr_sasrec_cfg& c, const float* params, const Workspace& w, const LayerOffsets& lo, cudaStream_t st) {
  const int D = c.D, F = c.F;
  if (!tc_enabled()) return DR4SR_OK;   // FFMA path needs none
  tc::ImageTable tab{};
  for (int l = 0; l < c.n_layer; ++l) {
    const float* lp = params + (size_t)c.L * D + (size_t)l * lo.total;
    const auto& m = w.img[l];
    auto add = [&](const float* src, int ld, int N, int K, int tr, const Img& im) {   // same predicate as use_tc()
      if (tc::tc_supported(N, K, false)) tab.job[tab.count++] = tc::ImageJob{src, ld, N, K, tr, im.hi, im.lo};
    };
    add(lp + lo.in_w, D, 3 * D, D, 0, m.in_f);
    add(lp + lo.out_w, D, D, D, 0, m.out_f);
    add(lp + lo.w1, D, F, D, 0, m.w1_f);
    add(lp + lo.w2, F, D, F, 0, m.w2_f);
    add(lp + lo.in_w, D, D, 3 * D, 1, m.in_b);     // dx  = dqkv Win : B'[d][j] = Win[j][d]
    add(lp + lo.out_w, D, D, D, 1, m.out_b);       // dO  = dz1 Wo   : B'[k][n] = Wo[n][k]
    add(lp + lo.w1, D, D, F, 1, m.w1_b);           // dx1 = dpre W1  : B'[d][f] = W1[f][d]
    add(lp + lo.w2, F, F, D, 1, m.w2_b);           // dh  = dz2 W2   : B'[f][d] = W2[d][f]
    if (tab.count + 8 > tc::kMaxImageJobs || l == c.n_layer - 1) {
      DR4SR_TRY(tc::launch_weight_images(tab, st));
      tab.count = 0;
    }
  }
  return DR4SR_OK;
}
// fixed-order reduction of every partial of a layer into the flat gradient buffer
int reduce_layer_partials(const float* part_w, size_t pw_in, size_t pw_out, size_t pw_w1, size_t pw_w2, const float* part_cs_in,
                          const float* part_cs_b1, const float* part_ln1, const float* part_ln2, float* lg, const LayerOffsets& lo, int D,
                          int F, cudaStream_t sw) {
  ReduceTable tab{};
  int k = 0;
  auto seg = [&](const float* src, float* dst, int ns, int64_t stride, int n) { tab.seg[k++] = ReduceSeg{src, dst, ns, stride, n}; };
  seg(part_w + pw_in, lg + lo.in_w, kSplit, (int64_t)3 * D * D, 3 * D * D);
  seg(part_w + pw_out, lg + lo.out_w, kSplit, (int64_t)D * D, D * D);
  seg(part_w + pw_w1, lg + lo.w1, kSplit, (int64_t)F * D, F * D);
  seg(part_w + pw_w2, lg + lo.w2, kSplit, (int64_t)D * F, D * F);
  seg(part_cs_in, lg + lo.in_b, kColsumBlocks, 3 * D, 3 * D);
  seg(part_cs_b1, lg + lo.b1, kColsumBlocks, F, F);
  seg(part_ln1, lg + lo.g1, kLnBwdBlocks, 3 * D, D);
  seg(part_ln1 + D, lg + lo.be1, kLnBwdBlocks, 3 * D, D);
  seg(part_ln1 + 2 * D, lg + lo.out_b, kLnBwdBlocks, 3 * D, D);
  seg(part_ln2, lg + lo.g2, kLnBwdBlocks, 3 * D, D);
  seg(part_ln2 + D, lg + lo.be2, kLnBwdBlocks, 3 * D, D);
  seg(part_ln2 + 2 * D, lg + lo.b2, kLnBwdBlocks, 3 * D, D);
  tab.count = k;
  return launch_reduce_segments(tab, sw);
}

// Side stream for the weight-gradient work of the backward (fork/join with events; also legal inside a
// CUDA-graph capture of the main stream).  One per device, created on first use, never destroyed.
struct SideStream { cudaStream_t s = nullptr; cudaEvent_t fork[16] = {}, join = nullptr; bool ok = false; };
SideStream& side_stream() {
  static SideStream per_dev[64];
  int dev = 0;
  cudaGetDevice(&dev);
  SideStream& x = per_dev[dev & 63];
  if (!x.ok) {
    bool good = cudaStreamCreateWithFlags(&x.s, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 16 && good; ++i) good = cudaEventCreateWithFlags(&x.fork[i], cudaEventDisableTiming) == cudaSuccess;
    good = good && cudaEventCreateWithFlags(&x.join, cudaEventDisableTiming) == cudaSuccess;
    x.ok = good;
  }
  return x;
}

__global__ void __launch_bounds__(256) gather_last_kernel(const float* __restrict__ x, const int32_t* __restrict__ tok_off, int B,
                                                          int D, float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = blockIdx.x * 8 + warp; b < B; b += gridDim.x * 8) {
    const int end = tok_off[b + 1], len = end - tok_off[b];
    for (int c = lane * 4; c < D; c += 128) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (len > 0) v = *reinterpret_cast<const float4*>(x + (size_t)(end - 1) * D + c);
      *reinterpret_cast<float4*>(out + (size_t)b * D + c) = v;
    }
  }
}

__global__ void __launch_bounds__(256) unpack_dense_kernel(const float* __restrict__ x, const int32_t* __restrict__ tok_off, int B,
                                                           int L, int D, float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int slot = blockIdx.x * 8 + warp; slot < B * L; slot += gridDim.x * 8) {
    const int b = slot / L, t = slot % L;
    const int off = tok_off[b], len = tok_off[b + 1] - off;
    for (int c = lane * 4; c < D; c += 128) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t < len) v = *reinterpret_cast<const float4*>(x + (size_t)(off + t) * D + c);
      *reinterpret_cast<float4*>(out + (size_t)slot * D + c) = v;
    }
  }
}

}  // namespace
}  // namespace dr4sr

using namespace dr4sr;

extern "C" int dr4sr_unpack_rows(const float* x_packed, const int32_t* tok_off, int32_t B, int32_t L, int32_t D, float* q_last,
                                 float* q_dense, dr4sr_stream_t stream) {
  if (!x_packed || !tok_off || D % 4) return DR4SR_EINVAL;
  cudaStream_t st = as_stream(stream);
  if (q_last) {
    ProfScope prof("gather_last", st);
    gather_last_kernel<<<ceil_div(B, 8), 256, 0, st>>>(x_packed, tok_off, B, D, q_last);
    DR4SR_LAUNCH_CHECK("gather_last_kernel");
  }
  if (q_dense) {
    const int T = B * L;
    const int blocks = ceil_div(T, 8) < 8 * kNumSMs ? ceil_div(T, 8) : 8 * kNumSMs;
    ProfScope prof("unpack_dense", st);
    unpack_dense_kernel<<<blocks, 256, 0, st>>>(x_packed, tok_off, B, L, D, q_dense);
    DR4SR_LAUNCH_CHECK("unpack_dense_kernel");
  }
  return DR4SR_OK;
}

extern "C" size_t dr4sr_sasrec_param_count(const dr4sr_sasrec_cfg* c) {
  if (check_cfg(c) != DR4SR_OK) return 0;
  return (size_t)c->L * c->D + (size_t)c->n_layer * layer_offsets(c->D, c->F).total;
}

extern "C" size_t dr4sr_sasrec_workspace_bytes(const dr4sr_sasrec_cfg* c) {
  if (check_cfg(c) != DR4SR_OK) return 0;
  return carve(*c, nullptr).bytes;
}

static int sasrec_fwd_impl(const dr4sr_sasrec_cfg* c, const ShardView& tv, const float* params,
                           const int64_t* in_item_id, const int32_t* tok_off, const int32_t* row_seq,
                           const int32_t* counts, void* ws, size_t ws_bytes, int32_t train, float* q_packed,
                           float* q_last, float* q_dense, dr4sr_stream_t stream) {
  const float* table = tv.table[0];
  DR4SR_TRY(check_cfg(c));
  if (!table || !params || !in_item_id || !tok_off || !row_seq || !counts || !ws) return DR4SR_EINVAL;
  Workspace w = carve(*c, ws);
  if (ws_bytes < w.bytes) return DR4SR_EWORKSPACE;
  cudaStream_t st = as_stream(stream);
  const int T = c->B * c->L, D = c->D, F = c->F;
  // a peer-sharded table is gathered by the fused forward only (the per-operator embedding kernel reads one local table)
  if (tv.world > 1 && !(fused_enabled() && fused_fwd_supported(c->L, D, F, c->n_head))) return DR4SR_EINVAL;
  const bool tr = train != 0;
  const float p = c->dropout_p;
  const LayerOffsets lo = layer_offsets(D, F);
  const float* pos = params;

  if (fused_enabled() && fused_fwd_supported(c->L, D, F, c->n_head)) {
    // one persistent kernel: gather + positions + dropout and every layer, per group of whole sequences.  Its two small
    // preparation kernels (weight images: parameters only; tiles: tok_off only) run side by side
    cudaStream_t sa = aux_fork(st);
    DR4SR_TRY(build_weight_images(*c, params, w, lo, sa));
    DR4SR_TRY(launch_fused_tiles(tok_off, c->B, w.fused_tiles, st));
    DR4SR_TRY(aux_join(sa, st));
    FusedFwdHost h{};
    h.table = tv; h.pos = pos; h.in_ids = in_item_id; h.tok_off = tok_off; h.row_seq = row_seq; h.tiles = w.fused_tiles;
    h.x0 = w.x0; h.B = c->B; h.L = c->L; h.n_layer = c->n_layer; h.ln_eps = c->ln_eps;
    h.d_embed = make_dropout(p, c->seed, c->step, SITE_EMBED, tr);
    for (int l = 0; l < c->n_layer; ++l) {
      const float* lp = params + (size_t)c->L * D + (size_t)l * lo.total;
      auto& y = w.layer[l];
      const auto& m = w.img[l];
      FusedLayerHost& d = h.layer[l];
      d.qkv = y.qkv; d.attn = y.attn; d.z1 = y.z1; d.st1 = y.st1; d.x1 = y.x1; d.hm = y.hm; d.gp = y.gp; d.z2 = y.z2; d.st2 = y.st2;
      d.x2 = (l == c->n_layer - 1 && q_packed) ? q_packed : y.x2;
      d.img[0] = m.in_f.hi; d.img[1] = m.in_f.lo; d.img[2] = m.out_f.hi; d.img[3] = m.out_f.lo;
      d.img[4] = m.w1_f.hi; d.img[5] = m.w1_f.lo; d.img[6] = m.w2_f.hi; d.img[7] = m.w2_f.lo;
      d.in_b = lp + lo.in_b; d.out_b = lp + lo.out_b; d.b1 = lp + lo.b1; d.b2 = lp + lo.b2;
      d.g1 = lp + lo.g1; d.be1 = lp + lo.be1; d.g2 = lp + lo.g2; d.be2 = lp + lo.be2;
      d.d_attn_p = make_dropout(p, c->seed, c->step, layer_site(SITE_ATTN_P, l), tr);
      d.d_attn_out = make_dropout(p, c->seed, c->step, layer_site(SITE_ATTN_OUT, l), tr);
      d.d_ffn_h = make_dropout(p, c->seed, c->step, layer_site(SITE_FFN_H, l), tr);
      d.d_ffn_out = make_dropout(p, c->seed, c->step, layer_site(SITE_FFN_OUT, l), tr);
    }
    DR4SR_TRY(launch_sasrec_fwd_fused(h, st));
    if (attn_tc_enabled() && attn_tc_supported(c->L, D, c->n_head))   // the per-op tcgen05 attention backward needs its own tiles
      DR4SR_TRY(launch_attn_tiles(tok_off, c->B, c->L, w.tile_first, st));
    const float* xl = (q_packed) ? q_packed : w.layer[c->n_layer - 1].x2;
    DR4SR_TRY(dr4sr_unpack_rows(xl, tok_off, c->B, c->L, D, q_last, q_dense, stream));
    return DR4SR_OK;
  }
  {  // K1: gather + positions + dropout
    const Dropout drop = make_dropout(p, c->seed, c->step, SITE_EMBED, tr);
    DR4SR_TRY(dr4sr_embed_fwd(table, pos, in_item_id, tok_off, row_seq, counts, c->B, c->L, D, tr ? p : 0.f, c->seed, c->step,
                              w.x0, stream));
    (void)drop;
  }
  DR4SR_TRY(build_weight_images(*c, params, w, lo, st));
  const bool attn_tc = attn_tc_enabled() && attn_tc_supported(c->L, D, c->n_head);
  if (attn_tc) DR4SR_TRY(launch_attn_tiles(tok_off, c->B, c->L, w.tile_first, st));
  if (attn_bwd_tc2_enabled() && attn_tc_supported(c->L, D, c->n_head) && D == 128 && c->n_head == 2)   // its backward runs on the greedy tiles
    DR4SR_TRY(launch_fused_tiles(tok_off, c->B, w.fused_tiles, st));
  const float* x = w.x0;
  for (int l = 0; l < c->n_layer; ++l) {
    const float* lp = params + (size_t)c->L * D + (size_t)l * lo.total;
    auto& y = w.layer[l];
    float* x2 = (l == c->n_layer - 1 && q_packed) ? q_packed : y.x2;
    {  // QKV projection
      GemmArgs g = gemm_args(x, D, lp + lo.in_w, D, y.qkv, 3 * D, T, 3 * D, D, counts);
      g.bias = lp + lo.in_b; g.tag = "gemm_qkv";
      DR4SR_TRY(gemm_nt(g, w.img[l].in_f, st));
    }
    if (attn_tc)
      DR4SR_TRY(launch_attn_tc_fwd(y.qkv, in_item_id, tok_off, row_seq, w.tile_first, y.attn, c->B, c->L, D, c->n_head,
                                   make_dropout(p, c->seed, c->step, layer_site(SITE_ATTN_P, l), tr), st));
    else
      DR4SR_TRY(launch_attn_fwd(y.qkv, in_item_id, tok_off, y.attn, c->B, c->L, D, c->n_head,
                                make_dropout(p, c->seed, c->step, layer_site(SITE_ATTN_P, l), tr), st));
    {  // out-proj + dropout + residual + LN1
      GemmArgs g = gemm_args(y.attn, D, lp + lo.out_w, D, y.x1, D, T, D, D, counts);
      g.bias = lp + lo.out_b; g.add = x; g.ldadd = D;
      g.dropE = make_dropout(p, c->seed, c->step, layer_site(SITE_ATTN_OUT, l), tr);
      g.gamma = lp + lo.g1; g.beta = lp + lo.be1; g.ln_eps = c->ln_eps; g.Z = y.z1; g.stats = y.st1; g.tag = "gemm_outproj_ln";
      DR4SR_TRY(gemm_ln(g, D, w.img[l].out_f, st));
    }
    {  // FFN up-projection (pre-activation kept for the backward)
      GemmArgs g = gemm_args(y.x1, D, lp + lo.w1, D, y.pre, F, T, F, D, counts);
      g.bias = lp + lo.b1; g.tag = "gemm_ffn1";
      DR4SR_TRY(gemm_nt(g, w.img[l].w1_f, st));
    }
    {  // gelu + dropout (prologue) -> down-projection + dropout + residual + LN2
      GemmArgs g = gemm_args(y.pre, F, lp + lo.w2, F, x2, D, T, D, F, counts);
      g.proA = PRO_GELU_DROP; g.dropA = make_dropout(p, c->seed, c->step, layer_site(SITE_FFN_H, l), tr);
      g.bias = lp + lo.b2; g.add = y.x1; g.ldadd = D;
      g.dropE = make_dropout(p, c->seed, c->step, layer_site(SITE_FFN_OUT, l), tr);
      g.gamma = lp + lo.g2; g.beta = lp + lo.be2; g.ln_eps = c->ln_eps; g.Z = y.z2; g.stats = y.st2; g.tag = "gemm_ffn2_ln";
      DR4SR_TRY(gemm_ln(g, D, w.img[l].w2_f, st));
    }
    x = x2;
  }
  DR4SR_TRY(dr4sr_unpack_rows(x, tok_off, c->B, c->L, D, q_last, q_dense, stream));
  return DR4SR_OK;
}

extern "C" int dr4sr_sasrec_fwd(const dr4sr_sasrec_cfg* c, const float* table, const float* params,
                                const int64_t* in_item_id, const int32_t* tok_off, const int32_t* row_seq,
                                const int32_t* counts, void* ws, size_t ws_bytes, int32_t train, float* q_packed,
                                float* q_last, float* q_dense, dr4sr_stream_t stream) {
  if (!c || !table) return DR4SR_EINVAL;
  return sasrec_fwd_impl(c, shard_view_local(table, nullptr, c->N), params, in_item_id, tok_off, row_seq, counts, ws, ws_bytes, train,
                         q_packed, q_last, q_dense, stream);
}

extern "C" int dr4sr_sasrec_fwd_sharded(const dr4sr_sasrec_cfg* c, const dr4sr_shard_map* map, const float* params,
                                        const int64_t* in_item_id, const int32_t* tok_off, const int32_t* row_seq,
                                        const int32_t* counts, void* ws, size_t ws_bytes, int32_t train, float* q_packed,
                                        float* q_last, float* q_dense, dr4sr_stream_t stream) {
  ShardView tv;
  if (!c || !shard_view_from(map, &tv) || tv.lo[tv.world] != c->N) return DR4SR_EINVAL;
  return sasrec_fwd_impl(c, tv, params, in_item_id, tok_off, row_seq, counts, ws, ws_bytes, train, q_packed, q_last, q_dense, stream);
}

static int sasrec_bwd_impl(const dr4sr_sasrec_cfg* c, const float* table, const float* params,
                           const int64_t* in_item_id, const int32_t* tok_off, const int32_t* row_seq,
                           const int32_t* counts, void* ws, size_t ws_bytes, float* dq_packed, float* grads,
                           float* dx0_packed, dr4sr_stream_t stream, bool join) {
  (void)table;
  DR4SR_TRY(check_cfg(c));
  if (!params || !in_item_id || !tok_off || !row_seq || !counts || !ws || !dq_packed || !grads || !dx0_packed) return DR4SR_EINVAL;
  Workspace w = carve(*c, ws);
  if (ws_bytes < w.bytes) return DR4SR_EWORKSPACE;
  cudaStream_t st = as_stream(stream);
  const int T = c->B * c->L, D = c->D, F = c->F;
  const float p = c->dropout_p;
  const bool tr = p > 0.f;   // the backward mirrors whatever masks the forward drew (p == 0 -> none)
  const LayerOffsets lo = layer_offsets(D, F);
  const size_t pw_in = 0, pw_out = (size_t)kSplit * 3 * D * D, pw_w1 = pw_out + (size_t)kSplit * D * D,
               pw_w2 = pw_w1 + (size_t)kSplit * F * D;

  // hm = dropout(gelu(pre)) was written by the fused forward (same predicate as in dr4sr_sasrec_fwd); the per-operator
  // forward does not materialise it, its FFN-down weight gradient recomputes GELU + dropout from `pre`
  const bool have_hm = fused_enabled() && fused_fwd_supported(c->L, D, F, c->n_head);
  SideStream& side = side_stream();
  cudaStream_t sw = side.ok ? side.s : st;            // weight-gradient stream (falls back to in-order if creation failed)
  float* gin = dq_packed;   // gradient w.r.t. the current layer's output
  for (int l = c->n_layer - 1; l >= 0; --l) {
    const float* lp = params + (size_t)c->L * D + (size_t)l * lo.total;
    float* lg = grads + (size_t)c->L * D + (size_t)l * lo.total;
    auto& y = w.layer[l];
    auto& s = w.bwd[l];
    const float* xin = l == 0 ? w.x0 : w.layer[l - 1].x2;
    const Dropout d_ffn_out = make_dropout(p, c->seed, c->step, layer_site(SITE_FFN_OUT, l), tr);
    const Dropout d_ffn_h = make_dropout(p, c->seed, c->step, layer_site(SITE_FFN_H, l), tr);
    const Dropout d_attn_out = make_dropout(p, c->seed, c->step, layer_site(SITE_ATTN_OUT, l), tr);
    const Dropout d_attn_p = make_dropout(p, c->seed, c->step, layer_site(SITE_ATTN_P, l), tr);

    if (fused_bwd_enabled() && fused_fwd_supported(c->L, D, F, c->n_head)) {
      // ---- position-wise half of the layer as one persistent kernel (main stream) ----
      FusedBwdFfnHost h{};
      h.gin = gin; h.z2 = y.z2; h.st2 = y.st2; h.gp = y.gp; h.z1 = y.z1; h.st1 = y.st1; h.gamma2 = lp + lo.g2; h.gamma1 = lp + lo.g1;
      h.img[0] = w.img[l].w2_b.hi; h.img[1] = w.img[l].w2_b.lo; h.img[2] = w.img[l].w1_b.hi; h.img[3] = w.img[l].w1_b.lo;
      h.img[4] = w.img[l].out_b.hi; h.img[5] = w.img[l].out_b.lo;
      h.g3 = s.g3; h.dpre = s.dpre; h.dx1 = s.dx1; h.g1 = s.g1; h.g2 = w.g2; h.part_ln2 = s.part_ln2; h.part_ln1 = s.part_ln1;
      h.counts = counts; h.T_cap = T;
      h.d_ffn_out = d_ffn_out; h.d_attn_out = d_attn_out;
      DR4SR_TRY(launch_sasrec_bwd_ffn_fused(h, st));
      const Dropout none = no_dropout();
      if (attn_bwd_tc2_enabled() && attn_tc_supported(c->L, D, c->n_head) && D == 128 && c->n_head == 2)
        DR4SR_TRY(launch_attn_bwd_tc2(y.qkv, w.g2, in_item_id, tok_off, row_seq, w.fused_tiles, fused_tiles_cap(c->B, c->L), s.dqkv, c->B,
                                      c->L, D, c->n_head, d_attn_p, st));
      else if (attn_tc_enabled() && attn_tc_supported(c->L, D, c->n_head))
        DR4SR_TRY(launch_attn_tc_bwd(y.qkv, w.g2, in_item_id, tok_off, row_seq, w.tile_first, s.dqkv, c->B, c->L, D, c->n_head,
                                     d_attn_p, st));
      else
        DR4SR_TRY(launch_attn_bwd(y.qkv, w.g2, in_item_id, tok_off, s.dqkv, c->B, c->L, D, c->n_head, d_attn_p, st));
      if (side.ok) {   // g3, dpre, g1 and dqkv are final: every weight gradient of the layer runs behind the attention backward
        if (cudaEventRecord(side.fork[2 * l + 1], st) != cudaSuccess || cudaStreamWaitEvent(sw, side.fork[2 * l + 1], 0) != cudaSuccess) {
          set_cuda_error(cudaGetLastError(), "backward fork");
          return DR4SR_ECUDA;
        }
      }
      DR4SR_TRY(launch_colsum(s.dpre, F, T, counts, s.part_cs_b1, sw));
      {
        tc::WgradTable tab{};
        tab.job[0] = tc::WgradJob{s.g3, D, PRO_DROPMASK, d_ffn_out, y.hm, F, PRO_NONE, none, D, F, s.part_w + pw_w2, 0};
        tab.job[1] = tc::WgradJob{s.dpre, F, PRO_NONE, none, y.x1, D, PRO_NONE, none, F, D, s.part_w + pw_w1, 0};
        tab.job[2] = tc::WgradJob{s.g1, D, PRO_DROPMASK, d_attn_out, y.attn, D, PRO_NONE, none, D, D, s.part_w + pw_out, 0};
        tab.count = 3; tab.T_cap = T; tab.tok_dev = counts; tab.n_split = kSplit;
        DR4SR_TRY(tc::launch_wgrad_tc(tab, sw));
      }
      {  // dx = dz1 + dqkv Win  (layer 0: times the embedding-dropout mask) -> this layer's own output buffer / dx0
        float* dst = l == 0 ? dx0_packed : s.gout;
        GemmArgs g = gemm_args(s.dqkv, 3 * D, lp + lo.in_w, D, dst, D, T, D, 3 * D, counts);
        g.add = s.g1; g.ldadd = D;
        if (l == 0) g.dropE = make_dropout(p, c->seed, c->step, SITE_EMBED, tr);
        g.tag = "gemm_bwd_dx";
        DR4SR_TRY(gemm_nn(g, w.img[l].in_b, st));
      }
      DR4SR_TRY(launch_colsum(s.dqkv, 3 * D, T, counts, s.part_cs_in, sw));
      {
        tc::WgradTable tab{};
        tab.job[0] = tc::WgradJob{s.dqkv, 3 * D, PRO_NONE, none, xin, D, PRO_NONE, none, 3 * D, D, s.part_w + pw_in, 0};
        tab.count = 1; tab.T_cap = T; tab.tok_dev = counts; tab.n_split = kSplit;
        DR4SR_TRY(tc::launch_wgrad_tc(tab, sw));
      }
      DR4SR_TRY(reduce_layer_partials(s.part_w, pw_in, pw_out, pw_w1, pw_w2, s.part_cs_in, s.part_cs_b1, s.part_ln1, s.part_ln2, lg, lo, D, F, sw));
      gin = s.gout;
      continue;
    }
    // ---- data-gradient chain (main stream) ----
    // LN2 backward: g3 = dz2 ; partials -> dgamma2, dbeta2, db2
    DR4SR_TRY(launch_ln_bwd(gin, y.z2, y.st2, lp + lo.g2, s.g3, s.part_ln2, D, T, counts, d_ffn_out, st));
    {  // dpre = ((dz2 * mask_out) W2) * mask_h * gelu'(pre)
      GemmArgs g = gemm_args(s.g3, D, lp + lo.w2, F, s.dpre, F, T, F, D, counts);
      g.proA = PRO_DROPMASK; g.dropA = d_ffn_out;
      g.epi = EPI_GELU_BWD; g.pre = y.pre; g.dropE = d_ffn_h; g.tag = "gemm_bwd_dpre";
      if (have_hm) { g.pre = nullptr; g.mul = y.gp; }       // the fused forward saved mask * gelu'(pre) instead of pre
      DR4SR_TRY(gemm_nn(g, w.img[l].w2_b, st));
    }
    {  // dx1 = dz2 + dpre W1   -> g0
      GemmArgs g = gemm_args(s.dpre, F, lp + lo.w1, D, w.g0, D, T, D, F, counts);
      g.add = s.g3; g.ldadd = D; g.tag = "gemm_bwd_dx1";
      DR4SR_TRY(gemm_nn(g, w.img[l].w1_b, st));
    }
    // LN1 backward: g1 = dz1 ; partials -> dgamma1, dbeta1, db_out
    DR4SR_TRY(launch_ln_bwd(w.g0, y.z1, y.st1, lp + lo.g1, s.g1, s.part_ln1, D, T, counts, d_attn_out, st));
    const bool wg_tc = tc_enabled() && tc::wgrad_supported(D, F) && tc::wgrad_supported(F, D) && tc::wgrad_supported(D, D) &&
                       tc::wgrad_supported(3 * D, D);
    const Dropout none = no_dropout();
    // Weight gradients that do not need dqkv (dW2, dW1, dWo, db1).  The attention backward holds a whole SM per CTA
    // (192 KB of shared memory, all 512 TMEM columns), so weight-gradient CTAs already resident on an SM keep it from
    // starting there: by default this work is queued on the side stream BEHIND the attention backward (it then overlaps
    // the dx GEMM and the next layer's position-wise kernels, which share SMs with it happily).
    static const bool early_fork = getenv("DR4SR_WGRAD_EARLY") != nullptr;
    auto side_A = [&]() -> int {
      DR4SR_TRY(launch_colsum(s.dpre, F, T, counts, s.part_cs_b1, sw));
      //   dW2[d,f]  = sum_m (dz2*mask_out)[m,d] * drop(gelu(pre))[m,f]      dW1[f,d]  = sum_m dpre[m,f] * x1[m,d]
      //   dWo[n,k]  = sum_m (dz1*mask)[m,n] * attn[m,k]                     dWin[j,d] = sum_m dqkv[m,j] * x[m,d]
      if (wg_tc) {
        tc::WgradTable tab{};
        tab.job[0] = have_hm ? tc::WgradJob{s.g3, D, PRO_DROPMASK, d_ffn_out, y.hm, F, PRO_NONE, none, D, F, s.part_w + pw_w2, 0}
                             : tc::WgradJob{s.g3, D, PRO_DROPMASK, d_ffn_out, y.pre, F, PRO_GELU_DROP, d_ffn_h, D, F, s.part_w + pw_w2, 0};
        tab.job[1] = tc::WgradJob{s.dpre, F, PRO_NONE, none, y.x1, D, PRO_NONE, none, F, D, s.part_w + pw_w1, 0};
        tab.job[2] = tc::WgradJob{s.g1, D, PRO_DROPMASK, d_attn_out, y.attn, D, PRO_NONE, none, D, D, s.part_w + pw_out, 0};
        tab.count = 3; tab.T_cap = T; tab.tok_dev = counts; tab.n_split = kSplit;
        DR4SR_TRY(tc::launch_wgrad_tc(tab, sw));
      }
      return DR4SR_OK;
    };
    if (early_fork) {
      if (side.ok) {   // dz2, dpre, dz1 are final
        if (cudaEventRecord(side.fork[2 * l], st) != cudaSuccess || cudaStreamWaitEvent(sw, side.fork[2 * l], 0) != cudaSuccess) {
          set_cuda_error(cudaGetLastError(), "backward fork");
          return DR4SR_ECUDA;
        }
      }
      DR4SR_TRY(side_A());
    }
    {  // d(attn) = (dz1 * mask) Wo -> g2
      GemmArgs g = gemm_args(s.g1, D, lp + lo.out_w, D, w.g2, D, T, D, D, counts);
      g.proA = PRO_DROPMASK; g.dropA = d_attn_out; g.tag = "gemm_bwd_dattn";
      DR4SR_TRY(gemm_nn(g, w.img[l].out_b, st));
    }
    if (attn_bwd_tc2_enabled() && attn_tc_supported(c->L, D, c->n_head) && D == 128 && c->n_head == 2)   // tiles were built by the forward
      DR4SR_TRY(launch_attn_bwd_tc2(y.qkv, w.g2, in_item_id, tok_off, row_seq, w.fused_tiles, fused_tiles_cap(c->B, c->L), s.dqkv, c->B,
                                    c->L, D, c->n_head, d_attn_p, st));
    else if (attn_tc_enabled() && attn_tc_supported(c->L, D, c->n_head))   // tiles were built by the forward (same workspace)
      DR4SR_TRY(launch_attn_tc_bwd(y.qkv, w.g2, in_item_id, tok_off, row_seq, w.tile_first, s.dqkv, c->B, c->L, D, c->n_head,
                                   d_attn_p, st));
    else
      DR4SR_TRY(launch_attn_bwd(y.qkv, w.g2, in_item_id, tok_off, s.dqkv, c->B, c->L, D, c->n_head, d_attn_p, st));
    if (side.ok) {   // everything the weight gradients of this layer read is now final: fork
      if (cudaEventRecord(side.fork[2 * l + 1], st) != cudaSuccess || cudaStreamWaitEvent(sw, side.fork[2 * l + 1], 0) != cudaSuccess) {
        set_cuda_error(cudaGetLastError(), "backward fork");
        return DR4SR_ECUDA;
      }
    }
    if (!early_fork) DR4SR_TRY(side_A());
    {  // dx = dz1 + dqkv Win  (layer 0: times the embedding-dropout mask) -> g0 / dx0
      float* dst = l == 0 ? dx0_packed : w.g0;
      GemmArgs g = gemm_args(s.dqkv, 3 * D, lp + lo.in_w, D, dst, D, T, D, 3 * D, counts);
      g.add = s.g1; g.ldadd = D;
      if (l == 0) g.dropE = make_dropout(p, c->seed, c->step, SITE_EMBED, tr);
      g.tag = "gemm_bwd_dx";
      DR4SR_TRY(gemm_nn(g, w.img[l].in_b, st));
    }

    // ---- remaining weight / bias gradients of this layer (side stream; overlaps the next layer's chain) ----
    DR4SR_TRY(launch_colsum(s.dqkv, 3 * D, T, counts, s.part_cs_in, sw));
    if (wg_tc) {
      tc::WgradTable tab{};
      tab.job[0] = tc::WgradJob{s.dqkv, 3 * D, PRO_NONE, none, xin, D, PRO_NONE, none, 3 * D, D, s.part_w + pw_in, 0};
      tab.count = 1; tab.T_cap = T; tab.tok_dev = counts; tab.n_split = kSplit;
      DR4SR_TRY(tc::launch_wgrad_tc(tab, sw));
    } else {
      {
        GemmArgs g = gemm_args(s.g3, D, y.pre, F, nullptr, F, D, F, T, counts);
        g.proA = PRO_DROPMASK; g.dropA = d_ffn_out; g.proB = PRO_GELU_DROP; g.dropB = d_ffn_h; g.tag = "gemm_wgrad_w2";
        DR4SR_TRY(gemm_tn(g, s.part_w + pw_w2, sw));
      }
      {
        GemmArgs g = gemm_args(s.dpre, F, y.x1, D, nullptr, D, F, D, T, counts);
        g.tag = "gemm_wgrad_w1";
        DR4SR_TRY(gemm_tn(g, s.part_w + pw_w1, sw));
      }
      {
        GemmArgs g = gemm_args(s.g1, D, y.attn, D, nullptr, D, D, D, T, counts);
        g.proA = PRO_DROPMASK; g.dropA = d_attn_out; g.tag = "gemm_wgrad_out";
        DR4SR_TRY(gemm_tn(g, s.part_w + pw_out, sw));
      }
      {
        GemmArgs g = gemm_args(s.dqkv, 3 * D, xin, D, nullptr, D, 3 * D, D, T, counts);
        g.tag = "gemm_wgrad_in";
        DR4SR_TRY(gemm_tn(g, s.part_w + pw_in, sw));
      }
    }
    DR4SR_TRY(reduce_layer_partials(s.part_w, pw_in, pw_out, pw_w1, pw_w2, s.part_cs_in, s.part_cs_b1, s.part_ln1, s.part_ln2, lg, lo, D, F, sw));
    gin = w.g0;
  }
  if (join) return dr4sr_sasrec_bwd_join(stream);
  return DR4SR_OK;
}

extern "C" int dr4sr_sasrec_bwd_join(dr4sr_stream_t stream) {
  SideStream& side = side_stream();
  if (side.ok) {   // the caller's stream sees every weight gradient after this point
    if (cudaEventRecord(side.join, side.s) != cudaSuccess || cudaStreamWaitEvent(as_stream(stream), side.join, 0) != cudaSuccess) {
      set_cuda_error(cudaGetLastError(), "backward join");
      return DR4SR_ECUDA;
    }
  }
  return DR4SR_OK;
}

extern "C" int dr4sr_sasrec_bwd(const dr4sr_sasrec_cfg* c, const float* table, const float* params,
                                const int64_t* in_item_id, const int32_t* tok_off, const int32_t* row_seq,
                                const int32_t* counts, void* ws, size_t ws_bytes, float* dq_packed, float* grads,
                                float* dx0_packed, dr4sr_stream_t stream) {
  return sasrec_bwd_impl(c, table, params, in_item_id, tok_off, row_seq, counts, ws, ws_bytes, dq_packed, grads, dx0_packed, stream, true);
}

extern "C" int dr4sr_sasrec_bwd_async(const dr4sr_sasrec_cfg* c, const float* table, const float* params,
                                      const int64_t* in_item_id, const int32_t* tok_off, const int32_t* row_seq,
                                      const int32_t* counts, void* ws, size_t ws_bytes, float* dq_packed, float* grads,
                                      float* dx0_packed, dr4sr_stream_t stream) {
  return sasrec_bwd_impl(c, table, params, in_item_id, tok_off, row_seq, counts, ws, ws_bytes, dq_packed, grads, dx0_packed, stream, false);
}

extern "C" int dr4sr_linear_fwd(const float* x, const float* w, const float* bias, float* y, int32_t M, int32_t N, int32_t K,
                                const int32_t* m_dev, dr4sr_stream_t stream) {
  if (!x || !w || !y || M <= 0 || N % 4 || K % 4) return DR4SR_EINVAL;
  GemmArgs g = gemm_args(x, K, w, K, y, N, M, N, K, m_dev);
  g.bias = bias;
  return gemm_nt(g, Img{nullptr, nullptr}, as_stream(stream));
}

