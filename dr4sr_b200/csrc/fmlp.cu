// fmlp.cu -- FMLP encoder (filter-enhanced MLP) forward / backward on dense [B, L, D] activations.
//
// Replaces FMLP.add_position_embedding + FMLPEncoder (reference model/fmlp.py:18-39,
// module/layers.py:740-808).  Spec: SURVEY.md Appendix C.1, C.4.
//   x0 = drop(LN(E[ids] + P))                                    (inputs are PRE-padded: every slot is live)
//   per layer:  f = irfft(rfft(x, time, ortho) * W, n=L, ortho);  y = LN(drop(f) + x)
//               z = LN(drop(W2 gelu(W1 y + b1) + b2) + y)         (inner width 4D)
//   query = z_last_layer[:, L-1, :]
// The learnable spectral filter is a per-channel circular convolution along time,
//   f[t,d] = sum_s h_d[(t - s) mod L] x[s,d],   h_d = irfft(W[:,d], n=L, norm='backward'),
// so the taps h are rebuilt from the complex weights once per step (26-point inverse real DFT per
// channel) and the filter runs as 50 x 50 register-resident FMAs per (sequence, channel): the spectrum
// never exists in memory, and irfft's silent drop of Im(W[0]), Im(W[L/2]) is inherited by construction.
// The dense layers reuse the tcgen05 / FFMA GEMMs of the SASRec path (dense.cuh).
#include "dense.cuh"

namespace dr4sr {
namespace {

enum : uint32_t { SITE_FILTER_OUT = SITE_ATTN_OUT };
constexpr int kFilterChunks = 4 * kNumSMs;   // sequence chunks of the tap-gradient reduction (4 CTAs of 4 warps per SM)

struct FmlpOffsets {   // floats inside one layer's slice of the flat parameter buffer (state_dict order)
  size_t cw, fg, fb, w1, b1, w2, b2, ig, ib, total;
};
FmlpOffsets fmlp_offsets(int L, int D) {
  FmlpOffsets o;
  size_t p = 0;
  o.cw = p; p += (size_t)(L / 2 + 1) * D * 2;
  o.fg = p; p += D;
  o.fb = p; p += D;
  o.w1 = p; p += (size_t)4 * D * D;
  o.b1 = p; p += (size_t)4 * D;
  o.w2 = p; p += (size_t)4 * D * D;
  o.b2 = p; p += D;
  o.ig = p; p += D;
  o.ib = p; p += D;
  o.total = p;
  return o;
}
inline size_t fmlp_head_params(int L, int D) { return (size_t)L * D + 2 * (size_t)D; }   // positions, LayerNorm

struct FmlpWs {
  float *z0, *st0, *x0;
  struct Layer { float *z1, *st1, *y, *pre, *z2, *st2, *x2, *taps; Img w1_f, w2_f, w1_b, w2_b; } layer[8];
  float *g0, *g1, *g3, *dpre;
  float *dh_part, *dh;            // [kFilterChunks][L][D], [L][D]
  float *part_w;                  // [kSplit][2 * 4D * D]
  float *part_ln_a, *part_ln_b;   // [kLnBwdBlocks][3D]
  float *part_cs_b1;              // [kColsumBlocks][4D]
  int32_t *tok_off, *counts;      // dense index: tok_off[b] = b * L, counts[0] = B * L
  size_t bytes;
};

FmlpWs carve(const dr4sr_fmlp_cfg& c, void* base) {
  FmlpWs w{};
  float* p = reinterpret_cast<float*>(base);
  size_t off = 0;
  const size_t T = (size_t)c.B * c.L, D = c.D, F = 4 * D;
  auto take = [&](size_t n) { float* r = p ? p + off : nullptr; off += ws_align(n); return r; };
  auto take_img = [&](size_t elems) { Img im; im.hi = reinterpret_cast<uint16_t*>(take((elems + 1) / 2)); im.lo = reinterpret_cast<uint16_t*>(take((elems + 1) / 2)); return im; };
  w.z0 = take(T * D); w.st0 = take(T * 2); w.x0 = take(T * D);
  for (int l = 0; l < c.n_layer; ++l) {
    auto& y = w.layer[l];
    y.z1 = take(T * D); y.st1 = take(T * 2); y.y = take(T * D); y.pre = take(T * F); y.z2 = take(T * D); y.st2 = take(T * 2);
    y.x2 = take(T * D); y.taps = take((size_t)c.L * D);
    y.w1_f = take_img(F * D); y.w2_f = take_img(D * F); y.w1_b = take_img(F * D); y.w2_b = take_img(D * F);
  }
  w.g0 = take(T * D); w.g1 = take(T * D); w.g3 = take(T * D); w.dpre = take(T * F);
  w.dh_part = take((size_t)kFilterChunks * c.L * D); w.dh = take((size_t)c.L * D);
  w.part_w = take((size_t)kSplit * 2 * F * D);
  w.part_ln_a = take((size_t)kLnBwdBlocks * 3 * D); w.part_ln_b = take((size_t)kLnBwdBlocks * 3 * D);
  w.part_cs_b1 = take((size_t)kColsumBlocks * F);
  w.tok_off = reinterpret_cast<int32_t*>(take((size_t)c.B + 1));
  w.counts = reinterpret_cast<int32_t*>(take(4));
  w.bytes = off * sizeof(float);
  return w;
}

int check_cfg(const dr4sr_fmlp_cfg* c) {
  if (!c) return DR4SR_EINVAL;
  if (c->B <= 0 || c->L != 50 || c->n_layer < 1 || c->n_layer > 8) return DR4SR_EINVAL;   // filter kernels are unrolled for L = 50
  if (c->D != 64 && c->D != 128) return DR4SR_EINVAL;
  if (c->dropout_p < 0.f || c->dropout_p >= 1.f) return DR4SR_EINVAL;
  return DR4SR_OK;
}

__global__ void dense_index_kernel(int B, int L, int32_t* tok_off, int32_t* counts) {
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b <= B; b += gridDim.x * blockDim.x) tok_off[b] = b * L;
  if (blockIdx.x == 0 && threadIdx.x == 0) { counts[0] = B * L; counts[1] = 0; counts[2] = 0; counts[3] = 0; }
}

// z0[b,t,:] = E[ids[b,t]] + P[t]   (dense rows; LayerNorm + dropout follow in ln_fwd)
__global__ void __launch_bounds__(256) embed_dense_kernel(const float* __restrict__ table, const float* __restrict__ pos,
                                                          const int64_t* __restrict__ ids, int T, int L, int D, float* __restrict__ z0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = blockIdx.x * 8 + warp; row < T; row += gridDim.x * 8) {
    const float* e = table + (size_t)ids[row] * D;
    const float* p = pos + (size_t)(row % L) * D;
    for (int c = lane * 4; c < D; c += 128) {
      const float4 a = *reinterpret_cast<const float4*>(e + c), b = *reinterpret_cast<const float4*>(p + c);
      *reinterpret_cast<float4*>(z0 + (size_t)row * D + c) = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
  }
}

// taps[tau, d] = irfft(W[:, d], n = L, norm = 'backward')[tau]; cw is [L/2+1, D, 2] (re, im)
__global__ void __launch_bounds__(128) filter_taps_kernel(const float* __restrict__ cw, int L, int D, float* __restrict__ taps) {
  const int d = blockIdx.y * 128 + threadIdx.x, tau = blockIdx.x;
  if (d >= D) return;
  const int half = L / 2;
  float acc = cw[(size_t)0 * D * 2 + d * 2];
  for (int k = 1; k < half; ++k) {
    float sn, cs;
    sincospif(2.0f * (float)((k * tau) % L) / (float)L, &sn, &cs);
    acc += 2.0f * (cw[((size_t)k * D + d) * 2] * cs - cw[((size_t)k * D + d) * 2 + 1] * sn);
  }
  acc += cw[((size_t)half * D + d) * 2] * ((tau & 1) ? -1.0f : 1.0f);
  taps[(size_t)tau * D + d] = acc / (float)L;
}

// d(cw) from d(taps): dRe_k = c_k/L sum_tau dh cos(2 pi k tau / L); dIm_k = -c_k/L sum_tau dh sin(...); c_0 = c_{L/2} = 1, else 2
__global__ void __launch_bounds__(128) filter_wgrad_kernel(const float* __restrict__ dh, int L, int D, float* __restrict__ dcw) {
  const int d = blockIdx.y * 128 + threadIdx.x, k = blockIdx.x;
  if (d >= D) return;
  const int half = L / 2;
  float re = 0.f, im = 0.f;
  for (int tau = 0; tau < L; ++tau) {
    float sn, cs;
    sincospif(2.0f * (float)((k * tau) % L) / (float)L, &sn, &cs);
    const float g = dh[(size_t)tau * D + d];
    re = fmaf(g, cs, re);
    im = fmaf(g, sn, im);
  }
  const float ck = (k == 0 || k == half) ? 1.0f : 2.0f;
  dcw[((size_t)k * D + d) * 2] = ck * re / (float)L;
  dcw[((size_t)k * D + d) * 2 + 1] = (k == 0 || k == half) ? 0.0f : -ck * im / (float)L;
}

// z1 = drop(circconv(h, x)) + x : thread = (sequence, channel); x and h live in registers, indices are static
template <int L>
__global__ void __launch_bounds__(128) filter_fwd_kernel(const float* __restrict__ x, const float* __restrict__ taps, int D,
                                                         Dropout drop, float* __restrict__ z1) {
  const int b = blockIdx.x, d = blockIdx.y * 128 + threadIdx.x;
  if (d >= D) return;
  float xs[L], h[L];
#pragma unroll
  for (int s = 0; s < L; ++s) { xs[s] = x[((size_t)b * L + s) * D + d]; h[s] = taps[(size_t)s * D + d]; }
#pragma unroll
  for (int t = 0; t < L; ++t) {
    float acc = 0.f;
#pragma unroll
    for (int s = 0; s < L; ++s) acc = fmaf(h[(t - s + L) % L], xs[s], acc);
    const size_t o = ((size_t)b * L + t) * D + d;
    z1[o] = drop.apply(acc, (uint32_t)o) + xs[t];
  }
}

// dx = dz1 + corr(h, dz1 * mask):  dx[s] = dz1[s] + sum_t h[(t - s) mod L] * df[t]
template <int L>
__global__ void __launch_bounds__(128) filter_bwd_dx_kernel(const float* __restrict__ dz1, const float* __restrict__ taps, int D,
                                                            Dropout drop, float* __restrict__ dx) {
  const int b = blockIdx.x, d = blockIdx.y * 128 + threadIdx.x;
  if (d >= D) return;
  float df[L], h[L], res[L];
#pragma unroll
  for (int t = 0; t < L; ++t) {
    const size_t o = ((size_t)b * L + t) * D + d;
    res[t] = dz1[o];
    df[t] = res[t] * drop.factor((uint32_t)o);
    h[t] = taps[(size_t)t * D + d];
  }
#pragma unroll
  for (int s = 0; s < L; ++s) {
    float acc = res[s];
#pragma unroll
    for (int t = 0; t < L; ++t) acc = fmaf(h[(t - s + L) % L], df[t], acc);
    dx[((size_t)b * L + s) * D + d] = acc;
  }
}

// dh[tau] = sum_b sum_t df[b,t] * x[b,(t - tau) mod L]; a CTA owns a slice of sequences and writes its partial
template <int L>
__global__ void __launch_bounds__(128) filter_bwd_dh_kernel(const float* __restrict__ dz1, const float* __restrict__ x, int B, int D,
                                                            Dropout drop, float* __restrict__ partial) {
  const int d = blockIdx.y * 128 + threadIdx.x;
  if (d >= D) return;
  float dh[L];
#pragma unroll
  for (int i = 0; i < L; ++i) dh[i] = 0.f;
  const int per = (B + gridDim.x - 1) / gridDim.x;
  const int b0 = min(B, (int)blockIdx.x * per), b1 = min(B, b0 + per);
  for (int b = b0; b < b1; ++b) {
    float df[L], xs[L];
#pragma unroll
    for (int t = 0; t < L; ++t) {
      const size_t o = ((size_t)b * L + t) * D + d;
      df[t] = dz1[o] * drop.factor((uint32_t)o);
      xs[t] = x[o];
    }
#pragma unroll
    for (int tau = 0; tau < L; ++tau) {
      float acc = dh[tau];
#pragma unroll
      for (int t = 0; t < L; ++t) acc = fmaf(df[t], xs[(t - tau + L) % L], acc);
      dh[tau] = acc;
    }
  }
#pragma unroll
  for (int tau = 0; tau < L; ++tau) partial[((size_t)blockIdx.x * L + tau) * D + d] = dh[tau];
}

__global__ void __launch_bounds__(256) take_last_kernel(const float* __restrict__ x, int B, int L, int D, float* __restrict__ out) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < B * (D / 4); e += gridDim.x * blockDim.x) {
    const int b = e / (D / 4), c = (e % (D / 4)) * 4;
    *reinterpret_cast<float4*>(out + (size_t)b * D + c) = *reinterpret_cast<const float4*>(x + ((size_t)b * L + L - 1) * D + c);
  }
}
// g[b, t, :] = (t == L-1) ? dq[b, :] : 0
__global__ void __launch_bounds__(256) put_last_kernel(const float* __restrict__ dq, int B, int L, int D, float* __restrict__ g) {
  const size_t total = (size_t)B * L * (D / 4);
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t row = e / (D / 4);
    const int c = (int)(e % (D / 4)) * 4;
    const int b = (int)(row / L), t = (int)(row % L);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t == L - 1) v = *reinterpret_cast<const float4*>(dq + (size_t)b * D + c);
    *reinterpret_cast<float4*>(g + row * D + c) = v;
  }
}

int build_images(const dr4sr_fmlp_cfg& c, const float* params, const FmlpWs& w, const FmlpOffsets& lo, cudaStream_t st) {
  if (!tc_enabled()) return DR4SR_OK;
  const int D = c.D, F = 4 * D;
  tc::ImageTable tab{};
  for (int l = 0; l < c.n_layer; ++l) {
    const float* lp = params + fmlp_head_params(c.L, D) + (size_t)l * lo.total;
    const auto& m = w.layer[l];
    auto add = [&](const float* src, int ld, int N, int K, int tr, const Img& im) {
      if (tc::tc_supported(N, K, false)) tab.job[tab.count++] = tc::ImageJob{src, ld, N, K, tr, im.hi, im.lo};
    };
    add(lp + lo.w1, D, F, D, 0, m.w1_f);
    add(lp + lo.w2, F, D, F, 0, m.w2_f);
    add(lp + lo.w1, D, D, F, 1, m.w1_b);     // dy  = dpre W1 : B'[d][f] = W1[f][d]
    add(lp + lo.w2, F, F, D, 1, m.w2_b);     // dh  = dz2 W2  : B'[f][d] = W2[d][f]
    if (tab.count + 4 > tc::kMaxImageJobs || l == c.n_layer - 1) {
      DR4SR_TRY(tc::launch_weight_images(tab, st));
      tab.count = 0;
    }
  }
  return DR4SR_OK;
}

}  // namespace
}  // namespace dr4sr

using namespace dr4sr;

extern "C" size_t dr4sr_fmlp_param_count(const dr4sr_fmlp_cfg* c) {
  if (check_cfg(c) != DR4SR_OK) return 0;
  return fmlp_head_params(c->L, c->D) + (size_t)c->n_layer * fmlp_offsets(c->L, c->D).total;
}

extern "C" size_t dr4sr_fmlp_workspace_bytes(const dr4sr_fmlp_cfg* c) {
  if (check_cfg(c) != DR4SR_OK) return 0;
  return carve(*c, nullptr).bytes;
}

extern "C" int dr4sr_fmlp_fwd(const dr4sr_fmlp_cfg* c, const float* table, const float* params, const int64_t* in_item_id,
                              void* ws, size_t ws_bytes, int32_t train, float* q_last, dr4sr_stream_t stream) {
  DR4SR_TRY(check_cfg(c));
  if (!table || !params || !in_item_id || !ws || !q_last) return DR4SR_EINVAL;
  FmlpWs w = carve(*c, ws);
  if (ws_bytes < w.bytes) return DR4SR_EWORKSPACE;
  cudaStream_t st = as_stream(stream);
  const int T = c->B * c->L, D = c->D, F = 4 * D, L = c->L;
  const bool tr = train != 0;
  const float p = c->dropout_p;
  const FmlpOffsets lo = fmlp_offsets(L, D);
  const float* pos = params;
  const float* ln0_g = params + (size_t)L * D;
  const float* ln0_b = ln0_g + D;
  const int row_blocks = ceil_div(T, 8) < 8 * kNumSMs ? ceil_div(T, 8) : 8 * kNumSMs;

  {
    ProfScope prof("fmlp_dense_index", st);
    dense_index_kernel<<<ceil_div(c->B + 1, 256), 256, 0, st>>>(c->B, L, w.tok_off, w.counts);
    DR4SR_LAUNCH_CHECK("dense_index_kernel");
  }
  {
    ProfScope prof("fmlp_embed", st);
    embed_dense_kernel<<<row_blocks, 256, 0, st>>>(table, pos, in_item_id, T, L, D, w.z0);
    DR4SR_LAUNCH_CHECK("embed_dense_kernel");
  }
  DR4SR_TRY(launch_ln_fwd(w.z0, ln0_g, ln0_b, c->ln_eps, w.x0, w.st0, D, T, nullptr, make_dropout(p, c->seed, c->step, SITE_EMBED, tr), st));
  DR4SR_TRY(build_images(*c, params, w, lo, st));

  const float* x = w.x0;
  for (int l = 0; l < c->n_layer; ++l) {
    const float* lp = params + fmlp_head_params(L, D) + (size_t)l * lo.total;
    auto& y = w.layer[l];
    {
      ProfScope prof("fmlp_filter_taps", st);
      filter_taps_kernel<<<dim3(L, ceil_div(D, 128)), 128, 0, st>>>(lp + lo.cw, L, D, y.taps);
      DR4SR_LAUNCH_CHECK("filter_taps_kernel");
    }
    {
      ProfScope prof("fmlp_filter_fwd", st);
      filter_fwd_kernel<50><<<dim3(c->B, ceil_div(D, 128)), 128, 0, st>>>(x, y.taps, D, make_dropout(p, c->seed, c->step, layer_site(SITE_FILTER_OUT, l), tr), y.z1);
      DR4SR_LAUNCH_CHECK("filter_fwd_kernel");
    }
    DR4SR_TRY(launch_ln_fwd(y.z1, lp + lo.fg, lp + lo.fb, c->ln_eps, y.y, y.st1, D, T, nullptr, no_dropout(), st));
    {  // dense_1 (pre-activation kept for the backward)
      GemmArgs g = gemm_args(y.y, D, lp + lo.w1, D, y.pre, F, T, F, D, w.counts);
      g.bias = lp + lo.b1; g.tag = "fmlp_gemm_ffn1";
      DR4SR_TRY(gemm_nt(g, y.w1_f, st));
    }
    {  // gelu (prologue) -> dense_2 + dropout + residual + LayerNorm
      GemmArgs g = gemm_args(y.pre, F, lp + lo.w2, F, y.x2, D, T, D, F, w.counts);
      g.proA = PRO_GELU_DROP; g.dropA = no_dropout();
      g.bias = lp + lo.b2; g.add = y.y; g.ldadd = D;
      g.dropE = make_dropout(p, c->seed, c->step, layer_site(SITE_FFN_OUT, l), tr);
      g.gamma = lp + lo.ig; g.beta = lp + lo.ib; g.ln_eps = c->ln_eps; g.Z = y.z2; g.stats = y.st2; g.tag = "fmlp_gemm_ffn2_ln";
      DR4SR_TRY(gemm_ln(g, D, y.w2_f, st));
    }
    x = y.x2;
  }
  {
    ProfScope prof("fmlp_take_last", st);
    take_last_kernel<<<ceil_div((int64_t)c->B * (D / 4), 256), 256, 0, st>>>(x, c->B, L, D, q_last);
    DR4SR_LAUNCH_CHECK("take_last_kernel");
  }
  return DR4SR_OK;
}

extern "C" int dr4sr_fmlp_bwd(const dr4sr_fmlp_cfg* c, const float* table, const float* params, const int64_t* in_item_id,
                              void* ws, size_t ws_bytes, const float* dq_last, float* grads, float* dz0_dense,
                              dr4sr_stream_t stream) {
  (void)table; (void)in_item_id;
  DR4SR_TRY(check_cfg(c));
  if (!params || !ws || !dq_last || !grads || !dz0_dense) return DR4SR_EINVAL;
  FmlpWs w = carve(*c, ws);
  if (ws_bytes < w.bytes) return DR4SR_EWORKSPACE;
  cudaStream_t st = as_stream(stream);
  const int T = c->B * c->L, D = c->D, F = 4 * D, L = c->L;
  const float p = c->dropout_p;
  const bool tr = p > 0.f;
  const FmlpOffsets lo = fmlp_offsets(L, D);
  const size_t pw_w1 = 0, pw_w2 = (size_t)kSplit * F * D;

  {
    ProfScope prof("fmlp_put_last", st);
    put_last_kernel<<<4 * kNumSMs, 256, 0, st>>>(dq_last, c->B, L, D, w.g0);
    DR4SR_LAUNCH_CHECK("put_last_kernel");
  }
  for (int l = c->n_layer - 1; l >= 0; --l) {
    const float* lp = params + fmlp_head_params(L, D) + (size_t)l * lo.total;
    float* lg = grads + fmlp_head_params(L, D) + (size_t)l * lo.total;
    auto& y = w.layer[l];
    const float* xin = l == 0 ? w.x0 : w.layer[l - 1].x2;
    const Dropout d_ffn_out = make_dropout(p, c->seed, c->step, layer_site(SITE_FFN_OUT, l), tr);
    const Dropout d_filter = make_dropout(p, c->seed, c->step, layer_site(SITE_FILTER_OUT, l), tr);

    // LayerNorm (intermediate) backward: g3 = dz2 ; partials -> d gamma, d beta, d b2
    DR4SR_TRY(launch_ln_bwd(w.g0, y.z2, y.st2, lp + lo.ig, w.g3, w.part_ln_a, D, T, w.counts, d_ffn_out, st));
    {  // dpre = ((dz2 * mask) W2) * gelu'(pre)
      GemmArgs g = gemm_args(w.g3, D, lp + lo.w2, F, w.dpre, F, T, F, D, w.counts);
      g.proA = PRO_DROPMASK; g.dropA = d_ffn_out;
      g.epi = EPI_GELU_BWD; g.pre = y.pre; g.dropE = no_dropout(); g.tag = "fmlp_gemm_bwd_dpre";
      DR4SR_TRY(gemm_nn(g, y.w2_b, st));
    }
    DR4SR_TRY(launch_colsum(w.dpre, F, T, w.counts, w.part_cs_b1, st));
    {  // dy = dz2 + dpre W1 -> g1
      GemmArgs g = gemm_args(w.dpre, F, lp + lo.w1, D, w.g1, D, T, D, F, w.counts);
      g.add = w.g3; g.ldadd = D; g.tag = "fmlp_gemm_bwd_dy";
      DR4SR_TRY(gemm_nn(g, y.w1_b, st));
    }
    // dW2[d,f] = sum_m (dz2*mask)[m,d] gelu(pre)[m,f] ; dW1[f,d] = sum_m dpre[m,f] y[m,d]
    if (tc_enabled() && tc::wgrad_supported(D, F) && tc::wgrad_supported(F, D)) {
      tc::WgradTable tab{};
      tab.job[0] = tc::WgradJob{w.g3, D, PRO_DROPMASK, d_ffn_out, y.pre, F, PRO_GELU_DROP, no_dropout(), D, F, w.part_w + pw_w2, 0};
      tab.job[1] = tc::WgradJob{w.dpre, F, PRO_NONE, no_dropout(), y.y, D, PRO_NONE, no_dropout(), F, D, w.part_w + pw_w1, 0};
      tab.count = 2; tab.T_cap = T; tab.tok_dev = w.counts; tab.n_split = kSplit;
      DR4SR_TRY(tc::launch_wgrad_tc(tab, st));
    } else {
      {
        GemmArgs g = gemm_args(w.g3, D, y.pre, F, nullptr, F, D, F, T, w.counts);
        g.proA = PRO_DROPMASK; g.dropA = d_ffn_out; g.proB = PRO_GELU_DROP; g.dropB = no_dropout(); g.tag = "fmlp_wgrad_w2";
        DR4SR_TRY(gemm_tn(g, w.part_w + pw_w2, st));
      }
      {
        GemmArgs g = gemm_args(w.dpre, F, y.y, D, nullptr, D, F, D, T, w.counts);
        g.tag = "fmlp_wgrad_w1";
        DR4SR_TRY(gemm_tn(g, w.part_w + pw_w1, st));
      }
    }
    // LayerNorm (filter) backward: g3 = dz1 ; partials -> d gamma, d beta
    DR4SR_TRY(launch_ln_bwd(w.g1, y.z1, y.st1, lp + lo.fg, w.g3, w.part_ln_b, D, T, w.counts, no_dropout(), st));
    {  // filter backward: taps gradient (partials per sequence slice) and input gradient
      ProfScope prof("fmlp_filter_bwd_dh", st);
      filter_bwd_dh_kernel<50><<<dim3(kFilterChunks, ceil_div(D, 128)), 128, 0, st>>>(w.g3, xin, c->B, D, d_filter, w.dh_part);
      DR4SR_LAUNCH_CHECK("filter_bwd_dh_kernel");
    }
    {
      ProfScope prof("fmlp_filter_bwd_dx", st);
      filter_bwd_dx_kernel<50><<<dim3(c->B, ceil_div(D, 128)), 128, 0, st>>>(w.g3, y.taps, D, d_filter, w.g0);
      DR4SR_LAUNCH_CHECK("filter_bwd_dx_kernel");
    }
    {
      ReduceTable tab{};
      int k = 0;
      auto seg = [&](const float* src, float* dst, int ns, int64_t stride, int n) { tab.seg[k++] = ReduceSeg{src, dst, ns, stride, n}; };
      seg(w.part_w + pw_w1, lg + lo.w1, kSplit, (int64_t)F * D, F * D);
      seg(w.part_w + pw_w2, lg + lo.w2, kSplit, (int64_t)D * F, D * F);
      seg(w.part_cs_b1, lg + lo.b1, kColsumBlocks, F, F);
      seg(w.part_ln_a, lg + lo.ig, kLnBwdBlocks, 3 * D, D);
      seg(w.part_ln_a + D, lg + lo.ib, kLnBwdBlocks, 3 * D, D);
      seg(w.part_ln_a + 2 * D, lg + lo.b2, kLnBwdBlocks, 3 * D, D);
      seg(w.part_ln_b, lg + lo.fg, kLnBwdBlocks, 3 * D, D);
      seg(w.part_ln_b + D, lg + lo.fb, kLnBwdBlocks, 3 * D, D);
      seg(w.dh_part, w.dh, kFilterChunks, (int64_t)L * D, L * D);
      tab.count = k;
      DR4SR_TRY(launch_reduce_segments(tab, st));
    }
    {
      ProfScope prof("fmlp_filter_wgrad", st);
      filter_wgrad_kernel<<<dim3(L / 2 + 1, ceil_div(D, 128)), 128, 0, st>>>(w.dh, L, D, lg + lo.cw);
      DR4SR_LAUNCH_CHECK("filter_wgrad_kernel");
    }
  }
  // embedding LayerNorm backward (through the embedding dropout): dz0 = d(E[ids] + P)
  DR4SR_TRY(launch_ln_bwd(w.g0, w.z0, w.st0, params + (size_t)L * D, dz0_dense, w.part_ln_a, D, T, w.counts, no_dropout(), st,
                          make_dropout(p, c->seed, c->step, SITE_EMBED, tr)));
  {
    ReduceTable tab{};
    tab.seg[0] = ReduceSeg{w.part_ln_a, grads + (size_t)L * D, kLnBwdBlocks, 3 * D, D};
    tab.seg[1] = ReduceSeg{w.part_ln_a + D, grads + (size_t)L * D + D, kLnBwdBlocks, 3 * D, D};
    tab.count = 2;
    DR4SR_TRY(launch_reduce_segments(tab, st));
  }
  return DR4SR_OK;
}
