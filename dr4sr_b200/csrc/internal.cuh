// internal.cuh -- launcher prototypes shared between the translation units of libdr4sr.
#pragma once
#include "common.cuh"

namespace dr4sr {

// auxiliary stream for short independent kernels  [api.cu]
cudaStream_t aux_fork(cudaStream_t st);
int aux_join(cudaStream_t aux, cudaStream_t st);
// background stream for one longer kernel that overlaps the main stream until bg_join  [api.cu]
cudaStream_t bg_fork(cudaStream_t st);
int bg_mark(cudaStream_t bg, cudaStream_t st);
int bg_join(cudaStream_t st);

// positional gradient (deterministic column reduction of dx0)  [loss_table.cu]
int launch_pos_grad(const float* dx0_packed, const int32_t* tok_off, int B, int L, int D, float* pos_grad, void* ws, cudaStream_t sa);

// GRU recurrence on tcgen05 (hidden size 256)  [gru_tc.cu]
bool gru_tc_supported(int H);
int launch_gru_fwd_tc(const float* gi, const uint16_t* whh_hi, const uint16_t* whh_lo, const int32_t* tok_off, const int32_t* order, int B,
                      float* h, float* hprev, float* gates, cudaStream_t st);

constexpr int kLnBwdBlocks = 2 * kNumSMs;   // CTAs (= column-partial slices) of the LayerNorm backward kernels

// attention over packed rows, one CTA per (sequence, head)  [attention.cu]
int launch_attn_fwd(const float* qkv, const int64_t* in_ids, const int32_t* tok_off, float* out, int B, int L, int D,
                    int n_head, Dropout drop, cudaStream_t st);
int launch_attn_bwd(const float* qkv, const float* d_out, const int64_t* in_ids, const int32_t* tok_off, float* d_qkv,
                    int B, int L, int D, int n_head, Dropout drop, cudaStream_t st);

// tcgen05 attention (head_dim 64): tiles of whole sequences with <= 128 rows  [attention_tc.cu]
bool attn_tc_supported(int L, int D, int n_head);
int attn_tc_num_tiles(int B, int L);
int launch_attn_tiles(const int32_t* tok_off, int B, int L, int32_t* tile_first, cudaStream_t st);
int launch_attn_tc_fwd(const float* qkv, const int64_t* in_ids, const int32_t* tok_off, const int32_t* row_seq,
                       const int32_t* tile_first, float* out, int B, int L, int D, int n_head, Dropout drop, cudaStream_t st);
int launch_attn_tc_bwd(const float* qkv, const float* d_out, const int64_t* in_ids, const int32_t* tok_off, const int32_t* row_seq,
                       const int32_t* tile_first, float* d_qkv, int B, int L, int D, int n_head, Dropout drop, cudaStream_t st);

// persistent tcgen05 attention backward over the greedy whole-sequence tiles of the fused forward  [attention_bwd_tc.cu]
int launch_attn_bwd_tc2(const float* qkv, const float* d_out, const int64_t* in_ids, const int32_t* tok_off, const int32_t* row_seq,
                        const int32_t* tiles, int tiles_cap, float* d_qkv, int B, int L, int D, int n_head, Dropout drop, cudaStream_t st);

// full-catalog scoring q @ E^T on tcgen05 with the top-k selection fused into the epilogue (no B x N logits)  [logits_tc.cu]
bool logits_tc_supported(int D, int32_t k);
size_t logits_tc_workspace_bytes(int B, int64_t N, int D, int H);
int launch_logits_topk_tc(const float* q, const float* table, const uint8_t* item_dead, const int64_t* user_hist, int B, int D, int64_t N,
                          int H, int k, float* out_scores, int64_t* out_ids, void* ws, size_t ws_bytes, cudaStream_t st);

// whole-encoder forward as one persistent tcgen05 kernel (D = F = 128, 2 heads)  [fused_fwd.cu]
struct FusedLayerHost {
  float *qkv, *attn, *z1, *st1, *x1, *hm, *gp, *z2, *st2, *x2;   // hm = mask * gelu(pre), gp = mask * gelu'(pre) (pre itself is not kept)
  const uint16_t* img[8];   // in_hi, in_lo, out_hi, out_lo, w1_hi, w1_lo, w2_hi, w2_lo (forward weight images)
  const float *in_b, *out_b, *b1, *b2, *g1, *be1, *g2, *be2;
  Dropout d_attn_p, d_attn_out, d_ffn_h, d_ffn_out;
};
struct FusedFwdHost {
  ShardView table; const float* pos;
  const int64_t* in_ids; const int32_t* tok_off; const int32_t* row_seq; const int32_t* tiles;
  float* x0;
  int B, L, n_layer;
  float ln_eps;
  Dropout d_embed;
  FusedLayerHost layer[8];
};
bool fused_fwd_supported(int L, int D, int F, int n_head);
int fused_tiles_cap(int B, int L);
int launch_fused_tiles(const int32_t* tok_off, int B, int32_t* tiles, cudaStream_t st);
int launch_sasrec_fwd_fused(const FusedFwdHost& h, cudaStream_t st);
int fused_fwd_set_trace(int* host_mapped);
int attn_bwd_set_trace(int* host_mapped);
// backward of the position-wise half of a layer (LN2', dpre, dx1, LN1', d(attn)) as one persistent kernel  [fused_fwd.cu]
struct FusedBwdFfnHost {
  const float *gin, *z2, *st2, *gp, *z1, *st1, *gamma2, *gamma1;   // gp = mask_ffn_h * gelu'(pre) from the fused forward
  const uint16_t* img[6];   // W2^T hi, lo, W1^T hi, lo, Wo^T hi, lo (backward-data weight images)
  float *g3, *dpre, *dx1, *g1, *g2;
  float *part_ln2, *part_ln1;   // [kLnBwdBlocks][3 * 128]: {sum dy xhat, sum dy, sum dz mask} column partials of LN2 / LN1
  const int32_t* counts;
  int T_cap;
  Dropout d_ffn_out, d_attn_out;
};
int launch_sasrec_bwd_ffn_fused(const FusedBwdFfnHost& h, cudaStream_t st);

// LayerNorm backward over packed rows + column partials  [rowops.cu]
//   dz = LN'(dy; z, stats, gamma);  partials[blk][0..D) = sum dy*xhat, [D..2D) = sum dy,
//   [2D..3D) = sum dz * bias_drop.factor (gradient of the bias that sits under the dropout before this LN)
int launch_ln_bwd(const float* dy, const float* z, const float* stats, const float* gamma, float* dz, float* partials,
                  int D, int T_cap, const int32_t* tok_dev, Dropout bias_drop, cudaStream_t st,
                  Dropout dy_drop = Dropout{0u, 0u, 1.0f});
// y = dropout(LayerNorm(z)) over packed rows, stats = {mean, rstd}
int launch_ln_fwd(const float* z, const float* gamma, const float* beta, float eps, float* y, float* stats, int D, int T_cap,
                  const int32_t* tok_dev, Dropout out_drop, cudaStream_t st);
// column sums of x[T, N] -> partials[blk][N]
constexpr int kColsumBlocks = 2 * kNumSMs;
int launch_colsum(const float* x, int N, int T_cap, const int32_t* tok_dev, float* partials, cudaStream_t st);

// out[e] = sum_s src[s * stride + e] for up to kMaxSeg segments in one launch
constexpr int kMaxSeg = 16;
struct ReduceSeg { const float* src; float* dst; int n_split; int64_t stride; int n; };
struct ReduceTable { ReduceSeg seg[kMaxSeg]; int count; };
int launch_reduce_segments(const ReduceTable& tab, cudaStream_t st);

}  // namespace dr4sr
