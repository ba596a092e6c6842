/* dr4sr.h -- C ABI of libdr4sr (B200 / sm_100a kernels for DR4SR's sequential-recommender hot path).
 *
 * The reference (USTC-StarTeam/DR4SR) has no FFI: its operator surface is Python duck typing
 * (`run.py -m <Model>` -> utils/utils.py:32-36 -> `model.<name>.<Name>`), and every kernel it runs is
 * an ATen/cuBLAS/cuDNN/cuFFT call dispatched from the functions cited below.  Each entry point here
 * replaces the set of library calls made by ONE of those reference functions; the Python classes in
 * dr4sr_b200/model/*.py (same names, constructor and hook signatures as the reference) bind them
 * through ctypes.  INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - plain C: raw device pointers, sizes, POD structs; no torch types; caller owns every buffer.
 *   - every function enqueues work on `stream` and returns immediately (no host sync, no allocation,
 *     no global mutable state); return 0 on success, a negative DR4SR_E* code otherwise.
 *   - ids are int64 (the reference batch contract, data/dataset.py:149-164); floats are fp32.
 *   - "packed" activations: only slots t < seqlen[b] of a post-padded batch are materialised; row
 *     tok_off[b] + t of a [T_cap, D] buffer holds (b, t), T_cap = B*L.  Pad slots are exact dead work
 *     in SASRec/GRU4Rec, so nothing is lost.
 */
#ifndef DR4SR_H_
#define DR4SR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* dr4sr_stream_t; /* == cudaStream_t */

#if defined(__GNUC__)
#define DR4SR_API __attribute__((visibility("default")))
#else
#define DR4SR_API
#endif

enum {
  DR4SR_OK = 0,
  DR4SR_EINVAL = -1,    /* unsupported shape / null pointer */
  DR4SR_EWORKSPACE = -2, /* workspace too small */
  DR4SR_ECUDA = -3       /* a CUDA launch failed; see dr4sr_last_cuda_error() */
};

#define DR4SR_ABI_VERSION 1
DR4SR_API int dr4sr_abi_version(void);
/* last CUDA error string seen by this thread's most recent failing call ("" if none) */
DR4SR_API const char* dr4sr_last_cuda_error(void);

/* Multi-GPU, peer memory: lets kernels launched on the CURRENT device dereference memory of `peer_device` (item-table
 * shards of the other ranks, mapped through CUDA IPC) over NVLink.  Idempotent; DR4SR_EINVAL when the pair has no P2P path. */
DR4SR_API int dr4sr_enable_peer_access(int peer_device);

/* Measurement hooks (bench / profiling only; not on the product path).
 * dr4sr_launch_count: kernels launched by this library since load.
 * dr4sr_prof_enable(1): bracket every launcher with CUDA events on its stream; dr4sr_prof_collect waits
 * for them and writes "name,launches,total_ms\n" lines sorted by total time into buf (returns bytes). */
DR4SR_API long long dr4sr_launch_count(void);
DR4SR_API int dr4sr_prof_enable(int on);
DR4SR_API size_t dr4sr_prof_collect(char* buf, size_t cap);
/* Debug only: install (null clears) a host-mapped buffer of 4 ints per CTA into which the persistent fused kernels
 * record the last phase each CTA reached and, if a bounded barrier wait expires (the kernel then traps instead of
 * hanging the device), which barrier it was: {phase, barrier tag, thread, parity}. */
DR4SR_API int dr4sr_debug_trace(int* host_mapped);

/* ------------------------------------------------------------------------------------------------
 * Batch preparation: packed-token index from `seqlen`, loss normaliser from `item_id`.
 * Replaces the mask construction at model/sasrec.py:48,58 and `(~padding_mask).sum()` at
 * model/loss_func.py:18,28.
 *   seqlen [B] i64, item_id [B,L] (or [B] when target_is_1d) i64
 *   tok_off [B+1] i32 (exclusive prefix sums of clamp(seqlen,0,L)); row_seq [B*L] i32 (row -> b)
 *   counts [4] i32: {T_valid, n_valid_targets, 0, 0}
 */
DR4SR_API int dr4sr_prep_batch(const int64_t* seqlen, const int64_t* item_id, int32_t B, int32_t L, int32_t target_is_1d,
                     int32_t* tok_off, int32_t* row_seq, int32_t* counts, dr4sr_stream_t stream);

/* Batch assembly, replaces SeparateDataset.__getitem__ + default_collate (data/dataset.py:149-164, one Python
 * call per sample): out[r, :] = src[idx[r], :] for a device-resident int64 column of `width` elements per row. */
DR4SR_API int dr4sr_gather_i64(const int64_t* src, int32_t width, const int64_t* idx, int64_t m, int64_t* out, dr4sr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Negative sampling, replaces BaseModel._neg_sampling (model/basemodel.py:50-61): one id per target
 * slot, i.i.d. uniform on {1..N-1}, with replacement (no B x N weight matrix).  Counter-based RNG:
 * out[i] = 1 + mulhi(hash(seed, step, i), N-1).
 */
DR4SR_API int dr4sr_neg_sample(int64_t* out, int64_t n, int64_t num_items, uint64_t seed, uint64_t step, dr4sr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * SASRec encoder, replaces SASRecQueryEncoder.forward (model/sasrec.py:39-75) with the
 * torch.nn.TransformerEncoder configured at model/sasrec.py:21-34 (post-norm, GELU-erf, causal +
 * key-padding mask) and SeqPoolingLayer 'origin' / 'last' (module/layers.py:41-50,69-73).
 *
 * Flat parameter layout (fp32, contiguous, same order as the reference state_dict):
 *   position_emb [L,D];  then per layer:
 *   in_proj_weight [3D,D], in_proj_bias [3D], out_proj.weight [D,D], out_proj.bias [D],
 *   linear1.weight [F,D], linear1.bias [F], linear2.weight [D,F], linear2.bias [D],
 *   norm1.weight [D], norm1.bias [D], norm2.weight [D], norm2.bias [D]
 * The gradient buffer has the same layout.
 */
typedef struct {
  int32_t B, L, D, F, n_head, n_layer;
  int64_t N;          /* table rows (incl. pad row 0) */
  float dropout_p;    /* applied only when train != 0 */
  float ln_eps;
  uint64_t seed;      /* dropout stream = (seed, step) */
  uint64_t step;
} dr4sr_sasrec_cfg;

/* Dense-layer backend: 0 (default) = tcgen05/TMEM with bf16 hi/lo split operands (3 UMMAs per product,
 * fp32 accumulate) wherever the shape allows (N % 128 == 0, K % 64 == 0), 1 = exact-fp32 FFMA kernels
 * everywhere.  Process-wide; used by the parity tests to check both. */
DR4SR_API int dr4sr_set_gemm_backend(int backend);
/* Attention backend: 0 = register-tiled FFMA kernels, one CTA per (sequence, head); 1 = tcgen05 tiles of whole sequences
 * (<= 128 rows, fixed windows) with bf16 hi/lo split operands, forward and backward; 2 (default) = the persistent tcgen05
 * backward over the fused forward's greedy tiles (csrc/attention_bwd_tc.cu; D = 128, 2 heads -- other shapes fall back to
 * the FFMA kernels), forward as in 0 when the per-operator schedule runs.  All are parity-tested. */
DR4SR_API int dr4sr_set_attn_backend(int backend);
/* Encoder schedule: 2 (default) = persistent fused forward (one CTA carries a group of whole sequences, <= 128 packed rows,
 * through every layer: gather, QKV, attention, out-proj+LN, FFN+LN on tcgen05) plus, in the backward, the position-wise
 * half of each layer (LN2', dpre, dx1, LN1', d(attn) and the LayerNorm / bias column partials) as one persistent kernel per
 * layer, where the shape allows (D = F = 128, 2 heads); 1 = fused forward, one kernel per operator in the backward;
 * 0 = one kernel per operator everywhere.  All three are parity-tested. */
DR4SR_API int dr4sr_set_fused_backend(int backend);
/* The tiling the fused kernels run on: greedy groups of whole sequences with <= 128 packed rows (no sequence is split,
 * so attention never leaves a tile).  tiles[0] = n_tiles, tiles[1 + k] = first sequence of tile k, tiles[1 + n_tiles] = B;
 * tiles_len (ints) must be >= B*L/64 + 6.  Exposed for tests; dr4sr_sasrec_fwd builds it in its workspace. */
DR4SR_API int dr4sr_fused_tiles(const int32_t* tok_off, int32_t B, int32_t L, int32_t* tiles, size_t tiles_len, dr4sr_stream_t stream);

DR4SR_API size_t dr4sr_sasrec_param_count(const dr4sr_sasrec_cfg* cfg);
DR4SR_API size_t dr4sr_sasrec_workspace_bytes(const dr4sr_sasrec_cfg* cfg);

/* Forward.  table [N,D]; params flat; in_item_id [B,L]; tok_off/row_seq/counts from dr4sr_prep_batch;
 * ws: workspace (activations stay there for the backward); train: 1 = dropout on.
 * q_packed (optional out) [B*L, D]: encoder output rows in packed order (== 'origin' pooling on
 * valid rows); q_last (optional out) [B, D]: row seqlen-1 of each sequence ('last' pooling);
 * q_dense (optional out) [B, L, D]: 'origin' pooling in the reference's dense layout. */
DR4SR_API int dr4sr_sasrec_fwd(const dr4sr_sasrec_cfg* cfg, const float* table, const float* params,
                     const int64_t* in_item_id, const int32_t* tok_off, const int32_t* row_seq,
                     const int32_t* counts, void* ws, size_t ws_bytes, int32_t train,
                     float* q_packed, float* q_last, float* q_dense, dr4sr_stream_t stream);

/* Backward of the forward whose activations are in `ws`.  dq_packed [B*L, D] (in: dLoss/dq rows,
 * clobbered).  Writes grads (flat layout, overwritten not accumulated) and dx0_packed [B*L, D]
 * (gradient w.r.t. the gathered input rows; consumed by dr4sr_table_grad). */
DR4SR_API int dr4sr_sasrec_bwd(const dr4sr_sasrec_cfg* cfg, const float* table, const float* params,
                     const int64_t* in_item_id, const int32_t* tok_off, const int32_t* row_seq,
                     const int32_t* counts, void* ws, size_t ws_bytes, float* dq_packed,
                     float* grads, float* dx0_packed, dr4sr_stream_t stream);
/* The weight gradients of the backward run on an internal side stream that overlaps the data-gradient chain.
 * dr4sr_sasrec_bwd joins it before returning (every gradient is ordered on `stream`).  dr4sr_sasrec_bwd_async leaves
 * the join to the caller: dx0_packed is ordered on `stream`, `grads` only after dr4sr_sasrec_bwd_join(stream) -- the
 * Python model enqueues the embedding scatter-add (dr4sr_table_grad) in between, under the tail of the side stream. */
DR4SR_API int dr4sr_sasrec_bwd_async(const dr4sr_sasrec_cfg* cfg, const float* table, const float* params,
                     const int64_t* in_item_id, const int32_t* tok_off, const int32_t* row_seq,
                     const int32_t* counts, void* ws, size_t ws_bytes, float* dq_packed,
                     float* grads, float* dx0_packed, dr4sr_stream_t stream);
DR4SR_API int dr4sr_sasrec_bwd_join(dr4sr_stream_t stream);

/* Packed rows -> the reference's layouts: q_last [B,D] = row seqlen-1 of every sequence ('last' pooling,
 * module/layers.py:69-73), q_dense [B,L,D] = rows scattered back with zeros at t >= seqlen ('origin'
 * pooling, module/layers.py:41-50).  Either output may be NULL. */
DR4SR_API int dr4sr_unpack_rows(const float* x_packed, const int32_t* tok_off, int32_t B, int32_t L, int32_t D, float* q_last,
                      float* q_dense, dr4sr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * GRU4Rec encoder, replaces GRU4Rec.query_encoder (model/gru4rec.py:12-22, module/layers.py:117-136):
 * x0 = drop(E[ids]); n_layer bias-free GRU layers (D -> H -> H, h0 = 0, gate rows [r; z; n]);
 * y = W_out h + b_out per slot; pooling as for SASRec.  Packed rows (pads are trailing).
 * Flat parameter layout (state_dict order): per layer gru.weight_ih_l{k} [3H, in], gru.weight_hh_l{k}
 * [3H, H]; then linear.weight [D, H], linear.bias [D].  H in {64, 128, 256}, D in {64, 128}.
 * Same calling convention as dr4sr_sasrec_fwd / dr4sr_sasrec_bwd.
 */
typedef struct {
  int32_t B, L, D, H, n_layer;
  int64_t N;
  float dropout_p;
  uint64_t seed;
  uint64_t step;
} dr4sr_gru_cfg;
DR4SR_API size_t dr4sr_gru_param_count(const dr4sr_gru_cfg* cfg);
DR4SR_API size_t dr4sr_gru_workspace_bytes(const dr4sr_gru_cfg* cfg);
DR4SR_API int dr4sr_gru_fwd(const dr4sr_gru_cfg* cfg, const float* table, const float* params, const int64_t* in_item_id,
                  const int32_t* tok_off, const int32_t* row_seq, const int32_t* counts, void* ws, size_t ws_bytes,
                  int32_t train, float* q_packed, float* q_last, float* q_dense, dr4sr_stream_t stream);
DR4SR_API int dr4sr_gru_bwd(const dr4sr_gru_cfg* cfg, const float* table, const float* params, const int64_t* in_item_id,
                  const int32_t* tok_off, const int32_t* row_seq, const int32_t* counts, void* ws, size_t ws_bytes,
                  float* dq_packed, float* grads, float* dx0_packed, dr4sr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * FMLP encoder, replaces FMLP.add_position_embedding + FMLPEncoder (model/fmlp.py:18-39,
 * module/layers.py:740-808): x0 = drop(LN(E[ids] + P)); per layer y = LN(drop(filter(x)) + x),
 * z = LN(drop(W2 gelu(W1 y + b1) + b2) + y) with inner width 4D; query = last position.  Inputs are
 * pre-padded, every slot is live (dense [B, L, D]); L must be 50 (the reference hard-codes it).
 * Flat parameter layout (state_dict order): position_embeddings [L,D], LayerNorm.weight [D], .bias [D];
 * per layer: filterlayer.complex_weight [L/2+1, D, 2], filterlayer.LayerNorm.weight/.bias [D],
 * intermediate.dense_1.weight [4D,D], .bias [4D], dense_2.weight [D,4D], .bias [D],
 * intermediate.LayerNorm.weight/.bias [D].
 * fwd: q_last [B,D] out.  bwd: dq_last [B,D] in; grads (flat, overwritten; the position block is left
 * to dr4sr_table_grad's pos_grad); dz0_dense [B*L, D] out = gradient w.r.t. E[ids] + P per slot.
 */
typedef struct {
  int32_t B, L, D, n_layer;
  int64_t N;
  float dropout_p;
  float ln_eps;
  uint64_t seed;
  uint64_t step;
} dr4sr_fmlp_cfg;
DR4SR_API size_t dr4sr_fmlp_param_count(const dr4sr_fmlp_cfg* cfg);
DR4SR_API size_t dr4sr_fmlp_workspace_bytes(const dr4sr_fmlp_cfg* cfg);
DR4SR_API int dr4sr_fmlp_fwd(const dr4sr_fmlp_cfg* cfg, const float* table, const float* params, const int64_t* in_item_id,
                   void* ws, size_t ws_bytes, int32_t train, float* q_last, dr4sr_stream_t stream);
DR4SR_API int dr4sr_fmlp_bwd(const dr4sr_fmlp_cfg* cfg, const float* table, const float* params, const int64_t* in_item_id,
                   void* ws, size_t ws_bytes, const float* dq_last, float* grads, float* dz0_dense,
                   dr4sr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Sampled scoring + BCE, replaces BaseModel.training_step (model/basemodel.py:205-210) and
 * BinaryCrossEntropyLoss.forward (model/loss_func.py:9-35) with one negative per slot, forward and
 * backward in one pass:
 *   s+ = <q, E[item_id]>, s- = <q, E[neg]>, loss_pos[b,t] = (-logsig(s+) + softplus(s-)) / n
 *   dq = ds+ E+ + ds- E-;  ds+ = -sigmoid(-s+) w/n, ds- = sigmoid(s-) w/n      (w = loss_weight or 1)
 * q_packed [B*L,D] packed rows; item_id/neg_item [B,L] i64; loss_pos [B,L] (out, 0 at pads);
 * dscore [B*L,2] packed (out: ds+, ds-); dq_packed (out).  upstream: optional device scalar that
 * multiplies every gradient (autograd's grad_output), NULL = 1.  loss_weight: optional [B,L].
 */
DR4SR_API int dr4sr_score_bce(const float* q_packed, const float* table, const int64_t* item_id, const int64_t* neg_item,
                    const int32_t* tok_off, const int32_t* row_seq, const int32_t* counts, int32_t B, int32_t L,
                    int32_t D, const float* loss_weight, const float* upstream, float* loss_pos, float* dscore,
                    float* dq_packed, dr4sr_stream_t stream);
/* The same sweep with the loss selected by `kind`: DR4SR_LOSS_BCE = dr4sr_score_bce; DR4SR_LOSS_BPR = the BPR loss as
 * model/loss_func.py:40-49 intends it with one negative, loss_pos[b,t] = -logsig(s+ - s-) / n, ds+ = -sigmoid(s- - s+) w/n,
 * ds- = -ds+ (the reference's training_step cannot reach it: it passes `reduce` to BPRLoss.forward, a TypeError at
 * model/basemodel.py:210 -- this is the "fixed BPR" extension, selected by config['model']['loss_fn'] = 'bpr'). */
enum { DR4SR_LOSS_BCE = 0, DR4SR_LOSS_BPR = 1 };
DR4SR_API int dr4sr_score_loss(int32_t kind, const float* q_packed, const float* table, const int64_t* item_id,
                     const int64_t* neg_item, const int32_t* tok_off, const int32_t* row_seq, const int32_t* counts, int32_t B,
                     int32_t L, int32_t D, const float* loss_weight, const float* upstream, float* loss_pos, float* dscore,
                     float* dq_packed, dr4sr_stream_t stream);
/* dq_packed and dscore of a dr4sr_score_bce call made with upstream = NULL, multiplied afterwards by the device scalar
 * `upstream` (autograd's grad_output arrives only at backward time): lets the forward pass compute the gradients in
 * the same sweep as the loss; the kernel exits at once when *upstream == 1. */
DR4SR_API int dr4sr_scale_grads(const float* upstream, const int32_t* counts, int32_t D, float* dq_packed, float* dscore,
                      dr4sr_stream_t stream);

/* Deterministic sum of loss_pos [n] into loss[0] (reduce=True path). */
DR4SR_API int dr4sr_sum(const float* x, int64_t n, float* out, dr4sr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Embedding-gradient scatter-add, replaces the autograd backward of the three table gathers
 * (nn.Embedding input lookup model/sasrec.py:46, weight[item_id] / weight[neg_item]
 * model/basemodel.py:206-207): table_grad[in_id] += dx0; [item_id] += ds+ q; [neg] += ds- q.
 * table_grad [N,D] is accumulated into (caller zeroes it; dr4sr_adam can zero it while reading).
 * pos_grad [L,D] (optional): dP[t] = sum_b dx0[b,t], overwritten; needs ws of
 * dr4sr_table_grad_workspace_bytes(L, D) bytes (per-CTA partials, summed in a fixed order).
 */
DR4SR_API size_t dr4sr_table_grad_workspace_bytes(int32_t L, int32_t D);
DR4SR_API int dr4sr_table_grad(const float* dx0_packed, const float* q_packed, const float* dscore,
                     const int64_t* in_item_id, const int64_t* item_id, const int64_t* neg_item,
                     const int32_t* tok_off, const int32_t* row_seq, const int32_t* counts, int32_t B, int32_t L,
                     int32_t D, int64_t N, float* table_grad, float* pos_grad, void* ws, size_t ws_bytes,
                     dr4sr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Row-sharded item table over PEER MEMORY (multi-GPU, one process per GPU; no reference counterpart -- the reference is
 * single-device).  Rank r owns rows [lo[r], lo[r+1]) of E and of its gradient accumulator G; every rank maps the other
 * ranks' shards into its address space (CUDA IPC opened in the local device's context + dr4sr_enable_peer_access) and
 * passes the resulting pointers here.  The kernels then read E rows and `red.global.add` gradient rows directly in the
 * owner's HBM over NVLink / NVSwitch: the row exchange is part of the gather / scatter kernels themselves, there is no
 * collective, no request planning and no host synchronisation (measured on 2 x B200: 550 GB/s of 512-byte row gathers
 * and 530 GB/s of 512-byte vector atomics per direction, profiles/r2_p2p_probe.md).  The caller orders the steps
 * across ranks (gathers after every rank's Adam, Adam after every rank's scatter) with stream-ordered barriers.
 *   world <= 8;  table[r] / grad[r]: base of rank r's shard (row lo[r] at offset 0);  lo[world] = N.
 * A null map (or world == 1) everywhere below means "the whole table is local". */
#define DR4SR_MAX_SHARDS 8
typedef struct {
  const float* table[DR4SR_MAX_SHARDS];
  float* grad[DR4SR_MAX_SHARDS];
  int64_t lo[DR4SR_MAX_SHARDS + 1];
  int32_t world;
  int32_t rank;
} dr4sr_shard_map;
/* Ranking metrics of one eval batch, accumulated on the device: for user u with the relevant item target[u] at position r of
 * topk_ids[u, 0..k) (absent: no contribution), sums[2 i] += 1 / log2(r + 2) and sums[2 i + 1] += 1 for every cutoff
 * cutoffs[i] > r (ndcg@c and recall@c numerators; divide by the number of users at the end of the epoch).  `cutoffs` is a HOST
 * array of n_cutoffs <= 8 values in 1..k; `sums` a device array of 2 * n_cutoffs doubles the caller zeroes once per epoch.
 * Reference: evaluation/__init__.py:9-36,107-134 via model/basemodel.py:337-352. */
DR4SR_API int dr4sr_rank_metrics(const int64_t* topk_ids, const int64_t* target, int32_t B, int32_t k, const int32_t* cutoffs,
                                 int32_t n_cutoffs, double* sums, dr4sr_stream_t stream);

/* Deterministic variant of dr4sr_table_grad: the (item id, source row) entries of the step are sorted by id (stable radix
 * sort) and every run of equal ids is reduced by warps in a fixed order -- a warp-segmented reduction, no float atomics, every
 * row of table_grad has one writer.  Same arguments and result contract; two calls on the same inputs give bit-identical
 * gradients, and a hot id (Zipf catalogs) costs run/16 row adds instead of `run` atomics serialised on one L2 line.
 * D <= 256.  `ws` needs dr4sr_table_grad_sorted_workspace_bytes(B, L, D, N) bytes.  Reference: the index_put / embedding
 * backward torch runs under model/basemodel.py:198 (which sorts, too). */
DR4SR_API size_t dr4sr_table_grad_sorted_workspace_bytes(int32_t B, int32_t L, int32_t D, int64_t N);
DR4SR_API int dr4sr_table_grad_sorted(const float* dx0_packed, const float* q_packed, const float* dscore,
                                      const int64_t* in_item_id, const int64_t* item_id, const int64_t* neg_item,
                                      const int32_t* tok_off, const int32_t* row_seq, const int32_t* counts, int32_t B, int32_t L,
                                      int32_t D, int64_t N, float* table_grad, float* pos_grad, void* ws, size_t ws_bytes,
                                      dr4sr_stream_t stream);

/* The target / negative part of dr4sr_table_grad (dE[item_id] += ds+ q, dE[neg] += ds- q), queued on the library's
 * background stream behind everything enqueued on `stream` so far: it depends on the loss kernel only and runs under the
 * encoder backward.  Follow with dr4sr_table_grad(dx0, NULL, NULL, in_item_id, NULL, NULL, ...) for the input rows and
 * dr4sr_table_grad_targets_join(stream), after which `stream` is ordered behind the background kernel.  The sum of the
 * two calls is dr4sr_table_grad's result (float atomics: equal to summation order).  Reference: the autograd scatter of
 * nn.Embedding's backward, model/basemodel.py:204-214 through torch. */
DR4SR_API int dr4sr_table_grad_targets_async(const float* q_packed, const float* dscore, const int64_t* item_id,
                                             const int64_t* neg_item, const int32_t* tok_off, const int32_t* row_seq,
                                             const int32_t* counts, int32_t B, int32_t L, int32_t D, int64_t N, float* table_grad,
                                             dr4sr_stream_t stream);
DR4SR_API int dr4sr_table_grad_targets_join(dr4sr_stream_t stream);

/* The per-step synchronisation of the peer-sharded layout as kernels over peer memory (no reference counterpart).
 * flags[r] / slots[r] / stage[r]: rank r's flag array (DR4SR_MAX_SHARDS int32, zero-initialised), count slots
 * (DR4SR_MAX_SHARDS int32) and staging buffer (floats), as pointers valid on THIS device (peer mappings for r != rank).
 * dr4sr_peer_barrier: stream-ordered barrier over the ranks; kernels enqueued after it on any rank start once the kernels
 *   enqueued before it on every rank have finished.  `epoch` must be positive and increase with every barrier / all-reduce
 *   call, identically on every rank.  count_inout (nullable): a device int32 that is replaced by its sum over the ranks.
 * dr4sr_peer_allreduce: the same barrier, then out[i] = sum_r stage[r][i] for i < n, added in rank order (bit-identical on
 *   every rank).  The staging buffers may be overwritten again after the NEXT barrier.
 * A rank that never arrives makes the others trap after 60 s instead of hanging. */
typedef struct dr4sr_peer_comm {
  int32_t* flags[DR4SR_MAX_SHARDS];
  int32_t* slots[DR4SR_MAX_SHARDS];
  float* stage[DR4SR_MAX_SHARDS];
  int32_t world, rank;
} dr4sr_peer_comm;
DR4SR_API int dr4sr_peer_barrier(const dr4sr_peer_comm* comm, int32_t epoch, int32_t* count_inout, dr4sr_stream_t stream);
DR4SR_API int dr4sr_peer_allreduce(const dr4sr_peer_comm* comm, int32_t epoch, int64_t n, float* out, dr4sr_stream_t stream);

/* dr4sr_sasrec_fwd / dr4sr_score_loss / dr4sr_table_grad with the table (resp. its gradient) given as a shard map. */
DR4SR_API int dr4sr_sasrec_fwd_sharded(const dr4sr_sasrec_cfg* cfg, const dr4sr_shard_map* map, const float* params,
                     const int64_t* in_item_id, const int32_t* tok_off, const int32_t* row_seq,
                     const int32_t* counts, void* ws, size_t ws_bytes, int32_t train,
                     float* q_packed, float* q_last, float* q_dense, dr4sr_stream_t stream);
DR4SR_API int dr4sr_score_loss_sharded(int32_t kind, const float* q_packed, const dr4sr_shard_map* map, const int64_t* item_id,
                     const int64_t* neg_item, const int32_t* tok_off, const int32_t* row_seq, const int32_t* counts, int32_t B,
                     int32_t L, int32_t D, const float* loss_weight, const float* upstream, float* loss_pos, float* dscore,
                     float* dq_packed, dr4sr_stream_t stream);
DR4SR_API int dr4sr_table_grad_sharded(const float* dx0_packed, const float* q_packed, const float* dscore,
                     const int64_t* in_item_id, const int64_t* item_id, const int64_t* neg_item,
                     const int32_t* tok_off, const int32_t* row_seq, const int32_t* counts, int32_t B, int32_t L,
                     int32_t D, const dr4sr_shard_map* map, float* pos_grad, void* ws, size_t ws_bytes,
                     dr4sr_stream_t stream);
DR4SR_API int dr4sr_table_grad_targets_async_sharded(const float* q_packed, const float* dscore, const int64_t* item_id,
                                                     const int64_t* neg_item, const int32_t* tok_off, const int32_t* row_seq,
                                                     const int32_t* counts, int32_t B, int32_t L, int32_t D,
                                                     const dr4sr_shard_map* map, dr4sr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Row-sharded item table, all-to-all variant (the rows travel through NCCL; kept for GRU4Rec and as the fallback when the
 * GPUs have no peer path).
 * Rank r owns rows [lo_r, hi_r) of E (balanced contiguous ranges: the first num_rows % world ranks hold one
 * extra row).  dr4sr_shard_plan turns the live slots of a batch into row requests bucketed by owner:
 *   request 3*row+k of packed row `row`: k=0 input id, k=1 positive target, k=2 negative (targets only
 *   where item_id != 0; id 0 is never requested)
 *   send_counts [world] i32 (out): requests per owner;  send_ids [3*B*L] i64 (out): ids grouped by owner;
 *   scratch [world+1] i32;  in_loc / item_loc / neg_loc [B,L] i64 (out): the ids remapped to rows of the
 *   staged local table (0 = pad row, 1 + position in send_ids otherwise) -- every other entry point then
 *   runs unchanged on (local table, remapped ids).
 * dr4sr_gather_rows: out[r] = src[ids[r] - lo];  dr4sr_scatter_add_rows: dst[ids[r] - lo] += rows[r].
 */
DR4SR_API int dr4sr_shard_plan(const int64_t* in_item_id, const int64_t* item_id, const int64_t* neg_item, const int32_t* tok_off,
                     const int32_t* row_seq, const int32_t* counts, int32_t B, int32_t L, int64_t num_rows, int32_t world,
                     int32_t* send_counts, int32_t* scratch, int64_t* send_ids, int64_t* in_loc, int64_t* item_loc,
                     int64_t* neg_loc, dr4sr_stream_t stream);
DR4SR_API int dr4sr_gather_rows(const float* src, const int64_t* ids, int64_t lo, int64_t m, int32_t D, float* out,
                      dr4sr_stream_t stream);
DR4SR_API int dr4sr_scatter_add_rows(float* dst, const int64_t* ids, int64_t lo, int64_t m, int32_t D, const float* rows,
                           dr4sr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Dense Adam, replaces torch.optim.Adam.step as configured at model/basemodel.py:85-86
 * (betas .9/.999, eps 1e-8, L2 weight decay added to the gradient, no amsgrad).  One pass over
 * (p, g, m, v); step is 1-based; zero_grad != 0 clears g after reading it.
 */
DR4SR_API int dr4sr_adam(float* p, float* g, float* m, float* v, int64_t n, int64_t step, float lr, float beta1, float beta2,
               float eps, float weight_decay, int32_t zero_grad, dr4sr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Full-catalog scoring + top-k, replaces BaseModel.topk (model/basemodel.py:354-365):
 * scores = q @ table[:N].T; -inf where item_dead[n] != 0 (ids outside the eval domain, always id 0)
 * and at every id in user_hist[b,:]; top-k values (descending) and ids.
 * q [B,D]; item_dead [N] u8; user_hist [B,H] i64 (may be NULL, H=0); out_scores [B,k] f32, out_ids [B,k] i64.
 * ws: at least dr4sr_topk_workspace_bytes(B,N,k).
 */
DR4SR_API size_t dr4sr_topk_workspace_bytes(int32_t B, int64_t N, int32_t k);
DR4SR_API int dr4sr_topk(const float* q, const float* table, const uint8_t* item_dead, const int64_t* user_hist, int32_t B,
               int32_t D, int64_t N, int32_t H, int32_t k, float* out_scores, int64_t* out_ids, void* ws,
               size_t ws_bytes, dr4sr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Building blocks exported for unit parity tests (each is also used inside the composites).
 */
/* x0_packed[row] = dropout(table[in_id] + pos[t]) (model/sasrec.py:42-46,64-66) */
DR4SR_API int dr4sr_embed_fwd(const float* table, const float* pos, const int64_t* in_item_id, const int32_t* tok_off,
                    const int32_t* row_seq, const int32_t* counts, int32_t B, int32_t L, int32_t D, float dropout_p,
                    uint64_t seed, uint64_t step, float* x0_packed, dr4sr_stream_t stream);
/* y[M,N] = x[M,K] @ w[N,K]^T + bias (torch.nn.functional.linear); M read from m_dev[0] if non-NULL */
DR4SR_API int dr4sr_linear_fwd(const float* x, const float* w, const float* bias, float* y, int32_t M, int32_t N, int32_t K,
                     const int32_t* m_dev, dr4sr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DR4SR_H_ */
