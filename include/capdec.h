/* capdec.h -- C ABI of the B200-native caption-decoding hot path (libcapdec.so).
 *
 * Drop-in boundary for thromel/Image-Captioning-ML-Project's batched decode step.  The reference
 * has no FFI of its own (it is pure PyTorch), so every entry point below cites the Python symbol
 * it replaces; the Python mirror of the reference classes binds these with ctypes
 * (image-captioning-ml-project_b200/_capi.py).
 *
 * Conventions
 *   - plain C: pointers + sizes only, no torch / C++ types.
 *   - every function returns 0 on success or a negative capdec_status; the message is available
 *     from capdec_last_error() (thread-local).  Nothing throws across the boundary.
 *   - all `*_dev` pointers are device pointers on the CURRENT cuda device, fp32 row-major unless
 *     stated; all work is enqueued on the caller's stream (`stream` is a cudaStream_t passed as
 *     void*), no hidden synchronisation except in the *_host entry points.
 *   - the caller owns inputs, outputs and the workspace; the handle owns only packed weight
 *     copies and GEMM scratch.  A handle is not thread-safe and carries ONE call in flight at a time:
 *     enqueue the next call on the same stream, or synchronise first (the *_host entry points drain
 *     the device before they start, so they may follow an asynchronous call directly).  Handles bind
 *     to the device that is current at capdec_create; one process may hold handles on several devices.
 *   - there is NO CPU fallback: without a CUDA device the calls fail with CAPDEC_ERR_CUDA.
 */
#ifndef CAPDEC_H
#define CAPDEC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAPDEC_VERSION 200

typedef enum {
  CAPDEC_OK = 0,
  CAPDEC_ERR_INVALID = -1,      /* bad argument / shape */
  CAPDEC_ERR_UNSUPPORTED = -2,  /* valid request this build has no kernel for (never falls back) */
  CAPDEC_ERR_CUDA = -3,         /* CUDA runtime error, text in capdec_last_error() */
  CAPDEC_ERR_STATE = -4,        /* call order (weights missing, not finalized, ...) */
  CAPDEC_ERR_WORKSPACE = -5     /* workspace too small */
} capdec_status;

/* decoder family: which reference class the handle mirrors */
typedef enum {
  CAPDEC_ARCH_LEGACY_SAT = 0,  /* models/decoder.py::Decoder (196x2048 ResNet features, LSTMCell) */
  CAPDEC_ARCH_LSTM = 1,        /* src/models/decoders.py::LSTMDecoder + src/models/attention.py */
  CAPDEC_ARCH_TRANSFORMER = 2, /* src/models/decoders.py::TransformerDecoder (6x post-LN nn.TransformerDecoderLayer), KV-cached */
  CAPDEC_ARCH_GPT2 = 3         /* src/models/decoders.py::GPT2Decoder: HF GPT-2 blocks (pre-LN, gelu_new, tied lm_head) over a
                                  10-token image prefix used as past K == V of every layer; feature_dim = pooled feature width;
                                  Conv1D weights are bound transposed ([out,in]) */
} capdec_arch;

/* src/config.py::AttentionType (LSTM arch only; legacy always uses its additive-ReLU attention) */
typedef enum {
  CAPDEC_ATT_SOFT = 0,
  CAPDEC_ATT_MULTI_HEAD = 1,
  CAPDEC_ATT_ADAPTIVE = 2,
  CAPDEC_ATT_AOA = 3
} capdec_attention;

/* arithmetic of the dense contractions.  Attention / softmax / cell math is always fp32; what changes with the mode
   is the storage the legacy attention kernel streams its region tiles from: fp32 (FP32, TF32X3, TF32), bf16 (BF16), or
   24-bit planes with 16 significant bits (BF16X3), and the soft-attention tanh (exact tanhf in FP32, an exp / reciprocal
   form accurate to ~1e-6 in the tensor-core modes). */
typedef enum {
  CAPDEC_PREC_FP32 = 0,    /* CUDA-core FFMA, IEEE fp32 accumulate: the exact mode            */
  CAPDEC_PREC_TF32X3 = 1,  /* tcgen05 kind::tf32, 3-term split (hi*hi + hi*lo + lo*hi), fp32 accumulate */
  CAPDEC_PREC_BF16 = 2,    /* tcgen05 kind::f16 on round-to-nearest bf16 operands, fp32 accumulate (north star "bf16 mode") */
  CAPDEC_PREC_TF32 = 3,    /* tcgen05 kind::tf32 single pass on round-to-nearest TF32 operands, fp32 accumulate */
  CAPDEC_PREC_BF16X3 = 4   /* tcgen05 kind::f16, 3-term bf16 split (hi*hi + hi*lo + lo*hi): ~16 mantissa bits per operand,
                              fp32 accumulate, at half the tensor time of TF32X3 */
} capdec_precision;

typedef struct {
  int32_t arch;           /* capdec_arch */
  int32_t attention;      /* capdec_attention */
  int32_t precision;      /* capdec_precision */
  int32_t vocab_size;     /* V */
  int32_t hidden_dim;     /* H: decoder_dim (legacy 512) / DecoderConfig.hidden_dim */
  int32_t embed_dim;      /* E: legacy 512 / LSTMDecoder.embedding_dim */
  int32_t feature_dim;    /* D: encoder feature dim (legacy 2048; LSTM arch == hidden_dim) */
  int32_t attention_dim;  /* A: legacy 512; LSTM arch == hidden_dim */
  int32_t num_layers;     /* nn.LSTM layers (legacy: 1) */
  int32_t num_heads;      /* AttentionConfig.num_heads */
  float temperature;      /* AttentionConfig.temperature */
  int32_t pad_token_id, bos_token_id, eos_token_id; /* models/constants.py / src/config.py:122-124 */
} capdec_config;

typedef struct capdec_handle capdec_handle;

const char* capdec_last_error(void);
int capdec_version(void);
/* number of CUDA kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t capdec_launch_count(void);

/* ---- lifecycle ---------------------------------------------------------------------------- */
int capdec_create(const capdec_config* cfg, capdec_handle** out);
void capdec_destroy(capdec_handle* h);

/* Bind one parameter by its reference state_dict name (e.g. "enc_att.weight",
 * "lstm.weight_ih_l0", "attention.base_attention.key_proj.bias").  The data (fp32, contiguous,
 * device) is copied into handle-owned storage on `stream`; the caller may free it afterwards.
 * Replaces nn.Module.load_state_dict for the classes in models/decoder.py:33-54 and
 * src/models/decoders.py:92-117, src/models/attention.py:48-52,133-137,232-239,311-320. */
int capdec_set_weight(capdec_handle* h, const char* name, const float* data_dev,
                      const int64_t* shape, int32_t rank, void* stream);
/* Validate that every parameter of the configured architecture is bound and build the packed
 * copies the kernels use (concatenated / gate-interleaved / split-precision weights). */
int capdec_finalize(capdec_handle* h, void* stream);

/* ---- decode entry points -------------------------------------------------------------------- */
size_t capdec_workspace_bytes(const capdec_handle* h, int32_t num_images, int32_t num_regions,
                              int32_t rows_per_image, int32_t max_length);

/* Beam search (HF static-shape semantics == the beam search the reference invokes at
 * src/models/decoders.py:645; applied to the legacy step models/decoder.py:148-173 or the
 * LSTMDecoder step src/models/decoders.py:274-303).
 *   features_dev [B,L,D]; pooled_dev [B,H] (LSTM arch; NULL for legacy);
 *   key_padding_mask_dev uint8 [B,L], 1 = padding (may be NULL);
 *   out_tokens_dev int32 [B,max_length] (filled with HF's fill value after the end),
 *   out_lengths_dev int32 [B], out_scores_dev float [B] (length-penalised score of the best beam);
 *   optional trace (may be NULL): per step s in [0,max_length-1): the 2k sorted candidates
 *   dbg_top_logprob_dev float [steps,B,2k], dbg_top_token_dev / dbg_top_beam_dev int32 [steps,B,2k]. */
int capdec_decode_beam(capdec_handle* h, const float* features_dev, const float* pooled_dev,
                       const uint8_t* key_padding_mask_dev, int32_t num_images, int32_t num_regions,
                       int32_t num_beams, int32_t max_length, float length_penalty,
                       int32_t* out_tokens_dev, int32_t* out_lengths_dev, float* out_scores_dev,
                       float* dbg_top_logprob_dev, int32_t* dbg_top_token_dev, int32_t* dbg_top_beam_dev,
                       void* workspace_dev, size_t workspace_bytes, void* stream);

/* Greedy decode with LSTMDecoder.generate's conventions (src/models/decoders.py:236-314):
 * out_tokens[:,0] = start token, exactly max_length steps, last argmax discarded, no EOS stop.
 *   out_alpha_dev float [B,max_length,L] attention weights (may be NULL). */
int capdec_decode_greedy(capdec_handle* h, const float* features_dev, const float* pooled_dev,
                         const uint8_t* key_padding_mask_dev, int32_t num_images, int32_t num_regions,
                         int32_t max_length, int32_t start_token_id, int32_t* out_tokens_dev,
                         float* out_alpha_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* Ancestral sampling rollout (src/train/trainer.py:383-438), `num_samples` rows per image sharing
 * the image's feature tiles, plus an optional greedy row (SCST baseline, trainer.py:353) as the last
 * row of each image.  uniforms_dev float [B*rows, max_length-1] drives the inverse-CDF draw.
 *   out_tokens_dev int32 [B*rows, max_length] (position 0 = bos), out_logprob_dev float [B*rows, max_length-1]. */
int capdec_decode_sample(capdec_handle* h, const float* features_dev, const float* pooled_dev,
                         const uint8_t* key_padding_mask_dev, int32_t num_images, int32_t num_regions,
                         int32_t num_samples, int32_t with_greedy_row, int32_t max_length,
                         const float* uniforms_dev, int32_t* out_tokens_dev, float* out_logprob_dev,
                         void* workspace_dev, size_t workspace_bytes, void* stream);

/* Teacher-forced forward of the legacy decoder, models/decoder.py:120-176.
 *   captions_dev int32 [B,cap_stride] sorted by decreasing length, dec_len_host[B] = caption_length-1,
 *   predictions_dev float [B,T,V], alphas_dev float [B,T,L] with T = max(dec_len) (rows past their
 *   length are left untouched, the caller zero-fills as the reference does). */
int capdec_forward_teacher(capdec_handle* h, const float* features_dev, int32_t num_images,
                           int32_t num_regions, const int32_t* captions_dev, int32_t cap_stride,
                           const int32_t* dec_len_host, float* predictions_dev, float* alphas_dev,
                           void* workspace_dev, size_t workspace_bytes, void* stream);

/* AttentionMechanism.forward for a 2-D query (src/models/attention.py:57,142,242,322):
 *   query_dev [R,H] with R = B*rows_per_image, features_dev [B,L,H] (key == value),
 *   memory_dev / cell_dev [R,H] (adaptive only) -> context_dev [R,H], weights_dev [R,L]. */
int capdec_attention_forward(capdec_handle* h, const float* query_dev, const float* features_dev,
                             const uint8_t* key_padding_mask_dev, const float* memory_dev,
                             const float* cell_dev, int32_t num_images, int32_t num_regions,
                             int32_t rows_per_image, float* context_dev, float* weights_dev,
                             void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- teacher-forced pass over given tokens (SCST boundary) ------------------------------------------
 * One batched pass of the decode step with FORCED tokens: position t consumes tokens[:, t] and produces the
 * distribution of position t+1.  This is what the reference's `decoder(encoder_features, captions=ids)["logits"]`
 * returns (src/models/decoders.py:137-234 LSTM, :377-438 transformer, :563-596 GPT-2) and what
 * CaptioningTrainer._sample_captions calls once per generated token (src/train/trainer.py:413-420); with
 * logprob_dev it is the single-pass re-scoring of sampled captions for the REINFORCE term (trainer.py:366-378).
 * Inference only (no autograd graph).
 *   tokens_dev int32 [R, tok_stride] with R = num_images * rows_per_image (rows of an image adjacent, sharing its tiles);
 *   logits_dev float [R, num_tokens, V] or NULL;  logprob_dev float [R, num_tokens-1] or NULL:
 *   logprob[r,t] = log_softmax(logits[r,t,:])[tokens[r,t+1]];  alpha_dev float [R, num_tokens, L] or NULL (LSTM / legacy);
 *   mask_pad_keys != 0 (transformer / GPT-2): positions whose token == pad_token_id are masked as attention KEYS, the
 *   reference's tgt_key_padding_mask (decoders.py:405) / attention_mask (decoders.py:581). */
int capdec_forward_tokens(capdec_handle* h, const float* features_dev, const float* pooled_dev,
                          const uint8_t* key_padding_mask_dev, int32_t num_images, int32_t num_regions,
                          int32_t rows_per_image, const int32_t* tokens_dev, int32_t tok_stride, int32_t num_tokens,
                          float* logits_dev, float* logprob_dev, float* alpha_dev, int32_t mask_pad_keys,
                          void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- token -> text boundary (src/train/trainer.py:546-547, src/evaluate/metrics.py:322-323) -----------
 * Per row: everything after the first EOS becomes pad_token_id, out_lengths = tokens kept (the EOS included when
 * keep_eos).  In place allowed (out_tokens_dev == tokens_dev).  Replaces the per-caption Python loop's implicit trimming. */
int capdec_trim_at_eos(const int32_t* tokens_dev, int64_t ld, int32_t rows, int32_t max_length, int32_t eos_token_id,
                       int32_t pad_token_id, int32_t keep_eos, int32_t* out_tokens_dev, int64_t ld_out,
                       int32_t* out_lengths_dev, void* stream);

/* ---- encoder -> decoder feature hand-off -----------------------------------------------------------------
 * What the encoders emit (models/encoder.py:12-16 before its permute; src/models/encoders.py:118-137, :209-230 before
 * the CLS drop) goes in ONE pass into the layout the decode kernels stream ("tiles"): the layout change
 * (permute(0,2,3,1) / [:,1:,:]), the widening from bf16/fp16, the region mean (models/decoder.py:137) and -- legacy
 * decoder in BF16X3 mode -- the p24 planes + the hoisted enc_att GEMM's lo operand are produced together, so the
 * decode prologue no longer re-reads and re-packs fp32 features. */
typedef enum {
  CAPDEC_LAYOUT_BLD = 0,      /* [B, L, D] row-major (what the decoders take) */
  CAPDEC_LAYOUT_BDL = 1,      /* [B, D, L]: NCHW feature map [B, D, h, w] with L = h*w (ResNet trunk output) */
  CAPDEC_LAYOUT_CLS_BLD = 2   /* [B, 1+L, D]: ViT / CLIP last_hidden_state, token 0 (CLS) dropped */
} capdec_layout;
typedef enum {
  CAPDEC_DT_F32 = 0,
  CAPDEC_DT_BF16 = 1,
  CAPDEC_DT_F16 = 2,
  CAPDEC_DT_P24 = 3           /* BLD only: per image a uint16 [L,D] plane (top 16 bits of the fp32) followed by a uint8 [L,D]
                                 plane (round(low16 / 257)): 3 bytes per element, 16 significant bits (csrc/common.cuh) */
} capdec_dtype;
size_t capdec_source_bytes(int32_t layout, int32_t dtype, int32_t num_images, int32_t num_regions, int32_t feature_dim);
size_t capdec_tiles_bytes(const capdec_handle* h, int32_t num_images, int32_t num_regions);
int capdec_ingest_features(capdec_handle* h, const void* src_dev, int32_t layout, int32_t dtype, int32_t num_images,
                           int32_t num_regions, void* tiles_dev, size_t tiles_bytes, void* stream);
/* capdec_decode_beam on a tile set written by capdec_ingest_features (same outputs, same results as decoding the
 * fp32 [B,L,D] features the tiles were made from) */
int capdec_decode_beam_tiles(capdec_handle* h, const void* tiles_dev, const float* pooled_dev,
                             const uint8_t* key_padding_mask_dev, int32_t num_images, int32_t num_regions,
                             int32_t num_beams, int32_t max_length, float length_penalty,
                             int32_t* out_tokens_dev, int32_t* out_lengths_dev, float* out_scores_dev,
                             float* dbg_top_logprob_dev, int32_t* dbg_top_token_dev, int32_t* dbg_top_beam_dev,
                             void* workspace_dev, size_t workspace_bytes, void* stream);
/* fp32 [L,D] features -> the CAPDEC_DT_P24 source format, on the HOST (for feature caches kept in host memory / on disk) */
int capdec_pack_p24_host(const float* src, int64_t num_images, int64_t elems_per_image, void* dst);

/* ---- host-buffer entry point (end-to-end path) ----------------------------------------------
 * Same as capdec_decode_beam but every buffer is HOST memory (pinned recommended): the call
 * allocates device staging on first use, streams the features host->device in image chunks
 * overlapped with the decode of the previous chunk, copies tokens/lengths/scores back and
 * synchronises before returning.  chunk_images <= 0 selects the default (two images per SM).  This is the call
 * bench.py times for the `e2e` figure. */
int capdec_decode_beam_host(capdec_handle* h, const float* features_host, const float* pooled_host,
                            int32_t num_images, int32_t num_regions, int32_t num_beams,
                            int32_t max_length, float length_penalty, int32_t chunk_images,
                            int32_t* out_tokens_host, int32_t* out_lengths_host, float* out_scores_host);

/* The same with the encoder hand-off formats above and an optional padding mask: features_host holds
 * capdec_source_bytes(layout, dtype, B, L, D) bytes; every chunk is copied in its source format (2-3 bytes per
 * element for bf16 / fp16 / p24 instead of 4) and ingested on the device.  key_padding_mask_host uint8 [B,L] or NULL. */
int capdec_decode_beam_host_ex(capdec_handle* h, const void* features_host, int32_t layout, int32_t dtype,
                               const float* pooled_host, const uint8_t* key_padding_mask_host,
                               int32_t num_images, int32_t num_regions, int32_t num_beams,
                               int32_t max_length, float length_penalty, int32_t chunk_images,
                               int32_t* out_tokens_host, int32_t* out_lengths_host, float* out_scores_host);

/* ---- per-stage device timing ---------------------------------------------------------------------
 * When enabled, every stage launch of subsequent decode calls is bracketed by a cudaEvent pair on the
 * caller's stream.  capdec_stage_times waits for them, returns the summed milliseconds and the number of
 * launches per stage (arrays of CAPDEC_STAGE_COUNT) and resets the log.  Stage ids: 0 prologue, 1 small
 * per-step GEMMs, 2 attention, 3 LSTM gate GEMM, 4 vocab GEMM, 5 top-k/argmax/sampling, 6 beam bookkeeping,
 * 7 reorder gathers. */
#define CAPDEC_STAGE_COUNT 8
int capdec_stage_timing(capdec_handle* h, int32_t enable);
int capdec_stage_times(capdec_handle* h, float* ms_out, int32_t* count_out);

/* ---- stage-level entry points (unit tests / profiling of single kernels) -------------------- */
/* C[M,N] = A[M,K] * W[N,K]^T + bias[N]   (nn.Linear), lda/ldw/ldc in elements. */
int capdec_linear(int32_t precision, const float* a_dev, int64_t lda, const float* w_dev, int64_t ldw,
                  const float* bias_dev, float* c_dev, int64_t ldc, int32_t m, int32_t n, int32_t k,
                  void* stream);
/* per-row log-sum-exp and sorted top-`topk` (logit - lse, index) of logits[R,V] */
int capdec_lse_topk(const float* logits_dev, int64_t ld, int32_t rows, int32_t vocab, int32_t topk,
                    float* out_logprob_dev, int32_t* out_index_dev, float* out_lse_dev, void* stream);

/* The fused form the beam / greedy decoders use on the tensor-core path (replaces fc / output_layer / lm_head followed
 * by log_softmax + topk: models/decoder.py:171, src/models/decoders.py:303,483, HF _beam_search): the GEMM epilogue
 * emits per-256-column-tile {max, sum-exp, top-k} partial records, a merge kernel combines them; the [m,n] logits are
 * never written.  Outputs as capdec_lse_topk.  precision must be a tensor-core mode. */
size_t capdec_linear_topk_workspace(int32_t m, int32_t n, int32_t topk);
int capdec_linear_topk(int32_t precision, const float* a_dev, int64_t lda, const float* w_dev, int64_t ldw,
                       const float* bias_dev, int32_t m, int32_t n, int32_t k, int32_t topk,
                       float* out_logprob_dev, int32_t* out_index_dev, float* out_lse_dev,
                       void* workspace_dev, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CAPDEC_H */
