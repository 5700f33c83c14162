// Attention stage kernels: argument blocks + launchers (attn_additive.cu, attn_mha.cu).
#pragma once
#include "common.cuh"

namespace capdec {

constexpr int kMaxRowsPerImage = 8;  // beams / samples of one image that share its feature tiles in one CTA

// ACT_TANH_FAST: tanh as 1 - 2/(1 + e^{2x}) on the MUFU pipe (~1e-6 absolute error), used by the tensor-core precision modes:
// with exact tanhf the 196 x H x k evaluations per image-step make the soft-attention kernel ALU-bound at 4x its HBM time
enum AddAct : int { ACT_RELU = 0, ACT_TANH = 1, ACT_TANH_FAST = 2 };

// Additive attention over one image's region tiles, all `k` rows (beams) of the image at once:
//   e[b,l]   = (w . act(att1[img,l,:] + att2[row_b,:]) + w_bias) / temperature   (masked -> -1e9)
//   alpha    = softmax_l(e)
//   ctx[b,:] = sum_l alpha[b,l] * feats[img,l,:]   (* gate[row_b,:] if gate != nullptr)
// legacy:  models/decoder.py:152-161 (act = relu, gate = sigmoid(f_beta(h)))
// soft:    src/models/attention.py:76-111 (act = tanh, value == key == features)
struct AddAttnArgs {
  const float* att1;                    // [B,L,A]  hoisted enc_att(enc) / key_proj(key), bias included
  const float* att2; int64_t ld_att2;   // [R,A]    dec_att(h) / query_proj(q), bias included
  const int32_t* row_src;               // [R] or nullptr: att2 / gate of row r are read from row row_src[r] (the projections were
                                        // computed before the beam reorder; the back-pointer is applied here instead of by a copy)
  const float* w;                       // [A]      att.weight / energy.weight
  float w_bias, temperature;
  const uint8_t* mask;                  // [B,L] 1 = padding, or nullptr
  const float* feats;                   // [B,L,D]
  int tile_fmt;                        // tile format (streaming kernel only).  0: fp32.  1: att1 and feats point at bf16 tiles of
                                        // the same shapes (bf16 mode: half the bytes).  2: "p24" planes (common.cuh; bf16x3 mode:
                                        // three quarters of the bytes): att1 / feats point at the 16-bit planes,
  const uint8_t* att1_b8;               //    att1_b8 / feats_b8 at the byte planes
  const uint8_t* feats_b8;
  const float* gate; int64_t ld_gate;   // [R,D] or nullptr
  float* ctx; int64_t ld_ctx;           // [R,D]  (may be nullptr when ctx_split carries the only consumer's copy)
  SplitDst ctx_split; int ctx_split_col; // optional: ctx also/only as the split GEMM operand, at column ctx_split_col
  float* alpha; int64_t ld_alpha;       // [R,L] (row stride ld_alpha) or nullptr
  int B, L, A, D, k;
};
int additive_attention(const AddAttnArgs& a, int act, cudaStream_t s);
// persistent TMA-streamed form (attn_stream.cu): returns 1 if it took the call, 0 if the shape is left to the generic
// kernel in attn_additive.cu, < 0 on error
int additive_attention_stream(const AddAttnArgs& a, int act, cudaStream_t s);
// whether the streaming kernel covers a shape (callers that want bf16 tiles must know before they lay out the workspace)
bool additive_attention_stream_supports(int A, int D, int L, int k, int tile_fmt);

// Multi-head dot-product attention on hoisted per-image K/V projections, all k rows of an image per CTA:
//   s[b,h,l] = q[row_b,h,:] . K[img,l,h,:] / denom      (masked -> -1e9)
//   p        = softmax_l(s);  out[row_b,h,:] = sum_l p[b,h,l] * V[img,l,h,:]
//   alpha[row_b,l] = mean_h p[b,h,l]
// src/models/attention.py:161-211 (the output_proj GEMM follows outside).
struct MhaArgs {
  const float* q; int64_t ld_q;         // [R,H] projected query
  const float* kproj;                   // [B,L,H] hoisted key_proj(key), row stride ld_kv
  const float* vproj;                   // [B,L,H] hoisted value_proj(value), row stride ld_kv
  int64_t ld_kv;
  const uint8_t* mask;                  // [B,L] or nullptr
  float denom;                          // temperature * sqrt(head_dim)
  float* out; int64_t ld_out;           // [R,H] concatenated heads
  float* alpha; int64_t ld_alpha;       // [R,L] head-mean weights or nullptr
  int B, L, H, heads, k;
};
int mha_attention(const MhaArgs& a, cudaStream_t s);
// persistent TMA-streamed form (attn_mha_stream.cu) for dense [L,H] K/V tiles: 1 = took the call, 0 = left to the generic kernel
int mha_attention_stream(const MhaArgs& a, cudaStream_t s);

}  // namespace capdec
