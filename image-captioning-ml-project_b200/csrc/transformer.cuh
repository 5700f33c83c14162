// Transformer decode-step kernels (transformer.cu).
#pragma once
#include "common.cuh"

namespace capdec {

struct SelfAttnArgs {
  const float* qkv; int64_t ld_qkv;     // [R, 3H]: q | k | v of the current position
  float* cache_k; float* cache_v;       // this layer's cache [R, T, H]
  const int32_t* anc;                   // [R, T] physical row of each earlier position, or nullptr (identity)
  const float* prefix_k; const float* prefix_v; int n_prefix;   // optional per-image prefix keys/values [B, n_prefix, H]
  int rows_per_image;
  float scale;                          // 1/sqrt(head_dim)
  float* out; int64_t ld_out;           // [R, H]
  SplitDst out_split;                   // optional: out also as the split operand of the projection GEMM that follows
  int rows, H, heads, T, t;
  // optional (teacher-forced pass): cache position p of row r is masked as a KEY when key_tok[r*ld_key_tok + p] == key_pad
  // (nn.TransformerDecoder tgt_key_padding_mask, decoders.py:405; HF attention_mask, decoders.py:581)
  const int32_t* key_tok; int64_t ld_key_tok; int key_pad;
};

int embed_pos(const int32_t* tok, const float* emb, const float* pos_row, float* x, int rows, int H, cudaStream_t s,
              const SplitDst* split = nullptr);
int add_layernorm(const float* x, const float* y, const float* gamma, const float* beta, float* sum_out, float* out,
                  int rows, int H, float eps, cudaStream_t s, const SplitDst* split = nullptr);
int self_attn_decode(const SelfAttnArgs& a, cudaStream_t s);
int reorder_ancestors(const int32_t* src, const int32_t* anc_old, int32_t* anc_new, int rows, int T, int t_done,
                      cudaStream_t s);

}  // namespace capdec
