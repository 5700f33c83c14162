// Token selection + hypothesis bookkeeping kernels (select.cu).
#pragma once
#include "common.cuh"

namespace capdec {

constexpr int kMaxTopK = 16;  // 2 * max beams

// per-row log-sum-exp and sorted top-K of logits[R,V]:  out_lp = (x - max) - log(sum exp(x - max))
int lse_topk(const float* logits, int64_t ld, int rows, int vocab, int topk, float* out_lp, int32_t* out_idx,
             float* out_lse, cudaStream_t s);

// fused path: merge the EPI_TOPK partial records [rows, tk_records(rows, vocab), tk_stride(part_k)] written by the vocabulary
// GEMM (gemm_tc.cu) into the same outputs as lse_topk; the logits themselves never exist in HBM
// (n_total = the GEMM's full N: vocabulary columns plus an optional projection tail, see GemmArgs::tk_vocab)
int topk_merge(const float* part, const float* lse_part, int rows, int vocab, int n_total, int part_k, int topk,
               float* out_lp, int32_t* out_idx, float* out_lse, cudaStream_t s);

// inverse-CDF draw per row: token = #{v : cdf[v] <= u}; row r uses uniforms[r*ld_u + step].
// rows whose (r % rows_per_image) == greedy_slot take the argmax instead (greedy_slot < 0: none).
int sample_rows(const float* logits, int64_t ld, int rows, int vocab, const float* uniforms, int64_t ld_u, int step,
                int rows_per_image, int greedy_slot, int32_t* out_tok, float* out_lp, cudaStream_t s);
// the same draw from the {max, sum exp} partials the logits GEMM leaves per (row, 128-column half tile) (tensor-core modes)
int sample_rows_partials(const float* logits, int64_t ld, int rows, int vocab, const float* lse_part, const float* uniforms,
                         int64_t ld_u, int step, int rows_per_image, int greedy_slot, int32_t* out_tok, float* out_lp, cudaStream_t s);

// out_lp[r*ld_out] = log_softmax(logits[r,:])[tok[r*ld_tok]]
int token_logprob(const float* logits, int64_t ld, int rows, int vocab, const int32_t* tok, int64_t ld_tok, float* out_lp,
                  int64_t ld_out, cudaStream_t s);
// pad everything after the first EOS of each row; out_len = tokens kept (out / out_len may be nullptr)
int trim_at_eos(const int32_t* tok, int64_t ld, int rows, int T, int eos, int pad, int keep_eos, int32_t* out, int64_t ld_out,
                int32_t* out_len, cudaStream_t s);

struct BeamState {          // all device pointers; [B,k,...] row-major
  int32_t* run_seq[2];      // [B,k,T] ping-pong
  int32_t* fin_seq[2];      // [B,k,T] ping-pong
  float* run_score;         // [B,k]
  float* fin_score;         // [B,k]
  int32_t* fin_len;         // [B,k]
  uint8_t* fin_flag;        // [B,k]
  uint8_t* unsatisfied;     // [B]
};
int beam_init(const BeamState& st, int B, int k, int T, int bos, int fill, cudaStream_t s);
// one HF-static beam step for every image (transformers _beam_search steps c-g); parity = step index
int beam_step(const BeamState& st, int B, int k, int T, int V, int cur_len, int eos, float len_div_finished,
              float len_div_heuristic, const float* cand_lp, const int32_t* cand_idx, int32_t* next_tok,
              int32_t* src_row, float* dbg_lp, int32_t* dbg_tok, int32_t* dbg_beam, cudaStream_t s);
int beam_finalize(const BeamState& st, int parity, int B, int k, int T, int32_t* out_tok, int32_t* out_len,
                  float* out_score, cudaStream_t s);

// next-step operands:  x[r, 0:E] = embedding[tok[r]];  h_dst[r] = h_src[src[r]] (per layer);  c likewise.
// src == nullptr means identity.  tok_out (optional) records tok at tok_out[r*ld_tok + pos].
struct GatherArgs {
  const int32_t* tok; const int32_t* src;
  const float* embedding; int E; float* x_emb; int64_t ld_x;       // embedding rows into X
  int n_state;                                                     // number of (src,dst) state pairs
  const float* state_src[16]; float* state_dst[16]; int64_t ld_src[16]; int64_t ld_dst[16]; int width[16];
  int32_t* tok_out; int64_t ld_tok; int pos;
  int rows;
  // optional split mirrors of X (the gate GEMM's operand): embedding at column 0, state i at column state_split_col[i]
  SplitDst x_split; int state_split_col[16];   // state_split_col[i] < 0: state i is not part of X
};
int gather_rows(const GatherArgs& a, cudaStream_t s);

// small-batch form of topk_merge + beam_step + gather_rows in one per-image kernel (fused top-k path; ga == nullptr: no
// gather, last step).  Same results as the three kernels.  Measured on B200 (ms per decode, fused vs three kernels):
// 64 images 3.71 vs 3.85, 256: 5.13 vs 5.27, 512: 8.16 vs 8.22, 1024: 14.77 vs 14.39, 2048: 27.7 vs 26.8.
constexpr int kFusedSelectMaxImages = 512;
int select_fused(const float* part, const float* lse_part, int vocab, int n_total, int part_k, const BeamState& st, int B, int k,
                 int T, int cur_len, int eos, float div_fin, float div_heur, int32_t* next_tok, int32_t* src_row,
                 float* dbg_lp, int32_t* dbg_tok, int32_t* dbg_beam, const GatherArgs* ga, cudaStream_t s);

// mean over regions: out[b,:] = mean_l feats[b,l,:]
int mean_regions(const float* feats, int B, int L, int D, float* out, cudaStream_t s);
// the same mean, fused with writing the hi/lo operand copies of feats ([B*L, D], row pitch split.ld) in one pass
int mean_regions_split(const float* feats, int B, int L, int D, float* out, const SplitDst& split, cudaStream_t s);
// dst[r, :] = src[r / k, :]  (expand per-image rows to per-beam rows), width % 4 == 0
int expand_rows(const float* src, int64_t ld_src, float* dst, int64_t ld_dst, int rows, int k, int width, cudaStream_t s);
int fill_i32(int32_t* p, int64_t n, int32_t v, cudaStream_t s);
int fill_f32(float* p, int64_t n, float v, cudaStream_t s);

}  // namespace capdec
