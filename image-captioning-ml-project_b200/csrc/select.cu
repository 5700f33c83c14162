// Token selection (log-softmax + top-k / argmax / inverse-CDF sampling), HF-static beam
// bookkeeping, and the in-place index gathers that reorder decoder state by back-pointer.
#include "select.cuh"
#include "attention.cuh"
#include "ingest.cuh"
#include <limits.h>

namespace capdec {
namespace {

constexpr int kThreads = 256;
constexpr float kNeg = -1.0e9f;

__device__ __forceinline__ bool better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

__device__ __forceinline__ void warp_argmax(float& v, int& i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (better(ov, oi, v, i)) { v = ov; i = oi; }
  }
}

__device__ __forceinline__ float block_max(float v, float* s_tmp) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) s_tmp[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = s_tmp[0];
#pragma unroll
  for (int w = 1; w < kThreads / 32; ++w) r = fmaxf(r, s_tmp[w]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* s_tmp) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) s_tmp[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) r += s_tmp[w];
  __syncthreads();
  return r;
}

// One CTA per row.  The row is staged in shared memory; element i is owned by thread i % 256, which
// caches its best remaining (value, index); each of the K rounds is a block arg-max after which only
// the winner's owner rescans its elements.
__global__ void __launch_bounds__(kThreads) lse_topk_kernel(const float* __restrict__ logits, int64_t ld, int V, int K,
                                                            float* __restrict__ out_lp, int32_t* __restrict__ out_idx,
                                                            float* __restrict__ out_lse) {
  extern __shared__ float row[];
  __shared__ float s_tmp[kThreads / 32];
  __shared__ float s_v[kThreads / 32];
  __shared__ int s_i[kThreads / 32];
  __shared__ int s_win;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* x = logits + (int64_t)blockIdx.x * ld;

  float bv = -INFINITY;
  int bi = INT_MAX;
  for (int i = tid; i < V; i += kThreads) {
    const float v = x[i];
    row[i] = v;
    if (v > bv) { bv = v; bi = i; }
  }
  const float M = block_max(bv, s_tmp);
  float sum = 0.f;
  for (int i = tid; i < V; i += kThreads) sum += expf(row[i] - M);
  const float S = block_sum(sum, s_tmp);
  const float logS = logf(S);
  if (tid == 0 && out_lse) out_lse[blockIdx.x] = M + logS;

  for (int r = 0; r < K; ++r) {
    float v = bv;
    int i = bi;
    warp_argmax(v, i);
    if (lane == 0) { s_v[warp] = v; s_i[warp] = i; }
    __syncthreads();
    if (tid == 0) {
      float wv = s_v[0];
      int wi = s_i[0];
#pragma unroll
      for (int w = 1; w < kThreads / 32; ++w)
        if (better(s_v[w], s_i[w], wv, wi)) { wv = s_v[w]; wi = s_i[w]; }
      const bool valid = wi != INT_MAX;
      out_lp[(int64_t)blockIdx.x * K + r] = valid ? (wv - M) - logS : -INFINITY;
      out_idx[(int64_t)blockIdx.x * K + r] = valid ? wi : -1;
      s_win = valid ? wi : -1;
    }
    __syncthreads();
    const int w = s_win;
    if (w >= 0 && (w % kThreads) == tid) {
      row[w] = -INFINITY;
      bv = -INFINITY;
      bi = INT_MAX;
      for (int j = tid; j < V; j += kThreads) {
        const float t = row[j];
        if (t > bv) { bv = t; bi = j; }
      }
      if (bv == -INFINITY) bi = INT_MAX;  // exhausted
    }
  }
}

// Combine the partial records of the fused vocabulary GEMM (EPI_TOPK, gemm_tc.cu) into the row's log-sum-exp and
// sorted top-K:  one warp per row; lane l owns records l, l+32, ... and keeps a read position per owned record in
// shared memory; each of the K rounds is a warp arg-max over the lanes' best list heads (ties -> lower vocabulary
// index, as torch.topk / the unfused kernel).  Which slots the GEMM wrote follows from its tile schedule (MergeArgs).
constexpr int kMergeMaxRecords = 1024;
struct MergeArgs {
  const float* part; const float* lse_part;
  int n_lse, vocab, n_rec, PS, TKB, rows, K;
  int n_tiles, quota, block_rows;   // the producing GEMM's schedule: tiles per row block, tiles per CTA-group run, rows per block
};
// one warp: merge the records of `row`; writes K sorted (log-prob, index) pairs through out_lp / out_idx (any address space)
constexpr int kMergeStageFloats = 256;   // per-warp staging area: rows with few records (4 x 24 floats at C2) are merged from shared memory
__device__ __forceinline__ void merge_row(const MergeArgs& a, int row, int lane, uint8_t* pos, float* out_lp, int32_t* out_idx,
                                          float* out_lse, float* stage = nullptr) {
  const int PS = a.PS, TKB = a.TKB;
  const float* pr = a.part + (int64_t)row * a.n_rec * PS;
  // slots this row's block actually wrote: one per CTA-group run that intersects its n tiles (x 2 column halves)
  const int blk = row / a.block_rows;
  const int n_rec = 2 * ((((blk + 1) * a.n_tiles - 1) / a.quota) - ((blk * a.n_tiles) / a.quota) + 1);
  // log-sum-exp from the per-(tile half) partials, in a fixed order: lane-strided sequential sums, then the shuffle tree
  const float2* lp2 = reinterpret_cast<const float2*>(a.lse_part) + (int64_t)row * a.n_lse;
  float M = -INFINITY;
  for (int t = lane; t < a.n_lse; t += 32)
    if (t * 128 < a.vocab) M = fmaxf(M, lp2[t].x);
  M = warp_max(M);
  float S = 0.f;
  for (int t = lane; t < a.n_lse; t += 32)
    if (t * 128 < a.vocab) { const float2 v = lp2[t]; S += v.y * expf(v.x - M); }
  S = warp_sum(S);
  for (int t = lane; t < n_rec; t += 32) pos[t] = 0;
  if (stage && n_rec * PS <= kMergeStageFloats) {
    // one coalesced read of the row's records instead of two dependent L2 reads per merge round
    for (int i = lane; i < n_rec * PS; i += 32) stage[i] = pr[i];
    pr = stage;
  }
  __syncwarp();
  const float logS = logf(S);
  if (lane == 0 && out_lse) *out_lse = M + logS;

  // lane-local best head, recomputed only by the lane whose head was taken
  auto scan = [&](float& bv, int& bi, int& bt) {
    bv = -INFINITY; bi = INT_MAX; bt = -1;
    for (int t = lane; t < n_rec; t += 32) {
      const int pp = pos[t];
      if (pp >= TKB) continue;
      const int idx = __float_as_int(pr[(int64_t)t * PS + 2 + TKB + pp]);
      const float v = pr[(int64_t)t * PS + 2 + pp];
      if (idx != INT_MAX && idx >= 0 && better(v, idx, bv, bi)) { bv = v; bi = idx; bt = t; }
    }
  };
  float bv; int bi, bt;
  scan(bv, bi, bt);
  for (int r = 0; r < a.K; ++r) {
    float wv = bv;
    int wi = bi;
    warp_argmax(wv, wi);
    const bool valid = wi != INT_MAX;
    if (lane == 0) {
      out_lp[r] = valid ? (wv - M) - logS : -INFINITY;
      out_idx[r] = valid ? wi : -1;
    }
    if (valid && bi == wi) {   // vocabulary indices are unique, so exactly one lane advances
      pos[bt] += 1;
      scan(bv, bi, bt);
    }
  }
}

__global__ void __launch_bounds__(128) topk_merge_kernel(const MergeArgs a, float* __restrict__ out_lp,
                                                         int32_t* __restrict__ out_idx, float* __restrict__ out_lse) {
  pdl_trigger();
  pdl_wait();
  __shared__ uint8_t s_pos[4][kMergeMaxRecords];
  __shared__ float s_stage[4][kMergeStageFloats];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * 4 + warp;
  if (row >= a.rows) return;
  merge_row(a, row, lane, s_pos[warp], out_lp + (int64_t)row * a.K, out_idx + (int64_t)row * a.K, out_lse ? out_lse + row : nullptr,
            s_stage[warp]);
}

// One CTA per row: inverse-CDF draw in vocabulary index order (double prefix sums), or argmax for the greedy slot.
//   token = #{v : cdf[v] <= u * S},  cdf[v] = sum_{i <= v} exp(x_i - M)            (src/train/trainer.py:423-425 with the draw
//   driven by a shared uniform instead of torch's RNG stream)
// The row is NOT staged in shared memory (a 50257-entry row is 200 KB: one CTA per SM, every phase serialised -- 837 us
// per launch at 3072 rows); it is read three times from L2 / HBM with coalesced loads instead: (1) row max (+ argmax for
// the greedy slot), (2) each warp sums its CONTIGUOUS eighth of the row, lanes striding it, (3) only the warp whose
// segment contains the target walks it in 32-element groups with a warp inclusive scan.  No shared-memory row means
// 8 CTAs per SM and ~0.6 MB of loads in flight per SM.
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__global__ void __launch_bounds__(kThreads) sample_kernel(const float* __restrict__ logits, int64_t ld, int V,
                                                          const float* __restrict__ uniforms, int64_t ld_u, int step,
                                                          int rows_per_image, int greedy_slot,
                                                          int32_t* __restrict__ out_tok, float* __restrict__ out_lp) {
  __shared__ float s_tmp[kThreads / 32];
  __shared__ float s_v[kThreads / 32];
  __shared__ int s_i[kThreads / 32];
  __shared__ double s_wsum[kThreads / 32];
  __shared__ int s_tok;
  constexpr int NW = kThreads / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r = blockIdx.x;
  const float* x = logits + (int64_t)r * ld;

  float bv = -INFINITY;
  int bi = INT_MAX;
  for (int i = tid; i < V; i += kThreads) {
    const float v = x[i];
    if (v > bv) { bv = v; bi = i; }
  }
  const float M = block_max(bv, s_tmp);
  const bool greedy = greedy_slot >= 0 && (r % rows_per_image) == greedy_slot;

  // contiguous segment per warp (a multiple of 32 elements), lanes striding it: prefix sums follow index order
  const int seg = ((V + NW - 1) / NW + 31) & ~31;
  const int beg = min(V, warp * seg), end = min(V, beg + seg);
  double local = 0.0;
  for (int i = beg + lane; i < end; i += 32) local += (double)expf(x[i] - M);
  local = warp_sum_d(local);
  if (lane == 0) s_wsum[warp] = local;
  __syncthreads();
  double S = 0.0, before = 0.0;
#pragma unroll
  for (int w = 0; w < NW; ++w) { if (w == warp) before = S; S += s_wsum[w]; }

  if (greedy) {  // block-uniform branch
    float v = bv; int i = bi;
    warp_argmax(v, i);
    if (lane == 0) { s_v[warp] = v; s_i[warp] = i; }
    __syncthreads();
    if (tid == 0) {
      float wv = s_v[0]; int wi = s_i[0];
#pragma unroll
      for (int w = 1; w < NW; ++w)
        if (better(s_v[w], s_i[w], wv, wi)) { wv = s_v[w]; wi = s_i[w]; }
      s_tok = wi;
    }
  } else {
    const double target = (double)uniforms[(int64_t)r * ld_u + step] * S;
    // the one warp whose segment holds the first cdf value above the target finds it; if no value is above it (u*S
    // rounds to >= the total) the draw is the last token, as torch's clamp does
    // (before(w+1) == before(w) + s_wsum[w] exactly, so at most one warp qualifies)
    const bool mine = before <= target && before + s_wsum[warp] > target;
    if (tid == 0) s_tok = V - 1;
    __syncthreads();
    if (mine && beg < end) {
      double base = before;
      int tok = -1;
      for (int g = beg; g < end && tok < 0; g += 32) {
        const int i = g + lane;
        double inc = i < end ? (double)expf(x[i] - M) : 0.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        const unsigned below = __ballot_sync(0xffffffffu, i < end && base + inc <= target);
        const int n_valid = min(32, end - g);
        const int cnt = __popc(below);
        if (cnt < n_valid) tok = g + cnt;       // cdf is non-decreasing: the first element above the target
        base += __shfl_sync(0xffffffffu, inc, 31);
      }
      // (the scan re-associates the segment's sum: if its last value lands a rounding below the target, the first element
      //  above it is the next segment's first)
      if (lane == 0) s_tok = min(tok >= 0 ? tok : end, V - 1);
    }
  }
  __syncthreads();
  if (tid == 0) {
    const int tok = s_tok;
    out_tok[r] = tok;
    if (out_lp) out_lp[r] = (x[tok] - M) - logf((float)S);
  }
}

// The same draw from the per-(row, 128-column half tile) {max, sum exp(x - max)} partials the logits GEMM's epilogue
// leaves next to the logits (tensor-core modes): the row is NOT re-read -- one warp per row sums the n_lse partials in
// index order (double), finds the half tile whose cdf range holds the target, and scans only that half tile's 128
// logits.  The greedy slot takes the first position of the row maximum the same way.  6.5 -> 0.4 ms per decode at
// configs[4] (3072 rows x 50257 logits per step: 617 MB read three times per step before).
__global__ void __launch_bounds__(128) sample_partials_kernel(const float* __restrict__ logits, int64_t ld, int V,
                                                              const float* __restrict__ lse_part, int n_lse,
                                                              const float* __restrict__ uniforms, int64_t ld_u, int step,
                                                              int rows, int rows_per_image, int greedy_slot,
                                                              int32_t* __restrict__ out_tok, float* __restrict__ out_lp) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* x = logits + (int64_t)r * ld;
  const float2* pp = reinterpret_cast<const float2*>(lse_part) + (int64_t)r * n_lse;
  const int n_half = (V + 127) / 128;                 // half tiles that hold columns (<= n_lse)
  const int nb = (n_half + 31) / 32;                  // contiguous half tiles per lane: prefix sums follow index order
  const int t0 = min(n_half, lane * nb), t1 = min(n_half, t0 + nb);
  float bm = -INFINITY;
  int bt = INT_MAX;
  for (int t = t0; t < t1; ++t) { const float m = pp[t].x; if (m > bm) { bm = m; bt = t; } }   // first half tile of the lane's maximum
  float M = bm; int mt = bt;
  warp_argmax(M, mt);                                 // row maximum and the first half tile that reaches it
  const bool greedy = greedy_slot >= 0 && (r % rows_per_image) == greedy_slot;
  double local = 0.0;
  for (int t = t0; t < t1; ++t) { const float2 v = pp[t]; local += (double)v.y * (double)expf(v.x - M); }
  double incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  const double S = __shfl_sync(0xffffffffu, incl, 31);
  int half_tile = mt;
  double base = 0.0, target = 0.0;
  if (!greedy) {
    target = (double)uniforms[(int64_t)r * ld_u + step] * S;
    const double before = incl - local;
    // the lane whose block holds the first cdf value above the target walks it; no lane (u * S rounds to >= the total):
    // the draw is the last token, as torch's clamp does
    int found = -1;
    double fbase = 0.0;
    if (before <= target && incl > target) {
      double b = before;
      for (int t = t0; t < t1; ++t) {
        const float2 v = pp[t];
        const double pt = (double)v.y * (double)expf(v.x - M);
        if (b + pt > target || t == t1 - 1) { found = t; fbase = b; break; }
        b += pt;
      }
    }
    const unsigned who = __ballot_sync(0xffffffffu, found >= 0);
    if (who == 0u) {
      if (lane == 0) { out_tok[r] = V - 1; if (out_lp) out_lp[r] = (x[V - 1] - M) - logf((float)S); }
      return;
    }
    const int src = __ffs(who) - 1;
    half_tile = __shfl_sync(0xffffffffu, found, src);
    base = __shfl_sync(0xffffffffu, fbase, src);
  }
  int tok = -1;
  const int beg = half_tile * 128, end = min(V, beg + 128);
  for (int g = beg; g < end && tok < 0; g += 32) {
    const int i = g + lane;
    const float xv = i < end ? x[i] : -INFINITY;
    if (greedy) {
      const unsigned hit = __ballot_sync(0xffffffffu, xv == M);
      if (hit) tok = g + __ffs(hit) - 1;
    } else {
      double inc = i < end ? (double)expf(xv - M) : 0.0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
      }
      const unsigned below = __ballot_sync(0xffffffffu, i < end && base + inc <= target);
      const int n_valid = min(32, end - g);
      const int cnt = __popc(below);
      if (cnt < n_valid) tok = g + cnt;       // cdf is non-decreasing: the first element above the target
      base += __shfl_sync(0xffffffffu, inc, 31);
    }
  }
  // (the half tile's partial is an fp32 sum in another order: if its re-summed last value lands a rounding below the
  //  target, the first element above it is the next half tile's first)
  tok = min(tok >= 0 ? tok : end, V - 1);
  if (lane == 0) {
    out_tok[r] = tok;
    if (out_lp) out_lp[r] = (x[tok] - M) - logf((float)S);
  }
}

// lp[r] = logits[r, tok[r]] - logsumexp(logits[r, :])   (log-prob of a GIVEN token: re-scoring of sampled captions,
// src/train/trainer.py:423-428 without the draw).  One CTA per row, two passes over the row (it sits in L2).
__global__ void __launch_bounds__(kThreads) token_logprob_kernel(const float* __restrict__ logits, int64_t ld, int V,
                                                                 const int32_t* __restrict__ tok, int64_t ld_tok,
                                                                 float* __restrict__ out_lp, int64_t ld_out) {
  __shared__ float s_tmp[kThreads / 32];
  const int r = blockIdx.x, tid = threadIdx.x;
  const float* x = logits + (int64_t)r * ld;
  float bv = -INFINITY;
  for (int i = tid; i < V; i += kThreads) bv = fmaxf(bv, x[i]);
  const float M = block_max(bv, s_tmp);
  float sum = 0.f;
  for (int i = tid; i < V; i += kThreads) sum += expf(x[i] - M);
  const float S = block_sum(sum, s_tmp);
  if (tid == 0) {
    const int t = tok[(int64_t)r * ld_tok];
    out_lp[(int64_t)r * ld_out] = (t >= 0 && t < V) ? (x[t] - M) - logf(S) : -INFINITY;
  }
}

// everything after a row's first EOS becomes pad; length = tokens kept.  One thread per row (T is 20-50).
__global__ void trim_at_eos_kernel(const int32_t* __restrict__ tok, int64_t ld, int rows, int T, int eos, int pad, int keep_eos,
                                   int32_t* __restrict__ out, int64_t ld_out, int32_t* __restrict__ out_len) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  int len = T;
  for (int t = 0; t < T; ++t)
    if (tok[(int64_t)r * ld + t] == eos) { len = min(T, t + (keep_eos ? 1 : 0)); break; }
  if (out)
    for (int t = 0; t < T; ++t) out[(int64_t)r * ld_out + t] = t < len ? tok[(int64_t)r * ld + t] : pad;
  if (out_len) out_len[r] = len;
}

// ---- beam bookkeeping: one thread per image -------------------------------------------------------
__global__ void beam_init_kernel(BeamState st, int B, int k, int T, int bos, int fill) {
  const int img = blockIdx.x * blockDim.x + threadIdx.x;
  if (img >= B) return;
  for (int b = 0; b < k; ++b) {
    const int64_t o = ((int64_t)img * k + b) * T;
    for (int t = 0; t < T; ++t) {
      const int v = t == 0 ? bos : fill;
      st.run_seq[0][o + t] = v; st.run_seq[1][o + t] = v;
      st.fin_seq[0][o + t] = v; st.fin_seq[1][o + t] = v;
    }
    st.run_score[img * k + b] = b == 0 ? 0.f : kNeg;
    st.fin_score[img * k + b] = kNeg;
    st.fin_len[img * k + b] = 1;
    st.fin_flag[img * k + b] = 0;
  }
  st.unsatisfied[img] = 1;
}

// one WARP per image: every lane evaluates the (tiny) selection logic redundantly on the same inputs, the token
// sequences are copied lane-parallel, lane 0 writes the scalars.  cand_lp / cand_idx: the image's k sorted lists of 2k
// candidates, cand[(b * 2k) + j] (shared memory).  The beam width is a template parameter and every per-image array is
// indexed by unrolled loop counters only, so the whole merge state lives in registers (the run-time-k form kept 656 bytes
// of it in local memory).
template <int K>
__device__ __forceinline__ void beam_step_image(const BeamState& st, int img, int lane, int T, int cur_len, int eos,
                                                float div_fin, float div_heur, const float* cand_lp, const int32_t* cand_idx,
                                                int32_t* next_tok, int32_t* src_row, float* dbg_lp, int32_t* dbg_tok,
                                                int32_t* dbg_beam) {
  constexpr int k = K, k2 = 2 * K;
  const int par = (cur_len - 1) & 1;  // read buffers [par], write [par ^ 1]
  // (ternaries, not st.run_seq[par]: a run-time index into a by-value kernel parameter forces a local-memory copy of it)
  const int32_t* run_old = (par ? st.run_seq[1] : st.run_seq[0]) + (int64_t)img * k * T;
  int32_t* run_new = (par ? st.run_seq[0] : st.run_seq[1]) + (int64_t)img * k * T;
  const int32_t* fin_old = (par ? st.fin_seq[1] : st.fin_seq[0]) + (int64_t)img * k * T;
  int32_t* fin_new = (par ? st.fin_seq[0] : st.fin_seq[1]) + (int64_t)img * k * T;

  // (c) top-2k continuations over the k*V accumulated log-probs: k-way merge of the per-row sorted lists
  float score[k];
  int ptr[k];
#pragma unroll
  for (int b = 0; b < k; ++b) { score[b] = st.run_score[img * k + b]; ptr[b] = 0; }
  float top_lp[k2];
  int top_tok[k2], top_beam[k2];
#pragma unroll
  for (int j = 0; j < k2; ++j) {
    float bv = -INFINITY;
    int bb = -1, bt = 0;
#pragma unroll
    for (int b = 0; b < k; ++b) {
      if (ptr[b] < k2) {
        const int o = b * k2 + ptr[b];
        const int tok = cand_idx[o];
        if (tok >= 0) {
          const float v = cand_lp[o] + score[b];
          if (bb < 0 || v > bv) { bv = v; bb = b; bt = tok; }  // ties keep the lower flat index (lower beam first)
        }
      }
    }
    top_lp[j] = bv; top_tok[j] = bt; top_beam[j] = bb < 0 ? 0 : bb;
#pragma unroll
    for (int b = 0; b < k; ++b) ptr[b] += (b == bb) ? 1 : 0;
    if (dbg_lp && lane == 0) {
      const int64_t o = (int64_t)img * k2 + j;
      dbg_lp[o] = bv; dbg_tok[o] = bt; dbg_beam[o] = top_beam[j];
    }
  }

  // (d) stopping criteria on the extended sequences: MaxLength | Eos
  const bool at_max = cur_len + 1 >= T;
  unsigned hits = 0u;
  float run_lp[k2];
#pragma unroll
  for (int j = 0; j < k2; ++j) {
    const bool hit = at_max || top_tok[j] == eos;
    hits |= (hit ? 1u : 0u) << j;
    run_lp[j] = top_lp[j] + (hit ? kNeg : 0.f);
  }

  // (e) next running beams: top-k of run_lp (stable)
  unsigned used = 0u;
  float new_score[k];
#pragma unroll
  for (int i = 0; i < k; ++i) {
    int bj = -1;
    float bvv = 0.f;
    int stok = 0, sbeam = 0;
#pragma unroll
    for (int j = 0; j < k2; ++j)
      if (!((used >> j) & 1u) && (bj < 0 || run_lp[j] > bvv)) { bj = j; bvv = run_lp[j]; stok = top_tok[j]; sbeam = top_beam[j]; }
    used |= 1u << bj;
    new_score[i] = bvv;
    const int32_t* srcp = run_old + (int64_t)sbeam * T;
    int32_t* dstp = run_new + (int64_t)i * T;
    for (int t = lane; t < T; t += 32) dstp[t] = t == cur_len ? stok : srcp[t];
    if (lane == 0) {
      next_tok[img * k + i] = stok;
      src_row[img * k + i] = img * k + sbeam;
    }
  }

  // (f) finished beams: merge the k finished with the 2k candidates, keep the best k (stable)
  const bool unsat = st.unsatisfied[img] != 0;
  float m_sc[k + k2];
#pragma unroll
  for (int b = 0; b < k; ++b) m_sc[b] = st.fin_score[img * k + b];
#pragma unroll
  for (int j = 0; j < k2; ++j) {
    const bool just_fin = ((hits >> j) & 1u) && j < k;
    float v = top_lp[j] / div_fin;
    v += unsat ? 0.f : kNeg;
    v += just_fin ? 0.f : kNeg;
    m_sc[k + j] = v;
  }
  unsigned m_used = 0u;
  float f_sc[k];
  int f_len[k];
  unsigned f_flags = 0u;
#pragma unroll
  for (int i = 0; i < k; ++i) {
    int bj = -1;
    float bvv = 0.f;
#pragma unroll
    for (int j = 0; j < k + k2; ++j)
      if (!((m_used >> j) & 1u) && (bj < 0 || m_sc[j] > bvv)) { bj = j; bvv = m_sc[j]; }
    m_used |= 1u << bj;
    f_sc[i] = bvv;
    int32_t* dstp = fin_new + (int64_t)i * T;
    if (bj < k) {
      const int32_t* srcp = fin_old + (int64_t)bj * T;
      for (int t = lane; t < T; t += 32) dstp[t] = srcp[t];
      f_len[i] = st.fin_len[img * k + bj];
      f_flags |= (st.fin_flag[img * k + bj] ? 1u : 0u) << i;
    } else {
      int stok = 0, sbeam = 0;
      bool fin = false;
#pragma unroll
      for (int j = 0; j < k2; ++j)
        if (j == bj - k) { stok = top_tok[j]; sbeam = top_beam[j]; fin = ((hits >> j) & 1u) && j < k; }
      const int32_t* srcp = run_old + (int64_t)sbeam * T;
      for (int t = lane; t < T; t += 32) dstp[t] = t == cur_len ? stok : srcp[t];
      f_len[i] = cur_len + 1;
      f_flags |= (fin ? 1u : 0u) << i;
    }
  }
  __syncwarp();   // every lane has read the old per-image state before lane 0 overwrites it
  float fmin = f_sc[0];
#pragma unroll
  for (int i = 0; i < k; ++i) {
    if (lane == 0) {
      st.fin_score[img * k + i] = f_sc[i];
      st.fin_len[img * k + i] = f_len[i];
      st.fin_flag[img * k + i] = (uint8_t)((f_flags >> i) & 1u);
      st.run_score[img * k + i] = new_score[i];
    }
    fmin = fminf(fmin, f_sc[i]);
  }

  // (g) early-stop heuristic (early_stopping=False): can the best running beam still beat the worst finished?
  const float best_possible = new_score[0] / div_heur;
  bool any = false;
#pragma unroll
  for (int i = 0; i < k; ++i) any = any || (best_possible > (((f_flags >> i) & 1u) ? fmin : kNeg));
  if (lane == 0) st.unsatisfied[img] = (unsat && any) ? 1 : 0;
}

template <int K>
__global__ void __launch_bounds__(128) beam_step_kernel(BeamState st, int B, int T, int cur_len, int eos, float div_fin,
                                                        float div_heur, const float* __restrict__ cand_lp,
                                                        const int32_t* __restrict__ cand_idx, int32_t* __restrict__ next_tok,
                                                        int32_t* __restrict__ src_row, float* dbg_lp, int32_t* dbg_tok,
                                                        int32_t* dbg_beam) {
  pdl_trigger();
  pdl_wait();
  // the image's k sorted candidate lists (k * 2k <= 128 entries) are read once, lane-parallel, into shared memory; the
  // sequential k-way merge then runs on shared-memory latency instead of ten rounds of dependent L2 reads
  __shared__ float s_lp[4][2 * K * K];
  __shared__ int32_t s_idx[4][2 * K * K];
  const int img = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (img >= B) return;
  constexpr int n = 2 * K * K;
  for (int i = lane; i < n; i += 32) {
    s_lp[w][i] = cand_lp[(int64_t)img * n + i];
    s_idx[w][i] = cand_idx[(int64_t)img * n + i];
  }
  __syncwarp();
  beam_step_image<K>(st, img, lane, T, cur_len, eos, div_fin, div_heur, s_lp[w], s_idx[w], next_tok, src_row, dbg_lp, dbg_tok,
                     dbg_beam);
}

__global__ void beam_finalize_kernel(BeamState st, int parity, int B, int k, int T, int32_t* out_tok, int32_t* out_len,
                                     float* out_score) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * T) return;
  const int img = i / T, t = i - img * T;
  out_tok[i] = (parity ? st.fin_seq[1] : st.fin_seq[0])[((int64_t)img * k) * T + t];
  if (t == 0) {
    out_len[img] = st.fin_len[img * k];
    out_score[img] = st.fin_score[img * k];
  }
}

// ---- gathers ----------------------------------------------------------------------------------------
__device__ __forceinline__ void gather_row(const GatherArgs& a, int r, int tid, int nthreads) {
  const int src = a.src ? a.src[r] : r;
  if (a.tok) {
    const int tok = a.tok[r];
    if (a.tok_out && tid == 0 && a.pos >= 0) a.tok_out[(int64_t)r * a.ld_tok + a.pos] = tok;
    if (a.embedding) {
      const float4* e = reinterpret_cast<const float4*>(a.embedding + (int64_t)tok * a.E);
      float4* d = reinterpret_cast<float4*>(a.x_emb + (int64_t)r * a.ld_x);
      for (int i = tid; i < a.E / 4; i += nthreads) {
        const float4 v = e[i];
        d[i] = v;
        split_store4(a.x_split, r, i * 4, v);
      }
    }
  }
  for (int sidx = 0; sidx < a.n_state; ++sidx) {
    const float4* sp = reinterpret_cast<const float4*>(a.state_src[sidx] + (int64_t)src * a.ld_src[sidx]);
    float4* dp = reinterpret_cast<float4*>(a.state_dst[sidx] + (int64_t)r * a.ld_dst[sidx]);
    const int scol = a.x_split.hi ? a.state_split_col[sidx] : -1;
    for (int i = tid; i < a.width[sidx] / 4; i += nthreads) {
      const float4 v = sp[i];
      dp[i] = v;
      if (scol >= 0) split_store4(a.x_split, r, scol + i * 4, v);
    }
  }
}

__global__ void __launch_bounds__(128) gather_rows_kernel(const GatherArgs a) {
  pdl_trigger();
  pdl_wait();
  gather_row(a, blockIdx.x, threadIdx.x, blockDim.x);
}

// Per-image fusion of the three bookkeeping kernels of a beam step (fused top-k path) for SMALL batches, where each of
// them sits on its launch-latency floor (15 + 16 + 10 us at 512 images): the image's k rows are merged from the
// vocabulary GEMM's records into k sorted candidate lists in shared memory (one warp per row), warp 0 runs the HF beam
// step on them, then the whole CTA gathers the image's new rows (state reorder by back-pointer + embedding).  Beams
// reorder only inside an image, so no other CTA's results are needed.  From 1024 images up the three specialised kernels
// win (the phases serialise inside a CTA), so the launcher uses this form only up to kFusedSelectMaxImages (select.cuh).
template <int K>
__global__ void __launch_bounds__(128) select_fused_kernel(const MergeArgs ma, const BeamState st, int T, int cur_len, int eos,
                                                           float div_fin, float div_heur, int32_t* next_tok, int32_t* src_row,
                                                           float* dbg_lp, int32_t* dbg_tok, int32_t* dbg_beam,
                                                           const GatherArgs ga, int do_gather) {
  pdl_trigger();
  pdl_wait();
  __shared__ uint8_t s_pos[4][kMergeMaxRecords];
  __shared__ float s_stage[4][kMergeStageFloats];
  __shared__ float s_lp[2 * K * K];
  __shared__ int32_t s_idx[2 * K * K];
  const int img = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int k2 = 2 * K;
  for (int b = warp; b < K; b += 4)
    merge_row(ma, img * K + b, lane, s_pos[warp], s_lp + b * k2, s_idx + b * k2, nullptr, s_stage[warp]);
  __syncthreads();
  if (warp == 0)
    beam_step_image<K>(st, img, lane, T, cur_len, eos, div_fin, div_heur, s_lp, s_idx, next_tok, src_row, dbg_lp, dbg_tok, dbg_beam);
  if (!do_gather) return;
  __threadfence_block();
  __syncthreads();   // next_tok / src_row of this image's rows are visible to the whole CTA
  for (int b = 0; b < K; ++b) gather_row(ga, img * K + b, threadIdx.x, blockDim.x);
}

__global__ void __launch_bounds__(256) mean_regions_kernel(const float* __restrict__ feats, int L, int D,
                                                           float* __restrict__ out) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;  // float4 column
  if (c >= D / 4) return;
  const float4* p = reinterpret_cast<const float4*>(feats + (int64_t)b * L * D) + c;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int l = 0; l < L; ++l) {
    const float4 v = ldg_stream(p + (int64_t)l * (D / 4));
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  const float fl = (float)L;
  reinterpret_cast<float4*>(out + (int64_t)b * D)[c] = make_float4(acc.x / fl, acc.y / fl, acc.z / fl, acc.w / fl);
}

__global__ void __launch_bounds__(128) expand_rows_kernel(const float* __restrict__ src, int64_t ld_src,
                                                          float* __restrict__ dst, int64_t ld_dst, int k, int width) {
  const int r = blockIdx.x;
  const float4* sp = reinterpret_cast<const float4*>(src + (int64_t)(r / k) * ld_src);
  float4* dp = reinterpret_cast<float4*>(dst + (int64_t)r * ld_dst);
  for (int i = threadIdx.x; i < width / 4; i += blockDim.x) dp[i] = sp[i];
}

__global__ void fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void fill_f32_kernel(float* p, int64_t n, float v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace

int lse_topk(const float* logits, int64_t ld, int rows, int vocab, int topk, float* out_lp, int32_t* out_idx,
             float* out_lse, cudaStream_t s) {
  CAPDEC_REQUIRE(topk >= 1 && topk <= kMaxTopK, CAPDEC_ERR_UNSUPPORTED, "lse_topk: topk %d not in [1,%d]", topk, kMaxTopK);
  CAPDEC_REQUIRE(vocab >= 1 && (size_t)vocab * 4 <= 220 * 1024, CAPDEC_ERR_UNSUPPORTED,
                 "lse_topk: vocab %d does not fit a shared-memory row", vocab);
  if (rows == 0) return CAPDEC_OK;
  const size_t smem = (size_t)vocab * sizeof(float);
  if (smem > 48 * 1024)
    CAPDEC_CHECK_CUDA(cudaFuncSetAttribute(lse_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  lse_topk_kernel<<<rows, kThreads, smem, s>>>(logits, ld, vocab, topk, out_lp, out_idx, out_lse);
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

int topk_merge(const float* part, const float* lse_part, int rows, int vocab, int n_total, int part_k, int topk,
               float* out_lp, int32_t* out_idx, float* out_lse, cudaStream_t s) {
  const int n_rec = tk_records(rows, n_total);
  CAPDEC_REQUIRE(tk_supported(vocab, part_k) && n_rec <= kMergeMaxRecords && topk >= 1 && topk <= tk_bucket(part_k), CAPDEC_ERR_INVALID,
                 "topk_merge: topk %d exceeds the partial list length %d (vocab %d)", topk, tk_bucket(part_k), vocab);
  if (rows == 0) return CAPDEC_OK;
  MergeArgs ma{part, lse_part, tk_lse_pairs(vocab), vocab, n_rec, tk_stride(part_k), tk_bucket(part_k), rows, topk, 0, 0, 0};
  tk_schedule(rows, n_total, &ma.n_tiles, &ma.quota, &ma.block_rows);
  CAPDEC_CHECK_CUDA(launch_k(topk_merge_kernel, dim3(ceil_div(rows, 4)), dim3(128), 0, s, true, ma, out_lp, out_idx, out_lse));
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

int sample_rows(const float* logits, int64_t ld, int rows, int vocab, const float* uniforms, int64_t ld_u, int step,
                int rows_per_image, int greedy_slot, int32_t* out_tok, float* out_lp, cudaStream_t s) {
  CAPDEC_REQUIRE(vocab >= 1, CAPDEC_ERR_INVALID, "sample_rows: empty vocabulary");
  if (rows == 0) return CAPDEC_OK;
  sample_kernel<<<rows, kThreads, 0, s>>>(logits, ld, vocab, uniforms, ld_u, step, rows_per_image, greedy_slot, out_tok, out_lp);
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

int sample_rows_partials(const float* logits, int64_t ld, int rows, int vocab, const float* lse_part, const float* uniforms,
                         int64_t ld_u, int step, int rows_per_image, int greedy_slot, int32_t* out_tok, float* out_lp, cudaStream_t s) {
  CAPDEC_REQUIRE(vocab >= 1 && lse_part, CAPDEC_ERR_INVALID, "sample_rows_partials: empty vocabulary / no partials");
  if (rows == 0) return CAPDEC_OK;
  CAPDEC_CHECK_CUDA(launch_k(sample_partials_kernel, dim3(ceil_div(rows, 4)), dim3(128), 0, s, true, logits, ld, vocab, lse_part,
                             tk_lse_pairs(vocab), uniforms, ld_u, step, rows, rows_per_image, greedy_slot, out_tok, out_lp));
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

int token_logprob(const float* logits, int64_t ld, int rows, int vocab, const int32_t* tok, int64_t ld_tok, float* out_lp,
                  int64_t ld_out, cudaStream_t s) {
  if (rows == 0) return CAPDEC_OK;
  token_logprob_kernel<<<rows, kThreads, 0, s>>>(logits, ld, vocab, tok, ld_tok, out_lp, ld_out);
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

int trim_at_eos(const int32_t* tok, int64_t ld, int rows, int T, int eos, int pad, int keep_eos, int32_t* out, int64_t ld_out,
                int32_t* out_len, cudaStream_t s) {
  if (rows == 0) return CAPDEC_OK;
  trim_at_eos_kernel<<<ceil_div(rows, 128), 128, 0, s>>>(tok, ld, rows, T, eos, pad, keep_eos, out, ld_out, out_len);
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

int beam_init(const BeamState& st, int B, int k, int T, int bos, int fill, cudaStream_t s) {
  if (B == 0) return CAPDEC_OK;
  beam_init_kernel<<<ceil_div(B, 128), 128, 0, s>>>(st, B, k, T, bos, fill);
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

int beam_step(const BeamState& st, int B, int k, int T, int V, int cur_len, int eos, float len_div_finished,
              float len_div_heuristic, const float* cand_lp, const int32_t* cand_idx, int32_t* next_tok,
              int32_t* src_row, float* dbg_lp, int32_t* dbg_tok, int32_t* dbg_beam, cudaStream_t s) {
  CAPDEC_REQUIRE(k >= 1 && k <= kMaxRowsPerImage, CAPDEC_ERR_UNSUPPORTED, "beam_step: num_beams %d not in [1,%d]", k,
                 kMaxRowsPerImage);
  (void)V;
  if (B == 0) return CAPDEC_OK;
#define CAPDEC_BEAM_CASE(KV)                                                                                              \
  case KV:                                                                                                                \
    CAPDEC_CHECK_CUDA(launch_k(beam_step_kernel<KV>, dim3(ceil_div(B, 4)), dim3(128), 0, s, true, st, B, T, cur_len, eos,  \
                               len_div_finished, len_div_heuristic, cand_lp, cand_idx, next_tok, src_row, dbg_lp, dbg_tok, \
                               dbg_beam));                                                                                \
    break;
  switch (k) {
    CAPDEC_BEAM_CASE(1) CAPDEC_BEAM_CASE(2) CAPDEC_BEAM_CASE(3) CAPDEC_BEAM_CASE(4)
    CAPDEC_BEAM_CASE(5) CAPDEC_BEAM_CASE(6) CAPDEC_BEAM_CASE(7) CAPDEC_BEAM_CASE(8)
  }
#undef CAPDEC_BEAM_CASE
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

int beam_finalize(const BeamState& st, int parity, int B, int k, int T, int32_t* out_tok, int32_t* out_len,
                  float* out_score, cudaStream_t s) {
  if (B == 0) return CAPDEC_OK;
  beam_finalize_kernel<<<ceil_div(B * T, 256), 256, 0, s>>>(st, parity, B, k, T, out_tok, out_len, out_score);
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

int select_fused(const float* part, const float* lse_part, int vocab, int n_total, int part_k, const BeamState& st, int B, int k,
                 int T, int cur_len, int eos, float div_fin, float div_heur, int32_t* next_tok, int32_t* src_row,
                 float* dbg_lp, int32_t* dbg_tok, int32_t* dbg_beam, const GatherArgs* ga, cudaStream_t s) {
  const int rows = B * k, n_rec = tk_records(rows, n_total);
  CAPDEC_REQUIRE(k >= 1 && k <= kMaxRowsPerImage && 2 * k <= tk_bucket(part_k) && n_rec <= kMergeMaxRecords, CAPDEC_ERR_INVALID,
                 "select_fused: num_beams %d / record layout unsupported", k);
  if (B == 0) return CAPDEC_OK;
  MergeArgs ma{part, lse_part, tk_lse_pairs(vocab), vocab, n_rec, tk_stride(part_k), tk_bucket(part_k), rows, 2 * k, 0, 0, 0};
  tk_schedule(rows, n_total, &ma.n_tiles, &ma.quota, &ma.block_rows);
  GatherArgs g{};
  if (ga) g = *ga;
#define CAPDEC_SF_CASE(KV)                                                                                              \
  case KV:                                                                                                              \
    CAPDEC_CHECK_CUDA(launch_k(select_fused_kernel<KV>, dim3(B), dim3(128), 0, s, true, ma, st, T, cur_len, eos, div_fin, \
                               div_heur, next_tok, src_row, dbg_lp, dbg_tok, dbg_beam, g, ga ? 1 : 0));                 \
    break;
  switch (k) {
    CAPDEC_SF_CASE(1) CAPDEC_SF_CASE(2) CAPDEC_SF_CASE(3) CAPDEC_SF_CASE(4)
    CAPDEC_SF_CASE(5) CAPDEC_SF_CASE(6) CAPDEC_SF_CASE(7) CAPDEC_SF_CASE(8)
  }
#undef CAPDEC_SF_CASE
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

int gather_rows(const GatherArgs& a, cudaStream_t s) {
  CAPDEC_REQUIRE(a.n_state <= 16, CAPDEC_ERR_INVALID, "gather_rows: too many state tensors");
  if (a.rows == 0) return CAPDEC_OK;
  CAPDEC_CHECK_CUDA(launch_k(gather_rows_kernel, dim3(a.rows), dim3(128), 0, s, true, a));
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

int mean_regions(const float* feats, int B, int L, int D, float* out, cudaStream_t s) {
  CAPDEC_REQUIRE(D % 4 == 0, CAPDEC_ERR_UNSUPPORTED, "mean_regions: D must be a multiple of 4");
  if (B == 0) return CAPDEC_OK;
  dim3 grid(ceil_div(D / 4, 256), B);
  mean_regions_kernel<<<grid, 256, 0, s>>>(feats, L, D, out);
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

int mean_regions_split(const float* feats, int B, int L, int D, float* out, const SplitDst& split, cudaStream_t s) {
  // the row-major fp32 case of the encoder hand-off pass (ingest.cu): region mean + operand copies in one read of feats
  CAPDEC_REQUIRE(D % 8 == 0 && split.ld % 8 == 0, CAPDEC_ERR_UNSUPPORTED, "mean_regions_split: D must be a multiple of 8");
  IngestArgs a{};
  a.src = feats; a.layout = CAPDEC_LAYOUT_BLD; a.dtype = CAPDEC_DT_F32; a.B = B; a.L = L; a.D = D; a.split = split; a.mean = out;
  return ingest_features(a, s);
}

int expand_rows(const float* src, int64_t ld_src, float* dst, int64_t ld_dst, int rows, int k, int width, cudaStream_t s) {
  CAPDEC_REQUIRE(width % 4 == 0 && ld_src % 4 == 0 && ld_dst % 4 == 0, CAPDEC_ERR_INVALID, "expand_rows: widths must be multiples of 4");
  if (rows == 0) return CAPDEC_OK;
  expand_rows_kernel<<<rows, 128, 0, s>>>(src, ld_src, dst, ld_dst, k, width);
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

int fill_i32(int32_t* p, int64_t n, int32_t v, cudaStream_t s) {
  if (n == 0) return CAPDEC_OK;
  fill_i32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, n, v);
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}
int fill_f32(float* p, int64_t n, float v, cudaStream_t s) {
  if (n == 0) return CAPDEC_OK;
  fill_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, n, v);
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

}  // namespace capdec
