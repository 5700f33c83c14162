// Fused GEMM epilogues shared by the CUDA-core (gemm_ffma.cu) and tcgen05 (gemm_tc.cu) kernels.
#pragma once
#include "common.cuh"

namespace capdec {

template <int EPI, bool FAST = false>
__device__ __forceinline__ void epilogue4(const GemmArgs& p, int m, int n, float v0, float v1, float v2, float v3,
                                          const float4* pre_bias = nullptr) {
  // pre_bias: bias[n..n+3] already in registers (the tcgen05 epilogue fetches a chunk's bias before it waits for the
  // accumulator, so the loads overlap the TMEM read instead of stalling every 4 columns)
  auto sig = [](float x) { return FAST ? sigmoid_fast_(x) : sigmoidf_(x); };
  auto th = [](float x) { return FAST ? tanh_fast_(x) : tanhf(x); };
  // (m, n..n+3) with n % 4 == 0.  Fused epilogues require N % 4 == 0 (checked by the launcher); the plain
  // store family also handles a ragged last group and unaligned C rows (e.g. V = 50257 logits).
  if (m >= p.M || n >= p.N) return;
  if (epi_is_store_family(EPI)) {
    const bool vec_ok = (n + 3 < p.N) && ((p.ldc & 3) == 0) && (!p.C2 || (p.ldc2 & 3) == 0);
    if (!vec_ok) {
      float v[4] = {v0, v1, v2, v3};
      for (int j = 0; j < 4 && n + j < p.N; ++j) {
        float t = v[j] + (p.bias ? p.bias[n + j] : 0.f);
        if (EPI == EPI_SIGMOID_TAIL && n + j >= p.n_split) t = sig(t);
        if (EPI == EPI_TANH) t = th(t);
        if (EPI == EPI_GELU) t = gelu_erf_(t);
        if (EPI == EPI_GELU_TANH) t = FAST ? gelu_tanh_fast_(t) : gelu_tanh_(t);
        if (p.C) p.C[(int64_t)m * p.ldc + n + j] = t;
        if (p.C2) p.C2[(int64_t)m * p.ldc2 + n + j] = t;
        split_store1(p.c_split, m, n + j, t);
      }
      return;
    }
  }
  if (pre_bias) {
    v0 += pre_bias->x; v1 += pre_bias->y; v2 += pre_bias->z; v3 += pre_bias->w;
  } else if (p.bias) {
    const float4 b = *reinterpret_cast<const float4*>(p.bias + n);
    v0 += b.x; v1 += b.y; v2 += b.z; v3 += b.w;
  }
  if (epi_is_store_family(EPI)) {
    if (EPI == EPI_SIGMOID_TAIL && n >= p.n_split) {
      v0 = sig(v0); v1 = sig(v1); v2 = sig(v2); v3 = sig(v3);
    }
    if (EPI == EPI_TANH) { v0 = th(v0); v1 = th(v1); v2 = th(v2); v3 = th(v3); }
    if (EPI == EPI_GELU) { v0 = gelu_erf_(v0); v1 = gelu_erf_(v1); v2 = gelu_erf_(v2); v3 = gelu_erf_(v3); }
    if (EPI == EPI_GELU_TANH) {
      if (FAST) { v0 = gelu_tanh_fast_(v0); v1 = gelu_tanh_fast_(v1); v2 = gelu_tanh_fast_(v2); v3 = gelu_tanh_fast_(v3); }
      else      { v0 = gelu_tanh_(v0); v1 = gelu_tanh_(v1); v2 = gelu_tanh_(v2); v3 = gelu_tanh_(v3); }
    }
    if (p.C) *reinterpret_cast<float4*>(p.C + (int64_t)m * p.ldc + n) = make_float4(v0, v1, v2, v3);
    if (p.C2) *reinterpret_cast<float4*>(p.C2 + (int64_t)m * p.ldc2 + n) = make_float4(v0, v1, v2, v3);
    split_store4(p.c_split, m, n, make_float4(v0, v1, v2, v3));
  } else if (EPI == EPI_LSTM) {
    // torch.nn.LSTMCell: c' = sigmoid(f)*c + sigmoid(i)*tanh(g);  h' = sigmoid(o)*tanh(c')
    if (p.row_table) {
      const float4 tb = *reinterpret_cast<const float4*>(p.row_table + (int64_t)p.row_index[m] * p.ld_table + n);
      v0 += tb.x; v1 += tb.y; v2 += tb.z; v3 += tb.w;
    }
    const int j = n >> 2;
    const float cp = p.c_in[(int64_t)m * p.ldcin + j];
    const float c2 = sig(v1) * cp + sig(v0) * th(v2);
    const float h2 = sig(v3) * th(c2);
    p.c_out[(int64_t)m * p.ldcout + j] = c2;
    p.C[(int64_t)m * p.ldc + j] = h2;
    if (p.C2) p.C2[(int64_t)m * p.ldc2 + j] = h2;
    split_store1(p.c_split, m, j, h2);
  } else if (EPI == EPI_AOA) {
    const int j = n >> 1;
    const float o0 = th(v0) * sig(v1);
    const float o1 = th(v2) * sig(v3);
    *reinterpret_cast<float2*>(p.C + (int64_t)m * p.ldc + j) = make_float2(o0, o1);
    if (p.C2) *reinterpret_cast<float2*>(p.C2 + (int64_t)m * p.ldc2 + j) = make_float2(o0, o1);
  }
}

}  // namespace capdec
