// Kernels for the KV-cached transformer decode step (src/models/decoders.py::TransformerDecoder, nn.TransformerDecoder
// post-LN layers; the same pieces serve the pre-LN GPT-2 block): embedding + learned positions, residual +
// LayerNorm, single-position causal self-attention over a back-pointer-indirected KV cache, and the ancestor-table
// update that replaces copying the cache on beam reorder.  The cross-attention over the hoisted per-layer K/V
// projections of the image regions reuses mha_attention_kernel (attn_mha.cu); all dense layers go through gemm().
#include <stdlib.h>

#include "transformer.cuh"

namespace capdec {
namespace {

// x[r,:] = embedding[tok[r],:] + pos[:]
__global__ void __launch_bounds__(128) embed_pos_kernel(const int32_t* __restrict__ tok, const float* __restrict__ emb,
                                                        const float* __restrict__ pos, float* __restrict__ x, int H,
                                                        const SplitDst split) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x;
  const float4* e = reinterpret_cast<const float4*>(emb + (int64_t)tok[r] * H);
  const float4* p = reinterpret_cast<const float4*>(pos);
  float4* o = reinterpret_cast<float4*>(x + (int64_t)r * H);
  for (int i = threadIdx.x; i < H / 4; i += blockDim.x) {
    const float4 a = e[i], b = p[i];
    const float4 v = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    o[i] = v;
    split_store4(split, r, i * 4, v);
  }
}

// s = x + y (y may be null); optionally store s; out = LayerNorm(s) * gamma + beta   (eps inside the sqrt, biased var).
// One WARP per row, the row held in registers (up to 32 floats per lane, H <= 1024): no shared memory and no block barrier
// -- the block-per-row form spent its time in two __syncthreads around 3 KB of work.  128-bit loads / stores; `split`
// (optional) also writes out as the hi/lo operand copies of the GEMM that consumes it.
constexpr int kLnRowsPerCta = 8;
template <int NV>   // float4 groups per lane: H <= 128 * NV
__global__ void __launch_bounds__(32 * kLnRowsPerCta) add_layernorm_warp_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                                float* __restrict__ sum_out, float* __restrict__ out,
                                                                                int rows, int H, float eps, const SplitDst split) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * kLnRowsPerCta + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int H4 = H >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + (int64_t)r * H);
  const float4* yr = y ? reinterpret_cast<const float4*>(y + (int64_t)r * H) : nullptr;
  float4 v[NV];
  float part = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = lane + 32 * j;
    v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < H4) {
      v[j] = xr[i];
      if (yr) { const float4 w = yr[i]; v[j].x += w.x; v[j].y += w.y; v[j].z += w.z; v[j].w += w.w; }
      part += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
  }
  const float mean = warp_sum(part) / (float)H;
  float var = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    if (lane + 32 * j < H4) {
      const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      var += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(var) / (float)H + eps);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = lane + 32 * j;
    if (i < H4) {
      const float4 g = reinterpret_cast<const float4*>(gamma)[i], bt = reinterpret_cast<const float4*>(beta)[i];
      if (sum_out) reinterpret_cast<float4*>(sum_out + (int64_t)r * H)[i] = v[j];
      const float4 o = make_float4((v[j].x - mean) * rstd * g.x + bt.x, (v[j].y - mean) * rstd * g.y + bt.y,
                                   (v[j].z - mean) * rstd * g.z + bt.z, (v[j].w - mean) * rstd * g.w + bt.w);
      if (out) reinterpret_cast<float4*>(out + (int64_t)r * H)[i] = o;   // (nullptr: only the operand mirror is consumed)
      split_store4(split, r, i * 4, o);
    }
  }
}

// block-per-row form for wider rows (H > 1024)
__global__ void __launch_bounds__(256) add_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float* __restrict__ sum_out, float* __restrict__ out, int H,
                                                            float eps, const SplitDst split) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float srow[];
  __shared__ float s_red[8];
  const int r = blockIdx.x, tid = threadIdx.x, H4 = H >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + (int64_t)r * H);
  const float4* yr = y ? reinterpret_cast<const float4*>(y + (int64_t)r * H) : nullptr;
  float part = 0.f;
  for (int i = tid; i < H4; i += 256) {
    float4 v = xr[i];
    if (yr) { const float4 w = yr[i]; v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w; }
    reinterpret_cast<float4*>(srow)[i] = v;
    part += (v.x + v.y) + (v.z + v.w);
  }
  part = warp_sum(part);
  if ((tid & 31) == 0) s_red[tid >> 5] = part;
  __syncthreads();
  float mean = 0.f;
  for (int w = 0; w < 8; ++w) mean += s_red[w];
  mean /= (float)H;
  __syncthreads();
  float var = 0.f;
  for (int i = tid; i < H4; i += 256) {
    const float4 v = reinterpret_cast<const float4*>(srow)[i];
    const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    var += (a * a + b * b) + (c * c + d * d);
  }
  var = warp_sum(var);
  if ((tid & 31) == 0) s_red[tid >> 5] = var;
  __syncthreads();
  float v = 0.f;
  for (int w = 0; w < 8; ++w) v += s_red[w];
  const float rstd = rsqrtf(v / (float)H + eps);
  for (int i = tid; i < H4; i += 256) {
    const float4 sv = reinterpret_cast<const float4*>(srow)[i];
    const float4 g = reinterpret_cast<const float4*>(gamma)[i], bt = reinterpret_cast<const float4*>(beta)[i];
    if (sum_out) reinterpret_cast<float4*>(sum_out + (int64_t)r * H)[i] = sv;
    const float4 o = make_float4((sv.x - mean) * rstd * g.x + bt.x, (sv.y - mean) * rstd * g.y + bt.y,
                                 (sv.z - mean) * rstd * g.z + bt.z, (sv.w - mean) * rstd * g.w + bt.w);
    if (out) reinterpret_cast<float4*>(out + (int64_t)r * H)[i] = o;
    split_store4(split, r, i * 4, o);
  }
}

// One CTA per row, one warp per head.  Appends this position's K/V to the cache, then attends over
//   [optional per-image prefix (GPT-2 image prefix, shared by the image's rows)] + cache positions 0..t
// where position p < t of row r lives in physical cache row anc[r][p] (back-pointer indirection; identity if null).
__global__ void self_attn_decode_kernel(const SelfAttnArgs a) {
  const int r = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = a.H, d = H / a.heads, T = a.T, t = a.t;
  const float* qkv = a.qkv + (int64_t)r * a.ld_qkv;
  float* kc = a.cache_k + ((int64_t)r * T + t) * H;
  float* vc = a.cache_v + ((int64_t)r * T + t) * H;
  for (int i = threadIdx.x; i < H; i += blockDim.x) { kc[i] = qkv[H + i]; vc[i] = qkv[2 * H + i]; }
  __syncthreads();

  const int hd = warp;
  if (hd >= a.heads) return;
  // lanes own up to 4 strided elements of the head dimension (d <= 128)
  float q[4], acc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int e = lane + 32 * j;
    q[j] = e < d ? qkv[hd * d + e] : 0.f;
    acc[j] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  const int img = r / a.rows_per_image;
  const int n_keys = a.n_prefix + t + 1;
  for (int p = 0; p < n_keys; ++p) {
    const float *kp, *vp;
    if (p < a.n_prefix) {
      kp = a.prefix_k + ((int64_t)img * a.n_prefix + p) * H;
      vp = a.prefix_v + ((int64_t)img * a.n_prefix + p) * H;
    } else {
      const int pos = p - a.n_prefix;
      const int prow = (pos == t || !a.anc) ? r : a.anc[(int64_t)r * T + pos];
      kp = a.cache_k + ((int64_t)prow * T + pos) * H;
      vp = a.cache_v + ((int64_t)prow * T + pos) * H;
    }
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = lane + 32 * j;
      if (e < d) dot = fmaf(q[j], kp[hd * d + e], dot);
    }
    dot = warp_sum(dot) * a.scale;
    if (a.key_tok && p >= a.n_prefix && a.key_tok[(int64_t)r * a.ld_key_tok + (p - a.n_prefix)] == a.key_pad) continue;   // masked key
    const float m_new = fmaxf(m, dot);
    const float corr = expf(m - m_new);       // exp(-inf) = 0 on the first key
    const float w = expf(dot - m_new);
    l = l * corr + w;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = lane + 32 * j;
      if (e < d) acc[j] = acc[j] * corr + w * vp[hd * d + e];
    }
    m = m_new;
  }
  const float inv = 1.f / l;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int e = lane + 32 * j;
    if (e < d) {
      a.out[(int64_t)r * a.ld_out + hd * d + e] = acc[j] * inv;
      split_store1(a.out_split, r, hd * d + e, acc[j] * inv);
    }
  }
}

// Two-pass form for up to 128 keys (every config of the path: prefix 10 + max_len <= 50), one WARP per (row, head) and
// nothing shared between the warps of a CTA: no staging of q through shared memory, no block barrier.  (The CTA-per-row
// form this replaces spent its time in dependent phases -- ancestor lookup, q/k/v staging + barrier, two score rounds,
// four value rounds of scalar loads -- at 16 % DRAM utilisation: 114 us per launch at 5120 rows x 12 heads; this form
// 88 us, issue-bound.  A single-pass online-softmax variant with K and V loads in flight together and no weight / pointer
// shuffles in the value pass was measured SLOWER, 112 us: 70 instead of 48 registers and every lane of a key group
// repeating the key's exp.  Eight-lane key groups with two float4s per lane (four keys per round, 40 % fewer instructions
// per key) were also measured slower, 107 us: the kernel is bound by how many distinct rows a load instruction touches,
// not by its instruction count.)
//   lane = key:  every lane resolves the address of ITS key once (prefix row, ancestor-indirected cache row, or -- for
//                the current position -- this step's projection output itself, so nothing waits for the cache append);
//   pass 1:      a group of GL lanes (8 / 16 / 32 >= head_dim / 4) reads one key's head slice with one 128-bit load per
//                lane, 32 / GL keys per round, 8 rounds in flight; the dot needs log2(GL) shuffle steps; scores pass
//                through the warp's shared-memory strip to land in lane = key; softmax across the lanes;
//   pass 2:      the same lane groups read the value slices (128-bit loads, 8 rounds in flight), each group accumulates
//                its keys, the groups are summed with shuffles at the end.
constexpr int kSaWarps = 4;
template <int NK, int D4T>   // keys per lane (n_keys <= 32 * NK); head_dim / 4 when known at compile time (0 = runtime)
__global__ void __launch_bounds__(32 * kSaWarps) self_attn_decode3_kernel(const SelfAttnArgs a) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_sc[kSaWarps][32 * NK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * kSaWarps + warp;
  if (item >= a.rows * a.heads) return;
  const int r = item / a.heads, hd = item - r * a.heads;
  const int H = a.H, d = H / a.heads, d4 = D4T > 0 ? D4T : (d >> 2), T = a.T, t = a.t;
  const float* qkv = a.qkv + (int64_t)r * a.ld_qkv + hd * d;   // this head's slice of q; k at + H, v at + 2H
  const int img = r / a.rows_per_image;
  const int n_keys = a.n_prefix + t + 1;
  const float* kptr[NK];
  const float* vptr[NK];
#pragma unroll
  for (int i = 0; i < NK; ++i) {
    const int p = lane + 32 * i;
    kptr[i] = qkv + H; vptr[i] = qkv + 2 * H;    // the current position (and a valid address for the unused lanes)
    if (p < n_keys - 1) {
      if (p < a.n_prefix) {
        const int64_t off = ((int64_t)img * a.n_prefix + p) * H + hd * d;
        kptr[i] = a.prefix_k + off; vptr[i] = a.prefix_v + off;
      } else {
        const int pos = p - a.n_prefix;
        const int prow = a.anc ? a.anc[(int64_t)r * T + pos] : r;
        const int64_t off = ((int64_t)prow * T + pos) * H + hd * d;
        kptr[i] = a.cache_k + off; vptr[i] = a.cache_v + off;
      }
    }
  }
  const int GL = d4 <= 8 ? 8 : d4 <= 16 ? 16 : 32;
  const int kpr = 32 / GL, sub = lane % GL, grp = lane / GL;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 qv = sub < d4 ? reinterpret_cast<const float4*>(qkv)[sub] : zero4;
  if (grp == 0 && sub < d4) {   // append this position's K / V head slice to the cache (read by LATER steps only)
    const int64_t off = ((int64_t)r * T + t) * H + hd * d;
    reinterpret_cast<float4*>(a.cache_k + off)[sub] = reinterpret_cast<const float4*>(qkv + H)[sub];
    reinterpret_cast<float4*>(a.cache_v + off)[sub] = reinterpret_cast<const float4*>(qkv + 2 * H)[sub];
  }
  // the address held by lane (p & 31), slot (p >> 5)
  auto ptr_of = [&](const float* const (&tab)[NK], int p) {
    const float* sel = nullptr;
#pragma unroll
    for (int i = 0; i < NK; ++i) {
      const float* cand = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, (unsigned long long)tab[i], p & 31));
      if (i == (p >> 5)) sel = cand;
    }
    return sel;
  };
  float* my_sc = s_sc[warp];
  constexpr int UR = 8;
  // ---- pass 1: scores
  for (int p0 = 0; p0 < n_keys; p0 += UR * kpr) {
    int pk[UR];
    float4 kv[UR];
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      pk[u] = p0 + u * kpr + grp;
      const float* kp = ptr_of(kptr, min(pk[u], n_keys - 1));
      kv[u] = sub < d4 ? reinterpret_cast<const float4*>(kp)[sub] : zero4;
    }
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      float part = fmaf(qv.x, kv[u].x, fmaf(qv.y, kv[u].y, fmaf(qv.z, kv[u].z, qv.w * kv[u].w)));
      for (int o = GL >> 1; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (sub == 0 && pk[u] < n_keys) my_sc[pk[u]] = part * a.scale;
    }
  }
  __syncwarp();
  float sc[NK];
#pragma unroll
  for (int i = 0; i < NK; ++i) {
    const int p = lane + 32 * i;
    sc[i] = p < n_keys ? my_sc[p] : -INFINITY;
    if (a.key_tok && p < n_keys && p >= a.n_prefix && a.key_tok[(int64_t)r * a.ld_key_tok + (p - a.n_prefix)] == a.key_pad)
      sc[i] = -INFINITY;   // masked key: weight exactly 0, as the reference's -inf additive mask
  }
  // ---- softmax across the lanes
  float m = sc[0];
#pragma unroll
  for (int i = 1; i < NK; ++i) m = fmaxf(m, sc[i]);
  m = warp_max(m);
  float l = 0.f;
#pragma unroll
  for (int i = 0; i < NK; ++i) { sc[i] = expf(sc[i] - m); l += sc[i]; }   // exp(-inf) = 0 for the unused slots
  l = warp_sum(l);
  const float inv = 1.f / l;
  // ---- pass 2: weighted values
  float4 acc = zero4;
  for (int p0 = 0; p0 < n_keys; p0 += UR * kpr) {
    float w[UR];
    float4 vv[UR];
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      const int pku = p0 + u * kpr + grp;
      const int p = min(pku, n_keys - 1);
      float wv = 0.f;
#pragma unroll
      for (int i = 0; i < NK; ++i) {
        const float cand = __shfl_sync(0xffffffffu, sc[i], p & 31);
        if (i == (p >> 5)) wv = cand;
      }
      w[u] = pku < n_keys ? wv : 0.f;
      const float* vp = ptr_of(vptr, p);
      vv[u] = sub < d4 ? reinterpret_cast<const float4*>(vp)[sub] : zero4;
    }
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      acc.x = fmaf(w[u], vv[u].x, acc.x); acc.y = fmaf(w[u], vv[u].y, acc.y);
      acc.z = fmaf(w[u], vv[u].z, acc.z); acc.w = fmaf(w[u], vv[u].w, acc.w);
    }
  }
  for (int o = GL; o < 32; o <<= 1) {   // sum the key groups
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
  }
  if (grp == 0 && sub < d4) {
    const float4 o4 = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
    reinterpret_cast<float4*>(a.out + (int64_t)r * a.ld_out + hd * d)[sub] = o4;
    split_store4(a.out_split, r, hd * d + sub * 4, o4);
  }
}

// anc_new[r][p] = (p < t_next) ? (p == t_next-1 ? src[r] : anc_old[src[r]][p]) : -  ; written for p < t_next
__global__ void reorder_ancestors_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ anc_old,
                                         int32_t* __restrict__ anc_new, int rows, int T, int t_done) {
  // after finishing position t_done (its K/V sit in physical row src[r]), build the table for the next step
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * (t_done + 1)) return;
  const int r = i / (t_done + 1), p = i - r * (t_done + 1);
  const int s = src[r];
  anc_new[(int64_t)r * T + p] = (p == t_done) ? s : anc_old[(int64_t)s * T + p];
}

}  // namespace

int embed_pos(const int32_t* tok, const float* emb, const float* pos_row, float* x, int rows, int H, cudaStream_t s,
              const SplitDst* split) {
  if (rows == 0) return CAPDEC_OK;
  CAPDEC_CHECK_CUDA(launch_k(embed_pos_kernel, dim3(rows), dim3(128), 0, s, true, tok, emb, pos_row, x, H, split ? *split : SplitDst{}));
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

int add_layernorm(const float* x, const float* y, const float* gamma, const float* beta, float* sum_out, float* out,
                  int rows, int H, float eps, cudaStream_t s, const SplitDst* split) {
  if (rows == 0) return CAPDEC_OK;
  CAPDEC_REQUIRE(H % 4 == 0, CAPDEC_ERR_UNSUPPORTED, "add_layernorm: H must be a multiple of 4");
  const SplitDst sp = split ? *split : SplitDst{};
  const dim3 grid(ceil_div(rows, kLnRowsPerCta)), block(32 * kLnRowsPerCta);
  if (H <= 256)       CAPDEC_CHECK_CUDA(launch_k(add_layernorm_warp_kernel<2>, grid, block, 0, s, true, x, y, gamma, beta, sum_out, out, rows, H, eps, sp));
  else if (H <= 512)  CAPDEC_CHECK_CUDA(launch_k(add_layernorm_warp_kernel<4>, grid, block, 0, s, true, x, y, gamma, beta, sum_out, out, rows, H, eps, sp));
  else if (H <= 1024) CAPDEC_CHECK_CUDA(launch_k(add_layernorm_warp_kernel<8>, grid, block, 0, s, true, x, y, gamma, beta, sum_out, out, rows, H, eps, sp));
  else CAPDEC_CHECK_CUDA(launch_k(add_layernorm_kernel, dim3(rows), dim3(256), (size_t)H * sizeof(float), s, true, x, y, gamma, beta,
                                  sum_out, out, H, eps, sp));
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

int self_attn_decode(const SelfAttnArgs& a, cudaStream_t s) {
  CAPDEC_REQUIRE(a.heads >= 1 && a.heads <= 32 && a.H % a.heads == 0 && a.H / a.heads <= 128, CAPDEC_ERR_UNSUPPORTED,
                 "self_attn_decode: heads=%d head_dim=%d unsupported (head_dim <= 128, heads <= 32)", a.heads,
                 a.heads ? a.H / a.heads : 0);
  if (a.rows == 0) return CAPDEC_OK;
  const int n_keys = a.n_prefix + a.t + 1;
  if (n_keys <= 128 && (a.H / a.heads) % 4 == 0 && a.ld_qkv % 4 == 0 && a.ld_out % 4 == 0 && a.H % 4 == 0 &&
      (a.out_split.hi == nullptr || a.out_split.ld % 4 == 0) &&
      ((((uintptr_t)a.qkv | (uintptr_t)a.cache_k | (uintptr_t)a.cache_v | (uintptr_t)a.prefix_k | (uintptr_t)a.prefix_v | (uintptr_t)a.out) & 15) == 0)) {
    const int d4 = a.H / a.heads / 4;
    const dim3 grid(ceil_div(a.rows * a.heads, kSaWarps)), block(32 * kSaWarps);
#define CAPDEC_SA_LAUNCH(NKV)                                                                               \
    if (d4 == 16)      CAPDEC_CHECK_CUDA(launch_k(self_attn_decode3_kernel<NKV, 16>, grid, block, 0, s, true, a));  \
    else if (d4 == 24) CAPDEC_CHECK_CUDA(launch_k(self_attn_decode3_kernel<NKV, 24>, grid, block, 0, s, true, a));  \
    else               CAPDEC_CHECK_CUDA(launch_k(self_attn_decode3_kernel<NKV, 0>, grid, block, 0, s, true, a));
    if (n_keys <= 32)      { CAPDEC_SA_LAUNCH(1) }
    else if (n_keys <= 64) { CAPDEC_SA_LAUNCH(2) }
    else                   { CAPDEC_SA_LAUNCH(4) }
#undef CAPDEC_SA_LAUNCH
    CAPDEC_LAUNCH_CHECK();
    return CAPDEC_OK;
  }
  self_attn_decode_kernel<<<a.rows, 32 * a.heads, 0, s>>>(a);
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

int reorder_ancestors(const int32_t* src, const int32_t* anc_old, int32_t* anc_new, int rows, int T, int t_done,
                      cudaStream_t s) {
  const int n = rows * (t_done + 1);
  if (n == 0) return CAPDEC_OK;
  reorder_ancestors_kernel<<<ceil_div(n, 256), 256, 0, s>>>(src, anc_old, anc_new, rows, T, t_done);
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

}  // namespace capdec
