// Kernels for the KV-cached transformer decode step (src/models/decoders.py::TransformerDecoder, nn.TransformerDecoder
// post-LN layers; the same pieces serve the pre-LN GPT-2 block): embedding + learned positions, residual +
// LayerNorm, single-position causal self-attention over a back-pointer-indirected KV cache, and the ancestor-table
// update that replaces copying the cache on beam reorder.  The cross-attention over the hoisted per-layer K/V
// projections of the image regions reuses mha_attention_kernel (attn_mha.cu); all dense layers go through gemm().
#include "transformer.cuh"

namespace capdec {
namespace {

// x[r,:] = embedding[tok[r],:] + pos[:]
__global__ void __launch_bounds__(128) embed_pos_kernel(const int32_t* __restrict__ tok, const float* __restrict__ emb,
                                                        const float* __restrict__ pos, float* __restrict__ x, int H) {
  const int r = blockIdx.x;
  const float4* e = reinterpret_cast<const float4*>(emb + (int64_t)tok[r] * H);
  const float4* p = reinterpret_cast<const float4*>(pos);
  float4* o = reinterpret_cast<float4*>(x + (int64_t)r * H);
  for (int i = threadIdx.x; i < H / 4; i += blockDim.x) {
    const float4 a = e[i], b = p[i];
    o[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}

// s = x + y (y may be null); optionally store s; out = LayerNorm(s) * gamma + beta   (eps inside the sqrt, biased var)
__global__ void __launch_bounds__(256) add_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float* __restrict__ sum_out, float* __restrict__ out, int H,
                                                            float eps) {
  extern __shared__ float srow[];
  __shared__ float s_red[8];
  const int r = blockIdx.x, tid = threadIdx.x;
  const float* xr = x + (int64_t)r * H;
  const float* yr = y ? y + (int64_t)r * H : nullptr;
  float part = 0.f;
  for (int i = tid; i < H; i += 256) {
    const float v = xr[i] + (yr ? yr[i] : 0.f);
    srow[i] = v;
    part += v;
  }
  part = warp_sum(part);
  if ((tid & 31) == 0) s_red[tid >> 5] = part;
  __syncthreads();
  float mean = 0.f;
  for (int w = 0; w < 8; ++w) mean += s_red[w];
  mean /= (float)H;
  __syncthreads();
  float var = 0.f;
  for (int i = tid; i < H; i += 256) { const float d = srow[i] - mean; var += d * d; }
  var = warp_sum(var);
  if ((tid & 31) == 0) s_red[tid >> 5] = var;
  __syncthreads();
  float v = 0.f;
  for (int w = 0; w < 8; ++w) v += s_red[w];
  const float rstd = rsqrtf(v / (float)H + eps);
  for (int i = tid; i < H; i += 256) {
    const float s = srow[i];
    if (sum_out) sum_out[(int64_t)r * H + i] = s;
    out[(int64_t)r * H + i] = (s - mean) * rstd * gamma[i] + beta[i];
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
    if (e < d) a.out[(int64_t)r * a.ld_out + hd * d + e] = acc[j] * inv;
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

int embed_pos(const int32_t* tok, const float* emb, const float* pos_row, float* x, int rows, int H, cudaStream_t s) {
  if (rows == 0) return CAPDEC_OK;
  embed_pos_kernel<<<rows, 128, 0, s>>>(tok, emb, pos_row, x, H);
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

int add_layernorm(const float* x, const float* y, const float* gamma, const float* beta, float* sum_out, float* out,
                  int rows, int H, float eps, cudaStream_t s) {
  if (rows == 0) return CAPDEC_OK;
  add_layernorm_kernel<<<rows, 256, (size_t)H * sizeof(float), s>>>(x, y, gamma, beta, sum_out, out, H, eps);
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

int self_attn_decode(const SelfAttnArgs& a, cudaStream_t s) {
  CAPDEC_REQUIRE(a.heads >= 1 && a.heads <= 32 && a.H % a.heads == 0 && a.H / a.heads <= 128, CAPDEC_ERR_UNSUPPORTED,
                 "self_attn_decode: heads=%d head_dim=%d unsupported (head_dim <= 128, heads <= 32)", a.heads,
                 a.heads ? a.H / a.heads : 0);
  if (a.rows == 0) return CAPDEC_OK;
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
