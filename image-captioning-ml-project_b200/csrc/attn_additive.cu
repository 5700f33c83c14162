// Additive attention (scores -> softmax -> context) for all rows of one image per CTA.
// HBM-bound: per image-step it streams att1 [L,A] and feats [L,D] exactly once (128-bit
// read-only no-L1-allocate loads, 4-8 independent loads in flight per thread); the k rows of an
// image reuse each loaded element from registers.  Several CTAs are resident per SM so one
// image's score phase overlaps another's context phase.
#include "attention.cuh"

namespace capdec {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kRowsPerIter = 4;  // region rows handled together by a warp in the score phase

template <int ACT>
__device__ __forceinline__ float act_fn(float x) {
  return ACT == ACT_RELU ? fmaxf(x, 0.f) : ACT == ACT_TANH_FAST ? tanh_fast_(x) : tanhf(x);
}

template <int KB, int ACT>
__global__ void __launch_bounds__(kThreads) additive_attention_kernel(const AddAttnArgs p) {
  extern __shared__ __align__(16) float smem[];
  const int A = p.A, L = p.L, D = p.D, k = p.k;
  const int Lp = (L + 3) & ~3;
  float* s_att2 = smem;                    // [KB][A]
  float* s_w = s_att2 + KB * A;            // [A]
  float* s_e = s_w + A;                    // [KB][Lp]   scores, then alpha
  float* s_red = s_e + KB * Lp;            // [G-1][KB][D] partial contexts (only when G > 1)

  const int img = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row0 = (int64_t)img * k;

  // ---- phase 0: stage the k query projections and the energy vector
  const int A4 = A >> 2;
  for (int i = tid; i < KB * A4; i += kThreads) {
    const int b = i / A4, a4 = i - b * A4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b < k) v = *reinterpret_cast<const float4*>(p.att2 + (p.row_src ? p.row_src[row0 + b] : row0 + b) * p.ld_att2 + a4 * 4);
    reinterpret_cast<float4*>(s_att2)[b * A4 + a4] = v;
  }
  for (int i = tid; i < A4; i += kThreads) reinterpret_cast<float4*>(s_w)[i] = reinterpret_cast<const float4*>(p.w)[i];
  __syncthreads();

  // ---- phase 1: scores.  A warp owns kRowsPerIter region rows; lanes split the attention dim.
  const float* att1 = p.att1 + (int64_t)img * L * A;
  for (int l0 = warp * kRowsPerIter; l0 < L; l0 += kWarps * kRowsPerIter) {
    float acc[kRowsPerIter][KB];
#pragma unroll
    for (int r = 0; r < kRowsPerIter; ++r)
#pragma unroll
      for (int b = 0; b < KB; ++b) acc[r][b] = 0.f;
    for (int a4 = lane; a4 < A4; a4 += 32) {
      float4 x[kRowsPerIter];
#pragma unroll
      for (int r = 0; r < kRowsPerIter; ++r) {
        const int l = min(l0 + r, L - 1);
        x[r] = ldg_stream(reinterpret_cast<const float4*>(att1 + (int64_t)l * A) + a4);
      }
      const float4 wv = reinterpret_cast<const float4*>(s_w)[a4];
#pragma unroll
      for (int b = 0; b < KB; ++b) {
        const float4 q = reinterpret_cast<const float4*>(s_att2)[b * A4 + a4];
#pragma unroll
        for (int r = 0; r < kRowsPerIter; ++r) {
          float t = acc[r][b];
          t = fmaf(wv.x, act_fn<ACT>(x[r].x + q.x), t);
          t = fmaf(wv.y, act_fn<ACT>(x[r].y + q.y), t);
          t = fmaf(wv.z, act_fn<ACT>(x[r].z + q.z), t);
          t = fmaf(wv.w, act_fn<ACT>(x[r].w + q.w), t);
          acc[r][b] = t;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kRowsPerIter; ++r) {
      const int l = l0 + r;
#pragma unroll
      for (int b = 0; b < KB; ++b) {
        const float v = warp_sum(acc[r][b]);
        if (lane == 0 && l < L) {
          float e = (v + p.w_bias) / p.temperature;
          if (p.mask && p.mask[(int64_t)img * L + l]) e = -1.0e9f;
          s_e[b * Lp + l] = e;
        }
      }
    }
  }
  __syncthreads();

  // ---- phase 2: softmax over regions, one warp per row
  for (int b = warp; b < k; b += kWarps) {
    float* e = s_e + b * Lp;
    float m = -INFINITY;
    for (int l = lane; l < L; l += 32) m = fmaxf(m, e[l]);
    m = warp_max(m);
    float sum = 0.f;
    for (int l = lane; l < L; l += 32) {
      const float v = expf(e[l] - m);
      e[l] = v;
      sum += v;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    float* aout = p.alpha ? p.alpha + (row0 + b) * p.ld_alpha : nullptr;
    for (int l = lane; l < L; l += 32) {
      const float v = e[l] * inv;
      e[l] = v;
      if (aout) aout[l] = v;
    }
    for (int l = L + lane; l < Lp; l += 32) e[l] = 0.f;
  }
  __syncthreads();

  // ---- phase 3: context.  Threads own float4 feature columns; if D/4 <= blockDim/2 the region
  // range is split over G thread groups and reduced through shared memory.
  const int D4 = D >> 2;
  const float* feats = p.feats + (int64_t)img * L * D;
  int G = 1;
  while (D4 * G * 2 <= kThreads) G *= 2;

  auto accumulate = [&](int c, int lstart, int lstride, float4 (&acc)[KB]) {
#pragma unroll
    for (int b = 0; b < KB; ++b) acc[b] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int l = lstart; l < L; l += lstride) {
      float4 x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int ll = min(l + u, L - 1);
        x[u] = ldg_stream(reinterpret_cast<const float4*>(feats + (int64_t)ll * D) + c);
      }
#pragma unroll
      for (int b = 0; b < KB; ++b) {
        const float4 al = *reinterpret_cast<const float4*>(s_e + b * Lp + l);  // alpha is zero beyond L
        acc[b].x = fmaf(al.x, x[0].x, acc[b].x); acc[b].y = fmaf(al.x, x[0].y, acc[b].y);
        acc[b].z = fmaf(al.x, x[0].z, acc[b].z); acc[b].w = fmaf(al.x, x[0].w, acc[b].w);
        acc[b].x = fmaf(al.y, x[1].x, acc[b].x); acc[b].y = fmaf(al.y, x[1].y, acc[b].y);
        acc[b].z = fmaf(al.y, x[1].z, acc[b].z); acc[b].w = fmaf(al.y, x[1].w, acc[b].w);
        acc[b].x = fmaf(al.z, x[2].x, acc[b].x); acc[b].y = fmaf(al.z, x[2].y, acc[b].y);
        acc[b].z = fmaf(al.z, x[2].z, acc[b].z); acc[b].w = fmaf(al.z, x[2].w, acc[b].w);
        acc[b].x = fmaf(al.w, x[3].x, acc[b].x); acc[b].y = fmaf(al.w, x[3].y, acc[b].y);
        acc[b].z = fmaf(al.w, x[3].z, acc[b].z); acc[b].w = fmaf(al.w, x[3].w, acc[b].w);
      }
    }
  };
  auto emit = [&](int c, int b, float4 v) {
    if (p.gate) {
      const float4 gt = *reinterpret_cast<const float4*>(p.gate + (p.row_src ? p.row_src[row0 + b] : row0 + b) * p.ld_gate + c * 4);
      v.x *= gt.x; v.y *= gt.y; v.z *= gt.z; v.w *= gt.w;
    }
    if (p.ctx) *reinterpret_cast<float4*>(p.ctx + (row0 + b) * p.ld_ctx + c * 4) = v;
    split_store4(p.ctx_split, row0 + b, p.ctx_split_col + c * 4, v);
  };

  if (G == 1) {
    for (int c = tid; c < D4; c += kThreads) {
      float4 acc[KB];
      accumulate(c, 0, 4, acc);
#pragma unroll
      for (int b = 0; b < KB; ++b)
        if (b < k) emit(c, b, acc[b]);
    }
  } else {
    const int g = tid / D4, c = tid - g * D4;
    const bool active = g < G;
    float4 acc[KB];
    if (active) {
      accumulate(c, g * 4, G * 4, acc);
      if (g > 0) {
#pragma unroll
        for (int b = 0; b < KB; ++b) reinterpret_cast<float4*>(s_red)[((g - 1) * KB + b) * D4 + c] = acc[b];
      }
    }
    __syncthreads();
    if (active && g == 0) {
#pragma unroll
      for (int b = 0; b < KB; ++b) {
        if (b >= k) continue;
        float4 v = acc[b];
        for (int gg = 1; gg < G; ++gg) {
          const float4 t = reinterpret_cast<const float4*>(s_red)[((gg - 1) * KB + b) * D4 + c];
          v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
        }
        emit(c, b, v);
      }
    }
  }
}

template <int KB>
int launch_kb(const AddAttnArgs& a, int act, cudaStream_t s) {
  const int Lp = (a.L + 3) & ~3;
  const int D4 = a.D / 4;
  int G = 1;
  while (D4 * G * 2 <= kThreads) G *= 2;
  size_t smem = sizeof(float) * ((size_t)KB * a.A + a.A + (size_t)KB * Lp + (size_t)(G - 1) * KB * a.D);
  CAPDEC_REQUIRE(smem <= 200 * 1024, CAPDEC_ERR_UNSUPPORTED, "additive_attention: shared memory %zu B too large", smem);
  auto kern = act == ACT_RELU ? additive_attention_kernel<KB, ACT_RELU>
            : act == ACT_TANH_FAST ? additive_attention_kernel<KB, ACT_TANH_FAST> : additive_attention_kernel<KB, ACT_TANH>;
  if (smem > 48 * 1024) CAPDEC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<a.B, kThreads, smem, s>>>(a);
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

}  // namespace

int additive_attention(const AddAttnArgs& a, int act, cudaStream_t s) {
  CAPDEC_REQUIRE(a.k >= 1 && a.k <= kMaxRowsPerImage, CAPDEC_ERR_UNSUPPORTED,
                 "additive_attention: rows per image %d not in [1,%d]", a.k, kMaxRowsPerImage);
  CAPDEC_REQUIRE(a.A % 4 == 0 && a.D % 4 == 0 && a.L >= 1, CAPDEC_ERR_UNSUPPORTED,
                 "additive_attention: A and D must be multiples of 4 (A=%d D=%d)", a.A, a.D);
  CAPDEC_REQUIRE(a.ld_att2 % 4 == 0 && a.ld_ctx % 4 == 0 && (!a.gate || a.ld_gate % 4 == 0), CAPDEC_ERR_INVALID,
                 "additive_attention: row strides must be multiples of 4");
  if (a.B == 0) return CAPDEC_OK;
  const int took = additive_attention_stream(a, act, s);   // persistent TMA-streamed kernel for the shapes it covers
  if (took != 0) return took > 0 ? CAPDEC_OK : took;
  CAPDEC_REQUIRE(!a.tile_fmt, CAPDEC_ERR_UNSUPPORTED, "additive_attention: bf16 / p24 tiles need the streaming kernel (A=%d D=%d L=%d k=%d)",
                 a.A, a.D, a.L, a.k);
  switch (a.k) {
    case 1: return launch_kb<1>(a, act, s);
    case 2: return launch_kb<2>(a, act, s);
    case 3: return launch_kb<3>(a, act, s);
    case 4: return launch_kb<4>(a, act, s);
    case 5: return launch_kb<5>(a, act, s);
    case 6: return launch_kb<6>(a, act, s);
    default: return launch_kb<8>(a, act, s);
  }
}

}  // namespace capdec
