// Multi-head dot-product cross-attention for one decode step, all k rows of an image per CTA
// (src/models/attention.py:161-211 with the key/value projections hoisted out of the step loop).
// HBM-bound: streams the image's projected K [L,H] and V [L,H] tiles once per image-step.
#include "attention.cuh"

namespace capdec {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kRowsPerIter = 4;

template <int KB>
__global__ void __launch_bounds__(kThreads) mha_attention_kernel(const MhaArgs p) {
  extern __shared__ __align__(16) float smem[];
  const int L = p.L, H = p.H, heads = p.heads, k = p.k;
  const int d = H / heads, d4 = d >> 2, H4 = H >> 2;
  const int Lp = (L + 3) & ~3;
  float* s_q = smem;                        // [KB][H]
  float* s_p = s_q + KB * H;                // [KB][heads][Lp]
  float* s_red = s_p + KB * heads * Lp;     // [G-1][KB][H]

  const int img = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row0 = (int64_t)img * k;

  for (int i = tid; i < KB * H4; i += kThreads) {
    const int b = i / H4, c = i - b * H4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b < k) v = *reinterpret_cast<const float4*>(p.q + (row0 + b) * p.ld_q + c * 4);
    reinterpret_cast<float4*>(s_q)[i] = v;
  }
  __syncthreads();

  // ---- scores
  const float* K = p.kproj + (int64_t)img * L * p.ld_kv;
  for (int l0 = warp * kRowsPerIter; l0 < L; l0 += kWarps * kRowsPerIter) {
    for (int hd = 0; hd < heads; ++hd) {
      float acc[kRowsPerIter][KB];
#pragma unroll
      for (int r = 0; r < kRowsPerIter; ++r)
#pragma unroll
        for (int b = 0; b < KB; ++b) acc[r][b] = 0.f;
      for (int c = lane; c < d4; c += 32) {
        float4 x[kRowsPerIter];
#pragma unroll
        for (int r = 0; r < kRowsPerIter; ++r) {
          const int l = min(l0 + r, L - 1);
          x[r] = ldg_stream(reinterpret_cast<const float4*>(K + (int64_t)l * p.ld_kv + hd * d) + c);
        }
#pragma unroll
        for (int b = 0; b < KB; ++b) {
          const float4 q = reinterpret_cast<const float4*>(s_q + b * H + hd * d)[c];
#pragma unroll
          for (int r = 0; r < kRowsPerIter; ++r) {
            float t = acc[r][b];
            t = fmaf(q.x, x[r].x, t); t = fmaf(q.y, x[r].y, t);
            t = fmaf(q.z, x[r].z, t); t = fmaf(q.w, x[r].w, t);
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
            float e = v / p.denom;
            if (p.mask && p.mask[(int64_t)img * L + l]) e = -1.0e9f;
            s_p[(b * heads + hd) * Lp + l] = e;
          }
        }
      }
    }
  }
  __syncthreads();

  // ---- softmax per (row, head)
  for (int i = warp; i < k * heads; i += kWarps) {
    float* e = s_p + i * Lp;
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
    for (int l = lane; l < L; l += 32) e[l] *= inv;
    for (int l = L + lane; l < Lp; l += 32) e[l] = 0.f;
  }
  __syncthreads();

  // ---- head-mean weights (attention.py:211)
  if (p.alpha) {
    for (int i = tid; i < k * L; i += kThreads) {
      const int b = i / L, l = i - b * L;
      float s = 0.f;
      for (int hd = 0; hd < heads; ++hd) s += s_p[(b * heads + hd) * Lp + l];
      p.alpha[(row0 + b) * p.ld_alpha + l] = s / (float)heads;
    }
  }

  // ---- attended values, heads concatenated
  const float* V = p.vproj + (int64_t)img * L * p.ld_kv;
  int G = 1;
  while (H4 * G * 2 <= kThreads) G *= 2;

  auto accumulate = [&](int c, int lstart, int lstride, float4 (&acc)[KB]) {
    const int hd = (c * 4) / d;
#pragma unroll
    for (int b = 0; b < KB; ++b) acc[b] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int l = lstart; l < L; l += lstride) {
      float4 x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int ll = min(l + u, L - 1);
        x[u] = ldg_stream(reinterpret_cast<const float4*>(V + (int64_t)ll * p.ld_kv) + c);
      }
#pragma unroll
      for (int b = 0; b < KB; ++b) {
        const float4 al = *reinterpret_cast<const float4*>(s_p + (b * heads + hd) * Lp + l);
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

  if (G == 1) {
    for (int c = tid; c < H4; c += kThreads) {
      float4 acc[KB];
      accumulate(c, 0, 4, acc);
#pragma unroll
      for (int b = 0; b < KB; ++b)
        if (b < k) *reinterpret_cast<float4*>(p.out + (row0 + b) * p.ld_out + c * 4) = acc[b];
    }
  } else {
    const int g = tid / H4, c = tid - g * H4;
    const bool active = g < G;
    float4 acc[KB];
    if (active) {
      accumulate(c, g * 4, G * 4, acc);
      if (g > 0) {
#pragma unroll
        for (int b = 0; b < KB; ++b) reinterpret_cast<float4*>(s_red)[((g - 1) * KB + b) * H4 + c] = acc[b];
      }
    }
    __syncthreads();
    if (active && g == 0) {
#pragma unroll
      for (int b = 0; b < KB; ++b) {
        if (b >= k) continue;
        float4 v = acc[b];
        for (int gg = 1; gg < G; ++gg) {
          const float4 t = reinterpret_cast<const float4*>(s_red)[((gg - 1) * KB + b) * H4 + c];
          v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
        }
        *reinterpret_cast<float4*>(p.out + (row0 + b) * p.ld_out + c * 4) = v;
      }
    }
  }
}

template <int KB>
int launch_kb(const MhaArgs& a, cudaStream_t s) {
  const int Lp = (a.L + 3) & ~3;
  const int H4 = a.H / 4;
  int G = 1;
  while (H4 * G * 2 <= kThreads) G *= 2;
  size_t smem = sizeof(float) * ((size_t)KB * a.H + (size_t)KB * a.heads * Lp + (size_t)(G - 1) * KB * a.H);
  CAPDEC_REQUIRE(smem <= 200 * 1024, CAPDEC_ERR_UNSUPPORTED, "mha_attention: shared memory %zu B too large", smem);
  auto kern = mha_attention_kernel<KB>;
  if (smem > 48 * 1024) CAPDEC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<a.B, kThreads, smem, s>>>(a);
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

}  // namespace

int mha_attention(const MhaArgs& a, cudaStream_t s) {
  CAPDEC_REQUIRE(a.k >= 1 && a.k <= kMaxRowsPerImage, CAPDEC_ERR_UNSUPPORTED,
                 "mha_attention: rows per image %d not in [1,%d]", a.k, kMaxRowsPerImage);
  CAPDEC_REQUIRE(a.heads >= 1 && a.H % a.heads == 0 && (a.H / a.heads) % 4 == 0, CAPDEC_ERR_UNSUPPORTED,
                 "mha_attention: head_dim must be a multiple of 4 (H=%d heads=%d)", a.H, a.heads);
  CAPDEC_REQUIRE(a.ld_q % 4 == 0 && a.ld_out % 4 == 0 && a.ld_kv % 4 == 0 && a.ld_kv >= a.H, CAPDEC_ERR_INVALID, "mha_attention: strides must be multiples of 4");
  if (a.B == 0) return CAPDEC_OK;
  const int took = mha_attention_stream(a, s);   // persistent TMA-streamed kernel for dense K/V tiles
  if (took != 0) return took > 0 ? CAPDEC_OK : took;
  switch (a.k) {
    case 1: return launch_kb<1>(a, s);
    case 2: return launch_kb<2>(a, s);
    case 3: return launch_kb<3>(a, s);
    case 4: return launch_kb<4>(a, s);
    case 5: return launch_kb<5>(a, s);
    case 6: return launch_kb<6>(a, s);
    default: return launch_kb<8>(a, s);
  }
}

}  // namespace capdec
