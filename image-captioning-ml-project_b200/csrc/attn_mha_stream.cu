// Streaming multi-head dot-product attention for one decode step: the persistent, warp-specialised form of
// attn_mha.cu (src/models/attention.py:161-211; also the cross-attention of nn.TransformerDecoderLayer,
// src/models/decoders.py:463-480 with K/V hoisted out of the step loop).
//
//   s[b,h,l] = q[row_b,h,:] . K[img,l,h,:] / denom      (masked -> -1e9)
//   p        = softmax_l(s);  out[row_b,h,:] = sum_l p[b,h,l] * V[img,l,h,:];  alpha[row_b,l] = mean_h p[b,h,l]
//
// HBM-bound: per image-step it must read the image's projected K [L,H] and V [L,H] exactly once (2*L*H*4 bytes).
// One CTA per SM walks images blockIdx.x, blockIdx.x + gridDim.x, ...:
//   warps 0, 1    producers: one thread each streams V / K through its own shared-memory ring with 1-D TMA bulk copies
//                 (dense [L,H] tiles: a ring stage is a run of whole rows)
//   warps 2-9     score warps: K stages of 8 rows are dealt round-robin to the warps; inside a warp lane = (row, quarter of
//                 the heads), so every lane computes COMPLETE dot products (no cross-lane reduction at all); the K rows sit
//                 at a padded pitch (H+4 floats, one bulk copy per row) so the 8 rows of a quarter-warp hit distinct banks;
//                 softmax per (row, head) into a double-buffered probability array; optional head-mean weights
//   warps 10-17   context warps: V stages x probabilities -> k output rows in registers -> global
// The score work of image i+1 overlaps the context work of image i; the K producer runs an image ahead.
#include <stdlib.h>

#include "attention.cuh"

namespace capdec {
namespace {

constexpr int kScoreWarps = 8, kCtxWarps = 8;
constexpr int kScoreThreads = 32 * kScoreWarps, kCtxThreads = 32 * kCtxWarps;
constexpr int kThreads = 64 + kScoreThreads + kCtxThreads;   // 576
constexpr int kStagesK = 3, kStagesV = 4;
constexpr int kRowsK = 8;                                    // K rows per stage = lanes per head-quarter
constexpr int kPBuf = 2;
constexpr uint32_t kSpinLimit = 1u << 24;

struct MhaLayout {
  int rowsK, rowsV, nK, nV;
  uint32_t stageK, stageV;
  uint32_t off_ringK, off_ringV, off_q, off_p, off_red, off_bar, total;
  int Lp, G;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (spin > kSpinLimit) __trap();   // a protocol bug traps instead of hanging the GPU
  }
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_barrier(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <int KB, int NC>
__global__ void __launch_bounds__(kThreads, 1) mha_attention_stream_kernel(const MhaArgs p, const MhaLayout y) {
  pdl_trigger();
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* ringK = smem + y.off_ringK;
  uint8_t* ringV = smem + y.off_ringV;
  float* s_q = reinterpret_cast<float*>(smem + y.off_q);       // [2][KB][H]
  float* s_p = reinterpret_cast<float*>(smem + y.off_p);       // [kPBuf][KB][heads][Lp]
  float* s_red = reinterpret_cast<float*>(smem + y.off_red);   // [G-1][KB][H]
  uint64_t* fullK = reinterpret_cast<uint64_t*>(smem + y.off_bar);
  uint64_t* emptyK = fullK + kStagesK;
  uint64_t* fullV = emptyK + kStagesK;
  uint64_t* emptyV = fullV + kStagesV;
  uint64_t* p_full = emptyV + kStagesV;
  uint64_t* p_empty = p_full + kPBuf;

  const int L = p.L, H = p.H, heads = p.heads, k = p.k, Lp = y.Lp;
  const int d = H / heads, d4 = d >> 2, H4 = H >> 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_img = ((int)blockIdx.x < p.B) ? (p.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const size_t row_bytes = (size_t)H * sizeof(float);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStagesK; ++s) { mbar_init(&fullK[s], 1); mbar_init(&emptyK[s], kScoreWarps); }
    for (int s = 0; s < kStagesV; ++s) { mbar_init(&fullV[s], 1); mbar_init(&emptyV[s], kCtxWarps); }
    for (int s = 0; s < kPBuf; ++s) { mbar_init(&p_full[s], kScoreWarps); mbar_init(&p_empty[s], kCtxWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_wait();

  if (warp == 0) {
    // ================================ V producer ================================
    if (lane == 0) {
      uint32_t it = 0;
      for (int i = 0; i < n_img; ++i) {
        const int img = blockIdx.x + i * gridDim.x;
        for (int c = 0; c < y.nV; ++c, ++it) {
          const int s = it % kStagesV;
          mbar_wait(&emptyV[s], ((it / kStagesV) & 1) ^ 1);
          const int rows = min(y.rowsV, L - c * y.rowsV);
          const uint32_t bytes = (uint32_t)(rows * row_bytes);
          mbar_expect_tx(&fullV[s], bytes);
          bulk_load(ringV + (size_t)s * y.stageV, p.vproj + ((size_t)img * L + (size_t)c * y.rowsV) * H, bytes, &fullV[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ================================ K producer ================================
    if (lane == 0) {
      uint32_t it = 0;
      const size_t pitch = row_bytes + 16;   // H + 4 floats: the 8 rows of a stage start 4 banks apart
      for (int i = 0; i < n_img; ++i) {
        const int img = blockIdx.x + i * gridDim.x;
        for (int c = 0; c < y.nK; ++c, ++it) {
          const int s = it % kStagesK;
          mbar_wait(&emptyK[s], ((it / kStagesK) & 1) ^ 1);
          const int rows = min(kRowsK, L - c * kRowsK);
          mbar_expect_tx(&fullK[s], (uint32_t)(rows * row_bytes));
          const float* src = p.kproj + ((size_t)img * L + (size_t)c * kRowsK) * H;
          for (int r = 0; r < rows; ++r)
            bulk_load(ringK + (size_t)s * y.stageK + r * pitch, src + (size_t)r * H, (uint32_t)row_bytes, &fullK[s]);
        }
      }
    }
  } else if (warp < 2 + kScoreWarps) {
    // ================================ score warps ================================
    const int sw = warp - 2, t = threadIdx.x - 64;
    uint32_t it = 0;
    for (int i = 0; i < n_img; ++i) {
      const int img = blockIdx.x + i * gridDim.x;
      const int64_t row0 = (int64_t)img * k;
      const int buf = i % kPBuf;
      float* q2 = s_q + (size_t)(i & 1) * KB * H;
      for (int j = t; j < KB * H4; j += kScoreThreads) {
        const int b = j / H4, c = j - b * H4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b < k) v = *reinterpret_cast<const float4*>(p.q + (row0 + b) * p.ld_q + c * 4);
        reinterpret_cast<float4*>(q2)[j] = v;
      }
      named_barrier(1, kScoreThreads);
      mbar_wait(&p_empty[buf], (((uint32_t)i / kPBuf) & 1) ^ 1);   // the context warps are done with this probability buffer
      float* pb = s_p + (size_t)buf * KB * heads * Lp;
      const int r = lane & 7, part = lane >> 3;            // lane = (row of the stage, quarter of the heads)
      const int hpp = (heads + 3) >> 2;                    // heads per quarter
      const int Hp4 = H4 + 1;                              // padded row pitch in float4
      for (int c = 0; c < y.nK; ++c, ++it) {
        // Stages are dealt round-robin to the score warps, but EVERY warp observes every phase of every stage barrier
        // (a parity wait is only meaningful for a waiter that is at most one phase behind) and arrives on its release.
        const int s = it % kStagesK;
        mbar_wait(&fullK[s], (it / kStagesK) & 1);
        const int rows = min(kRowsK, L - c * kRowsK);
        const float4* tile = reinterpret_cast<const float4*>(ringK + (size_t)s * y.stageK);
        if (c % kScoreWarps == sw && r < rows) {
          const int l = c * kRowsK + r;
          const bool masked = p.mask && p.mask[(int64_t)img * L + l];
          for (int hh = 0; hh < hpp; ++hh) {
            const int hd = part * hpp + hh;
            if (hd >= heads) break;
            float acc[KB];
#pragma unroll
            for (int b = 0; b < KB; ++b) acc[b] = 0.f;
            const float4* xr = tile + (size_t)r * Hp4 + hd * d4;
            const float4* qh = reinterpret_cast<const float4*>(q2) + hd * d4;
            // four K float4s and their 4 x KB query float4s are requested before the first FMA (the loop was bound by the
            // shared-memory load latency of one element at a time: 57 % of the stall cycles were short-scoreboard waits)
            int e = 0;
            for (; e + 4 <= d4; e += 4) {
              float4 x[4], q[4][KB];
#pragma unroll
              for (int u2 = 0; u2 < 4; ++u2) {
                x[u2] = xr[e + u2];
#pragma unroll
                for (int b = 0; b < KB; ++b) q[u2][b] = qh[b * H4 + e + u2];
              }
#pragma unroll
              for (int u2 = 0; u2 < 4; ++u2)
#pragma unroll
                for (int b = 0; b < KB; ++b) {
                  float u = acc[b];
                  u = fmaf(q[u2][b].x, x[u2].x, u); u = fmaf(q[u2][b].y, x[u2].y, u);
                  u = fmaf(q[u2][b].z, x[u2].z, u); u = fmaf(q[u2][b].w, x[u2].w, u);
                  acc[b] = u;
                }
            }
            for (; e < d4; ++e) {
              const float4 x = xr[e];
#pragma unroll
              for (int b = 0; b < KB; ++b) {
                const float4 q = qh[b * H4 + e];
                float u = acc[b];
                u = fmaf(q.x, x.x, u); u = fmaf(q.y, x.y, u); u = fmaf(q.z, x.z, u); u = fmaf(q.w, x.w, u);
                acc[b] = u;
              }
            }
#pragma unroll
            for (int b = 0; b < KB; ++b) pb[(b * heads + hd) * Lp + l] = masked ? -1.0e9f : acc[b] / p.denom;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&emptyK[s]);
      }
      named_barrier(1, kScoreThreads);   // every score of the image is in shared memory
      for (int j = sw; j < k * heads; j += kScoreWarps) {
        float* e = pb + j * Lp;
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
      }
      if (p.alpha) {   // head-mean weights (attention.py:211)
        named_barrier(1, kScoreThreads);
        for (int j = t; j < k * L; j += kScoreThreads) {
          const int b = j / L, l = j - b * L;
          float s = 0.f;
          for (int hd = 0; hd < heads; ++hd) s += pb[(b * heads + hd) * Lp + l];
          p.alpha[(row0 + b) * p.ld_alpha + l] = s / (float)heads;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[buf]);   // release: the probabilities of this image are ready
    }
  } else {
    // ================================ context warps ================================
    const int t = threadIdx.x - 64 - kScoreThreads;   // 0..255
    const int Hw = H4 < kCtxThreads ? H4 : kCtxThreads;
    const int g = t / Hw, c0 = t - g * Hw;
    const bool active = g < y.G;
    int hd_of[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) hd_of[j] = min(((c0 + j * kCtxThreads) * 4) / d, heads - 1);
    uint32_t it = 0;
    for (int i = 0; i < n_img; ++i) {
      const int img = blockIdx.x + i * gridDim.x;
      const int64_t row0 = (int64_t)img * k;
      const int buf = i % kPBuf;
      mbar_wait(&p_full[buf], ((uint32_t)i / kPBuf) & 1);
      const float* pb = s_p + (size_t)buf * KB * heads * Lp;
      float4 acc[NC][KB];
#pragma unroll
      for (int j = 0; j < NC; ++j)
#pragma unroll
        for (int b = 0; b < KB; ++b) acc[j][b] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c = 0; c < y.nV; ++c, ++it) {
        const int s = it % kStagesV;
        mbar_wait(&fullV[s], (it / kStagesV) & 1);
        const int rows = min(y.rowsV, L - c * y.rowsV);
        const float4* tile = reinterpret_cast<const float4*>(ringV + (size_t)s * y.stageV);
        if (active) {
          // two rows per iteration: their V float4s and 2 x KB probabilities are requested together
          int r = g;
          for (; r + y.G < rows; r += 2 * y.G) {
            const int l0 = c * y.rowsV + r, l1 = l0 + y.G;
#pragma unroll
            for (int j = 0; j < NC; ++j) {
              const int col = c0 + j * kCtxThreads;
              if (col < H4) {
                const float4 x0 = tile[(size_t)r * H4 + col], x1 = tile[(size_t)(r + y.G) * H4 + col];
                float a0[KB], a1[KB];
#pragma unroll
                for (int b = 0; b < KB; ++b) { a0[b] = pb[(b * heads + hd_of[j]) * Lp + l0]; a1[b] = pb[(b * heads + hd_of[j]) * Lp + l1]; }
#pragma unroll
                for (int b = 0; b < KB; ++b) {
                  acc[j][b].x = fmaf(a0[b], x0.x, acc[j][b].x); acc[j][b].y = fmaf(a0[b], x0.y, acc[j][b].y);
                  acc[j][b].z = fmaf(a0[b], x0.z, acc[j][b].z); acc[j][b].w = fmaf(a0[b], x0.w, acc[j][b].w);
                  acc[j][b].x = fmaf(a1[b], x1.x, acc[j][b].x); acc[j][b].y = fmaf(a1[b], x1.y, acc[j][b].y);
                  acc[j][b].z = fmaf(a1[b], x1.z, acc[j][b].z); acc[j][b].w = fmaf(a1[b], x1.w, acc[j][b].w);
                }
              }
            }
          }
          for (; r < rows; r += y.G) {
            const int l = c * y.rowsV + r;
#pragma unroll
            for (int j = 0; j < NC; ++j) {
              const int col = c0 + j * kCtxThreads;
              if (col < H4) {
                const float4 x = tile[(size_t)r * H4 + col];
#pragma unroll
                for (int b = 0; b < KB; ++b) {
                  const float al = pb[(b * heads + hd_of[j]) * Lp + l];
                  acc[j][b].x = fmaf(al, x.x, acc[j][b].x); acc[j][b].y = fmaf(al, x.y, acc[j][b].y);
                  acc[j][b].z = fmaf(al, x.z, acc[j][b].z); acc[j][b].w = fmaf(al, x.w, acc[j][b].w);
                }
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&emptyV[s]);
      }
      if (y.G > 1) {   // (NC == 1 here) fold the row groups through shared memory
        if (active && g > 0) {
#pragma unroll
          for (int b = 0; b < KB; ++b) reinterpret_cast<float4*>(s_red)[((size_t)(g - 1) * KB + b) * H4 + c0] = acc[0][b];
        }
        named_barrier(2, kCtxThreads);
        if (active && g == 0) {
#pragma unroll
          for (int b = 0; b < KB; ++b)
            for (int gg = 1; gg < y.G; ++gg) {
              const float4 v = reinterpret_cast<const float4*>(s_red)[((size_t)(gg - 1) * KB + b) * H4 + c0];
              acc[0][b].x += v.x; acc[0][b].y += v.y; acc[0][b].z += v.z; acc[0][b].w += v.w;
            }
        }
        named_barrier(2, kCtxThreads);
      }
      if (active && g == 0) {
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          const int col = c0 + j * kCtxThreads;
          if (col >= H4) continue;
#pragma unroll
          for (int b = 0; b < KB; ++b)
            if (b < k) *reinterpret_cast<float4*>(p.out + (row0 + b) * p.ld_out + col * 4) = acc[j][b];
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_empty[buf]);
    }
  }
}

int sm_count() { return num_sms(); }

bool plan(const MhaArgs& a, int KB, MhaLayout* y) {
  const int H4 = a.H / 4;
  if (a.H % 4 || H4 > 2 * kCtxThreads || a.L < 1 || a.ld_kv != a.H) return false;   // dense [L,H] tiles only
  if ((((uintptr_t)a.kproj | (uintptr_t)a.vproj | (uintptr_t)a.q | (uintptr_t)a.out) & 15) != 0) return false;
  if (a.ld_q % 4 || a.ld_out % 4) return false;
  int G = 1;
  if (H4 <= kCtxThreads / 2) { while (H4 * G * 2 <= kCtxThreads) G *= 2; }
  const size_t row = (size_t)a.H * 4;
  const int Lp = (a.L + 3) & ~3;
  size_t fixed = 0;
  auto take = [&](size_t bytes) { const size_t o = fixed; fixed = (fixed + bytes + 127) & ~(size_t)127; return (uint32_t)o; };
  y->off_q = take((size_t)2 * KB * a.H * 4);
  y->off_p = take((size_t)kPBuf * KB * a.heads * Lp * 4);
  y->off_red = take(G > 1 ? (size_t)(G - 1) * KB * a.H * 4 : 16);
  y->off_bar = take((size_t)(2 * kStagesK + 2 * kStagesV + 2 * kPBuf) * 8);
  const size_t budget = 220 * 1024;
  if (fixed + (kStagesK + kStagesV) * (row + 128) > budget) return false;
  const size_t left = budget - fixed;
  const size_t stageK = ((size_t)kRowsK * (row + 16) + 127) & ~(size_t)127;
  if ((size_t)kStagesK * stageK + kStagesV * (row + 128) > left) return false;
  const int rowsK = kRowsK;
  int rowsV = (int)(((left - kStagesK * stageK) / kStagesV - 128) / row);
  if (rowsV < 1) return false;
  rowsV = rowsV > a.L ? a.L : rowsV;
  if (rowsV > G) rowsV = rowsV / G * G;
  y->rowsK = rowsK; y->rowsV = rowsV;
  y->nK = (a.L + rowsK - 1) / rowsK; y->nV = (a.L + rowsV - 1) / rowsV;
  y->stageK = (uint32_t)stageK;
  y->stageV = (uint32_t)(((size_t)rowsV * row + 127) & ~(size_t)127);
  y->off_ringK = take((size_t)kStagesK * y->stageK);
  y->off_ringV = take((size_t)kStagesV * y->stageV);
  y->total = (uint32_t)fixed;
  y->Lp = Lp; y->G = G;
  return fixed <= 227 * 1024;
}

template <int KB>
int launch_stream(const MhaArgs& a, const MhaLayout& y, cudaStream_t s) {
  const int nc = (a.H / 4 + kCtxThreads - 1) / kCtxThreads;
  const int grid = a.B < sm_count() ? a.B : sm_count();
  if (nc == 1) {
    auto kern = mha_attention_stream_kernel<KB, 1>;
    CAPDEC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)y.total));
    CAPDEC_CHECK_CUDA(launch_k(kern, dim3(grid), dim3(kThreads), y.total, s, true, a, y));
  } else {
    auto kern = mha_attention_stream_kernel<KB, 2>;
    CAPDEC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)y.total));
    CAPDEC_CHECK_CUDA(launch_k(kern, dim3(grid), dim3(kThreads), y.total, s, true, a, y));
  }
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

}  // namespace

// returns 1 when the streaming kernel took the call, 0 when the shape is left to the generic kernel, < 0 on error
int mha_attention_stream(const MhaArgs& a, cudaStream_t s) {
  static const bool disabled = ab_switch("CAPDEC_ATTN_GENERIC");
  if (disabled || a.k < 1 || a.k > kMaxRowsPerImage || a.heads < 1 || a.H % a.heads || (a.H / a.heads) % 4) return 0;
  const int KB = a.k <= 6 ? a.k : 8;
  MhaLayout y{};
  if (!plan(a, KB, &y)) return 0;
  int st;
  switch (KB) {
    case 1: st = launch_stream<1>(a, y, s); break;
    case 2: st = launch_stream<2>(a, y, s); break;
    case 3: st = launch_stream<3>(a, y, s); break;
    case 4: st = launch_stream<4>(a, y, s); break;
    case 5: st = launch_stream<5>(a, y, s); break;
    case 6: st = launch_stream<6>(a, y, s); break;
    default: st = launch_stream<8>(a, y, s); break;
  }
  return st == CAPDEC_OK ? 1 : st;
}

}  // namespace capdec
