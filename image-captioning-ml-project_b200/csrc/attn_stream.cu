// Streaming additive attention: the HBM-bound stage of the decode step as a persistent, warp-specialised kernel.
//
//   e[b,l]   = (w . act(att1[img,l,:] + att2[row_b,:]) + w_bias) / temperature      (masked -> -1e9)
//   alpha    = softmax_l(e)
//   ctx[b,:] = sum_l alpha[b,l] * feats[img,l,:]   (* gate[row_b,:])
//   legacy models/decoder.py:152-161 (relu, gate) / src/models/attention.py:76-111 (tanh)
//
// Per image-step the kernel must read att1 [L,A] and feats [L,D] exactly once (2.0 MB as fp32 tiles for 196 x (512+2048);
// 1.5 MB as p24 planes in the bf16x3 mode, 1.0 MB as bf16 tiles in the bf16 mode -- template parameter BF = 0 / 2 / 1);
// everything else is noise.  One CTA per SM walks images blockIdx.x, blockIdx.x + gridDim.x, ...:
//   warps 0, 1    producers: one thread each streams feats / att1 through its own shared-memory ring with 1-D TMA bulk
//                 copies (cp.async.bulk, mbarrier complete_tx).  The rows of an image are contiguous in HBM (per plane), so
//                 a ring stage is simply a run of whole rows (p24: the 16-bit rows, then the byte rows).  The two streams
//                 are independent: the att1 producer runs up to two images ahead, so HBM requests never pause at an
//                 image boundary.
//   warps 2-9     score warps: att1 chunks -> scores (lanes split the attention dim, the k beams of the image reuse every
//                 decoded element; a transposing butterfly reduces the 2k partial sums of a row pair) -> softmax (bias,
//                 temperature and mask are applied here) -> alpha in a triple-buffered shared array laid out [l][beam].
//   warps 10-17   context warps: feats chunks x alpha -> k context rows in registers (packed FFMA2, two rows per
//                 iteration, adjacent columns per thread, constant strides) -> gate multiply -> written straight into
//                 the LSTM operand.
// The ALU-heavy score work of image i+1 therefore overlaps the load-heavy context work of image i inside one CTA, and
// up to ~170 KB of loads are in flight per SM without costing registers.  No loop contains a run-time division or
// modulo (ring positions are counters, the row-group count is a power of two).
#include <stdlib.h>

#include "attention.cuh"

namespace capdec {
namespace {

constexpr int kScoreWarps = 8, kCtxWarps = 8;
constexpr int kScoreThreads = 32 * kScoreWarps, kCtxThreads = 32 * kCtxWarps;
constexpr int kThreads = 64 + kScoreThreads + kCtxThreads;   // 576
constexpr int kMaxStagesA = 4, kMaxStagesF = 8;                // att1 ring: StreamLayout::stA stages (2: 3 and 4 measured slower, they shrink the feats ring)
constexpr int kEBuf = 3;                                     // alpha buffers: scores of image i+1 while context reads image i
constexpr uint32_t kSpinLimit = 1u << 24;

struct StreamLayout {
  int rowsA, rowsF, nA, nF;            // rows per ring stage, stages per image
  uint32_t stageA, stageF;             // bytes per stage
  uint32_t b8A, b8F;                   // p24 tiles: offset of the byte plane inside a stage
  uint32_t off_ringA, off_ringF, off_att2, off_w, off_e, off_red, off_bar, total;
  int Lp, G, gshift;                   // padded L; context row groups (threads split rows when D/4 <= 128), log2(G)
  int stA, stF;                        // ring stages
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

constexpr float kTwoLog2e = 2.8853900817779268f;   // e^{2x} = 2^{kTwoLog2e x}
constexpr float kTanhRange = 21.f;                 // e^{+-42}: the product of two such factors stays a normal fp32

template <int ACT>
__device__ __forceinline__ float act_fn(float x) {
  return ACT == ACT_RELU ? fmaxf(x, 0.f) : ACT == ACT_TANH_FAST ? tanh_fast_(x) : tanhf(x);
}

// 4 consecutive elements (index i4 counts groups of 4) of a shared-memory tile row holding fp32 or bf16 data
template <int BF>
__device__ __forceinline__ float4 tile_ld4(const void* row, int i4, const void* row_b8 = nullptr) {
  if (BF == 2) return p24_decode4_(reinterpret_cast<const uint2*>(row)[i4], reinterpret_cast<const uint32_t*>(row_b8)[i4]);
  if (BF) {
    const uint2 u = reinterpret_cast<const uint2*>(row)[i4];
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u),
                       __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
  }
  return reinterpret_cast<const float4*>(row)[i4];
}

template <int KB, int ACT, int NC, int BF>
__global__ void __launch_bounds__(kThreads, 1) additive_attention_stream_kernel(const AddAttnArgs p, const StreamLayout y) {
  constexpr size_t ES = BF ? 2 : 4;   // bytes per element of the (first) tile plane; BF == 2 adds a byte plane behind it
  pdl_trigger();
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* ringA = smem + y.off_ringA;
  uint8_t* ringF = smem + y.off_ringF;
  float* s_att2 = reinterpret_cast<float*>(smem + y.off_att2);   // [2][KB][A]
  float* s_w = reinterpret_cast<float*>(smem + y.off_w);         // [A]
  float* s_e = reinterpret_cast<float*>(smem + y.off_e);         // [kEBuf][Lp][KBP]: the beams of a region are adjacent
  constexpr int KBP = KB <= 4 ? 4 : 8;
  float* s_red = reinterpret_cast<float*>(smem + y.off_red);     // [G-1][KB][D]
  uint64_t* fullA = reinterpret_cast<uint64_t*>(smem + y.off_bar);
  uint64_t* emptyA = fullA + kMaxStagesA;
  uint64_t* fullF = emptyA + kMaxStagesA;
  const int kStagesA = y.stA;
  uint64_t* emptyF = fullF + kMaxStagesF;
  uint64_t* e_full = emptyF + kMaxStagesF;
  const int kStagesF = y.stF;
  uint64_t* e_empty = e_full + kEBuf;
  int* s_flag = reinterpret_cast<int*>(e_empty + kEBuf);         // [2] "an att2 value of this image is out of range"
  float* s_wsum = reinterpret_cast<float*>(s_flag + 2);          // sum_a w[a]                (both ACT_TANH_FAST only)

  const int A = p.A, L = p.L, D = p.D, k = p.k, Lp = y.Lp;
  const int A4 = A >> 2, D4 = D >> 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_img = ((int)blockIdx.x < p.B) ? (p.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStagesA; ++s) { mbar_init(&fullA[s], 1); mbar_init(&emptyA[s], kScoreWarps); }
    for (int s = 0; s < kMaxStagesF; ++s) { mbar_init(&fullF[s], 1); mbar_init(&emptyF[s], kCtxWarps); }
    for (int s = 0; s < kEBuf; ++s) { mbar_init(&e_full[s], kScoreWarps); mbar_init(&e_empty[s], kCtxWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_flag[0] = s_flag[1] = 0;
  }
  __syncthreads();
  pdl_wait();

  if (warp == 0) {
    // ================================ feats producer ================================
    if (lane == 0) {
      int sF = 0; uint32_t phF = 0;
      const size_t rowF = (size_t)D * ES;
      for (int i = 0; i < n_img; ++i) {
        const int img = blockIdx.x + i * gridDim.x;
        for (int c = 0; c < y.nF; ++c, phF ^= (++sF == kStagesF), sF = sF == kStagesF ? 0 : sF) {
          const int s = sF;
          mbar_wait(&emptyF[s], phF ^ 1);
          const int rows = min(y.rowsF, L - c * y.rowsF);
          const uint32_t bytes = (uint32_t)(rows * rowF);
          mbar_expect_tx(&fullF[s], BF == 2 ? bytes + bytes / 2 : bytes);
          bulk_load(ringF + (size_t)s * y.stageF,
                    reinterpret_cast<const char*>(p.feats) + ((size_t)img * L + (size_t)c * y.rowsF) * rowF, bytes, &fullF[s]);
          if (BF == 2)
            bulk_load(ringF + (size_t)s * y.stageF + y.b8F, p.feats_b8 + ((size_t)img * L + (size_t)c * y.rowsF) * D, bytes / 2,
                      &fullF[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ================================ att1 producer ================================
    if (lane == 0) {
      int sA = 0; uint32_t phA = 0;   // ring stage and its phase bit, kept as counters (the stage count is a run-time value)
      const size_t rowA = (size_t)A * ES;
      for (int i = 0; i < n_img; ++i) {
        const int img = blockIdx.x + i * gridDim.x;
        for (int c = 0; c < y.nA; ++c, phA ^= (++sA == kStagesA), sA = sA == kStagesA ? 0 : sA) {
          const int s = sA;
          mbar_wait(&emptyA[s], phA ^ 1);
          const int rows = min(y.rowsA, L - c * y.rowsA);
          const uint32_t bytes = (uint32_t)(rows * rowA);
          mbar_expect_tx(&fullA[s], BF == 2 ? bytes + bytes / 2 : bytes);
          bulk_load(ringA + (size_t)s * y.stageA,
                    reinterpret_cast<const char*>(p.att1) + ((size_t)img * L + (size_t)c * y.rowsA) * rowA, bytes, &fullA[s]);
          if (BF == 2)
            bulk_load(ringA + (size_t)s * y.stageA + y.b8A, p.att1_b8 + ((size_t)img * L + (size_t)c * y.rowsA) * A, bytes / 2,
                      &fullA[s]);
        }
      }
    }
  } else if (warp < 2 + kScoreWarps) {
    // ================================ score warps ================================
    const int sw = warp - 2, t = threadIdx.x - 64;
    for (int i = t; i < A4; i += kScoreThreads) reinterpret_cast<float4*>(s_w)[i] = reinterpret_cast<const float4*>(p.w)[i];
    if (ACT == ACT_TANH_FAST && sw == 0) {   // (p.w is a weight: not written by the preceding kernels)
      float ws = 0.f;
      for (int a = lane; a < A; a += 32) ws += p.w[a];
      ws = warp_sum(ws);
      if (lane == 0) *s_wsum = ws;
    }
    int sA = 0; uint32_t phA = 0;
    for (int i = 0; i < n_img; ++i) {
      const int img = blockIdx.x + i * gridDim.x;
      const int64_t row0 = (int64_t)img * k;
      const int buf = i % kEBuf;
      float* q2 = s_att2 + (size_t)(i & 1) * KB * A;
      for (int j = t; j < KB * A4; j += kScoreThreads) {
        const int b = j / A4, a4 = j - b * A4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b < k) {
          const int64_t rs = p.row_src ? p.row_src[row0 + b] : row0 + b;
          v = *reinterpret_cast<const float4*>(p.att2 + rs * p.ld_att2 + a4 * 4);
        }
        if (ACT == ACT_TANH_FAST) {
          // product form (see the score loop): the staged value is e^{2q}; out-of-range rows send the image down the
          // direct path instead
          if (fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))) > kTanhRange) s_flag[i & 1] = 1;
          v.x = ex2_approx_(kTwoLog2e * v.x); v.y = ex2_approx_(kTwoLog2e * v.y);
          v.z = ex2_approx_(kTwoLog2e * v.z); v.w = ex2_approx_(kTwoLog2e * v.w);
        }
        reinterpret_cast<float4*>(q2)[j] = v;
      }
      named_barrier(1, kScoreThreads);
      const bool qbig = ACT == ACT_TANH_FAST && s_flag[i & 1] != 0;
      if (ACT == ACT_TANH_FAST && t == 0) s_flag[(i + 1) & 1] = 0;   // its last readers (image i-1) are behind this barrier
      mbar_wait(&e_empty[buf], (((uint32_t)i / kEBuf) & 1) ^ 1);   // the context warps are done with this alpha buffer
      float* e = s_e + (size_t)buf * KBP * Lp;
      for (int c = 0; c < y.nA; ++c, phA ^= (++sA == kStagesA), sA = sA == kStagesA ? 0 : sA) {
        const int s = sA;
        mbar_wait(&fullA[s], phA);
        const int rows = min(y.rowsA, L - c * y.rowsA);
        const uint8_t* tile = ringA + (size_t)s * y.stageA;
        for (int r = sw * 2; r < rows; r += kScoreWarps * 2) {
          const int r1 = min(r + 1, rows - 1);
          const void* x0p = tile + (size_t)r * A * ES;
          const void* x1p = tile + (size_t)r1 * A * ES;
          const void* x0b = tile + y.b8A + (size_t)r * A;     // (BF == 2 only)
          const void* x1b = tile + y.b8A + (size_t)r1 * A;
          float acc0[KB], acc1[KB];
          float2 pacc0[KB], pacc1[KB];   // (packed relu path only)
#pragma unroll
          for (int b = 0; b < KB; ++b) { acc0[b] = 0.f; acc1[b] = 0.f; pacc0[b] = pacc1[b] = make_float2(0.f, 0.f); }
          for (int a4 = lane; a4 < A4; a4 += 32) {
            const float4 x0 = tile_ld4<BF>(x0p, a4, x0b), x1 = tile_ld4<BF>(x1p, a4, x1b);
            const float4 wv = reinterpret_cast<const float4*>(s_w)[a4];
            if constexpr (ACT == ACT_TANH_FAST) {
              // tanh(x + q) = 1 - 2 r,  r = 1 / (1 + e^{2x} e^{2q}):  e^{2x} once per tile element (shared by the k beams),
              // e^{2q} once per beam row (staged above), so each (element, beam) costs ONE MUFU op (the reciprocal)
              // instead of two.  acc accumulates sum_a w_a r_a; the score is sum(w) - 2 acc.  |x|, |q| <= 21 keeps the
              // product inside fp32 range; anything larger (never seen from a trained projection) takes the direct form.
              const float mx = fmaxf(fmaxf(fmaxf(fabsf(x0.x), fabsf(x0.y)), fmaxf(fabsf(x0.z), fabsf(x0.w))),
                                     fmaxf(fmaxf(fabsf(x1.x), fabsf(x1.y)), fmaxf(fabsf(x1.z), fabsf(x1.w))));
              if (!qbig && mx <= kTanhRange) {
                const float4 E0 = make_float4(ex2_approx_(kTwoLog2e * x0.x), ex2_approx_(kTwoLog2e * x0.y),
                                              ex2_approx_(kTwoLog2e * x0.z), ex2_approx_(kTwoLog2e * x0.w));
                const float4 E1 = make_float4(ex2_approx_(kTwoLog2e * x1.x), ex2_approx_(kTwoLog2e * x1.y),
                                              ex2_approx_(kTwoLog2e * x1.z), ex2_approx_(kTwoLog2e * x1.w));
#pragma unroll
                for (int b = 0; b < KB; ++b) {
                  const float4 q = reinterpret_cast<const float4*>(q2)[b * A4 + a4];
                  float u = acc0[b], v = acc1[b];
                  u = fmaf(wv.x, rcp_approx_(fmaf(E0.x, q.x, 1.f)), u); v = fmaf(wv.x, rcp_approx_(fmaf(E1.x, q.x, 1.f)), v);
                  u = fmaf(wv.y, rcp_approx_(fmaf(E0.y, q.y, 1.f)), u); v = fmaf(wv.y, rcp_approx_(fmaf(E1.y, q.y, 1.f)), v);
                  u = fmaf(wv.z, rcp_approx_(fmaf(E0.z, q.z, 1.f)), u); v = fmaf(wv.z, rcp_approx_(fmaf(E1.z, q.z, 1.f)), v);
                  u = fmaf(wv.w, rcp_approx_(fmaf(E0.w, q.w, 1.f)), u); v = fmaf(wv.w, rcp_approx_(fmaf(E1.w, q.w, 1.f)), v);
                  acc0[b] = u; acc1[b] = v;
                }
              } else {
                for (int b = 0; b < k; ++b) {
                  const int64_t rs = p.row_src ? p.row_src[row0 + b] : row0 + b;
                  const float4 q = *reinterpret_cast<const float4*>(p.att2 + rs * p.ld_att2 + a4 * 4);
                  const float xs0[4] = {x0.x, x0.y, x0.z, x0.w}, xs1[4] = {x1.x, x1.y, x1.z, x1.w};
                  const float qs[4] = {q.x, q.y, q.z, q.w}, ws[4] = {wv.x, wv.y, wv.z, wv.w};
                  float u = 0.f, v = 0.f;
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    u = fmaf(ws[j], rcp_approx_(1.f + ex2_approx_(kTwoLog2e * (xs0[j] + qs[j]))), u);
                    v = fmaf(ws[j], rcp_approx_(1.f + ex2_approx_(kTwoLog2e * (xs1[j] + qs[j]))), v);
                  }
#pragma unroll
                  for (int bb = 0; bb < KB; ++bb)
                    if (bb == b) { acc0[bb] += u; acc1[bb] += v; }
                }
              }
            } else if constexpr (ACT == ACT_RELU && BF != 0) {
              // tolerance-based modes: packed add / fma (FADD2, FFMA2), two interleaved partial sums per accumulator
#pragma unroll
              for (int b = 0; b < KB; ++b) {
                const float4 q = reinterpret_cast<const float4*>(q2)[b * A4 + a4];
                float2 s0 = __fadd2_rn(make_float2(x0.x, x0.y), make_float2(q.x, q.y));
                float2 s1 = __fadd2_rn(make_float2(x0.z, x0.w), make_float2(q.z, q.w));
                float2 t0 = __fadd2_rn(make_float2(x1.x, x1.y), make_float2(q.x, q.y));
                float2 t1 = __fadd2_rn(make_float2(x1.z, x1.w), make_float2(q.z, q.w));
                s0.x = fmaxf(s0.x, 0.f); s0.y = fmaxf(s0.y, 0.f); s1.x = fmaxf(s1.x, 0.f); s1.y = fmaxf(s1.y, 0.f);
                t0.x = fmaxf(t0.x, 0.f); t0.y = fmaxf(t0.y, 0.f); t1.x = fmaxf(t1.x, 0.f); t1.y = fmaxf(t1.y, 0.f);
                pacc0[b] = __ffma2_rn(make_float2(wv.x, wv.y), s0, pacc0[b]);
                pacc0[b] = __ffma2_rn(make_float2(wv.z, wv.w), s1, pacc0[b]);
                pacc1[b] = __ffma2_rn(make_float2(wv.x, wv.y), t0, pacc1[b]);
                pacc1[b] = __ffma2_rn(make_float2(wv.z, wv.w), t1, pacc1[b]);
              }
            } else {
#pragma unroll
              for (int b = 0; b < KB; ++b) {
                const float4 q = reinterpret_cast<const float4*>(q2)[b * A4 + a4];
                float u = acc0[b], v = acc1[b];
                u = fmaf(wv.x, act_fn<ACT>(x0.x + q.x), u); v = fmaf(wv.x, act_fn<ACT>(x1.x + q.x), v);
                u = fmaf(wv.y, act_fn<ACT>(x0.y + q.y), u); v = fmaf(wv.y, act_fn<ACT>(x1.y + q.y), v);
                u = fmaf(wv.z, act_fn<ACT>(x0.z + q.z), u); v = fmaf(wv.z, act_fn<ACT>(x1.z + q.z), v);
                u = fmaf(wv.w, act_fn<ACT>(x0.w + q.w), u); v = fmaf(wv.w, act_fn<ACT>(x1.w + q.w), v);
                acc0[b] = u; acc1[b] = v;
              }
            }
          }
          // 2 KB partial sums per lane -> one total per (row, beam): a transposing butterfly (each exchange halves the
          // values a lane still carries) needs N - 1 + log2(32 / N) shuffles instead of 5 per value, and leaves the
          // totals in different lanes, which store them in parallel.  Bias, temperature and mask are applied by the
          // softmax pass below, spread over all lanes.
          constexpr int N = KB <= 1 ? 2 : KB <= 2 ? 4 : KB <= 4 ? 8 : KB <= 8 ? 16 : 32;
          static_assert(2 * KB <= 32, "beam bucket too wide for the score reduction");
          if constexpr (ACT == ACT_RELU && BF != 0) {
#pragma unroll
            for (int b = 0; b < KB; ++b) { acc0[b] = pacc0[b].x + pacc0[b].y; acc1[b] = pacc1[b].x + pacc1[b].y; }
          }
          float red[N];
#pragma unroll
          for (int j = 0; j < N / 2; ++j) { red[j] = j < KB ? acc0[j] : 0.f; red[N / 2 + j] = j < KB ? acc1[j] : 0.f; }
          int n = N;
#pragma unroll
          for (int m = 16; m >= 1; m >>= 1) {
            if (n > 1) {
              const bool up = (lane & m) != 0;
#pragma unroll
              for (int j = 0; j < N / 2; ++j) {
                if (j < n / 2) {
                  const float send = up ? red[j] : red[j + n / 2];
                  const float keep = up ? red[j + n / 2] : red[j];
                  red[j] = keep + __shfl_xor_sync(0xffffffffu, send, m);
                }
              }
              n >>= 1;
            } else {
              red[0] += __shfl_xor_sync(0xffffffffu, red[0], m);
            }
          }
          constexpr int kIdxShift = N == 32 ? 0 : N == 16 ? 1 : N == 8 ? 2 : N == 4 ? 3 : 4;
          const int idx = lane >> kIdxShift, rsel = idx / (N / 2), b = idx - rsel * (N / 2);
          if ((lane & ((1 << kIdxShift) - 1)) == 0 && b < KB && r + rsel < rows) {
            float v = red[0];
            if (ACT == ACT_TANH_FAST) v = fmaf(-2.f, v, *s_wsum);
            e[(c * y.rowsA + r + rsel) * KBP + b] = v;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&emptyA[s]);
      }
      named_barrier(1, kScoreThreads);   // every score of the image is in shared memory
      for (int b = sw; b < k; b += kScoreWarps) {
        float* eb = e + b;   // element l at eb[l * KBP]
        float m = -INFINITY;
        for (int l = lane; l < L; l += 32) {
          float v = (eb[l * KBP] + p.w_bias) / p.temperature;
          if (p.mask && p.mask[(int64_t)img * L + l]) v = -1.0e9f;
          eb[l * KBP] = v;
          m = fmaxf(m, v);
        }
        m = warp_max(m);
        float sum = 0.f;
        for (int l = lane; l < L; l += 32) {
          const float v = expf(eb[l * KBP] - m);
          eb[l * KBP] = v;
          sum += v;
        }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
        float* aout = p.alpha ? p.alpha + (row0 + b) * p.ld_alpha : nullptr;
        for (int l = lane; l < L; l += 32) {
          const float v = eb[l * KBP] * inv;
          eb[l * KBP] = v;
          if (aout) aout[l] = v;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&e_full[buf]);   // release: alpha of this image is ready for the context warps
    }
  } else {
    // ================================ context warps ================================
    const int t = threadIdx.x - 64 - kScoreThreads;   // 0..255
    // NC == 1: thread = (row group g, float4 column c0), G row groups share a chunk's rows.  NC == 2 (D/4 > 256): every
    // thread owns the two ADJACENT float4 columns 2t, 2t+1 (one 16/32-byte shared-memory load per plane and row), G == 1.
    const int Dw = D4 < kCtxThreads ? D4 : kCtxThreads;
    const int g = NC == 2 ? 0 : t / Dw, c0 = NC == 2 ? 2 * t : t - g * Dw;
    const bool active = NC == 2 ? c0 < D4 : g < y.G;
    const uint32_t offH = (uint32_t)c0 * (BF ? 8u : 16u), offB = (uint32_t)c0 * 4u;   // byte offsets inside a tile row
    const uint32_t rowH = (uint32_t)D * (uint32_t)ES;
    const uint32_t stepH = (uint32_t)y.G * rowH, stepB = (uint32_t)y.G * (uint32_t)D, stepE = (uint32_t)y.G * KBP;
    int sF = 0; uint32_t phF = 0;
    for (int i = 0; i < n_img; ++i) {
      const int img = blockIdx.x + i * gridDim.x;
      const int64_t row0 = (int64_t)img * k;
      const int buf = i % kEBuf;
      mbar_wait(&e_full[buf], ((uint32_t)i / kEBuf) & 1);
      const float* e = s_e + (size_t)buf * KBP * Lp;
      // packed fp32x2 FMAs (FFMA2, sm_100): the same IEEE fma per element, half the issue slots
      float2 acc[NC][KB][2];
#pragma unroll
      for (int j = 0; j < NC; ++j)
#pragma unroll
        for (int b = 0; b < KB; ++b) acc[j][b][0] = acc[j][b][1] = make_float2(0.f, 0.f);
      for (int c = 0; c < y.nF; ++c, phF ^= (++sF == kStagesF), sF = sF == kStagesF ? 0 : sF) {
        const int s = sF;
        mbar_wait(&fullF[s], phF);
        const int rows = min(y.rowsF, L - c * y.rowsF);
        const uint8_t* tile = ringF + (size_t)s * y.stageF;
        if (active) {
          // rows g, g + G, ... of the chunk, two per iteration (the loads of both rows are in flight before the first
          // FMA needs them); all addresses advance by per-thread constants
          const uint8_t* ph = tile + (size_t)g * rowH + offH;
          const uint8_t* pb = tile + y.b8F + (size_t)g * D + offB;
          const float* ep = e + (size_t)(c * y.rowsF + g) * KBP;
          int n = (rows - g + y.G - 1) >> y.gshift;   // rows of this chunk that belong to row group g (G is a power of two)
          auto load = [&](float4 (&x)[NC], const uint8_t* h, const uint8_t* q8) {
            if (NC == 2) {
              if (BF == 2) {
                const uint4 hv = *reinterpret_cast<const uint4*>(h);
                const uint2 qv = *reinterpret_cast<const uint2*>(q8);
                x[0] = p24_decode4_(make_uint2(hv.x, hv.y), qv.x);
                x[NC - 1] = p24_decode4_(make_uint2(hv.z, hv.w), qv.y);
              } else if (BF == 1) {
                const uint4 hv = *reinterpret_cast<const uint4*>(h);
                x[0] = make_float4(__uint_as_float(hv.x << 16), __uint_as_float(hv.x & 0xffff0000u),
                                   __uint_as_float(hv.y << 16), __uint_as_float(hv.y & 0xffff0000u));
                x[NC - 1] = make_float4(__uint_as_float(hv.z << 16), __uint_as_float(hv.z & 0xffff0000u),
                                        __uint_as_float(hv.w << 16), __uint_as_float(hv.w & 0xffff0000u));
              } else {
                x[0] = reinterpret_cast<const float4*>(h)[0];
                x[NC - 1] = reinterpret_cast<const float4*>(h)[1];
              }
            } else {
              if (BF == 2) x[0] = p24_decode4_(*reinterpret_cast<const uint2*>(h), *reinterpret_cast<const uint32_t*>(q8));
              else if (BF == 1) x[0] = tile_ld4<1>(h, 0);
              else x[0] = *reinterpret_cast<const float4*>(h);
            }
          };
          auto fma_row = [&](const float4 (&x)[NC], const float* al) {
            float w[KBP];
            *reinterpret_cast<float4*>(w) = *reinterpret_cast<const float4*>(al);
            if (KBP == 8) *reinterpret_cast<float4*>(w + 4) = *reinterpret_cast<const float4*>(al + 4);
#pragma unroll
            for (int b = 0; b < KB; ++b) {
              const float2 a2 = make_float2(w[b], w[b]);
#pragma unroll
              for (int j = 0; j < NC; ++j) {
                acc[j][b][0] = __ffma2_rn(a2, make_float2(x[j].x, x[j].y), acc[j][b][0]);
                acc[j][b][1] = __ffma2_rn(a2, make_float2(x[j].z, x[j].w), acc[j][b][1]);
              }
            }
          };
          for (; n >= 2; n -= 2) {
            float4 x[NC], xb[NC];
            load(x, ph, pb);
            load(xb, ph + stepH, pb + stepB);
            fma_row(x, ep);
            fma_row(xb, ep + stepE);
            ph += 2 * stepH; pb += 2 * stepB; ep += 2 * stepE;
          }
          if (n == 1) {
            float4 x[NC];
            load(x, ph, pb);
            fma_row(x, ep);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&emptyF[s]);
      }
      if (y.G > 1) {   // (NC == 1 here) fold the row groups through shared memory
        if (active && g > 0) {
#pragma unroll
          for (int b = 0; b < KB; ++b)
            reinterpret_cast<float4*>(s_red)[((size_t)(g - 1) * KB + b) * D4 + c0] =
                make_float4(acc[0][b][0].x, acc[0][b][0].y, acc[0][b][1].x, acc[0][b][1].y);
        }
        named_barrier(2, kCtxThreads);
        if (active && g == 0) {
#pragma unroll
          for (int b = 0; b < KB; ++b)
            for (int gg = 1; gg < y.G; ++gg) {
              const float4 v = reinterpret_cast<const float4*>(s_red)[((size_t)(gg - 1) * KB + b) * D4 + c0];
              acc[0][b][0].x += v.x; acc[0][b][0].y += v.y; acc[0][b][1].x += v.z; acc[0][b][1].y += v.w;
            }
        }
        named_barrier(2, kCtxThreads);   // s_red may be overwritten by the next image only after everyone has read it
      }
      if (active && g == 0) {
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          const int col = c0 + j;
          if (col >= D4) continue;
#pragma unroll
          for (int b = 0; b < KB; ++b) {
            if (b >= k) continue;
            float4 v = make_float4(acc[j][b][0].x, acc[j][b][0].y, acc[j][b][1].x, acc[j][b][1].y);
            if (p.gate) {
              const int64_t rs = p.row_src ? p.row_src[row0 + b] : row0 + b;
              const float4 gt = *reinterpret_cast<const float4*>(p.gate + rs * p.ld_gate + col * 4);
              v.x *= gt.x; v.y *= gt.y; v.z *= gt.z; v.w *= gt.w;
            }
            if (p.ctx) *reinterpret_cast<float4*>(p.ctx + (row0 + b) * p.ld_ctx + col * 4) = v;
            split_store4(p.ctx_split, row0 + b, p.ctx_split_col + col * 4, v);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&e_empty[buf]);
    }
  }
}

int sm_count() { return num_sms(); }

// Shared-memory plan; returns false when the shape does not fit this kernel (the caller uses the generic kernel).
bool plan(const AddAttnArgs& a, int KB, StreamLayout* y) {
  const int A4 = a.A / 4, D4 = a.D / 4;
  const size_t es = a.tile_fmt == 2 ? 3 : a.tile_fmt ? 2 : 4;   // bytes per tile element over all planes
  if (a.tile_fmt == 2 && (a.A % 16 || a.D % 16)) return false;   // byte-plane rows must stay 16-byte multiples
  if (a.A % 8 || a.D % 8 || D4 > 2 * kCtxThreads || a.L < 1) return false;   // rows are multiples of 16 bytes in either tile type
  if ((((uintptr_t)a.att1 | (uintptr_t)a.feats | (uintptr_t)a.att2 | (uintptr_t)a.w) & 15) != 0) return false;
  int G = 1;
  if (D4 <= kCtxThreads / 2) { while (D4 * G * 2 <= kCtxThreads) G *= 2; }
  const size_t rowA = (size_t)a.A * es, rowF = (size_t)a.D * es;
  const int Lp = (a.L + 3) & ~3;
  size_t fixed = 0;
  auto take = [&](size_t bytes) { const size_t o = fixed; fixed = (fixed + bytes + 127) & ~(size_t)127; return (uint32_t)o; };
  y->off_att2 = take((size_t)2 * KB * a.A * 4);
  y->off_w = take((size_t)a.A * 4);
  y->off_e = take((size_t)kEBuf * (KB <= 4 ? 4 : 8) * Lp * 4);
  y->off_red = take(G > 1 ? (size_t)(G - 1) * KB * a.D * 4 : 16);
  y->off_bar = take((size_t)(2 * kMaxStagesA + 2 * kMaxStagesF + 2 * kEBuf) * 8 + 16);   // + s_misc: 2 range flags, sum(w)
  const size_t budget = 220 * 1024;
  constexpr int kStagesA = 2;
  y->stA = kStagesA;
  constexpr int kStagesF = 3;   // measured at C2: 3 stages of 7 rows beat 4 x 5, 5 x 4, 6 x 3, 8 x 2 (21.5 / 21.9 / 22.4 / 23.6 / 25.2 ms)
  y->stF = kStagesF;
  if (fixed + kStagesA * rowA + kStagesF * rowF > budget) return false;
  // att1 ring: one full pass of the score warps (2 rows each) per stage when it fits; the feats ring gets the rest
  const size_t left = budget - fixed;
  int rowsA = 2 * kScoreWarps, rowsF = 0;
  if (rowsA > a.L) rowsA = a.L;
  for (; rowsA >= 1; rowsA = rowsA > 1 ? rowsA / 2 : 0) {
    const size_t needA = (size_t)kStagesA * (((size_t)rowsA * rowA + 127) & ~(size_t)127);
    if (needA + kStagesF * (rowF + 128) > left) continue;
    rowsF = (int)(((left - needA) / kStagesF - 128) / rowF);
    if (rowsF >= 1) break;
  }
  if (rowsA < 1 || rowsF < 1) return false;
  rowsF = rowsF > a.L ? a.L : rowsF;
  if (rowsF > G) rowsF = rowsF / G * G;
  y->rowsA = rowsA; y->rowsF = rowsF;
  y->nA = (a.L + rowsA - 1) / rowsA; y->nF = (a.L + rowsF - 1) / rowsF;
  y->stageA = (uint32_t)(((size_t)rowsA * rowA + 127) & ~(size_t)127);
  y->stageF = (uint32_t)(((size_t)rowsF * rowF + 127) & ~(size_t)127);
  y->b8A = (uint32_t)rowsA * a.A * 2; y->b8F = (uint32_t)rowsF * a.D * 2;
  y->off_ringA = take((size_t)kStagesA * y->stageA);
  y->off_ringF = take((size_t)kStagesF * y->stageF);
  y->total = (uint32_t)fixed;
  y->Lp = Lp; y->G = G; y->gshift = 0;
  while ((1 << y->gshift) < G) ++y->gshift;
  (void)A4;
  return fixed <= 227 * 1024;
}

template <int KB>
int launch_stream(const AddAttnArgs& a, int act, const StreamLayout& y, cudaStream_t s) {
  const int nc = (a.D / 4 + kCtxThreads - 1) / kCtxThreads;
  const int grid = a.B < sm_count() ? a.B : sm_count();
#define CAPDEC_STREAM_LAUNCH(ACTV, NCV, BFV)                                                                              \
  {                                                                                                                       \
    auto kern = additive_attention_stream_kernel<KB, ACTV, NCV, BFV>;                                                     \
    CAPDEC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)y.total));             \
    CAPDEC_CHECK_CUDA(launch_k(kern, dim3(grid), dim3(kThreads), y.total, s, true, a, y));                                                                           \
  }
  if (a.tile_fmt == 2) {
    CAPDEC_REQUIRE(act == ACT_RELU && a.att1_b8 && a.feats_b8, CAPDEC_ERR_UNSUPPORTED, "p24 tiles: relu attention with both byte planes only");
    if (nc == 1) CAPDEC_STREAM_LAUNCH(ACT_RELU, 1, 2) else CAPDEC_STREAM_LAUNCH(ACT_RELU, 2, 2)
  } else if (a.tile_fmt) {
    if (act == ACT_RELU)           { if (nc == 1) CAPDEC_STREAM_LAUNCH(ACT_RELU, 1, 1) else CAPDEC_STREAM_LAUNCH(ACT_RELU, 2, 1) }
    else if (act == ACT_TANH_FAST) { if (nc == 1) CAPDEC_STREAM_LAUNCH(ACT_TANH_FAST, 1, 1) else CAPDEC_STREAM_LAUNCH(ACT_TANH_FAST, 2, 1) }
    else                           { if (nc == 1) CAPDEC_STREAM_LAUNCH(ACT_TANH, 1, 1) else CAPDEC_STREAM_LAUNCH(ACT_TANH, 2, 1) }
  } else {
    if (act == ACT_RELU)           { if (nc == 1) CAPDEC_STREAM_LAUNCH(ACT_RELU, 1, 0) else CAPDEC_STREAM_LAUNCH(ACT_RELU, 2, 0) }
    else if (act == ACT_TANH_FAST) { if (nc == 1) CAPDEC_STREAM_LAUNCH(ACT_TANH_FAST, 1, 0) else CAPDEC_STREAM_LAUNCH(ACT_TANH_FAST, 2, 0) }
    else                           { if (nc == 1) CAPDEC_STREAM_LAUNCH(ACT_TANH, 1, 0) else CAPDEC_STREAM_LAUNCH(ACT_TANH, 2, 0) }
  }
#undef CAPDEC_STREAM_LAUNCH
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

}  // namespace

bool additive_attention_stream_supports(int A, int D, int L, int k, int tile_fmt) {
  if (ab_switch("CAPDEC_ATTN_GENERIC") || k < 1 || k > kMaxRowsPerImage) return false;
  AddAttnArgs a{};
  a.A = A; a.D = D; a.L = L; a.k = k; a.tile_fmt = tile_fmt;
  StreamLayout y{};
  return plan(a, k <= 6 ? k : 8, &y);
}

// returns 1 when the streaming kernel took the call, 0 when the shape is left to the generic kernel, < 0 on error
int additive_attention_stream(const AddAttnArgs& a, int act, cudaStream_t s) {
  static const bool disabled = ab_switch("CAPDEC_ATTN_GENERIC");
  if (disabled || a.k < 1 || a.k > kMaxRowsPerImage) return 0;
  if (a.ld_att2 % 4 || a.ld_ctx % 4 || (a.gate && a.ld_gate % 4)) return 0;
  const int KB = a.k <= 6 ? a.k : 8;
  StreamLayout y{};
  if (!plan(a, KB, &y)) return 0;
  int st;
  switch (KB) {
    case 1: st = launch_stream<1>(a, act, y, s); break;
    case 2: st = launch_stream<2>(a, act, y, s); break;
    case 3: st = launch_stream<3>(a, act, y, s); break;
    case 4: st = launch_stream<4>(a, act, y, s); break;
    case 5: st = launch_stream<5>(a, act, y, s); break;
    case 6: st = launch_stream<6>(a, act, y, s); break;
    default: st = launch_stream<8>(a, act, y, s); break;
  }
  return st == CAPDEC_OK ? 1 : st;
}

}  // namespace capdec
