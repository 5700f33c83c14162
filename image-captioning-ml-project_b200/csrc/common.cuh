// Shared helpers for libcapdec (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include <atomic>
#include <string>

#include "../../include/capdec.h"

namespace capdec {

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launch_count;   // atomic: handles on different host threads launch concurrently

#define CAPDEC_CHECK_CUDA(expr)                                                          \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      capdec::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CAPDEC_ERR_CUDA;                                                            \
    }                                                                                    \
  } while (0)

#define CAPDEC_RETURN_IF(expr)          \
  do {                                  \
    int _s = (expr);                    \
    if (_s != CAPDEC_OK) return _s;     \
  } while (0)

#define CAPDEC_REQUIRE(cond, code, ...)     \
  do {                                      \
    if (!(cond)) {                          \
      capdec::set_error(__VA_ARGS__);       \
      return (code);                        \
    }                                       \
  } while (0)

// every kernel launch in the library goes through this so launches can be counted and checked
#define CAPDEC_LAUNCH_CHECK()                                                            \
  do {                                                                                   \
    ++capdec::g_launch_count;                                                            \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      capdec::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CAPDEC_ERR_CUDA;                                                            \
    }                                                                                    \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Per-device caches: one process may drive several GPUs (Engine(device=...)), and SM counts, occupancy answers and
// cudaFuncSetAttribute state belong to a device, not to the process.
constexpr int kMaxDevices = 64;
inline int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return (d < 0 || d >= kMaxDevices) ? 0 : d;
}
int num_sms();   // SM count of the current device (capdec.cu)
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- A/B switches ----------------------------------------------------------------------------------------------
// The library has a small, closed set of environment switches, each selecting an ALTERNATIVE TESTED CODE PATH so that a
// measurement or a parity test can compare the two in one process (DESIGN.md section 6 lists them).  They are never needed
// in production; everything else about the schedule (ring depths, grid sizes, chunking) is a compile-time constant.
//   CAPDEC_NO_FUSED_TOPK   vocabulary GEMM stores logits, lse_topk_kernel selects
//   CAPDEC_NO_PRESPLIT     GEMMs split their A operand themselves (no producer-written hi / lo mirrors)
//   CAPDEC_NO_P24_TILES    bf16x3 mode streams fp32 region tiles instead of the p24 planes
//   CAPDEC_NO_BF16_TILES   bf16 mode streams fp32 region tiles
//   CAPDEC_EXACT_TANH      soft attention uses tanhf in the tensor-core modes too
//   CAPDEC_ATTN_GENERIC    one-CTA-per-image attention kernels instead of the persistent TMA-streamed ones
//   CAPDEC_NO_STREAMK      tcgen05 GEMM always schedules whole tiles
//   CAPDEC_NO_PDL          launches without programmatic dependent launch
inline bool ab_switch(const char* name) { return getenv(name) != nullptr; }

// ---- programmatic dependent launch ------------------------------------------------------------------------
// The decode loop is a chain of dependent kernels on one stream.  Kernels launched through launch_k(..., pdl = true)
// may become resident while their predecessor is still draining (every kernel calls pdl_trigger() first thing) and do
// their own set-up (barrier init, TMEM allocation, descriptor prefetch); pdl_wait() then blocks until the predecessor
// has completed and its writes are visible.  No global memory is touched before pdl_wait().  CAPDEC_NO_PDL=1 turns the
// launch attribute off (the device-side calls are then no-ops).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
inline bool pdl_enabled() {
  static const bool on = !ab_switch("CAPDEC_NO_PDL");
  return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  int n = 0;
  if (pdl && pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = at; cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

// ---- device helpers -----------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
// MUFU-based forms for the tensor-core epilogues (ex2.approx + rcp: ~1e-6 relative / 2e-7 absolute error, far below the
// MMA accumulation noise of those modes); the exact-fp32 GEMM keeps expf / IEEE division
// (rcp.approx is the bare MUFU.RCP; __frcp_rn expands to a Newton fix-up plus a slow-path call and is 3x the instructions)
__device__ __forceinline__ float rcp_approx_(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ex2_approx_(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sigmoid_fast_(float x) { return rcp_approx_(1.f + ex2_approx_(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_fast_(float x) {
  return fmaf(-2.f, rcp_approx_(1.f + ex2_approx_(2.8853900817779268f * x)), 1.f);
}
__device__ __forceinline__ float gelu_erf_(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_tanh_(float x) {   // transformers.activations.NewGELUActivation
  return 0.5f * x * (1.f + tanhf(0.79788456080286535588f * (x + 0.044715f * x * x * x)));
}
// tensor-core modes: tanh through ex2.approx / rcp.approx (absolute error ~1e-7 on the (1 + tanh) factor; tanhf is ~30
// instructions per element, which made the GPT-2 c_fc epilogue -- 32 k elements per 128 x 256 tile -- twice as long as
// the tile's MMAs)
__device__ __forceinline__ float gelu_tanh_fast_(float x) {
  return 0.5f * x * (1.f + tanh_fast_(0.79788456080286535588f * (x + 0.044715f * x * x * x)));
}
__host__ __device__ constexpr bool epi_is_store_family(int e) { return e == 0 || e == 1 || e == 3 || e == 5 || e == 6; }

// streaming 128-bit load: read-only path, do not allocate in L1 (tiles are read once per CTA)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// ---- tensor-core operand formats ------------------------------------------------------------------------
// The tcgen05 GEMM reads K-major operands as hi (+ lo) copies: tf32 = two fp32 arrays (hi = round-to-nearest TF32,
// lo = exact residual); bf16 = two bf16 arrays (hi = bf16_rn(x), lo = bf16_rn(x - hi)).  Kernels that PRODUCE a GEMM
// operand (attention context, state gather, LSTM epilogue) can write these copies themselves through a SplitDst, so no
// separate split pass over the operand is needed.
enum Kind : int { KIND_TF32 = 0, KIND_BF16 = 1 };
struct SplitDst {
  void* hi; void* lo;      // lo == nullptr: single-term mode; hi == nullptr: no split copy wanted
  int64_t ld;              // row stride in ELEMENTS of the operand type
  int kind;
  uint8_t* b8;             // KIND_BF16 only, optional: the "p24" byte plane (below), same row stride
};
// ---- p24: planar 24-bit storage of an fp32 tile that is streamed every step (attention region tiles, bf16x3 mode) ----
// plane 1 = the top 16 bits of each fp32 (a truncated bf16: it doubles as the "hi" GEMM operand), plane 2 = one byte
// q = round(low16 / 257).  Decoding replicates the byte, (hi16 << 16) | (q << 8) | q = hi16:q*257, so it is a single
// byte-permute per element, needs no carry into the upper plane when encoding, and is within 128.5 fp32 ulps (2^-16
// relative) of the original -- the same 16 significant bits the 3-term bf16 GEMMs of that mode keep of every operand.
__device__ __forceinline__ uint32_t p24_q_(uint32_t bits) { return (((bits & 0xffffu) + 128u) * 65281u) >> 24; }   // == (low16 + 128) / 257
__device__ __forceinline__ float4 p24_decode4_(uint2 hi, uint32_t q) {
  return make_float4(__uint_as_float(__byte_perm(hi.x, q, 0x1044)), __uint_as_float(__byte_perm(hi.x, q, 0x3255)),
                     __uint_as_float(__byte_perm(hi.y, q, 0x1066)), __uint_as_float(__byte_perm(hi.y, q, 0x3277)));
}
// encode 4 values: returns the 4 hi16 words, the 4 bytes, and the remainders v - hi (for the GEMM's lo operand)
__device__ __forceinline__ void p24_encode4_(float4 v, uint2* hi, uint32_t* q, float4* rem) {
  const uint32_t b0 = __float_as_uint(v.x), b1 = __float_as_uint(v.y), b2 = __float_as_uint(v.z), b3 = __float_as_uint(v.w);
  hi->x = __byte_perm(b0, b1, 0x7632); hi->y = __byte_perm(b2, b3, 0x7632);
  *q = p24_q_(b0) | (p24_q_(b1) << 8) | (p24_q_(b2) << 16) | (p24_q_(b3) << 24);
  *rem = make_float4(v.x - __uint_as_float(b0 & 0xffff0000u), v.y - __uint_as_float(b1 & 0xffff0000u),
                     v.z - __uint_as_float(b2 & 0xffff0000u), v.w - __uint_as_float(b3 & 0xffff0000u));
}
// {bf16_rn(a), bf16_rn(b)} in one word (a in the low half) + the remainders a - hi, b - hi.  One packed conversion
// (F2FP.BF16.PACK_AB) instead of two scalar F2F.BF16.F32: the scalar form runs on the quarter-rate conversion / MUFU pipe,
// which the GELU epilogues and the attention kernels' operand mirrors saturate.
__device__ __forceinline__ uint32_t pack_bf16x2_(float a, float b, float* ra, float* rb) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  *ra = a - __uint_as_float(d << 16); *rb = b - __uint_as_float(d & 0xffff0000u);
  return d;
}
// write v (4 consecutive columns starting at `col`, col % 4 == 0) of row `row` into the split copies
__device__ __forceinline__ void split_store4(const SplitDst& d, int64_t row, int col, float4 v) {
  if (!d.hi) return;
  if (d.kind == KIND_TF32) {
    float4 h, l;
    uint32_t t;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.x)); h.x = __uint_as_float(t); l.x = v.x - h.x;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.y)); h.y = __uint_as_float(t); l.y = v.y - h.y;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.z)); h.z = __uint_as_float(t); l.z = v.z - h.z;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.w)); h.w = __uint_as_float(t); l.w = v.w - h.w;
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(d.hi) + row * d.ld + col) = h;
    if (d.lo) *reinterpret_cast<float4*>(reinterpret_cast<float*>(d.lo) + row * d.ld + col) = l;
  } else if (d.b8) {
    uint2 hp, lp; uint32_t q; float4 r; float dummy0, dummy1;
    p24_encode4_(v, &hp, &q, &r);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(d.hi) + row * d.ld + col) = hp;
    *reinterpret_cast<uint32_t*>(d.b8 + row * d.ld + col) = q;
    if (d.lo) {
      lp.x = pack_bf16x2_(r.x, r.y, &dummy0, &dummy1);
      lp.y = pack_bf16x2_(r.z, r.w, &dummy0, &dummy1);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(d.lo) + row * d.ld + col) = lp;
    }
  } else {
    float r0, r1, r2, r3, dummy0, dummy1;
    uint2 hp, lp;
    hp.x = pack_bf16x2_(v.x, v.y, &r0, &r1);
    hp.y = pack_bf16x2_(v.z, v.w, &r2, &r3);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(d.hi) + row * d.ld + col) = hp;
    if (d.lo) {
      lp.x = pack_bf16x2_(r0, r1, &dummy0, &dummy1);
      lp.y = pack_bf16x2_(r2, r3, &dummy0, &dummy1);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(d.lo) + row * d.ld + col) = lp;
    }
  }
}
// 8 consecutive columns (col % 8 == 0): one 16-byte store per bf16 copy
__device__ __forceinline__ void split_store8(const SplitDst& d, int64_t row, int col, const float (&v)[8]) {
  if (!d.hi) return;
  if (d.kind == KIND_TF32) {
    split_store4(d, row, col, make_float4(v[0], v[1], v[2], v[3]));
    split_store4(d, row, col + 4, make_float4(v[4], v[5], v[6], v[7]));
  } else if (d.b8) {
    uint2 h0, h1, q; float4 r0, r1; float dm0, dm1;
    p24_encode4_(make_float4(v[0], v[1], v[2], v[3]), &h0, &q.x, &r0);
    p24_encode4_(make_float4(v[4], v[5], v[6], v[7]), &h1, &q.y, &r1);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(d.hi) + row * d.ld + col) = make_uint4(h0.x, h0.y, h1.x, h1.y);
    *reinterpret_cast<uint2*>(d.b8 + row * d.ld + col) = q;
    if (d.lo) {
      uint4 lp;
      lp.x = pack_bf16x2_(r0.x, r0.y, &dm0, &dm1); lp.y = pack_bf16x2_(r0.z, r0.w, &dm0, &dm1);
      lp.z = pack_bf16x2_(r1.x, r1.y, &dm0, &dm1); lp.w = pack_bf16x2_(r1.z, r1.w, &dm0, &dm1);
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(d.lo) + row * d.ld + col) = lp;
    }
  } else {
    float r[8], dm0, dm1;
    uint4 hp, lp;
    hp.x = pack_bf16x2_(v[0], v[1], &r[0], &r[1]); hp.y = pack_bf16x2_(v[2], v[3], &r[2], &r[3]);
    hp.z = pack_bf16x2_(v[4], v[5], &r[4], &r[5]); hp.w = pack_bf16x2_(v[6], v[7], &r[6], &r[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(d.hi) + row * d.ld + col) = hp;
    if (d.lo) {
      lp.x = pack_bf16x2_(r[0], r[1], &dm0, &dm1); lp.y = pack_bf16x2_(r[2], r[3], &dm0, &dm1);
      lp.z = pack_bf16x2_(r[4], r[5], &dm0, &dm1); lp.w = pack_bf16x2_(r[6], r[7], &dm0, &dm1);
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(d.lo) + row * d.ld + col) = lp;
    }
  }
}
__device__ __forceinline__ void split_store1(const SplitDst& d, int64_t row, int col, float v) {
  if (!d.hi) return;
  if (d.kind == KIND_TF32) {
    uint32_t t;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v));
    const float h = __uint_as_float(t);
    reinterpret_cast<float*>(d.hi)[row * d.ld + col] = h;
    if (d.lo) reinterpret_cast<float*>(d.lo)[row * d.ld + col] = v - h;
  } else if (d.b8) {
    const uint32_t b = __float_as_uint(v);
    reinterpret_cast<uint16_t*>(d.hi)[row * d.ld + col] = (uint16_t)(b >> 16);
    d.b8[row * d.ld + col] = (uint8_t)p24_q_(b);
    if (d.lo) reinterpret_cast<__nv_bfloat16*>(d.lo)[row * d.ld + col] = __float2bfloat16_rn(v - __uint_as_float(b & 0xffff0000u));
  } else {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    reinterpret_cast<__nv_bfloat16*>(d.hi)[row * d.ld + col] = h;
    if (d.lo) reinterpret_cast<__nv_bfloat16*>(d.lo)[row * d.ld + col] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// ---- GEMM front door (gemm_ffma.cu / gemm_tc.cu) ---------------------------------------------------
enum Epilogue : int {
  EPI_STORE = 0,         // C = acc + bias
  EPI_SIGMOID_TAIL = 1,  // columns >= n_split get sigmoid(acc + bias)   (legacy [dec_att | f_beta] GEMM)
  EPI_LSTM = 2,          // N = 4*H gate-interleaved (n = 4*j + {i,f,g,o}); fused cell update
  EPI_TANH = 3,          // tanh(acc + bias)
  EPI_AOA = 4,           // N = 2*H interleaved (n = 2*j + {info, gate}); out = tanh(info)*sigmoid(gate)
  EPI_GELU = 5,          // exact (erf) GELU of acc + bias        (nn.TransformerDecoderLayer activation="gelu")
  EPI_GELU_TANH = 6,     // tanh-approximated GELU ("gelu_new")    (HF GPT-2 MLP)
  EPI_TOPK = 7           // vocabulary projection fused with log-softmax partials + per-row top-k: the logits never
                         // reach HBM.  Per (row, 128-column half tile) the epilogue emits {max, sum exp(x-max), top-TK values,
                         // top-TK indices}; topk_merge (select.cu) combines the records.  Tensor-core path only.
};

// ---- fused vocab top-k partial records (EPI_TOPK) ---------------------------------------------------------
// The GEMM's CTA groups own contiguous runs of tiles, so a row block's n tiles fall into at most a few runs; per
// (row, run, 128-column-parity half) the epilogue emits one candidate record {top-TK values, top-TK indices}; the
// log-sum-exp partials {max, sum exp(x-max)} are emitted per (row, tile half) so their combination order is canonical.
static inline int tk_bucket(int k) { return k <= 1 ? 1 : k <= 6 ? 6 : k <= 10 ? 10 : 16; }   // compiled list lengths
static inline int tk_stride(int k) { return (2 + 2 * tk_bucket(k) + 3) & ~3; }              // floats per record
int tk_records(int M, int N);                                                              // records per row (gemm_tc.cu)
void tk_schedule(int M, int N, int* n_tiles, int* quota, int* block_rows);                 // the launch's tile schedule
static inline int tk_lse_pairs(int vocab) { return 2 * ((vocab + 255) / 256); }                // 128-column halves of 256-column tiles
static inline bool tk_supported(int vocab, int k) { return k >= 1 && k <= 16 && vocab >= 1; }

struct GemmArgs {
  const float* A; int64_t lda;      // [M,K]
  const float* W; int64_t ldw;      // [N,K]  (nn.Linear weight layout)
  const float* bias;                // [N] or nullptr
  float* C; int64_t ldc;            // EPI_STORE/SIGMOID_TAIL/TANH: [M,N]; EPI_LSTM: h_out [M,H]; EPI_AOA: out [M,H]
  int M, N, K;
  int n_split;                      // EPI_SIGMOID_TAIL
  const float* c_in; int64_t ldcin; // EPI_LSTM: previous cell state [M,H]
  float* c_out; int64_t ldcout;     // EPI_LSTM: new cell state [M,H]
  float* C2; int64_t ldc2;          // optional second copy of the primary output (nullptr = none)
  const void* A_hi; const void* A_lo; int64_t ld_as;   // tensor-core path: A already split by its producer (row stride ld_as elements)
  // EPI_LSTM: per-row additive term of the gate pre-activations, looked up by row: gates[m, :] += row_table[row_index[m], :]
  // (the embedding part of the LSTM input folded into a [V, 4H] table: the GEMM's K loses the embedding columns)
  const float* row_table; const int32_t* row_index; int64_t ld_table;
  SplitDst c_split;                 // store-family / EPI_LSTM epilogues: also write the output as the split operand of its consumer GEMM
  float* tk_part; int tk_k;         // EPI_TOPK: candidate records [M, tk_records(M,N), tk_stride(tk_k)], requested list length
  float* tk_lse;                    // EPI_TOPK: {max, sum exp} per (row, 128-column tile half): [M, tk_lse_pairs(tk_vocab), 2]
  float* sk_part; int* sk_flag; int sk_epoch;   // stream-K (gemm_tc.cu, set by the launcher): part scratch, per-warp flags, launch tag
  int c_tma;                        // tensor-core path, set by the launcher: 1 = C (alone: no C2 / mirror), 2 = the single bf16 mirror plane (alone) is stored with TMA bulk tensor stores
  int tk_vocab;                     // EPI_TOPK: vocabulary columns V <= N.  Columns [ceil(V/256)*256, N) are a tail block stored
                                    // to C (ld ldc) as a plain projection, sigmoid on tail columns >= n_split
};

int gemm_ffma(const GemmArgs& a, int epilogue, cudaStream_t s);

}  // namespace capdec
