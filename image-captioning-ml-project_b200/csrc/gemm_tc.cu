// tcgen05 GEMM for the dense contractions of the decode step (all tensor-core precision modes):
//     C[M,N] = A[M,K] * W[N,K]^T + bias,  fused epilogues as in gemm_epilogue.cuh + the chunk-level ones below
// tcgen05 has no IEEE-fp32 MMA, so fp32-grade accuracy comes from a 3-term operand split
//     a = a_hi + a_lo,  w = w_hi + w_lo,    a*w ~= a_lo*w_hi + a_hi*w_lo + a_hi*w_hi   (lo*lo dropped)
// accumulated in fp32 in TMEM, either on kind::tf32 (hi = cvt.rna.tf32, lo = exact residual; CAPDEC_PREC_TF32X3) or
// on kind::f16 with bf16 operands (hi = bf16_rn(x), lo = bf16_rn(x - hi); CAPDEC_PREC_BF16X3, same three MMAs at
// twice the rate).  CAPDEC_PREC_TF32 / CAPDEC_PREC_BF16 are the single-term forms.
//
// Kernel (persistent, one CTA or one CTA PAIR per 128/256 x 256 output tile, K streamed in 128-byte SWIZZLE_128B blocks):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D loads of the A_hi/A_lo/W_hi/W_lo tiles, mbarrier tx
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (M = 128 or 256 with cta_group::2, N = 256),
//               tcgen05.commit releases shared-memory stages and publishes the (double-buffered) accumulator
//   warps 2-9   epilogue: tcgen05.ld 32 lanes x 32 columns -> registers -> fused epilogue -> global; warp w reads TMEM
//               lane quadrant w % 4 and column half (w - 2) / 4 of the tile, so every scheduler has two epilogue warps
// A operands may arrive pre-split (GemmArgs::A_hi/A_lo, written by the kernel that produced them); otherwise
// split_kernel makes the copies.  Weights are split once per handle.
#include <cuda.h>
#include <cuda_bf16.h>
#include <limits.h>
#include <stdlib.h>

#include "gemm_epilogue.cuh"
#include "handle.cuh"

namespace capdec {
namespace {

constexpr int BM = 128;
constexpr int kRowBytes = 128;        // one K block = one 128-byte swizzle row: 32 tf32 or 64 bf16 elements
constexpr int kThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quadrant)
constexpr uint32_t kSpinLimit = 1u << 22;  // bounded waits: a protocol bug traps instead of hanging the GPU

#ifdef CAPDEC_TIMELINE
// debug build only (scripts/gemm_timeline.py): per-CTA time stamps of the phases of one launch
__device__ unsigned long long g_timeline[2][320][64];
__device__ __forceinline__ void tl_stamp(int i) {
  unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  g_timeline[0][blockIdx.x][i] = t;
  g_timeline[1][blockIdx.x][i] = (unsigned long long)clock64();
}
#ifdef CAPDEC_TL_EPI
#define TL(i) do { if (EPI == CAPDEC_TL_EPI) tl_stamp(i); } while (0)   /* only launches with this epilogue (in-situ probe) */
#else
#define TL(i) tl_stamp(i)
#endif
#else
#define TL(i)
#endif

// ---- PTX wrappers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// explicit shared-space accesses for the epilogue's staging slab: its address comes out of integer alignment arithmetic on
// the dynamic shared-memory base, so plain dereferences compile to GENERIC ld / st (ST.E / LD.E in the SASS)
__device__ __forceinline__ void sts128(uint32_t a, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint4 lds128u(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, float x) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(x) : "memory"); }
__device__ __forceinline__ float lds32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
    if (spin > kSpinLimit) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// ---- CTA-pair (cta_group::2) variants: both CTAs of the pair run the same code; TMA completions of either CTA are
// counted on the LEADER's full barrier, tcgen05.commit multicasts its arrive to the same barrier offset in both CTAs
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t cta_rank) {   // address of the same offset in `cta_rank`
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint32_t leader_bar_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(leader_bar_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar_addr) {
  // relaxed: what is handed over is TMEM, ordered by tcgen05.wait::ld + tcgen05.fence::before_thread_sync; no generic
  // memory is published, so the memory barrier a release-arrive implies is not needed
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major operand tile, 128-byte rows, SWIZZLE_128B: 8-row groups are 1024 bytes apart (SBO), LBO unused (=1),
// descriptor version 1 (sm_100), layout type 2.  cute/arch/mma_sm100_desc.hpp::SmemDescriptor.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// cute/arch/mma_sm100_desc.hpp::InstrDescriptor: c_format F32 (1) @4, a/b format TF32 (2) @7/@10, K-major A and B,
// n_dim = N>>3 @17, m_dim = M>>4 @24.
// kind::tf32 operands use format code 2 (TF32); kind::f16 operands use 1 (BF16).
__host__ __device__ constexpr uint32_t make_idesc(int kind, int m, int n) {
  const uint32_t fmt = kind == KIND_BF16 ? 1u : 2u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// CG = CTAs cooperating on one output tile: 1 -> 128 x BN tile per CTA; 2 -> CTA pair (cta_group::2) on a 256 x BN tile,
// each CTA staging its own 128 rows of A and HALF of the W tile (BN/2 rows), which halves the shared-memory bytes per
// flop -- a single CTA's 3xTF32 main loop needs ~158 B/clk of shared-memory bandwidth (MMA operand reads + TMA fills),
// more than the SM's 128 B/clk; the pair needs ~106 B/clk.
template <int BN, int CG, int TERMS>
struct SmemLayout {
  static constexpr uint32_t kABytes = BM * kRowBytes;
  static constexpr uint32_t kWBytes = (BN / CG) * kRowBytes;
  // a stage holds one K block of A and W, as hi + lo copies in the 3-term modes.  The single-term modes have half the
  // bytes and a third of the MMA time per K block, so their ring is twice as deep: with 3 stages the loop waits on the
  // L2 -> shared-memory latency of every block (measured 46 us per GPT-2 projection GEMM against 5-17 us of MMA time)
  static constexpr uint32_t kStageBytes = TERMS == 3 ? 2 * kABytes + 2 * kWBytes : kABytes + kWBytes;
  static constexpr uint32_t kOffALo = kABytes;                                   // (3-term only)
  static constexpr uint32_t kOffWHi = TERMS == 3 ? 2 * kABytes : kABytes;
  static constexpr uint32_t kOffWLo = kOffWHi + kWBytes;                        // (3-term only)
  static constexpr int kStages = (CG == 2 ? 3 : 2) * (TERMS == 3 ? 1 : 2);
  // 8 epilogue warps x (32 x 32 floats): store transpose / top-k candidates.  The slabs sit on 1024-byte boundaries: a
  // slab written with the XOR swizzle of store_chunk IS a SWIZZLE_128B box, so a TMA tensor store can read it directly
  static constexpr uint32_t kStashOffset = kStages * kStageBytes;
  static constexpr uint32_t kStashBytes = 8 * 32 * 32 * 4;
  static constexpr uint32_t kBarOffset = kStashOffset + kStashBytes;
  static constexpr uint32_t kTotal = kBarOffset + 256 + 1024;  // + barriers + slack for manual 1024-byte alignment
};

// Persistent kernel: grid = min(#tiles, #SMs); every CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... (n fastest,
// so the CTAs running concurrently share A tiles in L2).  The accumulator is double-buffered in TMEM (2 x BN columns)
// so the epilogue of tile i overlaps the TMA/MMA main loop of tile i+1; the shared-memory ring runs across tiles.
template <int BN, int EPI, int TERMS, int TK, int CG, int KIND>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                    const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                    const __grid_constant__ CUtensorMap map_c, const GemmArgs p) {
  using SL = SmemLayout<BN, CG, TERMS>;
  constexpr int kStages = SL::kStages;
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SL::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;    // [2]
  uint64_t* tmem_init_bar = tmem_empty_bar + 2;    // [2]  stream-K: accumulator pre-loaded with the predecessor's partial sum
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_init_bar + 2);

  if (threadIdx.x == 0) TL(0);
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;   // position in the CTA pair; rank 0 issues the MMAs
  const int group = blockIdx.x / CG, num_groups = gridDim.x / CG;
  constexpr int TM = BM * CG;                                // output tile rows per CTA group
  const int n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = n_tiles * ((p.M + TM - 1) / TM);
  // Every CTA group owns one CONTIGUOUS run of `quota` tiles in m-major order (n fastest): consecutive tiles reuse the
  // same A rows from L2, and the epilogue threads (thread == output row) can carry per-row state across the n tiles of a
  // row block -- the fused vocabulary top-k keeps ONE candidate list per row and column half for the whole run.
  // The other epilogues keep the interleaved order (tile = group + i * num_groups): the groups running concurrently then
  // work on neighbouring tiles and their requests for the same A / W blocks coalesce in L2, which matters because the
  // 3-term main loop runs at the L2 -> shared-memory bandwidth limit.
  constexpr bool kContiguous = EPI == EPI_TOPK;
  const int quota = (num_tiles + num_groups - 1) / num_groups;
  const int tile_begin = kContiguous ? min(group * quota, num_tiles) : group;
  const int tile_end = kContiguous ? min(tile_begin + quota, num_tiles) : num_tiles;
  const int tile_step = kContiguous ? 1 : num_groups;
  constexpr int BK = KIND == KIND_BF16 ? 64 : 32;   // elements per 128-byte K block
  const int num_kb = (p.K + BK - 1) / BK;
  constexpr uint32_t kTmemCols = 2 * BN;           // 256 or 512: a power of two >= 32
  // ---- work items.  Default: whole tiles in the order above.  Stream-K (p.sk_part != nullptr; never with EPI_TOPK): the
  // launch's num_tiles * num_kb K blocks are cut into num_groups EQUAL contiguous ranges, so a shape whose tile count is
  // not a multiple of the resident CTA groups (80 gate-GEMM tiles on 74 pairs at 2560 rows: two waves, the second 8 %
  // full) costs total / groups instead of ceil(tiles / groups) tile times.  A range covers the tail of one tile, whole
  // tiles, and the head of another; a group walks its range BACKWARDS: the head part first (its raw accumulator goes to a
  // scratch slot + an epoch-tagged flag per epilogue warp), the tail part last (its epilogue adds the parts other groups
  // left for that tile -- they were those groups' FIRST items, so the wait is short and can never deadlock: every group
  // of the persistent grid is resident).
  const bool sk = EPI != EPI_TOPK && p.sk_part != nullptr;
  const int64_t sk_total = (int64_t)num_tiles * num_kb;
  auto sk_lo = [&](int g) { return (int)(sk_total * g / num_groups); };
  struct Work { int tile, kb0, kb1; };
  // cursor: next tile index (default) or the exclusive upper end of what is left of the range (stream-K)
  const int cursor0 = sk ? sk_lo(group + 1) : tile_begin;
  const int sk_begin = sk ? sk_lo(group) : 0;
  auto next_work = [&](int& cursor, Work& w) -> bool {
    if (sk) {
      if (cursor <= sk_begin) return false;
      w.tile = (cursor - 1) / num_kb;
      const int tb = w.tile * num_kb, s0 = max(sk_begin, tb);
      w.kb0 = s0 - tb; w.kb1 = cursor - tb;
      cursor = s0;
      return true;
    }
    if (cursor >= tile_end) return false;
    w.tile = cursor; w.kb0 = 0; w.kb1 = num_kb;
    cursor += tile_step;
    return true;
  };

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w_hi) : "memory");
    if (TERMS == 3) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_lo) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w_lo) : "memory");
    }
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], 8 * CG); mbar_init(&tmem_init_bar[b], 8 * CG); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // barriers initialised + TMEM allocated in BOTH CTAs of the pair
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) TL(1);
  pdl_wait();   // set-up above overlapped the predecessor's tail; its outputs (our operands) are visible from here on
  if (threadIdx.x == 0) TL(2);

  if (warp == 0) {
    // ===== TMA producer (both CTAs of a pair: own 128 rows of A, own BN/CG rows of W) =====
    if (lane == 0) {
      uint32_t it = 0;  // global k-block counter across tiles -> ring stage / phase
      int cursor = cursor0;
      Work wk;
      while (next_work(cursor, wk)) {
        const int tile = wk.tile;
        const int m_tile = tile / n_tiles, n_tile = tile - m_tile * n_tiles;
        const int a_row = m_tile * TM + (int)rank * BM;
        const int w_row = n_tile * BN + (int)rank * (BN / CG);
        for (int kb = wk.kb0; kb < wk.kb1; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* st = smem + s * SL::kStageBytes;
          constexpr uint32_t kTx = SL::kStageBytes;
          if (CG == 2) {
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * kTx);   // bytes of both CTAs land on the leader's barrier
            const uint32_t lbar = mapa_u32(smem_u32(&full_bar[s]), 0);
            tma_load_2d_pair(st, &map_a_hi, lbar, kb * BK, a_row);
            tma_load_2d_pair(st + SL::kOffWHi, &map_w_hi, lbar, kb * BK, w_row);
            if (TERMS == 3) {
              tma_load_2d_pair(st + SL::kOffALo, &map_a_lo, lbar, kb * BK, a_row);
              tma_load_2d_pair(st + SL::kOffWLo, &map_w_lo, lbar, kb * BK, w_row);
            }
          } else {
            mbar_arrive_expect_tx(&full_bar[s], kTx);
            tma_load_2d(st, &map_a_hi, &full_bar[s], kb * BK, a_row);
            tma_load_2d(st + SL::kOffWHi, &map_w_hi, &full_bar[s], kb * BK, w_row);
            if (TERMS == 3) {
              tma_load_2d(st + SL::kOffALo, &map_a_lo, &full_bar[s], kb * BK, a_row);
              tma_load_2d(st + SL::kOffWLo, &map_w_lo, &full_bar[s], kb * BK, w_row);
            }
          }
          if (it == 0) TL(3);
        }
      }
      TL(4);
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread; in pair mode only the leader CTA, its MMAs drive both SMs' tensor cores) =====
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc(KIND, TM, BN);
      auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t acc) {
        if (KIND == KIND_BF16) { if (CG == 2) mma_bf16_pair(d, a, b, idesc, acc); else mma_bf16(d, a, b, idesc, acc); }
        else                   { if (CG == 2) mma_tf32_pair(d, a, b, idesc, acc); else mma_tf32(d, a, b, idesc, acc); }
      };
      auto commit = [&](uint64_t* bar) { if (CG == 2) tcgen05_commit_pair(bar); else tcgen05_commit(bar); };
      uint32_t it = 0, local = 0;
      int cursor = cursor0;
      Work wk;
      for (; next_work(cursor, wk); ++local) {
        const uint32_t ab = local & 1;                       // accumulator buffer
        mbar_wait(&tmem_empty_bar[ab], ((local >> 1) & 1) ^ 1);  // the epilogue warps (of both CTAs) drained this buffer
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + ab * BN;
        const bool cont = sk && wk.kb0 > 0;      // stream-K: continue the sum the epilogue warps loaded into this buffer
        if (cont) { mbar_wait(&tmem_init_bar[ab], 0); tcgen05_fence_after(); }   // (at most one such item per group and launch)
        for (int kb = wk.kb0; kb < wk.kb1; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(&full_bar[s], ph);
          tcgen05_fence_after();
          if (it == 0) TL(5);
          const uint32_t a_hi = smem_u32(smem + s * SL::kStageBytes);
          const uint32_t a_lo = a_hi + SL::kOffALo;
          const uint32_t w_hi = a_hi + SL::kOffWHi;
          const uint32_t w_lo = a_hi + SL::kOffWLo;
#pragma unroll
          for (int k = 0; k < kRowBytes / 32; ++k) {
            const uint32_t koff = k * 32;  // one MMA consumes 32 bytes along K (8 tf32 / 16 bf16) inside the 128-byte swizzle span
            const uint32_t first = (kb == wk.kb0 && k == 0 && !cont) ? 0u : 1u;
            if (TERMS == 3) {
              mma(tmem_d, make_smem_desc(a_lo + koff), make_smem_desc(w_hi + koff), first);
              mma(tmem_d, make_smem_desc(a_hi + koff), make_smem_desc(w_lo + koff), 1u);
              mma(tmem_d, make_smem_desc(a_hi + koff), make_smem_desc(w_hi + koff), 1u);
            } else {
              mma(tmem_d, make_smem_desc(a_hi + koff), make_smem_desc(w_hi + koff), first);
            }
          }
          commit(&empty_bar[s]);                           // frees this shared-memory stage (in both CTAs) when the MMAs retire
          if (kb == wk.kb1 - 1) commit(&tmem_full_bar[ab]);  // accumulator (of this item's K range) complete
        }
        if (local < 8) TL(8 + local);
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue warps: TMEM -> registers -> fused epilogue -> global =====
    const int q = warp & 3;                      // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;            // which BN/2 column half of the tile this warp drains
    constexpr int HN = BN / 2;
    uint32_t local = 0;
    const uint32_t leader_empty_bar0 = CG == 2 ? mapa_u32(smem_u32(&tmem_empty_bar[0]), 0) : 0u;
    float* stash = reinterpret_cast<float*>(smem + SL::kStashOffset);   // 4 KB per epilogue warp
    // Coalesced chunk store.  tcgen05.ld hands every thread ONE ROW of the chunk (32 consecutive columns), so storing
    // from those registers makes each warp-wide 16-byte store touch 32 different rows: 32 half-filled sectors per
    // instruction.  At that request rate the L2 takes ~28 k cycles to absorb one 128 x 256 tile (measured with the phase
    // time stamps of scripts/gemm_timeline.py: 14.5 us per tile against 3.2 us of MMA time for a K = 768 GPT-2
    // projection -- every small-K GEMM was bound by its epilogue stores).  The chunk is therefore transposed through the
    // warp's 4 KB slab (XOR-swizzled 16-byte cells: conflict-free for the row-wise writes and the column-group reads) so
    // that 8 lanes cover one row's 128 bytes and a store instruction writes four complete 128-byte lines.
    const uint32_t slab_a = smem_u32(stash) + (uint32_t)(warp - 2) * 4096u;   // this warp's 4 KB slab (shared-space address)
    // Mirror-only output (C == nullptr: only the next GEMM reads it) in the bf16 operand type: the chunk is converted
    // first and transposed as packed bf16 -- a row of the chunk is 64 bytes, 4 lanes x 16 bytes, 8 rows per store.
    bool tma_pending = false;   // this warp's slab is the source of a bulk tensor store that may not have been read out yet
    auto store_chunk_bf16 = [&](const float (&o)[32], int row0, int col0) {
      uint32_t hw[16], lw[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float r0, r1, d0, d1;
        hw[j] = pack_bf16x2_(o[2 * j], o[2 * j + 1], &r0, &r1);
        lw[j] = TERMS == 3 ? pack_bf16x2_(r0, r1, &d0, &d1) : 0u;
      }
      // slab as [32 rows][4 cells of 16 bytes], cell index XOR-swizzled by (row >> 1) & 3 -- which is the SWIZZLE_64B box
      // layout, so a single-plane mirror (the single-term modes) leaves through one TMA bulk tensor store per chunk
      if (TERMS == 1 && p.c_tma == 2) {
        if (lane == 0 && tma_pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          sts128u(slab_a + (uint32_t)(lane * 4 + (j ^ ((lane >> 1) & 3))) * 16u, hw[4 * j], hw[4 * j + 1], hw[4 * j + 2], hw[4 * j + 3]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"(&map_c), "r"(slab_a), "r"(col0), "r"(row0) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        tma_pending = true;
        return;
      }
#pragma unroll
      for (int pass = 0; pass < (TERMS == 3 ? 2 : 1); ++pass) {
        const uint32_t* w = pass ? lw : hw;
        char* plane = reinterpret_cast<char*>(pass ? p.c_split.lo : p.c_split.hi);
        if (pass) __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          sts128u(slab_a + (uint32_t)(lane * 4 + (j ^ ((lane >> 1) & 3))) * 16u, w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = i * 8 + (lane >> 2), c = lane & 3;
          const uint4 t = lds128u(slab_a + (uint32_t)(r * 4 + (c ^ ((r >> 1) & 3))) * 16u);
          const int64_t row = row0 + r;
          if (row < p.M) *reinterpret_cast<uint4*>(plane + (row * p.c_split.ld + col0 + c * 8) * 2) = t;
        }
      }
      __syncwarp();
    };
    // C alone (no second copy, no operand mirror): the swizzled slab is handed to the TMA as a 32 x 32 box -- the warp
    // issues 8 shared stores and one bulk tensor store instead of 8 shared loads + 8 global stores, rows beyond M are
    // clipped by the tensor map.  The slab is rewritten only after the previous box has been read out.
    auto store_chunk_tma = [&](const float (&o)[32], int row0, int col0) {
      if (lane == 0 && tma_pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        sts128(slab_a + (uint32_t)(lane * 32 + ((j ^ (lane & 7)) << 2)) * 4u, o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                     ::"l"(&map_c), "r"(slab_a), "r"(col0), "r"(row0) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      tma_pending = true;
    };
    auto store_chunk = [&](const float (&o)[32], int row0, int col0, bool mirrors) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        sts128(slab_a + (uint32_t)(lane * 32 + ((j ^ (lane & 7)) << 2)) * 4u, o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + (lane >> 3), c4 = lane & 7;
        const float4 t = lds128(slab_a + (uint32_t)(r * 32 + ((c4 ^ (r & 7)) << 2)) * 4u);
        const int64_t row = row0 + r;
        if (row < p.M) {
          const int col = col0 + c4 * 4;
          if (p.C) *reinterpret_cast<float4*>(p.C + row * p.ldc + col) = t;
          if (mirrors) {
            if (p.C2) *reinterpret_cast<float4*>(p.C2 + row * p.ldc2 + col) = t;
            split_store4(p.c_split, row, col, t);
          }
        }
      }
      __syncwarp();
    };
    const int V = EPI == EPI_TOPK ? p.tk_vocab : p.N;                  // vocabulary columns; [n_vtiles*BN, N) is the tail block
    const int n_vtiles = (V + BN - 1) / BN;
    // EPI_TOPK state carried across the tiles of one row block: online log-sum-exp and a sorted top-TK list of this
    // thread's row over the columns of this warp's column half
    float rmax = -INFINITY, rsum = 0.f;
    float tv[TK > 0 ? TK : 1];
    int ti[TK > 0 ? TK : 1];
    int cur_block = -1;
    auto tk_reset = [&]() {
      rmax = -INFINITY; rsum = 0.f;
#pragma unroll
      for (int j = 0; j < (TK > 0 ? TK : 1); ++j) { tv[j] = -INFINITY; ti[j] = INT_MAX; }
    };
    auto tk_flush = [&](int block) {
      // record of (row, run segment of the block, column half): {0, 0, TK values, TK indices}
      constexpr int PS = (2 + 2 * (TK > 0 ? TK : 1) + 3) & ~3;
      const int row = block * TM + (int)rank * BM + q * 32 + lane;
      if (row >= p.M) return;
      const int slot = group - (block * n_tiles) / quota;   // which CTA group's run inside this row block
      const int n_rec = 2 * ((n_tiles + quota - 1) / quota + 1);
      float rec[PS];
      rec[0] = 0.f; rec[1] = 0.f;
#pragma unroll
      for (int j = 0; j < TK; ++j) { rec[2 + j] = tv[j]; rec[2 + TK + j] = __int_as_float(ti[j]); }
#pragma unroll
      for (int j = 2 + 2 * TK; j < PS; ++j) rec[j] = 0.f;
      float4* dst = reinterpret_cast<float4*>(p.tk_part + ((int64_t)row * n_rec + slot * 2 + half) * PS);
#pragma unroll
      for (int j = 0; j < PS / 4; ++j) dst[j] = make_float4(rec[4 * j], rec[4 * j + 1], rec[4 * j + 2], rec[4 * j + 3]);
    };
    int cursor = cursor0;
    Work wk;
    for (; next_work(cursor, wk); ++local) {
      const int tile = wk.tile;
      const int m_tile = tile / n_tiles, n_tile = tile - m_tile * n_tiles;
      const uint32_t ab = local & 1;
      const int m = m_tile * TM + (int)rank * BM + q * 32 + lane;   // accumulator row == TMEM lane of this CTA
      // stream-K roles of this item.  A part that starts inside the tile (kb0 > 0) CONTINUES the accumulation of the group
      // before this one: its epilogue warps copy that group's published accumulator into this item's TMEM buffer
      // (tcgen05.st) before the MMA warp issues with accumulate = 1, so the tile is summed in exactly the order a single
      // group would have used -- results stay bit-identical whatever the batch size / schedule.  A part that ends before
      // the tile's last K block (kb1 < num_kb) publishes its raw accumulator instead of running the epilogue.  Slot layout
      // (per group and CTA of the pair): [32-column chunk][column][row of the CTA]: lanes touch contiguous floats.
      const bool sk_publish = EPI != EPI_TOPK && sk && wk.kb1 < num_kb;
      if (EPI != EPI_TOPK && sk && wk.kb0 > 0) {
        if (lane == 0) {
          const volatile int* f = p.sk_flag + ((size_t)(group - 1) * CG + rank) * 8 + (warp - 2);
          for (uint32_t spin = 0; *f != p.sk_epoch; ++spin) if (spin > kSpinLimit) __trap();
          __threadfence();
        }
        __syncwarp();
#pragma unroll 1
        for (int c0 = half * HN; c0 < (half + 1) * HN; c0 += 32) {
          const float* src = p.sk_part + (((size_t)(group - 1) * CG + rank) * (BN / 32) + (c0 >> 5)) * (32 * BM) + q * 32 + lane;
          uint32_t v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__ldcg(src + j * BM));
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * BN + c0);
          asm volatile(
              "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
              "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
              "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
              :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                 "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
                 "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
                 "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
              : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {       // accumulator buffer `ab` of this CTA holds the predecessor's sum for this warp's lanes / columns
          if (CG == 2) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_init_bar[ab]), 0));
          else asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_init_bar[ab])) : "memory");
        }
      }
      mbar_wait(&tmem_full_bar[ab], (local >> 1) & 1);
      tcgen05_fence_after();
      if (warp == 2 && lane == 0 && local < 16) TL(16 + 2 * local);
      if constexpr (EPI == EPI_TOPK) {
        if (m_tile != cur_block) {
          if (cur_block >= 0) tk_flush(cur_block);
          tk_reset();
          cur_block = m_tile;
        }
        rmax = -INFINITY; rsum = 0.f;   // log-sum-exp partials are per (row, tile half): see tk_lse below
      }
      // EPI_STORE with tk_lse (the sampling path's logits GEMM): the same per-(row, tile half) {max, sum exp} partials next
      // to the stored logits, so the sampler locates its draw from 2 * ceil(V / 256) pairs instead of re-reading the row
      const bool store_lse = EPI == EPI_STORE && p.tk_lse != nullptr && !sk_publish;
      if (EPI == EPI_STORE) { rmax = -INFINITY; rsum = 0.f; }
      auto lse_chunk = [&](const float (&o)[32]) {   // o: this row's 32 logits of the chunk, -inf beyond the last column
        float t8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t8[j] = fmaxf(fmaxf(o[4 * j], o[4 * j + 1]), fmaxf(o[4 * j + 2], o[4 * j + 3]));
        const float cm = fmaxf(fmaxf(fmaxf(t8[0], t8[1]), fmaxf(t8[2], t8[3])), fmaxf(fmaxf(t8[4], t8[5]), fmaxf(t8[6], t8[7])));
        constexpr float kLog2e = 1.4426950408889634f;
        if (cm > rmax) { rsum *= ex2_approx_((rmax - cm) * kLog2e); rmax = cm; }
#pragma unroll
        for (int j = 0; j < 32; ++j) rsum += ex2_approx_((o[j] - rmax) * kLog2e);
      };
#pragma unroll 1
      for (int c0 = half * HN; c0 < (half + 1) * HN; c0 += 32) {
        uint32_t v[32];
        // this chunk's bias, requested before the accumulator load is waited for (full, 16-byte aligned chunks only)
        const int nb0 = n_tile * BN + c0;
        const bool pre_b = p.bias != nullptr && nb0 + 32 <= p.N;
        float4 bia[8];
        if (pre_b) {
#pragma unroll
          for (int j = 0; j < 8; ++j) bia[j] = __ldg(reinterpret_cast<const float4*>(p.bias + nb0) + j);
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * BN + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c0 + 32 >= (half + 1) * HN) {
          // every accumulator column this warp owns is now in registers: hand the TMEM buffer back to the MMA warp
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 2) mbar_arrive_cluster(leader_empty_bar0 + ab * 8);
            else asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[ab])) : "memory");
          }
        }
        if constexpr (EPI != EPI_TOPK) {
          if (sk_publish) {
            float* dst = p.sk_part + (((size_t)group * CG + rank) * (BN / 32) + (c0 >> 5)) * (32 * BM) + q * 32 + lane;
#pragma unroll
            for (int j = 0; j < 32; ++j) __stcg(dst + j * BM, __uint_as_float(v[j]));
            if (c0 + 32 >= (half + 1) * HN) {       // this warp's share of the part is written: publish it
              __threadfence();
              __syncwarp();
              if (lane == 0) *(volatile int*)(p.sk_flag + ((size_t)group * CG + rank) * 8 + (warp - 2)) = p.sk_epoch;
            }
            continue;
          }
        }
        const int n0 = n_tile * BN + c0;
        if (n0 < p.N) {
          if constexpr (EPI == EPI_TOPK) {
           if (n_tile >= n_vtiles) {
            // tail block behind the (padded) vocabulary columns: a plain projection of the same A rows, stored to C with
            // sigmoid on its columns >= n_split (the legacy step's [dec_att | f_beta] of the NEXT step rides here)
            const int c_tail = n0 - n_vtiles * BN;
            if (n0 + 32 <= p.N) {
              float o[32];
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b = pre_b ? bia[j >> 2] : make_float4(0.f, 0.f, 0.f, 0.f);
                o[j] = __uint_as_float(v[j]) + b.x;         o[j + 1] = __uint_as_float(v[j + 1]) + b.y;
                o[j + 2] = __uint_as_float(v[j + 2]) + b.z; o[j + 3] = __uint_as_float(v[j + 3]) + b.w;
                if (c_tail + j >= p.n_split) {   // (n_split % 4 == 0: the tail block's halves are multiples of 4 wide)
                  o[j] = sigmoid_fast_(o[j]); o[j + 1] = sigmoid_fast_(o[j + 1]); o[j + 2] = sigmoid_fast_(o[j + 2]); o[j + 3] = sigmoid_fast_(o[j + 3]);
                }
              }
              store_chunk(o, m - lane, c_tail, false);
            } else if (m < p.M) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const int col = c_tail + j;
                if (n0 + j < p.N) {
                  const float4 b = pre_b ? bia[j >> 2] : (p.bias ? *reinterpret_cast<const float4*>(p.bias + n0 + j) : make_float4(0.f, 0.f, 0.f, 0.f));
                  float4 t = make_float4(__uint_as_float(v[j]) + b.x, __uint_as_float(v[j + 1]) + b.y,
                                         __uint_as_float(v[j + 2]) + b.z, __uint_as_float(v[j + 3]) + b.w);
                  if (col >= p.n_split) { t.x = sigmoid_fast_(t.x); t.y = sigmoid_fast_(t.y); t.z = sigmoid_fast_(t.z); t.w = sigmoid_fast_(t.w); }
                  *reinterpret_cast<float4*>(p.C + (int64_t)m * p.ldc + col) = t;
                }
              }
            }
           } else if (n0 < V) {
            // logits of this 32-column chunk (bias added, columns beyond the vocabulary masked out)
            float x[32];
            const bool full = n0 + 32 <= V;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
              if (pre_b) b = bia[j >> 2];
              else if (p.bias) {
                if (full) b = *reinterpret_cast<const float4*>(p.bias + n0 + j);
                else {
                  b.x = n0 + j < V ? p.bias[n0 + j] : 0.f;         b.y = n0 + j + 1 < V ? p.bias[n0 + j + 1] : 0.f;
                  b.z = n0 + j + 2 < V ? p.bias[n0 + j + 2] : 0.f; b.w = n0 + j + 3 < V ? p.bias[n0 + j + 3] : 0.f;
                }
              }
              x[j] = __uint_as_float(v[j]) + b.x;         x[j + 1] = __uint_as_float(v[j + 1]) + b.y;
              x[j + 2] = __uint_as_float(v[j + 2]) + b.z; x[j + 3] = __uint_as_float(v[j + 3]) + b.w;
            }
            if (!full) {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (n0 + j >= V) x[j] = -INFINITY;
            }
            // The epilogue of a vocabulary tile has to fit under the tile's MMAs (K = 512, three terms: ~12 k cycles) with
            // two epilogue warps per scheduler, so dependent chains and issue slots both count (phase time stamps of
            // scripts/gemm_timeline.py: 8.5 us of epilogue against 7.7 us of MMA per tile before this form): the chunk
            // maximum is a tree and the exponentials are FADD + FMUL + MUFU.EX2 (expf's range fix-up is not needed: terms
            // below 2^-126 may flush to zero in a sum that contains exp(0) = 1).
            float t8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t8[j] = fmaxf(fmaxf(x[4 * j], x[4 * j + 1]), fmaxf(x[4 * j + 2], x[4 * j + 3]));
            const float cm = fmaxf(fmaxf(fmaxf(t8[0], t8[1]), fmaxf(t8[2], t8[3])), fmaxf(fmaxf(t8[4], t8[5]), fmaxf(t8[6], t8[7])));
            constexpr float kLog2e = 1.4426950408889634f;
            if (cm > rmax) { rsum *= ex2_approx_((rmax - cm) * kLog2e); rmax = cm; }   // x[0] is always a real column, so cm is finite
#pragma unroll
            for (int j = 0; j < 32; ++j) rsum += ex2_approx_((x[j] - rmax) * kLog2e);   // (column order: the sum's bits are part of the parity record)
            // Candidates of this chunk: columns above the row's current TK-th best.  Lanes (rows) find theirs at
            // different columns, so walking the columns in lockstep would make the warp pay for the union; instead the
            // chunk is parked in shared memory ([column][lane], conflict-free) and every lane pops ITS next candidate per
            // round -- the number of rounds is the largest per-lane count, not the size of the union.  A chunk in which no
            // row of the warp has a candidate (cm <= its threshold everywhere) is not parked at all.
            const float thr = tv[TK - 1];
            unsigned cand = 0u;
            const uint32_t st = slab_a + (uint32_t)lane * 4u;   // [column][lane]
            if (__any_sync(0xffffffffu, cm > thr)) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                sts32(st + (uint32_t)j * 128u, x[j]);
                if (x[j] > thr) cand |= 1u << j;
              }
            }
            while (__any_sync(0xffffffffu, cand != 0u)) {
              if (cand != 0u) {
                const int j = __ffs(cand) - 1;
                cand &= cand - 1u;
                const float xv = lds32(st + (uint32_t)j * 128u);
                if (xv > tv[TK - 1]) {
                  // branch-free sorted insert: x lands behind any equal value, so the lower vocabulary index stays
                  // ahead among equal logits (columns are popped in increasing order)
                  const int xi = n0 + j;
                  bool gt[TK];
#pragma unroll
                  for (int t = 0; t < TK; ++t) gt[t] = xv > tv[t];
#pragma unroll
                  for (int t = TK - 1; t > 0; --t) {
                    ti[t] = gt[t] ? (gt[t - 1] ? ti[t - 1] : xi) : ti[t];
                    tv[t] = fmaxf(tv[t], fminf(tv[t - 1], xv));
                  }
                  ti[0] = gt[0] ? xi : ti[0];
                  tv[0] = fmaxf(tv[0], xv);
                }
              }
            }
           }
          } else if (EPI == EPI_LSTM && pre_b && m < p.M && ((p.ldc | p.ldcin | p.ldcout | p.ldc2) & 3) == 0 &&
                     (p.c_split.hi == nullptr || ((p.c_split.ld & 7) == 0))) {
            // LSTM cell on the 8 hidden units of this 32-column chunk at once (gate columns are interleaved 4j + {i,f,g,o}):
            // 16-byte loads / stores of 8 consecutive units per row instead of 4-byte scattered ones
            const int j0 = n0 >> 2;
            const float4 ci0 = *reinterpret_cast<const float4*>(p.c_in + (int64_t)m * p.ldcin + j0);
            const float4 ci1 = *reinterpret_cast<const float4*>(p.c_in + (int64_t)m * p.ldcin + j0 + 4);
            const float cp[8] = {ci0.x, ci0.y, ci0.z, ci0.w, ci1.x, ci1.y, ci1.z, ci1.w};
            if (p.row_table) {   // embedding contribution of this row's token, pre-multiplied into gate space
              const float4* tb = reinterpret_cast<const float4*>(p.row_table + (int64_t)p.row_index[m] * p.ld_table + n0);
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const float4 tv4 = __ldg(tb + u);
                bia[u].x += tv4.x; bia[u].y += tv4.y; bia[u].z += tv4.z; bia[u].w += tv4.w;
              }
            }
            float hv[8], cv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float gi = __uint_as_float(v[4 * u]) + bia[u].x, gf = __uint_as_float(v[4 * u + 1]) + bia[u].y;
              const float gg = __uint_as_float(v[4 * u + 2]) + bia[u].z, go = __uint_as_float(v[4 * u + 3]) + bia[u].w;
              cv[u] = sigmoid_fast_(gf) * cp[u] + sigmoid_fast_(gi) * tanh_fast_(gg);
              hv[u] = sigmoid_fast_(go) * tanh_fast_(cv[u]);
            }
            float4* co = reinterpret_cast<float4*>(p.c_out + (int64_t)m * p.ldcout + j0);
            co[0] = make_float4(cv[0], cv[1], cv[2], cv[3]); co[1] = make_float4(cv[4], cv[5], cv[6], cv[7]);
            float4* ho = reinterpret_cast<float4*>(p.C + (int64_t)m * p.ldc + j0);
            ho[0] = make_float4(hv[0], hv[1], hv[2], hv[3]); ho[1] = make_float4(hv[4], hv[5], hv[6], hv[7]);
            if (p.C2) {
              float4* h2 = reinterpret_cast<float4*>(p.C2 + (int64_t)m * p.ldc2 + j0);
              h2[0] = make_float4(hv[0], hv[1], hv[2], hv[3]); h2[1] = make_float4(hv[4], hv[5], hv[6], hv[7]);
            }
            split_store8(p.c_split, m, j0, hv);
          } else if (epi_is_store_family(EPI) && nb0 + 32 <= p.N && (p.ldc & 3) == 0 && (!p.C2 || (p.ldc2 & 3) == 0) &&
                     (p.c_split.hi == nullptr || (p.c_split.ld & 3) == 0)) {
            // store-family epilogue on a full chunk: bias / activation in registers, then the coalesced store (fp32 copy
            // optional: C == nullptr when only the next GEMM reads the output, through its operand mirror)
            float o[32];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = pre_b ? bia[j >> 2] : make_float4(0.f, 0.f, 0.f, 0.f);
              o[j] = __uint_as_float(v[j]) + b.x;         o[j + 1] = __uint_as_float(v[j + 1]) + b.y;
              o[j + 2] = __uint_as_float(v[j + 2]) + b.z; o[j + 3] = __uint_as_float(v[j + 3]) + b.w;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (EPI == EPI_SIGMOID_TAIL && n0 + j >= p.n_split) o[j] = sigmoid_fast_(o[j]);
              if (EPI == EPI_TANH) o[j] = tanh_fast_(o[j]);
              if (EPI == EPI_GELU) o[j] = gelu_erf_(o[j]);
              if (EPI == EPI_GELU_TANH) o[j] = gelu_tanh_fast_(o[j]);   // (MUFU.TANH's 2^-11 was measured: it breaks the bf16 mode's 2e-2 bar at 124M)
            }
            if (store_lse) lse_chunk(o);
            if (KIND == KIND_BF16 && p.C == nullptr && p.C2 == nullptr && p.c_split.hi != nullptr && p.c_split.kind == KIND_BF16 &&
                p.c_split.b8 == nullptr && (TERMS == 3) == (p.c_split.lo != nullptr) && (p.c_split.ld & 7) == 0)
              store_chunk_bf16(o, m - lane, n0);
            else if (p.c_tma == 1)
              store_chunk_tma(o, m - lane, n0);
            else
              store_chunk(o, m - lane, n0, true);
          } else {
            if (store_lse) {   // (a ragged last chunk: columns beyond N do not exist)
              float o[32];
#pragma unroll
              for (int j = 0; j < 32; ++j)
                o[j] = n0 + j < p.N ? __uint_as_float(v[j]) + (p.bias ? __ldg(p.bias + n0 + j) : 0.f) : -INFINITY;
              lse_chunk(o);
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bj = pre_b ? bia[j >> 2] : make_float4(0.f, 0.f, 0.f, 0.f);
              epilogue4<EPI, true>(p, m, n0 + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                   __uint_as_float(v[j + 3]), pre_b ? &bj : nullptr);
            }
          }
        }
      }
      if (warp == 2 && lane == 0 && local < 16) TL(17 + 2 * local);
      if (EPI == EPI_STORE && store_lse && m < p.M && n_tile * BN + half * HN < p.N)
        *reinterpret_cast<float2*>(p.tk_lse + ((int64_t)m * (2 * n_tiles) + n_tile * 2 + half) * 2) = make_float2(rmax, rsum);
      if constexpr (EPI == EPI_TOPK) {
        // {max, sum exp(x - max)} of this row over this tile half.  Kept per tile (not folded along the run) so that the
        // merge kernel can combine them in one canonical order: the result does not depend on how the tiles were
        // distributed over CTA groups, i.e. a row decodes identically whatever the batch size.
        if (m < p.M && n_tile < n_vtiles && n_tile * BN + half * HN < V)
          *reinterpret_cast<float2*>(p.tk_lse + ((int64_t)m * (2 * n_vtiles) + n_tile * 2 + half) * 2) = make_float2(rmax, rsum);
      }
    }
    if constexpr (EPI == EPI_TOPK) {
      if (cur_block >= 0) tk_flush(cur_block);
    }
    if (lane == 0 && tma_pending) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the bulk stores are complete before the CTA retires
  }
  tcgen05_fence_before();
  if (threadIdx.x == 64) TL(62);
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // pair mode: the peer's shared memory / barriers stay alive until both are done
  if (threadIdx.x == 0) TL(63);
  if (warp == 1) {
    tcgen05_fence_after();
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// Operand split for the tensor-core modes.  Output is dense [rows, Kp] (Kp >= cols, zero padded so that a row is a
// multiple of 16 bytes for TMA).
//   tf32: hi = round-to-nearest TF32 of x (exactly representable, so the tensor core's own fp32->tf32 conversion is the
//         identity whatever its rounding), lo = x - hi (exact in fp32)
//   bf16: hi = bf16_rn(x), lo = bf16_rn(x - hi)
// lo == nullptr (single-pass modes) skips the residual.
template <int KIND>
__global__ void split_kernel(const float* __restrict__ x, int64_t ld, int rows, int cols, int Kp, void* __restrict__ hi,
                             void* __restrict__ lo) {
  pdl_trigger();
  pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // group of 4 columns
  const int c4 = Kp >> 2;
  if (i >= (int64_t)rows * c4) return;
  const int r = (int)(i / c4), c = (int)(i - (int64_t)r * c4);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c * 4 < cols) v = ldg_stream(reinterpret_cast<const float4*>(x + (int64_t)r * ld + c * 4));   // cols % 4 == 0
  if (KIND == KIND_TF32) {
    float4 h, l;
    uint32_t t;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.x)); h.x = __uint_as_float(t); l.x = v.x - h.x;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.y)); h.y = __uint_as_float(t); l.y = v.y - h.y;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.z)); h.z = __uint_as_float(t); l.z = v.z - h.z;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.w)); h.w = __uint_as_float(t); l.w = v.w - h.w;
    reinterpret_cast<float4*>(hi)[i] = h;
    if (lo) reinterpret_cast<float4*>(lo)[i] = l;
  } else {
    float r0, r1, r2, r3, d0, d1;
    uint2 hp;
    hp.x = pack_bf16x2_(v.x, v.y, &r0, &r1);
    hp.y = pack_bf16x2_(v.z, v.w, &r2, &r3);
    reinterpret_cast<uint2*>(hi)[i] = hp;
    if (lo) {
      uint2 lp;
      lp.x = pack_bf16x2_(r0, r1, &d0, &d1);
      lp.y = pack_bf16x2_(r2, r3, &d0, &d1);
      reinterpret_cast<uint2*>(lo)[i] = lp;
    }
  }
}

int split_operand(int kind, const float* x, int64_t ld, int rows, int cols, int Kp, void* hi, void* lo, cudaStream_t s) {
  const int64_t n = (int64_t)rows * (Kp / 4);
  if (n == 0) return CAPDEC_OK;
  const dim3 grid((unsigned)((n + 255) / 256));
  if (kind == KIND_BF16) CAPDEC_CHECK_CUDA(launch_k(split_kernel<KIND_BF16>, grid, dim3(256), 0, s, true, x, ld, rows, cols, Kp, hi, lo));
  else                   CAPDEC_CHECK_CUDA(launch_k(split_kernel<KIND_TF32>, grid, dim3(256), 0, s, true, x, ld, rows, cols, Kp, hi, lo));
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

// ---- host side --------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeTiledFn* out) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    CAPDEC_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q));
    CAPDEC_REQUIRE(ptr != nullptr && q == cudaDriverEntryPointSuccess, CAPDEC_ERR_CUDA,
                   "cuTensorMapEncodeTiled is not available from this driver");
    fn = (EncodeTiledFn)ptr;
  }
  *out = fn;
  return CAPDEC_OK;
}

// 2-D tensor [rows, cols] of tf32/bf16 elements with row stride ld (elements); box = [box_rows, one 128-byte K block];
// 128-byte swizzle; OOB -> 0
int make_map(CUtensorMap* m, int kind, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_row_bytes = kRowBytes) {
  EncodeTiledFn fn;
  CAPDEC_RETURN_IF(get_encode_fn(&fn));
  const size_t es = kind == KIND_BF16 ? 2 : 4;
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * es};
  const cuuint32_t box[2] = {(cuuint32_t)(box_row_bytes / es), (cuuint32_t)box_rows};   // box rows of 128 bytes (operands, fp32 outputs) or 64
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(m, kind == KIND_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                        const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        box_row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CAPDEC_REQUIRE(r == CUDA_SUCCESS, CAPDEC_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d (rows=%lld cols=%lld ld=%lld)",
                 (int)r, (long long)rows, (long long)cols, (long long)ld);
  return CAPDEC_OK;
}

// CTA groups one launch may keep resident: every SM for single CTAs; for pairs, what the occupancy calculator grants
// clusters of 2 (both SMs of a TPC).  All instantiations of a mode share thread count and shared-memory size, so one
// representative kernel answers for all of them.  tk_records() (the EPI_TOPK record layout) depends on this number.
int tc_max_groups(int cg) {
  static std::atomic<int> cache_dev[kMaxDevices][3];
  std::atomic<int>* cache = cache_dev[current_device()];
  if (cache[cg].load()) return cache[cg].load();
  if (cg == 1) { cache[1].store(num_sms()); return num_sms(); }
  auto kern = gemm_tcgen05_kernel<256, EPI_STORE, 3, 0, 2, KIND_TF32>;
  constexpr int smem = SmemLayout<256, 2, 3>::kTotal;
  int n = 0;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * num_sms()); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) n = 0;
  }
  cudaGetLastError();
  if (n <= 0 || n > num_sms() / 2) n = num_sms() / 2;
  cache[2].store(n);
  return n;
}
int tc_cta_group(int M) {
  return M > BM ? 2 : 1;   // CTA pairs (cta_group::2, 256-row tiles) whenever there is more than one 128-row tile
}

template <int BN, int TERMS, int CG, int KIND>
int launch_tc(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi, const CUtensorMap& w_lo,
              const CUtensorMap& c_map, const GemmArgs& g, int epi, cudaStream_t s) {
  const int num_tiles = ceil_div(g.N, BN) * ceil_div(g.M, BM * CG);
#define CAPDEC_TC_LAUNCH(E, TKV)                                                                                  \
  {                                                                                                               \
    auto kern = gemm_tcgen05_kernel<BN, E, TERMS, TKV, CG, KIND>;                                                 \
    constexpr int smem = SmemLayout<BN, CG, TERMS>::kTotal;                                                         \
    static std::atomic<bool> configured[kMaxDevices];   /* function attributes are per device */                 \
    const int dev_ = current_device();                                                                            \
    if (!configured[dev_].load()) {                                                                               \
      CAPDEC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));           \
      configured[dev_].store(true);                                                                               \
    }                                                                                                             \
    const int max_groups = tc_max_groups(CG);                                                                     \
    const int groups = (num_tiles < max_groups && !g.sk_part) ? num_tiles : max_groups;                           \
    cudaLaunchConfig_t cfg = {};                                                                                  \
    cfg.gridDim = dim3(groups * CG); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = s;  \
    cudaLaunchAttribute at[2];                                                                                    \
    at[0].id = cudaLaunchAttributeClusterDimension;                                                               \
    at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;                          \
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                                \
    at[1].val.programmaticStreamSerializationAllowed = 1;                                                         \
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 2 : 1;                                                         \
    CAPDEC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, a_hi, a_lo, w_hi, w_lo, c_map, g));                          \
  }
#define CAPDEC_TC_CASE(E) case E: CAPDEC_TC_LAUNCH(E, 0) break;
  switch (epi) {
    CAPDEC_TC_CASE(EPI_STORE)
    CAPDEC_TC_CASE(EPI_SIGMOID_TAIL)
    CAPDEC_TC_CASE(EPI_LSTM)
    CAPDEC_TC_CASE(EPI_TANH)
    CAPDEC_TC_CASE(EPI_AOA)
    CAPDEC_TC_CASE(EPI_GELU)
    CAPDEC_TC_CASE(EPI_GELU_TANH)
    case EPI_TOPK:
      if constexpr (BN == 256) {   // the record layout is defined on the column halves of 256-column tiles
        switch (tk_bucket(g.tk_k)) {
          case 1: CAPDEC_TC_LAUNCH(EPI_TOPK, 1) break;
          case 6: CAPDEC_TC_LAUNCH(EPI_TOPK, 6) break;
          case 10: CAPDEC_TC_LAUNCH(EPI_TOPK, 10) break;
          default: CAPDEC_TC_LAUNCH(EPI_TOPK, 16) break;
        }
        break;
      } else {
        CAPDEC_REQUIRE(false, CAPDEC_ERR_INVALID, "gemm_tc: EPI_TOPK needs the 256-column tile");
      }
    default: CAPDEC_REQUIRE(false, CAPDEC_ERR_INVALID, "gemm_tc: unknown epilogue %d", epi);
  }
#undef CAPDEC_TC_CASE
#undef CAPDEC_TC_LAUNCH
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

struct Scratch {
  float* p = nullptr;
  size_t bytes = 0;
};
Scratch g_scratch;  // handle-less path (capdec_linear): grow-only, process lifetime

int ensure(float** p, size_t* have, size_t need) {
  if (*have >= need) return CAPDEC_OK;
  if (*p) CAPDEC_CHECK_CUDA(cudaFree(*p));
  *p = nullptr; *have = 0;
  CAPDEC_CHECK_CUDA(cudaMalloc((void**)p, need));
  *have = need;
  return CAPDEC_OK;
}

}  // namespace

int tc_kind(int precision) { return (precision == CAPDEC_PREC_BF16 || precision == CAPDEC_PREC_BF16X3) ? KIND_BF16 : KIND_TF32; }
int tc_terms(int precision) { return (precision == CAPDEC_PREC_TF32X3 || precision == CAPDEC_PREC_BF16X3) ? 3 : 1; }
// one-off split of a dense fp32 matrix into caller-provided copies with row pitch `cols` (cols % 8 == 0)
int tc_split(int precision, const float* x, int64_t ld, int rows, int cols, void* hi, void* lo, cudaStream_t s) {
  return split_operand(tc_kind(precision), x, ld, rows, cols, cols, hi, tc_terms(precision) == 3 ? lo : nullptr, s);
}

// The EPI_TOPK launch's tile schedule for an [M,N] problem (the merge kernel derives from it which record slots of a row
// block were written): n tiles per row block, tiles per CTA group (contiguous runs), rows per block.
void tk_schedule(int M, int N, int* n_tiles_out, int* quota_out, int* block_rows_out) {
  const int cg = tc_cta_group(M > 0 ? M : 1);
  const int n_tiles = ceil_div(N, 256), num_tiles = n_tiles * ceil_div(M > 0 ? M : 1, BM * cg);
  const int max_groups = tc_max_groups(cg);
  const int groups = num_tiles < max_groups ? num_tiles : max_groups;
  *n_tiles_out = n_tiles; *quota_out = ceil_div(num_tiles, groups); *block_rows_out = BM * cg;
}
// records per row the EPI_TOPK epilogue may write: 2 column halves x (runs a row block can span)
int tk_records(int M, int N) {
  int n_tiles, quota, br;
  tk_schedule(M, N, &n_tiles, &quota, &br);
  return 2 * (ceil_div(n_tiles, quota) + 1);
}

int gemm_tc(const capdec_handle* h, int precision, const GemmArgs& a, int epilogue, cudaStream_t s) {
  CAPDEC_REQUIRE(precision == CAPDEC_PREC_TF32X3 || precision == CAPDEC_PREC_TF32 || precision == CAPDEC_PREC_BF16 ||
                     precision == CAPDEC_PREC_BF16X3,
                 CAPDEC_ERR_UNSUPPORTED, "precision mode %d has no kernel in this build", precision);
  CAPDEC_REQUIRE(a.M >= 0 && a.N > 0 && a.K > 0, CAPDEC_ERR_INVALID, "gemm: bad shape M=%d N=%d K=%d", a.M, a.N, a.K);
  if (a.M == 0) return CAPDEC_OK;
  CAPDEC_REQUIRE(a.K % 4 == 0 && a.lda % 4 == 0 && a.ldw % 4 == 0, CAPDEC_ERR_UNSUPPORTED,
                 "gemm: K, lda, ldw must be multiples of 4 (K=%d lda=%lld ldw=%lld)", a.K, (long long)a.lda, (long long)a.ldw);
  if (epilogue == EPI_LSTM || epilogue == EPI_AOA)
    CAPDEC_REQUIRE(a.N % 4 == 0, CAPDEC_ERR_UNSUPPORTED, "gemm: fused epilogue needs N %% 4 == 0 (N=%d)", a.N);
  CAPDEC_REQUIRE((((uintptr_t)a.A | (uintptr_t)a.W | (uintptr_t)a.bias | (uintptr_t)a.C) & 15) == 0, CAPDEC_ERR_INVALID,
                 "gemm: A/W/bias/C must be 16-byte aligned");
  if (epilogue == EPI_TOPK)
    CAPDEC_REQUIRE(a.tk_part && a.tk_lse && ((((uintptr_t)a.tk_part) | (uintptr_t)a.tk_lse) & 15) == 0 && tk_supported(a.N, a.tk_k), CAPDEC_ERR_INVALID,
                   "gemm: EPI_TOPK needs an aligned partial buffer, 1 <= k <= 16 and N <= 131072 (k=%d N=%d)", a.tk_k, a.N);
  const int terms = (precision == CAPDEC_PREC_TF32X3 || precision == CAPDEC_PREC_BF16X3) ? 3 : 1;
  const int kind = (precision == CAPDEC_PREC_BF16 || precision == CAPDEC_PREC_BF16X3) ? KIND_BF16 : KIND_TF32;
  const size_t es = kind == KIND_BF16 ? 2 : 4;
  const int K = a.K;
  const int Kp = kind == KIND_BF16 ? (K + 7) & ~7 : K;   // operand rows must be multiples of 16 bytes for TMA

  // ---- weight split: cached per handle (weights are immutable after capdec_finalize)
  char *w_hi = nullptr, *w_lo = nullptr;
  const size_t w_bytes = align_up((size_t)a.N * Kp * es, 256);
  if (h) {
    auto it = h->tc_weights.find(a.W);
    if (it == h->tc_weights.end()) {
      // first use of this weight: split it once and keep the copies.  The split is COMPLETE before this call returns
      // (stream sync), so a later call on another stream can never read half-written copies; inside a stream capture
      // neither the allocation nor the sync is legal, so a captured loop must have been run once eagerly before.
      cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
      CAPDEC_CHECK_CUDA(cudaStreamIsCapturing(s, &cap));
      CAPDEC_REQUIRE(cap == cudaStreamCaptureStatusNone, CAPDEC_ERR_STATE,
                     "gemm: weight operand copies are built on first use; run the call once outside stream capture first");
      float* buf = nullptr;
      CAPDEC_CHECK_CUDA(cudaMalloc((void**)&buf, 2 * w_bytes));
      CAPDEC_RETURN_IF(split_operand(kind, a.W, a.ldw, a.N, K, Kp, buf, terms == 3 ? (char*)buf + w_bytes : nullptr, s));
      CAPDEC_CHECK_CUDA(cudaStreamSynchronize(s));
      h->tc_weights[a.W] = buf;
      it = h->tc_weights.find(a.W);
    }
    w_hi = (char*)it->second; w_lo = w_hi + w_bytes;
  }
  // ---- activation split scratch, M chunked so the scratch stays <= ~1 GiB
  const size_t cap_bytes = (size_t)1 << 30;
  int m_chunk = (int)((cap_bytes / (2 * (size_t)Kp * es)) / (2 * BM) * (2 * BM));
  if (m_chunk < 2 * BM) m_chunk = 2 * BM;
  if (m_chunk > a.M) m_chunk = a.M;
  const size_t a_bytes = align_up((size_t)m_chunk * Kp * es, 256);
  char* scratch = nullptr;
  if (h) {
    CAPDEC_RETURN_IF(ensure(&h->tc_scratch, &h->tc_scratch_bytes, 2 * a_bytes));
    scratch = (char*)h->tc_scratch;
  } else {
    CAPDEC_RETURN_IF(ensure(&g_scratch.p, &g_scratch.bytes, 2 * a_bytes + 2 * w_bytes));
    scratch = (char*)g_scratch.p;
    w_hi = scratch + 2 * a_bytes; w_lo = w_hi + w_bytes;
    CAPDEC_RETURN_IF(split_operand(kind, a.W, a.ldw, a.N, K, Kp, w_hi, terms == 3 ? w_lo : nullptr, s));
  }
  char* a_hi = scratch;
  char* a_lo = scratch + a_bytes;

  constexpr int bn = 256;
  const int cg = tc_cta_group(a.M);
  if (epilogue == EPI_TOPK) {
    const int tail = a.N - ceil_div(a.tk_vocab, 256) * 256;
    CAPDEC_REQUIRE(a.tk_vocab >= 1 && a.tk_vocab <= a.N && (tail <= 0 || (tail % 4 == 0 && a.C && a.ldc % 4 == 0)), CAPDEC_ERR_INVALID,
                   "gemm: EPI_TOPK tail block must start at a 256-column boundary behind the vocabulary and be a multiple of 4 wide");
    CAPDEC_REQUIRE(a.M <= m_chunk, CAPDEC_ERR_UNSUPPORTED, "gemm: EPI_TOPK with %d rows exceeds the single-launch limit %d", a.M, m_chunk);
    // (record slots a row block does not use stay unwritten; topk_merge derives the used slots from tk_schedule)
  }
  CUtensorMap map_w_hi, map_w_lo;
  CAPDEC_RETURN_IF(make_map(&map_w_hi, kind, w_hi, a.N, Kp, Kp, bn / cg));
  CAPDEC_RETURN_IF(make_map(&map_w_lo, kind, terms == 3 ? w_lo : w_hi, a.N, Kp, Kp, bn / cg));

  const bool pre = a.A_hi != nullptr;   // the producer of A already wrote the split copies
  if (pre) {
    CAPDEC_REQUIRE((terms == 1 || a.A_lo) && (a.ld_as * es) % 16 == 0 && (((uintptr_t)a.A_hi | (uintptr_t)a.A_lo) & 15) == 0,
                   CAPDEC_ERR_INVALID, "gemm: pre-split A must be 16-byte aligned with a 16-byte row pitch and carry a lo copy");
    m_chunk = a.M;
  }
  for (int m0 = 0; m0 < a.M; m0 += m_chunk) {
    const int mc = a.M - m0 < m_chunk ? a.M - m0 : m_chunk;
    CUtensorMap map_a_hi, map_a_lo;
    if (pre) {
      CAPDEC_RETURN_IF(make_map(&map_a_hi, kind, a.A_hi, mc, K, a.ld_as, BM));
      CAPDEC_RETURN_IF(make_map(&map_a_lo, kind, terms == 3 ? a.A_lo : a.A_hi, mc, K, a.ld_as, BM));
    } else {
      CAPDEC_RETURN_IF(split_operand(kind, a.A + (int64_t)m0 * a.lda, a.lda, mc, K, Kp, a_hi, terms == 3 ? a_lo : nullptr, s));
      CAPDEC_RETURN_IF(make_map(&map_a_hi, kind, a_hi, mc, Kp, Kp, BM));
      CAPDEC_RETURN_IF(make_map(&map_a_lo, kind, terms == 3 ? a_lo : a_hi, mc, Kp, Kp, BM));
    }
    GemmArgs g = a;
    g.M = mc;
    g.K = pre ? K : Kp;
    // stream-K when whole tiles would leave the last wave mostly idle (see the kernel): needs the handle's scratch
    g.sk_part = nullptr; g.sk_flag = nullptr; g.sk_epoch = 0;
    static const bool no_sk = ab_switch("CAPDEC_NO_STREAMK");
    if (h && epilogue != EPI_TOPK && !no_sk && a.tk_lse == nullptr) {
      const int G = tc_max_groups(cg);
      const int tiles = ceil_div(a.N, bn) * ceil_div(mc, BM * cg);
      const int num_kb = ceil_div(g.K, kind == KIND_BF16 ? 64 : 32);
      const int waves = ceil_div(tiles, G);
      // cost model in K-block times: whole tiles take waves * num_kb; stream-K takes the even share plus ~30 blocks' worth of
      // un-overlapped part hand-over (publish the head part, copy the predecessor's sum into TMEM) -- measured on B200 for the
      // legacy gate GEMM (K = 2560): 80.1 vs 91.1 us at 2560 rows, 128.7 vs 136.1 at 5120, 244.8 vs 236.0 at 10240 rows
      if (tiles > G && (int64_t)waves * num_kb * G > (int64_t)tiles * num_kb + (int64_t)30 * G) {
        if (!h->sk_part) {
          cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
          CAPDEC_CHECK_CUDA(cudaStreamIsCapturing(s, &cap));
          if (cap == cudaStreamCaptureStatusNone) {
            const size_t slots = (size_t)tc_max_groups(2) * 2 > (size_t)tc_max_groups(1) ? (size_t)tc_max_groups(2) * 2 : (size_t)tc_max_groups(1);
            CAPDEC_CHECK_CUDA(cudaMalloc((void**)&h->sk_part, slots * BM * bn * sizeof(float)));
            CAPDEC_CHECK_CUDA(cudaMalloc((void**)&h->sk_flag, slots * 8 * sizeof(int)));
            CAPDEC_CHECK_CUDA(cudaMemsetAsync(h->sk_flag, 0, slots * 8 * sizeof(int), s));
          }
        }
        if (h->sk_part) {
          h->sk_epoch = (h->sk_epoch % 0x3fffffff) + 1;
          g.sk_part = h->sk_part; g.sk_flag = h->sk_flag; g.sk_epoch = h->sk_epoch;
        }
      }
    }
    g.C = a.C ? a.C + (int64_t)m0 * a.ldc : nullptr;
    if (a.C2) g.C2 = a.C2 + (int64_t)m0 * a.ldc2;
    if (a.c_in) g.c_in = a.c_in + (int64_t)m0 * a.ldcin;
    if (a.c_out) g.c_out = a.c_out + (int64_t)m0 * a.ldcout;
    if (a.c_split.hi && m0 > 0) {      // the output mirror advances with the row chunk too
      const size_t ces = a.c_split.kind == KIND_BF16 ? 2 : 4;
      g.c_split.hi = (char*)a.c_split.hi + (size_t)m0 * a.c_split.ld * ces;
      if (a.c_split.lo) g.c_split.lo = (char*)a.c_split.lo + (size_t)m0 * a.c_split.ld * ces;
      if (a.c_split.b8) g.c_split.b8 = a.c_split.b8 + (size_t)m0 * a.c_split.ld;
    }
    if (a.row_index) g.row_index = a.row_index + m0;
    if (a.tk_lse && epilogue != EPI_TOPK) g.tk_lse = a.tk_lse + (size_t)m0 * tk_lse_pairs(a.N) * 2;
    int st;
    // C stored by TMA (plain store-family output, nothing else written per element): a [rows, N] fp32 map with 32 x 32 boxes
    CUtensorMap map_c = map_w_hi;   // (placeholder when unused)
    g.c_tma = 0;
    if (epi_is_store_family(epilogue) && g.C && !g.C2 && !g.c_split.hi && a.N >= 32 && (a.ldc & 3) == 0 && (((uintptr_t)g.C) & 15) == 0) {
      CAPDEC_RETURN_IF(make_map(&map_c, KIND_TF32, g.C, mc, a.N, a.ldc, 32));
      g.c_tma = 1;
    } else if (epi_is_store_family(epilogue) && !g.C && !g.C2 && kind == KIND_BF16 && terms == 1 && g.c_split.hi && !g.c_split.lo && !g.c_split.b8 &&
               g.c_split.kind == KIND_BF16 && a.N >= 32 && (g.c_split.ld & 7) == 0 && (((uintptr_t)g.c_split.hi) & 15) == 0) {
      // mirror-only output of a single-term mode (one bf16 plane): 32 x 32 boxes of 64-byte rows
      CAPDEC_RETURN_IF(make_map(&map_c, KIND_BF16, g.c_split.hi, mc, a.N, g.c_split.ld, 32, 64));
      g.c_tma = 2;
    }
#define CAPDEC_TC_DISPATCH(TERMSV, CGV, KINDV) launch_tc<256, TERMSV, CGV, KINDV>(map_a_hi, map_a_lo, map_w_hi, map_w_lo, map_c, g, epilogue, s)
    if (kind == KIND_TF32) {
      if (cg == 1) st = terms == 3 ? CAPDEC_TC_DISPATCH(3, 1, KIND_TF32) : CAPDEC_TC_DISPATCH(1, 1, KIND_TF32);
      else         st = terms == 3 ? CAPDEC_TC_DISPATCH(3, 2, KIND_TF32) : CAPDEC_TC_DISPATCH(1, 2, KIND_TF32);
    } else {
      if (cg == 1) st = terms == 3 ? CAPDEC_TC_DISPATCH(3, 1, KIND_BF16) : CAPDEC_TC_DISPATCH(1, 1, KIND_BF16);
      else         st = terms == 3 ? CAPDEC_TC_DISPATCH(3, 2, KIND_BF16) : CAPDEC_TC_DISPATCH(1, 2, KIND_BF16);
    }
#undef CAPDEC_TC_DISPATCH
    CAPDEC_RETURN_IF(st);
  }
  return CAPDEC_OK;
}

int gemm_tc_prepare(capdec_handle* h, cudaStream_t) {
  for (auto& kv : h->tc_weights) cudaFree(kv.second);
  h->tc_weights.clear();
  return CAPDEC_OK;
}

void gemm_tc_release(capdec_handle* h) {
  for (auto& kv : h->tc_weights) cudaFree(kv.second);
  h->tc_weights.clear();
  if (h->tc_scratch) cudaFree(h->tc_scratch);
  h->tc_scratch = nullptr; h->tc_scratch_bytes = 0;
  if (h->sk_part) cudaFree(h->sk_part);
  if (h->sk_flag) cudaFree(h->sk_flag);
  h->sk_part = nullptr; h->sk_flag = nullptr;
}

}  // namespace capdec

#ifdef CAPDEC_TIMELINE
extern "C" int capdec_debug_timeline(unsigned long long* host_out, int clear) {
  if (clear) {
    void* p = nullptr;
    if (cudaGetSymbolAddress(&p, capdec::g_timeline) != cudaSuccess) return -1;
    return cudaMemset(p, 0, sizeof(capdec::g_timeline)) == cudaSuccess ? 0 : -1;
  }
  return cudaMemcpyFromSymbol(host_out, capdec::g_timeline, sizeof(capdec::g_timeline)) == cudaSuccess ? 0 : -1;
}
#endif
