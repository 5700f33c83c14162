// Encoder -> decoder feature hand-off: ONE pass from what an encoder emits to what the decode kernels stream.
//
// The reference hands features over as fp32 [B,L,D] after a layout change done by separate eager ops:
//   models/encoder.py:12-16            resnet trunk [B,2048,14,14] -> permute(0,2,3,1) -> (decoder.py:127) view [B,196,2048]
//   src/models/encoders.py:118-137     ViT last_hidden_state [B,197,768][:, 1:, :]   (CLS dropped)
//   src/models/encoders.py:209-230     CLIP last_hidden_state [B,50,768][:, 1:, :]
// and the decoder then takes the region mean (models/decoder.py:137) and re-projects the features.  Here the layout
// change, the widening from the encoder's bf16 / fp16 autocast output, the region mean and -- for the legacy decoder's
// BF16X3 mode -- the p24 planes the attention kernel streams plus the hi / lo operands of the hoisted enc_att GEMM are
// written by the same pass that reads the encoder output, so HBM sees the features once.
//
// Both kernels are HBM-bound: source bytes read once, each output written once.
#include "ingest.cuh"

#include <cuda_fp16.h>

namespace capdec {
namespace {

__device__ __forceinline__ uint2 ldg_stream_u2(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ldg_stream_u1(const void* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ float bf16_bits_to_f32(uint32_t b) { return __uint_as_float(b << 16); }
__device__ __forceinline__ float f16_bits_to_f32(uint32_t b) { return __half2float(__ushort_as_half((unsigned short)b)); }

// 4 consecutive elements starting at element offset `off` (off % 4 == 0) of a row-major image
template <int DT>
__device__ __forceinline__ float4 load4(const char* img, int64_t off, int64_t plane_elems) {
  if (DT == CAPDEC_DT_F32) return ldg_stream(reinterpret_cast<const float4*>(img) + (off >> 2));
  if (DT == CAPDEC_DT_P24) {
    const uint2 hi = ldg_stream_u2(img + off * 2);
    const uint32_t q = ldg_stream_u1(img + plane_elems * 2 + off);
    return p24_decode4_(hi, q);
  }
  const uint2 w = ldg_stream_u2(img + off * 2);
  if (DT == CAPDEC_DT_BF16)
    return make_float4(bf16_bits_to_f32(w.x & 0xffffu), bf16_bits_to_f32(w.x >> 16), bf16_bits_to_f32(w.y & 0xffffu),
                       bf16_bits_to_f32(w.y >> 16));
  return make_float4(f16_bits_to_f32(w.x & 0xffffu), f16_bits_to_f32(w.x >> 16), f16_bits_to_f32(w.y & 0xffffu),
                     f16_bits_to_f32(w.y >> 16));
}
template <int DT>
__device__ __forceinline__ float load1(const char* base, int64_t off) {
  if (DT == CAPDEC_DT_F32) return __ldg(reinterpret_cast<const float*>(base) + off);
  const uint32_t b = __ldg(reinterpret_cast<const unsigned short*>(base) + off);
  return DT == CAPDEC_DT_BF16 ? bf16_bits_to_f32(b) : f16_bits_to_f32(b);
}

// Row-major sources ([B,L,D], [B,1+L,D] with the CLS row skipped, p24-packed images): thread = 4 adjacent columns,
// walking the L regions, 4 rows of loads in flight; the mean is a per-thread register sum in region order.
template <int DT>
__global__ void __launch_bounds__(256) ingest_rows_kernel(const IngestArgs a, int64_t img_bytes, int64_t skip_bytes) {
  const int b = blockIdx.x;
  const int c = blockIdx.y * blockDim.x + threadIdx.x;   // float4 column
  const int D = a.D, L = a.L;
  if (c >= D / 4) return;
  const char* img = reinterpret_cast<const char*>(a.src) + (int64_t)b * img_bytes + skip_bytes;
  const int64_t plane = (int64_t)L * D;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int l0 = 0; l0 < L; l0 += 4) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = load4<DT>(img, (int64_t)min(l0 + u, L - 1) * D + c * 4, plane);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (l0 + u < L) {
        acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
        const int64_t row = (int64_t)b * L + l0 + u;
        if (a.out_f32) *reinterpret_cast<float4*>(a.out_f32 + row * D + c * 4) = v[u];
        split_store4(a.split, row, c * 4, v[u]);
      }
    }
  }
  if (a.mean) {
    const float fl = (float)L;
    reinterpret_cast<float4*>(a.mean + (int64_t)b * D)[c] = make_float4(acc.x / fl, acc.y / fl, acc.z / fl, acc.w / fl);
  }
}

// Channel-major source [B,D,L] (NCHW feature map): a CTA transposes a [64 channels, L] slab through shared memory --
// reads run along l (contiguous in the source), writes along d (contiguous in every output).
constexpr int kTC = 64;   // channels per slab: a region row of the slab is 128 B of hi plane, 64 B of byte plane, 128 B of lo
template <int DT>
__global__ void __launch_bounds__(256) ingest_transpose_kernel(const IngestArgs a, int pitch) {
  extern __shared__ float tile[];   // [kTC][pitch], pitch odd
  const int b = blockIdx.x, d0 = blockIdx.y * kTC;
  const int D = a.D, L = a.L;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t es = DT == CAPDEC_DT_F32 ? 4 : 2;
  const char* img = reinterpret_cast<const char*>(a.src) + (int64_t)b * D * L * es;
  // a warp owns channels warp, warp + 8, ...; it takes them four at a time and issues all 32 loads per lane before the
  // first shared-memory store, so a warp keeps 4 KB of the source in flight (one 128-byte row at a time made the kernel
  // latency-bound at 2.3 TB/s)
  constexpr int CH = 4;
  for (int c0 = warp; c0 < kTC; c0 += 8 * CH) {
    for (int l0 = 0; l0 < L; l0 += 256) {
      float v[CH][8];
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const int c = c0 + 8 * j;
        const bool live = c < kTC && d0 + c < D;
        const int64_t base = (int64_t)(d0 + c) * L;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int l = l0 + lane + 32 * u;
          v[j][u] = (live && l < L) ? load1<DT>(img, base + l) : 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const int c = c0 + 8 * j;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int l = l0 + lane + 32 * u;
          if (c < kTC && l < L) tile[c * pitch + l] = v[j][u];
        }
      }
    }
  }
  __syncthreads();
  if (a.mean && threadIdx.x < kTC && d0 + threadIdx.x < D) {
    float s = 0.f;
    for (int l = 0; l < L; ++l) s += tile[threadIdx.x * pitch + l];   // region order, as the row-major kernel
    a.mean[(int64_t)b * D + d0 + threadIdx.x] = s / (float)L;
  }
  // 16 threads per region row (4 channels each): a half-warp writes one row's 128 / 64 / 128 contiguous bytes
  for (int i = threadIdx.x; i < L * (kTC / 4); i += 256) {
    const int l = i / (kTC / 4), cg = i - l * (kTC / 4);
    const int d = d0 + cg * 4;
    if (d >= D) continue;
    const float4 v = make_float4(tile[(cg * 4) * pitch + l], tile[(cg * 4 + 1) * pitch + l], tile[(cg * 4 + 2) * pitch + l],
                                 tile[(cg * 4 + 3) * pitch + l]);
    const int64_t row = (int64_t)b * L + l;
    if (a.out_f32) *reinterpret_cast<float4*>(a.out_f32 + row * D + d) = v;
    split_store4(a.split, row, d, v);
  }
}

}  // namespace

size_t ingest_source_image_bytes(int layout, int dtype, int L, int D) {
  const size_t rows = layout == CAPDEC_LAYOUT_CLS_BLD ? (size_t)L + 1 : (size_t)L;
  const size_t es = dtype == CAPDEC_DT_F32 ? 4 : dtype == CAPDEC_DT_P24 ? 3 : 2;
  return rows * (size_t)D * es;
}

int ingest_features(const IngestArgs& a, cudaStream_t s) {
  CAPDEC_REQUIRE(a.layout >= CAPDEC_LAYOUT_BLD && a.layout <= CAPDEC_LAYOUT_CLS_BLD, CAPDEC_ERR_INVALID, "ingest: unknown layout %d", a.layout);
  CAPDEC_REQUIRE(a.dtype >= CAPDEC_DT_F32 && a.dtype <= CAPDEC_DT_P24, CAPDEC_ERR_INVALID, "ingest: unknown dtype %d", a.dtype);
  CAPDEC_REQUIRE(a.dtype != CAPDEC_DT_P24 || a.layout == CAPDEC_LAYOUT_BLD, CAPDEC_ERR_UNSUPPORTED, "ingest: p24 sources are [B,L,D] only");
  CAPDEC_REQUIRE(a.B >= 0 && a.L >= 1 && a.D >= 4 && a.D % 4 == 0, CAPDEC_ERR_INVALID, "ingest: bad sizes B=%d L=%d D=%d", a.B, a.L, a.D);
  CAPDEC_REQUIRE(a.src != nullptr || a.B == 0, CAPDEC_ERR_INVALID, "ingest: null source");
  CAPDEC_REQUIRE(!a.split.hi || (a.split.ld % 4 == 0 && (a.split.kind != KIND_BF16 || a.split.ld % 8 == 0)), CAPDEC_ERR_UNSUPPORTED,
                 "ingest: operand copies need a row pitch that is a multiple of 16 bytes");
  CAPDEC_REQUIRE(a.dtype != CAPDEC_DT_P24 || ((int64_t)a.L * a.D) % 8 == 0, CAPDEC_ERR_UNSUPPORTED, "ingest: p24 planes need L*D %% 8 == 0");
  if (a.B == 0) return CAPDEC_OK;
  if (a.layout == CAPDEC_LAYOUT_BDL) {
    const int pitch = a.L | 1;
    const size_t smem = (size_t)kTC * pitch * sizeof(float);
    CAPDEC_REQUIRE(smem <= 200 * 1024, CAPDEC_ERR_UNSUPPORTED, "ingest: %d regions do not fit the transpose slab", a.L);
    const dim3 grid(a.B, ceil_div(a.D, kTC));
#define CAPDEC_INGEST_T(DT)                                                                                         \
    {                                                                                                               \
      if (smem > 48 * 1024) CAPDEC_CHECK_CUDA(cudaFuncSetAttribute(ingest_transpose_kernel<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      ingest_transpose_kernel<DT><<<grid, 256, smem, s>>>(a, pitch);                                                \
    }
    if (a.dtype == CAPDEC_DT_F32) CAPDEC_INGEST_T(CAPDEC_DT_F32)
    else if (a.dtype == CAPDEC_DT_BF16) CAPDEC_INGEST_T(CAPDEC_DT_BF16)
    else CAPDEC_INGEST_T(CAPDEC_DT_F16)
#undef CAPDEC_INGEST_T
  } else {
    const size_t es = a.dtype == CAPDEC_DT_F32 ? 4 : 2;
    const int64_t img_bytes = (int64_t)ingest_source_image_bytes(a.layout, a.dtype, a.L, a.D);
    const int64_t skip = a.layout == CAPDEC_LAYOUT_CLS_BLD ? (int64_t)a.D * es : 0;
    CAPDEC_REQUIRE((((uintptr_t)a.src) & 15) == 0 && img_bytes % 8 == 0 && skip % 8 == 0 && (a.dtype != CAPDEC_DT_F32 || (img_bytes % 16 == 0 && skip % 16 == 0)),
                   CAPDEC_ERR_INVALID, "ingest: source must be 16-byte aligned with 16-byte image strides");
    const dim3 grid(a.B, ceil_div(a.D / 4, 256));
    switch (a.dtype) {
      case CAPDEC_DT_F32: ingest_rows_kernel<CAPDEC_DT_F32><<<grid, 256, 0, s>>>(a, img_bytes, skip); break;
      case CAPDEC_DT_BF16: ingest_rows_kernel<CAPDEC_DT_BF16><<<grid, 256, 0, s>>>(a, img_bytes, skip); break;
      case CAPDEC_DT_F16: ingest_rows_kernel<CAPDEC_DT_F16><<<grid, 256, 0, s>>>(a, img_bytes, skip); break;
      default: ingest_rows_kernel<CAPDEC_DT_P24><<<grid, 256, 0, s>>>(a, img_bytes, skip); break;
    }
  }
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

}  // namespace capdec
