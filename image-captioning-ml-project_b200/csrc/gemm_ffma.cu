// fp32 CUDA-core GEMM (the exact mode, CAPDEC_PREC_FP32):  C[M,N] = A[M,K] * W[N,K]^T + bias
// with the decode step's fused epilogues.  IEEE fp32 FFMA accumulation in a fixed k order, so
// results are reproducible run to run.  128x128x16 CTA tile, 8x8 register micro-tile, double-
// buffered shared memory with register prefetch.  The tensor-core modes live in gemm_tc.cu.
#include "common.cuh"
#include "gemm_epilogue.cuh"

namespace capdec {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

template <int EPI>
__global__ void __launch_bounds__(256, 2) gemm_ffma_kernel(const GemmArgs p) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Ws[2][BK][BN + PAD];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  // global->smem mapping: thread loads float4 (4 consecutive k) of rows lrow and lrow+64
  const int lrow = tid >> 2;
  const int lk = (tid & 3) * 4;
  const float* a_ptr[2];
  const float* w_ptr[2];
  bool a_ok[2], w_ok[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int m = m0 + lrow + 64 * i, n = n0 + lrow + 64 * i;
    a_ok[i] = m < p.M;
    w_ok[i] = n < p.N;
    a_ptr[i] = p.A + (int64_t)(a_ok[i] ? m : 0) * p.lda + lk;
    w_ptr[i] = p.W + (int64_t)(w_ok[i] ? n : 0) * p.ldw + lk;
  }

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rw[2];
  auto gload = [&](int k0) {
    const bool kin = (k0 + lk) < p.K;  // K % 4 == 0 guaranteed by launcher
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      ra[i] = (a_ok[i] && kin) ? *reinterpret_cast<const float4*>(a_ptr[i] + k0) : make_float4(0.f, 0.f, 0.f, 0.f);
      rw[i] = (w_ok[i] && kin) ? *reinterpret_cast<const float4*>(w_ptr[i] + k0) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = lrow + 64 * i;
      As[buf][lk + 0][r] = ra[i].x; As[buf][lk + 1][r] = ra[i].y; As[buf][lk + 2][r] = ra[i].z; As[buf][lk + 3][r] = ra[i].w;
      Ws[buf][lk + 0][r] = rw[i].x; Ws[buf][lk + 1][r] = rw[i].y; Ws[buf][lk + 2][r] = rw[i].z; Ws[buf][lk + 3][r] = rw[i].w;
    }
  };

  const int ktiles = (p.K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();

  for (int kt = 0; kt < ktiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < ktiles) gload((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Ws[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < ktiles) sstore(buf ^ 1);
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    epilogue4<EPI>(p, m, n0 + tx * 4, acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    epilogue4<EPI>(p, m, n0 + 64 + tx * 4, acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
  }
}

}  // namespace

int gemm_ffma(const GemmArgs& a, int epilogue, cudaStream_t s) {
  CAPDEC_REQUIRE(a.M >= 0 && a.N > 0 && a.K > 0, CAPDEC_ERR_INVALID, "gemm: bad shape M=%d N=%d K=%d", a.M, a.N, a.K);
  if (a.M == 0) return CAPDEC_OK;
  CAPDEC_REQUIRE(a.K % 4 == 0 && a.lda % 4 == 0 && a.ldw % 4 == 0, CAPDEC_ERR_UNSUPPORTED,
                 "gemm: K, lda, ldw must be multiples of 4 (K=%d lda=%lld ldw=%lld)", a.K, (long long)a.lda,
                 (long long)a.ldw);
  if (epilogue == EPI_LSTM || epilogue == EPI_AOA)
    CAPDEC_REQUIRE(a.N % 4 == 0, CAPDEC_ERR_UNSUPPORTED, "gemm: fused epilogue needs N %% 4 == 0 (N=%d)", a.N);
  CAPDEC_REQUIRE((((uintptr_t)a.A | (uintptr_t)a.W | (uintptr_t)a.bias) & 15) == 0, CAPDEC_ERR_INVALID,
                 "gemm: A/W/bias must be 16-byte aligned");
  if (epi_is_store_family(epilogue))
    CAPDEC_REQUIRE(((uintptr_t)a.C & 15) == 0 && (!a.C2 || ((uintptr_t)a.C2 & 15) == 0), CAPDEC_ERR_INVALID,
                   "gemm: C must be 16-byte aligned");
  if (epilogue == EPI_AOA)
    CAPDEC_REQUIRE(a.ldc % 2 == 0 && ((uintptr_t)a.C & 7) == 0 && (!a.C2 || (a.ldc2 % 2 == 0 && ((uintptr_t)a.C2 & 7) == 0)),
                   CAPDEC_ERR_INVALID, "gemm: AoA output must be 8-byte aligned");
  dim3 grid(ceil_div(a.N, BN), ceil_div(a.M, BM));
  switch (epilogue) {
    case EPI_STORE: gemm_ffma_kernel<EPI_STORE><<<grid, 256, 0, s>>>(a); break;
    case EPI_SIGMOID_TAIL: gemm_ffma_kernel<EPI_SIGMOID_TAIL><<<grid, 256, 0, s>>>(a); break;
    case EPI_LSTM: gemm_ffma_kernel<EPI_LSTM><<<grid, 256, 0, s>>>(a); break;
    case EPI_TANH: gemm_ffma_kernel<EPI_TANH><<<grid, 256, 0, s>>>(a); break;
    case EPI_AOA: gemm_ffma_kernel<EPI_AOA><<<grid, 256, 0, s>>>(a); break;
    case EPI_GELU: gemm_ffma_kernel<EPI_GELU><<<grid, 256, 0, s>>>(a); break;
    case EPI_GELU_TANH: gemm_ffma_kernel<EPI_GELU_TANH><<<grid, 256, 0, s>>>(a); break;
    default: CAPDEC_REQUIRE(false, CAPDEC_ERR_INVALID, "gemm: unknown epilogue %d", epilogue);
  }
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

}  // namespace capdec
