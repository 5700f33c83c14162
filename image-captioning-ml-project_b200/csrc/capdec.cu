// libcapdec C ABI + host-side decode programs (kernel sequencing on the caller's stream).
//
// Two decoder families share the stage kernels:
//   legacy  models/decoder.py::Decoder         attention(h_old) -> gate -> LSTMCell -> fc(h_new)
//   lstm    src/models/decoders.py::LSTMDecoder nn.LSTM([emb;prev_ctx]) -> attention(h_new) -> output_layer(ctx)
// Time-invariant projections of the region features (enc_att(enc), key_proj/value_proj(features)) are
// hoisted into a per-call prologue; the reference recomputes them every step (models/decoder.py:152,
// src/models/attention.py:77,172-174).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "handle.cuh"
#include "ingest.cuh"
#include "transformer.cuh"

namespace capdec {

static thread_local char g_err[1024] = "";
std::atomic<int64_t> g_launch_count{0};

int num_sms() {
  static std::atomic<int> cache[kMaxDevices];
  const int d = current_device();
  int n = cache[d].load(std::memory_order_relaxed);
  if (!n) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d);
    if (n <= 0) n = 148;
    cache[d].store(n, std::memory_order_relaxed);
  }
  return n;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int gemm(const capdec_handle* h, int precision, const GemmArgs& a, int epilogue, cudaStream_t s) {
  if (precision == CAPDEC_PREC_FP32) return gemm_ffma(a, epilogue, s);
  return gemm_tc(h, precision, a, epilogue, s);
}

namespace {

// dst[(G*j + g), col_off + c] (+)= src[j, c]   -- weight packing (concatenate / interleave gate rows)
__global__ void scatter_rows_kernel(float* dst, int64_t ld_dst, int G, int g, int64_t col_off, const float* src,
                                    int64_t ld_src, int nrows, int width, int accumulate) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)nrows * width) return;
  const int j = (int)(i / width), c = (int)(i - (int64_t)j * width);
  float* d = dst + ((int64_t)G * j + g) * ld_dst + col_off + c;
  const float v = src[(int64_t)j * ld_src + c];
  *d = accumulate ? (*d + v) : v;
}
int scatter_rows(float* dst, int64_t ld_dst, int G, int g, int64_t col_off, const float* src, int64_t ld_src,
                 int nrows, int width, bool accumulate, cudaStream_t s) {
  const int64_t n = (int64_t)nrows * width;
  scatter_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dst, ld_dst, G, g, col_off, src, ld_src, nrows, width,
                                                                  accumulate ? 1 : 0);
  CAPDEC_LAUNCH_CHECK();
  return CAPDEC_OK;
}

// adaptive attention elementwise pieces (src/models/attention.py:266-287)
__global__ void sentinel_pre_kernel(const float* gate, int64_t ld_g, const float* cell, int64_t ld_c, float* out,
                                    int64_t ld_o, int rows, int H) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * H) return;
  const int r = (int)(i / H), c = (int)(i - (int64_t)r * H);
  out[(int64_t)r * ld_o + c] = gate[(int64_t)r * ld_g + c] * tanhf(cell[(int64_t)r * ld_c + c]);  // gate already sigmoid
}
__global__ void __launch_bounds__(128) adaptive_mix_kernel(const float* ctx, const float* sent, const float* w,
                                                           float bias, float* out, int64_t ld_o, int H) {
  // beta = sigmoid(w . [ctx ; s] + b);  out = beta*ctx + (1-beta)*s      one CTA per row
  __shared__ float s_part[4];
  __shared__ float s_beta;
  const int r = blockIdx.x, tid = threadIdx.x;
  const float* c = ctx + (int64_t)r * H;
  const float* s = sent + (int64_t)r * H;
  float acc = 0.f;
  for (int i = tid; i < H; i += 128) acc += w[i] * c[i] + w[H + i] * s[i];
  acc = warp_sum(acc);
  if ((tid & 31) == 0) s_part[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) s_beta = sigmoidf_(s_part[0] + s_part[1] + s_part[2] + s_part[3] + bias);
  __syncthreads();
  const float beta = s_beta;
  for (int i = tid; i < H; i += 128) out[(int64_t)r * ld_o + i] = beta * c[i] + (1.f - beta) * s[i];
}

bool is_legacy(const capdec_handle* h) { return h->cfg.arch == CAPDEC_ARCH_LEGACY_SAT; }
bool is_transformer(const capdec_handle* h) { return h->cfg.arch == CAPDEC_ARCH_TRANSFORMER; }
bool is_gpt2(const capdec_handle* h) { return h->cfg.arch == CAPDEC_ARCH_GPT2; }
bool is_tf_family(const capdec_handle* h) { return is_transformer(h) || is_gpt2(h); }
bool base_is_mha(const capdec_handle* h) {
  const int a = h->cfg.attention;
  if (a == CAPDEC_ATT_MULTI_HEAD) return true;
  if (a == CAPDEC_ATT_SOFT) return false;
  return h->cfg.num_heads > 1;  // attention.py:229-230, 308-309
}
std::string base_prefix(const capdec_handle* h) {
  const int a = h->cfg.attention;
  return (a == CAPDEC_ATT_AOA || a == CAPDEC_ATT_ADAPTIVE) ? "attention.base_attention." : "attention.";
}

// tile set produced by capdec_ingest_features: what the decode streams instead of the caller's fp32 [B,L,D] features
struct TileSet {
  bool valid = false, p24 = false;
  float* f32 = nullptr;                                  // generic: dense fp32 [B,L,D]
  void* hi = nullptr; void* lo = nullptr; uint8_t* b8 = nullptr; float* mean = nullptr;   // p24: planes + GEMM lo operand + region mean
  size_t bytes = 0;
};

// ---- per-call workspace layout --------------------------------------------------------------------------
struct Session {
  int B = 0, L = 0, k = 1, R = 0, T = 0;
  // LSTM operands per layer: X[l] = [input_l | h_l], ld = in_l + H
  std::vector<float*> X, c, cnew, hnew;
  std::vector<int64_t> ldX;
  float* logits = nullptr;
  float* cand_lp = nullptr; int32_t* cand_idx = nullptr;
  int32_t* next_tok = nullptr; int32_t* src_row = nullptr; float* step_lp = nullptr;
  BeamState beam{};
  // legacy
  float* att1 = nullptr; float* meanb = nullptr; float* init = nullptr; float* hproj = nullptr;
  // lstm arch attention
  float* keyp = nullptr; float* valp = nullptr; float* qproj = nullptr; float* att_out = nullptr;
  float* ctx = nullptr; float* cat = nullptr; float* qq = nullptr; float* sgate = nullptr; float* spre = nullptr;
  float* sent = nullptr; float* base_ctx = nullptr;
  // transformer arch
  float* tx = nullptr; float* tqkv = nullptr; float* tsa = nullptr; float* ty = nullptr; float* tqc = nullptr;
  float* tca = nullptr; float* tff = nullptr; float* tmem = nullptr;
  std::vector<float*> tck, tcv, tcache_k, tcache_v;   // hoisted cross-attention K / V (dense [B,L,H] each), self-attention cache
  int32_t* anc[2] = {nullptr, nullptr};
  int anc_cur = -1;   // -1: identity (no reorder yet / greedy / sampling)
  float* txn = nullptr; float* gprefix = nullptr; int n_prefix = 0;   // GPT-2: pre-LN output, image prefix K == V
  // transformer family, tensor-core modes: hi/lo operand mirrors of the activations that feed GEMMs, written by the
  // kernels that produce them (embedding / LayerNorm / self-attention / GELU epilogue) instead of by a split pass
  SplitDst mx{}, msa{}, mff{};        // mirrors of tx (transformer) or txn (GPT-2), of tsa, of tff
  SplitDst mmem{};                    // transformer prologue: mirror of the projected memory (A operand of the 2 x layers hoisted K / V GEMMs)
  // fused vocabulary projection + log-softmax + top-k (EPI_TOPK): partial records instead of logits
  float* tk_part = nullptr; float* tk_lse = nullptr; int fuse_k = 0;
  float* samp_lse = nullptr;            // sampling on the tensor-core path: the logits GEMM's {max, sum exp} partials [R, tk_lse_pairs(V), 2]
  // legacy, tensor-core modes: the producers of the gate / vocabulary GEMM operands (attention context, state gather,
  // LSTM epilogue) write the hi/lo operand copies themselves, so no split pass runs inside the step loop
  void* xs_hi = nullptr; void* xs_lo = nullptr; void* hs_hi = nullptr; void* hs_lo = nullptr; bool presplit = false;
  // legacy, bf16 mode: the region tiles the attention kernel streams every step are kept in bf16 (half the bytes):
  // feats_h [B,L,D] is also the att1 GEMM's operand, att1_h [B,L,A] comes straight out of that GEMM's epilogue
  void* feats_h = nullptr; void* att1_h = nullptr;
  void* feats_l = nullptr;               // split modes: lo copy of the features (hi in feats_h); both feed the hoisted enc_att GEMM
  // bf16x3 mode, "p24" tiles (common.cuh): feats_h / att1_h are the 16-bit planes (feats_h still the GEMM's hi operand),
  // these the byte planes; the attention kernel streams 3 bytes per element instead of 4
  uint8_t* feats_b8 = nullptr; uint8_t* att1_b8 = nullptr;
  int tk_ntotal = 0;                    // N of the fused vocabulary GEMM (vocab, or vocab padded + the legacy [dec_att|f_beta] tail)
  bool hproj_ready = false;             // legacy: S.hproj already holds the projections of the CURRENT hidden state (pre-reorder rows)
  const int32_t* row_src = nullptr;     // back-pointers of the last commit (nullptr = identity)
  int64_t logits_ld = 0;                // row stride of S.logits (0 = vocab_size).  The workspace buffer pads rows to a multiple of
                                        // 4 floats so the GEMM epilogue stores whole float4s (V = 50257: the scalar path made the
                                        // GPT-2 logits GEMM store-bound); capdec_forward_tokens writes [R, t, V] blocks
  // teacher-forced pass (capdec_forward_tokens), transformer family: key positions whose forced token is pad are masked
  const int32_t* key_tok = nullptr; int64_t ld_key_tok = 0; int key_pad = -1;
  // encoder hand-off (capdec_ingest_features): the p24 planes / lo operand / region mean already exist in a caller-owned
  // tile set, so the prologue neither allocates nor recomputes them
  TileSet ext{};
};

enum Mode { MODE_BEAM, MODE_GREEDY, MODE_SAMPLE, MODE_TEACHER, MODE_ATTENTION };

// does this handle keep its region tiles as p24 planes (legacy decoder, BF16X3 mode)?  k-independent part of the test
bool tiles_p24(const capdec_handle* h) {
  const capdec_config& c = h->cfg;
  const int H = c.hidden_dim, E = c.embed_dim, D = c.feature_dim, A = c.attention_dim;
  return is_legacy(h) && c.precision == CAPDEC_PREC_BF16X3 && (E + D + H) % 8 == 0 && H % 8 == 0 && E % 4 == 0 && A % 16 == 0 &&
         D % 16 == 0 && !ab_switch("CAPDEC_NO_PRESPLIT") && !ab_switch("CAPDEC_NO_P24_TILES");
}
// carve-up of a tile set for B images (base == nullptr: sizes only)
TileSet tiles_layout(const capdec_handle* h, void* base, int B, int L) {
  const size_t n = (size_t)B * L * h->cfg.feature_dim;
  Arena ar(base, (size_t)-1);
  TileSet t{};
  t.valid = true;
  if (tiles_p24(h)) {
    t.p24 = true;
    t.hi = ar.take<char>(n * 2);
    t.b8 = ar.take<uint8_t>(n);
    t.lo = ar.take<char>(n * 2);
    t.mean = ar.take<float>((size_t)B * h->cfg.feature_dim);
  } else {
    t.f32 = ar.take<float>(n);
  }
  t.bytes = align_up(ar.off, 256);
  return t;
}

int carve(const capdec_handle* h, Arena& ar, Session& S, int B, int L, int k, int T, Mode mode) {
  const capdec_config& c = h->cfg;
  const int H = c.hidden_dim, E = c.embed_dim, D = c.feature_dim, A = c.attention_dim, V = c.vocab_size;
  S.B = B; S.L = L; S.k = k; S.R = B * k; S.T = T;
  const size_t R = (size_t)S.R;
  const int layers = c.num_layers;
  // beam / greedy on the tensor-core path: the vocabulary GEMM's epilogue keeps only per-tile log-sum-exp partials
  // and top-k candidates, so no [R,V] logits buffer exists (sampling and the teacher-forced forward need full rows)
  const int want_k = mode == MODE_BEAM ? 2 * k : mode == MODE_GREEDY ? 1 : 0;
  S.fuse_k = 0;
  if (want_k > 0 && c.precision != CAPDEC_PREC_FP32 && tk_supported(V, want_k) && !ab_switch("CAPDEC_NO_FUSED_TOPK"))
    S.fuse_k = want_k;
  auto take_logits = [&]() {
    if (S.fuse_k > 0) {
      S.tk_ntotal = (is_legacy(h) && h->w_vocab_cat) ? h->vocab_cat_n : V;
      S.tk_part = ar.take<float>(R * tk_records(S.R, S.tk_ntotal) * tk_stride(S.fuse_k));
      S.tk_lse = ar.take<float>(R * tk_lse_pairs(V) * 2);
    }
    else {
      S.logits_ld = (V + 3) & ~3; S.logits = ar.take<float>(R * S.logits_ld);
      if (mode == MODE_SAMPLE && c.precision != CAPDEC_PREC_FP32) S.samp_lse = ar.take<float>(R * tk_lse_pairs(V) * 2);
    }
  };
  if (is_tf_family(h)) {
    const DevTensor* f1 = h->find(is_gpt2(h) ? "model.transformer.h.0.mlp.c_fc.weight" : "transformer_decoder.layers.0.linear1.weight");
    const size_t F = f1 ? (size_t)f1->shape[0] : (size_t)4 * H;
    if (is_gpt2(h)) {
      const DevTensor* ip = h->find("image_to_prefix.weight");
      S.n_prefix = ip ? (int)(ip->shape[0] / H) : 10;
      S.txn = ar.take<float>(R * H);
      S.gprefix = ar.take<float>((size_t)B * S.n_prefix * H);
    }
    S.mx = S.msa = S.mff = S.mmem = SplitDst{};
    if (c.precision != CAPDEC_PREC_FP32 && H % 8 == 0 && F % 8 == 0 && !ab_switch("CAPDEC_NO_PRESPLIT")) {
      const int kind = tc_kind(c.precision);
      const size_t es = kind == KIND_BF16 ? 2 : 4;
      const bool lo = tc_terms(c.precision) == 3;
      auto mk = [&](size_t cols) {
        SplitDst d{};
        d.hi = ar.take<char>(R * cols * es); d.lo = lo ? ar.take<char>(R * cols * es) : nullptr; d.ld = (int64_t)cols; d.kind = kind;
        return d;
      };
      S.mx = mk(H); S.msa = mk(H); S.mff = mk(F);
      if (!is_gpt2(h)) {   // written once by the visual_projection epilogue instead of being re-split by each of the 2 x layers GEMMs
        S.mmem.hi = ar.take<char>((size_t)B * L * H * es); S.mmem.lo = lo ? ar.take<char>((size_t)B * L * H * es) : nullptr;
        S.mmem.ld = H; S.mmem.kind = kind;
      }
    }
    S.tx = ar.take<float>(R * H); S.tqkv = ar.take<float>(R * 3 * H); S.tsa = ar.take<float>(R * H);
    S.ty = ar.take<float>(R * H); S.tqc = ar.take<float>(R * H); S.tca = ar.take<float>(R * H);
    S.tff = ar.take<float>(R * F);
    if (!is_gpt2(h)) S.tmem = ar.take<float>((size_t)B * L * H);
    S.tck.resize(layers); S.tcv.resize(layers); S.tcache_k.resize(layers); S.tcache_v.resize(layers);
    for (int l = 0; l < layers; ++l) {
      if (!is_gpt2(h)) { S.tck[l] = ar.take<float>((size_t)B * L * H); S.tcv[l] = ar.take<float>((size_t)B * L * H); }
      S.tcache_k[l] = ar.take<float>(R * T * H);
      S.tcache_v[l] = ar.take<float>(R * T * H);
    }
    S.anc[0] = ar.take<int32_t>(R * T); S.anc[1] = ar.take<int32_t>(R * T);
    take_logits();
    S.next_tok = ar.take<int32_t>(R); S.src_row = ar.take<int32_t>(R); S.step_lp = ar.take<float>(R);
    S.cand_lp = ar.take<float>(R * 2 * k); S.cand_idx = ar.take<int32_t>(R * 2 * k);
    if (mode == MODE_BEAM) {
      for (int i = 0; i < 2; ++i) { S.beam.run_seq[i] = ar.take<int32_t>(R * T); S.beam.fin_seq[i] = ar.take<int32_t>(R * T); }
      S.beam.run_score = ar.take<float>(R); S.beam.fin_score = ar.take<float>(R); S.beam.fin_len = ar.take<int32_t>(R);
      S.beam.fin_flag = ar.take<uint8_t>(R); S.beam.unsatisfied = ar.take<uint8_t>(B);
    }
    return CAPDEC_OK;
  }
  if (mode != MODE_ATTENTION) {
    S.X.resize(layers); S.c.resize(layers); S.cnew.resize(layers); S.hnew.resize(layers); S.ldX.resize(layers);
    for (int l = 0; l < layers; ++l) {
      const int in = l == 0 ? (is_legacy(h) ? E + D : E + H) : H;
      S.ldX[l] = in + H;
      S.X[l] = ar.take<float>(R * S.ldX[l]);
      S.c[l] = ar.take<float>(R * H);
      S.cnew[l] = ar.take<float>(R * H);
      S.hnew[l] = ar.take<float>(R * (c.attention == CAPDEC_ATT_ADAPTIVE && l == layers - 1 && !is_legacy(h) ? 2 * H : H));
    }
    if (mode != MODE_TEACHER) take_logits();
    S.next_tok = ar.take<int32_t>(R);
    S.src_row = ar.take<int32_t>(R);
    S.step_lp = ar.take<float>(R);
    if (mode == MODE_BEAM) {
      S.cand_lp = ar.take<float>(R * 2 * k);
      S.cand_idx = ar.take<int32_t>(R * 2 * k);
      for (int i = 0; i < 2; ++i) {
        S.beam.run_seq[i] = ar.take<int32_t>(R * T);
        S.beam.fin_seq[i] = ar.take<int32_t>(R * T);
      }
      S.beam.run_score = ar.take<float>(R);
      S.beam.fin_score = ar.take<float>(R);
      S.beam.fin_len = ar.take<int32_t>(R);
      S.beam.fin_flag = ar.take<uint8_t>(R);
      S.beam.unsatisfied = ar.take<uint8_t>(B);
    } else if (mode == MODE_GREEDY) {
      S.cand_lp = ar.take<float>(R);
      S.cand_idx = ar.take<int32_t>(R);
    }
  }
  if (is_legacy(h)) {
    S.presplit = false;
    if (c.precision != CAPDEC_PREC_FP32 && mode != MODE_TEACHER && mode != MODE_ATTENTION && (E + D + H) % 8 == 0 && H % 8 == 0 &&
        E % 4 == 0 && !ab_switch("CAPDEC_NO_PRESPLIT")) {
      const size_t es = tc_kind(c.precision) == KIND_BF16 ? 2 : 4;
      const bool lo = tc_terms(c.precision) == 3;
      S.xs_hi = ar.take<char>(R * (E + D + H) * es);
      S.xs_lo = lo ? ar.take<char>(R * (E + D + H) * es) : nullptr;
      S.hs_hi = ar.take<char>(R * H * es);
      S.hs_lo = lo ? ar.take<char>(R * H * es) : nullptr;
      S.presplit = true;
      if (c.precision == CAPDEC_PREC_BF16 && A % 8 == 0 && D % 8 == 0 && !ab_switch("CAPDEC_NO_BF16_TILES") &&
          additive_attention_stream_supports(A, D, L, k, true)) {
        S.feats_h = ar.take<char>((size_t)B * L * D * 2);
        S.att1_h = ar.take<char>((size_t)B * L * A * 2);
      } else if (D % 8 == 0) {
        // operand copies of the features for the hoisted enc_att GEMM, written by the same pass that takes the region mean
        const bool p24 = c.precision == CAPDEC_PREC_BF16X3 && !ab_switch("CAPDEC_NO_P24_TILES") && additive_attention_stream_supports(A, D, L, k, 2);
        if (S.ext.valid && S.ext.p24) {
          if (!p24) return CAPDEC_ERR_UNSUPPORTED;   // the planes are all there is: no fp32 features to fall back to
          S.feats_h = S.ext.hi; S.feats_l = S.ext.lo; S.feats_b8 = S.ext.b8;
        } else {
          S.feats_h = ar.take<char>((size_t)B * L * D * es);
          S.feats_l = lo ? ar.take<char>((size_t)B * L * D * es) : nullptr;
          if (p24) S.feats_b8 = ar.take<uint8_t>((size_t)B * L * D);
        }
        if (p24) {
          S.att1_h = ar.take<char>((size_t)B * L * A * 2);
          S.att1_b8 = ar.take<uint8_t>((size_t)B * L * A);
        }
      }
    }
    if (S.ext.valid && S.ext.p24 && !S.feats_b8) return CAPDEC_ERR_UNSUPPORTED;
    if (!S.att1_h) S.att1 = ar.take<float>((size_t)B * L * A);
    S.meanb = (S.ext.valid && S.ext.p24) ? S.ext.mean : ar.take<float>((size_t)B * D);
    S.init = ar.take<float>((size_t)B * 2 * H);
    S.hproj = ar.take<float>(R * (A + D));
  } else {
    S.keyp = ar.take<float>((size_t)B * L * H);
    if (base_is_mha(h)) {
      S.valp = ar.take<float>((size_t)B * L * H);
      S.att_out = ar.take<float>(R * H);
    }
    S.qproj = ar.take<float>(R * H);
    S.ctx = ar.take<float>(R * H);
    if (mode != MODE_ATTENTION) S.init = ar.take<float>((size_t)B * 2 * H * layers);
    if (c.attention == CAPDEC_ATT_AOA) S.cat = ar.take<float>(R * 2 * H);
    if (c.attention == CAPDEC_ATT_ADAPTIVE) {
      S.qq = ar.take<float>(R * 2 * H);
      S.sgate = ar.take<float>(R * H);
      S.spre = ar.take<float>(R * H);
      S.sent = ar.take<float>(R * H);
      S.base_ctx = ar.take<float>(R * H);
    }
  }
  return CAPDEC_OK;
}

// ---- hoisted per-call projections ---------------------------------------------------------------------------
// a_pre: A's hi/lo mirror written by A's producer (hi == nullptr: none, the GEMM splits A itself);
// c_out: mirror to fill for the consumer of C (store-family and LSTM epilogues)
int linear(const capdec_handle* h, const float* A, int64_t lda, const std::string& name, float* C, int64_t ldc, int M,
           int epi, cudaStream_t s, float* C2 = nullptr, int64_t ldc2 = 0, const SplitDst* a_pre = nullptr,
           const SplitDst* c_out = nullptr) {
  const DevTensor* w = h->find(name + ".weight");
  const DevTensor* b = h->find(name + ".bias");
  CAPDEC_REQUIRE(w && w->shape.size() == 2, CAPDEC_ERR_STATE, "weight %s.weight missing", name.c_str());
  GemmArgs g{};
  g.A = A; g.lda = lda; g.W = w->p; g.ldw = w->shape[1]; g.bias = b ? b->p : nullptr;
  g.C = C; g.ldc = ldc; g.M = M; g.N = (int)w->shape[0]; g.K = (int)w->shape[1]; g.C2 = C2; g.ldc2 = ldc2;
  if (a_pre && a_pre->hi) { g.A_hi = a_pre->hi; g.A_lo = a_pre->lo; g.ld_as = a_pre->ld; }
  if (c_out && c_out->hi) g.c_split = *c_out;
  return gemm(h, h->cfg.precision, g, epi, s);
}

// vocabulary projection: logits [rows,V] into `logits`, or (logits == nullptr, beam/greedy on the tensor-core path) the
// fused EPI_TOPK partial records into S.tk_part
int vocab_project(const capdec_handle* h, Session& S, const float* A, int64_t lda, const float* W, const float* bias,
                  int rows, float* logits, int64_t ld_logits, cudaStream_t s) {
  const capdec_config& c = h->cfg;
  GemmArgs g{};
  g.A = A; g.lda = lda; g.W = W; g.ldw = c.hidden_dim; g.bias = bias; g.M = rows; g.N = c.vocab_size; g.K = c.hidden_dim;
  if (S.presplit && is_legacy(h)) { g.A_hi = S.hs_hi; g.A_lo = S.hs_lo; g.ld_as = c.hidden_dim; }   // written by the LSTM epilogue
  if (is_tf_family(h) && S.mx.hi) { g.A_hi = S.mx.hi; g.A_lo = S.mx.lo; g.ld_as = S.mx.ld; }        // written by the last LayerNorm
  if (logits == nullptr) {
    CAPDEC_REQUIRE(S.fuse_k > 0 && S.tk_part, CAPDEC_ERR_STATE, "vocab_project: no logits buffer and no fused top-k buffer");
    g.tk_part = S.tk_part; g.tk_k = S.fuse_k; g.tk_lse = S.tk_lse; g.tk_vocab = c.vocab_size;
    if (S.tk_ntotal != c.vocab_size) {
      // legacy: the same GEMM also projects the new hidden state for the next step's attention (models/decoder.py:153,160)
      g.W = h->w_vocab_cat; g.bias = h->b_vocab_cat; g.N = S.tk_ntotal;
      g.C = S.hproj; g.ldc = c.attention_dim + c.feature_dim; g.n_split = c.attention_dim;
      S.hproj_ready = true;
    }
    return gemm(h, c.precision, g, EPI_TOPK, s);
  }
  g.C = logits; g.ldc = ld_logits;
  if (S.samp_lse && logits == S.logits) { g.tk_lse = S.samp_lse; g.tk_vocab = c.vocab_size; }   // partials for the sampler
  return gemm(h, c.precision, g, EPI_STORE, s);
}

// per-row sorted top-`topk` log-probs of the step's vocabulary distribution
int select_topk(const capdec_handle* h, Session& S, int topk, float* out_lp, int32_t* out_idx, cudaStream_t s) {
  const int V = h->cfg.vocab_size;
  if (S.fuse_k > 0) return topk_merge(S.tk_part, S.tk_lse, S.R, V, S.tk_ntotal, S.fuse_k, topk, out_lp, out_idx, nullptr, s);
  return lse_topk(S.logits, S.logits_ld ? S.logits_ld : V, S.R, V, topk, out_lp, out_idx, nullptr, s);
}

int prologue_legacy(const capdec_handle* h, Session& S, const float* feats, bool expand, cudaStream_t s) {
  StageScope sc(h, STAGE_PROLOGUE, s);
  const capdec_config& c = h->cfg;
  const int H = c.hidden_dim, E = c.embed_dim, D = c.feature_dim, A = c.attention_dim;
  // att1 = enc_att(enc)  (models/decoder.py:152, hoisted)
  const int kind = tc_kind(c.precision);
  if (S.feats_h) {
    // one pass over the features: region mean (for h0 / c0, :137-139) + the hi/lo operand copies.  In the bf16 mode the
    // hi copy is also what the attention kernel streams every step and the GEMM's epilogue emits att1 directly as bf16.
    const SplitDst fsplit{S.feats_h, S.feats_l, D, kind, S.feats_b8};
    if (!(S.ext.valid && S.ext.p24)) CAPDEC_RETURN_IF(mean_regions_split(feats, S.B, S.L, D, S.meanb, fsplit, s));
    if (S.att1_h) {
      const SplitDst c_out{S.att1_h, nullptr, A, KIND_BF16, S.att1_b8};
      CAPDEC_RETURN_IF(linear(h, feats, D, "enc_att", nullptr, A, S.B * S.L, EPI_STORE, s, nullptr, 0, &fsplit, &c_out));
    } else {
      CAPDEC_RETURN_IF(linear(h, feats, D, "enc_att", S.att1, A, S.B * S.L, EPI_STORE, s, nullptr, 0, &fsplit));
    }
  } else {
    CAPDEC_RETURN_IF(linear(h, feats, D, "enc_att", S.att1, A, S.B * S.L, EPI_STORE, s));
    // h0, c0 = h_lin(mean), c_lin(mean)  (:137-139)
    CAPDEC_RETURN_IF(mean_regions(feats, S.B, S.L, D, S.meanb, s));
  }
  GemmArgs g{};
  g.A = S.meanb; g.lda = D; g.W = h->w_init; g.ldw = D; g.bias = h->b_init; g.C = S.init; g.ldc = 2 * H;
  g.M = S.B; g.N = 2 * H; g.K = D;
  CAPDEC_RETURN_IF(gemm(h, c.precision, g, EPI_STORE, s));
  if (expand) {
    CAPDEC_RETURN_IF(expand_rows(S.init, 2 * H, S.X[0] + E + D, S.ldX[0], S.R, S.k, H, s));
    CAPDEC_RETURN_IF(expand_rows(S.init + H, 2 * H, S.c[0], H, S.R, S.k, H, s));
    // one-off: bring the split mirror of X in line (h0); from here on its producers maintain it
    if (S.presplit)   // (the emb / ctx columns are rewritten by their producers before the first gate GEMM)
      CAPDEC_RETURN_IF(tc_split(c.precision, S.X[0], S.ldX[0], S.R, (int)S.ldX[0], S.xs_hi, S.xs_lo, s));
  }
  return CAPDEC_OK;
}

int prologue_attention(const capdec_handle* h, Session& S, const float* feats, cudaStream_t s) {
  const int H = h->cfg.hidden_dim;
  const std::string p = base_prefix(h);
  CAPDEC_RETURN_IF(linear(h, feats, H, p + "key_proj", S.keyp, H, S.B * S.L, EPI_STORE, s));
  if (base_is_mha(h)) CAPDEC_RETURN_IF(linear(h, feats, H, p + "value_proj", S.valp, H, S.B * S.L, EPI_STORE, s));
  return CAPDEC_OK;
}

int prologue_lstm(const capdec_handle* h, Session& S, const float* feats, const float* pooled, cudaStream_t s) {
  StageScope sc(h, STAGE_PROLOGUE, s);
  const capdec_config& c = h->cfg;
  const int H = c.hidden_dim, E = c.embed_dim, layers = c.num_layers;
  CAPDEC_RETURN_IF(prologue_attention(h, S, feats, s));
  // _init_hidden_states (decoders.py:122-135): [init_h ; init_c](pooled) -> [B, 2*layers*H]
  GemmArgs g{};
  g.A = pooled; g.lda = H; g.W = h->w_init; g.ldw = H; g.bias = h->b_init; g.C = S.init; g.ldc = 2 * layers * H;
  g.M = S.B; g.N = 2 * layers * H; g.K = H;
  CAPDEC_RETURN_IF(gemm(h, c.precision, g, EPI_STORE, s));
  for (int l = 0; l < layers; ++l) {
    const int in = l == 0 ? E + H : H;
    CAPDEC_RETURN_IF(expand_rows(S.init + (size_t)l * H, 2 * layers * H, S.X[l] + in, S.ldX[l], S.R, S.k, H, s));
    CAPDEC_RETURN_IF(expand_rows(S.init + (size_t)(layers + l) * H, 2 * layers * H, S.c[l], H, S.R, S.k, H, s));
  }
  // prev_ctx = 0 (decoders.py:265-266): both the attention output buffer and its slot in the LSTM operand
  CAPDEC_CHECK_CUDA(cudaMemsetAsync(S.ctx, 0, (size_t)S.R * H * sizeof(float), s));
  CAPDEC_CHECK_CUDA(cudaMemset2DAsync(S.X[0] + E, S.ldX[0] * sizeof(float), 0, (size_t)H * sizeof(float), S.R, s));
  return CAPDEC_OK;
}

// ---- attention for the lstm arch: q [rows,H] -> ctx [rows,H] (S.ctx), alpha ------------------------------------
int run_attention(const capdec_handle* h, Session& S, const float* feats, const uint8_t* mask, const float* q,
                  int64_t ld_q, const float* memory, int64_t ld_mem, const float* cell, int64_t ld_cell, int images,
                  float* ctx_out, float* alpha, int64_t ld_alpha, cudaStream_t s) {
  const capdec_config& c = h->cfg;
  const int H = c.hidden_dim, k = S.k, rows = images * k;
  const std::string p = base_prefix(h);
  float* base_dst = ctx_out;
  if (c.attention == CAPDEC_ATT_AOA) base_dst = S.cat;          // cat[:, 0:H], ld 2H
  if (c.attention == CAPDEC_ATT_ADAPTIVE) base_dst = S.base_ctx;
  const int64_t ld_base = c.attention == CAPDEC_ATT_AOA ? 2 * H : H;

  if (c.attention == CAPDEC_ATT_ADAPTIVE) {
    StageScope sc(h, STAGE_SMALL_GEMM, s);
    // sentinel (attention.py:266-272): s = sentinel_proj(sigmoid(W[q;mem]) * tanh(cell))
    CAPDEC_REQUIRE(memory && cell, CAPDEC_ERR_INVALID, "AdaptiveAttention requires memory_state and cell_state");
    GatherArgs ga{};
    ga.rows = rows; ga.n_state = 2; ga.pos = -1;
    ga.state_src[0] = q; ga.ld_src[0] = ld_q; ga.state_dst[0] = S.qq; ga.ld_dst[0] = 2 * H; ga.width[0] = H;
    ga.state_src[1] = memory; ga.ld_src[1] = ld_mem; ga.state_dst[1] = S.qq + H; ga.ld_dst[1] = 2 * H; ga.width[1] = H;
    if (!(q == S.qq && memory == S.qq + H)) CAPDEC_RETURN_IF(gather_rows(ga, s));
    const DevTensor* w = h->find("attention.sentinel_gate.weight");
    GemmArgs g{};
    g.A = S.qq; g.lda = 2 * H; g.W = w->p; g.ldw = 2 * H; g.bias = h->W("attention.sentinel_gate.bias");
    g.C = S.sgate; g.ldc = H; g.M = rows; g.N = H; g.K = 2 * H; g.n_split = 0;
    CAPDEC_RETURN_IF(gemm(h, c.precision, g, EPI_SIGMOID_TAIL, s));
    const int64_t n = (int64_t)rows * H;
    sentinel_pre_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(S.sgate, H, cell, ld_cell, S.spre, H, rows, H);
    CAPDEC_LAUNCH_CHECK();
    CAPDEC_RETURN_IF(linear(h, S.spre, H, "attention.sentinel_proj", S.sent, H, rows, EPI_STORE, s));
  }

  { StageScope sc(h, STAGE_SMALL_GEMM, s); CAPDEC_RETURN_IF(linear(h, q, ld_q, p + "query_proj", S.qproj, H, rows, EPI_STORE, s)); }
  if (base_is_mha(h)) {
    MhaArgs m{};
    m.q = S.qproj; m.ld_q = H; m.kproj = S.keyp; m.vproj = S.valp; m.ld_kv = H; m.mask = mask;
    m.denom = (float)((double)c.temperature * sqrt((double)(H / c.num_heads)));
    m.out = S.att_out; m.ld_out = H; m.alpha = alpha; m.ld_alpha = ld_alpha;
    m.B = images; m.L = S.L; m.H = H; m.heads = c.num_heads; m.k = k;
    { StageScope sc(h, STAGE_ATTENTION, s); CAPDEC_RETURN_IF(mha_attention(m, s)); }
    { StageScope sc(h, STAGE_SMALL_GEMM, s); CAPDEC_RETURN_IF(linear(h, S.att_out, H, p + "output_proj", base_dst, ld_base, rows, EPI_STORE, s)); }
  } else {
    AddAttnArgs a{};
    a.att1 = S.keyp; a.att2 = S.qproj; a.ld_att2 = H; a.w = h->W(p + "energy.weight");
    a.w_bias = h->energy_bias; a.temperature = c.temperature; a.mask = mask; a.feats = feats;
    a.gate = nullptr; a.ctx = base_dst; a.ld_ctx = ld_base; a.alpha = alpha; a.ld_alpha = ld_alpha;
    a.B = images; a.L = S.L; a.A = H; a.D = H; a.k = k;
    // exact tanhf in the fp32 mode; the MUFU form (1e-6) in the tensor-core modes, whose GEMM noise is 100x larger
    const int act = (c.precision == CAPDEC_PREC_FP32 || ab_switch("CAPDEC_EXACT_TANH")) ? ACT_TANH : ACT_TANH_FAST;
    { StageScope sc(h, STAGE_ATTENTION, s); CAPDEC_RETURN_IF(additive_attention(a, act, s)); }
  }

  StageScope sc_tail(h, STAGE_SMALL_GEMM, s);
  if (c.attention == CAPDEC_ATT_AOA) {
    // attention.py:343-353: cat = [ctx ; query_proj(q)];  out = tanh(W_i cat) * sigmoid(W_g cat)
    CAPDEC_RETURN_IF(linear(h, q, ld_q, "attention.query_proj", S.cat + H, 2 * H, rows, EPI_STORE, s));
    GemmArgs g{};
    g.A = S.cat; g.lda = 2 * H; g.W = h->w_aoa; g.ldw = 2 * H; g.bias = h->b_aoa; g.C = ctx_out; g.ldc = H;
    g.M = rows; g.N = 2 * H; g.K = 2 * H;
    CAPDEC_RETURN_IF(gemm(h, c.precision, g, EPI_AOA, s));
  } else if (c.attention == CAPDEC_ATT_ADAPTIVE) {
    adaptive_mix_kernel<<<rows, 128, 0, s>>>(S.base_ctx, S.sent, h->W("attention.adaptive_weight.weight"),
                                             h->adaptive_bias, ctx_out, H, H);
    CAPDEC_LAUNCH_CHECK();
  }
  return CAPDEC_OK;
}

// ---- one decode step: state in X/c -> logits (+ hnew/cnew) -----------------------------------------------------
int step_legacy(const capdec_handle* h, Session& S, const float* feats, int images, float* logits, int64_t ld_logits,
                float* alpha, int64_t ld_alpha, cudaStream_t s) {
  const capdec_config& c = h->cfg;
  const int H = c.hidden_dim, E = c.embed_dim, D = c.feature_dim, A = c.attention_dim, V = c.vocab_size;
  const int rows = images * S.k;
  // [dec_att(h) | sigmoid(f_beta(h))]   (models/decoder.py:153,160)
  GemmArgs g{};
  g.A = S.X[0] + E + D; g.lda = S.ldX[0]; g.W = h->w_hproj; g.ldw = H; g.bias = h->b_hproj;
  g.C = S.hproj; g.ldc = A + D; g.M = rows; g.N = A + D; g.K = H; g.n_split = A;
  // (steps >= 1 of the fused path: the previous step's vocabulary GEMM already produced these rows, indexed before the
  //  beam reorder, so the attention kernel applies the back-pointer instead)
  const bool reuse = S.hproj_ready;
  if (!reuse) { StageScope sc(h, STAGE_SMALL_GEMM, s); CAPDEC_RETURN_IF(gemm(h, c.precision, g, EPI_SIGMOID_TAIL, s)); }
  S.hproj_ready = false;
  // scores -> softmax -> gated context, written straight into the LSTM operand (:154-161)
  AddAttnArgs a{};
  a.row_src = reuse ? S.row_src : nullptr;
  a.att1 = S.att1; a.att2 = S.hproj;
  if (S.att1_h) {
    a.att1 = reinterpret_cast<const float*>(S.att1_h); a.tile_fmt = S.att1_b8 ? 2 : 1;
    a.att1_b8 = S.att1_b8; a.feats_b8 = S.feats_b8;
  }
  a.ld_att2 = A + D; a.w = h->W("att.weight"); a.w_bias = h->energy_bias;
  a.temperature = 1.f; a.mask = nullptr; a.feats = S.att1_h ? reinterpret_cast<const float*>(S.feats_h) : feats;
  a.gate = S.hproj + A; a.ld_gate = A + D;
  a.ctx = S.X[0] + E; a.ld_ctx = S.ldX[0]; a.alpha = alpha; a.ld_alpha = ld_alpha;
  a.B = images; a.L = S.L; a.A = A; a.D = D; a.k = S.k;
  const int kind = tc_kind(c.precision);
  if (S.presplit) {   // the gated context goes straight into the gate GEMM's split operand; nobody reads the fp32 copy
    a.ctx = nullptr;
    a.ctx_split = SplitDst{S.xs_hi, S.xs_lo, S.ldX[0], kind};
    a.ctx_split_col = E;
  }
  { StageScope sc(h, STAGE_ATTENTION, s); CAPDEC_RETURN_IF(additive_attention(a, ACT_RELU, s)); }
  // LSTMCell([emb ; ctx], (h, c)) with the cell update fused into the gate GEMM (:168)
  GemmArgs l{};
  l.A = S.X[0]; l.lda = S.ldX[0]; l.W = h->w_gates[0]; l.ldw = E + D + H; l.bias = h->b_gates[0];
  l.C = S.hnew[0]; l.ldc = H; l.M = rows; l.N = 4 * H; l.K = E + D + H;
  l.c_in = S.c[0]; l.ldcin = H; l.c_out = S.cnew[0]; l.ldcout = H;
  if (S.presplit) {
    l.A_hi = S.xs_hi; l.A_lo = S.xs_lo; l.ld_as = S.ldX[0];
    l.c_split = SplitDst{S.hs_hi, S.hs_lo, H, kind};   // new h as the vocabulary GEMM's operand
    if (h->emb_gates) {
      // the embedding columns of [emb | ctx | h] are served by the per-token table: K shrinks from E+D+H to D+H
      const size_t es = kind == KIND_BF16 ? 2 : 4;
      l.A = S.X[0] + E; l.W = h->w_gates[0] + E; l.K = D + H;
      l.A_hi = (const char*)S.xs_hi + (size_t)E * es; l.A_lo = S.xs_lo ? (const char*)S.xs_lo + (size_t)E * es : nullptr;
      l.row_table = h->emb_gates; l.row_index = S.next_tok; l.ld_table = 4 * H;
    }
  }
  { StageScope sc(h, STAGE_GATE_GEMM, s); CAPDEC_RETURN_IF(gemm(h, c.precision, l, EPI_LSTM, s)); }
  // fc(h)  (:171; dropout is the identity in eval)
  { StageScope sc(h, STAGE_VOCAB_GEMM, s);
    CAPDEC_RETURN_IF(vocab_project(h, S, S.hnew[0], H, h->W("fc.weight"), h->W("fc.bias"), rows, logits, ld_logits, s)); }
  (void)V;
  return CAPDEC_OK;
}

int step_lstm(const capdec_handle* h, Session& S, const float* feats, const uint8_t* mask, float* alpha,
              int64_t ld_alpha, cudaStream_t s) {
  const capdec_config& c = h->cfg;
  const int H = c.hidden_dim, E = c.embed_dim, layers = c.num_layers;
  const int rows = S.R;
  const bool adaptive = c.attention == CAPDEC_ATT_ADAPTIVE;
  // nn.LSTM single step (decoders.py:281): stacked cells, layer l+1 input = new h of layer l
  for (int l = 0; l < layers; ++l) {
    const int in = l == 0 ? E + H : H;
    const bool top = l == layers - 1;
    GemmArgs g{};
    g.A = S.X[l]; g.lda = S.ldX[l]; g.W = h->w_gates[l]; g.ldw = in + H; g.bias = h->b_gates[l];
    g.C = S.hnew[l]; g.ldc = (top && adaptive) ? 2 * H : H; g.M = rows; g.N = 4 * H; g.K = in + H;
    g.c_in = S.c[l]; g.ldcin = H; g.c_out = S.cnew[l]; g.ldcout = H;
    if (!top) { g.C2 = S.X[l + 1]; g.ldc2 = S.ldX[l + 1]; }
    else if (adaptive) { g.C2 = S.hnew[l] + H; g.ldc2 = 2 * H; }  // [q | memory_state] with memory_state == q
    StageScope sc(h, STAGE_GATE_GEMM, s);
    CAPDEC_RETURN_IF(gemm(h, c.precision, g, EPI_LSTM, s));
  }
  const float* q = S.hnew[layers - 1];
  const int64_t ld_q = adaptive ? 2 * H : H;
  if (adaptive) {
    // run_attention expects [q|mem] in S.qq: alias it onto the top layer's output
    float* saved = S.qq;
    S.qq = S.hnew[layers - 1];
    const int st = run_attention(h, S, feats, mask, q, ld_q, q + H, ld_q, S.cnew[layers - 1], H, S.B, S.ctx, alpha,
                                 ld_alpha, s);
    S.qq = saved;
    CAPDEC_RETURN_IF(st);
  } else {
    CAPDEC_RETURN_IF(run_attention(h, S, feats, mask, q, ld_q, nullptr, 0, nullptr, 0, S.B, S.ctx, alpha, ld_alpha, s));
  }
  // logits = output_layer(context)  (decoders.py:303)
  StageScope sc(h, STAGE_VOCAB_GEMM, s);
  CAPDEC_RETURN_IF(vocab_project(h, S, S.ctx, H, h->W("output_layer.weight"), h->W("output_layer.bias"), rows, S.logits,
                                 S.logits_ld ? S.logits_ld : c.vocab_size, s));
  return CAPDEC_OK;
}


// ---- transformer arch (src/models/decoders.py:317-493) ----------------------------------------------------------
std::string tl(int l, const char* name) { return "transformer_decoder.layers." + std::to_string(l) + "." + name; }

int gemm_w(const capdec_handle* h, const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, float* C,
           int64_t ldc, int M, int N, int K, int epi, cudaStream_t s, const SplitDst* a_pre = nullptr) {
  GemmArgs g{};
  g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.bias = bias; g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K;
  if (a_pre && a_pre->hi) { g.A_hi = a_pre->hi; g.A_lo = a_pre->lo; g.ld_as = a_pre->ld; }
  return gemm(h, h->cfg.precision, g, epi, s);
}

int prologue_transformer(const capdec_handle* h, Session& S, const float* feats, cudaStream_t s) {
  StageScope sc(h, STAGE_PROLOGUE, s);
  const int H = h->cfg.hidden_dim, rows = S.B * S.L;
  // memory = visual_projection(features)  (decoders.py:453), then every layer's cross-attention K|V projection of
  // it, hoisted out of the step loop (nn.MultiheadAttention in_proj rows [H:3H) are the packed k and v projections)
  CAPDEC_RETURN_IF(linear(h, feats, H, "visual_projection", S.tmem, H, rows, EPI_STORE, s, nullptr, 0, nullptr, &S.mmem));
  for (int l = 0; l < h->cfg.num_layers; ++l) {
    const float* w = h->W(tl(l, "multihead_attn.in_proj_weight"));
    const float* b = h->W(tl(l, "multihead_attn.in_proj_bias"));
    // K and V into separate dense [B,L,H] buffers (what the streaming attention kernel's bulk copies want)
    CAPDEC_RETURN_IF(gemm_w(h, S.tmem, H, w + (size_t)H * H, H, b + H, S.tck[l], H, rows, H, H, EPI_STORE, s, &S.mmem));
    CAPDEC_RETURN_IF(gemm_w(h, S.tmem, H, w + (size_t)2 * H * H, H, b + 2 * H, S.tcv[l], H, rows, H, H, EPI_STORE, s, &S.mmem));
  }
  S.anc_cur = -1;
  return CAPDEC_OK;
}

// one KV-cached decode step at position t: tokens S.next_tok -> S.logits
int step_transformer(const capdec_handle* h, Session& S, const uint8_t* mask, int t, cudaStream_t s) {
  const capdec_config& c = h->cfg;
  const int H = c.hidden_dim, rows = S.R, heads = c.num_heads;
  const DevTensor* pos = h->find("position_encoding.weight");
  CAPDEC_REQUIRE(t < pos->shape[0], CAPDEC_ERR_INVALID, "position %d exceeds position_encoding rows %lld", t, (long long)pos->shape[0]);
  const int F = (int)h->find(tl(0, "linear1.weight"))->shape[0];
  { StageScope sc(h, STAGE_GATHER, s);
    CAPDEC_RETURN_IF(embed_pos(S.next_tok, h->W("embedding.weight"), pos->p + (size_t)t * H, S.tx, rows, H, s, &S.mx)); }
  const float eps = 1e-5f;
  for (int l = 0; l < c.num_layers; ++l) {
    // self-attention block: x = norm1(x + out_proj(attn(in_proj(x))))   (post-LN, norm_first=False)
    { StageScope sc(h, STAGE_SMALL_GEMM, s);
      CAPDEC_RETURN_IF(gemm_w(h, S.tx, H, h->W(tl(l, "self_attn.in_proj_weight")), H, h->W(tl(l, "self_attn.in_proj_bias")),
                              S.tqkv, 3 * H, rows, 3 * H, H, EPI_STORE, s, &S.mx)); }
    { StageScope sc(h, STAGE_ATTENTION, s);
      SelfAttnArgs a{};
      a.qkv = S.tqkv; a.ld_qkv = 3 * H; a.cache_k = S.tcache_k[l]; a.cache_v = S.tcache_v[l];
      a.anc = S.anc_cur >= 0 ? S.anc[S.anc_cur] : nullptr; a.n_prefix = 0; a.rows_per_image = S.k;
      a.scale = (float)(1.0 / sqrt((double)(H / heads))); a.out = S.tsa; a.ld_out = H; a.out_split = S.msa;
      a.rows = rows; a.H = H; a.heads = heads; a.T = S.T; a.t = t;
      a.key_tok = S.key_tok; a.ld_key_tok = S.ld_key_tok; a.key_pad = S.key_pad;
      CAPDEC_RETURN_IF(self_attn_decode(a, s)); }
    { StageScope sc(h, STAGE_SMALL_GEMM, s);
      CAPDEC_RETURN_IF(linear(h, S.tsa, H, tl(l, "self_attn.out_proj"), S.ty, H, rows, EPI_STORE, s, nullptr, 0, &S.msa));
      CAPDEC_RETURN_IF(add_layernorm(S.tx, S.ty, h->W(tl(l, "norm1.weight")), h->W(tl(l, "norm1.bias")), nullptr, S.tx, rows, H, eps, s, &S.mx));
      // cross-attention block over the hoisted K|V of the image regions: x = norm2(x + out_proj(attn(q(x), K, V)))
      CAPDEC_RETURN_IF(gemm_w(h, S.tx, H, h->W(tl(l, "multihead_attn.in_proj_weight")), H, h->W(tl(l, "multihead_attn.in_proj_bias")),
                              S.tqc, H, rows, H, H, EPI_STORE, s, &S.mx)); }
    { StageScope sc(h, STAGE_ATTENTION, s);
      MhaArgs m{};
      // memory_key_padding_mask: masked region keys get -1e9 (decoders.py:393-398; attention.py:97-100,183-186 semantics)
      m.q = S.tqc; m.ld_q = H; m.kproj = S.tck[l]; m.vproj = S.tcv[l]; m.ld_kv = H; m.mask = mask;
      m.denom = (float)sqrt((double)(H / heads)); m.out = S.tca; m.ld_out = H; m.alpha = nullptr; m.ld_alpha = 0;
      m.B = S.B; m.L = S.L; m.H = H; m.heads = heads; m.k = S.k;
      CAPDEC_RETURN_IF(mha_attention(m, s)); }
    { StageScope sc(h, STAGE_SMALL_GEMM, s);
      CAPDEC_RETURN_IF(linear(h, S.tca, H, tl(l, "multihead_attn.out_proj"), S.ty, H, rows, EPI_STORE, s));
      CAPDEC_RETURN_IF(add_layernorm(S.tx, S.ty, h->W(tl(l, "norm2.weight")), h->W(tl(l, "norm2.bias")), nullptr, S.tx, rows, H, eps, s, &S.mx)); }
    // feed-forward block: x = norm3(x + linear2(gelu(linear1(x))))
    { StageScope sc(h, STAGE_GATE_GEMM, s);
      CAPDEC_RETURN_IF(linear(h, S.tx, H, tl(l, "linear1"), S.mff.hi ? nullptr : S.tff, F, rows, EPI_GELU, s, nullptr, 0, &S.mx, &S.mff));
      CAPDEC_RETURN_IF(linear(h, S.tff, F, tl(l, "linear2"), S.ty, H, rows, EPI_STORE, s, nullptr, 0, &S.mff));
      CAPDEC_RETURN_IF(add_layernorm(S.tx, S.ty, h->W(tl(l, "norm3.weight")), h->W(tl(l, "norm3.bias")), nullptr, S.tx, rows, H, eps, s, &S.mx)); }
  }
  StageScope sc(h, STAGE_VOCAB_GEMM, s);
  return vocab_project(h, S, S.tx, H, h->W("output_layer.weight"), h->W("output_layer.bias"), rows, S.logits,
                       S.logits_ld ? S.logits_ld : c.vocab_size, s);
}

// after position t_done: record tokens, and (beam) re-point every row's earlier cache positions at its parent's
int commit_transformer(const capdec_handle* h, Session& S, const int32_t* src, int32_t* tok_out, int64_t ld_tok, int pos,
                       int t_done, cudaStream_t s) {
  StageScope sc(h, STAGE_GATHER, s);
  if (tok_out && pos >= 0 && pos < S.T) {
    GatherArgs g{};
    g.rows = S.R; g.tok = S.next_tok; g.src = nullptr; g.embedding = nullptr; g.tok_out = tok_out; g.ld_tok = ld_tok;
    g.pos = pos; g.n_state = 0;
    CAPDEC_RETURN_IF(gather_rows(g, s));
  }
  if (src && t_done >= 0) {
    const int nxt = S.anc_cur < 0 ? 0 : (S.anc_cur ^ 1);
    // first reorder: the old table is the identity; build it lazily by pointing anc_old at a table never read (p == t_done only)
    CAPDEC_REQUIRE(S.anc_cur >= 0 || t_done == 0, CAPDEC_ERR_STATE, "ancestor table must start at position 0");
    CAPDEC_RETURN_IF(reorder_ancestors(src, S.anc_cur >= 0 ? S.anc[S.anc_cur] : S.anc[1], S.anc[nxt], S.R, S.T, t_done, s));
    S.anc_cur = nxt;
  }
  return CAPDEC_OK;
}


// ---- GPT-2 arch (src/models/decoders.py:496-656; block math = transformers GPT2Block, pre-LN, gelu_new) -------------
std::string gl(int l, const char* name) { return "model.transformer.h." + std::to_string(l) + "." + name; }

int prologue_gpt2(const capdec_handle* h, Session& S, const float* pooled, cudaStream_t s) {
  StageScope sc(h, STAGE_PROLOGUE, s);
  const capdec_config& c = h->cfg;
  // image_prefix = image_to_prefix(pooled).view(B, P, n_embd)  (decoders.py:634-637); the SAME tensor is the past
  // key and the past value of every layer (decoders.py:608-615)
  CAPDEC_RETURN_IF(linear(h, pooled, c.feature_dim, "image_to_prefix", S.gprefix, (int64_t)S.n_prefix * c.hidden_dim, S.B, EPI_STORE, s));
  S.anc_cur = -1;
  return CAPDEC_OK;
}

int step_gpt2(const capdec_handle* h, Session& S, int t, cudaStream_t s) {
  const capdec_config& c = h->cfg;
  const int H = c.hidden_dim, rows = S.R, heads = c.num_heads, P = S.n_prefix;
  const DevTensor* wpe = h->find("model.transformer.wpe.weight");
  CAPDEC_REQUIRE(P + t < wpe->shape[0], CAPDEC_ERR_INVALID, "position %d exceeds n_positions %lld", P + t, (long long)wpe->shape[0]);
  const int F = (int)h->find(gl(0, "mlp.c_fc.weight"))->shape[0];
  const float eps = 1e-5f;
  // position ids continue after the prefix: past_len + t  (HF prepare_inputs_for_generation with a cache of length P)
  { StageScope sc(h, STAGE_GATHER, s);
    CAPDEC_RETURN_IF(embed_pos(S.next_tok, h->W("model.transformer.wte.weight"), wpe->p + (size_t)(P + t) * H, S.tx, rows, H, s)); }
  const float* pending = nullptr;
  // the pre-LN outputs feed GEMMs only: when those read the operand mirror (S.mx) the fp32 copy is never written
  float* const txn_out = S.mx.hi ? nullptr : S.txn;
  for (int l = 0; l < c.num_layers; ++l) {
    { StageScope sc(h, STAGE_SMALL_GEMM, s);
      CAPDEC_RETURN_IF(add_layernorm(S.tx, pending, h->W(gl(l, "ln_1.weight")), h->W(gl(l, "ln_1.bias")), pending ? S.tx : nullptr, txn_out, rows, H, eps, s, &S.mx));
      CAPDEC_RETURN_IF(linear(h, S.txn, H, gl(l, "attn.c_attn"), S.tqkv, 3 * H, rows, EPI_STORE, s, nullptr, 0, &S.mx)); }
    { StageScope sc(h, STAGE_ATTENTION, s);
      SelfAttnArgs a{};
      a.qkv = S.tqkv; a.ld_qkv = 3 * H; a.cache_k = S.tcache_k[l]; a.cache_v = S.tcache_v[l];
      a.anc = S.anc_cur >= 0 ? S.anc[S.anc_cur] : nullptr;
      a.prefix_k = S.gprefix; a.prefix_v = S.gprefix; a.n_prefix = P; a.rows_per_image = S.k;
      a.scale = (float)(1.0 / sqrt((double)(H / heads))); a.out = S.tsa; a.ld_out = H; a.out_split = S.msa;
      a.rows = rows; a.H = H; a.heads = heads; a.T = S.T; a.t = t;
      a.key_tok = S.key_tok; a.ld_key_tok = S.ld_key_tok; a.key_pad = S.key_pad;
      CAPDEC_RETURN_IF(self_attn_decode(a, s)); }
    { StageScope sc(h, STAGE_SMALL_GEMM, s);
      CAPDEC_RETURN_IF(linear(h, S.tsa, H, gl(l, "attn.c_proj"), S.ty, H, rows, EPI_STORE, s, nullptr, 0, &S.msa));
      CAPDEC_RETURN_IF(add_layernorm(S.tx, S.ty, h->W(gl(l, "ln_2.weight")), h->W(gl(l, "ln_2.bias")), S.tx, txn_out, rows, H, eps, s, &S.mx)); }
    { StageScope sc(h, STAGE_GATE_GEMM, s);
      CAPDEC_RETURN_IF(linear(h, S.txn, H, gl(l, "mlp.c_fc"), S.mff.hi ? nullptr : S.tff, F, rows, EPI_GELU_TANH, s, nullptr, 0, &S.mx, &S.mff));
      CAPDEC_RETURN_IF(linear(h, S.tff, F, gl(l, "mlp.c_proj"), S.ty, H, rows, EPI_STORE, s, nullptr, 0, &S.mff)); }
    pending = S.ty;
  }
  { StageScope sc(h, STAGE_SMALL_GEMM, s);
    CAPDEC_RETURN_IF(add_layernorm(S.tx, pending, h->W("model.transformer.ln_f.weight"), h->W("model.transformer.ln_f.bias"), S.tx, txn_out, rows, H, eps, s, &S.mx)); }
  StageScope sc(h, STAGE_VOCAB_GEMM, s);
  // lm_head is tied to wte and has no bias
  return vocab_project(h, S, S.txn, H, h->W("model.lm_head.weight"), nullptr, rows, S.logits, S.logits_ld ? S.logits_ld : c.vocab_size, s);
}

int step_any(const capdec_handle* h, Session& S, const float* feats, const uint8_t* mask, float* alpha,
             int64_t ld_alpha, int t, cudaStream_t s) {
  if (is_transformer(h)) return step_transformer(h, S, mask, t, s);
  if (is_gpt2(h)) return step_gpt2(h, S, t, s);
  if (is_legacy(h)) return step_legacy(h, S, feats, S.B, S.logits, S.logits_ld ? S.logits_ld : h->cfg.vocab_size, alpha, ld_alpha, s);
  return step_lstm(h, S, feats, mask, alpha, ld_alpha, s);
}

// the step's commit as gather arguments: reorder state by back-pointer, embed the chosen tokens, optionally record them
void build_gather(const capdec_handle* h, Session& S, const int32_t* src, int32_t* tok_out, int64_t ld_tok, int pos,
                  bool with_state, GatherArgs* out) {
  S.row_src = src;
  const capdec_config& c = h->cfg;
  const int H = c.hidden_dim, E = c.embed_dim, D = c.feature_dim;
  GatherArgs g{};
  g.rows = S.R; g.tok = S.next_tok; g.src = src;
  g.embedding = h->W("embedding.weight"); g.E = E; g.x_emb = S.X[0]; g.ld_x = S.ldX[0];
  g.tok_out = (tok_out && pos >= 0 && pos < S.T) ? tok_out : nullptr; g.ld_tok = ld_tok; g.pos = pos;
  int n = 0;
  if (with_state) {
    for (int l = 0; l < c.num_layers; ++l) {
      const int in = l == 0 ? (is_legacy(h) ? E + D : E + H) : H;
      const bool wide = !is_legacy(h) && c.attention == CAPDEC_ATT_ADAPTIVE && l == c.num_layers - 1;
      g.state_src[n] = S.hnew[l]; g.ld_src[n] = wide ? 2 * H : H; g.state_dst[n] = S.X[l] + in; g.ld_dst[n] = S.ldX[l]; g.width[n] = H; ++n;
      g.state_src[n] = S.cnew[l]; g.ld_src[n] = H; g.state_dst[n] = S.c[l]; g.ld_dst[n] = H; g.width[n] = H; ++n;
    }
    if (!is_legacy(h)) {  // prev_ctx (decoders.py:297)
      g.state_src[n] = S.ctx; g.ld_src[n] = H; g.state_dst[n] = S.X[0] + E; g.ld_dst[n] = S.ldX[0]; g.width[n] = H; ++n;
    }
  }
  g.n_state = n;
  if (S.presplit && is_legacy(h)) {
    if (h->emb_gates) g.embedding = nullptr;   // the gate GEMM reads the token's row of emb_gates instead of X[:, :E]
    g.x_split = SplitDst{S.xs_hi, S.xs_lo, S.ldX[0], tc_kind(c.precision)};
    for (int i = 0; i < 16; ++i) g.state_split_col[i] = -1;
    if (with_state) g.state_split_col[0] = E + D;   // state 0 = new h of the single legacy layer -> X[:, E+D:]
  }
  *out = g;
}

int commit(const capdec_handle* h, Session& S, const int32_t* src, int32_t* tok_out, int64_t ld_tok, int pos,
           bool with_state, cudaStream_t s, int t_done = -1) {
  if (is_tf_family(h)) return commit_transformer(h, S, with_state ? src : nullptr, tok_out, ld_tok, pos, t_done, s);
  GatherArgs g{};
  build_gather(h, S, src, tok_out, ld_tok, pos, with_state, &g);
  StageScope sc(h, STAGE_GATHER, s);
  return gather_rows(g, s);
}

int prologue_any(const capdec_handle* h, Session& S, const float* feats, const float* pooled, cudaStream_t s) {
  if (is_transformer(h)) return prologue_transformer(h, S, feats, s);
  if (is_gpt2(h)) return prologue_gpt2(h, S, pooled, s);
  if (is_legacy(h)) return prologue_legacy(h, S, feats, true, s);
  return prologue_lstm(h, S, feats, pooled, s);
}

int check_common(const capdec_handle* h, const void* feats, const float* pooled, int B, int L, int k, int T) {
  CAPDEC_REQUIRE(h != nullptr, CAPDEC_ERR_INVALID, "null handle");
  CAPDEC_REQUIRE(h->finalized, CAPDEC_ERR_STATE, "capdec_finalize has not been called");
  CAPDEC_REQUIRE(B == 0 || is_gpt2(h) || feats != nullptr, CAPDEC_ERR_INVALID, "features pointer is null");
  CAPDEC_REQUIRE(B == 0 || (h->cfg.arch != CAPDEC_ARCH_LSTM && !is_gpt2(h)) || pooled != nullptr, CAPDEC_ERR_INVALID,
                 "pooled_features pointer is null");
  CAPDEC_REQUIRE(B >= 0 && L >= 1 && T >= 2, CAPDEC_ERR_INVALID, "bad sizes B=%d L=%d max_length=%d", B, L, T);
  CAPDEC_REQUIRE(k >= 1 && k <= kMaxRowsPerImage, CAPDEC_ERR_UNSUPPORTED, "rows per image %d not in [1,%d]", k,
                 kMaxRowsPerImage);
  return CAPDEC_OK;
}

static int need(const capdec_handle* h, const std::string& n, std::vector<int64_t> shape) {
  const DevTensor* t = h->find(n);
  CAPDEC_REQUIRE(t != nullptr, CAPDEC_ERR_STATE, "missing parameter '%s'", n.c_str());
  bool ok = t->shape.size() == shape.size();
  for (size_t i = 0; ok && i < shape.size(); ++i) ok = t->shape[i] == shape[i];
  if (!ok) {
    std::string got, want;
    for (auto v : t->shape) got += std::to_string(v) + ",";
    for (auto v : shape) want += std::to_string(v) + ",";
    set_error("parameter '%s' has shape [%s] but the config implies [%s]", n.c_str(), got.c_str(), want.c_str());
    return CAPDEC_ERR_INVALID;
  }
  return CAPDEC_OK;
}

template <typename T>
static int dev_alloc(capdec_handle* h, T** p, size_t n) {
  CAPDEC_CHECK_CUDA(cudaMalloc((void**)p, n * sizeof(T)));
  h->owned.push_back(*p);
  return CAPDEC_OK;
}

static int pack_gates(capdec_handle* h, const std::string& wih, const std::string& whh, const std::string& bih,
                      const std::string& bhh, int in, cudaStream_t s) {
  const int H = h->cfg.hidden_dim;
  float *w = nullptr, *b = nullptr;
  CAPDEC_RETURN_IF(dev_alloc(h, &w, (size_t)4 * H * (in + H)));
  CAPDEC_RETURN_IF(dev_alloc(h, &b, (size_t)4 * H));
  for (int g = 0; g < 4; ++g) {
    CAPDEC_RETURN_IF(scatter_rows(w, in + H, 4, g, 0, h->W(wih) + (size_t)g * H * in, in, H, in, false, s));
    CAPDEC_RETURN_IF(scatter_rows(w, in + H, 4, g, in, h->W(whh) + (size_t)g * H * H, H, H, H, false, s));
    CAPDEC_RETURN_IF(scatter_rows(b, 1, 4, g, 0, h->W(bih) + (size_t)g * H, 1, H, 1, false, s));
    CAPDEC_RETURN_IF(scatter_rows(b, 1, 4, g, 0, h->W(bhh) + (size_t)g * H, 1, H, 1, true, s));
  }
  h->w_gates.push_back(w);
  h->b_gates.push_back(b);
  h->gate_in.push_back(in);
  return CAPDEC_OK;
}

}  // namespace
}  // namespace capdec

using namespace capdec;

// ================================================ C ABI =================================================
extern "C" {

const char* capdec_last_error(void) { return g_err; }
int capdec_version(void) { return CAPDEC_VERSION; }
int64_t capdec_launch_count(void) { return g_launch_count.load(); }

int capdec_create(const capdec_config* cfg, capdec_handle** out) {
  CAPDEC_REQUIRE(cfg && out, CAPDEC_ERR_INVALID, "capdec_create: null argument");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  CAPDEC_REQUIRE(e == cudaSuccess && ndev > 0, CAPDEC_ERR_CUDA,
                 "capdec_create: no CUDA device (%s); libcapdec has no CPU fallback", cudaGetErrorString(e));
  CAPDEC_REQUIRE(cfg->arch == CAPDEC_ARCH_LEGACY_SAT || cfg->arch == CAPDEC_ARCH_LSTM ||
                     cfg->arch == CAPDEC_ARCH_TRANSFORMER || cfg->arch == CAPDEC_ARCH_GPT2,
                 CAPDEC_ERR_UNSUPPORTED, "unsupported decoder arch %d", cfg->arch);
  CAPDEC_REQUIRE(cfg->attention >= CAPDEC_ATT_SOFT && cfg->attention <= CAPDEC_ATT_AOA, CAPDEC_ERR_UNSUPPORTED,
                 "Unsupported attention type: %d", cfg->attention);
  CAPDEC_REQUIRE(cfg->precision >= CAPDEC_PREC_FP32 && cfg->precision <= CAPDEC_PREC_BF16X3, CAPDEC_ERR_UNSUPPORTED,
                 "unsupported precision %d", cfg->precision);
  CAPDEC_REQUIRE(cfg->vocab_size > 0 && cfg->hidden_dim > 0 && cfg->embed_dim > 0 && cfg->num_layers >= 1 &&
                     cfg->num_layers <= (cfg->arch >= CAPDEC_ARCH_TRANSFORMER ? 64 : 7),
                 CAPDEC_ERR_INVALID, "bad dimensions in config");
  CAPDEC_REQUIRE(cfg->hidden_dim % 4 == 0 && cfg->embed_dim % 4 == 0 && cfg->feature_dim % 4 == 0 &&
                     cfg->attention_dim % 4 == 0,
                 CAPDEC_ERR_UNSUPPORTED, "hidden/embed/feature/attention dims must be multiples of 4");
  if (cfg->arch == CAPDEC_ARCH_TRANSFORMER || cfg->arch == CAPDEC_ARCH_GPT2) {
    CAPDEC_REQUIRE(cfg->embed_dim == cfg->hidden_dim && (cfg->arch == CAPDEC_ARCH_GPT2 || cfg->feature_dim == cfg->hidden_dim),
                   CAPDEC_ERR_INVALID, "transformer arch requires feature_dim == embed_dim == hidden_dim (decoders.py:343-375)");
    CAPDEC_REQUIRE(cfg->num_heads >= 1 && cfg->hidden_dim % cfg->num_heads == 0 && cfg->hidden_dim / cfg->num_heads <= 128 &&
                       (cfg->hidden_dim / cfg->num_heads) % 4 == 0,
                   CAPDEC_ERR_UNSUPPORTED, "transformer arch needs head_dim <= 128 and a multiple of 4");
  } else if (cfg->arch == CAPDEC_ARCH_LSTM) {
    CAPDEC_REQUIRE(cfg->feature_dim == cfg->hidden_dim && cfg->attention_dim == cfg->hidden_dim, CAPDEC_ERR_INVALID,
                   "LSTM arch requires feature_dim == attention_dim == hidden_dim (attention.py:45-51)");
    CAPDEC_REQUIRE(cfg->num_heads >= 1 && cfg->hidden_dim % cfg->num_heads == 0, CAPDEC_ERR_INVALID,
                   "Hidden dim must be divisible by num heads");
  } else {
    CAPDEC_REQUIRE(cfg->num_layers == 1, CAPDEC_ERR_INVALID, "legacy decoder has a single LSTMCell");
  }
  capdec_handle* h = new capdec_handle();
  h->cfg = *cfg;
  *out = h;
  return CAPDEC_OK;
}

void capdec_destroy(capdec_handle* h) {
  if (!h) return;
  for (auto& kv : h->w) cudaFree(kv.second.p);
  for (void* p : h->owned) cudaFree(p);
  gemm_tc_release(h);
  if (h->stage_dev) cudaFree(h->stage_dev);
  for (int i = 0; i < kHostBufs; ++i) {
    if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]);
    if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
  }
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  if (h->stream_compute) cudaStreamDestroy(h->stream_compute);
  if (h->stream_copy) cudaStreamDestroy(h->stream_copy);
  delete h;
}

int capdec_set_weight(capdec_handle* h, const char* name, const float* data_dev, const int64_t* shape, int32_t rank,
                      void* stream) {
  CAPDEC_REQUIRE(h && name && data_dev && shape && rank >= 1 && rank <= 4, CAPDEC_ERR_INVALID,
                 "capdec_set_weight: bad argument");
  DevTensor t;
  t.numel = 1;
  for (int i = 0; i < rank; ++i) { t.shape.push_back(shape[i]); t.numel *= shape[i]; }
  CAPDEC_REQUIRE(t.numel > 0, CAPDEC_ERR_INVALID, "capdec_set_weight(%s): empty tensor", name);
  auto it = h->w.find(name);
  if (it != h->w.end()) { cudaFree(it->second.p); h->w.erase(it); }
  CAPDEC_CHECK_CUDA(cudaMalloc((void**)&t.p, (size_t)t.numel * sizeof(float)));
  CAPDEC_CHECK_CUDA(cudaMemcpyAsync(t.p, data_dev, (size_t)t.numel * sizeof(float), cudaMemcpyDeviceToDevice,
                                    (cudaStream_t)stream));
  h->w[name] = t;
  h->finalized = false;
  return CAPDEC_OK;
}

int capdec_finalize(capdec_handle* h, void* stream) {
  CAPDEC_REQUIRE(h, CAPDEC_ERR_INVALID, "null handle");
  cudaStream_t s = (cudaStream_t)stream;
  const capdec_config& c = h->cfg;
  const int64_t H = c.hidden_dim, E = c.embed_dim, D = c.feature_dim, A = c.attention_dim, V = c.vocab_size;
  for (void* p : h->owned) cudaFree(p);
  h->owned.clear(); h->w_gates.clear(); h->b_gates.clear(); h->gate_in.clear();
  h->w_hproj = h->b_hproj = h->w_init = h->b_init = h->w_aoa = h->b_aoa = nullptr;
  h->w_vocab_cat = h->b_vocab_cat = nullptr; h->vocab_cat_n = 0; h->emb_gates = nullptr;

  if (is_gpt2(h)) {
    // src/models/decoders.py:513-561 + transformers GPT2LMHeadModel parameter names; Conv1D weights are bound
    // TRANSPOSED, i.e. in nn.Linear layout [out, in] (the Python binding does the transpose)
    const int64_t Din = c.feature_dim;
    CAPDEC_RETURN_IF(need(h, "model.transformer.wte.weight", {V, H}));
    CAPDEC_RETURN_IF(need(h, "model.lm_head.weight", {V, H}));
    const DevTensor* wpe = h->find("model.transformer.wpe.weight");
    CAPDEC_REQUIRE(wpe && wpe->shape.size() == 2 && wpe->shape[1] == H, CAPDEC_ERR_STATE, "missing parameter 'model.transformer.wpe.weight'");
    const DevTensor* ip = h->find("image_to_prefix.weight");
    CAPDEC_REQUIRE(ip && ip->shape.size() == 2 && ip->shape[1] == Din && ip->shape[0] % H == 0, CAPDEC_ERR_STATE,
                   "missing or mis-shaped parameter 'image_to_prefix.weight'");
    CAPDEC_RETURN_IF(need(h, "image_to_prefix.bias", {ip->shape[0]}));
    CAPDEC_RETURN_IF(need(h, "model.transformer.ln_f.weight", {H})); CAPDEC_RETURN_IF(need(h, "model.transformer.ln_f.bias", {H}));
    const DevTensor* f1 = h->find(gl(0, "mlp.c_fc.weight"));
    CAPDEC_REQUIRE(f1 && f1->shape.size() == 2, CAPDEC_ERR_STATE, "missing parameter '%s'", gl(0, "mlp.c_fc.weight").c_str());
    const int64_t F = f1->shape[0];
    for (int l = 0; l < c.num_layers; ++l) {
      CAPDEC_RETURN_IF(need(h, gl(l, "ln_1.weight"), {H})); CAPDEC_RETURN_IF(need(h, gl(l, "ln_1.bias"), {H}));
      CAPDEC_RETURN_IF(need(h, gl(l, "ln_2.weight"), {H})); CAPDEC_RETURN_IF(need(h, gl(l, "ln_2.bias"), {H}));
      CAPDEC_RETURN_IF(need(h, gl(l, "attn.c_attn.weight"), {3 * H, H})); CAPDEC_RETURN_IF(need(h, gl(l, "attn.c_attn.bias"), {3 * H}));
      CAPDEC_RETURN_IF(need(h, gl(l, "attn.c_proj.weight"), {H, H})); CAPDEC_RETURN_IF(need(h, gl(l, "attn.c_proj.bias"), {H}));
      CAPDEC_RETURN_IF(need(h, gl(l, "mlp.c_fc.weight"), {F, H})); CAPDEC_RETURN_IF(need(h, gl(l, "mlp.c_fc.bias"), {F}));
      CAPDEC_RETURN_IF(need(h, gl(l, "mlp.c_proj.weight"), {H, F})); CAPDEC_RETURN_IF(need(h, gl(l, "mlp.c_proj.bias"), {H}));
    }
  } else if (is_transformer(h)) {
    // src/models/decoders.py:343-375 (nn.TransformerDecoderLayer parameter names)
    CAPDEC_RETURN_IF(need(h, "embedding.weight", {V, H}));
    const DevTensor* pos = h->find("position_encoding.weight");
    CAPDEC_REQUIRE(pos && pos->shape.size() == 2 && pos->shape[1] == H, CAPDEC_ERR_STATE, "missing parameter 'position_encoding.weight'");
    CAPDEC_RETURN_IF(need(h, "output_layer.weight", {V, H})); CAPDEC_RETURN_IF(need(h, "output_layer.bias", {V}));
    CAPDEC_RETURN_IF(need(h, "visual_projection.weight", {H, H})); CAPDEC_RETURN_IF(need(h, "visual_projection.bias", {H}));
    const DevTensor* f1 = h->find(tl(0, "linear1.weight"));
    CAPDEC_REQUIRE(f1 && f1->shape.size() == 2, CAPDEC_ERR_STATE, "missing parameter '%s'", tl(0, "linear1.weight").c_str());
    const int64_t F = f1->shape[0];
    for (int l = 0; l < c.num_layers; ++l) {
      for (const char* att : {"self_attn", "multihead_attn"}) {
        const std::string a = att;
        CAPDEC_RETURN_IF(need(h, tl(l, (a + ".in_proj_weight").c_str()), {3 * H, H}));
        CAPDEC_RETURN_IF(need(h, tl(l, (a + ".in_proj_bias").c_str()), {3 * H}));
        CAPDEC_RETURN_IF(need(h, tl(l, (a + ".out_proj.weight").c_str()), {H, H}));
        CAPDEC_RETURN_IF(need(h, tl(l, (a + ".out_proj.bias").c_str()), {H}));
      }
      CAPDEC_RETURN_IF(need(h, tl(l, "linear1.weight"), {F, H})); CAPDEC_RETURN_IF(need(h, tl(l, "linear1.bias"), {F}));
      CAPDEC_RETURN_IF(need(h, tl(l, "linear2.weight"), {H, F})); CAPDEC_RETURN_IF(need(h, tl(l, "linear2.bias"), {H}));
      for (const char* n : {"norm1", "norm2", "norm3"}) {
        const std::string nn = n;
        CAPDEC_RETURN_IF(need(h, tl(l, (nn + ".weight").c_str()), {H})); CAPDEC_RETURN_IF(need(h, tl(l, (nn + ".bias").c_str()), {H}));
      }
    }
  } else if (is_legacy(h)) {
    // models/decoder.py:33-54
    CAPDEC_RETURN_IF(need(h, "enc_att.weight", {A, D})); CAPDEC_RETURN_IF(need(h, "enc_att.bias", {A}));
    CAPDEC_RETURN_IF(need(h, "dec_att.weight", {A, H})); CAPDEC_RETURN_IF(need(h, "dec_att.bias", {A}));
    CAPDEC_RETURN_IF(need(h, "att.weight", {1, A}));     CAPDEC_RETURN_IF(need(h, "att.bias", {1}));
    CAPDEC_RETURN_IF(need(h, "decode_step.weight_ih", {4 * H, E + D}));
    CAPDEC_RETURN_IF(need(h, "decode_step.weight_hh", {4 * H, H}));
    CAPDEC_RETURN_IF(need(h, "decode_step.bias_ih", {4 * H})); CAPDEC_RETURN_IF(need(h, "decode_step.bias_hh", {4 * H}));
    CAPDEC_RETURN_IF(need(h, "h_lin.weight", {H, D})); CAPDEC_RETURN_IF(need(h, "h_lin.bias", {H}));
    CAPDEC_RETURN_IF(need(h, "c_lin.weight", {H, D})); CAPDEC_RETURN_IF(need(h, "c_lin.bias", {H}));
    CAPDEC_RETURN_IF(need(h, "f_beta.weight", {D, H})); CAPDEC_RETURN_IF(need(h, "f_beta.bias", {D}));
    CAPDEC_RETURN_IF(need(h, "fc.weight", {V, H})); CAPDEC_RETURN_IF(need(h, "fc.bias", {V}));
    CAPDEC_RETURN_IF(need(h, "embedding.weight", {V, E}));
    CAPDEC_RETURN_IF(pack_gates(h, "decode_step.weight_ih", "decode_step.weight_hh", "decode_step.bias_ih",
                                "decode_step.bias_hh", (int)(E + D), s));
    CAPDEC_RETURN_IF(dev_alloc(h, &h->w_hproj, (size_t)(A + D) * H));
    CAPDEC_RETURN_IF(dev_alloc(h, &h->b_hproj, (size_t)(A + D)));
    CAPDEC_RETURN_IF(scatter_rows(h->w_hproj, H, 1, 0, 0, h->W("dec_att.weight"), H, (int)A, (int)H, false, s));
    CAPDEC_RETURN_IF(scatter_rows(h->w_hproj + A * H, H, 1, 0, 0, h->W("f_beta.weight"), H, (int)D, (int)H, false, s));
    CAPDEC_RETURN_IF(scatter_rows(h->b_hproj, 1, 1, 0, 0, h->W("dec_att.bias"), 1, (int)A, 1, false, s));
    CAPDEC_RETURN_IF(scatter_rows(h->b_hproj + A, 1, 1, 0, 0, h->W("f_beta.bias"), 1, (int)D, 1, false, s));
    if (c.precision != CAPDEC_PREC_FP32) {
      CAPDEC_RETURN_IF(dev_alloc(h, &h->emb_gates, (size_t)V * 4 * H));
      GemmArgs g{};
      g.A = h->W("embedding.weight"); g.lda = E; g.W = h->w_gates[0]; g.ldw = E + D + H; g.bias = nullptr;
      g.C = h->emb_gates; g.ldc = 4 * H; g.M = (int)V; g.N = (int)(4 * H); g.K = (int)E;
      CAPDEC_RETURN_IF(gemm_ffma(g, EPI_STORE, s));
    }
    if (c.precision != CAPDEC_PREC_FP32) {
      const int64_t Vp = (V + 255) / 256 * 256, Nc = Vp + A + D;
      CAPDEC_RETURN_IF(dev_alloc(h, &h->w_vocab_cat, (size_t)Nc * H));
      CAPDEC_RETURN_IF(dev_alloc(h, &h->b_vocab_cat, (size_t)Nc));
      CAPDEC_CHECK_CUDA(cudaMemsetAsync(h->w_vocab_cat, 0, (size_t)Nc * H * sizeof(float), s));
      CAPDEC_CHECK_CUDA(cudaMemsetAsync(h->b_vocab_cat, 0, (size_t)Nc * sizeof(float), s));
      CAPDEC_RETURN_IF(scatter_rows(h->w_vocab_cat, H, 1, 0, 0, h->W("fc.weight"), H, (int)V, (int)H, false, s));
      CAPDEC_RETURN_IF(scatter_rows(h->b_vocab_cat, 1, 1, 0, 0, h->W("fc.bias"), 1, (int)V, 1, false, s));
      CAPDEC_RETURN_IF(scatter_rows(h->w_vocab_cat + Vp * H, H, 1, 0, 0, h->w_hproj, H, (int)(A + D), (int)H, false, s));
      CAPDEC_RETURN_IF(scatter_rows(h->b_vocab_cat + Vp, 1, 1, 0, 0, h->b_hproj, 1, (int)(A + D), 1, false, s));
      h->vocab_cat_n = (int)Nc;
    }
    CAPDEC_RETURN_IF(dev_alloc(h, &h->w_init, (size_t)2 * H * D));
    CAPDEC_RETURN_IF(dev_alloc(h, &h->b_init, (size_t)2 * H));
    CAPDEC_RETURN_IF(scatter_rows(h->w_init, D, 1, 0, 0, h->W("h_lin.weight"), D, (int)H, (int)D, false, s));
    CAPDEC_RETURN_IF(scatter_rows(h->w_init + H * D, D, 1, 0, 0, h->W("c_lin.weight"), D, (int)H, (int)D, false, s));
    CAPDEC_RETURN_IF(scatter_rows(h->b_init, 1, 1, 0, 0, h->W("h_lin.bias"), 1, (int)H, 1, false, s));
    CAPDEC_RETURN_IF(scatter_rows(h->b_init + H, 1, 1, 0, 0, h->W("c_lin.bias"), 1, (int)H, 1, false, s));
    CAPDEC_CHECK_CUDA(cudaMemcpyAsync(&h->energy_bias, h->W("att.bias"), sizeof(float), cudaMemcpyDeviceToHost, s));
  } else {
    // src/models/decoders.py:92-117
    const int64_t Ln = c.num_layers;
    CAPDEC_RETURN_IF(need(h, "embedding.weight", {V, E}));
    CAPDEC_RETURN_IF(need(h, "output_layer.weight", {V, H})); CAPDEC_RETURN_IF(need(h, "output_layer.bias", {V}));
    CAPDEC_RETURN_IF(need(h, "init_h.weight", {H * Ln, H})); CAPDEC_RETURN_IF(need(h, "init_h.bias", {H * Ln}));
    CAPDEC_RETURN_IF(need(h, "init_c.weight", {H * Ln, H})); CAPDEC_RETURN_IF(need(h, "init_c.bias", {H * Ln}));
    for (int l = 0; l < Ln; ++l) {
      const int64_t in = l == 0 ? E + H : H;
      const std::string sfx = "_l" + std::to_string(l);
      CAPDEC_RETURN_IF(need(h, "lstm.weight_ih" + sfx, {4 * H, in}));
      CAPDEC_RETURN_IF(need(h, "lstm.weight_hh" + sfx, {4 * H, H}));
      CAPDEC_RETURN_IF(need(h, "lstm.bias_ih" + sfx, {4 * H}));
      CAPDEC_RETURN_IF(need(h, "lstm.bias_hh" + sfx, {4 * H}));
      CAPDEC_RETURN_IF(pack_gates(h, "lstm.weight_ih" + sfx, "lstm.weight_hh" + sfx, "lstm.bias_ih" + sfx,
                                  "lstm.bias_hh" + sfx, (int)in, s));
    }
    CAPDEC_RETURN_IF(dev_alloc(h, &h->w_init, (size_t)2 * Ln * H * H));
    CAPDEC_RETURN_IF(dev_alloc(h, &h->b_init, (size_t)2 * Ln * H));
    CAPDEC_RETURN_IF(scatter_rows(h->w_init, H, 1, 0, 0, h->W("init_h.weight"), H, (int)(Ln * H), (int)H, false, s));
    CAPDEC_RETURN_IF(scatter_rows(h->w_init + Ln * H * H, H, 1, 0, 0, h->W("init_c.weight"), H, (int)(Ln * H), (int)H, false, s));
    CAPDEC_RETURN_IF(scatter_rows(h->b_init, 1, 1, 0, 0, h->W("init_h.bias"), 1, (int)(Ln * H), 1, false, s));
    CAPDEC_RETURN_IF(scatter_rows(h->b_init + Ln * H, 1, 1, 0, 0, h->W("init_c.bias"), 1, (int)(Ln * H), 1, false, s));
    // attention parameters (src/models/attention.py:48-52,133-137,232-239,311-320)
    const std::string p = base_prefix(h);
    CAPDEC_RETURN_IF(need(h, p + "query_proj.weight", {H, H})); CAPDEC_RETURN_IF(need(h, p + "query_proj.bias", {H}));
    CAPDEC_RETURN_IF(need(h, p + "key_proj.weight", {H, H}));   CAPDEC_RETURN_IF(need(h, p + "key_proj.bias", {H}));
    if (base_is_mha(h)) {
      CAPDEC_RETURN_IF(need(h, p + "value_proj.weight", {H, H}));  CAPDEC_RETURN_IF(need(h, p + "value_proj.bias", {H}));
      CAPDEC_RETURN_IF(need(h, p + "output_proj.weight", {H, H})); CAPDEC_RETURN_IF(need(h, p + "output_proj.bias", {H}));
    } else {
      CAPDEC_RETURN_IF(need(h, p + "energy.weight", {1, H})); CAPDEC_RETURN_IF(need(h, p + "energy.bias", {1}));
      CAPDEC_CHECK_CUDA(cudaMemcpyAsync(&h->energy_bias, h->W(p + "energy.bias"), sizeof(float), cudaMemcpyDeviceToHost, s));
    }
    if (c.attention == CAPDEC_ATT_AOA) {
      CAPDEC_RETURN_IF(need(h, "attention.query_proj.weight", {H, H})); CAPDEC_RETURN_IF(need(h, "attention.query_proj.bias", {H}));
      CAPDEC_RETURN_IF(need(h, "attention.info_vector_proj.0.weight", {H, 2 * H}));
      CAPDEC_RETURN_IF(need(h, "attention.info_vector_proj.0.bias", {H}));
      CAPDEC_RETURN_IF(need(h, "attention.info_gate_proj.0.weight", {H, 2 * H}));
      CAPDEC_RETURN_IF(need(h, "attention.info_gate_proj.0.bias", {H}));
      CAPDEC_RETURN_IF(dev_alloc(h, &h->w_aoa, (size_t)2 * H * 2 * H));
      CAPDEC_RETURN_IF(dev_alloc(h, &h->b_aoa, (size_t)2 * H));
      CAPDEC_RETURN_IF(scatter_rows(h->w_aoa, 2 * H, 2, 0, 0, h->W("attention.info_vector_proj.0.weight"), 2 * H, (int)H, (int)(2 * H), false, s));
      CAPDEC_RETURN_IF(scatter_rows(h->w_aoa, 2 * H, 2, 1, 0, h->W("attention.info_gate_proj.0.weight"), 2 * H, (int)H, (int)(2 * H), false, s));
      CAPDEC_RETURN_IF(scatter_rows(h->b_aoa, 1, 2, 0, 0, h->W("attention.info_vector_proj.0.bias"), 1, (int)H, 1, false, s));
      CAPDEC_RETURN_IF(scatter_rows(h->b_aoa, 1, 2, 1, 0, h->W("attention.info_gate_proj.0.bias"), 1, (int)H, 1, false, s));
    }
    if (c.attention == CAPDEC_ATT_ADAPTIVE) {
      CAPDEC_RETURN_IF(need(h, "attention.sentinel_gate.weight", {H, 2 * H})); CAPDEC_RETURN_IF(need(h, "attention.sentinel_gate.bias", {H}));
      CAPDEC_RETURN_IF(need(h, "attention.sentinel_proj.weight", {H, H}));     CAPDEC_RETURN_IF(need(h, "attention.sentinel_proj.bias", {H}));
      CAPDEC_RETURN_IF(need(h, "attention.adaptive_weight.weight", {1, 2 * H})); CAPDEC_RETURN_IF(need(h, "attention.adaptive_weight.bias", {1}));
      CAPDEC_CHECK_CUDA(cudaMemcpyAsync(&h->adaptive_bias, h->W("attention.adaptive_weight.bias"), sizeof(float), cudaMemcpyDeviceToHost, s));
    }
  }
  CAPDEC_RETURN_IF(gemm_tc_prepare(h, s));
  CAPDEC_CHECK_CUDA(cudaStreamSynchronize(s));
  h->finalized = true;
  return CAPDEC_OK;
}

size_t capdec_workspace_bytes(const capdec_handle* h, int32_t B, int32_t L, int32_t k, int32_t T) {
  if (!h) return 0;
  size_t need = 0;
  for (Mode m : {MODE_BEAM, MODE_GREEDY, MODE_SAMPLE, MODE_TEACHER}) {   // one workspace serves every entry point
    Arena ar(nullptr, 0);
    Session S;
    carve(h, ar, S, B, L, k, T, m);
    need = ar.off > need ? ar.off : need;
  }
  return align_up(need + 4096, 4096);
}

static int decode_beam_impl(capdec_handle* h, const float* feats, const TileSet* ext, const float* pooled, const uint8_t* mask,
                            int32_t B, int32_t L, int32_t k, int32_t T, float length_penalty, int32_t* out_tok, int32_t* out_len,
                            float* out_score, float* dbg_lp, int32_t* dbg_tok, int32_t* dbg_beam, void* ws, size_t ws_bytes,
                            void* stream) {
  CAPDEC_RETURN_IF(check_common(h, ext ? (const void*)ext : (const void*)feats, pooled, B, L, k, T));
  CAPDEC_REQUIRE(B == 0 || (out_tok && out_len && out_score), CAPDEC_ERR_INVALID, "output pointers must not be null");
  CAPDEC_REQUIRE(is_legacy(h) ? mask == nullptr : true, CAPDEC_ERR_UNSUPPORTED, "legacy decoder takes no padding mask");
  CAPDEC_REQUIRE(is_gpt2(h) ? mask == nullptr : true, CAPDEC_ERR_UNSUPPORTED,
                 "GPT-2 decoder reads only pooled_features (decoders.py:629-637): a region padding mask has nothing to mask");
  cudaStream_t s = (cudaStream_t)stream;
  Arena ar(ws, ws_bytes);
  Session S;
  if (ext) S.ext = *ext;
  CAPDEC_REQUIRE(carve(h, ar, S, B, L, k, T, MODE_BEAM) == CAPDEC_OK, CAPDEC_ERR_UNSUPPORTED,
                 "p24 tile set cannot be decoded with %d rows per image (no streaming-attention plan); decode the fp32 features instead", k);
  CAPDEC_REQUIRE(ws != nullptr && ar.ok(), CAPDEC_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", ar.off, ws_bytes);
  if (B == 0) return CAPDEC_OK;
  const capdec_config& c = h->cfg;
  CAPDEC_RETURN_IF(prologue_any(h, S, feats, pooled, s));
  // HF: output_fill_value = pad_token_id or eos_token_id  (generation/utils.py:3187)
  const int fill = c.pad_token_id ? c.pad_token_id : c.eos_token_id;
  CAPDEC_RETURN_IF(beam_init(S.beam, B, k, T, c.bos_token_id, fill, s));
  CAPDEC_RETURN_IF(fill_i32(S.next_tok, S.R, c.bos_token_id, s));
  CAPDEC_RETURN_IF(commit(h, S, nullptr, nullptr, 0, -1, false, s));
  const int k2 = 2 * k;
  for (int cur_len = 1; cur_len < T; ++cur_len) {
    CAPDEC_RETURN_IF(step_any(h, S, feats, mask, nullptr, 0, cur_len - 1, s));
    // prompt length is 1 (BOS): finished score / (cur_len+1-1)^lp ; heuristic uses ((cur_len+1)-1)^lp
    const float div_fin = (float)pow((double)cur_len, (double)length_penalty);
    const float div_heur = div_fin;
    const size_t o = (size_t)(cur_len - 1) * B * k2;
    const bool more = cur_len + 1 < T;
    if (S.fuse_k > 0 && !is_tf_family(h) && B <= kFusedSelectMaxImages) {
      // small batches: candidate merge + beam bookkeeping + state reorder / embedding gather of an image in one CTA
      GatherArgs ga{};
      if (more) build_gather(h, S, S.src_row, nullptr, 0, -1, true, &ga);
      StageScope sc(h, STAGE_BEAM, s);
      CAPDEC_RETURN_IF(select_fused(S.tk_part, S.tk_lse, c.vocab_size, S.tk_ntotal, S.fuse_k, S.beam, B, k, T, cur_len,
                                    c.eos_token_id, div_fin, div_heur, S.next_tok, S.src_row, dbg_lp ? dbg_lp + o : nullptr,
                                    dbg_tok ? dbg_tok + o : nullptr, dbg_beam ? dbg_beam + o : nullptr, more ? &ga : nullptr, s));
      continue;
    }
    { StageScope sc(h, STAGE_SELECT, s);
      CAPDEC_RETURN_IF(select_topk(h, S, k2, S.cand_lp, S.cand_idx, s)); }
    { StageScope sc(h, STAGE_BEAM, s);
      CAPDEC_RETURN_IF(beam_step(S.beam, B, k, T, c.vocab_size, cur_len, c.eos_token_id, div_fin, div_heur, S.cand_lp,
                                 S.cand_idx, S.next_tok, S.src_row, dbg_lp ? dbg_lp + o : nullptr,
                                 dbg_tok ? dbg_tok + o : nullptr, dbg_beam ? dbg_beam + o : nullptr, s)); }
    if (more) CAPDEC_RETURN_IF(commit(h, S, S.src_row, nullptr, 0, -1, true, s, cur_len - 1));
  }
  CAPDEC_RETURN_IF(beam_finalize(S.beam, (T - 1) & 1, B, k, T, out_tok, out_len, out_score, s));
  return CAPDEC_OK;
}

int capdec_decode_beam(capdec_handle* h, const float* feats, const float* pooled, const uint8_t* mask, int32_t B,
                       int32_t L, int32_t k, int32_t T, float length_penalty, int32_t* out_tok, int32_t* out_len,
                       float* out_score, float* dbg_lp, int32_t* dbg_tok, int32_t* dbg_beam, void* ws, size_t ws_bytes,
                       void* stream) {
  return decode_beam_impl(h, feats, nullptr, pooled, mask, B, L, k, T, length_penalty, out_tok, out_len, out_score, dbg_lp,
                          dbg_tok, dbg_beam, ws, ws_bytes, stream);
}

// ---- encoder -> decoder feature hand-off -------------------------------------------------------------------
size_t capdec_source_bytes(int32_t layout, int32_t dtype, int32_t B, int32_t L, int32_t D) {
  if (B < 0 || L < 1 || D < 1) return 0;
  return (size_t)B * ingest_source_image_bytes(layout, dtype, L, D);
}

size_t capdec_tiles_bytes(const capdec_handle* h, int32_t B, int32_t L) {
  if (!h || B < 0 || L < 1) return 0;
  return tiles_layout(h, nullptr, B, L).bytes;
}

int capdec_ingest_features(capdec_handle* h, const void* src, int32_t layout, int32_t dtype, int32_t B, int32_t L,
                           void* tiles, size_t tiles_bytes, void* stream) {
  CAPDEC_REQUIRE(h != nullptr && h->finalized, CAPDEC_ERR_STATE, "capdec_ingest_features: handle not finalized");
  CAPDEC_REQUIRE(!is_gpt2(h), CAPDEC_ERR_UNSUPPORTED, "GPT-2 decoder consumes pooled_features only (decoders.py:629-637)");
  CAPDEC_REQUIRE(B >= 0 && L >= 1, CAPDEC_ERR_INVALID, "bad sizes B=%d L=%d", B, L);
  const TileSet t = tiles_layout(h, tiles, B, L);
  CAPDEC_REQUIRE(B == 0 || (tiles != nullptr && tiles_bytes >= t.bytes && (((uintptr_t)tiles) & 255) == 0), CAPDEC_ERR_WORKSPACE,
                 "tile buffer too small or not 256-byte aligned: need %zu bytes, got %zu", t.bytes, tiles_bytes);
  IngestArgs a{};
  a.src = src; a.layout = layout; a.dtype = dtype; a.B = B; a.L = L; a.D = h->cfg.feature_dim;
  if (t.p24) { a.split = SplitDst{t.hi, t.lo, h->cfg.feature_dim, KIND_BF16, t.b8}; a.mean = t.mean; }
  else a.out_f32 = t.f32;
  StageScope sc(h, STAGE_PROLOGUE, (cudaStream_t)stream);
  return ingest_features(a, (cudaStream_t)stream);
}

int capdec_decode_beam_tiles(capdec_handle* h, const void* tiles, const float* pooled, const uint8_t* mask, int32_t B, int32_t L,
                             int32_t k, int32_t T, float length_penalty, int32_t* out_tok, int32_t* out_len, float* out_score,
                             float* dbg_lp, int32_t* dbg_tok, int32_t* dbg_beam, void* ws, size_t ws_bytes, void* stream) {
  CAPDEC_REQUIRE(h != nullptr && h->finalized, CAPDEC_ERR_STATE, "capdec_decode_beam_tiles: handle not finalized");
  CAPDEC_REQUIRE(B == 0 || tiles != nullptr, CAPDEC_ERR_INVALID, "tiles pointer is null");
  const TileSet t = tiles_layout(h, const_cast<void*>(tiles), B, L);
  if (!t.p24)
    return decode_beam_impl(h, t.f32, nullptr, pooled, mask, B, L, k, T, length_penalty, out_tok, out_len, out_score, dbg_lp,
                            dbg_tok, dbg_beam, ws, ws_bytes, stream);
  return decode_beam_impl(h, nullptr, &t, pooled, mask, B, L, k, T, length_penalty, out_tok, out_len, out_score, dbg_lp, dbg_tok,
                          dbg_beam, ws, ws_bytes, stream);
}

int capdec_pack_p24_host(const float* src, int64_t num_images, int64_t elems, void* dst) {
  CAPDEC_REQUIRE(src && dst && num_images >= 0 && elems >= 0, CAPDEC_ERR_INVALID, "capdec_pack_p24_host: bad argument");
  for (int64_t b = 0; b < num_images; ++b) {
    const uint32_t* in = reinterpret_cast<const uint32_t*>(src) + b * elems;
    uint16_t* hi = reinterpret_cast<uint16_t*>(reinterpret_cast<char*>(dst) + b * elems * 3);
    uint8_t* q = reinterpret_cast<uint8_t*>(hi + elems);
    for (int64_t i = 0; i < elems; ++i) {
      const uint32_t bits = in[i];
      hi[i] = (uint16_t)(bits >> 16);
      q[i] = (uint8_t)((((bits & 0xffffu) + 128u) * 65281u) >> 24);   // == p24_q_ in common.cuh
    }
  }
  return CAPDEC_OK;
}

int capdec_trim_at_eos(const int32_t* tok, int64_t ld, int32_t rows, int32_t T, int32_t eos, int32_t pad, int32_t keep_eos,
                       int32_t* out, int64_t ld_out, int32_t* out_len, void* stream) {
  CAPDEC_REQUIRE(tok && rows >= 0 && T >= 1 && (out || out_len), CAPDEC_ERR_INVALID, "capdec_trim_at_eos: bad argument");
  return trim_at_eos(tok, ld, rows, T, eos, pad, keep_eos, out, ld_out, out_len, (cudaStream_t)stream);
}

int capdec_forward_tokens(capdec_handle* h, const float* feats, const float* pooled, const uint8_t* mask, int32_t B, int32_t L,
                          int32_t k, const int32_t* tokens, int32_t tok_stride, int32_t n_tok, float* logits_out,
                          float* logprob_out, float* alpha_out, int32_t mask_pad_keys, void* ws, size_t ws_bytes, void* stream) {
  const int T = n_tok + 1;   // cache / workspace sized like a decode of n_tok generated positions
  CAPDEC_RETURN_IF(check_common(h, feats, pooled, B, L, k, T));
  CAPDEC_REQUIRE(tokens && n_tok >= 1 && tok_stride >= n_tok, CAPDEC_ERR_INVALID, "capdec_forward_tokens: bad token block");
  CAPDEC_REQUIRE(is_legacy(h) || is_gpt2(h) ? mask == nullptr : true, CAPDEC_ERR_UNSUPPORTED, "this decoder takes no region padding mask");
  CAPDEC_REQUIRE(!alpha_out || !is_tf_family(h), CAPDEC_ERR_UNSUPPORTED,
                 "attention weights are not exposed by the transformer decoders (decoders.py:435)");
  cudaStream_t s = (cudaStream_t)stream;
  Arena ar(ws, ws_bytes);
  Session S;
  carve(h, ar, S, B, L, k, T, MODE_SAMPLE);
  CAPDEC_REQUIRE(ws != nullptr && ar.ok(), CAPDEC_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", ar.off, ws_bytes);
  if (B == 0) return CAPDEC_OK;
  const capdec_config& c = h->cfg;
  const int V = c.vocab_size;
  if (mask_pad_keys && is_tf_family(h)) { S.key_tok = tokens; S.ld_key_tok = tok_stride; S.key_pad = c.pad_token_id; }
  CAPDEC_RETURN_IF(prologue_any(h, S, feats, pooled, s));
  float* const logits_buf = S.logits;
  const int64_t ld_buf = S.logits_ld ? S.logits_ld : V;
  const bool direct = logits_out != nullptr && V % 4 == 0;   // step t's GEMM writes straight into logits_out[:, t, :]
  for (int t = 0; t < n_tok; ++t) {
    CAPDEC_CHECK_CUDA(cudaMemcpy2DAsync(S.next_tok, sizeof(int32_t), tokens + t, (size_t)tok_stride * sizeof(int32_t),
                                        sizeof(int32_t), S.R, cudaMemcpyDeviceToDevice, s));
    CAPDEC_RETURN_IF(commit(h, S, nullptr, nullptr, 0, -1, t > 0, s));
    if (direct) { S.logits = logits_out + (size_t)t * V; S.logits_ld = (int64_t)n_tok * V; }
    else        { S.logits = logits_buf; S.logits_ld = ld_buf; }
    CAPDEC_RETURN_IF(step_any(h, S, feats, mask, alpha_out ? alpha_out + (size_t)t * L : nullptr, (int64_t)n_tok * L, t, s));
    const int64_t ld = S.logits_ld ? S.logits_ld : V;
    if (logits_out && !direct)
      CAPDEC_CHECK_CUDA(cudaMemcpy2DAsync(logits_out + (size_t)t * V, (size_t)n_tok * V * sizeof(float), logits_buf,
                                          (size_t)ld_buf * sizeof(float), (size_t)V * sizeof(float), S.R, cudaMemcpyDeviceToDevice, s));
    if (logprob_out && t + 1 < n_tok) {
      StageScope sc(h, STAGE_SELECT, s);
      CAPDEC_RETURN_IF(token_logprob(S.logits, ld, S.R, V, tokens + t + 1, tok_stride, logprob_out + t, n_tok - 1, s));
    }
  }
  return CAPDEC_OK;
}

int capdec_decode_greedy(capdec_handle* h, const float* feats, const float* pooled, const uint8_t* mask, int32_t B,
                         int32_t L, int32_t T, int32_t start_token_id, int32_t* out_tok, float* out_alpha, void* ws,
                         size_t ws_bytes, void* stream) {
  CAPDEC_RETURN_IF(check_common(h, feats, pooled, B, L, 1, T));
  CAPDEC_REQUIRE(out_tok, CAPDEC_ERR_INVALID, "output pointer must not be null");
  cudaStream_t s = (cudaStream_t)stream;
  Arena ar(ws, ws_bytes);
  Session S;
  carve(h, ar, S, B, L, 1, T, MODE_GREEDY);
  CAPDEC_REQUIRE(ws != nullptr && ar.ok(), CAPDEC_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", ar.off, ws_bytes);
  if (B == 0) return CAPDEC_OK;
  CAPDEC_RETURN_IF(prologue_any(h, S, feats, pooled, s));
  CAPDEC_RETURN_IF(fill_i32(S.next_tok, S.R, start_token_id, s));
  CAPDEC_RETURN_IF(commit(h, S, nullptr, out_tok, T, 0, false, s));  // out[:,0] = start (decoders.py:271)
  // LSTMDecoder.generate evaluates max_length steps and discards the last argmax (decoders.py:269-306);
  // TransformerDecoder.generate runs max_length-1 steps and keeps every token (decoders.py:461-487)
  const int n_steps = is_tf_family(h) ? T - 1 : T;
  for (int t = 0; t < n_steps; ++t) {
    float* alpha = out_alpha ? out_alpha + (size_t)t * L : nullptr;
    CAPDEC_RETURN_IF(step_any(h, S, feats, mask, alpha, (int64_t)T * L, t, s));
    if (t + 1 == T) break;  // the last argmax is discarded (decoders.py:306 after the final store at :271)
    { StageScope sc(h, STAGE_SELECT, s);
      CAPDEC_RETURN_IF(select_topk(h, S, 1, S.cand_lp, S.next_tok, s)); }
    CAPDEC_RETURN_IF(commit(h, S, nullptr, out_tok, T, t + 1, true, s));
  }
  return CAPDEC_OK;
}

int capdec_decode_sample(capdec_handle* h, const float* feats, const float* pooled, const uint8_t* mask, int32_t B,
                         int32_t L, int32_t num_samples, int32_t with_greedy, int32_t T, const float* uniforms,
                         int32_t* out_tok, float* out_lp, void* ws, size_t ws_bytes, void* stream) {
  const int k = num_samples + (with_greedy ? 1 : 0);
  CAPDEC_RETURN_IF(check_common(h, feats, pooled, B, L, k, T));
  CAPDEC_REQUIRE(num_samples >= 0 && out_tok && (num_samples == 0 || uniforms), CAPDEC_ERR_INVALID, "bad sampling arguments");
  cudaStream_t s = (cudaStream_t)stream;
  Arena ar(ws, ws_bytes);
  Session S;
  carve(h, ar, S, B, L, k, T, MODE_SAMPLE);
  CAPDEC_REQUIRE(ws != nullptr && ar.ok(), CAPDEC_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", ar.off, ws_bytes);
  if (B == 0) return CAPDEC_OK;
  const capdec_config& c = h->cfg;
  CAPDEC_RETURN_IF(prologue_any(h, S, feats, pooled, s));
  CAPDEC_RETURN_IF(fill_i32(S.next_tok, S.R, c.bos_token_id, s));
  CAPDEC_RETURN_IF(commit(h, S, nullptr, out_tok, T, 0, false, s));
  for (int t = 0; t + 1 < T; ++t) {  // trainer.py:413
    CAPDEC_RETURN_IF(step_any(h, S, feats, mask, nullptr, 0, t, s));
    { StageScope sc(h, STAGE_SELECT, s);
      if (S.samp_lse)
        CAPDEC_RETURN_IF(sample_rows_partials(S.logits, S.logits_ld ? S.logits_ld : c.vocab_size, S.R, c.vocab_size, S.samp_lse, uniforms,
                                              T - 1, t, k, with_greedy ? k - 1 : -1, S.next_tok, S.step_lp, s));
      else
        CAPDEC_RETURN_IF(sample_rows(S.logits, S.logits_ld ? S.logits_ld : c.vocab_size, S.R, c.vocab_size, uniforms, T - 1, t, k,
                                     with_greedy ? k - 1 : -1, S.next_tok, S.step_lp, s)); }
    if (out_lp) {
      // out_lp[r, t] = step_lp[r]
      CAPDEC_CHECK_CUDA(cudaMemcpy2DAsync(out_lp + t, (size_t)(T - 1) * sizeof(float), S.step_lp, sizeof(float),
                                          sizeof(float), S.R, cudaMemcpyDeviceToDevice, s));
    }
    CAPDEC_RETURN_IF(commit(h, S, nullptr, out_tok, T, t + 1, t + 2 < T, s));
  }
  return CAPDEC_OK;
}

int capdec_forward_teacher(capdec_handle* h, const float* feats, int32_t B, int32_t L, const int32_t* captions,
                           int32_t cap_stride, const int32_t* dec_len, float* preds, float* alphas, void* ws,
                           size_t ws_bytes, void* stream) {
  CAPDEC_RETURN_IF(check_common(h, feats, feats, B, L, 1, 2));
  CAPDEC_REQUIRE(is_legacy(h), CAPDEC_ERR_UNSUPPORTED, "capdec_forward_teacher mirrors models/decoder.py::Decoder.forward only");
  CAPDEC_REQUIRE(captions && dec_len && preds, CAPDEC_ERR_INVALID, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  int T = 0;
  for (int b = 0; b < B; ++b) {
    CAPDEC_REQUIRE(b == 0 || dec_len[b] <= dec_len[b - 1], CAPDEC_ERR_INVALID,
                   "captions must be sorted by decreasing length (data_loader.py:65 collate_fn)");
    T = dec_len[b] > T ? dec_len[b] : T;
  }
  CAPDEC_REQUIRE(T >= 1 && T <= cap_stride, CAPDEC_ERR_INVALID, "bad caption lengths");
  Arena ar(ws, ws_bytes);
  Session S;
  carve(h, ar, S, B, L, 1, T, MODE_TEACHER);
  CAPDEC_REQUIRE(ws != nullptr && ar.ok(), CAPDEC_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", ar.off, ws_bytes);
  if (B == 0) return CAPDEC_OK;
  const capdec_config& c = h->cfg;
  const int H = c.hidden_dim, E = c.embed_dim, D = c.feature_dim, V = c.vocab_size;
  CAPDEC_RETURN_IF(prologue_legacy(h, S, feats, true, s));
  for (int t = 0; t < T; ++t) {
    int bt = 0;
    for (int b = 0; b < B; ++b) bt += dec_len[b] > t ? 1 : 0;  // models/decoder.py:149
    // next_tok[r] = captions[r, t]
    CAPDEC_CHECK_CUDA(cudaMemcpy2DAsync(S.next_tok, sizeof(int32_t), captions + t, (size_t)cap_stride * sizeof(int32_t),
                                        sizeof(int32_t), bt, cudaMemcpyDeviceToDevice, s));
    GatherArgs g{};
    g.rows = bt; g.tok = S.next_tok; g.src = nullptr; g.embedding = h->W("embedding.weight"); g.E = E;
    g.x_emb = S.X[0]; g.ld_x = S.ldX[0]; g.pos = -1; g.n_state = 0;
    if (t > 0) {
      g.state_src[0] = S.hnew[0]; g.ld_src[0] = H; g.state_dst[0] = S.X[0] + E + D; g.ld_dst[0] = S.ldX[0]; g.width[0] = H;
      g.state_src[1] = S.cnew[0]; g.ld_src[1] = H; g.state_dst[1] = S.c[0]; g.ld_dst[1] = H; g.width[1] = H;
      g.n_state = 2;
    }
    CAPDEC_RETURN_IF(gather_rows(g, s));
    CAPDEC_RETURN_IF(step_legacy(h, S, feats, bt, preds + (size_t)t * V, (int64_t)T * V,
                                 alphas ? alphas + (size_t)t * L : nullptr, (int64_t)T * L, s));
  }
  return CAPDEC_OK;
}

int capdec_attention_forward(capdec_handle* h, const float* query, const float* feats, const uint8_t* mask,
                             const float* memory, const float* cell, int32_t B, int32_t L, int32_t k, float* context,
                             float* weights, void* ws, size_t ws_bytes, void* stream) {
  CAPDEC_RETURN_IF(check_common(h, feats, feats, B, L, k, 2));
  CAPDEC_REQUIRE(!is_legacy(h), CAPDEC_ERR_UNSUPPORTED, "capdec_attention_forward mirrors src/models/attention.py");
  CAPDEC_REQUIRE(query && context, CAPDEC_ERR_INVALID, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  Arena ar(ws, ws_bytes);
  Session S;
  carve(h, ar, S, B, L, k, 2, MODE_ATTENTION);
  CAPDEC_REQUIRE(ws != nullptr && ar.ok(), CAPDEC_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", ar.off, ws_bytes);
  if (B == 0) return CAPDEC_OK;
  const int H = h->cfg.hidden_dim;
  CAPDEC_RETURN_IF(prologue_attention(h, S, feats, s));
  return run_attention(h, S, feats, mask, query, H, memory, H, cell, H, B, context, weights, L, s);
}

static int decode_beam_host_impl(capdec_handle* h, const void* feats_host, int layout, int dtype, const float* pooled_host,
                                 const uint8_t* mask_host, int32_t B, int32_t L, int32_t k, int32_t T, float length_penalty,
                                 int32_t chunk, int32_t* out_tok_host, int32_t* out_len_host, float* out_score_host) {
  CAPDEC_RETURN_IF(check_common(h, feats_host, pooled_host, B, L, k, T));
  CAPDEC_REQUIRE(out_tok_host && out_len_host && out_score_host, CAPDEC_ERR_INVALID, "output pointers must not be null");
  CAPDEC_REQUIRE(layout >= CAPDEC_LAYOUT_BLD && layout <= CAPDEC_LAYOUT_CLS_BLD && dtype >= CAPDEC_DT_F32 && dtype <= CAPDEC_DT_P24,
                 CAPDEC_ERR_INVALID, "unknown feature layout %d / dtype %d", layout, dtype);
  if (B == 0) return CAPDEC_OK;
  const capdec_config& c = h->cfg;
  if (chunk <= 0) {
    // default: whole rounds of the persistent attention kernel.  fp32 sources are PCIe-bound (6.6 GB per 4096 images at
    // ~55 GB/s): two images per SM keep the exposed first copy / last decode short (124 ms vs 128 ms at four per SM).
    // bf16 / fp16 / p24 sources halve the copy, so the call becomes decode-bound and larger chunks -- more efficient
    // GEMM waves per step -- win: measured 84.0 / 74.5 / 74.6 / 76.4 ms at 2 / 4 / 6 / 8 images per SM for bf16.
    chunk = (dtype == CAPDEC_DT_F32 ? 2 : 4) * num_sms();
  }
  if (chunk > B) chunk = B;
  // One in-flight call per handle: the GEMM activation scratch and the lazily split weights are handle state, and this
  // call runs on private streams that have no ordering with the caller's.  Draining the device first makes a preceding
  // asynchronous decode on the caller's stream safe (the call is synchronous by contract anyway).
  CAPDEC_CHECK_CUDA(cudaDeviceSynchronize());
  const bool have_feats = feats_host != nullptr;
  const bool direct = layout == CAPDEC_LAYOUT_BLD && dtype == CAPDEC_DT_F32 && !tiles_p24(h);   // decode straight from the staged chunk
  const size_t img_src = have_feats ? ingest_source_image_bytes(layout, dtype, L, c.feature_dim) : 0;
  const int pooled_w = is_gpt2(h) ? c.feature_dim : c.hidden_dim;
  const size_t feat_chunk = align_up((size_t)chunk * img_src + 16, 256);
  const size_t pool_chunk = align_up((size_t)chunk * pooled_w * sizeof(float), 256);
  const size_t mask_chunk = mask_host ? align_up((size_t)chunk * L, 256) : 0;
  const size_t tiles_bytes = (have_feats && !direct) ? tiles_layout(h, nullptr, chunk, L).bytes : 0;
  // Chunk schedule: equal chunks.  (Ramping the chunk size up at the start and down at the end, to shorten the exposed
  // first copy / last decode, was measured slower on B200: small chunks decode too inefficiently for the ramp to pay.)
  std::vector<int> sizes;
  for (int left = B; left > 0;) { const int n = left < chunk ? left : chunk; sizes.push_back(n); left -= n; }
  // the fused top-k record count is not monotone in the row count, so size the shared workspace for every chunk size used
  size_t ws_bytes = 0;
  {
    std::vector<int> seen;
    for (int n : sizes) {
      bool dup = false;
      for (int m : seen) dup = dup || m == n;
      if (dup) continue;
      seen.push_back(n);
      const size_t w = capdec_workspace_bytes(h, n, L, k, T);
      ws_bytes = w > ws_bytes ? w : ws_bytes;
    }
  }
  const size_t out_bytes = align_up((size_t)B * T * 4, 256) + 2 * align_up((size_t)B * 4, 256);
  const size_t total = kHostBufs * (feat_chunk + pool_chunk + mask_chunk) + tiles_bytes + ws_bytes + out_bytes;
  if (h->stage_bytes < total) {
    if (h->stage_dev) CAPDEC_CHECK_CUDA(cudaFree(h->stage_dev));
    h->stage_dev = nullptr; h->stage_bytes = 0;
    CAPDEC_CHECK_CUDA(cudaMalloc(&h->stage_dev, total));
    h->stage_bytes = total;
  }
  if (!h->stream_compute) {
    CAPDEC_CHECK_CUDA(cudaStreamCreateWithFlags(&h->stream_compute, cudaStreamNonBlocking));
    CAPDEC_CHECK_CUDA(cudaStreamCreateWithFlags(&h->stream_copy, cudaStreamNonBlocking));
    for (int i = 0; i < kHostBufs; ++i) {
      CAPDEC_CHECK_CUDA(cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming));
      CAPDEC_CHECK_CUDA(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
    }
  }
  char* base = (char*)h->stage_dev;
  char* d_feat[kHostBufs];
  float* d_pool[kHostBufs];
  uint8_t* d_mask[kHostBufs];
  for (int i = 0; i < kHostBufs; ++i) {
    d_feat[i] = base + (size_t)i * feat_chunk;
    d_pool[i] = (float*)(base + (size_t)kHostBufs * feat_chunk + (size_t)i * pool_chunk);
    d_mask[i] = (uint8_t*)(base + (size_t)kHostBufs * (feat_chunk + pool_chunk) + (size_t)i * mask_chunk);
  }
  char* d_tiles = base + (size_t)kHostBufs * (feat_chunk + pool_chunk + mask_chunk);
  void* d_ws = d_tiles + tiles_bytes;
  int32_t* d_tok = (int32_t*)((char*)d_ws + ws_bytes);
  int32_t* d_len = (int32_t*)((char*)d_tok + align_up((size_t)B * T * 4, 256));
  float* d_score = (float*)((char*)d_len + align_up((size_t)B * 4, 256));
  auto run = [&]() -> int {
    int idx = 0, b0 = 0;
    for (size_t ci = 0; ci < sizes.size(); b0 += sizes[ci], ++ci, ++idx) {
      const int nb = sizes[ci];
      const int buf = idx % kHostBufs;   // three staging buffers: the copy engine never waits on the decode of the previous chunk
      if (idx >= kHostBufs) CAPDEC_CHECK_CUDA(cudaStreamWaitEvent(h->stream_copy, h->ev_done[buf], 0));
      if (have_feats)
        CAPDEC_CHECK_CUDA(cudaMemcpyAsync(d_feat[buf], (const char*)feats_host + (size_t)b0 * img_src, (size_t)nb * img_src,
                                          cudaMemcpyHostToDevice, h->stream_copy));
      if (pooled_host)
        CAPDEC_CHECK_CUDA(cudaMemcpyAsync(d_pool[buf], pooled_host + (size_t)b0 * pooled_w, (size_t)nb * pooled_w * sizeof(float),
                                          cudaMemcpyHostToDevice, h->stream_copy));
      if (mask_host)
        CAPDEC_CHECK_CUDA(cudaMemcpyAsync(d_mask[buf], mask_host + (size_t)b0 * L, (size_t)nb * L, cudaMemcpyHostToDevice, h->stream_copy));
      CAPDEC_CHECK_CUDA(cudaEventRecord(h->ev_copied[buf], h->stream_copy));
      CAPDEC_CHECK_CUDA(cudaStreamWaitEvent(h->stream_compute, h->ev_copied[buf], 0));
      const float* pool = pooled_host ? d_pool[buf] : nullptr;
      const uint8_t* msk = mask_host ? d_mask[buf] : nullptr;
      if (have_feats && !direct) {
        // hand-off pass on the device: the chunk arrives in the encoder's format and leaves as the tiles the decode streams
        CAPDEC_RETURN_IF(capdec_ingest_features(h, d_feat[buf], layout, dtype, nb, L, d_tiles, tiles_bytes, h->stream_compute));
        CAPDEC_RETURN_IF(capdec_decode_beam_tiles(h, d_tiles, pool, msk, nb, L, k, T, length_penalty, d_tok + (size_t)b0 * T,
                                                  d_len + b0, d_score + b0, nullptr, nullptr, nullptr, d_ws, ws_bytes, h->stream_compute));
      } else {
        CAPDEC_RETURN_IF(capdec_decode_beam(h, have_feats ? (const float*)d_feat[buf] : (const float*)d_feat[0], pool, msk, nb, L, k, T,
                                            length_penalty, d_tok + (size_t)b0 * T, d_len + b0, d_score + b0, nullptr, nullptr,
                                            nullptr, d_ws, ws_bytes, h->stream_compute));
      }
      CAPDEC_CHECK_CUDA(cudaEventRecord(h->ev_done[buf], h->stream_compute));
    }
    CAPDEC_CHECK_CUDA(cudaMemcpyAsync(out_tok_host, d_tok, (size_t)B * T * 4, cudaMemcpyDeviceToHost, h->stream_compute));
    CAPDEC_CHECK_CUDA(cudaMemcpyAsync(out_len_host, d_len, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream_compute));
    CAPDEC_CHECK_CUDA(cudaMemcpyAsync(out_score_host, d_score, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream_compute));
    return CAPDEC_OK;
  };
  const int st = run();
  // success or failure, nothing of this call is left in flight: copies from caller memory included
  const cudaError_t e1 = cudaStreamSynchronize(h->stream_copy);
  const cudaError_t e2 = cudaStreamSynchronize(h->stream_compute);
  CAPDEC_RETURN_IF(st);
  CAPDEC_CHECK_CUDA(e1);
  CAPDEC_CHECK_CUDA(e2);
  return CAPDEC_OK;
}

int capdec_decode_beam_host(capdec_handle* h, const float* feats_host, const float* pooled_host, int32_t B, int32_t L,
                            int32_t k, int32_t T, float length_penalty, int32_t chunk, int32_t* out_tok_host,
                            int32_t* out_len_host, float* out_score_host) {
  return decode_beam_host_impl(h, feats_host, CAPDEC_LAYOUT_BLD, CAPDEC_DT_F32, pooled_host, nullptr, B, L, k, T, length_penalty,
                               chunk, out_tok_host, out_len_host, out_score_host);
}

int capdec_decode_beam_host_ex(capdec_handle* h, const void* feats_host, int32_t layout, int32_t dtype, const float* pooled_host,
                               const uint8_t* mask_host, int32_t B, int32_t L, int32_t k, int32_t T, float length_penalty,
                               int32_t chunk, int32_t* out_tok_host, int32_t* out_len_host, float* out_score_host) {
  return decode_beam_host_impl(h, feats_host, layout, dtype, pooled_host, mask_host, B, L, k, T, length_penalty, chunk,
                               out_tok_host, out_len_host, out_score_host);
}

int capdec_stage_timing(capdec_handle* h, int32_t enable) {
  CAPDEC_REQUIRE(h, CAPDEC_ERR_INVALID, "null handle");
  h->timing = enable != 0;
  h->ev_used = 0;
  return CAPDEC_OK;
}

int capdec_stage_times(capdec_handle* h, float* ms_out, int32_t* count_out) {
  CAPDEC_REQUIRE(h && ms_out && count_out, CAPDEC_ERR_INVALID, "null argument");
  for (int i = 0; i < STAGE_COUNT; ++i) { ms_out[i] = 0.f; count_out[i] = 0; }
  for (size_t i = 0; i < h->ev_used; ++i) {
    CAPDEC_CHECK_CUDA(cudaEventSynchronize(h->ev_pool[2 * i + 1]));
    float ms = 0.f;
    CAPDEC_CHECK_CUDA(cudaEventElapsedTime(&ms, h->ev_pool[2 * i], h->ev_pool[2 * i + 1]));
    ms_out[h->ev_stage[i]] += ms;
    count_out[h->ev_stage[i]] += 1;
  }
  h->ev_used = 0;
  return CAPDEC_OK;
}

int capdec_linear(int32_t precision, const float* a, int64_t lda, const float* w, int64_t ldw, const float* bias,
                  float* c, int64_t ldc, int32_t m, int32_t n, int32_t k, void* stream) {
  GemmArgs g{};
  g.A = a; g.lda = lda; g.W = w; g.ldw = ldw; g.bias = bias; g.C = c; g.ldc = ldc; g.M = m; g.N = n; g.K = k;
  return gemm(nullptr, precision, g, EPI_STORE, (cudaStream_t)stream);
}

size_t capdec_linear_topk_workspace(int32_t m, int32_t n, int32_t topk) {
  if (m < 0 || !tk_supported(n, topk)) return 0;
  return align_up((size_t)m * tk_records(m, n) * tk_stride(topk) * sizeof(float), 256) +
         align_up((size_t)m * tk_lse_pairs(n) * 2 * sizeof(float), 256) + 256;
}

int capdec_linear_topk(int32_t precision, const float* a, int64_t lda, const float* w, int64_t ldw, const float* bias,
                       int32_t m, int32_t n, int32_t k, int32_t topk, float* out_lp, int32_t* out_idx, float* out_lse,
                       void* ws, size_t ws_bytes, void* stream) {
  CAPDEC_REQUIRE(precision != CAPDEC_PREC_FP32, CAPDEC_ERR_UNSUPPORTED,
                 "capdec_linear_topk: the fused epilogue exists on the tensor-core path only");
  CAPDEC_REQUIRE(tk_supported(n, topk), CAPDEC_ERR_UNSUPPORTED, "capdec_linear_topk: topk %d / vocab %d unsupported", topk, n);
  CAPDEC_REQUIRE(ws && ws_bytes >= capdec_linear_topk_workspace(m, n, topk), CAPDEC_ERR_WORKSPACE,
                 "capdec_linear_topk: workspace too small");
  GemmArgs g{};
  g.A = a; g.lda = lda; g.W = w; g.ldw = ldw; g.bias = bias; g.M = m; g.N = n; g.K = k;
  g.tk_part = (float*)ws; g.tk_k = topk; g.tk_vocab = n;
  g.tk_lse = (float*)((char*)ws + align_up((size_t)m * tk_records(m, n) * tk_stride(topk) * sizeof(float), 256));
  CAPDEC_RETURN_IF(gemm(nullptr, precision, g, EPI_TOPK, (cudaStream_t)stream));
  return topk_merge(g.tk_part, g.tk_lse, m, n, n, topk, topk, out_lp, out_idx, out_lse, (cudaStream_t)stream);
}

int capdec_lse_topk(const float* logits, int64_t ld, int32_t rows, int32_t vocab, int32_t topk, float* out_lp,
                    int32_t* out_idx, float* out_lse, void* stream) {
  return lse_topk(logits, ld, rows, vocab, topk, out_lp, out_idx, out_lse, (cudaStream_t)stream);
}

}  // extern "C"
