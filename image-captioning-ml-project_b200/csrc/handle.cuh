// capdec_handle: bound weights, packed copies, and the per-call workspace carve-up.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "attention.cuh"
#include "select.cuh"

struct DevTensor {
  float* p = nullptr;
  std::vector<int64_t> shape;
  int64_t numel = 0;
};

constexpr int kHostBufs = 3;   // staging buffers of the host-buffer entry point

struct capdec_handle {
  capdec_config cfg{};
  std::map<std::string, DevTensor> w;  // reference state_dict name -> owned device copy
  std::vector<void*> owned;            // packed buffers (cudaFree on destroy)
  bool finalized = false;

  // ---- packed weights (built by capdec_finalize) ----
  // gate GEMM per LSTM layer: rows interleaved n = 4*j + {i,f,g,o}; columns = [input | h]
  std::vector<float*> w_gates, b_gates;
  std::vector<int> gate_in;            // input width of each layer (without the h part)
  // legacy: h -> [dec_att | f_beta] and mean(enc) -> [h_lin ; c_lin]
  float* w_hproj = nullptr; float* b_hproj = nullptr;
  // legacy, tensor-core beam/greedy path: [fc ; 0-pad to a 256 multiple ; dec_att ; f_beta] so that the vocabulary GEMM of
  // step t also emits [dec_att | sigmoid(f_beta)] of the new hidden state for step t+1 (same A operand)
  float* w_vocab_cat = nullptr; float* b_vocab_cat = nullptr; int vocab_cat_n = 0;
  // legacy, tensor-core modes: emb_gates[v, n] = sum_e embedding[v, e] * W_ih[n, e] in the interleaved gate order (exact fp32):
  // the embedding columns leave the gate GEMM's K, the LSTM epilogue adds the row of the step's token instead
  float* emb_gates = nullptr;
  float* w_init = nullptr;  float* b_init = nullptr;   // legacy [2H,D]; lstm arch [2*H*layers, H] = [init_h ; init_c]
  // aoa: [info ; gate] rows interleaved n = 2*j + {info, gate}
  float* w_aoa = nullptr;   float* b_aoa = nullptr;
  float energy_bias = 0.f;    // att.bias / energy.bias (scalar, host copy)
  float adaptive_bias = 0.f;  // attention.adaptive_weight.bias

  // ---- host-API staging (capdec_decode_beam_host) ----
  void* stage_dev = nullptr; size_t stage_bytes = 0;
  cudaStream_t stream_compute = nullptr, stream_copy = nullptr;
  cudaEvent_t ev_copied[kHostBufs] = {}, ev_done[kHostBufs] = {};

  // ---- optional per-stage device timing (capdec_stage_timing): cudaEvent pairs around every stage launch
  mutable bool timing = false;
  mutable std::vector<cudaEvent_t> ev_pool;      // [2*i] start, [2*i+1] stop
  mutable std::vector<int> ev_stage;             // stage id of pair i
  mutable size_t ev_used = 0;

  // ---- tensor-core GEMM state (gemm_tc.cu): split weight copies keyed by weight pointer, activation scratch
  mutable std::map<const float*, float*> tc_weights;
  mutable float* tc_scratch = nullptr;
  mutable size_t tc_scratch_bytes = 0;
  // stream-K: one accumulator-part slot per CTA of the persistent grid + epoch-tagged flags (8 epilogue warps per CTA)
  mutable float* sk_part = nullptr;
  mutable int* sk_flag = nullptr;
  mutable int sk_epoch = 0;

  const DevTensor* find(const std::string& n) const {
    auto it = w.find(n);
    return it == w.end() ? nullptr : &it->second;
  }
  const float* W(const std::string& n) const { return w.at(n).p; }
};

namespace capdec {

// bump allocator over the caller's workspace
struct Arena {
  char* base; size_t cap; size_t off = 0; bool dry;
  Arena(void* p, size_t bytes) : base((char*)p), cap(bytes), dry(p == nullptr) {}
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* r = dry ? nullptr : reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return r;
  }
  bool ok() const { return dry || off <= cap; }
};

enum Stage : int {
  STAGE_PROLOGUE = 0,   // hoisted region projections + init state
  STAGE_SMALL_GEMM = 1, // per-step query-side projections (dec_att|f_beta, query_proj, output_proj, AoA ...)
  STAGE_ATTENTION = 2,  // scores + softmax + context over the region tiles
  STAGE_GATE_GEMM = 3,  // LSTM gate GEMM with fused cell update
  STAGE_VOCAB_GEMM = 4, // vocabulary projection
  STAGE_SELECT = 5,     // log-softmax + top-k / argmax / sampling
  STAGE_BEAM = 6,       // per-image beam bookkeeping
  STAGE_GATHER = 7,     // back-pointer reorder + embedding gather
  STAGE_COUNT = 8
};

// RAII: record an event pair around the launches issued while in scope (no-op unless timing is enabled)
struct StageScope {
  const capdec_handle* h; cudaStream_t s; cudaEvent_t stop = nullptr;
  StageScope(const capdec_handle* h_, int stage, cudaStream_t s_) : h(h_), s(s_) {
    if (!h->timing) return;
    if (h->ev_used * 2 + 2 > h->ev_pool.size()) {
      cudaEvent_t a, b;
      cudaEventCreate(&a); cudaEventCreate(&b);
      h->ev_pool.push_back(a); h->ev_pool.push_back(b);
      h->ev_stage.push_back(stage);
    }
    h->ev_stage[h->ev_used] = stage;
    cudaEventRecord(h->ev_pool[h->ev_used * 2], s);
    stop = h->ev_pool[h->ev_used * 2 + 1];
    h->ev_used++;
  }
  ~StageScope() { if (stop) cudaEventRecord(stop, s); }
};

// tensor-core GEMM modes (gemm_tc.cu)
int gemm_tc(const capdec_handle* h, int precision, const GemmArgs& a, int epilogue, cudaStream_t s);
int gemm_tc_prepare(capdec_handle* h, cudaStream_t s);
int tc_kind(int precision);
int tc_terms(int precision);
int tc_split(int precision, const float* x, int64_t ld, int rows, int cols, void* hi, void* lo, cudaStream_t s);
void gemm_tc_release(capdec_handle* h);
int gemm(const capdec_handle* h, int precision, const GemmArgs& a, int epilogue, cudaStream_t s);

}  // namespace capdec
