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
  float* w_init = nullptr;  float* b_init = nullptr;   // legacy [2H,D]; lstm arch [2*H*layers, H] = [init_h ; init_c]
  // aoa: [info ; gate] rows interleaved n = 2*j + {info, gate}
  float* w_aoa = nullptr;   float* b_aoa = nullptr;
  float energy_bias = 0.f;    // att.bias / energy.bias (scalar, host copy)
  float adaptive_bias = 0.f;  // attention.adaptive_weight.bias

  // ---- host-API staging (capdec_decode_beam_host) ----
  void* stage_dev = nullptr; size_t stage_bytes = 0;
  cudaStream_t stream_compute = nullptr, stream_copy = nullptr;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};

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

// tensor-core GEMM modes (gemm_tc.cu)
int gemm_tc(int precision, const GemmArgs& a, int epilogue, cudaStream_t s);
int gemm_tc_prepare(capdec_handle* h, cudaStream_t s);

}  // namespace capdec
