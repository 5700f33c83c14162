// Encoder -> decoder feature hand-off (ingest.cu): one pass from what an encoder emits to what the decode kernels stream.
#pragma once
#include "common.cuh"

namespace capdec {

struct IngestArgs {
  const void* src;        // source features, layout / dtype below
  int layout;             // capdec_layout
  int dtype;              // capdec_dtype
  int B, L, D;
  float* out_f32;         // [B,L,D] dense fp32 (or nullptr)
  SplitDst split;         // hi / lo (/ p24 byte plane) operand copies, row = b*L + l, ld = D (hi == nullptr: none)
  float* mean;            // [B,D] mean over the L regions (or nullptr)
};
int ingest_features(const IngestArgs& a, cudaStream_t s);
// bytes one image occupies in a source format
size_t ingest_source_image_bytes(int layout, int dtype, int L, int D);

}  // namespace capdec
