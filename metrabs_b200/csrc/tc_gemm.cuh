// tcgen05 tensor-core path (placeholder until the GEMM kernels land): nothing is eligible, so BF16_TC mode runs
// the CUDA-core kernels over bf16 storage.
#pragma once
#include <vector>

#include "common.cuh"
#include "conv_simt.cuh"
#include "decode.cuh"

namespace mtb {

struct TcWeights {
  bool ready = false;
};

inline bool tc_eligible(bool is_conv, bool depthwise, bool small_io, int k, int stride, int cin, int cout) { return false; }
inline const char* tc_prepare_weights(TcWeights&, const float*, const float*, int, int, int, int, int, std::vector<void*>&) { return nullptr; }
inline const char* tc_prepare_head(TcWeights&, const float*, const float*, int, int, std::vector<void*>&) { return nullptr; }
inline const char* tc_conv_launch(const TcWeights&, const ConvParams&, cudaStream_t) { return "not built"; }
inline const char* tc_head_launch(const TcWeights&, const void*, int, int, int, int, int, DecodeScale, float*, float*, cudaStream_t) { return "not built"; }

}  // namespace mtb
