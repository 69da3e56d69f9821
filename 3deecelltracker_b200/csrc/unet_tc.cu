// tcgen05 implicit-GEMM convolution (placeholder until the tensor-core engine lands; the CUDA-core
// engine in unet_direct.cu takes every layer while this returns "unsupported").
#include "unet_common.cuh"

namespace ct {
size_t tc_weight_floats(int, int) { return 0; }
void tc_pack_weights(const float*, int, int, int, float*) {}
int launch_conv_tc(const CtUNet*, const Op&, float*, size_t, int, cudaStream_t) { return 2; }
}  // namespace ct
