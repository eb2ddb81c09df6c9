// nn_tc.cu — placeholder until the tcgen05 chain lands (next commit): reports "unsupported" so that
// AGPU_NN_BF16_TC contexts fail loudly instead of silently using another evaluator.
#include "nn.cuh"
namespace ag {
int tc_supported(int, int, int, int) { return 0; }
size_t tc_image_bytes(int, int, int, int) { return 0; }
void tc_build_image(const float*, const float* const*, const float*, const float*, const float*, const float*, int, int, int, int, void*, float*) {}
cudaError_t tc_forward(const NetDev&, const NNInput&, int, float*, int, cudaStream_t) { return cudaErrorNotSupported; }
}
