// nn.cuh — snetwork2 forward (DenseNet.jl:294-304) on the device.
//
// Two evaluators behind one interface (NetDev):
//   * nn_fp32_kernel  — CUDA-core fp32, the evaluation order of the reference formula (dot products with k
//                       ascending, separate multiply and add).  Bit-identical to the CPU oracle's fp32
//                       mode; this is the parity mode (AGPU_NN_FP32).
//   * the bf16 tcgen05/TMEM chain in nn_tc.cu (AGPU_NN_BF16_TC), the product path.
// Both read the leaf states straight from the tree records (the reference's `decoder` kernel and its
// fp32 `batch` tensor, mcts_gpu.jl:202-223, never touch HBM) and write logits[A] | value per game.
#pragma once
#include "common.cuh"

namespace ag {

struct NetDev {
  int in, n, k, A;
  // fp32, Julia column-major (what convert_back hands over, DenseNet.jl:331-333)
  const float* base;    // n x in
  const float* res;     // k x (n x n)
  const float* pol_w;   // A x n
  const float* pol_b;   // A
  const float* val_w;   // n
  const float* val_b;   // 1
  // bf16 operand images for the tensor-core chain (nn_tc.cu): per layer one K-major, 128B-swizzled
  // shared-memory image, ready for a 1-D bulk copy
  const void* tc_img;
  const float* tc_bias; // [NH] head biases (policy then value), zero padded
};

// where the evaluator finds its input
struct NNInput {
  const char* tree;        // node records, or null when x_direct is used
  size_t game_stride;
  int rec, off_state, nc;  // stride between the states of consecutive nodes of a game, offset of node 0's state in the game's block, 64-bit chunks per board
  int VS;
  const int32_t* leaf;     // [L] 0-based node per game
  const float* x_direct;   // [L][2VS] already encoded (agpu_forward)
  const int* seg;          // optional {off, len} in device memory: evaluate slots [off, off+len) (graph replay); else [0, L)
};

AG_D bool nn_input_bit(const NNInput& I, int g, int j) {
  const char* st = I.tree + (size_t)g * I.game_stride + (size_t)I.leaf[g] * I.rec + I.off_state;
  const u64* b = reinterpret_cast<const u64*>(st) + (j < I.VS ? 0 : I.nc);
  const int jj = j < I.VS ? j : j - I.VS;
  return (b[jj >> 6] >> (jj & 63)) & 1;
}

// blockDim.x == n; GT games per block
template <int GT>
__global__ void nn_fp32_kernel(NetDev net, NNInput I, int L, float* __restrict__ out, int outs) {
  extern __shared__ float sm[];
  const int n = net.n, in = net.in;
  float* xin = sm;             // [GT][in]
  float* b = sm + GT * in;     // [GT][n]
  int g0 = blockIdx.x * GT;
  if (I.seg) { const int off = I.seg[0], len = I.seg[1]; if (g0 >= len) return; L = off + len; g0 += off; }
  const int o = threadIdx.x;
  for (int t = o; t < GT * in; t += n) {
    const int gg = t / in, j = t % in, g = g0 + gg;
    float v = 0.f;
    if (g < L) v = I.x_direct ? I.x_direct[(size_t)g * in + j] : (nn_input_bit(I, g, j) ? 1.f : 0.f);
    xin[t] = v;
  }
  __syncthreads();
  float acc[GT];
#pragma unroll
  for (int gg = 0; gg < GT; gg++) acc[gg] = 0.f;
  for (int i = 0; i < in; i++) {                              // b = relu.(base*x)      DenseNet.jl:295
    const float w = net.base[o + (size_t)n * i];
#pragma unroll
    for (int gg = 0; gg < GT; gg++) acc[gg] = fadd(acc[gg], fmul(w, xin[gg * in + i]));
  }
#pragma unroll
  for (int gg = 0; gg < GT; gg++) b[gg * n + o] = fmaxf(acc[gg], 0.f);
  __syncthreads();
  for (int l = 0; l < net.k; l++) {                           // b .= relu.(b .+ relu.(w*b))   :297-299
    const float* w_l = net.res + (size_t)l * n * n;
#pragma unroll
    for (int gg = 0; gg < GT; gg++) acc[gg] = 0.f;
    for (int i = 0; i < n; i++) {
      const float w = w_l[o + (size_t)n * i];
#pragma unroll
      for (int gg = 0; gg < GT; gg++) acc[gg] = fadd(acc[gg], fmul(w, b[gg * n + i]));
    }
    __syncthreads();
#pragma unroll
    for (int gg = 0; gg < GT; gg++) b[gg * n + o] = fmaxf(fadd(b[gg * n + o], fmaxf(acc[gg], 0.f)), 0.f);
    __syncthreads();
  }
  if (o <= net.A) {                                           // policy*b .+ bias, σ.(value*b .+ bias)   :301
    const bool is_v = (o == net.A);
#pragma unroll
    for (int gg = 0; gg < GT; gg++) acc[gg] = 0.f;
    for (int i = 0; i < n; i++) {
      const float w = is_v ? net.val_w[i] : net.pol_w[o + (size_t)net.A * i];
#pragma unroll
      for (int gg = 0; gg < GT; gg++) acc[gg] = fadd(acc[gg], fmul(w, b[gg * n + i]));
    }
    const float bias = is_v ? net.val_b[0] : net.pol_b[o];
#pragma unroll
    for (int gg = 0; gg < GT; gg++) {
      const int g = g0 + gg;
      if (g < L) {
        const float z = fadd(acc[gg], bias);
        out[(size_t)g * outs + o] = is_v ? c_sigmoidf(z) : z;
      }
    }
  }
}

// launchers implemented in nn_tc.cu (tensor-core chain)
struct TcPlan;   // opaque per-network plan
int tc_supported(int in, int n, int k, int A);
size_t tc_image_bytes(int in, int n, int k, int A);
// builds the bf16 swizzled operand images from the fp32 column-major weights (host side), returns bytes written
// fmt: 0 = bf16 operands, 1 = fp16 operands (both kind::f16 MMAs with fp32 accumulation)
void tc_build_image(const float* base, const float* const* res, const float* pol_w, const float* pol_b, const float* val_w,
                    const float* val_b, int in, int n, int k, int A, void* img_host, float* bias_host, int fmt);
cudaError_t tc_forward(const NetDev& net, const NNInput& I, int L, float* out, int outs, cudaStream_t stream, int fmt);
cudaError_t tc_init();
extern long long* g_tc_dbg;

}  // namespace ag
