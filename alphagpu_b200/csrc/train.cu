// train.cu — the training step behind include/alphagpu_train.h (reference train.jl:12-15,47-162 on a
// `networkf`, DenseNet.jl:161-198).
//
// Data layout (HBM, fp32, sample-major): X [B][in]; H_l [B][n] (l = 0..k, output of base / block l); R_l [B][n] (relu of the
// block's product, kept for the backward masks); Zh [B][NH] head pre-activations, NH = A + 1 + FS.  Parameters, gradient and
// Adam moments are ONE flat array each, in Flux.params order: base (n x in), res[0..k) (n x n), heads packed (NH x n) — rows
// policy, value, feature — and the NH head biases; every matrix Julia column-major, i.e. element (out o, in i) at o + outs*i,
// which is exactly the row-major [K = in][N = out] operand the forward product wants.
//
// Arithmetic contract (the CPU checker of the tests restates it bit for bit): every product is an fma chain ascending in k starting from
// +0; weight gradients are accumulated per slice of TRAIN_KSLICE = 256 samples and the slices added in ascending order;
// element-wise steps are single IEEE operations in the order written here; exp/sigmoid/tanh are the canonical c_expf forms;
// Adam + weight decay run per element in fp64 and round once per stored value, as Flux 0.12.6's broadcasts over Float64
// hyper-parameters do.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/alphagpu.h"
#include "../../include/alphagpu_train.h"
#include "common.cuh"

namespace agt {
using namespace ag;

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;
constexpr int TRAIN_KSLICE = 256;

// tanh from c_expf: odd polynomial below 1/4 (truncation < 1e-8 relative), 1 - 2/(e^{2|x|}+1) above; +-1 beyond 9
AG_HD float c_tanhf(float x) {
  const float ax = x < 0.f ? -x : x;
  float t;
  if (ax < 0.25f) {
    const float s = fmul(ax, ax);
    float p = 2.18694885361552e-2f;                       // 62/2835
    p = fadd(fmul(p, s), -5.39682539682540e-2f);          // -17/315
    p = fadd(fmul(p, s), 1.33333333333333e-1f);           // 2/15
    p = fadd(fmul(p, s), -3.33333333333333e-1f);          // -1/3
    p = fadd(fmul(p, s), 1.0f);
    t = fmul(ax, p);
  } else if (ax > 9.0f) {
    t = 1.0f;
  } else {
    const float e = c_expf(fmul(2.0f, ax));
    t = fsub(1.0f, fdiv(2.0f, fadd(e, 1.0f)));
  }
  return x < 0.f ? -t : t;
}

enum { EPI_STORE = 0, EPI_RELU = 1, EPI_RES = 2, EPI_BIAS = 3, EPI_BWD = 4 };

struct Epi {
  const float* bias;    // EPI_BIAS: [N]
  const float* hprev;   // EPI_RES: H_{l-1} [M][N]
  float* r_out;         // EPI_RES: R_l [M][N]
  const float* add;     // EPI_BWD: dS of the layer above [M][N] or null
  const float* hmask;   // EPI_BWD: H_l (gradient passes where H_l > 0)
  const float* rmask;   // EPI_BWD: R_l or null (base layer)
  float* ds_out;        // EPI_BWD: dS_l (only with rmask)
};

// C[i][j] = sum_k A(i,k) * B(k,j), k ascending, fma chain from +0.
//   AT = 0: A(i,k) = A[i*lda + k]      AT = 1: A(i,k) = A[k*lda + i]
//   BT = 0: B(k,j) = B[k*ldb + j]      BT = 1: B(k,j) = B[j*ldb + k]
// blockIdx.z = K slice [z*kslice, min(K, (z+1)*kslice)) written to C + z*c_slice (split-K partials, EPI_STORE only).
// 64x64 tile, 256 threads x (4x4), BK = 16, next tile prefetched into registers during the fma block.
template <int AT, int BT, int EPI>
__global__ void __launch_bounds__(256) gemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                                   float* __restrict__ C, int ldc, long long c_slice, int kslice, Epi e) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * kslice, kend = min(K, kbeg + kslice);
  C += (long long)blockIdx.z * c_slice;

  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      int i, kk;
      if (AT == 0) { i = (tid >> 4) + 16 * r; kk = tid & 15; } else { kk = (tid >> 6) + 4 * r; i = tid & 63; }
      const int gi = m0 + i, gk = k0 + kk;
      ra[r] = (gi < M && gk < kend) ? __ldg(AT == 0 ? A + (size_t)gi * lda + gk : A + (size_t)gk * lda + gi) : 0.f;
      int j;
      if (BT == 0) { kk = (tid >> 6) + 4 * r; j = tid & 63; } else { j = (tid >> 4) + 16 * r; kk = tid & 15; }
      const int gj = n0 + j, gk2 = k0 + kk;
      rb[r] = (gj < N && gk2 < kend) ? __ldg(BT == 0 ? B + (size_t)gk2 * ldb + gj : B + (size_t)gj * ldb + gk2) : 0.f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      if (AT == 0) As[tid & 15][(tid >> 4) + 16 * r] = ra[r]; else As[(tid >> 6) + 4 * r][tid & 63] = ra[r];
      if (BT == 0) Bs[(tid >> 6) + 4 * r][tid & 63] = rb[r]; else Bs[tid & 15][(tid >> 4) + 16 * r] = rb[r];
    }
  };

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  fetch(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    stash();
    __syncthreads();
    if (k0 + BK < kend) fetch(k0 + BK);
    // rows beyond kend inside the last tile hold zeros: fma(0, 0, acc) = acc exactly
    const int kn = min(BK, kend - k0);
#pragma unroll 4
    for (int kk = 0; kk < kn; kk++) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = __fmaf_rn(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int gi = m0 + ty * 4 + i;
    if (gi >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int gj = n0 + tx * 4 + j;
      if (gj >= N) continue;
      const size_t o = (size_t)gi * ldc + gj;
      const float v = acc[i][j];
      if (EPI == EPI_STORE) {
        C[o] = v;
      } else if (EPI == EPI_RELU) {
        C[o] = v > 0.f ? v : 0.f;
      } else if (EPI == EPI_RES) {                           // resnets training branch, DenseNet.jl:37-39
        const float r = v > 0.f ? v : 0.f;
        e.r_out[o] = r;
        const float sres = fadd(e.hprev[o], r);
        C[o] = sres > 0.f ? sres : 0.f;
      } else if (EPI == EPI_BIAS) {
        C[o] = fadd(v, e.bias[gj]);
      } else {                                               // EPI_BWD
        const float dh = e.add ? fadd(v, e.add[o]) : v;
        const float ds = e.hmask[o] > 0.f ? dh : 0.f;
        if (e.rmask) {
          e.ds_out[o] = ds;
          C[o] = e.rmask[o] > 0.f ? ds : 0.f;
        } else {
          C[o] = ds;
        }
      }
    }
  }
}

__global__ void cast_i8_kernel(const int8_t* __restrict__ x, float* __restrict__ y, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = (float)x[i];
}

// Heads + loss + their gradient for one sample per thread (train.jl:12-15).  Zh [B][NH] -> dZh [B][NH], terms [B][3].
__global__ void __launch_bounds__(128) loss_kernel(const float* __restrict__ Zh, const float* __restrict__ ypol, const float* __restrict__ yval,
                                                  const float* __restrict__ yfeat, float* __restrict__ dZh, float* __restrict__ terms, int B, int A,
                                                  int FS, float fweight, int want_grad) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int NH = A + 1 + FS;
  const float* z = Zh + (size_t)b * NH;
  float* dz = dZh + (size_t)b * NH;
  const float invB = fdiv(1.0f, (float)B);
  // policy: logitcrossentropy = mean_b( -sum_a y_a * logsoftmax(z)_a )
  float m = z[0];
  for (int a = 1; a < A; a++) m = z[a] > m ? z[a] : m;
  float s = 0.f;
  for (int a = 0; a < A; a++) s = fadd(s, c_expf(fsub(z[a], m)));
  const float ls = logf(s);                                  // loss value only (compared by tolerance)
  float lp = 0.f, sumy = 0.f;
  const float* y = ypol + (size_t)b * A;
  for (int a = 0; a < A; a++) {
    lp = fsub(lp, fmul(y[a], fsub(fsub(z[a], m), ls)));
    sumy = fadd(sumy, y[a]);
  }
  if (want_grad) {
    const float sy = fmul(sumy, invB);
    for (int a = 0; a < A; a++) dz[a] = fsub(fmul(fdiv(c_expf(fsub(z[a], m)), s), sy), fmul(y[a], invB));
  }
  // value: mse(sigmoid(z), r)
  const float v = c_sigmoidf(z[A]);
  const float dv = fsub(v, yval[b]);
  if (want_grad) dz[A] = fmul(fmul(fmul(2.0f, dv), invB), fmul(v, fsub(1.0f, v)));
  // feature: fweight * mse(tanh(z), f) over B*FS elements
  float lf = 0.f;
  const float invBF = fdiv(1.0f, fmul((float)B, (float)FS));
  const float* yf = yfeat + (size_t)b * FS;
  for (int j = 0; j < FS; j++) {
    const float f = c_tanhf(z[A + 1 + j]);
    const float df = fsub(f, yf[j]);
    lf = fadd(lf, fmul(df, df));
    if (want_grad) dz[A + 1 + j] = fmul(fmul(fmul(fmul(fweight, 2.0f), df), invBF), fsub(1.0f, fmul(f, f)));
  }
  terms[(size_t)b * 3 + 0] = lp;
  terms[(size_t)b * 3 + 1] = fmul(dv, dv);
  terms[(size_t)b * 3 + 2] = lf;
}

// out[4] = total, policy, value, feature; fp64 accumulation in a fixed order (one block)
__global__ void __launch_bounds__(256) loss_reduce_kernel(const float* __restrict__ terms, int B, int FS, float fweight, float* __restrict__ out) {
  __shared__ double sh[3][256];
  double a[3] = {0, 0, 0};
  for (int b = threadIdx.x; b < B; b += 256)
    for (int c = 0; c < 3; c++) a[c] += (double)terms[(size_t)b * 3 + c];
  for (int c = 0; c < 3; c++) sh[c][threadIdx.x] = a[c];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int c = 0; c < 3; c++) sh[c][threadIdx.x] += sh[c][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double lp = sh[0][0] / B, lv = sh[1][0] / B, lf = sh[2][0] / ((double)B * FS);
    out[0] = (float)(lp + lv + (double)fweight * lf);
    out[1] = (float)lp; out[2] = (float)lv; out[3] = (float)lf;
  }
}

// bias gradient partials: part[z][o] = sum over the slice's samples (ascending) of dZh[b][o]
__global__ void bias_partial_kernel(const float* __restrict__ dZh, int B, int NH, float* __restrict__ part, long long c_slice, int kslice) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  const int z = blockIdx.y;
  if (o >= NH) return;
  const int b0 = z * kslice, b1 = min(B, b0 + kslice);
  float s = 0.f;
  for (int b = b0; b < b1; b++) s = fadd(s, dZh[(size_t)b * NH + o]);
  part[(long long)z * c_slice + o] = s;
}

// g[p] = sum_z part[z][p], ascending from +0
__global__ void reduce_partials_kernel(const float* __restrict__ part, long long P, int nslices, float* __restrict__ g) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float s = 0.f;
  for (int z = 0; z < nslices; z++) s = fadd(s, part[(long long)z * P + p]);
  g[p] = s;
}

// Flux 0.12.6 Optimiser(ADAM(eta, (b1, b2)), WeightDecay(wd)): apply!(ADAM) then apply!(WeightDecay) then x .-= delta.
// Hyper-parameters are Float64, arrays Float32: each broadcast evaluates in Float64 and rounds on assignment.
__global__ void adam_kernel(float* __restrict__ x, const float* __restrict__ g, float* __restrict__ mt, float* __restrict__ vt, long long P,
                            float gscale, double eta, double b1, double b2, double bp1, double bp2, double eps, double wd) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float d = fmul(g[p], gscale);
  const float m = (float)(b1 * (double)mt[p] + (1.0 - b1) * (double)d);
  const float v = (float)(b2 * (double)vt[p] + (1.0 - b2) * (double)fmul(d, d));
  mt[p] = m; vt[p] = v;
  float delta = (float)((double)m / (1.0 - bp1) / (sqrt((double)v / (1.0 - bp2)) + eps) * eta);
  delta = (float)((double)delta + wd * (double)x[p]);
  x[p] = fsub(x[p], delta);
}

}  // namespace agt

using namespace agt;

struct agpu_trainer {
  agpu_train_config cfg;
  int NH = 0, maxB = 0;
  long long P = 0, off_res = 0, off_heads = 0, off_bias = 0;
  float *params = nullptr, *grads = nullptr, *mt = nullptr, *vt = nullptr, *part = nullptr;
  float *X = nullptr, *H = nullptr, *R = nullptr, *Zh = nullptr, *dZh = nullptr, *dS[2] = {nullptr, nullptr}, *dZ[2] = {nullptr, nullptr};
  float *ypol = nullptr, *yval = nullptr, *yfeat = nullptr, *terms = nullptr, *loss_dev = nullptr;
  int8_t *st_i8 = nullptr, *fs_i8 = nullptr;
  double betap[2] = {0, 0};
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  float last_ms[2] = {0, 0};
  long long launches = 0;
  std::string err;
  bool has_grad = false;
};

static thread_local std::string g_trainer_create_error;

#define TR_CK(call)                                                                                      \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) {                                                                             \
      char buf_[512];                                                                                    \
      snprintf(buf_, sizeof(buf_), "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      tr->err = buf_;                                                                                    \
      return AGPU_ERR_CUDA;                                                                              \
    }                                                                                                    \
  } while (0)
#define TR_REQUIRE(cond, code, msg) \
  do {                              \
    if (!(cond)) {                  \
      tr->err = msg;                \
      return code;                  \
    }                               \
  } while (0)

namespace {

template <int AT, int BT, int EPI>
cudaError_t gemm(agpu_trainer* tr, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, const Epi& e,
                 int nslices = 1, long long c_slice = 0, int kslice = 0) {
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, nslices);
  gemm_kernel<AT, BT, EPI><<<grid, 256, 0, tr->stream>>>(M, N, K, A, lda, B, ldb, C, ldc, c_slice, nslices == 1 ? K : kslice, e);
  tr->launches++;
  return cudaGetLastError();
}

int stage_batch(agpu_trainer* tr, const int8_t* state, const float* policy, const float* value, const int8_t* fstate, int64_t B) {
  const auto& c = tr->cfg;
  TR_REQUIRE(state && policy && value && fstate, AGPU_ERR_INVALID, "null batch pointer");
  TR_REQUIRE(B >= 1 && B <= tr->maxB, AGPU_ERR_INVALID, "batch size outside [1, max_batch]");
  TR_CK(cudaMemcpyAsync(tr->st_i8, state, (size_t)B * c.in, cudaMemcpyDefault, tr->stream));
  TR_CK(cudaMemcpyAsync(tr->ypol, policy, sizeof(float) * B * c.actions, cudaMemcpyDefault, tr->stream));
  TR_CK(cudaMemcpyAsync(tr->yval, value, sizeof(float) * B, cudaMemcpyDefault, tr->stream));
  TR_CK(cudaMemcpyAsync(tr->fs_i8, fstate, (size_t)B * c.fsize, cudaMemcpyDefault, tr->stream));
  const long long nx = (long long)B * c.in, nf = (long long)B * c.fsize;
  cast_i8_kernel<<<(unsigned)((nx + 255) / 256), 256, 0, tr->stream>>>(tr->st_i8, tr->X, nx);
  cast_i8_kernel<<<(unsigned)((nf + 255) / 256), 256, 0, tr->stream>>>(tr->fs_i8, tr->yfeat, nf);
  tr->launches += 2;
  TR_CK(cudaGetLastError());
  return AGPU_OK;
}

// networkf(x; training=true) (DenseNet.jl:171-186) + lossTot (train.jl:12-15)
int forward_loss(agpu_trainer* tr, int B, bool want_grad) {
  const auto& c = tr->cfg;
  const int n = c.width, k = c.blocks, NH = tr->NH;
  const size_t bn = (size_t)tr->maxB * n;
  Epi e{};
  TR_CK((gemm<0, 0, EPI_RELU>(tr, B, n, c.in, tr->X, c.in, tr->params, n, tr->H, n, e)));
  for (int l = 1; l <= k; l++) {
    Epi r{};
    r.hprev = tr->H + (l - 1) * bn;
    r.r_out = tr->R + (l - 1) * bn;
    TR_CK((gemm<0, 0, EPI_RES>(tr, B, n, n, tr->H + (l - 1) * bn, n, tr->params + tr->off_res + (size_t)(l - 1) * n * n, n, tr->H + l * bn, n, r)));
  }
  Epi hb{};
  hb.bias = tr->params + tr->off_bias;
  TR_CK((gemm<0, 0, EPI_BIAS>(tr, B, NH, n, tr->H + k * bn, n, tr->params + tr->off_heads, NH, tr->Zh, NH, hb)));
  loss_kernel<<<(B + 127) / 128, 128, 0, tr->stream>>>(tr->Zh, tr->ypol, tr->yval, tr->yfeat, tr->dZh, tr->terms, B, c.actions, c.fsize,
                                                       c.feature_weight, want_grad ? 1 : 0);
  loss_reduce_kernel<<<1, 256, 0, tr->stream>>>(tr->terms, B, c.fsize, c.feature_weight, tr->loss_dev);
  tr->launches += 2;
  TR_CK(cudaGetLastError());
  return AGPU_OK;
}

int backward(agpu_trainer* tr, int B) {
  const auto& c = tr->cfg;
  const int n = c.width, k = c.blocks, NH = tr->NH;
  const size_t bn = (size_t)tr->maxB * n;
  const int ns = (B + TRAIN_KSLICE - 1) / TRAIN_KSLICE;
  const long long P = tr->P;
  Epi none{};
  // heads: dW = dZh^T H_k per slice, db = column sums per slice
  TR_CK((gemm<1, 0, EPI_STORE>(tr, n, NH, B, tr->H + k * bn, n, tr->dZh, NH, tr->part + tr->off_heads, NH, none, ns, P, TRAIN_KSLICE)));
  bias_partial_kernel<<<dim3((NH + 127) / 128, ns), 128, 0, tr->stream>>>(tr->dZh, B, NH, tr->part + tr->off_bias, P, TRAIN_KSLICE);
  tr->launches++;
  // dH_k = dZh Wh, masked by block k (or by the base layer when there are no blocks)
  int cur = 0;
  {
    Epi e{};
    e.hmask = tr->H + k * bn;
    if (k > 0) { e.rmask = tr->R + (k - 1) * bn; e.ds_out = tr->dS[cur]; }
    TR_CK((gemm<0, 1, EPI_BWD>(tr, B, n, NH, tr->dZh, NH, tr->params + tr->off_heads, NH, tr->dZ[cur], n, e)));
  }
  for (int l = k; l >= 1; l--) {
    const float* W = tr->params + tr->off_res + (size_t)(l - 1) * n * n;
    // dW_l (i, o) = sum_b H_{l-1}[b][i] dZ_l[b][o]
    TR_CK((gemm<1, 0, EPI_STORE>(tr, n, n, B, tr->H + (l - 1) * bn, n, tr->dZ[cur], n, tr->part + tr->off_res + (size_t)(l - 1) * n * n, n, none, ns,
                                 P, TRAIN_KSLICE)));
    // dH_{l-1} = dS_l + dZ_l W_l, masked by layer l-1
    Epi e{};
    e.add = tr->dS[cur];
    e.hmask = tr->H + (l - 1) * bn;
    if (l - 1 >= 1) { e.rmask = tr->R + (l - 2) * bn; e.ds_out = tr->dS[cur ^ 1]; }
    TR_CK((gemm<0, 1, EPI_BWD>(tr, B, n, n, tr->dZ[cur], n, W, n, tr->dZ[cur ^ 1], n, e)));
    cur ^= 1;
  }
  // base: dW_0 (i, o) = sum_b X[b][i] dZ_0[b][o]
  TR_CK((gemm<1, 0, EPI_STORE>(tr, c.in, n, B, tr->X, c.in, tr->dZ[cur], n, tr->part, n, none, ns, P, TRAIN_KSLICE)));
  reduce_partials_kernel<<<(unsigned)((P + 255) / 256), 256, 0, tr->stream>>>(tr->part, P, ns, tr->grads);
  tr->launches++;
  TR_CK(cudaGetLastError());
  tr->has_grad = true;
  return AGPU_OK;
}

int read_loss(agpu_trainer* tr, float loss_out[4]) {
  float h[4];
  TR_CK(cudaMemcpyAsync(h, tr->loss_dev, sizeof(h), cudaMemcpyDeviceToHost, tr->stream));
  TR_CK(cudaStreamSynchronize(tr->stream));
  if (loss_out) memcpy(loss_out, h, sizeof(h));
  return AGPU_OK;
}

// flat <-> Flux.params arrays.  dir = 0: arrays -> flat (host to device); 1: flat -> arrays.
int exchange(agpu_trainer* tr, float* flat_dev, int dir, float* base, float* const* res, float* pol_w, float* pol_b, float* val_w, float* val_b,
             float* feat_w, float* feat_b) {
  const auto& c = tr->cfg;
  const int n = c.width, k = c.blocks, A = c.actions, FS = c.fsize, NH = tr->NH;
  std::vector<float> h((size_t)tr->P);
  if (dir == 1) {
    TR_CK(cudaMemcpyAsync(h.data(), flat_dev, sizeof(float) * tr->P, cudaMemcpyDeviceToHost, tr->stream));
    TR_CK(cudaStreamSynchronize(tr->stream));
  }
  auto mv = [&](float* flat, float* arr, size_t cnt) {
    if (!arr) return;
    if (dir == 0) memcpy(flat, arr, cnt * sizeof(float)); else memcpy(arr, flat, cnt * sizeof(float));
  };
  mv(h.data(), base, (size_t)n * c.in);
  for (int l = 0; l < k; l++) mv(h.data() + tr->off_res + (size_t)l * n * n, res ? res[l] : nullptr, (size_t)n * n);
  // heads: packed (NH x n) column-major <-> policy (A x n), value (1 x n), feature (FS x n)
  float* hp = h.data() + tr->off_heads;
  for (int i = 0; i < n; i++) {
    float* col = hp + (size_t)i * NH;
    if (pol_w) for (int a = 0; a < A; a++) { if (dir == 0) col[a] = pol_w[a + (size_t)A * i]; else pol_w[a + (size_t)A * i] = col[a]; }
    if (val_w) { if (dir == 0) col[A] = val_w[i]; else val_w[i] = col[A]; }
    if (feat_w) for (int j = 0; j < FS; j++) { if (dir == 0) col[A + 1 + j] = feat_w[j + (size_t)FS * i]; else feat_w[j + (size_t)FS * i] = col[A + 1 + j]; }
  }
  float* hb = h.data() + tr->off_bias;
  mv(hb, pol_b, A);
  mv(hb + A, val_b, 1);
  mv(hb + A + 1, feat_b, FS);
  if (dir == 0) {
    TR_CK(cudaMemcpyAsync(flat_dev, h.data(), sizeof(float) * tr->P, cudaMemcpyHostToDevice, tr->stream));
    TR_CK(cudaStreamSynchronize(tr->stream));
  }
  return AGPU_OK;
}

}  // namespace

extern "C" {

int agpu_trainer_create(agpu_trainer** out, const agpu_train_config* cfg) {
  if (!out || !cfg) { g_trainer_create_error = "null argument"; return AGPU_ERR_INVALID; }
  *out = nullptr;
  if (cfg->in < 1 || cfg->width < 1 || cfg->blocks < 0 || cfg->actions < 1 || cfg->fsize < 1 || cfg->max_batch < 1) {
    g_trainer_create_error = "bad network or batch dimensions";
    return AGPU_ERR_INVALID;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    g_trainer_create_error = "no CUDA device: the training step has no CPU fallback";
    return AGPU_ERR_NO_DEVICE;
  }
  if (cfg->device < 0 || cfg->device >= ndev) { g_trainer_create_error = "device ordinal out of range"; return AGPU_ERR_INVALID; }
  agpu_trainer* tr = new agpu_trainer();
  tr->cfg = *cfg;
  const int n = cfg->width, k = cfg->blocks;
  tr->NH = cfg->actions + 1 + cfg->fsize;
  tr->maxB = cfg->max_batch;
  tr->off_res = (long long)n * cfg->in;
  tr->off_heads = tr->off_res + (long long)k * n * n;
  tr->off_bias = tr->off_heads + (long long)tr->NH * n;
  tr->P = tr->off_bias + tr->NH;
  tr->betap[0] = cfg->beta1; tr->betap[1] = cfg->beta2;
  auto fail = [&](cudaError_t e, const char* what) {
    g_trainer_create_error = std::string(what) + ": " + cudaGetErrorString(e);
    agpu_trainer_destroy(tr);
    return AGPU_ERR_CUDA;
  };
  cudaError_t e;
  if ((e = cudaSetDevice(cfg->device)) != cudaSuccess) return fail(e, "cudaSetDevice");
  if ((e = cudaStreamCreateWithFlags(&tr->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "cudaStreamCreate");
  for (auto& ev : tr->ev) if ((e = cudaEventCreate(&ev)) != cudaSuccess) return fail(e, "cudaEventCreate");
  const size_t P = (size_t)tr->P, B = (size_t)tr->maxB, bn = B * n;
  const size_t ns = (B + TRAIN_KSLICE - 1) / TRAIN_KSLICE;
  struct { float** p; size_t cnt; } allocs[] = {
      {&tr->params, P}, {&tr->grads, P}, {&tr->mt, P}, {&tr->vt, P}, {&tr->part, ns * P},
      {&tr->X, B * cfg->in}, {&tr->H, (size_t)(k + 1) * bn}, {&tr->R, (size_t)(k > 0 ? k : 1) * bn}, {&tr->Zh, B * tr->NH}, {&tr->dZh, B * tr->NH},
      {&tr->dS[0], bn}, {&tr->dS[1], bn}, {&tr->dZ[0], bn}, {&tr->dZ[1], bn},
      {&tr->ypol, B * cfg->actions}, {&tr->yval, B}, {&tr->yfeat, B * cfg->fsize}, {&tr->terms, B * 3}, {&tr->loss_dev, 4}};
  for (auto& a : allocs) {
    if ((e = cudaMalloc((void**)a.p, a.cnt * sizeof(float))) != cudaSuccess) return fail(e, "cudaMalloc");
    if ((e = cudaMemset(*a.p, 0, a.cnt * sizeof(float))) != cudaSuccess) return fail(e, "cudaMemset");
  }
  if ((e = cudaMalloc((void**)&tr->st_i8, B * cfg->in)) != cudaSuccess) return fail(e, "cudaMalloc");
  if ((e = cudaMalloc((void**)&tr->fs_i8, B * cfg->fsize)) != cudaSuccess) return fail(e, "cudaMalloc");
  *out = tr;
  return AGPU_OK;
}

void agpu_trainer_destroy(agpu_trainer* tr) {
  if (!tr) return;
  cudaSetDevice(tr->cfg.device);
  float* bufs[] = {tr->params, tr->grads, tr->mt, tr->vt, tr->part, tr->X, tr->H, tr->R, tr->Zh, tr->dZh, tr->dS[0], tr->dS[1], tr->dZ[0], tr->dZ[1],
                   tr->ypol, tr->yval, tr->yfeat, tr->terms, tr->loss_dev};
  for (float* b : bufs) if (b) cudaFree(b);
  if (tr->st_i8) cudaFree(tr->st_i8);
  if (tr->fs_i8) cudaFree(tr->fs_i8);
  for (auto ev : tr->ev) if (ev) cudaEventDestroy(ev);
  if (tr->stream) cudaStreamDestroy(tr->stream);
  delete tr;
}

const char* agpu_trainer_last_error(const agpu_trainer* tr) { return tr ? tr->err.c_str() : g_trainer_create_error.c_str(); }

int agpu_trainer_set_params(agpu_trainer* tr, const float* base, const float* const* res, const float* pol_w, const float* pol_b,
                            const float* val_w, const float* val_b, const float* feat_w, const float* feat_b, int32_t reset_optimizer) {
  if (!tr) return AGPU_ERR_INVALID;
  TR_REQUIRE(base && pol_w && pol_b && val_w && val_b && feat_w && feat_b && (tr->cfg.blocks == 0 || res), AGPU_ERR_INVALID, "null weight pointer");
  for (int l = 0; l < tr->cfg.blocks; l++) TR_REQUIRE(res[l] != nullptr, AGPU_ERR_INVALID, "null residual weight pointer");
  TR_CK(cudaSetDevice(tr->cfg.device));
  int rc = exchange(tr, tr->params, 0, const_cast<float*>(base), const_cast<float* const*>(res), const_cast<float*>(pol_w), const_cast<float*>(pol_b),
                    const_cast<float*>(val_w), const_cast<float*>(val_b), const_cast<float*>(feat_w), const_cast<float*>(feat_b));
  if (rc != AGPU_OK) return rc;
  if (reset_optimizer) {
    TR_CK(cudaMemsetAsync(tr->mt, 0, sizeof(float) * tr->P, tr->stream));
    TR_CK(cudaMemsetAsync(tr->vt, 0, sizeof(float) * tr->P, tr->stream));
    TR_CK(cudaStreamSynchronize(tr->stream));
    tr->betap[0] = tr->cfg.beta1; tr->betap[1] = tr->cfg.beta2;
  }
  return AGPU_OK;
}

int agpu_trainer_get_params(agpu_trainer* tr, float* base, float* const* res, float* pol_w, float* pol_b, float* val_w, float* val_b,
                            float* feat_w, float* feat_b) {
  if (!tr) return AGPU_ERR_INVALID;
  TR_CK(cudaSetDevice(tr->cfg.device));
  return exchange(tr, tr->params, 1, base, res, pol_w, pol_b, val_w, val_b, feat_w, feat_b);
}

int agpu_trainer_get_grads(agpu_trainer* tr, float* base, float* const* res, float* pol_w, float* pol_b, float* val_w, float* val_b,
                           float* feat_w, float* feat_b) {
  if (!tr) return AGPU_ERR_INVALID;
  TR_REQUIRE(tr->has_grad, AGPU_ERR_STATE, "no gradient yet: call agpu_trainer_loss_grad first");
  TR_CK(cudaSetDevice(tr->cfg.device));
  return exchange(tr, tr->grads, 1, base, res, pol_w, pol_b, val_w, val_b, feat_w, feat_b);
}

int agpu_trainer_loss_grad(agpu_trainer* tr, const int8_t* state, const float* policy, const float* value, const int8_t* fstate, int64_t B,
                           float loss_out[4]) {
  if (!tr) return AGPU_ERR_INVALID;
  TR_CK(cudaSetDevice(tr->cfg.device));
  int rc = stage_batch(tr, state, policy, value, fstate, B);
  if (rc != AGPU_OK) return rc;
  TR_CK(cudaEventRecord(tr->ev[0], tr->stream));
  if ((rc = forward_loss(tr, (int)B, true)) != AGPU_OK) return rc;
  if ((rc = backward(tr, (int)B)) != AGPU_OK) return rc;
  TR_CK(cudaEventRecord(tr->ev[1], tr->stream));
  if ((rc = read_loss(tr, loss_out)) != AGPU_OK) return rc;
  TR_CK(cudaEventElapsedTime(&tr->last_ms[0], tr->ev[0], tr->ev[1]));
  return AGPU_OK;
}

int agpu_trainer_loss(agpu_trainer* tr, const int8_t* state, const float* policy, const float* value, const int8_t* fstate, int64_t B,
                      float loss_out[4]) {
  if (!tr) return AGPU_ERR_INVALID;
  TR_CK(cudaSetDevice(tr->cfg.device));
  int rc = stage_batch(tr, state, policy, value, fstate, B);
  if (rc != AGPU_OK) return rc;
  if ((rc = forward_loss(tr, (int)B, false)) != AGPU_OK) return rc;
  return read_loss(tr, loss_out);
}

int agpu_trainer_grad_buffer(agpu_trainer* tr, void** device_ptr, int64_t* count) {
  if (!tr) return AGPU_ERR_INVALID;
  TR_REQUIRE(device_ptr && count, AGPU_ERR_INVALID, "null argument");
  *device_ptr = tr->grads;
  *count = tr->P;
  return AGPU_OK;
}

int agpu_trainer_apply(agpu_trainer* tr, float grad_scale) {
  if (!tr) return AGPU_ERR_INVALID;
  TR_REQUIRE(tr->has_grad, AGPU_ERR_STATE, "no gradient yet: call agpu_trainer_loss_grad first");
  const auto& c = tr->cfg;
  TR_CK(cudaSetDevice(c.device));
  TR_CK(cudaEventRecord(tr->ev[2], tr->stream));
  adam_kernel<<<(unsigned)((tr->P + 255) / 256), 256, 0, tr->stream>>>(tr->params, tr->grads, tr->mt, tr->vt, tr->P, grad_scale, c.lr, c.beta1, c.beta2,
                                                                       tr->betap[0], tr->betap[1], c.eps, c.weight_decay);
  tr->launches++;
  TR_CK(cudaGetLastError());
  TR_CK(cudaEventRecord(tr->ev[3], tr->stream));
  TR_CK(cudaStreamSynchronize(tr->stream));
  TR_CK(cudaEventElapsedTime(&tr->last_ms[1], tr->ev[2], tr->ev[3]));
  tr->betap[0] *= c.beta1;                                   // βp .= βp .* β
  tr->betap[1] *= c.beta2;
  return AGPU_OK;
}

int agpu_trainer_step(agpu_trainer* tr, const int8_t* state, const float* policy, const float* value, const int8_t* fstate, int64_t B,
                      float loss_out[4]) {
  int rc = agpu_trainer_loss_grad(tr, state, policy, value, fstate, B, loss_out);
  if (rc != AGPU_OK) return rc;
  return agpu_trainer_apply(tr, 1.0f);
}

int agpu_trainer_opt_state(agpu_trainer* tr, float* m, float* v, double beta_pow[2], int32_t set) {
  if (!tr) return AGPU_ERR_INVALID;
  TR_CK(cudaSetDevice(tr->cfg.device));
  const cudaMemcpyKind kind = set ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
  if (m) TR_CK(set ? cudaMemcpyAsync(tr->mt, m, sizeof(float) * tr->P, kind, tr->stream) : cudaMemcpyAsync(m, tr->mt, sizeof(float) * tr->P, kind, tr->stream));
  if (v) TR_CK(set ? cudaMemcpyAsync(tr->vt, v, sizeof(float) * tr->P, kind, tr->stream) : cudaMemcpyAsync(v, tr->vt, sizeof(float) * tr->P, kind, tr->stream));
  TR_CK(cudaStreamSynchronize(tr->stream));
  if (beta_pow) {
    if (set) { tr->betap[0] = beta_pow[0]; tr->betap[1] = beta_pow[1]; } else { beta_pow[0] = tr->betap[0]; beta_pow[1] = tr->betap[1]; }
  }
  return AGPU_OK;
}

int agpu_trainer_last_ms(agpu_trainer* tr, float ms[2]) {
  if (!tr || !ms) return AGPU_ERR_INVALID;
  ms[0] = tr->last_ms[0]; ms[1] = tr->last_ms[1];
  return AGPU_OK;
}

long long agpu_trainer_launches(agpu_trainer* tr) { return tr ? tr->launches : 0; }

}  // extern "C"
