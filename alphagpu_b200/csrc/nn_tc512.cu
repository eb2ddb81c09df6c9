// nn_tc512.cu — the tcgen05/TMEM network chain for WIDTH 512 (the 512x8 nets of BASELINE configs 3-5: Hex, Gobang 9x9, Reversi).
//
// One CTA = one M=128 tile of leaf positions through all layers; nothing but the 16-byte leaf states and the logits touches HBM.
//   * Activations: 128 x 512 16-bit operands = 128 KB of shared memory, K-major, 128B-swizzled, eight K tiles of 64.
//   * Accumulators: 128 x 512 fp32 = all 512 TMEM columns (four N blocks of 128 columns).
//   * Weights: 512 KB per layer do not fit on chip; they stream as 16 KB chunks [128 outputs x 64 K] through a 5-stage
//     bulk-copy ring (cp.async.bulk + mbarrier complete_tx), pre-swizzled on the host in exactly the order the MMAs consume them
//     (N block major, K tile minor).  Each chunk feeds four M128xN128xK16 MMAs; tcgen05.commit frees its ring stage.
//   * Roles: warps 0-15 epilogue (TMEM lane quarter w%4, 128-column slice w/4), warp 16 lane 0 issues every MMA, warp 17 lane 0
//     is the weight producer.  The residual stream is the 16-bit activation tile itself (TMEM is full of accumulators), i.e.
//     b <- round16(b + relu(acc)) — the oracle's *_RESID modes mirror this.
// Roofline intent: per tile a layer is 128 MMAs x ~107 cycles of tensor pipe against ~1.5k cycles of epilogue and 512 KB of L2->SM
// weight traffic (37 B/clk/SM), so the chain is tensor-bound; one CTA per SM (208 KB shared memory).
#include <cuda_fp16.h>

#include <cstring>
#include <vector>

#include "tc_ptx.cuh"

namespace ag {

using namespace tc;

namespace {

constexpr int W5_N = 512;
constexpr int W5_KT = W5_N / 64;                      // 8 K tiles per trunk layer
constexpr int W5_STAGES = 5;
constexpr int W5_CHUNK = 128 * 128;                   // 16 KB: 128 outputs x 64 K operands
constexpr int W5_A_BYTES = 128 * W5_N * 2;            // 128 KB
constexpr int W5_THREADS = 32 * 18;
// register cap of the 576 threads.  (Capped at 64 the kernel leaves room for three 128-thread blocks of the search kernels next to a
// network CTA; measured on B200 that changes nothing — Hex 7: 8.07e7 sims/s at 64 registers, 8.18e7 at 112, with or without the
// search kernels' shared-memory carve-out set to the network kernel's — so the cap stays where the compiler is unconstrained.)
constexpr int W5_MAXREG = 112;
constexpr int W5_SMEM = W5_A_BYTES + W5_STAGES * W5_CHUNK + 1024 + 1024;

struct Tc512Args {
  const unsigned char* img;
  const float* bias;
  int nlayers, kt0, A, NH, in;
};

AG_D void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

// 64 bits [start, start+64) of an nc-word bit field (bits outside the field read as zero; start may be negative)
AG_D u64 field64(const u64* b, int nc, int start) {
  u64 r = 0;
  for (int w = 0; w < nc; w++) {
    const int sh = 64 * w - start;                    // position of word w inside the window
    if (sh > -64 && sh < 64) r |= sh >= 0 ? (b[w] << sh) : (b[w] >> (-sh));
  }
  return r;
}

template <int FMT> AG_D float2 unpack2(uint32_t u) {
  if (FMT == 0) return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xFFFF0000u));
  const __half2 h = *reinterpret_cast<const __half2*>(&u);
  return __half22float2(h);
}

template <int FMT>
__global__ void __maxnreg__(W5_MAXREG) tc_mlp512_kernel(Tc512Args T, NNInput I, int L, float* __restrict__ out, int outs) {
  int seg_off = 0;
  if (I.seg) {
    seg_off = I.seg[0];
    const int len = I.seg[1];
    if ((int)blockIdx.x * TC_TILE_M >= len) return;
    L = seg_off + len;
  }
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = smem;                                            // [8 K tiles][128 rows x 128 B]
  unsigned char* sW = smem + W5_A_BYTES;                               // [5][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + W5_STAGES * W5_CHUNK);
  // bars[0..4] full, [5..9] empty, [10] layer done, [11] activations ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  float* sbias = reinterpret_cast<float*>(bars + 13);                  // [128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);               // the same value, visibly warp-uniform to the compiler
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + 5), bar_done = smem_u32(bars + 10), bar_act = smem_u32(bars + 11);

  if (threadIdx.x == 0) {
    for (int s = 0; s < W5_STAGES; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_done, 1);
    mbar_init(bar_act, 16);                                            // one arrival per epilogue warp
    fence_barrier_init();
  }
  if (threadIdx.x < 128) sbias[threadIdx.x] = T.bias[threadIdx.x];
  if (warp == 16) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int last_layer = T.nlayers - 1;

  if (warp == 17) {
    // ===================== weight producer =====================
    if (lane == 0) {
      int c = 0;
      size_t off = 0;
      for (int l = 0; l <= last_layer; l++) {
        const int nchunks = (l == 0) ? 4 * T.kt0 : (l < last_layer ? 4 * W5_KT : W5_KT);
        const uint32_t bytes = (l == last_layer) ? (uint32_t)(T.NH * 128) : (uint32_t)W5_CHUNK;
        for (int j = 0; j < nchunks; j++, c++) {
          const int s = c % W5_STAGES;
          if (c >= W5_STAGES) mbar_wait(bar_empty + 8 * s, ((c / W5_STAGES) - 1) & 1);
          mbar_expect_tx(bar_full + 8 * s, bytes);
          bulk_g2s(smem_u32(sW + s * W5_CHUNK), T.img + off, bytes, bar_full + 8 * s);
          off += bytes;
        }
      }
    }
  } else if (warp_u == 16) {
    // ===================== MMA issuer =====================
    // the whole warp walks the schedule (uniform control flow, uniform operands); one elected lane issues
    {
      int c = 0;
      const uint32_t a0 = smem_u32(sA);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      for (int l = 0; l <= last_layer; l++) {
        mbar_wait(bar_act, l & 1);                                     // A operand of this layer is in shared memory
        tc_fence_after();
        const int KT = (l == 0) ? T.kt0 : W5_KT;
        const int NB = (l == last_layer) ? 1 : 4;
        const uint32_t idesc = umma_idesc<FMT>(l == last_layer ? T.NH : 128);
        for (int nb = 0; nb < NB; nb++) {
          for (int kt = 0; kt < KT; kt++, c++) {
            const int s = c % W5_STAGES;
            mbar_wait(bar_full + 8 * s, (c / W5_STAGES) & 1);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t ad0 = umma_desc(a0 + kt * TC_KTILE_BYTES_A);
              const uint64_t bd0 = umma_desc(smem_u32(sW) + (uint32_t)(s * W5_CHUNK));
#pragma unroll
              for (int ks = 0; ks < 4; ks++)
                umma_bf16(tmem_u + (uint32_t)(nb * 128), ad0 + (uint64_t)((ks * 32) >> 4), bd0 + (uint64_t)((ks * 32) >> 4), idesc, (kt | ks) ? 1u : 0u);
              umma_commit(bar_empty + 8 * s);                          // chunk consumed -> producer may refill the stage
            }
            __syncwarp();
          }
        }
        if (elect_one()) umma_commit(bar_done);                        // the whole layer has drained
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int wq = warp & 3, cs = warp >> 2;                           // TMEM lane quarter, 128-column slice
    const int r = wq * 32 + lane;
    const int g = seg_off + (int)blockIdx.x * TC_TILE_M + r;
    const uint32_t tmem_row = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(cs * 128);
    // operand address of the 16-byte chunk holding columns [8*c8, 8*c8+8) of row r
    auto chunk_ptr = [&](int c8) { return sA + (c8 >> 3) * TC_KTILE_BYTES_A + r * 128 + (((c8 & 7) ^ (r & 7)) << 4); };

    // ---- A operand of the base layer (decoder, mcts_gpu.jl:202-223): this warp's K tile = cs, if the input reaches it ----
    if (cs < T.kt0) {
      if (I.x_direct) {
        const float* xd = I.x_direct + (size_t)g * T.in;
#pragma unroll 1
        for (int i = 0; i < 8; i++) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const int k = 64 * cs + 8 * i + 2 * e;
            const float f0 = (k < T.in && g < L) ? xd[k] : 0.f;
            const float f1 = (k + 1 < T.in && g < L) ? xd[k + 1] : 0.f;
            w[e] = pack2<FMT>(f0, f1);
          }
          *reinterpret_cast<uint4*>(chunk_ptr(8 * cs + i)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      } else {
        u64 bits = 0;
        if (g < L) {
          const u64* st = reinterpret_cast<const u64*>(I.tree + (size_t)g * I.game_stride + (size_t)I.leaf[g] * I.rec + I.off_state);
          u64 bp[3] = {0, 0, 0}, bo[3] = {0, 0, 0};
          for (int w = 0; w < I.nc; w++) { bp[w] = st[w]; bo[w] = st[I.nc + w]; }
          bits = field64(bp, I.nc, 64 * cs) | field64(bo, I.nc, 64 * cs - I.VS);      // x = [bplayer bits | bopponent bits]
        }
        const uint32_t one = (FMT == 0) ? 0x3F80u : 0x3C00u;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const uint32_t byte = (uint32_t)(bits >> (8 * i)) & 0xFFu;
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; e++) w[e] = bits2_to_operands(byte >> (2 * e), one);
          *reinterpret_cast<uint4*>(chunk_ptr(8 * cs + i)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_act);

    for (int l = 0; l <= last_layer; l++) {
      mbar_wait(bar_done, l & 1);
      tc_fence_after();
      if (l < last_layer) {
        // b = relu(acc) (base) or round16(b + relu(acc)); the activation tile is overwritten in place (all MMAs of the layer are done).
        // The tensor-memory load of chunk cb + 1 is in flight while chunk cb is processed (tcgen05.wait::ld waits for ALL outstanding
        // loads: without the double buffer the load latency is exposed eight times per layer).
        uint32_t v[2][16];
        const f32x2 half2 = pack2f(0.5f, 0.5f);
        tmem_ld16(tmem_row, v[0]);
#pragma unroll
        for (int cb = 0; cb < 8; cb++) {
          const int b = cb & 1;
          unsigned char* p0 = chunk_ptr(16 * cs + 2 * cb);
          unsigned char* p1 = chunk_ptr(16 * cs + 2 * cb + 1);
          uint4 old0 = make_uint4(0, 0, 0, 0), old1 = make_uint4(0, 0, 0, 0);
          if (l > 0) { old0 = *reinterpret_cast<const uint4*>(p0); old1 = *reinterpret_cast<const uint4*>(p1); }
          tmem_ld_wait();                                              // chunk cb has arrived
          if (cb + 1 < 8) tmem_ld16(tmem_row + 16 * (cb + 1), v[b ^ 1]);
          const uint32_t ow[8] = {old0.x, old0.y, old0.z, old0.w, old1.x, old1.y, old1.z, old1.w};
          uint32_t nw[8];
#pragma unroll
          for (int e = 0; e < 8; e++) {
            // b + relu(a) as fma(a + |a|, 0.5, b): a + |a| is 2 relu(a) exactly and the fma rounds once, like the add — FMA-pipe
            // instructions only (an FADD per column, a packed FFMA per pair) instead of an FMNMX and an FADD per column
            const float2 o = unpack2<FMT>(ow[e]);
            const float a0 = __uint_as_float(v[b][2 * e]), a1 = __uint_as_float(v[b][2 * e + 1]);
            float h0, h1;
            unpack2f(fma2(pack2f(__fadd_rn(a0, fabsf(a0)), __fadd_rn(a1, fabsf(a1))), half2, pack2f(o.x, o.y)), h0, h1);
            nw[e] = pack2<FMT>(h0, h1);
          }
          *reinterpret_cast<uint4*>(p0) = make_uint4(nw[0], nw[1], nw[2], nw[3]);
          *reinterpret_cast<uint4*>(p1) = make_uint4(nw[4], nw[5], nw[6], nw[7]);
        }
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_act);
      } else if (cs == 0) {
        // heads: logits = acc + bias, value = σ(acc[A] + bias[A])   (DenseNet.jl:301)
        float* o = out + (size_t)g * outs;
        for (int cb = 0; cb < 8 && 16 * cb < T.NH; cb++) {
          uint32_t v[16];
          tmem_ld16(tmem_row + 16 * cb, v);
          tmem_ld_wait();
          const int a0 = 16 * cb;
          float z[16];
#pragma unroll
          for (int e = 0; e < 16; e++) z[e] = __uint_as_float(v[e]) + sbias[a0 + e];
          if (T.A >= a0 && T.A < a0 + 16) {
#pragma unroll
            for (int e = 0; e < 16; e++) if (a0 + e == T.A) z[e] = c_sigmoidf(z[e]);
          }
          if (g < L) {
#pragma unroll
            for (int q4 = 0; q4 < 4; q4++)
              if (a0 + 4 * q4 < outs) *reinterpret_cast<float4*>(o + a0 + 4 * q4) = make_float4(z[4 * q4], z[4 * q4 + 1], z[4 * q4 + 2], z[4 * q4 + 3]);
          }
        }
        tc_fence_before();
      }
    }
  }
  __syncthreads();
  if (warp == 16) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

inline uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) != 0x7F800000u) u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
inline uint16_t f2h(float f) {
  const __half h = __float2half_rn(f);      // host path of cuda_fp16: round to nearest even
  uint16_t u;
  memcpy(&u, &h, 2);
  if ((u & 0x7FFF) >= 0x7C00) u = (u & 0x8000) | 0x7BFF;   // saturate to +-65504 like cvt.rn.satfinite
  return u;
}

// one chunk: W[n][k] for n in [n0, n0+rows_pad), k in [k0, k0+64); rows beyond `rows` / columns beyond `kreal` are zero
void put_chunk(unsigned char* img, int rows_pad, const std::vector<float>& w, int rows, int kreal, int n0, int k0, int fmt) {
  memset(img, 0, (size_t)rows_pad * 128);
  for (int nl = 0; nl < rows_pad; nl++) {
    const int n = n0 + nl;
    if (n >= rows) continue;
    for (int kk = 0; kk < 64; kk++) {
      const int k = k0 + kk;
      if (k >= kreal) continue;
      const size_t off = (size_t)nl * 128 + (size_t)(((kk >> 3) ^ (nl & 7)) << 4) + (size_t)(kk & 7) * 2;
      const float x = w[(size_t)n * kreal + k];
      const uint16_t v = fmt == 0 ? f2bf(x) : f2h(x);
      memcpy(img + off, &v, 2);
    }
  }
}

}  // namespace

int tc512_supported(int in, int n, int k, int A) { return n == W5_N && in <= 256 && k >= 0 && head_n(A) <= 128; }

size_t tc512_image_bytes(int in, int n, int k, int A) {
  const int kt0 = (in + 63) / 64;
  (void)n;
  return (size_t)4 * kt0 * W5_CHUNK + (size_t)k * 4 * W5_KT * W5_CHUNK + (size_t)W5_KT * head_n(A) * 128;
}

void tc512_build_image(const float* base, const float* const* res, const float* pol_w, const float* pol_b, const float* val_w,
                       const float* val_b, int in, int n, int k, int A, void* img_host, float* bias_host, int fmt) {
  unsigned char* img = (unsigned char*)img_host;
  const int kt0 = (in + 63) / 64;
  std::vector<float> w((size_t)n * in);
  for (int o = 0; o < n; o++) for (int i = 0; i < in; i++) w[(size_t)o * in + i] = base[o + (size_t)n * i];
  for (int nb = 0; nb < 4; nb++) for (int kt = 0; kt < kt0; kt++) { put_chunk(img, 128, w, n, in, nb * 128, kt * 64, fmt); img += W5_CHUNK; }
  for (int l = 0; l < k; l++) {
    w.assign((size_t)n * n, 0.f);
    for (int o = 0; o < n; o++) for (int i = 0; i < n; i++) w[(size_t)o * n + i] = res[l][o + (size_t)n * i];
    for (int nb = 0; nb < 4; nb++) for (int kt = 0; kt < W5_KT; kt++) { put_chunk(img, 128, w, n, n, nb * 128, kt * 64, fmt); img += W5_CHUNK; }
  }
  const int NH = head_n(A);
  w.assign((size_t)(A + 1) * n, 0.f);
  for (int a = 0; a < A; a++) for (int i = 0; i < n; i++) w[(size_t)a * n + i] = pol_w[a + (size_t)A * i];
  for (int i = 0; i < n; i++) w[(size_t)A * n + i] = val_w[i];
  for (int kt = 0; kt < W5_KT; kt++) { put_chunk(img, NH, w, A + 1, n, 0, kt * 64, fmt); img += (size_t)NH * 128; }
  for (int a = 0; a < 256; a++) bias_host[a] = 0.f;
  for (int a = 0; a < A; a++) bias_host[a] = pol_b[a];
  bias_host[A] = val_b[0];
}

cudaError_t tc512_init() {
  cudaError_t e = cudaFuncSetAttribute(tc_mlp512_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, W5_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_mlp512_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, W5_SMEM);
  return e;
}

cudaError_t tc512_forward(const NetDev& net, const NNInput& I, int L, float* out, int outs, cudaStream_t stream, int fmt) {
  Tc512Args T;
  T.img = (const unsigned char*)net.tc_img; T.bias = net.tc_bias; T.nlayers = net.k + 2; T.kt0 = (net.in + 63) / 64; T.A = net.A;
  T.NH = head_n(net.A); T.in = net.in;
  const int grid = (L + TC_TILE_M - 1) / TC_TILE_M;
  if (fmt == 0) tc_mlp512_kernel<0><<<grid, W5_THREADS, W5_SMEM, stream>>>(T, I, L, out, outs);
  else tc_mlp512_kernel<1><<<grid, W5_THREADS, W5_SMEM, stream>>>(T, I, L, out, outs);
  return cudaGetLastError();
}

}  // namespace ag
