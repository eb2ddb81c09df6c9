// nn_tc.cu — snetwork2 forward (DenseNet.jl:294-304) as a bf16 tcgen05/TMEM GEMM chain for sm_100a.
//
// One CTA evaluates 256 leaf positions (two M=128 tiles) through ALL layers without touching HBM in
// between:  x -> relu(W0 x) -> k x [ b = relu(b + relu(W b)) ] -> (policy | value) heads.
//
//   warps 0-7  : tile 0   } each group of 8 warps owns one 128-row tile: builds the A operand (0/1 encoding of the
//   warps 8-15 : tile 1   } leaf bitboards, straight from the tree record), one elected lane issues the
//                           tcgen05.mma chain of the layer, all 256 threads run the epilogue: warp w reads TMEM lane
//                           quarter w%4 (its rows) and column half w/4 with tcgen05.ld, keeps its 64 columns of the
//                           residual stream in fp32 registers, and repacks them into the swizzled A operand of the
//                           next layer.  While one tile is in its epilogue the tensor pipe runs the other tile's MMAs.
//   weight producer: the issuing lane of tile 0 also streams the per-layer weight images global->shared with 1-D
//                bulk copies (cp.async.bulk, mbarrier complete_tx) through a 3-stage ring shared by both tiles, two
//                layers ahead of the MMAs (no extra producer warp: 17 warps would put 5 on one SM sub-partition and cut
//                the register budget below what the fp32 residual columns need).
//
// Operands: A (activations) and B (weights) are K-major bf16 with the 128-byte swizzle the UMMA shared-memory
// descriptor expects (8-row x 128 B atoms, SBO = 1024 B); weights are pre-swizzled on the host into exactly
// that image (tc_build_image), so a plain bulk copy lands them ready for the MMA.  Accumulators: fp32 in TMEM,
// 128 columns per tile.  The residual stream stays in fp32 registers (one row per thread); only the MMA operand
// is rounded to bf16 — the CPU oracle's "bf16-faithful" mode mirrors exactly these roundings.
#include <cstring>
#include <vector>

#include "tc_ptx.cuh"

namespace ag {

using namespace tc;

namespace {

constexpr int TC_WARPS_PER_TILE = 8;                                 // 4 TMEM lane quarters x 2 column halves
constexpr int TC_THREADS = 32 * TC_WARPS_PER_TILE * TC_TILES;        // 512: 4 warps per SM sub-partition, 128 registers per thread
constexpr int TC_SMEM = TC_TILES * TC_A_BYTES + TC_STAGES * TC_W_STAGE_BYTES + 1024 + 1024;   // + barriers + alignment slack


template <int FMT>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_mlp128_kernel(TcArgs T, NNInput I, int L, float* __restrict__ out, int outs) {
  int seg_off = 0;
  if (I.seg) {                                     // graph replay: this launch covers slots [off, off+len); surplus CTAs leave at once
    seg_off = I.seg[0];
    const int len = I.seg[1];
    if ((int)blockIdx.x * (TC_TILES * TC_TILE_M) >= len) return;
    L = seg_off + len;
  }
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = smem;                                          // [2][32 KB]
  unsigned char* sW = smem + TC_TILES * TC_A_BYTES;                  // [3][32 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + TC_STAGES * TC_W_STAGE_BYTES);
  // bars[0..2] full, [3..5] empty, [6..7] mma_done, [8] stagger (one-shot), then the TMEM base word
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  float* sbias = reinterpret_cast<float*>(bars + 10);               // [128] head biases

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* trp = (T.dbg && (warp % TC_WARPS_PER_TILE) == 0 && lane == 0) ? T.dbg + (((size_t)blockIdx.x * TC_TILES + warp / TC_WARPS_PER_TILE) * 16 + 12) * 4 : nullptr;
  if (trp) trp[0] = clock64();
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + 3), bar_done = smem_u32(bars + 6);

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, TC_TILES); }
    for (int t = 0; t < TC_TILES; t++) mbar_init(bar_done + 8 * t, 1);
    mbar_init(bar_full + 8 * 8, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < TC_N) sbias[threadIdx.x] = T.bias[threadIdx.x];
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (trp) trp[1] = clock64();

  // weight image of layer l -> ring stage l % 3 (called by one thread)
  auto load_layer = [&](int l) {
    const int s = l % TC_STAGES;
    const uint32_t bytes = (l == T.nlayers - 1) ? (uint32_t)(T.NH * TC_N * 2) : (uint32_t)TC_W_STAGE_BYTES;
    if (l >= TC_STAGES) mbar_wait(bar_empty + 8 * s, ((l / TC_STAGES) - 1) & 1);    // both tiles' MMAs of layer l-3 have drained
    mbar_expect_tx(bar_full + 8 * s, bytes);
    bulk_g2s(smem_u32(sW + s * TC_W_STAGE_BYTES), T.img + (size_t)l * TC_W_STAGE_BYTES, bytes, bar_full + 8 * s);
  };
  if (threadIdx.x == 0) {
    load_layer(0);
    if (T.nlayers > 1) load_layer(1);
  }
  {
    // ===================== tile warpgroups =====================
    const int t = warp / TC_WARPS_PER_TILE;        // tile of this warp group
    const int wq = warp & 3;                       // TMEM lane quarter (a warp may only touch lanes 32*(warp%4)..+31)
    const int ch = (warp >> 2) & 1;                // column half handled by this warp
    const int r = wq * 32 + lane;                  // row in tile == TMEM lane
    const int g = seg_off + blockIdx.x * (TC_TILES * TC_TILE_M) + t * TC_TILE_M + r;
    unsigned char* At = sA + t * TC_A_BYTES;
    const uint32_t tmem_acc = tmem_base + (uint32_t)(t * TC_N);                       // column offset of this tile
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0), t_u = warp_u / TC_WARPS_PER_TILE;   // the same values, visibly warp-uniform
    const uint32_t tmem_acc_u = __shfl_sync(0xffffffffu, tmem_base, 0) + (uint32_t)(t_u * TC_N);
    const uint32_t tmem_row = tmem_acc + ((uint32_t)(wq * 32) << 16) + (uint32_t)(ch * 64);   // lane base of this warp's quarter, its column half

    // ---- A operand of the base layer: 0/1 encoding of the leaf position (decoder, mcts_gpu.jl:202-223) ----
    {
      u64 b0 = 0, b1 = 0;
      const float* xd = nullptr;
      if (g < L) {
        if (I.x_direct) xd = I.x_direct + (size_t)g * (2 * I.VS);
        else {
          const u64* st = reinterpret_cast<const u64*>(I.tree + (size_t)g * I.game_stride + (size_t)I.leaf[g] * I.rec + I.off_state);
          const int nch = 2 * I.nc;
          (void)nch;
          b0 = st[0]; b1 = st[1];                           // nc == 1 (tc_supported): bplayer, bopponent
        }
      }
      const int VS = I.VS, nc = I.nc;
      if (!xd) {
        // tree path (2*VS <= 128 => one 64-bit chunk per board): x = [bplayer bits 0..VS-1 | bopponent bits 0..VS-1] as a
        // 128-bit vector; each pair of bits becomes one packed pair of 0.0/1.0 operands
        const u64 bp = b0, bo = b1;
        const u64 x0 = (VS < 64) ? (bp | (bo << VS)) : bp;
        const u64 x1 = (VS < 64) ? (bo >> (64 - VS)) : bo;
        const u64 xh = ch ? x1 : x0;                       // this warp's K tile = operand bits 64*ch .. 64*ch+63
        const uint32_t one = (FMT == 0) ? 0x3F80u : 0x3C00u;
#pragma unroll
        for (int cl = 0; cl < 8; cl++) {
          const uint32_t byte = (uint32_t)(xh >> (8 * cl)) & 0xFFu;
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; e++) w[e] = bits2_to_operands(byte >> (2 * e), one);
          *reinterpret_cast<uint4*>(At + ch * TC_KTILE_BYTES_A + r * 128 + ((cl ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      } else {
#pragma unroll 1
        for (int c = 8 * ch; c < 8 * ch + 8; c++) {  // this warp's 8 chunks of 8 operands (K tile `ch`)
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const int k = c * 8 + e * 2;
            const float f0 = (k < 2 * VS && g < L) ? xd[k] : 0.f;
            const float f1 = (k + 1 < 2 * VS && g < L) ? xd[k + 1] : 0.f;
            w[e] = pack2<FMT>(f0, f1);
          }
          const int ktile = c >> 3, cc = c & 7;
          *reinterpret_cast<uint4*>(At + ktile * TC_KTILE_BYTES_A + r * 128 + ((cc ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      (void)nc;
    }
    fence_proxy_async();
    named_bar_sync(1 + t, 32 * TC_WARPS_PER_TILE);
    if (trp) trp[2] = clock64();

    float h[TC_N / 2];                                   // fp32 residual stream of this row
#pragma unroll
    for (int i = 0; i < TC_N / 2; i++) h[i] = 0.f;

    for (int l = 0; l < T.nlayers; l++) {
      const int s = l % TC_STAGES;
      const bool is_head = (l == T.nlayers - 1);
      const int nl = is_head ? T.NH : TC_N;
      const bool tracer = T.dbg && (warp % TC_WARPS_PER_TILE) == 0 && lane == 0;
      long long* tr = T.dbg ? T.dbg + (((size_t)blockIdx.x * TC_TILES + t) * 16 + l) * 4 : nullptr;
      if (tracer) tr[0] = clock64();
      if ((warp_u % TC_WARPS_PER_TILE) == 0) {
        // ---- MMA issue: D[128 x nl] = A[128 x K] * W_l[nl x K]^T ----
        // (warp-uniform branch, one elected lane issues: see elect_one in tc_ptx.cuh)
        mbar_wait(bar_full + 8 * s, (l / TC_STAGES) & 1);
        // stagger: tile 1 issues its first layer only after tile 0's first layer has drained, so that from then on one tile's
        // MMAs run while the other tile is in its epilogue (in lockstep both would share the tensor pipe, then both leave it idle)
        // (a one-shot barrier: its phase 0 completes once and never flips back, so a late tile 1 can not miss it)
        if (l == 0 && t_u == 1) mbar_wait(bar_full + 8 * 8, 0);
        tc_fence_after();
        if (elect_one()) {
          // 8 K-steps of 16 (the base layer's operands are zero-padded to K = 128), fully unrolled with precomputed descriptor increments
          const uint64_t ad0 = umma_desc(smem_u32(sA) + (uint32_t)(t_u * TC_A_BYTES));
          const uint64_t bd0 = umma_desc(smem_u32(sW) + (uint32_t)(s * TC_W_STAGE_BYTES));
          const uint32_t idesc = umma_idesc<FMT>(nl);
          const uint64_t bstep = (uint64_t)((nl * 128) >> 4);            // second K tile of the weight image
#pragma unroll
          for (int ks = 0; ks < TC_N / 16; ks++) {
            const uint64_t ainc = (uint64_t)(((ks >> 2) * TC_KTILE_BYTES_A + (ks & 3) * 32) >> 4);
            const uint64_t binc = (uint64_t)(((ks & 3) * 32) >> 4) + ((ks >> 2) ? bstep : 0);
            umma_bf16(tmem_acc_u, ad0 + ainc, bd0 + binc, idesc, ks > 0 ? 1u : 0u);
          }
          umma_commit(bar_done + 8 * t_u);             // accumulator ready -> epilogue of this tile
          umma_commit(bar_empty + 8 * s);              // weight stage consumed by this tile
          if (l == 0 && t_u == 0) umma_commit(bar_full + 8 * 8);
          if (t_u == 0 && l + 2 < T.nlayers) load_layer(l + 2);   // producer role: keep the ring two layers ahead
        }
        __syncwarp();
        if (tracer) tr[1] = clock64();
      }
      mbar_wait(bar_done + 8 * t, l & 1);
      tc_fence_after();
      if (tracer) tr[2] = clock64();

      if (!is_head) {
        // ---- epilogue: b = relu(acc) (base) or relu(b + relu(acc)); next A operand = bf16/fp16(b) ----
        uint32_t va[16], vb[16];
        auto process = [&](const int cb, const uint32_t (&cur)[16]) {      // this warp's columns 64*ch + 16*cb .. +15
#pragma unroll
          for (int i = 0; i < 16; i++) {
            const float ra = fmaxf(__uint_as_float(cur[i]), 0.f);
            // relu(b + relu(acc)) == b + relu(acc): b >= 0 by induction from b0 = relu(.), so the outer relu is the identity
            h[cb * 16 + i] = (l == 0) ? ra : h[cb * 16 + i] + ra;
          }
#pragma unroll
          for (int c2 = 0; c2 < 2; c2++) {           // 2 chunks of 8 columns
            const int cl = cb * 2 + c2;              // chunk within this warp's half, 0..7
            const uint4 pk = make_uint4(pack2<FMT>(h[cl * 8 + 0], h[cl * 8 + 1]), pack2<FMT>(h[cl * 8 + 2], h[cl * 8 + 3]),
                                        pack2<FMT>(h[cl * 8 + 4], h[cl * 8 + 5]), pack2<FMT>(h[cl * 8 + 6], h[cl * 8 + 7]));
            *reinterpret_cast<uint4*>(At + ch * TC_KTILE_BYTES_A + r * 128 + ((cl ^ (r & 7)) << 4)) = pk;   // K tile == column half
          }
        };
        // the TMEM load of the next 16 columns is in flight while the current ones are processed
        tmem_ld16(tmem_row, va);
#pragma unroll
        for (int cb = 0; cb < 4; cb += 2) {
          tmem_ld_wait();
          tmem_ld16(tmem_row + (cb + 1) * 16, vb);
          process(cb, va);
          tmem_ld_wait();
          if (cb + 2 < 4) tmem_ld16(tmem_row + (cb + 2) * 16, va);
          process(cb + 1, vb);
        }
        tc_fence_before();
        fence_proxy_async();
        named_bar_sync(1 + t, 32 * TC_WARPS_PER_TILE);
        if (tracer) tr[3] = clock64();
      } else {
        // ---- heads: logits = acc + bias, value = σ(acc[A] + bias[A])   (DenseNet.jl:301) ----
        float* o = out + (size_t)g * outs;
        for (int cb = 0; cb < 4 && ch * 64 + cb * 16 < T.NH; cb++) {     // warp-uniform bounds
          uint32_t v[16];
          tmem_ld16(tmem_row + cb * 16, v);
          tmem_ld_wait();
          const int a0 = ch * 64 + cb * 16;
          float z[16];
#pragma unroll
          for (int i = 0; i < 16; i++) z[i] = __uint_as_float(v[i]) + sbias[a0 + i];
          if (T.A >= a0 && T.A < a0 + 16) {                              // the value column lives in this group of 16 (warp-uniform)
#pragma unroll
            for (int i = 0; i < 16; i++) if (a0 + i == T.A) z[i] = c_sigmoidf(z[i]);
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
  if (trp) trp[3] = clock64();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
  if (trp) trp[4] = clock64();
}

inline uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) != 0x7F800000u) u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
// fp32 -> fp16, round to nearest even, saturating to +-65504 (matches cvt.rn.satfinite.f16.f32)
inline uint16_t f2h(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  const uint16_t sign = (uint16_t)((u >> 16) & 0x8000u);
  const uint32_t a = u & 0x7FFFFFFFu;
  if (a >= 0x477FF000u) return sign | 0x7BFF;                 // >= 65520 (or inf/nan): saturate to 65504
  if (a < 0x33000001u) return sign;                           // < 2^-25 (rounds to zero)
  int e = (int)(a >> 23) - 127;
  uint32_t m = (a & 0x7FFFFFu) | 0x800000u;
  if (e < -14) {                                              // subnormal half
    const int sh = 13 + (-14 - e);
    uint32_t q = m >> sh, rem = m & ((1u << sh) - 1u), half = 1u << (sh - 1);
    if (rem > half || (rem == half && (q & 1u))) q++;
    return sign | (uint16_t)q;
  }
  uint32_t q = m >> 13, rem = m & 0x1FFFu;
  uint32_t h = ((uint32_t)(e + 15) << 10) + (q - 0x400u);
  if (rem > 0x1000u || (rem == 0x1000u && (q & 1u))) h++;
  return sign | (uint16_t)h;
}

// image of one layer: W[n][k] (n < rows, k < kreal; zero elsewhere), K padded to 128, as two K-tiles of
// [rows_pad x 64] bf16 with the 128B swizzle: chunk' = chunk ^ (n & 7)
void put_layer(unsigned char* img, int rows_pad, int rows, int kreal, const std::vector<float>& w /* row-major [rows][kreal] */, int fmt) {
  memset(img, 0, (size_t)rows_pad * TC_N * 2);
  for (int n = 0; n < rows; n++)
    for (int k = 0; k < kreal; k++) {
      const int ktile = k >> 6, kk = k & 63, chunk = kk >> 3;
      const size_t off = (size_t)ktile * rows_pad * 128 + (size_t)n * 128 + (size_t)((chunk ^ (n & 7)) << 4) + (size_t)(kk & 7) * 2;
      const uint16_t v = fmt == 0 ? f2bf(w[(size_t)n * kreal + k]) : f2h(w[(size_t)n * kreal + k]);
      memcpy(img + off, &v, 2);
    }
}

}  // namespace

// width-512 variant (nn_tc512.cu)
int tc512_supported(int in, int n, int k, int A);
size_t tc512_image_bytes(int in, int n, int k, int A);
void tc512_build_image(const float* base, const float* const* res, const float* pol_w, const float* pol_b, const float* val_w,
                       const float* val_b, int in, int n, int k, int A, void* img_host, float* bias_host, int fmt);
cudaError_t tc512_init();
cudaError_t tc512_forward(const NetDev& net, const NNInput& I, int L, float* out, int outs, cudaStream_t stream, int fmt);

int tc_supported(int in, int n, int k, int A) {
  return (n == TC_N && in <= TC_N && k >= 0 && head_n(A) <= TC_N) || tc512_supported(in, n, k, A);
}

size_t tc_image_bytes(int in, int n, int k, int A) {
  if (n != TC_N) return tc512_image_bytes(in, n, k, A);
  return (size_t)(1 + k) * n * n * 2 + (size_t)head_n(A) * n * 2;
}

void tc_build_image(const float* base, const float* const* res, const float* pol_w, const float* pol_b, const float* val_w,
                    const float* val_b, int in, int n, int k, int A, void* img_host, float* bias_host, int fmt) {
  if (n != TC_N) { tc512_build_image(base, res, pol_w, pol_b, val_w, val_b, in, n, k, A, img_host, bias_host, fmt); return; }
  unsigned char* img = (unsigned char*)img_host;
  std::vector<float> w;
  // base: Julia (n x in) column-major -> row-major [n][in]
  w.assign((size_t)n * in, 0.f);
  for (int o = 0; o < n; o++) for (int i = 0; i < in; i++) w[(size_t)o * in + i] = base[o + (size_t)n * i];
  put_layer(img, n, n, in, w, fmt);
  img += (size_t)n * n * 2;
  for (int l = 0; l < k; l++) {
    w.assign((size_t)n * n, 0.f);
    for (int o = 0; o < n; o++) for (int i = 0; i < n; i++) w[(size_t)o * n + i] = res[l][o + (size_t)n * i];
    put_layer(img, n, n, n, w, fmt);
    img += (size_t)n * n * 2;
  }
  const int NH = head_n(A);
  w.assign((size_t)(A + 1) * n, 0.f);
  for (int a = 0; a < A; a++) for (int i = 0; i < n; i++) w[(size_t)a * n + i] = pol_w[a + (size_t)A * i];
  for (int i = 0; i < n; i++) w[(size_t)A * n + i] = val_w[i];
  put_layer(img, NH, A + 1, n, w, fmt);
  for (int a = 0; a < 256; a++) bias_host[a] = 0.f;
  for (int a = 0; a < A; a++) bias_host[a] = pol_b[a];
  bias_host[A] = val_b[0];
}

long long* g_tc_dbg = nullptr;   // development: set through agpu_debug_tc_trace

cudaError_t tc_init() {
  cudaError_t e = cudaFuncSetAttribute(tc_mlp128_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_mlp128_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
  if (e == cudaSuccess) e = tc512_init();
  return e;
}

cudaError_t tc_forward(const NetDev& net, const NNInput& I, int L, float* out, int outs, cudaStream_t stream, int fmt) {
  if (net.n != TC_N) return tc512_forward(net, I, L, out, outs, stream, fmt);
  TcArgs T;
  T.img = (const unsigned char*)net.tc_img; T.bias = net.tc_bias; T.nlayers = net.k + 2; T.k0_steps = (net.in + 15) / 16; T.A = net.A;
  T.NH = head_n(net.A); T.in = net.in;
  T.dbg = g_tc_dbg;
  const int grid = (L + TC_TILES * TC_TILE_M - 1) / (TC_TILES * TC_TILE_M);
  if (fmt == 0) tc_mlp128_kernel<0><<<grid, TC_THREADS, TC_SMEM, stream>>>(T, I, L, out, outs);
  else tc_mlp128_kernel<1><<<grid, TC_THREADS, TC_SMEM, stream>>>(T, I, L, out, outs);
  return cudaGetLastError();
}

}  // namespace ag
