// tc_ptx.cuh — thin inline-PTX wrappers for the sm_100a tensor-core path (tcgen05 / TMEM / mbarrier / bulk copy) and the
// constants of the operand layout shared by the stand-alone network kernel (nn_tc.cu) and the fused per-ply kernel (fused.cuh).
#pragma once
#include <cuda_bf16.h>
#include "nn.cuh"

namespace ag {
namespace tc {

constexpr int TC_N = 128;              // MLP width handled by the tensor-core chain
constexpr int TC_TILE_M = 128;         // rows (games) per tile = UMMA M
constexpr int TC_TILES = 2;            // tiles per CTA
constexpr int TC_STAGES = 3;           // weight ring depth
constexpr int TC_W_STAGE_BYTES = TC_N * TC_N * 2;                    // 32 KB: one 128x128 16-bit layer image
constexpr int TC_A_BYTES = TC_TILE_M * TC_N * 2;                     // 32 KB per tile
constexpr int TC_KTILE_BYTES_A = TC_TILE_M * 128;                    // 16 KB: 128 rows x 64 operands

__host__ __device__ inline int head_n(int A) { return (A + 1 + 15) / 16 * 16; }

// ---------------- PTX wrappers ----------------
AG_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

AG_D void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
AG_D void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
AG_D void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
AG_D void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
AG_D void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}
AG_D void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
AG_D void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
AG_D void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
AG_D void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// bar.sync is an .aligned barrier: every thread of the warp must execute it together (a warp that reaches it diverged would be counted once
// per fragment), and the compiler does not know that about inline assembly — hence the explicit reconvergence.
AG_D void named_bar_sync(int id, int nthreads) { __syncwarp(); asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

AG_D void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
AG_D void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, single CTA
AG_D void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T: the A operand (M x K, K-major: row m in TMEM lane m, K elements packed two per 32-bit column) is read
// from tensor memory instead of shared memory — an SS-mode MMA with a small N is bound by the shared-memory read of its 4 KB A operand
AG_D void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one lane of a converged warp (issue tcgen05.mma from a warp-UNIFORM branch and elect inside it: from a divergent `lane == 0` branch
// every operand goes through an ELECT / R2UR / BRA.U.ANY waterfall, ~75 cycles per instruction)
AG_D bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
AG_D void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread
AG_D void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread
AG_D void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
AG_D void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B: start>>4 | LBO(16B units)=1 | SBO=1024B | version 1 | layout 2
AG_D uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D fp32, A/B bf16, both K-major, M=128, N
// (a_format/b_format: 0 = F16, 1 = BF16)
template <int FMT> AG_D uint32_t umma_idesc(int n) {
  const uint32_t f = (FMT == 0) ? 1u : 0u;
  return (1u << 4) | (f << 7) | (f << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_TILE_M >> 4) << 24);
}

// two fp32 -> one packed pair of MMA operands (element `lo` in the low half).  FMT 0: bf16 (RNE).  FMT 1: fp16 (RNE, saturating
// to +-65504 so that an outlier activation cannot become inf)
template <int FMT> AG_D uint32_t pack2(float lo, float hi) {
  uint32_t d;
  if (FMT == 0) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

// bits 0 and 1 of t (t < 2^16) -> one packed pair of MMA operands {bit 0 ? 1.0 : 0, bit 1 ? 1.0 : 0}; `one` = the 16-bit pattern of 1.0.
// t * 0x8001 puts bit 0 at position 0 and bit 1 at position 16; the mask keeps those two; the product with `one` cannot carry.
AG_D uint32_t bits2_to_operands(uint32_t t, uint32_t one) { return ((t * 0x8001u) & 0x00010001u) * one; }

struct TcArgs {
  const unsigned char* img;   // layer images, back to back
  const float* bias;          // [NH] head biases
  int nlayers;                // 1 + k + 1 (base, k residual blocks, heads)
  int k0_steps;               // UMMA K-steps of the base layer = ceil(2VS/16)
  int A;                      // actions
  int NH;                     // head N (multiple of 16)
  int in;                     // 2*VS
  long long* dbg;             // optional clock64 trace [cta][tile][layer][4] (development)
};


}  // namespace tc
}  // namespace ag
