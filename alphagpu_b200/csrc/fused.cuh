// fused.cuh — one persistent kernel per ply: the whole R-rollout loop of mcts_single (mcts_gpu.jl:396-439) on chip.
//
// A CTA of 512 threads (16 warps x 128 registers) owns up to 256 games for the entire search of a ply and alternates, per rollout,
//   search phase : ONE POOL of warp-sized work units drawn from a shared-memory counter — the (game, ancestor) items of backUp + the
//                  α re-solve (search.cuh: backup_item), listed game by game by the descents that produced them, and the expansion of
//                  the leaves (expand_game1), 32 games per unit — then, behind one barrier, the descent of the next rollout, one thread
//                  per game (select_game1).  Everything a phase hands to the next lives in shared memory (RolloutShared).
//   network phase: the tcgen05/TMEM chain of DenseNet.jl:294-304 on the leaves, 128 games per tile.  Two tiles (129..256 games): 8 warps
//                  per tile — TMEM lane quarter w%4, two 32-column slices each — with their own MMA issuer, tile 1 trailing tile 0 by one
//                  MMA phase so that one tile's epilogue runs under the other's MMAs.  One tile (33..128 games): all 16 warps on it, one
//                  slice each.  In both the A operand of every layer lives in TENSOR memory (encoder and epilogues write it with
//                  tcgen05.st, every layer is tcgen05.mma with A from TMEM): with both operands in shared memory a K-step was bound by
//                  the 8 KB it read there (~120 cycles against 64 of tensor time).  The fp32 residual stream lives in registers (a
//                  thread owns the same row and columns in every layer); weights stream global -> shared through a bulk-copy ring that
//                  never drains between rollouts.  Up to 32 games per CTA: the swapped kernel (SW = 1, below).
//                  (Measured alternatives, B200: all 16 warps alternating between the two tiles — 32 k cycles per rollout against 25 k,
//                  because issuing a layer's eight tcgen05.mma occupies the issuing warp for 600-1200 cycles and the other 15 wait for
//                  its share of the epilogue; the same with a 17th, issue-only warp — the register file then holds 20 warps x 96
//                  registers and the search phases pay for it; two independent 256-thread CTAs of one tile each per SM, rounds 1 and 2 —
//                  2.22 ms per full-load ply against 2.13: the co-resident CTAs do not hide each other's phases.)
// Games of a CTA depend on each other only through their shared GEMM tile, so there is no grid-wide barrier and no kernel boundary
// inside a ply: the per-rollout cost is the on-chip critical path instead of three launches plus their tails.
#pragma once
#include "search.cuh"
#include "tc_ptx.cuh"

namespace ag {

namespace fused {
using namespace tc;

// -DAG_TRACE=1 builds the development library whose per-ply kernel records clock64 phase traces into the buffer set by
// agpu_debug_tc_trace (scripts/fused_trace.py): per-warp busy time in the search pool and in the descent, phase totals of thread 0.
// -DAG_TRACE=2 adds the fine-grained stamps of thread 0 inside the descent, the backup items and the layers (they slow its warp down).
// The product library contains none of that code.
#ifndef AG_TRACE
#define AG_TRACE 0
#endif

// NT = tiles per CTA.  NT = 2: the full-load kernel (129..256 games per CTA).  NT = 1: at most 128 games per CTA, with a node cache
// (search.cuh: CacheSlot) in the shared memory the second tile would have used.  All stream the weights through a 2-stage ring (a third
// stage bought nothing and cost 32 KB of L1: the search phases live on the number of load/store requests and on where they hit, see
// search.cuh).
// SW = 1: the swapped kernel (one tile, at most 64 games per CTA).  SW = 0: ordinary orientation — every layer takes its A operand from
// tensor memory, so shared memory only holds the network's outputs of a tile, not its 32 KB activation image.
template <class G, int NT, int SW> struct FCfg {
  static_assert(SW == 0 || NT == 1, "the swapped orientation is a one-tile kernel");
  static constexpr int THREADS = 512;
  static constexpr int WPT = 16 / NT;                                  // warps per tile
  static constexpr int CPW = 16 / WPT;                                 // 32-column slices per warp
  static constexpr int GAMES = NT * TC_TILE_M;                         // games per CTA (capacity)
  static constexpr int STAGES = 2;
  static constexpr int PER_GAME = 2 * (int)sizeof(typename G::State) + 16 + 8 + 4 + 2 * ITEMS_PER_GAME + 2 + 2 * PATH_SMEM_DEPTH;
  static constexpr int WORK = 1024 + GAMES * ((PER_GAME + 15) / 16 * 16);   // barriers, counters, biases + the per-game hand-off
  static constexpr int TREE_BYTES = NT == 1 ? 64 * 1024 : 0;
  static constexpr int OUT_BYTES = (TC_TILE_M * Layout<G>::OUTS * 4 + 1023) / 1024 * 1024;   // the network's outputs of one tile
  static constexpr int A_BYTES = SW ? TC_A_BYTES : OUT_BYTES;          // per tile: activation image (+ outputs parked in it) or the outputs alone
  static constexpr int SMEM = NT * A_BYTES + STAGES * TC_W_STAGE_BYTES + 1024 + WORK + TREE_BYTES;   // + 1 KB alignment slack
  // tensor memory: 128 accumulator columns per tile (the fp32 residual stream is in registers); the small-batch kernel, in the swapped
  // orientation, keeps the trunk weights — the A operand there — resident in columns 64..511 (TW_COL0 + 64 per layer)
  static constexpr int TMEM_COLS = 512;
  static constexpr int TW_COL0 = 64, TW_MAX_LAYERS = 7;
  static_assert(SMEM <= 227 * 1024, "shared memory per CTA");
  // the carve-out steps of the SM (…, 100, 132, 164, 196, 228 KB; 1 KB of each CTA is reserved): the search phases live on the L1 that is
  // left — one step more cost the two-tile kernel 8 % per full-load ply when it happened
  static_assert(NT != 2 || SMEM + 1024 <= 132 * 1024, "two-tile kernel: stay inside the 132 KB shared-memory carve-out");
  static_assert(NT != 1 || SW != 0 || SMEM + 1024 <= 164 * 1024, "one-tile ordinary kernel: stay inside the 164 KB carve-out");
};

AG_D void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
AG_D void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
AG_D void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
AG_D void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
AG_D void tmem_ldn(uint32_t taddr, uint32_t (&v)[8]) { tmem_ld8(taddr, v); }
AG_D void tmem_ldn(uint32_t taddr, uint32_t (&v)[16]) { tmem_ld16(taddr, v); }

// Epilogue of a trunk layer computed in the SWAPPED orientation (few games per CTA): the accumulator holds out-feature f in TMEM
// lane f and game n in column n.  This thread owns feature 32*wq + lane and the NC games of slice cs in every layer, so the fp32
// residual stream stays in its registers (sres); it applies relu / the residual and scatters the 16-bit operand of the next layer into
// the ordinary games x features, K-major, 128B-swizzled tile (so the head layer and the encoder need no second layout).
template <int FMT, int NC>
AG_D void epilogue_swapped(uint32_t tmem_acc, int wq, int cs, int lane, unsigned char* At, uint32_t (&sres)[16]) {
  uint32_t va[NC];
  const uint32_t taddr = ((uint32_t)(wq * 32) << 16) + (uint32_t)(cs * NC);
  tmem_ldn(tmem_acc + taddr, va);
  tmem_ld_wait();
#pragma unroll
  for (int e = 0; e < NC; e++) {
    sres[e] = __float_as_uint(__uint_as_float(sres[e]) + fmaxf(__uint_as_float(va[e]), 0.f));   // (sres starts at zero before the base layer)
  }
  const int f = 32 * wq + lane;
  unsigned char* base = At + (f >> 6) * TC_KTILE_BYTES_A + (f & 7) * 2;
  const int c = (f & 63) >> 3;
#pragma unroll
  for (int e = 0; e < NC; e++) {
    const int n = cs * NC + e;
    *reinterpret_cast<uint16_t*>(base + n * 128 + ((c ^ (n & 7)) << 4)) = (uint16_t)(pack2<FMT>(__uint_as_float(sres[e]), 0.f) & 0xFFFFu);
  }
}

// Epilogue of a trunk layer in the ordinary orientation: row r = 32*wq + lane of the tile, NSL consecutive 32-column slices from cs0.
// b = relu(acc) (base layer) or b + relu(acc); the fp32 residual stream stays in this thread's registers for the whole chain (res: it
// owns the same row and columns in every layer; they are dead registers of the search phases); next A operand = fp16/bf16(b), written
// to TENSOR memory: in the ordinary orientation every layer is tcgen05.mma with A from TMEM.  With both operands in shared memory an
// M128 x N128 x K16 step is bound by the 8 KB it reads from there (~120 cycles per step measured, against 64 of tensor time); with A
// from tensor memory the shared-memory side only delivers the 4 KB of weights.
// The tensor-memory loads of a 16-column chunk are issued before the previous chunk is processed (tcgen05.wait::ld waits for ALL
// outstanding loads, so without this the load latency is exposed once per chunk: 4 or 8 times per layer).
template <int FMT, int NSL>
AG_D void epilogue_ordinary(uint32_t tmem_acc, uint32_t tmem_a, int wq, int cs0, f32x2 (&res)[16 * NSL]) {
  const uint32_t lane_sel = ((uint32_t)(wq * 32) << 16);
  const uint32_t acc0 = tmem_acc + lane_sel + (uint32_t)(cs0 * 32);
  const uint32_t a0 = tmem_a + lane_sel + (uint32_t)(cs0 * 16);          // two operands per 32-bit column
  constexpr int NCH = 2 * NSL;                                          // 16-column chunks
  const f32x2 half2 = pack2f(0.5f, 0.5f);
  uint32_t va[2][16];
  tmem_ld16(acc0, va[0]);
#pragma unroll
  for (int i = 0; i < NCH; i++) {
    const int b = i & 1;
    tmem_ld_wait();                                                     // chunk i has arrived
    if (i + 1 < NCH) tmem_ld16(acc0 + 16 * (i + 1), va[b ^ 1]);         // chunk i + 1 in flight while chunk i is processed
    // res += relu(acc), two columns per FFMA2: a + |a| is 2 relu(a) exactly, and fma(2 relu(a), 0.5, res) rounds once, like the add — one
    // FADD per column and one packed FFMA per pair on the FMA pipe instead of an FMNMX (half-rate ALU pipe) and an FADD per column.
    // (The caller zeroes res before the base layer: no select per element is spent on l == 0.)
    uint32_t pk[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
      const float x0 = __uint_as_float(va[b][2 * e]), x1 = __uint_as_float(va[b][2 * e + 1]);
      res[8 * i + e] = fma2(pack2f(__fadd_rn(x0, fabsf(x0)), __fadd_rn(x1, fabsf(x1))), half2, res[8 * i + e]);
      float lo, hi;
      unpack2f(res[8 * i + e], lo, hi);
      pk[e] = pack2<FMT>(lo, hi);
    }
    tmem_st8(a0 + 8 * i, pk);                                           // the next layer's A operand: row r in lane r, two operands per column
  }
  tmem_st_wait();
}

// development trace: BAR.SYNC.DEFER_BLOCKING does not block at issue but at the next consumer, so a clock read placed right behind a
// barrier is early; a volatile shared-memory load in between makes the stamp mean "after the barrier"
AG_D long long clock_after_barrier(const void* smem_word) {
  uint32_t d;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(d) : "r"(smem_u32(smem_word)) : "memory");
  return clock64() + (long long)(d & 0u);
}

template <class G, int FMT, int NT, int SW>
__global__ void __launch_bounds__(FCfg<G, NT, SW>::THREADS, 1) ply_kernel(SearchParams P, TcArgs T, SegParams S, int visits, int gpc) {
  typedef Layout<G> Lay;
  typedef FCfg<G, NT, SW> C;
  typedef typename G::State State;
  constexpr bool SMALL = NT == 1;                                      // the small-batch kernel: swapped orientation, node cache
  constexpr int W = Lay::W;
  constexpr int STAGES = C::STAGES;
  static_assert(Lay::FAST && G::Geo::NC == 1 && 2 * G::VS <= TC_N, "fused ply kernel: small boards only");
  static_assert(C::GAMES <= 256, "items carry the local game in 8 bits");
  // gpc = games per CTA (<= 256), chosen by the host so that the live games spread over all SMs: late plies run many lightly filled
  // CTAs rather than a few full ones.
  const int cta_first = (int)blockIdx.x * gpc;                         // first local slot of this CTA
  if (cta_first >= S.len) return;
  const int count = min(gpc, S.len - cta_first);                       // games of this CTA
  const int g0 = S.off + cta_first;                                    // global slot of local game 0
  const int ntiles = (NT == 2 && count > TC_TILE_M) ? 2 : 1;           // a CTA with <= 128 games runs a single tile
  // Few games per CTA (the long tail of a generation): the trunk layers run as D^T = W * X^T — out-features on the M = 128 side, the
  // NS = 32 games on the N side — so the epilogue shrinks with the batch instead of paying for 128 rows (up to 32 games per CTA; from 33
  // on the ordinary one-tile kernel, whose A operand also lives in tensor memory, is faster: 1.17 against 1.28 ms per ply at 9 k games).
  // The weight image (out x in, K-major) serves as the A operand unchanged and the activation tile as the B operand unchanged.
  constexpr bool swapped = SW != 0;                                    // (the host launches this kernel for at most 32 games per CTA)
  constexpr int NS = 32;
  // ... and then the weights are the A operand: resident in tensor memory for the whole ply (tcgen05.mma with A from TMEM) when the trunk
  // fits its 448 spare columns, instead of streamed through shared memory for every rollout
  const bool ts_mode = swapped && T.nlayers - 1 <= C::TW_MAX_LAYERS;
  if (swapped && count > NS) return;                                    // (never launched that way)

  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = smem;                                            // [NT][32 KB] activations (A operands)
  unsigned char* sW = smem + NT * C::A_BYTES;                          // [STAGES][32 KB] weight ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + STAGES * TC_W_STAGE_BYTES);
  // bars[0..2] full, [3..5] empty, [6..7] mma_done per tile, [8] stagger (one-shot)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  int* s_next = reinterpret_cast<int*>(bars + 17);                     // [2] work-unit counters of the search pool, by rollout parity (+ [2] development)
  int* s_lvcnt = reinterpret_cast<int*>(bars + 19);                    // [2] items listed by the last descent, by rollout parity
  float* sbias = reinterpret_cast<float*>(bars + 20);                  // [128] head biases
  static_assert(20 * 8 + TC_N * 4 <= 1024, "fixed part of the work area");
  // hand-off between the phases of a rollout (search.cuh: RolloutShared)
  RolloutShared<G> SH;
  SH.state = reinterpret_cast<State*>(reinterpret_cast<unsigned char*>(bars) + 1024);      // 16-byte aligned
  SH.root = SH.state + C::GAMES;
  SH.rnd = reinterpret_cast<Philox4*>(SH.root + C::GAMES);
  SH.hdr = reinterpret_cast<NodeHdr*>(SH.rnd + C::GAMES);
  SH.d = reinterpret_cast<int*>(SH.hdr + C::GAMES);
  SH.lv_item = reinterpret_cast<uint16_t*>(SH.d + C::GAMES);
  SH.lv_cap = ITEMS_PER_GAME * C::GAMES;
  SH.lv_cnt = s_lvcnt;
  SH.leaf = reinterpret_cast<uint8_t*>(SH.lv_item + ITEMS_PER_GAME * C::GAMES);
  SH.ovf = SH.leaf + C::GAMES;
  SH.path = reinterpret_cast<uint16_t*>(SH.ovf + C::GAMES);
  static_assert(sizeof(State) % 8 == 0 && sizeof(Philox4) == 16 && sizeof(NodeHdr) == 8, "hand-off layout");
  // the network's outputs go where the tile's A operand lived: it is dead from the head MMA until the next rollout's encoder, and
  // the search phase reads the outputs in between
  SH.out = reinterpret_cast<float*>(sA);
  SH.out_tile_stride = C::A_BYTES / 4;
  static_assert(TC_TILE_M * Lay::OUTS * 4 <= TC_KTILE_BYTES_A, "the network's outputs live in the first K tile of the idle A tile");
  // node cache (small-batch kernel): the first nc_nodes nodes of each of this CTA's games; the fewer games, the deeper the cache
  SH.nc_base = reinterpret_cast<unsigned char*>(bars) + C::WORK;
  SH.nc_nodes = SMALL ? min(P.R, C::TREE_BYTES / (CacheSlot<Lay::APAD>::BYTES * count)) : 0;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool dbg_on = AG_TRACE && T.dbg != nullptr;
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + 3), bar_done = smem_u32(bars + 6), bar_stagger = smem_u32(bars + 8);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, ntiles); }
    for (int t = 0; t < NT; t++) mbar_init(bar_done + 8 * t, 1);
    mbar_init(bar_stagger, 1);
    fence_barrier_init();
    s_next[0] = s_next[1] = s_next[2] = s_next[3] = 0;
    s_lvcnt[0] = s_lvcnt[1] = 0;
  }
  if (threadIdx.x < TC_N) sbias[threadIdx.x] = T.bias[threadIdx.x];
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), C::TMEM_COLS);        // NT x 128 accumulator columns + NT x 128 residual columns
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int nlayers = T.nlayers;
  const int total_layers = visits * nlayers;
  // weight image of global layer index wl (= rollout * nlayers + layer) -> ring stage wl % STAGES
  auto load_layer = [&](int wl) {
    const int s = wl % STAGES, l = wl % nlayers;
    const uint32_t bytes = (l == nlayers - 1) ? (uint32_t)(T.NH * TC_N * 2) : (uint32_t)TC_W_STAGE_BYTES;
    if (wl >= STAGES) mbar_wait(bar_empty + 8 * s, ((wl / STAGES) - 1) & 1);
    mbar_expect_tx(bar_full + 8 * s, bytes);
    bulk_g2s(smem_u32(sW + s * TC_W_STAGE_BYTES), T.img + (size_t)l * TC_W_STAGE_BYTES, bytes, bar_full + 8 * s);
  };

  // ---- roles ----
  // network: tile, TMEM lane quarter, first 32-column slice of this warp; its thread carries row r of its tile
  constexpr int WPT = C::WPT, CPW = C::CPW;
  const int t = warp / WPT, wq = warp & 3, csb = ((warp >> 2) & (WPT / 4 - 1)) * CPW;
  const int r = wq * 32 + lane;
  unsigned char* At = sA + t * C::A_BYTES;                             // (the activation image: swapped kernel only)
  const uint32_t tmem_acc = tmem_base + (uint32_t)(t * TC_N);
  // ordinary orientation: the A operand (activations, 128 rows x 128 operands = 64 columns) of tile t lives in tensor memory behind the
  // accumulators (the small-batch kernel's resident weights occupy those columns only in the swapped orientation)
  constexpr int TA_COL0 = NT * TC_N;
  const uint32_t tmem_a = tmem_base + (uint32_t)(TA_COL0 + 64 * t);
  // The MMA-issuing warp of a tile takes a WARP-UNIFORM branch and elects one lane inside it; every operand of tcgen05.mma is derived
  // from values the compiler can see as uniform (the broadcast warp index, the broadcast TMEM base).  Issued from a divergent
  // `lane == 0` branch each MMA went through an ELECT / 5 x R2UR / BRA.U.ANY waterfall: ~75 cycles per instruction, 600 per layer.
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const int t_u = warp_u / WPT;
  const bool issuer_warp = (warp_u % WPT) == 0;
  const uint32_t tmem_base_u = __shfl_sync(0xffffffffu, tmem_base, 0);
  const uint32_t one = (FMT == 0) ? 0x3F80u : 0x3C00u;

  if (ts_mode) {
    // trunk weights -> tensor memory.  Layer l lives in columns TW_COL0 + 64 l .. + 63: out-feature (row) m in lane m, the 128 inputs packed
    // two per column.  The global image is the shared-memory operand image (K-major, 128B-swizzled, two 64-input tiles of 128 rows x
    // 128 bytes): this thread un-swizzles row 32 wq + lane of the layers l = warp / 4, warp / 4 + 4.
    for (int l = warp >> 2; l < nlayers - 1; l += 4) {
      const unsigned char* img = T.img + (size_t)l * TC_W_STAGE_BYTES;
#pragma unroll
      for (int q = 0; q < 4; q++) {                                    // 16 columns = 32 inputs = 4 chunks of 16 bytes
        uint32_t v[16];
#pragma unroll
        for (int c4 = 0; c4 < 4; c4++) {
          const int c = (q & 1) * 4 + c4;                              // chunk within the 64-input tile q >> 1
          const uint4 x = *reinterpret_cast<const uint4*>(img + (q >> 1) * TC_KTILE_BYTES_A + r * 128 + ((c ^ (r & 7)) << 4));
          v[4 * c4] = x.x; v[4 * c4 + 1] = x.y; v[4 * c4 + 2] = x.z; v[4 * c4 + 3] = x.w;
        }
        tmem_st16(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(C::TW_COL0 + 64 * l + 16 * q), v);
      }
    }
    tmem_st_wait();
    tc_fence_before();
  }

  // one thread per game for the descent: the game's uid and node count stay in its registers for the whole ply; the root's state and the
  // first Philox block of the coming descent wait in shared memory
  // Up to 112 games per CTA they are dealt round-robin to the warps (local game lane * 16 + warp), so that all 16 warps descend: the
  // phase lasts as long as the deepest game of the CTA, and with a few warps of 32 games its levels take longer (B200, per ply:
  // -3..4 % below 12 k live games, +2 % at full load, where the packed assignment stays)
  const int my_gl = (SMALL && gpc <= 112) ? lane * (C::THREADS / 32) + warp : (int)threadIdx.x;
  const bool has_game = my_gl < count;
  const unsigned game_mask = __ballot_sync(0xffffffffu, has_game);      // the lanes of this warp that descend (a prefix of the warp)
  const int my_g = g0 + my_gl;
  const u32 my_uid = has_game ? P.uid[my_g] : 0u;
  int my_nn = has_game ? P.nnodes[my_g] : 0;
  if (has_game) {
    SH.root[my_gl] = *reinterpret_cast<const State*>(P.tree + (size_t)my_g * P.game_stride + Lay::state_off(P.R, 0));
    SH.rnd[my_gl] = philox4x32_10(my_uid, S.ply, 0u, 0u, (u32)S.seed, (u32)(S.seed >> 32));
    SH.d[my_gl] = 0;
    SH.ovf[my_gl] = 0;
    if (SMALL) {
      // fill the node cache with the nodes that exist when the ply starts (the root alone after root_reset); read back only by this thread
      typedef CacheSlot<Lay::APAD> CS;
      const char* gb = P.tree + (size_t)my_g * P.game_stride;
      for (int nd = 0; nd < min(my_nn, SH.nc_nodes); nd++) {
        unsigned char* sl = node_cache_slot<G, true>(SH, my_gl, nd);
        const char* rec = gb + (size_t)nd * Lay::REC;
        *reinterpret_cast<uint2*>(sl) = *reinterpret_cast<const uint2*>(rec + Lay::OFF_HDR);
        for (int c = 0; c < Lay::APAD / 8; c++) *reinterpret_cast<uint2*>(sl + CS::OFF_CHILD + 8 * c) = *reinterpret_cast<const uint2*>(rec + Lay::OFF_CHILD + 8 * c);
        for (int c = 0; c < Lay::APAD / 4; c++) *reinterpret_cast<float4*>(sl + CS::OFF_POLICY + 16 * c) = *reinterpret_cast<const float4*>(rec + Lay::OFF_POLICY + 16 * c);
      }
    }
  }
  __syncthreads();

  // issue of one layer's MMAs for tile t (called by the issuer warp only, all lanes, warp-uniform arguments)
  auto issue_layer = [&](const int t, const int l, const int wl_) {
    const int s = ts_mode ? 0 : wl_ % STAGES;
    const bool is_head = (l == nlayers - 1);
    const int nl = is_head ? T.NH : TC_N;
    tc_fence_after();
    if (elect_one()) {
      const uint32_t tmem_acc_u = tmem_base_u + (uint32_t)(t * TC_N);
      const uint64_t ad0 = umma_desc(smem_u32(sA) + (uint32_t)(t * C::A_BYTES));
      const uint64_t bd0 = umma_desc(smem_u32(sW) + (uint32_t)(s * TC_W_STAGE_BYTES));
      const uint64_t bstep = (uint64_t)((nl * 128) >> 4);
      if (SMALL && ts_mode && !is_head) {
        const uint32_t idesc = umma_idesc<FMT>(NS);                     // M = 128 out-features, N = NS games
        const uint32_t tw = tmem_base_u + (uint32_t)(C::TW_COL0 + 64 * l);
#pragma unroll
        for (int ks = 0; ks < TC_N / 16; ks++) {
          const uint64_t ainc = (uint64_t)(((ks >> 2) * TC_KTILE_BYTES_A + (ks & 3) * 32) >> 4);
          umma_ts(tmem_acc_u, tw + (uint32_t)(8 * ks), ad0 + ainc, idesc, ks > 0 ? 1u : 0u);   // weights (TMEM) as A, activations as B
        }
      } else if (SMALL && swapped && !is_head) {
        const uint32_t idesc = umma_idesc<FMT>(NS);                     // M = 128 out-features, N = NS games
#pragma unroll
        for (int ks = 0; ks < TC_N / 16; ks++) {
          const uint64_t ainc = (uint64_t)(((ks >> 2) * TC_KTILE_BYTES_A + (ks & 3) * 32) >> 4);
          const uint64_t binc = (uint64_t)(((ks & 3) * 32) >> 4) + ((ks >> 2) ? bstep : 0);
          umma_bf16(tmem_acc_u, bd0 + binc, ad0 + ainc, idesc, ks > 0 ? 1u : 0u);   // weights as A, activations as B
        }
      } else if (!swapped) {
        const uint32_t idesc = umma_idesc<FMT>(nl);                     // ordinary orientation, every layer: activations (TMEM) as A, weights as B
        const uint32_t ta = tmem_base_u + (uint32_t)(TA_COL0 + 64 * t);
#pragma unroll
        for (int ks = 0; ks < TC_N / 16; ks++) {
          const uint64_t binc = (uint64_t)(((ks & 3) * 32) >> 4) + ((ks >> 2) ? bstep : 0);
          umma_ts(tmem_acc_u, ta + (uint32_t)(8 * ks), bd0 + binc, idesc, ks > 0 ? 1u : 0u);
        }
      } else {
        const uint32_t idesc = umma_idesc<FMT>(nl);                     // the head layer behind a swapped trunk: both operands in shared memory
#pragma unroll
        for (int ks = 0; ks < TC_N / 16; ks++) {
          const uint64_t ainc = (uint64_t)(((ks >> 2) * TC_KTILE_BYTES_A + (ks & 3) * 32) >> 4);
          const uint64_t binc = (uint64_t)(((ks & 3) * 32) >> 4) + ((ks >> 2) ? bstep : 0);
          umma_bf16(tmem_acc_u, ad0 + ainc, bd0 + binc, idesc, ks > 0 ? 1u : 0u);
        }
      }
      umma_commit(bar_done + 8 * t);
      umma_commit(bar_empty + 8 * s);                                   // the stage is free once every tile's MMAs of this layer are done
    }
    __syncwarp();
  };

  if (threadIdx.x == 32) {
    if (ts_mode) {                                                     // only the head layer is read from shared memory: loaded once, stage 0
      const uint32_t bytes = (uint32_t)(T.NH * TC_N * 2);
      mbar_expect_tx(bar_full, bytes);
      bulk_g2s(smem_u32(sW), T.img + (size_t)(nlayers - 1) * TC_W_STAGE_BYTES, bytes, bar_full);
    } else {                                                           // fill the ring: STAGES - 1 layers ahead
      for (int i = 0; i < STAGES - 1 && i < total_layers; i++) load_layer(i);
    }
  }
  int wl = 0;                                                          // global layer counter (ring / barrier phases)
  long long t_ly[5] = {0, 0, 0, 0, 0};                                  // development trace of thread 0: weights wait, MMA issue, MMA done, epilogue, barrier
  long long t_ph[5] = {0, 0, 0, 0, 0}, t_mark = dbg_on ? clock64() : 0;   // development trace (agpu_debug_tc_trace): -, -, search pool, descent, network
  for (int k = 0; k < visits; k++) {
    const int last = (k == visits - 1);
    // ================= search phase =================
    if (k > 0) {
      // The pool: backup items first (the longer units), then the expansions.  The items were listed, level by level, by the descents of
      // rollout k-1 (SH.lv_item); a unit is 32 consecutive items or the leaves of 32 consecutive games.  A (game, ancestor) item and the
      // expansion of that game's leaf touch different nodes, and both only read what the network and the descent left in shared memory,
      // so the units are independent.
      const int par = (k - 1) & 1;
      if (threadIdx.x == 0) { s_next[par ^ 1] = 0; s_lvcnt[par ^ 1] = 0; }   // counters of the other parity: for the coming descent / the next pool
      const long long w_t0 = dbg_on ? clock64() : 0;                   // development trace: per-warp busy time in the pool
      const int n_items = min(s_lvcnt[par], SH.lv_cap);
      const int UB = (n_items + 31) >> 5, UE = (count + 31) >> 5;
      while (true) {
        int u = 0;
        if (lane == 0) u = atomicAdd(&s_next[par], 1);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= UB + UE) break;
        if (dbg_on && lane == 0) atomicAdd(&s_next[2], 1);                // development check: units started ...
        if (u < UB) {
          // backUp + re-solve of π̄: one (game, ancestor) item per thread
          const int i = u * 32 + lane;
          if (i < n_items) {
            const int item = SH.lv_item[i];
            const int gl = item & 0xFF, jj = item >> 8;
            const LeafEval E = leaf_eval1<G>(SH, gl);
            backup_item<G>(P, g0 + gl, jj, SH.d[gl], E, 0, S.cpuct, (AG_TRACE >= 2 && dbg_on && threadIdx.x == 0) ? T.dbg + blockIdx.x * 128 + 8 : nullptr,
                           SH.path + gl * PATH_SMEM_DEPTH,
                           SH.nc_nodes > 0 ? SH.nc_base + (size_t)gl * SH.nc_nodes * CacheSlot<Lay::APAD>::BYTES : nullptr, SH.nc_nodes);
          }
        } else {
          const int gl = (u - UB) * 32 + lane;
          if (gl < count) expand_game1<G, SMALL>(P, g0 + gl, gl, SH, S.training, 0);
        }
        __syncwarp();
        if (dbg_on && lane == 0) atomicAdd(&s_next[3], 1);                // ... and finished
      }
      if (has_game && SH.ovf[my_gl] < SH.d[my_gl]) {                   // the list was full: this game backs up the rest of its path itself
        const int gl = my_gl;
        const LeafEval E = leaf_eval1<G>(SH, gl);
        for (int jj = SH.ovf[gl]; jj < SH.d[gl]; jj++)
          backup_item<G>(P, g0 + gl, jj, SH.d[gl], E, 0, S.cpuct, nullptr, SH.path + gl * PATH_SMEM_DEPTH,
                         SH.nc_nodes > 0 ? SH.nc_base + (size_t)gl * SH.nc_nodes * CacheSlot<Lay::APAD>::BYTES : nullptr, SH.nc_nodes);
      }
      const long long w_d0 = dbg_on ? clock64() - w_t0 : 0;
      named_bar_sync(1, C::THREADS);
      if (dbg_on && lane == 0) T.dbg[blockIdx.x * 128 + 32 + warp] += w_d0;
      if (dbg_on && threadIdx.x == 0) { const long long c = clock64(); t_ph[2] += c - t_mark; t_mark = c; }
    }
    // descent of this rollout; its path nodes are listed as the items of the next pool (parity k)
    SH.lv_cnt = s_lvcnt + (k & 1);
    const long long w_t1 = dbg_on ? clock64() : 0;                      // development trace: per-warp time in the descent
    if (dbg_on && has_game && lane == 0 && s_next[2] != s_next[3]) T.dbg[blockIdx.x * 128 + 7] += 1;   // a descent started while a pool unit was still running
    if (has_game) select_game1<G, SMALL>(P, my_g, my_gl, SH, my_uid, my_nn, k, last, S.seed, S.ply, game_mask, (AG_TRACE >= 2 && dbg_on && threadIdx.x == 0) ? T.dbg + blockIdx.x * 128 + 8 : nullptr);
    const long long w_d1 = dbg_on ? clock64() - w_t1 : 0;
    named_bar_sync(1, C::THREADS);
    if (dbg_on && lane == 0) T.dbg[blockIdx.x * 128 + 48 + warp] += w_d1;
    if (dbg_on && threadIdx.x == 0) { const long long c = clock64(); t_ph[3] += c - t_mark; t_mark = c; }

    // ================= network phase =================
    const bool obs = AG_TRACE >= 2 && dbg_on && threadIdx.x == 256;    // development trace: a non-issuing warp's view of the network phase
    long long ob0 = 0;
    if (obs) ob0 = clock_after_barrier(s_next);
    {
      // A operand of the base layer (decoder, mcts_gpu.jl:202-223): this thread's operand columns of its row
      u64 x0 = 0, x1 = 0;
      const int gl = t * TC_TILE_M + r;
      if (gl < count) {
        const u64* st = reinterpret_cast<const u64*>(SH.state + gl);      // left there by this rollout's descent
        const u64 bp = st[0], bo = st[1];
        constexpr int VS = G::VS;
        x0 = (VS < 64) ? (bp | (bo << VS)) : bp;
        x1 = (VS < 64) ? (bo >> (64 - VS)) : bo;
      }
#pragma unroll
      for (int j = 0; j < CPW; j++) {
        const int cs = csb + j;
        // (warp-uniform) columns beyond the input: zero weights in the base layer's image, zeros written by the first rollout, finite
        // activations afterwards — nothing to encode.  In shared memory columns 0..63 are always rewritten: the network's outputs
        // (fp32, any bit pattern) were parked in the first K tile.  (Skipping the K-steps in the MMA sequence instead costs far more than
        // it saves: a branch between two tcgen05.mma stalls the issue, +2.8 ms per generation on B200.)
        if (k > 0 && 32 * cs >= max(16 * T.k0_steps, 64)) continue;
        const uint32_t bits = (uint32_t)(((cs & 2) ? x1 : x0) >> (32 * (cs & 1)));
        uint32_t w[16];                                                // the 32 operands of the slice, two per word
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const uint32_t byte = (bits >> (8 * i)) & 0xFFu;
#pragma unroll
          for (int e = 0; e < 4; e++) w[4 * i + e] = bits2_to_operands(byte >> (2 * e), one);
        }
        if (!swapped) {                                                // ordinary orientation: the A operand lives in tensor memory
          tmem_st16(tmem_a + ((uint32_t)(wq * 32) << 16) + (uint32_t)(16 * cs), w);
        } else {
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const int c = 4 * cs + i;                                  // chunk of 8 operands in the row, 0..15
            *reinterpret_cast<uint4*>(At + (c >> 3) * TC_KTILE_BYTES_A + r * 128 + (((c & 7) ^ (r & 7)) << 4)) = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
          }
        }
      }
      if (!swapped) { tmem_st_wait(); tc_fence_before(); }
    }
    if (t < ntiles) {                                                  // an idle tile rejoins at the end-of-rollout barrier
      fence_proxy_async();
      named_bar_sync(NT == 1 ? 2 : 2 + t, 32 * WPT);   // (a compile-time barrier id where there is one tile)
      if (obs) { const long long c = clock_after_barrier(s_next); T.dbg[blockIdx.x * 128 + 64] += c - ob0; ob0 = c; }
      uint32_t sres[16];                                               // this thread's residual values (swapped orientation)
#pragma unroll
      for (int e = 0; e < 16; e++) sres[e] = 0u;
      f32x2 rres[16 * CPW];                                            // ... and in the ordinary orientation, as pairs (starts at zero)
#pragma unroll
      for (int e = 0; e < 16 * CPW; e++) rres[e] = 0ull;
      for (int l = 0; l < nlayers; l++) {
        const int wll = wl + l;
        const int s = wll % STAGES;
        const bool is_head = (l == nlayers - 1);
        long long lt0 = 0, lt1 = 0, lt2 = 0;
        const bool ltr = AG_TRACE >= 2 && dbg_on && threadIdx.x == 0 && !is_head;
        if (ltr) lt0 = clock64();
        if (issuer_warp) {
          if (!ts_mode) mbar_wait(bar_full + 8 * s, (wll / STAGES) & 1);
          else if (is_head) mbar_wait(bar_full, 0);                      // the resident head image (complete after the first wait)
          if (ltr) lt1 = clock64();
          if (wll == 0 && t_u == 1) mbar_wait(bar_stagger, 0);          // tile 1 trails tile 0 by one MMA phase
          issue_layer(t_u, l, wll);
          if (wll == 0 && t_u == 0 && elect_one()) umma_commit(bar_stagger);
          __syncwarp();
          if (ltr) lt2 = clock64();
        }
        // the weights STAGES - 1 layers ahead are requested by a lane that would otherwise just wait for this layer's MMAs (on the
        // issuer the request sat on the critical path: 2 k cycles per rollout)
        if (!ts_mode && threadIdx.x == 32 && wll + STAGES - 1 < total_layers) load_layer(wll + STAGES - 1);
        // the Philox block of depths 0..3 of the NEXT descent, while the first MMAs run
        if (l == 0 && has_game) SH.rnd[my_gl] = philox4x32_10(my_uid, S.ply, (u32)(k + 1), 0u, (u32)S.seed, (u32)(S.seed >> 32));
        mbar_wait(bar_done + 8 * t, wll & 1);
        tc_fence_after();
        if (obs) { const long long c = clock64(); T.dbg[blockIdx.x * 128 + 66 + l] += c - ob0; ob0 = c; }
        long long lt3 = 0;
        if (ltr) { lt3 = clock64(); t_ly[0] += lt1 - lt0; t_ly[1] += lt2 - lt1; t_ly[2] += lt3 - lt2; }
        if (!is_head) {
          if (SMALL && swapped) {
            epilogue_swapped<FMT, 8>(tmem_acc, wq, csb, lane, At, sres);
          } else {
            epilogue_ordinary<FMT, CPW>(tmem_acc, tmem_a, wq, csb, rres);
          }
          tc_fence_before();
          fence_proxy_async();
          long long lt4 = 0;
          if (ltr) { lt4 = clock64(); t_ly[3] += lt4 - lt3; }
          if (obs) { const long long c = clock64(); T.dbg[blockIdx.x * 128 + 74 + l] += c - ob0; ob0 = c; }
          named_bar_sync(NT == 1 ? 2 : 2 + t, 32 * WPT);   // (a compile-time barrier id where there is one tile)
          if (obs) { const long long c = clock_after_barrier(s_next); T.dbg[blockIdx.x * 128 + 82 + l] += c - ob0; ob0 = c; }
          if (ltr) t_ly[4] += clock64() - lt4;
        } else {
          // heads: logits = acc + bias, value = σ(acc[A] + bias[A])   (DenseNet.jl:301) -> shared memory (read by the next search phase)
          const uint32_t lane_sel = ((uint32_t)(wq * 32) << 16) + (uint32_t)(csb * 32);
#pragma unroll
          for (int i = 0; i < 2; i++) {
            const int a0 = csb * 32 + i * 16;                            // NH <= 32: only a warp's first slice can hold head columns
            if (a0 < T.NH) {                                            // warp-uniform
              uint32_t v[16];
              tmem_ld16(tmem_acc + lane_sel + 16 * i, v);
              tmem_ld_wait();
              float z[16];
#pragma unroll
              for (int e = 0; e < 16; e++) z[e] = __uint_as_float(v[e]) + sbias[a0 + e];
              if (T.A >= a0 && T.A < a0 + 16) {
#pragma unroll
                for (int e = 0; e < 16; e++) if (a0 + e == T.A) z[e] = c_sigmoidf(z[e]);
              }
              const int gl = t * TC_TILE_M + r;
              if (gl < count) {
                float* so = SH.out + t * SH.out_tile_stride + r * Lay::OUTS;
                float* o = P.nn_out + (size_t)(g0 + gl) * Lay::OUTS;
#pragma unroll
                for (int q4 = 0; q4 < 4; q4++)
                  if (a0 + 4 * q4 < Lay::OUTS) {
                    const float4 zv = make_float4(z[4 * q4], z[4 * q4 + 1], z[4 * q4 + 2], z[4 * q4 + 3]);
                    *reinterpret_cast<float4*>(so + a0 + 4 * q4) = zv;
                    if (last) *reinterpret_cast<float4*>(o + a0 + 4 * q4) = zv;   // the expand of the last rollout reads it from global memory
                  }
              }
            }
          }
          tc_fence_before();
        }
      }
    } else if (has_game) {                                             // (a game thread in a warp of the idle tile)
      SH.rnd[my_gl] = philox4x32_10(my_uid, S.ply, (u32)(k + 1), 0u, (u32)S.seed, (u32)(S.seed >> 32));
    }
    wl += nlayers;
    named_bar_sync(1, C::THREADS);                                     // the outputs are visible to the search phase
    if (obs) { const long long c = clock_after_barrier(s_next); T.dbg[blockIdx.x * 128 + 65] += c - ob0; }
    if (dbg_on && threadIdx.x == 0) { const long long c = clock64(); t_ph[4] += c - t_mark; t_mark = c; }
  }
  if (dbg_on && threadIdx.x == 0) {
    for (int i = 0; i < 5; i++) T.dbg[blockIdx.x * 128 + i] = t_ph[i];
    T.dbg[blockIdx.x * 128 + 5] = count; T.dbg[blockIdx.x * 128 + 6] = visits;
    for (int i = 0; i < 5; i++) T.dbg[blockIdx.x * 128 + 24 + i] = t_ly[i];
  }

  // expand + backUp of the last rollout (publishes nothing new for the root: policy_final was written by its descent)
  {
    const int sg = threadIdx.x / W, sl = threadIdx.x & (W - 1);
    const unsigned gm = group_mask<W>();
    constexpr int GROUPS = C::THREADS / W;
#pragma unroll 1
    for (int p = 0; p * GROUPS < count; p++) {
      const int gl = p * GROUPS + sg;
      if (gl < count) expand_backup_game<G, false>(P, g0 + gl, sl, gm, S.training, 1, nullptr, nullptr, S.cpuct);
    }
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

}  // namespace fused
}  // namespace ag
