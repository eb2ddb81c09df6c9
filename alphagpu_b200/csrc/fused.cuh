// fused.cuh — one persistent kernel per ply: the whole R-rollout loop of mcts_single (mcts_gpu.jl:396-439) on chip.
//
// A CTA owns 256 games (two 128-row tiles) for the entire search of a ply and alternates, per rollout,
//   search phase : all 1024 threads, 8 lanes per game, two passes of 128 games: expand+backUp of the previous rollout, then the
//                  descent of this one (the same device functions as the stand-alone kernels, search.cuh);
//   network phase: the tcgen05/TMEM chain of nn_tc.cu re-organised for 32 warps: warp w serves tile w/16, TMEM lane quarter w%4
//                  and the 32-column slice (w/4)%4; one lane per tile issues the MMAs; the fp32 residual stream lives in TMEM
//                  (columns 256..511) so the epilogue needs few registers; weights stream global->shared through the 3-stage
//                  bulk-copy ring without ever draining between rollouts.
// Games of a CTA depend on each other only through their shared GEMM tile, so there is no grid-wide barrier and no kernel
// boundary inside a ply: the per-rollout cost is the on-chip critical path instead of three launches plus their tails
// (profiles/r01_ply_trace_*.txt: 70 us -> per rollout at 32768 games, 25-40 us floor at small L with separate launches).
#pragma once
#include "search.cuh"
#include "tc_ptx.cuh"

namespace ag {

namespace fused {
using namespace tc;

// -DAG_NHALF=1 (development variant, NOT YET RUN ON A GPU): with 8 warps per tile a trunk layer in the ordinary orientation is issued as
// two chains of 8 MMAs, output columns 0..63 and 64..127 (two barriers), and every warp owns one 32-column slice in EACH half: it runs
// the TMEM loads, the arithmetic and the residual store of its first slice while the tensor core still works on the second half; the
// operands computed from the first slice are stored only once that second chain — which reads the same tile as its A operand — is done.
// Every column is accumulated over K in the same order as before, so the results are bit-identical.
#ifndef AG_NHALF
#define AG_NHALF 0
#endif

// NT = tiles per CTA.  NT = 2: 1024 threads, one CTA per SM, 3-stage weight ring.  NT = 1: 512 threads, 2-stage ring, TWO CTAs per
// SM (98 KB shared memory, 256 TMEM columns, 64 registers each): while one CTA waits on its MMA chain the other's search phase
// uses the issue slots.
template <int NT, int WPT = 16> struct FCfg {
  static constexpr int THREADS = 32 * WPT * NT;                        // WPT warps per tile: 16 (one 32-column slice per warp) or 8 (two)
  static constexpr int GAMES = NT * TC_TILE_M;                         // games per CTA (capacity)
  static constexpr int STAGES = NT == 1 ? 2 : 3;
  static constexpr int ITEM_MAP = 1024;                                // backup items whose game is looked up in a byte map
  static constexpr int WORK = 2048 + GAMES * (44 + 72); // barriers/bias/backup work list + rollout hand-off
  static constexpr int WORK_USED = 640 + GAMES * 32 + GAMES * 4 + (GAMES + 1) * 4 + 16 + GAMES * 68 + ITEM_MAP;   // as laid out in the kernel
  static_assert(WORK_USED <= WORK, "shared-memory work area overflows its budget");
  // AG_TREE_SMEM (KB): node cache of the one-tile, 16-warp kernels (the small-batch variant of the tail), after the work area
  static constexpr int TREE_BYTES = (NT == 1 && WPT == 16) ? AG_TREE_SMEM * 1024 : 0;
  static constexpr int SMEM = NT * TC_A_BYTES + STAGES * TC_W_STAGE_BYTES + 1024 + WORK + TREE_BYTES;   // + 1 KB alignment slack
  static constexpr int TMEM_COLS = 256 * NT;                           // accumulators + fp32 residual stream
  static constexpr int CTAS_PER_SM = NT == 1 ? 2 : 1;
};

AG_D void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
AG_D void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
AG_D void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
AG_D void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
AG_D void tmem_ldn(uint32_t taddr, uint32_t (&v)[8]) { tmem_ld8(taddr, v); }
AG_D void tmem_ldn(uint32_t taddr, uint32_t (&v)[16]) { tmem_ld16(taddr, v); }
AG_D void tmem_stn(uint32_t taddr, const uint32_t (&v)[8]) { tmem_st8(taddr, v); }
AG_D void tmem_stn(uint32_t taddr, const uint32_t (&v)[16]) { tmem_st16(taddr, v); }

// Epilogue of a trunk layer computed in the SWAPPED orientation (few games per CTA): the accumulator holds out-feature f in TMEM
// lane f and game n in column n.  This thread owns feature 32*wq + lane and the NC games of slice cs in every layer, so the fp32
// residual stream stays in its registers (sres); it applies relu / the residual and scatters the 16-bit operand of the next layer into
// the ordinary games x features, K-major, 128B-swizzled tile (so the head layer and the encoder need no second layout).
// KS accumulator chains (AG_KSPLIT): the layer's 8 K-steps are dealt round-robin to KS accumulators, NS columns apart, so that
// consecutive tcgen05.mma instructions do not wait on each other's accumulator; the partial sums are added here, pairwise.
#ifndef AG_KSPLIT
#define AG_KSPLIT 1
#endif
template <int FMT, int NC, int KS>
AG_D void epilogue_swapped(uint32_t tmem_acc, int NS, int wq, int cs, int lane, int l, unsigned char* At, uint32_t (&sres)[16]) {
  static_assert(KS == 1 || KS == 2 || KS == 4, "accumulator chains");
  uint32_t va[KS][NC], vh[NC];
  const uint32_t taddr = ((uint32_t)(wq * 32) << 16) + (uint32_t)(cs * NC);
#pragma unroll
  for (int c = 0; c < KS; c++) tmem_ldn(tmem_acc + taddr + (uint32_t)(c * NS), va[c]);
  tmem_ld_wait();
#pragma unroll
  for (int e = 0; e < NC; e++) {
    float acc = __uint_as_float(va[0][e]);
    if (KS == 2) acc = acc + __uint_as_float(va[1][e]);
    if (KS == 4) acc = (acc + __uint_as_float(va[1][e])) + (__uint_as_float(va[2][e]) + __uint_as_float(va[3][e]));
    const float ra = fmaxf(acc, 0.f);
    const float hv = (l == 0) ? ra : __uint_as_float(sres[e]) + ra;
    sres[e] = vh[e] = __float_as_uint(hv);
  }
  const int f = 32 * wq + lane;
  unsigned char* base = At + (f >> 6) * TC_KTILE_BYTES_A + (f & 7) * 2;
  const int c = (f & 63) >> 3;
#pragma unroll
  for (int e = 0; e < NC; e++) {
    const int n = cs * NC + e;
    *reinterpret_cast<uint16_t*>(base + n * 128 + ((c ^ (n & 7)) << 4)) = (uint16_t)(pack2<FMT>(__uint_as_float(vh[e]), 0.f) & 0xFFFFu);
  }
}


// SW: the small-batch variant (host: games per CTA <= 128): one tile, 512 threads, one CTA per SM — 128 registers per thread instead
// of 64 — and, up to 64 games, the trunk layers in the swapped orientation (below).
template <class G, int FMT, int NT, bool SW = false, int WPT = 16>
__global__ void __launch_bounds__(FCfg<NT, WPT>::THREADS, (SW || WPT == 8) ? 1 : FCfg<NT, WPT>::CTAS_PER_SM) ply_kernel(SearchParams P, TcArgs T, SegParams S, int visits, int gpc) {
  static_assert(!SW || (NT == 1 && WPT == 16), "the swapped variant runs a single tile with 16 warps");
  static_assert(WPT == 16 || WPT == 8, "warps per tile");
  typedef Layout<G> Lay;
  typedef FCfg<NT, WPT> C;
  constexpr int CPW = 16 / WPT;                                        // 32-column slices per warp
  constexpr int W = Lay::W;
  constexpr int STAGES = C::STAGES;
  static_assert(Lay::FAST && G::Geo::NC == 1 && 2 * G::VS <= TC_N, "fused ply kernel: small boards only");
  // gpc = games per CTA (<= 256), chosen by the host so that the live games spread over all SMs: the search phase of a CTA is
  // issue-bound on its one SM, so late plies run many lightly filled CTAs rather than a few full ones.
  const int cta_first = (int)blockIdx.x * gpc;                         // first local slot of this CTA
  if (cta_first >= S.len) return;
  const int count = min(gpc, S.len - cta_first);                       // games of this CTA
  const int L_end = S.off + cta_first + count;                         // one past this CTA's last slot
  const int ntiles = (NT == 2 && count > TC_TILE_M) ? 2 : 1;           // a CTA with <= 128 games runs a single tile
  // Few games per CTA (the long tail of a generation): the trunk layers run as D^T = W * X^T — out-features on the M = 128 side, the
  // NS = 32 / 64 games on the N side — so the tensor time and the epilogue shrink with the batch instead of paying for 128 rows.
  // The weight image (out x in, K-major) serves as the A operand unchanged and the activation tile as the B operand unchanged.
  const bool swapped = SW && count <= 64;                              // CTA-uniform; 65..128 games keep the ordinary orientation
  const int NS = count <= 32 ? 32 : 64;

  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = smem;                                            // [2][32 KB] activations (A operands)
  unsigned char* sW = smem + NT * TC_A_BYTES;                          // [STAGES][32 KB] weight ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + STAGES * TC_W_STAGE_BYTES);
  // bars[0..2] full, [3..5] empty, [6..7] mma_done, [8] stagger (one-shot)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  float* sbias = reinterpret_cast<float*>(bars + 10);                  // [128] head biases
  // work list of the backup phase: per game of the CTA the leaf evaluation, the path length and its exclusive prefix sum
  LeafEval* s_eval = reinterpret_cast<LeafEval*>(bars + 80);           // [256]
  int* s_d = reinterpret_cast<int*>(s_eval + C::GAMES);                // [256]
  int* s_off = s_d + C::GAMES;                                         // [257]
  // hand-off between the phases of a rollout (search.cuh: RolloutShared), 128 bytes per game
  RolloutShared<G> SH;
  SH.state = reinterpret_cast<typename G::State*>((reinterpret_cast<uintptr_t>(s_off + C::GAMES + 1) + 15) & ~uintptr_t(15));   // 16-byte aligned (float4 reads of `out`)
  SH.hdr = reinterpret_cast<NodeHdr*>(SH.state + C::GAMES);
  SH.d = s_d;
  SH.leaf = reinterpret_cast<uint8_t*>(SH.hdr + C::GAMES);
  SH.pn = SH.leaf + C::GAMES;
  SH.pm = SH.pn + C::GAMES * PATH_SMEM_DEPTH;
  constexpr int ITEM_MAP = C::ITEM_MAP;                                // backup items whose game is looked up in a byte map (the rest: binary search)
  static_assert(sizeof(LeafEval) == 32 && sizeof(NodeHdr) == 8, "FCfg::WORK_USED assumes these sizes");
  uint8_t* s_item = SH.pm + C::GAMES * PATH_SMEM_DEPTH;                // [ITEM_MAP] item -> local game
  // the network's outputs go where the tile's A operand lived: it is dead from the head MMA until the next rollout's encoder, and
  // expand reads the outputs in between.  (Staying under 196 KB of shared memory keeps the next carve-out step — 32 KB of L1 — free.)
  SH.out = reinterpret_cast<float*>(sA);
  SH.out_tile_stride = TC_A_BYTES / 4;
  // node cache (AG_TREE_SMEM): the first nc_nodes nodes of each of this CTA's games; the fewer games, the deeper the cache
  SH.nc_base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(sW + STAGES * TC_W_STAGE_BYTES + C::WORK) + 15) & ~uintptr_t(15));
  SH.nc_nodes = C::TREE_BYTES > 0 ? min(P.R, (C::TREE_BYTES - 16) / (RootSlot<Lay::APAD>::BYTES * count)) : 0;
  static_assert(TC_TILE_M * Lay::OUTS * 4 <= TC_A_BYTES, "the network's outputs live in the idle A tile");
  static_assert(sizeof(typename G::State) + 8 + 1 + 2 * PATH_SMEM_DEPTH <= 68, "rollout hand-off budget per game");
  static_assert(FCfg<2>::SMEM <= 195 * 1024, "shared memory beyond the 196 KB carve-out costs 32 KB of L1");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + 3), bar_done = smem_u32(bars + 6), bar_stagger = smem_u32(bars + 8);
#if AG_NHALF
  constexpr bool NHALF = CPW == 2;                                     // only where a warp owns a slice in each column half
  const uint32_t bar_done2 = smem_u32(bars + 74);                      // [2] all MMAs of the layer (the 48 bytes behind sbias are free)
#endif

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, ntiles); }
    for (int t = 0; t < NT; t++) mbar_init(bar_done + 8 * t, 1);
#if AG_NHALF
    if (NHALF) for (int t = 0; t < NT; t++) mbar_init(bar_done2 + 8 * t, 1);
#endif
    mbar_init(bar_stagger, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < TC_N) sbias[threadIdx.x] = T.bias[threadIdx.x];
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), C::TMEM_COLS);        // NT x 128 accumulator columns + NT x 128 residual columns
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int nlayers = T.nlayers;
  const int total_layers = visits * nlayers;
  // weight image of global layer index wl (= rollout * nlayers + layer) -> ring stage wl % 3
  auto load_layer = [&](int wl) {
    const int s = wl % STAGES, l = wl % nlayers;
    const uint32_t bytes = (l == nlayers - 1) ? (uint32_t)(T.NH * TC_N * 2) : (uint32_t)TC_W_STAGE_BYTES;
    if (NT == 2 && wl >= STAGES) mbar_wait(bar_empty + 8 * s, ((wl / STAGES) - 1) & 1);
    mbar_expect_tx(bar_full + 8 * s, bytes);
    bulk_g2s(smem_u32(sW + s * TC_W_STAGE_BYTES), T.img + (size_t)l * TC_W_STAGE_BYTES, bytes, bar_full + 8 * s);
  };
  if (threadIdx.x == 0) {                                              // fill the ring: STAGES - 1 layers ahead
    load_layer(0);
    if (STAGES > 2 && total_layers > 1) load_layer(1);
  }

  // ---- roles ----
  // search: group of W lanes per game, pass p covers local games p*128 .. p*128+127
  const int sg = threadIdx.x / W, sl = threadIdx.x & (W - 1);
  const unsigned gm = group_mask<W>();
  constexpr int GROUPS = C::THREADS / W;                               // games per pass
  constexpr int PASSES = C::GAMES / GROUPS;
  // network: tile, TMEM lane quarter, 32-column slice
  const int t = warp / WPT, wq = warp & 3, csb = ((warp >> 2) & (WPT / 4 - 1)) * CPW;   // first column slice of this warp
  const int r = wq * 32 + lane;
  const int g_row = S.off + cta_first + t * TC_TILE_M + r;            // the game whose activations this thread carries
  unsigned char* At = sA + t * TC_A_BYTES;
  const uint32_t tmem_acc = tmem_base + (uint32_t)(t * TC_N);
  const uint32_t tmem_res = tmem_base + (uint32_t)(NT * TC_N + t * TC_N);
  const uint32_t lane_row = (uint32_t)(wq * 32) << 16;
  // The MMA-issuing warp of a tile takes a WARP-UNIFORM branch and elects one lane inside it; every operand of tcgen05.mma is derived
  // from values the compiler can see as uniform (the broadcast warp index, the broadcast TMEM base).  Issued from a divergent
  // `lane == 0` branch each MMA went through an ELECT / 5 x R2UR / BRA.U.ANY waterfall: ~75 cycles per instruction, 600 per layer.
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const int t_u = warp_u / WPT;
  const bool issuer_warp = (warp_u % WPT) == 0;
  const uint32_t tmem_acc_u = __shfl_sync(0xffffffffu, tmem_base, 0) + (uint32_t)(t_u * TC_N);
  const uint32_t a_smem_u = smem_u32(sA) + (uint32_t)(t_u * TC_A_BYTES);
  const uint32_t one = (FMT == 0) ? 0x3F80u : 0x3C00u;

  // one thread per game for the descent and the expansion: its uid and node count stay in registers for the whole ply
  const bool has_game = (int)threadIdx.x < count;
  const int my_g = S.off + cta_first + (int)threadIdx.x;
  const u32 my_uid = has_game ? P.uid[my_g] : 0u;
  int my_nn = has_game ? P.nnodes[my_g] : 0;
  typename G::State my_root = G::init();                               // the root's state: constant for the whole ply
  if (has_game) my_root = *reinterpret_cast<const typename G::State*>(P.tree + (size_t)my_g * P.game_stride + Lay::OFF_STATE);
  // Philox block (depths 0..3) of the NEXT descent, computed while this thread would otherwise wait for the first MMA of a network phase
  Philox4 rnd_next = philox4x32_10(my_uid, S.ply, 0u, 0u, (u32)S.seed, (u32)(S.seed >> 32));
  if (threadIdx.x < C::GAMES) s_d[threadIdx.x] = 0;
#if AG_TREE_SMEM
  // fill the node cache with the nodes that exist when the ply starts (the root alone after root_reset); read back only by this thread
  if (has_game) {
    const char* gb = P.tree + (size_t)my_g * P.game_stride;
    for (int nd = 0; nd < min(my_nn, SH.nc_nodes); nd++) {
      unsigned char* sl = node_cache_slot<G>(SH, (int)threadIdx.x, nd);
      const char* rec = gb + (size_t)nd * Lay::REC;
      *reinterpret_cast<uint2*>(sl) = *reinterpret_cast<const uint2*>(rec + Lay::OFF_HDR);
      for (int c = 0; c < Lay::APAD / 8; c++) *reinterpret_cast<uint2*>(sl + RootSlot<Lay::APAD>::OFF_CHILD + 8 * c) = *reinterpret_cast<const uint2*>(rec + Lay::OFF_CHILD + 8 * c);
      for (int c = 0; c < Lay::APAD / 4; c++) *reinterpret_cast<float4*>(sl + RootSlot<Lay::APAD>::OFF_POLICY + 16 * c) = *reinterpret_cast<const float4*>(rec + Lay::OFF_POLICY + 16 * c);
    }
  }
#endif

  int wl = 0;                                                          // global layer counter (ring / barrier phases)
  long long t_ly[5] = {0, 0, 0, 0, 0};                                  // development trace of the issuer: weights wait, MMA issue, MMA done, epilogue, barrier
  long long t_ph[5] = {0, 0, 0, 0, 0}, t_mark = T.dbg ? clock64() : 0;   // development trace (agpu_debug_tc_trace): expand, scan, backup, select, network
  for (int k = 0; k < visits; k++) {
    const int last = (k == visits - 1);
    // ================= search phase =================
    if (k > 0) {
      // (a) expand every game of the CTA (softmax, legal mask, prior), one thread per game, leaving the leaf evaluation in shared memory;
      // (b) meanwhile the last warp — never a game thread at these sizes — prefix-sums the path lengths of the previous descent and
      //     fills the item -> game map of the backup phase
      if (has_game) s_eval[threadIdx.x] = expand_game1<G>(P, my_g, (int)threadIdx.x, SH, S.training, 0);
      if (warp == C::THREADS / 32 - 1) {
        constexpr int PER = C::GAMES / 32;
        int loc[PER], dd[PER], sum = 0;
#pragma unroll
        for (int i = 0; i < PER; i++) { dd[i] = s_d[lane * PER + i]; loc[i] = sum; sum += dd[i]; }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        const int excl = incl - sum;
#pragma unroll
        for (int i = 0; i < PER; i++) {
          const int off = excl + loc[i];
          s_off[lane * PER + i] = off;
          for (int j = 0; j < dd[i]; j++) if (off + j < ITEM_MAP) s_item[off + j] = (uint8_t)(lane * PER + i);
        }
        if (lane == 31) s_off[C::GAMES] = incl;
      }
      __syncthreads();
      if (T.dbg && threadIdx.x == 0) { const long long c = clock64(); t_ph[0] += c - t_mark; t_mark = c; }
      // (c) backUp + re-solve of π̄: one (game, ancestor) item per THREAD, packed densely over the CTA — with a lane group per
      //     game only d of its 8 lanes (46 % on average) had an ancestor to work on, and the solve is 40 % of the search time
      const int items = s_off[C::GAMES];
      for (int i = threadIdx.x; i < items; i += C::THREADS) {
        int lo = 0;
        if (i < ITEM_MAP) {
          lo = s_item[i];
        } else {                                                       // largest gl with s_off[gl] <= i
          int hi = C::GAMES;
          while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_off[mid] <= i) lo = mid; else hi = mid; }
        }
        backup_item<G, SW>(P, S.off + cta_first + lo, i - s_off[lo], s_d[lo], s_eval[lo], 0, S.cpuct, (T.dbg && threadIdx.x == 0) ? T.dbg + blockIdx.x * 32 + 8 : nullptr,
                       SH.pn + lo * PATH_SMEM_DEPTH, SH.pm + lo * PATH_SMEM_DEPTH,
                       SH.nc_nodes > 0 ? SH.nc_base + (size_t)lo * SH.nc_nodes * RootSlot<Lay::APAD>::BYTES : nullptr, SH.nc_nodes);
      }
      __syncthreads();
      if (T.dbg && threadIdx.x == 0) { const long long c = clock64(); t_ph[2] += c - t_mark; t_mark = c; }
    }
    // (d) descent of this rollout
    if (has_game) select_game1<G>(P, my_g, (int)threadIdx.x, SH, my_uid, my_nn, k, last, S.seed, S.ply, my_root, rnd_next, (T.dbg && threadIdx.x == 0) ? T.dbg + blockIdx.x * 32 + 8 : nullptr);
    __syncthreads();                                                   // leaves (global) visible to the encoders of this CTA
    if (T.dbg && threadIdx.x == 0) { const long long c = clock64(); t_ph[3] += c - t_mark; t_mark = c; }

    // ================= network phase =================
    {
      // A operand of the base layer: this thread's 32 operand columns of its row (decoder, mcts_gpu.jl:202-223)
      u64 x0 = 0, x1 = 0;
      if (g_row < L_end) {
        const u64* st = reinterpret_cast<const u64*>(SH.state + t * TC_TILE_M + r);     // left there by this rollout's descent
        const u64 bp = st[0], bo = st[1];
        constexpr int VS = G::VS;
        x0 = (VS < 64) ? (bp | (bo << VS)) : bp;
        x1 = (VS < 64) ? (bo >> (64 - VS)) : bo;
      }
#pragma unroll
      for (int j = 0; j < CPW; j++) {
        const int cs = csb + j;
        const uint32_t bits = (uint32_t)(((cs & 2) ? x1 : x0) >> (32 * (cs & 1)));
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const uint32_t byte = (bits >> (8 * i)) & 0xFFu;
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; e++) w[e] = ((byte >> (2 * e)) & 1u) * one | (((byte >> (2 * e + 1)) & 1u) * one) << 16;
          const int c = 4 * cs + i;                                    // chunk of 8 operands in the row, 0..15
          *reinterpret_cast<uint4*>(At + (c >> 3) * TC_KTILE_BYTES_A + r * 128 + (((c & 7) ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
    if (t >= ntiles) { wl += nlayers; __syncthreads(); continue; }      // idle tile: rejoin at the end-of-rollout barrier
    fence_proxy_async();
    named_bar_sync(1 + t, 32 * WPT);

    uint32_t sres[16];                                                 // this thread's residual values (swapped orientation)
    for (int l = 0; l < nlayers; l++, wl++) {
      const int s = wl % STAGES;
      const bool is_head = (l == nlayers - 1);
      const int nl = is_head ? T.NH : TC_N;
      long long lt0 = 0, lt1 = 0, lt2 = 0;
      const bool ltr = T.dbg != nullptr && threadIdx.x == 0 && !is_head;
      if (ltr) lt0 = clock64();
      if (issuer_warp) {
        mbar_wait(bar_full + 8 * s, (wl / STAGES) & 1);
        if (ltr) lt1 = clock64();
        if (wl == 0 && t_u == 1) mbar_wait(bar_stagger, 0);            // tile 1 trails tile 0 by one MMA phase
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad0 = umma_desc(a_smem_u);
          const uint64_t bd0 = umma_desc(smem_u32(sW) + (uint32_t)(s * TC_W_STAGE_BYTES));
          const uint64_t bstep = (uint64_t)((nl * 128) >> 4);
          if (SW && swapped && !is_head) {
            const uint32_t idesc = umma_idesc<FMT>(NS);                 // M = 128 out-features, N = NS games
#pragma unroll
            for (int ks = 0; ks < TC_N / 16; ks++) {
              const uint64_t ainc = (uint64_t)(((ks >> 2) * TC_KTILE_BYTES_A + (ks & 3) * 32) >> 4);
              const uint64_t binc = (uint64_t)(((ks & 3) * 32) >> 4) + ((ks >> 2) ? bstep : 0);
              umma_bf16(tmem_acc_u + (uint32_t)((ks % AG_KSPLIT) * NS), bd0 + binc, ad0 + ainc, idesc, ks >= AG_KSPLIT ? 1u : 0u);   // weights as A, activations as B
            }
          }
#if AG_NHALF
          else if (NHALF && !is_head) {
            const uint32_t idesc = umma_idesc<FMT>(TC_N / 2);           // 64 output columns per chain
#pragma unroll
            for (int ks = 0; ks < TC_N / 16; ks++) {
              const uint64_t ainc = (uint64_t)(((ks >> 2) * TC_KTILE_BYTES_A + (ks & 3) * 32) >> 4);
              const uint64_t binc = (uint64_t)(((ks & 3) * 32) >> 4) + ((ks >> 2) ? bstep : 0);
              umma_bf16(tmem_acc_u, ad0 + ainc, bd0 + binc, idesc, ks > 0 ? 1u : 0u);
            }
            umma_commit(bar_done + 8 * t_u);                             // columns 0..63 are complete
            const uint64_t brow = (uint64_t)(((TC_N / 2) * 128) >> 4);   // weight rows 64..127: 8 swizzle atoms further in each K tile
#pragma unroll
            for (int ks = 0; ks < TC_N / 16; ks++) {
              const uint64_t ainc = (uint64_t)(((ks >> 2) * TC_KTILE_BYTES_A + (ks & 3) * 32) >> 4);
              const uint64_t binc = (uint64_t)(((ks & 3) * 32) >> 4) + ((ks >> 2) ? bstep : 0);
              umma_bf16(tmem_acc_u + (uint32_t)(TC_N / 2), ad0 + ainc, bd0 + brow + binc, idesc, ks > 0 ? 1u : 0u);
            }
            umma_commit(bar_done2 + 8 * t_u);                            // all of the layer's MMAs
          }
#endif
          else {
            const uint32_t idesc = umma_idesc<FMT>(nl);
#pragma unroll
            for (int ks = 0; ks < TC_N / 16; ks++) {
              const uint64_t ainc = (uint64_t)(((ks >> 2) * TC_KTILE_BYTES_A + (ks & 3) * 32) >> 4);
              const uint64_t binc = (uint64_t)(((ks & 3) * 32) >> 4) + ((ks >> 2) ? bstep : 0);
              umma_bf16(tmem_acc_u, ad0 + ainc, bd0 + binc, idesc, ks > 0 ? 1u : 0u);
            }
          }
#if AG_NHALF
          // unsplit layers (head, swapped) complete both barriers at once, so that both advance one phase per layer
          if (!(NHALF && !is_head && !(SW && swapped))) { umma_commit(bar_done + 8 * t_u); if (NHALF) umma_commit(bar_done2 + 8 * t_u); }
#else
          umma_commit(bar_done + 8 * t_u);
#endif
          if (NT == 2) umma_commit(bar_empty + 8 * s);                  // one tile: the requesting lane has itself seen the previous layer complete
          if (wl == 0 && t_u == 0) umma_commit(bar_stagger);
        }
        __syncwarp();
        if (ltr) lt2 = clock64();
      }
      // the weights two layers ahead are requested by a lane that would otherwise just wait for this layer's MMAs (on the issuer
      // the request sat on the critical path: 2 k cycles per rollout)
      if (warp == 1 && lane == 0 && wl + STAGES - 1 < total_layers) load_layer(wl + STAGES - 1);
      if (l == 0 && has_game) rnd_next = philox4x32_10(my_uid, S.ply, (u32)(k + 1), 0u, (u32)S.seed, (u32)(S.seed >> 32));
      mbar_wait(bar_done + 8 * t, wl & 1);
#if AG_NHALF
      const bool split = NHALF && !is_head && !(SW && swapped);         // CTA-uniform: this layer was issued as two column halves
      if (NHALF && !split) mbar_wait(bar_done2 + 8 * t, wl & 1);        // keeps the second barrier's phase in step on unsplit layers
#endif
      tc_fence_after();
      long long lt3 = 0;
      if (ltr) { lt3 = clock64(); t_ly[0] += lt1 - lt0; t_ly[1] += lt2 - lt1; t_ly[2] += lt3 - lt2; }

      if (SW && swapped && !is_head) {
        if constexpr (SW) {
          if (NS == 32) epilogue_swapped<FMT, 8, AG_KSPLIT>(tmem_acc, NS, wq, csb, lane, l, At, sres);
          else epilogue_swapped<FMT, 16, AG_KSPLIT>(tmem_acc, NS, wq, csb, lane, l, At, sres);
        }
        tc_fence_before();
        fence_proxy_async();
        long long lt4 = 0;
        if (ltr) { lt4 = clock64(); t_ly[3] += lt4 - lt3; }
        named_bar_sync(1 + t, 32 * WPT);
        if (ltr) t_ly[4] += clock64() - lt4;
      } else if (!is_head) {
        // epilogue: b = relu(acc) (base) or b + relu(acc); fp32 residual in TMEM; next A operand = fp16/bf16(b)
        const bool keep = (l + 2 < nlayers);                            // the last trunk layer's residual is not read again
#if AG_NHALF
        uint4 pend[4];                                                  // operands of a first-half slice, held back
        int pend_cs = -1;
        bool waited2 = !split;
        auto flush_pending = [&] {
          if (!waited2) { mbar_wait(bar_done2 + 8 * t, wl & 1); tc_fence_after(); waited2 = true; }
          if (pend_cs >= 0) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
              const int c = 4 * pend_cs + q;
              *reinterpret_cast<uint4*>(At + (c >> 3) * TC_KTILE_BYTES_A + r * 128 + (((c & 7) ^ (r & 7)) << 4)) = pend[q];
            }
            pend_cs = -1;
          }
        };
#pragma unroll
        for (int ji = 0; ji < 2 * CPW; ji++) {
          // slices of this warp: one in each column half when the layer is split, else csb, csb + 1
          const int cs = NHALF ? (csb >> 1) + 2 * (ji >> 1) : csb + (ji >> 1);
          const int i = ji & 1;
          const bool hold = split && cs < 2;                            // a slice of the first half: its stores wait
          if (split && !hold) flush_pending();                          // first touch of the second half: wait for its MMAs
          const uint32_t lane_sel = lane_row + (uint32_t)(cs * 32);
          uint32_t va[16], vh[16];
          tmem_ld16(tmem_acc + lane_sel + 16 * i, va);
          if (l > 0) tmem_ld16(tmem_res + lane_sel + 16 * i, vh);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; e++) {
            const float ra = fmaxf(__uint_as_float(va[e]), 0.f);
            const float hv = (l == 0) ? ra : __uint_as_float(vh[e]) + ra;
            vh[e] = __float_as_uint(hv);
          }
          if (keep) tmem_st16(tmem_res + lane_sel + 16 * i, vh);
#pragma unroll
          for (int c2 = 0; c2 < 2; c2++) {
            const int c = 4 * cs + 2 * i + c2;
            const uint4 pk = make_uint4(pack2<FMT>(__uint_as_float(vh[8 * c2 + 0]), __uint_as_float(vh[8 * c2 + 1])),
                                        pack2<FMT>(__uint_as_float(vh[8 * c2 + 2]), __uint_as_float(vh[8 * c2 + 3])),
                                        pack2<FMT>(__uint_as_float(vh[8 * c2 + 4]), __uint_as_float(vh[8 * c2 + 5])),
                                        pack2<FMT>(__uint_as_float(vh[8 * c2 + 6]), __uint_as_float(vh[8 * c2 + 7])));
            if (hold) { pend[2 * i + c2] = pk; pend_cs = cs; }
            else *reinterpret_cast<uint4*>(At + (c >> 3) * TC_KTILE_BYTES_A + r * 128 + (((c & 7) ^ (r & 7)) << 4)) = pk;
          }
        }
        flush_pending();
#else
#pragma unroll
        for (int ji = 0; ji < 2 * CPW; ji++) {
          const int cs = csb + (ji >> 1), i = ji & 1;
          const uint32_t lane_sel = lane_row + (uint32_t)(cs * 32);
          uint32_t va[16], vh[16];
          tmem_ld16(tmem_acc + lane_sel + 16 * i, va);
          if (l > 0) tmem_ld16(tmem_res + lane_sel + 16 * i, vh);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; e++) {
            const float ra = fmaxf(__uint_as_float(va[e]), 0.f);
            const float hv = (l == 0) ? ra : __uint_as_float(vh[e]) + ra;
            vh[e] = __float_as_uint(hv);
          }
          if (keep) tmem_st16(tmem_res + lane_sel + 16 * i, vh);
#pragma unroll
          for (int c2 = 0; c2 < 2; c2++) {
            const int c = 4 * cs + 2 * i + c2;
            const uint4 pk = make_uint4(pack2<FMT>(__uint_as_float(vh[8 * c2 + 0]), __uint_as_float(vh[8 * c2 + 1])),
                                        pack2<FMT>(__uint_as_float(vh[8 * c2 + 2]), __uint_as_float(vh[8 * c2 + 3])),
                                        pack2<FMT>(__uint_as_float(vh[8 * c2 + 4]), __uint_as_float(vh[8 * c2 + 5])),
                                        pack2<FMT>(__uint_as_float(vh[8 * c2 + 6]), __uint_as_float(vh[8 * c2 + 7])));
            *reinterpret_cast<uint4*>(At + (c >> 3) * TC_KTILE_BYTES_A + r * 128 + (((c & 7) ^ (r & 7)) << 4)) = pk;
          }
        }
#endif
        tmem_st_wait();
        tc_fence_before();
        fence_proxy_async();
        long long lt4 = 0;
        if (ltr) { lt4 = clock64(); t_ly[3] += lt4 - lt3; }
        named_bar_sync(1 + t, 32 * WPT);
        if (ltr) t_ly[4] += clock64() - lt4;
      } else {
        // heads: logits = acc + bias, value = σ(acc[A] + bias[A])   (DenseNet.jl:301) -> nn_out (global, read by the next search phase)
        float* o = P.nn_out + (size_t)g_row * Lay::OUTS;
#pragma unroll
        for (int i = 0; i < 2; i++) {
          const int cs = csb;                                           // NH <= 32: only a warp's first slice can hold head columns
          const uint32_t lane_sel = lane_row + (uint32_t)(cs * 32);
          const int a0 = cs * 32 + i * 16;
          if (a0 < T.NH) {                                              // warp-uniform
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
            if (g_row < L_end) {
              float* so = SH.out + t * SH.out_tile_stride + r * Lay::OUTS;
#pragma unroll
              for (int q4 = 0; q4 < 4; q4++)
                if (a0 + 4 * q4 < Lay::OUTS) {
                  const float4 zv = make_float4(z[4 * q4], z[4 * q4 + 1], z[4 * q4 + 2], z[4 * q4 + 3]);
                  *reinterpret_cast<float4*>(so + a0 + 4 * q4) = zv;
                  if (last) *reinterpret_cast<float4*>(o + a0 + 4 * q4) = zv;     // the expand of the last rollout reads it from global memory
                }
            }
          }
        }
        tc_fence_before();
      }
    }
    __syncthreads();                                                   // nn_out (global) visible to the search phase
    if (T.dbg && threadIdx.x == 0) { const long long c = clock64(); t_ph[4] += c - t_mark; t_mark = c; }
  }
  if (T.dbg && threadIdx.x == 0) {
    for (int i = 0; i < 5; i++) T.dbg[blockIdx.x * 32 + i] = t_ph[i];
    T.dbg[blockIdx.x * 32 + 5] = count; T.dbg[blockIdx.x * 32 + 6] = visits;
    for (int i = 0; i < 5; i++) T.dbg[blockIdx.x * 32 + 24 + i] = t_ly[i];
  }

  // expand + backUp of the last rollout (publishes nothing new for the root: policy_final was written by its descent)
#pragma unroll 1
  for (int p = 0; p < PASSES; p++) {
    const int g = S.off + cta_first + p * GROUPS + sg;
    if (g < L_end) expand_backup_game<G, false>(P, g, sl, gm, S.training, 1, nullptr, nullptr, S.cpuct);
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

}  // namespace fused
}  // namespace ag
