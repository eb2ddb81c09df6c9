// search.cuh — tree layout in HBM and the search kernels (select / expand+backup / ply bookkeeping).
//
// Replaces kdescendTree!, expand, backUp, copy_pol, reinit and the per-ply fills of mcts_gpu.jl
// (:100-199, :250-339, :359-387).  Results are bit-identical to a sequential IEEE-fp32 reading of those
// functions (the CPU oracle); what changes is where the work runs:
//
//  * Layout.  One contiguous block of R node records per game; a record holds, struct-of-arrays over the
//    (padded) actions, prior f32 | q f32 | visits u16 | child u8 | order u8, then the packed State and an
//    8-byte header.  Connect4: exactly one 128-byte line per node.  The reference's Achild/childID double
//    indirection (R*R int64 per game) becomes child[a] (node id) + order[slot] (action, creation order);
//    `policy`, `uptodate`, `expanded` arrays disappear (π̄ is recomputed from q/visits when a node has
//    visits — the reference's sticky-dirty rule, mcts_gpu.jl:114,321 — and flags live in the header).
//  * Mapping.  W = min(32, pow2ceil(A)) lanes cooperate on one game (Connect4: 8 lanes, 4 games per warp):
//    lane l owns actions l, l+W, …  Loads of a node's statistics are one coalesced segment per field.
//    Divisions/sqrt of the α-solve run lane-parallel; every order-sensitive fp32 sum is then accumulated
//    by broadcast in the reference's order (ascending action; Newton terms in child-creation order), so
//    parallelism never changes a rounding.
//  * No per-ply memset: a node's q/visits/child are zeroed when the node is allocated.
#pragma once
#include "games.cuh"

namespace ag {

// F_PRIOR_BOX (FAST layouts, set by expand): every prior of the node is +0 or >= 2^-50, which — with λ in [2^-10, 2^10] — puts every
// numerator of the α-solve (λ·prior, λ·Σ prior) inside the operand box of fdiv_fast without looking at them again
enum : uint8_t { F_EXPANDED = 1, F_TERMINAL = 2, F_PRIOR_BOX = 4 };
constexpr float PRIOR_BOX_LO = 8.881784197001252e-16f;   // 2^-50

struct alignas(8) NodeHdr {
  uint8_t parent;   // 1-based node id, 0 = root has none            (vnodes.parent)
  uint8_t action;   // 1-based action from parent                    (vnodes.actionFromParent)
  uint8_t nchild;   //                                               (vnodesStats.childnbr)
  uint8_t flags;    // F_EXPANDED (vnodes.expanded) | F_TERMINAL (isOver cached at creation)
  int8_t result;    // winner colour if terminal
  uint8_t pad[3];
};
static_assert(sizeof(NodeHdr) == 8, "header");
// a header as ONE 8-byte word: field-by-field assignment through a pointer compiles to eight byte stores, i.e. eight requests to the
// memory pipeline per new node (ncu, round 1: 8 of the 129 L1 requests of a simulation)
AG_HD u64 hdr_word(int parent, int action, int nchild, int flags, int result) {
  return (u64)(uint8_t)parent | ((u64)(uint8_t)action << 8) | ((u64)(uint8_t)nchild << 16) | ((u64)(uint8_t)flags << 24) | ((u64)(uint8_t)(int8_t)result << 32);
}
AG_D void hdr_store(void* p, u64 w) { *reinterpret_cast<u64*>(p) = w; }
struct alignas(8) NodeAux {
  float rem;         // Σ_{a: no child} prior[a], ascending action order
  uint16_t acount;   // #{a: prior[a] > 0}
  uint16_t nvis;     // Σ_a visits[a]
};
static_assert(sizeof(NodeAux) == 8, "aux");
AG_D NodeHdr hdr_from_word(u64 w) { NodeHdr h; *reinterpret_cast<u64*>(&h) = w; return h; }

constexpr int pow2ceil(int x) { int p = 1; while (p < x) p <<= 1; return p; }

// Lanes cooperating on one game: W = pow2ceil(A) capped at a warp.  Measured on B200 (profiles/r01_lanes_sweep.txt,
// Connect4): 8 lanes 64.6 us per select launch, 4 lanes 72.7, 2 lanes 96.5, 1 lane 138.8 — the descent is a serial
// dependency chain per game, so spreading a node's actions over more lanes shortens it.  AG_LANES_SHIFT halves W per step.
#ifndef AG_LANES_SHIFT
#define AG_LANES_SHIFT 0
#endif
// lanes per game where A > 32 (the descent there is bound by issue slots: the ordered sums are executed once per group, not per lane)
#ifndef AG_LANES_BIG
#define AG_LANES_BIG 32
#endif
// resident 256-thread blocks per SM the search kernels are compiled for (register cap = 65536 / (256 * AG_MINBLOCKS))
#ifndef AG_MINBLOCKS
#define AG_MINBLOCKS 4
#endif
// threads per block of the search kernels for large action sets (a block's slot is held until its slowest game ends its descent)
#ifndef AG_BLOCK_BIG
#define AG_BLOCK_BIG 128
#endif
template <class G>
struct Layout {
  static constexpr int A = G::A;
  static constexpr int W0 = pow2ceil(A) < 32 ? pow2ceil(A) : (A > 32 ? AG_LANES_BIG : 32);
  static constexpr int W = (A <= 32 && (W0 >> AG_LANES_SHIFT) >= 1) ? (W0 >> AG_LANES_SHIFT) : W0;   // lanes per game
  static constexpr int APL = (A + W - 1) / W;                      // actions per lane
  static constexpr int APAD = (W * APL + 7) / 8 * 8;
  static constexpr int GPW = 32 / W;                               // games per warp
  // Small action sets (Connect4, tic-tac-toe): π̄ is stored per node and re-solved in the BACKUP phase, one lane per ancestor,
  // all ancestors of a path in parallel; the descent then only samples.  Equivalent to the reference's solve-at-descent because
  // a node's statistics change only when a backup passes through it (sticky `uptodate`, mcts_gpu.jl:114,321) and the solve is a
  // pure function of them.  Large action sets keep the cooperative solve-at-descent (the network dominates there).
  static constexpr bool FAST = (A <= 9);
  // FAST records with at most 7 actions keep the creation order of the children in the header's spare bytes — slot k's 1-based action in
  // bits 40 + 3k of the header word — so that header, order and child ids are one 16-byte access for the backup and one 16-byte store
  // when a child is created (the order byte array of the other layouts stays unused)
  static constexpr bool ORD_IN_HDR = FAST && A <= 7;
  // ... and the visit counts (one byte each: a game's tree has at most 255 nodes) in the eighth, unused slots of the q and prior vectors —
  // actions 0..3 behind q, 4..6 behind prior — so that the backup needs no separate load for them.  Both slots are zero when the node
  // is created (q) / expanded (prior), before any of its children can be visited.
  static constexpr bool VIS_IN_PAD = ORD_IN_HDR;
  static constexpr int vis_byte(int a) { return a < 4 ? OFF_Q_F + 28 + a : OFF_PRIOR_F + 24 + a; }   // (FAST offsets; VIS_IN_PAD only)
  static constexpr int SB = FAST ? 256 : AG_BLOCK_BIG;            // threads per block of select / expand+backup / step kernels
  static constexpr int SB_MIN = AG_MINBLOCKS * 256 / SB;          // resident blocks per SM they are compiled for
  static constexpr int a16(int x) { return (x + 15) / 16 * 16; }
  // FAST record: what the descent reads (header, child ids, π̄) sits in the first 64 bytes, so one 64-byte-aligned sector pair
  // serves a level of the descent; the backup's lane reads the rest.  Other layouts: prior | q | visits | child | order | state | header.
  // Records with the order in the header and the visits in the vectors need neither array, and keep the node's STATE outside the record
  // (SPLIT_STATE: the states of a game follow its R records) — nothing on the hot path reads a stored state: header + child ids |
  // π̄ | prior | q = 112 bytes, ONE 128-byte line per node (192 bytes before: five sectors per backup item instead of four, a third
  // more tree per game in L1 / L2).
  static constexpr bool SPLIT_STATE = ORD_IN_HDR;
  static constexpr int OFF_HDR_F = 0;
  static constexpr int OFF_CHILD_F = 8;
  static constexpr int OFF_POLICY_F = a16(8 + APAD);
  static constexpr int OFF_ORDER_F = OFF_POLICY_F + 4 * APAD;                                   // (unused with ORD_IN_HDR)
  static constexpr int OFF_VIS_F = a16(OFF_ORDER_F + APAD);                                     // (unused with VIS_IN_PAD)
  static constexpr int OFF_PRIOR_F = SPLIT_STATE ? OFF_POLICY_F + 4 * APAD : a16(OFF_VIS_F + 2 * APAD);
  static constexpr int OFF_Q_F = OFF_PRIOR_F + 4 * APAD;
  static constexpr int OFF_STATE_F = (OFF_Q_F + 4 * APAD + 7) / 8 * 8;                          // (end of the statistics; the state itself only if !SPLIT_STATE)

  static constexpr int OFF_PRIOR = FAST ? OFF_PRIOR_F : 0;
  static constexpr int OFF_Q = FAST ? OFF_Q_F : 4 * APAD;
  static constexpr int OFF_VIS = FAST ? OFF_VIS_F : 8 * APAD;
  static constexpr int OFF_CHILD = FAST ? OFF_CHILD_F : 10 * APAD;
  static constexpr int OFF_ORDER = FAST ? OFF_ORDER_F : 11 * APAD;
  static constexpr int OFF_POLICY = FAST ? OFF_POLICY_F : 0;           // (FAST only)
  static constexpr int OFF_STATE = FAST ? OFF_STATE_F : (12 * APAD + 7) / 8 * 8;
  static constexpr int OFF_HDR = FAST ? OFF_HDR_F : OFF_STATE + (int)sizeof(typename G::State);
  static constexpr int STATS_BEGIN = FAST ? 8 : 0;                     // [STATS_BEGIN, STATS_END): zeroed when a root is installed
  static constexpr int STATS_END = OFF_STATE;
  // Large action sets: 8 more bytes behind the header cache what the α-solve would otherwise recompute from A-long arrays at every
  // visited level — prior_rem = Σ prior over the actions without a child (ascending action order, mcts_gpu.jl:122-124; it changes only
  // when a child is created), #{prior > 0} (:128-130; fixed by expand) and Σ visits (:118-120; +1 per backup through the node).
  static constexpr int OFF_AUX = OFF_HDR + 8;
  static constexpr int REC = SPLIT_STATE ? (OFF_STATE + 63) / 64 * 64
                             : FAST ? (OFF_STATE + (int)sizeof(typename G::State) + 63) / 64 * 64
                                    : (OFF_AUX + 8 + 31) / 32 * 32;     // record size
  // bytes of one game's block for R nodes, and where the state of a node lives inside it
  static constexpr int STATE_STRIDE = SPLIT_STATE ? ((int)sizeof(typename G::State) + 15) / 16 * 16 : REC;   // (16-byte aligned: state_store)
  static constexpr size_t game_bytes(int R) { return (size_t)R * (REC + (SPLIT_STATE ? STATE_STRIDE : 0)); }
  static constexpr size_t state_off(int R, int node) {
    return SPLIT_STATE ? (size_t)R * REC + (size_t)node * STATE_STRIDE : (size_t)node * REC + OFF_STATE;
  }
  static constexpr int OUTS = (A + 1 + 3) / 4 * 4;                 // floats per game of network output: logits[A], value
  static_assert(sizeof(typename G::State) % 8 == 0, "state alignment");
};

struct SearchParams {
  char* tree;            // [L][R] node records
  size_t game_stride;    // R * REC
  int R;
  int32_t* nnodes;       // [L]  newindex
  int32_t* leaf;         // [L]  0-based leaf node
  uint32_t* uid;         // [L]  global game id (RNG key)
  float* policy_final;   // [L][A]
  float* nn_out;         // [L][OUTS] logits then value
  unsigned long long* counters;   // [2] nodes traversed, descents (profiling; may be null)
  // path of the last descent (FAST layouts): node ids (0-based) and actions (0-based) from the root down to the leaf's parent
  uint8_t* path_node;    // [L][R]
  uint8_t* path_move;    // [L][R]
  uint8_t* path_len;     // [L]
};

// Shared-memory hand-off between the phases of one rollout inside the fused per-ply kernel: what one phase leaves for the next
// (leaf node, its state and header, the path, the network's output) would otherwise make a round trip through L2 — about a
// thousand cycles each on the serial chain descent -> encode -> network -> expand -> backup -> descent.  Global memory keeps a copy
// of everything (the stores are off the critical path), so the stand-alone kernels and the read-back entry points see the same data.
constexpr int PATH_SMEM_DEPTH = 16;          // path entries per game held in shared memory; deeper levels are read from global

// loads of the records on a path (descent and backup)
AG_D uint2 hot_ld_u2(const void* a) { return *reinterpret_cast<const uint2*>(a); }
AG_D uint4 hot_ld_u4(const void* a) { return *reinterpret_cast<const uint4*>(a); }
AG_D float4 hot_ld_f4(const void* a) { return *reinterpret_cast<const float4*>(a); }
constexpr int ITEMS_PER_GAME = 8;            // capacity of the backup phase's item list, per game of the CTA (mean path length: 3-5)
template <class G>
struct RolloutShared {
  typename G::State* state;  // [GAMES] state of the leaf
  typename G::State* root;   // [GAMES] state of the root: constant for the whole ply
  NodeHdr* hdr;              // [GAMES] header of the leaf as the descent saw it
  Philox4* rnd;              // [GAMES] Philox block (depths 0..3) of the NEXT descent, computed under a network phase
  float* out;                // logits, value of the leaf: row gl at out + (gl / 128) * out_tile_stride + (gl % 128) * OUTS
  int out_tile_stride;       // (floats) the rows of a 128-game tile live in that tile's idle A-operand buffer
  int* d;                    // [GAMES] path length
  uint8_t* leaf;             // [GAMES]
  uint16_t* path;            // [GAMES][PATH_SMEM_DEPTH] path entries: node | move << 8 (one request per level for the descent and per item
                             // for the backup — the search phases are bound by the number of load/store requests)
  int* lv_cnt;               // items of the last descent: entries reserved in lv_item (may exceed lv_cap)
  uint16_t* lv_item;         // [lv_cap] item = local game | path index << 8
  int lv_cap;                // capacity of the item list
  uint8_t* ovf;              // [GAMES] number of leading path indices of the game that ARE in the list (== path length unless the list was full)
  unsigned char* nc_base;    // node cache of the small-batch kernel: write-through copy of the descent fields (header, child ids, π̄) of the
  int nc_nodes;              //   first nc_nodes nodes of every game of the CTA; entry (gl, node) at nc_base + (gl * nc_nodes + node) * BYTES
};

// Node cache (small-batch per-ply kernel only).  In the tail of a generation a CTA holds <= 32 games and a rollout is a latency chain; the
// descent fields of the first nodes of each game are kept in shared memory so that a level of the descent is a shared-memory read
// instead of an L2 round trip (and the root, whose π̄ the backup has just rewritten, does not wait for that store to reach L2 and come
// back).  Every writer of those fields (node allocation in the descent, expand, the π̄ re-solve of the backup) updates global memory as
// before AND the cache entry; global memory stays authoritative for everything else.
template <int AP> struct CacheSlot {
  static constexpr int OFF_CHILD = 8;                                  // after the 8-byte header
  static constexpr int OFF_POLICY = (8 + AP + 15) & ~15;
  static constexpr int BYTES = OFF_POLICY + 4 * AP;                    // 48 (AP = 8), 96 (AP = 16)
};
// cache entry of (local game gl, node) or nullptr
template <class G, bool CACHE>
AG_D unsigned char* node_cache_slot(const RolloutShared<G>& SH, const int gl, const int node) {
  if (!CACHE) return nullptr;
  return node < SH.nc_nodes ? SH.nc_base + (size_t)(gl * SH.nc_nodes + node) * CacheSlot<Layout<G>::APAD>::BYTES : nullptr;
}

// One independent slice of the live games.  The rollout loop of a slice is replayed from a CUDA graph on its own stream, so
// everything that changes from ply to ply is read from this device-resident record instead of being a kernel argument.
struct SegParams {
  int off, len;          // slots [off, off+len)
  u32 ply;
  int training;
  u64 seed;
  float cpuct;
  int pad;
};

template <int W> AG_D unsigned group_mask() {
  return W == 32 ? 0xffffffffu : (((1u << W) - 1u) << ((threadIdx.x & 31) & ~(W - 1)));
}
template <int W> AG_D float gshfl(unsigned m, float v, int src) { return __shfl_sync(m, v, src, W); }
template <int W> AG_D int gshfl(unsigned m, int v, int src) { return __shfl_sync(m, v, src, W); }
template <int W> AG_D float gmax(unsigned m, float v) {
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(m, v, o, W));
  return v;
}
template <int W> AG_D int gsum(unsigned m, int v) {
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(m, v, o, W);
  return v;
}
template <int W> AG_D int gcount(unsigned m, bool pred) { return __popc(__ballot_sync(m, pred) & m); }

// value of slot j (uniform across the group) out of a small per-lane array, without dynamic register indexing
template <int APL, class T> AG_D T pick(const T (&v)[APL], int j) {
  T r = v[0];
#pragma unroll
  for (int k = 1; k < APL; k++) if (j == k) r = v[k];
  return r;
}

// a State as 8-byte words (one or a few wide stores instead of a store per field)
template <class S> AG_D void state_store(void* dst, const S& st) {
  static_assert(sizeof(S) % 8 == 0, "state size");
  union { S s; u64 w[sizeof(S) / 8]; } u;
  u.w[sizeof(S) / 8 - 1] = 0;
  u.s = st;
  if constexpr (sizeof(S) == 24) {
    *reinterpret_cast<uint4*>(dst) = make_uint4((u32)u.w[0], (u32)(u.w[0] >> 32), (u32)u.w[1], (u32)(u.w[1] >> 32));
    *reinterpret_cast<u64*>(reinterpret_cast<char*>(dst) + 16) = u.w[2];
  } else {
#pragma unroll
    for (int i = 0; i < (int)(sizeof(S) / 8); i++) reinterpret_cast<u64*>(dst)[i] = u.w[i];
  }
}

// ------------------------------------------------------------------------------------------------
// root install: re_init (mcts_gpu.jl:359-373) + the fills of mcts_single (:380-387) for node 1 only.
// src == null keeps the stored root state (agpu_search_begin).
// ------------------------------------------------------------------------------------------------
template <class G>
__global__ void __launch_bounds__(256) root_reset_kernel(SearchParams P, int L, const typename G::State* __restrict__ src,
                                                         const uint32_t* __restrict__ uid) {
  typedef Layout<G> Lay;
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= L) return;
  char* rec = P.tree + (size_t)g * P.game_stride;
  if (src) *reinterpret_cast<typename G::State*>(rec + Lay::state_off(P.R, 0)) = src[g];
  if (uid) P.uid[g] = uid[g];
  // zero the statistics (prior, q, visits, child, order, π̄)
  for (int o = Lay::STATS_BEGIN; o < Lay::STATS_END; o += 8) *reinterpret_cast<uint2*>(rec + o) = make_uint2(0, 0);
  hdr_store(rec + Lay::OFF_HDR, 0ull);
  if (!Lay::FAST) hdr_store(rec + Lay::OFF_AUX, 0ull);
  P.nnodes[g] = 1;
  P.leaf[g] = 0;
}

// ------------------------------------------------------------------------------------------------
// The α-solve of kdescendTree! (mcts_gpu.jl:116-169) for ONE node by ONE lane, everything in registers.  Same operations in the
// same order as the reference loop: sums ascending in the action, Newton terms in child-creation order starting from prior_rem/α.
// vis = visits after the update; ord[k] = 1-based action of the k-th created child.
// ------------------------------------------------------------------------------------------------
// v[idx] for a small register array without dynamic indexing (which would put the array in local memory): a binary tree of selects
template <int AP> AG_D float sel_reg(const float (&v)[AP], const int idx) {
  static_assert(AP == 8 || AP == 16, "select tree over 8 or 16 registers");
  float t[AP / 2];
#pragma unroll
  for (int i = 0; i < AP / 2; i++) t[i] = (idx & 1) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
  for (int i = 0; i < AP / 4; i++) t[i] = (idx & 2) ? t[2 * i + 1] : t[2 * i];
#pragma unroll
  for (int i = 0; i < AP / 8; i++) t[i] = (idx & 4) ? t[2 * i + 1] : t[2 * i];
  if (AP == 16) return (idx & 8) ? t[1] : t[0];
  return t[0];
}

template <int A, int AP>
AG_D void solve_node(const float (&p)[AP], const float (&q)[AP], const int (&vis)[AP], const int (&ch)[AP], const int (&ord)[AP],
                     const int nchild, const float cpuct, const bool prior_box, float (&pol)[AP], long long* tr = nullptr) {
  int nv = 0, acount = 0;
  float rem = 0.f;
#pragma unroll
  for (int a = 0; a < A; a++) {
    nv += vis[a];
    rem = fadd(rem, ch[a] == 0 ? p[a] : 0.f);
    acount += p[a] > 0.f ? 1 : 0;
  }
  const float n = (float)(1 + nv);
  const float lambda = fdiv(fmul(cpuct, fsqrt(n)), fadd((float)acount, n));      // :132
  rem = fmul(rem, lambda);                                                          // :134
  float alpha = 0.f;
  float top[AP];
#pragma unroll
  for (int a = 0; a < AP; a++) top[a] = a < A ? fmul(lambda, p[a]) : 0.f;
#pragma unroll
  for (int a = 0; a < A; a++) alpha = fmaxf(alpha, fadd(q[a], fmaxf(top[a], 1e-4f)));   // :135-138
  // statistics of the children in slot (creation) order — the order of the Newton sums — picked out of the registers
  float tops[AP], qs[AP];
#pragma unroll
  for (int k = 0; k < A; k++) {
    tops[k] = 0.f; qs[k] = 0.f;
    if (k < nchild) {
      const int a = ord[k] - 1;
      qs[k] = sel_reg<AP>(q, a);
      tops[k] = sel_reg<AP>(top, a);
    }
  }
  // Fast-division plan (common.cuh: fdiv_fast): every numerator below is loop-invariant and non-negative, every denominator positive.
  // The derivative is summed with all signs flipped, G = rem/α² + Σ tops/bot² = -gs: rounding is sign-symmetric, so -G is bit for bit the
  // reference's gs, and α - err/gs = α + err/G.
  // Operand checks in O(1): the numerators are λ·prior and λ·Σ prior — in the box when the node's priors are (F_PRIOR_BOX, decided once
  // by expand) and λ is in [2^-10, 2^10].  The denominators of an iteration are α and α - q_k: rounding is monotone, so the smallest and
  // the largest of them are α - max(0, q) and α - min(0, q), with the q range taken once per solve (an action without a child has q = 0).
  constexpr float SQ_LO = 9.313225746154785e-10f, SQ_HI = 1073741824.0f;             // 2^-30, 2^30: the squares stay inside the box
  float qmin = 0.f, qmax = 0.f;
#pragma unroll
  for (int a = 0; a < A; a++) { qmin = fminf(qmin, q[a]); qmax = fmaxf(qmax, q[a]); }
  const bool num_ok = prior_box && lambda >= 0.0009765625f && lambda <= 1024.0f;
  float err = __int_as_float(0x7f800000);
  const long long trs = tr ? clock64() + (__float_as_int(alpha) & 0) + (__float_as_int(tops[0]) & 0) + (__float_as_int(qs[A - 1]) & 0): 0;
  for (int it = 0; it < 100; it++) {                                                 // :141-162
    float bot[AP];
#pragma unroll
    for (int k = 0; k < A; k++) bot[k] = fsub(alpha, qs[k]);
    const bool fast = num_ok && fsub(alpha, qmax) >= SQ_LO && fsub(alpha, qmin) <= SQ_HI;   // (a NaN fails the comparisons)
    float S;
    if (fast) {
      // all quotients first — independent, branch-free, staged so that they overlap in the pipeline — then the adds in reference order
      float na[A + 1], nb[A + 1], nq[A + 1];
      na[0] = rem; nb[0] = alpha;
#pragma unroll
      for (int k = 0; k < A; k++) { na[k + 1] = tops[k]; nb[k + 1] = bot[k]; }
      fdiv_fast_n<A + 1>(na, nb, nq);
      S = nq[0];
#pragma unroll
      for (int k = 0; k < A; k++) if (k < nchild) S = fadd(S, nq[k + 1]);
    } else {
      S = fdiv(rem, alpha);
#pragma unroll
      for (int k = 0; k < A; k++) if (k < nchild) S = fadd(S, fdiv(tops[k], bot[k]));
    }
    const float newerr = fsub(S, 1.f);
    if (newerr < 0.001f || newerr == err) break;
    // the derivative is only needed when the iteration continues (the reference computes it in the same loop and drops it on exit)
    if (fast && alpha == alpha) {
      float na[A + 1], nb[A + 1], nq[A + 1];
      na[0] = rem; nb[0] = fmul(alpha, alpha);
#pragma unroll
      for (int k = 0; k < A; k++) { na[k + 1] = tops[k]; nb[k + 1] = fmul(bot[k], bot[k]); }
      fdiv_fast_n<A + 1>(na, nb, nq);
      float G = nq[0];
#pragma unroll
      for (int k = 0; k < A; k++) if (k < nchild) G = fadd(G, nq[k + 1]);
      alpha = fadd(alpha, (fdiv_box_num_bits(newerr) && fdiv_box_den_bits(G)) ? fdiv_fast(newerr, G) : fdiv(newerr, G));
    } else {
      float gs = fdiv(-rem, fmul(alpha, alpha));
#pragma unroll
      for (int k = 0; k < A; k++) if (k < nchild) gs = fadd(gs, fdiv(-tops[k], fmul(bot[k], bot[k])));
      alpha = fsub(alpha, fdiv(newerr, gs));
    }
    err = newerr;
  }
  if (tr) tr[3] += clock64() + (__float_as_int(alpha) & 0) - trs;
  {
    // π̄_a = λ p_a / (α - q_a) (:165-169)
    float num[A], den[A];
#pragma unroll
    for (int a = 0; a < A; a++) { num[a] = top[a]; den[a] = fsub(alpha, q[a]); }
    if (num_ok && fsub(alpha, qmax) >= FDIV_BOX_LO && fsub(alpha, qmin) <= FDIV_BOX_HI) {
      float nq[A];
      fdiv_fast_n<A>(num, den, nq);
#pragma unroll
      for (int a = 0; a < A; a++) pol[a] = nq[a];
    } else {
#pragma unroll
      for (int a = 0; a < A; a++) pol[a] = fdiv(num[a], den[a]);
    }
  }
#pragma unroll
  for (int a = A; a < AP; a++) pol[a] = 0.f;
}

// ------------------------------------------------------------------------------------------------
// select: kdescendTree! (mcts_gpu.jl:100-199).  One group of W lanes per game.
// ------------------------------------------------------------------------------------------------
// Large action sets: the ORDERED sums of a level (Newton sums over the children in creation order, the inverse-CDF scan and
// prior_rem in ascending action order) are the same dependent chain on every lane of the group, so their cost is issue slots, not
// latency: the addends are staged in a per-group shared-memory scratch and read back four (two pairs) per broadcast LDS.128, which
// leaves ~1.5 instructions per addend instead of a shuffle + select + add on every lane.
template <class G> struct SelScratch {
  static constexpr int FLOATS = Layout<G>::FAST ? 4 : 2 * Layout<G>::W * Layout<G>::APL;   // per game: (t1, t2) pairs of every child slot
  static constexpr int A4 = (G::A + 3) / 4 * 4;
};
// ascending-index fp32 sum of n4 (a multiple of 4) staged addends, continued from acc
AG_D float ordered_sum4(const float* __restrict__ sc, const int n4, float acc) {
  const float4* sc4 = reinterpret_cast<const float4*>(sc);
#pragma unroll
  for (int k = 0; k < n4; k += 4) {
    const float4 v = sc4[k >> 2];
    acc = fadd(fadd(fadd(fadd(acc, v.x), v.y), v.z), v.w);
  }
  return acc;
}
template <class G>
AG_D void select_game(const SearchParams& P, const int g, const int l, const unsigned gm, int L, int rollout, int last_rollout, float cpuct,
                      const float* __restrict__ prob, u64 seed, u32 ply, float* __restrict__ sc) {
  typedef Layout<G> Lay;
  constexpr int W = Lay::W, APL = Lay::APL, A = G::A, REC = Lay::REC;
  char* gbase = P.tree + (size_t)g * P.game_stride;
  int nn = P.nnodes[g];
  const u32 uid = P.uid[g];
  int node = 0, depth = 0, rblock = -1;
  Philox4 rnd; rnd.v[0] = rnd.v[1] = rnd.v[2] = rnd.v[3] = 0;

  while (true) {
    char* rec = gbase + (size_t)node * REC;
    const NodeHdr h = *reinterpret_cast<const NodeHdr*>(rec + Lay::OFF_HDR);
    float p[APL], q[APL], pol[APL];
    int vis[APL], ch[APL], ord[APL];
    if constexpr (Lay::FAST) {
      // header, child ids and π̄ share the record's first 64 bytes: all three loads are issued before the flag is tested
      static_assert(APL == 1, "FAST layouts have one action per lane");
      pol[0] = l < A ? *reinterpret_cast<const float*>(rec + Lay::OFF_POLICY + 4 * l) : 0.f;   // π̄ as left by expand / the last backup
      ch[0] = l < A ? (int)*reinterpret_cast<const uint8_t*>(rec + Lay::OFF_CHILD + l) : 0;
    }
    if (!(h.flags & F_EXPANDED)) break;                                   // while expanded[nindex]==1  (:110)
    if constexpr (Lay::FAST) {
      // the next node is one of the children: pull their first sectors towards L1 while this level is sampled
      if (ch[0] != 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(gbase + (size_t)(ch[0] - 1) * REC));
    } else {
    // what the solve needs from A-long sums is cached behind the header (NodeAux): Σ visits, prior_rem, #{prior > 0}
    const NodeAux ax = *reinterpret_cast<const NodeAux*>(rec + Lay::OFF_AUX);
#pragma unroll
    for (int j = 0; j < APL; j++) {
      const int a = j * W + l;
      const bool in = a < A;
      p[j] = in ? *reinterpret_cast<const float*>(rec + Lay::OFF_PRIOR + 4 * a) : 0.f;
      q[j] = in ? *reinterpret_cast<const float*>(rec + Lay::OFF_Q + 4 * a) : 0.f;
      ch[j] = in ? (int)*reinterpret_cast<const uint8_t*>(rec + Lay::OFF_CHILD + a) : 0;
      ord[j] = in ? (int)*reinterpret_cast<const uint8_t*>(rec + Lay::OFF_ORDER + a) : 0;
    }
    if (ax.nvis > 0) {                                                     // uptodate != 1  (:114): some backup passed through
      // n = 1 + Σ visits (exact), prior_rem = Σ_{no child} prior in ascending action order, A = #{prior > 0}   (:116-131)
      const float n = (float)(1 + (int)ax.nvis);
      const float lambda = fdiv(fmul(cpuct, fsqrt(n)), fadd((float)ax.acount, n));   // :132
      const float rem = fmul(ax.rem, lambda);                                        // :134
      float alpha = 0.f;                                                             // :135-138
      float top[APL];
#pragma unroll
      for (int j = 0; j < APL; j++) {
        top[j] = fmul(lambda, p[j]);
        if (j * W + l < A) alpha = fmaxf(alpha, fadd(q[j], fmaxf(top[j], 1e-4f)));
      }
      alpha = gmax<W>(gm, alpha);
      // The Newton sums run over the CHILDREN only, in creation (slot) order: lane l owns the children in slots l, l + W, … and reads
      // their statistics once (the record's q / prior lines were just loaded), so that an iteration costs two divisions per lane and
      // slot group instead of two per lane and action group.
      const int nchild = h.nchild;
      const int ngroups = (nchild + W - 1) / W;                                      // group-uniform, <= APL
      float ctop[APL], cq[APL];
#pragma unroll
      for (int j = 0; j < APL; j++) {
        ctop[j] = 0.f; cq[j] = 0.f;                                                  // an empty slot contributes +0 to both sums
        if (j * W + l < nchild) {
          const int a = ord[j] - 1;
          cq[j] = *reinterpret_cast<const float*>(rec + Lay::OFF_Q + 4 * a);
          ctop[j] = fmul(lambda, *reinterpret_cast<const float*>(rec + Lay::OFF_PRIOR + 4 * a));
        }
      }
      // Divisions: the branch-free fast path (common.cuh: fdiv_fast — correctly rounded inside its operand box, checked against
      // __fdiv_rn on 2^34 pairs) whenever every operand of every lane is in the box, else the IEEE division; same quotients either way.
      // A denominator b that is also squared is held to [2^-30, 2^30], which puts b*b in the box as well.  The numerators of the child
      // terms are a subset of top[].  The derivative is accumulated with all signs flipped (G = -gs: rounding is sign-symmetric), so
      // every numerator is >= 0.
      bool num_ok = fdiv_box_num_bits(rem);
#pragma unroll
      for (int j = 0; j < APL; j++) num_ok = num_ok && fdiv_box_num_bits(top[j]);
      num_ok = __all_sync(gm, num_ok);
      float2* sc2 = reinterpret_cast<float2*>(sc);
      const float4* sc4 = reinterpret_cast<const float4*>(sc);
      const f32x2 one2 = pack2f(1.0f, 1.0f);
      float err = __int_as_float(0x7f800000);
      for (int it = 0; it < 100; it++) {                                             // :141-162
        const float a2 = fmul(alpha, alpha);
        float bot[APL], bot2[APL];
        bool ok = num_ok && fdiv_box_den_sq(alpha);
#pragma unroll
        for (int j = 0; j < APL; j++) {
          bot[j] = fsub(alpha, cq[j]); bot2[j] = fmul(bot[j], bot[j]);
          if (j < ngroups) ok = ok && fdiv_box_den_sq(bot[j]);
        }
        ok = __all_sync(gm, ok);
        float S, G;
        if (ok) {
          S = fdiv_fast(rem, alpha);
          G = fdiv_fast(rem, a2);
#pragma unroll
          for (int j = 0; j < APL; j++)
            if (j < ngroups) sc2[j * W + l] = make_float2(fdiv_fast(ctop[j], bot[j]), fdiv_fast(ctop[j], bot2[j]));
        } else {
          S = fdiv(rem, alpha);
          G = fdiv(rem, a2);
#pragma unroll
          for (int j = 0; j < APL; j++)
            if (j < ngroups) sc2[j * W + l] = make_float2(fdiv(ctop[j], bot[j]), fdiv(ctop[j], bot2[j]));
        }
        __syncwarp(gm);
        // children in creation (slot) order, both sums side by side: fma(x, 1, t) is the correctly rounded x + t.  An odd count reads
        // one empty slot of the same group: + (+0) changes nothing.
        f32x2 SG = pack2f(S, G);
        for (int k = 0; k < nchild; k += 2) {
          const float4 v = sc4[k >> 1];
          SG = fma2(SG, one2, pack2f(v.x, v.y));
          SG = fma2(SG, one2, pack2f(v.z, v.w));
        }
        unpack2f(SG, S, G);
        __syncwarp(gm);                                                              // the scratch is rewritten by the next iteration / the scan
        const float newerr = fsub(S, 1.f);
        if (newerr < 0.001f || newerr == err) break;
        // α - err/gs with gs = -G
        alpha = fadd(alpha, (fdiv_box_num(newerr) && fdiv_box_den(G)) ? fdiv_fast(newerr, G) : fdiv(newerr, G));
        err = newerr;
      }
      {
        float den[APL];
        bool ok = num_ok;
#pragma unroll
        for (int j = 0; j < APL; j++) { den[j] = fsub(alpha, q[j]); ok = ok && fdiv_box_den_bits(den[j]); }
        ok = __all_sync(gm, ok);
#pragma unroll
        for (int j = 0; j < APL; j++) pol[j] = ok ? fdiv_fast(top[j], den[j]) : fdiv(top[j], den[j]);      // :165-169
      }
    } else {
#pragma unroll
      for (int j = 0; j < APL; j++) pol[j] = p[j];                                  // policy == prior until the first backup (:297-299)
    }
    }   // !FAST

    if (node == 0 && last_rollout) {                                                 // copy_pol (:330-339): π̄_root of the last descent
#pragma unroll
      for (int j = 0; j < APL; j++) if (j * W + l < A) P.policy_final[(size_t)g * A + j * W + l] = pol[j];
    }

    // uniform for this depth: injected prob[depth, game] (:178) or Philox
    float u;
    if (prob) u = prob[((size_t)rollout * L + g) * G::MAXLEN + depth];
    else {
      if ((depth >> 2) != rblock) { rblock = depth >> 2; rnd = philox4x32_10(uid, ply, (u32)rollout, (u32)rblock, (u32)seed, (u32)(seed >> 32)); }
      const int w = depth & 3;
      u = u01(w == 0 ? rnd.v[0] : w == 1 ? rnd.v[1] : w == 2 ? rnd.v[2] : rnd.v[3]);
    }
    // inverse-CDF scan in ascending action order (:172-182): the sample is the last action with π̄ > 0 at or before the first
    // index whose running sum reaches u
    int best = -1;
    if constexpr (Lay::FAST) {
      float cum = 0.f;
      bool done = false;                                                              // the same on every lane of the group
#pragma unroll
      for (int s = 0; s < W; s++) {
        if (s < A) {
          const float d = gshfl<W>(gm, pol[0], s);
          if (!done) {
            cum = fadd(cum, d);
            if (d > 0.f) best = s;
            if (cum >= u) done = true;
          }
        }
      }
    } else {
      constexpr int A4 = SelScratch<G>::A4;
#pragma unroll
      for (int j = 0; j < APL; j++) sc[j * W + l] = (j * W + l < A) ? pol[j] : 0.f;
      __syncwarp(gm);
      const float4* sc4 = reinterpret_cast<const float4*>(sc);
      float cum = 0.f;
      int kstar = 0;                                                                  // number of leading indices with sum < u
      bool open = true;
#pragma unroll
      for (int k = 0; k < A4; k += 4) {
        if ((k & 7) == 0 && k > 0 && !open) break;                                    // the rest of the scan changes nothing
        const float4 v = sc4[k >> 2];
        cum = fadd(cum, v.x); open = open && !(cum >= u); kstar += open ? 1 : 0;
        cum = fadd(cum, v.y); open = open && !(cum >= u); kstar += open ? 1 : 0;
        cum = fadd(cum, v.z); open = open && !(cum >= u); kstar += open ? 1 : 0;
        cum = fadd(cum, v.w); open = open && !(cum >= u); kstar += open ? 1 : 0;
      }
      __syncwarp(gm);
      const int lane0 = (int)(threadIdx.x & 31) - l;                                  // first lane of the group inside its warp
#pragma unroll
      for (int j = 0; j < APL; j++) {
        const int a = j * W + l;
        const unsigned b = __ballot_sync(gm, a < A && a <= kstar && pol[j] > 0.f) & gm;
        if (b) best = j * W + (31 - __clz(b)) - lane0;
      }
    }
    if (best < 0) best = 0;
    if (l == 0) {                                                                     // remember the path for the lane-parallel backup
      P.path_node[(size_t)g * P.R + depth] = (uint8_t)node;
      P.path_move[(size_t)g * P.R + depth] = (uint8_t)best;
    }
    int c;
    if constexpr (Lay::FAST) c = gshfl<W>(gm, ch[0], best);
    else c = (int)*reinterpret_cast<const uint8_t*>(rec + Lay::OFF_CHILD + best);      // (the line was loaded at the top of the level)
    if (c == 0) {                                                                     // allocate the child (:183-191)
      nn += 1;
      c = nn;
      if (l == best % W) *reinterpret_cast<uint8_t*>(rec + Lay::OFF_CHILD + best) = (uint8_t)c;
      const typename G::State ps = *reinterpret_cast<const typename G::State*>(gbase + Lay::state_off(P.R, node));
      const typename G::State ns = G::play(ps, best + 1);
      int res = 0;
      const bool term = G::is_over(ns, res);
      char* nrec = gbase + (size_t)(c - 1) * REC;
#pragma unroll
      for (int j = 0; j < APL; j++) {
        const int a = j * W + l;
        *reinterpret_cast<float*>(nrec + Lay::OFF_Q + 4 * a) = 0.f;
        *reinterpret_cast<uint16_t*>(nrec + Lay::OFF_VIS + 2 * a) = 0;
        *reinterpret_cast<uint8_t*>(nrec + Lay::OFF_CHILD + a) = 0;
      }
      if constexpr (!Lay::FAST) {
        // prior_rem of the parent now excludes the new child: the same ascending sum over the actions that still have none
        constexpr int A4 = SelScratch<G>::A4;
#pragma unroll
        for (int j = 0; j < APL; j++) sc[j * W + l] = (j * W + l < A && ch[j] == 0 && j * W + l != best) ? p[j] : 0.f;
        __syncwarp(gm);
        const float rem = ordered_sum4(sc, A4, 0.f);
        if (l == 0) {
          reinterpret_cast<NodeAux*>(rec + Lay::OFF_AUX)->rem = rem;
          hdr_store(nrec + Lay::OFF_AUX, 0ull);
        }
      }
      if (l == 0) {
        if constexpr (Lay::ORD_IN_HDR) {
          const u64 w = *reinterpret_cast<const u64*>(rec + Lay::OFF_HDR);
          hdr_store(rec + Lay::OFF_HDR, (w & ~(0xFFull << 16)) | ((u64)(h.nchild + 1) << 16) | ((u64)(best + 1) << (40 + 3 * h.nchild)));
        } else {
          *reinterpret_cast<uint8_t*>(rec + Lay::OFF_ORDER + h.nchild) = (uint8_t)(best + 1);
          reinterpret_cast<NodeHdr*>(rec + Lay::OFF_HDR)->nchild = (uint8_t)(h.nchild + 1);
        }
        *reinterpret_cast<typename G::State*>(gbase + Lay::state_off(P.R, c - 1)) = ns;
        hdr_store(nrec + Lay::OFF_HDR, hdr_word(node + 1, best + 1, 0, term ? F_TERMINAL : 0, res));
      }
      node = c - 1;
      depth += 1;
      break;                                                                           // a new node is unexpanded: the while ends
    }
    node = c - 1;                                                                      // :192
    depth += 1;
  }
  if (l == 0) {
    P.leaf[g] = node;                                                                  // :195
    P.nnodes[g] = nn;
    P.path_len[g] = (uint8_t)depth;
    if (P.counters) { atomicAdd(&P.counters[0], (unsigned long long)depth); atomicAdd(&P.counters[1], 1ull); }
  }
}

template <class G>
__global__ void __launch_bounds__(Layout<G>::SB, Layout<G>::SB_MIN) select_kernel(SearchParams P, int L, int rollout, int last_rollout, float cpuct,
                                                     const float* __restrict__ prob, u64 seed, u32 ply) {
  constexpr int W = Layout<G>::W;
  __shared__ __align__(16) float s_sel[(Layout<G>::SB / W) * SelScratch<G>::FLOATS];
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) / W;
  if (g >= L) return;
  select_game<G>(P, g, threadIdx.x & (W - 1), group_mask<W>(), L, rollout, last_rollout, cpuct, prob, seed, ply,
                 s_sel + (threadIdx.x / W) * SelScratch<G>::FLOATS);
}

// ------------------------------------------------------------------------------------------------
// expand + backUp (mcts_gpu.jl:250-328), with softmax! (:417) when the prior comes from the network.
// INJECT: prior_in[L][A] is already softmaxed (what `expand` receives), value_in[L].
// ------------------------------------------------------------------------------------------------
// what expand leaves for the backup of the same game
struct LeafEval {
  float v;            // network value of the leaf (non-terminal)
  double value0_d;    // terminal value (1 + player*r)/2, a Float64 in the reference (:314)
  int term;           // leaf is terminal
  int parent, action; // vnodes.parent / actionFromParent of the leaf (1-based, 0 = none)
};

template <class G, bool INJECT>
AG_D LeafEval expand_game(const SearchParams& P, const int g, const int l, const unsigned gm, int training, int last_rollout,
                          const float* __restrict__ prior_in, const float* __restrict__ value_in, float* __restrict__ sc = nullptr) {
  typedef Layout<G> Lay;
  constexpr int W = Lay::W, APL = Lay::APL, A = G::A, REC = Lay::REC;
  char* gbase = P.tree + (size_t)g * P.game_stride;
  const int leaf = P.leaf[g];
  char* rec = gbase + (size_t)leaf * REC;
  const NodeHdr h = *reinterpret_cast<const NodeHdr*>(rec + Lay::OFF_HDR);
  const typename G::State st = *reinterpret_cast<const typename G::State*>(gbase + Lay::state_off(P.R, leaf));
  const bool term = (h.flags & F_TERMINAL) != 0;
  float v = 0.f;

  if (!term) {                                                                         // expand: :258-296
    float pin[APL];
    if (INJECT) {
#pragma unroll
      for (int j = 0; j < APL; j++) { const int a = j * W + l; pin[j] = a < A ? prior_in[(size_t)g * A + a] : 0.f; }
      v = value_in[g];
    } else {
      const float* o = P.nn_out + (size_t)g * Lay::OUTS;
      float x[APL];
      float m = -__int_as_float(0x7f800000);
#pragma unroll
      for (int j = 0; j < APL; j++) { const int a = j * W + l; x[j] = a < A ? o[a] : 0.f; if (a < A) m = fmaxf(m, x[j]); }
      m = gmax<W>(gm, m);
      float e[APL];
      float ssum = 0.f;
#pragma unroll
      for (int j = 0; j < APL; j++) {
        e[j] = (j * W + l < A) ? c_expf(fsub(x[j], m)) : 0.f;
        if constexpr (Lay::FAST) {
#pragma unroll
          for (int s = 0; s < W; s++) if (j * W + s < A) ssum = fadd(ssum, gshfl<W>(gm, e[j], s));
        } else {
          sc[j * W + l] = e[j];                                                          // ordered sums: staged (see SelScratch)
        }
      }
      if constexpr (!Lay::FAST) {
        __syncwarp(gm);
        ssum = ordered_sum4(sc, SelScratch<G>::A4, 0.f);
        __syncwarp(gm);
      }
#pragma unroll
      for (int j = 0; j < APL; j++) pin[j] = fdiv(e[j], ssum);
      v = o[A];
    }
    bool legal[APL];
    float normalize = 0.f;
    int acount = 0;
#pragma unroll
    for (int j = 0; j < APL; j++) {
      const int a = j * W + l;
      legal[j] = a < A && G::can_play(st, a + 1);
      const float contrib = legal[j] ? pin[j] : 0.f;
      acount += gcount<W>(gm, legal[j]);
      if constexpr (Lay::FAST) {
#pragma unroll
        for (int s = 0; s < W; s++) if (j * W + s < A) normalize = fadd(normalize, gshfl<W>(gm, contrib, s));
      } else {
        sc[a] = contrib;
      }
    }
    if constexpr (!Lay::FAST) {
      __syncwarp(gm);
      normalize = ordered_sum4(sc, SelScratch<G>::A4, 0.f);
      __syncwarp(gm);
    }
    const bool rootmix = (leaf == 0) && training;                                       // :259-275
    const float unif = fdiv(0.25f, (float)acount);
    float prv[APL];
#pragma unroll
    for (int j = 0; j < APL; j++) {
      const int a = j * W + l;
      float pr = 0.f;
      if (legal[j]) pr = rootmix ? fadd(fdiv(fmul(0.75f, pin[j]), normalize), unif) : fdiv(pin[j], normalize);
      prv[j] = pr;
      *reinterpret_cast<float*>(rec + Lay::OFF_PRIOR + 4 * a) = pr;
      if (Lay::FAST) *reinterpret_cast<float*>(rec + Lay::OFF_POLICY + 4 * a) = pr;      // policy[:,leaf] = prior[:,leaf]  (:297-299)
      if (leaf == 0 && last_rollout && a < A) P.policy_final[(size_t)g * A + a] = pr;    // R == 1: policy[:,1] is the prior itself
    }
    uint8_t nflags = (uint8_t)(h.flags | F_EXPANDED);
    if constexpr (Lay::FAST) {
      bool pbox = true;
#pragma unroll
      for (int j = 0; j < APL; j++) pbox = pbox && (__float_as_uint(prv[j]) == 0u || prv[j] >= PRIOR_BOX_LO);
      if (__all_sync(gm, pbox)) nflags |= F_PRIOR_BOX;
    }
    if (l == 0) reinterpret_cast<NodeHdr*>(rec + Lay::OFF_HDR)->flags = nflags;         // :256
    if constexpr (!Lay::FAST) {
      // NodeAux of the freshly expanded node: no child yet, so prior_rem is the ascending sum of the whole prior; #{prior > 0}; no visits
      int pos = 0;
#pragma unroll
      for (int j = 0; j < APL; j++) {
        pos += gcount<W>(gm, prv[j] > 0.f);
        sc[j * W + l] = prv[j];
      }
      __syncwarp(gm);
      const float rem = ordered_sum4(sc, SelScratch<G>::A4, 0.f);
      __syncwarp(gm);
      if (l == 0) {
        NodeAux ax; ax.rem = rem; ax.acount = (uint16_t)pos; ax.nvis = 0;
        *reinterpret_cast<NodeAux*>(rec + Lay::OFF_AUX) = ax;
      }
    }
  }

  LeafEval E;
  E.v = v; E.term = term ? 1 : 0; E.parent = h.parent; E.action = h.action;
  E.value0_d = (double)(1 + (int)(int8_t)(st.player * h.result)) * 0.5;
  (void)gbase; (void)REC;
  return E;
}

// One ancestor of one game, by one lane (FAST layouts): running mean of the child's value (:319-320) and the re-solve of π̄.
// jj = index in the recorded path (0 = root), d = path length.  Each ancestor is a different node, so items are independent;
// the value an ancestor receives is the leaf value flipped once per level below it (value = 1 - value, :324), evaluated as that
// literal chain.  π̄ is not re-solved after the last rollout: nobody reads it (policy_final is the root policy of the last DESCENT, :443).
template <class G>
AG_D void backup_item(const SearchParams& P, const int g, const int jj, const int d, const LeafEval& E, int last_rollout, const float cpuct,
                      long long* tr = nullptr, const uint16_t* s_path = nullptr,
                      unsigned char* s_cache_row = nullptr, const int nc_nodes = 0) {
  const long long tr0 = tr ? clock64() : 0;
  typedef Layout<G> Lay;
  constexpr int A = G::A, REC = Lay::REC, AP = Lay::APAD;
  char* gbase = P.tree + (size_t)g * P.game_stride;
  const bool term = E.term != 0;
  const float v = E.v;
  const double value0_d = E.value0_d;
  {
    {
      {
        const int flips = d - 1 - jj;
        // s_path: this game's path in shared memory (fused kernel), PATH_SMEM_DEPTH entries of node | move << 8
        int nd, mv;
        if (s_path != nullptr && jj < PATH_SMEM_DEPTH) { const int e = s_path[jj]; nd = e & 0xFF; mv = e >> 8; }
        else { nd = P.path_node[(size_t)g * P.R + jj]; mv = P.path_move[(size_t)g * P.R + jj]; }
        char* nrec = gbase + (size_t)nd * REC;
        float p[AP], q[AP], pol[AP];
        int vis[AP], ch[AP], ord[AP];
#pragma unroll
        for (int c = 0; c < AP / 4; c++) {
          const float4 pv = hot_ld_f4(nrec + Lay::OFF_PRIOR + 16 * c);
          const float4 qv = hot_ld_f4(nrec + Lay::OFF_Q + 16 * c);
          p[4 * c] = pv.x; p[4 * c + 1] = pv.y; p[4 * c + 2] = pv.z; p[4 * c + 3] = pv.w;
          q[4 * c] = qv.x; q[4 * c + 1] = qv.y; q[4 * c + 2] = qv.z; q[4 * c + 3] = qv.w;
        }
        if (Lay::VIS_IN_PAD) {
          const uint32_t v03 = __float_as_uint(q[7]), v46 = __float_as_uint(p[7]);
#pragma unroll
          for (int a = 0; a < 8; a++) vis[a] = a < 4 ? (int)((v03 >> (8 * a)) & 0xFFu) : a < 7 ? (int)((v46 >> (8 * (a - 4))) & 0xFFu) : 0;
        } else {
#pragma unroll
          for (int c = 0; c < AP / 8; c++) {
            const uint4 vv = hot_ld_u4(nrec + Lay::OFF_VIS + 16 * c);
            const uint32_t w4[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
            for (int e = 0; e < 4; e++) { vis[8 * c + 2 * e] = (int)(w4[e] & 0xFFFFu); vis[8 * c + 2 * e + 1] = (int)(w4[e] >> 16); }
          }
        }
        uint32_t hw;                                                                   // parent | action | nchild | flags
        if (Lay::ORD_IN_HDR) {                                                         // header (+ creation order) + child ids: one 16-byte request
          const uint4 hc = hot_ld_u4(nrec + Lay::OFF_HDR);
          hw = hc.x;
          const uint32_t cw[2] = {hc.z, hc.w};
#pragma unroll
          for (int e = 0; e < 8; e++) { ch[e] = (int)((cw[e >> 2] >> (8 * (e & 3))) & 0xFFu); ord[e] = e < 7 ? (int)((hc.y >> (8 + 3 * e)) & 7u) : 0; }
        } else {
#pragma unroll
          for (int c = 0; c < AP / 8; c++) {          // child bytes then order bytes, AP bytes each, contiguous
            const uint2 cv = hot_ld_u2(nrec + Lay::OFF_CHILD + 8 * c);
            const uint2 ov = hot_ld_u2(nrec + Lay::OFF_ORDER + 8 * c);
            const uint32_t cw[2] = {cv.x, cv.y}, ow[2] = {ov.x, ov.y};
#pragma unroll
            for (int e = 0; e < 8; e++) { ch[8 * c + e] = (int)((cw[e >> 2] >> (8 * (e & 3))) & 0xFFu); ord[8 * c + e] = (int)((ow[e >> 2] >> (8 * (e & 3))) & 0xFFu); }
          }
          hw = *reinterpret_cast<const uint32_t*>(nrec + Lay::OFF_HDR);
        }
        const int nchild = (int)((hw >> 16) & 0xFFu);
        const bool prior_box = ((hw >> 24) & F_PRIOR_BOX) != 0;
        // running mean of the child's value from this node's point of view (:319-320)
        float qold = 0.f; int vold = 0;
#pragma unroll
        for (int a = 0; a < A; a++) if (a == mv) { qold = q[a]; vold = vis[a]; }
        const float vf = (float)vold;
        float qnew;
        if (term) {
          double val = value0_d;
          for (int t = 0; t < flips; t++) val = __dsub_rn(1.0, val);
          qnew = (float)__ddiv_rn(__dadd_rn((double)fmul(vf, qold), __dsub_rn(1.0, val)), (double)fadd(vf, 1.f));
        } else {
          float val = v;
          for (int t = 0; t < flips; t++) val = fsub(1.f, val);
          qnew = fdiv(fadd(fmul(vf, qold), fsub(1.f, val)), fadd(vf, 1.f));
        }
#pragma unroll
        for (int a = 0; a < A; a++) if (a == mv) { q[a] = qnew; vis[a] = vold + 1; }
        *reinterpret_cast<float*>(nrec + Lay::OFF_Q + 4 * mv) = qnew;
        if (Lay::VIS_IN_PAD) *reinterpret_cast<uint8_t*>(nrec + (mv < 4 ? Lay::OFF_Q + 28 + mv : Lay::OFF_PRIOR + 24 + mv)) = (uint8_t)(vold + 1);
        else *reinterpret_cast<uint16_t*>(nrec + Lay::OFF_VIS + 2 * mv) = (uint16_t)(vold + 1);
        long long tr1 = 0;
        if (tr) { tr1 = clock64() + (__float_as_int(qnew) & 0); tr[0] += tr1 - tr0; }
        if (!last_rollout) {
          solve_node<A, AP>(p, q, vis, ch, ord, nchild, cpuct, prior_box, pol, tr);
          if (tr) { tr[1] += clock64() + (__float_as_int(pol[0]) & 0) - tr1; tr[2] += 1; }
#pragma unroll
          for (int c = 0; c < AP / 4; c++)
            *reinterpret_cast<float4*>(nrec + Lay::OFF_POLICY + 16 * c) = make_float4(pol[4 * c], pol[4 * c + 1], pol[4 * c + 2], pol[4 * c + 3]);
          if (s_cache_row != nullptr && nd < nc_nodes) {                              // write-through: π̄ of a cached node
#pragma unroll
            for (int c = 0; c < AP / 4; c++)
              *reinterpret_cast<float4*>(s_cache_row + nd * CacheSlot<AP>::BYTES + CacheSlot<AP>::OFF_POLICY + 16 * c) =
                  make_float4(pol[4 * c], pol[4 * c + 1], pol[4 * c + 2], pol[4 * c + 3]);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// FAST layouts, ONE THREAD per game (the fused per-ply kernel): with π̄ stored in the node the descent only samples, and almost
// everything a lane group did was the same scalar work repeated on every lane (header decode, Philox, move generation for the new
// node) — measured on B200 the 8-lane descent of 223 games kept an SM's issue slots busy for 28 k cycles per rollout.  One thread
// per game issues an eighth of the instructions; the seven-term prefix scan it serialises is shorter than the shuffles it replaces.
// Same operations in the same order as select_game / expand_game: results are bit-identical.
//
// What the descent needs besides the tree comes from shared memory (RolloutShared): the root's state (constant for the ply; every
// other state on the path is play() of it — a child's stored state IS play(parent state, action) — so no state is ever loaded), and
// the Philox block of depths 0..3, computed ahead of time under the previous network phase.  The copies in global memory of what the
// descent hands to the next phases (leaf, node count, path) are written on the last rollout of a ply — that is when the stand-alone
// kernels and the read-back entry points look at them — and for path entries beyond the shared-memory window.
// The walk only reads; the allocation of the new leaf (move generation, terminal test, nine stores) happens once, after the loop, when
// the lanes of the warp have reconverged — inside the loop it was executed divergently, level after level, by whichever lanes had
// reached their leaf.
// CACHE: the small-batch kernel's node cache (above).
// ------------------------------------------------------------------------------------------------
template <class G, bool CACHE>
AG_D void select_game1(const SearchParams& P, const int g, const int gl, const RolloutShared<G>& SH, const u32 uid, int& nn, int rollout, int last_rollout,
                       u64 seed, u32 ply, const unsigned gmask, long long* tr = nullptr) {
  const long long tr0 = tr ? clock64() : 0;
  long long trA = tr0, trB = tr0, trC = tr0, trD = tr0;
  typedef Layout<G> Lay;
  typedef CacheSlot<Lay::APAD> CS;
  static_assert(Lay::FAST, "thread-per-game descent needs the stored policy");
  constexpr int A = G::A, REC = Lay::REC, AP = Lay::APAD;
  char* gbase = P.tree + (size_t)g * P.game_stride;
  int node = 0, depth = 0, rblock = 0;
  Philox4 rnd = SH.rnd[gl];
  uint8_t* pnode = P.path_node + (size_t)g * P.R;
  uint8_t* pmove = P.path_move + (size_t)g * P.R;
  uint2 hw;
  typename G::State cur = SH.root[gl];
  int best = 0, nchild = 0;
  bool create = false;
  u64 cw0 = 0, cw1 = 0;                                                              // child ids of the current node, 8 per word

  while (true) {
    char* rec = gbase + (size_t)node * REC;
    // header, child ids and π̄ are the record's first bytes: every load of the level is issued before the flag is tested
    static_assert(AP == 8 || AP == 16, "child ids are read as one or two 64-bit words");
    float pol[AP];
    const unsigned char* sl = node_cache_slot<G, CACHE>(SH, gl, node);                 // this node's cache entry, if it has one
    if (CACHE && sl != nullptr) {
      if (AP == 8) {                                                                   // header + child ids: one 16-byte request
        const uint4 hc = *reinterpret_cast<const uint4*>(sl);
        hw = make_uint2(hc.x, hc.y);
        cw0 = (u64)hc.z | ((u64)hc.w << 32);
      } else {
        hw = *reinterpret_cast<const uint2*>(sl);
        const uint2 cv = *reinterpret_cast<const uint2*>(sl + CS::OFF_CHILD);
        cw0 = (u64)cv.x | ((u64)cv.y << 32);
      }
      if (AP == 16) {
        const uint2 cv1 = *reinterpret_cast<const uint2*>(sl + CS::OFF_CHILD + 8);
        cw1 = (u64)cv1.x | ((u64)cv1.y << 32);
      }
#pragma unroll
      for (int c = 0; c < AP / 4; c++) {
        const float4 pv = *reinterpret_cast<const float4*>(sl + CS::OFF_POLICY + 16 * c);
        pol[4 * c] = pv.x; pol[4 * c + 1] = pv.y; pol[4 * c + 2] = pv.z; pol[4 * c + 3] = pv.w;
      }
    } else {
      // child ids, 8 per 64-bit word (plain scalars: an indexed array would live in local memory, and its store would stall on the load)
      // (the search phases are bound by the number of load/store requests: header and child ids, adjacent in the record, travel together)
      if (AP == 8) {
        static_assert(AP != 8 || (Lay::OFF_HDR % 16 == 0 && Lay::OFF_CHILD == Lay::OFF_HDR + 8 && Lay::REC % 16 == 0), "header + child ids as one 16-byte load");
        const uint4 hc = hot_ld_u4(rec + Lay::OFF_HDR);
        hw = make_uint2(hc.x, hc.y);
        cw0 = (u64)hc.z | ((u64)hc.w << 32);
      } else {
        hw = hot_ld_u2(rec + Lay::OFF_HDR);
        const uint2 cv = hot_ld_u2(rec + Lay::OFF_CHILD);
        cw0 = (u64)cv.x | ((u64)cv.y << 32);
      }
      if (AP == 16) {
        const uint2 cv1 = hot_ld_u2(rec + Lay::OFF_CHILD + 8);
        cw1 = (u64)cv1.x | ((u64)cv1.y << 32);
      }
#pragma unroll
      for (int c = 0; c < AP / 4; c++) {
        const float4 pv = hot_ld_f4(rec + Lay::OFF_POLICY + 16 * c);
        pol[4 * c] = pv.x; pol[4 * c + 1] = pv.y; pol[4 * c + 2] = pv.z; pol[4 * c + 3] = pv.w;
      }
    }
    // the uniform of this depth does not depend on the loads above: a Philox block beyond the first runs while they are in flight
    if ((depth >> 2) != rblock) { rblock = depth >> 2; rnd = philox4x32_10(uid, ply, (u32)rollout, (u32)rblock, (u32)seed, (u32)(seed >> 32)); }
    const int w = depth & 3;
    const float u = u01(w == 0 ? rnd.v[0] : w == 1 ? rnd.v[1] : w == 2 ? rnd.v[2] : rnd.v[3]);
    nchild = (int)((hw.x >> 16) & 0xFFu);
    const int flags = (int)(hw.x >> 24);
    if (tr && depth == 0) trA = clock64() + (hw.x & 0) + (__float_as_int(pol[A - 1]) & 0) + ((u32)cw0 & 0);
    if (!(flags & F_EXPANDED)) break;                                                 // while expanded[nindex]==1  (:110): the root before its
                                                                                      // first evaluation, or a terminal node
    if (node == 0 && last_rollout) {                                                  // copy_pol (:330-339)
#pragma unroll
      for (int a = 0; a < A; a++) P.policy_final[(size_t)g * A + a] = pol[a];
    }
    if (tr && depth == 0) trB = clock64() + (__float_as_int(u) & 0);
    // inverse-CDF scan in ascending action order (:172-182)
    float cum = 0.f;
    best = -1;
    bool done = false;
#pragma unroll
    for (int a = 0; a < A; a++) {
      const float d = pol[a];
      if (!done) {
        cum = fadd(cum, d);
        if (d > 0.f) best = a;
        if (cum >= u) done = true;
      }
    }
    if (best < 0) best = 0;
    if (tr && depth == 0) trC = clock64() + (best & 0);
    if (last_rollout || depth >= PATH_SMEM_DEPTH) { pnode[depth] = (uint8_t)node; pmove[depth] = (uint8_t)best; }
    if (depth < PATH_SMEM_DEPTH) SH.path[gl * PATH_SMEM_DEPTH + depth] = (uint16_t)(node | (best << 8));
    const int c = (int)(((AP == 16 && best >= 8 ? cw1 : cw0) >> (8 * (best & 7))) & 0xFFu);
    if (c == 0) { create = true; break; }                                              // the child does not exist yet: allocate it below
    node = c - 1;                                                                      // :192
    depth += 1;
    cur = G::play(cur, best + 1);
    if (tr && depth == 1) trD = clock64() + (node & 0);
  }

  if (create) {                                                                        // allocate the child (:183-191)
    char* rec = gbase + (size_t)node * REC;
    nn += 1;
    const int c = nn;
    uint4 phc = make_uint4(0, 0, 0, 0);
    if (Lay::ORD_IN_HDR) {                           // the parent's header (child count, creation order) and child ids are in registers: one 16-byte store
      const u64 ncw = cw0 | ((u64)(uint32_t)c << (8 * best));
      phc = make_uint4((hw.x & 0xFF00FFFFu) | ((uint32_t)(nchild + 1) << 16), hw.y | ((uint32_t)(best + 1) << (8 + 3 * nchild)), (uint32_t)ncw, (uint32_t)(ncw >> 32));
      *reinterpret_cast<uint4*>(rec + Lay::OFF_HDR) = phc;
    } else {
      *reinterpret_cast<uint8_t*>(rec + Lay::OFF_CHILD + best) = (uint8_t)c;
      reinterpret_cast<NodeHdr*>(rec + Lay::OFF_HDR)->nchild = (uint8_t)(nchild + 1);
      *reinterpret_cast<uint8_t*>(rec + Lay::OFF_ORDER + nchild) = (uint8_t)(best + 1);
    }
    const typename G::State ns = G::play(cur, best + 1);
    int res = 0;
    const bool term = G::is_over(ns, res);
    char* nrec = gbase + (size_t)(c - 1) * REC;
#pragma unroll
    for (int k = 0; k < AP / 4; k++) *reinterpret_cast<float4*>(nrec + Lay::OFF_Q + 16 * k) = make_float4(0.f, 0.f, 0.f, 0.f);
    const u64 nhw = hdr_word(node + 1, best + 1, 0, term ? F_TERMINAL : 0, res);
    if (!Lay::VIS_IN_PAD) {                          // (else the counts live in the zeroed q vector / the prior vector written by expand)
#pragma unroll
      for (int k = 0; k < AP / 8; k++) *reinterpret_cast<uint4*>(nrec + Lay::OFF_VIS + 16 * k) = make_uint4(0, 0, 0, 0);
    }
    if (Lay::ORD_IN_HDR) {                           // header + (empty) child ids: one 16-byte store
      *reinterpret_cast<uint4*>(nrec + Lay::OFF_HDR) = make_uint4((uint32_t)nhw, (uint32_t)(nhw >> 32), 0u, 0u);
    } else {
#pragma unroll
      for (int k = 0; k < AP / 8; k++) *reinterpret_cast<uint2*>(nrec + Lay::OFF_CHILD + 8 * k) = make_uint2(0, 0);
      hdr_store(nrec + Lay::OFF_HDR, nhw);
    }
    state_store(gbase + Lay::state_off(P.R, c - 1), ns);
    if (CACHE) {
      if (unsigned char* csl = node_cache_slot<G, CACHE>(SH, gl, node)) {              // the parent's entry: child id, child count, order
        if (Lay::ORD_IN_HDR) *reinterpret_cast<uint4*>(csl) = phc;
        else { csl[CS::OFF_CHILD + best] = (uint8_t)c; csl[2] = (uint8_t)(nchild + 1); }
      }
      if (unsigned char* nsl = node_cache_slot<G, CACHE>(SH, gl, c - 1)) {             // the new node's entry (π̄ is written by expand)
        hdr_store(nsl, nhw);
#pragma unroll
        for (int k = 0; k < AP / 8; k++) *reinterpret_cast<uint2*>(nsl + CS::OFF_CHILD + 8 * k) = make_uint2(0, 0);
      }
    }
    SH.state[gl] = ns;
    hdr_store(&SH.hdr[gl], nhw | ((u64)(uint8_t)(int8_t)(ns.player * res) << 40));   // + player * result for the backup items (leaf_eval1)
    node = c - 1;
    depth += 1;
  } else {
    SH.state[gl] = cur;
    *reinterpret_cast<uint2*>(&SH.hdr[gl]) = make_uint2(hw.x, (hw.y & 0xFFu) | ((uint32_t)(uint8_t)(int8_t)(cur.player * (int8_t)(hw.y & 0xFFu)) << 8));
  }
  if (tr) { tr[4] += clock64() + (node & 0) - tr0; tr[5] += depth; tr[8] += trA - tr0; tr[9] += trB - tr0; tr[10] += trC - tr0; tr[6] += trD - tr0; }
  SH.leaf[gl] = (uint8_t)node;
  SH.d[gl] = depth;
  {
    // The (game, path index) pairs of this descent are the items of the next backup phase.  One list per CTA, filled with one
    // shared-memory atomic per warp: an inclusive scan of the path lengths over the lanes of the warp (gmask — they have reconverged
    // here) gives every game its range.  A game whose range does not fit the list keeps its tail for itself: SH.ovf[gl] = first path
    // index that is not listed, and its own thread backs those up before the barrier of the pool.
    __syncwarp(gmask);
    const int lane = (int)(threadIdx.x & 31);
    int incl = depth;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(gmask, incl, o); if (lane >= o && ((gmask >> (lane - o)) & 1u)) incl += y; }
    const int top = 31 - __clz((int)gmask);                                           // gmask is a prefix of the warp: its last lane holds the total
    const int total = __shfl_sync(gmask, incl, top);
    int base = 0;
    if (lane == top) base = atomicAdd(SH.lv_cnt, total);
    base = __shfl_sync(gmask, base, top) + incl - depth;
    int listed = depth;
    if (base + depth > SH.lv_cap) listed = base < SH.lv_cap ? SH.lv_cap - base : 0;
    for (int j = 0; j < listed; j++) SH.lv_item[base + j] = (uint16_t)(gl | (j << 8));
    SH.ovf[gl] = (uint8_t)listed;
  }
  if (last_rollout) {
    P.leaf[g] = node;                                                                  // :195
    P.nnodes[g] = nn;
    P.path_len[g] = (uint8_t)depth;
  }
  if (P.counters) { atomicAdd(&P.counters[0], (unsigned long long)depth); atomicAdd(&P.counters[1], 1ull); }
}

// expand (mcts_gpu.jl:250-302) + softmax! (:417) of the leaf of local game gl, by one thread; everything it needs about the leaf is in
// shared memory (RolloutShared).  Returns nothing: the backup items read the leaf's value and terminal flag from the same place.
template <class G, bool CACHE>
AG_D void expand_game1(const SearchParams& P, const int g, const int gl, const RolloutShared<G>& SH, int training, int last_rollout) {
  typedef Layout<G> Lay;
  typedef CacheSlot<Lay::APAD> CS;
  static_assert(Lay::FAST, "thread-per-game expand is written for the FAST record");
  constexpr int A = G::A, REC = Lay::REC, AP = Lay::APAD;
  const int leaf = SH.leaf[gl];
  char* rec = P.tree + (size_t)g * P.game_stride + (size_t)leaf * REC;
  const NodeHdr h = SH.hdr[gl];
  if ((h.flags & F_TERMINAL) != 0) return;
  const typename G::State st = SH.state[gl];
  {                                                                                    // expand: :258-296, softmax! :417
    float x[Lay::OUTS];
#pragma unroll
    for (int c = 0; c < Lay::OUTS / 4; c++) {
      const float4 ov = *reinterpret_cast<const float4*>(SH.out + (gl >> 7) * SH.out_tile_stride + (gl & 127) * Lay::OUTS + 4 * c);
      x[4 * c] = ov.x; x[4 * c + 1] = ov.y; x[4 * c + 2] = ov.z; x[4 * c + 3] = ov.w;
    }
    float m = -__int_as_float(0x7f800000);
#pragma unroll
    for (int a = 0; a < A; a++) m = fmaxf(m, x[a]);
    float e[A], ssum = 0.f;
#pragma unroll
    for (int a = 0; a < A; a++) { e[a] = c_expf(fsub(x[a], m)); ssum = fadd(ssum, e[a]); }
    // the two rounds of divisions (softmax, renormalisation over the legal moves) as packed fast-path batches when every operand is
    // in the box of fdiv_fast (common.cuh), else the plain IEEE division: same quotients either way
    {
      float den[A], qn[A];
      bool ok = fdiv_box_den(ssum);
#pragma unroll
      for (int a = 0; a < A; a++) { den[a] = ssum; ok = ok && fdiv_box_num(e[a]); }
      if (ok) {
        fdiv_fast_n<A>(e, den, qn);
#pragma unroll
        for (int a = 0; a < A; a++) e[a] = qn[a];
      } else {
#pragma unroll
        for (int a = 0; a < A; a++) e[a] = fdiv(e[a], ssum);
      }
    }
    float normalize = 0.f;
    int acount = 0;
    bool legal[A];
#pragma unroll
    for (int a = 0; a < A; a++) {
      legal[a] = G::can_play(st, a + 1);
      normalize = fadd(normalize, legal[a] ? e[a] : 0.f);
      acount += legal[a] ? 1 : 0;
    }
    const bool rootmix = (leaf == 0) && training;                                       // :259-275
    const float unif = fdiv(0.25f, (float)acount);
    float pr[AP];
#pragma unroll
    for (int a = 0; a < AP; a++) pr[a] = 0.f;
    {
      float num[A], den[A], qn[A];
      bool ok = fdiv_box_den(normalize);
#pragma unroll
      for (int a = 0; a < A; a++) { num[a] = rootmix ? fmul(0.75f, e[a]) : e[a]; den[a] = normalize; ok = ok && fdiv_box_num(num[a]); }
      if (ok) {
        fdiv_fast_n<A>(num, den, qn);
      } else {
#pragma unroll
        for (int a = 0; a < A; a++) qn[a] = legal[a] ? fdiv(num[a], normalize) : 0.f;
      }
#pragma unroll
      for (int a = 0; a < A; a++)
        if (legal[a]) pr[a] = rootmix ? fadd(qn[a], unif) : qn[a];
    }
#pragma unroll
    for (int c = 0; c < AP / 4; c++) {
      const float4 pv = make_float4(pr[4 * c], pr[4 * c + 1], pr[4 * c + 2], pr[4 * c + 3]);
      *reinterpret_cast<float4*>(rec + Lay::OFF_PRIOR + 16 * c) = pv;
      *reinterpret_cast<float4*>(rec + Lay::OFF_POLICY + 16 * c) = pv;                 // policy[:,leaf] = prior[:,leaf]  (:297-299)
    }
    if (leaf == 0 && last_rollout) {
#pragma unroll
      for (int a = 0; a < A; a++) P.policy_final[(size_t)g * A + a] = pr[a];
    }
    bool pbox = true;
#pragma unroll
    for (int a = 0; a < A; a++) pbox = pbox && (__float_as_uint(pr[a]) == 0u || pr[a] >= PRIOR_BOX_LO);
    const uint8_t nflags = (uint8_t)(h.flags | F_EXPANDED | (pbox ? F_PRIOR_BOX : 0));
    reinterpret_cast<NodeHdr*>(rec + Lay::OFF_HDR)->flags = nflags;                     // :256
    if (CACHE) {
      if (unsigned char* sl = node_cache_slot<G, CACHE>(SH, gl, leaf)) {                 // write-through: π̄ = prior, expanded flag
#pragma unroll
        for (int c = 0; c < AP / 4; c++)
          *reinterpret_cast<float4*>(sl + CS::OFF_POLICY + 16 * c) = make_float4(pr[4 * c], pr[4 * c + 1], pr[4 * c + 2], pr[4 * c + 3]);
        sl[3] = nflags;                                                                  // NodeHdr::flags
      }
    }
  }
}

// what a backup item needs to know about the leaf of its game, from the shared-memory hand-off
template <class G>
AG_D LeafEval leaf_eval1(const RolloutShared<G>& SH, const int gl) {
  typedef Layout<G> Lay;
  const NodeHdr h = SH.hdr[gl];
  LeafEval E;
  E.term = (h.flags & F_TERMINAL) ? 1 : 0;
  E.v = SH.out[(gl >> 7) * SH.out_tile_stride + (gl & 127) * Lay::OUTS + G::A];
  E.parent = h.parent; E.action = h.action;
  E.value0_d = (double)(1 + (int)(int8_t)h.pad[0]) * 0.5;                           // pad[0] of the hand-off copy: player * result (select_game1)
  return E;
}

template <class G, bool INJECT>
AG_D void expand_backup_game(const SearchParams& P, const int g, const int l, const unsigned gm, int training, int last_rollout,
                             const float* __restrict__ prior_in, const float* __restrict__ value_in, const float cpuct,
                             float* __restrict__ sc = nullptr) {
  typedef Layout<G> Lay;
  constexpr int W = Lay::W, REC = Lay::REC;
  char* gbase = P.tree + (size_t)g * P.game_stride;
  const LeafEval E = expand_game<G, INJECT>(P, g, l, gm, training, last_rollout, prior_in, value_in, sc);
  // backUp: :306-328
  if constexpr (Lay::FAST) {
    // lane-parallel: lane jj takes the jj-th node of the recorded path
    const int d = P.path_len[g];
    for (int base = 0; base < d; base += W) {
      const int jj = base + l;
      if (jj < d) backup_item<G>(P, g, jj, d, E, last_rollout, cpuct);
    }
    return;
  }
  // Large action sets: the same lane-parallel walk over the recorded path (one ancestor per lane: the loads of all levels are in
  // flight together instead of one parent pointer after the other); only the visited action's statistics change here, π̄ is solved
  // by the next descent.  The value an ancestor receives is the leaf value flipped once per level below it, as that literal chain.
  const int d = P.path_len[g];
  for (int base = 0; base < d; base += W) {
    const int jj = base + l;
    if (jj < d) {
      const int nd = P.path_node[(size_t)g * P.R + jj], mv = P.path_move[(size_t)g * P.R + jj];
      char* nrec = gbase + (size_t)nd * REC;
      float* qp = reinterpret_cast<float*>(nrec + Lay::OFF_Q + 4 * mv);
      uint16_t* vp = reinterpret_cast<uint16_t*>(nrec + Lay::OFF_VIS + 2 * mv);
      uint16_t* nv = reinterpret_cast<uint16_t*>(nrec + Lay::OFF_AUX + 6);              // NodeAux::nvis
      const int vold = *vp, nvold = *nv;
      const float qold = *qp;
      const float vf = (float)vold;
      const int flips = d - 1 - jj;
      float qnew;
      if (E.term) {
        // value = (1 + player*r)/2 is a Float64 in the reference (:314): the running mean on this path is evaluated in double
        double val = E.value0_d;
        for (int t = 0; t < flips; t++) val = __dsub_rn(1.0, val);
        qnew = (float)__ddiv_rn(__dadd_rn((double)fmul(vf, qold), __dsub_rn(1.0, val)), (double)fadd(vf, 1.f));
      } else {
        float val = E.v;
        for (int t = 0; t < flips; t++) val = fsub(1.f, val);                            // :324
        qnew = fdiv(fadd(fmul(vf, qold), fsub(1.f, val)), fadd(vf, 1.f));               // :319
      }
      *qp = qnew;
      *vp = (uint16_t)(vold + 1);                                                        // :320
      *nv = (uint16_t)(nvold + 1);
    }
  }
}

template <class G, bool INJECT>
__global__ void __launch_bounds__(Layout<G>::SB) expand_backup_kernel(SearchParams P, int L, int training, int last_rollout,
                                                            const float* __restrict__ prior_in, const float* __restrict__ value_in, float cpuct) {
  constexpr int W = Layout<G>::W;
  __shared__ __align__(16) float s_sel[(Layout<G>::SB / W) * SelScratch<G>::FLOATS];
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) / W;
  if (g >= L) return;
  expand_backup_game<G, INJECT>(P, g, threadIdx.x & (W - 1), group_mask<W>(), training, last_rollout, prior_in, value_in, cpuct,
                                s_sel + (threadIdx.x / W) * SelScratch<G>::FLOATS);
}

// One launch per rollout for the search side: expand + backUp of rollout k-1 (its network output is ready) followed at once by the
// descent of rollout k.  The group that just walked a game's path back up re-descends through the same, still cached, records.
template <class G>
__global__ void __launch_bounds__(Layout<G>::SB, Layout<G>::SB_MIN) step_kernel(SearchParams P, int L, int rollout, int last_rollout, int training, float cpuct, u64 seed,
                                                   u32 ply) {
  constexpr int W = Layout<G>::W;
  __shared__ __align__(16) float s_sel[(Layout<G>::SB / W) * SelScratch<G>::FLOATS];
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) / W;
  if (g >= L) return;
  const int l = threadIdx.x & (W - 1);
  const unsigned gm = group_mask<W>();
  expand_backup_game<G, false>(P, g, l, gm, training, 0, nullptr, nullptr, cpuct, s_sel + (threadIdx.x / W) * SelScratch<G>::FLOATS);
  __syncwarp(gm);                                      // the group's global writes (q, visits, prior, flags) are ordered before its reads
  select_game<G>(P, g, l, gm, L, rollout, last_rollout, cpuct, nullptr, seed, ply, s_sel + (threadIdx.x / W) * SelScratch<G>::FLOATS);
}

// ---- the same three launches, addressed through a SegParams record (graph replay, one stream per slice) ----
template <class G>
__global__ void __launch_bounds__(Layout<G>::SB, Layout<G>::SB_MIN) select_seg_kernel(SearchParams P, const SegParams* __restrict__ sp, int rollout, int last_rollout) {
  constexpr int W = Layout<G>::W;
  __shared__ __align__(16) float s_sel[(Layout<G>::SB / W) * SelScratch<G>::FLOATS];
  const int gl = (blockIdx.x * blockDim.x + threadIdx.x) / W;
  const SegParams S = *sp;
  if (gl >= S.len) return;
  select_game<G>(P, S.off + gl, threadIdx.x & (W - 1), group_mask<W>(), 0, rollout, last_rollout, S.cpuct, nullptr, S.seed, S.ply,
                 s_sel + (threadIdx.x / W) * SelScratch<G>::FLOATS);
}
template <class G>
__global__ void __launch_bounds__(Layout<G>::SB, Layout<G>::SB_MIN) step_seg_kernel(SearchParams P, const SegParams* __restrict__ sp, int rollout, int last_rollout) {
  constexpr int W = Layout<G>::W;
  __shared__ __align__(16) float s_sel[(Layout<G>::SB / W) * SelScratch<G>::FLOATS];
  const int gl = (blockIdx.x * blockDim.x + threadIdx.x) / W;
  const SegParams S = *sp;
  if (gl >= S.len) return;
  const int g = S.off + gl, l = threadIdx.x & (W - 1);
  const unsigned gm = group_mask<W>();
  expand_backup_game<G, false>(P, g, l, gm, S.training, 0, nullptr, nullptr, S.cpuct, s_sel + (threadIdx.x / W) * SelScratch<G>::FLOATS);
  __syncwarp(gm);
  select_game<G>(P, g, l, gm, 0, rollout, last_rollout, S.cpuct, nullptr, S.seed, S.ply, s_sel + (threadIdx.x / W) * SelScratch<G>::FLOATS);
}
template <class G>
__global__ void __launch_bounds__(Layout<G>::SB) expand_seg_kernel(SearchParams P, const SegParams* __restrict__ sp, int last_rollout) {
  constexpr int W = Layout<G>::W;
  __shared__ __align__(16) float s_sel[(Layout<G>::SB / W) * SelScratch<G>::FLOATS];
  const int gl = (blockIdx.x * blockDim.x + threadIdx.x) / W;
  const SegParams S = *sp;
  if (gl >= S.len) return;
  expand_backup_game<G, false>(P, S.off + gl, threadIdx.x & (W - 1), group_mask<W>(), S.training, last_rollout, nullptr, nullptr, S.cpuct,
                               s_sel + (threadIdx.x / W) * SelScratch<G>::FLOATS);
}

// ------------------------------------------------------------------------------------------------
// decoder / decoder_roots (mcts_gpu.jl:202-246) to fp32 for the API, one thread per (game, feature)
// ------------------------------------------------------------------------------------------------
template <class G>
__global__ void encode_nodes_kernel(SearchParams P, int L, int use_leaf, float* __restrict__ batch) {
  typedef Layout<G> Lay;
  const int F = 2 * G::VS;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)L * F) return;
  int g = (int)(t / F), j = (int)(t % F);
  int node = use_leaf ? P.leaf[g] : 0;
  const typename G::State* st = reinterpret_cast<const typename G::State*>(P.tree + (size_t)g * P.game_stride + Lay::state_off(P.R, node));
  batch[t] = enc_bit<G>(*st, j) ? 1.f : 0.f;
}

// plugin surface on plain state arrays (agpu_can_play / agpu_play / agpu_is_over / agpu_encode)
template <class G>
__global__ void game_ops_kernel(const typename G::State* __restrict__ in, const int32_t* __restrict__ actions, int n,
                                typename G::State* __restrict__ played, uint8_t* __restrict__ legal, uint8_t* __restrict__ over,
                                int8_t* __restrict__ result, float* __restrict__ enc, int init_only) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (init_only) { played[i] = G::init(); return; }
  const typename G::State s = in[i];
  if (legal) for (int a = 1; a <= G::A; a++) legal[(size_t)i * G::A + a - 1] = G::can_play(s, a) ? 1 : 0;
  if (played) played[i] = G::play(s, actions[i]);
  if (over) { int r = 0; bool f = G::is_over(s, r); over[i] = f ? 1 : 0; result[i] = (int8_t)r; }
  if (enc) for (int j = 0; j < 2 * G::VS; j++) enc[(size_t)i * 2 * G::VS + j] = enc_bit<G>(s, j) ? 1.f : 0.f;
}

// ------------------------------------------------------------------------------------------------
// Ply bookkeeping of the self-play / duel loops (mcts_gpu.jl:506-561, 600-639), one thread per game.
// ------------------------------------------------------------------------------------------------
struct SampleBufs {          // device SoA (Game.Sample, main4IARow.jl:29-37)
  int8_t* state; float* policy; int8_t* player; float* value; int8_t* fstate; int32_t* game; int32_t* ply;
  long long capacity;
};
struct PlyState {
  void* next_state;          // [L] State after the move
  uint8_t* alive;            // [L] 1 = game continues
  int32_t* block_count;      // [blocks] survivors per block
  int8_t* game_result;       // [ngames] winner colour
  void* game_final;          // [ngames] final State (for fstate)
  unsigned long long* tallies;  // [0..2] v,n,d  [3] total_length  [4] faults
};

// StatsBase.sample over the non-zero weights (mcts_gpu.jl:518-521) / over all entries in the duel (:605-606);
// argmax (first maximal index) afterwards (:523, :608).  Returns a 1-based action.
template <class G, bool DUEL>
AG_D int choose_move(const float* __restrict__ pol, u32 round, float u) {
  constexpr int A = G::A;
  if (round < (DUEL ? 15u : 25u)) {
    float wsum = 0.f;
    int last = A;
    if (!DUEL) last = 1;
    for (int c = 0; c < A; c++) {
      const float w = pol[c];
      if (DUEL) wsum = fadd(wsum, w);
      else if (w != 0.f) { wsum = fadd(wsum, w); last = c + 1; }
    }
    const float t = fmul(u, wsum);
    float cw = 0.f;
    bool first = true;
    for (int c = 0; c < A; c++) {
      const float w = pol[c];
      if (!DUEL && w == 0.f) continue;
      if (first) { cw = w; first = false; } else cw = fadd(cw, w);
      if (!(cw < t) || c + 1 == last) return c + 1;
    }
    return last;
  }
  int best = 0;
  float bv = pol[0];
  for (int c = 1; c < A; c++) if (pol[c] > bv) { bv = pol[c]; best = c; }
  return best + 1;
}

template <class G, bool DUEL>
__global__ void __launch_bounds__(256) finish_ply_kernel(SearchParams P, int L, u32 round, u64 seed, u32 uid_base, SampleBufs S,
                                                         long long sample_base, PlyState Y) {
  typedef Layout<G> Lay;
  typedef typename G::State State;
  constexpr int A = G::A;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  bool alive = false;
  if (g < L) {
    const State st = *reinterpret_cast<const State*>(P.tree + (size_t)g * P.game_stride + Lay::state_off(P.R, 0));
    const u32 uid = P.uid[g];
    const float* pol = P.policy_final + (size_t)g * A;
    if (!DUEL) {                                                                        // push_buffer (main4IARow.jl:49-63)
      const long long row = sample_base + g;
      if (row < S.capacity) {
        for (int j = 0; j < 2 * G::VS; j++) S.state[row * 2 * G::VS + j] = enc_bit<G>(st, j) ? 1 : 0;
        for (int a = 0; a < A; a++) S.policy[row * A + a] = pol[a];
        S.player[row] = st.player;
        S.game[row] = (int32_t)uid;
        S.ply[row] = (int32_t)round;
      }
    }
    const Philox4 r = philox4x32_10(uid, round, ROLLOUT_MOVE, 0u, (u32)seed, (u32)(seed >> 32));
    const int c = choose_move<G, DUEL>(pol, round, u01(r.v[0]));
    if (!G::can_play(st, c)) atomicAdd(&Y.tallies[4], 1ull);                             // "faute" (:526-529)
    const State ns = G::play(st, c);                                                     // :530
    int res = 0;
    const bool f = G::is_over(ns, res);                                                  // :531
    reinterpret_cast<State*>(Y.next_state)[g] = ns;
    alive = !f;
    if (f) {
      const u32 local = uid - uid_base;
      Y.game_result[local] = (int8_t)res;
      reinterpret_cast<State*>(Y.game_final)[local] = ns;
      atomicAdd(&Y.tallies[res == 1 ? 0 : (res == 0 ? 1 : 2)], 1ull);                   // :541-547
      atomicAdd(&Y.tallies[3], (unsigned long long)round);                               // :535
    }
    Y.alive[g] = alive ? 1 : 0;
  }
  const int cnt = __syncthreads_count(alive);
  if (threadIdx.x == 0) Y.block_count[blockIdx.x] = cnt;
}

// exclusive scan of the per-block survivor counts (<= 1024 blocks per pass, looped), total -> *total_out
static __global__ void __launch_bounds__(1024) scan_blocks_kernel(int32_t* __restrict__ block_count, int nblocks, int32_t* __restrict__ total_out) {
  __shared__ int sm[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nblocks ? block_count[i] : 0;
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      int t = threadIdx.x >= o ? sm[threadIdx.x - o] : 0;
      __syncthreads();
      sm[threadIdx.x] += t;
      __syncthreads();
    }
    const int incl = sm[threadIdx.x];
    if (i < nblocks) block_count[i] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = carry;
}

// order-preserving compaction (deleteat!, mcts_gpu.jl:550-553) fused with re_init of the next ply (:557-561)
template <class G>
__global__ void __launch_bounds__(256) compact_kernel(SearchParams P, int L, PlyState Y, typename G::State* __restrict__ state_out,
                                                      uint32_t* __restrict__ uid_out) {
  typedef typename G::State State;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const bool alive = g < L && Y.alive[g];
  // block-level exclusive prefix of `alive`
  __shared__ int wsum[8];
  const unsigned b = __ballot_sync(0xffffffffu, alive);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int inwarp = __popc(b & ((1u << lane) - 1u));
  if (lane == 0) wsum[w] = __popc(b);
  __syncthreads();
  int off = Y.block_count[blockIdx.x];
  for (int k = 0; k < w; k++) off += wsum[k];
  if (alive) {
    const int ng = off + inwarp;
    state_out[ng] = reinterpret_cast<const State*>(Y.next_state)[g];
    uid_out[ng] = P.uid[g];
  }
}

// update_buffer (main4IARow.jl:65-75) for every sample once all games have ended
template <class G>
__global__ void finalize_samples_kernel(SampleBufs S, long long count, u32 uid_base, const int8_t* __restrict__ game_result,
                                        const typename G::State* __restrict__ game_final) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= count || row >= S.capacity) return;
  const u32 local = (u32)S.game[row] - uid_base;
  const int res = game_result[local];
  const int player = S.player[row];
  S.value[row] = (float)((double)(1 + res * player) / 2.0);
  const typename G::State fs = game_final[local];
  for (int j = 0; j < G::FS; j++) {
    const int f = bb_get0<typename G::Geo>(fs.bp, j) ? fs.player : -fs.player;           // decode (mcts_gpu.jl:464-474)
    S.fstate[row * G::FS + j] = (int8_t)(f * player);
  }
}

}  // namespace ag
