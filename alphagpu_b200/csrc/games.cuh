// games.cuh — the game-plugin surface of AlphaGPU (Position / canPlay / play / isOver / encoding and
// the constants VectorizedState, FeatureSize, maxActions, maxLengthGame) as static-polymorphic C++
// types usable in kernels.  Actions are 1-based as in the reference.  Semantics follow, and results
// are bit-identical to, 4IARow.jl, Gobang.jl, Hex.jl, Reversi8x8.jl and Reversi6x6.jl; only the
// evaluation is reorganised (constant masks instead of per-column loops, ctz instead of a row loop).
//
// A game type G provides:
//   Geo, A, VS, FS, MAXLEN, WIRE_BYTES, HAS_LEGAL, struct State, init(), can_play(), legal_mask() (Reversi),
//   play(), is_over().
#pragma once
#include "common.cuh"

namespace ag {

// Julia isbits layouts crossing the C ABI (Bitboard.jl:5-9, 4IARow.jl:16-21, Reversi8x8.jl:73-78)
struct WireBB { u64 chunks[3]; int64_t len; int64_t dims[2]; };
struct WirePos2 { WireBB bplayer, bopponent; int8_t player, aux; int8_t pad[6]; };
struct WirePos3 { WireBB bplayer, bopponent, legalplay; int8_t player; int8_t pad[7]; };
static_assert(sizeof(WireBB) == 48 && sizeof(WirePos2) == 104 && sizeof(WirePos3) == 152, "Julia Position layout");

template <class Geo>
struct State2 {       // Connect4 / Gobang / Hex (two boards + player + round|lp)
  BB<Geo> bp, bo;     // bplayer = side to move, bopponent = side that just moved
  int8_t player, aux;
};
template <class Geo>
struct State3 {       // Reversi (cached legal moves of the side to move)
  BB<Geo> bp, bo, lp;
  int8_t player, aux;
};

// k-in-a-row test on one board: 4IARow.jl:47-78 / Gobang.jl:36-67
template <class Geo, int NV>
AG_HD bool row_test(const BB<Geo>& b0) {
  BB<Geo> b = b0;
#pragma unroll
  for (int j = 1; j < NV; j++) b = b & bb_right<Geo>(b);
  if (bb_any<Geo>(b)) return true;
  b = b0;
#pragma unroll
  for (int j = 1; j < NV; j++) b = b & bb_down<Geo>(b);
  if (bb_any<Geo>(b)) return true;
  b = b0;
#pragma unroll
  for (int j = 1; j < NV; j++) b = b & bb_down<Geo>(bb_right<Geo>(b));
  if (bb_any<Geo>(b)) return true;
  b = b0;
#pragma unroll
  for (int j = 1; j < NV; j++) b = b & bb_left<Geo>(bb_down<Geo>(b));
  return bb_any<Geo>(b);
}

// ---------------------------------------------------------------- Connect4 (4IARow.jl)
struct Connect4 {
  typedef Geom<6, 7> Geo;
  typedef State2<Geo> State;
  static constexpr int A = 7, VS = 42, FS = 42, MAXLEN = 42, WIRE_BYTES = 104;
  static constexpr bool HAS_LEGAL = false;
  static constexpr int GAME_ID = 0, PN = 0, PNV = 4;
  AG_HD static State init() { State s; s.bp = bb_zero<Geo>(); s.bo = bb_zero<Geo>(); s.player = 1; s.aux = 1; return s; }   // :23
  AG_HD static bool can_play(const State& s, int col) { return !(((s.bp.c[0] | s.bo.c[0]) >> (6 * (col - 1))) & 1); }      // :25-27
  AG_HD static State play(const State& s, int col) {                                                                         // :30-44
    u32 colbits = (u32)(((s.bp.c[0] | s.bo.c[0]) >> (6 * (col - 1))) & 0x3F);
#ifdef __CUDA_ARCH__
    int fr = __ffs((int)(colbits | 0x40)) - 1;   // empty rows counted from the top (row 1)
#else
    int fr = __builtin_ctz(colbits | 0x40);
#endif
    if (fr == 0) fr = 1;                           // full column: `free` keeps its initial 1 (:31)
    State r;
    r.bp = s.bo;
    r.bo = s.bp; r.bo.c[0] |= u64(1) << (6 * (col - 1) + fr - 1);
    r.player = (int8_t)(-s.player); r.aux = (int8_t)(s.aux + 1);
    return r;
  }
  AG_HD static bool is_over(const State& s, int& res) {                                                                      // :47-81
    if (row_test<Geo, 4>(s.bo)) { res = -s.player; return true; }
    res = 0;
    return bb_count<Geo>(s.bp) + bb_count<Geo>(s.bo) == MAXLEN;
  }
};

// ---------------------------------------------------------------- Gobang (Gobang.jl), N x N, NV in a row
template <int N, int NV>
struct Gobang {
  typedef Geom<N, N> Geo;
  typedef State2<Geo> State;
  static constexpr int A = N * N, VS = N * N, FS = N * N, MAXLEN = N * N, WIRE_BYTES = 104;
  static constexpr bool HAS_LEGAL = false;
  static constexpr int GAME_ID = 1, PN = N, PNV = NV;
  AG_HD static State init() { State s; s.bp = bb_zero<Geo>(); s.bo = bb_zero<Geo>(); s.player = 1; s.aux = 0; return s; }   // :23
  AG_HD static bool can_play(const State& s, int col) { return !bb_get0<Geo>(s.bp | s.bo, col - 1); }                       // :25-27
  AG_HD static State play(const State& s, int col) {                                                                         // :30-33
    State r; r.bp = s.bo; r.bo = bb_set0<Geo>(s.bp, col - 1); r.player = (int8_t)(-s.player); r.aux = (int8_t)(s.aux + 1);
    return r;
  }
  AG_HD static bool is_over(const State& s, int& res) {                                                                      // :36-70
    if (row_test<Geo, NV>(s.bo)) { res = -s.player; return true; }
    res = 0;
    return bb_count<Geo>(s.bp) + bb_count<Geo>(s.bo) == N * N;
  }
};

// ---------------------------------------------------------------- Hex (Hex.jl), N x N cells on an (N+1)^2 bordered board
template <int N>
struct Hex {
  typedef Geom<N + 1, N + 1> Geo;
  typedef State2<Geo> State;
  static constexpr int A = N * N, VS = (N + 1) * (N + 1), FS = VS, MAXLEN = N * N, WIRE_BYTES = 104;
  static constexpr bool HAS_LEGAL = false;
  static constexpr int GAME_ID = 2, PN = N, PNV = 0;
  // 0-based bit of [r, c] (1-based)
  static constexpr int bit_rc(int r, int c) { return (N + 1) * (c - 1) + r - 1; }
  // Hex.jl:24-33: x owns [3..N+1, 1], o owns [1, 3..N+1]
  static constexpr u64 startx(int ch) { u64 m = 0; for (int i = 3; i <= N + 1; i++) { int b = bit_rc(i, 1); if (b / 64 == ch) m |= u64(1) << (b % 64); } return m; }
  static constexpr u64 starto(int ch) { u64 m = 0; for (int i = 3; i <= N + 1; i++) { int b = bit_rc(1, i); if (b / 64 == ch) m |= u64(1) << (b % 64); } return m; }
  // border re-injected after reduction round j when player == 1: [1, k] for k = 3+j..N+1 (Hex.jl:60-64)
  static constexpr u64 border(int j, int ch) { u64 m = 0; for (int k = 3 + j; k <= N + 1; k++) { int b = bit_rc(1, k); if (b / 64 == ch) m |= u64(1) << (b % 64); } return m; }
  AG_HD static State init() {                                                                                                // :22-35
    State s;
#pragma unroll
    for (int k = 0; k < Geo::NC; k++) { s.bp.c[k] = startx(k); s.bo.c[k] = starto(k); }
    s.player = 1; s.aux = (int8_t)(N * N);
    return s;
  }
  AG_HD static int cell_bit(int col) { int x = (col - 1) / N, y = col - N * x; return (N + 1) * (x + 1) + y; }               // newcol-1, :38-40
  AG_HD static bool can_play(const State& s, int col) { return !bb_get0<Geo>(s.bp | s.bo, cell_bit(col)); }                  // :37-42
  AG_HD static State play(const State& s, int col) {                                                                         // :45-51
    State r; r.bp = s.bo; r.bo = bb_set0<Geo>(s.bp, cell_bit(col)); r.player = (int8_t)(-s.player); r.aux = (int8_t)(s.aux - 1);
    return r;
  }
  AG_HD static bool is_over(const State& s, int& res) {                                                                      // :54-67
    BB<Geo> a = s.bo;
    const bool inject = (s.player == 1);
#pragma unroll
    for (int j = 1; j <= 2 * N - 2; j++) {
      BB<Geo> b = bb_up<Geo>(a);
      BB<Geo> c = bb_right<Geo>(b);
      a = bb_down<Geo>((a & (b | c)) | (b & c));
      if (inject) {
#pragma unroll
        for (int k = 0; k < Geo::NC; k++) a.c[k] |= border(j, k);
      }
    }
    res = -s.player;
    return bb_get0<Geo>(a, bit_rc(N + 1, N + 1));
  }
};

// ---------------------------------------------------------------- Reversi (Reversi8x8.jl / Reversi6x6.jl)
template <int N>
struct Reversi {
  typedef Geom<N, N> Geo;
  typedef State3<Geo> State;
  static constexpr int A = N * N + 1, VS = N * N, FS = N * N, MAXLEN = (N == 8 ? 70 : 50), WIRE_BYTES = 152;
  static constexpr bool HAS_LEGAL = true;
  static constexpr int GAME_ID = (N == 8 ? 3 : 4), PN = 0, PNV = 0;
  typedef BB<Geo> B;
  // directions (Reversi8x8.jl:17-23): 0 up 1 down 2 left 3 right 4 diaghd 5 diaghg 6 diagbd 7 diagbg
  template <int D> AG_HD static B dir(const B& x) {
    if (D == 0) return bb_up<Geo>(x);
    if (D == 1) return bb_down<Geo>(x);
    if (D == 2) return bb_left<Geo>(x);
    if (D == 3) return bb_right<Geo>(x);
    if (D == 4) return bb_up<Geo>(bb_right<Geo>(x));
    if (D == 5) return bb_up<Geo>(bb_left<Geo>(x));
    if (D == 6) return bb_down<Geo>(bb_right<Geo>(x));
    return bb_down<Geo>(bb_left<Geo>(x));
  }
  template <int D> AG_HD static B legal_dir(const B& me, const B& opp, const B& vide) {                                      // :26-35
    B moves = bb_zero<Geo>();
    B cand = dir<D>(me) & opp;
    while (bb_any<Geo>(cand)) {
      B nxt = dir<D>(cand);
      moves = moves | (vide & nxt);
      cand = opp & nxt;
    }
    return moves;
  }
  AG_HD static B legalplay(const B& me, const B& opp) {                                                                      // :37-40 (OR is order-free)
    B vide = bb_not<Geo>(me) & bb_not<Geo>(opp);
    return legal_dir<0>(me, opp, vide) | legal_dir<1>(me, opp, vide) | legal_dir<2>(me, opp, vide) | legal_dir<3>(me, opp, vide) |
           legal_dir<4>(me, opp, vide) | legal_dir<5>(me, opp, vide) | legal_dir<6>(me, opp, vide) | legal_dir<7>(me, opp, vide);
  }
  template <int D> AG_HD static B flip_dir(const B& me, const B& opp, const B& mv) {                                         // :44-56
    B cand = dir<D>(mv) & opp;
    B toflip = cand;
    while (bb_any<Geo>(cand)) { cand = opp & dir<D>(cand); toflip = toflip | cand; }
    return bb_any<Geo>(dir<D>(toflip) & me) ? toflip : bb_zero<Geo>();
  }
  AG_HD static B flip(const B& me, const B& opp, int c) {                                                                    // :58-70
    B mv = bb_bit0<Geo>(c - 1);
    return flip_dir<0>(me, opp, mv) | flip_dir<1>(me, opp, mv) | flip_dir<2>(me, opp, mv) | flip_dir<3>(me, opp, mv) |
           flip_dir<4>(me, opp, mv) | flip_dir<5>(me, opp, mv) | flip_dir<6>(me, opp, mv) | flip_dir<7>(me, opp, mv);
  }
  // 0-based bit of [r, c]
  static constexpr int bit_rc(int r, int c) { return N * (c - 1) + r - 1; }
  AG_HD static State init() {                                                                                                // :10-14,80-82
    State s;
    s.bp = bb_zero<Geo>(); s.bo = bb_zero<Geo>();
    if (N == 8) { s.bp = bb_set0<Geo>(bb_set0<Geo>(s.bp, bit_rc(4, 5)), bit_rc(5, 4)); s.bo = bb_set0<Geo>(bb_set0<Geo>(s.bo, bit_rc(5, 5)), bit_rc(4, 4)); }
    else { s.bp = bb_set0<Geo>(bb_set0<Geo>(s.bp, bit_rc(4, 3)), bit_rc(3, 4)); s.bo = bb_set0<Geo>(bb_set0<Geo>(s.bo, bit_rc(3, 3)), bit_rc(4, 4)); }
    s.lp = legalplay(s.bp, s.bo);
    s.player = 1; s.aux = 0;
    return s;
  }
  AG_HD static bool can_play(const State& s, int c) {                                                                        // :84-90
    if (c == A) return !bb_any<Geo>(s.lp);
    return bb_get0<Geo>(s.lp, c - 1);
  }
  AG_HD static State play(const State& s, int c) {                                                                           // :93-106
    State r; r.aux = 0; r.player = (int8_t)(-s.player);
    if (c == A) { r.bp = s.bo; r.bo = s.bp; r.lp = legalplay(s.bo, s.bp); return r; }
    B h = flip(s.bp, s.bo, c);
    B me = bb_set0<Geo>(s.bp ^ h, c - 1);
    B opp = s.bo ^ h;
    r.bp = opp; r.bo = me; r.lp = legalplay(opp, me);
    return r;
  }
  AG_HD static bool is_over(const State& s, int& res) {                                                                      // Reversi8x8.jl:109-131, Reversi6x6.jl:109-130
    bool over = !bb_any<Geo>(s.lp) && !bb_any<Geo>(legalplay(s.bo, s.bp));
    int test = (int)(int8_t)(bb_count<Geo>(s.bp) - bb_count<Geo>(s.bo));
    int sg = (test > 0) - (test < 0);
    res = (N == 6 && !over) ? 0 : sg * s.player;
    return over;
  }
};

// encoding of one side for the network input / samples (mcts_gpu.jl:202-223): bit j-1 of a board, j = 1..VS
template <class G> AG_HD bool enc_bit(const typename G::State& s, int j /*0-based in 0..2VS-1*/) {
  return j < G::VS ? bb_get0<typename G::Geo>(s.bp, j) : bb_get0<typename G::Geo>(s.bo, j - G::VS);
}

// ---- wire (Julia isbits) <-> device state ----
template <class G> inline void wire_bb(const BB<typename G::Geo>& b, WireBB* w) {
  for (int k = 0; k < 3; k++) w->chunks[k] = k < G::Geo::NC ? b.c[k] : 0;
  w->len = G::Geo::LEN; w->dims[0] = G::Geo::H; w->dims[1] = G::Geo::W;
}
template <class G> inline BB<typename G::Geo> unwire_bb(const WireBB* w) {
  BB<typename G::Geo> b;
  for (int k = 0; k < G::Geo::NC; k++) b.c[k] = w->chunks[k];
  return b;
}
template <class G> inline void to_wire(const typename G::State& s, void* out) {
  if constexpr (G::HAS_LEGAL) {
    WirePos3 w; memset(&w, 0, sizeof(w));
    wire_bb<G>(s.bp, &w.bplayer); wire_bb<G>(s.bo, &w.bopponent); wire_bb<G>(s.lp, &w.legalplay); w.player = s.player;
    memcpy(out, &w, sizeof(w));
  } else {
    WirePos2 w; memset(&w, 0, sizeof(w));
    wire_bb<G>(s.bp, &w.bplayer); wire_bb<G>(s.bo, &w.bopponent); w.player = s.player; w.aux = s.aux;
    memcpy(out, &w, sizeof(w));
  }
}
template <class G> inline typename G::State from_wire(const void* in) {
  typename G::State s;
  if constexpr (G::HAS_LEGAL) {
    WirePos3 w; memcpy(&w, in, sizeof(w));
    s.bp = unwire_bb<G>(&w.bplayer); s.bo = unwire_bb<G>(&w.bopponent); s.lp = unwire_bb<G>(&w.legalplay); s.player = w.player; s.aux = 0;
  } else {
    WirePos2 w; memcpy(&w, in, sizeof(w));
    s.bp = unwire_bb<G>(&w.bplayer); s.bo = unwire_bb<G>(&w.bopponent); s.player = w.player; s.aux = w.aux;
  }
  return s;
}

}  // namespace ag
