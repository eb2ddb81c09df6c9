// oracle/oracle.cpp — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
//
// A line-by-line CPU restatement of the AlphaGPU self-play hot path
// (Bitboard.jl, 4IARow.jl, Gobang.jl, Hex.jl, Reversi6x6.jl, Reversi8x8.jl,
// mcts_gpu.jl, DenseNet.jl:294-304, main4IARow.jl:29-77).  Every function cites
// the reference file:line it follows.  Nothing in the product path
// (alphagpu_b200/, libalphagpu.so) may include, link or call this file; only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs do, and there only as the checker / CPU baseline.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures
//  and Julia is not installed here, so this restatement cannot
// be checked against reference outputs.  It is pinned instead against
// (i) an independent big-integer Python restatement (oracle/pyref.py),
// (ii) naive rule implementations (tests/naive_rules.py) and (iii) the derived
// known-answer vectors of SURVEY.md Appendix B (tests/golden/).
//
// Float semantics: IEEE-754 binary32, round-to-nearest-even, the operations
// and their order exactly as written in the Julia source, no FMA contraction,
// no reassociation (compile with -ffp-contract=off, never -ffast-math).  The
// reference itself is launched with --math-mode=fast (README.md:23) so its own
// last-ulp behaviour is not reproducible from source; this is the canonical
// reading of the source text.
//
// Third-party arithmetic the reference pulls from un-vendored packages
// (Manifest.toml pins) and how it is restated here:
//   CUDA.jl 3.3.5 CUDA.rand (mcts_gpu.jl:397)  -> Philox4x32-10 keyed by
//       (seed, game uid, ply, rollout, depth), mapped to (0,1] like CURAND;
//       or an injected `prob` tensor.
//   cuBLAS SGEMM via `*` (DenseNet.jl:295-301) -> fp32 dot products, k ascending.
//   NNlib 0.7.27 softmax! (mcts_gpu.jl:417)    -> exp(x-max)/sum, ascending sum.
//   StatsBase 0.33.9 sample(.., Weights) (mcts_gpu.jl:520,606) -> inverse CDF
//       on t = u*sum(w), `while cw < t && i < n` (StatsBase sampling.jl), u from Philox.

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

// ----------------------------------------------------------------------------
// Bitboard.jl
// ----------------------------------------------------------------------------
typedef uint64_t u64;

// Bitboard.jl:5-9 — immutable 3x UInt64 + len + dims (Julia isbits layout, 48 B)
struct bitboard {
  u64 chunks[3];
  int64_t len;
  int64_t dims[2];
};
static_assert(sizeof(bitboard) == 48, "Julia bitboard{2} layout");

// Julia shifts saturate: a UInt64 shifted by >= 64 is 0 (Bitboard.jl:100-101,126-127 rely on it)
static inline u64 jshl(u64 x, int64_t n) { return (n >= 64) ? 0 : (n < 0 ? 0 : x << n); }
static inline u64 jshr(u64 x, int64_t n) { return (n >= 64) ? 0 : (n < 0 ? 0 : x >> n); }

static const u64 _msk64 = ~u64(0);                                     // Bitboard.jl:27
static inline int64_t _div64(int64_t l) { return l >> 6; }             // :28
static inline int64_t _mod64(int64_t l) { return l & 63; }             // :29
static inline u64 _msk_end(int64_t l) { return jshr(_msk64, _mod64(-l)); }  // :31

// Bitboard.jl:33-41
static inline void _msk(const bitboard& bb, u64 m[3]) {
  if (bb.len <= 64) { m[0] = _msk_end(bb.len); m[1] = 0; m[2] = 0; }
  else if (bb.len <= 128) { m[0] = _msk64; m[1] = _msk_end(bb.len); m[2] = 0; }
  else { m[0] = _msk64; m[1] = _msk64; m[2] = _msk_end(bb.len); }
}

// Bitboard.jl:14-25
static inline bitboard bb_new(int64_t d1, int64_t d2) {
  bitboard b; b.chunks[0] = b.chunks[1] = b.chunks[2] = 0; b.len = d1 * d2; b.dims[0] = d1; b.dims[1] = d2;
  return b;
}

// Bitboard.jl:45-52 — 1-based linear index
static inline bool bb_get(const bitboard& bb, int64_t i) {
  int64_t i1 = _div64(i - 1), i2 = _mod64(i - 1);
  return (bb.chunks[i1] & (u64(1) << i2)) != 0;
}
// Bitboard.jl:54-57 — [row, col], column-major
static inline bool bb_get2(const bitboard& bb, int64_t r, int64_t c) { return bb_get(bb, bb.dims[0] * (c - 1) + r); }

// Bitboard.jl:60-74 — functional setindex
static inline bitboard bb_set(const bitboard& bb, bool x, int64_t i) {
  int64_t i1 = _div64(i - 1), i2 = _mod64(i - 1);
  u64 u = u64(1) << i2;
  bitboard r = bb;
  r.chunks[i1] = x ? (bb.chunks[i1] | u) : (bb.chunks[i1] & ~u);
  return r;
}
// Bitboard.jl:76-79
static inline bitboard bb_set2(const bitboard& bb, bool x, int64_t r, int64_t c) { return bb_set(bb, x, bb.dims[0] * (c - 1) + r); }

// Bitboard.jl:85-107
static inline bitboard bb_shl(const bitboard& bb, int64_t n) {
  int64_t i1 = _div64(n), i2 = _mod64(n);
  u64 x = bb.chunks[0], y = bb.chunks[1], z = bb.chunks[2];
  if (1 <= i1 && i1 < 2) { z = y; y = x; x = 0; }
  else if (i1 >= 2) { z = x; x = 0; y = x; }
  u64 newx = jshl(x, n);
  u64 headx = jshr(x, 64 - i2);
  u64 heady = jshr(y, 64 - i2);
  u64 newy = jshl(y, i2) | headx;
  u64 newz = jshl(z, i2) | heady;
  u64 m[3]; _msk(bb, m);
  bitboard r = bb; r.chunks[0] = newx & m[0]; r.chunks[1] = newy & m[1]; r.chunks[2] = newz & m[2];
  return r;
}

// Bitboard.jl:110-133 — the 1<=i1<2 branch is restated as written (y=z; z=0; x=y),
// unreachable for every shipped game (max shift = dims[1] <= 14).
static inline bitboard bb_shr(const bitboard& bb, int64_t n) {
  int64_t i1 = _div64(n), i2 = _mod64(n);
  u64 x = bb.chunks[0], y = bb.chunks[1], z = bb.chunks[2];
  if (1 <= i1 && i1 < 2) { y = z; z = 0; x = y; }
  else if (i1 >= 2) { x = z; z = 0; y = z; }
  u64 newz = jshr(z, n);
  u64 headz = jshl(z, 64 - i2);
  u64 heady = jshl(y, 64 - i2);
  u64 newy = jshr(y, i2) | headz;
  u64 newx = jshr(x, i2) | heady;
  u64 m[3]; _msk(bb, m);
  bitboard r = bb; r.chunks[0] = newx & m[0]; r.chunks[1] = newy & m[1]; r.chunks[2] = newz & m[2];
  return r;
}

static inline bitboard bb_right(const bitboard& bb) { return bb_shl(bb, bb.dims[0]); }  // :135-138
static inline bitboard bb_left(const bitboard& bb) { return bb_shr(bb, bb.dims[0]); }   // :141-144

// Bitboard.jl:146-160 — shift by one row, clear the bit at every column start
static inline bitboard bb_down(const bitboard& bb) {
  bitboard d = bb_shl(bb, 1);
  for (int64_t i = 1; i <= bb.len; i += bb.dims[0]) {
    int64_t i1 = _div64(i - 1), i2 = _mod64(i - 1);
    d.chunks[i1] &= ~(u64(1) << i2);
  }
  return d;
}
// Bitboard.jl:162-176 — shift back one row, clear the bit at every column end
static inline bitboard bb_up(const bitboard& bb) {
  bitboard d = bb_shr(bb, 1);
  for (int64_t i = bb.dims[0]; i <= bb.len; i += bb.dims[0]) {
    int64_t i1 = _div64(i - 1), i2 = _mod64(i - 1);
    d.chunks[i1] &= ~(u64(1) << i2);
  }
  return d;
}
// Bitboard.jl:177-180
static inline int num_bit(const bitboard& bb) {
  return __builtin_popcountll(bb.chunks[0]) + __builtin_popcountll(bb.chunks[1]) + __builtin_popcountll(bb.chunks[2]);
}
// Bitboard.jl:182-187 (masked)
static inline bitboard bb_not(const bitboard& bb) {
  u64 m[3]; _msk(bb, m);
  bitboard r = bb; for (int k = 0; k < 3; k++) r.chunks[k] = (~bb.chunks[k]) & m[k];
  return r;
}
// Bitboard.jl:189-205 (unmasked)
static inline bitboard bb_and(const bitboard& a, const bitboard& b) { bitboard r = a; for (int k = 0; k < 3; k++) r.chunks[k] = a.chunks[k] & b.chunks[k]; return r; }
static inline bitboard bb_or(const bitboard& a, const bitboard& b) { bitboard r = a; for (int k = 0; k < 3; k++) r.chunks[k] = a.chunks[k] | b.chunks[k]; return r; }
static inline bitboard bb_xor(const bitboard& a, const bitboard& b) { bitboard r = a; for (int k = 0; k < 3; k++) r.chunks[k] = a.chunks[k] ^ b.chunks[k]; return r; }

// ----------------------------------------------------------------------------
// Game plugins.  One runtime Spec stands for the module constants
// (Main.N / NN / Nvict, VectorizedState, FeatureSize, maxActions, maxLengthGame).
// ----------------------------------------------------------------------------
enum { G_CONNECT4 = 0, G_GOBANG = 1, G_HEX = 2, G_REVERSI8 = 3, G_REVERSI6 = 4 };

struct Spec {
  int game, N, NN, Nvict;
  int A, VS, FS, maxLen;
  int pos_bytes;   // Julia isbits size of Position: 104 (two bitboards) or 152 (three)
};

static bool make_spec(int game, int N, int Nvict, Spec* s) {
  s->game = game; s->N = N; s->NN = N * N; s->Nvict = Nvict;
  switch (game) {
    case G_CONNECT4:  // 4IARow.jl:6-12
      s->N = 0; s->NN = 0; s->Nvict = 4; s->VS = 42; s->FS = 42; s->A = 7; s->maxLen = 42; s->pos_bytes = 104; return true;
    case G_GOBANG:    // Gobang.jl:8-11
      if (N < 1 || N * N > 192 || Nvict < 2) return false;
      s->VS = s->NN; s->FS = s->NN; s->A = s->NN; s->maxLen = s->NN; s->pos_bytes = 104; return true;
    case G_HEX:       // Hex.jl:8-11
      if (N < 2 || (N + 1) * (N + 1) > 192) return false;
      s->Nvict = 0; s->VS = (N + 1) * (N + 1); s->FS = s->VS; s->A = s->NN; s->maxLen = s->NN; s->pos_bytes = 104; return true;
    case G_REVERSI8:  // Reversi8x8.jl:5-8
      s->N = 8; s->NN = 64; s->Nvict = 0; s->FS = 64; s->VS = 64; s->A = 65; s->maxLen = 70; s->pos_bytes = 152; return true;
    case G_REVERSI6:  // Reversi6x6.jl:6-9
      s->N = 6; s->NN = 36; s->Nvict = 0; s->FS = 36; s->VS = 36; s->A = 37; s->maxLen = 50; s->pos_bytes = 152; return true;
  }
  return false;
}

// Internal position: superset of both Julia structs.
struct Pos {
  bitboard bplayer, bopponent, legalplay;
  int8_t player;
  int8_t aux;   // round (4IARow.jl:20, Gobang.jl:20) or lp (Hex.jl:20); unused for Reversi
};

// Julia isbits layouts
struct WirePos2 { bitboard bplayer, bopponent; int8_t player; int8_t aux; int8_t pad[6]; };
struct WirePos3 { bitboard bplayer, bopponent, legalplay; int8_t player; int8_t pad[7]; };
static_assert(sizeof(WirePos2) == 104, "Position (Connect4/Gobang/Hex) = 104 B");
static_assert(sizeof(WirePos3) == 152, "Position (Reversi) = 152 B");

static void from_wire(const Spec& s, const void* w, Pos* p) {
  if (s.pos_bytes == 104) { const WirePos2* q = (const WirePos2*)w; p->bplayer = q->bplayer; p->bopponent = q->bopponent; p->legalplay = bb_new(0, 0); p->player = q->player; p->aux = q->aux; }
  else { const WirePos3* q = (const WirePos3*)w; p->bplayer = q->bplayer; p->bopponent = q->bopponent; p->legalplay = q->legalplay; p->player = q->player; p->aux = 0; }
}
static void to_wire(const Spec& s, const Pos& p, void* w) {
  memset(w, 0, s.pos_bytes);
  if (s.pos_bytes == 104) { WirePos2* q = (WirePos2*)w; q->bplayer = p.bplayer; q->bopponent = p.bopponent; q->player = p.player; q->aux = p.aux; }
  else { WirePos3* q = (WirePos3*)w; q->bplayer = p.bplayer; q->bopponent = p.bopponent; q->legalplay = p.legalplay; q->player = p.player; }
}

// ---- Reversi helpers (Reversi8x8.jl:17-71; Reversi6x6.jl identical but for sizes) ----
typedef bitboard (*DirFn)(const bitboard&);
static bitboard d_up(const bitboard& x) { return bb_up(x); }
static bitboard d_down(const bitboard& x) { return bb_down(x); }
static bitboard d_left(const bitboard& x) { return bb_left(x); }
static bitboard d_right(const bitboard& x) { return bb_right(x); }
static bitboard diaghd(const bitboard& x) { return bb_up(bb_right(x)); }    // :17
static bitboard diaghg(const bitboard& x) { return bb_up(bb_left(x)); }     // :19
static bitboard diagbd(const bitboard& x) { return bb_down(bb_right(x)); }  // :21
static bitboard diagbg(const bitboard& x) { return bb_down(bb_left(x)); }   // :23

// Reversi8x8.jl:26-35
static bitboard rv_legal_play(const bitboard& tabjoueur, const bitboard& tabadversaire, DirFn dir) {
  bitboard tabvide = bb_and(bb_not(tabjoueur), bb_not(tabadversaire));
  bitboard moves = bb_new(tabjoueur.dims[0], tabjoueur.dims[1]);
  bitboard candidats = bb_and(dir(tabjoueur), tabadversaire);
  while (num_bit(candidats) != 0) {
    moves = bb_or(moves, bb_and(tabvide, dir(candidats)));
    candidats = bb_and(tabadversaire, dir(candidats));
  }
  return moves;
}
// Reversi8x8.jl:37-40 (order of the OR chain as written)
static bitboard rv_legalplay(const bitboard& j, const bitboard& a) {
  bitboard m = rv_legal_play(j, a, d_up);
  m = bb_or(m, rv_legal_play(j, a, d_down));
  m = bb_or(m, rv_legal_play(j, a, d_left));
  m = bb_or(m, rv_legal_play(j, a, d_right));
  m = bb_or(m, rv_legal_play(j, a, diaghg));
  m = bb_or(m, rv_legal_play(j, a, diagbg));
  m = bb_or(m, rv_legal_play(j, a, diaghd));
  m = bb_or(m, rv_legal_play(j, a, diagbd));
  return m;
}
// Reversi8x8.jl:44-56
static bitboard rv_flippar(const bitboard& tabjoueur, const bitboard& tabadversaire, const bitboard& play, DirFn dir) {
  bitboard candidats = bb_and(dir(play), tabadversaire);
  bitboard toflip = candidats;
  while (num_bit(candidats) != 0) {
    candidats = bb_and(tabadversaire, dir(candidats));
    toflip = bb_or(toflip, candidats);
  }
  if (num_bit(bb_and(dir(toflip), tabjoueur)) != 0) return toflip;
  return bb_new(tabjoueur.dims[0], tabjoueur.dims[1]);
}
// Reversi8x8.jl:58-70
static bitboard rv_flip(const bitboard& j, const bitboard& a, int64_t play) {
  bitboard test = bb_set(bb_new(j.dims[0], j.dims[1]), true, play);
  bitboard h = rv_flippar(j, a, test, d_up);
  h = bb_or(h, rv_flippar(j, a, test, d_down));
  h = bb_or(h, rv_flippar(j, a, test, d_left));
  h = bb_or(h, rv_flippar(j, a, test, d_right));
  h = bb_or(h, rv_flippar(j, a, test, diaghd));
  h = bb_or(h, rv_flippar(j, a, test, diaghg));
  h = bb_or(h, rv_flippar(j, a, test, diagbd));
  h = bb_or(h, rv_flippar(j, a, test, diagbg));
  return h;
}

// ---- Position() ----
static Pos pos_init(const Spec& s) {
  Pos p; p.legalplay = bb_new(0, 0); p.aux = 0;
  switch (s.game) {
    case G_CONNECT4:  // 4IARow.jl:23  Position(bitboard(6,7), bitboard(6,7), 1, 1)
      p.bplayer = bb_new(6, 7); p.bopponent = bb_new(6, 7); p.player = 1; p.aux = 1; break;
    case G_GOBANG:    // Gobang.jl:23  (N,N), player 1, round 0
      p.bplayer = bb_new(s.N, s.N); p.bopponent = bb_new(s.N, s.N); p.player = 1; p.aux = 0; break;
    case G_HEX: {     // Hex.jl:22-35  borders [3..N+1,1] for x and [1,3..N+1] for o, lp = NN
      bitboard startx = bb_new(s.N + 1, s.N + 1), starto = bb_new(s.N + 1, s.N + 1);
      for (int i = 3; i <= s.N + 1; i++) { startx = bb_set2(startx, true, i, 1); starto = bb_set2(starto, true, 1, i); }
      p.bplayer = startx; p.bopponent = starto; p.player = 1; p.aux = (int8_t)s.NN; break;
    }
    case G_REVERSI8: {  // Reversi8x8.jl:10-14,80-82
      bitboard empty = bb_new(8, 8);
      bitboard start = bb_set2(empty, true, 4, 5); bitboard starto = bb_set2(start, true, 5, 4);
      start = bb_set2(empty, true, 5, 5); bitboard startp = bb_set2(start, true, 4, 4);
      p.bplayer = starto; p.bopponent = startp; p.legalplay = rv_legalplay(starto, startp); p.player = 1; break;
    }
    case G_REVERSI6: {  // Reversi6x6.jl:10-14
      bitboard empty = bb_new(6, 6);
      bitboard start = bb_set2(empty, true, 4, 3); bitboard starto = bb_set2(start, true, 3, 4);
      start = bb_set2(empty, true, 3, 3); bitboard startp = bb_set2(start, true, 4, 4);
      p.bplayer = starto; p.bopponent = startp; p.legalplay = rv_legalplay(starto, startp); p.player = 1; break;
    }
  }
  return p;
}

// ---- canPlay ----
static bool can_play(const Spec& s, const Pos& pos, int col) {
  switch (s.game) {
    case G_CONNECT4:  // 4IARow.jl:25-27  top cell of the column free
      return !bb_get2(pos.bplayer, 1, col) & !bb_get2(pos.bopponent, 1, col);
    case G_GOBANG:    // Gobang.jl:25-27
      return !bb_get(pos.bplayer, col) & !bb_get(pos.bopponent, col);
    case G_HEX: {     // Hex.jl:37-42
      int x = (col - 1) / s.N, y = col - s.N * x, newcol = (s.N + 1) * (x + 1) + y + 1;
      return !bb_get(pos.bplayer, newcol) & !bb_get(pos.bopponent, newcol);
    }
    case G_REVERSI8: case G_REVERSI6:  // Reversi8x8.jl:84-90 — last action = pass, legal only when stuck
      if (col == s.A) return num_bit(pos.legalplay) == 0;
      return bb_get(pos.legalplay, col);
  }
  return false;
}

// ---- play ----
static Pos play(const Spec& s, const Pos& pos, int col) {
  Pos r; r.legalplay = pos.legalplay; r.aux = 0;
  switch (s.game) {
    case G_CONNECT4: {  // 4IARow.jl:30-44  gravity: deepest free row reached from the top
      int free_ = 1;
      bitboard empty = bb_not(bb_or(pos.bplayer, pos.bopponent));
      for (int i = 1; i <= 6; i++) { if (bb_get2(empty, i, col)) free_ = i; else break; }
      int c = 6 * (col - 1) + free_;
      bitboard bplayer = bb_set(pos.bplayer, true, c);
      r.bplayer = pos.bopponent; r.bopponent = bplayer; r.player = (int8_t)(-pos.player); r.aux = (int8_t)(pos.aux + 1);
      return r;
    }
    case G_GOBANG: {    // Gobang.jl:30-33
      bitboard bplayer = bb_set(pos.bplayer, true, col);
      r.bplayer = pos.bopponent; r.bopponent = bplayer; r.player = (int8_t)(-pos.player); r.aux = (int8_t)(pos.aux + 1);
      return r;
    }
    case G_HEX: {       // Hex.jl:45-51
      int x = (col - 1) / s.N, y = col - s.N * x, newcol = (s.N + 1) * (x + 1) + y + 1;
      bitboard bplayer = bb_set(pos.bplayer, true, newcol);
      r.bplayer = pos.bopponent; r.bopponent = bplayer; r.player = (int8_t)(-pos.player); r.aux = (int8_t)(pos.aux - 1);
      return r;
    }
    case G_REVERSI8: case G_REVERSI6: {  // Reversi8x8.jl:93-106
      bitboard tabjoueur = pos.bplayer, tabadversaire = pos.bopponent;
      if (col == s.A) {
        r.bplayer = pos.bopponent; r.bopponent = pos.bplayer; r.legalplay = rv_legalplay(tabadversaire, tabjoueur); r.player = (int8_t)(-pos.player);
        return r;
      }
      bitboard h = rv_flip(tabjoueur, tabadversaire, col);
      tabjoueur = bb_xor(tabjoueur, h);
      tabadversaire = bb_xor(tabadversaire, h);
      tabjoueur = bb_set(tabjoueur, true, col);
      r.bplayer = tabadversaire; r.bopponent = tabjoueur; r.legalplay = rv_legalplay(tabadversaire, tabjoueur); r.player = (int8_t)(-pos.player);
      return r;
    }
  }
  return pos;
}

// 4IARow.jl:47-81 / Gobang.jl:36-70 — k-in-a-row on the previous mover's stones
static bool row_test(const Spec& s, const Pos& pos) {
  int nv = s.Nvict;
  bitboard board = pos.bopponent;
  for (int j = 1; j <= nv - 1; j++) board = bb_and(board, bb_right(board));
  if (num_bit(board) != 0) return true;
  board = pos.bopponent;
  for (int j = 1; j <= nv - 1; j++) board = bb_and(board, bb_down(board));
  if (num_bit(board) != 0) return true;
  board = pos.bopponent;
  for (int j = 1; j <= nv - 1; j++) board = bb_and(board, bb_down(bb_right(board)));
  if (num_bit(board) != 0) return true;
  board = pos.bopponent;
  for (int j = 1; j <= nv - 1; j++) board = bb_and(board, bb_left(bb_down(board)));
  if (num_bit(board) != 0) return true;
  return false;
}

// ---- isOver -> (over, winner in {+1,-1,0}) ----
static bool is_over(const Spec& s, const Pos& pos, int8_t* res) {
  switch (s.game) {
    case G_CONNECT4:  // 4IARow.jl:47-81
      if (row_test(s, pos)) { *res = (int8_t)(-pos.player); return true; }
      *res = 0; return num_bit(pos.bplayer) + num_bit(pos.bopponent) == s.maxLen;
    case G_GOBANG:    // Gobang.jl:36-70
      if (row_test(s, pos)) { *res = (int8_t)(-pos.player); return true; }
      *res = 0; return num_bit(pos.bplayer) + num_bit(pos.bopponent) == s.NN;
    case G_HEX: {     // Hex.jl:54-67  majority-of-3 Y reduction, border re-injected when player==1
      bitboard a = pos.bopponent;
      for (int j = 1; j <= 2 * s.N - 2; j++) {
        bitboard b = bb_up(a);
        bitboard c = bb_right(b);
        a = bb_down(bb_or(bb_and(a, bb_or(b, c)), bb_and(b, c)));
        if (pos.player == 1) for (int k = 3 + j; k <= s.N + 1; k++) a = bb_set2(a, true, 1, k);
      }
      *res = (int8_t)(-pos.player);
      return bb_get2(a, s.N + 1, s.N + 1);
    }
    case G_REVERSI8: {  // Reversi8x8.jl:109-131
      int8_t test = (int8_t)(num_bit(pos.bplayer) - num_bit(pos.bopponent));
      int sg = (test > 0) - (test < 0);
      *res = (int8_t)(sg * pos.player);
      return num_bit(pos.legalplay) == 0 && num_bit(rv_legalplay(pos.bopponent, pos.bplayer)) == 0;
    }
    case G_REVERSI6: {  // Reversi6x6.jl:109-130 (branchy form)
      if (num_bit(pos.legalplay) != 0 || num_bit(rv_legalplay(pos.bopponent, pos.bplayer)) != 0) { *res = 0; return false; }
      int8_t test = (int8_t)(num_bit(pos.bplayer) - num_bit(pos.bopponent));
      if (test > 0) *res = pos.player; else if (test == 0) *res = 0; else *res = (int8_t)(-pos.player);
      return true;
    }
  }
  *res = 0; return false;
}

// mcts_gpu.jl:202-223 (decoder) / :225-246 (decoder_roots): [bplayer bits 1..VS ; bopponent bits 1..VS]
static void encode(const Spec& s, const Pos& pos, float* out) {
  for (int j = 1; j <= s.VS; j++) {
    out[j - 1] = bb_get(pos.bplayer, j) ? 1.f : 0.f;
    out[j - 1 + s.VS] = bb_get(pos.bopponent, j) ? 1.f : 0.f;
  }
}
// mcts_gpu.jl:464-474 (decode): fstate[j] = bplayer[j] ? player : -player
static void decode_fstate(const Spec& s, const Pos& pos, int8_t* fstate) {
  for (int j = 1; j <= s.VS; j++) fstate[j - 1] = bb_get(pos.bplayer, j) ? pos.player : (int8_t)(-pos.player);
}

// ----------------------------------------------------------------------------
// RNG: Philox4x32-10 (Salmon et al. 2011), standing in for CUDA.rand (mcts_gpu.jl:397)
// and the host RNG behind StatsBase.sample (mcts_gpu.jl:520,606).
// key = (seed lo, seed hi); counter = (game uid, ply, rollout, depth/4); word depth%4.
// ----------------------------------------------------------------------------
static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
// CURAND's curand_uniform mapping: x * 2^-32 + 2^-33, in (0, 1]
static inline float u01(uint32_t x) { return (float)x * 2.3283064365386963e-10f + 1.1641532182693481e-10f; }

static const uint32_t ROLLOUT_MOVE = 0xFFFFFFFFu;  // counter word 2 for the per-ply move draw

static inline float rng_uniform(uint64_t seed, uint32_t uid, uint32_t ply, uint32_t rollout, uint32_t depth) {
  uint32_t c[4] = {uid, ply, rollout, depth >> 2};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  return u01(c[depth & 3]);
}

// ----------------------------------------------------------------------------
// snetwork2 forward (DenseNet.jl:294-304), weights as convert_back hands them
// over (DenseNet.jl:331-333): Julia column-major fp32.
// ----------------------------------------------------------------------------
struct Net {
  int in, n, k, A;
  std::vector<float> base;               // n x in
  std::vector<std::vector<float>> res;   // k times n x n
  std::vector<float> policy, policy_bias, value;  // A x n, A, 1 x n
  float value_bias;
  // bf16-rounded copies of the matrices (modes 1, 2)
  std::vector<float> base_r, policy_r, value_r;
  std::vector<std::vector<float>> res_r;
  // fp16-rounded copies (mode 3)
  std::vector<float> base_h, policy_h, value_h;
  std::vector<std::vector<float>> res_h;
};

static inline float bf16_round(float x) {  // round-to-nearest-even to bfloat16, back to fp32
  uint32_t u; memcpy(&u, &x, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return x;
  u += 0x7FFFu + ((u >> 16) & 1u); u &= 0xFFFF0000u;
  float r; memcpy(&r, &u, 4); return r;
}
// round-to-nearest-even to IEEE binary16, saturating to +-65504 (PTX cvt.rn.satfinite.f16.f32), back to fp32
static inline float f16_round(float x) {
  uint32_t u; memcpy(&u, &x, 4);
  uint32_t sign = u & 0x80000000u, a = u & 0x7FFFFFFFu;
  float r;
  if (a >= 0x477FF000u) { r = 65504.0f; }                       // >= 65520: saturate
  else if (a < 0x33000001u) { r = 0.0f; }                       // <= 2^-25: rounds to zero
  else {
    int e = (int)(a >> 23) - 127;
    int drop = e < -14 ? 13 + (-14 - e) : 13;                   // mantissa bits that do not fit (more for subnormal halves)
    uint32_t m = (a & 0x7FFFFFu) | 0x800000u;
    uint32_t q = m >> drop, rem = m & ((1u << drop) - 1u), half = 1u << (drop - 1);
    if (rem > half || (rem == half && (q & 1u))) q++;
    r = ldexpf((float)q, e - 23 + drop);
  }
  uint32_t ru; memcpy(&ru, &r, 4); ru |= sign; memcpy(&r, &ru, 4);
  return r;
}
static inline float relu(float x) { return x > 0.f ? x : 0.f; }

// exp as SPECIFIED for this path (DESIGN.md "canonical exp"): the reference takes exp from the platform
// (NNlib softmax!/σ -> CUDA libdevice / Julia Base), which is not reproducible across CPU and GPU, so the
// oracle and the kernels both evaluate this fixed sequence of exactly-rounded fp32 operations:
// Cody–Waite reduction x = k ln2 + r, degree-7 Taylor/Horner for e^r, scale by 2^k.  <= 2 ulp from exp.
static inline float c_expf(float x) {
  if (x < -87.0f) return 0.0f;
  if (x > 88.0f) x = 88.0f;
  const float log2e = 1.44269504088896341f, ln2_hi = 0.693145751953125f, ln2_lo = 1.42860682030941723212e-6f;
  float t = x * log2e;
  float kf = (t + 12582912.0f) - 12582912.0f;     // nearest integer (ties to even), exact for |t| < 2^22
  float r = (x - kf * ln2_hi) - kf * ln2_lo;
  const float c[8] = {1.9841269841e-4f, 1.3888888889e-3f, 8.3333333333e-3f, 4.1666666667e-2f, 1.6666666667e-1f, 0.5f, 1.0f, 1.0f};
  float p = c[0];
  for (int i = 1; i < 8; i++) p = p * r + c[i];    // compiled with -ffp-contract=off: mul then add, two roundings
  int k = (int)kf;
  uint32_t bits = (uint32_t)(k + 127) << 23;
  float scale; memcpy(&scale, &bits, 4);
  return p * scale;
}
// NNlib sigmoid, numerically stable form
static inline float sigmoidf(float x) { float t = c_expf(-fabsf(x)); return x >= 0.f ? 1.f / (1.f + t) : t / (1.f + t); }

// mode 0: fp32 as the reference.  mode 1: "bf16-faithful" — weights and every MMA
// operand rounded to bf16, fp32 accumulate, fp32 residual stream (what the
// tcgen05 chain computes, up to accumulation order).
// mode 2: as mode 1 but the residual stream itself is stored in bf16 (wide nets).
// mode 3: as mode 1 with fp16 operands (saturating) instead of bf16.
// mode 4: as mode 3 with the residual stream stored in fp16 (wide nets).
static void net_forward_one(const Net& net, const float* x, float* logits, float* v, int mode, float* b, float* t, float* op) {
  const int n = net.n;
  const bool f16m = (mode == 3 || mode == 4);
  const std::vector<float>& Wbase = f16m ? net.base_h : mode ? net.base_r : net.base;
  const std::vector<float>& Wpol = f16m ? net.policy_h : mode ? net.policy_r : net.policy;
  const std::vector<float>& Wval = f16m ? net.value_h : mode ? net.value_r : net.value;
  auto rnd = [mode, f16m](float v) { return f16m ? f16_round(v) : mode ? bf16_round(v) : v; };
  // b = relu.(base * x)        DenseNet.jl:295
  for (int o = 0; o < n; o++) t[o] = 0.f;
  for (int i = 0; i < net.in; i++) {
    float xi = x[i];
    if (xi == 0.f) continue;  // inputs are 0/1: adding 0*w is exact, skipping is bit-identical
    for (int o = 0; o < n; o++) t[o] += Wbase[o + n * i] * xi;
  }
  for (int o = 0; o < n; o++) b[o] = relu(t[o]);
  if (mode == 2) for (int o = 0; o < n; o++) b[o] = bf16_round(b[o]);
  if (mode == 4) for (int o = 0; o < n; o++) b[o] = f16_round(b[o]);
  // for w in res: b .= relu.(b .+ relu.(w*b))   DenseNet.jl:297-299
  for (int l = 0; l < net.k; l++) {
    const std::vector<float>& w = f16m ? net.res_h[l] : mode ? net.res_r[l] : net.res[l];
    for (int o = 0; o < n; o++) { t[o] = 0.f; op[o] = rnd(b[o]); }
    for (int i = 0; i < n; i++) {
      float bi = op[i];
      for (int o = 0; o < n; o++) t[o] += w[o + n * i] * bi;
    }
    for (int o = 0; o < n; o++) b[o] = relu(b[o] + relu(t[o]));
    if (mode == 2) for (int o = 0; o < n; o++) b[o] = bf16_round(b[o]);
    if (mode == 4) for (int o = 0; o < n; o++) b[o] = f16_round(b[o]);
  }
  // policy*b .+ policy_bias , σ.(value*b .+ value_bias)     DenseNet.jl:301
  for (int o = 0; o < n; o++) op[o] = rnd(b[o]);
  for (int a = 0; a < net.A; a++) {
    float acc = 0.f;
    for (int i = 0; i < n; i++) acc += Wpol[a + net.A * i] * op[i];
    logits[a] = acc + net.policy_bias[a];
  }
  float acc = 0.f;
  for (int i = 0; i < n; i++) acc += Wval[i] * op[i];
  *v = sigmoidf(acc + net.value_bias);
}

// softmax! over the action axis (mcts_gpu.jl:417; NNlib: exp.(x .- max) ./ sum)
static void softmax_inplace(float* x, int A) {
  float m = x[0];
  for (int a = 1; a < A; a++) m = std::max(m, x[a]);
  float ssum = 0.f;
  for (int a = 0; a < A; a++) { x[a] = c_expf(x[a] - m); ssum += x[a]; }
  for (int a = 0; a < A; a++) x[a] = x[a] / ssum;
}

// ----------------------------------------------------------------------------
// Tree storage (mcts_gpu.jl:35-53) for L games x `visits` nodes.  Indices are
// kept 1-based in the stored values (0 = none) as in the reference.
// ----------------------------------------------------------------------------
struct Tree {
  Spec s;
  int R;        // `visits`: node capacity per game
  int64_t L;    // capacity
  // vnodesStats  [(a, node, game)] -> idx3
  std::vector<float> prior, policy, q, visits;
  std::vector<int32_t> Achild;      // (A, R, L)
  std::vector<int32_t> childID;     // (R, R, L)
  std::vector<int32_t> childnbr;    // (R, L)
  std::vector<float> policy_final;  // (A, L)
  std::vector<float> batch;         // (2VS, L)
  // vnodes
  std::vector<int32_t> parent, actionFromParent;  // (R, L)
  std::vector<Pos> state;                          // (R, L)
  std::vector<int8_t> expanded, uptodate;          // (R, L)
  std::vector<int32_t> leaf, newindex;             // (L)
  std::vector<uint32_t> uid;                       // (L) global game id (RNG key)
  // scratch for evaluation
  std::vector<float> nn_prior;  // (A, L)
  std::vector<float> nn_v;      // (L)
  // counters (reported by the bench: mean descent depth d̄, Newton iterations)
  int64_t cnt_descents, cnt_nodes_traversed, cnt_newton_solves, cnt_newton_iters;
  int64_t max_newton_iters, max_depth, hist_iters[101];

  inline size_t i3(int a, int node, int64_t g) const { return (size_t)a + (size_t)s.A * ((size_t)node + (size_t)R * (size_t)g); }
  inline size_t i2(int node, int64_t g) const { return (size_t)node + (size_t)R * (size_t)g; }
  inline size_t ic(int slot, int node, int64_t g) const { return (size_t)slot + (size_t)R * ((size_t)node + (size_t)R * (size_t)g); }
};

static Tree* tree_create(const Spec& s, int R, int64_t L) {
  Tree* t = new Tree();
  t->s = s; t->R = R; t->L = L;
  size_t arl = (size_t)s.A * R * L, rl = (size_t)R * L;
  t->prior.assign(arl, 0.f); t->policy.assign(arl, 0.f); t->q.assign(arl, 0.f); t->visits.assign(arl, 0.f);
  t->Achild.assign(arl, 0); t->childID.assign((size_t)R * R * L, 0); t->childnbr.assign(rl, 0);
  t->policy_final.assign((size_t)s.A * L, 0.f); t->batch.assign((size_t)2 * s.VS * L, 0.f);
  t->parent.assign(rl, 0); t->actionFromParent.assign(rl, 0);
  Pos p0 = pos_init(s);
  t->state.assign(rl, p0);
  t->expanded.assign(rl, 0); t->uptodate.assign(rl, 1);
  t->leaf.assign(L, 0); t->newindex.assign(L, 1);
  t->uid.resize(L); for (int64_t g = 0; g < L; g++) t->uid[g] = (uint32_t)g;
  t->nn_prior.assign((size_t)s.A * L, 0.f); t->nn_v.assign(L, 0.f);
  t->cnt_descents = t->cnt_nodes_traversed = t->cnt_newton_solves = t->cnt_newton_iters = 0;
  t->max_newton_iters = t->max_depth = 0; memset(t->hist_iters, 0, sizeof(t->hist_iters));
  return t;
}

// re_init (mcts_gpu.jl:359-373): install roots, expanded .= 0, uptodate .= 1
static void tree_reinit(Tree* t, const Pos* positions, const uint32_t* uids, int64_t L) {
  for (int64_t g = 0; g < L; g++) {
    t->state[t->i2(0, g)] = positions[g];
    if (uids) t->uid[g] = uids[g];
  }
  std::fill(t->expanded.begin(), t->expanded.end(), 0);
  std::fill(t->uptodate.begin(), t->uptodate.end(), 1);
}

// the eight fills at the top of mcts_single (mcts_gpu.jl:380-387)
static void search_begin(Tree* t, int64_t L) {
  const Spec& s = t->s;
  size_t arl = (size_t)s.A * t->R * L, rl = (size_t)t->R * L;
  std::fill(t->q.begin(), t->q.begin() + arl, 0.f);
  std::fill(t->Achild.begin(), t->Achild.begin() + arl, 0);
  std::fill(t->childID.begin(), t->childID.begin() + (size_t)t->R * t->R * L, 0);
  std::fill(t->visits.begin(), t->visits.begin() + arl, 0.f);
  std::fill(t->prior.begin(), t->prior.begin() + arl, 0.f);
  std::fill(t->policy.begin(), t->policy.begin() + arl, 0.f);
  std::fill(t->childnbr.begin(), t->childnbr.begin() + rl, 0);
  std::fill(t->newindex.begin(), t->newindex.begin() + L, 1);
}

// kdescendTree! (mcts_gpu.jl:100-199) for one game i.  `prob` gives the uniform for depth cpt (1-based).
template <class ProbFn>
static void descend_one(Tree* t, int64_t i, float cpuct, ProbFn prob, int64_t* n_trav, int64_t* n_solve, int64_t* n_iter) {
  const int A = t->s.A;
  int nindex = 1;   // 1-based node id
  int cpt = 1;
  while (t->expanded[t->i2(nindex - 1, i)] == 1) {
    (*n_trav)++;
    int bestmove = -1;
    float pr = 0.f;
    float* prior = &t->prior[t->i3(0, nindex - 1, i)];
    float* q = &t->q[t->i3(0, nindex - 1, i)];
    float* vis = &t->visits[t->i3(0, nindex - 1, i)];
    float* policy = &t->policy[t->i3(0, nindex - 1, i)];
    int32_t* Achild = &t->Achild[t->i3(0, nindex - 1, i)];
    if (t->uptodate[t->i2(nindex - 1, i)] != 1) {                    // :114
      float Acount = 0.f, n = 1.f, prior_rem = 0.f;
      int childnbr = t->childnbr[t->i2(nindex - 1, i)];
      for (int k = 0; k < A; k++) {                                   // :120-131
        n += vis[k];
        if (Achild[k] == 0) prior_rem += prior[k];
        if (prior[k] > 0.f) Acount += 1.f;
      }
      float lambda = cpuct * sqrtf(n) / (Acount + n);                 // :132
      float alpha = 0.f;
      prior_rem *= lambda;                                            // :134
      for (int k = 0; k < A; k++) {                                   // :135-138
        float gap = std::max(lambda * prior[k], 1e-4f);
        alpha = std::max(alpha, q[k] + gap);
      }
      float err = INFINITY, newerr = INFINITY;
      (*n_solve)++;
      int iters_here = 0;
      for (int j = 1; j <= 100; j++) {                                // :141-162
        (*n_iter)++; iters_here++;
        float S = prior_rem / alpha;
        float g = -prior_rem / (alpha * alpha);
        for (int k = 0; k < childnbr; k++) {
          int CID = t->childID[t->ic(k, nindex - 1, i)];
          int action = t->actionFromParent[t->i2(CID - 1, i)];
          float top = lambda * prior[action - 1];
          float bot = (alpha - q[action - 1]);
          S += top / bot;
          g += -top / (bot * bot);
        }
        newerr = S - 1.f;
        if (newerr < 0.001f || newerr == err) break;
        alpha -= newerr / g;
        err = newerr;
      }
      if (iters_here > t->max_newton_iters) t->max_newton_iters = iters_here;   // (racy under OpenMP; diagnostics only)
      t->hist_iters[iters_here]++;
      for (int k = 0; k < A; k++) policy[k] = lambda * prior[k] / (alpha - q[k]);   // :165-169
    }
    float u = prob(cpt);
    for (int k = 1; k <= A; k++) {                                    // :172-182
      float delta = policy[k - 1];
      pr += delta;
      if (delta > 0.f) bestmove = k;
      if (pr >= u) break;
    }
    if (bestmove < 1) bestmove = 1;  // reference would index [-1] (UB under --check-bounds=no); never reached with a positive policy
    if (Achild[bestmove - 1] == 0) {                                  // :183-191
      t->newindex[i] += 1;
      int ni = t->newindex[i];
      t->childnbr[t->i2(nindex - 1, i)] += 1;
      int cn = t->childnbr[t->i2(nindex - 1, i)];
      t->childID[t->ic(cn - 1, nindex - 1, i)] = ni;
      Achild[bestmove - 1] = cn;
      t->parent[t->i2(ni - 1, i)] = nindex;
      t->actionFromParent[t->i2(ni - 1, i)] = bestmove;
      t->state[t->i2(ni - 1, i)] = play(t->s, t->state[t->i2(nindex - 1, i)], bestmove);
    }
    nindex = t->childID[t->ic(Achild[bestmove - 1] - 1, nindex - 1, i)];   // :192
    cpt += 1;
  }
  if (cpt - 1 > t->max_depth) t->max_depth = cpt - 1;
  t->leaf[i] = nindex;                                                // :195
}

// expand (mcts_gpu.jl:250-302) for one game; `prior_in` = softmaxed network output (A)
static void expand_one(Tree* t, int64_t i, const float* prior_in, bool training) {
  const int A = t->s.A;
  int nindex = t->leaf[i];
  const Pos& st = t->state[t->i2(nindex - 1, i)];
  int8_t r; bool f = is_over(t->s, st, &r);
  t->expanded[t->i2(nindex - 1, i)] = (int8_t)(1 - (f ? 1 : 0));     // :256
  float* prior = &t->prior[t->i3(0, nindex - 1, i)];
  float* policy = &t->policy[t->i3(0, nindex - 1, i)];
  if (!f) {
    if (nindex == 1) {                                                // :259-281
      float normalize = 0.f, Acount = 0.f;
      for (int j = 1; j <= A; j++) if (can_play(t->s, st, j)) { prior[j - 1] = prior_in[j - 1]; normalize += prior[j - 1]; Acount += 1.f; }
      if (training) {
        for (int j = 1; j <= A; j++) if (can_play(t->s, st, j)) prior[j - 1] = 0.75f * prior[j - 1] / normalize + 0.25f / Acount;   // :273
      } else {
        for (int j = 1; j <= A; j++) prior[j - 1] /= normalize;
      }
    } else {                                                          // :283-295
      float normalize = 0.f;
      for (int j = 1; j <= A; j++) if (can_play(t->s, st, j)) { prior[j - 1] = prior_in[j - 1]; normalize += prior[j - 1]; }
      for (int j = 1; j <= A; j++) prior[j - 1] /= normalize;
    }
  }
  for (int k = 0; k < A; k++) policy[k] = prior[k];                   // :297-299
}

// backUp (mcts_gpu.jl:306-328) for one game.  The terminal branch produces a Float64
// (`(1+player*r)/2`, :314) so the running mean on that path is evaluated in double
// and rounded on store, exactly as the Julia promotion rules give.
static void backup_one(Tree* t, int64_t i, float v) {
  int leaf = t->leaf[i];
  int nindex = t->parent[t->i2(leaf - 1, i)];
  int move = t->actionFromParent[t->i2(leaf - 1, i)];
  const Pos& st = t->state[t->i2(leaf - 1, i)];
  int8_t r; bool f = is_over(t->s, st, &r);
  if (f) {
    double value = (double)(1 + (int)(int8_t)(st.player * r)) / 2.0;
    while (nindex != 0) {
      float* q = &t->q[t->i3(move - 1, nindex - 1, i)];
      float* vis = &t->visits[t->i3(move - 1, nindex - 1, i)];
      float prod = (*vis) * (*q);                       // Float32 * Float32
      float den = (*vis) + 1.f;                         // Float32 + Int
      *q = (float)(((double)prod + (1.0 - value)) / (double)den);
      *vis += 1.f;
      t->uptodate[t->i2(nindex - 1, i)] = 0;
      move = t->actionFromParent[t->i2(nindex - 1, i)];
      nindex = t->parent[t->i2(nindex - 1, i)];
      value = 1.0 - value;
    }
  } else {
    float value = v;
    while (nindex != 0) {
      float* q = &t->q[t->i3(move - 1, nindex - 1, i)];
      float* vis = &t->visits[t->i3(move - 1, nindex - 1, i)];
      *q = ((*vis) * (*q) + (1.f - value)) / ((*vis) + 1.f);          // :319
      *vis += 1.f;                                                     // :320
      t->uptodate[t->i2(nindex - 1, i)] = 0;                           // :321
      move = t->actionFromParent[t->i2(nindex - 1, i)];
      nindex = t->parent[t->i2(nindex - 1, i)];
      value = 1.f - value;                                             // :324
    }
  }
}

// One rollout's descent + decoder over all L games (mcts_gpu.jl:397-407).
// prob (optional): injected uniforms laid out [(rollout*L + i)*maxLen + depth].
static void select_all(Tree* t, int64_t L, int rollout, float cpuct, const float* prob, uint64_t seed, uint32_t ply) {
  const int maxLen = t->s.maxLen;
  int64_t trav = 0, solve = 0, iters = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : trav, solve, iters)
  for (int64_t i = 0; i < L; i++) {
    if (prob) {
      const float* p = prob + ((size_t)rollout * L + i) * maxLen;
      descend_one(t, i, cpuct, [&](int cpt) { return p[cpt - 1]; }, &trav, &solve, &iters);
    } else {
      uint32_t uid = t->uid[i];
      descend_one(t, i, cpuct, [&](int cpt) { return rng_uniform(seed, uid, ply, (uint32_t)rollout, (uint32_t)(cpt - 1)); }, &trav, &solve, &iters);
    }
    encode(t->s, t->state[t->i2(t->leaf[i] - 1, i)], &t->batch[(size_t)2 * t->s.VS * i]);   // decoder :202-223
  }
  t->cnt_descents += L; t->cnt_nodes_traversed += trav; t->cnt_newton_solves += solve; t->cnt_newton_iters += iters;
}

// actor + softmax! over all L leaves (mcts_gpu.jl:414-417)
static void eval_all(Tree* t, const Net& net, int64_t L, int mode, float* logits_out) {
  const int A = t->s.A, in = 2 * t->s.VS;
#pragma omp parallel
  {
    std::vector<float> b(net.n), tt(net.n), op(net.n);
#pragma omp for schedule(static)
    for (int64_t i = 0; i < L; i++) {
      float* pr = &t->nn_prior[(size_t)A * i];
      net_forward_one(net, &t->batch[(size_t)in * i], pr, &t->nn_v[i], mode, b.data(), tt.data(), op.data());
      if (logits_out) memcpy(logits_out + (size_t)A * i, pr, sizeof(float) * A);
      softmax_inplace(pr, A);
    }
  }
}

// expand + backUp over all games (mcts_gpu.jl:424-431)
static void expand_backup_all(Tree* t, int64_t L, const float* prior, const float* v, bool training) {
  const int A = t->s.A;
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < L; i++) {
    expand_one(t, i, prior + (size_t)A * i, training);
    backup_one(t, i, v[i]);
  }
}

// tail of mcts_single (mcts_gpu.jl:441-443): decoder_roots + copy_pol
static void finish_search(Tree* t, int64_t L) {
  const int A = t->s.A;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < L; i++) {
    encode(t->s, t->state[t->i2(0, i)], &t->batch[(size_t)2 * t->s.VS * i]);
    for (int k = 0; k < A; k++) t->policy_final[(size_t)A * i + k] = t->policy[t->i3(k, 0, i)];
  }
}

// mcts_single (mcts_gpu.jl:376-462).  Evaluator: `net` (mode fp32 / bf16-faithful) or, when
// inj_prior/inj_v are given, injected softmaxed priors [(rollout*L+i)*A + a] and values [rollout*L+i].
static void mcts_single(Tree* t, const Net* net, int visits, int64_t L, bool training, float cpuct,
                        const float* prob, const float* inj_prior, const float* inj_v,
                        uint64_t seed, uint32_t ply, int nn_mode) {
  search_begin(t, L);
  const int A = t->s.A;
  for (int k = 0; k < visits; k++) {
    select_all(t, L, k, cpuct, prob, seed, ply);
    const float* pr; const float* vv;
    if (inj_prior) { pr = inj_prior + (size_t)k * L * A; vv = inj_v + (size_t)k * L; }
    else { eval_all(t, *net, L, nn_mode, nullptr); pr = t->nn_prior.data(); vv = t->nn_v.data(); }
    expand_backup_all(t, L, pr, vv, training);
  }
  finish_search(t, L);
}

// Move choice of the self-play loop (mcts_gpu.jl:518-524): ply<25 -> StatsBase sample over the
// non-zero entries weighted by pol, else argmax (first maximal index).  Returns 1-based action.
static int choose_move_selfplay(const float* pol, int A, uint32_t round, float u) {
  if (round < 25) {
    float wsum = 0.f; int n = 0; int last = 1;
    for (int c = 0; c < A; c++) if (pol[c] != 0.f) { wsum += pol[c]; n++; last = c + 1; }
    if (n == 0) return 1;
    float tt = u * wsum;
    // StatsBase.sample(rng, wv): i=1; cw=wv[1]; while cw < t && i < n: i+=1; cw+=wv[i]
    float cw = 0.f; bool first = true; int chosen = last;
    for (int c = 0; c < A; c++) {
      if (pol[c] == 0.f) continue;
      if (first) { cw = pol[c]; first = false; } else cw += pol[c];
      if (!(cw < tt) || c + 1 == last) { chosen = c + 1; break; }
    }
    return chosen;
  }
  int best = 0;
  for (int c = 1; c < A; c++) if (pol[c] > pol[best]) best = c;
  return best + 1;
}
// Duel variant (mcts_gpu.jl:605-609): ply<15 -> sample(1:maxActions, Weights(policy)) over ALL entries
static int choose_move_duel(const float* pol, int A, uint32_t round, float u) {
  if (round < 15) {
    float wsum = 0.f;
    for (int c = 0; c < A; c++) wsum += pol[c];
    float tt = u * wsum;
    int i = 1; float cw = pol[0];
    while (cw < tt && i < A) { i += 1; cw += pol[i - 1]; }
    return i;
  }
  int best = 0;
  for (int c = 1; c < A; c++) if (pol[c] > pol[best]) best = c;
  return best + 1;
}

}  // namespace orc

// ============================================================================
// C entry points (ctypes).  Prefix orc_.
// ============================================================================
using namespace orc;

extern "C" {

struct orc_game_info_t { int32_t A, VS, FS, maxLen, pos_bytes; };

int orc_game_info(int game, int N, int Nvict, orc_game_info_t* out) {
  Spec s; if (!make_spec(game, N, Nvict, &s)) return -1;
  out->A = s.A; out->VS = s.VS; out->FS = s.FS; out->maxLen = s.maxLen; out->pos_bytes = s.pos_bytes; return 0;
}

int orc_position_init(int game, int N, int Nvict, void* pos_out) {
  Spec s; if (!make_spec(game, N, Nvict, &s)) return -1;
  Pos p = pos_init(s); to_wire(s, p, pos_out); return 0;
}

// batch plugin ops over n wire positions
int orc_can_play(int game, int N, int Nvict, const void* pos, const int32_t* action, int64_t n, uint8_t* out) {
  Spec s; if (!make_spec(game, N, Nvict, &s)) return -1;
  for (int64_t i = 0; i < n; i++) { Pos p; from_wire(s, (const char*)pos + i * s.pos_bytes, &p); out[i] = can_play(s, p, action[i]) ? 1 : 0; }
  return 0;
}
int orc_legal(int game, int N, int Nvict, const void* pos, int64_t n, uint8_t* out /* n x A */) {
  Spec s; if (!make_spec(game, N, Nvict, &s)) return -1;
  for (int64_t i = 0; i < n; i++) { Pos p; from_wire(s, (const char*)pos + i * s.pos_bytes, &p); for (int a = 1; a <= s.A; a++) out[i * s.A + a - 1] = can_play(s, p, a) ? 1 : 0; }
  return 0;
}
int orc_play(int game, int N, int Nvict, const void* pos, const int32_t* action, int64_t n, void* pos_out) {
  Spec s; if (!make_spec(game, N, Nvict, &s)) return -1;
  for (int64_t i = 0; i < n; i++) { Pos p; from_wire(s, (const char*)pos + i * s.pos_bytes, &p); Pos q = play(s, p, action[i]); to_wire(s, q, (char*)pos_out + i * s.pos_bytes); }
  return 0;
}
int orc_is_over(int game, int N, int Nvict, const void* pos, int64_t n, uint8_t* over, int8_t* result) {
  Spec s; if (!make_spec(game, N, Nvict, &s)) return -1;
  for (int64_t i = 0; i < n; i++) { Pos p; from_wire(s, (const char*)pos + i * s.pos_bytes, &p); int8_t r; over[i] = is_over(s, p, &r) ? 1 : 0; result[i] = r; }
  return 0;
}
int orc_encode(int game, int N, int Nvict, const void* pos, int64_t n, float* out /* n x 2VS */) {
  Spec s; if (!make_spec(game, N, Nvict, &s)) return -1;
  for (int64_t i = 0; i < n; i++) { Pos p; from_wire(s, (const char*)pos + i * s.pos_bytes, &p); encode(s, p, out + (size_t)i * 2 * s.VS); }
  return 0;
}
int orc_decode_fstate(int game, int N, int Nvict, const void* pos, int64_t n, int8_t* out /* n x VS */) {
  Spec s; if (!make_spec(game, N, Nvict, &s)) return -1;
  for (int64_t i = 0; i < n; i++) { Pos p; from_wire(s, (const char*)pos + i * s.pos_bytes, &p); decode_fstate(s, p, out + (size_t)i * s.VS); }
  return 0;
}

// raw bitboard ops for unit tests: op 0 '<<' 1 '>>>' 2 right 3 left 4 down 5 up 6 '~'
int orc_bb_op(int op, const void* bb_in, int64_t n, void* bb_out) {
  bitboard b = *(const bitboard*)bb_in, r;
  switch (op) {
    case 0: r = bb_shl(b, n); break; case 1: r = bb_shr(b, n); break; case 2: r = bb_right(b); break; case 3: r = bb_left(b); break;
    case 4: r = bb_down(b); break; case 5: r = bb_up(b); break; case 6: r = bb_not(b); break; default: return -1;
  }
  *(bitboard*)bb_out = r; return 0;
}

float orc_uniform(uint64_t seed, uint32_t uid, uint32_t ply, uint32_t rollout, uint32_t depth) { return rng_uniform(seed, uid, ply, rollout, depth); }
void orc_philox(uint32_t* ctr4, uint32_t k0, uint32_t k1) { philox4x32_10(ctr4, k0, k1); }

// ---- network ----
void* orc_net_create(int in, int n, int k, int A, const float* base, const float* const* res,
                     const float* pol_w, const float* pol_b, const float* val_w, const float* val_b) {
  Net* net = new Net(); net->in = in; net->n = n; net->k = k; net->A = A;
  net->base.assign(base, base + (size_t)n * in);
  for (int l = 0; l < k; l++) net->res.emplace_back(res[l], res[l] + (size_t)n * n);
  net->policy.assign(pol_w, pol_w + (size_t)A * n); net->policy_bias.assign(pol_b, pol_b + A);
  net->value.assign(val_w, val_w + n); net->value_bias = val_b[0];
  auto rr = [](const std::vector<float>& w) { std::vector<float> r(w.size()); for (size_t i = 0; i < w.size(); i++) r[i] = bf16_round(w[i]); return r; };
  net->base_r = rr(net->base); net->policy_r = rr(net->policy); net->value_r = rr(net->value);
  for (int l = 0; l < k; l++) net->res_r.push_back(rr(net->res[l]));
  auto rh = [](const std::vector<float>& w) { std::vector<float> r(w.size()); for (size_t i = 0; i < w.size(); i++) r[i] = f16_round(w[i]); return r; };
  net->base_h = rh(net->base); net->policy_h = rh(net->policy); net->value_h = rh(net->value);
  for (int l = 0; l < k; l++) net->res_h.push_back(rh(net->res[l]));
  return net;
}
void orc_net_destroy(void* net) { delete (Net*)net; }
// x: (in, L) column-major as the reference's `batch`; logits (A, L); v (L). softmax optional.
int orc_net_forward(void* netp, const float* x, int64_t L, float* logits, float* v, int mode, int apply_softmax) {
  const Net& net = *(Net*)netp;
#pragma omp parallel
  {
    std::vector<float> b(net.n), t(net.n), op(net.n);
#pragma omp for schedule(static)
    for (int64_t i = 0; i < L; i++) {
      net_forward_one(net, x + (size_t)net.in * i, logits + (size_t)net.A * i, v + i, mode, b.data(), t.data(), op.data());
      if (apply_softmax) softmax_inplace(logits + (size_t)net.A * i, net.A);
    }
  }
  return 0;
}
void orc_expf(const float* x, int64_t n, float* y) { for (int64_t i = 0; i < n; i++) y[i] = c_expf(x[i]); }
void orc_sigmoid(const float* x, int64_t n, float* y) { for (int64_t i = 0; i < n; i++) y[i] = sigmoidf(x[i]); }
void orc_round(const float* x, int64_t n, float* y, int fmt) { for (int64_t i = 0; i < n; i++) y[i] = fmt == 1 ? f16_round(x[i]) : bf16_round(x[i]); }
void orc_softmax(float* x, int A, int64_t L) { for (int64_t i = 0; i < L; i++) softmax_inplace(x + (size_t)A * i, A); }

// ---- tree / search ----
void* orc_tree_create(int game, int N, int Nvict, int R, int64_t L) {
  Spec s; if (!make_spec(game, N, Nvict, &s)) return nullptr;
  return tree_create(s, R, L);
}
void orc_tree_destroy(void* t) { delete (Tree*)t; }
int orc_tree_reinit(void* tp, const void* positions, const uint32_t* uids, int64_t L) {
  Tree* t = (Tree*)tp; if (L > t->L) return -1;
  std::vector<Pos> ps(L);
  for (int64_t i = 0; i < L; i++) from_wire(t->s, (const char*)positions + i * t->s.pos_bytes, &ps[i]);
  tree_reinit(t, ps.data(), uids, L); return 0;
}
int orc_search_begin(void* tp, int64_t L) { search_begin((Tree*)tp, L); return 0; }
int orc_select(void* tp, int64_t L, int rollout, float cpuct, const float* prob, uint64_t seed, uint32_t ply) {
  select_all((Tree*)tp, L, rollout, cpuct, prob, seed, ply); return 0;
}
// leaf ids (1-based) and the decoder output (2VS, L)
int orc_get_leaf_batch(void* tp, int64_t L, int32_t* leaf, float* batch) {
  Tree* t = (Tree*)tp;
  if (leaf) memcpy(leaf, t->leaf.data(), sizeof(int32_t) * L);
  if (batch) memcpy(batch, t->batch.data(), sizeof(float) * 2 * t->s.VS * L);
  return 0;
}
int orc_eval(void* tp, void* net, int64_t L, int mode, float* logits_out) { eval_all((Tree*)tp, *(Net*)net, L, mode, logits_out); return 0; }
int orc_get_eval(void* tp, int64_t L, float* prior, float* v) {
  Tree* t = (Tree*)tp; memcpy(prior, t->nn_prior.data(), sizeof(float) * t->s.A * L); memcpy(v, t->nn_v.data(), sizeof(float) * L); return 0;
}
// prior/v == NULL -> use the result of orc_eval
int orc_expand_backup(void* tp, int64_t L, const float* prior, const float* v, int training) {
  Tree* t = (Tree*)tp;
  expand_backup_all(t, L, prior ? prior : t->nn_prior.data(), v ? v : t->nn_v.data(), training != 0); return 0;
}
int orc_finish_search(void* tp, int64_t L) { finish_search((Tree*)tp, L); return 0; }
int orc_mcts_single(void* tp, void* net, int visits, int64_t L, int training, float cpuct,
                    const float* prob, const float* inj_prior, const float* inj_v, uint64_t seed, uint32_t ply, int nn_mode) {
  Tree* t = (Tree*)tp; if (visits > t->R || L > t->L) return -1;
  mcts_single(t, (Net*)net, visits, L, training != 0, cpuct, prob, inj_prior, inj_v, seed, ply, nn_mode); return 0;
}
int orc_get_roots(void* tp, int64_t L, float* policy_final, float* batch) {
  Tree* t = (Tree*)tp;
  if (policy_final) memcpy(policy_final, t->policy_final.data(), sizeof(float) * t->s.A * L);
  if (batch) memcpy(batch, t->batch.data(), sizeof(float) * 2 * t->s.VS * L);
  return 0;
}

// Tree dump in a layout-neutral form, all [game][node](...)[action], ids 1-based, 0 = none.
//   child[g][node][a] = node id reached by action a+1 (Achild/childID composed), order[g][node][slot] = action of the slot-th created child
int orc_tree_dump(void* tp, int64_t L, int32_t* nnodes, int32_t* parent, int32_t* action, int32_t* child, int32_t* order,
                  int32_t* nchild, int8_t* expanded, float* prior, float* q, float* visits, float* policy, void* states) {
  Tree* t = (Tree*)tp; const int A = t->s.A, R = t->R;
  for (int64_t g = 0; g < L; g++) {
    if (nnodes) nnodes[g] = t->newindex[g];
    for (int nd = 0; nd < R; nd++) {
      size_t o = (size_t)g * R + nd;
      bool live = nd < t->newindex[g];
      if (parent) parent[o] = live ? t->parent[t->i2(nd, g)] : 0;
      if (action) action[o] = live ? t->actionFromParent[t->i2(nd, g)] : 0;
      if (nchild) nchild[o] = live ? t->childnbr[t->i2(nd, g)] : 0;
      if (expanded) expanded[o] = live ? t->expanded[t->i2(nd, g)] : 0;
      if (states) { if (live) to_wire(t->s, t->state[t->i2(nd, g)], (char*)states + o * t->s.pos_bytes); else memset((char*)states + o * t->s.pos_bytes, 0, t->s.pos_bytes); }
      for (int a = 0; a < A; a++) {
        size_t oa = o * A + a;
        int slot = live ? t->Achild[t->i3(a, nd, g)] : 0;
        if (child) child[oa] = slot ? t->childID[t->ic(slot - 1, nd, g)] : 0;
        if (order) { int cn = live ? t->childnbr[t->i2(nd, g)] : 0; order[oa] = (a < cn && a < R) ? t->actionFromParent[t->i2(t->childID[t->ic(a, nd, g)] - 1, g)] : 0; }
        if (prior) prior[oa] = live ? t->prior[t->i3(a, nd, g)] : 0.f;
        if (q) q[oa] = live ? t->q[t->i3(a, nd, g)] : 0.f;
        if (visits) visits[oa] = live ? t->visits[t->i3(a, nd, g)] : 0.f;
        if (policy) policy[oa] = live ? t->policy[t->i3(a, nd, g)] : 0.f;
      }
    }
  }
  return 0;
}
// Test hook: overwrite the statistics of node `node` (1-based) of game g and create its children in
// the given action order, as if those descents/backups had happened (used by the α-solve KAT).
int orc_tree_poke(void* tp, int64_t g, int node, const float* prior, const float* q, const float* visits,
                  const int32_t* child_order, int nchild) {
  Tree* t = (Tree*)tp; const int A = t->s.A;
  for (int a = 0; a < A; a++) {
    t->prior[t->i3(a, node - 1, g)] = prior[a]; t->q[t->i3(a, node - 1, g)] = q[a]; t->visits[t->i3(a, node - 1, g)] = visits[a];
    t->policy[t->i3(a, node - 1, g)] = prior[a];
  }
  for (int k = 0; k < nchild; k++) {
    int a = child_order[k];
    t->newindex[g] += 1; int ni = t->newindex[g];
    t->childnbr[t->i2(node - 1, g)] += 1; int cn = t->childnbr[t->i2(node - 1, g)];
    t->childID[t->ic(cn - 1, node - 1, g)] = ni; t->Achild[t->i3(a - 1, node - 1, g)] = cn;
    t->parent[t->i2(ni - 1, g)] = node; t->actionFromParent[t->i2(ni - 1, g)] = a;
    t->state[t->i2(ni - 1, g)] = play(t->s, t->state[t->i2(node - 1, g)], a);
  }
  t->expanded[t->i2(node - 1, g)] = 1; t->uptodate[t->i2(node - 1, g)] = 0;
  return 0;
}
int orc_get_extremes(void* tp, int64_t* out /* [2 + 101] */) {
  Tree* t = (Tree*)tp; out[0] = t->max_newton_iters; out[1] = t->max_depth; for (int i = 0; i <= 100; i++) out[2 + i] = t->hist_iters[i]; return 0;
}
int orc_get_counters(void* tp, int64_t* out4) {
  Tree* t = (Tree*)tp; out4[0] = t->cnt_descents; out4[1] = t->cnt_nodes_traversed; out4[2] = t->cnt_newton_solves; out4[3] = t->cnt_newton_iters; return 0;
}

// ---- self-play loop (mcts_gpu.jl:477-579) and samples (main4IARow.jl:29-77) ----
// Samples are emitted in push order (ply-major, live games in order), SoA:
//   state int8 (2VS), policy f32 (A), player i8, value f32, fstate i8 (FS), game uid i32, ply i32.
struct orc_samples_t {
  int64_t capacity;   // in: rows available
  int64_t count;      // out
  int8_t* state; float* policy; int8_t* player; float* value; int8_t* fstate; int32_t* game; int32_t* ply;
};
// stats: [0] sims (sum L_ply*R) [1] positions (sum L_ply) [2] plies [3] total game length (sum of `round` at end) [4] faults
int orc_selfplay(int game, int N, int Nvict, void* netp, int visits, int64_t ngames, uint32_t uid_base, float cpuct, uint64_t seed,
                 int nn_mode, orc_samples_t* out, int64_t results[3], int64_t stats[5]) {
  Spec s; if (!make_spec(game, N, Nvict, &s)) return -1;
  const Net* net = (const Net*)netp;
  Tree* t = tree_create(s, visits, ngames);
  std::vector<Pos> positions(ngames, pos_init(s));                          // :479
  std::vector<uint32_t> uids(ngames);
  for (int64_t g = 0; g < ngames; g++) uids[g] = uid_base + (uint32_t)g;
  std::vector<std::vector<int64_t>> rtemp(ngames);                           // :482
  tree_reinit(t, positions.data(), uids.data(), ngames);
  uint32_t round = 0; int64_t v = 0, n = 0, d = 0, L = ngames, tot_length = 0, sims = 0, npos = 0, faults = 0;
  int64_t count = 0;
  std::vector<int8_t> fstate(s.VS);
  while (!positions.empty()) {
    mcts_single(t, net, visits, L, true, cpuct, nullptr, nullptr, nullptr, seed, round, nn_mode);   // :503
    sims += L * visits; npos += L;
    std::vector<int64_t> finished;
    for (int64_t i = 0; i < (int64_t)positions.size(); i++) {               // :513-549
      int64_t index = count++;                                              // push_buffer (main4IARow.jl:49-63), ring wrap not modelled
      if (out && index < out->capacity) {
        for (int j = 0; j < 2 * s.VS; j++) out->state[index * 2 * s.VS + j] = (int8_t)t->batch[(size_t)2 * s.VS * i + j];
        for (int a = 0; a < s.A; a++) out->policy[index * s.A + a] = t->policy_final[(size_t)s.A * i + a];
        out->player[index] = positions[i].player;
        out->game[index] = (int32_t)uids[i]; out->ply[index] = (int32_t)round;
      }
      rtemp[i].push_back(index);
      const float* pol = &t->policy_final[(size_t)s.A * i];
      float u = rng_uniform(seed, uids[i], round, ROLLOUT_MOVE, 0);
      int c = choose_move_selfplay(pol, s.A, round, u);
      if (!can_play(s, positions[i], c)) { faults++; }                       // "faute" :526-529 (counted; the loop stops after this ply)
      positions[i] = play(s, positions[i], c);
      int8_t res; bool f = is_over(s, positions[i], &res);
      if (f) {
        decode_fstate(s, positions[i], fstate.data());
        finished.push_back(i);
        tot_length += round;
        for (int64_t id : rtemp[i]) {                                       // update_buffer (main4IARow.jl:65-75)
          if (out && id < out->capacity) {
            int8_t player = out->player[id];
            out->value[id] = (float)((1 + res * player) / 2.0);
            for (int j = 0; j < s.FS; j++) out->fstate[id * s.FS + j] = (int8_t)(fstate[j] * player);
          }
        }
        if (res == 1) v++; else if (res == 0) n++; else d++;
      }
    }
    for (size_t k = 0; k < finished.size(); k++) {                          // :550-553 order-preserving deleteat!
      int64_t c = finished[k] - (int64_t)k;
      rtemp.erase(rtemp.begin() + c); positions.erase(positions.begin() + c); uids.erase(uids.begin() + c);
    }
    round++;
    L = (int64_t)positions.size();
    // "faute" (:526-529): the reference returns valid=false at the first illegal move.  Restated per ply (the ply that produced it is
    // completed, then the generation stops); the ply cap guards against positions an illegal move leaves unchanged.
    if (faults > 0 || round > (uint32_t)s.maxLen + 8) break;
    if (L > 0) tree_reinit(t, positions.data(), uids.data(), L);             // :557-561
  }
  if (out) out->count = count;
  results[0] = v; results[1] = n; results[2] = d;
  if (stats) { stats[0] = sims; stats[1] = npos; stats[2] = round; stats[3] = tot_length; stats[4] = faults; }
  delete t;
  return 0;
}

// duel (mcts_gpu.jl:581-651): actor alternates by ply parity, training=false
int orc_duel(int game, int N, int Nvict, void* net1, void* net2, int visits, int64_t ngames, uint32_t uid_base, float cpuct, uint64_t seed,
             int nn_mode, int64_t results[3], int64_t stats[5]) {
  Spec s; if (!make_spec(game, N, Nvict, &s)) return -1;
  Tree* t = tree_create(s, visits, ngames);
  std::vector<Pos> positions(ngames, pos_init(s));
  std::vector<uint32_t> uids(ngames);
  for (int64_t g = 0; g < ngames; g++) uids[g] = uid_base + (uint32_t)g;
  tree_reinit(t, positions.data(), uids.data(), ngames);
  uint32_t round = 0; int64_t v = 0, n = 0, d = 0, L = ngames, sims = 0, npos = 0, faults = 0;
  while (!positions.empty()) {
    const Net* actor = (round % 2 == 0) ? (const Net*)net1 : (const Net*)net2;   // :592-596
    mcts_single(t, actor, visits, L, false, cpuct, nullptr, nullptr, nullptr, seed, round, nn_mode);
    sims += L * visits; npos += L;
    std::vector<int64_t> finished;
    for (int64_t i = 0; i < (int64_t)positions.size(); i++) {
      const float* pol = &t->policy_final[(size_t)s.A * i];
      float u = rng_uniform(seed, uids[i], round, ROLLOUT_MOVE, 0);
      int c = choose_move_duel(pol, s.A, round, u);
      if (!can_play(s, positions[i], c)) { faults++; }
      positions[i] = play(s, positions[i], c);
      int8_t res; bool f = is_over(s, positions[i], &res);
      if (f) { finished.push_back(i); if (res == 1) v++; else if (res == 0) n++; else d++; }
    }
    for (size_t k = 0; k < finished.size(); k++) { int64_t c = finished[k] - (int64_t)k; positions.erase(positions.begin() + c); uids.erase(uids.begin() + c); }
    round++;
    L = (int64_t)positions.size();
    if (faults > 0 || round > (uint32_t)s.maxLen + 8) break;                 // "faute" (:611-614)
    if (L > 0) tree_reinit(t, positions.data(), uids.data(), L);
  }
  results[0] = v; results[1] = n; results[2] = d;
  if (stats) { stats[0] = sims; stats[1] = npos; stats[2] = round; stats[3] = 0; stats[4] = faults; }
  delete t;
  return 0;
}

int orc_choose_move(const float* pol, int A, uint32_t round, float u, int duel) { return duel ? choose_move_duel(pol, A, round, u) : choose_move_selfplay(pol, A, round, u); }

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

}  // extern "C"
