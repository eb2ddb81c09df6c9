/* alphagpu.h — C ABI of libalphagpu.so, the B200-native self-play MCTS engine.
 *
 * Drop-in boundary for AlphaGPU's batched self-play search.  The reference has no FFI of its
 * own: its seam is the Julia-level API of `module mcts_gpu` (mcts_gpu.jl).  Each entry point
 * below names the reference interface it replaces (file:line in fabricerosay/AlphaGPU); the
 * Julia `ccall` stubs and the Python ctypes twin that bind them are shown in INTEGRATION.md.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; no C++/torch types cross the boundary.
 *  - every call returns 0 (AGPU_OK) or a negative agpu_status; nothing throws across the ABI;
 *    the message is available from agpu_last_error().
 *  - all pointer arguments are HOST pointers unless the name ends in _dev; the library owns all
 *    device memory.  Calls are synchronous on return (the reference synchronises after every
 *    phase, mcts_gpu.jl:398-444).
 *  - a context is bound to one CUDA device and is not thread-safe.
 *  - actions, node ids and plies cross the boundary 1-based where the reference is 1-based
 *    (actions 1..maxActions, node ids 1..rollouts, 0 = none).
 *  - "Position" buffers use the Julia isbits layout of the reference structs:
 *      bitboard{2}  = { uint64 chunks[3]; int64 len; int64 dims[2]; }           48 B  (Bitboard.jl:5-9)
 *      Position     = { bitboard bplayer, bopponent; int8 player; int8 round|lp; }  -> 104 B
 *                      (4IARow.jl:16-21, Gobang.jl:16-21, Hex.jl:16-21)
 *      Position     = { bitboard bplayer, bopponent, legalplay; int8 player; }  -> 152 B
 *                      (Reversi8x8.jl:73-78, Reversi6x6.jl)
 *  - batched float outputs are game-major: policy_final[L][A], batch[L][2*VS] — the same bytes as
 *    the reference's column-major (A, L) / (2VS, L) CuArrays (mcts_gpu.jl:38).
 */
#ifndef ALPHAGPU_H
#define ALPHAGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGPU_ABI_VERSION 1

typedef enum agpu_status {
  AGPU_OK = 0,
  AGPU_ERR_INVALID = -1,     /* bad argument / unsupported configuration */
  AGPU_ERR_CUDA = -2,        /* CUDA runtime error (message in agpu_last_error) */
  AGPU_ERR_NO_DEVICE = -3,   /* no usable CUDA device: there is no CPU fallback */
  AGPU_ERR_STATE = -4,       /* call out of order (e.g. search before reinit / weights) */
  AGPU_ERR_ILLEGAL_MOVE = -5 /* the reference's "faute" (mcts_gpu.jl:526-529) */
} agpu_status;

/* game plugins (module names in the reference) */
typedef enum agpu_game {
  AGPU_CONNECT4 = 0, /* FourIARow, 4IARow.jl */
  AGPU_GOBANG = 1,   /* GoBang, Gobang.jl — needs n (board side, Main.N) and nvict (Main.Nvict) */
  AGPU_HEX = 2,      /* Hex, Hex.jl — needs n (Main.N) */
  AGPU_REVERSI8 = 3, /* RevSix in Reversi8x8.jl */
  AGPU_REVERSI6 = 4  /* RevSix in Reversi6x6.jl */
} agpu_game;

typedef enum agpu_nn_mode {
  AGPU_NN_BF16_TC = 0,  /* tcgen05/TMEM GEMM chain, bf16 operands, fp32 accumulate and residual stream */
  AGPU_NN_FP32 = 1,     /* fp32 CUDA-core chain, evaluation order of DenseNet.jl:294-304 (bit-exact parity mode) */
  AGPU_NN_FP16_TC = 2   /* same tcgen05 chain with fp16 operands (3 more mantissa bits than bf16, same MMA rate):
                           the mode that meets the 1e-3 policy/value tolerance against the fp32 formula */
} agpu_nn_mode;

typedef struct agpu_config {
  int32_t game;       /* agpu_game */
  int32_t n;          /* Main.N for Gobang / Hex, else 0 */
  int32_t nvict;      /* Main.Nvict for Gobang, else 0 */
  int32_t rollouts;   /* `visits`: node capacity per game tree (mcts_gpu.jl:342-357), 1..255 */
  int64_t max_games;  /* capacity L of the tree arrays (== ngames of mcts(), mcts_gpu.jl:481) */
  int32_t width;      /* MLP width n  (ressimplesf n_filter, DenseNet.jl:193) */
  int32_t blocks;     /* MLP residual blocks k (n_tower) */
  int32_t device;     /* CUDA device ordinal */
  int32_t nn_mode;    /* agpu_nn_mode */
} agpu_config;

/* plugin constants: VectorizedState, FeatureSize, maxActions, maxLengthGame (4IARow.jl:2,8-11 …) */
typedef struct agpu_game_info {
  int32_t max_actions;
  int32_t vectorized_state;
  int32_t feature_size;
  int32_t max_length_game;
  int32_t position_bytes; /* 104 or 152 */
} agpu_game_info;

typedef struct agpu_ctx agpu_ctx;

/* ---- life cycle -------------------------------------------------------------------------- */
/* host-only: plugin constants for a (game, n, nvict); works without a GPU. */
int agpu_game_info_get(int32_t game, int32_t n, int32_t nvict, agpu_game_info* out);
int agpu_abi_version(void);
/* replaces mcts_gpu.init(positions, visits) (mcts_gpu.jl:342-357) + create_cunodes_stats/create_roots (:35-53) */
int agpu_create(agpu_ctx** out, const agpu_config* cfg);
void agpu_destroy(agpu_ctx* ctx);
/* message of the last failing call on ctx (ctx may be NULL for agpu_create failures) */
const char* agpu_last_error(const agpu_ctx* ctx);

/* ---- network ------------------------------------------------------------------------------ */
/* replaces convert_back(net)::snetwork2 (DenseNet.jl:279-286,331-333).  Julia column-major fp32:
 * base (width x 2VS), res[k] (width x width) each, pol_w (A x width), pol_b (A), val_w (1 x width), val_b (1).
 * slot 0/1: two resident networks (self-play uses 0; the duel uses both, mcts_gpu.jl:592-596). */
int agpu_set_weights(agpu_ctx* ctx, int32_t slot, const float* base, const float* const* res,
                     const float* pol_w, const float* pol_b, const float* val_w, const float* val_b);
/* actor(x; training=false) -> (logits, value) on device for L host-encoded inputs x[L][2VS]
 * (DenseNet.jl:294-304).  logits[L][A] are pre-softmax; value[L] is after the sigmoid. */
int agpu_forward(agpu_ctx* ctx, int32_t slot, const float* x, int64_t L, float* logits, float* value);

/* ---- game plugin surface, batched on the device (Position / canPlay / play / isOver) ------ */
int agpu_position_init(agpu_ctx* ctx, void* positions_out, int64_t n);                      /* Position() */
int agpu_can_play(agpu_ctx* ctx, const void* positions, int64_t n, uint8_t* legal /* [n][A] */);
int agpu_play(agpu_ctx* ctx, const void* positions, const int32_t* actions, int64_t n, void* positions_out);
int agpu_is_over(agpu_ctx* ctx, const void* positions, int64_t n, uint8_t* over, int8_t* result);
/* decoder (mcts_gpu.jl:202-223): batch[n][2VS] of 0/1 */
int agpu_encode(agpu_ctx* ctx, const void* positions, int64_t n, float* batch);

/* ---- one search: the mcts_single seam ---------------------------------------------------- */
/* replaces re_init(cu(positions), vnodes, L, …) (mcts_gpu.jl:359-373).  uids (optional, length L)
 * are the global game ids that key the RNG; default 0..L-1. */
int agpu_reinit(agpu_ctx* ctx, const void* positions, int64_t L, const uint32_t* uids);
/* replaces mcts_single(actor, visits, 256, vnodes, vnodesStats, leaf, newindex, L; training, cpuct, noise)
 * (mcts_gpu.jl:376-462).  prob: optional injected uniforms [visits][L][maxLengthGame] standing in
 * for CUDA.rand (:397); NULL = Philox4x32-10 keyed (seed, uid, ply, rollout, depth).
 * `noise` is accepted and ignored exactly as the reference ignores it (mcts_gpu.jl:250,273). */
int agpu_search(agpu_ctx* ctx, int64_t L, int32_t slot, int32_t visits, int32_t training, float cpuct, float noise,
                const float* prob, uint64_t seed, uint32_t ply);
/* vnodesStats.policy_final (A x L) and vnodesStats.batch (2VS x L) after a search (mcts_gpu.jl:441-443,506) */
int agpu_get_roots(agpu_ctx* ctx, int64_t L, float* policy_final, float* batch);

/* ---- the kernels of one rollout, individually (lower seam; injected-evaluator parity) ----- */
int agpu_search_begin(agpu_ctx* ctx, int64_t L);                                 /* the 8 fills, mcts_gpu.jl:380-387 */
/* kdescendTree! + decoder (mcts_gpu.jl:100-223); last_rollout!=0 publishes the root policy (copy_pol, :330-339) */
int agpu_select(agpu_ctx* ctx, int64_t L, int32_t rollout, int32_t last_rollout, float cpuct, const float* prob,
                uint64_t seed, uint32_t ply);
int agpu_get_leaves(agpu_ctx* ctx, int64_t L, int32_t* leaf /* 1-based node ids */, float* batch /* [L][2VS] */);
/* actor on the current leaves (mcts_gpu.jl:414); results stay on the device, optionally copied out */
int agpu_eval(agpu_ctx* ctx, int64_t L, int32_t slot, float* logits /* [L][A] or NULL */, float* value /* [L] or NULL */);
/* softmax! + expand + backUp (mcts_gpu.jl:417-431).  prior/value NULL = use agpu_eval's device result
 * (softmax applied here); non-NULL prior[L][A] is taken as already softmaxed, as expand's input is.
 * (Where the engine re-solves the regularised policy during the backup it uses the cpuct of the preceding agpu_select.) */
int agpu_expand_backup(agpu_ctx* ctx, int64_t L, int32_t training, int32_t last_rollout, const float* prior, const float* value);

/* Tree tables for parity checks, layout-neutral: all [game][node]([action]); ids 1-based, 0 = none.
 * Any pointer may be NULL.  states: Position wire structs [L][R]. */
typedef struct agpu_tree_dump {
  int32_t* nnodes;   /* [L]            newindex */
  int32_t* parent;   /* [L][R]         vnodes.parent */
  int32_t* action;   /* [L][R]         vnodes.actionFromParent */
  int32_t* child;    /* [L][R][A]      node reached by the action (Achild∘childID) */
  int32_t* order;    /* [L][R][A]      action of the slot-th created child (childID order), 0 past nchild */
  int32_t* nchild;   /* [L][R]         childnbr */
  int8_t* expanded;  /* [L][R] */
  float* prior;      /* [L][R][A] */
  float* q;          /* [L][R][A] */
  float* visits;     /* [L][R][A] */
  void* states;      /* [L][R] Position */
} agpu_tree_dump;
int agpu_get_tree(agpu_ctx* ctx, int64_t L, agpu_tree_dump* out);

/* ---- whole loops --------------------------------------------------------------------------- */
/* Samples in push order (ply-major, live games in slot order), SoA — the fields of
 * Game.Sample (main4IARow.jl:29-37) after update_buffer (main4IARow.jl:65-75). */
typedef struct agpu_samples {
  int64_t capacity;  /* in: rows available in every array below */
  int64_t count;     /* out: rows produced (may exceed capacity: then only `capacity` rows were written) */
  int8_t* state;     /* [capacity][2VS] */
  float* policy;     /* [capacity][A] */
  int8_t* player;    /* [capacity] */
  float* value;      /* [capacity] */
  int8_t* fstate;    /* [capacity][FS] */
  int32_t* game;     /* [capacity] global game uid (optional, may be NULL) */
  int32_t* ply;      /* [capacity] (optional, may be NULL) */
} agpu_samples;

typedef struct agpu_run_stats {
  int64_t sims;          /* sum over plies of live games x rollouts */
  int64_t positions;     /* sum over plies of live games */
  int64_t plies;
  int64_t total_length;  /* sum over games of the ply index at which they ended (mcts_gpu.jl:535) */
  int64_t faults;        /* illegal moves chosen ("faute") */
  int64_t kernel_launches;
  double device_ms;      /* CUDA-event time of the whole loop on the library's stream */
  double search_ms;      /* … of the rollout kernels only */
} agpu_run_stats;

/* replaces mcts_gpu.mcts(actor, visits, ngames, buffer; cpuct, noise) (mcts_gpu.jl:477-579) with the
 * whole ply loop on the device; games get uids uid_base..uid_base+ngames-1 (sharding key).
 * samples may be NULL (device-resident run: nothing is copied back but results/stats).
 * results = [v, n, d] as printed at mcts_gpu.jl:574. */
int agpu_selfplay(agpu_ctx* ctx, int32_t slot, int32_t visits, int64_t ngames, uint32_t uid_base, float cpuct, float noise,
                  uint64_t seed, agpu_samples* samples, int64_t results[3], agpu_run_stats* stats);
/* replaces mcts(actor1, actor2, visits, ngames; cpuct) (mcts_gpu.jl:581-651): slot_a moves on even plies */
int agpu_duel(agpu_ctx* ctx, int32_t slot_a, int32_t slot_b, int32_t visits, int64_t ngames, uint32_t uid_base, float cpuct,
              uint64_t seed, int64_t results[3], agpu_run_stats* stats);

/* ---- several GPUs behind one call ------------------------------------------------------------
 * The reference is one process on one GPU (selfplay.jl:34 calls mcts_gpu.mcts once per generation).  The same single call can drive
 * `ngpus` devices: one context per device inside the library, one host thread and one stream per device, the `ngames` games
 * block-partitioned over the devices by uid (results do not depend on the partition: the RNG is keyed by uid), no collective on the
 * search path, and the sample gather done here — every device copies its block to its place in the caller's arrays (device order =
 * ascending uid blocks; within a block, push order).  cfg->max_games is the TOTAL number of games; cfg->device is ignored
 * (devices[r], or r when devices is NULL). */
typedef struct agpu_multi agpu_multi;
int agpu_multi_create(agpu_multi** out, const agpu_config* cfg, int32_t ngpus, const int32_t* devices);
void agpu_multi_destroy(agpu_multi* m);
const char* agpu_multi_last_error(const agpu_multi* m);
int agpu_multi_ngpus(const agpu_multi* m);
agpu_ctx* agpu_multi_context(agpu_multi* m, int32_t index);   /* the per-device context, e.g. for agpu_get_kernel_times */
int agpu_multi_set_weights(agpu_multi* m, int32_t slot, const float* base, const float* const* res, const float* pol_w, const float* pol_b,
                           const float* val_w, const float* val_b);
/* == agpu_selfplay / agpu_duel over all devices; stats: sums, except plies and device_ms = max over devices */
int agpu_multi_selfplay(agpu_multi* m, int32_t slot, int32_t visits, int64_t ngames, uint32_t uid_base, float cpuct, float noise,
                        uint64_t seed, agpu_samples* samples, int64_t results[3], agpu_run_stats* stats);
int agpu_multi_duel(agpu_multi* m, int32_t slot_a, int32_t slot_b, int32_t visits, int64_t ngames, uint32_t uid_base, float cpuct,
                    uint64_t seed, int64_t results[3], agpu_run_stats* stats);

/* ---- measurement --------------------------------------------------------------------------- */
#define AGPU_NKERNELS 8
typedef struct agpu_kernel_times {
  /* per kernel class: launches and summed CUDA-event milliseconds since the last reset, measured on
   * the library's own stream when profiling is enabled.  Classes: 0 select 1 nn 2 expand_backup
   * 3 begin/reinit 4 finish_ply 5 compact 6 finalize 7 fused per-ply kernel (whole rollout loop) and misc */
  int64_t launches[AGPU_NKERNELS];
  double ms[AGPU_NKERNELS];
  int64_t nodes_traversed;   /* expanded nodes visited by descents (for d-bar), when profiling */
  int64_t descents;
} agpu_kernel_times;
int agpu_profile(agpu_ctx* ctx, int32_t enable);   /* enable=1: bracket every launch with events (slows the loop) */
int agpu_get_kernel_times(agpu_ctx* ctx, agpu_kernel_times* out, int32_t reset);
/* bytes of one node record / one game tree in HBM for this configuration */
int agpu_layout_info(agpu_ctx* ctx, int64_t* node_bytes, int64_t* game_bytes, int64_t* lanes_per_game);

/* Page-locked host memory for the caller's sample arrays (agpu_samples): with page-locked destinations agpu_selfplay streams the rows of
 * every ply out while the next ply searches, and the final copies run at the speed of the link (B200: 96 MB in 2 ms; pageable arrays
 * take 25 ms).  Any host language can bind these two instead of a CUDA runtime of its own.  The reference keeps its samples in ordinary
 * Julia vectors (main4IARow.jl:29-47); the glue allocates once per context and copies into its PoolSample. */
int agpu_host_alloc(void** out, uint64_t bytes);
int agpu_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* ALPHAGPU_H */
