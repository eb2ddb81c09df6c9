// api.cu — the C ABI of include/alphagpu.h: argument checking and run-time dispatch on the game plugin.
#include <mutex>
#include <thread>
#include <vector>

#include "engine.cuh"

using namespace ag;

struct agpu_ctx {
  EngineBase* eng;
};

static thread_local std::string g_create_error;

namespace ag {
__global__ void debug_expf_kernel(const float* x, long long n, float* y, int sigmoid) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = sigmoid ? c_sigmoidf(x[i]) : c_expf(x[i]);
}
// fdiv_fast against __fdiv_rn on random operands of its box (and a band outside it, where `bad` must be raised whenever they differ)
__global__ void debug_fdiv_kernel(unsigned long long n_per_thread, unsigned long long seed, unsigned long long* out) {
  const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long mism = 0, flagged = 0;
  for (unsigned long long i = 0; i < n_per_thread; i += 2) {
    const Philox4 r = philox4x32_10((u32)tid, (u32)(tid >> 32), (u32)i, (u32)(i >> 32), (u32)seed, (u32)(seed >> 32));
    for (int h = 0; h < 2; h++) {
      const u32 wa = r.v[2 * h], wb = r.v[2 * h + 1];
      // exponent field 60..194: the box is 67..187, so ~10 % of the operands fall outside it
      const u32 ea = 60u + (wa >> 9) % 135u, eb = 60u + (wb >> 9) % 135u;
      // mostly positive operands (the box), a negative sign one time in sixteen
      float a = __uint_as_float((wa & 0x007FFFFFu) | (ea << 23) | ((wa >> 28) == 0 ? 0x80000000u : 0u));
      const float b = __uint_as_float((wb & 0x007FFFFFu) | (eb << 23) | ((wb >> 28) == 0 ? 0x80000000u : 0u));
      if ((wa & 0x7F000u) == 0) a = (wa & 0x08000000u) ? -0.f : 0.f;         // zeros of both signs now and then
      const bool bad = !(fdiv_box_num(a) && fdiv_box_den(b));
      const float f = fdiv_fast(a, b);
      const float t = __fdiv_rn(a, b);
      if (bad) flagged++;
      else if (__float_as_uint(f) != __float_as_uint(t)) mism++;
    }
  }
  atomicAdd(&out[0], mism);
  atomicAdd(&out[1], flagged);
}
}  // namespace ag

static bool fill_info(int32_t game, int32_t n, int32_t nvict, agpu_game_info* o) {
  switch (game) {
    case AGPU_CONNECT4: *o = {7, 42, 42, 42, 104}; return true;                                   // 4IARow.jl:6-12
    case AGPU_GOBANG:                                                                              // Gobang.jl:8-11
      if (n < 1 || n * n > 192 || nvict < 2 || nvict > n) return false;
      *o = {n * n, n * n, n * n, n * n, 104}; return true;
    case AGPU_HEX:                                                                                 // Hex.jl:8-11
      if (n < 2 || (n + 1) * (n + 1) > 192) return false;
      *o = {n * n, (n + 1) * (n + 1), (n + 1) * (n + 1), n * n, 104}; return true;
    case AGPU_REVERSI8: *o = {65, 64, 64, 70, 152}; return true;                                   // Reversi8x8.jl:5-8
    case AGPU_REVERSI6: *o = {37, 36, 36, 50, 152}; return true;                                   // Reversi6x6.jl:6-9
  }
  return false;
}

// ---- several GPUs behind ONE call from ONE caller thread (the caller is selfplay.jl:34 / :56) ----
struct agpu_multi {
  std::vector<agpu_ctx*> ctx;
  std::string err;
};

// games [0, n) block-partitioned over `parts` shards: shard r plays [base, base + count)
static void shard_block(int64_t n, int parts, int r, int64_t* base, int64_t* count) {
  const int64_t q = n / parts, rem = n % parts;
  *count = q + (r < rem ? 1 : 0);
  *base = r * q + (r < rem ? r : rem);
}

// one host thread per device runs f(r); returns the first non-OK status (AGPU_ERR_ILLEGAL_MOVE ranks last: the run itself completed)
template <class F> static int multi_for_each(agpu_multi* m, F&& f) {
  const int n = (int)m->ctx.size();
  std::vector<int> rc(n, AGPU_OK);
  std::vector<std::thread> th;
  for (int r = 1; r < n; r++) th.emplace_back([&, r] { rc[r] = f(r); });
  rc[0] = f(0);
  for (auto& t : th) t.join();
  int out = AGPU_OK;
  for (int r = 0; r < n; r++)
    if (rc[r] != AGPU_OK && (out == AGPU_OK || out == AGPU_ERR_ILLEGAL_MOVE)) { out = rc[r]; m->err = m->ctx[r]->eng->err; }
  return out;
}

static int multi_run(agpu_multi* m, int32_t slot, int32_t visits, int64_t ngames, uint32_t uid_base, float cpuct, uint64_t seed,
                     agpu_samples* samples, int64_t results[3], agpu_run_stats* stats, bool duel, int32_t slot_b) {
  if (!m || !results || ngames < 1) return AGPU_ERR_INVALID;
  const int n = (int)m->ctx.size();
  std::vector<int64_t> res(3 * (size_t)n, 0);
  std::vector<agpu_run_stats> st((size_t)n);
  for (auto& s : st) memset(&s, 0, sizeof(s));
  // phase 1: every device plays its block of games to the end; the samples stay in its memory
  const int rc1 = multi_for_each(m, [&](int r) {
    int64_t base, count;
    shard_block(ngames, n, r, &base, &count);
    if (count == 0) return (int)AGPU_OK;
    return m->ctx[r]->eng->selfplay(slot, visits, count, uid_base + (uint32_t)base, cpuct, seed, nullptr, &res[3 * (size_t)r], &st[(size_t)r], duel, slot_b);
  });
  if (rc1 != AGPU_OK && rc1 != AGPU_ERR_ILLEGAL_MOVE) return rc1;
  results[0] = results[1] = results[2] = 0;
  agpu_run_stats tot;
  memset(&tot, 0, sizeof(tot));
  std::vector<int64_t> offset((size_t)n + 1, 0);
  for (int r = 0; r < n; r++) {
    for (int i = 0; i < 3; i++) results[i] += res[3 * (size_t)r + i];
    tot.sims += st[r].sims; tot.positions += st[r].positions; tot.total_length += st[r].total_length; tot.faults += st[r].faults;
    tot.kernel_launches += st[r].kernel_launches;
    tot.plies = std::max(tot.plies, st[r].plies); tot.device_ms = std::max(tot.device_ms, st[r].device_ms);
    int64_t base, count;
    shard_block(ngames, n, r, &base, &count);
    offset[(size_t)r + 1] = offset[(size_t)r] + (count > 0 && !duel ? m->ctx[r]->eng->last_samples() : 0);
  }
  if (stats) *stats = tot;
  // phase 2: the gather — every device copies its block to its place in the caller's arrays (device order = ascending uid blocks)
  if (samples && !duel) {
    samples->count = offset[(size_t)n];
    const int rc2 = multi_for_each(m, [&](int r) { return m->ctx[r]->eng->fetch_samples(samples, offset[(size_t)r]); });
    if (rc2 != AGPU_OK) return rc2;
  }
  return rc1;
}

extern "C" {

int agpu_abi_version(void) { return AGPU_ABI_VERSION; }

int agpu_game_info_get(int32_t game, int32_t n, int32_t nvict, agpu_game_info* out) {
  if (!out) return AGPU_ERR_INVALID;
  return fill_info(game, n, nvict, out) ? AGPU_OK : AGPU_ERR_INVALID;
}

int agpu_create(agpu_ctx** out, const agpu_config* cfg) {
  if (!out || !cfg) { g_create_error = "null argument"; return AGPU_ERR_INVALID; }
  *out = nullptr;
  agpu_game_info info;
  if (!fill_info(cfg->game, cfg->n, cfg->nvict, &info)) { g_create_error = "unknown game or board size"; return AGPU_ERR_INVALID; }
  EngineBase* e = nullptr;
  switch (cfg->game) {
    case AGPU_CONNECT4: e = make_engine_connect4(); break;
    case AGPU_GOBANG: e = make_engine_gobang(cfg->n, cfg->nvict); break;
    case AGPU_HEX: e = make_engine_hex(cfg->n); break;
    case AGPU_REVERSI8: e = make_engine_reversi(8); break;
    case AGPU_REVERSI6: e = make_engine_reversi(6); break;
  }
  if (!e) {
    g_create_error = "this (game, n, nvict) is not compiled in; see AG_GOBANG_SIZES / AG_HEX_SIZES in engine_games.cu";
    return AGPU_ERR_INVALID;
  }
  e->cfg = *cfg;
  e->info = info;
  int rc = e->init();
  if (rc != AGPU_OK) { g_create_error = e->err; delete e; return rc; }
  *out = new agpu_ctx{e};
  return AGPU_OK;
}

void agpu_destroy(agpu_ctx* ctx) {
  if (!ctx) return;
  delete ctx->eng;
  delete ctx;
}

const char* agpu_last_error(const agpu_ctx* ctx) { return ctx ? ctx->eng->err.c_str() : g_create_error.c_str(); }

#define CTX_OR_FAIL() \
  if (!ctx) return AGPU_ERR_INVALID;

int agpu_set_weights(agpu_ctx* ctx, int32_t slot, const float* base, const float* const* res, const float* pol_w, const float* pol_b,
                     const float* val_w, const float* val_b) {
  CTX_OR_FAIL();
  return ctx->eng->set_weights(slot, base, res, pol_w, pol_b, val_w, val_b);
}
int agpu_forward(agpu_ctx* ctx, int32_t slot, const float* x, int64_t L, float* logits, float* value) {
  CTX_OR_FAIL();
  return ctx->eng->forward(slot, x, L, logits, value);
}
int agpu_position_init(agpu_ctx* ctx, void* positions_out, int64_t n) {
  CTX_OR_FAIL();
  return ctx->eng->position_init(positions_out, n);
}
int agpu_can_play(agpu_ctx* ctx, const void* positions, int64_t n, uint8_t* legal) {
  CTX_OR_FAIL();
  if (!legal) return AGPU_ERR_INVALID;
  return ctx->eng->game_ops(positions, nullptr, n, nullptr, legal, nullptr, nullptr, nullptr);
}
int agpu_play(agpu_ctx* ctx, const void* positions, const int32_t* actions, int64_t n, void* positions_out) {
  CTX_OR_FAIL();
  if (!positions_out || !actions) return AGPU_ERR_INVALID;
  return ctx->eng->game_ops(positions, actions, n, positions_out, nullptr, nullptr, nullptr, nullptr);
}
int agpu_is_over(agpu_ctx* ctx, const void* positions, int64_t n, uint8_t* over, int8_t* result) {
  CTX_OR_FAIL();
  if (!over || !result) return AGPU_ERR_INVALID;
  return ctx->eng->game_ops(positions, nullptr, n, nullptr, nullptr, over, result, nullptr);
}
int agpu_encode(agpu_ctx* ctx, const void* positions, int64_t n, float* batch) {
  CTX_OR_FAIL();
  if (!batch) return AGPU_ERR_INVALID;
  return ctx->eng->game_ops(positions, nullptr, n, nullptr, nullptr, nullptr, nullptr, batch);
}
int agpu_reinit(agpu_ctx* ctx, const void* positions, int64_t L, const uint32_t* uids) {
  CTX_OR_FAIL();
  return ctx->eng->reinit(positions, L, uids);
}
int agpu_search(agpu_ctx* ctx, int64_t L, int32_t slot, int32_t visits, int32_t training, float cpuct, float noise, const float* prob,
                uint64_t seed, uint32_t ply) {
  CTX_OR_FAIL();
  (void)noise;   // parsed and ignored by the reference as well (mcts_gpu.jl:250,273)
  return ctx->eng->search(L, slot, visits, training, cpuct, prob, seed, ply);
}
int agpu_get_roots(agpu_ctx* ctx, int64_t L, float* policy_final, float* batch) {
  CTX_OR_FAIL();
  return ctx->eng->get_roots(L, policy_final, batch);
}
int agpu_search_begin(agpu_ctx* ctx, int64_t L) {
  CTX_OR_FAIL();
  return ctx->eng->search_begin(L);
}
int agpu_select(agpu_ctx* ctx, int64_t L, int32_t rollout, int32_t last_rollout, float cpuct, const float* prob, uint64_t seed, uint32_t ply) {
  CTX_OR_FAIL();
  return ctx->eng->select(L, rollout, last_rollout, cpuct, prob, seed, ply);
}
int agpu_get_leaves(agpu_ctx* ctx, int64_t L, int32_t* leaf, float* batch) {
  CTX_OR_FAIL();
  return ctx->eng->get_leaves(L, leaf, batch);
}
int agpu_eval(agpu_ctx* ctx, int64_t L, int32_t slot, float* logits, float* value) {
  CTX_OR_FAIL();
  return ctx->eng->eval(L, slot, logits, value);
}
int agpu_expand_backup(agpu_ctx* ctx, int64_t L, int32_t training, int32_t last_rollout, const float* prior, const float* value) {
  CTX_OR_FAIL();
  return ctx->eng->expand_backup(L, training, last_rollout, prior, value);
}
int agpu_get_tree(agpu_ctx* ctx, int64_t L, agpu_tree_dump* out) {
  CTX_OR_FAIL();
  return ctx->eng->get_tree(L, out);
}
int agpu_selfplay(agpu_ctx* ctx, int32_t slot, int32_t visits, int64_t ngames, uint32_t uid_base, float cpuct, float noise, uint64_t seed,
                  agpu_samples* samples, int64_t results[3], agpu_run_stats* stats) {
  CTX_OR_FAIL();
  (void)noise;
  return ctx->eng->selfplay(slot, visits, ngames, uid_base, cpuct, seed, samples, results, stats, false, 0);
}
int agpu_duel(agpu_ctx* ctx, int32_t slot_a, int32_t slot_b, int32_t visits, int64_t ngames, uint32_t uid_base, float cpuct, uint64_t seed,
              int64_t results[3], agpu_run_stats* stats) {
  CTX_OR_FAIL();
  return ctx->eng->selfplay(slot_a, visits, ngames, uid_base, cpuct, seed, nullptr, results, stats, true, slot_b);
}
int agpu_profile(agpu_ctx* ctx, int32_t enable) {
  CTX_OR_FAIL();
  return ctx->eng->profile(enable);
}
int agpu_get_kernel_times(agpu_ctx* ctx, agpu_kernel_times* out, int32_t reset) {
  CTX_OR_FAIL();
  return ctx->eng->kernel_times(out, reset);
}
int agpu_host_alloc(void** out, uint64_t bytes) {
  if (!out || bytes == 0) return AGPU_ERR_INVALID;
  *out = nullptr;
  return cudaHostAlloc(out, (size_t)bytes, cudaHostAllocPortable) == cudaSuccess ? AGPU_OK : AGPU_ERR_CUDA;
}
int agpu_host_free(void* p) { return (p == nullptr || cudaFreeHost(p) == cudaSuccess) ? AGPU_OK : AGPU_ERR_CUDA; }

int agpu_layout_info(agpu_ctx* ctx, int64_t* node_bytes, int64_t* game_bytes, int64_t* lanes_per_game) {
  CTX_OR_FAIL();
  return ctx->eng->layout_info(node_bytes, game_bytes, lanes_per_game);
}
int agpu_multi_create(agpu_multi** out, const agpu_config* cfg, int32_t ngpus, const int32_t* devices) {
  if (!out || !cfg || ngpus < 1 || ngpus > 64) { g_create_error = "bad arguments"; return AGPU_ERR_INVALID; }
  *out = nullptr;
  agpu_multi* m = new agpu_multi;
  for (int r = 0; r < ngpus; r++) {
    agpu_config c = *cfg;
    c.device = devices ? devices[r] : r;
    c.max_games = (cfg->max_games + ngpus - 1) / ngpus;                 // the largest shard
    agpu_ctx* x = nullptr;
    const int rc = agpu_create(&x, &c);
    if (rc != AGPU_OK) {
      for (agpu_ctx* y : m->ctx) agpu_destroy(y);
      delete m;
      return rc;                                                        // message in agpu_last_error(NULL)
    }
    m->ctx.push_back(x);
  }
  *out = m;
  return AGPU_OK;
}
void agpu_multi_destroy(agpu_multi* m) {
  if (!m) return;
  for (agpu_ctx* x : m->ctx) agpu_destroy(x);
  delete m;
}
const char* agpu_multi_last_error(const agpu_multi* m) { return m ? m->err.c_str() : g_create_error.c_str(); }
int agpu_multi_ngpus(const agpu_multi* m) { return m ? (int)m->ctx.size() : 0; }
agpu_ctx* agpu_multi_context(agpu_multi* m, int32_t index) { return (m && index >= 0 && index < (int)m->ctx.size()) ? m->ctx[index] : nullptr; }

int agpu_multi_set_weights(agpu_multi* m, int32_t slot, const float* base, const float* const* res, const float* pol_w, const float* pol_b,
                           const float* val_w, const float* val_b) {
  if (!m) return AGPU_ERR_INVALID;
  for (agpu_ctx* x : m->ctx) {
    const int rc = x->eng->set_weights(slot, base, res, pol_w, pol_b, val_w, val_b);
    if (rc != AGPU_OK) { m->err = x->eng->err; return rc; }
  }
  return AGPU_OK;
}

int agpu_multi_selfplay(agpu_multi* m, int32_t slot, int32_t visits, int64_t ngames, uint32_t uid_base, float cpuct, float noise, uint64_t seed,
                        agpu_samples* samples, int64_t results[3], agpu_run_stats* stats) {
  (void)noise;
  return multi_run(m, slot, visits, ngames, uid_base, cpuct, seed, samples, results, stats, false, 0);
}
int agpu_multi_duel(agpu_multi* m, int32_t slot_a, int32_t slot_b, int32_t visits, int64_t ngames, uint32_t uid_base, float cpuct, uint64_t seed,
                    int64_t results[3], agpu_run_stats* stats) {
  return multi_run(m, slot_a, visits, ngames, uid_base, cpuct, seed, nullptr, results, stats, true, slot_b);
}

/* development hook: device buffer receiving clock64 stamps of the tensor-core chain ([cta][tile][16 layers][4]); NULL disables */
int agpu_debug_tc_trace(void* dev_buf) { ag::g_tc_dbg = (long long*)dev_buf; return AGPU_OK; }
/* test hook: the canonical exp / sigmoid evaluated on the device */
int agpu_debug_fdiv_check(uint64_t pairs, uint64_t seed, uint64_t out[2]) {
  if (!out) return AGPU_ERR_INVALID;
  unsigned long long* d = nullptr;
  if (cudaMalloc((void**)&d, 16) != cudaSuccess) return AGPU_ERR_CUDA;
  cudaMemset(d, 0, 16);
  const int blocks = 148 * 8, threads = 256;
  const unsigned long long per = (pairs + (unsigned long long)blocks * threads - 1) / ((unsigned long long)blocks * threads);
  ag::debug_fdiv_kernel<<<blocks, threads>>>((per + 1) & ~1ull, seed, d);
  unsigned long long h[2] = {0, 0};
  const cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return AGPU_ERR_CUDA;
  out[0] = h[0]; out[1] = h[1];
  return AGPU_OK;
}
int agpu_debug_expf(agpu_ctx* ctx, const float* x, int64_t n, float* y, int32_t sigmoid) {
  CTX_OR_FAIL();
  if (!x || !y || n < 1) return AGPU_ERR_INVALID;
  return ctx->eng->debug_expf(x, n, y, sigmoid);
}

}  // extern "C"
