// engine.cuh — host side of one context: device memory, launch sequencing, the self-play / duel loops.
// Type-erased through EngineBase so api.cu can dispatch on the game at run time; one EngineT<G> per
// game plugin is instantiated in engine_*.cu.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <cstring>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges show up in nsys / ncu --nvtx timelines, no-ops without a tool attached

#include "../../include/alphagpu.h"
#include "nn.cuh"
#include "search.cuh"
#include "fused.cuh"

namespace ag {

enum { K_SELECT = 0, K_NN = 1, K_EXPAND = 2, K_BEGIN = 3, K_FINISH = 4, K_COMPACT = 5, K_FINALIZE = 6, K_OTHER = 7 };

#define AG_CK(call)                                                                                      \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) {                                                                             \
      char buf_[512];                                                                                    \
      snprintf(buf_, sizeof(buf_), "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      err = buf_;                                                                                        \
      return AGPU_ERR_CUDA;                                                                              \
    }                                                                                                    \
  } while (0)

// NVTX range for the enclosing scope (the reference brackets its phases with time() stamps, mcts_gpu.jl:377-445)
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

#define AG_REQUIRE(cond, code, msg) \
  do {                              \
    if (!(cond)) {                  \
      err = msg;                    \
      return code;                  \
    }                               \
  } while (0)

struct EngineBase {
  agpu_config cfg;
  agpu_game_info info;
  std::string err;
  virtual ~EngineBase() {}
  virtual int init() = 0;
  virtual int set_weights(int slot, const float* base, const float* const* res, const float* pol_w, const float* pol_b,
                          const float* val_w, const float* val_b) = 0;
  virtual int forward(int slot, const float* x, int64_t L, float* logits, float* value) = 0;
  virtual int position_init(void* out, int64_t n) = 0;
  virtual int game_ops(const void* pos, const int32_t* actions, int64_t n, void* played, uint8_t* legal, uint8_t* over, int8_t* result,
                       float* enc) = 0;
  virtual int reinit(const void* positions, int64_t L, const uint32_t* uids) = 0;
  virtual int search(int64_t L, int slot, int visits, int training, float cpuct, const float* prob, uint64_t seed, uint32_t ply) = 0;
  virtual int get_roots(int64_t L, float* policy_final, float* batch) = 0;
  virtual int search_begin(int64_t L) = 0;
  virtual int select(int64_t L, int rollout, int last, float cpuct, const float* prob, uint64_t seed, uint32_t ply) = 0;
  virtual int get_leaves(int64_t L, int32_t* leaf, float* batch) = 0;
  virtual int eval(int64_t L, int slot, float* logits, float* value) = 0;
  virtual int expand_backup(int64_t L, int training, int last, const float* prior, const float* value) = 0;
  virtual int get_tree(int64_t L, agpu_tree_dump* out) = 0;
  virtual int selfplay(int slot, int visits, int64_t ngames, uint32_t uid_base, float cpuct, uint64_t seed, agpu_samples* samples,
                       int64_t results[3], agpu_run_stats* stats, bool duel, int slot_b) = 0;
  virtual int fetch_samples(agpu_samples* out, int64_t row_offset) = 0;
  virtual int64_t last_samples() const = 0;
  virtual int profile(int enable) = 0;
  virtual int kernel_times(agpu_kernel_times* out, int reset) = 0;
  virtual int layout_info(int64_t* node_bytes, int64_t* game_bytes, int64_t* lanes) = 0;
  virtual int debug_expf(const float* x, int64_t n, float* y, int sigmoid) = 0;
};

__global__ void debug_expf_kernel(const float* x, long long n, float* y, int sigmoid);

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t ensure(size_t count) {
    if (count <= n) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; n = 0;
    cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }              // temporaries of forward / game_ops / ... are freed on every early return
};

struct NetSlot {
  bool set = false;
  DevBuf<float> base, res, pol_w, pol_b, val_w, val_b, tc_bias;
  DevBuf<unsigned char> tc_img;
  NetDev dev;
};

template <class G>
struct EngineT : EngineBase {
  typedef Layout<G> Lay;
  typedef typename G::State State;
  static constexpr int A = G::A;

  cudaStream_t stream = nullptr;
  // Sample streaming: the rows a ply appends to the sample arrays (state, policy, player, game, ply — final when pushed) are copied to the
  // caller's buffers on a second stream while the next ply searches; only value and fstate, which need the game's end, are copied after
  // the loop (26 of 96 MB at the bench size; B200: 66.1 -> 65.2 ms per generation end to end).  Used when the caller's arrays are
  // page-locked (a pageable destination would make every copy a blocking one); AGPU_STREAM_SAMPLES=0 switches it off.
  cudaStream_t copy_stream = nullptr;
  bool stream_samples = true;
  int64_t L_cap = 0;
  int R = 0;
  int64_t L_live = 0;
  SearchParams P{};
  DevBuf<char> tree;
  DevBuf<int32_t> nnodes, leaf, block_count, total_dev;
  DevBuf<uint32_t> uid, uid_b;
  DevBuf<float> policy_final, nn_out, d_prob, d_prior, d_value, d_fscratch;
  DevBuf<State> st_a, st_b, game_final;
  DevBuf<uint8_t> alive, d_u8;
  DevBuf<int8_t> game_result, d_i8;
  DevBuf<int32_t> d_i32;
  DevBuf<unsigned long long> tallies, counters;
  DevBuf<uint8_t> path_node, path_move, path_len;
  int64_t last_sample_count = 0;   // rows the last self-play run left in the device sample arrays (agpu_multi_selfplay gathers them)
  float last_cpuct = 2.0f;   // cpuct of the most recent descent: the backup re-solves π̄ with it (FAST layouts)
  int32_t* total_host = nullptr;   // pinned
  unsigned long long* fault_host = nullptr;   // pinned: tallies[4] of the ply just played ("faute", mcts_gpu.jl:526-529)
  unsigned long long* tallies_host = nullptr;
  NetSlot nets[2];
  // samples
  DevBuf<int8_t> s_state, s_player, s_fstate;
  DevBuf<float> s_policy, s_value;
  DevBuf<int32_t> s_game, s_ply;
  // segmented graph replay of the rollout loop: the live games are cut into `nseg` independent slices, each replays the whole
  // R-rollout loop from a CUDA graph on its own stream, so that one slice's network chain overlaps another slice's descents
  // and the per-launch latency floor is shared instead of paid serially.
  static constexpr int MAX_SEG = 8;
  int nseg = 4;
  cudaStream_t seg_stream[MAX_SEG] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[MAX_SEG] = {};
  SegParams* seg_host = nullptr;              // pinned [MAX_SEG]
  DevBuf<SegParams> seg_dev;
  int64_t seg_cap = 0;                        // slots a slice's graph is sized for
  struct GraphSet { int visits, slot, training; std::vector<cudaGraphExec_t> exec; };
  std::vector<GraphSet> graphs;
  // fused per-ply kernel (fused.cuh): available for small boards with the tensor-core chain
  static constexpr bool FUSED_OK = Lay::FAST && G::Geo::NC == 1 && 2 * G::VS <= tc::TC_N;   // and width 128, checked at run time
  bool use_fused = false;
  int num_sms = 148, fused_min_gpc = 8;
  // profiling
  bool profiling = false;
  struct Ev { cudaEvent_t a, b; int cls; };
  std::vector<Ev> ev_pending;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
  agpu_kernel_times kt{};
  int64_t launch_count = 0;

  ~EngineT() override {
    if (stream) cudaStreamSynchronize(stream);
    drop_graphs();
    for (int i = 0; i < MAX_SEG; i++) { if (seg_stream[i]) cudaStreamDestroy(seg_stream[i]); if (ev_join[i]) cudaEventDestroy(ev_join[i]); }
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (seg_host) cudaFreeHost(seg_host);
    seg_dev.release();
    for (auto& e : ev_pending) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    for (auto& e : ev_pool) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    path_node.release(); path_move.release(); path_len.release(); tree.release(); nnodes.release(); leaf.release(); block_count.release(); total_dev.release(); uid.release(); uid_b.release();
    policy_final.release(); nn_out.release(); d_prob.release(); d_prior.release(); d_value.release(); d_fscratch.release();
    st_a.release(); st_b.release(); game_final.release(); alive.release(); d_u8.release(); game_result.release(); d_i8.release();
    d_i32.release(); tallies.release(); counters.release();
    for (auto& n : nets) { n.base.release(); n.res.release(); n.pol_w.release(); n.pol_b.release(); n.val_w.release(); n.val_b.release(); n.tc_bias.release(); n.tc_img.release(); }
    s_state.release(); s_player.release(); s_fstate.release(); s_policy.release(); s_value.release(); s_game.release(); s_ply.release();
    if (total_host) cudaFreeHost(total_host);
    if (fault_host) cudaFreeHost(fault_host);
    if (tallies_host) cudaFreeHost(tallies_host);
    if (stream) cudaStreamDestroy(stream);
    if (copy_stream) cudaStreamDestroy(copy_stream);
  }

  // ---- launch bracket: counts launches, optionally times them with events on the library stream ----
  template <class F> void launch(int cls, F&& f) {
    launch_count++;
    kt.launches[cls]++;
    if (profiling) {
      std::pair<cudaEvent_t, cudaEvent_t> e;
      if (!ev_pool.empty()) { e = ev_pool.back(); ev_pool.pop_back(); }
      else { cudaEventCreate(&e.first); cudaEventCreate(&e.second); }
      cudaEventRecord(e.first, stream);
      f();
      cudaEventRecord(e.second, stream);
      ev_pending.push_back({e.first, e.second, cls});
    } else {
      f();
    }
  }
  void harvest() {
    if (ev_pending.empty()) return;
    cudaStreamSynchronize(stream);
    for (auto& e : ev_pending) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e.a, e.b);
      kt.ms[e.cls] += ms;
      ev_pool.push_back({e.a, e.b});
    }
    ev_pending.clear();
    if (counters.p) {
      unsigned long long c[2];
      cudaMemcpy(c, counters.p, sizeof(c), cudaMemcpyDeviceToHost);
      kt.nodes_traversed = (int64_t)c[0]; kt.descents = (int64_t)c[1];
    }
  }

  int init() override {
    AG_REQUIRE(cfg.rollouts >= 1 && cfg.rollouts <= 255, AGPU_ERR_INVALID, "rollouts must be in 1..255 (node ids are stored in 8 bits)");
    AG_REQUIRE(cfg.max_games >= 1 && cfg.max_games <= (1ll << 30), AGPU_ERR_INVALID, "max_games out of range");
    AG_REQUIRE(cfg.width >= 1 && cfg.width <= 1024 && cfg.width >= A + 1 && cfg.blocks >= 0 && cfg.blocks <= 64, AGPU_ERR_INVALID, "unsupported MLP shape");
    AG_REQUIRE(cfg.nn_mode == AGPU_NN_FP32 || cfg.nn_mode == AGPU_NN_BF16_TC || cfg.nn_mode == AGPU_NN_FP16_TC, AGPU_ERR_INVALID, "bad nn_mode");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { err = "no CUDA device (there is no CPU fallback)"; return AGPU_ERR_NO_DEVICE; }
    if (is_tc())
      AG_REQUIRE(tc_supported(2 * G::VS, cfg.width, cfg.blocks, A), AGPU_ERR_INVALID, "the tensor-core chain supports width 128 (2*VS <= 128) and width 512 (2*VS <= 256), heads up to 127 actions");
    AG_REQUIRE(cfg.device >= 0 && cfg.device < ndev, AGPU_ERR_INVALID, "device ordinal out of range");
    AG_CK(cudaSetDevice(cfg.device));
    AG_CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    if (const char* e = getenv("AGPU_STREAM_SAMPLES")) stream_samples = atoi(e) != 0;
    AG_CK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    L_cap = cfg.max_games; R = cfg.rollouts;
    AG_CK(tree.ensure((size_t)L_cap * Lay::game_bytes(R)));
    AG_CK(nnodes.ensure(L_cap)); AG_CK(leaf.ensure(L_cap)); AG_CK(uid.ensure(L_cap)); AG_CK(uid_b.ensure(L_cap));
    AG_CK(policy_final.ensure((size_t)L_cap * A)); AG_CK(nn_out.ensure((size_t)L_cap * Lay::OUTS));
    AG_CK(st_a.ensure(L_cap)); AG_CK(st_b.ensure(L_cap)); AG_CK(alive.ensure(L_cap));
    AG_CK(block_count.ensure((L_cap + 255) / 256 + 1)); AG_CK(total_dev.ensure(1));
    AG_CK(tallies.ensure(8)); AG_CK(counters.ensure(2));
    AG_CK(path_node.ensure((size_t)L_cap * R)); AG_CK(path_move.ensure((size_t)L_cap * R)); AG_CK(path_len.ensure(L_cap));
    AG_CK(cudaMemsetAsync(path_len.p, 0, L_cap, stream));
    AG_CK(cudaMallocHost((void**)&total_host, sizeof(int32_t)));
    AG_CK(cudaMallocHost((void**)&fault_host, sizeof(unsigned long long)));
    AG_CK(cudaMallocHost((void**)&tallies_host, 8 * sizeof(unsigned long long)));
    AG_CK(cudaMemsetAsync(tree.p, 0, (size_t)L_cap * Lay::game_bytes(R), stream));
    AG_CK(cudaMemsetAsync(policy_final.p, 0, (size_t)L_cap * A * sizeof(float), stream));
    AG_CK(cudaMemsetAsync(nn_out.p, 0, (size_t)L_cap * Lay::OUTS * sizeof(float), stream));
    AG_CK(cudaMemsetAsync(counters.p, 0, 2 * sizeof(unsigned long long), stream));
    P.tree = tree.p; P.game_stride = Lay::game_bytes(R); P.R = R; P.nnodes = nnodes.p; P.leaf = leaf.p; P.uid = uid.p;
    P.policy_final = policy_final.p; P.nn_out = nn_out.p; P.counters = nullptr;
    P.path_node = path_node.p; P.path_move = path_move.p; P.path_len = path_len.p;
    if (const char* e = getenv("AGPU_SEGMENTS")) nseg = atoi(e);
    if (nseg < 1) nseg = 1;
    if (nseg > MAX_SEG) nseg = MAX_SEG;
    for (int i = 0; i < nseg; i++) { AG_CK(cudaStreamCreateWithFlags(&seg_stream[i], cudaStreamNonBlocking)); AG_CK(cudaEventCreateWithFlags(&ev_join[i], cudaEventDisableTiming)); }
    AG_CK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    AG_CK(cudaMallocHost((void**)&seg_host, sizeof(SegParams) * MAX_SEG));
    AG_CK(seg_dev.ensure(MAX_SEG));
    seg_cap = ((L_cap + nseg - 1) / nseg + 255) / 256 * 256;
    if (is_tc()) AG_CK(tc_init());
    if constexpr (FUSED_OK) {
      if (is_tc() && cfg.width == tc::TC_N) {
        AG_CK(cudaFuncSetAttribute(fused::ply_kernel<G, 0, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, fused::FCfg<G, 1, 1>::SMEM));
        AG_CK(cudaFuncSetAttribute(fused::ply_kernel<G, 1, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, fused::FCfg<G, 1, 1>::SMEM));
        AG_CK(cudaFuncSetAttribute(fused::ply_kernel<G, 0, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, fused::FCfg<G, 1, 0>::SMEM));
        AG_CK(cudaFuncSetAttribute(fused::ply_kernel<G, 1, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, fused::FCfg<G, 1, 0>::SMEM));
        AG_CK(cudaFuncSetAttribute(fused::ply_kernel<G, 0, 2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, fused::FCfg<G, 2, 0>::SMEM));
        AG_CK(cudaFuncSetAttribute(fused::ply_kernel<G, 1, 2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, fused::FCfg<G, 2, 0>::SMEM));
        use_fused = true;
        if (const char* e = getenv("AGPU_FUSED")) use_fused = atoi(e) != 0;
        if (const char* e = getenv("AGPU_FUSED_MIN_GPC")) fused_min_gpc = atoi(e);
        cudaDeviceProp prop;
        AG_CK(cudaGetDeviceProperties(&prop, cfg.device));
        num_sms = prop.multiProcessorCount;
      }
    }
    AG_CK(cudaStreamSynchronize(stream));
    return AGPU_OK;
  }

  static bool host_pinned(const void* p) {
    cudaPointerAttributes a;
    if (p == nullptr || cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
  }

  // one launch for the whole rollout loop of a ply (fused.cuh)
  int enqueue_search_fused(int64_t L, int slot, int visits, int training, float cpuct, uint64_t seed, uint32_t ply) {
    if constexpr (FUSED_OK) {
      const NetSlot& ns = nets[slot];
      tc::TcArgs T;
      T.img = (const unsigned char*)ns.dev.tc_img; T.bias = ns.dev.tc_bias; T.nlayers = ns.dev.k + 2; T.k0_steps = (ns.dev.in + 15) / 16;
      T.A = ns.dev.A; T.NH = tc::head_n(ns.dev.A); T.in = ns.dev.in; T.dbg = g_tc_dbg;
      SegParams S; S.off = 0; S.len = (int)L; S.ply = ply; S.training = training; S.seed = seed; S.cpuct = cpuct; S.pad = 0;
      // games per CTA: spread the live games over all SMs, at least fused_min_gpc and at most 256 (two 128-row tiles)
      int gpc = (int)((L + num_sms - 1) / num_sms);
      gpc = (gpc + 7) / 8 * 8;
      if (gpc < fused_min_gpc) gpc = fused_min_gpc;
      if (gpc > 256) gpc = 256;
      const int grid = (int)((L + gpc - 1) / gpc);
      const int fmt = tc_fmt();
      launch(K_OTHER, [&] {
        if (gpc <= 32) {
          // the tail of a generation: the small-batch kernel in the swapped orientation (weights resident in tensor memory, node cache).
          // (33..64 games per CTA: the ordinary one-tile kernel is 8 % faster since its A operand moved to tensor memory.)
          typedef fused::FCfg<G, 1, 1> C1;
          if (fmt == 0) fused::ply_kernel<G, 0, 1, 1><<<grid, C1::THREADS, C1::SMEM, stream>>>(P, T, S, visits, gpc);
          else fused::ply_kernel<G, 1, 1, 1><<<grid, C1::THREADS, C1::SMEM, stream>>>(P, T, S, visits, gpc);
        } else if (gpc <= 128) {
          // one 128-row tile per CTA, ordinary orientation, all 16 warps on it (node cache)
          typedef fused::FCfg<G, 1, 0> C1;
          if (fmt == 0) fused::ply_kernel<G, 0, 1, 0><<<grid, C1::THREADS, C1::SMEM, stream>>>(P, T, S, visits, gpc);
          else fused::ply_kernel<G, 1, 1, 0><<<grid, C1::THREADS, C1::SMEM, stream>>>(P, T, S, visits, gpc);
        } else {
          // two 128-row tiles per CTA, 8 warps each
          typedef fused::FCfg<G, 2, 0> C2;
          if (fmt == 0) fused::ply_kernel<G, 0, 2, 0><<<grid, C2::THREADS, C2::SMEM, stream>>>(P, T, S, visits, gpc);
          else fused::ply_kernel<G, 1, 2, 0><<<grid, C2::THREADS, C2::SMEM, stream>>>(P, T, S, visits, gpc);
        }
      });
      AG_CK(cudaGetLastError());
      last_cpuct = cpuct;
      return AGPU_OK;
    } else {
      (void)L; (void)slot; (void)visits; (void)training; (void)cpuct; (void)seed; (void)ply;
      err = "fused ply kernel not available for this game";
      return AGPU_ERR_INVALID;
    }
  }

  void drop_graphs() {
    for (auto& gs : graphs) for (auto e : gs.exec) if (e) cudaGraphExecDestroy(e);
    graphs.clear();
  }

  bool is_tc() const { return cfg.nn_mode == AGPU_NN_BF16_TC || cfg.nn_mode == AGPU_NN_FP16_TC; }
  int tc_fmt() const { return cfg.nn_mode == AGPU_NN_FP16_TC ? 1 : 0; }
  static int blocks_for_groups(int64_t L) { return (int)((L * Lay::W + Lay::SB - 1) / Lay::SB); }
  static int blocks_for_threads(int64_t n) { return (int)((n + 255) / 256); }

  NNInput nn_input_tree() const {
    NNInput I; I.tree = tree.p; I.game_stride = P.game_stride; I.rec = Lay::STATE_STRIDE; I.off_state = (int)Lay::state_off(R, 0); I.nc = G::Geo::NC; I.VS = G::VS;
    I.leaf = leaf.p; I.x_direct = nullptr; I.seg = nullptr;
    return I;
  }

  // ---- network ----
  int set_weights(int slot, const float* base, const float* const* res, const float* pol_w, const float* pol_b, const float* val_w,
                  const float* val_b) override {
    AG_REQUIRE(slot == 0 || slot == 1, AGPU_ERR_INVALID, "slot must be 0 or 1");
    AG_REQUIRE(base && pol_w && pol_b && val_w && val_b && (cfg.blocks == 0 || res), AGPU_ERR_INVALID, "null weight pointer");
    NetSlot& s = nets[slot];
    const int in = 2 * G::VS, n = cfg.width, k = cfg.blocks;
    AG_CK(cudaSetDevice(cfg.device));
    const NetDev before = s.dev;
    const bool was_set = s.set;
    AG_CK(s.base.ensure((size_t)n * in)); AG_CK(s.res.ensure((size_t)std::max(1, k) * n * n)); AG_CK(s.pol_w.ensure((size_t)A * n));
    AG_CK(s.pol_b.ensure(A)); AG_CK(s.val_w.ensure(n)); AG_CK(s.val_b.ensure(1));
    AG_CK(cudaMemcpyAsync(s.base.p, base, sizeof(float) * n * in, cudaMemcpyHostToDevice, stream));
    for (int l = 0; l < k; l++) {
      AG_REQUIRE(res[l] != nullptr, AGPU_ERR_INVALID, "null residual weight pointer");
      AG_CK(cudaMemcpyAsync(s.res.p + (size_t)l * n * n, res[l], sizeof(float) * n * n, cudaMemcpyHostToDevice, stream));
    }
    AG_CK(cudaMemcpyAsync(s.pol_w.p, pol_w, sizeof(float) * A * n, cudaMemcpyHostToDevice, stream));
    AG_CK(cudaMemcpyAsync(s.pol_b.p, pol_b, sizeof(float) * A, cudaMemcpyHostToDevice, stream));
    AG_CK(cudaMemcpyAsync(s.val_w.p, val_w, sizeof(float) * n, cudaMemcpyHostToDevice, stream));
    AG_CK(cudaMemcpyAsync(s.val_b.p, val_b, sizeof(float), cudaMemcpyHostToDevice, stream));
    s.dev.in = in; s.dev.n = n; s.dev.k = k; s.dev.A = A;
    s.dev.base = s.base.p; s.dev.res = s.res.p; s.dev.pol_w = s.pol_w.p; s.dev.pol_b = s.pol_b.p; s.dev.val_w = s.val_w.p; s.dev.val_b = s.val_b.p;
    s.dev.tc_img = nullptr; s.dev.tc_bias = nullptr;
    if (is_tc()) {
      const size_t bytes = tc_image_bytes(in, n, k, A);
      std::vector<unsigned char> img(bytes);
      std::vector<float> bias(256, 0.f);
      tc_build_image(base, res, pol_w, pol_b, val_w, val_b, in, n, k, A, img.data(), bias.data(), tc_fmt());
      AG_CK(s.tc_img.ensure(bytes)); AG_CK(s.tc_bias.ensure(256));
      AG_CK(cudaMemcpyAsync(s.tc_img.p, img.data(), bytes, cudaMemcpyHostToDevice, stream));
      AG_CK(cudaMemcpyAsync(s.tc_bias.p, bias.data(), 256 * sizeof(float), cudaMemcpyHostToDevice, stream));
      AG_CK(cudaStreamSynchronize(stream));
      s.dev.tc_img = s.tc_img.p; s.dev.tc_bias = s.tc_bias.p;
    }
    AG_CK(cudaStreamSynchronize(stream));
    s.set = true;
    if (!was_set || memcmp(&before, &s.dev, sizeof(NetDev)) != 0) drop_graphs();   // graphs bake the weight pointers
    return AGPU_OK;
  }

  cudaError_t nn_on(cudaStream_t st, int slot, const NNInput& I, int64_t L, float* out, int outs) {
    const NetSlot& s = nets[slot];
    if (is_tc()) return tc_forward(s.dev, I, (int)L, out, outs, st, tc_fmt());
    constexpr int GT = 8;
    const size_t smem = sizeof(float) * GT * (s.dev.in + s.dev.n);
    nn_fp32_kernel<GT><<<(int)((L + GT - 1) / GT), s.dev.n, smem, st>>>(s.dev, I, (int)L, out, outs);
    return cudaGetLastError();
  }
  int run_nn(int slot, const NNInput& I, int64_t L, float* out, int outs) {
    cudaError_t e = cudaSuccess;
    launch(K_NN, [&] { e = nn_on(stream, slot, I, L, out, outs); });
    AG_CK(e);
    return AGPU_OK;
  }

  // ---- segmented graph replay ----
  // records the rollout loop of slice i (capacity seg_cap slots) on its stream and instantiates it
  int build_graph(int i, int visits, int slot, cudaGraphExec_t* out) {
    cudaStream_t st = seg_stream[i];
    const SegParams* sp = seg_dev.p + i;
    NNInput I = nn_input_tree();
    I.seg = reinterpret_cast<const int*>(sp);
    const int gb = blocks_for_groups(seg_cap);
    AG_CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < visits && e == cudaSuccess; k++) {
      const int last = (k == visits - 1);
      if (k == 0) select_seg_kernel<G><<<gb, Lay::SB, 0, st>>>(P, sp, 0, last);
      else step_seg_kernel<G><<<gb, Lay::SB, 0, st>>>(P, sp, k, last);
      e = nn_on(st, slot, I, seg_cap, nn_out.p, Lay::OUTS);
      if (last) expand_seg_kernel<G><<<gb, Lay::SB, 0, st>>>(P, sp, 1);
    }
    cudaGraph_t graph = nullptr;
    cudaError_t e2 = cudaStreamEndCapture(st, &graph);
    AG_CK(e);
    AG_CK(e2);
    AG_CK(cudaGraphInstantiate(out, graph, 0));
    AG_CK(cudaGraphDestroy(graph));
    return AGPU_OK;
  }

  // enqueues one mcts_single over L live games as nseg concurrent slices; returns with the main stream ordered after all of them
  int enqueue_search_segmented(int64_t L, int slot, int visits, int training, float cpuct, uint64_t seed, uint32_t ply) {
    GraphSet* gs = nullptr;
    for (auto& g : graphs) if (g.visits == visits && g.slot == slot && g.training == training) gs = &g;
    if (!gs) {
      GraphSet n; n.visits = visits; n.slot = slot; n.training = training; n.exec.assign(nseg, nullptr);
      for (int i = 0; i < nseg; i++) { int rc = build_graph(i, visits, slot, &n.exec[i]); if (rc != AGPU_OK) return rc; }
      graphs.push_back(n);
      gs = &graphs.back();
    }
    int64_t per = ((L + nseg - 1) / nseg + 255) / 256 * 256;        // tile-aligned slices
    if (per > seg_cap) per = seg_cap;
    int used = 0;
    for (int i = 0; i < nseg; i++) {
      const int64_t off = (int64_t)i * per;
      const int64_t len = off >= L ? 0 : std::min<int64_t>(per, L - off);
      SegParams& S = seg_host[i];
      S.off = (int)off; S.len = (int)len; S.ply = ply; S.training = training; S.seed = seed; S.cpuct = cpuct; S.pad = 0;
      if (len > 0) used = i + 1;
    }
    AG_CK(cudaMemcpyAsync(seg_dev.p, seg_host, sizeof(SegParams) * nseg, cudaMemcpyHostToDevice, stream));
    AG_CK(cudaEventRecord(ev_fork, stream));
    for (int i = 0; i < used; i++) {
      AG_CK(cudaStreamWaitEvent(seg_stream[i], ev_fork, 0));
      AG_CK(cudaGraphLaunch(gs->exec[i], seg_stream[i]));
      AG_CK(cudaEventRecord(ev_join[i], seg_stream[i]));
      AG_CK(cudaStreamWaitEvent(stream, ev_join[i], 0));
      launch_count += 2 * visits + 1;
      kt.launches[K_SELECT] += visits; kt.launches[K_NN] += visits; kt.launches[K_EXPAND] += 1;
    }
    last_cpuct = cpuct;
    return AGPU_OK;
  }

  int forward(int slot, const float* x, int64_t L, float* logits, float* value) override {
    AG_REQUIRE((slot == 0 || slot == 1) && nets[slot].set, AGPU_ERR_STATE, "weights not set for this slot");
    AG_REQUIRE(L >= 1 && x, AGPU_ERR_INVALID, "bad arguments");
    AG_CK(cudaSetDevice(cfg.device));
    const int in = 2 * G::VS;
    DevBuf<float> dx, dout;
    AG_CK(dx.ensure((size_t)L * in)); AG_CK(dout.ensure((size_t)L * Lay::OUTS));
    AG_CK(cudaMemcpyAsync(dx.p, x, sizeof(float) * L * in, cudaMemcpyHostToDevice, stream));
    NNInput I = nn_input_tree(); I.tree = nullptr; I.leaf = nullptr; I.x_direct = dx.p;
    int rc = run_nn(slot, I, L, dout.p, Lay::OUTS);
    if (rc != AGPU_OK) { dx.release(); dout.release(); return rc; }
    std::vector<float> h((size_t)L * Lay::OUTS);
    AG_CK(cudaMemcpyAsync(h.data(), dout.p, sizeof(float) * h.size(), cudaMemcpyDeviceToHost, stream));
    AG_CK(cudaStreamSynchronize(stream));
    for (int64_t g = 0; g < L; g++) {
      if (logits) memcpy(logits + g * A, &h[g * Lay::OUTS], sizeof(float) * A);
      if (value) value[g] = h[g * Lay::OUTS + A];
    }
    dx.release(); dout.release();
    if (profiling) harvest();
    return AGPU_OK;
  }

  // ---- plugin surface ----
  int position_init(void* out, int64_t n) override {
    AG_REQUIRE(out && n >= 1, AGPU_ERR_INVALID, "bad arguments");
    AG_CK(cudaSetDevice(cfg.device));
    DevBuf<State> d; AG_CK(d.ensure(n));
    launch(K_OTHER, [&] { game_ops_kernel<G><<<blocks_for_threads(n), 256, 0, stream>>>(nullptr, nullptr, (int)n, d.p, nullptr, nullptr, nullptr, nullptr, 1); });
    std::vector<State> h(n);
    AG_CK(cudaMemcpyAsync(h.data(), d.p, sizeof(State) * n, cudaMemcpyDeviceToHost, stream));
    AG_CK(cudaStreamSynchronize(stream));
    for (int64_t i = 0; i < n; i++) to_wire<G>(h[i], (char*)out + i * G::WIRE_BYTES);
    d.release();
    return AGPU_OK;
  }

  int game_ops(const void* pos, const int32_t* actions, int64_t n, void* played, uint8_t* legal, uint8_t* over, int8_t* result,
               float* enc) override {
    AG_REQUIRE(pos && n >= 1, AGPU_ERR_INVALID, "bad arguments");
    AG_REQUIRE(!played || actions, AGPU_ERR_INVALID, "play needs actions");
    AG_CK(cudaSetDevice(cfg.device));
    std::vector<State> h(n);
    for (int64_t i = 0; i < n; i++) h[i] = from_wire<G>((const char*)pos + i * G::WIRE_BYTES);
    DevBuf<State> din, dout; DevBuf<int32_t> dact; DevBuf<uint8_t> dlegal, dover; DevBuf<int8_t> dres; DevBuf<float> denc;
    AG_CK(din.ensure(n));
    AG_CK(cudaMemcpyAsync(din.p, h.data(), sizeof(State) * n, cudaMemcpyHostToDevice, stream));
    if (played) {
      for (int64_t i = 0; i < n; i++) AG_REQUIRE(actions[i] >= 1 && actions[i] <= A, AGPU_ERR_INVALID, "action out of range");
      AG_CK(dout.ensure(n)); AG_CK(dact.ensure(n));
      AG_CK(cudaMemcpyAsync(dact.p, actions, sizeof(int32_t) * n, cudaMemcpyHostToDevice, stream));
    }
    if (legal) AG_CK(dlegal.ensure((size_t)n * A));
    if (over) { AG_CK(dover.ensure(n)); AG_CK(dres.ensure(n)); }
    if (enc) AG_CK(denc.ensure((size_t)n * 2 * G::VS));
    launch(K_OTHER, [&] { game_ops_kernel<G><<<blocks_for_threads(n), 256, 0, stream>>>(din.p, dact.p, (int)n, dout.p, dlegal.p, dover.p, dres.p, denc.p, 0); });
    AG_CK(cudaGetLastError());
    std::vector<State> ho(played ? n : 0);
    if (played) AG_CK(cudaMemcpyAsync(ho.data(), dout.p, sizeof(State) * n, cudaMemcpyDeviceToHost, stream));
    if (legal) AG_CK(cudaMemcpyAsync(legal, dlegal.p, (size_t)n * A, cudaMemcpyDeviceToHost, stream));
    if (over) { AG_CK(cudaMemcpyAsync(over, dover.p, n, cudaMemcpyDeviceToHost, stream)); AG_CK(cudaMemcpyAsync(result, dres.p, n, cudaMemcpyDeviceToHost, stream)); }
    if (enc) AG_CK(cudaMemcpyAsync(enc, denc.p, sizeof(float) * n * 2 * G::VS, cudaMemcpyDeviceToHost, stream));
    AG_CK(cudaStreamSynchronize(stream));
    if (played) for (int64_t i = 0; i < n; i++) to_wire<G>(ho[i], (char*)played + i * G::WIRE_BYTES);
    din.release(); dout.release(); dact.release(); dlegal.release(); dover.release(); dres.release(); denc.release();
    return AGPU_OK;
  }

  // ---- search seam ----
  int reinit(const void* positions, int64_t L, const uint32_t* uids) override {
    AG_REQUIRE(positions && L >= 1 && L <= L_cap, AGPU_ERR_INVALID, "L out of range");
    AG_CK(cudaSetDevice(cfg.device));
    std::vector<State> h(L);
    for (int64_t i = 0; i < L; i++) h[i] = from_wire<G>((const char*)positions + i * G::WIRE_BYTES);
    std::vector<uint32_t> u(L);
    for (int64_t i = 0; i < L; i++) u[i] = uids ? uids[i] : (uint32_t)i;
    AG_CK(cudaMemcpyAsync(st_a.p, h.data(), sizeof(State) * L, cudaMemcpyHostToDevice, stream));
    AG_CK(cudaMemcpyAsync(uid_b.p, u.data(), sizeof(uint32_t) * L, cudaMemcpyHostToDevice, stream));
    launch(K_BEGIN, [&] { root_reset_kernel<G><<<blocks_for_threads(L), 256, 0, stream>>>(P, (int)L, st_a.p, uid_b.p); });
    AG_CK(cudaGetLastError());
    AG_CK(cudaStreamSynchronize(stream));
    L_live = L;
    return AGPU_OK;
  }

  int search_begin(int64_t L) override {
    AG_REQUIRE(L >= 1 && L <= L_live, AGPU_ERR_STATE, "search_begin: L exceeds the games installed by reinit");
    AG_CK(cudaSetDevice(cfg.device));
    launch(K_BEGIN, [&] { root_reset_kernel<G><<<blocks_for_threads(L), 256, 0, stream>>>(P, (int)L, nullptr, nullptr); });
    AG_CK(cudaGetLastError());
    AG_CK(cudaStreamSynchronize(stream));
    return AGPU_OK;
  }

  int upload_prob(const float* prob, int64_t L, int visits, const float** dev) {
    *dev = nullptr;
    if (!prob) return AGPU_OK;
    const size_t n = (size_t)visits * L * G::MAXLEN;
    AG_CK(d_prob.ensure(n));
    AG_CK(cudaMemcpyAsync(d_prob.p, prob, sizeof(float) * n, cudaMemcpyHostToDevice, stream));
    *dev = d_prob.p;
    return AGPU_OK;
  }

  void launch_select(int64_t L, int rollout, int last, float cpuct, const float* dprob, uint64_t seed, uint32_t ply) {
    last_cpuct = cpuct;
    launch(K_SELECT, [&] { select_kernel<G><<<blocks_for_groups(L), Lay::SB, 0, stream>>>(P, (int)L, rollout, last, cpuct, dprob, seed, ply); });
  }
  void launch_expand(int64_t L, int training, int last, const float* dprior, const float* dvalue) {
    if (dprior) launch(K_EXPAND, [&] { expand_backup_kernel<G, true><<<blocks_for_groups(L), Lay::SB, 0, stream>>>(P, (int)L, training, last, dprior, dvalue, last_cpuct); });
    else launch(K_EXPAND, [&] { expand_backup_kernel<G, false><<<blocks_for_groups(L), Lay::SB, 0, stream>>>(P, (int)L, training, last, nullptr, nullptr, last_cpuct); });
  }

  void launch_step(int64_t L, int rollout, int last, int training, float cpuct, uint64_t seed, uint32_t ply) {
    last_cpuct = cpuct;
    launch(K_SELECT, [&] { step_kernel<G><<<blocks_for_groups(L), Lay::SB, 0, stream>>>(P, (int)L, rollout, last, training, cpuct, seed, ply); });
  }

  // the rollout loop of mcts_single (mcts_gpu.jl:396-439), no host synchronisation inside.  With the in-kernel RNG the
  // search side is one launch per rollout: [expand+backUp of rollout k-1 | descent of rollout k].
  int enqueue_search(int64_t L, int slot, int visits, int training, float cpuct, const float* dprob, uint64_t seed, uint32_t ply) {
    NNInput I = nn_input_tree();
    for (int k = 0; k < visits; k++) {
      const int last = (k == visits - 1);
      if (dprob) {
        launch_select(L, k, last, cpuct, dprob, seed, ply);
      } else {
        if (k == 0) launch_select(L, 0, last, cpuct, nullptr, seed, ply);
        else launch_step(L, k, last, training, cpuct, seed, ply);
      }
      int rc = run_nn(slot, I, L, nn_out.p, Lay::OUTS);
      if (rc != AGPU_OK) return rc;
      if (dprob || last) launch_expand(L, training, last, nullptr, nullptr);
    }
    AG_CK(cudaGetLastError());
    return AGPU_OK;
  }

  int search(int64_t L, int slot, int visits, int training, float cpuct, const float* prob, uint64_t seed, uint32_t ply) override {
    AG_REQUIRE((slot == 0 || slot == 1) && nets[slot].set, AGPU_ERR_STATE, "weights not set for this slot");
    AG_REQUIRE(L >= 1 && L <= L_live, AGPU_ERR_STATE, "search: L exceeds the games installed by reinit");
    AG_REQUIRE(visits >= 1 && visits <= R, AGPU_ERR_INVALID, "visits exceeds the rollouts the context was created with");
    AG_CK(cudaSetDevice(cfg.device));
    const float* dprob;
    int rc = upload_prob(prob, L, visits, &dprob);
    if (rc != AGPU_OK) return rc;
    NvtxRange nv_search("agpu_search (mcts_single)");
    launch(K_BEGIN, [&] { root_reset_kernel<G><<<blocks_for_threads(L), 256, 0, stream>>>(P, (int)L, nullptr, nullptr); });
    rc = (use_fused && !dprob) ? enqueue_search_fused(L, slot, visits, training, cpuct, seed, ply)
         : (dprob || profiling || nseg <= 1) ? enqueue_search(L, slot, visits, training, cpuct, dprob, seed, ply)
                                             : enqueue_search_segmented(L, slot, visits, training, cpuct, seed, ply);
    if (rc != AGPU_OK) return rc;
    AG_CK(cudaStreamSynchronize(stream));
    if (profiling) harvest();
    return AGPU_OK;
  }

  int get_roots(int64_t L, float* pol, float* batch) override {
    AG_REQUIRE(L >= 1 && L <= L_live, AGPU_ERR_STATE, "L exceeds live games");
    AG_CK(cudaSetDevice(cfg.device));
    if (pol) AG_CK(cudaMemcpyAsync(pol, policy_final.p, sizeof(float) * L * A, cudaMemcpyDeviceToHost, stream));
    if (batch) {
      const size_t n = (size_t)L * 2 * G::VS;
      AG_CK(d_fscratch.ensure(n));
      launch(K_OTHER, [&] { encode_nodes_kernel<G><<<blocks_for_threads(n), 256, 0, stream>>>(P, (int)L, 0, d_fscratch.p); });
      AG_CK(cudaMemcpyAsync(batch, d_fscratch.p, sizeof(float) * n, cudaMemcpyDeviceToHost, stream));
    }
    AG_CK(cudaStreamSynchronize(stream));
    return AGPU_OK;
  }

  int select(int64_t L, int rollout, int last, float cpuct, const float* prob, uint64_t seed, uint32_t ply) override {
    AG_REQUIRE(L >= 1 && L <= L_live, AGPU_ERR_STATE, "L exceeds live games");
    AG_REQUIRE(rollout >= 0 && rollout < R, AGPU_ERR_INVALID, "rollout out of range");
    AG_CK(cudaSetDevice(cfg.device));
    // prob here is the slice for this rollout only: [L][maxLen]
    const float* dprob = nullptr;
    if (prob) {
      const size_t n = (size_t)L * G::MAXLEN;
      AG_CK(d_prob.ensure(n));
      AG_CK(cudaMemcpyAsync(d_prob.p, prob, sizeof(float) * n, cudaMemcpyHostToDevice, stream));
      dprob = d_prob.p;
    }
    // the kernel indexes prob by (rollout*L + g): pass rollout 0 for the slice, keep the RNG counter separately
    last_cpuct = cpuct;
    if (dprob) launch(K_SELECT, [&] { select_kernel<G><<<blocks_for_groups(L), Lay::SB, 0, stream>>>(P, (int)L, 0, last, cpuct, dprob, seed, ply); });
    else launch_select(L, rollout, last, cpuct, nullptr, seed, ply);
    AG_CK(cudaGetLastError());
    AG_CK(cudaStreamSynchronize(stream));
    if (profiling) harvest();
    return AGPU_OK;
  }

  int get_leaves(int64_t L, int32_t* leaf_out, float* batch) override {
    AG_REQUIRE(L >= 1 && L <= L_live, AGPU_ERR_STATE, "L exceeds live games");
    AG_CK(cudaSetDevice(cfg.device));
    if (leaf_out) {
      AG_CK(cudaMemcpyAsync(leaf_out, leaf.p, sizeof(int32_t) * L, cudaMemcpyDeviceToHost, stream));
    }
    if (batch) {
      const size_t n = (size_t)L * 2 * G::VS;
      AG_CK(d_fscratch.ensure(n));
      launch(K_OTHER, [&] { encode_nodes_kernel<G><<<blocks_for_threads(n), 256, 0, stream>>>(P, (int)L, 1, d_fscratch.p); });
      AG_CK(cudaMemcpyAsync(batch, d_fscratch.p, sizeof(float) * n, cudaMemcpyDeviceToHost, stream));
    }
    AG_CK(cudaStreamSynchronize(stream));
    if (leaf_out) for (int64_t i = 0; i < L; i++) leaf_out[i] += 1;   // 1-based over the ABI
    return AGPU_OK;
  }

  int eval(int64_t L, int slot, float* logits, float* value) override {
    AG_REQUIRE((slot == 0 || slot == 1) && nets[slot].set, AGPU_ERR_STATE, "weights not set for this slot");
    AG_REQUIRE(L >= 1 && L <= L_live, AGPU_ERR_STATE, "L exceeds live games");
    AG_CK(cudaSetDevice(cfg.device));
    int rc = run_nn(slot, nn_input_tree(), L, nn_out.p, Lay::OUTS);
    if (rc != AGPU_OK) return rc;
    if (logits || value) {
      std::vector<float> h((size_t)L * Lay::OUTS);
      AG_CK(cudaMemcpyAsync(h.data(), nn_out.p, sizeof(float) * h.size(), cudaMemcpyDeviceToHost, stream));
      AG_CK(cudaStreamSynchronize(stream));
      for (int64_t g = 0; g < L; g++) {
        if (logits) memcpy(logits + g * A, &h[g * Lay::OUTS], sizeof(float) * A);
        if (value) value[g] = h[g * Lay::OUTS + A];
      }
    }
    AG_CK(cudaStreamSynchronize(stream));
    if (profiling) harvest();
    return AGPU_OK;
  }

  int expand_backup(int64_t L, int training, int last, const float* prior, const float* value) override {
    AG_REQUIRE(L >= 1 && L <= L_live, AGPU_ERR_STATE, "L exceeds live games");
    AG_REQUIRE((prior == nullptr) == (value == nullptr), AGPU_ERR_INVALID, "prior and value must be given together");
    AG_CK(cudaSetDevice(cfg.device));
    const float *dp = nullptr, *dv = nullptr;
    if (prior) {
      AG_CK(d_prior.ensure((size_t)L * A)); AG_CK(d_value.ensure(L));
      AG_CK(cudaMemcpyAsync(d_prior.p, prior, sizeof(float) * L * A, cudaMemcpyHostToDevice, stream));
      AG_CK(cudaMemcpyAsync(d_value.p, value, sizeof(float) * L, cudaMemcpyHostToDevice, stream));
      dp = d_prior.p; dv = d_value.p;
    }
    launch_expand(L, training, last, dp, dv);
    AG_CK(cudaGetLastError());
    AG_CK(cudaStreamSynchronize(stream));
    if (profiling) harvest();
    return AGPU_OK;
  }

  int get_tree(int64_t L, agpu_tree_dump* out) override {
    AG_REQUIRE(out && L >= 1 && L <= L_live, AGPU_ERR_INVALID, "bad arguments");
    AG_CK(cudaSetDevice(cfg.device));
    std::vector<char> h((size_t)L * Lay::game_bytes(R));
    std::vector<int32_t> nn(L);
    AG_CK(cudaMemcpyAsync(h.data(), tree.p, h.size(), cudaMemcpyDeviceToHost, stream));
    AG_CK(cudaMemcpyAsync(nn.data(), nnodes.p, sizeof(int32_t) * L, cudaMemcpyDeviceToHost, stream));
    AG_CK(cudaStreamSynchronize(stream));
    for (int64_t g = 0; g < L; g++) {
      if (out->nnodes) out->nnodes[g] = nn[g];
      for (int nd = 0; nd < R; nd++) {
        const size_t o = (size_t)g * R + nd;
        const char* gbase_h = h.data() + (size_t)g * Lay::game_bytes(R);
        const char* rec = gbase_h + (size_t)nd * Lay::REC;
        const bool live = nd < nn[g];
        NodeHdr hd; memcpy(&hd, rec + Lay::OFF_HDR, sizeof(hd));
        if (out->parent) out->parent[o] = live ? hd.parent : 0;
        if (out->action) out->action[o] = live ? hd.action : 0;
        if (out->nchild) out->nchild[o] = live ? hd.nchild : 0;
        if (out->expanded) out->expanded[o] = live ? ((hd.flags & F_EXPANDED) ? 1 : 0) : 0;
        if (out->states) {
          if (live) { State st; memcpy(&st, gbase_h + Lay::state_off(R, nd), sizeof(st)); to_wire<G>(st, (char*)out->states + o * G::WIRE_BYTES); }
          else memset((char*)out->states + o * G::WIRE_BYTES, 0, G::WIRE_BYTES);
        }
        // prior is only defined once a node was expanded (the reference zero-fills; terminal/unexpanded nodes report 0)
        const bool has_prior = live && (hd.flags & F_EXPANDED);
        for (int a = 0; a < A; a++) {
          const size_t oa = o * A + a;
          float f; uint16_t v16; uint8_t b8;
          if (out->prior) { memcpy(&f, rec + Lay::OFF_PRIOR + 4 * a, 4); out->prior[oa] = has_prior ? f : 0.f; }
          if (out->q) { memcpy(&f, rec + Lay::OFF_Q + 4 * a, 4); out->q[oa] = live ? f : 0.f; }
          if (out->visits) {
            if (Lay::VIS_IN_PAD) { memcpy(&b8, rec + Lay::vis_byte(a), 1); v16 = b8; }
            else memcpy(&v16, rec + Lay::OFF_VIS + 2 * a, 2);
            out->visits[oa] = live ? (float)v16 : 0.f;
          }
          if (out->child) { memcpy(&b8, rec + Lay::OFF_CHILD + a, 1); out->child[oa] = live ? b8 : 0; }
          if (out->order) {
            if (Lay::ORD_IN_HDR) { uint64_t w; memcpy(&w, rec + Lay::OFF_HDR, 8); b8 = a < 7 ? (uint8_t)((w >> (40 + 3 * a)) & 7u) : 0; }
            else memcpy(&b8, rec + Lay::OFF_ORDER + a, 1);
            out->order[oa] = (live && a < hd.nchild) ? b8 : 0;
          }
        }
      }
    }
    return AGPU_OK;
  }

  // ---- the self-play loop (mcts_gpu.jl:477-579) and the duel loop (:581-651), all plies on the device ----
  int selfplay(int slot, int visits, int64_t ngames, uint32_t uid_base, float cpuct, uint64_t seed, agpu_samples* samples,
               int64_t results[3], agpu_run_stats* stats, bool duel, int slot_b) override {
    AG_REQUIRE((slot == 0 || slot == 1) && nets[slot].set, AGPU_ERR_STATE, "weights not set for this slot");
    if (duel) AG_REQUIRE((slot_b == 0 || slot_b == 1) && nets[slot_b].set, AGPU_ERR_STATE, "weights not set for the second slot");
    AG_REQUIRE(ngames >= 1 && ngames <= L_cap, AGPU_ERR_INVALID, "ngames exceeds max_games");
    AG_REQUIRE(visits >= 1 && visits <= R, AGPU_ERR_INVALID, "visits exceeds the rollouts the context was created with");
    AG_REQUIRE(results != nullptr, AGPU_ERR_INVALID, "results is null");
    AG_CK(cudaSetDevice(cfg.device));
    const long long cap = duel ? 0 : (long long)ngames * G::MAXLEN;
    SampleBufs S{};
    if (!duel) {
      AG_CK(s_state.ensure((size_t)cap * 2 * G::VS)); AG_CK(s_policy.ensure((size_t)cap * A)); AG_CK(s_player.ensure(cap));
      AG_CK(s_value.ensure(cap)); AG_CK(s_fstate.ensure((size_t)cap * G::FS)); AG_CK(s_game.ensure(cap)); AG_CK(s_ply.ensure(cap));
      S.state = s_state.p; S.policy = s_policy.p; S.player = s_player.p; S.value = s_value.p; S.fstate = s_fstate.p; S.game = s_game.p; S.ply = s_ply.p;
      S.capacity = cap;
    }
    AG_CK(game_result.ensure(ngames)); AG_CK(game_final.ensure(ngames));
    PlyState Y; Y.next_state = st_a.p; Y.alive = alive.p; Y.block_count = block_count.p; Y.game_result = game_result.p; Y.game_final = game_final.p;
    Y.tallies = tallies.p;

    struct EvPair { cudaEvent_t a = nullptr, b = nullptr; ~EvPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); } } evp;
    AG_CK(cudaEventCreate(&evp.a)); AG_CK(cudaEventCreate(&evp.b));
    const cudaEvent_t ev0 = evp.a, ev1 = evp.b;
    const int64_t launches0 = launch_count;
    AG_CK(cudaEventRecord(ev0, stream));
    AG_CK(cudaMemsetAsync(tallies.p, 0, 8 * sizeof(unsigned long long), stream));
    // positions = [Position() for k in 1:ngames]; init(positions, visits)   (:479-481)
    launch(K_OTHER, [&] { game_ops_kernel<G><<<blocks_for_threads(ngames), 256, 0, stream>>>(nullptr, nullptr, (int)ngames, st_b.p, nullptr, nullptr, nullptr, nullptr, 1); });
    {
      std::vector<uint32_t> u(ngames);
      for (int64_t i = 0; i < ngames; i++) u[i] = uid_base + (uint32_t)i;
      AG_CK(cudaMemcpyAsync(uid_b.p, u.data(), sizeof(uint32_t) * ngames, cudaMemcpyHostToDevice, stream));
      AG_CK(cudaStreamSynchronize(stream));
    }
    launch(K_BEGIN, [&] { root_reset_kernel<G><<<blocks_for_threads(ngames), 256, 0, stream>>>(P, (int)ngames, st_b.p, uid_b.p); });
    L_live = ngames;
    int64_t L = ngames, sims = 0, npos = 0, count = 0;
    uint32_t round = 0;
    bool aborted = false;
    const bool trace_plies = getenv("AGPU_TRACE_PLIES") != nullptr;   // development: per-ply wall time on stderr
    double t_prev = 0;
    auto now_ms = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    if (trace_plies) t_prev = now_ms();
    const bool streaming = stream_samples && !duel && samples != nullptr && samples->capacity > 0 && host_pinned(samples->state) &&
                           host_pinned(samples->policy) && host_pinned(samples->player);
    if (streaming) AG_REQUIRE(samples->state && samples->policy && samples->player && samples->value && samples->fstate, AGPU_ERR_INVALID, "null sample array");
    NvtxRange nv_loop(duel ? "agpu_duel" : "agpu_selfplay");
    while (L > 0) {
      char nv_name[48];
      snprintf(nv_name, sizeof(nv_name), "ply %u (%lld games)", round, (long long)L);
      NvtxRange nv_ply(nv_name);
      const int actor = duel ? ((round % 2 == 0) ? slot : slot_b) : slot;                  // :592-596
      int rc = use_fused ? enqueue_search_fused(L, actor, visits, duel ? 0 : 1, cpuct, seed, round)                     // mcts_single (:503, :599)
               : (profiling || nseg <= 1) ? enqueue_search(L, actor, visits, duel ? 0 : 1, cpuct, nullptr, seed, round)
                                          : enqueue_search_segmented(L, actor, visits, duel ? 0 : 1, cpuct, seed, round);
      if (rc != AGPU_OK) return rc;
      sims += L * visits; npos += L;
      const int nb = blocks_for_threads(L);
      nvtxRangePushA("move + compaction");
      if (duel) launch(K_FINISH, [&] { finish_ply_kernel<G, true><<<nb, 256, 0, stream>>>(P, (int)L, round, seed, uid_base, S, count, Y); });
      else launch(K_FINISH, [&] { finish_ply_kernel<G, false><<<nb, 256, 0, stream>>>(P, (int)L, round, seed, uid_base, S, count, Y); });
      launch(K_COMPACT, [&] { scan_blocks_kernel<<<1, 1024, 0, stream>>>(block_count.p, nb, total_dev.p); });
      launch(K_COMPACT, [&] { compact_kernel<G><<<nb, 256, 0, stream>>>(P, (int)L, Y, st_b.p, uid_b.p); });
      AG_CK(cudaMemcpyAsync(total_host, total_dev.p, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
      AG_CK(cudaMemcpyAsync(fault_host, tallies.p + 4, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
      AG_CK(cudaStreamSynchronize(stream));                                                  // one host sync per ply (the reference: 6·R+3)
      nvtxRangePop();
      if (streaming) {                                                                       // this ply's rows [count, count + L) are final
        const long long lo = count, hi = std::min<long long>(count + L, std::min<long long>(samples->capacity, cap));
        if (hi > lo) {
          const size_t n = (size_t)(hi - lo);
          AG_CK(cudaMemcpyAsync(samples->state + lo * 2 * G::VS, s_state.p + lo * 2 * G::VS, n * 2 * G::VS, cudaMemcpyDeviceToHost, copy_stream));
          AG_CK(cudaMemcpyAsync(samples->policy + lo * A, s_policy.p + lo * A, sizeof(float) * n * A, cudaMemcpyDeviceToHost, copy_stream));
          AG_CK(cudaMemcpyAsync(samples->player + lo, s_player.p + lo, n, cudaMemcpyDeviceToHost, copy_stream));
          if (samples->game) AG_CK(cudaMemcpyAsync(samples->game + lo, s_game.p + lo, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, copy_stream));
          if (samples->ply) AG_CK(cudaMemcpyAsync(samples->ply + lo, s_ply.p + lo, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, copy_stream));
        }
      }
      if (trace_plies) { const double t = now_ms(); fprintf(stderr, "ply %u L %lld ms %.3f us/rollout %.2f\n", round, (long long)L, t - t_prev, 1e3 * (t - t_prev) / visits); t_prev = t; }
      count += L;
      L = *total_host;
      round += 1;
      // "faute" (mcts_gpu.jl:526-529, :611-614): the reference stops the generation at the first illegal move and returns valid=false.
      // Here the ply that produced it is completed (all its games in one launch) and the loop ends; the call returns
      // AGPU_ERR_ILLEGAL_MOVE with the samples pushed so far.  Without this a degenerate policy (NaN weights) never terminates:
      // an illegal Connect4 move leaves the position unchanged.  The ply cap is a second guard for plug-in games.
      if (*fault_host != 0 || round > (uint32_t)G::MAXLEN + 8) { aborted = true; break; }
      if (L > 0) launch(K_BEGIN, [&] { root_reset_kernel<G><<<blocks_for_threads(L), 256, 0, stream>>>(P, (int)L, st_b.p, uid_b.p); });   // re_init (:557-561)
      L_live = L;
    }
    if (!duel && count > 0)
      launch(K_FINALIZE, [&] { finalize_samples_kernel<G><<<(int)((count + 255) / 256), 256, 0, stream>>>(S, count, uid_base, game_result.p, game_final.p); });
    AG_CK(cudaMemcpyAsync(tallies_host, tallies.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    AG_CK(cudaEventRecord(ev1, stream));
    AG_CK(cudaStreamSynchronize(stream));
    AG_CK(cudaGetLastError());
    float ms = 0.f;
    AG_CK(cudaEventElapsedTime(&ms, ev0, ev1));
    results[0] = (int64_t)tallies_host[0]; results[1] = (int64_t)tallies_host[1]; results[2] = (int64_t)tallies_host[2];
    if (stats) {
      stats->sims = sims; stats->positions = npos; stats->plies = round; stats->total_length = (int64_t)tallies_host[3];
      stats->faults = (int64_t)tallies_host[4]; stats->kernel_launches = launch_count - launches0; stats->device_ms = ms; stats->search_ms = 0.0;
    }
    last_sample_count = duel ? 0 : std::min<long long>(count, cap);
    if (samples) {
      samples->count = count;
      const long long rows = std::min<long long>(count, std::min<long long>(samples->capacity, cap));
      if (rows > 0) {
        AG_REQUIRE(samples->state && samples->policy && samples->player && samples->value && samples->fstate, AGPU_ERR_INVALID, "null sample array");
        if (!streaming) {
          AG_CK(cudaMemcpyAsync(samples->state, s_state.p, (size_t)rows * 2 * G::VS, cudaMemcpyDeviceToHost, stream));
          AG_CK(cudaMemcpyAsync(samples->policy, s_policy.p, sizeof(float) * rows * A, cudaMemcpyDeviceToHost, stream));
          AG_CK(cudaMemcpyAsync(samples->player, s_player.p, rows, cudaMemcpyDeviceToHost, stream));
        }
        AG_CK(cudaMemcpyAsync(samples->value, s_value.p, sizeof(float) * rows, cudaMemcpyDeviceToHost, stream));
        AG_CK(cudaMemcpyAsync(samples->fstate, s_fstate.p, (size_t)rows * G::FS, cudaMemcpyDeviceToHost, stream));
        if (!streaming) {
          if (samples->game) AG_CK(cudaMemcpyAsync(samples->game, s_game.p, sizeof(int32_t) * rows, cudaMemcpyDeviceToHost, stream));
          if (samples->ply) AG_CK(cudaMemcpyAsync(samples->ply, s_ply.p, sizeof(int32_t) * rows, cudaMemcpyDeviceToHost, stream));
        }
        AG_CK(cudaStreamSynchronize(stream));
        if (streaming) AG_CK(cudaStreamSynchronize(copy_stream));
      }
    }
    if (profiling) harvest();
    if (aborted || tallies_host[4] != 0) {
      err = tallies_host[4] != 0 ? "illegal move chosen from the search policy (\"faute\", mcts_gpu.jl:526-529): generation stopped"
                                 : "ply cap exceeded: games do not terminate";
      return AGPU_ERR_ILLEGAL_MOVE;
    }
    return AGPU_OK;
  }

  // the rows of the last self-play run, still resident in the device sample arrays, copied to out's arrays starting at row_offset
  // (agpu_multi_selfplay: every device's block lands at its place in the caller's buffers over that device's own PCIe link)
  int fetch_samples(agpu_samples* out, int64_t row_offset) override {
    AG_REQUIRE(out && row_offset >= 0, AGPU_ERR_INVALID, "bad arguments");
    AG_CK(cudaSetDevice(cfg.device));
    const long long rows = std::min<long long>(last_sample_count, out->capacity - row_offset);
    if (rows <= 0) return AGPU_OK;
    AG_REQUIRE(out->state && out->policy && out->player && out->value && out->fstate, AGPU_ERR_INVALID, "null sample array");
    const long long o = row_offset;
    AG_CK(cudaMemcpyAsync(out->state + o * 2 * G::VS, s_state.p, (size_t)rows * 2 * G::VS, cudaMemcpyDeviceToHost, stream));
    AG_CK(cudaMemcpyAsync(out->policy + o * A, s_policy.p, sizeof(float) * rows * A, cudaMemcpyDeviceToHost, stream));
    AG_CK(cudaMemcpyAsync(out->player + o, s_player.p, rows, cudaMemcpyDeviceToHost, stream));
    AG_CK(cudaMemcpyAsync(out->value + o, s_value.p, sizeof(float) * rows, cudaMemcpyDeviceToHost, stream));
    AG_CK(cudaMemcpyAsync(out->fstate + o * G::FS, s_fstate.p, (size_t)rows * G::FS, cudaMemcpyDeviceToHost, stream));
    if (out->game) AG_CK(cudaMemcpyAsync(out->game + o, s_game.p, sizeof(int32_t) * rows, cudaMemcpyDeviceToHost, stream));
    if (out->ply) AG_CK(cudaMemcpyAsync(out->ply + o, s_ply.p, sizeof(int32_t) * rows, cudaMemcpyDeviceToHost, stream));
    AG_CK(cudaStreamSynchronize(stream));
    return AGPU_OK;
  }

  int64_t last_samples() const override { return last_sample_count; }

  int profile(int enable) override {
    AG_CK(cudaSetDevice(cfg.device));
    harvest();
    profiling = enable != 0;
    P.counters = profiling ? counters.p : nullptr;
    return AGPU_OK;
  }
  int kernel_times(agpu_kernel_times* out, int reset) override {
    AG_REQUIRE(out != nullptr, AGPU_ERR_INVALID, "null out");
    AG_CK(cudaSetDevice(cfg.device));
    harvest();
    *out = kt;
    if (reset) {
      memset(&kt, 0, sizeof(kt));
      AG_CK(cudaMemsetAsync(counters.p, 0, 2 * sizeof(unsigned long long), stream));
      AG_CK(cudaStreamSynchronize(stream));
    }
    return AGPU_OK;
  }
  int layout_info(int64_t* node_bytes, int64_t* game_bytes, int64_t* lanes) override {
    if (node_bytes) *node_bytes = Lay::REC;
    if (game_bytes) *game_bytes = (int64_t)Lay::game_bytes(R);
    if (lanes) *lanes = Lay::W;
    return AGPU_OK;
  }
  int debug_expf(const float* x, int64_t n, float* y, int sigmoid) override {
    AG_CK(cudaSetDevice(cfg.device));
    DevBuf<float> dx, dy;
    AG_CK(dx.ensure(n)); AG_CK(dy.ensure(n));
    AG_CK(cudaMemcpyAsync(dx.p, x, sizeof(float) * n, cudaMemcpyHostToDevice, stream));
    debug_expf_kernel<<<(int)((n + 255) / 256), 256, 0, stream>>>(dx.p, n, dy.p, sigmoid);
    AG_CK(cudaMemcpyAsync(y, dy.p, sizeof(float) * n, cudaMemcpyDeviceToHost, stream));
    AG_CK(cudaStreamSynchronize(stream));
    dx.release(); dy.release();
    return AGPU_OK;
  }
};

// factories, one per translation unit
EngineBase* make_engine_connect4();
EngineBase* make_engine_gobang(int n, int nvict);
EngineBase* make_engine_hex(int n);
EngineBase* make_engine_reversi(int n);

}  // namespace ag
