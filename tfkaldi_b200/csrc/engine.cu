// C-ABI engine (include/tfkaldi_b200.h): owns parameters, optimizer state, activations and TMA plans
// for one GPU and strings the sm_100a kernels into the reference's train / eval / decode steps.
//
// Reference semantics restated here (paths relative to the reference repo):
//   layer  = matmul + bias -> [batch_norm] -> nonlin -> [dropout]      classifiers/layer.py:52-56,
//                                                                      classifiers/activation.py:22-42, nnet.py:42-72
//   output layer: identity, no BN / dropout                            classifiers/dnn.py:67-68, 108-109
//   loss   = SUM over frames of softmax-CE                             trainer.py:526-531
//   update = accumulate over micro-batches, /num_frames, clip, Adam    trainer.py:165-184
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/tfkaldi_b200.h"
#include "gemm.cuh"
#include "kernels.cuh"

using namespace tfk;

namespace {

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
template <typename T>
inline T* opt(T* base, size_t off) { return base ? base + off : nullptr; }

// ------------------------------------------------------------------ NCCL (resolved at run time)
struct NcclApi {
  typedef struct { char internal[128]; } UniqueId;
  int (*GetUniqueId)(UniqueId*) = nullptr;
  int (*CommInitRank)(void**, int, UniqueId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*ReduceScatter)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
NcclApi& nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  // Prefer the copy already mapped into the process (the one torch.distributed loaded).
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return api;
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(lib, "ncclAllReduce"));
  api.ReduceScatter = reinterpret_cast<decltype(api.ReduceScatter)>(dlsym(lib, "ncclReduceScatter"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(lib, "ncclAllGather"));
  api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(dlsym(lib, "ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(dlsym(lib, "ncclGroupEnd"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.ReduceScatter &&
           api.AllGather && api.GroupStart && api.GroupEnd;
  return api;
}
constexpr int kNcclFloat = 7, kNcclDouble = 8, kNcclBf16 = 9, kNcclSum = 0;

thread_local std::string g_create_error;

struct Layer {
  int K = 0, N = 0;   // in / out dimension
  int Kpad = 0;       // rows reserved for W in the arenas (K rounded up to 512)
  int ldk = 0;        // pitch (elements) of the bf16 input activations
  int ldn = 0;        // pitch of W rows, output activations, gradient rows: round_up(N, 8)
  int npad = 0;       // padded length of per-column vectors: round_up(N, 256)
  bool hidden = false, bn = false;
  size_t off_w = 0, off_b = 0, off_beta = 0;  // offsets (floats) into the parameter arenas
  size_t w_count = 0;                         // floats reserved for W (K*ldn rounded up to 1024)
  bool peer_reduce = false;                   // wgrad epilogue adds straight into the owner GPU's slice
  CUtensorMap* peer_tm = nullptr;             // device [nranks]: every rank's copy of this layer's gradient matrix
  float *moving_mean = nullptr, *moving_var = nullptr;  // [npad]
  float *bn_mean = nullptr, *bn_rstd = nullptr;          // statistics used by the last forward
  float* bn_sums = nullptr;                              // [2*ldn] backward column sums
  __nv_bfloat16 *z_hi = nullptr, *z_lo = nullptr;        // pre-BN linear output [maxB, ldn]
  __nv_bfloat16 *u_hi = nullptr, *u_lo = nullptr;        // L2Norm: output of the nonlinearity [maxB, ldn]
  float* l2_s = nullptr;                                 // L2Norm: per-frame mean square [maxB]
  uint32_t* maskbits = nullptr;  // [ceil(ldn/32), maxB] gradient-pass bits written by the forward epilogue
  float* db_part = nullptr;  // [maxB/32 + 8, ldn] bias-gradient partials written by the dgrad epilogue above
};

struct Plan {  // TMA descriptors for one (frames, active layers) shape; the work lists they point at live in list_cache
  std::vector<GemmParams> fwd_train, fwd_eval, bwd;
};

}  // namespace

struct tfk_handle {
  tfk_config cfg;
  int L = 0;        // hidden layers
  int active = 0;   // hidden layers in use (layer-wise growth)
  bool x3 = false;
  bool two_cta = true;  // cta_group::2 CTA-pair GEMM (TFK_GEMM_2CTA=0 selects the 1-CTA kernel)
  int num_sms = 148;
  std::vector<Layer> layers;  // L+1
  size_t arena_n = 0;
  size_t nW = 0;  // [0, nW): weight regions of all layers; [nW, arena_n): biases and batch-norm betas
  bool sharded = false;       // data parallel with reduce-scatter -> sharded Adam -> all-gather
  bool fused_rs = false;      // weight-gradient reduce-scatter fused into the wgrad epilogue (peer memory)
  std::vector<float*> peer_G; // [nranks] every rank's gradient arena (IPC-mapped; own pointer for self)
  std::vector<__nv_bfloat16*> peer_Sh, peer_Sl;  // every rank's bf16 operand arenas
  int* dp_flags = nullptr;                        // [TFK_DP_FLAG_WORDS] published epochs [slot][source rank] (peers write here)
  int** d_peer_flags = nullptr;                   // device array [nranks-1] of the OTHER ranks' flag arrays
  int xchg_stride = 0;                            // floats per rank slot of the exchange area behind the flag words
  bool fused_ag = false;                          // Adam stores the refreshed operands into every peer (no NCCL all-gather)
  int dp_epoch = 0;
  std::vector<void*> ipc_opened;
  bool params_synced = true;  // fp32 master weights / Adam slots identical on every rank
  float *P = nullptr, *G = nullptr, *M = nullptr, *V = nullptr;
  __nv_bfloat16 *Sh = nullptr, *Sl = nullptr;
  std::vector<__nv_bfloat16*> act_hi, act_lo;  // [L+1]: act[0] = input, act[l+1] = output of hidden l
  float* logits = nullptr;
  int ldo = 0, ldh = 0, ld0 = 0, ldmax = 0;
  __nv_bfloat16 *dzo_hi = nullptr, *dzo_lo = nullptr;
  __nv_bfloat16 *dA_hi[2] = {nullptr, nullptr}, *dA_lo[2] = {nullptr, nullptr};
  float* row_loss = nullptr;
  double* acc = nullptr;  // device {loss_sum, num_frames}
  double* acc_host = nullptr;
  bool loss_pending = false;  // a step's {loss, frames} copy into acc_host is queued and has not been read (tfk_last_loss)
  float *bn_ps = nullptr, *bn_pq = nullptr;
  float* ws = nullptr;
  float* ws_colsum = nullptr;
  unsigned int* bn_counters = nullptr;  // 512 self-resetting counters of the BN backward reduction
  float* tmp_f32 = nullptr;
  int* sched = nullptr;
  std::map<long long, Plan> plans;
  struct TileLists { int* d; int pairs, stride; };
  std::map<std::vector<int>, TileLists> list_cache;  // CTA-pair work lists per launch shape (finish_params)
  // trainer scalars
  long long global_step = 0;
  // Adam's own step count t (TF keeps beta1_power / beta2_power inside the optimizer: they are NOT in the
  // `train_variables` saver, so restore_trainer never rewinds them and init_op resets them; trainer.py:115,204-205)
  long long adam_step = 0;
  double lr_fact = 1.0;
  unsigned long long drop_seed = 0;
  // DP
  void* comm = nullptr;
  bool own_comm = false;
  int rank = 0, nranks = 1;
  cudaStream_t comm_stream = nullptr;
  cudaStream_t adam_stream = nullptr;      // tfk_train_step: per-layer Adam overlapped with the rest of the backward pass
  std::vector<cudaEvent_t> layer_events;
  std::vector<cudaEvent_t> colsum_events;  // tfk_train_step: column sums of layer l taken (side stream)
  int pending_accumulates = 0;             // tfk_accumulate calls since the last tfk_apply
  // data parallel: the host may queue at most `runahead` optimizer steps ahead of the device (TFK_DP_RUNAHEAD,
  // default 2, 0 = unbounded).  Ranks whose hosts run many steps ahead of each other keep NCCL work, peer stores
  // and flag waits of different steps in flight at once; bounding it costs nothing (the device stays 2 steps deep).
  int runahead = 2;
  std::vector<cudaEvent_t> step_events;
  long long steps_queued = 0;
  cudaEvent_t ev_compute = nullptr, ev_comm = nullptr;
  // timers
  bool timers_on = false;
  struct TimerRec { cudaEvent_t a, b; int cat; };
  std::vector<TimerRec> timer_recs;
  std::vector<cudaEvent_t> event_pool;
  double timer_ms[TFK_NUM_TIMERS] = {0};
  int64_t timer_launches[TFK_NUM_TIMERS] = {0};
  int64_t launches = 0;
  std::string err;
  std::vector<void*> allocs;
};

namespace {

int fail(tfk_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  else g_create_error = buf;
  return code;
}

#define TFK_CUDA(h, expr)                                                                       \
  do {                                                                                          \
    cudaError_t e_ = (expr);                                                                    \
    if (e_ != cudaSuccess)                                                                      \
      return fail(h, TFK_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define TFK_LAUNCH(h, expr)                                                                     \
  do {                                                                                          \
    int rc_ = (expr);                                                                           \
    if (rc_ != 0)                                                                               \
      return fail(h, TFK_ECUDA, "%s failed: %s (%s:%d)", #expr,                                 \
                  cudaGetErrorString(static_cast<cudaError_t>(rc_)), __FILE__, __LINE__);       \
  } while (0)
#define TFK_TRY(expr)          \
  do {                         \
    int rc_ = (expr);          \
    if (rc_ != TFK_OK) return rc_; \
  } while (0)

template <typename T>
int dev_alloc(tfk_handle* h, T** p, size_t count, bool zero = true) {
  void* q = nullptr;
  const size_t bytes = (count ? count : 1) * sizeof(T);
  TFK_CUDA(h, cudaMalloc(&q, bytes));
  h->allocs.push_back(q);
  if (zero) TFK_CUDA(h, cudaMemset(q, 0, bytes));
  *p = static_cast<T*>(q);
  return TFK_OK;
}

struct TimerScope {  // brackets one (or a few) launches with an event pair when timing is on
  tfk_handle* h;
  cudaStream_t st;
  int cat;
  cudaEvent_t a = nullptr, b = nullptr;
  TimerScope(tfk_handle* h_, cudaStream_t st_, int cat_, int nlaunch = 1) : h(h_), st(st_), cat(cat_) {
    h->launches += nlaunch;
    h->timer_launches[cat] += nlaunch;
    if (!h->timers_on) return;
    a = take();
    b = take();
    cudaEventRecord(a, st);
  }
  ~TimerScope() {
    if (!a) return;
    cudaEventRecord(b, st);
    h->timer_recs.push_back({a, b, cat});
  }
  cudaEvent_t take() {
    if (!h->event_pool.empty()) {
      cudaEvent_t e = h->event_pool.back();
      h->event_pool.pop_back();
      return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
};

int drain_timers(tfk_handle* h) {
  for (auto& r : h->timer_recs) {
    TFK_CUDA(h, cudaEventSynchronize(r.b));
    float ms = 0.f;
    TFK_CUDA(h, cudaEventElapsedTime(&ms, r.a, r.b));
    h->timer_ms[r.cat] += ms;
    h->event_pool.push_back(r.a);
    h->event_pool.push_back(r.b);
  }
  h->timer_recs.clear();
  return TFK_OK;
}

// ------------------------------------------------------------------ plans
void base_operands(tfk_handle* h, GemmSpec& s) { s.nsplit = h->x3 ? 3 : 1; }

// Tensor maps are rebuilt for every new frame count (host-only, microseconds).  The per-pair work lists are not:
// computing them (longest-first assignment + the half-tile search) costs ~2 ms per launch shape and uploading them a
// cudaMalloc + synchronous copy, which with utterance batches of a different length every step (the reference's
// normal case, trainer.py:276-307) was 40 ms per step against a 1.2 ms step.  The lists depend on the frame count
// only through the tile / k-block counts, so they are cached per launch SHAPE and owned by the handle.
int finish_params(tfk_handle* h, Plan& plan, const GemmSpec* s, int n, GemmParams* gp, const char* what, int l) {
  (void)plan;
  char err[256] = {0};
  if (gemm_build_params(s, n, h->sched, gp, err, sizeof(err), h->two_cta ? 1 : 0))
    return fail(h, TFK_ECUDA, "%s plan layer %d: %s", what, l, err);
  std::vector<int> key;
  if (gp->two_cta) {  // everything gemm_upload_tile_lists reads
    key = {gp->total_tiles, gp->nprob, gp->a_resident};
    for (int i = 0; i < gp->nprob; ++i) {
      const GemmProblem& p = gp->p[i];
      const int f[7] = {p.tile_begin, p.tiles_m, p.tiles_n, p.kb_per_split, p.num_kb, p.nsplit, p.N};
      key.insert(key.end(), f, f + 7);
    }
    auto it = h->list_cache.find(key);
    if (it != h->list_cache.end()) {
      gp->num_pairs = it->second.pairs;
      gp->list_stride = it->second.stride;
      gp->tile_list = it->second.d;
      return TFK_OK;
    }
  }
  int* d = nullptr;
  if (gemm_upload_tile_lists(gp, h->num_sms, &d, err, sizeof(err)))
    return fail(h, TFK_ECUDA, "%s plan layer %d: %s", what, l, err);
  if (d) h->list_cache[key] = {d, gp->num_pairs, gp->list_stride};
  return TFK_OK;
}

void free_plan(Plan& plan) { (void)plan; }  // plans hold host-side tensor maps only; device lists / peer maps are the handle's

int build_plan(tfk_handle* h, int B, Plan& plan) {
  const int L = h->L, act = h->active;
  const bool relu = h->cfg.nonlin == TFK_NONLIN_RELU;
  const bool smooth = h->cfg.nonlin == TFK_NONLIN_SIGMOID || h->cfg.nonlin == TFK_NONLIN_TANH;
  const int act_code = relu ? 1 : h->cfg.nonlin == TFK_NONLIN_SIGMOID ? 2 : h->cfg.nonlin == TFK_NONLIN_TANH ? 3 : 0;
  const bool drop = h->cfg.keep_prob < 1.0f;
  const bool l2 = h->cfg.l2_norm != 0;
  plan.fwd_train.assign(L + 1, GemmParams());
  plan.fwd_eval.assign(L + 1, GemmParams());
  plan.bwd.assign(L + 1, GemmParams());
  auto in_of = [&](int l) { return l == L ? act : l; };  // activation index feeding layer l
  for (int l = 0; l <= L; ++l) {
    if (l < L && l >= act) continue;
    const Layer& ly = h->layers[l];
    const int ai = in_of(l);
    for (int train = 0; train < 2; ++train) {
      GemmSpec s;
      base_operands(h, s);
      s.M = B; s.N = ly.N; s.K = ly.K;
      s.A_hi = h->act_hi[ai]; s.A_lo = h->act_lo[ai]; s.lda = ly.ldk; s.a_mn = 0;
      s.B_hi = h->Sh + ly.off_w; s.B_lo = opt(h->Sl, ly.off_w); s.ldb = ly.ldn; s.b_mn = 1;
      s.bias = h->P + ly.off_b;
      if (!ly.hidden) {
        s.out_kind = OUT_F32; s.D_hi = h->logits; s.ldd = h->ldo;
      } else if (ly.bn) {
        s.out_kind = h->x3 ? OUT_BF16_SPLIT : OUT_BF16;
        s.D_hi = ly.z_hi; s.D_lo = ly.z_lo; s.ldd = ly.ldn;
        if (train) { s.stat_sum = h->bn_ps; s.stat_sq = h->bn_pq; s.stat_ld = ly.npad; }
      } else {
        s.out_kind = h->x3 ? OUT_BF16_SPLIT : OUT_BF16;
        s.D_hi = h->act_hi[l + 1]; s.D_lo = h->act_lo[l + 1]; s.ldd = ly.ldn;
        s.act = act_code;
        if (l2) {  // the L2Norm kernel finishes the chain (normalise + dropout) from u
          s.D_hi = ly.u_hi; s.D_lo = ly.u_lo;
        }
        if (!l2 && train && drop) { s.keep = h->cfg.keep_prob; s.seed = 0; }
        if (!l2 && train && !smooth && (relu || drop)) {
          s.mask_bits_out = ly.maskbits; s.mask_bits_ld = h->cfg.max_frames; s.mask_nonzero = relu ? 0 : 1;
        }
      }
      GemmParams& gp = train ? plan.fwd_train[l] : plan.fwd_eval[l];
      TFK_TRY(finish_params(h, plan, &s, 1, &gp, "forward", l));
    }
    // backward: problem 0 = wgrad (long K first), problem 1 = dgrad into the layer below
    const __nv_bfloat16* dz_hi = ly.hidden ? h->dA_hi[(L - 1 - l) & 1] : h->dzo_hi;
    const __nv_bfloat16* dz_lo = ly.hidden ? h->dA_lo[(L - 1 - l) & 1] : h->dzo_lo;
    const int ldz = ly.hidden ? ly.ldn : h->ldo;
    GemmSpec s[2];
    base_operands(h, s[0]);
    s[0].M = ly.K; s[0].N = ly.N; s[0].K = B;
    s[0].A_hi = h->act_hi[ai]; s[0].A_lo = h->act_lo[ai]; s[0].lda = ly.ldk; s[0].a_mn = 1;
    s[0].B_hi = dz_hi; s[0].B_lo = dz_lo; s[0].ldb = ldz; s[0].b_mn = 1;
    s[0].out_kind = OUT_F32_REDADD; s[0].D_hi = h->G + ly.off_w; s[0].ldd = ly.ldn;
    {
      // split the frame (reduction) dimension so a wgrad alone can fill the SMs; partials meet in the
      // TMA reduce-add.  Keep >= 16 k-blocks (1024 frames) per work item.
      const int tile_m = h->two_cta ? 256 : BM;
      const int tiles = ((ly.K + tile_m - 1) / tile_m) * ((ly.N + BN - 1) / BN);
      const int kb = (B + BK - 1) / BK;
      const int units = h->two_cta ? h->num_sms / 2 : h->num_sms;  // concurrent tiles
      // only when the launch cannot fill the machine otherwise (layer 0: its wgrad runs alone and has
      // few output tiles); hidden layers share their launch with 4x more dgrad tiles and measured slower
      // with split-K (extra reduce-add traffic)
      int ks = (ai > 0 || tiles >= units) ? 1 : (2 * units + tiles - 1) / tiles;
      if (ks > kb / 16) ks = kb / 16;
      s[0].ksplit = ks < 1 ? 1 : ks;
    }
    if (h->fused_rs && ly.peer_reduce) {  // add each output slab into its owner GPU's accumulator over NVLink
      s[0].peer_tm = ly.peer_tm;
      s[0].num_peers = h->nranks;
      s[0].rows_per_owner = ly.Kpad / h->nranks;
    }
    int nspec = 1;
    if (ai > 0) {
      // dX[B, K] = dZ[B, N] . W[K, N]^T, masked by the forward output of the layer below
      const int lower = ai - 1;  // hidden layer that produced act[ai]
      const int dst = (L - 1 - lower) & 1;  // where layer `lower` expects its dZ
      base_operands(h, s[1]);
      s[1].M = B; s[1].N = ly.K; s[1].K = ly.N;
      s[1].A_hi = dz_hi; s[1].A_lo = dz_lo; s[1].lda = ldz; s[1].a_mn = 0;
      s[1].B_hi = h->Sh + ly.off_w; s[1].B_lo = opt(h->Sl, ly.off_w); s[1].ldb = ly.ldn; s[1].b_mn = 0;
      s[1].out_kind = h->x3 ? OUT_BF16_SPLIT : OUT_BF16;
      s[1].D_hi = h->dA_hi[dst]; s[1].D_lo = h->dA_lo[dst]; s[1].ldd = h->layers[lower].ldn;
      // Where the dropout mask has to be told from a stored activation, "stored value == 0" proves a drop only if the
      // chain cannot produce an exact 0 on a kept unit.  relu can (and its slope there is 0 anyway: nothing to tell);
      // sigmoid cannot; identity and tanh can, on pre-activations that are exactly 0 — rare, but 8192 x 2048 draws per
      // layer find them — so those chains re-draw the keep decisions from the layer's Philox key instead.
      const bool redraw = drop && (h->cfg.nonlin == TFK_NONLIN_LINEAR || h->cfg.nonlin == TFK_NONLIN_TANH);
      if (l2) {  // the layer below ends in L2Norm: only its dropout is undone here, the rest in k_l2norm_bwd
        if (drop) {
          s[1].mask_src = h->act_hi[ai]; s[1].mask_ld = h->layers[lower].ldn;
          s[1].mask_nonzero = 1; s[1].scale = 1.0f / h->cfg.keep_prob;
          s[1].bwd_drop_keep = h->cfg.keep_prob;  // (relu + L2Norm: a kept 0 has slope 0 downstream, but the re-draw is exact for every chain)
        }
      } else if (smooth) {  // sigmoid / tanh: slope from the stored forward output
        s[1].mask_src = h->act_hi[ai]; s[1].mask_src_lo = h->act_lo[ai]; s[1].mask_ld = h->layers[lower].ldn;
        s[1].deriv = h->cfg.nonlin == TFK_NONLIN_SIGMOID ? 1 : 2;
        s[1].dropout_in_chain = drop ? 1 : 0;
        s[1].scale = drop ? 1.0f / h->cfg.keep_prob : 1.0f;
        if (redraw) s[1].bwd_drop_keep = h->cfg.keep_prob;
      } else if (relu || drop) {
        if (h->layers[lower].bn) {  // activation written by bn_apply: test the stored forward output
          s[1].mask_src = h->act_hi[ai]; s[1].mask_ld = h->layers[lower].ldn;
          if (redraw) s[1].bwd_drop_keep = h->cfg.keep_prob;
        } else {                    // 1 bit per unit, written by the forward epilogue
          s[1].mask_bits_in = h->layers[lower].maskbits; s[1].mask_bits_ld = h->cfg.max_frames;
        }
        s[1].mask_nonzero = relu ? 0 : 1;
        s[1].scale = drop ? 1.0f / h->cfg.keep_prob : 1.0f;
      }
      if (!h->layers[lower].bn && !l2) {  // dz of the layer below is final here: take its column sums for free
        s[1].colsum_part = h->layers[lower].db_part;
        s[1].colsum_ld = h->layers[lower].ldn;
      } else if (h->layers[lower].bn && !l2) {
        // dY of the batch-norm layer below is final here: emit the partials of sum dY and sum dY*xhat (bn_bwd_finalize
        // turns them into the two column means of the batch-norm gradient and the beta gradient).  The forward's
        // statistics partials are idle during the backward pass and have the same [groups, npad] shape.
        const Layer& lw = h->layers[lower];
        s[1].colsum_part = h->bn_ps;
        s[1].colsum2_part = h->bn_pq;
        s[1].colsum_ld = lw.npad;
        s[1].bn_z_hi = lw.z_hi; s[1].bn_z_lo = lw.z_lo; s[1].bn_z_ld = lw.ldn;
        s[1].bn_mean = lw.bn_mean; s[1].bn_rstd = lw.bn_rstd;
        // relu / linear chains recover xhat from the stored output the mask is read from anyway (no z traffic);
        // in bf16x3 mode that output is hi + lo
        static const bool from_y = [] {  // TFK_BN_FROM_Y=0: always read z (A/B measurements)
          const char* e = getenv("TFK_BN_FROM_Y");
          return !(e && e[0] == '0');
        }();
        if (from_y && !smooth && (relu || drop)) { s[1].bn_beta = h->P + lw.off_beta; s[1].mask_src_lo = h->act_lo[ai]; }
      }
      nspec = 2;
    }
    TFK_TRY(finish_params(h, plan, s, nspec, &plan.bwd[l], "backward", l));
  }
  return TFK_OK;
}

int get_plan(tfk_handle* h, int B, Plan** out) {
  const long long key = static_cast<long long>(B) * 64 + h->active;
  auto it = h->plans.find(key);
  if (it == h->plans.end()) {
    if (h->plans.size() > 512) {  // tensor maps only: ~40 KB of host memory each
      cudaDeviceSynchronize();
      for (auto& kv : h->plans) free_plan(kv.second);
      h->plans.clear();
    }
    Plan p;
    if (int rc = build_plan(h, B, p)) {
      free_plan(p);
      return rc;
    }
    it = h->plans.emplace(key, std::move(p)).first;
  }
  *out = &it->second;
  return TFK_OK;
}

// data parallel, no host sync requested: keep the host at most `runahead` optimizer steps ahead of the device
int bound_runahead(tfk_handle* h, cudaStream_t st) {
  if (!(h->comm && h->nranks > 1 && h->runahead > 0)) return TFK_OK;
  if (h->step_events.empty()) {
    h->step_events.resize(h->runahead);
    for (auto& e : h->step_events) TFK_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  cudaEvent_t e = h->step_events[h->steps_queued % h->runahead];
  if (h->steps_queued >= h->runahead) TFK_CUDA(h, cudaEventSynchronize(e));  // step (k - runahead) has finished
  TFK_CUDA(h, cudaEventRecord(e, st));
  h->steps_queued += 1;
  return TFK_OK;
}

int check_frames(tfk_handle* h, int B, const char* what) {
  if (B <= 0) return fail(h, TFK_ESHAPE, "%s: B=%d must be positive", what, B);
  if (B > h->cfg.max_frames)
    return fail(h, TFK_ESHAPE, "%s: B=%d exceeds max_frames=%d", what, B, h->cfg.max_frames);
  return TFK_OK;
}

// Philox key of layer l's dropout mask in the micro-batch about to run.  `drop_seed` counts GLOBAL micro-batches:
// K data-parallel ranks stand for K accumulated micro-batches of one step (trainer.py:310-332), micro-batch j of
// which uses drop_seed + j (L+1); rank r is micro-batch r, so every rank draws a different mask.
inline unsigned long long layer_seed(const tfk_handle* h, int l) {
  return h->drop_seed + static_cast<unsigned long long>(h->rank) * (h->L + 1) + static_cast<unsigned long long>(l);
}
inline void advance_drop_seed(tfk_handle* h) {
  h->drop_seed += static_cast<unsigned long long>(h->nranks) * (h->L + 1);
}

// forward of hidden layers [first, active) and optionally the output layer; input must be in act[first]
int forward_range(tfk_handle* h, Plan& plan, int B, bool training, int first, bool with_output,
                  cudaStream_t st) {
  const int L = h->L;
  const int act_code = h->cfg.nonlin == TFK_NONLIN_RELU ? 1 : h->cfg.nonlin == TFK_NONLIN_SIGMOID ? 2
                       : h->cfg.nonlin == TFK_NONLIN_TANH ? 3 : 0;
  for (int l = first; l <= L; ++l) {
    if (l < L && l >= h->active) continue;
    if (l == L && !with_output) break;
    Layer& ly = h->layers[l];
    GemmParams gp = training ? plan.fwd_train[l] : plan.fwd_eval[l];
    if (training && gp.p[0].drop_thr != 0u) gp.p[0].seed = layer_seed(h, l);
    {
      TimerScope ts(h, st, TFK_TIMER_GEMM_FWD);
      TFK_LAUNCH(h, gemm_launch(gp, h->num_sms, st));
    }
    if (ly.hidden && ly.bn) {
      TimerScope ts(h, st, TFK_TIMER_BN, 1);
      const bool l2 = h->cfg.l2_norm != 0;
      const float keep = (!l2 && training && h->cfg.keep_prob < 1.0f) ? h->cfg.keep_prob : 1.0f;
      // statistics + normalise + beta + non-linearity + dropout in one column-strip launch
      TFK_LAUNCH(h, k_bn_fwd_strip(h->bn_ps, h->bn_pq, (B + 31) / 32, ly.npad, ly.z_hi, ly.z_lo, ly.ldn, B, ly.N, h->cfg.bn_eps,
                                   h->cfg.bn_decay, training ? 1 : 0, ly.bn_mean, ly.bn_rstd, ly.moving_mean, ly.moving_var,
                                   h->P + ly.off_beta, act_code, keep, layer_seed(h, l), l2 ? ly.u_hi : h->act_hi[l + 1],
                                   l2 ? ly.u_lo : h->act_lo[l + 1], st));
    }
    if (ly.hidden && h->cfg.l2_norm) {
      TimerScope ts(h, st, TFK_TIMER_BN);
      const float keep = (training && h->cfg.keep_prob < 1.0f) ? h->cfg.keep_prob : 1.0f;
      TFK_LAUNCH(h, k_l2norm_fwd(ly.u_hi, ly.u_lo, ly.ldn, B, ly.N, keep, layer_seed(h, l),
                                 h->act_hi[l + 1], h->act_lo[l + 1], ly.l2_s, st));
    }
  }
  return TFK_OK;
}

// backward of layer l given dZ_l (hidden: in dA[(L-1-l)&1], as d(loss)/d(layer OUTPUT after mask) for BN)
// `side`: tfk_train_step's second stream.  Bias-gradient column sums that need their own kernel are only consumed by
// the final bias update, so they run there, next to the backward GEMM, fenced by per-layer events.
int backward_layer(tfk_handle* h, Plan& plan, int B, int l, cudaStream_t st,
                   const GemmParams* gemm_override = nullptr, bool fused_colsum = false, cudaStream_t side = nullptr) {
  const int L = h->L;
  Layer& ly = h->layers[l];
  __nv_bfloat16* dz_hi = ly.hidden ? h->dA_hi[(L - 1 - l) & 1] : h->dzo_hi;
  __nv_bfloat16* dz_lo = ly.hidden ? h->dA_lo[(L - 1 - l) & 1] : h->dzo_lo;
  const int ldz = ly.hidden ? ly.ldn : h->ldo;
  if (ly.hidden && h->cfg.l2_norm) {  // d(L2Norm output) -> d(pre-nonlinearity)
    TimerScope ts(h, st, TFK_TIMER_BN);
    const int act_code = h->cfg.nonlin == TFK_NONLIN_RELU ? 1 : h->cfg.nonlin == TFK_NONLIN_SIGMOID ? 2
                         : h->cfg.nonlin == TFK_NONLIN_TANH ? 3 : 0;
    TFK_LAUNCH(h, k_l2norm_bwd(dz_hi, dz_lo, ly.u_hi, ly.u_lo, ly.l2_s, ly.ldn, B, ly.N, act_code, st));
  }
  if (ly.hidden && ly.bn) {  // d(bn output) -> d(linear output), plus dbeta
    TimerScope ts(h, st, TFK_TIMER_BN, 2);
    if (fused_colsum && !h->cfg.l2_norm) {  // the dgrad epilogue of the layer above left the partial sums: one strip launch
      TFK_LAUNCH(h, k_bn_bwd_strip(h->bn_ps, h->bn_pq, (B + 31) / 32, ly.npad, dz_hi, dz_lo, ly.z_hi, ly.z_lo, ly.ldn, B, ly.N,
                                   ly.bn_mean, ly.bn_rstd, h->G + ly.off_beta, st));
    } else {
      TFK_LAUNCH(h, k_bn_bwd_reduce(dz_hi, dz_lo, ly.z_hi, ly.z_lo, ly.ldn, B, ly.N, ly.bn_mean, ly.bn_rstd,
                                    h->ws, h->bn_counters, ly.bn_sums, h->G + ly.off_beta, st));
      TFK_LAUNCH(h, k_bn_bwd_apply(dz_hi, dz_lo, ly.z_hi, ly.z_lo, ly.ldn, B, ly.N, ly.bn_mean, ly.bn_rstd,
                                   ly.bn_sums, st));
    }
  }
  // Bias gradient = column sums of dZ.  Under batch-norm it is identically zero: z = xW + b enters only through
  // z - mean_B(z), so sum_B dz = rstd * (sum dy - sum dy - mean(dy xhat) * sum xhat) = 0 (layer.py:52 feeding
  // activation.py:159).  The reference's fp32 graph evaluates that zero as round-off noise and lets Adam normalise
  // it into a random walk of a parameter the training-mode output does not depend on; no two implementations share
  // that noise, so the exact value is used here and in the oracle: the bias of a batch-normalised layer never moves.
  if ((!fused_colsum || !ly.hidden || h->cfg.l2_norm) && !(ly.hidden && ly.bn)) {
    cudaStream_t cs = side ? side : st;
    if (side) {  // dZ_l is final on `st` here
      TFK_CUDA(h, cudaEventRecord(h->colsum_events[l], st));
      TFK_CUDA(h, cudaStreamWaitEvent(side, h->colsum_events[l], 0));
    }
    {
      TimerScope ts(h, cs, TFK_TIMER_COLSUM);
      TFK_LAUNCH(h, k_colsum_bf16(dz_hi, dz_lo, ldz, B, ly.N, h->ws_colsum, h->G + ly.off_b, cs));
    }
    if (side) TFK_CUDA(h, cudaEventRecord(h->colsum_events[l], side));
  }
  // this layer's GEMM writes dZ_{l-1} into the buffer that held dZ_{l+1}: its column sums must have been taken
  if (side && l + 1 < L) TFK_CUDA(h, cudaStreamWaitEvent(st, h->colsum_events[l + 1], 0));
  {
    TimerScope ts(h, st, TFK_TIMER_GEMM_BWD);
    GemmParams gp = gemm_override ? *gemm_override : plan.bwd[l];
    // a dgrad epilogue that re-draws the dropout mask of the layer below needs that layer's Philox key of THIS micro-batch
    if (gp.nprob > 1 && gp.p[1].bwd_drop_thr != 0u) gp.p[1].seed = layer_seed(h, (l == L ? h->active : l) - 1);
    TFK_LAUNCH(h, gemm_launch(gp, h->num_sms, st));
  }
  return TFK_OK;
}

int load_input(tfk_handle* h, const float* x, int B, int layer_in, cudaStream_t st) {
  const Layer& ly = h->layers[layer_in == h->L ? h->L : layer_in];
  const int ai = layer_in == h->L ? h->active : layer_in;
  TimerScope ts(h, st, TFK_TIMER_CONVERT);
  TFK_LAUNCH(h, k_split_f32(x, ly.K, h->act_hi[ai], h->act_lo[ai], ly.ldk, B, ly.K, st));
  return TFK_OK;
}

int tensor_view(tfk_handle* h, int kind, int layer, float** base, int* rows, int* cols, int* ld) {
  if (layer < 0 || layer > h->L) return fail(h, TFK_EINVAL, "layer %d out of range [0,%d]", layer, h->L);
  Layer& ly = h->layers[layer];
  float* arena = nullptr;
  int which = 0;  // 0 = W, 1 = b, 2 = beta
  switch (kind) {
    case TFK_T_WEIGHTS: arena = h->P; which = 0; break;
    case TFK_T_BIASES: arena = h->P; which = 1; break;
    case TFK_T_BN_BETA: arena = h->P; which = 2; break;
    case TFK_T_ADAM_M_W: arena = h->M; which = 0; break;
    case TFK_T_ADAM_V_W: arena = h->V; which = 0; break;
    case TFK_T_ADAM_M_B: arena = h->M; which = 1; break;
    case TFK_T_ADAM_V_B: arena = h->V; which = 1; break;
    case TFK_T_ADAM_M_BETA: arena = h->M; which = 2; break;
    case TFK_T_ADAM_V_BETA: arena = h->V; which = 2; break;
    case TFK_T_GRAD_W: arena = h->G; which = 0; break;
    case TFK_T_GRAD_B: arena = h->G; which = 1; break;
    case TFK_T_GRAD_BETA: arena = h->G; which = 2; break;
    case TFK_T_BN_MOVING_MEAN:
    case TFK_T_BN_MOVING_VAR:
      if (!ly.bn) return fail(h, TFK_EINVAL, "layer %d has no batch-norm state", layer);
      *base = kind == TFK_T_BN_MOVING_MEAN ? ly.moving_mean : ly.moving_var;
      *rows = 1; *cols = ly.N; *ld = ly.N;
      return TFK_OK;
    default: return fail(h, TFK_EINVAL, "unknown tensor kind %d", kind);
  }
  if (which == 0) { *base = arena + ly.off_w; *rows = ly.K; *cols = ly.N; *ld = ly.ldn; }
  else if (which == 1) { *base = arena + ly.off_b; *rows = 1; *cols = ly.N; *ld = ly.N; }
  else {
    if (!ly.bn) return fail(h, TFK_EINVAL, "layer %d has no batch-norm beta", layer);
    *base = arena + ly.off_beta; *rows = 1; *cols = ly.N; *ld = ly.N;
  }
  return TFK_OK;
}

int refresh_shadow(tfk_handle* h, const Layer& ly, cudaStream_t st) {
  TimerScope ts(h, st, TFK_TIMER_CONVERT);
  TFK_LAUNCH(h, k_split_f32(h->P + ly.off_w, ly.ldn, h->Sh + ly.off_w, h->x3 ? h->Sl + ly.off_w : nullptr, ly.ldn,
                            ly.K, ly.ldn, st));
  return TFK_OK;
}

int allreduce_grads(tfk_handle* h, cudaStream_t st) {
  if (!h->comm || h->nranks <= 1) return TFK_OK;
  NcclApi& api = nccl();
  TFK_CUDA(h, cudaEventRecord(h->ev_compute, st));
  TFK_CUDA(h, cudaStreamWaitEvent(h->comm_stream, h->ev_compute, 0));
  {
    TimerScope ts(h, h->comm_stream, TFK_TIMER_ALLREDUCE, 2);
    int rc = api.GroupStart();
    // one bucket per layer, output layer first (the order the backward pass finishes them)
    for (int l = h->L; l >= 0 && rc == 0; --l) {
      const Layer& ly = h->layers[l];
      rc = api.AllReduce(h->G + ly.off_w, h->G + ly.off_w, ly.w_count, kNcclFloat, kNcclSum, h->comm, h->comm_stream);
    }
    if (rc == 0) rc = api.AllReduce(h->G + h->nW, h->G + h->nW, h->arena_n - h->nW, kNcclFloat, kNcclSum, h->comm, h->comm_stream);
    if (rc == 0) rc = api.AllReduce(h->acc, h->acc, 2, kNcclDouble, kNcclSum, h->comm, h->comm_stream);
    const int rc2 = api.GroupEnd();
    if (rc || rc2)
      return fail(h, TFK_ENCCL, "ncclAllReduce failed: %s", api.GetErrorString ? api.GetErrorString(rc ? rc : rc2) : "?");
  }
  TFK_CUDA(h, cudaEventRecord(h->ev_comm, h->comm_stream));
  TFK_CUDA(h, cudaStreamWaitEvent(st, h->ev_comm, 0));
  return TFK_OK;
}

// Sharded data parallelism (ZeRO-1 style): every layer's weight gradient is reduce-scattered (rank r ends
// up with the global sum of slice r), the small vectors and {loss, frames} are all-reduced.
int reduce_scatter_grads(tfk_handle* h, cudaStream_t st) {
  NcclApi& api = nccl();
  TFK_CUDA(h, cudaEventRecord(h->ev_compute, st));
  TFK_CUDA(h, cudaStreamWaitEvent(h->comm_stream, h->ev_compute, 0));
  {
    TimerScope ts(h, h->comm_stream, TFK_TIMER_ALLREDUCE, h->L + 3);
    int rc = api.GroupStart();
    for (int l = h->L; l >= 0 && rc == 0; --l) {
      const Layer& ly = h->layers[l];
      if (h->fused_rs && ly.peer_reduce) continue;  // already summed into the owners by the wgrad epilogues
      const size_t cnt = ly.w_count / h->nranks;
      rc = api.ReduceScatter(h->G + ly.off_w, h->G + ly.off_w + cnt * h->rank, cnt, kNcclFloat, kNcclSum, h->comm,
                             h->comm_stream);
    }
    if (rc == 0) rc = api.AllReduce(h->G + h->nW, h->G + h->nW, h->arena_n - h->nW, kNcclFloat, kNcclSum, h->comm, h->comm_stream);
    if (rc == 0) rc = api.AllReduce(h->acc, h->acc, 2, kNcclDouble, kNcclSum, h->comm, h->comm_stream);
    const int rc2 = api.GroupEnd();
    if (rc || rc2)
      return fail(h, TFK_ENCCL, "ncclReduceScatter failed: %s", api.GetErrorString ? api.GetErrorString(rc ? rc : rc2) : "?");
  }
  TFK_CUDA(h, cudaEventRecord(h->ev_comm, h->comm_stream));
  TFK_CUDA(h, cudaStreamWaitEvent(st, h->ev_comm, 0));
  return TFK_OK;
}

// after the sharded Adam step: every rank publishes the bf16 operand copies of its weight slices
int all_gather_shadows(tfk_handle* h, cudaStream_t st) {
  NcclApi& api = nccl();
  TFK_CUDA(h, cudaEventRecord(h->ev_compute, st));
  TFK_CUDA(h, cudaStreamWaitEvent(h->comm_stream, h->ev_compute, 0));
  {
    TimerScope ts(h, h->comm_stream, TFK_TIMER_ALLREDUCE, (h->L + 1) * (h->x3 ? 2 : 1));
    int rc = api.GroupStart();
    for (int l = 0; l <= h->L && rc == 0; ++l) {  // layer 0 first: the next forward needs it first
      const Layer& ly = h->layers[l];
      const size_t cnt = ly.w_count / h->nranks;
      rc = api.AllGather(h->Sh + ly.off_w + cnt * h->rank, h->Sh + ly.off_w, cnt, kNcclBf16, h->comm, h->comm_stream);
      if (rc == 0 && h->x3)
        rc = api.AllGather(h->Sl + ly.off_w + cnt * h->rank, h->Sl + ly.off_w, cnt, kNcclBf16, h->comm, h->comm_stream);
    }
    const int rc2 = api.GroupEnd();
    if (rc || rc2)
      return fail(h, TFK_ENCCL, "ncclAllGather failed: %s", api.GetErrorString ? api.GetErrorString(rc ? rc : rc2) : "?");
  }
  TFK_CUDA(h, cudaEventRecord(h->ev_comm, h->comm_stream));
  TFK_CUDA(h, cudaStreamWaitEvent(st, h->ev_comm, 0));
  return TFK_OK;
}

// collective: make the fp32 master weights and Adam slots of the weight regions identical everywhere
int sync_sharded_params(tfk_handle* h, cudaStream_t st) {
  if (!h->sharded || h->params_synced) return TFK_OK;
  NcclApi& api = nccl();
  TFK_CUDA(h, cudaEventRecord(h->ev_compute, st));
  TFK_CUDA(h, cudaStreamWaitEvent(h->comm_stream, h->ev_compute, 0));
  int rc = api.GroupStart();
  float* arenas[3] = {h->P, h->M, h->V};
  for (int a = 0; a < 3 && rc == 0; ++a)
    for (int l = 0; l <= h->L && rc == 0; ++l) {
      const Layer& ly = h->layers[l];
      const size_t cnt = ly.w_count / h->nranks;
      rc = api.AllGather(arenas[a] + ly.off_w + cnt * h->rank, arenas[a] + ly.off_w, cnt, kNcclFloat, h->comm, h->comm_stream);
    }
  const int rc2 = api.GroupEnd();
  if (rc || rc2) return fail(h, TFK_ENCCL, "parameter all-gather failed: %s", api.GetErrorString ? api.GetErrorString(rc ? rc : rc2) : "?");
  TFK_CUDA(h, cudaEventRecord(h->ev_comm, h->comm_stream));
  TFK_CUDA(h, cudaStreamWaitEvent(st, h->ev_comm, 0));
  h->params_synced = true;
  return TFK_OK;
}

int ce_and_backward(tfk_handle* h, Plan& plan, const int32_t* labels, int B, bool backward, cudaStream_t st) {
  {
    TimerScope ts(h, st, TFK_TIMER_SOFTMAX_CE, 2);
    TFK_LAUNCH(h, k_softmax_ce(h->logits, h->ldo, labels, B, h->cfg.output_dim, h->row_loss,
                               backward ? h->dzo_hi : nullptr, (backward && h->x3) ? h->dzo_lo : nullptr, st));
    TFK_LAUNCH(h, k_accum_loss(h->row_loss, B, h->acc, st));
  }
  if (!backward) return TFK_OK;
  TFK_TRY(backward_layer(h, plan, B, h->L, st, nullptr, true));
  for (int l = h->active - 1; l >= 0; --l) TFK_TRY(backward_layer(h, plan, B, l, st, nullptr, true));
  {
    // bias gradients of the non-BN hidden layers: finish the column sums their dgrad epilogues started
    const float* parts[64];
    float* outs[64];
    int n = 0;
    for (int l = 0; l < h->active; ++l)
      if (!h->layers[l].bn && !h->cfg.l2_norm) {
        parts[n] = h->layers[l].db_part;
        outs[n] = h->G + h->layers[l].off_b;
        ++n;
      }
    if (n > 0) {
      TimerScope ts(h, st, TFK_TIMER_COLSUM);
      TFK_LAUNCH(h, k_colsum_finalize(parts, outs, n, (B + 31) / 32, h->ldh, h->cfg.hidden_dim, st));
    }
  }
  return TFK_OK;
}

}  // namespace

// ================================================================== C ABI
extern "C" {

int tfk_abi_version(void) { return TFK_ABI_VERSION; }

int tfk_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

void tfk_default_config(tfk_config* cfg) {
  memset(cfg, 0, sizeof(*cfg));
  cfg->abi_version = TFK_ABI_VERSION;
  cfg->num_layers = 6;
  cfg->input_dim = 440;
  cfg->hidden_dim = 2048;
  cfg->output_dim = 1936;
  cfg->max_frames = 8192;
  cfg->nonlin = TFK_NONLIN_RELU;
  cfg->batch_norm = 0;
  cfg->keep_prob = 1.0f;
  cfg->bn_eps = 1e-3f;
  cfg->bn_decay = 0.999f;
  cfg->adam_beta1 = 0.9f;
  cfg->adam_beta2 = 0.999f;
  cfg->adam_eps = 1e-8f;
  cfg->precision = TFK_PREC_BF16;
  cfg->device = 0;
  cfg->seed = 0;
}

const char* tfk_last_error(const tfk_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int tfk_destroy(tfk_handle* h) {
  if (!h) return TFK_OK;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  if (h->comm && h->own_comm && nccl().ok) nccl().CommDestroy(h->comm);
  for (auto& r : h->timer_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto e : h->event_pool) cudaEventDestroy(e);
  if (h->ev_compute) cudaEventDestroy(h->ev_compute);
  if (h->ev_comm) cudaEventDestroy(h->ev_comm);
  if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
  if (h->adam_stream) cudaStreamDestroy(h->adam_stream);
  for (auto e : h->layer_events) cudaEventDestroy(e);
  for (auto e : h->colsum_events) cudaEventDestroy(e);
  for (auto e : h->step_events) cudaEventDestroy(e);
  for (auto& kv : h->plans) free_plan(kv.second);
  for (auto& kv : h->list_cache) cudaFree(kv.second.d);
  for (void* p : h->ipc_opened) cudaIpcCloseMemHandle(p);
  for (void* p : h->allocs) cudaFree(p);
  if (h->acc_host) cudaFreeHost(h->acc_host);
  delete h;
  return TFK_OK;
}

int tfk_create(const tfk_config* cfg, tfk_handle** out) {
  if (!cfg || !out) return fail(nullptr, TFK_EINVAL, "tfk_create: null argument");
  *out = nullptr;
  if (cfg->abi_version != TFK_ABI_VERSION)
    return fail(nullptr, TFK_EINVAL, "tfk_create: abi_version %d != %d", cfg->abi_version, TFK_ABI_VERSION);
  if (cfg->num_layers < 1 || cfg->num_layers > 62 || cfg->input_dim < 1 || cfg->hidden_dim < 1 ||
      cfg->output_dim < 1 || cfg->max_frames < 1)
    return fail(nullptr, TFK_EINVAL, "tfk_create: bad dimensions (L=%d I=%d H=%d O=%d maxB=%d)", cfg->num_layers,
                cfg->input_dim, cfg->hidden_dim, cfg->output_dim, cfg->max_frames);
  if (!(cfg->keep_prob > 0.0f))
    return fail(nullptr, TFK_EINVAL, "tfk_create: keep_prob must be in (0,1] (classifiers/activation.py:127)");
  if (cfg->nonlin < TFK_NONLIN_RELU || cfg->nonlin > TFK_NONLIN_TANH)
    return fail(nullptr, TFK_EINVAL, "tfk_create: unknown nonlinearity %d (nnet.py:65)", cfg->nonlin);
  if (cfg->precision != TFK_PREC_BF16 && cfg->precision != TFK_PREC_BF16X3)
    return fail(nullptr, TFK_EINVAL, "tfk_create: unknown precision %d", cfg->precision);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(nullptr, TFK_ECUDA, "tfk_create: no CUDA device available (this engine has no CPU path)");
  }
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, TFK_EINVAL, "tfk_create: device %d of %d", cfg->device, ndev);

  tfk_handle* h = new tfk_handle();
  h->cfg = *cfg;
  if (h->cfg.keep_prob > 1.0f) h->cfg.keep_prob = 1.0f;
  h->L = cfg->num_layers;
  h->active = h->L;
  h->x3 = cfg->precision == TFK_PREC_BF16X3;
  h->drop_seed = cfg->seed;
  if (const char* e = getenv("TFK_GEMM_2CTA")) h->two_cta = atoi(e) != 0;
  if (const char* e = getenv("TFK_DP_RUNAHEAD")) h->runahead = atoi(e) < 0 ? 0 : (atoi(e) > 64 ? 64 : atoi(e));
  auto bail = [&](int rc) {
    g_create_error = h->err;
    tfk_destroy(h);
    return rc;
  };
#define CREATE_TRY(expr)               \
  do {                                 \
    int rc__ = (expr);                 \
    if (rc__ != TFK_OK) return bail(rc__); \
  } while (0)
  {
    cudaError_t e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) return bail(fail(h, TFK_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e)));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, cfg->device);
    if (e != cudaSuccess) return bail(fail(h, TFK_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)));
    if (prop.major != 10)
      return bail(fail(h, TFK_ECUDA, "device %s is sm_%d%d; this library contains sm_100a code only", prop.name,
                       prop.major, prop.minor));
    h->num_sms = prop.multiProcessorCount;
    const int rc = gemm_init();
    if (rc) return bail(fail(h, TFK_ECUDA, "gemm_init: %s", cudaGetErrorString(static_cast<cudaError_t>(rc))));
  }
  const int L = h->L, maxB = cfg->max_frames;
  h->layers.resize(L + 1);
  size_t off = 0;
  for (int l = 0; l <= L; ++l) {
    Layer& ly = h->layers[l];
    ly.hidden = l < L;
    ly.K = (l == 0) ? cfg->input_dim : cfg->hidden_dim;
    ly.N = ly.hidden ? cfg->hidden_dim : cfg->output_dim;
    ly.ldk = round_up(ly.K, 8);
    ly.ldn = round_up(ly.N, 8);
    ly.npad = round_up(ly.N, 256);
    ly.bn = ly.hidden && cfg->batch_norm;
    // weight regions first (each a multiple of 1024 floats so it splits evenly over 2^k ranks), then the
    // small per-column vectors: under sharded data parallelism the first part is reduce-scattered and
    // its Adam step sharded, the second part is all-reduced and replicated
    // Rows are padded to a multiple of 512 (layer 0: 440 -> 512; the padding rows are zero and stay zero) so that every
    // layer's rows split over 2..16 ranks in multiples of the 32-row store slab: the fused GEMM -> reduce-scatter
    // epilogue and the sharded update then cover ALL layers and no NCCL reduce-scatter is left on the step's tail.
    ly.Kpad = round_up(ly.K, 512);
    ly.w_count = static_cast<size_t>(ly.Kpad) * ly.ldn;
    ly.off_w = off; off += ly.w_count;
  }
  h->nW = off;
  for (int l = 0; l <= L; ++l) {
    Layer& ly = h->layers[l];
    ly.off_b = off; off += ly.npad;
    if (ly.bn) { ly.off_beta = off; off += ly.npad; }
  }
  h->arena_n = (off + 7) / 8 * 8;  // the Adam kernel works on 8-float granules
  h->ld0 = h->layers[0].ldk;
  h->ldh = round_up(cfg->hidden_dim, 8);
  h->ldo = round_up(cfg->output_dim, 8);
  h->ldmax = h->ldh > h->ldo ? h->ldh : h->ldo;
  if (h->ld0 > h->ldmax) h->ldmax = h->ld0;
  CREATE_TRY(dev_alloc(h, &h->P, h->arena_n));
  CREATE_TRY(dev_alloc(h, &h->G, h->arena_n));
  CREATE_TRY(dev_alloc(h, &h->M, h->arena_n));
  CREATE_TRY(dev_alloc(h, &h->V, h->arena_n));
  CREATE_TRY(dev_alloc(h, &h->Sh, h->arena_n));
  if (h->x3) CREATE_TRY(dev_alloc(h, &h->Sl, h->arena_n));
  h->act_hi.assign(L + 1, nullptr);
  h->act_lo.assign(L + 1, nullptr);
  for (int l = 0; l <= L; ++l) {
    const size_t n = static_cast<size_t>(maxB) * (l == 0 ? h->ld0 : h->ldh);
    CREATE_TRY(dev_alloc(h, &h->act_hi[l], n));
    if (h->x3) CREATE_TRY(dev_alloc(h, &h->act_lo[l], n));
  }
  if (cfg->l2_norm) {
    for (int l = 0; l < L; ++l) {
      Layer& ly = h->layers[l];
      CREATE_TRY(dev_alloc(h, &ly.u_hi, static_cast<size_t>(maxB) * ly.ldn));
      if (h->x3) CREATE_TRY(dev_alloc(h, &ly.u_lo, static_cast<size_t>(maxB) * ly.ldn));
      CREATE_TRY(dev_alloc(h, &ly.l2_s, maxB));
    }
  }
  for (int l = 0; l < L; ++l) {
    if (!h->layers[l].bn) {
      CREATE_TRY(dev_alloc(h, &h->layers[l].db_part, static_cast<size_t>(maxB / 32 + 8) * h->layers[l].ldn));
      CREATE_TRY(dev_alloc(h, &h->layers[l].maskbits, static_cast<size_t>((h->layers[l].ldn + 31) / 32 + 8) * maxB));
    }
  }
  for (int l = 0; l < L; ++l) {
    Layer& ly = h->layers[l];
    if (!ly.bn) continue;
    CREATE_TRY(dev_alloc(h, &ly.z_hi, static_cast<size_t>(maxB) * ly.ldn));
    if (h->x3) CREATE_TRY(dev_alloc(h, &ly.z_lo, static_cast<size_t>(maxB) * ly.ldn));
    CREATE_TRY(dev_alloc(h, &ly.moving_mean, ly.npad));
    CREATE_TRY(dev_alloc(h, &ly.moving_var, ly.npad));
    CREATE_TRY(dev_alloc(h, &ly.bn_mean, ly.npad));
    CREATE_TRY(dev_alloc(h, &ly.bn_rstd, ly.npad));
    CREATE_TRY(dev_alloc(h, &ly.bn_sums, 2 * static_cast<size_t>(ly.ldn)));
    std::vector<float> ones(ly.npad, 1.0f);  // moving_variance initialises to 1 (tf.contrib.layers.batch_norm)
    cudaError_t e = cudaMemcpy(ly.moving_var, ones.data(), ly.npad * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return bail(fail(h, TFK_ECUDA, "init moving_var: %s", cudaGetErrorString(e)));
  }
  CREATE_TRY(dev_alloc(h, &h->logits, static_cast<size_t>(maxB) * h->ldo));
  CREATE_TRY(dev_alloc(h, &h->dzo_hi, static_cast<size_t>(maxB) * h->ldo));
  if (h->x3) CREATE_TRY(dev_alloc(h, &h->dzo_lo, static_cast<size_t>(maxB) * h->ldo));
  for (int i = 0; i < 2; ++i) {
    CREATE_TRY(dev_alloc(h, &h->dA_hi[i], static_cast<size_t>(maxB) * h->ldh));
    if (h->x3) CREATE_TRY(dev_alloc(h, &h->dA_lo[i], static_cast<size_t>(maxB) * h->ldh));
  }
  CREATE_TRY(dev_alloc(h, &h->row_loss, maxB));
  CREATE_TRY(dev_alloc(h, &h->acc, 2));
  if (cfg->batch_norm) {
    const size_t n = static_cast<size_t>(maxB / 32 + 8) * round_up(cfg->hidden_dim, 256);
    CREATE_TRY(dev_alloc(h, &h->bn_ps, n));
    CREATE_TRY(dev_alloc(h, &h->bn_pq, n));
  }
  CREATE_TRY(dev_alloc(h, &h->ws, 256 * static_cast<size_t>(h->ldmax)));       // bn backward partials [128][2][ld]
  CREATE_TRY(dev_alloc(h, &h->ws_colsum, 1024 + 64 * static_cast<size_t>(h->ldmax)));  // colsum counters + partials
  CREATE_TRY(dev_alloc(h, &h->bn_counters, 512));
  CREATE_TRY(dev_alloc(h, &h->tmp_f32, static_cast<size_t>(maxB) * h->ldmax));
  CREATE_TRY(dev_alloc(h, &h->sched, 2));
  // flag words + the exchange area of the hand-rolled small-vector all-reduce (16 rank slots), one IPC-exported allocation
  h->xchg_stride = (static_cast<int>(h->arena_n - h->nW) + 4 + 3) / 4 * 4;
  CREATE_TRY(dev_alloc(h, &h->dp_flags, TFK_DP_FLAG_WORDS + 16 * static_cast<size_t>(h->xchg_stride)));
  {
    cudaError_t e = cudaMallocHost(reinterpret_cast<void**>(&h->acc_host), 2 * sizeof(double));
    if (e != cudaSuccess) return bail(fail(h, TFK_ECUDA, "cudaMallocHost: %s", cudaGetErrorString(e)));
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return bail(fail(h, TFK_ECUDA, "create sync: %s", cudaGetErrorString(e)));
  }
#undef CREATE_TRY
  *out = h;
  return TFK_OK;
}

int tfk_set_tensor(tfk_handle* h, int kind, int layer, const float* src, size_t count, void* stream) {
  if (!h || !src) return fail(h, TFK_EINVAL, "tfk_set_tensor: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* base; int rows, cols, ld;
  TFK_TRY(tensor_view(h, kind, layer, &base, &rows, &cols, &ld));
  if (count != static_cast<size_t>(rows) * cols)
    return fail(h, TFK_ESHAPE, "tfk_set_tensor: kind %d layer %d expects %d x %d elements, got %zu", kind, layer, rows, cols, count);
  TFK_CUDA(h, cudaSetDevice(h->cfg.device));
  TFK_CUDA(h, cudaMemcpy2DAsync(base, static_cast<size_t>(ld) * 4, src, static_cast<size_t>(cols) * 4,
                                static_cast<size_t>(cols) * 4, rows, cudaMemcpyDefault, st));
  if (kind == TFK_T_WEIGHTS) TFK_TRY(refresh_shadow(h, h->layers[layer], st));
  TFK_CUDA(h, cudaStreamSynchronize(st));  // src may be pageable host memory owned by the caller
  return TFK_OK;
}

int tfk_get_tensor(tfk_handle* h, int kind, int layer, float* dst, size_t count, void* stream) {
  if (!h || !dst) return fail(h, TFK_EINVAL, "tfk_get_tensor: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* base; int rows, cols, ld;
  TFK_TRY(tensor_view(h, kind, layer, &base, &rows, &cols, &ld));
  if (count != static_cast<size_t>(rows) * cols)
    return fail(h, TFK_ESHAPE, "tfk_get_tensor: kind %d layer %d holds %d x %d elements, asked %zu", kind, layer, rows, cols, count);
  TFK_CUDA(h, cudaSetDevice(h->cfg.device));
  if (kind == TFK_T_WEIGHTS || kind == TFK_T_ADAM_M_W || kind == TFK_T_ADAM_V_W)
    TFK_TRY(sync_sharded_params(h, st));  // collective under sharded data parallelism (see header)
  TFK_CUDA(h, cudaMemcpy2DAsync(dst, static_cast<size_t>(cols) * 4, base, static_cast<size_t>(ld) * 4,
                                static_cast<size_t>(cols) * 4, rows, cudaMemcpyDefault, st));
  TFK_CUDA(h, cudaStreamSynchronize(st));
  return TFK_OK;
}

int tfk_set_scalar(tfk_handle* h, int kind, double value) {
  if (!h) return TFK_EINVAL;
  switch (kind) {
    case TFK_S_GLOBAL_STEP: h->global_step = static_cast<long long>(value); return TFK_OK;
    case TFK_S_ADAM_STEP: h->adam_step = static_cast<long long>(value); return TFK_OK;
    case TFK_S_LR_FACT: h->lr_fact = value; return TFK_OK;
    case TFK_S_ACTIVE_LAYERS: return tfk_set_active_layers(h, static_cast<int>(value));
    default: return fail(h, TFK_EINVAL, "tfk_set_scalar: kind %d is not writable", kind);
  }
}

int tfk_get_scalar(tfk_handle* h, int kind, double* value_host, void* stream) {
  if (!h || !value_host) return fail(h, TFK_EINVAL, "tfk_get_scalar: null argument");
  switch (kind) {
    case TFK_S_GLOBAL_STEP: *value_host = static_cast<double>(h->global_step); return TFK_OK;
    case TFK_S_ADAM_STEP: *value_host = static_cast<double>(h->adam_step); return TFK_OK;
    case TFK_S_LR_FACT: *value_host = h->lr_fact; return TFK_OK;
    case TFK_S_ACTIVE_LAYERS: *value_host = h->active; return TFK_OK;
    case TFK_S_LOSS_SUM:
    case TFK_S_NUM_FRAMES: {
      cudaStream_t st = static_cast<cudaStream_t>(stream);
      TFK_CUDA(h, cudaMemcpyAsync(h->acc_host, h->acc, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
      TFK_CUDA(h, cudaStreamSynchronize(st));
      h->loss_pending = false;  // acc_host now holds the running sums, not the last step's
      *value_host = h->acc_host[kind == TFK_S_LOSS_SUM ? 0 : 1];
      return TFK_OK;
    }
    default: return fail(h, TFK_EINVAL, "tfk_get_scalar: unknown kind %d", kind);
  }
}

int tfk_halve_lr(tfk_handle* h) {
  if (!h) return TFK_EINVAL;
  h->lr_fact *= 0.5;  // learning_rate_fact.assign(learning_rate_fact/2)   trainer.py:141-142
  return TFK_OK;
}

int tfk_set_active_layers(tfk_handle* h, int n) {
  if (!h) return TFK_EINVAL;
  if (n < 1 || n > h->L) return fail(h, TFK_EINVAL, "tfk_set_active_layers: %d not in [1,%d]", n, h->L);
  h->active = n;
  return TFK_OK;
}

int tfk_set_dropout_seed(tfk_handle* h, uint64_t seed) {
  if (!h) return TFK_EINVAL;
  h->drop_seed = seed;
  return TFK_OK;
}

int tfk_fflayer_fwd(tfk_handle* h, int layer, const float* x, float* y, int B, int training, void* stream) {
  if (!h || !x || !y) return fail(h, TFK_EINVAL, "tfk_fflayer_fwd: null argument");
  if (layer < 0 || layer > h->L) return fail(h, TFK_EINVAL, "tfk_fflayer_fwd: layer %d", layer);
  if (layer < h->L && layer >= h->active) return fail(h, TFK_EINVAL, "tfk_fflayer_fwd: layer %d is not active", layer);
  if (h->cfg.l2_norm && layer < h->L) return fail(h, TFK_EINVAL, "tfk_fflayer_fwd: per-layer entry points do not cover L2Norm chains");
  TFK_TRY(check_frames(h, B, "tfk_fflayer_fwd"));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TFK_CUDA(h, cudaSetDevice(h->cfg.device));
  Plan* plan;
  TFK_TRY(get_plan(h, B, &plan));
  TFK_TRY(load_input(h, x, B, layer, st));
  Layer& ly = h->layers[layer];
  // run exactly this layer
  const int saved_active = h->active;
  if (layer < h->L) {
    GemmParams gp = training ? plan->fwd_train[layer] : plan->fwd_eval[layer];
    // reuse forward_range for the BN tail by temporarily narrowing the range
    h->active = layer + 1;
    int rc = forward_range(h, *plan, B, training != 0, layer, false, st);
    h->active = saved_active;
    (void)gp;
    TFK_TRY(rc);
    TimerScope ts(h, st, TFK_TIMER_CONVERT);
    TFK_LAUNCH(h, k_merge_bf16(h->act_hi[layer + 1], h->act_lo[layer + 1], ly.ldn, y, ly.N, B, ly.N, st));
  } else {
    {
      TimerScope ts(h, st, TFK_TIMER_GEMM_FWD);
      TFK_LAUNCH(h, gemm_launch(training ? plan->fwd_train[layer] : plan->fwd_eval[layer], h->num_sms, st));
    }
    TFK_CUDA(h, cudaMemcpy2DAsync(y, static_cast<size_t>(ly.N) * 4, h->logits, static_cast<size_t>(h->ldo) * 4,
                                  static_cast<size_t>(ly.N) * 4, B, cudaMemcpyDeviceToDevice, st));
  }
  return TFK_OK;
}

int tfk_fflayer_bwd(tfk_handle* h, int layer, const float* dy, float* dx, int B, void* stream) {
  if (!h || !dy) return fail(h, TFK_EINVAL, "tfk_fflayer_bwd: null argument");
  if (layer < 0 || layer > h->L) return fail(h, TFK_EINVAL, "tfk_fflayer_bwd: layer %d", layer);
  if (layer < h->L && layer >= h->active) return fail(h, TFK_EINVAL, "tfk_fflayer_bwd: layer %d is not active", layer);
  if (h->cfg.l2_norm && layer < h->L) return fail(h, TFK_EINVAL, "tfk_fflayer_bwd: per-layer entry points do not cover L2Norm chains");
  TFK_TRY(check_frames(h, B, "tfk_fflayer_bwd"));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TFK_CUDA(h, cudaSetDevice(h->cfg.device));
  Plan* plan;
  TFK_TRY(get_plan(h, B, &plan));
  Layer& ly = h->layers[layer];
  const int L = h->L;
  const int ai = layer == L ? h->active : layer;
  if (dx && ai == 0) return fail(h, TFK_EINVAL, "tfk_fflayer_bwd: layer 0 has no dx (the reference never forms d/d input)");
  __nv_bfloat16* dz_hi = ly.hidden ? h->dA_hi[(L - 1 - layer) & 1] : h->dzo_hi;
  __nv_bfloat16* dz_lo = ly.hidden ? h->dA_lo[(L - 1 - layer) & 1] : h->dzo_lo;
  const int ldz = ly.hidden ? ly.ldn : h->ldo;
  {
    // (1) the layer's own activation chain, backward: relu / dropout mask from the stored forward output
    TimerScope ts(h, st, TFK_TIMER_CONVERT, 3);
    const float* src = dy;
    const bool relu = h->cfg.nonlin == TFK_NONLIN_RELU, drop = h->cfg.keep_prob < 1.0f;
    const bool smooth = h->cfg.nonlin == TFK_NONLIN_SIGMOID || h->cfg.nonlin == TFK_NONLIN_TANH;
    if (ly.hidden && (relu || drop || smooth)) {
      TFK_LAUNCH(h, k_merge_bf16(h->act_hi[layer + 1], h->act_lo[layer + 1], ly.ldn, h->tmp_f32, ly.N, B, ly.N, st));
      TFK_LAUNCH(h, k_mask_scale_f32(dy, h->tmp_f32, h->tmp_f32, static_cast<size_t>(B) * ly.N,
                                     drop ? 1.0f / h->cfg.keep_prob : 1.0f,
                                     (relu ? 0 : h->cfg.nonlin == TFK_NONLIN_SIGMOID ? 2 : h->cfg.nonlin == TFK_NONLIN_TANH ? 3 : 1) | (drop ? 4 : 0), st));
      src = h->tmp_f32;
    }
    TFK_LAUNCH(h, k_split_f32(src, ly.N, dz_hi, h->x3 ? dz_lo : nullptr, ldz, B, ly.N, st));
  }
  // (2) [BN backward] + dbias + wgrad, and (3) the raw dgrad dx = dz . W^T (no mask of the layer below)
  GemmSpec s[2];
  s[0].nsplit = s[1].nsplit = h->x3 ? 3 : 1;
  s[0].M = ly.K; s[0].N = ly.N; s[0].K = B;
  s[0].A_hi = h->act_hi[ai]; s[0].A_lo = h->act_lo[ai]; s[0].lda = ly.ldk; s[0].a_mn = 1;
  s[0].B_hi = dz_hi; s[0].B_lo = dz_lo; s[0].ldb = ldz; s[0].b_mn = 1;
  s[0].out_kind = OUT_F32_REDADD; s[0].D_hi = h->G + ly.off_w; s[0].ldd = ly.ldn;
  const int dst = (L - layer) & 1;  // any dA buffer other than the one holding dz
  if (dx) {
    s[1].M = B; s[1].N = ly.K; s[1].K = ly.N;
    s[1].A_hi = dz_hi; s[1].A_lo = dz_lo; s[1].lda = ldz; s[1].a_mn = 0;
    s[1].B_hi = h->Sh + ly.off_w; s[1].B_lo = opt(h->Sl, ly.off_w); s[1].ldb = ly.ldn; s[1].b_mn = 0;
    s[1].out_kind = h->x3 ? OUT_BF16_SPLIT : OUT_BF16;
    s[1].D_hi = h->dA_hi[dst]; s[1].D_lo = h->dA_lo[dst]; s[1].ldd = ly.ldk;
  }
  GemmParams gp;
  Plan scratch;
  if (int rc = finish_params(h, scratch, s, dx ? 2 : 1, &gp, "tfk_fflayer_bwd", layer)) {
    free_plan(scratch);
    return rc;
  }
  TFK_TRY(backward_layer(h, *plan, B, layer, st, &gp));  // (the work list it uses is owned by h->list_cache)
  if (dx) {
    TimerScope ts(h, st, TFK_TIMER_CONVERT);
    TFK_LAUNCH(h, k_merge_bf16(h->dA_hi[dst], h->dA_lo[dst], ly.ldk, dx, ly.K, B, ly.K, st));
  }
  return TFK_OK;
}

int tfk_softmax_ce(tfk_handle* h, const float* logits, const int32_t* labels, int B, float* loss_sum,
                   float* dlogits, void* stream) {
  if (!h || !logits || !labels || !loss_sum) return fail(h, TFK_EINVAL, "tfk_softmax_ce: null argument");
  TFK_TRY(check_frames(h, B, "tfk_softmax_ce"));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TFK_CUDA(h, cudaSetDevice(h->cfg.device));
  const int O = h->cfg.output_dim;
  TFK_CUDA(h, cudaMemcpy2DAsync(h->logits, static_cast<size_t>(h->ldo) * 4, logits, static_cast<size_t>(O) * 4,
                                static_cast<size_t>(O) * 4, B, cudaMemcpyDeviceToDevice, st));
  {
    TimerScope ts(h, st, TFK_TIMER_SOFTMAX_CE);
    // full-precision gradient for the stand-alone entry: always emit hi+lo
    __nv_bfloat16* lo = h->x3 ? h->dzo_lo : reinterpret_cast<__nv_bfloat16*>(h->tmp_f32);
    TFK_LAUNCH(h, k_softmax_ce(h->logits, h->ldo, labels, B, O, h->row_loss, h->dzo_hi, lo, st));
    if (dlogits) TFK_LAUNCH(h, k_merge_bf16(h->dzo_hi, lo, h->ldo, dlogits, O, B, O, st));
  }
  // loss_sum = sum(row_loss): reuse the deterministic reducer on a scratch accumulator
  double* scratch = reinterpret_cast<double*>(h->ws);
  TFK_CUDA(h, cudaMemsetAsync(scratch, 0, 2 * sizeof(double), st));
  TFK_LAUNCH(h, k_accum_loss(h->row_loss, B, scratch, st));
  TFK_LAUNCH(h, k_double_to_float(scratch, loss_sum, st));
  h->launches += 2;
  return TFK_OK;
}

int tfk_accumulate(tfk_handle* h, const float* x, const int32_t* labels, int B, void* stream) {
  if (!h || !x || !labels) return fail(h, TFK_EINVAL, "tfk_accumulate: null argument");
  TFK_TRY(check_frames(h, B, "tfk_accumulate"));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TFK_CUDA(h, cudaSetDevice(h->cfg.device));
  Plan* plan;
  TFK_TRY(get_plan(h, B, &plan));
  TFK_TRY(load_input(h, x, B, 0, st));
  TFK_TRY(forward_range(h, *plan, B, true, 0, true, st));
  TFK_TRY(ce_and_backward(h, *plan, labels, B, true, st));
  advance_drop_seed(h);
  h->pending_accumulates += 1;
  return TFK_OK;
}

static int check_raw(tfk_handle* h, const char* what, int feat_dim, int context, int num_utts) {
  if (feat_dim < 1 || context < 0 || num_utts < 1)
    return fail(h, TFK_EINVAL, "%s: feat_dim=%d context=%d num_utts=%d", what, feat_dim, context, num_utts);
  if (feat_dim * (2 * context + 1) != h->cfg.input_dim)
    return fail(h, TFK_ESHAPE, "%s: feat_dim %d x (2*%d+1) != input_dim %d", what, feat_dim, context, h->cfg.input_dim);
  return TFK_OK;
}

static int load_raw(tfk_handle* h, const float* raw, const int32_t* utt_off, int num_utts, const float* cmvn,
                    int feat_dim, int context, int row_begin, int rows, cudaStream_t st) {
  TimerScope ts(h, st, TFK_TIMER_CONVERT);
  TFK_LAUNCH(h, k_splice_cmvn(raw, utt_off, num_utts, cmvn, feat_dim, context, row_begin, rows, h->act_hi[0],
                              h->act_lo[0], h->ld0, st));
  return TFK_OK;
}

int tfk_accumulate_raw(tfk_handle* h, const float* raw, const int32_t* utt_offsets, int num_utts, const float* cmvn,
                       const int32_t* labels, int R, int feat_dim, int context, void* stream) {
  if (!h || !raw || !utt_offsets || !cmvn || !labels) return fail(h, TFK_EINVAL, "tfk_accumulate_raw: null argument");
  TFK_TRY(check_frames(h, R, "tfk_accumulate_raw"));
  TFK_TRY(check_raw(h, "tfk_accumulate_raw", feat_dim, context, num_utts));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TFK_CUDA(h, cudaSetDevice(h->cfg.device));
  Plan* plan;
  TFK_TRY(get_plan(h, R, &plan));
  TFK_TRY(load_raw(h, raw, utt_offsets, num_utts, cmvn, feat_dim, context, 0, R, st));
  TFK_TRY(forward_range(h, *plan, R, true, 0, true, st));
  TFK_TRY(ce_and_backward(h, *plan, labels, R, true, st));
  advance_drop_seed(h);
  h->pending_accumulates += 1;
  return TFK_OK;
}

int tfk_forward_loglik_raw_rows(tfk_handle* h, const float* raw, const int32_t* utt_offsets, int num_utts, const float* cmvn,
                                int row_begin, int rows, int feat_dim, int context, const float* prior, float* out,
                                void* stream) {
  if (!h || !raw || !utt_offsets || !cmvn || !out) return fail(h, TFK_EINVAL, "tfk_forward_loglik_raw: null argument");
  if (rows <= 0 || row_begin < 0) return fail(h, TFK_ESHAPE, "tfk_forward_loglik_raw: rows [%d, +%d) must be a non-empty range", row_begin, rows);
  TFK_TRY(check_raw(h, "tfk_forward_loglik_raw", feat_dim, context, num_utts));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TFK_CUDA(h, cudaSetDevice(h->cfg.device));
  const int O = h->cfg.output_dim, maxB = h->cfg.max_frames;
  const float* log_prior = nullptr;
  if (prior) {
    TFK_LAUNCH(h, k_log_vector(prior, h->tmp_f32, O, st));
    h->launches += 1;
    log_prior = h->tmp_f32;
  }
  for (int t0 = 0; t0 < rows; t0 += maxB) {  // the splice reads across chunk borders straight from `raw`
    const int B = (rows - t0 < maxB) ? (rows - t0) : maxB;
    Plan* plan;
    TFK_TRY(get_plan(h, B, &plan));
    TFK_TRY(load_raw(h, raw, utt_offsets, num_utts, cmvn, feat_dim, context, row_begin + t0, B, st));
    TFK_TRY(forward_range(h, *plan, B, false, 0, true, st));
    TimerScope ts(h, st, TFK_TIMER_DECODE_OUT);
    TFK_LAUNCH(h, k_decode_out(h->logits, h->ldo, B, O, log_prior, out + static_cast<size_t>(t0) * O, st));
  }
  return TFK_OK;
}

int tfk_forward_loglik_raw(tfk_handle* h, const float* raw, const int32_t* utt_offsets, int num_utts, const float* cmvn,
                           int R, int feat_dim, int context, const float* prior, float* out, void* stream) {
  return tfk_forward_loglik_raw_rows(h, raw, utt_offsets, num_utts, cmvn, 0, R, feat_dim, context, prior, out, stream);
}

int tfk_apply(tfk_handle* h, float lr, float* mean_loss_host, void* stream) {
  if (!h) return TFK_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TFK_CUDA(h, cudaSetDevice(h->cfg.device));
  h->pending_accumulates = 0;
  // Fused transports on every layer: nothing is left for NCCL but the small replicated vectors and {loss, frames} —
  // those go over peer memory too (push into every GPU's exchange slot, flags, rank-ordered sum: ~15 us against the
  // 35-60 us of a latency-bound NCCL all-reduce on the step's tail).  The flag wait doubles as the barrier that makes
  // every peer's remote reduce-adds into this rank's accumulators complete before Adam reads them.
  bool all_fused = h->sharded && h->fused_rs && h->fused_ag && h->nranks <= 16;
  for (int l = 0; l <= h->L && all_fused; ++l) all_fused = h->layers[l].peer_reduce;
  static const bool nccl_small = [] {  // TFK_DP_SMALL=nccl keeps the NCCL all-reduce (A/B measurements)
    const char* e = getenv("TFK_DP_SMALL");
    return e && strcmp(e, "nccl") == 0;
  }();
  const bool peer_small = all_fused && !nccl_small;
  if (peer_small) {
    TimerScope ts(h, st, TFK_TIMER_ALLREDUCE, 4);
    h->dp_epoch += 1;
    const int small_n = static_cast<int>(h->arena_n - h->nW);
    TFK_LAUNCH(h, k_dp_small_push(h->d_peer_flags, h->nranks - 1, h->dp_flags, h->rank, h->G + h->nW, h->acc, small_n, h->xchg_stride, st));
    TFK_LAUNCH(h, k_dp_publish(h->d_peer_flags, h->nranks - 1, TFK_DP_FLAG_SLOTS - 1, h->rank, h->dp_epoch, st));
    TFK_LAUNCH(h, k_dp_wait(h->dp_flags, h->nranks, TFK_DP_FLAG_SLOTS - 1, h->rank, h->dp_epoch, st));
    TFK_LAUNCH(h, k_dp_small_reduce(h->dp_flags, h->nranks, h->G + h->nW, h->acc, small_n, h->xchg_stride, st));
  } else if (h->sharded) TFK_TRY(reduce_scatter_grads(h, st));
  else TFK_TRY(allreduce_grads(h, st));
  h->global_step += 1;  // apply_gradients(global_step=...)   trainer.py:182-184
  h->adam_step += 1;    // beta1_power / beta2_power advance with every apply, whatever global_step is restored to
  const double t = static_cast<double>(h->adam_step);
  const double b1 = h->cfg.adam_beta1, b2 = h->cfg.adam_beta2;
  const double lr_eff = static_cast<double>(lr) * h->lr_fact;
  const float lr_t = static_cast<float>(lr_eff * std::sqrt(1.0 - std::pow(b2, t)) / (1.0 - std::pow(b1, t)));
  {
    TimerScope ts(h, st, TFK_TIMER_ADAM);
    size_t off[80], cnt[80];
    int n = 0;
    if (h->sharded) {  // this rank's slice of every layer's weights + the replicated small vectors
      for (int l = 0; l <= h->L; ++l) {
        const size_t c = h->layers[l].w_count / h->nranks;
        off[n] = h->layers[l].off_w + c * h->rank; cnt[n] = c; ++n;
      }
      off[n] = h->nW; cnt[n] = h->arena_n - h->nW; ++n;
    } else {
      off[0] = 0; cnt[0] = h->arena_n; n = 1;
    }
    std::vector<__nv_bfloat16*> phi, plo;
    if (h->sharded && h->fused_ag)
      for (int r = 0; r < h->nranks; ++r)
        if (r != h->rank) { phi.push_back(h->peer_Sh[r]); plo.push_back(h->peer_Sl[r]); }
    // fused update -> all-gather: the weight slices (all segments but the last) are also stored into the peers
    TFK_LAUNCH(h, k_adam(h->P, h->G, h->M, h->V, h->Sh, h->x3 ? h->Sl : nullptr, off, cnt, n, h->acc, lr_t,
                         h->cfg.adam_beta1, h->cfg.adam_beta2, h->cfg.adam_eps, st, phi.empty() ? 0 : n - 1,
                         static_cast<int>(phi.size()), phi.data(), h->x3 ? plo.data() : nullptr));
  }
  if (h->sharded) {
    // non-owned gradient slices still hold this rank's local sums: clear them for the next accumulation
    for (int l = 0; l <= h->L; ++l) {
      const Layer& ly = h->layers[l];
      if (h->fused_rs && ly.peer_reduce) continue;  // non-owned slices were never written locally
      const size_t c = ly.w_count / h->nranks, lo = c * h->rank;
      if (lo) TFK_CUDA(h, cudaMemsetAsync(h->G + ly.off_w, 0, lo * sizeof(float), st));
      if (lo + c < ly.w_count)
        TFK_CUDA(h, cudaMemsetAsync(h->G + ly.off_w + lo + c, 0, (ly.w_count - lo - c) * sizeof(float), st));
    }
    if (h->fused_ag) {
      // every rank has stored its refreshed operand slices into all peers: publish, then wait for the others
      TimerScope ts(h, st, TFK_TIMER_ALLREDUCE, 2);
      if (!peer_small) h->dp_epoch += 1;
      TFK_LAUNCH(h, k_dp_publish(h->d_peer_flags, h->nranks - 1, 0, h->rank, h->dp_epoch, st));
      TFK_LAUNCH(h, k_dp_wait(h->dp_flags, h->nranks, 0, h->rank, h->dp_epoch, st));
    } else {
      TFK_TRY(all_gather_shadows(h, st));
    }
    h->params_synced = false;
  }
  TFK_CUDA(h, cudaMemcpyAsync(h->acc_host, h->acc, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  TFK_CUDA(h, cudaMemsetAsync(h->acc, 0, 2 * sizeof(double), st));  // init_loss / init_num_frames
  h->loss_pending = mean_loss_host == nullptr;
  if (mean_loss_host) {
    TFK_CUDA(h, cudaStreamSynchronize(st));
    *mean_loss_host = static_cast<float>(h->acc_host[0] / h->acc_host[1]);  // average_loss   trainer.py:198
  } else {
    TFK_TRY(bound_runahead(h, st));
  }
  return TFK_OK;
}

int tfk_last_loss(tfk_handle* h, float* mean_loss_host, void* stream) {
  if (!h || !mean_loss_host) return fail(h, TFK_EINVAL, "tfk_last_loss: null argument");
  if (!h->loss_pending) return fail(h, TFK_EINVAL, "tfk_last_loss: no step with an unread loss is outstanding");
  TFK_CUDA(h, cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  *mean_loss_host = static_cast<float>(h->acc_host[0] / h->acc_host[1]);
  h->loss_pending = false;
  return TFK_OK;
}

// One whole optimizer step for the common case of ONE micro-batch on ONE GPU
// (== tfk_accumulate + tfk_apply, same arithmetic): the clip+Adam update of a layer's weights is launched
// on a side stream as soon as that layer's fused wgrad/dgrad kernel has finished, so the HBM-bound Adam
// pass overlaps the tensor-bound backward kernels of the layers below instead of following them.
// the step proper, once layer 0's operand is in place (load_input or load_raw)
static int train_step_loaded(tfk_handle* h, Plan* plan, const int32_t* labels, int B, float lr, float* mean_loss_host,
                             cudaStream_t st) {
  if (!h->adam_stream) {
    TFK_CUDA(h, cudaStreamCreateWithFlags(&h->adam_stream, cudaStreamNonBlocking));
    h->layer_events.resize(h->L + 2);
    for (auto& e : h->layer_events) TFK_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->colsum_events.resize(h->L + 1);
    for (auto& e : h->colsum_events) TFK_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  TFK_TRY(forward_range(h, *plan, B, true, 0, true, st));
  {
    TimerScope ts(h, st, TFK_TIMER_SOFTMAX_CE, 2);
    TFK_LAUNCH(h, k_softmax_ce(h->logits, h->ldo, labels, B, h->cfg.output_dim, h->row_loss, h->dzo_hi,
                               h->x3 ? h->dzo_lo : nullptr, st));
    TFK_LAUNCH(h, k_accum_loss(h->row_loss, B, h->acc, st));  // acc = {loss_sum, frames}: Adam reads frames
  }
  const bool dp = h->comm != nullptr && h->nranks > 1;  // (dp_overlap_ready() held: fused reduce-scatter + all-gather, all layers)
  std::vector<__nv_bfloat16*> phi, plo;
  if (dp) {
    // the mean gradient divides by the GLOBAL frame count (trainer.py:174): sum {loss, frames} over the ranks now, on the
    // communication stream, so that it overlaps the output layer's backward kernel; every update below waits for it
    NcclApi& api = nccl();
    TFK_CUDA(h, cudaEventRecord(h->ev_compute, st));
    TFK_CUDA(h, cudaStreamWaitEvent(h->comm_stream, h->ev_compute, 0));
    {
      TimerScope ts(h, h->comm_stream, TFK_TIMER_ALLREDUCE);
      const int rc = api.AllReduce(h->acc, h->acc, 2, kNcclDouble, kNcclSum, h->comm, h->comm_stream);
      if (rc) return fail(h, TFK_ENCCL, "ncclAllReduce {loss, frames} failed: %s", api.GetErrorString ? api.GetErrorString(rc) : "?");
    }
    TFK_CUDA(h, cudaEventRecord(h->ev_comm, h->comm_stream));
    TFK_CUDA(h, cudaStreamWaitEvent(h->adam_stream, h->ev_comm, 0));
    h->dp_epoch += 1;
    for (int r = 0; r < h->nranks; ++r)
      if (r != h->rank) { phi.push_back(h->peer_Sh[r]); plo.push_back(h->peer_Sl[r]); }
  }
  h->global_step += 1;
  h->adam_step += 1;
  const double t = static_cast<double>(h->adam_step);
  const double b1 = h->cfg.adam_beta1, b2 = h->cfg.adam_beta2;
  const float lr_t = static_cast<float>(static_cast<double>(lr) * h->lr_fact * std::sqrt(1.0 - std::pow(b2, t)) /
                                        (1.0 - std::pow(b1, t)));
  auto adam_layer = [&](int l) -> int {  // weights of layer l, after its backward kernel, on the side stream
    TFK_CUDA(h, cudaEventRecord(h->layer_events[l], st));
    TFK_CUDA(h, cudaStreamWaitEvent(h->adam_stream, h->layer_events[l], 0));
    if (!dp) {
      TimerScope ts(h, h->adam_stream, TFK_TIMER_ADAM);
      const size_t off = h->layers[l].off_w, cnt = h->layers[l].w_count;
      TFK_LAUNCH(h, k_adam(h->P, h->G, h->M, h->V, h->Sh, h->x3 ? h->Sl : nullptr, &off, &cnt, 1, h->acc, lr_t,
                           h->cfg.adam_beta1, h->cfg.adam_beta2, h->cfg.adam_eps, h->adam_stream));
      return TFK_OK;
    }
    // Data parallel.  This rank's backward kernel of layer l is done: its wgrad epilogue has reduce-added every slab into
    // the slice owner's accumulator over NVLink.  Tell the peers, wait until all of THEIR layer-l kernels are done too —
    // from then on this rank's slice of dW_l is the global sum and nobody reads W_l any more in this step — then update
    // the slice and store its refreshed bf16 operands into every peer, all under the backward kernels of the layers below.
    {
      TimerScope ts(h, h->adam_stream, TFK_TIMER_ALLREDUCE, 2);
      TFK_LAUNCH(h, k_dp_publish(h->d_peer_flags, h->nranks - 1, 1 + l, h->rank, h->dp_epoch, h->adam_stream));
      TFK_LAUNCH(h, k_dp_wait(h->dp_flags, h->nranks, 1 + l, h->rank, h->dp_epoch, h->adam_stream));
    }
    TimerScope ts(h, h->adam_stream, TFK_TIMER_ADAM);
    const size_t cnt = h->layers[l].w_count / h->nranks, off = h->layers[l].off_w + cnt * h->rank;
    TFK_LAUNCH(h, k_adam(h->P, h->G, h->M, h->V, h->Sh, h->x3 ? h->Sl : nullptr, &off, &cnt, 1, h->acc, lr_t,
                         h->cfg.adam_beta1, h->cfg.adam_beta2, h->cfg.adam_eps, h->adam_stream, 1,
                         static_cast<int>(phi.size()), phi.data(), h->x3 ? plo.data() : nullptr));
    return TFK_OK;
  };
  // bias-gradient column sums that are not produced by a GEMM epilogue (output layer; BN / L2Norm layers) go to the
  // side stream: they are only needed by the final small Adam launch
  TFK_TRY(backward_layer(h, *plan, B, h->L, st, nullptr, true, h->adam_stream));
  TFK_TRY(adam_layer(h->L));
  for (int l = h->active - 1; l >= 0; --l) {
    TFK_TRY(backward_layer(h, *plan, B, l, st, nullptr, true, h->adam_stream));
    TFK_TRY(adam_layer(l));
  }
  {
    const float* parts[64];
    float* outs[64];
    int n = 0;
    for (int l = 0; l < h->active; ++l)
      if (!h->layers[l].bn && !h->cfg.l2_norm) {
        parts[n] = h->layers[l].db_part;
        outs[n] = h->G + h->layers[l].off_b;
        ++n;
      }
    if (n > 0) {
      TimerScope ts(h, st, TFK_TIMER_COLSUM);
      TFK_LAUNCH(h, k_colsum_finalize(parts, outs, n, (B + 31) / 32, h->ldh, h->cfg.hidden_dim, st));
    }
  }
  // join the side stream (output-layer column sums, per-layer Adam launches) before the bias update
  TFK_CUDA(h, cudaEventRecord(h->layer_events[h->L + 1], h->adam_stream));
  TFK_CUDA(h, cudaStreamWaitEvent(st, h->layer_events[h->L + 1], 0));
  if (dp) {  // the small replicated vectors (biases, betas): one NCCL all-reduce, identical sums on every rank
    NcclApi& api = nccl();
    TimerScope ts(h, st, TFK_TIMER_ALLREDUCE);
    const int rc = api.AllReduce(h->G + h->nW, h->G + h->nW, h->arena_n - h->nW, kNcclFloat, kNcclSum, h->comm, st);
    if (rc) return fail(h, TFK_ENCCL, "ncclAllReduce (biases) failed: %s", api.GetErrorString ? api.GetErrorString(rc) : "?");
  }
  {  // biases / betas (small), then inactive layers' weight regions (zero gradients: a no-op update, kept for
     // exact equivalence with tfk_apply, which always covers the whole arena)
    TimerScope ts(h, st, TFK_TIMER_ADAM);
    size_t off[70], cnt[70];
    int n = 0;
    off[n] = h->nW; cnt[n] = h->arena_n - h->nW; ++n;
    if (!dp)
      for (int l = h->active; l < h->L; ++l) { off[n] = h->layers[l].off_w; cnt[n] = h->layers[l].w_count; ++n; }
    TFK_LAUNCH(h, k_adam(h->P, h->G, h->M, h->V, h->Sh, h->x3 ? h->Sl : nullptr, off, cnt, n, h->acc, lr_t,
                         h->cfg.adam_beta1, h->cfg.adam_beta2, h->cfg.adam_eps, st));
  }
  if (dp) {
    // every rank has stored the refreshed operands of its slices into all peers: publish, then wait for the others —
    // the next forward pass (and the next step's remote reduce-adds into the accumulators) start behind this
    TimerScope ts(h, st, TFK_TIMER_ALLREDUCE, 2);
    TFK_LAUNCH(h, k_dp_publish(h->d_peer_flags, h->nranks - 1, 0, h->rank, h->dp_epoch, st));
    TFK_LAUNCH(h, k_dp_wait(h->dp_flags, h->nranks, 0, h->rank, h->dp_epoch, st));
    h->params_synced = false;
  }
  advance_drop_seed(h);
  TFK_CUDA(h, cudaMemcpyAsync(h->acc_host, h->acc, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  TFK_CUDA(h, cudaMemsetAsync(h->acc, 0, 2 * sizeof(double), st));
  h->loss_pending = mean_loss_host == nullptr;
  if (mean_loss_host) {
    TFK_CUDA(h, cudaStreamSynchronize(st));
    *mean_loss_host = static_cast<float>(h->acc_host[0] / h->acc_host[1]);
  } else {
    TFK_TRY(bound_runahead(h, st));
  }
  return TFK_OK;
}

// Data parallel: the overlapped step needs the fused transports on every layer in use (peer memory on one node,
// TFK_DP_MODE unset); anything else takes tfk_accumulate + tfk_apply.  It is OPT-IN (TFK_DP_OVERLAP=1): measured at two
// GPUs (profiles/r2dp_bench_2gpu_*.json) the serial schedule is faster, 1.18 vs 1.28 ms/step — with half of every layer
// to update per rank the side-stream Adam + peer stores cost the backward GEMMs more (62 us) than the tail they hide;
// the trade reverses only when the per-rank slices are small and the NVLink-bound tail is long (8 ranks).
static bool dp_overlap_ready(const tfk_handle* h) {
  if (!(h->sharded && h->fused_rs && h->fused_ag)) return false;
  static const bool on = [] {
    const char* e = getenv("TFK_DP_OVERLAP");
    return e && e[0] == '1';
  }();
  if (!on) return false;
  for (int l = 0; l <= h->L; ++l) {
    if (l < h->L && l >= h->active) continue;
    if (!h->layers[l].peer_reduce) return false;
  }
  return true;
}

int tfk_train_step(tfk_handle* h, const float* x, const int32_t* labels, int B, float lr, float* mean_loss_host,
                   void* stream) {
  if (!h || !x || !labels) return fail(h, TFK_EINVAL, "tfk_train_step: null argument");
  if ((h->comm != nullptr && !dp_overlap_ready(h)) || h->pending_accumulates > 0) {  // plain sequence
    TFK_TRY(tfk_accumulate(h, x, labels, B, stream));
    return tfk_apply(h, lr, mean_loss_host, stream);
  }
  TFK_TRY(check_frames(h, B, "tfk_train_step"));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TFK_CUDA(h, cudaSetDevice(h->cfg.device));
  Plan* plan;
  TFK_TRY(get_plan(h, B, &plan));
  TFK_TRY(load_input(h, x, B, 0, st));
  return train_step_loaded(h, plan, labels, B, lr, mean_loss_host, st);
}

// tfk_train_step fed by the device-side CMVN + splice feeder (== tfk_accumulate_raw + tfk_apply)
int tfk_train_step_raw(tfk_handle* h, const float* raw, const int32_t* utt_offsets, int num_utts, const float* cmvn,
                       const int32_t* labels, int R, int feat_dim, int context, float lr, float* mean_loss_host,
                       void* stream) {
  if (!h || !raw || !utt_offsets || !cmvn || !labels) return fail(h, TFK_EINVAL, "tfk_train_step_raw: null argument");
  if ((h->comm != nullptr && !dp_overlap_ready(h)) || h->pending_accumulates > 0) {
    TFK_TRY(tfk_accumulate_raw(h, raw, utt_offsets, num_utts, cmvn, labels, R, feat_dim, context, stream));
    return tfk_apply(h, lr, mean_loss_host, stream);
  }
  TFK_TRY(check_frames(h, R, "tfk_train_step_raw"));
  TFK_TRY(check_raw(h, "tfk_train_step_raw", feat_dim, context, num_utts));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TFK_CUDA(h, cudaSetDevice(h->cfg.device));
  Plan* plan;
  TFK_TRY(get_plan(h, R, &plan));
  TFK_TRY(load_raw(h, raw, utt_offsets, num_utts, cmvn, feat_dim, context, 0, R, st));
  return train_step_loaded(h, plan, labels, R, lr, mean_loss_host, st);
}

int tfk_eval_accumulate(tfk_handle* h, const float* x, const int32_t* labels, int B, void* stream) {
  if (!h || !x || !labels) return fail(h, TFK_EINVAL, "tfk_eval_accumulate: null argument");
  TFK_TRY(check_frames(h, B, "tfk_eval_accumulate"));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TFK_CUDA(h, cudaSetDevice(h->cfg.device));
  Plan* plan;
  TFK_TRY(get_plan(h, B, &plan));
  TFK_TRY(load_input(h, x, B, 0, st));
  TFK_TRY(forward_range(h, *plan, B, false, 0, true, st));
  TFK_TRY(ce_and_backward(h, *plan, labels, B, false, st));
  return TFK_OK;
}

int tfk_eval_finish(tfk_handle* h, float* mean_loss_host, void* stream) {
  if (!h || !mean_loss_host) return fail(h, TFK_EINVAL, "tfk_eval_finish: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TFK_CUDA(h, cudaSetDevice(h->cfg.device));
  if (h->comm && h->nranks > 1) {
    // data parallel: every rank holds a different share of the validation utterances; the decisions Nnet.train
    // takes on this number (rollback, halving, stop: nnet.py:168-207) must be identical on all ranks
    NcclApi& api = nccl();
    TimerScope ts(h, st, TFK_TIMER_ALLREDUCE);
    const int rc = api.AllReduce(h->acc, h->acc, 2, kNcclDouble, kNcclSum, h->comm, st);
    if (rc) return fail(h, TFK_ENCCL, "ncclAllReduce (validation loss) failed: %s", api.GetErrorString ? api.GetErrorString(rc) : "?");
  }
  TFK_CUDA(h, cudaMemcpyAsync(h->acc_host, h->acc, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  TFK_CUDA(h, cudaMemsetAsync(h->acc, 0, 2 * sizeof(double), st));
  TFK_CUDA(h, cudaStreamSynchronize(st));
  h->loss_pending = false;
  *mean_loss_host = static_cast<float>(h->acc_host[0] / h->acc_host[1]);
  return TFK_OK;
}

static int forward_decode(tfk_handle* h, const float* x, int T, const float* prior, float* out, cudaStream_t st) {
  if (T <= 0) return fail(h, TFK_ESHAPE, "decode: T=%d must be positive", T);
  TFK_CUDA(h, cudaSetDevice(h->cfg.device));
  const int I = h->cfg.input_dim, O = h->cfg.output_dim, maxB = h->cfg.max_frames;
  const float* log_prior = nullptr;
  if (prior) {  // log(prior) once per call, into the (otherwise idle) gradient-side scratch
    float* lp = h->tmp_f32;
    TFK_LAUNCH(h, k_log_vector(prior, lp, O, st));
    h->launches += 1;
    log_prior = lp;
  }
  for (int t0 = 0; t0 < T; t0 += maxB) {  // frames are independent: tile long utterances over the workspace
    const int B = (T - t0 < maxB) ? (T - t0) : maxB;
    Plan* plan;
    TFK_TRY(get_plan(h, B, &plan));
    TFK_TRY(load_input(h, x + static_cast<size_t>(t0) * I, B, 0, st));
    TFK_TRY(forward_range(h, *plan, B, false, 0, true, st));
    TimerScope ts(h, st, TFK_TIMER_DECODE_OUT);
    TFK_LAUNCH(h, k_decode_out(h->logits, h->ldo, B, O, log_prior, out + static_cast<size_t>(t0) * O, st));
  }
  return TFK_OK;
}

int tfk_forward_posteriors(tfk_handle* h, const float* x, int T, float* out, void* stream) {
  if (!h || !x || !out) return fail(h, TFK_EINVAL, "tfk_forward_posteriors: null argument");
  return forward_decode(h, x, T, nullptr, out, static_cast<cudaStream_t>(stream));
}

int tfk_forward_loglik(tfk_handle* h, const float* x, int T, const float* prior, float* out, void* stream) {
  if (!h || !x || !out || !prior) return fail(h, TFK_EINVAL, "tfk_forward_loglik: null argument");
  return forward_decode(h, x, T, prior, out, static_cast<cudaStream_t>(stream));
}

int tfk_get_activation(tfk_handle* h, int layer, float* dst, int B, void* stream) {
  if (!h || !dst) return fail(h, TFK_EINVAL, "tfk_get_activation: null argument");
  if (layer < 0 || layer >= h->L) return fail(h, TFK_EINVAL, "tfk_get_activation: hidden layer %d of %d", layer, h->L);
  TFK_TRY(check_frames(h, B, "tfk_get_activation"));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TFK_CUDA(h, cudaSetDevice(h->cfg.device));
  const Layer& ly = h->layers[layer];
  TimerScope ts(h, st, TFK_TIMER_CONVERT);
  TFK_LAUNCH(h, k_merge_bf16(h->act_hi[layer + 1], h->act_lo[layer + 1], ly.ldn, dst, ly.N, B, ly.N, st));
  return TFK_OK;
}

int tfk_comm_unique_id(uint8_t* id128_host) {
  if (!id128_host) return TFK_EINVAL;
  NcclApi& api = nccl();
  if (!api.ok) return fail(nullptr, TFK_ENCCL, "libnccl.so.2 not available");
  NcclApi::UniqueId id;
  const int rc = api.GetUniqueId(&id);
  if (rc) return fail(nullptr, TFK_ENCCL, "ncclGetUniqueId failed (%d)", rc);
  memcpy(id128_host, id.internal, 128);
  return TFK_OK;
}

static int setup_comm_streams(tfk_handle* h) {
  if (!h->comm_stream) {
    TFK_CUDA(h, cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
    TFK_CUDA(h, cudaEventCreateWithFlags(&h->ev_compute, cudaEventDisableTiming));
    TFK_CUDA(h, cudaEventCreateWithFlags(&h->ev_comm, cudaEventDisableTiming));
  }
  return TFK_OK;
}

int tfk_comm_init(tfk_handle* h, const uint8_t* id128_host, int rank, int nranks) {
  if (!h || !id128_host) return fail(h, TFK_EINVAL, "tfk_comm_init: null argument");
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(h, TFK_EINVAL, "tfk_comm_init: rank %d of %d", rank, nranks);
  NcclApi& api = nccl();
  if (!api.ok) return fail(h, TFK_ENCCL, "libnccl.so.2 not available");
  TFK_CUDA(h, cudaSetDevice(h->cfg.device));
  NcclApi::UniqueId id;
  memcpy(id.internal, id128_host, 128);
  void* comm = nullptr;
  const int rc = api.CommInitRank(&comm, nranks, id, rank);
  if (rc) return fail(h, TFK_ENCCL, "ncclCommInitRank failed: %s", api.GetErrorString ? api.GetErrorString(rc) : "?");
  h->comm = comm;
  h->own_comm = true;
  h->rank = rank;
  h->nranks = nranks;
  const char* mode = getenv("TFK_DP_MODE");  // "allreduce" keeps the replicated-Adam path
  h->sharded = nranks > 1 && (nranks & (nranks - 1)) == 0 && nranks <= 256 && !(mode && strcmp(mode, "allreduce") == 0);
  return setup_comm_streams(h);
}

int tfk_set_comm(tfk_handle* h, void* nccl_comm, int rank, int nranks) {
  if (!h) return TFK_EINVAL;
  if (nccl_comm && !nccl().ok) return fail(h, TFK_ENCCL, "libnccl.so.2 not available");
  h->comm = nccl_comm;
  h->own_comm = false;
  h->rank = rank;
  h->nranks = nccl_comm ? nranks : 1;
  const char* mode = getenv("TFK_DP_MODE");
  h->sharded = h->nranks > 1 && (h->nranks & (h->nranks - 1)) == 0 && h->nranks <= 256 && !(mode && strcmp(mode, "allreduce") == 0);
  return nccl_comm ? setup_comm_streams(h) : TFK_OK;
}

int tfk_ipc_export(tfk_handle* h, uint8_t* handle256_host) {
  if (!h || !handle256_host) return fail(h, TFK_EINVAL, "tfk_ipc_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  TFK_CUDA(h, cudaSetDevice(h->cfg.device));
  memset(handle256_host, 0, 256);
  void* bufs[4] = {h->G, h->Sh, h->Sl, h->dp_flags};  // gradient arena, bf16 operand arenas, publish flags
  for (int i = 0; i < 4; ++i) {
    if (!bufs[i]) continue;
    cudaIpcMemHandle_t hd;
    TFK_CUDA(h, cudaIpcGetMemHandle(&hd, bufs[i]));
    memcpy(handle256_host + 64 * i, &hd, 64);
  }
  return TFK_OK;
}

int tfk_ipc_import(tfk_handle* h, const uint8_t* handles_host, int nranks) {
  if (!h || !handles_host) return fail(h, TFK_EINVAL, "tfk_ipc_import: null argument");
  if (!h->sharded || nranks != h->nranks) return fail(h, TFK_EINVAL, "tfk_ipc_import: needs the sharded mode and %d handles", h->nranks);
  TFK_CUDA(h, cudaSetDevice(h->cfg.device));
  h->peer_G.assign(nranks, nullptr);
  h->peer_Sh.assign(nranks, nullptr);
  h->peer_Sl.assign(nranks, nullptr);
  std::vector<int*> peer_flags;
  for (int r = 0; r < nranks; ++r) {
    if (r == h->rank) {
      h->peer_G[r] = h->G; h->peer_Sh[r] = h->Sh; h->peer_Sl[r] = h->Sl;
      continue;
    }
    void* mapped[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < 4; ++i) {
      if (i == 2 && !h->x3) continue;
      cudaIpcMemHandle_t hd;
      memcpy(&hd, handles_host + 256 * r + 64 * i, 64);
      TFK_CUDA(h, cudaIpcOpenMemHandle(&mapped[i], hd, cudaIpcMemLazyEnablePeerAccess));
      h->ipc_opened.push_back(mapped[i]);
    }
    h->peer_G[r] = static_cast<float*>(mapped[0]);
    h->peer_Sh[r] = static_cast<__nv_bfloat16*>(mapped[1]);
    h->peer_Sl[r] = static_cast<__nv_bfloat16*>(mapped[2]);
    peer_flags.push_back(static_cast<int*>(mapped[3]));
  }
  if (!h->d_peer_flags) TFK_TRY(dev_alloc(h, &h->d_peer_flags, 256));
  TFK_CUDA(h, cudaMemcpy(h->d_peer_flags, peer_flags.data(), peer_flags.size() * sizeof(int*), cudaMemcpyHostToDevice));
  // a layer is eligible when its rows split evenly over the ranks in multiples of the 32-row store slab
  int eligible = 0;
  for (int l = 0; l <= h->L; ++l) {
    Layer& ly = h->layers[l];
    ly.peer_reduce = ly.Kpad % nranks == 0 && (ly.Kpad / nranks) % 32 == 0;
    eligible += ly.peer_reduce ? 1 : 0;
    if (ly.peer_reduce && !ly.peer_tm) {  // per layer, once: the maps do not depend on the frame count
      std::vector<void*> peers;
      for (int r = 0; r < nranks; ++r) peers.push_back(h->peer_G[r] + ly.off_w);
      char err[256] = {0};
      if (gemm_build_peer_maps(peers.data(), nranks, ly.K, ly.N, ly.ldn, &ly.peer_tm, err, sizeof(err)))
        return fail(h, TFK_ECUDA, "tfk_ipc_import layer %d: %s", l, err);
      h->allocs.push_back(ly.peer_tm);
    }
  }
  const char* mode = getenv("TFK_DP_MODE");
  const bool nccl_only = mode && strcmp(mode, "sharded_nccl") == 0;
  h->fused_rs = eligible > 0 && !nccl_only;
  h->fused_ag = nranks <= 16 && !nccl_only && !(mode && strcmp(mode, "fused_nccl_ag") == 0);
  cudaDeviceSynchronize();
  for (auto& kv : h->plans) free_plan(kv.second);  // plans built before the import lack the peer maps
  h->plans.clear();
  return TFK_OK;
}

int tfk_enable_timers(tfk_handle* h, int on) {
  if (!h) return TFK_EINVAL;
  TFK_TRY(drain_timers(h));
  h->timers_on = on != 0;
  if (on) {
    for (int i = 0; i < TFK_NUM_TIMERS; ++i) { h->timer_ms[i] = 0; h->timer_launches[i] = 0; }
  }
  return TFK_OK;
}

int tfk_get_timers(tfk_handle* h, double* ms_total, int64_t* launches) {
  if (!h) return TFK_EINVAL;
  TFK_CUDA(h, cudaSetDevice(h->cfg.device));
  TFK_TRY(drain_timers(h));
  for (int i = 0; i < TFK_NUM_TIMERS; ++i) {
    if (ms_total) ms_total[i] = h->timer_ms[i];
    if (launches) launches[i] = h->timer_launches[i];
  }
  return TFK_OK;
}

int64_t tfk_kernel_launches(const tfk_handle* h) { return h ? h->launches : 0; }

}  // extern "C"
