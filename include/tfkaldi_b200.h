/* tfkaldi_b200 — C-ABI of the B200-native DNN / CrossEnthropyTrainer / Decoder engine.
 *
 * This is the drop-in boundary for the ONE hot path of vrenkens/tfkaldi: the place where the
 * reference crosses from Python into the TensorFlow runtime (`tf.Session.run` / `op.run()` /
 * `tensor.eval()`).  The reference has no FFI of its own; every entry point below names the
 * reference call it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - every function returns 0 (TFK_OK) or a negative TFK_E* code; tfk_last_error() gives the text.
 *   - `x`, `labels`, `out`, `log_prior` are DEVICE pointers (caller-owned, e.g. torch tensors);
 *     names ending in `_host` are host pointers.  tfk_set_tensor/tfk_get_tensor accept either
 *     (cudaMemcpyDefault).
 *   - all work is enqueued on the caller's `stream` (a cudaStream_t passed as void*); calls that
 *     return a host value synchronise that stream.
 *   - a handle is bound to one GPU and is not re-entrant; one process (rank) per GPU.
 *   - no C++ exceptions cross the boundary; plain pointers and sizes only.
 */
#ifndef TFKALDI_B200_H_
#define TFKALDI_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TFK_OK 0
#define TFK_EINVAL (-1) /* bad argument / configuration */
#define TFK_ECUDA (-2)  /* CUDA runtime / driver error */
#define TFK_ENCCL (-3)  /* NCCL error or NCCL unavailable */
#define TFK_ESHAPE (-4) /* shape outside the configured workspace */

#define TFK_ABI_VERSION 1

/* numeric modes */
#define TFK_PREC_BF16 0   /* bf16 operands, fp32 accumulate, fp32 master weights ("perf") */
#define TFK_PREC_BF16X3 1 /* every operand carried as bf16 hi+lo, 3 MMAs per product: fp32-equivalent ("parity") */

/* hidden-layer non-linearity (reference: neuralNetworks/nnet.py:47-65) */
#define TFK_NONLIN_RELU 0
#define TFK_NONLIN_LINEAR 1
#define TFK_NONLIN_SIGMOID 2
#define TFK_NONLIN_TANH 3

/* tensor kinds for tfk_set_tensor / tfk_get_tensor.  `layer` = 0..L-1 hidden, L = output layer.
 * Names in comments are the reference's TF variable names. */
#define TFK_T_WEIGHTS 0     /* Classifier/layer{l}/parameters/weights  [in, out] row-major fp32 (classifiers/layer.py:42-44) */
#define TFK_T_BIASES 1      /* Classifier/layer{l}/parameters/biases   [out]                     (classifiers/layer.py:46-48) */
#define TFK_T_BN_BETA 2     /* .../activation/batch_norm/beta            [out]  (hidden layers, batch_norm only) */
#define TFK_T_BN_MOVING_MEAN 3
#define TFK_T_BN_MOVING_VAR 4
#define TFK_T_ADAM_M_W 5    /* Adam slots (the reference never checkpoints them; exposed as a superset) */
#define TFK_T_ADAM_V_W 6
#define TFK_T_ADAM_M_B 7
#define TFK_T_ADAM_V_B 8
#define TFK_T_ADAM_M_BETA 9
#define TFK_T_ADAM_V_BETA 10
#define TFK_T_GRAD_W 11     /* gradient accumulators `gradients/...` (trainer.py:118-122), read-mostly */
#define TFK_T_GRAD_B 12
#define TFK_T_GRAD_BETA 13

/* scalar kinds for tfk_set_scalar / tfk_get_scalar */
#define TFK_S_GLOBAL_STEP 0   /* train_variables/global_step        (trainer.py:98-100) */
#define TFK_S_LR_FACT 1       /* train_variables/learning_rate_fact (trainer.py:104-106) */
#define TFK_S_ACTIVE_LAYERS 2 /* Classifier/initialisedlayers + 1   (classifiers/dnn.py:85-89) */
#define TFK_S_LOSS_SUM 3      /* batch_loss accumulator             (trainer.py:91-93), read-only */
#define TFK_S_NUM_FRAMES 4    /* train/num_frames accumulator       (trainer.py:126-128), read-only */
#define TFK_S_ADAM_STEP 5     /* Adam's own step count t (TF: beta1_power/beta2_power inside AdamOptimizer, trainer.py:115).
                                 Advanced by every tfk_apply / tfk_train_step, NOT touched by setting TFK_S_GLOBAL_STEP:
                                 the reference's `train_variables` saver (trainer.py:204-205) does not hold the beta powers,
                                 so a restored / rolled-back trainer keeps counting and a freshly initialised one restarts at 0 */

typedef struct tfk_handle tfk_handle;

typedef struct tfk_config {
  int32_t abi_version;  /* TFK_ABI_VERSION */
  int32_t num_layers;   /* L hidden layers (DNN num_layers, classifiers/dnn.py:13) */
  int32_t input_dim;    /* spliced feature dimension (nnet.py:40) */
  int32_t hidden_dim;   /* DNN num_units */
  int32_t output_dim;   /* number of pdf-ids */
  int32_t max_frames;   /* workspace size: largest micro-batch (frames) of accumulate/eval calls */
  int32_t nonlin;       /* TFK_NONLIN_* */
  int32_t batch_norm;   /* 0/1 (nnet.py:42) */
  float keep_prob;      /* conf['dropout'] is a KEEP probability; >= 1 disables (nnet.py:70-72) */
  float bn_eps;         /* 1e-3  (tf.contrib.layers.batch_norm default) */
  float bn_decay;       /* 0.999 */
  float adam_beta1;     /* 0.9   (tf.train.AdamOptimizer defaults, trainer.py:115) */
  float adam_beta2;     /* 0.999 */
  float adam_eps;       /* 1e-8 */
  int32_t precision;    /* TFK_PREC_* */
  int32_t device;       /* CUDA device ordinal */
  uint64_t seed;        /* dropout Philox key base */
  int32_t l2_norm;      /* 0/1: L2Norm after the nonlinearity (nnet.py:67-68, activation.py:87-111) */
  int32_t reserved;
} tfk_config;

/* Fill `cfg` with the reference's defaults (ReLU, no BN, keep 1, Adam/BN TF defaults, bf16). */
void tfk_default_config(tfk_config* cfg);

/* Replaces graph construction + `init_op.run()` (trainer.py:37-215, 244-247; decoder.py:20-47).
 * All parameters start at ZERO; the caller injects the initial weights (classifiers/layer.py:39-48)
 * with tfk_set_tensor so the CPU oracle and the GPU start from identical values. */
int tfk_create(const tfk_config* cfg, tfk_handle** out);
int tfk_destroy(tfk_handle* h);
const char* tfk_last_error(const tfk_handle* h); /* h may be NULL: last error of tfk_create */

/* Replaces tf.train.Saver save/restore of single variables (dnn.py:129, trainer.py:448-486,
 * decoder.py:73-81).  `count` = number of fp32 elements (in*out for weights, out for vectors). */
int tfk_set_tensor(tfk_handle* h, int kind, int layer, const float* src, size_t count, void* stream);
int tfk_get_tensor(tfk_handle* h, int kind, int layer, float* dst, size_t count, void* stream);
int tfk_set_scalar(tfk_handle* h, int kind, double value);
int tfk_get_scalar(tfk_handle* h, int kind, double* value_host, void* stream);

/* One fused FFLayer forward on engine-owned activations: unit-test / roofline entry.
 * x: fp32 [B, in_dim(layer)] device, y: fp32 [B, out_dim(layer)] device (post activation chain;
 * raw logits for the output layer).  Replaces FFLayer.__call__ (classifiers/layer.py:24-58). */
int tfk_fflayer_fwd(tfk_handle* h, int layer, const float* x, float* y, int B, int training, void* stream);
/* Backward of the same layer: dy fp32 [B, out_dim] is d(loss)/d(layer output); accumulates dW, db
 * (and dbeta) into the gradient accumulators and writes dx fp32 [B, in_dim] (may be NULL).
 * Must follow a tfk_fflayer_fwd(training=1) of the same layer and B.  Replaces the layer's slice of
 * tf.gradients (trainer.py:155). */
int tfk_fflayer_bwd(tfk_handle* h, int layer, const float* dy, float* dx, int B, void* stream);

/* Softmax cross-entropy (trainer.py:526-531): loss_sum (device float[1]) = sum of per-frame CE,
 * dlogits (device fp32 [B,O], may be NULL) = softmax - onehot. */
int tfk_softmax_ce(tfk_handle* h, const float* logits, const int32_t* labels, int B, float* loss_sum,
                   float* dlogits, void* stream);

/* == `update_gradients_op.run(feed_dict)` (trainer.py:165-169, 328): forward in training mode,
 * sum-CE, backward; grads += g, batch_loss += loss, num_frames += B, BN moving averages updated.
 * x: fp32 [B, input_dim] packed frames (utterance-major, no padding), labels: int32 [B]. */
int tfk_accumulate(tfk_handle* h, const float* x, const int32_t* labels, int B, void* stream);

/* Device-side feeder: the same as tfk_accumulate / tfk_forward_loglik, but from RAW
 * (un-normalised, un-spliced) frames; CMVN and the +-context splice of FeatureReader.get_utt
 * (processing/feature_reader.py:42-60, 91-156) run on the device, fused with the operand conversion.
 *   raw          fp32 [R, feat_dim], the utterances packed one after the other
 *   utt_offsets  int32 [num_utts + 1] first row of every utterance (utt_offsets[num_utts] == R)
 *   cmvn         fp32 [num_utts, 2, feat_dim]: per utterance the speaker's mean and 1/sqrt(variance)
 *   labels       int32 [R];   feat_dim * (2*context + 1) must equal input_dim
 * Rows beyond an utterance's edges are zero, exactly as the reference's splice.  (Utterances shorter than
 * 2*context+1 frames are dropped by the reference's dispenser; the caller filters them the same way.) */
int tfk_accumulate_raw(tfk_handle* h, const float* raw, const int32_t* utt_offsets, int num_utts, const float* cmvn,
                       const int32_t* labels, int R, int feat_dim, int context, void* stream);
int tfk_forward_loglik_raw(tfk_handle* h, const float* raw, const int32_t* utt_offsets, int num_utts, const float* cmvn,
                           int R, int feat_dim, int context, const float* prior, float* out, void* stream);
/* The same for rows [row_begin, row_begin + rows) of the packed utterances only (out: fp32 [rows, O]); the splice
 * still reads its context across the range borders from `raw`.  Lets Nnet.decode (nnet.py:267-289) pipeline a long
 * utterance tile by tile: device->host copy and archive write of one tile while the next one is computed. */
int tfk_forward_loglik_raw_rows(tfk_handle* h, const float* raw, const int32_t* utt_offsets, int num_utts, const float* cmvn,
                                int row_begin, int rows, int feat_dim, int context, const float* prior, float* out,
                                void* stream);

/* == `run([average_loss, apply_gradients_op])` + init_grads/init_loss/init_num_frames
 * (trainer.py:174-184, 337-352): [allreduce over ranks] -> grads / num_frames -> clip[-1,1] -> Adam
 * (TF form) -> global_step += 1 -> accumulators re-zeroed.  `lr` is the decayed learning rate
 * lr0 * decay^(global_step/num_steps) (trainer.py:110-112); the engine multiplies by
 * learning_rate_fact.  mean_loss_host (may be NULL => no host sync) receives batch_loss/num_frames
 * evaluated with the pre-update weights. */
int tfk_apply(tfk_handle* h, float lr, float* mean_loss_host, void* stream);
/* The mean loss of the most recent tfk_apply / tfk_train_step / tfk_train_step_raw that was called with
 * mean_loss_host == NULL: waits for that step and returns batch_loss/num_frames exactly as the step itself would have.
 * It lets the caller put host work (queueing the next batch's copy, trainer.py's logging) between launching a step
 * and blocking on its loss.  TFK_EINVAL when no such step is outstanding. */
int tfk_last_loss(tfk_handle* h, float* mean_loss_host, void* stream);

/* == tfk_accumulate + tfk_apply for ONE micro-batch (the reference's update() with
 * numutterances_per_minibatch covering the whole batch, trainer.py:310-352), same arithmetic; on a single
 * GPU each layer's clip+Adam update runs on a side stream right after that layer's backward kernel so it
 * overlaps the remaining backward pass.  Falls back to the plain sequence under data parallelism or when
 * gradients are already being accumulated. */
int tfk_train_step(tfk_handle* h, const float* x, const int32_t* labels, int B, float lr, float* mean_loss_host,
                   void* stream);
/* tfk_train_step from RAW frames (== tfk_accumulate_raw + tfk_apply; arguments as tfk_accumulate_raw): what the
 * prefetching feeder (processing/feeder.py) calls once per batch, so that FeatureReader.get_utt's CMVN + splice
 * (feature_reader.py:42-60) and Trainer.update's padding (trainer.py:276-307) never run on the host. */
int tfk_train_step_raw(tfk_handle* h, const float* raw, const int32_t* utt_offsets, int num_utts, const float* cmvn,
                       const int32_t* labels, int R, int feat_dim, int context, float lr, float* mean_loss_host,
                       void* stream);

/* == `update_valid_loss.run(feed_dict)` (trainer.py:186-195, 428): eval-mode forward (moving-stat BN,
 * no dropout) + CE; batch_loss += loss, num_frames += B. */
int tfk_eval_accumulate(tfk_handle* h, const float* x, const int32_t* labels, int B, void* stream);
/* == `average_loss.eval()` + init_loss/init_num_frames (trainer.py:435-439). Synchronises. */
int tfk_eval_finish(tfk_handle* h, float* mean_loss_host, void* stream);

/* == Decoder.__call__ (decoder.py:49-71): eval-mode forward + softmax -> out fp32 dense [T, O]. */
int tfk_forward_posteriors(tfk_handle* h, const float* x, int T, float* out, void* stream);
/* == Decoder.__call__ + Nnet.decode post-processing (nnet.py:277-286): out = log(softmax / prior),
 * no flooring (the reference discards its np.where).  prior: device fp32 [O]. */
int tfk_forward_loglik(tfk_handle* h, const float* x, int T, const float* prior, float* out, void* stream);

/* == `halve_learningrate_op.run()` (trainer.py:141-142, 443-446) */
int tfk_halve_lr(tfk_handle* h);
/* == control_ops['add'] (classifiers/dnn.py:92): use the first n hidden layers, 1 <= n <= L */
int tfk_set_active_layers(tfk_handle* h, int n);
/* Philox key for the NEXT tfk_accumulate's dropout masks (layer l uses seed + l); afterwards the
 * engine advances it by num_layers + 1 per call.  Under data parallelism a rank is one micro-batch of the step
 * (trainer.py:310-332): rank r uses seed + r (num_layers + 1) + l and a call advances it by nranks (num_layers + 1),
 * the sequence one GPU accumulating the ranks' shards as micro-batches would draw. */
int tfk_set_dropout_seed(tfk_handle* h, uint64_t seed);
/* Diagnostic read-back (tests): the stored output of hidden layer `layer` from the LAST forward pass — what the
 * backward pass derives the ReLU / dropout gradient mask from (activation.py:84, 140-141) — as fp32 [B, hidden_dim]
 * into the DEVICE buffer dst.  Lets a checker replay the backward pass with exactly the engine's activation pattern. */
int tfk_get_activation(tfk_handle* h, int layer, float* dst, int B, void* stream);

/* Data parallel (no reference equivalent: a rank plays the role of one utterance micro-batch of
 * trainer.py:310-332).  tfk_comm_unique_id: rank 0 fills 128 bytes; every rank then calls
 * tfk_comm_init with the same id.  After that tfk_apply sums {gradients, batch_loss, num_frames}
 * over ranks before the (mathematically identical) mean -> clip -> Adam step:
 *   - default for 2^k ranks ("sharded"): per-layer NCCL reduce-scatter of the weight gradients, each
 *     rank runs Adam on its 1/n slice of every layer, then all-gathers the bf16 operand copies; the
 *     small vectors (biases, betas) and {loss, frames} are all-reduced and updated everywhere;
 *   - TFK_DP_MODE=allreduce (or a non-power-of-two world): all-reduce everything, replicated Adam.
 * Under the sharded mode the fp32 master weights / Adam slots of a slice live on its owner only:
 * tfk_get_tensor(TFK_T_WEIGHTS / TFK_T_ADAM_*_W) is then COLLECTIVE (it all-gathers them first), so
 * every rank must call it, in the same order (checkpointing code that all ranks run does). */
int tfk_comm_unique_id(uint8_t* id128_host);
int tfk_comm_init(tfk_handle* h, const uint8_t* id128_host, int rank, int nranks);
/* Collectives fused into the compute kernels over NVLink peer memory (single node, after tfk_comm_init in
 * the sharded mode): every rank exports 256 bytes (four CUDA IPC handles: gradient arena, bf16 operand arenas
 * hi / lo, publish flags); the 256*nranks bytes of all ranks (rank order) are imported by every rank.  Then
 *   - GEMM -> reduce-scatter: the wgrad epilogue TMA-reduce-adds each output slab directly into the slice
 *     owner's accumulator (layers whose rows split evenly over the ranks in multiples of 32; others keep NCCL);
 *   - update -> all-gather: the sharded Adam kernel stores the refreshed bf16 operand slices into every peer's
 *     arena as it computes them, followed by a flag publish/wait between the GPUs (no NCCL all-gather).
 * TFK_DP_MODE=sharded_nccl disables both, fused_nccl_ag only the second. */
int tfk_ipc_export(tfk_handle* h, uint8_t* handle256_host);
int tfk_ipc_import(tfk_handle* h, const uint8_t* handles_host, int nranks);
/* Alternative: adopt an existing ncclComm_t (not destroyed by tfk_destroy). */
int tfk_set_comm(tfk_handle* h, void* nccl_comm, int rank, int nranks);

/* Per-kernel CUDA-event timing (measurement only): categories in TFK_TIMER_*; enabling inserts event
 * pairs around every kernel launched by the calls above.  tfk_get_timers synchronises. */
#define TFK_TIMER_GEMM_FWD 0
#define TFK_TIMER_GEMM_BWD 1
#define TFK_TIMER_SOFTMAX_CE 2
#define TFK_TIMER_ADAM 3
#define TFK_TIMER_COLSUM 4
#define TFK_TIMER_BN 5
#define TFK_TIMER_CONVERT 6
#define TFK_TIMER_DECODE_OUT 7
#define TFK_TIMER_ALLREDUCE 8
#define TFK_NUM_TIMERS 9
int tfk_enable_timers(tfk_handle* h, int on);
int tfk_get_timers(tfk_handle* h, double* ms_total /*[TFK_NUM_TIMERS]*/, int64_t* launches /*[TFK_NUM_TIMERS]*/);
/* number of kernels launched by this handle so far (bench.py's gpu_launches) */
int64_t tfk_kernel_launches(const tfk_handle* h);

/* Library / device probes (safe without a GPU). */
int tfk_abi_version(void);
int tfk_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* TFKALDI_B200_H_ */
