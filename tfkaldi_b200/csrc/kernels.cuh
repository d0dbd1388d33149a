// HBM-bound kernels of the tfkaldi hot path (everything that is not a GEMM).  sm_100a only.
// Each launcher returns cudaGetLastError() as int (0 = ok).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace tfk {

// fp32 [rows, cols] (pitch ld_src) -> bf16 hi (+ lo = bf16(x - hi) when lo != null), pitch ld_dst.
int k_split_f32(const float* src, int ld_src, __nv_bfloat16* hi, __nv_bfloat16* lo, int ld_dst,
                int rows, int cols, cudaStream_t st);

// Device-side feeder (reference: processing/feature_reader.py:91-156 apply_cmvn + splice, done on the host
// there): raw fp32 frames [R, D] of `num_utts` packed utterances (row offsets utt_off[num_utts+1]) ->
// CMVN with the utterance's (mean, 1/std) from cmvn[num_utts, 2, D] -> +-k spliced rows
// [x[t-k] .. x[t] .. x[t+k]] with ZEROS beyond the utterance edges -> bf16 hi(+lo) [rows, ld_dst] for rows
// [row_begin, row_begin + rows).  Writes layer 0's GEMM operand directly: the spliced fp32 matrix (11x the
// raw bytes for k = 5) never exists.
int k_splice_cmvn(const float* raw, const int32_t* utt_off, int num_utts, const float* cmvn, int D, int k,
                  int row_begin, int rows, __nv_bfloat16* hi, __nv_bfloat16* lo, int ld_dst, cudaStream_t st);

// bf16 hi (+lo) [rows, ld_src] -> fp32 [rows, cols] (pitch ld_dst): dst = hi + lo
int k_merge_bf16(const __nv_bfloat16* hi, const __nv_bfloat16* lo, int ld_src, float* dst, int ld_dst,
                 int rows, int cols, cudaStream_t st);

// activation-chain backward on fp32: out = dy * f'(.) * dropmask / keep reconstructed from the stored output y
// mode 0: relu (y > 0), 1: linear+dropout (y != 0), 2: sigmoid, 3: tanh; +4 when dropout is in the chain
int k_mask_scale_f32(const float* dy, const float* y, float* out, size_t n, float scale, int mode,
                     cudaStream_t st);
// dst[0] = (float)src[0]
int k_double_to_float(const double* src, float* dst, cudaStream_t st);

// Softmax cross-entropy forward+backward in one pass over the logits
// (reference: neuralNetworks/trainer.py:526-531: one_hot + softmax_cross_entropy_with_logits + reduce_sum).
//   row_loss[r] = logsumexp(z_r) - z_r[label_r]        (0 if the label is outside [0,O): empty one-hot row; the
//                                                       gradient of such a row is softmax - 0, as TensorFlow's op gives)
//   d = softmax(z_r) - onehot(label_r)  -> bf16 hi (+lo), pad columns [O, ld) zeroed.  d_hi may be null.
int k_softmax_ce(const float* logits, int ld, const int32_t* labels, int B, int O, float* row_loss,
                 __nv_bfloat16* d_hi, __nv_bfloat16* d_lo, cudaStream_t st);

// acc[0] += sum(row_loss[0..B)) (fixed-order, double);  acc[1] += B
int k_accum_loss(const float* row_loss, int B, double* acc, cudaStream_t st);

// out[n] += sum over rows of (hi + lo)[row, n]; deterministic, one launch.
// ws: zero-initialised scratch of >= 1024 + 64*ld floats (self-resetting counters, then partials).
int k_colsum_bf16(const __nv_bfloat16* hi, const __nv_bfloat16* lo, int ld, int rows, int cols,
                  float* ws, float* out, cudaStream_t st);

// outs[j][c] += sum_g parts[j][g*ld + c], g < groups: finishes the per-32-row column sums written by the
// GEMM epilogue (GemmSpec::colsum_part) for njobs layers in one launch.  parts/outs are HOST arrays.
int k_colsum_finalize(const float* const* parts, float* const* outs, int njobs, int groups, int ld, int cols,
                      cudaStream_t st);

// Mean -> clip[-1,1] -> TF-form Adam -> refresh bf16 shadows -> re-zero the accumulator
// (reference: neuralNetworks/trainer.py:174-184 and :350).  frames = acc[1] (device double).
// Works on `nseg` segments {offset, count} (floats, multiples of 4) of the flat arenas: the whole arena on
// one GPU, a rank's slice of every layer under sharded data parallelism.
int k_adam(float* w, float* g, float* m, float* v, __nv_bfloat16* w_hi, __nv_bfloat16* w_lo, const size_t* seg_off,
           const size_t* seg_cnt, int nseg, const double* acc, float lr_t, float beta1, float beta2, float eps,
           cudaStream_t st, int n_bcast = 0, int n_peers = 0, __nv_bfloat16* const* peer_hi = nullptr,
           __nv_bfloat16* const* peer_lo = nullptr);
// Fused update -> all-gather (data parallel, peer memory): with n_bcast > 0 the refreshed bf16 copies of the
// first n_bcast segments are ALSO stored into the n_peers other GPUs' shadow arenas (same offsets) over NVLink.
// k_dp_publish then writes `value` into slot `me` of every peer's flag array (d_peer_flags: DEVICE array of
// peer-mapped pointers); k_dp_wait spins until every other rank's slot in the local array reached `value`.
// The flag array holds [slot][16 ranks] words (TFK_DP_FLAG_WORDS ints): slot 0 paces whole steps, slot 1 + l layers.
constexpr int TFK_DP_FLAG_SLOTS = 66;
constexpr int TFK_DP_FLAG_WORDS = TFK_DP_FLAG_SLOTS * 16;
// Exchange area behind the flag words (same allocation, so the peers have it mapped): [16 ranks][stride floats].
// k_dp_small_push stores this rank's small_n floats (a multiple of 4) + {loss, frames} into slot `me` on every GPU;
// after a publish / wait pair k_dp_small_reduce adds the n slots in rank order into g_small / acc.
int k_dp_small_push(int* const* d_peer_flags, int n_peers, int* own_flags, int me, const float* g_small, const double* acc,
                    int small_n, int stride, cudaStream_t st);
int k_dp_small_reduce(const int* own_flags, int n_ranks, float* g_small, double* acc, int small_n, int stride, cudaStream_t st);
int k_dp_publish(int* const* d_peer_flags, int n_peers, int slot, int me, int value, cudaStream_t st);
int k_dp_wait(const int* flags, int n_ranks, int slot, int me, int value, cudaStream_t st);

// Batch-norm (reference: classifiers/activation.py:159 -> tf.contrib.layers.batch_norm defaults: decay 0.999, eps 1e-3,
// center without scale, biased variance in the moving average).
// backward: sums[0..N) = sum_B dy, sums[ld..ld+N) = sum_B dy*xhat; g_beta += sum_B dy.  ws >= 256*ld floats
int k_bn_bwd_reduce(const __nv_bfloat16* dy_hi, const __nv_bfloat16* dy_lo, const __nv_bfloat16* z_hi,
                    const __nv_bfloat16* z_lo, int ld, int B, int N, const float* mean,
                    const float* rstd, float* ws, unsigned int* counters, float* sums, float* g_beta, cudaStream_t st);
// Column-strip batch-norm passes (one launch per layer and direction): statistics from the GEMM epilogue's per-32-row
// partials (part_* [groups, pld]; training: batch statistics + moving-average update, else the moving statistics),
// then y = f((z - mean) * rstd + beta) [* dropout] over the whole [B, N] matrix; backward: dbeta += sum dy and
// dz = rstd * (dy - mean(dy) - xhat * mean(dy * xhat)) in place.  mean / rstd [N] are written by the forward for the backward.
int k_bn_fwd_strip(const float* part_sum, const float* part_sq, int groups, int pld, const __nv_bfloat16* z_hi,
                   const __nv_bfloat16* z_lo, int ld, int B, int N, float eps, float decay, int training, float* mean, float* rstd,
                   float* moving_mean, float* moving_var, const float* beta, int act, float keep, unsigned long long seed,
                   __nv_bfloat16* y_hi, __nv_bfloat16* y_lo, cudaStream_t st);
int k_bn_bwd_strip(const float* part_sum, const float* part_dot, int groups, int pld, __nv_bfloat16* dy_hi, __nv_bfloat16* dy_lo,
                   const __nv_bfloat16* z_hi, const __nv_bfloat16* z_lo, int ld, int B, int N, const float* mean, const float* rstd,
                   float* g_beta, cudaStream_t st);
// dz = rstd * (dy - mean_B(dy) - xhat * mean_B(dy*xhat)), written in place over dy, from k_bn_bwd_reduce's sums
int k_bn_bwd_apply(__nv_bfloat16* dy_hi, __nv_bfloat16* dy_lo, const __nv_bfloat16* z_hi,
                   const __nv_bfloat16* z_lo, int ld, int B, int N, const float* mean,
                   const float* rstd, const float* sums, cudaStream_t st);

// L2Norm (reference: classifiers/activation.py:87-111, quirk kept): s = mean over columns of u^2 (per frame);
// y = u / s if s > 1 else u  (divides by the mean SQUARE, not its root), then dropout.  u: output of the
// nonlinearity, bf16 hi(+lo) [B, ld]; y likewise; s_out fp32 [B].
int k_l2norm_fwd(const __nv_bfloat16* u_hi, const __nv_bfloat16* u_lo, int ld, int B, int N, float keep,
                 unsigned long long seed, __nv_bfloat16* y_hi, __nv_bfloat16* y_lo, float* s_out, cudaStream_t st);
// backward, in place over d (= d loss / d y, dropout already undone): for s > 1
//   du_j = d_j / s - u_j * (2 / (N s^2)) * sum_k d_k u_k ;  else du = d ;  then the nonlinearity's slope taken
// from u (act 1 relu: u > 0, 2 sigmoid: u(1-u), 3 tanh: 1-u^2, 0 none).
int k_l2norm_bwd(__nv_bfloat16* d_hi, __nv_bfloat16* d_lo, const __nv_bfloat16* u_hi, const __nv_bfloat16* u_lo,
                 const float* s_in, int ld, int B, int N, int act, cudaStream_t st);

// Decoder output (reference: neuralNetworks/decoder.py:44 softmax; nnet.py:280-286 log(P/prior)):
//   prior == null : out = softmax(z)          (Decoder.__call__)
//   prior != null : out = log(softmax(z)/prior)   (Nnet.decode pseudo log-likelihood, no flooring);
//                   `prior` must hold LOG(prior) (k_log_vector), so the kernel needs no per-element log/div
// out is dense [T, O] fp32.
int k_log_vector(const float* x, float* y, int n, cudaStream_t st);
int k_decode_out(const float* logits, int ld, int T, int O, const float* prior, float* out,
                 cudaStream_t st);

}  // namespace tfk
