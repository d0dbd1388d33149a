// Persistent grouped GEMM for sm_100a: tcgen05.mma (bf16 x bf16 -> fp32 in TMEM), operands staged by
// TMA into 128B-swizzled shared memory, fused epilogue, TMA store / TMA reduce-add.
//
// One launch runs up to two "problems" D = A.B^T (UMMA convention: A is [M,K], B is [N,K], both may
// be K-major or MN-major in memory).  The three uses on the tfkaldi hot path
// (reference: neuralNetworks/classifiers/layer.py:52  `tf.matmul(inputs, weights) + biases`,
//  and its tf.gradients twin, neuralNetworks/trainer.py:155):
//   forward  Y[B,N]   = X[B,K] . W[K,N]      A = X  (K-major)   B = W  (MN-major)
//   dgrad    dX[B,K]  = dZ[B,N] . W[K,N]^T   A = dZ (K-major)   B = W  (K-major, W row = output col)
//   wgrad    dW[K,N] += X[B,K]^T . dZ[B,N]   A = X  (MN-major)  B = dZ (MN-major), TMA reduce-add
#pragma once
#include <vector>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace tfk {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int STAGES = 4;
constexpr int GEMM_THREADS = 320;  // warp0 TMA, warp1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)
constexpr int TMEM_COLS = 512;     // 2 accumulator stages x 256 fp32 columns

enum OutKind : int {
  OUT_BF16 = 0,        // D_hi = bf16(v)
  OUT_BF16_SPLIT = 1,  // D_hi = bf16(v), D_lo = bf16(v - D_hi)   (fp32-equivalent storage)
  OUT_F32 = 2,         // D_hi(f32) = v
  OUT_F32_REDADD = 3,  // D_hi(f32) += v   (TMA reduce-add; split-K and micro-batch accumulation)
};

// Host-side description of one GEMM.  All matrices are row-major with a leading dimension in
// elements; bf16 leading dimensions must be multiples of 8, fp32 ones multiples of 4 (16-byte TMA pitch).
struct GemmSpec {
  int M = 0, N = 0, K = 0;
  const __nv_bfloat16* A_hi = nullptr;
  const __nv_bfloat16* A_lo = nullptr;
  int lda = 0;
  int a_mn = 0;  // 0: A stored [M,K];  1: A stored [K,M]
  const __nv_bfloat16* B_hi = nullptr;
  const __nv_bfloat16* B_lo = nullptr;
  int ldb = 0;
  int b_mn = 0;    // 0: B stored [N,K];  1: B stored [K,N]
  int nsplit = 1;  // 1: plain bf16;  3: bf16x3 (Ah.Bh + Ah.Bl + Al.Bh), ~fp32 accuracy
  int ksplit = 1;  // split the reduction over this many work items (OUT_F32_REDADD only: partials add up)
  int out_kind = OUT_BF16;
  void* D_hi = nullptr;
  void* D_lo = nullptr;
  int ldd = 0;
  const float* bias = nullptr;  // [>= roundup(N,256)] added per column, or null
  int act = 0;  // element-wise nonlinearity after bias: 0 none, 1 relu, 2 sigmoid, 3 tanh (nnet.py:47-62)
  const __nv_bfloat16* mask_src = nullptr;  // [M, mask_ld]: v = pass(mask_src) ? v*scale : 0
  const __nv_bfloat16* mask_src_lo = nullptr;  // optional low half (bf16x3 storage), used by deriv != 0
  int mask_ld = 0;
  // deriv != 0: mask_src holds the stored activation a = f(z) * dropmask / keep and the epilogue multiplies by
  // f'(z) * dropmask / keep reconstructed from it: 1 sigmoid y(1-y), 2 tanh 1-y^2, y = a / scale... (scale = 1/keep)
  int deriv = 0;
  int dropout_in_chain = 0;  // deriv == 2 only: a == 0 means "dropped" (otherwise tanh(0) = 0 keeps slope 1)
  int mask_nonzero = 0;  // 0: pass where mask_src > 0 (relu);  1: pass where mask_src != 0 (linear+dropout)
  float scale = 1.0f;
  // Compact alternative to mask_src: 1 bit per element, word (col/32, row) at bits[(col/32)*mask_bits_ld + row]
  // (chunk-major so a warp's 32 rows read/write 128 contiguous bytes).  The forward epilogue WRITES it
  // (bit = stored activation passes gradient: > 0, or != 0 with mask_nonzero), the dgrad epilogue READS it.
  uint32_t* mask_bits_out = nullptr;
  const uint32_t* mask_bits_in = nullptr;
  int mask_bits_ld = 0;
  // forward dropout (reference: classifiers/activation.py:140-141): keep<1 => v = v/keep * floor(keep+u)
  float keep = 1.0f;
  unsigned long long seed = 0;  // Philox key; counter = (col / 8, row)
  // backward through a chain whose stored output does not prove a drop (identity or tanh before the dropout: a kept
  // unit may hold an exact 0): < 1 makes the epilogue re-draw the forward pass's keep decisions from `seed` (the key of
  // the layer below's forward dropout) instead of testing the stored value
  float bwd_drop_keep = 1.0f;
  // per-(128-row tile, column) partial sums of v and v*v taken BEFORE relu/dropout (batch-norm stats)
  float* stat_sum = nullptr;  // [tiles_m, stat_ld]
  float* stat_sq = nullptr;
  int stat_ld = 0;
  // per-(32-row group, column) sums of the FINAL stored value (after mask / relu / dropout): the bias
  // gradient of the layer below falls out of the dgrad epilogue.  [ceil(M/128)*4, colsum_ld]
  float* colsum_part = nullptr;
  int colsum_ld = 0;
  // dgrad into a batch-norm layer (reference: tf.contrib.layers.batch_norm's gradient, classifiers/activation.py:159):
  // with colsum_part (= sum_B dY, also the beta gradient) the epilogue emits, in the same layout, the per-32-row
  // sums of dY * xhat, xhat = (z - mean) * rstd read from the layer's stored linear output z [M, bn_z_ld] — the two
  // column reductions the batch-norm backward needs, so no separate pass over dY and z is made for them.
  const __nv_bfloat16* bn_z_hi = nullptr;
  const __nv_bfloat16* bn_z_lo = nullptr;
  int bn_z_ld = 0;
  const float* bn_mean = nullptr;  // [>= roundup(N, 256)]
  const float* bn_rstd = nullptr;
  float* colsum2_part = nullptr;   // [ceil(M/128)*4, colsum_ld]
  // relu / linear chains: with bn_beta ([>= roundup(N,256)]) and mask_src (the stored y = f(xhat + beta) * keepmask / keep)
  // xhat is recovered as y * keep - beta on every element where dY != 0, and z is not read at all
  const float* bn_beta = nullptr;
  // OUT_F32_REDADD only: fused reduce-scatter.  Output rows [o*rows_per_owner, (o+1)*rows_per_owner) are
  // reduce-added through peer_tm[o] (tensor map of the same [M, ldd] matrix on GPU o, mapped over NVLink; the
  // local matrix for this rank) instead of D_hi.  rows_per_owner must be a multiple of 32.
  const CUtensorMap* peer_tm = nullptr;  // DEVICE array [num_peers] from gemm_build_peer_maps (depends on the matrix only)
  int num_peers = 0;
  int rows_per_owner = 0;
};

struct alignas(64) GemmProblem {
  CUtensorMap tmA[2];
  CUtensorMap tmB[2];
  CUtensorMap tmD[2];
  int M, N, K;
  int a_mn, b_mn, nsplit, out_kind;
  int epi_variant;  // which epilogue instantiation serves this problem (gemm.cu: EpiVariant)
  int act, mask_ld, mask_nonzero, deriv;
  float scale, keep_inv;
  unsigned int drop_thr;  // keep element iff its 16-bit Philox field >= drop_thr; 0 => no dropout
  unsigned int bwd_drop_thr;  // backward: re-draw the keep bits with this threshold (0 => derive the mask from stored values)
  const float* bias;
  const __nv_bfloat16* mask_src;
  const __nv_bfloat16* mask_src_lo;
  int dropout_in_chain;
  uint32_t* mask_bits_out;
  const uint32_t* mask_bits_in;
  int mask_bits_ld;
  unsigned long long seed;
  float* stat_sum;
  float* stat_sq;
  int stat_ld;
  float* colsum_part;
  int colsum_ld;
  const __nv_bfloat16* bn_z_hi;
  const __nv_bfloat16* bn_z_lo;
  int bn_z_ld;
  const float* bn_mean;
  const float* bn_rstd;
  float* colsum2_part;
  const float* bn_beta;
  int bn_from_y;
  const CUtensorMap* peer_tm;  // device array [num_peers] (gemm_build_peer_maps; owned by the caller)
  int num_peers, rows_per_owner;
  int tiles_m, tiles_n, tile_begin, num_kb;
  int ksplit, kb_per_split;
};

struct alignas(64) GemmParams {
  GemmProblem p[2];
  int nprob;
  int total_tiles;
  int* sched;  // [2]: {next tile counter, finished-CTA counter}; self-resetting (1-CTA kernel)
  // CTA-pair kernel (cta_group::2, 256x256 tiles): static longest-first work lists, one per pair
  int two_cta;
  int bias_shfl;   // forward tiles: bias loaded once per tile before the accumulator wait and broadcast by shuffles
  int a_resident;  // CTA-pair kernel: A-stationary launch (one short-K problem; consecutive column tiles reuse the A panel)
  int num_pairs;
  int list_stride;
  const int* tile_list;  // device [num_pairs, list_stride], -1 terminated
};

// Build the kernel parameters (tensor maps) for up to two problems.  Long-K problems should come first.
// Returns 0 on success, negative on error (message in err).
int gemm_build_params(const GemmSpec* specs, int nspec, int* sched, GemmParams* out, char* err,
                      int errlen, int two_cta = 0);
// CTA-pair plans only: compute the per-pair longest-processing-time-first work lists and upload them
// (cudaMalloc; the caller owns *d_list and frees it with cudaFree).
int gemm_upload_tile_lists(GemmParams* params, int num_sms, int** d_list, char* err, int errlen);
// the host-only half of it: flat [pairs, stride] lists, -1 padded (deterministic; cached per launch shape by the engine)
void gemm_schedule_tile_lists(const GemmParams* params, int num_sms, std::vector<int>* flat, int* pairs, int* stride);
// fp32 output tensor maps of the SAME [M, N] matrix (pitch ldd) at num_peers addresses (one per GPU), uploaded to a
// device array the caller owns (cudaFree): the per-owner targets of the fused GEMM -> reduce-scatter epilogue.
int gemm_build_peer_maps(void* const* peer_D, int num_peers, int M, int N, int ldd, CUtensorMap** d_out, char* err,
                         int errlen);
// Launch on `stream`.  `num_sms` = multiprocessor count of the current device.
int gemm_launch(const GemmParams& params, int num_sms, cudaStream_t stream);
// One-time (per process) kernel attribute setup; returns cudaError_t as int.
int gemm_init();
size_t gemm_smem_bytes();

}  // namespace tfk
