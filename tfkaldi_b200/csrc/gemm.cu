// Persistent grouped tcgen05 GEMM with fused epilogues (see gemm.cuh for the role on the hot path).
//
// CTA = 320 threads, one CTA per SM (smem-limited):
//   warp 0   : tile scheduler + TMA producer (one elected lane)
//   warp 1   : TMEM allocator + tcgen05.mma issuer (one lane)
//   warps 2-9: epilogue; warp w owns TMEM lanes / tile rows 32*(w%4) .. +31 and half of the columns
// Pipelines: smem ring full/empty (TMA <-> MMA), TMEM accumulator ring full/empty (MMA <-> epilogue,
// 2 x 256 columns so the next tile's MMAs overlap this tile's epilogue), tile-id ring (scheduler ->
// MMA/epilogue; tiles are claimed with an atomic counter so long wgrad tiles and short dgrad tiles of
// one launch balance across the 148 SMs).
#include "gemm.cuh"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "philox.cuh"
#include "ptx.cuh"

namespace tfk {

namespace {

constexpr uint32_t A_TILE_BYTES = BM * BK * 2;  // 16 KB
constexpr uint32_t B_TILE_BYTES = BN * BK * 2;  // 32 KB
constexpr uint32_t STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
constexpr uint32_t SLAB_BYTES = 32 * 128;  // 32 rows x 128 B, one epilogue warp's staging buffer
constexpr uint32_t SLABS_OFF = STAGES * STAGE_BYTES;
constexpr uint32_t BARS_OFF = SLABS_OFF + 8 * SLAB_BYTES;  // one slab per epilogue warp
constexpr int SCHED_DEPTH = 4;
constexpr uint32_t SMEM_USED = BARS_OFF + 256;
constexpr uint32_t SMEM_BYTES = SMEM_USED + 1024;  // slack for manual 1024-byte alignment
constexpr uint32_t MN_ATOM_BYTES = 64 * BK * 2;    // one 64(MN) x 64(K) TMA box = 8 KB

// CTA-pair kernel: per CTA a stage holds its 128 A rows and its 128 of the 256 B rows (16 KB each)
constexpr int STAGES2 = 6;
constexpr uint32_t STAGE2_BYTES = 2 * A_TILE_BYTES;
static_assert(STAGES2 * STAGE2_BYTES == SLABS_OFF, "both kernels share one smem carve-up");
struct SmemBars2 {
  uint64_t full[STAGES2];
  uint64_t empty[STAGES2];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t a_full, a_empty;  // A-stationary launches: the resident A panel of the current run of tiles
  uint32_t tmem_base;
};
// A-stationary launches (one problem, K <= 7 k-blocks, e.g. the layer-0 forward 8192 x 440 -> 2048): a pair works through
// CONSECUTIVE column tiles of one 256-row block, so its A panel (7 x 16 KB per CTA) is loaded once per run and stays in
// the first 112 KB of the operand area; only B streams through a ring of RES_STAGES 16 KB slots behind it.  Such a tile
// moves 224 KB instead of 448 KB per pair through L2 -> SM, which is what paces these short-K tiles.
constexpr int RES_MAX_KB = 7;
constexpr int RES_STAGES = 5;
static_assert(RES_MAX_KB * A_TILE_BYTES + RES_STAGES * A_TILE_BYTES <= SLABS_OFF, "resident layout must fit the operand area");
static_assert(RES_STAGES <= STAGES2, "the B ring reuses the full/empty barriers");
static_assert(sizeof(SmemBars2) <= 256, "barrier block too large");

struct SmemBars {
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t sched_full[SCHED_DEPTH];
  uint64_t sched_empty[SCHED_DEPTH];
  int tile_ring[SCHED_DEPTH];
  uint32_t tmem_base;
};
static_assert(sizeof(SmemBars) <= 256, "barrier block too large");

struct TileCoord {
  int p, m_blk, n_blk;
  int kb_begin, kb_count;  // k-block range of this work item (split-K)
};

// tile-list entry of the CTA-pair kernel: tile index | (half << 28); half 0 = all 256 columns of the tile,
// 1 / 2 = its left / right 128 columns (an M256 N128 MMA: same tensor throughput, half the time), -1 ends a list
constexpr int kHalfShift = 28;
constexpr int kTileMask = (1 << kHalfShift) - 1;

__device__ __forceinline__ TileCoord decode_tile(const GemmParams& P, int tile) {
  TileCoord t;
  t.p = (P.nprob > 1 && tile >= P.p[1].tile_begin) ? 1 : 0;
  const GemmProblem& pr = P.p[t.p];
  int local = tile - pr.tile_begin;
  const int per_split = pr.tiles_m * pr.tiles_n;
  const int split = local / per_split;
  local -= split * per_split;
  t.m_blk = local / pr.tiles_n;
  t.n_blk = local - t.m_blk * pr.tiles_n;
  t.kb_begin = split * pr.kb_per_split;
  t.kb_count = min(pr.kb_per_split, pr.num_kb - t.kb_begin);
  return t;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// Sum v[0..31] over the 32 lanes; lane L returns the total of column L (31 shuffles, not 160).
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], uint32_t lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int j = 0; j < s; ++j) {
      const float keep = up ? v[j + s] : v[j];
      const float send = up ? v[j] : v[j + s];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

// ------------------------------------------------------------------------------------------------
// Epilogue for one 128 x 256 accumulator tile.  Eight epilogue warps: warp w owns TMEM lanes / tile rows
// 32*(w%4).. and the 32-column chunks [c_begin, c_end) (half of the tile each; a lone warp per scheduler
// runs this dependent instruction stream at low IPC, two per scheduler nearly double the drain rate).
// One 4 KB staging slab per warp (slab_a); the bf16x3 store needs two (hi, lo): there warps 2-5 do the whole
// tile with their partner's slab as slab_b and warps 6-9 sit the tile out.
// ------------------------------------------------------------------------------------------------
// Compile-time feature mask of an epilogue instantiation: a feature outside the mask is compiled out, one inside it
// is still switched by the problem's run-time fields.  The frequent launches get lean instantiations (a few hundred
// instructions per 32-column chunk instead of ~6000 with every branch present: the generic body does not fit the
// instruction caches and its skipped blocks cost a fetch miss at every reconvergence point; profiles/r2_ncu_l0fwd_*),
// everything else runs the generic one.
enum EpiFeat : uint32_t {
  F_BIAS = 1u << 0,        // + bias
  F_STATS = 1u << 1,       // batch-norm statistics partials of z
  F_RELU = 1u << 2,        // act == 1
  F_ACT_SMOOTH = 1u << 3,  // act == 2 / 3 (sigmoid / tanh)
  F_DROPOUT = 1u << 4,     // forward dropout (Philox)
  F_BITS_OUT = 1u << 5,    // write the 1-bit gradient-pass mask
  F_BITS_IN = 1u << 6,     // backward: mask from the 1-bit mask
  F_MASK_DERIV = 1u << 7,  // backward: sigmoid / tanh slope from the stored output
  F_MASK_RELU = 1u << 8,   // backward: mask from the stored output (> 0 or != 0)
  F_COLSUM = 1u << 9,      // column-sum partials of the stored value
  F_COLSUM2 = 1u << 10,    // batch-norm backward: column-sum partials of value * xhat
  F_BWD_PHILOX = 1u << 11, // backward: the dropout keep bits are re-drawn (chains where a stored 0 is not proof of a drop)
  F_ALL = (1u << 12) - 1,
};

// The problem description lives in the kernel's parameter block and is indexed by a run-time problem number: every
// `pr.x` inside the chunk loop compiles to an indexed constant load, per element where the use is predicated.  The lean
// instantiations read each field once per tile into a local copy (fields an instantiation does not use are dead and cost
// nothing); the generic one, which needs every register it has, keeps reading through the reference.
struct EpiFields {
  int M;
  int N;
  int act;
  const float* bias;
  const float* bn_beta;
  int bn_from_y;
  const float* bn_mean;
  const float* bn_rstd;
  const __nv_bfloat16* bn_z_hi;
  int bn_z_ld;
  const __nv_bfloat16* bn_z_lo;
  float* colsum2_part;
  int colsum_ld;
  float* colsum_part;
  int deriv;
  unsigned int drop_thr;
  unsigned int bwd_drop_thr;
  int dropout_in_chain;
  float keep_inv;
  const uint32_t* mask_bits_in;
  int mask_bits_ld;
  uint32_t* mask_bits_out;
  int mask_ld;
  int mask_nonzero;
  const __nv_bfloat16* mask_src;
  const __nv_bfloat16* mask_src_lo;
  int num_peers;
  const CUtensorMap* peer_tm;
  int rows_per_owner;
  float scale;
  unsigned long long seed;
  int stat_ld;
  float* stat_sq;
  float* stat_sum;
  __device__ __forceinline__ explicit EpiFields(const GemmProblem& p)
      : M(p.M),
        N(p.N),
        act(p.act),
        bias(p.bias),
        bn_beta(p.bn_beta),
        bn_from_y(p.bn_from_y),
        bn_mean(p.bn_mean),
        bn_rstd(p.bn_rstd),
        bn_z_hi(p.bn_z_hi),
        bn_z_ld(p.bn_z_ld),
        bn_z_lo(p.bn_z_lo),
        colsum2_part(p.colsum2_part),
        colsum_ld(p.colsum_ld),
        colsum_part(p.colsum_part),
        deriv(p.deriv),
        drop_thr(p.drop_thr),
        bwd_drop_thr(p.bwd_drop_thr),
        dropout_in_chain(p.dropout_in_chain),
        keep_inv(p.keep_inv),
        mask_bits_in(p.mask_bits_in),
        mask_bits_ld(p.mask_bits_ld),
        mask_bits_out(p.mask_bits_out),
        mask_ld(p.mask_ld),
        mask_nonzero(p.mask_nonzero),
        mask_src(p.mask_src),
        mask_src_lo(p.mask_src_lo),
        num_peers(p.num_peers),
        peer_tm(p.peer_tm),
        rows_per_owner(p.rows_per_owner),
        scale(p.scale),
        seed(p.seed),
        stat_ld(p.stat_ld),
        stat_sq(p.stat_sq),
        stat_sum(p.stat_sum) {}
};
template <uint32_t FEAT>
struct EpiView {
  using type = const EpiFields;
};
template <>
struct EpiView<F_ALL> {
  using type = const GemmProblem&;
};

// Which instantiation serves a problem (decided on the host, gemm_build_params): a lean one when the problem uses no
// feature outside its mask, else the generic one.
enum EpiVariant : int { EV_GENERIC = 0, EV_FWD_RELU_BITS, EV_FWD_STATS, EV_DGRAD_BITS_COLSUM, EV_DGRAD_BN, EV_PLAIN };
constexpr uint32_t kFeatOf[] = {F_ALL, F_BIAS | F_RELU | F_BITS_OUT, F_BIAS | F_STATS, F_BITS_IN | F_COLSUM,
                                F_MASK_RELU | F_COLSUM | F_COLSUM2, F_BIAS};

// Everything between the accumulator read and the store of one 32-column chunk: the fused FFLayer arithmetic
// (classifiers/layer.py:52-56 forward; its tf.gradients twin backward) on v[j] = element (row, col0 + j).
template <uint32_t FEAT, typename E>
__device__ __forceinline__ void epilogue_math(const E& e, float (&v)[32], int col0, int row, bool row_ok, uint32_t lane,
                                              uint32_t q, int m0) {
  float yv[32];  // stored forward output of the layer below (live only in instantiations with F_COLSUM2)
#pragma unroll
  for (int j = 0; j < 32; ++j) yv[j] = 0.f;
  if ((FEAT & F_BIAS) && e.bias != nullptr) {  // bias buffers are padded to a multiple of 256 floats
    const float4* bp = reinterpret_cast<const float4*>(e.bias + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = __ldg(bp + j);
      v[4 * j + 0] += b.x;
      v[4 * j + 1] += b.y;
      v[4 * j + 2] += b.z;
      v[4 * j + 3] += b.w;
    }
  }
  if ((FEAT & F_STATS) && e.stat_sum != nullptr) {  // batch-norm statistics of z = xW + b over this warp's 32 rows
    float s1[32], s2[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float z = row_ok ? v[j] : 0.f;
      s1[j] = z;
      s2[j] = z * z;
    }
    const float t1 = warp_transpose_reduce(s1, lane);
    const float t2 = warp_transpose_reduce(s2, lane);
    const int col = col0 + static_cast<int>(lane);
    if (col < e.N) {
      const size_t o = static_cast<size_t>((m0 >> 5) + q) * e.stat_ld + col;
      e.stat_sum[o] = t1;
      e.stat_sq[o] = t2;
    }
  }
  if ((FEAT & F_RELU) && e.act == 1) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  } else if ((FEAT & F_ACT_SMOOTH) && e.act == 2) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 1.0f / (1.0f + expf(-v[j]));
  } else if ((FEAT & F_ACT_SMOOTH) && e.act == 3) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
  }
  uint32_t keepword = 0xFFFFFFFFu;  // dropout keep decisions of this chunk's 32 columns (all ones without dropout)
  if ((FEAT & F_DROPOUT) && e.drop_thr != 0u) {
    keepword = 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // one Philox call per 8 columns
      const Philox4 rnd =
          philox4x32_10(static_cast<uint32_t>(col0 >> 3) + j, static_cast<uint32_t>(row), 0u, 0u,
                        static_cast<uint32_t>(e.seed), static_cast<uint32_t>(e.seed >> 32));
      const uint32_t keep = dropout_keep_bits(rnd, e.drop_thr);
      keepword |= keep << (8 * j);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[8 * j + k] = ((keep >> k) & 1u) ? v[8 * j + k] * e.keep_inv : 0.f;
    }
  }
  if ((FEAT & F_BITS_OUT) && e.mask_bits_out != nullptr) {  // remember which units pass gradient: 1 bit each, coalesced store
    uint32_t word;
    if (e.mask_nonzero) {
      // identity non-linearity: the gradient passes wherever dropout kept the unit — the keep decision itself, not
      // "stored value != 0" (a kept pre-activation that is exactly 0 would otherwise lose its gradient)
      word = keepword;
    } else {
      uint32_t b4[4] = {0u, 0u, 0u, 0u};  // four independent chains (two epilogue warps per scheduler: ILP is all there is)
#pragma unroll
      for (int j = 0; j < 32; ++j) b4[j & 3] |= (v[j] > 0.f ? 1u : 0u) << j;
      word = (b4[0] | b4[1]) | (b4[2] | b4[3]);
    }
    if (row_ok) e.mask_bits_out[static_cast<size_t>(col0 >> 5) * e.mask_bits_ld + row] = word;
  }
  if ((FEAT & F_BWD_PHILOX) && e.bwd_drop_thr != 0u) {  // backward: re-draw the forward pass's keep decisions
    keepword = 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      keepword |= dropout_keep_bits(philox4x32_10(static_cast<uint32_t>(col0 >> 3) + j, static_cast<uint32_t>(row), 0u, 0u,
                                                  static_cast<uint32_t>(e.seed), static_cast<uint32_t>(e.seed >> 32)),
                                    e.bwd_drop_thr) << (8 * j);
  }
  if ((FEAT & F_BITS_IN) && e.mask_bits_in != nullptr) {  // backward of relu(+dropout) from the forward pass's bit mask
    const uint32_t bits = row_ok ? __ldg(e.mask_bits_in + static_cast<size_t>(col0 >> 5) * e.mask_bits_ld + row) : 0u;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = ((bits >> j) & 1u) ? v[j] * e.scale : 0.f;
  }
  if ((FEAT & F_MASK_DERIV) && e.mask_src != nullptr && e.deriv != 0) {
    // backward of sigmoid / tanh (+dropout) from the stored forward output a = f(z) * dropmask / keep
    const float keep = 1.0f / e.scale;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const bool ok = row_ok && (col0 + j) < e.N;
      const size_t o = static_cast<size_t>(row) * e.mask_ld + col0 + j;
      float a = ok ? __bfloat162float(e.mask_src[o]) : 0.f;
      if (ok && e.mask_src_lo != nullptr) a += __bfloat162float(e.mask_src_lo[o]);
      const float y = a * keep;
      const float d = e.deriv == 1 ? y * (1.0f - y) : 1.0f - y * y;
      const bool dropped = e.dropout_in_chain && (((FEAT & F_BWD_PHILOX) && e.bwd_drop_thr != 0u) ? ((keepword >> j) & 1u) == 0u : a == 0.f);
      v[j] = (ok && !dropped) ? v[j] * d * e.scale : 0.f;
    }
  } else
  if ((FEAT & F_MASK_RELU) && e.mask_src != nullptr) {  // backward of relu(+dropout): pass where the forward output was > 0
    const __nv_bfloat16* mp = e.mask_src + static_cast<size_t>(row) * e.mask_ld + col0;
    const bool keep_y = (FEAT & F_COLSUM2) && e.bn_from_y;  // the batch-norm sums below want the values, not just the signs
    if (row_ok && col0 + 32 <= e.N) {
      const uint4* mp4 = reinterpret_cast<const uint4*>(mp);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 m = __ldg(mp4 + j);
        const uint32_t w[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // bf16 > 0  <=>  sign bit clear and magnitude non-zero
          const uint32_t lo = w[k] & 0xFFFFu, hi = w[k] >> 16;
          bool plo = (e.mask_nonzero || (lo & 0x8000u) == 0) && (lo & 0x7FFFu) != 0;
          bool phi = (e.mask_nonzero || (hi & 0x8000u) == 0) && (hi & 0x7FFFu) != 0;
          if ((FEAT & F_BWD_PHILOX) && e.bwd_drop_thr != 0u && e.mask_nonzero) {  // identity + dropout: the keep bit decides
            plo = (keepword >> (8 * j + 2 * k)) & 1u;
            phi = (keepword >> (8 * j + 2 * k + 1)) & 1u;
          }
          v[8 * j + 2 * k + 0] = plo ? v[8 * j + 2 * k + 0] * e.scale : 0.f;
          v[8 * j + 2 * k + 1] = phi ? v[8 * j + 2 * k + 1] * e.scale : 0.f;
          if (keep_y) {
            yv[8 * j + 2 * k + 0] = __uint_as_float(w[k] << 16);
            yv[8 * j + 2 * k + 1] = __uint_as_float(w[k] & 0xFFFF0000u);
          }
        }
      }
      if (keep_y && e.mask_src_lo != nullptr) {  // bf16x3: the stored output is hi + lo
        const uint4* lp4 = reinterpret_cast<const uint4*>(e.mask_src_lo + static_cast<size_t>(row) * e.mask_ld + col0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 m = __ldg(lp4 + j);
          const uint32_t w[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            yv[8 * j + 2 * k + 0] += __uint_as_float(w[k] << 16);
            yv[8 * j + 2 * k + 1] += __uint_as_float(w[k] & 0xFFFF0000u);
          }
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const bool ok = row_ok && (col0 + j) < e.N;
        float mv = ok ? __bfloat162float(mp[j]) : 0.f;
        bool pass = e.mask_nonzero ? (mv != 0.f) : (mv > 0.f);
        if ((FEAT & F_BWD_PHILOX) && e.bwd_drop_thr != 0u && e.mask_nonzero) pass = ok && ((keepword >> j) & 1u);
        v[j] = pass ? v[j] * e.scale : 0.f;
        if (keep_y) {
          if (ok && e.mask_src_lo != nullptr)
            mv += __bfloat162float(e.mask_src_lo[static_cast<size_t>(row) * e.mask_ld + col0 + j]);
          yv[j] = mv;
        }
      }
    }
  }

  if ((FEAT & F_COLSUM) && e.colsum_part != nullptr) {  // column sums of what is about to be stored (bias gradient below)
    float t[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) t[j] = row_ok ? v[j] : 0.f;
    const float tot = warp_transpose_reduce(t, lane);
    const int col = col0 + static_cast<int>(lane);
    if (col < e.N) e.colsum_part[static_cast<size_t>((m0 >> 5) + q) * e.colsum_ld + col] = tot;
  }
  if ((FEAT & F_COLSUM2) && e.colsum2_part != nullptr && e.bn_from_y) {
    // batch-norm backward, column sums of dY * xhat WITHOUT reading z: the layer below stored y = f(xhat + beta) * keepmask
    // / keep with f = relu or identity, and dY (v, already masked) is zero wherever y is, so on every element that
    // counts xhat = y * keep - beta — from the values the mask was just derived from (no extra operand traffic in a
    // kernel whose main loop is bound by L2 -> SM delivery)
    float t[32];
    const float keep = 1.0f / e.scale;
    const float4* bp = reinterpret_cast<const float4*>(e.bn_beta + col0);  // padded to a multiple of 256 floats
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 be = __ldg(bp + j);
      t[4 * j + 0] = v[4 * j + 0] * (yv[4 * j + 0] * keep - be.x);
      t[4 * j + 1] = v[4 * j + 1] * (yv[4 * j + 1] * keep - be.y);
      t[4 * j + 2] = v[4 * j + 2] * (yv[4 * j + 2] * keep - be.z);
      t[4 * j + 3] = v[4 * j + 3] * (yv[4 * j + 3] * keep - be.w);
    }
    const float tot = warp_transpose_reduce(t, lane);
    const int col = col0 + static_cast<int>(lane);
    if (col < e.N) e.colsum2_part[static_cast<size_t>((m0 >> 5) + q) * e.colsum_ld + col] = tot;
  } else if ((FEAT & F_COLSUM2) && e.colsum2_part != nullptr) {  // general form: xhat = (z - mean) * rstd from the stored z
    float t[32];
    const bool full = row_ok && col0 + 32 <= e.N;
    const __nv_bfloat16* zp = e.bn_z_hi + static_cast<size_t>(row) * e.bn_z_ld + col0;
    if (full) {
      const uint4* z4 = reinterpret_cast<const uint4*>(zp);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 zz = __ldg(z4 + j);
        const uint32_t w[4] = {zz.x, zz.y, zz.z, zz.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          t[8 * j + 2 * k + 0] = __uint_as_float(w[k] << 16);
          t[8 * j + 2 * k + 1] = __uint_as_float(w[k] & 0xFFFF0000u);
        }
      }
      if (e.bn_z_lo != nullptr) {
        const uint4* l4 = reinterpret_cast<const uint4*>(e.bn_z_lo + static_cast<size_t>(row) * e.bn_z_ld + col0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 zz = __ldg(l4 + j);
          const uint32_t w[4] = {zz.x, zz.y, zz.z, zz.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            t[8 * j + 2 * k + 0] += __uint_as_float(w[k] << 16);
            t[8 * j + 2 * k + 1] += __uint_as_float(w[k] & 0xFFFF0000u);
          }
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const bool ok = row_ok && (col0 + j) < e.N;
        float z = ok ? __bfloat162float(zp[j]) : 0.f;
        if (ok && e.bn_z_lo != nullptr)
          z += __bfloat162float(e.bn_z_lo[static_cast<size_t>(row) * e.bn_z_ld + col0 + j]);
        t[j] = z;
      }
    }
    const float4* mp = reinterpret_cast<const float4*>(e.bn_mean + col0);  // padded to a multiple of 256 floats
    const float4* rp = reinterpret_cast<const float4*>(e.bn_rstd + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 mu = __ldg(mp + j), rs = __ldg(rp + j);
      t[4 * j + 0] = row_ok ? v[4 * j + 0] * ((t[4 * j + 0] - mu.x) * rs.x) : 0.f;
      t[4 * j + 1] = row_ok ? v[4 * j + 1] * ((t[4 * j + 1] - mu.y) * rs.y) : 0.f;
      t[4 * j + 2] = row_ok ? v[4 * j + 2] * ((t[4 * j + 2] - mu.z) * rs.z) : 0.f;
      t[4 * j + 3] = row_ok ? v[4 * j + 3] * ((t[4 * j + 3] - mu.w) * rs.w) : 0.f;
    }
    const float tot = warp_transpose_reduce(t, lane);
    const int col = col0 + static_cast<int>(lane);
    if (col < e.N) e.colsum2_part[static_cast<size_t>((m0 >> 5) + q) * e.colsum_ld + col] = tot;
  }

}

template <int OUT, uint32_t FEAT>
__device__ __forceinline__ void epilogue_tile(const GemmProblem& pr, uint32_t tmem_acc, int m0,
                                              int n0, uint32_t q, uint32_t lane, uint32_t slab_a,
                                              uint32_t slab_b, int c_begin, int c_end,
                                              float4 bias_pre = make_float4(0.f, 0.f, 0.f, 0.f), bool has_bias_pre = false) {
  typename EpiView<FEAT>::type e(pr);
  const int row = m0 + static_cast<int>(q * 32 + lane);
  const bool row_ok = row < e.M;
  const uint32_t lane_taddr = tmem_acc + ((q * 32u) << 16);
  const uint32_t row_off = lane * 128u;
  const uint32_t sw = lane & 7u;

  if constexpr (OUT == OUT_BF16 && (FEAT == kFeatOf[EV_FWD_RELU_BITS] || FEAT == kFeatOf[EV_DGRAD_BITS_COLSUM])) {
    // The two hottest bf16 instantiations (pairing the batch-norm ones spills): the two chunks of a 64-column store group are read from TMEM together and their
    // arithmetic is one straight-line block, so the compiler interleaves two independent instruction streams — with two
    // epilogue warps per scheduler, instruction-level parallelism is the only latency hiding there is
    // (profiles/r2g_ncu_l0fwd_stalls.txt: a third of the epilogue's samples were fixed-latency dependency stalls).
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 2) {
      const int col0 = n0 + c * 32;
      if (col0 >= e.N) break;  // warp-uniform
      uint32_t r0[32], r1[32];
      tmem_ld_32x32(lane_taddr + static_cast<uint32_t>(c * 32), r0);
      tmem_ld_32x32(lane_taddr + static_cast<uint32_t>(c * 32 + 32), r1);
      tmem_ld_wait();
      float v0[32], v1[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v0[j] = __uint_as_float(r0[j]);
        v1[j] = __uint_as_float(r1[j]);
      }
      if ((FEAT & F_BIAS) && has_bias_pre) {
        // The warp's 128 bias values were loaded before the accumulator was waited for, four consecutive floats per lane
        // (lane l: columns 4 l .. 4 l + 3 of the warp's column range); every lane needs all 32 of a chunk: broadcast.
        // (Per-chunk loads of the bias stalled on first-touch L2 latency — each tile uses another 1 KB of it — and
        // neither an L1 prefetch nor issuing them a tile ahead removed that: profiles/r2i_ncu_l0fwd_stalls.txt.)
        const int lb = (c - c_begin) * 8;  // lane holding column 0 of chunk c
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v0[4 * j + 0] += __shfl_sync(0xffffffffu, bias_pre.x, lb + j);
          v0[4 * j + 1] += __shfl_sync(0xffffffffu, bias_pre.y, lb + j);
          v0[4 * j + 2] += __shfl_sync(0xffffffffu, bias_pre.z, lb + j);
          v0[4 * j + 3] += __shfl_sync(0xffffffffu, bias_pre.w, lb + j);
          v1[4 * j + 0] += __shfl_sync(0xffffffffu, bias_pre.x, lb + 8 + j);
          v1[4 * j + 1] += __shfl_sync(0xffffffffu, bias_pre.y, lb + 8 + j);
          v1[4 * j + 2] += __shfl_sync(0xffffffffu, bias_pre.z, lb + 8 + j);
          v1[4 * j + 3] += __shfl_sync(0xffffffffu, bias_pre.w, lb + 8 + j);
        }
        epilogue_math<FEAT & ~F_BIAS>(e, v0, col0, row, row_ok, lane, q, m0);
        epilogue_math<FEAT & ~F_BIAS>(e, v1, col0 + 32, row, row_ok, lane, q, m0);
      } else {
        epilogue_math<FEAT>(e, v0, col0, row, row_ok, lane, q, m0);
        epilogue_math<FEAT>(e, v1, col0 + 32, row, row_ok, lane, q, m0);
      }
      if (lane == 0) tma_wait_group_read<0>();  // the previous store has finished reading the slab
      __syncwarp();
      const uint32_t slab = slab_a + row_off;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slab + ((static_cast<uint32_t>(j) ^ sw) << 4)),
                     "r"(pack_bf16x2(v0[8 * j + 0], v0[8 * j + 1])), "r"(pack_bf16x2(v0[8 * j + 2], v0[8 * j + 3])),
                     "r"(pack_bf16x2(v0[8 * j + 4], v0[8 * j + 5])), "r"(pack_bf16x2(v0[8 * j + 6], v0[8 * j + 7]))
                     : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slab + ((static_cast<uint32_t>(4 + j) ^ sw) << 4)),
                     "r"(pack_bf16x2(v1[8 * j + 0], v1[8 * j + 1])), "r"(pack_bf16x2(v1[8 * j + 2], v1[8 * j + 3])),
                     "r"(pack_bf16x2(v1[8 * j + 4], v1[8 * j + 5])), "r"(pack_bf16x2(v1[8 * j + 6], v1[8 * j + 7]))
                     : "memory");
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&pr.tmD[0], slab_a, col0, m0 + static_cast<int>(q * 32));
        tma_commit_group();
      }
    }
    return;
  }

#pragma unroll 1
  for (int c = c_begin; c < c_end; ++c) {
    const int col0 = n0 + c * 32;
    // bf16 outputs are emitted in 64-column groups (two chunks), fp32 outputs per 32-column chunk
    const int group_col0 = (OUT == OUT_BF16 || OUT == OUT_BF16_SPLIT) ? (n0 + (c & ~1) * 32) : col0;
    if (group_col0 >= e.N) break;  // warp-uniform

    uint32_t r[32];
    tmem_ld_32x32(lane_taddr + static_cast<uint32_t>(c * 32), r);
    tmem_ld_wait();
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    epilogue_math<FEAT>(e, v, col0, row, row_ok, lane, q, m0);

    if constexpr (OUT == OUT_F32 || OUT == OUT_F32_REDADD) {
      if (lane == 0) tma_wait_group_read<0>();  // the previous store has finished reading the slab
      __syncwarp();
      const uint32_t slab = slab_a + row_off;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t addr = slab + ((static_cast<uint32_t>(j) ^ sw) << 4);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[4 * j]),
                     "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                     : "memory");
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        const uint32_t src = slab_a;
        if constexpr (OUT == OUT_F32) {
          tma_store_2d(&pr.tmD[0], src, col0, m0 + static_cast<int>(q * 32));
        } else {
          // fused reduce-scatter: this 32-row slab belongs to one owner GPU; add it there over NVLink
          const CUtensorMap* tm = &pr.tmD[0];
          if (e.peer_tm != nullptr) {
            int owner = (m0 + static_cast<int>(q * 32)) / e.rows_per_owner;
            owner = owner < e.num_peers ? owner : e.num_peers - 1;
            tm = e.peer_tm + owner;
          }
          tma_reduce_add_2d(tm, src, col0, m0 + static_cast<int>(q * 32));
        }
        tma_commit_group();
      }
    } else if constexpr (OUT == OUT_BF16) {
      const int half = c & 1;
      if (half == 0) {
        if (lane == 0) tma_wait_group_read<0>();
        __syncwarp();
      }
      const uint32_t slab = slab_a + row_off;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t addr = slab + ((static_cast<uint32_t>(half * 4 + j) ^ sw) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr),
                     "r"(pack_bf16x2(v[8 * j + 0], v[8 * j + 1])),
                     "r"(pack_bf16x2(v[8 * j + 2], v[8 * j + 3])),
                     "r"(pack_bf16x2(v[8 * j + 4], v[8 * j + 5])),
                     "r"(pack_bf16x2(v[8 * j + 6], v[8 * j + 7]))
                     : "memory");
      }
      if (half == 1) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&pr.tmD[0], slab_a, group_col0, m0 + static_cast<int>(q * 32));
          tma_commit_group();
        }
      }
    } else {  // OUT_BF16_SPLIT: slab_a = hi, slab_b = lo
      const int half = c & 1;
      if (half == 0) {
        if (lane == 0) tma_wait_group_read<0>();
        __syncwarp();
      }
      const uint32_t slab_hi = slab_a + row_off;
      const uint32_t slab_lo = slab_b + row_off;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float a = v[8 * j + 2 * k], b = v[8 * j + 2 * k + 1];
          const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
          const float ar = a - __bfloat162float(ah), br = b - __bfloat162float(bh);
          __nv_bfloat162 h2;
          h2.x = ah;
          h2.y = bh;
          hi[k] = *reinterpret_cast<uint32_t*>(&h2);
          lo[k] = pack_bf16x2(ar, br);
        }
        const uint32_t off = ((static_cast<uint32_t>(half * 4 + j) ^ sw) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slab_hi + off), "r"(hi[0]),
                     "r"(hi[1]), "r"(hi[2]), "r"(hi[3])
                     : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slab_lo + off), "r"(lo[0]),
                     "r"(lo[1]), "r"(lo[2]), "r"(lo[3])
                     : "memory");
      }
      if (half == 1) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&pr.tmD[0], slab_a, group_col0, m0 + static_cast<int>(q * 32));
          tma_store_2d(&pr.tmD[1], slab_b, group_col0, m0 + static_cast<int>(q * 32));
          tma_commit_group();
        }
      }
    }
  }
}

// Tile-level dispatch shared by both kernels.  `ew` = epilogue warp index 0..7.
__device__ __forceinline__ void epilogue_dispatch(const GemmProblem& pr, uint32_t tmem_acc, int m0, int n0,
                                                  uint32_t ew, uint32_t lane, uint32_t slabs, int chunks = BN / 32,
                                                  float4 bias_pre = make_float4(0.f, 0.f, 0.f, 0.f), bool has_bias_pre = false) {
  // the hardware ties a warp to TMEM lanes 32 * (warp id % 4): epilogue warp ew is CTA warp ew + 2
  const uint32_t q = (ew + 2) & 3, part = ew >> 2;
  const uint32_t slab_a = slabs + ew * SLAB_BYTES;
  if (pr.out_kind == OUT_BF16_SPLIT) {
    // hi and lo need two slabs: warps 0-3 take the whole tile with their partner's (ew+4) slab; the pair
    // meets on a named barrier afterwards so the partner never writes a slab that is still being read
    if (part == 1) {  // entry: my slab may still be read by my previous (non-split) store
      if (lane == 0) tma_wait_group_read<0>();
      __syncwarp();
    }
    asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
    if (part == 0) {
      epilogue_tile<OUT_BF16_SPLIT, F_ALL>(pr, tmem_acc, m0, n0, q, lane, slab_a, slab_a + 4 * SLAB_BYTES, 0, chunks);
      if (lane == 0) tma_wait_group_read<0>();
      __syncwarp();
    }
    asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
    return;
  }
  const int c0 = static_cast<int>(part) * (chunks / 2), c1 = c0 + chunks / 2;  // chunks = 8 (256 cols) or 4 (half tile)
  switch (pr.out_kind * 8 + pr.epi_variant) {
    case OUT_BF16 * 8 + EV_FWD_RELU_BITS:
      epilogue_tile<OUT_BF16, kFeatOf[EV_FWD_RELU_BITS]>(pr, tmem_acc, m0, n0, q, lane, slab_a, 0u, c0, c1, bias_pre, has_bias_pre);
      break;
    case OUT_BF16 * 8 + EV_FWD_STATS:
      epilogue_tile<OUT_BF16, kFeatOf[EV_FWD_STATS]>(pr, tmem_acc, m0, n0, q, lane, slab_a, 0u, c0, c1);
      break;
    case OUT_BF16 * 8 + EV_DGRAD_BITS_COLSUM:
      epilogue_tile<OUT_BF16, kFeatOf[EV_DGRAD_BITS_COLSUM]>(pr, tmem_acc, m0, n0, q, lane, slab_a, 0u, c0, c1);
      break;
    case OUT_BF16 * 8 + EV_DGRAD_BN:
      epilogue_tile<OUT_BF16, kFeatOf[EV_DGRAD_BN]>(pr, tmem_acc, m0, n0, q, lane, slab_a, 0u, c0, c1);
      break;
    case OUT_F32 * 8 + EV_PLAIN:
      epilogue_tile<OUT_F32, kFeatOf[EV_PLAIN]>(pr, tmem_acc, m0, n0, q, lane, slab_a, 0u, c0, c1);
      break;
    case OUT_F32_REDADD * 8 + EV_PLAIN:
      epilogue_tile<OUT_F32_REDADD, 0u>(pr, tmem_acc, m0, n0, q, lane, slab_a, 0u, c0, c1);
      break;
    default:
      if (pr.out_kind == OUT_BF16)
        epilogue_tile<OUT_BF16, F_ALL>(pr, tmem_acc, m0, n0, q, lane, slab_a, 0u, c0, c1);
      else if (pr.out_kind == OUT_F32)
        epilogue_tile<OUT_F32, F_ALL>(pr, tmem_acc, m0, n0, q, lane, slab_a, 0u, c0, c1);
      else
        epilogue_tile<OUT_F32_REDADD, F_ALL>(pr, tmem_acc, m0, n0, q, lane, slab_a, 0u, c0, c1);
      break;
  }
}

// Called by every lane of epilogue warp `ew` before it waits for the accumulator: L1 prefetch of the lines the
// epilogue will read with plain loads for this tile (layout as in epilogue_tile / epilogue_dispatch).
__device__ __forceinline__ void epilogue_prefetch(const GemmProblem& pr, int m0, int n0, uint32_t ew, uint32_t lane, int chunks) {
  if (pr.out_kind == OUT_BF16_SPLIT) return;  // (the hi+lo path is bound by its three MMA passes, not by the epilogue)
  const uint32_t q = (ew + 2) & 3, part = ew >> 2;
  const int c0 = static_cast<int>(part) * (chunks / 2), nch = chunks / 2;
  if (static_cast<int>(lane) >= nch) return;
  const int col0 = n0 + (c0 + static_cast<int>(lane)) * 32;  // lane i covers chunk c0 + i: 32 columns = one 128-byte line of floats
  if (col0 >= pr.N) return;
  if (pr.bias != nullptr) asm volatile("prefetch.global.L1 [%0];" ::"l"(pr.bias + col0));
  if (pr.mask_bits_in != nullptr) {
    const int row0 = m0 + static_cast<int>(q) * 32;
    if (row0 < pr.M) asm volatile("prefetch.global.L1 [%0];" ::"l"(pr.mask_bits_in + static_cast<size_t>(col0 >> 5) * pr.mask_bits_ld + row0));
  }
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
tfk_gemm_kernel(const __grid_constant__ GemmParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  SmemBars* bars = reinterpret_cast<SmemBars*>(smem_gen + BARS_OFF);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int p = 0; p < P.nprob; ++p) {
      tma_prefetch_desc(&P.p[p].tmA[0]);
      tma_prefetch_desc(&P.p[p].tmB[0]);
      tma_prefetch_desc(&P.p[p].tmD[0]);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < STAGES; ++i) {
        mbar_init(&bars->full[i], 1);
        mbar_init(&bars->empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bars->tmem_full[i], 1);
        mbar_init(&bars->tmem_empty[i], 8);
      }
      for (int i = 0; i < SCHED_DEPTH; ++i) {
        mbar_init(&bars->sched_full[i], 1);
        mbar_init(&bars->sched_empty[i], 9);  // MMA thread + 8 epilogue warps
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(&bars->tmem_base, TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ============================ scheduler + TMA producer ============================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      int tile = blockIdx.x;
      for (int it = 0;; ++it) {
        const int slot = it % SCHED_DEPTH;
        mbar_wait(&bars->sched_empty[slot], (((it / SCHED_DEPTH) & 1) ^ 1));
        bars->tile_ring[slot] = tile;
        mbar_arrive(&bars->sched_full[slot]);
        if (tile >= P.total_tiles) break;

        const TileCoord tc = decode_tile(P, tile);
        const GemmProblem& pr = P.p[tc.p];
        const int m0 = tc.m_blk * BM, n0 = tc.n_blk * BN;
        // bf16x3: a k-block is staged ONCE as two consecutive slots, (A_hi, B_hi) then (A_lo, B_lo); the MMA warp
        // forms the three products A_hi.B_hi + A_hi.B_lo + A_lo.B_hi from them (4 operand tiles per k-block
        // instead of the 6 that three separate (A, B) pairs would move)
        const int loads = pr.nsplit == 3 ? 2 : 1;
        const int iters = tc.kb_count * loads;
        int kb = tc.kb_begin, s = 0;
        for (int i = 0; i < iters; ++i) {
          mbar_wait(&bars->empty[stage], phase ^ 1);
          mbar_expect_tx(&bars->full[stage], STAGE_BYTES);
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint32_t sb = sa + A_TILE_BYTES;
          const CUtensorMap* ta = &pr.tmA[s];
          const CUtensorMap* tb = &pr.tmB[s];
          const int k0 = kb * BK;
          if (pr.a_mn) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_2d(sa + j * MN_ATOM_BYTES, ta, &bars->full[stage], m0 + 64 * j, k0);
          } else {
            tma_load_2d(sa, ta, &bars->full[stage], k0, m0);
          }
          if (pr.b_mn) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sb + j * MN_ATOM_BYTES, tb, &bars->full[stage], n0 + 64 * j, k0);
          } else {
            tma_load_2d(sb, tb, &bars->full[stage], k0, n0);
          }
          if (++s == loads) {
            s = 0;
            ++kb;
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        tile = atomicAdd(&P.sched[0], 1) + static_cast<int>(gridDim.x);
      }
      // self-resetting scheduler counters: the last CTA to finish claiming tiles zeroes them
      __threadfence();
      const int done = atomicAdd(&P.sched[1], 1);
      if (done == static_cast<int>(gridDim.x) - 1) {
        P.sched[0] = 0;
        P.sched[1] = 0;
        __threadfence();
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int it = 0;; ++it) {
        const int slot = it % SCHED_DEPTH;
        mbar_wait(&bars->sched_full[slot], (it / SCHED_DEPTH) & 1);
        const int tile = bars->tile_ring[slot];
        mbar_arrive(&bars->sched_empty[slot]);
        if (tile >= P.total_tiles) break;

        const TileCoord tc = decode_tile(P, tile);
        const GemmProblem& pr = P.p[tc.p];
        const uint32_t as = it & 1, aphase = (it >> 1) & 1;
        mbar_wait(&bars->tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + as * BN;
        const uint32_t idesc = make_idesc_bf16(BM, BN, pr.a_mn, pr.b_mn);
        // K-major: 8-row x 128B atoms, SBO = 1024 B, advance 32 B per UMMA_K inside the swizzled row.
        // MN-major: 64(MN) x 8(K) atoms, SBO = 1024 B between K groups, LBO = 8 KB between MN atoms,
        //           advance 2 K-groups (2048 B) per UMMA_K.
        const uint32_t a_lbo = pr.a_mn ? MN_ATOM_BYTES : 16u, b_lbo = pr.b_mn ? MN_ATOM_BYTES : 16u;
        const uint32_t a_kadv = pr.a_mn ? 2048u : 32u, b_kadv = pr.b_mn ? 2048u : 32u;
        for (int i = 0; i < tc.kb_count; ++i) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint32_t sb = sa + A_TILE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = make_smem_desc_sw128(sa + k * a_kadv, a_lbo, 1024u);
            const uint64_t db = make_smem_desc_sw128(sb + k * b_kadv, b_lbo, 1024u);
            umma_bf16(tmem_acc, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
          if (pr.nsplit == 3) {  // + A_hi.B_lo + A_lo.B_hi with the low halves from the next slot
            const uint32_t hi_stage = stage;
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
            mbar_wait(&bars->full[stage], phase);
            tc_fence_after();
            const uint32_t la = smem_base + stage * STAGE_BYTES;
            const uint32_t lb = la + A_TILE_BYTES;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16(tmem_acc, make_smem_desc_sw128(sa + k * a_kadv, a_lbo, 1024u),
                        make_smem_desc_sw128(lb + k * b_kadv, b_lbo, 1024u), idesc, 1u);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16(tmem_acc, make_smem_desc_sw128(la + k * a_kadv, a_lbo, 1024u),
                        make_smem_desc_sw128(sb + k * b_kadv, b_lbo, 1024u), idesc, 1u);
            umma_commit(&bars->empty[hi_stage]);
          }
          umma_commit(&bars->empty[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&bars->tmem_full[as]);  // accumulator complete -> epilogue
      }
    }
  } else {
    // ============================ epilogue warps ============================
    const uint32_t ew = warp - 2;
    const uint32_t slabs = smem_base + SLABS_OFF;
    for (int it = 0;; ++it) {
      const int slot = it % SCHED_DEPTH;
      mbar_wait(&bars->sched_full[slot], (it / SCHED_DEPTH) & 1);
      const int tile = bars->tile_ring[slot];
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->sched_empty[slot]);
      if (tile >= P.total_tiles) break;

      const TileCoord tc = decode_tile(P, tile);
      const GemmProblem& pr = P.p[tc.p];
      const uint32_t as = it & 1, aphase = (it >> 1) & 1;
      mbar_wait(&bars->tmem_full[as], aphase);
      tc_fence_after();
      epilogue_dispatch(pr, tmem_base + as * BN, tc.m_blk * BM, tc.n_blk * BN, ew, lane, slabs);
      // all tcgen05.ld of this accumulator are complete (tmem_ld_wait) -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tmem_empty[as]);
    }
    if (lane == 0) tma_wait_group<0>();  // smem must outlive the bulk stores; make them complete
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


// ------------------------------------------------------------------------------------------------
// CTA-pair kernel: two CTAs (one cluster, two SMs of a TPC) own one 256 x 256 tile.  Each CTA stages
// its own 128 rows of A and 128 of the 256 B rows, so a k-block costs 32 KB of L2->smem traffic per SM
// instead of 48 KB and 8 KB instead of 12 KB of operand reads per MMA; the leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256) for both.  Accumulator rows 0-127 live in the leader's TMEM,
// rows 128-255 in the peer's; each CTA runs the same epilogue on its half.
// Work distribution is a static, host-computed longest-first list per pair (tile durations are known
// from their k-extent), read by both CTAs, so no cross-CTA scheduling traffic is needed.
// ------------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
tfk_gemm2_kernel(const __grid_constant__ GemmParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  SmemBars2* bars = reinterpret_cast<SmemBars2*>(smem_gen + BARS_OFF);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int* __restrict__ my_list = P.tile_list + static_cast<size_t>(pair) * P.list_stride;
  pdl_launch_dependents();  // the next kernel in the stream may set itself up as soon as SMs free up

  if (warp == 0 && lane == 0) {
    for (int p = 0; p < P.nprob; ++p) {
      tma_prefetch_desc(&P.p[p].tmA[0]);
      tma_prefetch_desc(&P.p[p].tmB[0]);
      tma_prefetch_desc(&P.p[p].tmD[0]);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < STAGES2; ++i) {
        mbar_init(&bars->full[i], 1);   // leader: one arrive.expect_tx; bytes arrive from both CTAs' TMA
        mbar_init(&bars->empty[i], 1);  // multicast tcgen05.commit from the leader
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bars->tmem_full[i], 1);   // multicast tcgen05.commit from the leader
        mbar_init(&bars->tmem_empty[i], 16);  // 8 epilogue warps x 2 CTAs (used in the leader only)
      }
      mbar_init(&bars->a_full, 1);
      mbar_init(&bars->a_empty, 1);
      fence_mbar_init();
    }
    __syncwarp();
  }
  cluster_sync_all();  // barrier inits of both CTAs visible before any remote arrive / TMA signal
  if (warp == 1) tmem_alloc_2sm(&bars->tmem_base, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_wait();  // everything above overlapped the previous kernel's tail; its results are visible from here on

  if (warp == 0 && P.a_resident) {
    // ============================ TMA producer, A-stationary launch ============================
    if (lane == 0) {
      const GemmProblem& pr = P.p[0];
      const uint32_t ring = smem_base + RES_MAX_KB * A_TILE_BYTES;
      const uint32_t tx_b = 2 * 2 * MN_ATOM_BYTES;  // 128 B rows per CTA (two 64-row atoms or one 128-row box), both CTAs
      uint32_t stage = 0, phase = 0;
      int cur_m = -1, run = -1;
      int next_entry = __ldg(my_list);
      for (int it = 0;; ++it) {
        const int entry = next_entry;
        if (entry < 0) break;
        next_entry = __ldg(my_list + it + 1);
        const TileCoord tc = decode_tile(P, entry & kTileMask);
        const int m0 = tc.m_blk * 256 + static_cast<int>(rank) * 128;
        const int nb = tc.n_blk * BN + static_cast<int>(rank) * 128;
        if (tc.m_blk != cur_m) {  // a new run: replace the resident A panel once the MMAs of the previous run have read it
          cur_m = tc.m_blk;
          ++run;
          mbar_wait(&bars->a_empty, (run & 1) ^ 1);
          if (rank == 0) mbar_expect_tx(&bars->a_full, 2 * pr.num_kb * A_TILE_BYTES);
          for (int kb = 0; kb < pr.num_kb; ++kb)
            tma_load_2d_2sm(smem_base + kb * A_TILE_BYTES, &pr.tmA[0], &bars->a_full, kb * BK, m0);
        }
        for (int kb = 0; kb < pr.num_kb; ++kb) {
          mbar_wait(&bars->empty[stage], phase ^ 1);
          if (rank == 0) mbar_expect_tx(&bars->full[stage], tx_b);
          const uint32_t sb = ring + stage * A_TILE_BYTES;
          if (pr.b_mn) {
#pragma unroll
            for (int j = 0; j < 2; ++j) tma_load_2d_2sm(sb + j * MN_ATOM_BYTES, &pr.tmB[0], &bars->full[stage], nb + 64 * j, kb * BK);
          } else {
            tma_load_2d_2sm(sb, &pr.tmB[0], &bars->full[stage], kb * BK, nb);
          }
          if (++stage == RES_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 0) {
    // ============================ TMA producer (both CTAs) ============================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      int next_entry = __ldg(my_list);
      for (int it = 0;; ++it) {
        const int entry = next_entry;
        if (entry < 0) break;
        next_entry = __ldg(my_list + it + 1);  // in flight while this tile is processed (a list ends with -1 inside its row)
        const TileCoord tc = decode_tile(P, entry & kTileMask);
        const GemmProblem& pr = P.p[tc.p];
        const int half = entry >> kHalfShift;  // 0: 256 columns, 1 / 2: left / right 128 columns
        const int m0 = tc.m_blk * 256 + static_cast<int>(rank) * 128;  // this CTA's A / D rows
        // this CTA's half of the tile's B rows (128 of 256, or 64 of 128)
        const int nb = tc.n_blk * BN + (half == 2 ? 128 : 0) + static_cast<int>(rank) * (half ? 64 : 128);
        const int b_atoms = half ? 1 : 2;
        // a K-major B box is always 128 rows (a half tile uses the first 64); MN-major B is loaded per 64-row atom
        const uint32_t tx = 2 * (A_TILE_BYTES + (pr.b_mn ? b_atoms * MN_ATOM_BYTES : 2 * MN_ATOM_BYTES));
        const int loads = pr.nsplit == 3 ? 2 : 1;  // bf16x3: slots (A_hi, B_hi), (A_lo, B_lo) per k-block
        const int iters = tc.kb_count * loads;
        int kb = tc.kb_begin, s = 0;
        for (int i = 0; i < iters; ++i) {
          mbar_wait(&bars->empty[stage], phase ^ 1);
          if (rank == 0) mbar_expect_tx(&bars->full[stage], tx);
          const uint32_t sa = smem_base + stage * STAGE2_BYTES;
          const uint32_t sb = sa + A_TILE_BYTES;
          const CUtensorMap* ta = &pr.tmA[s];
          const CUtensorMap* tb = &pr.tmB[s];
          const int k0 = kb * BK;
          if (pr.a_mn) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
              tma_load_2d_2sm(sa + j * MN_ATOM_BYTES, ta, &bars->full[stage], m0 + 64 * j, k0);
          } else {
            tma_load_2d_2sm(sa, ta, &bars->full[stage], k0, m0);
          }
          if (pr.b_mn) {
            for (int j = 0; j < b_atoms; ++j)
              tma_load_2d_2sm(sb + j * MN_ATOM_BYTES, tb, &bars->full[stage], nb + 64 * j, k0);
          } else {
            tma_load_2d_2sm(sb, tb, &bars->full[stage], k0, nb);
          }
          if (++s == loads) {
            s = 0;
            ++kb;
          }
          if (++stage == STAGES2) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1 && P.a_resident) {
    // ============================ MMA issuer, A-stationary launch (leader CTA only) ============================
    if (lane == 0 && rank == 0) {
      const GemmProblem& pr = P.p[0];
      const uint32_t ring = smem_base + RES_MAX_KB * A_TILE_BYTES;
      const uint32_t idesc = make_idesc_bf16(256, BN, 0, pr.b_mn);
      const uint32_t b_lbo = pr.b_mn ? MN_ATOM_BYTES : 16u, b_kadv = pr.b_mn ? 2048u : 32u;
      uint32_t stage = 0, phase = 0;
      int cur_m = -1, run = -1;
      int next_entry = __ldg(my_list);
      for (int it = 0;; ++it) {
        const int entry = next_entry;
        if (entry < 0) break;
        next_entry = __ldg(my_list + it + 1);
        const TileCoord tc = decode_tile(P, entry & kTileMask);
        if (tc.m_blk != cur_m) {
          cur_m = tc.m_blk;
          ++run;
          mbar_wait(&bars->a_full, run & 1);
          tc_fence_after();
        }
        const uint32_t as = it & 1, aphase = (it >> 1) & 1;
        mbar_wait(&bars->tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + as * BN;
        for (int kb = 0; kb < pr.num_kb; ++kb) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_base + kb * A_TILE_BYTES;
          const uint32_t sb = ring + stage * A_TILE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16_2sm(tmem_acc, make_smem_desc_sw128(sa + k * 32u, 16u, 1024u),
                          make_smem_desc_sw128(sb + k * b_kadv, b_lbo, 1024u), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_2sm(&bars->empty[stage]);
          if (++stage == RES_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm(&bars->tmem_full[as]);
        // last tile of its run: once these MMAs are done the producers may overwrite the A panel
        const bool run_ends = next_entry < 0 || decode_tile(P, next_entry & kTileMask).m_blk != cur_m;
        if (run_ends) umma_commit_2sm(&bars->a_empty);
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer (leader CTA only) ============================
    if (lane == 0 && rank == 0) {
      uint32_t stage = 0, phase = 0;
      int next_entry = __ldg(my_list);
      for (int it = 0;; ++it) {
        const int entry = next_entry;
        if (entry < 0) break;
        next_entry = __ldg(my_list + it + 1);
        const TileCoord tc = decode_tile(P, entry & kTileMask);
        const GemmProblem& pr = P.p[tc.p];
        const uint32_t as = it & 1, aphase = (it >> 1) & 1;
        mbar_wait(&bars->tmem_empty[as], aphase ^ 1);  // both CTAs' epilogues drained this stage
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + as * BN;
        const uint32_t idesc = make_idesc_bf16(256, (entry >> kHalfShift) ? BN / 2 : BN, pr.a_mn, pr.b_mn);
        const uint32_t a_lbo = pr.a_mn ? MN_ATOM_BYTES : 16u, b_lbo = pr.b_mn ? MN_ATOM_BYTES : 16u;
        const uint32_t a_kadv = pr.a_mn ? 2048u : 32u, b_kadv = pr.b_mn ? 2048u : 32u;
        for (int i = 0; i < tc.kb_count; ++i) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE2_BYTES;
          const uint32_t sb = sa + A_TILE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = make_smem_desc_sw128(sa + k * a_kadv, a_lbo, 1024u);
            const uint64_t db = make_smem_desc_sw128(sb + k * b_kadv, b_lbo, 1024u);
            umma_bf16_2sm(tmem_acc, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
          if (pr.nsplit == 3) {  // + A_hi.B_lo + A_lo.B_hi: the low halves of this k-block sit in the next slot
            const uint32_t hi_stage = stage;
            if (++stage == STAGES2) {
              stage = 0;
              phase ^= 1;
            }
            mbar_wait(&bars->full[stage], phase);
            tc_fence_after();
            const uint32_t la = smem_base + stage * STAGE2_BYTES;
            const uint32_t lb = la + A_TILE_BYTES;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_2sm(tmem_acc, make_smem_desc_sw128(sa + k * a_kadv, a_lbo, 1024u),
                            make_smem_desc_sw128(lb + k * b_kadv, b_lbo, 1024u), idesc, 1u);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_2sm(tmem_acc, make_smem_desc_sw128(la + k * a_kadv, a_lbo, 1024u),
                            make_smem_desc_sw128(sb + k * b_kadv, b_lbo, 1024u), idesc, 1u);
            umma_commit_2sm(&bars->empty[hi_stage]);
          }
          umma_commit_2sm(&bars->empty[stage]);  // frees this smem slot in BOTH CTAs
          if (++stage == STAGES2) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm(&bars->tmem_full[as]);  // accumulator halves ready in both CTAs
      }
    }
  } else {
    // ============================ epilogue warps (both CTAs, own 128 rows) ============================
    const uint32_t ew = warp - 2;
    const uint32_t slabs = smem_base + SLABS_OFF;
    // Pull this warp's share of a tile's per-column / per-row epilogue operands (bias, gradient-pass bits) towards L1 one
    // whole tile ahead: their first-touch L2 latency was the largest single stall of the epilogue warps
    // (profiles/r2f_ncu_l0fwd_stalls.txt), and when the epilogue is the pacing stage there is no idle wait to hide it in.
    auto prefetch_for = [&](int entry_) {
      if (entry_ < 0) return;
      const TileCoord t_ = decode_tile(P, entry_ & kTileMask);
      const int half_ = entry_ >> kHalfShift;
      epilogue_prefetch(P.p[t_.p], t_.m_blk * 256 + static_cast<int>(rank) * 128, t_.n_blk * BN + (half_ == 2 ? 128 : 0), ew, lane,
                        half_ ? BN / 64 : BN / 32);
    };
    // list entries are read two tiles ahead so that neither the tile loop nor the prefetch below ever waits for one
    int next_entry = __ldg(my_list);
    int next2_entry = next_entry >= 0 ? __ldg(my_list + 1) : -1;
    prefetch_for(next_entry);
    for (int it = 0;; ++it) {
      const int entry = next_entry;
      if (entry < 0) break;
      next_entry = next2_entry;
      next2_entry = next_entry >= 0 ? __ldg(my_list + it + 2) : -1;  // (a list ends with -1 inside its row)
      const TileCoord tc = decode_tile(P, entry & kTileMask);
      const GemmProblem& pr = P.p[tc.p];
      const int half = entry >> kHalfShift;
      const uint32_t as = it & 1, aphase = (it >> 1) & 1;
      const int m0 = tc.m_blk * 256 + static_cast<int>(rank) * 128, n0 = tc.n_blk * BN + (half == 2 ? 128 : 0);
      prefetch_for(next_entry);
      // forward tiles: this warp's bias values (its half of the tile's columns), in flight while the MMAs finish
      const int chunks = half ? BN / 64 : BN / 32;
      const bool bias_early = P.bias_shfl && pr.epi_variant == EV_FWD_RELU_BITS && pr.out_kind == OUT_BF16 && pr.bias != nullptr;
      float4 bias_pre = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bias_early && static_cast<int>(lane) < chunks * 4)  // chunks/2 chunks x 8 lanes per chunk
        bias_pre = __ldg(reinterpret_cast<const float4*>(pr.bias + n0 + static_cast<int>(ew >> 2) * (chunks / 2) * 32) + lane);
      mbar_wait(&bars->tmem_full[as], aphase);
      tc_fence_after();
      if (m0 < pr.M)  // a ragged last pair-tile may leave the peer CTA without rows (CTA-uniform)
        epilogue_dispatch(pr, tmem_base + as * BN, m0, n0, ew, lane, slabs, chunks, bias_pre, bias_early);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&bars->tmem_empty[as]);
    }
    if (lane == 0) tma_wait_group<0>();
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA may exit (or free TMEM) while its peer can still signal / read it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  // resolved at run time so the library has no link-time dependency on libcuda.so
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) !=
          cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

int make_tmap(CUtensorMap* tm, CUtensorMapDataType dt, int elem_bytes, const void* ptr,
              uint64_t inner, uint64_t outer, uint64_t pitch_elems, uint32_t box_inner,
              uint32_t box_outer, char* err, int errlen) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return -2;
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (pitch_elems * elem_bytes) % 16 != 0) {
    snprintf(err, errlen, "TMA operand %p pitch %llu B is not 16-byte aligned", ptr,
             (unsigned long long)(pitch_elems * elem_bytes));
    return -1;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * static_cast<uint64_t>(elem_bytes)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled failed (%d) inner=%llu outer=%llu pitch=%llu", (int)r,
             (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_elems);
    return -2;
  }
  return 0;
}

}  // namespace

size_t gemm_smem_bytes() { return SMEM_BYTES; }

int gemm_init() {
  static int done = 0;
  if (done) return 0;
  cudaError_t e = cudaFuncSetAttribute(tfk_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(tfk_gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  done = 1;
  return 0;
}

int gemm_build_peer_maps(void* const* peer_D, int num_peers, int M, int N, int ldd, CUtensorMap** d_out, char* err,
                         int errlen) {
  *d_out = nullptr;
  std::vector<CUtensorMap> maps(num_peers);
  for (int o = 0; o < num_peers; ++o) {
    const int rc = make_tmap(&maps[o], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, peer_D[o], N, M, ldd, 32, 32, err, errlen);
    if (rc) return rc;
  }
  CUtensorMap* d = nullptr;
  cudaError_t ce = cudaMalloc(&d, maps.size() * sizeof(CUtensorMap));
  if (ce == cudaSuccess) ce = cudaMemcpy(d, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) {
    snprintf(err, errlen, "peer tensor-map upload failed: %s", cudaGetErrorString(ce));
    if (d) cudaFree(d);
    return -2;
  }
  *d_out = d;
  return 0;
}

int gemm_build_params(const GemmSpec* specs, int nspec, int* sched, GemmParams* out, char* err,
                      int errlen, int two_cta) {
  if (nspec < 1 || nspec > 2) {
    snprintf(err, errlen, "gemm_build_params: nspec must be 1 or 2");
    return -1;
  }
  memset(out, 0, sizeof(GemmParams));
  out->nprob = nspec;
  out->sched = sched;
  out->two_cta = two_cta ? 1 : 0;
  const uint32_t b_box_rows = two_cta ? 128 : BN;  // a CTA of a pair stages half of the 256 B rows
  int tile_begin = 0;
  for (int i = 0; i < nspec; ++i) {
    const GemmSpec& s = specs[i];
    GemmProblem& p = out->p[i];
    if (s.M <= 0 || s.N <= 0 || s.K <= 0) {
      snprintf(err, errlen, "gemm: empty problem M=%d N=%d K=%d", s.M, s.N, s.K);
      return -1;
    }
    if (s.nsplit != 1 && s.nsplit != 3) {
      snprintf(err, errlen, "gemm: nsplit must be 1 or 3");
      return -1;
    }
    int rc;
    for (int h = 0; h < (s.nsplit == 3 ? 2 : 1); ++h) {
      const __nv_bfloat16* a = h ? s.A_lo : s.A_hi;
      const __nv_bfloat16* b = h ? s.B_lo : s.B_hi;
      if (!a || !b) {
        snprintf(err, errlen, "gemm: missing operand pointer");
        return -1;
      }
      if (s.a_mn)
        rc = make_tmap(&p.tmA[h], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a, s.M, s.K, s.lda, 64, BK, err,
                       errlen);
      else
        rc = make_tmap(&p.tmA[h], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a, s.K, s.M, s.lda, BK, BM, err,
                       errlen);
      if (rc) return rc;
      if (s.b_mn)
        rc = make_tmap(&p.tmB[h], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, b, s.N, s.K, s.ldb, 64, BK, err,
                       errlen);
      else
        rc = make_tmap(&p.tmB[h], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, b, s.K, s.N, s.ldb, BK, b_box_rows,
                       err, errlen);
      if (rc) return rc;
    }
    if (s.out_kind == OUT_F32 || s.out_kind == OUT_F32_REDADD) {
      rc = make_tmap(&p.tmD[0], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, s.D_hi, s.N, s.M, s.ldd, 32, 32,
                     err, errlen);
      if (rc) return rc;
    } else {
      rc = make_tmap(&p.tmD[0], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, s.D_hi, s.N, s.M, s.ldd, 64, 32,
                     err, errlen);
      if (rc) return rc;
      if (s.out_kind == OUT_BF16_SPLIT) {
        rc = make_tmap(&p.tmD[1], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, s.D_lo, s.N, s.M, s.ldd, 64, 32,
                       err, errlen);
        if (rc) return rc;
      }
    }
    p.M = s.M;
    p.N = s.N;
    p.K = s.K;
    p.a_mn = s.a_mn;
    p.b_mn = s.b_mn;
    p.nsplit = s.nsplit;
    p.out_kind = s.out_kind;
    p.act = s.act;
    p.deriv = s.deriv;
    p.mask_src_lo = s.mask_src_lo;
    p.dropout_in_chain = s.dropout_in_chain;
    p.mask_ld = s.mask_ld;
    p.mask_nonzero = s.mask_nonzero;
    p.mask_bits_out = s.mask_bits_out;
    p.mask_bits_in = s.mask_bits_in;
    p.mask_bits_ld = s.mask_bits_ld;
    p.scale = s.scale;
    p.keep_inv = 1.0f / s.keep;
    p.drop_thr = (s.keep < 1.0f) ? dropout_threshold(s.keep) : 0u;
    p.bwd_drop_thr = (s.bwd_drop_keep < 1.0f) ? dropout_threshold(s.bwd_drop_keep) : 0u;
    p.bias = s.bias;
    p.mask_src = s.mask_src;
    p.seed = s.seed;
    p.stat_sum = s.stat_sum;
    p.stat_sq = s.stat_sq;
    p.stat_ld = s.stat_ld;
    p.colsum_part = s.colsum_part;
    p.colsum_ld = s.colsum_ld;
    p.bn_z_hi = s.bn_z_hi;
    p.bn_z_lo = s.bn_z_lo;
    p.bn_z_ld = s.bn_z_ld;
    p.bn_mean = s.bn_mean;
    p.bn_rstd = s.bn_rstd;
    p.colsum2_part = s.colsum2_part;
    p.bn_beta = s.bn_beta;
    p.bn_from_y = (s.colsum2_part != nullptr && s.bn_beta != nullptr && s.mask_src != nullptr && s.deriv == 0) ? 1 : 0;
    if (s.colsum2_part != nullptr && s.colsum_part == nullptr) {
      snprintf(err, errlen, "gemm: colsum2_part needs colsum_part");
      return -1;
    }
    if (s.colsum2_part != nullptr && !p.bn_from_y && (s.bn_z_hi == nullptr || s.bn_mean == nullptr || s.bn_rstd == nullptr)) {
      snprintf(err, errlen, "gemm: colsum2_part needs either (bn_beta, mask_src of a relu/linear chain) or (bn_z_hi, bn_mean, bn_rstd)");
      return -1;
    }
    p.peer_tm = nullptr;
    p.num_peers = 0;
    p.rows_per_owner = 0;
    if (s.peer_tm != nullptr && s.num_peers > 1) {
      if (s.out_kind != OUT_F32_REDADD || s.rows_per_owner <= 0 || s.rows_per_owner % 32 != 0) {
        snprintf(err, errlen, "gemm: peer reduce needs OUT_F32_REDADD and rows_per_owner %% 32 == 0");
        return -1;
      }
      p.peer_tm = s.peer_tm;
      p.num_peers = s.num_peers;
      p.rows_per_owner = s.rows_per_owner;
    }
    {
      // features this problem uses -> the leanest epilogue instantiation whose mask covers them
      uint32_t used = 0;
      if (s.bias) used |= F_BIAS;
      if (s.stat_sum) used |= F_STATS;
      if (s.act == 1) used |= F_RELU;
      if (s.act >= 2) used |= F_ACT_SMOOTH;
      if (s.keep < 1.0f) used |= F_DROPOUT;
      if (s.mask_bits_out) used |= F_BITS_OUT;
      if (s.mask_bits_in) used |= F_BITS_IN;
      if (s.mask_src && s.deriv != 0) used |= F_MASK_DERIV;
      if (s.mask_src && s.deriv == 0) used |= F_MASK_RELU;
      if (s.colsum_part) used |= F_COLSUM;
      if (s.colsum2_part) used |= F_COLSUM2;
      if (s.bwd_drop_keep < 1.0f) used |= F_BWD_PHILOX;
      p.epi_variant = EV_GENERIC;
      static const char* force_generic = getenv("TFK_GEMM_GENERIC_EPILOGUE");  // tests: every launch through the generic body
      if (!(force_generic && force_generic[0] == '1')) {
        const int candidates[] = {EV_PLAIN, EV_FWD_RELU_BITS, EV_FWD_STATS, EV_DGRAD_BITS_COLSUM, EV_DGRAD_BN};
        for (int v : candidates) {
          const bool out_ok = v == EV_PLAIN ? (s.out_kind == OUT_F32 || (s.out_kind == OUT_F32_REDADD && used == 0))
                                            : s.out_kind == OUT_BF16;
          if (out_ok && (used & ~kFeatOf[v]) == 0) {
            p.epi_variant = v;
            break;
          }
        }
      }
    }
    p.tiles_m = two_cta ? (s.M + 255) / 256 : (s.M + BM - 1) / BM;
    p.tiles_n = (s.N + BN - 1) / BN;
    p.tile_begin = tile_begin;
    p.num_kb = (s.K + BK - 1) / BK;
    p.ksplit = 1;
    if (s.ksplit > 1) {
      if (s.out_kind != OUT_F32_REDADD) {
        snprintf(err, errlen, "gemm: ksplit needs OUT_F32_REDADD");
        return -1;
      }
      p.ksplit = s.ksplit < p.num_kb ? s.ksplit : p.num_kb;
    }
    p.kb_per_split = (p.num_kb + p.ksplit - 1) / p.ksplit;
    p.ksplit = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;  // no empty splits
    tile_begin += p.tiles_m * p.tiles_n * p.ksplit;
  }
  out->total_tiles = tile_begin;
  static const char* no_shfl = getenv("TFK_GEMM_BIAS_SHFL");  // "0": per-chunk bias loads (A/B measurements)
  out->bias_shfl = (no_shfl && no_shfl[0] == '0') ? 0 : 1;
  out->a_resident = 0;
  if (two_cta && nspec == 1) {
    const GemmSpec& s0 = specs[0];
    const GemmProblem& p0 = out->p[0];
    // Measured on the layer-0 forward (8192 x 440 -> 2048, profiles/r2i_selftest_l0.txt): 21.9 us A-stationary against
    // 20.1 us with the ordinary ring + half-tile balancing — that launch is paced by its epilogue (stall summaries
    // profiles/r2h_ncu_l0fwd_stalls.txt), not by operand delivery, so halving the operand traffic buys nothing while the
    // contiguous tile ranges cost the half-tile load balancing.  Opt-in (TFK_GEMM_A_RESIDENT=1) until the epilogue is faster.
    static const char* on = getenv("TFK_GEMM_A_RESIDENT");
    if ((on && on[0] == '1') && !s0.a_mn && s0.nsplit == 1 && p0.ksplit == 1 && p0.num_kb <= RES_MAX_KB && s0.N % BN == 0 &&
        p0.tiles_n >= 2)
      out->a_resident = 1;
  }
  return 0;
}

int gemm_launch(const GemmParams& params, int num_sms, cudaStream_t stream) {
  int rc = gemm_init();
  if (rc) return rc;
  if (params.two_cta) {
    if (params.tile_list == nullptr || params.num_pairs < 1) return (int)cudaErrorInvalidValue;
    static const bool pdl = [] {
      const char* e = getenv("TFK_PDL");
      return !(e && e[0] == '0');
    }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * params.num_pairs, 1, 1);
    cfg.blockDim = dim3(GEMM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return (int)cudaLaunchKernelEx(&cfg, tfk_gemm2_kernel, params);
  }
  const int grid = params.total_tiles < num_sms ? params.total_tiles : num_sms;
  tfk_gemm_kernel<<<grid, GEMM_THREADS, SMEM_BYTES, stream>>>(params);
  return (int)cudaGetLastError();
}

// Host-only half of gemm_upload_tile_lists: the per-pair work lists as a flat [pairs, stride] array (-1 padded).
void gemm_schedule_tile_lists(const GemmParams* params, int num_sms, std::vector<int>* flat_out, int* pairs_out,
                              int* stride_out) {
  const int total = params->total_tiles;
  int pairs = num_sms / 2;
  if (pairs > 2 * total) pairs = 2 * total;
  if (pairs < 1) pairs = 1;
  if (params->a_resident) {
    // A-stationary launch: tile index = m_blk * tiles_n + n_blk, so CONSECUTIVE indices share their A panel; every pair
    // gets one contiguous range (sizes differ by at most one tile), whole tiles only
    if (pairs > total) pairs = total;
    const int base = total / pairs, rem = total % pairs;
    const int stride = base + 2;
    std::vector<int> flat(static_cast<size_t>(pairs) * stride, -1);
    int t = 0;
    for (int p = 0; p < pairs; ++p) {
      const int n = base + (p < rem ? 1 : 0);
      for (int i = 0; i < n; ++i) flat[static_cast<size_t>(p) * stride + i] = t++;
    }
    *pairs_out = pairs;
    *stride_out = stride;
    flat_out->swap(flat);
    return;
  }
  // Cost model, in half-k-block units: k-extent (MMA time) + a constant for the epilogue.  A tile may be scheduled
  // as two 128-column halves (M256 N128 MMAs): used for tiles whose right half lies outside N, to shorten a partly
  // filled last round (256 equal tiles on 74 pairs: 3 + 0.8 rounds instead of 4) and for launches with few tiles.
  struct Item {
    long long cost;
    int entry;
  };
  const char* force = getenv("TFK_GEMM_HALF_TILES");  // "all": every tile as halves (tests); "0": never;
                                                      // "auto": halves cost half, any modelled gain is taken
  const bool optimistic = force && strcmp(force, "auto") == 0;
  std::vector<Item> full;   // tiles that may be scheduled whole or as two halves
  std::vector<Item> fixed;  // tiles whose right 128 columns lie outside N: always a single left half
  auto tile_cost = [&](int t, bool half) -> long long {
    const int pi = (params->nprob > 1 && t >= params->p[1].tile_begin) ? 1 : 0;
    const GemmProblem& pr = params->p[pi];
    const int local = t - pr.tile_begin;
    const int split = local / (pr.tiles_m * pr.tiles_n);
    const int kb0 = split * pr.kb_per_split;
    const int cnt = pr.kb_per_split < pr.num_kb - kb0 ? pr.kb_per_split : pr.num_kb - kb0;
    const long long mma = static_cast<long long>(cnt) * pr.nsplit;
    // measured (profiles/r1_halftile_policies.txt): a half tile takes 0.8 of a whole one, not 0.5 - the mainloop is
    // paced by operand delivery from L2 (~70 GB/s per SM) and a half tile still stages all of A
    if (optimistic) return half ? mma + 7 : 2 * mma + 12;
    return half ? (8 * (2 * mma + 12)) / 10 : 2 * mma + 12;
  };
  for (int t = 0; t < total; ++t) {
    const int pi = (params->nprob > 1 && t >= params->p[1].tile_begin) ? 1 : 0;
    const GemmProblem& pr = params->p[pi];
    const int n_blk = (t - pr.tile_begin) % pr.tiles_n;
    if (n_blk * BN + BN / 2 >= pr.N)
      fixed.push_back({tile_cost(t, true), t | (1 << kHalfShift)});
    else
      full.push_back({tile_cost(t, false), t});
  }
  auto by_cost = [](const Item& a, const Item& b) { return a.cost > b.cost; };
  std::stable_sort(full.begin(), full.end(), by_cost);
  // longest-processing-time-first assignment of `items`; returns the makespan
  auto lpt = [&](std::vector<Item> items, std::vector<std::vector<int>>* lists) -> long long {
    std::stable_sort(items.begin(), items.end(), by_cost);
    if (lists) lists->assign(pairs, std::vector<int>());
    // every item goes to the least loaded pair, the lowest-numbered one among equals: a min-heap on (load, pair)
    typedef std::pair<long long, int> Slot;
    std::vector<Slot> heap(pairs);
    for (int p = 0; p < pairs; ++p) heap[p] = Slot(0, p);  // ascending (load, pair): already a valid min-heap
    auto later = [](const Slot& a, const Slot& b) { return a > b; };
    long long mk = 0;
    for (const auto& it : items) {
      std::pop_heap(heap.begin(), heap.end(), later);
      Slot& best = heap.back();
      if (lists) (*lists)[best.second].push_back(it.entry);
      best.first += it.cost;
      mk = best.first > mk ? best.first : mk;
      std::push_heap(heap.begin(), heap.end(), later);
    }
    return mk;
  };
  auto with_split = [&](int nsplit_tiles) {  // the `nsplit_tiles` cheapest whole tiles become two halves each
    std::vector<Item> items(fixed);
    const int nfull = static_cast<int>(full.size());
    for (int i = 0; i < nfull; ++i) {
      const int t = full[i].entry;
      if (i >= nfull - nsplit_tiles) {
        const long long c = tile_cost(t, true);
        items.push_back({c, t | (1 << kHalfShift)});
        items.push_back({c, t | (2 << kHalfShift)});
      } else {
        items.push_back(full[i]);
      }
    }
    return items;
  };
  int best_split = 0;
  {
    const int nfull = static_cast<int>(full.size());
    if (force && strcmp(force, "all") == 0) {
      best_split = nfull;
    } else if (!(force && strcmp(force, "0") == 0)) {
      const long long whole = lpt(with_split(0), nullptr);
      long long best = optimistic ? whole : whole - whole / 50;  // default: only for a modelled gain of 2 % or more
      const int limit = nfull < 2 * pairs ? nfull : 2 * pairs;
      for (int sp = 1; sp <= limit; ++sp) {
        const long long mk = lpt(with_split(sp), nullptr);
        if (mk < best) {
          best = mk;
          best_split = sp;
        }
      }
    }
  }
  std::vector<std::vector<int>> lists;
  const long long makespan = lpt(with_split(best_split), &lists);
  if (getenv("TFK_GEMM_DEBUG"))
    fprintf(stderr, "[tfk gemm] %d tiles (%zu left-half only), %d pairs: %d tiles split in two, modelled makespan %lld\n",
            total, fixed.size(), pairs, best_split, makespan);
  if (static_cast<int>(fixed.size()) + static_cast<int>(full.size()) + best_split < pairs) {
    // fewer work items than pairs: launch only the pairs that got one
    int used = 0;
    for (const auto& l : lists) used += l.empty() ? 0 : 1;
    std::stable_sort(lists.begin(), lists.end(),
                     [](const std::vector<int>& a, const std::vector<int>& b) { return a.size() > b.size(); });
    pairs = used < 1 ? 1 : used;
    lists.resize(pairs);
  }
  size_t stride = 1;
  for (const auto& l : lists) stride = l.size() + 1 > stride ? l.size() + 1 : stride;
  std::vector<int> flat(static_cast<size_t>(pairs) * stride, -1);
  for (int p = 0; p < pairs; ++p)
    for (size_t i = 0; i < lists[p].size(); ++i) flat[p * stride + i] = lists[p][i];
  *pairs_out = pairs;
  *stride_out = static_cast<int>(stride);
  flat_out->swap(flat);
}

int gemm_upload_tile_lists(GemmParams* params, int num_sms, int** d_list, char* err, int errlen) {
  *d_list = nullptr;
  if (!params->two_cta) return 0;
  std::vector<int> flat;
  int pairs = 0, stride = 0;
  gemm_schedule_tile_lists(params, num_sms, &flat, &pairs, &stride);
  int* d = nullptr;
  cudaError_t e = cudaMalloc(&d, flat.size() * sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpy(d, flat.data(), flat.size() * sizeof(int), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    snprintf(err, errlen, "tile list upload failed: %s", cudaGetErrorString(e));
    if (d) cudaFree(d);
    return -2;
  }
  params->num_pairs = pairs;
  params->list_stride = stride;
  params->tile_list = d;
  *d_list = d;
  return 0;
}

}  // namespace tfk
