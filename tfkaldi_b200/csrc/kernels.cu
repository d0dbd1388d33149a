// HBM-bound kernels of the hot path: coalesced, 128-bit vectorised accesses, one pass where possible.
#include "kernels.cuh"

#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "philox.cuh"

namespace tfk {

namespace {

// Programmatic dependent launch for the small kernels between the GEMMs of a step: OPT-IN (TFK_PDL_SMALL=1).  A kernel
// launched through launch_pdl() with it may be scheduled while its predecessor in the stream is still running (the GEMM
// kernels release their dependents at their very start); it calls pdl_enter() before touching anything the predecessor
// wrote, so what overlaps is its launch latency.  Measured (one box, two repetitions, profiles/r2k_ab_pdl_small.txt): the
// C4 step got SLOWER, 1.056 -> 1.144 ms, the C2 step 1-2 % slower — early-resident blocks of the small kernels sit on the
// SMs the GEMM's last tiles still need and delay the next GEMM's CTAs — so the default is the plain launch.
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args... args) {
  static const bool pdl = [] {
    const char* e = getenv("TFK_PDL_SMALL");
    return e && e[0] == '1';
  }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
  __nv_bfloat162 h;
  h.x = ah;
  h.y = bh;
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = pack2(a - __bfloat162float(ah), b - __bfloat162float(bh));
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// ------------------------------------------------------------------------------------------------
__global__ void split_f32_kernel(const float* __restrict__ src, int ld_src, __nv_bfloat16* __restrict__ hi,
                                 __nv_bfloat16* __restrict__ lo, int ld_dst, int rows, int cols,
                                 int vec_ok) {
  pdl_enter();
  const int groups = ld_dst >> 2;
  const size_t total = static_cast<size_t>(rows) * groups;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / groups);
    const int c = static_cast<int>(i - static_cast<size_t>(r) * groups) << 2;
    float x[4];
    const float* sp = src + static_cast<size_t>(r) * ld_src + c;
    if (vec_ok && c + 4 <= cols) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(sp));
      x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) x[k] = (c + k < cols) ? __ldg(sp + k) : 0.f;
    }
    uint2 h, l;
    split2(x[0], x[1], h.x, l.x);
    split2(x[2], x[3], h.y, l.y);
    const size_t o = static_cast<size_t>(r) * ld_dst + c;
    *reinterpret_cast<uint2*>(hi + o) = h;
    if (lo) *reinterpret_cast<uint2*>(lo + o) = l;
  }
}

// one warp per output row: find the row's utterance (binary search, warp-uniform), then lanes stride over
// the (2k+1)*D output columns; neighbouring rows re-read the same raw frames from L1/L2.
__global__ void __launch_bounds__(256)
splice_cmvn_kernel(const float* __restrict__ raw, const int32_t* __restrict__ utt_off, int num_utts,
                   const float* __restrict__ cmvn, int D, int k, int row_begin, int rows,
                   __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int ld_dst) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= rows) return;
  const int r = row_begin + i;
  int a = 0, b = num_utts;  // utt_off[a] <= r < utt_off[b]
  while (b - a > 1) {
    const int m = (a + b) >> 1;
    if (__ldg(utt_off + m) <= r) a = m; else b = m;
  }
  const int u0 = __ldg(utt_off + a), u1 = __ldg(utt_off + a + 1);
  const float* __restrict__ mean = cmvn + static_cast<size_t>(a) * 2 * D;
  const float* __restrict__ istd = mean + D;
  const int width = (2 * k + 1) * D;
  const size_t o = static_cast<size_t>(i) * ld_dst;
  for (int c = lane; c < ld_dst; c += 32) {
    float v = 0.f;
    if (c < width) {
      const int j = c / D, d = c - j * D;
      const int src = r + j - k;
      if (src >= u0 && src < u1) v = (__ldg(raw + static_cast<size_t>(src) * D + d) - __ldg(mean + d)) * __ldg(istd + d);
    }
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[o + c] = h;
    if (lo) lo[o + c] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

__global__ void merge_bf16_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                  int ld_src, float* __restrict__ dst, int ld_dst, int rows, int cols) {
  const size_t total = static_cast<size_t>(rows) * cols;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / cols);
    const int c = static_cast<int>(i - static_cast<size_t>(r) * cols);
    float v = __bfloat162float(hi[static_cast<size_t>(r) * ld_src + c]);
    if (lo) v += __bfloat162float(lo[static_cast<size_t>(r) * ld_src + c]);
    dst[static_cast<size_t>(r) * ld_dst + c] = v;
  }
}

__global__ void mask_scale_f32_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                      float* __restrict__ out, size_t n, float scale, int mode) {
  // mode 0: pass where y > 0; 1: pass where y != 0; 2: sigmoid; 3: tanh; +4: dropout in the chain (y == 0 => dropped)
  const int kind = mode & 3;
  const bool drop = (mode & 4) != 0;
  const float keep = 1.0f / scale;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float a = y[i];
    float f;
    if (kind == 0) f = a > 0.f ? scale : 0.f;
    else if (kind == 1) f = a != 0.f ? scale : 0.f;
    else {
      const float yy = a * keep;
      const float d = kind == 2 ? yy * (1.0f - yy) : 1.0f - yy * yy;
      f = (drop && a == 0.f) ? 0.f : d * scale;
    }
    out[i] = dy[i] * f;
  }
}
__global__ void double_to_float_kernel(const double* __restrict__ src, float* __restrict__ dst) {
  dst[0] = static_cast<float>(src[0]);
}

// ------------------------------------------------------------------------------------------------
// One warp per row, two sweeps: (1) running max and sum of exp in ONE pass over the logits (online softmax), (2) the
// row is read again — it was just brought into L1/L2 — to emit loss and gradient.  Nothing but a float4 lives in
// registers between loads (~32 registers per thread; the round-1 kernels kept the whole row in registers: 80 registers at
// 1936 pdf-ids, 174 at 3401 with 12 % occupancy), so six to eight blocks are resident per SM and enough loads are in
// flight to stream the logits at HBM speed.
__global__ void __launch_bounds__(256)
softmax_ce_stream_kernel(const float* __restrict__ logits, int ld, const int32_t* __restrict__ labels, int B,
                         int O, float* __restrict__ row_loss, __nv_bfloat16* __restrict__ d_hi,
                         __nv_bfloat16* __restrict__ d_lo) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const int nvec = ld >> 2;
  const float4* zp = reinterpret_cast<const float4*>(logits + static_cast<size_t>(row) * ld);
  float mx = -INFINITY, sum = 0.f;
  for (int idx0 = lane; idx0 < nvec; idx0 += 128) {  // four independent 16-byte loads in flight per lane
    float4 t[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = idx0 + 32 * u;
      t[u] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      if (idx < nvec) {
        t[u] = __ldg(zp + idx);
        const int e = idx << 2;
        if (e + 1 > O) t[u].x = -INFINITY;
        if (e + 2 > O) t[u].y = -INFINITY;
        if (e + 3 > O) t[u].z = -INFINITY;
        if (e + 4 > O) t[u].w = -INFINITY;
      }
    }
    float m4 = mx;
#pragma unroll
    for (int u = 0; u < 4; ++u) m4 = fmaxf(m4, fmaxf(fmaxf(t[u].x, t[u].y), fmaxf(t[u].z, t[u].w)));
    if (m4 > -INFINITY) {  // (a lane whose elements so far are all padding keeps sum = 0)
      float part = 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        part += (__expf(t[u].x - m4) + __expf(t[u].y - m4)) + (__expf(t[u].z - m4) + __expf(t[u].w - m4));
      sum = sum * __expf(mx - m4) + part;  // exp(-inf) = 0 on the first block
      mx = m4;
    }
  }
  const float gmx = warp_max(mx);
  sum = warp_sum(mx > -INFINITY ? sum * __expf(mx - gmx) : 0.f);
  const int label = labels[row];
  const bool label_ok = label >= 0 && label < O;
  if (lane == 0) row_loss[row] = label_ok ? (logf(sum) + gmx - __ldg(logits + static_cast<size_t>(row) * ld + label)) : 0.f;
  if (d_hi == nullptr) return;
  const float inv = 1.0f / sum;
  for (int idx0 = lane; idx0 < nvec; idx0 += 128) {
    float4 t[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (idx0 + 32 * u < nvec) t[u] = __ldg(zp + idx0 + 32 * u);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = idx0 + 32 * u;
      if (idx >= nvec) continue;
      const int e = idx << 2;
      float p[4] = {e + 0 < O ? __expf(t[u].x - gmx) * inv : 0.f, e + 1 < O ? __expf(t[u].y - gmx) * inv : 0.f,
                    e + 2 < O ? __expf(t[u].z - gmx) * inv : 0.f, e + 3 < O ? __expf(t[u].w - gmx) * inv : 0.f};
      // softmax - onehot; an out-of-range label has an all-zero one-hot row (tf.one_hot): no loss term, but the op's
      // gradient is still prob - labels = softmax (SoftmaxCrossEntropyWithLogits computes backprop = prob - labels)
      if (label_ok && (label >> 2) == idx) p[label & 3] -= 1.0f;
      uint2 h, l;
      split2(p[0], p[1], h.x, l.x);
      split2(p[2], p[3], h.y, l.y);
      const size_t o = static_cast<size_t>(row) * ld + e;
      *reinterpret_cast<uint2*>(d_hi + o) = h;
      if (d_lo) *reinterpret_cast<uint2*>(d_lo + o) = l;
    }
  }
}

__global__ void __launch_bounds__(1024) accum_loss_kernel(const float* __restrict__ row_loss, int B,
                                                          double* __restrict__ acc) {
  pdl_enter();
  __shared__ double sm[1024];
  double s = 0.0;
  for (int i = threadIdx.x; i < B; i += 1024) s += static_cast<double>(row_loss[i]);
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (static_cast<int>(threadIdx.x) < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    acc[0] += sm[0];
    acc[1] += static_cast<double>(B);
  }
}

// ------------------------------------------------------------------------------------------------
constexpr int COLSUM2_RS = 64;  // row splits of colsum_kernel
// Single launch, deterministic.  A warp reads 256 consecutive columns (16 B per lane) of one row per
// load; a block's 8 warps stride over its row range; the LAST block to finish a 256-column group adds
// the COLSUM2_RS partials in fixed order into out[] (threadfence + self-resetting counter).
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, int ld, int rows,
              int cols, float* __restrict__ ws, unsigned int* __restrict__ counters, float* __restrict__ out) {
  __shared__ float sm[8][256];
  __shared__ unsigned int last;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  const int rs = blockIdx.y;
  const int per = (rows + COLSUM2_RS - 1) / COLSUM2_RS;
  const int r0 = rs * per, r1 = min(rows, r0 + per);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < ld) {
#pragma unroll 4
    for (int r = r0 + w; r < r1; r += 8) {
      const size_t o = static_cast<size_t>(r) * ld + col;
      const uint4 h = __ldg(reinterpret_cast<const uint4*>(hi + o));
      acc[0] += bf_lo(h.x); acc[1] += bf_hi(h.x); acc[2] += bf_lo(h.y); acc[3] += bf_hi(h.y);
      acc[4] += bf_lo(h.z); acc[5] += bf_hi(h.z); acc[6] += bf_lo(h.w); acc[7] += bf_hi(h.w);
      if (lo) {
        const uint4 l = __ldg(reinterpret_cast<const uint4*>(lo + o));
        acc[0] += bf_lo(l.x); acc[1] += bf_hi(l.x); acc[2] += bf_lo(l.y); acc[3] += bf_hi(l.y);
        acc[4] += bf_lo(l.z); acc[5] += bf_hi(l.z); acc[6] += bf_lo(l.w); acc[7] += bf_hi(l.w);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) sm[w][lane * 8 + k] = acc[k];
  __syncthreads();
  const int t = threadIdx.x;
  {
    float s = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) s += sm[y][t];
    if (blockIdx.x * 256 + t < ld) ws[static_cast<size_t>(rs) * ld + blockIdx.x * 256 + t] = s;
  }
  __threadfence();
  __syncthreads();
  if (t == 0) last = atomicAdd(&counters[blockIdx.x], 1u);
  __syncthreads();
  if (last != COLSUM2_RS - 1) return;
  __threadfence();
  const int c = blockIdx.x * 256 + t;
  if (c < cols) {
    float part[COLSUM2_RS];
#pragma unroll
    for (int r = 0; r < COLSUM2_RS; ++r) part[r] = __ldcg(ws + static_cast<size_t>(r) * ld + c);  // all in flight
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < COLSUM2_RS; ++r) s += part[r];  // fixed order
    out[c] += s;
  }
  if (t == 0) counters[blockIdx.x] = 0u;
}

// out_j[c] += sum over the 32-row groups of part_j[g][c]: finishes the column sums the GEMM epilogue
// started (bias gradients), for up to 16 layers in one launch (blockIdx.y = job).
struct ColsumJobs {
  const float* part[16];
  float* out[16];
  int groups, ld, cols, njobs;
};
__global__ void __launch_bounds__(256) colsum_finalize_kernel(const ColsumJobs J) {
  pdl_enter();
  __shared__ float sm[8][32];
  const float* __restrict__ part = J.part[blockIdx.y];
  const int cl = threadIdx.x & 31, gl = threadIdx.x >> 5;  // 32 columns x 8 group lanes
  const int c = blockIdx.x * 32 + cl;
  float s = 0.f;
  if (c < J.cols) {
    int g = gl;
    for (; g + 56 < J.groups; g += 64) {  // 8 independent loads in flight
      float t[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) t[k] = part[static_cast<size_t>(g + 8 * k) * J.ld + c];
#pragma unroll
      for (int k = 0; k < 8; ++k) s += t[k];
    }
    for (; g < J.groups; g += 8) s += part[static_cast<size_t>(g) * J.ld + c];
  }
  sm[gl][cl] = s;
  __syncthreads();
  if (gl == 0 && c < J.cols) {
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot += sm[k][cl];
    J.out[blockIdx.y][c] += tot;
  }
}

// ------------------------------------------------------------------------------------------------
struct AdamSegs {
  unsigned long long off4[64];  // segment start, in float4 units
  unsigned long long cnt4[64];  // segment length, in float4 units
  int n;
  // fused update -> all-gather: the refreshed bf16 operand copies of the first `n_bcast` segments are also
  // stored into every peer GPU's shadow arena over NVLink (peer_hi[r] / peer_lo[r], r != self)
  int n_bcast, n_peers;
  uint4* peer_hi[15];
  uint4* peer_lo[15];
};
// blockIdx.y = segment (the whole arena on one GPU; a rank's slice of every layer + the small
// replicated region under sharded data parallelism)
// 128-thread blocks at <= 64 registers: one such block fits on an SM NEXT TO a resident GEMM CTA (320 threads x 168
// registers leave 11.7 K of the 64 K registers; the GEMM takes all the shared memory, this kernel uses none), so the
// per-layer updates launched on the side stream really run under the backward kernels instead of in the gaps between them.
__global__ void __launch_bounds__(128)
adam_kernel(float4* __restrict__ w, float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
            uint4* __restrict__ w_hi, uint4* __restrict__ w_lo, const AdamSegs segs, const double* __restrict__ acc,
            float lr_t, float b1, float b2, float eps) {
  const float frames = static_cast<float>(acc[1]);
  const float omb1 = 1.0f - b1, omb2 = 1.0f - b2;
  // eight parameters per thread and iteration: eight independent 16-byte loads in flight, and the refreshed bf16
  // operands leave as ONE 16-byte store per destination (the peer copies travel over NVLink: half as many packets as
  // with 8-byte stores)
  const size_t base = segs.off4[blockIdx.y] >> 1, n8 = segs.cnt4[blockIdx.y] >> 1;
  for (size_t k = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; k < n8;
       k += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t i = base + k;
    const float4 g0 = g[2 * i], g1 = g[2 * i + 1], m0 = m[2 * i], m1 = m[2 * i + 1];
    const float4 v0 = v[2 * i], v1 = v[2 * i + 1], w0 = w[2 * i], w1 = w[2 * i + 1];
    float gx[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    float mx[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
    float vx[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    float wx[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) {
      float gh = gx[k2] / frames;                      // tf.div(grad, num_frames)   trainer.py:174
      gh = fminf(fmaxf(gh, -1.0f), 1.0f);              // tf.clip_by_value(-1, 1)    trainer.py:178
      mx[k2] += (gh - mx[k2]) * omb1;                  // TF ApplyAdam
      vx[k2] += (gh * gh - vx[k2]) * omb2;
      wx[k2] -= (mx[k2] * lr_t) / (sqrtf(vx[k2]) + eps);
    }
    m[2 * i] = make_float4(mx[0], mx[1], mx[2], mx[3]);
    m[2 * i + 1] = make_float4(mx[4], mx[5], mx[6], mx[7]);
    v[2 * i] = make_float4(vx[0], vx[1], vx[2], vx[3]);
    v[2 * i + 1] = make_float4(vx[4], vx[5], vx[6], vx[7]);
    w[2 * i] = make_float4(wx[0], wx[1], wx[2], wx[3]);
    w[2 * i + 1] = make_float4(wx[4], wx[5], wx[6], wx[7]);
    g[2 * i] = make_float4(0.f, 0.f, 0.f, 0.f);        // init_grads                 trainer.py:350
    g[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    uint4 h, l;
    split2(wx[0], wx[1], h.x, l.x);
    split2(wx[2], wx[3], h.y, l.y);
    split2(wx[4], wx[5], h.z, l.z);
    split2(wx[6], wx[7], h.w, l.w);
    w_hi[i] = h;
    if (w_lo) w_lo[i] = l;
    if (static_cast<int>(blockIdx.y) < segs.n_bcast) {
#pragma unroll 1
      for (int r = 0; r < segs.n_peers; ++r) {
        segs.peer_hi[r][i] = h;
        if (w_lo) segs.peer_lo[r][i] = l;
      }
    }
  }
  if (segs.n_bcast > 0) __threadfence_system();  // remote stores visible before the kernel is seen as done
}

// publish / wait: flag words [slot][source rank] in every GPU's flag array (peer-mapped); slot 0 = "my operand stores
// of this step have landed", slot 1 + l = "my backward kernel of layer l has finished (its remote reduce-adds are done)"
constexpr int kFlagStride = 16;  // ranks per slot
__global__ void dp_publish_kernel(int* const* __restrict__ peer_flags, int n_peers, int slot, int me, int value) {
  const int r = threadIdx.x;
  if (r < n_peers) {
    __threadfence_system();
    *(reinterpret_cast<volatile int*>(peer_flags[r]) + slot * kFlagStride + me) = value;
  }
}
__global__ void dp_wait_kernel(const int* __restrict__ flags, int n_ranks, int slot, int me, int value) {
  const int r = threadIdx.x;
  if (r < n_ranks && r != me) {
    const volatile int* f = flags + slot * kFlagStride + r;
    const long long t0 = clock64();
    while (*f < value) {
      if (clock64() - t0 > 6000000000LL) {
        printf("tfk: timeout waiting for rank %d to publish step %d in slot %d (have %d)\n", r, value, slot, *f);
        __trap();
      }
    }
  }
  __threadfence_system();
}

__device__ __forceinline__ void load8(const __nv_bfloat16* hi, const __nv_bfloat16* lo, size_t o, float (&x)[8]) {
  const uint4 h = __ldg(reinterpret_cast<const uint4*>(hi + o));
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    x[2 * k] = bf_lo(hw[k]);
    x[2 * k + 1] = bf_hi(hw[k]);
  }
  if (lo) {
    const uint4 l = __ldg(reinterpret_cast<const uint4*>(lo + o));
    const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      x[2 * k] += bf_lo(lw[k]);
      x[2 * k + 1] += bf_hi(lw[k]);
    }
  }
}
__device__ __forceinline__ void store8(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t o, const float (&x)[8]) {
  uint4 h, l;
  split2(x[0], x[1], h.x, l.x);
  split2(x[2], x[3], h.y, l.y);
  split2(x[4], x[5], h.z, l.z);
  split2(x[6], x[7], h.w, l.w);
  *reinterpret_cast<uint4*>(hi + o) = h;
  if (lo) *reinterpret_cast<uint4*>(lo + o) = l;
}

// Raw 16-byte row fragments (bf16 hi [+ lo]) are kept in flight as loaded; converted only when consumed.
__device__ __forceinline__ void unpack8(const uint4& h, const uint4& l, bool has_lo, float (&x)[8]) {
  x[0] = bf_lo(h.x); x[1] = bf_hi(h.x); x[2] = bf_lo(h.y); x[3] = bf_hi(h.y);
  x[4] = bf_lo(h.z); x[5] = bf_hi(h.z); x[6] = bf_lo(h.w); x[7] = bf_hi(h.w);
  if (has_lo) {
    x[0] += bf_lo(l.x); x[1] += bf_hi(l.x); x[2] += bf_lo(l.y); x[3] += bf_hi(l.y);
    x[4] += bf_lo(l.z); x[5] += bf_hi(l.z); x[6] += bf_lo(l.w); x[7] += bf_hi(l.w);
  }
}

constexpr int BN_SPAN = 2048;  // columns per block of bn_bwd_apply_kernel (256 threads x 8)

// BN backward column reductions sum_B(dy) and sum_B(dy * xhat), one launch, deterministic.  ws layout [RS][2][ld].
// A warp reads 256 consecutive columns (16 B per lane per array) of one row per load; a block's 8 warps stride over
// its row range; the LAST block to finish a 256-column group adds the BN_RS partials in fixed order
// (threadfence + self-resetting counter) into sums[] and the beta gradient.  grid = (ld/256, 74): one wave.
constexpr int BN_RS = 74;
template <bool X3>
__global__ void __launch_bounds__(256, 4)
bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dy_hi, const __nv_bfloat16* __restrict__ dy_lo,
                     const __nv_bfloat16* __restrict__ z_hi, const __nv_bfloat16* __restrict__ z_lo, int ld,
                     int B, int N, const float* __restrict__ mean, const float* __restrict__ rstd,
                     float* __restrict__ ws, unsigned int* __restrict__ counters, float* __restrict__ sums,
                     float* __restrict__ g_beta) {
  __shared__ float sm[8][256];
  __shared__ unsigned int last;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  const int rs = blockIdx.y;
  const int per = (B + BN_RS - 1) / BN_RS;
  const int r0 = rs * per, r1 = min(B, r0 + per);
  constexpr bool has_lo = X3;
  float a[8], b[8], mu[8], rsd[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    a[k] = 0.f;
    b[k] = 0.f;
    mu[k] = (col + k < N) ? mean[col + k] : 0.f;
    rsd[k] = (col + k < N) ? rstd[col + k] : 0.f;
  }
  if (col < ld) {
    for (int r = r0 + w; r < r1; r += 16) {  // two rows x two arrays in flight
      uint4 dh[2], dl[2], zh[2], zl[2];
#pragma unroll
      for (int u = 0; u < 2; ++u)
        if (r + 8 * u < r1) {
          const size_t o = static_cast<size_t>(r + 8 * u) * ld + col;
          dh[u] = __ldg(reinterpret_cast<const uint4*>(dy_hi + o));
          zh[u] = __ldg(reinterpret_cast<const uint4*>(z_hi + o));
          if (has_lo) {
            dl[u] = __ldg(reinterpret_cast<const uint4*>(dy_lo + o));
            zl[u] = __ldg(reinterpret_cast<const uint4*>(z_lo + o));
          }
        }
#pragma unroll
      for (int u = 0; u < 2; ++u)
        if (r + 8 * u < r1) {
          float d[8], z[8];
          unpack8(dh[u], dl[u], has_lo, d);
          unpack8(zh[u], zl[u], has_lo, z);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            a[k] += d[k];
            b[k] += d[k] * ((z[k] - mu[k]) * rsd[k]);
          }
        }
    }
  }
  const int t = threadIdx.x;
  const int c = blockIdx.x * 256 + t;
#pragma unroll
  for (int which = 0; which < 2; ++which) {
#pragma unroll
    for (int k = 0; k < 8; ++k) sm[w][lane * 8 + k] = which ? b[k] : a[k];
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) s += sm[y][t];
    if (c < ld) ws[(static_cast<size_t>(rs) * 2 + which) * ld + c] = s;
    __syncthreads();
  }
  __threadfence();
  __syncthreads();
  if (t == 0) last = atomicAdd(&counters[blockIdx.x], 1u);
  __syncthreads();
  if (last != BN_RS - 1) return;
  __threadfence();
  // final sums of this 256-column group: warp w adds partial rows w, w+8, ... (lane = 8 columns, loads batched),
  // then the eight warp totals are added in fixed order
  float f1[8], f2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    f1[k] = 0.f;
    f2[k] = 0.f;
  }
  if (col < ld) {
#pragma unroll 5
    for (int r = w; r < BN_RS; r += 8) {
      const float4* pa = reinterpret_cast<const float4*>(ws + (static_cast<size_t>(r) * 2 + 0) * ld + col);
      const float4* pb = reinterpret_cast<const float4*>(ws + (static_cast<size_t>(r) * 2 + 1) * ld + col);
      const float4 a0 = __ldcg(pa), a1 = __ldcg(pa + 1), b0 = __ldcg(pb), b1 = __ldcg(pb + 1);
      f1[0] += a0.x; f1[1] += a0.y; f1[2] += a0.z; f1[3] += a0.w; f1[4] += a1.x; f1[5] += a1.y; f1[6] += a1.z; f1[7] += a1.w;
      f2[0] += b0.x; f2[1] += b0.y; f2[2] += b0.z; f2[3] += b0.w; f2[4] += b1.x; f2[5] += b1.y; f2[6] += b1.z; f2[7] += b1.w;
    }
  }
  float tot[2];
#pragma unroll
  for (int which = 0; which < 2; ++which) {
#pragma unroll
    for (int k = 0; k < 8; ++k) sm[w][lane * 8 + k] = which ? f2[k] : f1[k];
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) s += sm[y][t];
    tot[which] = s;
    __syncthreads();
  }
  if (c < N) {
    sums[c] = tot[0];
    sums[ld + c] = tot[1];
    g_beta[c] += tot[0];
  }
  if (t == 0) counters[blockIdx.x] = 0u;
}

// ------------------------------------------------------------------------------------------------
// Column-strip batch-norm kernels: one block owns 16 columns and ALL rows.  It first finishes the column reductions of
// its 16 columns from the per-32-row partials a GEMM epilogue wrote (a few KB, fixed summation order), then streams the
// strip (16 bf16 = 32 bytes per row: whole sectors) through the element-wise transform.  One launch per layer and
// direction instead of a tiny latency-bound `finalize` launch followed by an `apply` launch (the two finalize kernels
// were 78 us of a C4 step: profiles/r2b_ncu_full_summary_c4_small_kernels.txt).
constexpr int STRIP_COLS = 16;
constexpr int STRIP_ROWS_IN_FLIGHT = 4;       // bf16x3 (hi + lo arrays)
constexpr int STRIP_ROWS_IN_FLIGHT_BF16 = 6;  // bf16: a thread's 10-11 rows of a C4 strip (4096 rows, 3 row splits) in two batches
constexpr int STRIP_BLOCKS_PER_SM = 3;

// eight consecutive floats of a 16-byte-aligned shared array; `volatile` keeps the loads where they are written (hoisted
// out of the row loop they would pin 24-32 registers for the whole kernel)
__device__ __forceinline__ void lds8(const float* p, float (&v)[8]) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(a) : "memory");
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(a + 16u) : "memory");
}

// forward: statistics (training: from the partials, with the moving-average update; eval: the moving statistics), then
// y = f((z - mean) * rstd + beta) [* dropout]   (reference: classifiers/activation.py:159-161 -> nonlinearity -> 140-141)
template <bool X3>
__global__ void __launch_bounds__(256, STRIP_BLOCKS_PER_SM)
bn_fwd_strip_kernel(const float* __restrict__ ps, const float* __restrict__ pq, int groups, int pld,
                    const __nv_bfloat16* __restrict__ z_hi, const __nv_bfloat16* __restrict__ z_lo, int ld, int B, int N,
                    float eps, float decay, int training, float* __restrict__ mean, float* __restrict__ rstd,
                    float* __restrict__ mm, float* __restrict__ mv, const float* __restrict__ beta, int act,
                    unsigned int drop_thr, float keep_inv, unsigned long long seed, __nv_bfloat16* __restrict__ y_hi,
                    __nv_bfloat16* __restrict__ y_lo) {
  pdl_enter();
  __shared__ double sm[2][16][STRIP_COLS + 1];
  __shared__ __align__(16) float s_mu[STRIP_COLS], s_rs[STRIP_COLS], s_be[STRIP_COLS];
  const int t = threadIdx.x;
  const int c0 = blockIdx.x * STRIP_COLS;
  // streaming role of this thread: 8 of the strip's 16 columns, one row per 128-row sweep.  Its first rows are
  // requested (and their dropout decisions drawn) BEFORE the statistics are reduced: neither depends on them, and the
  // reduction below is two dependent round trips to L2 during which the block would otherwise have nothing in flight.
  const int h = t & 1, rr = t >> 1;
  const int c = c0 + 8 * h;
  const bool active = c < ld;
  const int rstep = 128 * static_cast<int>(gridDim.y);
  const int r_first = rr + 128 * static_cast<int>(blockIdx.y);
  constexpr int RIF = X3 ? STRIP_ROWS_IN_FLIGHT : STRIP_ROWS_IN_FLIGHT_BF16;
  uint4 hv[RIF], lv[RIF];
  uint32_t keepb[RIF];
  auto load_batch = [&](int r0) {
#pragma unroll
    for (int u = 0; u < RIF; ++u) {
      const int r = r0 + rstep * u;
      keepb[u] = 0xFFu;
      if (active && r < B) {
        const size_t o = static_cast<size_t>(r) * ld + c;
        hv[u] = __ldg(reinterpret_cast<const uint4*>(z_hi + o));
        if (X3) lv[u] = __ldg(reinterpret_cast<const uint4*>(z_lo + o));
        if (drop_thr != 0u)
          keepb[u] = dropout_keep_bits(philox4x32_10(static_cast<uint32_t>(c >> 3), static_cast<uint32_t>(r), 0u, 0u,
                                                     static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)), drop_thr);
      }
    }
  };
  load_batch(r_first);
  // the finishing threads' own operands are requested now too, not behind the reduction's barrier
  float pre_beta = 0.f, pre_mm = 0.f, pre_mv = 1.f;
  if (t < STRIP_COLS && c0 + t < N) {
    pre_beta = beta[c0 + t];
    if (!training || blockIdx.y == 0) {
      pre_mm = mm[c0 + t];
      pre_mv = mv[c0 + t];
    }
  }
  {
    const int cl = t & (STRIP_COLS - 1), gl = t >> 4;  // 16 columns x 16 group lanes
    const int c = c0 + cl;
    double s1 = 0.0, s2 = 0.0;
    if (training && c < N) {
      for (int g0 = gl; g0 < groups; g0 += 64) {  // four independent loads per array in flight
        float a[4], b[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int g = g0 + 16 * k;
          a[k] = g < groups ? ps[static_cast<size_t>(g) * pld + c] : 0.f;
          b[k] = g < groups ? pq[static_cast<size_t>(g) * pld + c] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          s1 += static_cast<double>(a[k]);
          s2 += static_cast<double>(b[k]);
        }
      }
    }
    sm[0][gl][cl] = s1;
    sm[1][gl][cl] = s2;
    __syncthreads();
    if (t < STRIP_COLS) {
      const int cc = c0 + t;
      float muf = 0.f, rsf = 0.f, bef = 0.f;
      if (cc < N) {
        if (training) {
          double a1 = 0.0, a2 = 0.0;
#pragma unroll
          for (int k = 0; k < 16; ++k) {  // fixed order
            a1 += sm[0][k][t];
            a2 += sm[1][k][t];
          }
          const double mu = a1 / B;
          double var = a2 / B - mu * mu;  // biased (no Bessel), as tf.nn.moments
          if (var < 0.0) var = 0.0;
          muf = static_cast<float>(mu);
          const float varf = static_cast<float>(var);
          rsf = rsqrtf(varf + eps);
          if (blockIdx.y == 0) {  // every row split of a strip derives the same statistics; one of them records them
            mm[cc] = pre_mm - (1.0f - decay) * (pre_mm - muf);  // assign_moving_average
            mv[cc] = pre_mv - (1.0f - decay) * (pre_mv - varf);
          }
        } else {
          muf = pre_mm;
          rsf = rsqrtf(pre_mv + eps);
        }
        if (blockIdx.y == 0) {
          mean[cc] = muf;  // kept for the backward pass
          rstd[cc] = rsf;
        }
        bef = pre_beta;
      }
      s_mu[t] = muf;
      s_rs[t] = rsf;
      s_be[t] = bef;
    }
    __syncthreads();
  }
  if (!active) return;
  for (int r0 = r_first; r0 < B; r0 += rstep * RIF) {
    if (r0 != r_first) load_batch(r0);
#pragma unroll
    for (int u = 0; u < RIF; ++u) {
      const int r = r0 + rstep * u;
      if (r >= B) continue;
      float x[8], mu[8], rs[8], be[8];
      unpack8(hv[u], lv[u], X3, x);
      // the strip's constants are re-read from shared memory per row: registers go to rows in flight instead
      lds8(s_mu + 8 * h, mu);
      lds8(s_rs + 8 * h, rs);
      lds8(s_be + 8 * h, be);
      const uint32_t keep = keepb[u];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float y = (x[k] - mu[k]) * rs[k] + be[k];
        if (act == 1) y = fmaxf(y, 0.f);
        else if (act == 2) y = 1.0f / (1.0f + expf(-y));
        else if (act == 3) y = tanhf(y);
        y = (c + k < N) ? y : 0.f;
        x[k] = (drop_thr != 0u) ? (((keep >> k) & 1u) ? y * keep_inv : 0.f) : y;
      }
      store8(y_hi, X3 ? y_lo : nullptr, static_cast<size_t>(r) * ld + c, x);
    }
  }
}

// backward: m1 = mean_B(dy), m2 = mean_B(dy * xhat) from the dgrad epilogue's partials (and dbeta += sum_B dy), then
// dz = rstd * (dy - m1 - xhat * m2) in place over dy
template <bool X3>
__global__ void __launch_bounds__(256, STRIP_BLOCKS_PER_SM)
bn_bwd_strip_kernel(const float* __restrict__ p1, const float* __restrict__ p2, int groups, int pld,
                    __nv_bfloat16* __restrict__ dy_hi, __nv_bfloat16* __restrict__ dy_lo, const __nv_bfloat16* __restrict__ z_hi,
                    const __nv_bfloat16* __restrict__ z_lo, int ld, int B, int N, const float* __restrict__ mean,
                    const float* __restrict__ rstd, float* __restrict__ g_beta) {
  pdl_enter();
  __shared__ float sm[2][16][STRIP_COLS + 1];
  __shared__ __align__(16) float s_mu[STRIP_COLS], s_rs[STRIP_COLS], s_m1[STRIP_COLS], s_m2[STRIP_COLS];
  const int t = threadIdx.x;
  const int c0 = blockIdx.x * STRIP_COLS;
  // as in the forward kernel: the thread's first rows are requested before the reductions are finished
  const int h = t & 1, rr = t >> 1;
  const int c = c0 + 8 * h;
  const bool active = c < ld;
  constexpr int R = X3 ? 2 : STRIP_ROWS_IN_FLIGHT_BF16;  // (two arrays per row)
  const int rstep = 128 * static_cast<int>(gridDim.y);
  const int r_first = rr + 128 * static_cast<int>(blockIdx.y);
  uint4 dh[R], dl[R], zh[R], zl[R];
  auto load_batch = [&](int r0) {
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int r = r0 + rstep * u;
      if (active && r < B) {
        const size_t o = static_cast<size_t>(r) * ld + c;
        dh[u] = *reinterpret_cast<const uint4*>(dy_hi + o);
        zh[u] = __ldg(reinterpret_cast<const uint4*>(z_hi + o));
        if (X3) {
          dl[u] = *reinterpret_cast<const uint4*>(dy_lo + o);
          zl[u] = __ldg(reinterpret_cast<const uint4*>(z_lo + o));
        }
      }
    }
  };
  load_batch(r_first);
  float pre_mu = 0.f, pre_rs = 0.f, pre_gb = 0.f;  // the finishing threads' operands: requested before the reduction
  if (t < STRIP_COLS && c0 + t < N) {
    pre_mu = mean[c0 + t];
    pre_rs = rstd[c0 + t];
    if (blockIdx.y == 0) pre_gb = g_beta[c0 + t];
  }
  {
    const int cl = t & (STRIP_COLS - 1), gl = t >> 4;
    const int c = c0 + cl;
    float s1 = 0.f, s2 = 0.f;
    if (c < N) {
      for (int g0 = gl; g0 < groups; g0 += 64) {
        float a[4], b[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int g = g0 + 16 * k;
          a[k] = g < groups ? p1[static_cast<size_t>(g) * pld + c] : 0.f;
          b[k] = g < groups ? p2[static_cast<size_t>(g) * pld + c] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          s1 += a[k];
          s2 += b[k];
        }
      }
    }
    sm[0][gl][cl] = s1;
    sm[1][gl][cl] = s2;
    __syncthreads();
    if (t < STRIP_COLS) {
      const int cc = c0 + t;
      float t1 = 0.f, t2 = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        t1 += sm[0][k][t];
        t2 += sm[1][k][t];
      }
      const bool ok = cc < N;
      if (ok && blockIdx.y == 0) g_beta[cc] = pre_gb + t1;
      const float invB = 1.0f / static_cast<float>(B);
      s_m1[t] = ok ? t1 * invB : 0.f;
      s_m2[t] = ok ? t2 * invB : 0.f;
      s_mu[t] = pre_mu;
      s_rs[t] = pre_rs;
    }
    __syncthreads();
  }
  if (!active) return;
  for (int r0 = r_first; r0 < B; r0 += rstep * R) {
    if (r0 != r_first) load_batch(r0);
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int r = r0 + rstep * u;
      if (r >= B) continue;
      float d[8], z[8], mu[8], rs[8], m1[8], m2[8];
      unpack8(dh[u], dl[u], X3, d);
      unpack8(zh[u], zl[u], X3, z);
      lds8(s_mu + 8 * h, mu);
      lds8(s_rs + 8 * h, rs);
      lds8(s_m1 + 8 * h, m1);
      lds8(s_m2 + 8 * h, m2);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float xh = (z[k] - mu[k]) * rs[k];
        d[k] = (c + k < N) ? rs[k] * (d[k] - m1[k] - xh * m2[k]) : 0.f;
      }
      store8(dy_hi, X3 ? dy_lo : nullptr, static_cast<size_t>(r) * ld + c, d);
    }
  }
}

// dz = rstd * (dy - mean_B(dy) - xhat * mean_B(dy * xhat)), in place over dy, from the sums of bn_bwd_reduce_kernel
// (the unit entry point tfk_fflayer_bwd and L2Norm chains; the training step uses bn_bwd_strip_kernel)
template <bool X3>
__global__ void __launch_bounds__(256, 4)
bn_bwd_apply_kernel(__nv_bfloat16* __restrict__ dy_hi, __nv_bfloat16* __restrict__ dy_lo,
                    const __nv_bfloat16* __restrict__ z_hi, const __nv_bfloat16* __restrict__ z_lo, int ld,
                    int B, int N, const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ sums) {
  __shared__ float s_mu[8][256], s_rs[8][256], s_m1[8][256], s_m2[8][256];
  const int cb = blockIdx.x * BN_SPAN;
  const float invB = 1.0f / static_cast<float>(B);
  const int t = threadIdx.x;
  const int c = cb + (t << 3);
  if (c >= ld) return;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const bool ok = c + k < N;
    s_mu[k][t] = ok ? __ldg(mean + c + k) : 0.f;
    s_rs[k][t] = ok ? __ldg(rstd + c + k) : 0.f;
    s_m1[k][t] = ok ? __ldg(sums + c + k) * invB : 0.f;
    s_m2[k][t] = ok ? __ldg(sums + ld + c + k) * invB : 0.f;
  }
  const int step = static_cast<int>(gridDim.y);
  constexpr int R = X3 ? 2 : 4;  // rows in flight (two arrays each)
  for (int r0 = blockIdx.y; r0 < B; r0 += R * step) {
    uint32_t dh[R][4], dl[R][4], zh[R][4], zl[R][4];
#pragma unroll
    for (int u = 0; u < R; ++u)
      if (r0 + u * step < B) {
        const size_t o = static_cast<size_t>(r0 + u * step) * ld + c;
        const uint4 a = *reinterpret_cast<const uint4*>(dy_hi + o);
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(z_hi + o));
        dh[u][0] = a.x; dh[u][1] = a.y; dh[u][2] = a.z; dh[u][3] = a.w;
        zh[u][0] = b.x; zh[u][1] = b.y; zh[u][2] = b.z; zh[u][3] = b.w;
        if (X3) {
          const uint4 e = *reinterpret_cast<const uint4*>(dy_lo + o);
          const uint4 f = __ldg(reinterpret_cast<const uint4*>(z_lo + o));
          dl[u][0] = e.x; dl[u][1] = e.y; dl[u][2] = e.z; dl[u][3] = e.w;
          zl[u][0] = f.x; zl[u][1] = f.y; zl[u][2] = f.z; zl[u][3] = f.w;
        }
      }
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      float mu[4], rs[4], m1[4], m2[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        mu[i] = s_mu[4 * jj + i][t];
        rs[i] = s_rs[4 * jj + i][t];
        m1[i] = s_m1[4 * jj + i][t];
        m2[i] = s_m2[4 * jj + i][t];
      }
#pragma unroll
      for (int u = 0; u < R; ++u) {
        if (r0 + u * step < B) {
#pragma unroll
          for (int w = 0; w < 2; ++w) {
            const int j = 2 * jj + w;
            float d[2] = {bf_lo(dh[u][j]), bf_hi(dh[u][j])};
            float z[2] = {bf_lo(zh[u][j]), bf_hi(zh[u][j])};
            if (X3) {
              d[0] += bf_lo(dl[u][j]);
              d[1] += bf_hi(dl[u][j]);
              z[0] += bf_lo(zl[u][j]);
              z[1] += bf_hi(zl[u][j]);
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int i = 2 * w + e;
              const float xh = (z[e] - mu[i]) * rs[i];
              d[e] = (c + 4 * jj + i < N) ? rs[i] * (d[e] - m1[i] - xh * m2[i]) : 0.f;
            }
            split2(d[0], d[1], dh[u][j], dl[u][j]);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int r = r0 + u * step;
      if (r < B) {
        const size_t o = static_cast<size_t>(r) * ld + c;
        *reinterpret_cast<uint4*>(dy_hi + o) = make_uint4(dh[u][0], dh[u][1], dh[u][2], dh[u][3]);
        if (X3) *reinterpret_cast<uint4*>(dy_lo + o) = make_uint4(dl[u][0], dl[u][1], dl[u][2], dl[u][3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// L2Norm: one warp per frame, the row is walked twice (second pass hits L1/L2); 8 columns per lane per step.
__global__ void __launch_bounds__(256)
l2norm_fwd_kernel(const __nv_bfloat16* __restrict__ u_hi, const __nv_bfloat16* __restrict__ u_lo, int ld, int B,
                  int N, unsigned int drop_thr, float keep_inv, unsigned long long seed,
                  __nv_bfloat16* __restrict__ y_hi, __nv_bfloat16* __restrict__ y_lo, float* __restrict__ s_out) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const size_t base = static_cast<size_t>(row) * ld;
  float ss = 0.f;
  for (int c = lane * 8; c < ld; c += 256) {
    float x[8];
    load8(u_hi, u_lo, base + c, x);
#pragma unroll
    for (int k = 0; k < 8; ++k) ss += (c + k < N) ? x[k] * x[k] : 0.f;
  }
  ss = warp_sum(ss);
  const float sig = ss / static_cast<float>(N);  // reduce_mean(square(activations), 1)
  if (lane == 0) s_out[row] = sig;
  const bool norm = sig > 1.0f;
  for (int c = lane * 8; c < ld; c += 256) {
    float x[8];
    load8(u_hi, u_lo, base + c, x);
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = (c + k < N) ? (norm ? x[k] / sig : x[k]) : 0.f;
    if (drop_thr != 0u) {
      const uint32_t keep = dropout_keep_bits(philox4x32_10(static_cast<uint32_t>(c >> 3), static_cast<uint32_t>(row), 0u, 0u,
                                                            static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)), drop_thr);
#pragma unroll
      for (int k = 0; k < 8; ++k) x[k] = ((keep >> k) & 1u) ? x[k] * keep_inv : 0.f;
    }
    store8(y_hi, y_lo, base + c, x);
  }
}
__global__ void __launch_bounds__(256)
l2norm_bwd_kernel(__nv_bfloat16* __restrict__ d_hi, __nv_bfloat16* __restrict__ d_lo,
                  const __nv_bfloat16* __restrict__ u_hi, const __nv_bfloat16* __restrict__ u_lo,
                  const float* __restrict__ s_in, int ld, int B, int N, int act) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const size_t base = static_cast<size_t>(row) * ld;
  const float sig = s_in[row];
  const bool norm = sig > 1.0f;
  float dot = 0.f;
  if (norm) {
    for (int c = lane * 8; c < ld; c += 256) {
      float d[8], u[8];
      load8(d_hi, d_lo, base + c, d);
      load8(u_hi, u_lo, base + c, u);
#pragma unroll
      for (int k = 0; k < 8; ++k) dot += (c + k < N) ? d[k] * u[k] : 0.f;
    }
    dot = warp_sum(dot);
  }
  const float inv = norm ? 1.0f / sig : 1.0f;
  const float coef = norm ? 2.0f * dot / (static_cast<float>(N) * sig * sig) : 0.f;
  for (int c = lane * 8; c < ld; c += 256) {
    float d[8], u[8];
    load8(d_hi, d_lo, base + c, d);
    load8(u_hi, u_lo, base + c, u);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float g = (c + k < N) ? d[k] * inv - u[k] * coef : 0.f;
      if (act == 1) g = u[k] > 0.f ? g : 0.f;
      else if (act == 2) g *= u[k] * (1.0f - u[k]);
      else if (act == 3) g *= 1.0f - u[k] * u[k];
      d[k] = g;
    }
    store8(d_hi, d_lo, base + c, d);
  }
}

// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256)
decode_out_kernel(const float* __restrict__ logits, int ld, int T, int O, const float* __restrict__ prior,
                  float* __restrict__ out) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= T) return;
  const int nvec = ld >> 2;
  const float4* zp = reinterpret_cast<const float4*>(logits + static_cast<size_t>(row) * ld);
  float4 z[NV];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = lane + 32 * i;
    float4 t = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    if (idx < nvec) {
      t = __ldg(zp + idx);
      const int e = idx << 2;
      if (e + 1 > O) t.x = -INFINITY;
      if (e + 2 > O) t.y = -INFINITY;
      if (e + 3 > O) t.z = -INFINITY;
      if (e + 4 > O) t.w = -INFINITY;
    }
    z[i] = t;
    mx = fmaxf(mx, fmaxf(fmaxf(t.x, t.y), fmaxf(t.z, t.w)));
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    z[i].x = __expf(z[i].x - mx);
    z[i].y = __expf(z[i].y - mx);
    z[i].z = __expf(z[i].z - mx);
    z[i].w = __expf(z[i].w - mx);
    sum += (z[i].x + z[i].y) + (z[i].z + z[i].w);
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  const float lsum = logf(sum);
  float* op = out + static_cast<size_t>(row) * O;
  const bool vec_store = (O & 3) == 0;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int e = (lane + 32 * i) << 2;
    if (e >= O) continue;
    float p[4] = {z[i].x * inv, z[i].y * inv, z[i].z * inv, z[i].w * inv};
    if (prior) {
      // np.log(output/prior) (nnet.py:280-286) evaluated as log(e^(z-max)) - log(sum) - log(prior): one log per
      // element less; `prior` here is log(prior) (k_decode_out precomputes it); an underflowed posterior keeps
      // the reference's -inf
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (e + k < O) {
          const float ez = k == 0 ? z[i].x : k == 1 ? z[i].y : k == 2 ? z[i].z : z[i].w;
          p[k] = (p[k] > 0.f) ? __logf(ez) - lsum - __ldg(prior + e + k) : -INFINITY - __ldg(prior + e + k);
        }
    }
    if (vec_store) {
      *reinterpret_cast<float4*>(op + e) = make_float4(p[0], p[1], p[2], p[3]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (e + k < O) op[e + k] = p[k];
    }
  }
}
__global__ void __launch_bounds__(256)
decode_out_generic_kernel(const float* __restrict__ logits, int ld, int T, int O,
                          const float* __restrict__ prior, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= T) return;
  const float* zp = logits + static_cast<size_t>(row) * ld;
  float mx = -INFINITY;
  for (int e = lane; e < O; e += 32) mx = fmaxf(mx, zp[e]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int e = lane; e < O; e += 32) sum += expf(zp[e] - mx);
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  for (int e = lane; e < O; e += 32) {
    const float ez = expf(zp[e] - mx);
    float p = ez * inv;
    if (prior) p = (p > 0.f) ? logf(ez) - logf(sum) - prior[e] : -INFINITY - prior[e];  // prior = log(prior)
    out[static_cast<size_t>(row) * O + e] = p;
  }
}

inline int grid_for(size_t work_items, int block, int cap = 148 * 16) {
  size_t g = (work_items + block - 1) / block;
  if (g > static_cast<size_t>(cap)) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace

int k_split_f32(const float* src, int ld_src, __nv_bfloat16* hi, __nv_bfloat16* lo, int ld_dst, int rows,
                int cols, cudaStream_t st) {
  if (rows <= 0) return 0;
  const int vec_ok = (ld_src % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  const size_t total = static_cast<size_t>(rows) * (ld_dst >> 2);
  launch_pdl(split_f32_kernel, dim3(grid_for(total, 256)), dim3(256), st, src, ld_src, hi, lo, ld_dst, rows, cols, vec_ok);
  return static_cast<int>(cudaGetLastError());
}

int k_splice_cmvn(const float* raw, const int32_t* utt_off, int num_utts, const float* cmvn, int D, int k,
                  int row_begin, int rows, __nv_bfloat16* hi, __nv_bfloat16* lo, int ld_dst, cudaStream_t st) {
  if (rows <= 0) return 0;
  launch_pdl(splice_cmvn_kernel, dim3((rows + 7) / 8), dim3(256), st, raw, utt_off, num_utts, cmvn, D, k, row_begin, rows, hi, lo, ld_dst);
  return static_cast<int>(cudaGetLastError());
}

int k_merge_bf16(const __nv_bfloat16* hi, const __nv_bfloat16* lo, int ld_src, float* dst, int ld_dst,
                 int rows, int cols, cudaStream_t st) {
  if (rows <= 0) return 0;
  merge_bf16_kernel<<<grid_for(static_cast<size_t>(rows) * cols, 256), 256, 0, st>>>(hi, lo, ld_src, dst,
                                                                                   ld_dst, rows, cols);
  return static_cast<int>(cudaGetLastError());
}

int k_mask_scale_f32(const float* dy, const float* y, float* out, size_t n, float scale, int mode,
                     cudaStream_t st) {
  if (n == 0) return 0;
  mask_scale_f32_kernel<<<grid_for(n, 256), 256, 0, st>>>(dy, y, out, n, scale, mode);
  return static_cast<int>(cudaGetLastError());
}
int k_double_to_float(const double* src, float* dst, cudaStream_t st) {
  double_to_float_kernel<<<1, 1, 0, st>>>(src, dst);
  return static_cast<int>(cudaGetLastError());
}

int k_softmax_ce(const float* logits, int ld, const int32_t* labels, int B, int O, float* row_loss,
                 __nv_bfloat16* d_hi, __nv_bfloat16* d_lo, cudaStream_t st) {
  if (B <= 0) return 0;
  const int grid = (B + 7) / 8;
  if ((ld & 3) != 0) return static_cast<int>(cudaErrorInvalidValue);  // rows are float4-aligned (ld is a multiple of 8)
  launch_pdl(softmax_ce_stream_kernel, dim3(grid), dim3(256), st, logits, ld, labels, B, O, row_loss, d_hi, d_lo);
  return static_cast<int>(cudaGetLastError());
}

int k_accum_loss(const float* row_loss, int B, double* acc, cudaStream_t st) {
  launch_pdl(accum_loss_kernel, dim3(1), dim3(1024), st, row_loss, B, acc);
  return static_cast<int>(cudaGetLastError());
}

int k_colsum_bf16(const __nv_bfloat16* hi, const __nv_bfloat16* lo, int ld, int rows, int cols, float* ws,
                  float* out, cudaStream_t st) {
  if (rows <= 0) return 0;
  // ws layout: 1024 self-resetting counters (one per 256-column group, zero-initialised) at a FIXED place,
  // then [COLSUM2_RS * ld] partials (calls with different ld must not clobber the counters)
  if (ld > 65536) return static_cast<int>(cudaErrorInvalidValue);
  unsigned int* counters = reinterpret_cast<unsigned int*>(ws);
  dim3 grid((ld + 255) / 256, COLSUM2_RS);
  colsum_kernel<<<grid, 256, 0, st>>>(hi, lo, ld, rows, cols, ws + 1024, counters, out);
  return static_cast<int>(cudaGetLastError());
}

int k_colsum_finalize(const float* const* parts, float* const* outs, int njobs, int groups, int ld, int cols,
                      cudaStream_t st) {
  for (int j0 = 0; j0 < njobs; j0 += 16) {
    ColsumJobs J;
    J.njobs = njobs - j0 < 16 ? njobs - j0 : 16;
    for (int j = 0; j < J.njobs; ++j) {
      J.part[j] = parts[j0 + j];
      J.out[j] = outs[j0 + j];
    }
    J.groups = groups; J.ld = ld; J.cols = cols;
    dim3 grid((cols + 31) / 32, J.njobs);
    launch_pdl(colsum_finalize_kernel, grid, dim3(256), st, J);
  }
  return static_cast<int>(cudaGetLastError());
}

int k_adam(float* w, float* g, float* m, float* v, __nv_bfloat16* w_hi, __nv_bfloat16* w_lo, const size_t* seg_off,
           const size_t* seg_cnt, int nseg, const double* acc, float lr_t, float beta1, float beta2, float eps,
           cudaStream_t st, int n_bcast, int n_peers, __nv_bfloat16* const* peer_hi, __nv_bfloat16* const* peer_lo) {
  for (int s0 = 0; s0 < nseg; s0 += 64) {
    AdamSegs segs;
    segs.n = nseg - s0 < 64 ? nseg - s0 : 64;
    segs.n_bcast = n_bcast - s0 > 0 ? (n_bcast - s0 < segs.n ? n_bcast - s0 : segs.n) : 0;
    segs.n_peers = segs.n_bcast > 0 ? n_peers : 0;
    if (segs.n_peers > 15) return static_cast<int>(cudaErrorInvalidValue);
    for (int r = 0; r < segs.n_peers; ++r) {
      segs.peer_hi[r] = reinterpret_cast<uint4*>(peer_hi[r]);
      segs.peer_lo[r] = peer_lo ? reinterpret_cast<uint4*>(peer_lo[r]) : nullptr;
    }
    size_t longest = 0;
    for (int i = 0; i < segs.n; ++i) {
      if ((seg_off[s0 + i] | seg_cnt[s0 + i]) & 7) return static_cast<int>(cudaErrorInvalidValue);  // 8-float granules
      segs.off4[i] = seg_off[s0 + i] >> 2;
      segs.cnt4[i] = seg_cnt[s0 + i] >> 2;
      longest = segs.cnt4[i] > longest ? segs.cnt4[i] : longest;
    }
    if (longest == 0) continue;
    int per_seg = 148 * 16 / segs.n;
    if (per_seg < 16) per_seg = 16;
    dim3 grid(grid_for(longest >> 1, 128, per_seg), segs.n);
    adam_kernel<<<grid, 128, 0, st>>>(reinterpret_cast<float4*>(w), reinterpret_cast<float4*>(g),
                                      reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v),
                                      reinterpret_cast<uint4*>(w_hi), reinterpret_cast<uint4*>(w_lo), segs, acc, lr_t,
                                      beta1, beta2, eps);
  }
  return static_cast<int>(cudaGetLastError());
}

// Hand-rolled all-reduce of the small replicated vectors (bias / beta gradients) and {loss, frames} over peer memory:
// every rank stores its copy into slot `me` of the exchange area behind every GPU's flag words (its own included),
// publishes, waits for the others, and every rank then adds the n slots in RANK ORDER — identical sums everywhere.
__global__ void __launch_bounds__(256)
dp_small_push_kernel(int* const* __restrict__ peer_flags, int n_peers, int* __restrict__ own_flags, int flag_words, int me,
                     const float4* __restrict__ g_small, const double* __restrict__ acc, int small_n4, int stride) {
  int* base = static_cast<int>(blockIdx.y) < n_peers ? peer_flags[blockIdx.y] : own_flags;
  float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(base + flag_words) + static_cast<size_t>(me) * stride);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < small_n4; i += gridDim.x * blockDim.x) dst[i] = g_small[i];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double2 a = make_double2(acc[0], acc[1]);
    *reinterpret_cast<double2*>(dst + small_n4) = a;  // (stride keeps this 16-byte aligned)
  }
  __threadfence_system();
}
__global__ void __launch_bounds__(256)
dp_small_reduce_kernel(const int* __restrict__ own_flags, int flag_words, int n_ranks, float4* __restrict__ g_small,
                       double* __restrict__ acc, int small_n4, int stride) {
  const float* x = reinterpret_cast<const float*>(own_flags + flag_words);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < small_n4; i += gridDim.x * blockDim.x) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < n_ranks; ++r) {  // fixed order
      const float4 v = __ldcg(reinterpret_cast<const float4*>(x + static_cast<size_t>(r) * stride) + i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    g_small[i] = s;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double l = 0.0, f = 0.0;
    for (int r = 0; r < n_ranks; ++r) {
      const double2 v = __ldcg(reinterpret_cast<const double2*>(reinterpret_cast<const float4*>(x + static_cast<size_t>(r) * stride) + small_n4));
      l += v.x;
      f += v.y;
    }
    acc[0] = l;
    acc[1] = f;
  }
}
int k_dp_small_push(int* const* d_peer_flags, int n_peers, int* own_flags, int me, const float* g_small, const double* acc,
                    int small_n, int stride, cudaStream_t st) {
  const int n4 = small_n >> 2;
  dim3 grid((n4 + 255) / 256 > 16 ? 16 : ((n4 + 255) / 256 < 1 ? 1 : (n4 + 255) / 256), n_peers + 1);
  dp_small_push_kernel<<<grid, 256, 0, st>>>(d_peer_flags, n_peers, own_flags, TFK_DP_FLAG_WORDS, me,
                                             reinterpret_cast<const float4*>(g_small), acc, n4, stride);
  return static_cast<int>(cudaGetLastError());
}
int k_dp_small_reduce(const int* own_flags, int n_ranks, float* g_small, double* acc, int small_n, int stride, cudaStream_t st) {
  const int n4 = small_n >> 2;
  int blocks = (n4 + 255) / 256;
  blocks = blocks > 64 ? 64 : (blocks < 1 ? 1 : blocks);
  dp_small_reduce_kernel<<<blocks, 256, 0, st>>>(own_flags, TFK_DP_FLAG_WORDS, n_ranks, reinterpret_cast<float4*>(g_small), acc, n4, stride);
  return static_cast<int>(cudaGetLastError());
}

int k_dp_publish(int* const* d_peer_flags, int n_peers, int slot, int me, int value, cudaStream_t st) {
  dp_publish_kernel<<<1, 32, 0, st>>>(d_peer_flags, n_peers, slot, me, value);
  return static_cast<int>(cudaGetLastError());
}
int k_dp_wait(const int* flags, int n_ranks, int slot, int me, int value, cudaStream_t st) {
  dp_wait_kernel<<<1, 32, 0, st>>>(flags, n_ranks, slot, me, value);
  return static_cast<int>(cudaGetLastError());
}

int k_bn_bwd_reduce(const __nv_bfloat16* dy_hi, const __nv_bfloat16* dy_lo, const __nv_bfloat16* z_hi,
                    const __nv_bfloat16* z_lo, int ld, int B, int N, const float* mean, const float* rstd,
                    float* ws, unsigned int* counters, float* sums, float* g_beta, cudaStream_t st) {
  if (B <= 0) return 0;
  if (ld > 256 * 512) return static_cast<int>(cudaErrorInvalidValue);  // 512 counters
  dim3 grid((ld + 255) / 256, BN_RS);
  if (dy_lo)
    bn_bwd_reduce_kernel<true><<<grid, 256, 0, st>>>(dy_hi, dy_lo, z_hi, z_lo, ld, B, N, mean, rstd, ws, counters,
                                                     sums, g_beta);
  else
    bn_bwd_reduce_kernel<false><<<grid, 256, 0, st>>>(dy_hi, dy_lo, z_hi, z_lo, ld, B, N, mean, rstd, ws, counters,
                                                      sums, g_beta);
  return static_cast<int>(cudaGetLastError());
}
// strips x row splits (every row split re-derives its strip's statistics from the same partials: a few KB from L2,
// identical results)
static dim3 strip_grid(int ld, int B) {
  const int strips = (ld + STRIP_COLS - 1) / STRIP_COLS;
  // ONE wave of resident blocks: these kernels are latency chains (reduce the statistics, then one or two batches of rows),
  // so a partly filled second wave costs a whole chain more (640 blocks on 444 slots: 18 us per launch)
  int rs = 148 * STRIP_BLOCKS_PER_SM / strips;
  const int max_rs = (B + 127) / 128;
  if (rs > max_rs) rs = max_rs;
  if (rs < 1) rs = 1;
  return dim3(strips, rs);
}
int k_bn_fwd_strip(const float* part_sum, const float* part_sq, int groups, int pld, const __nv_bfloat16* z_hi,
                   const __nv_bfloat16* z_lo, int ld, int B, int N, float eps, float decay, int training, float* mean, float* rstd,
                   float* moving_mean, float* moving_var, const float* beta, int act, float keep, unsigned long long seed,
                   __nv_bfloat16* y_hi, __nv_bfloat16* y_lo, cudaStream_t st) {
  if (B <= 0) return 0;
  const unsigned int thr = keep < 1.0f ? dropout_threshold(keep) : 0u;
  const dim3 grid = strip_grid(ld, B);
  if (z_lo)
    launch_pdl(bn_fwd_strip_kernel<true>, grid, dim3(256), st, part_sum, part_sq, groups, pld, z_hi, z_lo, ld, B, N, eps, decay,
               training, mean, rstd, moving_mean, moving_var, beta, act, thr, 1.0f / keep, seed, y_hi, y_lo);
  else
    launch_pdl(bn_fwd_strip_kernel<false>, grid, dim3(256), st, part_sum, part_sq, groups, pld, z_hi, z_lo, ld, B, N, eps, decay,
               training, mean, rstd, moving_mean, moving_var, beta, act, thr, 1.0f / keep, seed, y_hi, y_lo);
  return static_cast<int>(cudaGetLastError());
}
int k_bn_bwd_strip(const float* part_sum, const float* part_dot, int groups, int pld, __nv_bfloat16* dy_hi, __nv_bfloat16* dy_lo,
                   const __nv_bfloat16* z_hi, const __nv_bfloat16* z_lo, int ld, int B, int N, const float* mean, const float* rstd,
                   float* g_beta, cudaStream_t st) {
  if (B <= 0) return 0;
  const dim3 grid = strip_grid(ld, B);
  if (dy_lo)
    launch_pdl(bn_bwd_strip_kernel<true>, grid, dim3(256), st, part_sum, part_dot, groups, pld, dy_hi, dy_lo, z_hi, z_lo, ld, B, N, mean, rstd, g_beta);
  else
    launch_pdl(bn_bwd_strip_kernel<false>, grid, dim3(256), st, part_sum, part_dot, groups, pld, dy_hi, dy_lo, z_hi, z_lo, ld, B, N, mean, rstd, g_beta);
  return static_cast<int>(cudaGetLastError());
}
int k_bn_bwd_apply(__nv_bfloat16* dy_hi, __nv_bfloat16* dy_lo, const __nv_bfloat16* z_hi,
                   const __nv_bfloat16* z_lo, int ld, int B, int N, const float* mean, const float* rstd,
                   const float* sums, cudaStream_t st) {
  if (B <= 0) return 0;
  const int gx = (ld + BN_SPAN - 1) / BN_SPAN;
  int gy = 148 * 4 / gx;
  gy = gy < 1 ? 1 : (gy > B ? B : gy);
  if (dy_lo)
    bn_bwd_apply_kernel<true><<<dim3(gx, gy), 256, 0, st>>>(dy_hi, dy_lo, z_hi, z_lo, ld, B, N, mean, rstd, sums);
  else
    bn_bwd_apply_kernel<false><<<dim3(gx, gy), 256, 0, st>>>(dy_hi, dy_lo, z_hi, z_lo, ld, B, N, mean, rstd, sums);
  return static_cast<int>(cudaGetLastError());
}

int k_l2norm_fwd(const __nv_bfloat16* u_hi, const __nv_bfloat16* u_lo, int ld, int B, int N, float keep,
                 unsigned long long seed, __nv_bfloat16* y_hi, __nv_bfloat16* y_lo, float* s_out, cudaStream_t st) {
  if (B <= 0) return 0;
  const unsigned int thr = keep < 1.0f ? dropout_threshold(keep) : 0u;
  l2norm_fwd_kernel<<<(B + 7) / 8, 256, 0, st>>>(u_hi, u_lo, ld, B, N, thr, 1.0f / keep, seed, y_hi, y_lo, s_out);
  return static_cast<int>(cudaGetLastError());
}
int k_l2norm_bwd(__nv_bfloat16* d_hi, __nv_bfloat16* d_lo, const __nv_bfloat16* u_hi, const __nv_bfloat16* u_lo,
                 const float* s_in, int ld, int B, int N, int act, cudaStream_t st) {
  if (B <= 0) return 0;
  l2norm_bwd_kernel<<<(B + 7) / 8, 256, 0, st>>>(d_hi, d_lo, u_hi, u_lo, s_in, ld, B, N, act);
  return static_cast<int>(cudaGetLastError());
}

__global__ void log_vector_kernel(const float* __restrict__ x, float* __restrict__ y, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = logf(x[i]);  // log(0) = -inf  ->  +inf log-likelihood, as the reference
}
int k_log_vector(const float* x, float* y, int n, cudaStream_t st) {
  log_vector_kernel<<<(n + 255) / 256, 256, 0, st>>>(x, y, n);
  return static_cast<int>(cudaGetLastError());
}

int k_decode_out(const float* logits, int ld, int T, int O, const float* prior, float* out, cudaStream_t st) {
  if (T <= 0) return 0;
  const int grid = (T + 7) / 8;
  if (ld <= 1024)
    launch_pdl(decode_out_kernel<8>, dim3(grid), dim3(256), st, logits, ld, T, O, prior, out);
  else if (ld <= 2048)
    launch_pdl(decode_out_kernel<16>, dim3(grid), dim3(256), st, logits, ld, T, O, prior, out);
  else if (ld <= 4096)
    launch_pdl(decode_out_kernel<32>, dim3(grid), dim3(256), st, logits, ld, T, O, prior, out);
  else
    decode_out_generic_kernel<<<grid, 256, 0, st>>>(logits, ld, T, O, prior, out);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace tfk
