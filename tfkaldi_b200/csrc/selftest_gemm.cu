// Stand-alone GPU self-test + micro-benchmark of the tcgen05 GEMM (no torch, no Python):
//   nvcc ... selftest_gemm.cu gemm.cu -o build/selftest_gemm ; ./build/selftest_gemm [quick|full|bench]
// Compares every operand-major / epilogue combination used by the engine against a double-precision
// host reference on small, ragged and full-size (sampled) shapes.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "gemm.cuh"
#include "philox.cuh"

using namespace tfk;

#define CK(x)                                                                            \
  do {                                                                                   \
    cudaError_t e_ = (x);                                                                \
    if (e_ != cudaSuccess) {                                                             \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);    \
      exit(2);                                                                           \
    }                                                                                    \
  } while (0)

static int g_sms = 148;
static int* g_sched = nullptr;
static int g_fail = 0;
static int g_two_cta = 0;
static std::vector<int*> g_lists;

static int build(const GemmSpec* s, int n, GemmParams* P, char* err, int errlen) {
  int rc = gemm_build_params(s, n, g_sched, P, err, errlen, g_two_cta);
  if (rc) return rc;
  int* d = nullptr;
  rc = gemm_upload_tile_lists(P, g_sms, &d, err, errlen);
  if (d) g_lists.push_back(d);
  return rc;
}

static float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

struct HostMat {  // row-major fp32 master + bf16 hi/lo device copies with leading dimension ld
  int rows, cols, ld;
  std::vector<float> h;  // [rows*cols] dense
  __nv_bfloat16 *d_hi = nullptr, *d_lo = nullptr;
};

static HostMat make_mat(int rows, int cols, int ld, std::mt19937& rng, bool exact_bf16, float scale = 1.f) {
  HostMat m;
  m.rows = rows; m.cols = cols; m.ld = ld;
  m.h.resize((size_t)rows * cols);
  std::normal_distribution<float> nd(0.f, scale);
  for (auto& x : m.h) { x = nd(rng); if (exact_bf16) x = bf16r(x); }
  std::vector<__nv_bfloat16> hi((size_t)rows * ld), lo((size_t)rows * ld);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < ld; ++c) {
      // padding columns hold garbage on purpose: TMA must clip them
      float x = c < cols ? m.h[(size_t)r * cols + c] : 777.f;
      __nv_bfloat16 h = __float2bfloat16_rn(x);
      hi[(size_t)r * ld + c] = h;
      lo[(size_t)r * ld + c] = __float2bfloat16_rn(x - __bfloat162float(h));
    }
  CK(cudaMalloc(&m.d_hi, hi.size() * 2));
  CK(cudaMalloc(&m.d_lo, lo.size() * 2));
  CK(cudaMemcpy(m.d_hi, hi.data(), hi.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(m.d_lo, lo.data(), lo.size() * 2, cudaMemcpyHostToDevice));
  return m;
}
static void free_mat(HostMat& m) { cudaFree(m.d_hi); cudaFree(m.d_lo); }

struct Case {
  const char* name;
  int M, N, K;
  int a_mn, b_mn;
  int nsplit;
  int out_kind;
  bool bias, relu, mask, stats, dropout;
  int ksplit = 1;
};

// reference element (double accumulate). A(m,k), B(n,k) accessors honour storage majorness.
static double ref_dot(const HostMat& A, const HostMat& B, const Case& c, int m, int n, bool use_bf16_inputs) {
  double acc = 0;
  for (int k = 0; k < c.K; ++k) {
    float a = c.a_mn ? A.h[(size_t)k * c.M + m] : A.h[(size_t)m * c.K + k];
    float b = c.b_mn ? B.h[(size_t)k * c.N + n] : B.h[(size_t)n * c.K + k];
    if (use_bf16_inputs) { a = bf16r(a); b = bf16r(b); }
    acc += (double)a * (double)b;
  }
  return acc;
}

static bool run_case(const Case& c, std::mt19937& rng, int nsamples) {
  const int ldA = ((c.a_mn ? c.M : c.K) + 7) / 8 * 8 + (c.K % 64 ? 8 : 0);
  const int ldB = ((c.b_mn ? c.N : c.K) + 7) / 8 * 8 + (c.K % 64 ? 8 : 0);
  const bool split = c.nsplit == 3;
  HostMat A = c.a_mn ? make_mat(c.K, c.M, ldA, rng, !split) : make_mat(c.M, c.K, ldA, rng, !split);
  HostMat B = c.b_mn ? make_mat(c.K, c.N, ldB, rng, !split, 0.05f) : make_mat(c.N, c.K, ldB, rng, !split, 0.05f);
  const bool f32out = c.out_kind == OUT_F32 || c.out_kind == OUT_F32_REDADD;
  const int ldd = f32out ? (c.N + 3) / 4 * 4 + 4 : (c.N + 7) / 8 * 8 + 8;
  const size_t dbytes = (size_t)c.M * ldd * (f32out ? 4 : 2);
  void *D_hi = nullptr, *D_lo = nullptr;
  CK(cudaMalloc(&D_hi, dbytes));
  CK(cudaMalloc(&D_lo, dbytes));
  std::vector<float> dinit;
  if (c.out_kind == OUT_F32_REDADD) {
    dinit.resize((size_t)c.M * ldd);
    for (auto& x : dinit) x = (float)(rng() % 1000) * 0.001f;
    CK(cudaMemcpy(D_hi, dinit.data(), dbytes, cudaMemcpyHostToDevice));
  } else {
    CK(cudaMemset(D_hi, 0x7f, dbytes));  // poison
    CK(cudaMemset(D_lo, 0x7f, dbytes));
  }
  const int npad = (c.N + 255) / 256 * 256;
  std::vector<float> hbias(npad, 0.f);
  float* dbias = nullptr;
  if (c.bias) {
    std::normal_distribution<float> nd(0.f, 0.5f);
    for (int i = 0; i < c.N; ++i) hbias[i] = nd(rng);
    CK(cudaMalloc(&dbias, npad * 4));
    CK(cudaMemcpy(dbias, hbias.data(), npad * 4, cudaMemcpyHostToDevice));
  }
  HostMat Mk;
  const int ldm = (c.N + 7) / 8 * 8;
  if (c.mask) Mk = make_mat(c.M, c.N, ldm, rng, true);
  const int tiles_m = (c.M + BM - 1) / BM;
  float *dsum = nullptr, *dsq = nullptr;
  const int stat_ld = npad;
  if (c.stats) {
    CK(cudaMalloc(&dsum, (size_t)tiles_m * 4 * stat_ld * 4));
    CK(cudaMalloc(&dsq, (size_t)tiles_m * 4 * stat_ld * 4));
    CK(cudaMemset(dsum, 0, (size_t)tiles_m * 4 * stat_ld * 4));
    CK(cudaMemset(dsq, 0, (size_t)tiles_m * 4 * stat_ld * 4));
  }

  GemmSpec s;
  s.M = c.M; s.N = c.N; s.K = c.K;
  s.A_hi = A.d_hi; s.A_lo = A.d_lo; s.lda = ldA; s.a_mn = c.a_mn;
  s.B_hi = B.d_hi; s.B_lo = B.d_lo; s.ldb = ldB; s.b_mn = c.b_mn;
  s.nsplit = c.nsplit; s.out_kind = c.out_kind; s.ksplit = c.ksplit;
  s.D_hi = D_hi; s.D_lo = D_lo; s.ldd = ldd;
  s.bias = dbias; s.act = c.relu ? 1 : 0;
  if (c.mask) { s.mask_src = Mk.d_hi; s.mask_ld = ldm; s.scale = 2.0f; }
  if (c.dropout) { s.keep = 0.5f; s.seed = 0x1234567890abcdefULL; }
  if (c.stats) { s.stat_sum = dsum; s.stat_sq = dsq; s.stat_ld = stat_ld; }

  GemmParams P;
  char err[256] = {0};
  int rc = build(&s, 1, &P, err, sizeof(err));
  if (rc) { printf("[%s] build_params failed: %s\n", c.name, err); g_fail++; return false; }
  const int reps = c.out_kind == OUT_F32_REDADD ? 2 : 1;
  for (int i = 0; i < reps; ++i) {
    rc = gemm_launch(P, g_sms, 0);
    if (rc) { printf("[%s] launch failed: %d %s\n", c.name, rc, cudaGetErrorString((cudaError_t)rc)); exit(3); }
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("[%s] kernel failed: %s\n", c.name, cudaGetErrorString(e)); exit(3); }

  std::vector<uint8_t> out(dbytes), out_lo(dbytes);
  CK(cudaMemcpy(out.data(), D_hi, dbytes, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(out_lo.data(), D_lo, dbytes, cudaMemcpyDeviceToHost));
  auto get = [&](int m, int n) -> double {
    if (f32out) return reinterpret_cast<float*>(out.data())[(size_t)m * ldd + n];
    double v = __bfloat162float(reinterpret_cast<__nv_bfloat16*>(out.data())[(size_t)m * ldd + n]);
    if (c.out_kind == OUT_BF16_SPLIT)
      v += __bfloat162float(reinterpret_cast<__nv_bfloat16*>(out_lo.data())[(size_t)m * ldd + n]);
    return v;
  };
  const uint32_t thr = dropout_threshold(0.5);
  auto expect = [&](int m, int n, double* z_out) -> double {
    double v = ref_dot(A, B, c, m, n, !split);
    if (c.bias) v += hbias[n];
    if (z_out) *z_out = v;
    if (c.relu) v = v > 0 ? v : 0;
    if (c.dropout) {
      Philox4 r = philox4x32_10((uint32_t)(n >> 3), (uint32_t)m, 0, 0, (uint32_t)s.seed, (uint32_t)(s.seed >> 32));
      v = ((dropout_keep_bits(r, thr) >> (n & 7)) & 1u) ? v * 2.0 : 0.0;
    }
    if (c.mask) v = Mk.h[(size_t)m * c.N + n] > 0 ? v * 2.0 : 0.0;
    if (c.out_kind == OUT_F32_REDADD) v = 2.0 * v + dinit[(size_t)m * ldd + n];
    return v;
  };
  double max_err = 0, max_ref = 0;
  int bad = 0;
  const bool full = (size_t)c.M * c.N <= (size_t)nsamples;
  const int total = full ? c.M * c.N : nsamples;
  // error model: fp32 accumulation + output rounding (bf16: 2^-9 rel; split/f32: ~1e-6 rel)
  const double out_rel = (c.out_kind == OUT_BF16) ? 4.0e-3 : 2.0e-5;
  const double in_abs = split ? 4e-5 : 2e-5;  // bf16x3 drops the lo*lo term (2^-16 rel per product)
  for (int i = 0; i < total; ++i) {
    int m, n;
    if (full) { m = i / c.N; n = i % c.N; }
    else {
      m = rng() % c.M; n = rng() % c.N;
      if (i % 7 == 0) m = c.M - 1 - (rng() % std::min(c.M, 130));  // stress the ragged edges
      if (i % 5 == 0) n = c.N - 1 - (rng() % std::min(c.N, 70));
    }
    const double ex = expect(m, n, nullptr), got = get(m, n);
    const double scale = std::sqrt((double)c.K) * 0.05 + 1.0;
    const double tol = out_rel * std::fabs(ex) + in_abs * scale * (c.out_kind == OUT_F32_REDADD ? 4 : 2);
    const double er = std::fabs(ex - got);
    if (!(er <= tol)) {
      if (bad < 5) printf("   mismatch (%d,%d): expect %.6f got %.6f\n", m, n, ex, got);
      ++bad;
    }
    if (er > max_err) max_err = er;
    if (std::fabs(ex) > max_ref) max_ref = std::fabs(ex);
  }
  // padding columns of the output must be untouched (TMA clips at N)
  int pad_bad = 0;
  if (c.out_kind != OUT_F32_REDADD) {
    // TMA clips stores at 16-byte granularity: the tail of the last 16-byte chunk may be written
    const int nclip = f32out ? (c.N + 3) / 4 * 4 : (c.N + 7) / 8 * 8;
    for (int m = 0; m < c.M; m += std::max(1, c.M / 64))
      for (int n = nclip; n < ldd; ++n) {
        if (f32out) { uint32_t u = reinterpret_cast<uint32_t*>(out.data())[(size_t)m * ldd + n]; if (u != 0x7f7f7f7fu) ++pad_bad; }
        else { uint16_t u = reinterpret_cast<uint16_t*>(out.data())[(size_t)m * ldd + n]; if (u != 0x7f7fu) ++pad_bad; }
      }
  }
  int stat_bad = 0;
  if (c.stats) {
    std::vector<float> hs((size_t)tiles_m * 4 * stat_ld), hq((size_t)tiles_m * 4 * stat_ld);
    CK(cudaMemcpy(hs.data(), dsum, hs.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hq.data(), dsq, hq.size() * 4, cudaMemcpyDeviceToHost));
    for (int t = 0; t < 24; ++t) {
      const int g = rng() % ((c.M + 31) / 32), n = rng() % c.N;
      double s1 = 0, s2 = 0;
      for (int m = g * 32; m < std::min(c.M, g * 32 + 32); ++m) { double z; expect(m, n, &z); s1 += z; s2 += z * z; }
      const double g1 = hs[(size_t)g * stat_ld + n], g2 = hq[(size_t)g * stat_ld + n];
      if (std::fabs(g1 - s1) > 1e-3 * (1 + std::fabs(s1)) || std::fabs(g2 - s2) > 1e-3 * (1 + std::fabs(s2))) {
        if (stat_bad < 3) printf("   stat mismatch g=%d n=%d: sum %.5f vs %.5f, sq %.5f vs %.5f\n", g, n, g1, s1, g2, s2);
        ++stat_bad;
      }
    }
  }
  const bool ok = bad == 0 && pad_bad == 0 && stat_bad == 0;
  printf("[%s]%s %-44s M=%d N=%d K=%d  checked=%d max_err=%.3e (max|ref|=%.2f) bad=%d pad_bad=%d stat_bad=%d\n",
         ok ? "PASS" : "FAIL", g_two_cta ? "[2cta]" : "      ", c.name, c.M, c.N, c.K, total, max_err, max_ref, bad, pad_bad, stat_bad);
  if (!ok) g_fail++;
  free_mat(A); free_mat(B);
  if (c.mask) free_mat(Mk);
  cudaFree(D_hi); cudaFree(D_lo); cudaFree(dbias); cudaFree(dsum); cudaFree(dsq);
  return ok;
}

// fused backward launch: problem 0 = wgrad (MN/MN, REDADD), problem 1 = dgrad (K/K, bf16 + mask)
static void run_fused_bwd(int Bsz, int Kin, int Nout, int nsplit, std::mt19937& rng, int nsamples, bool time_it) {
  const bool split = nsplit == 3;
  HostMat X = make_mat(Bsz, Kin, Kin, rng, !split);          // activations of the layer below [B,K]
  HostMat dZ = make_mat(Bsz, Nout, Nout, rng, !split, 0.05f);  // [B,N]
  HostMat W = make_mat(Kin, Nout, Nout, rng, !split, 0.05f);   // [K,N]
  float* G; CK(cudaMalloc(&G, (size_t)Kin * Nout * 4)); CK(cudaMemset(G, 0, (size_t)Kin * Nout * 4));
  __nv_bfloat16 *dX_hi, *dX_lo;
  CK(cudaMalloc(&dX_hi, (size_t)Bsz * Kin * 2)); CK(cudaMalloc(&dX_lo, (size_t)Bsz * Kin * 2));
  GemmSpec s[2];
  s[0].M = Kin; s[0].N = Nout; s[0].K = Bsz;
  s[0].A_hi = X.d_hi; s[0].A_lo = X.d_lo; s[0].lda = Kin; s[0].a_mn = 1;
  s[0].B_hi = dZ.d_hi; s[0].B_lo = dZ.d_lo; s[0].ldb = Nout; s[0].b_mn = 1;
  s[0].nsplit = nsplit; s[0].out_kind = OUT_F32_REDADD; s[0].D_hi = G; s[0].ldd = Nout;
  s[1].M = Bsz; s[1].N = Kin; s[1].K = Nout;
  s[1].A_hi = dZ.d_hi; s[1].A_lo = dZ.d_lo; s[1].lda = Nout; s[1].a_mn = 0;
  s[1].B_hi = W.d_hi; s[1].B_lo = W.d_lo; s[1].ldb = Nout; s[1].b_mn = 0;
  s[1].nsplit = nsplit; s[1].out_kind = split ? OUT_BF16_SPLIT : OUT_BF16; s[1].D_hi = dX_hi; s[1].D_lo = dX_lo; s[1].ldd = Kin;
  s[1].mask_src = X.d_hi; s[1].mask_ld = Kin; s[1].scale = 1.0f;
  GemmParams P; char err[256];
  if (build(s, 2, &P, err, sizeof(err))) { printf("fused build failed: %s\n", err); g_fail++; return; }
  int rc = gemm_launch(P, g_sms, 0);
  cudaError_t e = cudaDeviceSynchronize();
  if (rc || e != cudaSuccess) { printf("fused bwd kernel failed: %d %s\n", rc, cudaGetErrorString(e)); exit(3); }
  std::vector<float> hG((size_t)Kin * Nout);
  std::vector<__nv_bfloat16> hX((size_t)Bsz * Kin), hXl((size_t)Bsz * Kin);
  CK(cudaMemcpy(hG.data(), G, hG.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hX.data(), dX_hi, hX.size() * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hXl.data(), dX_lo, hXl.size() * 2, cudaMemcpyDeviceToHost));
  int bad = 0; double maxe = 0;
  for (int i = 0; i < nsamples; ++i) {
    { const int k = rng() % Kin, n = rng() % Nout; double acc = 0;
      for (int b = 0; b < Bsz; ++b) acc += (double)X.h[(size_t)b * Kin + k] * dZ.h[(size_t)b * Nout + n];
      const double er = std::fabs(acc - hG[(size_t)k * Nout + n]);
      if (er > 2e-5 * std::fabs(acc) + 1e-4 * std::sqrt((double)Bsz) * 0.05) { if (bad < 5) printf("   wgrad mismatch (%d,%d) %.6f vs %.6f\n", k, n, acc, hG[(size_t)k * Nout + n]); ++bad; }
      maxe = std::max(maxe, er); }
    { const int b = rng() % Bsz, k = rng() % Kin; double acc = 0;
      for (int n = 0; n < Nout; ++n) acc += (double)dZ.h[(size_t)b * Nout + n] * W.h[(size_t)k * Nout + n];
      if (!(bf16r(X.h[(size_t)b * Kin + k]) > 0)) acc = 0;
      double got = __bfloat162float(hX[(size_t)b * Kin + k]); if (split) got += __bfloat162float(hXl[(size_t)b * Kin + k]);
      const double er = std::fabs(acc - got);
      if (er > (split ? 2e-5 : 4e-3) * std::fabs(acc) + 1e-4) { if (bad < 5) printf("   dgrad mismatch (%d,%d) %.6f vs %.6f\n", b, k, acc, got); ++bad; }
      maxe = std::max(maxe, er); }
  }
  printf("[%s]%s fused wgrad+dgrad B=%d K=%d N=%d nsplit=%d  max_err=%.3e bad=%d\n", bad ? "FAIL" : "PASS", g_two_cta ? "[2cta]" : "      ", Bsz, Kin, Nout, nsplit, maxe, bad);
  if (bad) g_fail++;
  if (time_it) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) gemm_launch(P, g_sms, 0);
    cudaEventRecord(e0);
    const int it = 20;
    for (int i = 0; i < it; ++i) gemm_launch(P, g_sms, 0);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= it;
    const double fl = 2.0 * 2.0 * Bsz * (double)Kin * Nout * nsplit;
    printf("       fused bwd: %.1f us  %.1f TFLOP/s (bf16 MMA flops, nsplit=%d)\n", ms * 1e3, fl / ms / 1e9, nsplit);
    // the same two problems as two back-to-back launches
    GemmParams P0, P1;
    if (!build(&s[0], 1, &P0, err, sizeof(err)) && !build(&s[1], 1, &P1, err, sizeof(err))) {
      for (int i = 0; i < 3; ++i) { gemm_launch(P0, g_sms, 0); gemm_launch(P1, g_sms, 0); }
      cudaEventRecord(e0);
      for (int i = 0; i < it; ++i) { gemm_launch(P0, g_sms, 0); gemm_launch(P1, g_sms, 0); }
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      cudaEventElapsedTime(&ms, e0, e1); ms /= it;
      printf("       separate wgrad + dgrad launches: %.1f us  %.1f TFLOP/s\n", ms * 1e3, fl / ms / 1e9);
      cudaEventRecord(e0);
      for (int i = 0; i < it; ++i) { gemm_launch(P1, g_sms, 0); gemm_launch(P0, g_sms, 0); }
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      cudaEventElapsedTime(&ms, e0, e1); ms /= it;
      printf("       separate dgrad + wgrad launches: %.1f us  %.1f TFLOP/s\n", ms * 1e3, fl / ms / 1e9);
    }
  }
  free_mat(X); free_mat(dZ); free_mat(W); cudaFree(G); cudaFree(dX_hi); cudaFree(dX_lo);
}

static void bench_case(const char* name, int M, int N, int K, int a_mn, int b_mn, int out_kind, int nsplit, std::mt19937& rng,
                       int ksplit = 1, bool maskbits = false, int iters = 30) {
  HostMat A = a_mn ? make_mat(K, M, (M + 7) / 8 * 8, rng, true) : make_mat(M, K, (K + 7) / 8 * 8, rng, true);
  HostMat B = b_mn ? make_mat(K, N, (N + 7) / 8 * 8, rng, true, 0.05f) : make_mat(N, K, (K + 7) / 8 * 8, rng, true, 0.05f);
  const bool f32out = out_kind >= OUT_F32;
  void *D, *Dl; CK(cudaMalloc(&D, (size_t)M * N * 4)); CK(cudaMalloc(&Dl, (size_t)M * N * 4));
  CK(cudaMemset(D, 0, (size_t)M * N * 4));
  float* bias; CK(cudaMalloc(&bias, (N + 255) / 256 * 256 * 4)); CK(cudaMemset(bias, 0, (N + 255) / 256 * 256 * 4));
  GemmSpec s; s.M = M; s.N = N; s.K = K;
  s.A_hi = A.d_hi; s.A_lo = A.d_lo; s.lda = A.ld; s.a_mn = a_mn;
  s.B_hi = B.d_hi; s.B_lo = B.d_lo; s.ldb = B.ld; s.b_mn = b_mn;
  s.nsplit = nsplit; s.out_kind = out_kind; s.D_hi = D; s.D_lo = Dl; s.ldd = N; s.bias = bias; s.act = f32out ? 0 : 1;
  s.ksplit = ksplit;
  if (out_kind == OUT_F32_REDADD) s.bias = nullptr;
  uint32_t* bits = nullptr;
  if (maskbits) {  // as the training forward runs it: 1 gradient-pass bit per output element
    CK(cudaMalloc(&bits, (size_t)((N + 31) / 32 + 8) * M * 4));
    s.mask_bits_out = bits; s.mask_bits_ld = M;
  }
  GemmParams P; char err[256];
  if (build(&s, 1, &P, err, sizeof(err))) { printf("bench build failed %s\n", err); return; }
  for (int i = 0; i < 3; ++i) gemm_launch(P, g_sms, 0);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int it = iters;
  cudaEventRecord(e0);
  for (int i = 0; i < it; ++i) gemm_launch(P, g_sms, 0);
  cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= it;
  if (bits) cudaFree(bits);
  printf("[BENCH]%s %-34s M=%d N=%d K=%d nsplit=%d: %8.1f us  %7.1f TFLOP/s (algorithmic 2MNK: %.1f)\n", g_two_cta ? "[2cta]" : "      ", name, M, N, K, nsplit,
         ms * 1e3, 2.0 * M * N * (double)K * nsplit / ms / 1e9, 2.0 * M * N * (double)K / ms / 1e9);
  free_mat(A); free_mat(B); cudaFree(D); cudaFree(Dl); cudaFree(bias);
}

// "sched": the CTA-pair work-list scheduler alone (host code, runs without a GPU).  For a sweep of launch shapes
// (forward, fused wgrad+dgrad, split-K layer-0 wgrad; ragged N; bf16 and bf16x3) every output tile must be covered
// exactly once - whole, as its two 128-column halves, or (right half outside N) as its left half only - the lists must
// be -1 terminated within their stride, no pair may be idle while another holds two items more than it needs to, and
// the result must be deterministic.
static int sched_selftest() {
  constexpr int kHalfShift = 28;  // entry = tile | half << 28 (gemm.cu)
  int shapes = 0, bad = 0;
  const int Ns[] = {2048, 1936, 3401, 183, 256, 200};
  for (int B = 1; B <= 20000; B = B < 600 ? B + 37 : B + 611)
    for (int N : Ns)
      for (int kind = 0; kind < 4; ++kind)
        for (int nsplit = 1; nsplit <= 3; nsplit += 2) {
          const int H = N > 1000 ? 2048 : 256;
          GemmParams gp;
          memset(&gp, 0, sizeof(gp));
          gp.two_cta = 1;
          int begin = 0;
          auto add = [&](int M, int Nn, int K, int ks) {
            GemmProblem& p = gp.p[gp.nprob++];
            p.M = M; p.N = Nn; p.K = K; p.nsplit = nsplit;
            p.tiles_m = (M + 255) / 256; p.tiles_n = (Nn + BN - 1) / BN; p.tile_begin = begin;
            p.num_kb = (K + BK - 1) / BK; p.ksplit = ks; p.kb_per_split = (p.num_kb + ks - 1) / ks;
            begin += p.tiles_m * p.tiles_n * ks;
          };
          if (kind == 0) add(B, N, H, 1);
          else if (kind == 1) { add(H, N, B, 1); add(B, H, N, 1); }
          else if (kind == 2) { const int kb = (B + BK - 1) / BK; add(440, N, B, std::max(1, std::min(9, kb / 16))); }
          else add(B, N, 440, 1);
          gp.total_tiles = begin;
          std::vector<int> flat, again;
          int pairs = 0, stride = 0, p2 = 0, s2 = 0;
          gemm_schedule_tile_lists(&gp, 148, &flat, &pairs, &stride);
          gemm_schedule_tile_lists(&gp, 148, &again, &p2, &s2);
          ++shapes;
          bool ok = flat == again && pairs == p2 && stride == s2 && pairs >= 1 && pairs <= 74 &&
                    flat.size() == static_cast<size_t>(pairs) * stride;
          std::vector<int> seen(begin, 0);  // bit 0: whole, bit 1: left half, bit 2: right half
          size_t longest = 0, shortest = 1u << 30;
          for (int p = 0; ok && p < pairs; ++p) {
            size_t n = 0;
            while (n < static_cast<size_t>(stride) && flat[p * stride + n] >= 0) ++n;
            ok = ok && n < static_cast<size_t>(stride);  // terminator inside the row
            for (size_t i = n; ok && i < static_cast<size_t>(stride); ++i) ok = flat[p * stride + i] == -1;
            for (size_t i = 0; ok && i < n; ++i) {
              const int e = flat[p * stride + i], t = e & ((1 << kHalfShift) - 1), half = e >> kHalfShift;
              ok = t < begin && half >= 0 && half <= 2 && !(seen[t] & (1 << half));
              if (ok) seen[t] |= 1 << half;
            }
            longest = std::max(longest, n);
            shortest = std::min(shortest, n);
          }
          for (int t = 0; ok && t < begin; ++t) {
            const GemmProblem& pr = gp.p[(gp.nprob > 1 && t >= gp.p[1].tile_begin) ? 1 : 0];
            const int n_blk = (t - pr.tile_begin) % pr.tiles_n;
            const bool left_only = n_blk * BN + BN / 2 >= pr.N;
            ok = left_only ? seen[t] == 2 : (seen[t] == 1 || seen[t] == 6);
          }
          ok = ok && shortest >= 1;  // only pairs with work are launched
          if (!ok && ++bad <= 5) printf("[FAIL] sched B=%d N=%d kind=%d nsplit=%d (pairs %d stride %d)\n", B, N, kind, nsplit, pairs, stride);
        }
  printf("sched: %d launch shapes checked, %d failures\n", shapes, bad);
  printf(bad ? "selftest_gemm sched: FAILED\n" : "selftest_gemm sched: ALL PASS\n");
  return bad ? 1 : 0;
}

int main(int argc, char** argv) {
  std::string mode = argc > 1 ? argv[1] : "quick";
  if (mode == "sched") return sched_selftest();
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  g_sms = prop.multiProcessorCount;
  printf("device: %s, %d SMs, cc %d.%d, smem/block optin %zu, gemm smem %zu\n", prop.name, g_sms, prop.major, prop.minor,
         prop.sharedMemPerBlockOptin, gemm_smem_bytes());
  CK(cudaMalloc(&g_sched, 8)); CK(cudaMemset(g_sched, 0, 8));
  int rc = gemm_init();
  if (rc) { printf("gemm_init failed: %s\n", cudaGetErrorString((cudaError_t)rc)); return 2; }
  std::mt19937 rng(1234);
  const int NS = 20000;

  if (mode == "l0") {  // the north-star shape alone (8192 x 440 -> 2048), CTA-pair kernel: forward as the training step runs
    g_two_cta = 1;     // it (bias + relu + gradient-pass bits) and the split-K weight gradient; few launches, for ncu
    const int it = argc > 2 ? atoi(argv[2]) : 30;
    bench_case("fwd layer0 bias+relu+bits", 8192, 2048, 440, 0, 1, OUT_BF16, 1, rng, 1, true, it);
    bench_case("wgrad layer0 MN/MN redadd ksplit 9", 440, 2048, 8192, 1, 1, OUT_F32_REDADD, 1, rng, 9, false, it);
    bench_case("fwd hidden bias+relu+bits", 8192, 2048, 2048, 0, 1, OUT_BF16, 1, rng, 1, true, it);
    return 0;
  }
  for (g_two_cta = (argc > 2 ? atoi(argv[2]) : 0); g_two_cta <= (argc > 3 ? atoi(argv[3]) : 1); ++g_two_cta) {
  if (mode != "bench") {
    std::vector<Case> cases = {
        // name, M,N,K, a_mn,b_mn, nsplit, out, bias,relu,mask,stats,dropout
        {"fwd-like K/K one tile bf16", 128, 256, 64, 0, 0, 1, OUT_BF16, false, false, false, false, false},
        {"K/K multi-k f32", 128, 256, 256, 0, 0, 1, OUT_F32, false, false, false, false, false},
        {"fwd K/MN one tile bf16", 128, 256, 64, 0, 1, 1, OUT_BF16, false, false, false, false, false},
        {"fwd K/MN bias+relu", 256, 512, 192, 0, 1, 1, OUT_BF16, true, true, false, false, false},
        {"wgrad MN/MN redadd", 128, 256, 128, 1, 1, 1, OUT_F32_REDADD, false, false, false, false, false},
        {"wgrad MN/MN ragged", 440, 183, 300, 1, 1, 1, OUT_F32_REDADD, false, false, false, false, false},
        {"dgrad K/K mask", 384, 256, 1936, 0, 0, 1, OUT_BF16, false, false, true, false, false},
        {"fwd ragged K=440 N=1936 f32 logits", 300, 1936, 440, 0, 1, 1, OUT_F32, true, false, false, false, false},
        {"fwd ragged N=183", 256, 183, 256, 0, 1, 1, OUT_F32, true, false, false, false, false},
        {"fwd ragged hidden=200", 300, 200, 136, 0, 1, 1, OUT_BF16, true, true, false, false, false},
        {"fwd split x3 bf16-split out", 256, 512, 192, 0, 1, 3, OUT_BF16_SPLIT, true, true, false, false, false},
        {"fwd split x3 f32", 300, 1936, 440, 0, 1, 3, OUT_F32, true, false, false, false, false},
        {"wgrad split x3", 440, 256, 512, 1, 1, 3, OUT_F32_REDADD, false, false, false, false, false},
        {"dgrad split x3 mask", 384, 256, 1936, 0, 0, 3, OUT_BF16_SPLIT, false, false, true, false, false},
        {"fwd bn-stats split", 300, 512, 440, 0, 1, 3, OUT_BF16_SPLIT, true, false, false, true, false},
        {"fwd relu+dropout", 256, 512, 128, 0, 1, 1, OUT_BF16, true, true, false, false, true},
        {"fwd many tiles (persistent loop)", 2048, 2048, 256, 0, 1, 1, OUT_BF16, true, true, false, false, false},
        {"fwd A-stationary ragged M, K=440", 700, 768, 440, 0, 1, 1, OUT_BF16, true, true, false, false, false},
        {"fwd A-stationary K/K f32, 5 runs", 1100, 512, 300, 0, 0, 1, OUT_F32, true, false, false, false, false},
        {"wgrad split-K x5 ragged", 440, 300, 8192 + 40, 1, 1, 1, OUT_F32_REDADD, false, false, false, false, false, 5},
        {"wgrad split-K x3 bf16x3", 256, 256, 2048, 1, 1, 3, OUT_F32_REDADD, false, false, false, false, false, 3},
    };
    for (auto& c : cases) run_case(c, rng, NS);
    run_fused_bwd(512, 256, 512, 1, rng, 300, false);
    run_fused_bwd(512, 440, 183 + 1, 3, rng, 300, false);  // N even for the bf16 pitch
    if (mode == "full") {
      std::vector<Case> big = {
          {"C2 layer0 fwd 8192x440->2048", 8192, 2048, 440, 0, 1, 1, OUT_BF16, true, true, false, false, false},
          {"C2 hidden fwd 8192x2048->2048", 8192, 2048, 2048, 0, 1, 1, OUT_BF16, true, true, false, false, false},
          {"C2 out fwd 8192x2048->1936 x3", 8192, 1936, 2048, 0, 1, 3, OUT_F32, true, false, false, false, false},
          {"C4 out fwd 4096x2048->3401", 4096, 3401, 2048, 0, 1, 1, OUT_F32, true, false, false, false, false},
      };
      for (auto& c : big) run_case(c, rng, 3000);
      run_fused_bwd(8192, 2048, 2048, 1, rng, 200, true);
    }
  }
  if (mode == "bench" || mode == "full") {
    bench_case("fwd layer0 K/MN bf16", 8192, 2048, 440, 0, 1, OUT_BF16, 1, rng);
    bench_case("fwd hidden K/MN bf16", 8192, 2048, 2048, 0, 1, OUT_BF16, 1, rng);
    bench_case("fwd out K/MN f32", 8192, 1936, 2048, 0, 1, OUT_F32, 1, rng);
    bench_case("dgrad hidden K/K bf16", 8192, 2048, 2048, 0, 0, OUT_BF16, 1, rng);
    bench_case("wgrad hidden MN/MN redadd", 2048, 2048, 8192, 1, 1, OUT_F32_REDADD, 1, rng);
    bench_case("fwd hidden K/MN x3 split", 8192, 2048, 2048, 0, 1, OUT_BF16_SPLIT, 3, rng);
    bench_case("square 8192^3 K/K bf16 (vs cuBLAS peak)", 8192, 8192, 8192, 0, 0, OUT_BF16, 1, rng);
  }
  }
  printf("selftest_gemm: %s (%d failures)\n", g_fail ? "FAILED" : "ALL PASS", g_fail);
  return g_fail ? 1 : 0;
}
